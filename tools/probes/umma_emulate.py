#!/usr/bin/env python
"""NumPy emulation of the data flow of csrc/pyrdown_umma.cu (tile plan, baked weight slices, lagged strips, register
carries, border patches) against oracle.evm.pyrdown_cascade.  Test infrastructure: it validates the index algebra on
the CPU so that only the hardware encodings (descriptors, swizzle) are left to debug on the GPU.

    python tools/probes/umma_emulate.py [H W]
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

W5 = np.array([1, 4, 6, 4, 1], dtype=np.int64)
W13 = np.array([1, 4, 10, 20, 31, 40, 44, 40, 31, 20, 10, 4, 1], dtype=np.int64)
PX = 20           # level-2 pixels per strip
NB = 12 * PX      # fresh bytes per strip (MMA N)
CARRY = 30        # D columns carried from the previous strip
MAX_N4 = 29


def refl(i, n):
    if n == 1:
        return 0
    i = abs(i)
    if i >= n:
        i = 2 * (n - 1) - i
    return abs(i)


def down_matrix(n):
    no = (n + 1) // 2
    m = np.zeros((no, n), dtype=np.int64)
    for r in range(no):
        for d in range(5):
            m[r, refl(2 * r - 2 + d, n)] += W5[d]
    return m


def make_plan(H, W):
    """The split with the fewest 64-row stages per frame (ties: the smaller tile height), as csrc/pyrdown_umma.cu."""
    h4 = H
    for _ in range(4):
        h4 = (h4 + 1) // 2
    ntiles = -(-h4 // MAX_N4)
    best = None
    for n4 in range(-(-h4 // ntiles), MAX_N4 + 1):
        p = make_plan_n4(H, W, n4)
        cost = sum(t["nks"] for t in p["tiles"])
        if best is None or cost < best[0]:
            best = (cost, p)
    return best[1]


def make_plan_n4(H, W, n4):
    assert W % 80 == 0
    h = [H]
    w = [W]
    for _ in range(4):
        h.append((h[-1] + 1) // 2)
        w.append((w[-1] + 1) // 2)
    c2v = down_matrix(h[1]) @ down_matrix(h[0])          # h2 x H, rows sum to 256
    c2h = down_matrix(w[1]) @ down_matrix(w[0])          # w2 x W
    ntiles = -(-h[4] // n4)
    tiles = []
    for t in range(ntiles):
        a = t * n4
        n4t = min(n4, h[4] - a)
        l3 = sorted({refl(2 * r - 2 + d, h[3]) for r in range(a, a + n4t) for d in range(5)})
        g0, g1 = l3[0], l3[-1]
        l2 = sorted({refl(2 * g - 2 + d, h[2]) for g in range(g0, g1 + 1) for d in range(5)})
        r0, r1 = l2[0], l2[-1]
        nr = r1 - r0 + 1
        assert nr <= 128 and g1 - g0 + 1 <= 64 and n4t <= 32
        nz = np.nonzero(c2v[r0:r1 + 1].sum(axis=0))[0]
        i0 = 4 * r0 - 8
        assert i0 <= nz[0]
        nks = 2 * int(-(-(nz[-1] - i0 + 1) // 64))
        slices, codes = [], []
        for ks in range(nks):
            s = np.zeros((128, 32), dtype=np.int64)
            gen = np.zeros((128, 32), dtype=np.int64)
            for m in range(nr):
                for k in range(32):
                    row = i0 + 32 * ks + k
                    if 0 <= row < H:
                        s[m, k] = c2v[r0 + m, row]
                    j = 32 * ks + k - 4 * m - 2
                    if 0 <= j <= 12:
                        gen[m, k] = W13[j]
            codes.append("g" if np.array_equal(s, gen) and ks <= 16 else "s")
            slices.append(s)
        tiles.append(dict(a=a, n4=n4t, g0=g0, n3=g1 - g0 + 1, r0=r0, nr=nr, i0=i0, nks=nks, slices=slices, codes=codes))
    # horizontal specials in window coordinates (L0 px 4x-6+j)
    special = {}
    for x in range(w[2]):
        v = np.zeros(13, dtype=np.int64)
        inside = 0
        for j in range(13):
            p = 4 * x - 6 + j
            if 0 <= p < W:
                v[j] = c2h[x, p]
                inside += c2h[x, p]
        assert inside == 256, (x, inside)          # every folded weight lands inside the window
        if not np.array_equal(v, W13):
            special[x] = v
    assert set(special) <= {0, 1, w[2] - 1}, sorted(special)
    return dict(h=h, w=w, tiles=tiles, special=special, nstrips=w[2] // PX + 1)


def emulate(frame, plan):
    """frame (H, W, 3) uint8 -> level 4 (h4, w4, 3) float64 following the kernel's data flow."""
    H, W, _ = frame.shape
    h, w = plan["h"], plan["w"]
    img = frame.reshape(H, W * 3).astype(np.int64)
    S = plan["nstrips"]
    out = np.full((h[4], w[4], 3), np.nan)
    sp = plan["special"]
    for tile in plan["tiles"]:
        nr, n3, n4, r0, g0, a = tile["nr"], tile["n3"], tile["n4"], tile["r0"], tile["g0"], tile["a"]
        carry = np.zeros((128, CARRY), dtype=np.int64)
        carry_l2 = np.zeros((128, 3, 3))                 # L2 slots -3..-1
        carry_l3 = np.zeros((64, 3, 3))
        rows3 = np.array([[refl(2 * (g0 + i) - 2 + d, h[2]) - r0 for d in range(5)] for i in range(n3)])
        rows4 = np.array([[refl(2 * (a + j) - 2 + d, h[3]) - g0 for d in range(5)] for j in range(n4)])
        assert rows3.min() >= 0 and rows3.max() < nr and rows4.min() >= 0 and rows4.max() < n3
        for s in range(S):
            # tensor core: D (128 x 240) = sum over k-steps of slice (128 x 32) x image rows (32 x 240)
            D = np.zeros((128, NB), dtype=np.int64)
            for ks in range(tile["nks"]):
                B = np.zeros((32, NB), dtype=np.int64)
                for k in range(32):
                    row = tile["i0"] + 32 * ks + k
                    if 0 <= row < H:
                        seg = img[row, NB * s: NB * s + NB]
                        B[k, :len(seg)] = seg
                D += tile["slices"][ks] @ B
            cols = np.concatenate([carry, D], axis=1)     # index = strip-local column + 30
            l2 = np.zeros((128, PX, 3))
            for e in range(PX):
                x = PX * s - 1 + e
                wv = sp.get(x, W13)
                for ch in range(3):
                    l2[:, e, ch] = sum(wv[j] * cols[:, 12 * e + 3 * j + ch] for j in range(13))
            if s == 0:
                l2[:, 0] = l2[:, 2]
                carry_l2[:, 2] = l2[:, 3]
            if s == S - 1:
                l2[:, 1] = carry_l2[:, 2]
            ext = np.concatenate([carry_l2, l2], axis=1)  # slot e -> ext[e + 3]
            l3h = np.zeros((128, 10, 3))
            for e3 in range(10):
                l3h[:, e3] = sum(W5[d] * ext[:, 2 * e3 + d] for d in range(5))
            carry = cols[:, -CARRY:]
            carry_l2 = l2[:, PX - 3:]
            # level-3 warps: vertical pass over the L2 rows, then the horizontal pass of level 4
            l3 = np.zeros((64, 10, 3))
            for i in range(n3):
                l3[i] = sum(W5[d] * l3h[rows3[i, d]] for d in range(5))
            if s == 0:
                l3[:, 0] = l3[:, 2]
                carry_l3[:, 2] = l3[:, 3]
            if s == S - 1:
                l3[:, 1] = carry_l3[:, 2]
            ext3 = np.concatenate([carry_l3, l3], axis=1)
            l4h = np.zeros((64, 5, 3))
            for e4 in range(5):
                l4h[:, e4] = sum(W5[d] * ext3[:, 2 * e4 + d] for d in range(5))
            carry_l3 = l3[:, 7:]
            # level-4 warp
            for j in range(n4):
                v = sum(W5[d] * l4h[rows4[j, d]] for d in range(5)) * 2.0 ** -32
                for e4 in range(5):
                    x4 = 5 * s - 1 + e4
                    if 0 <= x4 < w[4] and (s < S - 1 or e4 == 0):
                        out[a + j, x4] = v[e4]
    return out


def main():
    from oracle import evm as oevm
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
    plan = make_plan(H, W)
    for t in plan["tiles"]:
        print({k: v for k, v in t.items() if k != "slices"})
    print("specials", {k: v.tolist() for k, v in plan["special"].items()}, "strips", plan["nstrips"])
    rng = np.random.default_rng(1)
    fr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    got = emulate(fr, plan)
    ref = oevm.pyrdown_cascade(fr[None], 4)[0]
    assert not np.isnan(got).any()
    print("max abs err vs oracle", np.abs(got - ref).max(), "max", ref.max())


if __name__ == "__main__":
    main()
