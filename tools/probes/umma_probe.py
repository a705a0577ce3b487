#!/usr/bin/env python
"""Bring-up probe of csrc/pyrdown_umma.cu on a GPU box: raw accumulators of one (item, strip) against the NumPy
emulation (descriptor / swizzle check), the level-4 output against the oracle and the streaming kernel, then timing.

    python tools/probes/umma_probe.py [H W T] [--time]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    import umma_emulate as em
    import video_heart_rate_b200 as vhr
    from oracle import evm as oevm
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    H, W, T = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (1080, 1920, 2)
    eng = vhr.Engine(0)
    rng = np.random.default_rng(7)
    fr = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    frd = torch.as_tensor(fr, device=eng.tdev)
    plan = em.make_plan(H, W)
    res = {"H": H, "W": W, "T": T}
    # 1. raw accumulators
    for item, strip in ((0, 1), (len(plan["tiles"]) - 1, 0), (0, plan["nstrips"] - 2)):
        out, dbg = eng.pyrdown_tc_accumulators(frd, item, strip)
        torch.cuda.synchronize()
        f, t = divmod(item, len(plan["tiles"]))
        tile = plan["tiles"][t]
        img = fr[f].reshape(H, W * 3).astype(np.int64)
        D = np.zeros((128, 240), dtype=np.int64)
        for ks in range(tile["nks"]):
            B = np.zeros((32, 240), dtype=np.int64)
            for k in range(32):
                row = tile["i0"] + 32 * ks + k
                if 0 <= row < H:
                    seg = img[row, 240 * strip: 240 * strip + 240]
                    B[k, :len(seg)] = seg
            D += tile["slices"][ks] @ B
        got = dbg.cpu().numpy().astype(np.int64)
        nr = tile["nr"]
        bad = int((got[:nr] != D[:nr]).sum())
        res[f"acc_item{item}_strip{strip}_mismatch"] = bad
        if bad:
            rows, cols = np.nonzero(got[:nr] != D[:nr])
            res[f"acc_item{item}_strip{strip}_first"] = [int(rows[0]), int(cols[0]), int(got[rows[0], cols[0]]), int(D[rows[0], cols[0]])]
            res[f"acc_item{item}_strip{strip}_rows_bad"] = sorted(set(int(r) for r in rows))[:20]
            res[f"acc_item{item}_strip{strip}_cols_bad"] = sorted(set(int(c) for c in cols))[:40]
            np.save(os.path.join(ROOT, "gpurun_out", f"umma_dbg_got_{item}_{strip}.npy"), got)
            np.save(os.path.join(ROOT, "gpurun_out", f"umma_dbg_ref_{item}_{strip}.npy"), D)
    # 2. output
    got = out.cpu().numpy()
    ref = oevm.pyrdown_cascade(fr[: min(T, 3)], 4)
    res["rel_err_oracle"] = float(np.abs(got[: min(T, 3)] - ref).max() / np.abs(ref).max())
    os.environ["VHR_PYRDOWN_IMPL"] = "stream"
    st = eng.pyrdown(frd, 4).cpu().numpy()
    res["rel_err_stream"] = float(np.abs(got - st).max() / np.abs(st).max())
    res["nan"] = int(np.isnan(got).sum())
    print(json.dumps(res))
    if "--time" in sys.argv:
        Tt = 1800
        frt = eng.synth_clip(vhr.SynthSpec(T=Tt, H=H, W=W, fps=30.0, pulse_hz=1.2, seed=0, clip=0))
        tm = {}
        for impl in ("umma", "stream"):
            os.environ["VHR_PYRDOWN_IMPL"] = impl
            o = eng.pyrdown(frt, 4)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                eng.pyrdown(frt, 4, out=o)
            e1.record()
            torch.cuda.synchronize()
            tm[impl + "_ms"] = e0.elapsed_time(e1) / 5
            tm[impl + "_sum"] = float(o.double().sum().item())
        print(json.dumps(tm))


if __name__ == "__main__":
    main()
