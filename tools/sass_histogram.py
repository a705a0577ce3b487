#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libvhr_b200.so (cuobjdump -sass), written to
profiles/sass_opcodes.txt: the evidence that the data movement is TMA / mbarrier based
(UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops) and which math
pipes each kernel uses (UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, IMMA = mma.sync integer tensor tiles, IDP = dp4a/dp2a, FFMA ...).

    python tools/sass_histogram.py [--top 14]
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-heart-rate_b200", "csrc", "libvhr_b200.so")
KEY = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCIMMA", "UTCHMMA", "UTCBAR", "LDTM", "IMMA", "HMMA", "IDP", "LDGSTS")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=14)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sass_opcodes.txt"))
    args = ap.parse_args()
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = [f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a), {len(kernels)} kernels",
             "# key opcodes: UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk, SYNCS = mbarrier, IMMA = mma.sync integer",
             ""]
    for (name, cnt), pretty in zip(kernels.items(), demangle):
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", pretty)
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)[:150]
        total = sum(cnt.values())
        keys = " ".join(f"{k}={cnt[k]}" for k in KEY if cnt[k])
        top = " ".join(f"{k}:{v}" for k, v in cnt.most_common(args.top))
        lines.append(f"{short}\n    total={total}  {keys}\n    {top}")
    open(args.out, "w").write("\n".join(lines) + "\n")
    print(f"wrote {args.out}: {len(kernels)} kernels")


if __name__ == "__main__":
    sys.exit(main())
