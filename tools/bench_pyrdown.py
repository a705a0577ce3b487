#!/usr/bin/env python
"""pyrDown cascade alone on one 1080p / 1800-frame clip resident in HBM: CUDA-event time per kernel variant and level
count (VHR_PYRDOWN_IMPL = umma | stream).  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import video_heart_rate_b200 as vhr
    eng = vhr.Engine(0)
    T = int(os.environ.get("PYR_T", 1800))
    W, H = int(os.environ.get("PYR_W", 1920)), int(os.environ.get("PYR_H", 1080))
    fr = eng.synth_clip(vhr.SynthSpec(T=T, H=H, W=W, fps=30.0, pulse_hz=1.2, seed=0, clip=0))
    res = {"T": T, "W": W, "H": H}
    for impl in os.environ.get("PYR_IMPLS", "umma,stream").split(","):
        os.environ["VHR_PYRDOWN_IMPL"] = impl
        for L in (2, 4):
            out = eng.pyrdown(fr, L)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                eng.pyrdown(fr, L, out=out)
            e1.record(); torch.cuda.synchronize()
            res[f"{impl}_L{L}_ms"] = e0.elapsed_time(e1) / 5
    print(json.dumps(res))


if __name__ == "__main__":
    main()
