#!/usr/bin/env python
"""Time csrc/pyrdown_umma.cu on one 1080p / 1800-frame clip under the measurement hooks in VHR_UMMA_MODES (comma list)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import video_heart_rate_b200 as vhr
eng = vhr.Engine(0)
T = int(os.environ.get("PYR_T", 1800)); W = int(os.environ.get("PYR_W", 1920)); H = int(os.environ.get("PYR_H", 1080))
fr = eng.synth_clip(vhr.SynthSpec(T=T, H=H, W=W, fps=30.0, pulse_hz=1.2, seed=0, clip=0))
res = {}
for rep, mode in enumerate(os.environ.get("VHR_UMMA_MODES", "0").split(",")):
    os.environ["VHR_UMMA_MODE"] = mode
    o = eng.pyrdown(fr, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.pyrdown(fr, 4, out=o)
    e1.record(); torch.cuda.synchronize()
    res[f"mode{mode}_{rep}"] = round(e0.elapsed_time(e1) / 5, 4)
print(json.dumps(res))
