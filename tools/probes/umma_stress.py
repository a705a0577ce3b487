#!/usr/bin/env python
"""Soak test of csrc/pyrdown_umma.cu: random eligible shapes and clip lengths, random frames generated on the device,
every result compared with the streaming / generic kernel (which the oracle tests hold separately), repeated launches of
the same input compared bit for bit.  Runs for UMMA_STRESS_SECONDS (default 60); prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import video_heart_rate_b200 as vhr

eng = vhr.Engine(0)
g = torch.Generator(device=eng.tdev); g.manual_seed(1)
shapes = [(1080, 1920), (720, 1280), (480, 640), (64, 160), (90, 720), (135, 240), (1081, 160), (2160, 320), (270, 480), (1440, 2560),
          (99, 400), (777, 80 * 7), (540, 960), (360, 640)]
t_end = time.time() + float(os.environ.get("UMMA_STRESS_SECONDS", 60))
n = 0; worst = 0.0; frames_total = 0
import random
random.seed(3)
while time.time() < t_end:
    H, W = random.choice(shapes)
    budget = 600_000_000 // (H * W * 3)
    T = max(1, min(budget, random.choice([1, 2, 3, 7, 50, 149, 300, 1000])))
    fr = torch.randint(0, 256, (T, H, W, 3), dtype=torch.uint8, device=eng.tdev, generator=g)
    os.environ["VHR_PYRDOWN_IMPL"] = "umma"
    a = eng.pyrdown(fr, 4)
    b = eng.pyrdown(fr, 4)
    os.environ["VHR_PYRDOWN_IMPL"] = "stream"
    ref = eng.pyrdown(fr, 4)
    torch.cuda.synchronize()
    assert torch.equal(a, b), ("nondeterministic", H, W, T)
    err = float((a - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, ("mismatch", H, W, T, err)
    worst = max(worst, err); n += 1; frames_total += T
print(json.dumps({"launch_triples": n, "frames": frames_total, "worst_rel_err_vs_cuda_core_kernel": worst, "shapes": len(shapes)}))
