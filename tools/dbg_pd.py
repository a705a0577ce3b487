import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import video_heart_rate_b200 as vhr
from oracle import evm as oevm
eng = vhr.Engine(0)
for (T,H,W,L) in [(3,144,256,3),(2,64,1920,3),(2,270,1920,4),(2,1080,1920,4),(700,40,64,2),(400,70,96,4),(40,1080,1920,4)]:
    rng = np.random.default_rng(1)
    fr = rng.integers(0,256,(T,H,W,3),dtype=np.uint8)
    got = eng.pyrdown(torch.as_tensor(fr, device=eng.tdev), L)
    torch.cuda.synchronize()
    ref = oevm.pyrdown_cascade(fr[:4], L)
    err = np.abs(got[:4].cpu().numpy()-ref).max()/np.abs(ref).max()
    print(T,H,W,L,'rel err',err, flush=True)
