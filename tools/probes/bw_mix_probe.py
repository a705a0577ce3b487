import torch, json
dev = torch.device("cuda:0")
n = 1 << 32   # 4 Gi floats? too big; use 8 GB buffers
a = torch.empty(2 << 30, dtype=torch.float32, device=dev)   # 8 GB
b = torch.empty(2 << 30, dtype=torch.float32, device=dev)
def t(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / best / 1e6
res = {}
res["fill_GBs"] = t(lambda: a.fill_(1.0), a.numel() * 4)
res["copy_GBs"] = t(lambda: b.copy_(a), 2 * a.numel() * 4)
res["sum_GBs"] = t(lambda: a.sum(), a.numel() * 4)
u = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
res["u8_to_f32_GBs"] = t(lambda: torch.add(u, 0, out=a) if False else a.copy_(u), u.numel() * 5)   # 1 B read : 4 B written
print(json.dumps(res))
