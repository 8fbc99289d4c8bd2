import torch
t = torch.empty(540_672_000, dtype=torch.float32, device='cuda')   # 2.16 GB
for f in (lambda: t.fill_(1.0), lambda: t.zero_(), lambda: torch.cuda.current_stream().synchronize()):
    pass
def bw(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return t.numel() * 4 / (e0.elapsed_time(e1) / n * 1e-3) / 1e9
print("fill_ GB/s", bw(lambda: t.fill_(1.0)))
print("zero_ GB/s", bw(lambda: t.zero_()))
s = torch.empty_like(t)
print("copy GB/s (r+w)", 2 * bw(lambda: s.copy_(t)))
print("sum  GB/s (read)", bw(lambda: t.sum()))
