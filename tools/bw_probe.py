import torch, time
x = torch.empty(1 << 30, dtype=torch.float16, device="cuda")  # 2 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms = t(lambda: x.fill_(1.0)); print(f"fill (write only): {2**31/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy (r+w): {2*2**31/ms/1e6:.0f} GB/s")
ms = t(lambda: x.sum()); print(f"sum (read only): {2**31/ms/1e6:.0f} GB/s")
