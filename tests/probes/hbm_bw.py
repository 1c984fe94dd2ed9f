import torch
x = torch.empty(2 * 1024**3, dtype=torch.bfloat16, device="cuda")   # 4 GiB
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gb = x.numel() * 2 / 1e9
ms = t(lambda: x.zero_()); print(f"memset  {gb/ms:.2f} TB/s (write only)")
ms = t(lambda: x.fill_(1.5)); print(f"fill    {gb/ms:.2f} TB/s (write only)")
ms = t(lambda: x.sum()); print(f"sum     {gb/ms:.2f} TB/s (read only)")
ms = t(lambda: y.copy_(x)); print(f"copy    {2*gb/ms:.2f} TB/s (read+write)")
ms = t(lambda: torch.add(x, 1.0, out=y)); print(f"add     {2*gb/ms:.2f} TB/s (read+write)")
