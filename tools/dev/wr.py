import torch
dev = torch.device("cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n = 1 << 28   # 1 GiB of fp32
x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
t = timed(lambda: x.fill_(1.5)); print(f"fill   {4*n/t/1e6:.0f} GB/s written")
t = timed(lambda: x.zero_()); print(f"memset {4*n/t/1e6:.0f} GB/s written")
t = timed(lambda: y.copy_(x)); print(f"copy   {8*n/t/1e6:.0f} GB/s read+written")
t = timed(lambda: x.sum()); print(f"sum    {4*n/t/1e6:.0f} GB/s read")
t = timed(lambda: torch.add(x, 1.0, out=y)); print(f"add    {8*n/t/1e6:.0f} GB/s read+written")
