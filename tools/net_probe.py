"""How fast can PyTorch run the stand-in network of the rollout leg (dev tool, GPU box)?  Times the forward pass at the
rollout batch sizes inside a CUDA graph under a few cuDNN / layout settings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ppnn_b200.nets import ConvResNet
dev = torch.device("cuda")

def bench(tag, B, channels_last=False, benchmark=False, tf32=True, dtype=None):
    torch.backends.cudnn.benchmark = benchmark
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.manual_seed(1)
    net = ConvResNet(True).to(dev).eval()
    u = torch.randn(B, 2, 256, 256, device=dev); p = torch.randn(B, 2, 256, 256, device=dev); mu = torch.rand(B, 1, 1, 1, device=dev)
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    def fwd():
        with torch.no_grad():
            if dtype is not None:
                with torch.autocast("cuda", dtype=dtype):
                    return net(u, mu, p)
            return net(u, mu, p)
    for _ in range(3): y = fwd()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            y = fwd()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    print(f"B={B:3d} {tag:34s} {a.elapsed_time(e) / 20:.3f} ms / forward", flush=True)

for B in (12, 96):
    bench("default (tf32 convs)", B)
    bench("cudnn.benchmark", B, benchmark=True)
    bench("channels_last", B, channels_last=True)
    bench("channels_last + benchmark", B, channels_last=True, benchmark=True)
    bench("strict fp32 (no tf32)", B, tf32=False)
    bench("bf16 autocast (numerics change!)", B, dtype=torch.bfloat16)
