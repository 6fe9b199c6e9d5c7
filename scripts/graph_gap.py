"""Cost of a kernel boundary inside a CUDA graph on this GPU: a chain of n dependent tiny kernels, per-kernel time."""
import torch
dev = torch.device("cuda", 0)
x = torch.zeros(32, device=dev)
s = torch.cuda.Stream()
for n in (1, 10, 30, 60):
    with torch.cuda.stream(s):
        for _ in range(3):
            for _ in range(n):
                x.add_(1.0)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(n):
            x.add_(1.0)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 200 * 1e3
    print(f"chain of {n:3d} tiny kernels: {t:8.2f} us per replay = {t / n:6.2f} us per kernel")
