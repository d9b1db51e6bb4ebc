"""Kernel-to-kernel latency inside a CUDA graph on this GPU: a chain of N dependent tiny kernels (zero of 4 floats), and the same
chain alternating between two streams (fork/join per link), timed with CUDA events over graph replays."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
x = torch.zeros(1024, device="cuda")
N = 200
def chain():
    for i in range(N):
        ops.zero(x[:4])
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        chain()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("graph chain of %d tiny kernels: %.2f us per kernel" % (N, e0.elapsed_time(e1) * 1e3 / (10 * N)))
    big = torch.zeros(64 << 20, device="cuda")
    def chain2():
        for i in range(20):
            ops.zero(big)
    chain2()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        chain2()
    g2.replay()
    e0.record()
    for _ in range(5):
        g2.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / 100
    print("graph chain of 20 x zero(256 MB): %.1f us per kernel = %.0f GB/s" % (t, 256e6 * 1.048576 / t / 1e3))
