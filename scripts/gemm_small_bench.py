"""Batch-row GEMMs of the step (M = B = 2048: heads, factor MLPs, decoder step 0) on the tcgen05 and CUDA-core paths;
times a chain of 20 identical launches inside a CUDA graph (what the step sees: launch gap + kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
B = 2048
shapes = [("nt", B, 32, 200, "zy from h"), ("nt", B, 32, 64, "zy from mem"), ("nt", B, 88, 32, "zl->fl fc1"), ("nt", B, 88, 88, "fl fc2"),
          ("nt", B, 16, 32, "zy->fy"), ("nt", B, 416, 104, "dec step-0 proj"), ("nn", B, 104, 416, "dEMB"), ("nn", B, 32, 88, "dz"),
          ("nt", B, 32, 32, "enc fc1"), ("nn", B, 200, 32, "dHlast"), ("tn", 32, 200, B, "dWzy")]
for path, name in ((1, "tcgen05"), (0, "cuda-core")):
    ops.set_gemm_path(path, min_work=(1 << 20))
    for mode, M, N, K, what in shapes:
        A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
        Bm = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        bias = torch.randn(N, device="cuda") if mode == "nt" else None
        acc = mode == "tn"
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ops.gemm(mode, A, Bm, C, bias=bias, accumulate=acc)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    ops.gemm(mode, A, Bm, C, bias=bias, accumulate=acc)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        print("%-9s %-3s %5d x %4d x %5d  %-16s %6.2f us per launch" % (name, mode, M, N, K, what, e0.elapsed_time(e1) * 10))
