"""A few representative tcgen05 GEMM launches for an ncu capture (one warm-up + one profiled launch per shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
ops.set_gemm_path(1, min_work=0)
TB = 40960
shapes = (("nt", TB, 128, 400), ("nt", TB, 400, 128), ("tn", 128, 400, TB), ("nn", TB, 400, 384))
bufs = []
for mode, M, N, K in shapes:
    A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
    B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    bufs.append((A, B, C))
for rep in range(2):
    for (mode, M, N, K), (A, B, C) in zip(shapes, bufs):
        ops.gemm(mode, A, B, C, accumulate=(mode == "tn"))
torch.cuda.synchronize()
