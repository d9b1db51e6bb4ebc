"""One GEMM shape for a dense-sampling ncu capture: python scripts/gemm_prof1.py nt 40960 400 128"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
ops.set_gemm_path(1, min_work=0)
mode, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
C = torch.zeros(M, N, device="cuda")
bias = torch.randn(N, device="cuda") if mode == "nt" else None
for _ in range(3):
    ops.gemm(mode, A, B, C, bias=bias, accumulate=(mode == "tn"))
torch.cuda.synchronize()
