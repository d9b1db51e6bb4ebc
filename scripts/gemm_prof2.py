import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps(); ops.set_gemm_path(1, min_work=0)
TB = 40960
for mode, M, N, K in (("nn", TB, 400, 128),):
    A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
    B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    for _ in range(2):
        ops.gemm(mode, A, B, C)
torch.cuda.synchronize()
