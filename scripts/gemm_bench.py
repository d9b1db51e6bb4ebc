"""Time the step's main GEMM shapes on each math path (CUDA events); prints TFLOP/s and effective GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps

ops = CudaOps()
TB = 40960
shapes = [("nt", TB, 480, 300, "x_l input proj (enc+mfn)"), ("nt", TB, 128, 400, "att1_fc1"), ("nt", TB, 400, 128, "att1_fc2"),
          ("nt", TB, 300, 104, "dec fc1"), ("nn", TB, 400, 128, "dcStar = dH1 W11"), ("nn", TB, 128, 400, "dH1 = dL W12"),
          ("tn", 128, 400, TB, "dW11 = dH1^T cStar"), ("tn", 352, 300, TB, "dW_ih mfn_l"), ("tn", 416, 104, TB, "dW dec_l")]
for path in (0, 1, 2):
    ops.set_gemm_path(path, min_work=0)
    for mode, M, N, K, what in shapes:
        A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
        B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        acc = mode == "tn"
        for _ in range(2):
            ops.gemm(mode, A, B, C, accumulate=acc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gemm(mode, A, B, C, accumulate=acc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * M * N * K
        by = 4.0 * (A.numel() + B.numel() + C.numel())
        print("path %d %-3s %6d x %4d x %6d  %-24s %8.3f ms  %7.1f TFLOP/s  %7.0f GB/s" % (path, mode, M, N, K, what, ms, fl / ms / 1e9, by / ms / 1e6))
