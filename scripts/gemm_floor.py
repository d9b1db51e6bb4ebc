import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps(); ops.set_gemm_path(1, min_work=0)
for mode, M, N, K in (("nt", 40960, 128, 400), ("nt", 40960, 128, 32), ("nt", 2048, 128, 32), ("nt", 2048, 128, 400), ("nt", 2048, 16, 16), ("nt", 128, 16, 16), ("tn", 128, 128, 2048), ("tn", 8, 8, 2048)):
    A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
    B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    acc = mode == "tn"
    for _ in range(3): ops.gemm(mode, A, B, C, accumulate=acc)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): ops.gemm(mode, A, B, C, accumulate=acc)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print("dbg=%s %-3s %6d x %4d x %6d  %7.2f us per launch (graph of 20)" % (os.environ.get("MFM_TC_DEBUG", "0"), mode, M, N, K, e0.elapsed_time(e1) * 1000 / 20))
