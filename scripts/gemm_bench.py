"""Time the step's main GEMM shapes (CUDA events); prints TFLOP/s and algorithmic GB/s (4*(|A|+|B|+|C|) per launch).
usage: gemm_bench.py [path ...]   (0 = fp32 CUDA cores, 1 = tcgen05 split-bf16 x3, 2 = tcgen05 plain bf16)
env MFM_TCP=0 forces the register-prefetch kernel (gemm_tc.cu); MFM_TCP_BK=16|32 picks the pipelined kernel's K chunk."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps

ops = CudaOps()
TB = 40960
shapes = [("nt", TB, 480, 300, "x_l input proj (enc+mfn)"), ("nt", TB, 128, 400, "att1_fc1"), ("nt", TB, 400, 128, "att1_fc2"),
          ("nt", TB, 300, 104, "dec fc1"), ("nt", TB, 64, 128, "att2_fc2"), ("nt", 2048, 2048, 80, "mmd pair matrix"),
          ("nn", TB, 400, 128, "dcStar = dH1 W11"), ("nn", TB, 128, 400, "dH1 = dL W12"), ("nn", TB, 400, 384, "dAtt = dUcat Wcat"),
          ("tn", 128, 400, TB, "dW11 = dH1^T cStar"), ("tn", 400, 128, TB, "dW12 = dL^T H1"), ("tn", 352, 300, TB, "dW_ih mfn_l"),
          ("tn", 416, 104, TB, "dW dec_l"), ("tn", 352, 88, TB, "dW_hh mfn_l")]
paths = [int(v) for v in sys.argv[1:]] or [1]
tag = "tcp=%s bk=%s" % (os.environ.get("MFM_TCP", "1"), os.environ.get("MFM_TCP_BK", "16"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for path in paths:
    ops.set_gemm_path(path, min_work=0)
    for mode, M, N, K, what in shapes:
        A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
        B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        acc = mode == "tn"
        for _ in range(2):
            ops.gemm(mode, A, B, C, accumulate=acc)
        tot = 0.0
        for _ in range(5):
            flush.zero_()                      # operands start cold in L2, as they do inside the training step
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(mode, A, B, C, accumulate=acc)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / 5
        fl = 2.0 * M * N * K
        by = 4.0 * (A.numel() + B.numel() + C.numel())
        print("[%s] path %d %-3s %6d x %4d x %6d  %-24s %8.3f ms  %7.1f TFLOP/s  %7.0f GB/s" % (tag, path, mode, M, N, K, what, ms, fl / ms / 1e9, by / ms / 1e6))
