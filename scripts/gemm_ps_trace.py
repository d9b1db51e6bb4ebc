"""Role timeline of the persistent GEMM (library built with make EXTRA=-DPS_DEBUG=1): per chunk the producer issue, converter
wake-up / arrival, MMA wake-up / commit; per tile the epilogue's wake-up, accumulator release and last store."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from factorized_b200.cuda_ops import CudaOps

ops = CudaOps()
ops.set_gemm_path(1, min_work=0)
TB = 40960
NC, NT_ = 64, 16
W = 4 + 6 * NC + 4 * NT_
shapes = [("nt", TB, 400, 128), ("nt", TB, 128, 400), ("nt", TB, 480, 300), ("nt", TB, 64, 128)]
if len(sys.argv) > 1:
    shapes = [shapes[int(v)] for v in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode, M, N, K in shapes:
    A = torch.randn(M, K, device="cuda")
    B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    ops.gemm(mode, A, B, C)
    buf = torch.zeros(148 * W, dtype=torch.int64, device="cuda")
    flush.zero_()
    torch.cuda.synchronize()
    assert ops.lib.mfm_debug_set_gemm_ps_trace(buf.data_ptr()) == 0, "library not built with -DPS_DEBUG=1"
    ops.gemm(mode, A, B, C)
    torch.cuda.synchronize()
    ops.lib.mfm_debug_set_gemm_ps_trace(None)
    t = buf.cpu().numpy().reshape(148, W).astype(np.float64)
    t = t[t[:, 0] > 0]
    g0 = t[:, 0].min()
    print("== %s %dx%dx%d: %d CTAs; start spread %.1f us; CTA life p50 %.1f us max %.1f us (%.0f cycles p50)" % (
        mode, M, N, K, len(t), (t[:, 0] - g0).max() / 1e3, np.median(t[:, 2] - t[:, 0]) / 1e3, (t[:, 2] - t[:, 0]).max() / 1e3, np.median(t[:, 3])))
    ch = np.median(t[:, 4:4 + 6 * NC].reshape(-1, NC, 6), axis=0)
    nck = (K + 15) // 16
    print("   chunk  issue   conv0-woke conv0-arr conv7-arr  mma-woke  mma-commit | load-lat conv  conv->mma  mma")
    show = list(range(0, min(NC, 3 * nck + 2)))
    for g in show:
        i, w0, a0, a7, mw, mc = ch[g]
        if mc == 0:
            break
        print("   %3d %8.0f %8.0f %8.0f %8.0f %8.0f %8.0f   | %6.0f %5.0f %6.0f %5.0f" % (g, i, w0, a0, a7, mw, mc, w0 - i, max(a0, a7) - w0, mw - max(a0, a7), mc - mw))
    valid = ch[:, 5] > 0
    n = int(valid.sum())
    if n > 12:
        print("   steady state: %.0f cycles per chunk (chunks 8..%d)" % ((ch[n - 1, 5] - ch[8, 5]) / (n - 9), n - 1))
    tl = np.median(t[:, 4 + 6 * NC:].reshape(-1, NT_, 4), axis=0)
    print("   tile  mma-got-acc  epi-woke  acc-released  last-store-issued")
    for i in range(NT_):
        if tl[i, 0] == 0:
            break
        print("   %3d %10.0f %10.0f %10.0f %10.0f" % (i, tl[i, 3], tl[i, 0], tl[i, 1], tl[i, 2]))
