"""Poor man's timeline of the pipelined GEMM (no nsys in this image): every CTA records clock stamps of its producer /
converter / MMA roles (mfm_debug_set_gemm_trace); this prints the wave structure and the per-chunk stage latencies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from factorized_b200.cuda_ops import CudaOps

ops = CudaOps()
ops.set_gemm_path(1, min_work=0)
TB = 40960
W = 4 + 6 * 32
shapes = [("nt", TB, 128, 400), ("nt", TB, 400, 128), ("tn", 128, 400, TB), ("nn", TB, 400, 384)]
if len(sys.argv) > 1:
    shapes = [shapes[int(v)] for v in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode, M, N, K in shapes:
    A = torch.randn((M, K) if mode != "tn" else (K, M), device="cuda")
    B = torch.randn((N, K) if mode == "nt" else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    acc = mode == "tn"
    ops.gemm(mode, A, B, C, accumulate=acc)
    buf = torch.zeros(4096 * W, dtype=torch.int64, device="cuda")
    flush.zero_()
    torch.cuda.synchronize()
    ops.lib.mfm_debug_set_gemm_trace(buf.data_ptr(), buf.numel() * 8)
    ops.gemm(mode, A, B, C, accumulate=acc)
    torch.cuda.synchronize()
    ops.lib.mfm_debug_set_gemm_trace(None, 0)
    t = buf.cpu().numpy().reshape(-1, W)
    t = t[t[:, 0] > 0]
    n = len(t)
    g0 = t[:, 0].min()
    start, end = (t[:, 0] - g0) / 1e3, (t[:, 3] - g0) / 1e3
    nch = t[:, 2]
    print("== %s %dx%dx%d: %d CTAs, chunks/CTA %d..%d, kernel span %.1f us" % (mode, M, N, K, n, nch.min(), nch.max(), end.max()))
    print("   CTA start (us): p0 %.1f p50 %.1f p90 %.1f max %.1f | CTA duration (us): p10 %.1f p50 %.1f p90 %.1f max %.1f" % (
        start.min(), np.percentile(start, 50), np.percentile(start, 90), start.max(),
        *np.percentile(end - start, [10, 50, 90]), (end - start).max()))
    late = start > 1.0
    print("   CTAs starting after 1 us: %d (their start p50 %.1f us)" % (late.sum(), np.percentile(start[late], 50) if late.any() else 0))
    nc = int(min(nch.max(), 32))
    st = t[:, 4:4 + 6 * nc].reshape(n, nc, 6).astype(np.float64)
    first = ~late
    def med(x):
        return np.median(x[first], axis=0)
    issue, fullw, arr, mmaw, commit, pfreew = [med(st[:, :, i]) for i in range(6)]
    print("   chunk: issue  full  conv_done  mma_start  mma_commit  (cycles since CTA start, median over first-wave CTAs)")
    for c in list(range(min(nc, 8))) + list(range(max(8, nc - 3), nc)):
        print("   %3d  %7.0f %7.0f %7.0f %7.0f %7.0f   tma-lat %6.0f  convert %5.0f  conv->mma %5.0f" % (
            c, issue[c], fullw[c], arr[c], mmaw[c], commit[c], fullw[c] - issue[c], arr[c] - fullw[c], mmaw[c] - arr[c]))
    if nch.max() < 31:
        ep = t[:, 4 + 6 * 31:4 + 6 * 31 + 3].astype(np.float64)[first]
        last_commit = st[first, nc - 1, 4]
        print("   after the last commit (cycles, median): accumulator ready +%.0f, epilogue %.0f, final barrier +%.0f" % (
            np.median(ep[:, 0] - last_commit), np.median(ep[:, 1] - ep[:, 0]), np.median(ep[:, 2] - ep[:, 1])))
    if nch.max() < 30:
        e6 = t[:, 4 + 6 * 30:4 + 6 * 30 + 6].astype(np.float64)[first]
        d = np.median(e6[:, 1:] - e6[:, :-1], axis=0)
        print("   epilogue of warp 0 (cycles, median): first tcgen05.ld %.0f, to scratch %.0f, first block stored %.0f, block 2 %.0f, block 3 %.0f" % tuple(d))
    if nc > 4:
        per_chunk = (commit[nc - 1] - commit[2]) / (nc - 3)
        print("   steady state: %.0f cycles per chunk per CTA; epilogue+teardown %.1f us" % (
            per_chunk, np.median((end - start)[first]) - commit[nc - 1] / 1.9e3))
