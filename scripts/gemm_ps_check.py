"""Persistent GEMM (gemm_ps.cu): correctness against torch fp64 on the shapes / epilogues it accepts, then CUDA-event timings
next to the per-tile kernel (MFM_PS=0 in a second process)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from factorized_b200.cuda_ops import CudaOps
from emu_ops import keep_mask

ops = CudaOps()
ops.set_gemm_path(1, min_work=0)
torch.manual_seed(0)
dev = "cuda"
lib = ops.lib
ok = True
cases = [("nt", 40960, 400, 128, 0, None), ("nt", 40960, 128, 400, 1, None), ("nn", 40960, 400, 384, 0, None),
         ("nt", 4100, 36, 20, 2, None), ("nn", 5000, 300, 104, 3, None), ("nt", 8192, 64, 128, 2, (0.3, 5)),
         ("nt", 40960, 480, 300, 0, None), ("nt", 4096, 1000, 50, 1, (0.5, 2)), ("nn", 40960, 128, 400, 0, None),
         ("nt", 12800, 352, 5, 0, None)]
if len(sys.argv) > 1 and sys.argv[1] == "time":
    cases = []
for mode, M, N, K, act, drop in cases:
    A = torch.randn(M, K, device=dev)
    B = torch.randn((N, K) if mode == "nt" else (K, N), device=dev) / K ** 0.5
    ldc = (N + 3) // 4 * 4
    Cfull = torch.full((M, ldc), 7.0, device=dev)
    C = Cfull[:, :N]
    bias = torch.randn(N, device=dev)
    rng = torch.tensor([1234, 3], dtype=torch.int64, device=dev)
    n0 = lib.mfm_debug_gemm_ps_count()
    ops.gemm(mode, A, B, C, bias=bias, act=act, drop=drop, rng=rng)
    torch.cuda.synchronize()
    used = lib.mfm_debug_gemm_ps_count() - n0
    ref = A.double() @ (B.double().t() if mode == "nt" else B.double()) + bias.double()
    ref = [ref, ref.clamp_min(0), torch.tanh(ref), torch.sigmoid(ref)][act]
    if drop:
        ref = ref * keep_mask(rng.cpu(), drop[1], drop[0], M, N).to(dev).double() / (1 - drop[0])
    err = float((C.double() - ref).norm() / ref.norm())
    pad_ok = bool((Cfull[:, N:] == 7.0).all())
    good = err < 2e-5 and pad_ok and used == 1
    ok &= good
    print("%s %s %dx%dx%d act %d drop %s: rel err %.2e, padding untouched %s, persistent kernel used %d" % (
        "ok  " if good else "FAIL", mode, M, N, K, act, drop, err, pad_ok, used))
if cases:
    print("ALL OK" if ok else "FAILURES")

TB = 40960
shapes = [("nt", TB, 480, 300, "x_l input proj"), ("nt", TB, 128, 400, "att1_fc1"), ("nt", TB, 400, 128, "att1_fc2"),
          ("nt", TB, 300, 104, "dec fc1"), ("nt", TB, 64, 128, "att2_fc2"), ("nt", TB, 384, 400, "attended -> g1|g2|h2"),
          ("nn", TB, 400, 128, "dcStar = dH1 W11"), ("nn", TB, 128, 400, "dH1 = dL W12"), ("nn", TB, 400, 384, "dAtt = dUcat Wcat"),
          ("nn", TB, 104, 300, "dHd = dXhat W")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode, M, N, K, what in shapes:
    A = torch.randn(M, K, device=dev)
    B = torch.randn((N, K) if mode == "nt" else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    for _ in range(2):
        ops.gemm(mode, A, B, C, bias=bias)
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(mode, A, B, C, bias=bias)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 5
    by = 4.0 * (A.numel() + B.numel() + C.numel())
    print("[MFM_PS=%s] %-3s %6d x %4d x %4d  %-24s %7.1f us  %7.0f GB/s" % (os.environ.get("MFM_PS", "1"), mode, M, N, K, what, ms * 1e3, by / ms / 1e6))
