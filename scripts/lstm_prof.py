"""Stand-alone launches of the recurrence kernels at the bench's sizes (MOSI, batch 2048): the 6-cell forward / backward
launch (3 encoder + 3 MFN cells) and the three decoder cells -- for ncu captures and CUDA-event timing."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
T, B = int(os.environ.get("T", 20)), int(os.environ.get("B", 2048))
dev = "cuda"


def fwd_cell(h, gx_steps):
    return dict(T=T, B=B, h=h, gx=torch.randn(gx_steps * B, 4 * h, device=dev), gx_steps=gx_steps,
                bias_rest=torch.randn(4 * h, device=dev) * 0.1 if gx_steps < T else None,
                W=torch.randn(4 * h, h, device=dev) * 0.1, hs=torch.zeros((T + 1) * B, h, device=dev),
                cs=torch.zeros((T + 1) * B, h, device=dev), gates=torch.zeros(T * B, 4 * h, device=dev))


def bwd_cell(c, dec):
    h = c["h"]
    return dict(T=T, B=B, h=h, gates=c["gates"], cs=c["cs"], W=c["W"],
                dh_all=torch.randn(T * B, h, device=dev) * 0.01 if dec else None,
                dh_last=None if dec else torch.randn(B, h, device=dev) * 0.01,
                dc_ext=None if dec else torch.randn(T * B, h, device=dev) * 0.01,
                dG=torch.zeros(T * B, 4 * h, device=dev), dc_scratch=torch.zeros(B, h, device=dev))


enc = [fwd_cell(h, T) for h in (88, 80, 64, 48, 32, 8)]
dec = [[fwd_cell(h, 1)] for h in (104, 24, 24)]
enc_b = [bwd_cell(c, False) for c in enc]
dec_b = [[bwd_cell(c[0], True)] for c in dec]


def timed(name, fn, algo_bytes, flops, reps=5):
    if os.environ.get("LSTM_PROF_ONCE"):          # ncu capture: one launch per shape, no timing
        fn()
        torch.cuda.synchronize()
        return
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print("%-28s %8.1f us   %7.1f GB/s algorithmic   %6.2f TFLOP/s recurrent gate GEMM" %
          (name, ms * 1e3, algo_bytes / ms / 1e6, flops / ms / 1e9))


def fwd_bytes(cells):      # read G_x, write gates + h + c
    return sum(4.0 * c["T"] * c["B"] * (4 * c["h"] * (2 if c["gx_steps"] == c["T"] else 1) + 2 * c["h"]) for c in cells)


def bwd_bytes(cells):      # read gates + c (twice: L2) + external gradients, write dG
    return sum(4.0 * c["T"] * c["B"] * (8 * c["h"] + 3 * c["h"]) for c in cells)


def flops(cells):
    return sum(2.0 * c["T"] * c["B"] * 4 * c["h"] * c["h"] for c in cells)


timed("fwd 6 cells (enc+mfn)", lambda: ops.lstm_fwd(enc), fwd_bytes(enc), flops(enc))
timed("bwd 6 cells (enc+mfn)", lambda: ops.lstm_bwd(enc_b), bwd_bytes(enc_b), flops(enc_b))
for d, db in zip(dec, dec_b):
    timed("fwd decoder h=%d" % d[0]["h"], lambda: ops.lstm_fwd(d), fwd_bytes(d), flops(d))
    timed("bwd decoder h=%d" % d[0]["h"], lambda: ops.lstm_bwd(db), bwd_bytes(db), flops(db))
torch.cuda.synchronize()
