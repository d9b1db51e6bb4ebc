"""Clock-stamp timeline of CTA 0 of the forward recurrence kernel.  Needs a library built with -DWS_DEBUG=1:
   touch factorized_b200/csrc/lstm_ws.cu && make -C factorized_b200/csrc EXTRA=-DWS_DEBUG=1
Per step: compute warps -- wait start / accumulator ready / cell update done for chain A and B; issuer warp -- operand
ready / GEMM issued for chain A and B.  Cycles since the first stamp of the CTA."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
T, B = 20, int(os.environ.get("B", 2048))
dev = "cuda"
hs = [int(v) for v in os.environ.get("HS", "88,80,64,48,32,8").split(",")]
cells = [dict(T=T, B=B, h=h, gx=torch.randn(T * B, 4 * h, device=dev), gx_steps=T, bias_rest=None,
              W=torch.randn(4 * h, h, device=dev) * 0.1, hs=torch.zeros((T + 1) * B, h, device=dev),
              cs=torch.zeros((T + 1) * B, h, device=dev), gates=torch.zeros(T * B, 4 * h, device=dev)) for h in hs]
ops.lstm_fwd(cells)
ops.lstm_fwd(cells)
torch.cuda.synchronize()
host = (ctypes.c_longlong * (17 * 32 * 8))()
rc = ops.lib.mfm_debug_set_lstm_trace(ctypes.cast(host, ctypes.c_void_p))
assert rc == 0, "library was not built with -DWS_DEBUG=1 (rc %d)" % rc
tr = torch.tensor(list(host), dtype=torch.int64).view(17, 32, 8)
nz = tr[tr > 0]
t0 = int(nz.min())
print("cells", hs, "B", B, "-- CTA 0 (first CTA of the widest cell); cycles since the CTA's first stamp")
for t in (0, 1, 2, 3, 10, 11):
    print("step %d" % t)
    r = tr[16, t]
    print("  issuer   A: ready %7d issued %7d   B: ready %7d issued %7d" % tuple(int(v) - t0 if int(v) else -1 for v in r[:4]))
    for w in range(16):
        r = tr[w, t]
        if int(r[0]) == 0:
            continue
        print("  warp %2d  A: wait %7d acc %7d done %7d   B: wait %7d acc %7d done %7d   A first tcgen05.ld: issue %7d data %7d" %
              tuple([w] + [int(v) - t0 if int(v) else -1 for v in r[:8]]))
