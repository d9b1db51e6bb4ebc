"""Clock-stamp timeline of one CTA of the forward recurrence kernel (mfm_debug_set_lstm_trace): per warp and step,
cycles from the step's first stamp: wait start, wait done (accumulator ready), epilogue done, MMA issue done."""
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
torch.cuda.synchronize()
buf = torch.zeros(16 * 32 * 4, dtype=torch.int64, device=dev)
ops.lib.mfm_debug_set_lstm_trace(buf.data_ptr())
ops.lstm_fwd(cells)
torch.cuda.synchronize()
ops.lib.mfm_debug_set_lstm_trace(None)
tr = buf.cpu().view(16, 32, 4)
t0 = int(tr[:, 0, 0][tr[:, 0, 0] > 0].min())
print("cells", hs, "B", B, "(CTA 0 = first CTA of the widest cell); cycles since the CTA's first stamp")
for t in (0, 1, 2, 3, 10, 19):
    print("step %d" % t)
    for w in range(16):
        r = tr[w, t]
        if int(r[0]) == 0:
            continue
        print("  warp %2d  wait_start %7d  acc_ready %7d  epi_done %7d  issued %s" %
              (w, int(r[0]) - t0, int(r[1]) - t0, int(r[2]) - t0 if int(r[2]) else -1, (int(r[3]) - t0) if int(r[3]) else "-"))
