"""Step timeline of the tensor-core memory recurrence (library built with -DMW_DEBUG=1 for mem_ws.cu): clock stamps of CTA 0."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
T, B, mem, g1, g2 = 20, 2048, 64, 128, 128
TB = T * B
dev = "cuda"
r = lambda *s: torch.randn(*s, device=dev) * 0.2
Wg1, Wg2 = r(g1, 400 + mem), r(g2, 400 + mem)
drop = (0.5, 3) if os.environ.get("DROP") else None
a = dict(T=T, B=B, mem=mem, g1=g1, g2=g2, G1pre=r(TB, g1), G2pre=r(TB, g2), cHat=torch.tanh(r(TB, mem)), W1m=Wg1[:, 400:], W2m=Wg2[:, 400:],
         W12=r(mem, g1), b12=r(mem), W22=r(mem, g2), b22=r(mem), mems=torch.zeros((T + 1) * B, mem, device=dev),
         U1=torch.zeros(TB, g1, device=dev), U2=torch.zeros(TB, g2, device=dev), Gam1=torch.zeros(TB, mem, device=dev),
         Gam2=torch.zeros(TB, mem, device=dev), drop1=drop, drop2=drop, rng=torch.tensor([5, 2], dtype=torch.int64, device=dev))
for _ in range(3):
    ops.mfn_mem_fwd(a)
torch.cuda.synchronize()
buf = np.zeros(32 * 16, dtype=np.int64)
assert ops.lib.mfm_debug_mem_ws_trace(buf.ctypes.data) == 0, "build mem_ws.cu with -DMW_DEBUG=1"
t = buf.reshape(32, 16)[:T].astype(np.float64)
t0 = t[0, 4]
names = ["iss:A woke", "iss:A issued", "iss:B woke", "iss:B issued", "cw:wait A", "cw:A woke", "cw:A ld done", "cw:stage1 done", "cw:published",
         "cw:wait B", "cw:B woke", "cw:stage2 done", "cw:published"]
print("step " + " ".join("%14s" % n for n in names))
for s in range(2, 8):
    print("%4d " % s + " ".join("%14.0f" % (t[s, k] - t[s, 4]) for k in range(13)) + "   | step length %.0f" % (t[s + 1, 4] - t[s, 4]))
