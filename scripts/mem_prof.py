"""CUDA-event timing of the MFN memory recurrence, forward and backward, tensor-core and CUDA-core forms (env T, B)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
T, B = int(os.environ.get("T", 20)), int(os.environ.get("B", 2048))
mem, g1, g2 = 64, 128, 128
TB = T * B
dev = "cuda"
r = lambda *s: torch.randn(*s, device=dev) * 0.2
Wg1, Wg2 = r(g1, 400 + mem), r(g2, 400 + mem)
a = dict(T=T, B=B, mem=mem, g1=g1, g2=g2, G1pre=r(TB, g1), G2pre=r(TB, g2), cHat=torch.tanh(r(TB, mem)), W1m=Wg1[:, 400:], W2m=Wg2[:, 400:],
         W12=r(mem, g1), b12=r(mem), W22=r(mem, g2), b22=r(mem), mems=torch.zeros((T + 1) * B, mem, device=dev),
         U1=torch.zeros(TB, g1, device=dev), U2=torch.zeros(TB, g2, device=dev), Gam1=torch.zeros(TB, mem, device=dev),
         Gam2=torch.zeros(TB, mem, device=dev), drop1=None, drop2=None, rng=torch.tensor([5, 2], dtype=torch.int64, device=dev))
b = dict(a)
b.update(scale1=1.0, scale2=1.0, dmem_last=r(B, mem), dU1=torch.zeros(TB, g1, device=dev), dU2=torch.zeros(TB, g2, device=dev),
         dP1=torch.zeros(TB, mem, device=dev), dP2=torch.zeros(TB, mem, device=dev), dPc=torch.zeros(TB, mem, device=dev))
for simt, flags in ((0, 0), (1, 0)):
    ops.lib.mfm_debug_mem_force_simt(simt)
    for name, fn, arg in (("fwd", ops.mfn_mem_fwd, a), ("bwd", ops.mfn_mem_bwd, b)):
        for _ in range(3):
            fn(arg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(arg)
        e1.record()
        torch.cuda.synchronize()
        print("%s flags %d %s T=%d B=%d: %.1f us" % ("cuda-core  " if simt else "tensor-core", flags, name, T, B, e0.elapsed_time(e1) * 100))
