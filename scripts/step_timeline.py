"""Milestones of one CUDA-graph replay of the fused training step (device clock stamps on the main stream).
There is no nsys in this image; this is the step-level timeline: where the main stream is at which microsecond, with all
cross-stream overlap (weight-gradient GEMMs on side streams, MMD on the auxiliary stream) included."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import factorized_b200 as F
from factorized_b200.train import MFMTrainer
from oracle import mfm_oracle as O

T, B = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
configs = O.best_acc_configs(dropout=True)
torch.manual_seed(123)
dev = torch.device("cuda", 0)
model = F.MFM(*configs).to(dev).train()
tr = MFMTrainer(model, T, B, head="l1", use_graph=True)
stamps = torch.zeros(64, dtype=torch.int64, device=dev)
tr.eng.stamps = stamps
x = torch.randn(T, B, tr.eng.dm.D, device=dev)
y = torch.randn(B, device=dev)
for _ in range(6):
    tr.step(x, y)
torch.cuda.synchronize()
acc = None
N = 10
for _ in range(N):
    tr.step(x, y)
    torch.cuda.synchronize()
    s = stamps.cpu().double()
    acc = s - s[0] if acc is None else acc + (s - s[0])
acc /= N
names = tr.eng.stamp_names
order = sorted(range(len(names)), key=lambda i: acc[i].item())
prev = 0.0
print("%-26s %10s %10s" % ("milestone (main stream)", "at [us]", "delta [us]"))
for i in order:
    t = acc[i].item() / 1e3
    print("%-26s %10.1f %10.1f" % (names[i], t, t - prev))
    prev = t
