import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factorized_b200.cuda_ops import CudaOps
ops = CudaOps()
T, B = 20, 2048
cells = []
for h in (88, 80, 64, 48, 32, 8):
    cells.append(dict(T=T, B=B, h=h, gx=torch.randn(T * B, 4 * h, device="cuda"), gx_steps=T, bias_rest=None,
                      W=torch.randn(4 * h, h, device="cuda") * 0.1, hs=torch.zeros((T + 1) * B, h, device="cuda"),
                      cs=torch.zeros((T + 1) * B, h, device="cuda"), gates=torch.zeros(T * B, 4 * h, device="cuda")))
for _ in range(2):
    ops.lstm_fwd(cells)
torch.cuda.synchronize()
