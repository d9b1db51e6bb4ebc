"""Where does the parity error of the measured configuration sit?  One train-mode step at MOSI batch 2048 after N
training steps, against the oracle (masks / noise / branches replayed), on both math paths: exact-fp32 CUDA cores and
tcgen05 split-bf16.  Prints the largest relative errors."""
import os
import sys
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import factorized_b200 as F
from factorized_b200.train import MFMTrainer
from factorized_b200.configs import best_acc_configs
from factorized_b200.cuda_ops import CudaOps, PATH_SIMT_FP32, PATH_TC_BF16X3
from oracle import mfm_oracle as O
from oracle.rng_replay import train_masks_and_branches

T, B, N = 20, int(os.environ.get("B", 2048)), int(os.environ.get("N", 40))
configs = best_acc_configs(dropout=True)
for path, name in ((PATH_TC_BF16X3, "tcgen05 bf16x3"), (PATH_SIMT_FP32, "cuda-core fp32")):
    CudaOps().set_gemm_path(path)
    torch.manual_seed(123)
    model = F.MFM(*configs).cuda().train()
    tr = MFMTrainer(model, T, B, use_graph=os.environ.get("GRAPH", "0") == "1", seed=123)
    rot = int(os.environ.get("ROTATE", "0"))
    for i in range(N):
        x, y = O.synthetic_batch(configs, T, B, 100 + (i % rot if rot else i))
        tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    x, y = O.synthetic_batch(configs, T, B, 4321)
    lb = tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    masks, br = train_masks_and_branches(tr.eng, tr.rng.cpu())
    noise = [t.detach().cpu().clone() for t in tr.noise]
    del O.RELU_REPLAY_VIOLATIONS[:]
    _, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, train=True, masks=masks, branches=br)
    Pd = OrderedDict((k, v.double()) for k, v in P.items())
    _, _, Go64, _ = O.train_step(Pd, x.double(), y.double(), configs, [t.double() for t in noise], {}, train=True,
                                 masks={k: v.double() for k, v in masks.items()}, branches=br)
    rel = lambda a, b: float((a.detach().cpu().double() - b.double()).norm() / (b.double().norm() + 1e-30))
    rep = {k: rel(tr.G[k], go) for k, go in Go.items() if go is not None}
    rep64 = {k: rel(tr.G[k], Go64[k]) for k, go in Go.items() if go is not None}
    ora = {k: rel(go, Go64[k]) for k, go in Go.items() if go is not None}
    top = sorted(rep, key=rep.get, reverse=True)[:8]
    print("== %s after %d steps (violations %d): rel-L2 error of gradients vs fp32 oracle | vs fp64 oracle | fp32 oracle vs fp64 oracle"
          % (name, N, len(O.RELU_REPLAY_VIOLATIONS)))
    for k in top:
        print("   %-36s %.2e | %.2e | %.2e   |g| = %.3g" % (k, rep[k], rep64[k], ora[k], float(Go[k].norm())))
CudaOps().set_gemm_path(PATH_TC_BF16X3)
