"""The ``MFM`` of the reference's classification script, /root/reference/mfm_mosi_acc.py:311-394 -- the one script that carries
its own copy of the model classes instead of importing ``mfm_model``.  Same network as ``mfm_model.MFM`` with two differences:
``output_dim`` is hard-coded to 2 (:331, :211) and ``forward`` returns ``(zl, za, zv, zy, x_l_hat, x_a_hat, x_v_hat, y_hat)``
(:394) -- the training loop applies ``loss_MMD`` to the latents itself (:441), so the latents must carry gradient.

Here: ``MFM.forward`` runs the same kernel schedule without the in-step MMD and hands the four latents out as differentiable
outputs; whatever the caller does with them (``factorized_b200.functional.loss_MMD`` is the drop-in for the script's own
function) comes back as ``d_latents`` into ``Engine.backward``.  The fused path for this script is ``train_mfm(..., head="ce")``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import mfm_model as M
from .mfm_model import _ops, _require_cuda, eng_generation


class _LatentOutFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        T, B, _ = x.shape
        eng = module._engine(T, B, x.device)
        P = OrderedDict(zip(module._param_names, params))
        rng = module._rng_state(x.device)
        if module.training:
            _ops().rng_tick(rng)
        eng.want_mmd = False                               # the caller regularises the latents itself
        try:
            out = eng.forward(P, x, [None] * 4, train=module.training, rng=rng)
        finally:
            eng.want_mmd = True
        ctx.eng, ctx.P, ctx.gen = eng, P, eng_generation(eng, bump=True)
        dm = eng.dm
        return (out["zl"].clone(), out["za"].clone(), out["zv"].clone(), out["zy"].clone(),
                out["x_l_hat"].reshape(T, B, dm.d[0]).clone(), out["x_a_hat"].reshape(T, B, dm.d[1]).clone(),
                out["x_v_hat"].reshape(T, B, dm.d[2]).clone(), out["y_hat"].clone())

    @staticmethod
    def backward(ctx, dzl, dza, dzv, dzy, dxl, dxa, dxv, dy):
        eng, P = ctx.eng, ctx.P
        if eng_generation(eng) != ctx.gen:
            raise RuntimeError("MFM backward called after another forward of the same (T,B) shape overwrote the kernel "
                               "workspace; run backward before the next forward")
        dm = eng.dm
        TB = dm.T * dm.B

        def dense(g, shape):
            if g is None:
                return torch.zeros(shape, dtype=torch.float32, device=eng.device)
            return g.contiguous().float().view(shape)
        dX = [dense(dxl, (TB, dm.d[0])), dense(dxa, (TB, dm.d[1])), dense(dxv, (TB, dm.d[2]))]
        dY = dense(dy, (dm.B, dm.out))
        dlat = [dense(dzl, (dm.B, dm.z[0])), dense(dza, (dm.B, dm.z[1])), dense(dzv, (dm.B, dm.z[2])), dense(dzy, (dm.B, dm.zy))]
        sizes = [p.numel() for p in P.values()]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=eng.device)
        _ops().zero(flat)
        G, o = OrderedDict(), 0
        for (k, p), n in zip(P.items(), sizes):
            G[k] = flat[o:o + n].view(p.shape)
            o += n
        eng.backward(P, G, dX, dY, 0.0, d_latents=dlat)
        return (None, None) + tuple(G.values())


class MFM(M.MFM):
    """mfm_mosi_acc.py:311-394.  ``forward(x[T,N,D]) -> (zl, za, zv, zy, x_l_hat, x_a_hat, x_v_hat, y_hat)``, ``y_hat`` [N, 2]."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        config = dict(config)
        config["output_dim"] = 2                           # hard-coded there (:331; the MFN's unused head too, :211)
        super(MFM, self).__init__(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)

    def forward(self, x):
        _require_cuda(x, "MFM.forward")
        if x.dim() != 3:
            raise ValueError("MFM.forward expects x[T,N,D]")
        if x.requires_grad:
            raise RuntimeError("MFM.forward: gradient w.r.t. the input is not provided (the reference never asks for it)")
        pd = dict(self.named_parameters())
        return _LatentOutFunction.apply(self, x.contiguous().float(), *[pd[k] for k in self._param_names])
