"""The module-level functions of the reference's ``mfm_model`` -- ``compute_kernel`` (:14-23), ``loss_MMD`` (:25-34),
``loss_KLD`` (:36-38) -- on the CUDA primitives, with autograd, for scripts and models that call them directly (``seq2seq`` and
``basic_missing`` do).  ``MFM.forward`` does not come through here: its schedule computes the same statistics inside the step."""
from __future__ import annotations

import torch

from .mfm_model import _ops, _require_cuda


class _MMDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, g):
        out = torch.zeros(1, dtype=torch.float32, device=z.device)
        _ops().mmd_fwd(z, g, out)
        ctx.save_for_backward(z, g)
        return out[0]

    @staticmethod
    def backward(ctx, d):
        z, g = ctx.saved_tensors
        dz = torch.zeros_like(z)
        _ops().mmd_bwd(z, g, 1.0, dz, scale_dev=d.contiguous().view(1).float())
        return dz, None


def loss_MMD(zy):
    """mfm_model.py:25-34: biased MMD (diagonal included) between ``zy`` [n, dim] and a standard Gaussian sample of the same
    shape, drawn like the reference on the CPU default generator and copied to the device (:26-29)."""
    _require_cuda(zy, "loss_MMD")
    g = torch.randn(zy.size()).to(zy.device)
    return _MMDFn.apply(zy.contiguous().float(), g)


class _KLDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        out = torch.zeros(1, dtype=torch.float32, device=mu.device)
        _ops().kld_fwd(mu, logvar, out)
        ctx.save_for_backward(mu, logvar)
        return out[0]

    @staticmethod
    def backward(ctx, d):
        mu, logvar = ctx.saved_tensors
        dmu, dlv = torch.zeros_like(mu), torch.zeros_like(logvar)
        _ops().kld_bwd(mu, logvar, 1.0, dmu, dlv, d.contiguous().view(1).float())
        return dmu, dlv


def loss_KLD(mu, logvar):
    """mfm_model.py:36-38: -0.5 * sum(1 + logvar - mu^2 - exp(logvar))."""
    _require_cuda(mu, "loss_KLD")
    return _KLDFn.apply(mu.contiguous().float(), logvar.contiguous().float())


def compute_kernel(x, y):
    """mfm_model.py:14-23: K[i, j] = exp(-mean_k((x_ik - y_jk)^2) / dim), as |x|^2 + |y|^2 - 2 x y^T on the tensor cores -- no
    [n, m, dim] tensor.  Forward only (loss_MMD carries the gradient of the statistic the reference builds from it)."""
    _require_cuda(x, "compute_kernel")
    if x.requires_grad or y.requires_grad:
        raise RuntimeError("compute_kernel: forward only; use loss_MMD for a differentiable statistic")
    ops = _ops()
    x, y = x.contiguous().float(), y.contiguous().float()
    nx = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
    ny = torch.zeros(y.shape[0], dtype=torch.float32, device=x.device)
    K = torch.zeros(x.shape[0], y.shape[0], dtype=torch.float32, device=x.device)
    ops.rownorm2(x, nx)
    ops.rownorm2(y, ny)
    ops.gemm("nt", x, y, K)
    ops.mmd_kexp(K, nx, ny, x.shape[1], 0.0, torch.zeros(1, dtype=torch.float32, device=x.device))
    return K
