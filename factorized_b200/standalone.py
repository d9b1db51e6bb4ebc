"""Standalone forwards (with autograd) of encoderLSTM / decoderLSTM / MFN on the CUDA primitive set.

MFM.forward does not go through these (it runs the whole step as one schedule, engine.Engine); they
exist so the three building-block modules keep working on their own, as in the reference where the
ablation models re-wire them (mfm_model.py:201-467).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import engine as E


def _ops():
    from .mfm_model import _ops as f
    return f()


def _rows2d(x: torch.Tensor) -> torch.Tensor:
    """[T,N,d] (possibly a last-axis slice of a wider tensor, mfm_model.py:523-525) -> 2-D [T*N,d] view."""
    T, N, d = x.shape
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(2) == 1 and x.stride(0) == N * x.stride(1) and x.stride(1) >= d:
        return x.as_strided((T * N, d), (x.stride(1), 1))
    return x.contiguous().view(T * N, d)


def _new(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, fw, fb):
        ops = _ops()
        T, N, d = x.shape
        h = w_hh.shape[1]
        dev = x.device
        X2 = _rows2d(x)
        Gx = _new((T * N, 4 * h), dev)
        ops.gemm("nt", X2, w_ih, Gx, bias=b_ih, bias2=b_hh)
        hs, cs, gates = _new(((T + 1) * N, h), dev), _new(((T + 1) * N, h), dev), _new((T * N, 4 * h), dev)
        ops.lstm_fwd([dict(T=T, B=N, h=h, gx=Gx, gx_steps=T, bias_rest=None, W=w_hh, hs=hs, cs=cs, gates=gates)])
        z = _new((N, h), dev)
        ops.gemm("nt", hs[T * N:], fw, z, bias=fb)
        ctx.save_for_backward(X2, w_ih, w_hh, fw, hs, cs, gates)
        ctx.dims = (T, N, d, h)
        ctx.x_grad = x.requires_grad
        return z

    @staticmethod
    def backward(ctx, dz):
        ops = _ops()
        X2, w_ih, w_hh, fw, hs, cs, gates = ctx.saved_tensors
        T, N, d, h = ctx.dims
        dev = dz.device
        dz = dz.contiguous()
        z0 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        g_fw, g_fb = z0(h, h), z0(h)
        ops.gemm("tn", dz, hs[T * N:], g_fw, accumulate=True)
        ops.colsum(dz, g_fb)
        dh = _new((N, h), dev)
        ops.gemm("nn", dz, fw, dh)
        dG = _new((T * N, 4 * h), dev)
        ops.lstm_bwd([dict(T=T, B=N, h=h, gates=gates, cs=cs, W=w_hh, dh_all=None, dh_last=dh, dc_ext=None, dG=dG,
                           dc_scratch=_new((N, h), dev))])
        g_ih, g_hh, g_b = z0(4 * h, d), z0(4 * h, h), z0(4 * h)
        ops.gemm("tn", dG, X2, g_ih, accumulate=True)
        ops.gemm("tn", dG, hs[:T * N], g_hh, accumulate=True)
        ops.colsum(dG, g_b)
        dx = None
        if ctx.x_grad:
            dx2 = _new((T * N, d), dev)
            ops.gemm("nn", dG, w_ih, dx2)
            dx = dx2.view(T, N, d)
        return dx, g_ih, g_hh, g_b, g_b.clone(), g_fw, g_fb


def encoder_forward(mod, x):
    """encoderLSTM.forward, mfm_model.py:47-62."""
    return _EncoderFn.apply(x, mod.lstm.weight_ih, mod.lstm.weight_hh, mod.lstm.bias_ih, mod.lstm.bias_hh,
                            mod.fc1.weight, mod.fc1.bias)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, T, w_ih, w_hh, b_ih, b_hh, fw, fb):
        ops = _ops()
        N, h = emb.shape
        d = fw.shape[0]
        dev = emb.device
        emb = emb.contiguous().float()
        G0 = _new((N, 4 * h), dev)
        ops.gemm("nt", emb, w_ih, G0, bias=b_ih, bias2=b_hh)
        Wm, bs = _new((4 * h, h), dev), _new((4 * h,), dev)
        ops.add(w_ih, w_hh, Wm)                  # input == h_{t-1} for t >= 1 (mfm_model.py:85)
        ops.add(b_ih, b_hh, bs)
        hs, cs, gates = _new(((T + 1) * N, h), dev), _new(((T + 1) * N, h), dev), _new((T * N, 4 * h), dev)
        ops.lstm_fwd([dict(T=T, B=N, h=h, gx=G0, gx_steps=1, bias_rest=bs, W=Wm, hs=hs, cs=cs, gates=gates)])
        xh = _new((T * N, d), dev)
        ops.gemm("nt", hs[N:], fw, xh, bias=fb)
        ctx.save_for_backward(emb, w_ih, fw, Wm, hs, cs, gates)
        ctx.dims = (T, N, d, h)
        return xh.view(T, N, d)

    @staticmethod
    def backward(ctx, dxh):
        ops = _ops()
        emb, w_ih, fw, Wm, hs, cs, gates = ctx.saved_tensors
        T, N, d, h = ctx.dims
        dev = dxh.device
        dX = dxh.contiguous().view(T * N, d)
        z0 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        g_fw, g_fb = z0(d, h), z0(d)
        ops.gemm("tn", dX, hs[N:], g_fw, accumulate=True)
        ops.colsum(dX, g_fb)
        dH = _new((T * N, h), dev)
        ops.gemm("nn", dX, fw, dH)
        dG = _new((T * N, 4 * h), dev)
        ops.lstm_bwd([dict(T=T, B=N, h=h, gates=gates, cs=cs, W=Wm, dh_all=dH, dh_last=None, dc_ext=None, dG=dG,
                           dc_scratch=_new((N, h), dev))])
        g_ih, g_hh, g_b = z0(4 * h, h), z0(4 * h, h), z0(4 * h)
        ops.gemm("tn", dG, hs[:T * N], g_hh, accumulate=True)
        ops.gemm("tn", dG, hs[:T * N], g_ih, accumulate=True)
        ops.gemm("tn", dG[:N], emb, g_ih, accumulate=True)
        ops.colsum(dG, g_b)
        demb = _new((N, h), dev)
        ops.gemm("nn", dG[:N], w_ih, demb)
        return demb, None, g_ih, g_hh, g_b, g_b.clone(), g_fw, g_fb


def decoder_forward(mod, hT, t):
    """decoderLSTM.forward, mfm_model.py:72-91."""
    return _DecoderFn.apply(hT, int(t), mod.lstm.weight_ih, mod.lstm.weight_hh, mod.lstm.bias_ih, mod.lstm.bias_hh,
                            mod.fc1.weight, mod.fc1.bias)


_MFN_UNUSED = ("out_fc1.weight", "out_fc1.bias", "out_fc2.weight", "out_fc2.bias")


class _MFNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        T, N, _ = x.shape
        engs = mod.__dict__.setdefault("_engines", {})
        key = (T, N, str(x.device))
        if key not in engs:
            if len(engs) >= 4:
                engs.pop(next(iter(engs)))
            engs[key] = E.Engine(mod._cfg, T, N, x.device, _ops(), mfn_only=True, mfn_prefix="")
        eng = engs[key]
        P = OrderedDict(zip(mod._names, params))
        rng = mod.__dict__.get("_rng")
        if rng is None or rng.device != x.device:
            rng = torch.tensor([123, 0], dtype=torch.int64, device=x.device)
            mod.__dict__["_rng"] = rng
        if mod.training:
            _ops().rng_tick(rng)
        out = eng.forward(P, x, [], train=mod.training, rng=rng)
        ctx.eng, ctx.P = eng, P
        return out["mfn_last"].clone()

    @staticmethod
    def backward(ctx, dlast):
        eng, P = ctx.eng, ctx.P
        G = OrderedDict((k, torch.zeros_like(p)) for k, p in P.items())
        eng.backward(P, G, None, None, 0.0, d_mfn_last=dlast.contiguous())
        return (None, None) + tuple(G.values())


def mfn_forward(mod, x):
    """MFN.forward, mfm_model.py:140-199."""
    if "_names" not in mod.__dict__:
        mod.__dict__["_names"] = [k for k, _ in mod.named_parameters() if k not in _MFN_UNUSED]
    pd = dict(mod.named_parameters())
    return _MFNFn.apply(mod, x.contiguous().float(), *[pd[k] for k in mod._names])
