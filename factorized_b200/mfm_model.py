"""Drop-in replacements for the reference's ``mfm_model`` classes on the MFM path.

Same class names, constructor signatures, submodule / parameter names and
construction order as /root/reference/mfm_model.py (encoderLSTM :40-62,
decoderLSTM :64-91, MFN :93-199, MFM :469-555), so ``torch.manual_seed(s);
MFM(*configs)`` draws the same initial weights, ``state_dict()`` has the same 90
keys and ``torch.save(model)`` round-trips.  The parameters live in ordinary
``nn.LSTMCell`` / ``nn.Linear`` containers, but those containers are never
*called*: ``forward`` runs the hand-written sm_100a kernels through the C ABI
(``cuda_ops``) under one ``torch.autograd.Function`` whose backward is the
hand-derived adjoint (``engine.Engine.backward``).  CUDA only -- like the
reference, which hard-codes ``.cuda()`` -- and no fallback.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch
import torch.nn as nn

from . import engine as E

_OPS = None


def _ops():
    """The CUDA primitive binding, created on first use (keeps modules picklable)."""
    global _OPS
    if _OPS is None:
        from .cuda_ops import CudaOps
        _OPS = CudaOps()
    return _OPS


def _require_cuda(x: torch.Tensor, who: str):
    if not x.is_cuda:
        raise RuntimeError("%s: input is on %s. factorized_b200 runs on CUDA (sm_100a) only; there is no CPU path "
                           "(the reference itself calls .cuda() inside forward)." % (who, x.device))


UNUSED = ("mfn_encoder.out_fc1.weight", "mfn_encoder.out_fc1.bias",
          "mfn_encoder.out_fc2.weight", "mfn_encoder.out_fc2.bias")


class _MFMFunction(torch.autograd.Function):
    """forward = Engine.forward, backward = Engine.backward; inputs (x, 4 noise tensors, *params)."""

    @staticmethod
    def forward(ctx, module, x, n0, n1, n2, n3, *params):
        T, B, _ = x.shape
        eng = module._engine(T, B, x.device)
        names = module._param_names
        P = OrderedDict(zip(names, params))
        rng = module._rng_state(x.device)
        if module.training:
            _ops().rng_tick(rng)
        # forward-only inference (evaluate / predict, mfm_mosi.py:445-465, call forward on the WHOLE validation / test set under
        # no_grad and discard the MMD): with ``module.eval_skip_mmd`` set the O(n^2) statistic is not computed and reads 0
        eng.want_mmd = not module.__dict__.get("_skip_mmd_now", False)     # (decided in MFM.forward: grad mode is off in here)
        if not eng.want_mmd:
            _ops().zero(eng.loss_buf[4:8])
        try:
            out = eng.forward(P, x, [n0, n1, n2, n3], train=module.training, rng=rng)
        finally:
            eng.want_mmd = True
        ctx.eng, ctx.P, ctx.gen = eng, P, eng_generation(eng, bump=True)
        dm = eng.dm
        mmd = eng.loss_buf[4:8].sum()
        res = (out["x_l_hat"].reshape(T, B, dm.d[0]).clone(), out["x_a_hat"].reshape(T, B, dm.d[1]).clone(),
               out["x_v_hat"].reshape(T, B, dm.d[2]).clone(), out["y_hat"].clone(), mmd)
        # (the ablation models of factorized_b200.ablations lack some of the four latents)
        module._latents = {k: out[k].clone() for k in ("zl", "za", "zv", "zy") if out.get(k) is not None}
        return res

    @staticmethod
    def backward(ctx, dxl, dxa, dxv, dy, dmmd):
        eng, P = ctx.eng, ctx.P
        if eng_generation(eng) != ctx.gen:
            raise RuntimeError("MFM backward called after another forward of the same (T,B) shape overwrote the "
                               "kernel workspace; run backward before the next forward")
        dm = eng.dm
        TB = dm.T * dm.B
        ops = _ops()

        def dense(g, shape):
            if g is None:
                return torch.zeros(shape, dtype=torch.float32, device=eng.device)
            return g.contiguous().view(shape)
        dX = [dense(dxl, (TB, dm.d[0])), dense(dxa, (TB, dm.d[1])), dense(dxv, (TB, dm.d[2]))]
        dY = dense(dy, (dm.B, dm.out))
        sizes = [p.numel() for p in P.values()]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=eng.device)
        ops.zero(flat)
        G, o = OrderedDict(), 0
        for (k, p), n in zip(P.items(), sizes):
            G[k] = flat[o:o + n].view(p.shape)
            o += n
        if dmmd is None:
            eng.backward(P, G, dX, dY, 0.0)
        else:
            eng.backward(P, G, dX, dY, 1.0, mmd_scale_dev=dmmd.contiguous().view(1))
        return (None, None, None, None, None, None) + tuple(G.values())


def eng_generation(eng, bump=False):
    g = getattr(eng, "_generation", 0)
    if bump:
        g += 1
        eng._generation = g
    return g


class encoderLSTM(nn.Module):
    """mfm_model.py:40-62.  ``forward(x[T,N,d]) -> fc1(h_T) [N,h]``."""

    def __init__(self, d, h):
        super(encoderLSTM, self).__init__()
        self.lstm = nn.LSTMCell(d, h)
        self.fc1 = nn.Linear(h, h)
        self.h = h

    def forward(self, x):
        from .standalone import encoder_forward
        _require_cuda(x, "encoderLSTM.forward")
        return encoder_forward(self, x)


class EFLSTM(nn.Module):
    """The early-fusion LSTM baseline of the reference's MOSI script (test_mosi.py:130-157): one LSTMCell over the
    concatenated input, ``fc2(dropout(relu(fc1(h_T))))``.  Same parameter names (``lstm``, ``fc1``, ``fc2``); the recurrence
    and ``fc1`` run on the CUDA kernels of ``encoderLSTM``, the [N, h] head is three small torch ops."""

    def __init__(self, d, h, output_dim, dropout):
        super(EFLSTM, self).__init__()
        self.h = h
        self.lstm = nn.LSTMCell(d, h)
        self.fc1 = nn.Linear(h, h)
        self.fc2 = nn.Linear(h, output_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        from .standalone import encoder_forward
        _require_cuda(x, "EFLSTM.forward")
        output = torch.relu(encoder_forward(self, x))       # encoder_forward = fc1(h_T), no activation
        return self.fc2(self.dropout(output))


class decoderLSTM(nn.Module):
    """mfm_model.py:64-91.  ``forward(hT[N,h], t) -> [t,N,d]``."""

    def __init__(self, h, d):
        super(decoderLSTM, self).__init__()
        self.lstm = nn.LSTMCell(h, h)
        self.fc1 = nn.Linear(h, d)
        self.d = d
        self.h = h

    def forward(self, hT, t):
        from .standalone import decoder_forward
        _require_cuda(hT, "decoderLSTM.forward")
        return decoder_forward(self, hT, t)


class MFN(nn.Module):
    """mfm_model.py:93-199 (parameter container + standalone forward).  ``out_fc1/out_fc2`` are
    constructed and never used, exactly like the reference (:136-138)."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        super(MFN, self).__init__()
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        [self.dh_l, self.dh_a, self.dh_v] = config["h_dims"]
        total_h_dim = self.dh_l + self.dh_a + self.dh_v
        self.mem_dim = config["memsize"]
        window_dim = config["windowsize"]
        output_dim = config["output_dim"]
        attInShape = total_h_dim * window_dim
        gammaInShape = attInShape + self.mem_dim
        final_out = total_h_dim + self.mem_dim
        self.lstm_l = nn.LSTMCell(self.d_l, self.dh_l)
        self.lstm_a = nn.LSTMCell(self.d_a, self.dh_a)
        self.lstm_v = nn.LSTMCell(self.d_v, self.dh_v)
        self.att1_fc1 = nn.Linear(attInShape, NN1Config["shapes"])
        self.att1_fc2 = nn.Linear(NN1Config["shapes"], attInShape)
        self.att1_dropout = nn.Dropout(NN1Config["drop"])
        self.att2_fc1 = nn.Linear(attInShape, NN2Config["shapes"])
        self.att2_fc2 = nn.Linear(NN2Config["shapes"], self.mem_dim)
        self.att2_dropout = nn.Dropout(NN2Config["drop"])
        self.gamma1_fc1 = nn.Linear(gammaInShape, gamma1Config["shapes"])
        self.gamma1_fc2 = nn.Linear(gamma1Config["shapes"], self.mem_dim)
        self.gamma1_dropout = nn.Dropout(gamma1Config["drop"])
        self.gamma2_fc1 = nn.Linear(gammaInShape, gamma2Config["shapes"])
        self.gamma2_fc2 = nn.Linear(gamma2Config["shapes"], self.mem_dim)
        self.gamma2_dropout = nn.Dropout(gamma2Config["drop"])
        self.out_fc1 = nn.Linear(final_out, outConfig["shapes"])
        self.out_fc2 = nn.Linear(outConfig["shapes"], output_dim)
        self.out_dropout = nn.Dropout(outConfig["drop"])
        self._cfg = [dict(config), dict(NN1Config), dict(NN2Config), dict(gamma1Config), dict(gamma2Config), dict(outConfig)]

    def __getstate__(self):
        st = self.__dict__.copy()
        for k in ("_engines", "_rng", "_names"):
            st.pop(k, None)
        return st

    def forward(self, x):
        from .standalone import mfn_forward
        _require_cuda(x, "MFN.forward")
        return mfn_forward(self, x)


class MFM(nn.Module):
    """mfm_model.py:469-555.  ``forward(x[T,N,D]) -> ([x_l_hat, x_a_hat, x_v_hat, y_hat], mmd_loss, 0.0)``."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        super(MFM, self).__init__()
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        [self.dh_l, self.dh_a, self.dh_v] = config["h_dims"]
        zy_size, zl_size, za_size, zv_size = config["zy_size"], config["zl_size"], config["za_size"], config["zv_size"]
        fy_size, fl_size, fa_size, fv_size = config["fy_size"], config["fl_size"], config["fa_size"], config["fv_size"]
        total_h_dim = self.dh_l + self.dh_a + self.dh_v
        last_mfn_size = total_h_dim + config["memsize"]
        output_dim = config["output_dim"]
        # construction order fixes the init RNG stream (mfm_model.py:491-520)
        self.encoder_l = encoderLSTM(self.d_l, zl_size)
        self.encoder_a = encoderLSTM(self.d_a, za_size)
        self.encoder_v = encoderLSTM(self.d_v, zv_size)
        self.decoder_l = decoderLSTM(fy_size + fl_size, self.d_l)
        self.decoder_a = decoderLSTM(fy_size + fa_size, self.d_a)
        self.decoder_v = decoderLSTM(fy_size + fv_size, self.d_v)
        self.mfn_encoder = MFN(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)
        self.last_to_zy_fc1 = nn.Linear(last_mfn_size, zy_size)
        self.zy_to_fy_fc1 = nn.Linear(zy_size, fy_size)
        self.zy_to_fy_fc2 = nn.Linear(fy_size, fy_size)
        self.zy_to_fy_dropout = nn.Dropout(config["zy_to_fy_dropout"])
        self.zl_to_fl_fc1 = nn.Linear(zl_size, fl_size)
        self.zl_to_fl_fc2 = nn.Linear(fl_size, fl_size)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za_size, fa_size)
        self.za_to_fa_fc2 = nn.Linear(fa_size, fa_size)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv_size, fv_size)
        self.zv_to_fv_fc2 = nn.Linear(fv_size, fv_size)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fy_size, fy_size)
        self.fy_to_y_fc2 = nn.Linear(fy_size, output_dim)
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._cfg = [dict(config), dict(NN1Config), dict(NN2Config), dict(gamma1Config), dict(gamma2Config), dict(outConfig)]
        self._param_names = [k for k, _ in self.named_parameters() if k not in UNUSED]
        self.mmd_noise = "cpu"      # "cpu": torch.randn on the CPU default generator then H2D, bit-compatible with
        #                             loss_MMD (mfm_model.py:26-29); "cuda": generated on the device
        self.dropout_seed = 123

    # transient kernel state is rebuilt on demand and never pickled (torch.save(model) must work, mfm_mosi.py:477)
    def __getstate__(self):
        st = self.__dict__.copy()
        for k in ("_engines", "_rng", "_latents"):
            st.pop(k, None)
        return st

    def _engine(self, T, B, device):
        engs = self.__dict__.setdefault("_engines", {})
        key = (int(T), int(B), str(device))
        if key not in engs:
            if len(engs) >= 4:                       # bound workspace growth across odd batch sizes
                engs.pop(next(iter(engs)))
            engs[key] = E.Engine(self._cfg, T, B, device, _ops(), head="l1", variant=getattr(self, "_variant", "mfm"))
        return engs[key]

    def _rng_state(self, device):
        r = self.__dict__.get("_rng")
        if r is None or r.device != torch.device(device):
            r = torch.tensor([int(self.dropout_seed), 0], dtype=torch.int64, device=device)
            self.__dict__["_rng"] = r
        return r

    def draw_mmd_noise(self, n, device):
        """The four Gaussian samples of loss_MMD, drawn in the reference's order zl, za, zv, zy (:536)."""
        c = self._cfg[0]
        sizes = (c["zl_size"], c["za_size"], c["zv_size"], c["zy_size"])
        if getattr(self, "_variant", "mfm") in ("kl", "kl_ef"):   # MFM_KL / MFM_KL_EF draw nothing (no sampling, KL regulariser)
            return [torch.zeros(1, 1, device=device) for _ in sizes]
        if self.mmd_noise == "cpu":
            return [torch.randn(n, k).to(device) for k in sizes]
        return [torch.randn(n, k, device=device) for k in sizes]

    def forward(self, x):
        _require_cuda(x, "MFM.forward")
        if x.dim() != 3:
            raise ValueError("MFM.forward expects x[T,N,D]")
        if x.requires_grad:
            raise RuntimeError("MFM.forward: gradient w.r.t. the input is not provided (the reference never asks for it)")
        x = x.contiguous().float()
        noise = self.draw_mmd_noise(x.shape[1], x.device)
        pd = dict(self.named_parameters())
        params = [pd[k] for k in self._param_names]
        self.__dict__["_skip_mmd_now"] = bool(getattr(self, "eval_skip_mmd", False) and not self.training and not torch.is_grad_enabled())
        x_l_hat, x_a_hat, x_v_hat, y_hat, mmd_loss = _MFMFunction.apply(self, x, *noise, *params)
        missing_loss = 0.0
        decoded = [x_l_hat, x_a_hat, x_v_hat, y_hat]
        return decoded, mmd_loss, missing_loss

    @property
    def latents(self) -> Dict[str, torch.Tensor]:
        """zl, za, zv, zy of the last forward (the reference computes but does not return them)."""
        return self.__dict__.get("_latents", {})


class MFM_KL(MFM):
    """mfm_model.py:662-764 -- the variant ``train_mfm`` builds for ``config['type'] == 'kl'`` (mfm_mosi.py:398-399).
    ``forward(x) -> ([x_l_hat, x_a_hat, x_v_hat, y_hat], kld_loss, 0.0)``: the encoder outputs pass one more Linear to
    the means (which ARE the latents: the reference does not sample) and another to the log-variances, and the
    regulariser is ``loss_KLD`` (:36-38) summed over zl, za, zv, zy.  Same kernels, one more reduction."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        [self.dh_l, self.dh_a, self.dh_v] = config["h_dims"]
        zy_size, zl_size, za_size, zv_size = config["zy_size"], config["zl_size"], config["za_size"], config["zv_size"]
        fy_size, fl_size, fa_size, fv_size = config["fy_size"], config["fl_size"], config["fa_size"], config["fv_size"]
        total_h_dim = self.dh_l + self.dh_a + self.dh_v
        last_mfn_size = total_h_dim + config["memsize"]
        output_dim = config["output_dim"]
        # construction order fixes the init RNG stream (mfm_model.py:683-721)
        self.encoder_l = encoderLSTM(self.d_l, zl_size)
        self.encoder_a = encoderLSTM(self.d_a, za_size)
        self.encoder_v = encoderLSTM(self.d_v, zv_size)
        self.decoder_l = decoderLSTM(fy_size + fl_size, self.d_l)
        self.decoder_a = decoderLSTM(fy_size + fa_size, self.d_a)
        self.decoder_v = decoderLSTM(fy_size + fv_size, self.d_v)
        self.mfn_encoder = MFN(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)
        self.last_to_zy_fc1 = nn.Linear(last_mfn_size, zy_size)
        self.last_to_logvarzy_fc1 = nn.Linear(last_mfn_size, zy_size)
        self.last_to_zl_fc1 = nn.Linear(zl_size, zl_size)
        self.last_to_za_fc1 = nn.Linear(za_size, za_size)
        self.last_to_zv_fc1 = nn.Linear(zv_size, zv_size)
        self.last_to_logvarzl_fc1 = nn.Linear(zl_size, zl_size)
        self.last_to_logvarza_fc1 = nn.Linear(za_size, za_size)
        self.last_to_logvarzv_fc1 = nn.Linear(zv_size, zv_size)
        self.zy_to_fy_fc1 = nn.Linear(zy_size, fy_size)
        self.zy_to_fy_fc2 = nn.Linear(fy_size, fy_size)
        self.zy_to_fy_dropout = nn.Dropout(config["zy_to_fy_dropout"])
        self.zl_to_fl_fc1 = nn.Linear(zl_size, fl_size)
        self.zl_to_fl_fc2 = nn.Linear(fl_size, fl_size)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za_size, fa_size)
        self.za_to_fa_fc2 = nn.Linear(fa_size, fa_size)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv_size, fv_size)
        self.zv_to_fv_fc2 = nn.Linear(fv_size, fv_size)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fy_size, fy_size)
        self.fy_to_y_fc2 = nn.Linear(fy_size, output_dim)
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._cfg = [dict(config), dict(NN1Config), dict(NN2Config), dict(gamma1Config), dict(gamma2Config), dict(outConfig)]
        self._param_names = [k for k, _ in self.named_parameters() if k not in UNUSED]
        self._variant = "kl"
        self.mmd_noise = "cpu"
        self.dropout_seed = 123


class MFM_KL_EF(MFM):
    """mfm_model.py:557-660 -- MFM_KL with the MFN encoder replaced by ONE early-fusion ``encoderLSTM`` over the
    concatenated input (``ef_encoder``, hidden size zl+za+zv); z_y and its log-variance are Linears of its output.
    ``forward(x) -> ([x_l_hat, x_a_hat, x_v_hat, y_hat], kld_loss, 0.0)``.  Same kernels as MFM_KL minus the MFN."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        [self.dh_l, self.dh_a, self.dh_v] = config["h_dims"]
        zy_size, zl_size, za_size, zv_size = config["zy_size"], config["zl_size"], config["za_size"], config["zv_size"]
        fy_size, fl_size, fa_size, fv_size = config["fy_size"], config["fl_size"], config["fa_size"], config["fv_size"]
        output_dim = config["output_dim"]
        # construction order fixes the init RNG stream (mfm_model.py:579-619)
        self.encoder_l = encoderLSTM(self.d_l, zl_size)
        self.encoder_a = encoderLSTM(self.d_a, za_size)
        self.encoder_v = encoderLSTM(self.d_v, zv_size)
        self.decoder_l = decoderLSTM(fy_size + fl_size, self.d_l)
        self.decoder_a = decoderLSTM(fy_size + fa_size, self.d_a)
        self.decoder_v = decoderLSTM(fy_size + fv_size, self.d_v)
        last_ef_size = zl_size + za_size + zv_size
        self.ef_encoder = encoderLSTM(self.d_l + self.d_a + self.d_v, last_ef_size)
        self.last_to_zy_fc1 = nn.Linear(last_ef_size, zy_size)
        self.last_to_logvarzy_fc1 = nn.Linear(last_ef_size, zy_size)
        self.last_to_zl_fc1 = nn.Linear(zl_size, zl_size)
        self.last_to_za_fc1 = nn.Linear(za_size, za_size)
        self.last_to_zv_fc1 = nn.Linear(zv_size, zv_size)
        self.last_to_logvarzl_fc1 = nn.Linear(zl_size, zl_size)
        self.last_to_logvarza_fc1 = nn.Linear(za_size, za_size)
        self.last_to_logvarzv_fc1 = nn.Linear(zv_size, zv_size)
        self.zy_to_fy_fc1 = nn.Linear(zy_size, fy_size)
        self.zy_to_fy_fc2 = nn.Linear(fy_size, fy_size)
        self.zy_to_fy_dropout = nn.Dropout(config["zy_to_fy_dropout"])
        self.zl_to_fl_fc1 = nn.Linear(zl_size, fl_size)
        self.zl_to_fl_fc2 = nn.Linear(fl_size, fl_size)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za_size, fa_size)
        self.za_to_fa_fc2 = nn.Linear(fa_size, fa_size)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv_size, fv_size)
        self.zv_to_fv_fc2 = nn.Linear(fv_size, fv_size)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fy_size, fy_size)
        self.fy_to_y_fc2 = nn.Linear(fy_size, output_dim)
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._cfg = [dict(config), dict(NN1Config), dict(NN2Config), dict(gamma1Config), dict(gamma2Config), dict(outConfig)]
        self._param_names = [k for k, _ in self.named_parameters()]
        self._variant = "kl_ef"
        self.mmd_noise = "cpu"
        self.dropout_seed = 123
