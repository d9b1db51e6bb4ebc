"""The baselines of the reference's MOSI script on the same kernels (/root/reference/test_mosi.py): the early-fusion LSTM
(``EFLSTM``, :130-157) and the Memory Fusion Network with its output head (``MFN``, :158-265).

``test_mosi.MFN`` is ``mfm_model.MFN`` (same submodules, same construction order, so the same initial weights for the same seed)
whose forward goes on through the head ``out_fc2(out_dropout(relu(out_fc1(last_hs))))`` that ``mfm_model.MFN`` only constructs.
The recurrences, the attention and the memory run on the CUDA schedule of ``MFN.forward``; the [N, H+mem] head is three small
torch ops (as in ``EFLSTM``).  CUDA only.
"""
from __future__ import annotations

import torch

from . import mfm_model as M
from .mfm_model import EFLSTM  # noqa: F401


class MFN(M.MFN):
    """test_mosi.py:158-265.  ``forward(x[T,N,D]) -> [N,1]``."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        config = dict(config)
        config["output_dim"] = 1                        # hard-coded there (:166)
        super(MFN, self).__init__(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)

    def forward(self, x):
        last_hs = M.MFN.forward(self, x)                # cat(h_T^l, h_T^a, h_T^v, mem_T)  (:258-263)
        return self.out_fc2(self.out_dropout(torch.relu(self.out_fc1(last_hs))))     # :264


# ---------------------------------------------------------------------------------------------------------------------
# the two missing-modality baselines of mfm_model.py (train_seq2seq / train_basic_missing, mfm_mosi.py:769, :1108)
# ---------------------------------------------------------------------------------------------------------------------
import torch.nn as nn  # noqa: E402

from .functional import loss_MMD  # noqa: E402
from .mfm_model import decoderLSTM, encoderLSTM  # noqa: E402


def _factor_mlp(z, fc1, fc2, drop):
    return torch.relu(fc2(drop(torch.relu(fc1(z)))))


class seq2seq(nn.Module):
    """mfm_model.py:887-958: each modality reconstructed from the other two -- a cross-modal ``encoderLSTM`` over
    cat(x_a, x_b), a factor MLP, a ``decoderLSTM``.  ``forward(x) -> ([x_l_hat_nol], [x_a_hat_noa], [x_v_hat_nov], mmd_loss)``.
    Recurrences, input projections, reconstructions and the MMD run on the CUDA kernels (the standalone encoder / decoder
    forwards and ``functional.loss_MMD``, each with autograd); the [N, f] factor MLPs are torch ops."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        super(seq2seq, self).__init__()
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        zl, za, zv = config["zl_size"], config["za_size"], config["zv_size"]
        fl, fa, fv = config["fl_size"], config["fa_size"], config["fv_size"]
        # construction order: mfm_model.py:909-927
        self.encoder_la_to_v = encoderLSTM(self.d_l + self.d_a, zv)
        self.encoder_lv_to_a = encoderLSTM(self.d_l + self.d_v, za)
        self.encoder_av_to_l = encoderLSTM(self.d_a + self.d_v, zl)
        self.decoder_l = decoderLSTM(fl, self.d_l)
        self.decoder_a = decoderLSTM(fa, self.d_a)
        self.decoder_v = decoderLSTM(fv, self.d_v)
        self.zl_to_fl_fc1 = nn.Linear(zl, fl)
        self.zl_to_fl_fc2 = nn.Linear(fl, fl)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za, fa)
        self.za_to_fa_fc2 = nn.Linear(fa, fa)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv, fv)
        self.zv_to_fv_fc2 = nn.Linear(fv, fv)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])

    def forward(self, x):
        d_l, d_a = self.d_l, self.d_a
        x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
        t = x.shape[0]
        zv_nov = self.encoder_la_to_v.forward(torch.cat([x_l, x_a], dim=2))           # :938-940
        za_noa = self.encoder_lv_to_a.forward(torch.cat([x_l, x_v], dim=2))
        zl_nol = self.encoder_av_to_l.forward(torch.cat([x_a, x_v], dim=2))
        mmd_loss = loss_MMD(zv_nov) + loss_MMD(za_noa) + loss_MMD(zl_nol)              # :942
        fl = _factor_mlp(zl_nol, self.zl_to_fl_fc1, self.zl_to_fl_fc2, self.zl_to_fl_dropout)
        fa = _factor_mlp(za_noa, self.za_to_fa_fc1, self.za_to_fa_fc2, self.za_to_fa_dropout)
        fv = _factor_mlp(zv_nov, self.zv_to_fv_fc1, self.zv_to_fv_fc2, self.zv_to_fv_dropout)
        return [self.decoder_l.forward(fl, t)], [self.decoder_a.forward(fa, t)], [self.decoder_v.forward(fv, t)], mmd_loss


class basic_missing(nn.Module):
    """mfm_model.py:960-1017: the label predicted from two modalities -- a cross-modal ``encoderLSTM`` and a two-layer head per
    missing modality.  ``forward(x) -> (y_hat_nol, y_hat_noa, y_hat_nov, mmd_loss)``.  Same split as ``seq2seq``."""

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        super(basic_missing, self).__init__()
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        zy, fy, od, p = config["zy_size"], config["fy_size"], config["output_dim"], config["zy_to_fy_dropout"]
        # construction order: mfm_model.py:982-996
        self.encoder_la_to_y = encoderLSTM(self.d_l + self.d_a, zy)
        self.encoder_lv_to_y = encoderLSTM(self.d_l + self.d_v, zy)
        self.encoder_av_to_y = encoderLSTM(self.d_a + self.d_v, zy)
        self.zy_nol_to_y_fc1 = nn.Linear(zy, fy)
        self.zy_nol_to_y_fc2 = nn.Linear(fy, od)
        self.zy_nol_to_y_dropout = nn.Dropout(p)
        self.zy_noa_to_y_fc1 = nn.Linear(zy, fy)
        self.zy_noa_to_y_fc2 = nn.Linear(fy, od)
        self.zy_noa_to_y_dropout = nn.Dropout(p)
        self.zy_nov_to_y_fc1 = nn.Linear(zy, fy)
        self.zy_nov_to_y_fc2 = nn.Linear(fy, od)
        self.zy_nov_to_y_dropout = nn.Dropout(p)

    def forward(self, x):
        d_l, d_a = self.d_l, self.d_a
        x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
        zy_nov = self.encoder_la_to_y.forward(torch.cat([x_l, x_a], dim=2))           # :1007-1009
        zy_noa = self.encoder_lv_to_y.forward(torch.cat([x_l, x_v], dim=2))
        zy_nol = self.encoder_av_to_y.forward(torch.cat([x_a, x_v], dim=2))
        mmd_loss = loss_MMD(zy_nov) + loss_MMD(zy_noa) + loss_MMD(zy_nol)              # :1011
        y_hat_nol = self.zy_nol_to_y_fc2(self.zy_nol_to_y_dropout(torch.relu(self.zy_nol_to_y_fc1(zy_nol))))
        y_hat_noa = self.zy_noa_to_y_fc2(self.zy_noa_to_y_dropout(torch.relu(self.zy_noa_to_y_fc1(zy_noa))))
        y_hat_nov = self.zy_nov_to_y_fc2(self.zy_nov_to_y_dropout(torch.relu(self.zy_nov_to_y_fc1(zy_nov))))
        return y_hat_nol, y_hat_noa, y_hat_nov, mmd_loss
