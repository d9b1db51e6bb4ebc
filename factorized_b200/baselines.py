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
