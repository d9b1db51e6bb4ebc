"""factorized_b200 -- the MFM (Multimodal Factorization Model) training step on B200 (sm_100a).

Drop-in for the hot path of pliang279/factorized: ``encoderLSTM / decoderLSTM / MFN / MFM / MFM_KL / MFM_KL_EF`` and
``train_mfm``; arithmetic in hand-written CUDA behind ``include/mfm_b200.h``.  CUDA only.
"""
from .mfm_model import encoderLSTM, decoderLSTM, MFN, MFM, MFM_KL, MFM_KL_EF, EFLSTM  # noqa: F401


def train_mfm(*a, **k):
    from .train import train_mfm as f
    return f(*a, **k)
