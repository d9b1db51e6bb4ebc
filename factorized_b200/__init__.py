"""factorized_b200 -- the MFM (Multimodal Factorization Model) training step on B200 (sm_100a).

Drop-in for the hot path of pliang279/factorized: ``encoderLSTM / decoderLSTM / MFN / MFM`` and ``train_mfm``, and the variants
either side of it -- ``MFM_KL / MFM_KL_EF / M_A..M_D / MFM_missing`` with ``train_mfm_ablation / train_mfm_missing /
train_mfm_test_zeros`` (``baselines``: the MFN and early-fusion LSTM of test_mosi.py, ``seq2seq``, ``basic_missing``;
``functional``: ``compute_kernel / loss_MMD / loss_KLD``); arithmetic in hand-written CUDA behind ``include/mfm_b200.h``.
CUDA only.
"""
from .mfm_model import encoderLSTM, decoderLSTM, MFN, MFM, MFM_KL, MFM_KL_EF, EFLSTM  # noqa: F401
from .ablations import M_A, M_B, M_C, M_D  # noqa: F401
from .missing import MFM_missing  # noqa: F401


def train_mfm(*a, **k):
    from .train import train_mfm as f
    return f(*a, **k)


def train_mfm_ablation(*a, **k):
    from .train import train_mfm_ablation as f
    return f(*a, **k)


def train_mfm_test_zeros(*a, **k):
    from .train import train_mfm_test_zeros as f
    return f(*a, **k)


def train_mfm_missing(*a, **k):
    from .train import train_mfm_missing as f
    return f(*a, **k)
