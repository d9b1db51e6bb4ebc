"""Import shim so the reference's ``from mfm_model import encoderLSTM, decoderLSTM, MFN, MFM`` (mfm_mosi.py:30-31)
resolves to the B200 implementation."""
from factorized_b200.mfm_model import encoderLSTM, decoderLSTM, MFN, MFM, MFM_KL, MFM_KL_EF, EFLSTM  # noqa: F401
from factorized_b200.ablations import M_A, M_B, M_C, M_D  # noqa: F401,E402  (mfm_mosi.py:30)
from factorized_b200.missing import MFM_missing  # noqa: F401,E402
from factorized_b200.baselines import seq2seq, basic_missing  # noqa: F401,E402
from factorized_b200.functional import compute_kernel, loss_MMD, loss_KLD  # noqa: F401,E402
