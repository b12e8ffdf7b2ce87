"""B200-native inference hot path of ishine/Prior-DiffuSE."""
from .pipeline import Enhancer, inference_schedule, write_wav  # noqa: F401
from .modules import GCRN, DiffUNet, DiffUNet1, DiffWave, Nocon, aia_complex_trans_ri  # noqa: F401
from .diffwave import DiffWaveSampler  # noqa: F401
from . import metrics, signal  # noqa: F401
