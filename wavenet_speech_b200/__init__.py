"""wavenet_speech_b200: B200-native (sm_100a) implementation of the wavenet-speech hot path --
dilated residual stack, RawCTCNet and WaveNetClassifier -- behind the reference's nn.Module API."""
from . import functional, ops  # noqa: F401
from .modules import *  # noqa: F401,F403

__version__ = "0.1.0"
from . import _lib, fastpath, optim  # noqa: F401,E402
from .functional import CTCLoss  # noqa: F401,E402
from .fastpath import invalidate_packs, reduced_precision, tc_precision  # noqa: F401,E402
