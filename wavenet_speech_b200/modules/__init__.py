"""Drop-in modules: same class names, constructor/forward signatures, attributes and state_dict keys as
the reference's `modules/` package (paultsw/wavenet-speech), computed by libwnb200's sm_100a kernels.

To run the reference's scripts unchanged, put this package's parent directory first on sys.path
(`sys.path.insert(0, "<repo>/wavenet_speech_b200")`) so that `from modules.wavenet import WaveNet`
resolves here; see INTEGRATION.md."""
from .block import GatedActivationUnit, MultiplicativeUnit, ResidualBlock, ResidualMUBlock, ResidualReLUBlock
from .classifier import WaveNetClassifier
from .conv_ops import CausalConv1d, NonCausalConv1d, autopad, compute_new_length, reshape_in, reshape_out
from .layernorm import LayerNorm
from .linear_conv_ops import LinearConv1d
from .raw_ctcnet import RawCTCNet
from .sequence_decoders import Decoder, argmax_decode, greedy_ctc_decode, labels2strings
from .wavenet import WaveNet

__all__ = ["CausalConv1d", "NonCausalConv1d", "ResidualBlock", "GatedActivationUnit", "MultiplicativeUnit",
           "ResidualMUBlock", "ResidualReLUBlock", "WaveNet", "RawCTCNet", "WaveNetClassifier", "LayerNorm",
           "LinearConv1d", "autopad", "compute_new_length", "reshape_in", "reshape_out", "argmax_decode", "labels2strings",
           "greedy_ctc_decode", "Decoder"]
