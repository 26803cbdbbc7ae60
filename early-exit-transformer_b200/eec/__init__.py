"""eec -- B200-native early-exit conformer (CTC hot path) behind the reference's module interface.

    from eec import Early_conformer, Splitformer, CTCLoss, multi_exit_ctc_loss, greedy_decode, GraphedTrainStep, GraphedEarlyExit
"""
from .lib import EecError, load, LIB_PATH, EXPORTS  # noqa: F401
from .early_exit import Early_conformer, Splitformer, greedy_decode  # noqa: F401
from .ctc import CTCLoss, multi_exit_ctc_loss  # noqa: F401
from .aed import full_conformer, multi_exit_cross_entropy  # noqa: F401
from .graph import GraphedEarlyExit, GraphedForward, GraphedTrainStep  # noqa: F401
from .optim import FusedNoamAdamW  # noqa: F401
from .features import Fbank  # noqa: F401
from .decoder import CUCTCDecoder, CUCTCHypothesis, ctc_cuda_predict, cuda_ctc_decoder  # noqa: F401
from . import batching, distributed  # noqa: F401,E402
