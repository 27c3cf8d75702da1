"""Drop-in mirror of the hot-path part of the reference's `nnet` package (reference nnet/__init__.py:19-49)."""
import torch.nn as nn

from .layers import *          # noqa: F401,F403
from .modules import *         # noqa: F401,F403
from .blocks import *          # noqa: F401,F403
from .networks import *        # noqa: F401,F403
from .models_zoo import *      # noqa: F401,F403
from .preprocessing import AudioPreprocessing, SpecAugment  # noqa: F401
from .transforms import VideoAugment, align_video_to_audio  # noqa: F401
from .losses import CTCLoss    # noqa: F401
from .decoders import CTCGreedySearchDecoder  # noqa: F401
from . import optimizers, schedulers  # noqa: F401


def zero_dropout(module):
    """p = 0 for every dropout layer and SpecAugment bypassed: the deterministic parity configuration (SURVEY section 0
    item 10).  Without this call a model in train() runs the reference's training graph (dropout 0.1, SpecAugment)."""
    for m in module.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, SpecAugment):
            m.enabled = False
    return module
