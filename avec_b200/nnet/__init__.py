"""Drop-in mirror of the hot-path part of the reference's `nnet` package (reference nnet/__init__.py:19-49)."""
import torch.nn as nn

from .layers import *          # noqa: F401,F403
from .modules import *         # noqa: F401,F403
from .blocks import *          # noqa: F401,F403
from .networks import *        # noqa: F401,F403
from .models_zoo import *      # noqa: F401,F403
from .preprocessing import AudioPreprocessing  # noqa: F401
from .losses import CTCLoss    # noqa: F401


def zero_dropout(module):
    """p = 0 for every dropout layer (the deterministic parity / throughput configuration)."""
    for m in module.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    return module
