"""avec_b200 - B200-native (sm_100a) hot path of AVEC's Efficient-Conformer encoders.

  avec_b200.nnet          drop-in mirror of the reference's nnet modules on the hot path (same names / state_dict keys)
  avec_b200.functional    fused autograd Functions (one per reference module) over the C ABI
  avec_b200.ops           tensor-level wrappers of include/avec_b200.h
  avec_b200.patch_reference()   swap these encoders (+ CTC loss, fused Adam) under the unmodified reference launcher
"""
from . import _lib  # noqa: F401
from . import data  # noqa: F401
from .functional import set_compute_dtype, compute_dtype, new_step, invalidate_weights, manual_seed  # noqa: F401
from .ops import set_gemm_impl, launch_count, reset_launch_count  # noqa: F401
from .dropin import patch_reference, unpatch_reference  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
