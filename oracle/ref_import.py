"""TEST INFRASTRUCTURE ONLY - imports the unmodified reference (burchim/AVEC) as the live CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package avec_b200 never does.  The reference lives at /root/reference in the authoring container (read-only) or
at baseline/_ref on a GPU box if the driver installed it; it needs four non-arithmetic stub modules (matplotlib, jiwer,
skimage, gdown - see oracle/ref_stubs) ahead of it on sys.path.
"""
import contextlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ["/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def reference_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "nnet")):
            return c
    return None


def available():
    return reference_root() is not None


def import_reference():
    """returns the reference's `nnet` package (raises ImportError if the reference tree is absent)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference tree not found (looked in %s)" % _CANDIDATES)
    if "nnet" in sys.modules and getattr(sys.modules["nnet"], "__file__", "").startswith(root):
        return sys.modules["nnet"]
    stubs = os.path.join(_HERE, "ref_stubs")
    for p in (root, stubs):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, root)
    sys.path.insert(0, stubs)
    with contextlib.redirect_stdout(io.StringIO()):
        import nnet  # noqa: E402  (prints two optional-dependency notices)
    return nnet


def zero_dropout(module):
    """parity runs compare deterministic graphs: p = 0 for every dropout layer."""
    import torch.nn as nn
    for m in module.modules():
        if isinstance(m, nn.Dropout) or m.__class__.__name__ == "Dropout":
            m.p = 0.0
    return module
