"""avec_b200.patch_reference(): swap the hot path under the UNMODIFIED reference launcher.

`main.py`, `functions.py`, the config files and `nnet.Model` (fit / train_step / evaluate / save / load / swa / DDP wrap) stay
the reference's own code (main.py:49,66-89, functions.py:46-103, nnet/model.py:59-65,346-409).  Calling

    import nnet, avec_b200
    avec_b200.patch_reference()          # before the config builds `model`

rebinds, inside the reference's `nnet` package,

  * networks.{Audio,Visual,AudioVisual}EfficientConformerEncoder -> the sm_100a encoders of avec_b200.nnet.networks.  The zoo models
    (nnet/models_zoo.py:64-182) look them up by name at construction time, so `model.encoder` becomes the fused implementation
    with identical state_dict keys (released checkpoints load unchanged).  Each encoder opens its own forward pass
    (avec_b200.functional.forward_scope: zero arena, dropout-site counter, RNG step), so nothing depends on which `Model` calls it;
  * CTCLoss -> the fused log-softmax + CTC kernel with device-side lengths (losses=True);
  * optimizers.Adam -> the fused flat-buffer Adam, same constructor / state_dict / `scheduler.model_step` contract (optimizer=True);
  * Model.distribute_strategy -> same DDP wrap; its sync_batch_norm argument (default True, model.py:59-61) additionally selects
    the all-reduced BatchNorm statistics of avec_b200.ops.set_sync_batchnorm inside the fused kernels' BatchNorm.

The compute dtype follows the reference's `precision`: Model.train_step wraps the forward in torch autocast when `precision` is a
half type (model.py:356-360) -> bf16 tensor-core path; no autocast -> fp32 parity path (set_compute_dtype("auto")).
"""
import functools

_PATCHED = {}


def patch_reference(ref_nnet=None, losses=True, optimizer=True, compute_dtype="auto"):
    if ref_nnet is None:
        import nnet as ref_nnet          # the reference package must already be importable (it is when main.py runs)
    if getattr(ref_nnet, "__avec_b200_patched__", False):
        return ref_nnet
    from . import functional as AF
    from . import nnet as fast
    from . import ops

    saved = {}
    for name in ("AudioEfficientConformerEncoder", "VisualEfficientConformerEncoder", "AudioVisualEfficientConformerEncoder"):
        saved[("networks", name)] = getattr(ref_nnet.networks, name)
        setattr(ref_nnet.networks, name, getattr(fast.networks, name))
        if hasattr(ref_nnet, name):
            setattr(ref_nnet, name, getattr(fast.networks, name))
    if losses:
        saved[("losses", "CTCLoss")] = ref_nnet.losses.CTCLoss
        ref_nnet.losses.CTCLoss = fast.CTCLoss
        ref_nnet.CTCLoss = fast.CTCLoss
    if optimizer:
        saved[("optimizers", "Adam")] = ref_nnet.optimizers.Adam
        ref_nnet.optimizers.Adam = fast.optimizers.Adam
        if hasattr(ref_nnet, "Adam"):
            ref_nnet.Adam = fast.optimizers.Adam
        ref_nnet.optimizers.optim_dict["Adam"] = fast.optimizers.Adam

    Model = ref_nnet.model.Model
    orig_ds = Model.distribute_strategy
    saved[("model", "distribute_strategy")] = orig_ds

    @functools.wraps(orig_ds)
    def distribute_strategy(self, rank, sync_batch_norm=True):
        ops.set_sync_batchnorm(bool(sync_batch_norm))
        return orig_ds(self, rank, sync_batch_norm=sync_batch_norm)

    Model.distribute_strategy = distribute_strategy
    AF.set_compute_dtype(compute_dtype)
    ref_nnet.__avec_b200_patched__ = True
    _PATCHED[id(ref_nnet)] = saved
    return ref_nnet


def unpatch_reference(ref_nnet=None):
    """undo patch_reference (tests)"""
    if ref_nnet is None:
        import nnet as ref_nnet
    saved = _PATCHED.pop(id(ref_nnet), None)
    if saved is None:
        return
    for (mod, name), obj in saved.items():
        if mod == "model":
            setattr(ref_nnet.model.Model, name, obj)
            continue
        setattr(getattr(ref_nnet, mod), name, obj)
        if hasattr(ref_nnet, name):
            setattr(ref_nnet, name, obj)
        if mod == "optimizers":
            ref_nnet.optimizers.optim_dict["Adam"] = obj
    ref_nnet.__avec_b200_patched__ = False
