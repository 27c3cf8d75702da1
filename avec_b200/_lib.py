"""ctypes binding of libavec_b200.so (the C ABI declared in include/avec_b200.h).

The product path has no CPU or PyTorch-eager fallback: if the shared library is missing, or a call returns a negative
status, a RuntimeError is raised (reference error convention: Python exceptions, nnet/model.py:819-828).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVEC_LIB selects another build of the same sources (e.g. the -DAVEC_TIMELINE diagnostics build used by tools/ts_probe.py)
LIB_PATH = os.environ.get("AVEC_LIB") or os.path.join(_HERE, "libavec_b200.so")

F32, BF16 = 0, 1
GEMM_PLAIN, GEMM_CONV_FWD, GEMM_CONV_DGRAD, GEMM_CONV_WGRAD = 0, 1, 2, 3
EPI_LINEAR, EPI_SWISH, EPI_RESIDUAL, EPI_DSWISH, EPI_ACCUM, EPI_RELU = 0, 1, 2, 3, 4, 5
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SWISH = 0, 1, 2
COPY_CHUNK = 4096      # AVEC_COPY_CHUNK
STATS_REPLICAS = 32   # copies of the GEMM-epilogue BatchNorm statistics accumulator (AVEC_STATS_REPLICAS in common.cuh)


class ConvGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "N", "Ti", "Hi", "Wi", "C", "To", "Ho", "Wo", "Co", "KT", "KH", "KW", "st", "sh", "sw", "pt", "ph", "pw")]


class GemmArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("impl", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("sam", C.c_longlong), ("sak", C.c_longlong),
        ("B", C.c_void_p), ("sbn", C.c_longlong), ("sbk", C.c_longlong),
        ("ab_dtype", C.c_int), ("g", ConvGeom),
        ("epi", C.c_int), ("alpha", C.c_float), ("bias", C.c_void_p),
        ("out", C.c_void_p), ("out_dtype", C.c_int), ("ldo", C.c_longlong),
        ("out2", C.c_void_p), ("out2_dtype", C.c_int), ("ldo2", C.c_longlong),
        ("aux", C.c_void_p), ("aux_dtype", C.c_int), ("ldaux", C.c_longlong),
        ("colstats", C.c_void_p), ("split_k", C.c_int),
        ("drop_p", C.c_float), ("drop_site", C.c_int), ("drop_rng", C.c_void_p),
    ]


class CopyJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("start", C.c_longlong), ("ss", C.c_longlong * 4), ("ds", C.c_longlong * 4),
                ("n", C.c_int * 4), ("src_dtype", C.c_int), ("dst_dtype", C.c_int)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (every symbol include/avec_b200.h declares; tests/test_abi.py checks the two lists agree)
PROTOTYPES = {
    "avec_strerror": ([_I], C.c_char_p),
    "avec_last_cuda_error": ([], _I),
    "avec_version": ([], _I),
    "avec_launch_count": ([], _L),
    "avec_reset_launch_count": ([], None),
    "avec_gemm": ([C.POINTER(GemmArgs), _P], _I),
    "avec_set_tma": ([_I], None),
    "avec_set_pdl": ([_I], None),
    "avec_pdl_exclude_stream": ([_P, _I], None),
    "avec_set_debug_timestamps": ([_P], None),
    "avec_colsum": ([_P, _I, _L, _I, _L, _F, _P, _I, _P], _I),
    "avec_layernorm_fwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _L, _P], _I),
    "avec_layernorm_bwd": ([_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "avec_upsample_add": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_pool_sum": ([_P, _P, _I, _I, _I, _I, _I, _I, _L, _P], _I),
    "avec_set_attention_long": ([_I], None),
    "avec_relpos_attn_fwd": ([_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P], _I),
    "avec_relpos_attn_bwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P], _I),
    "avec_relpos_attn_tc_fwd": ([_P, _L, _P, _L, _P, _I, _P, _L, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_relpos_attn_tc_bwd": ([_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _I, _P, _L, _P, _P, _L, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_attn_group_pack": ([_P, _I, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_attn_group_unpack": ([_P, _I, _L, _P, _I, _L, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_attn_group_unpack_dqkv": ([_P, _L, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_softmax_fwd": ([_P, _I, _P, _I, _L, _I, _P], _I),
    "avec_softmax_bwd": ([_P, _P, _I, _P, _P, _I, _L, _I, _P], _I),
    "avec_glu_dwconv_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_glu_dwconv_bwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_bn_finalize": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _I, _P], _I),
    "avec_bn_eval_affine": ([_P, _P, _P, _P, _P, _P, _I, _F, _P], _I),
    "avec_bn_stats": ([_P, _I, _L, _I, _P, _P], _I),
    "avec_bn_apply": ([_P, _P, _P, _P, _P, _L, _I, _I, _I, _L, _P], _I),
    "avec_bn_bwd_reduce": ([_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P], _I),
    "avec_bn_bwd_apply": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P], _I),
    "avec_stft_mel_log": ([_P, _P, _P, _I, _I, _I, _I, _P], _I),
    "avec_im2col_c1": ([_P, _P, C.POINTER(ConvGeom), _I, _I, _P], _I),
    "avec_bn_bwd_pool": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_stem2d_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "avec_stem2d_wgrad": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "avec_stem3d_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "avec_stem3d_wgrad": ([_P, _P, _P, _I, _I, _I, _I, _P], _I),
    "avec_bn_relu_maxpool_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_bn_relu_maxpool_bwd": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_avgpool_fwd": ([_P, _P, _I, _I, _I, _I, _P], _I),
    "avec_avgpool_bwd": ([_P, _P, _I, _I, _I, _I, _P], _I),
    "avec_zero_upsample": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_ctc_loss": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "avec_counter_advance": ([_P, _P], _I),
    "avec_dropout": ([_P, _P, _P, _L, _I, _I, _F, _F, _P, _I, _I, _I, _I, _L, _P], _I),
    "avec_spec_augment": ([_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _P], _I),
    "avec_video_augment": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _F, _F, _P, _I, _P, _P], _I),
    "avec_ctc_greedy_decode": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "avec_sumsq": ([_P, _L, _P, _P], _I),
    "avec_adam_step": ([_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _F, _F, _F, _P, _P, _P, _P], _I),
    "avec_convert_multi": ([_P, _P, _I, _P], _I),
    "avec_unpad_heads": ([_P, _P, _L, _I, _I, _L, _P], _I),
    "avec_convert": ([_P, _I, _L, _P, _I, _L, _L, _I, _P], _I),
}

_lib = None


def load():
    """Load the shared library (no CUDA context is needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"avec_b200: {LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or avec_b200/csrc/build.sh). There is no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        raise RuntimeError(f"avec_b200: {what} failed: {lib.avec_strerror(rc).decode()} "
                           f"(status {rc}, cuda error {lib.avec_last_cuda_error()})")
