"""ctypes binding of libeec.so (the C ABI declared in include/eec.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EEC_LIB") or os.path.join(_HERE, "libeec.so")   # EEC_LIB: A/B-test another build of the same ABI

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_GLU, ACT_DSILU, ACT_RELU, ACT_DRELU = 0, 1, 2, 3, 4, 5

vp, i32, i64, f32, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32),
        ("A", vp), ("lda", i32), ("a_kmajor", i32),
        ("B", vp), ("ldb", i32), ("b_kmajor", i32),
        ("in_dtype", i32),
        ("bias", vp),
        ("act", i32),
        ("preact", vp), ("ldp", i32),
        ("preact_dtype", i32),
        ("alpha", f32),
        ("residual", vp), ("ldr", i32), ("res_row_mod", i32),
        ("C", vp), ("ldc", i32), ("out_dtype", i32),
        ("accumulate", i32),
        ("ln_gamma", vp), ("ln_beta", vp), ("ln_out", vp), ("ln_dtype", i32), ("ld_ln", i32),
        ("ln_mean", vp), ("ln_rstd", vp),
        ("ln2_gamma", vp), ("ln2_beta", vp), ("ln2_mean", vp), ("ln2_rstd", vp),
        ("x_pre", vp),
        ("a_colsum", vp),
        ("a_colsum_scale", f32),
        ("drop_state", vp), ("drop_p", f32), ("drop_site", C.c_uint32),
        ("drop_bits", vp),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("dh", i32), ("Tq", i32), ("Tk", i32),
        ("q", vp), ("ldq", i32),
        ("k", vp), ("ldk", i32),
        ("v", vp), ("ldv", i32),
        ("dtype", i32),
        ("key_len", vp),
        ("key_valid_bits", vp),
        ("causal", i32),
        ("drop_state", vp), ("drop_p", f32), ("drop_site", u32), ("drop_bits", vp),
    ]


# name -> argtypes (stream last); every function returns int unless noted
_SIGS = {
    "eec_gemm": [C.POINTER(GemmDesc), vp],
    "eec_layernorm_fwd": [vp, vp, vp, vp, i32, vp, vp, i32, i32, vp],
    "eec_layernorm_bwd": [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, f32, vp, f32, u32, i32, i32, vp],
    "eec_layernorm_bwd_dy": [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, f32, vp, f32, u32, i32, i32, vp],
    "eec_attn_fwd": [vp, i32, vp, vp, vp, i32, i32, i32, i32, vp, f32, u32, vp, vp],
    "eec_attn_bwd": [vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, f32, u32, vp, vp],
    "eec_attn_general_fwd": [C.POINTER(AttnDesc), vp, i32, vp, vp],
    "eec_attn_general_bwd": [C.POINTER(AttnDesc), vp, vp, i32, vp, vp, i32, vp, i32, vp, i32, vp, vp, vp],
    "eec_key_bits_from_tokens": [vp, i32, i32, i64, vp, vp],
    "eec_embed_pe": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "eec_embed_bwd": [vp, vp, vp, i32, i32, i32, i32, vp],
    "eec_cross_entropy": [vp, vp, i32, i32, vp, vp, vp],
    "eec_dropout_bits": [vp, f32, u32, i64, i32, i64, i32, vp, vp],
    "eec_dropout": [vp, i32, vp, i32, i64, vp, f32, u32, vp],
    "eec_dropout_advance": [vp, vp],
    "eec_dwconv_bn_silu_eval": [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "eec_dwconv_stats": [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "eec_bn_silu_train": [vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, i32, i32, i32, i64, vp],
    "eec_bn_silu_bwd_stats": [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp],
    "eec_bn_silu_bwd_apply": [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, vp],
    "eec_dwconv_bwd": [vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp],
    "eec_glu_bwd": [vp, vp, vp, i32, i32, i32, vp],
    "eec_logsoftmax_fwd": [vp, vp, vp, vp, i32, i32, vp],
    "eec_logsoftmax_bwd": [vp, vp, vp, i32, i32, vp],
    "eec_head_logsoftmax": [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "eec_ctc_fwd_bwd": [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp],
    "eec_greedy_collapse": [vp, vp, vp, i32, i32, i32, vp],
    "eec_ctc_beam_search": [vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp],
    "eec_im2col_k3s2": [vp, i32, i64, i64, i64, vp, i32, i32, i32, i32, i32, vp],
    "eec_col2im_k3s2": [vp, i32, vp, i32, i32, i32, i32, vp],
    "eec_encoder_lengths": [vp, vp, i32, i32, i32, i32, vp],
    "eec_cast": [vp, i32, vp, i32, i64, vp],
    "eec_noam_adamw_step": [vp, vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, f32, f32, vp],
    "eec_colsum": [vp, i32, i32, vp, f32, i32, i32, vp],
    "eec_cast_colsum": [vp, i32, vp, i32, vp, f32, i32, i32, vp],
    "eec_axpy": [vp, f32, vp, i64, vp],
    "eec_scale_dev": [vp, vp, vp, i64, vp],
    "eec_scale_rows_dev": [vp, vp, vp, i32, i64, vp],
    "eec_exit_select": [vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "eec_gather_rows": [vp, vp, vp, vp, i32, i64, vp],
    "eec_set_active_items": [vp, i32, i32, vp],
    "eec_fbank_frames": [vp, vp, i64, vp, vp, i32, i32, i32, i32, vp],
    "eec_fbank_power": [vp, i32, vp, i64, i32, i32, vp],
    "eec_fbank_finish": [vp, i32, vp, i32, i32, i32, vp],
    "eec_fbank_split_operand": [vp, i32, vp, i64, i32, vp],
    "eec_gather_i64": [vp, vp, vp, i32, vp],
    "eec_stride2_gather": [vp, vp, i32, i32, i32, vp],
    "eec_repeat2_add": [vp, vp, i32, i32, i32, vp],
    "eec_repeat2_bwd": [vp, vp, i32, i32, i32, vp],
    "eec_stride2_scatter_add": [vp, vp, i32, i32, i32, vp],
}
EXPORTS = sorted(list(_SIGS) + ["eec_last_error", "eec_version", "eec_device_ok", "eec_ctc_workspace_bytes", "eec_dwconv_bwd_workspace_bytes",
                             "eec_launch_count", "eec_ctc_beam_workspace_bytes"])

# entry points of eec/libeec_exp.so only (`make experiments`, include/eec_experiments.h): bound when the loaded library has them
_SIGS_EXPERIMENTAL = {
    "eec_ffn_fwd": [vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, vp],
}

_lib = None


class EecError(RuntimeError):
    pass


def load():
    """Load libeec.so; raises if it was not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EecError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C early-exit-transformer_b200`. There is no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    for name, args in _SIGS_EXPERIMENTAL.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.argtypes = args
            fn.restype = i32
    lib.eec_last_error.restype = C.c_char_p
    lib.eec_last_error.argtypes = []
    lib.eec_version.restype = i32
    lib.eec_device_ok.restype = i32
    lib.eec_launch_count.restype = C.c_longlong
    lib.eec_launch_count.argtypes = []
    lib.eec_ctc_workspace_bytes.restype = i64
    lib.eec_ctc_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.eec_ctc_beam_workspace_bytes.restype = i64
    lib.eec_ctc_beam_workspace_bytes.argtypes = [i32, i32, i32]
    lib.eec_dwconv_bwd_workspace_bytes.restype = i64
    lib.eec_dwconv_bwd_workspace_bytes.argtypes = [i32, i32, i32]
    _lib = lib
    return lib


def call(name: str, *args):
    lib = load()
    fn = getattr(lib, name, None)
    if fn is None:
        raise EecError(f"{name} is not in {LIB_PATH} (experimental entry points: make -C early-exit-transformer_b200 experiments; EEC_LIB=eec/libeec_exp.so)")
    rc = fn(*args)
    if rc != 0:
        raise EecError(f"{name} failed ({rc}): {lib.eec_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    """torch's current stream of the CURRENT device.  Every tensor handed to a kernel must live on that device: `call` is always
    reached through a model / op whose entry point has switched to the tensors' device (eec.early_exit._on_device), so a model on
    cuda:1 works without a global torch.cuda.set_device(1)."""
    return torch.cuda.current_stream().cuda_stream


def on_device(dev):
    """Context: make `dev` torch's current CUDA device (so `stream()` is the stream of the tensors' device); a CPU device passes through
    -- the entry point it guards raises the "no CPU path" error itself."""
    import contextlib
    dev = torch.device(dev)
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise EecError(f"unsupported dtype {t.dtype}")
