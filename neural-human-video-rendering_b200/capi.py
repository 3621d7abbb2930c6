"""ctypes binding of the C-ABI in include/nhvr.h (libnhvr_sm100.so).

This is the ONLY way the Python host reaches the kernels: plain pointers and sizes, the stream of
``torch.cuda.current_stream()``.  Loading fails loudly if the library has not been built
(``__graft_entry__.build()`` / ``make -C neural-human-video-rendering_b200/csrc``); there is no
fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnhvr_sm100.so")

# enums (include/nhvr.h)
HALO_ZERO, HALO_REFLECT = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU02, ACT_TANH, ACT_TANH_SIGMOID_LAST = 0, 1, 2, 3, 4
CONV, CONV_TRANSPOSE, CONV_DGRAD_S1 = 0, 1, 2
EPI_RAW_STATS, EPI_BIAS_ACT_F32, EPI_BIAS_ACT_P8, EPI_RAW_P8 = 0, 1, 2, 3


class ActDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("N", "C8", "H", "W", "pad_t", "pad_l", "pad_b", "pad_r", "split", "halo", "hilo")]

    def copy(self) -> "ActDesc":
        d = ActDesc()
        C.memmove(C.byref(d), C.byref(self), C.sizeof(ActDesc))
        return d


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "Cin", "Cout", "kh", "kw", "stride", "pad", "N", "H", "W", "halo", "epilogue", "act", "in_extra_rows", "in_extra_cols", "out_h", "out_w", "flags")]


# every symbol include/nhvr.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "nhvr_version": (C.c_int, []),
    "nhvr_strerror": (C.c_char_p, [C.c_int]),
    "nhvr_last_cuda_error": (C.c_char_p, []),
    "nhvr_arch_ok": (C.c_int, []),
    "nhvr_launch_count": (C.c_uint64, []),
    "nhvr_set_operand_dtype": (C.c_int, [C.c_int]),
    "nhvr_get_operand_dtype": (C.c_int, []),
    "nhvr_set_overflow_flag": (C.c_int, [_P]),
    "nhvr_act_bytes": (C.c_size_t, [C.POINTER(ActDesc)]),
    "nhvr_pack_nchw": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_int32), C.c_int32, _P, C.POINTER(ActDesc), _P]),
    "nhvr_stem_stat_shift": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, _P, _P]),
    "nhvr_unpack_nchw": (C.c_int, [_P, C.POINTER(ActDesc), _P, C.c_int32, _P]),
    "nhvr_conv_plan_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_P)]),
    "nhvr_conv_plan_destroy": (None, [_P]),
    "nhvr_conv_input_desc": (C.c_int, [_P, C.POINTER(ActDesc)]),
    "nhvr_conv_output_dims": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "nhvr_conv_plan_set_input_desc": (C.c_int, [_P, C.POINTER(ActDesc)]),
    "nhvr_wgrad_plan_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_P)]),
    "nhvr_wgrad_plan_destroy": (None, [_P]),
    "nhvr_wgrad_grad_desc": (C.c_int, [_P, C.POINTER(ActDesc)]),
    "nhvr_wgrad_workspace_bytes": (C.c_size_t, [_P]),
    "nhvr_wgrad": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, C.c_int32, _P]),
    "nhvr_conv_weight_bytes": (C.c_size_t, [_P]),
    "nhvr_conv_flops": (C.c_double, [_P]),
    "nhvr_conv_plan_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int32]),
    "nhvr_conv_pack_weights": (C.c_int, [_P, _P, _P, _P]),
    "nhvr_conv_pack_record_bytes": (C.c_size_t, []),
    "nhvr_conv_pack_record_fill": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_int64)]),
    "nhvr_conv_pack_weights_batched": (C.c_int, [_P, C.c_int32, C.c_int64, _P]),
    "nhvr_conv_forward": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(ActDesc), _P, _P]),
    "nhvr_conv_in_fused_supported": (C.c_int, [_P]),
    "nhvr_conv_forward_in_fused": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int32, _P, C.POINTER(ActDesc), _P, C.POINTER(ActDesc),
                                             _P, _P]),
    "nhvr_in_apply": (C.c_int, [_P, C.POINTER(ActDesc), _P, C.c_float, C.c_int32, _P, C.POINTER(ActDesc), _P,
                                C.POINTER(ActDesc), _P]),
    "nhvr_pose_rasterize": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, _P, _P, _P]),
    "nhvr_texture_sample": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P,
                                      _P, _P]),
    "nhvr_texture_unfold": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_texture_unfold_finish": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_float, _P, _P]),
    "nhvr_composite": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_texture_sample_bwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "nhvr_composite_bwd": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "nhvr_loss_pair_bwd": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_float, C.c_float, _P, C.c_int32, _P, _P]),
    "nhvr_avgpool3s2_bwd": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_in_bwd": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.POINTER(ActDesc), _P, C.c_float,
                            C.c_int32, _P, _P, C.POINTER(ActDesc), _P, _P]),
    "nhvr_act_bwd": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.POINTER(ActDesc), C.c_int32, _P,
                             C.POINTER(ActDesc), _P, _P]),
    "nhvr_fold_unpack": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(ActDesc), _P, C.c_int32,
                                 C.c_float, _P]),
    "nhvr_head_bwd": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P]),
    "nhvr_bias_grad": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P, _P]),
    "nhvr_loss_sum_sq_diff": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "nhvr_loss_sum_abs_diff": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "nhvr_loss_sum_sq_const": (C.c_int, [_P, C.c_float, C.c_int64, _P, _P]),
    "nhvr_loss_uv_prob": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_loss_uv_prob_bwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_float, C.c_float, _P, _P, _P]),
    "nhvr_loss_temporal": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_loss_temporal_bwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P]),
    "nhvr_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _P]),
    "nhvr_avgpool3s2": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_maxpool2": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nhvr_maxpool2_bwd": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
}

# fp16 operands are the default: measured on B200 at 512^2 against the fp32 oracle the temporal generator is
# within max-abs 1e-2 / 63 dB with fp16 operands but 8e-2 / 45 dB with bf16 (profiles/r01_precision.log), and
# only the former meets the 2e-2 / 45 dB parity bar.  Same tcgen05 kind::f16 rate.  NHVR_OPERAND=bf16 selects bf16.
DEFAULT_OPERAND = "f16"
_lib = None


class NhvrError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library and bind every declared symbol; raises if it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NhvrError(
                "libnhvr_sm100.so not built at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        lib.nhvr_set_operand_dtype(1 if os.environ.get("NHVR_OPERAND", DEFAULT_OPERAND) == "f16" else 0)
    return _lib


def set_operand_dtype(name: str) -> None:
    """'bf16' or 'f16': element type of activations / packed weights (see include/nhvr.h).  Engines built
    under one setting must not be reused under the other (weights are re-packed only on version change)."""
    assert name in ("bf16", "f16")
    load().nhvr_set_operand_dtype(1 if name == "f16" else 0)


def operand_dtype() -> str:
    return "f16" if load().nhvr_get_operand_dtype() else "bf16"


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.nhvr_strerror(status).decode()
        cuda = lib.nhvr_last_cuda_error().decode()
        raise NhvrError("%s failed: %s%s" % (what or "nhvr call", msg, (" [" + cuda + "]") if cuda else ""))


def require_device() -> None:
    """The product path needs a CUDA sm_100 device; fail loudly otherwise."""
    if not torch.cuda.is_available():
        raise NhvrError("nhvr_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    check(load().nhvr_arch_ok(), "nhvr_arch_ok")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


_ovf = {}


def overflow_flag(device=None) -> torch.Tensor:
    """The device int32 the kernels OR with 1 when they read a non-finite 16-bit value (an fp16 overflow of a conv
    output or gradient; include/nhvr.h nhvr_set_overflow_flag).  One flag per device, registered on first use."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _ovf:
        _ovf[idx] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", idx))
        check(load().nhvr_set_overflow_flag(_ovf[idx].data_ptr()), "nhvr_set_overflow_flag")
    return _ovf[idx]


def check_overflow(device=None, what: str = "") -> None:
    """Synchronising read of the range guard: raises if a 16-bit conv output / gradient overflowed since the last check."""
    flag = overflow_flag(device)
    if int(flag.item()) != 0:
        flag.zero_()
        raise NhvrError("16-bit operand overflow%s: a conv output or gradient exceeded the %s range (|x| > 65504 for fp16); "
                        "use NHVR_OPERAND=bf16 or rescale the weights" % ((" in " + what) if what else "", operand_dtype()))


def launch_count() -> int:
    return int(load().nhvr_launch_count())
