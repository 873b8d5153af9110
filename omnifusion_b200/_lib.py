"""ctypes binding of libofb.so (the C ABI declared in include/ofb.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call
fails, an exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFB_LIB") or os.path.join(_HERE, "libofb.so")   # OFB_LIB: A/B experiments with another build

LAYOUT_REF, LAYOUT_FOLDED, LAYOUT_STEM16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
FMT_F32, FMT_SPLIT16 = 0, 1


class OfbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [("in0", C.c_void_p), ("in1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
                ("n", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("wgt", C.c_void_p), ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("cout", C.c_int),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
                ("act", C.c_int), ("out", C.c_void_p), ("engine", C.c_int),
                ("in_fmt", C.c_int), ("out_fmt", C.c_int), ("wgt_split", C.c_void_p), ("wgt_unscale", C.c_float),
                ("ups2x", C.c_int), ("ksplit", C.c_int), ("partial", C.c_void_p)]


class Geometry(C.Structure):
    _fields_ = [("n_patch", C.c_int), ("patch", C.c_int), ("erp_h", C.c_int), ("erp_w", C.c_int),
                ("grid_hi", C.c_void_p), ("grid_lo", C.c_void_p), ("pts", C.c_void_p), ("pts_c", C.c_int),
                ("blend_rowptr", C.c_void_p), ("blend_idx", C.c_void_p), ("blend_w", C.c_void_p)]


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_int64 * 5)]


_P, _I = C.c_void_p, C.c_int
_SIGNATURES = {
    "ofb_version": (C.c_int, []),
    "ofb_last_error": (C.c_char_p, []),
    "ofb_set_device": (_I, [_I]),
    "ofb_equi2pers_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _I, _P]),
    "ofb_equi2pers_taps": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "ofb_pers2equi_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P]),
    "ofb_equi2pers_backward_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P]),
    "ofb_pers2equi_backward_f32": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P]),
    "ofb_blend_conf_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P]),
    "ofb_blend_conf_pairs_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P]),
    "ofb_heads_tc_pairs_f16": (_I, [_P, _I, _I, _I, _P, C.c_float, C.c_float, C.c_float, _P, _P]),
    "ofb_conv_f32": (_I, [C.POINTER(ConvDesc), _P]),
    "ofb_split_f16": (_I, [_P, C.c_size_t, C.c_float, _P, _P]),
    "ofb_merge_f16": (_I, [_P, C.c_size_t, _P, _P]),
    "ofb_stem_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "ofb_stem_tc_f16": (_I, [_P, _I, _I, _I, _P, C.c_float, _P, _P, _P, _P]),
    "ofb_maxpool3x3s2_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _P]),
    "ofb_upsample2x_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "ofb_point_embed_f32": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "ofb_token_pack_f32": (_I, [_P, _P, _I, _I, _P, _I, _P]),
    "ofb_layernorm_f32": (_I, [_P, _P, _P, _I, _I, C.c_float, _P, _I, _I, _P]),
    "ofb_splitk_finish_ln_f32": (_I, [_P, _I, C.c_float, _P, _P, _I, _I, _P, _P, _P, C.c_float, _P, _I, _P]),
    "ofb_attention_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "ofb_attention_qkv_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _P]),
    "ofb_attention_tc_f16": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "ofb_splitk_finish_conv_f16": (_I, [_P, _I, C.c_longlong, _I, _P, _P, C.c_float, _P, _I, _P, _P]),
    "ofb_token_stack_f32": (_I, [_P, _P, _P, C.c_longlong, _P, _I, _I, _I, _I, _P]),
    "ofb_token_stack_scratch_floats": (C.c_longlong, [_I, _I]),
    "ofb_token_stack_resident_groups": (_I, [_I]),
    "ofb_debug_token_stamps": (_I, [_P]),
    "ofb_heads_tc_f16": (_I, [_P, _I, _I, _I, _P, C.c_float, C.c_float, C.c_float, _I, _P, _P, _P]),
    "ofb_heads_f32": (_I, [_P, _I, _I, _I, _P, C.c_float, _P, C.c_float, _I, _P, _P, _I, _P]),
    "ofb_u8hwc_to_f32chw": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "ofb_loss_work_bytes": (C.c_longlong, [_I]),
    "ofb_depth_loss_f32": (_I, [_P, _P, _P, _P, _I, C.c_longlong, _I, _P, _P, _P, _P]),
    "ofb_depth_loss_backward_f32": (_I, [_P, _P, _P, _P, _I, C.c_longlong, _I, _P, _P, _P, _P]),
    "ofb_absrel_partial": (_I, [_P, _P, _P, C.c_size_t, C.c_float, _P, _P]),
    "ofb_depth_metrics_partial": (_I, [_P, _P, _P, C.c_size_t, C.c_float, _P, _P]),
    "ofb_depth_metrics_partial_ds": (_I, [_P, _P, _P, C.c_size_t, _P, _P, _P]),
    "ofb_area_resize_u8": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "ofb_masked_median_f32": (_I, [_P, _P, C.c_size_t, _P, _P, _P]),
    "ofb_median_scale_f32": (_I, [_P, _P, _P, C.c_size_t, _P, _P, _P]),
    "ofb_depth_to_points_f32": (_I, [_P, _P, _I, _I, _I, C.c_float, _P, _P]),
    "ofb_create": (_I, [_I, C.POINTER(_P)]),
    "ofb_destroy": (_I, [_P]),
    "ofb_set_geometry": (_I, [_P, C.POINTER(Geometry)]),
    "ofb_load_weights": (_I, [_P, C.POINTER(TensorDesc), _I, _I]),
    "ofb_forward_f32": (_I, [_P, _P, _I, _I, _I, C.POINTER(_P), _P]),
    "ofb_set_option": (_I, [_P, C.c_char_p, _I]),
    "ofb_get_activation": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64, C.POINTER(_I * 4), _P]),
    "ofb_launch_count": (C.c_int64, [_I]),
    "ofb_workspace_generation": (C.c_longlong, [_P]),
    "ofb_last_conv_variant": (C.c_char_p, []),
    "ofb_debug_stamps": (_I, [_P]),
    "ofb_debug_set": (_I, [_I]),
    "ofb_debug_nstack": (_I, [_I]),
    "ofb_debug_timeline": (_I, [_P, _I]),
    "ofb_debug_timeline_raw": (_I, [_P]),
    "ofb_range_report": (_I, [_P, C.c_char_p, _I]),
    "ofb_profile_enable": (_I, [_P, _I]),
    "ofb_profile_report": (_I, [_P, C.c_char_p, _I]),
}

_lib = None


def lib():
    """Loads libofb.so once.  Raises if it has not been built (python -c
    'import __graft_entry__ as g; g.build()' or make -C omnifusion_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfbError(f"{LIB_PATH} not found: build it with `make -C omnifusion_b200/csrc` "
                           "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                if os.environ.get("OFB_LIB"):      # A/B run against an older build: newer entry points are absent
                    continue
                raise
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc < 0:
        raise OfbError(lib().ofb_last_error().decode("utf-8", "replace"))
    return rc


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name, dtype=torch.float32, allow_grad=False):
    if not t.is_cuda:
        raise OfbError(f"{name} must be a CUDA tensor: omnifusion_b200 has no CPU path (got {t.device})")
    if t.dtype != dtype:
        raise OfbError(f"{name} must be {dtype} (got {t.dtype})")
    if torch.is_grad_enabled() and t.requires_grad and not allow_grad:
        raise OfbError(f"{name} requires grad: omnifusion_b200 implements inference only; use torch.no_grad()")
    return t.contiguous()


def use_device(device):
    check(lib().ofb_set_device(device.index if device.index is not None else torch.cuda.current_device()))
