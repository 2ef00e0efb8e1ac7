"""ctypes binding of ``libconvnp_b200.so`` (the C-ABI declared in ``include/convnp_b200.h``).

The library is the product: there is no CPU or PyTorch fallback.  ``lib()`` raises if the shared
object is missing (run ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
deepsensornz_b200/csrc``) and ``check_device()`` raises if the current GPU is not sm_100.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libconvnp_b200.so")

c_stream = C.c_void_p
c_fp = C.c_void_p  # device pointers are passed as integers (tensor.data_ptr())


class CnpMlpParams(C.Structure):
    MAX_LAYERS = 6
    _fields_ = [
        ("W", C.c_void_p * 6),
        ("b", C.c_void_p * 6),
        ("dW", C.c_void_p * 6),
        ("db", C.c_void_p * 6),
        ("dims", C.c_int * 7),
        ("n_layers", C.c_int),
        ("likelihood", C.c_int),
    ]


class CnpBlk(C.Structure):
    """View of a blocked bf16 activation tensor [B][C/8][H+4][W+4][8]."""
    _fields_ = [
        ("base", C.c_void_p),
        ("bstride", C.c_longlong),
        ("cb_off", C.c_int),
        ("H", C.c_int),
        ("W", C.c_int),
    ]


class CnpConvOut(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("blk", CnpBlk),
        ("f32", C.c_void_p),
        ("f32_bstride", C.c_longlong),
        ("f32_ch_off", C.c_int),
        ("sy", C.c_int), ("ay", C.c_int), ("sx", C.c_int), ("ax", C.c_int),
        ("bias", C.c_void_p),
        ("relu", C.c_int),
        ("mask", C.POINTER(CnpBlk)),
        ("accumulate", C.c_int),
        ("s2d", C.POINTER(CnpBlk)),
        ("s2d_c0", C.c_int),
        ("s2d_band", C.c_int),
    ]


class CnpEncSet(C.Structure):
    _fields_ = [("kind", C.c_int), ("C", C.c_int), ("ch_off", C.c_int), ("batched", C.c_int),
                ("x1", C.c_void_p), ("x2", C.c_void_p), ("y", C.c_void_p), ("mask", C.c_void_p),
                ("N1", C.c_int), ("N2", C.c_int), ("mono1", C.c_int), ("mono2", C.c_int),
                ("scale2", C.c_float), ("KB", C.c_int), ("tab_i", C.c_void_p), ("tab_w", C.c_void_p),
                ("T", C.c_void_p), ("V", C.c_void_p), ("V_bs", C.c_longlong)]


class CnpEncSets(C.Structure):
    _fields_ = [("n_sets", C.c_int), ("pad_", C.c_int), ("s", CnpEncSet * 8)]


LIKELIHOODS = {"cnp": 0, "het": 0, "bernoulli-gamma": 1, "cnp-spikes-beta": 2, "spikes-beta": 2}
LIK_CHANNELS = {0: 2, 1: 4, 2: 5}      # head inputs per target variable

# conv_tc kinds (must match conv_bf16.cu)
KIND_K5S1, KIND_K1, KIND_K5S2, KIND_K5S1_DGRAD, KIND_K1_DGRAD, KIND_K5S2_DGRAD, KIND_UP_PHASE, KIND_UP_PHASE_DGRAD = range(8)
# conv_tc_wgrad kinds (must match wgrad_bf16.cu)
WG_K5S1, WG_K1, WG_K5S2, WG_K5S1_NARROW, WG_UP_PHASE, WG_K5S1_T = range(6)

_i, _d, _f, _ll = C.c_int, C.c_double, C.c_float, C.c_longlong
_GRID = [_d, _i, _d, _i, _d]  # start1, n1, start2, n2, res

_SIGS = {
    "cnp_version": (C.c_int, []),
    "cnp_last_error": (C.c_char_p, []),
    "cnp_check_device": (C.c_int, []),
    # (1) SetConv encoder
    "cnp_setconv_enc_offgrid_fwd": (C.c_int, [c_fp, c_fp, c_fp, _i, _i, _i] + _GRID + [_f, _f, c_fp, _i, _i, c_stream]),
    "cnp_setconv_enc_grid_workspace_bytes": (_ll, [_i, _i, _i, _i, _i, _i]),
    "cnp_setconv_enc_grid_fwd": (C.c_int, [c_fp, c_fp, _i, c_fp, c_fp, _i, _i, _i, _i, _i, _i] + _GRID +
                                 [_f, _f, c_fp, _i, _i, _i, c_fp, _ll, c_stream]),
    "cnp_encode_vpass": (C.c_int, [C.POINTER(CnpEncSets), _i, _i, _i, _f, c_stream]),
    "cnp_encode_hpass": (C.c_int, [C.POINTER(CnpEncSets), _i, _i, _i, c_stream]),
    "cnp_encode_tables": (C.c_int, [c_fp, c_fp, _i, _i, _i, _i] + _GRID + [_f, _i, c_fp, c_fp, c_stream]),
    "cnp_encode_fused": (C.c_int, [C.POINTER(CnpEncSets), _i] + _GRID + [_f, _i, c_fp, _ll, _i, C.POINTER(CnpBlk), _i,
                                   c_stream]),
    # (3) SetConv decoder
    "cnp_setconv_dec_offgrid_fwd": (C.c_int, [c_fp, _ll, c_fp, _i, _i, _i] + _GRID + [_f, c_fp, _i, c_stream]),
    "cnp_setconv_dec_offgrid_bwd": (C.c_int, [c_fp, _i, c_fp, _i, _i, _i] + _GRID + [_f, c_fp, _ll, c_stream]),
    # (2) fp32 UNet blocks
    "cnp_conv2d_fwd_f32": (C.c_int, [c_fp, _ll, c_fp, c_fp, c_fp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, c_stream]),
    "cnp_conv2d_dgrad_f32": (C.c_int, [c_fp, _ll, c_fp, c_fp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, c_stream]),
    "cnp_conv2d_wgrad_f32": (C.c_int, [c_fp, _ll, c_fp, _ll, c_fp, c_fp, _i, _i, _i, _i, _i, _i, _i, c_stream]),
    "cnp_relu_bwd_f32": (C.c_int, [c_fp, _ll, c_fp, _ll, _i, _ll, c_stream]),
    "cnp_upsample2x_fwd_f32": (C.c_int, [c_fp, _ll, c_fp, _ll, _i, _i, _i, _i, c_stream]),
    "cnp_upsample2x_bwd_f32": (C.c_int, [c_fp, _ll, c_fp, _ll, _i, _i, _i, _i, _i, c_stream]),
    "cnp_setconv_dec_grid_workspace_bytes": (_ll, [_i, _i, _i, _i, _i]),
    "cnp_setconv_dec_grid_fwd": (C.c_int, [c_fp, _ll, c_fp, c_fp, _i, _i, _i, _i] + _GRID + [_f, c_fp, _ll, c_fp, _ll,
                                           c_stream]),
    "cnp_mlp_head_points_fwd": (C.c_int, [C.POINTER(CnpMlpParams), c_fp, _ll, _i, c_fp, _ll, _i, _i, _ll, c_fp, c_fp,
                                          c_stream]),
    "cnp_dec_blk_fwd": (C.c_int, [C.POINTER(CnpBlk), c_fp, _i, _i, _d, _d, _d, _f, c_fp, c_fp, _i, c_fp, c_fp, c_fp,
                                  c_stream]),
    "cnp_dec_blk_bwd_params": (C.c_int, [c_fp, c_fp, c_fp, c_fp, _i, _i, _i, c_fp, c_fp, c_fp, c_stream]),
    "cnp_dec_blk_bwd_data": (C.c_int, [c_fp, c_fp, _i, _i, _d, _d, _d, _f, C.POINTER(CnpBlk), C.POINTER(CnpBlk),
                                       c_stream]),
    "cnp_decode_grid_fused_workspace_bytes": (_ll, [_i, _i, _i, _i]),
    "cnp_decode_grid_fused_fwd": (C.c_int, [C.POINTER(CnpBlk), c_fp, c_fp, _i, _i, _i, _d, _d, _d, _f, c_fp, c_fp,
                                            C.POINTER(CnpMlpParams), c_fp, _ll, _i, c_fp, c_fp, c_fp, _ll, c_stream]),
    "cnp_decode_grid_tc_workspace_bytes": (_ll, [_i, _i, _i, _i]),
    "cnp_decode_grid_tc_fwd": (C.c_int, [C.POINTER(CnpBlk), c_fp, c_fp, _i, _i, _i, _d, _d, _d, _f, c_fp, c_fp,
                                         C.POINTER(CnpMlpParams), c_fp, _ll, _i, c_fp, c_fp, c_fp, _ll, c_stream]),
    # (4) MLP + Gaussian head + NLL
    "cnp_mlp_head_fwd": (C.c_int, [C.POINTER(CnpMlpParams), c_fp, _i, _i, c_fp, _i, c_fp, _i, _i, c_fp, c_fp, c_fp, c_fp,
                                   c_fp, c_stream]),
    "cnp_mlp_head_bwd_workspace_bytes": (_ll, [C.POINTER(CnpMlpParams), _i, _i]),
    "cnp_mlp_head_bwd": (C.c_int, [C.POINTER(CnpMlpParams), c_fp, _i, _i, c_fp, _i, c_fp, _i, _i, c_fp, c_fp, c_fp, _ll,
                                   c_stream]),
    "cnp_loss_mean": (C.c_int, [c_fp, c_fp, _i, _i, c_fp, c_fp, c_stream]),
    # (2) bf16 tensor-core UNet blocks
    "cnp_conv_tc2_debug": (C.c_int, [c_fp, _i]),
    "cnp_conv_tc2_set_cluster": (C.c_int, [_i]),
    "cnp_conv_tc2_packed_bytes": (_ll, [_i, _i, _i]),
    "cnp_conv_tc2_pack": (C.c_int, [c_fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, c_fp, c_stream]),
    "cnp_conv_tc2": (C.c_int, [C.POINTER(CnpBlk), _i, c_fp, _i, _i, _i, _i, C.POINTER(CnpConvOut), _i, c_stream]),
    "cnp_conv_tc2_w2": (C.c_int, [C.POINTER(CnpBlk), _i, c_fp, c_fp, _i, _i, _i, _i, _i, C.POINTER(CnpConvOut), _i, c_stream]),
    "cnp_blk_from_nchw_f32": (C.c_int, [c_fp, _ll, _i, _i, _i, _i, C.POINTER(CnpBlk), c_stream]),
    "cnp_blk_from_nchw_f32_ones": (C.c_int, [c_fp, _ll, _i, _i, _i, _i, C.POINTER(CnpBlk), _i, C.c_ulonglong, c_stream]),
    "cnp_up_phase_weights": (C.c_int, [c_fp, _i, _i, c_fp, c_stream]),
    "cnp_up_strips_fwd": (C.c_int, [C.POINTER(CnpBlk), _i, C.POINTER(CnpBlk), C.POINTER(CnpBlk), _i, c_stream]),
    "cnp_up_strips_scatter": (C.c_int, [C.POINTER(CnpBlk), C.POINTER(CnpBlk), C.POINTER(CnpBlk), _i, c_stream]),
    "cnp_up_dy_split": (C.c_int, [C.POINTER(CnpBlk), C.POINTER(CnpBlk), C.POINTER(CnpBlk), C.POINTER(CnpBlk), _i, c_stream]),
    "cnp_up_strips_bwd_fold": (C.c_int, [C.POINTER(CnpBlk), C.POINTER(CnpBlk), C.POINTER(CnpBlk), C.POINTER(CnpBlk), _i, _i,
                                         c_stream]),
    "cnp_up_wgrad_fold": (C.c_int, [c_fp, c_fp, _i, _i, c_fp, c_stream]),
    "cnp_fold_in_fwd": (C.c_int, [c_fp, c_fp, c_fp, _i, _i, _i, _i, _i, c_fp, c_stream]),
    "cnp_fold_in_bwd": (C.c_int, [c_fp, c_fp, c_fp, c_fp, _i, _i, _i, _i, _i, c_fp, c_fp, c_fp, c_stream]),
    "cnp_blk_to_nchw_f32": (C.c_int, [C.POINTER(CnpBlk), _i, _i, c_fp, _ll, c_stream]),
    "cnp_conv1x1_in_bf16": (C.c_int, [c_fp, _ll, c_fp, c_fp, _i, _i, _i, C.POINTER(CnpBlk), c_stream]),
    "cnp_blk_upsample2x_fwd": (C.c_int, [C.POINTER(CnpBlk), _i, C.POINTER(CnpBlk), _i, c_stream]),
    "cnp_blk_upsample2x_bwd": (C.c_int, [C.POINTER(CnpBlk), _i, C.POINTER(CnpBlk), C.POINTER(CnpBlk), _i, _i, c_stream]),
    "cnp_blk_space_to_depth": (C.c_int, [C.POINTER(CnpBlk), _i, C.POINTER(CnpBlk), _i, c_stream]),
    "cnp_conv_tc_wgrad_workspace_bytes": (_ll, []),
    "cnp_conv_tc_wgrad": (C.c_int, [C.POINTER(CnpBlk), _i, C.POINTER(CnpBlk), _i, c_fp, c_fp, _i, _i, c_fp, _ll, c_stream]),
    "cnp_conv_tc_wgrad_pair": (C.c_int, [C.POINTER(CnpBlk), _i, C.POINTER(CnpBlk), c_fp, c_fp, _i, c_fp, _i, _i, c_fp, _ll,
                                         c_stream]),
    "cnp_conv1x1_in_wgrad": (C.c_int, [c_fp, _ll, _i, C.POINTER(CnpBlk), _i, c_fp, c_fp, c_stream]),
    "cnp_blk_channel_sum": (C.c_int, [C.POINTER(CnpBlk), _i, _i, c_fp, c_stream]),
}

# only in a `make LEGACY=1` build (first convolution formulation, A/B runs and its own tests)
_OPTIONAL_SIGS = {
    "cnp_conv_tc_packed_bytes": (_ll, [_i, _i]),
    "cnp_conv_tc_pack": (C.c_int, [c_fp, _i, _i, _i, _i, _i, _i, _i, _i, c_fp, c_stream]),
    "cnp_conv_tc": (C.c_int, [C.POINTER(CnpBlk), _i, c_fp, _i, _i, _i, C.POINTER(CnpConvOut), _i, c_stream]),
}

_lib: Optional[C.CDLL] = None


class CnpError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CnpError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                           "(there is no CPU fallback for the ConvNP hot path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in _OPTIONAL_SIGS.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().cnp_last_error().decode("utf-8", "replace")
        raise CnpError(f"{what or 'libconvnp_b200'} failed (rc={rc}): {msg}")


def check_device() -> None:
    check(lib().cnp_check_device(), "cnp_check_device")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)
