"""ctypes binding of libb200bd.so (the C-ABI declared in include/b200bd.h).

The product path has no CPU / PyTorch fallback: if the shared library is missing or the device is not an
sm_100 part, importing the compute entry points raises immediately ("fail loudly").
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200bd.so")

BD_OK, BD_ERR_INVALID, BD_ERR_CUDA, BD_ERR_UNSUPPORTED = 0, -1, -2, -3
BD_CONV_S1, BD_CONV_S2_PAD01 = 0, 1
BD_OUT_F16, BD_OUT_F32 = 0, 1
BD_IMPL_AUTO, BD_IMPL_SIMT, BD_IMPL_UMMA, BD_IMPL_UMMA_TILE = 0, 1, 2, 3

vp, i64, i32, f32, u64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_size_t


class ConvArgs(C.Structure):
    """struct bd_conv_args (include/b200bd.h)."""

    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32), ("ksize", i32), ("mode", i32), ("pad", i32),
        ("x", vp), ("ld_x", i64),
        ("w", vp),
        ("x2", vp), ("ld_x2", i64), ("Cin2", i32), ("w2", vp),
        ("bias", vp), ("bias2", vp),
        ("rowbias", vp), ("ld_rowbias", i64),
        ("residual", vp), ("ld_res", i64),
        ("out_scale", f32),
        ("y", vp), ("ld_y", i64), ("out_dtype", i32),
        ("impl", i32),
        ("gn_sums", vp), ("ld_sums", i64),
    ]


_SIGS = {
    "bd_version": (i32, []),
    "bd_init": (i32, []),
    "bd_last_error": (C.c_char_p, []),
    "bd_umma_error": (i32, []),
    "bd_device_supported": (i32, []),
    "bd_launch_count": (u64, []),
    "bd_batch_prep": (i32, [vp] * 12 + [i32] * 5 + [u64, u64, vp, vp]),
    "bd_batch_prep_u8": (i32, [vp] * 13 + [i32] * 5 + [u64, u64, vp, vp]),
    "bd_mse_workspace_floats": (sz, []),
    "bd_mse_fwd_bwd": (i32, [vp] * 6 + [sz, vp]),
    "bd_ddpm_step": (i32, [vp] * 6 + [sz, u64, u64, vp]),
    "bd_ddim_step": (i32, [vp] * 6 + [sz, u64, u64, vp]),
    "bd_pndm_step": (i32, [vp] * 6 + [sz, vp]),
    "bd_sampler_advance": (i32, [vp, vp, vp, i32, i32, vp]),
    "bd_finalize_images": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "bd_image_metrics": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "bd_temb_mlp": (i32, [vp] * 9 + [i32, i32, i32, i32, vp, vp]),
    "bd_sgemm": (i32, [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i32, i32, i32, i32, i32, vp]),
    "bd_gn_workspace_floats": (sz, [i32, i32]),
    "bd_groupnorm_fwd": (i32, [vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]),
    "bd_groupnorm_bwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, i32, vp]),
    "bd_conv_fwd": (i32, [C.POINTER(ConvArgs), vp]),
    "bd_conv_fwd_gn_sums_supported": (i32, [C.POINTER(ConvArgs)]),
    "bd_groupnorm_apply_sums": (i32, [vp, i64, vp, i64, vp, vp, vp, i64, vp, i32, i32, i32, i32, f32, i32, vp]),
    "bd_conv_dgrad": (i32, [C.POINTER(ConvArgs), vp]),
    "bd_conv_wgrad": (i32, [vp, i64, vp, i64, vp, vp] + [i32] * 10 + [vp]),
    "bd_pack_conv_weight": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "bd_cast_f32_to_f16": (i32, [vp, vp, sz, vp]),
    "bd_colsum_f16": (i32, [vp, i64, vp, i64, i32, i64, i32, i32, vp]),
    "bd_bias_from_gsum": (i32, [vp, i32, i32, i32, vp]),
    "bd_silu_bwd_f32": (i32, [vp, vp, vp, sz, vp]),
    "bd_silu_f32_to_f16": (i32, [vp, vp, sz, vp]),
    "bd_conv_in_fwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
    "bd_conv_in_fwd_gn_sums_supported": (i32, [i32, i32, i32, i32]),
    "bd_conv_in_fwd_sums": (i32, [vp, vp, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp]),
    "bd_conv_in_wgrad": (i32, [vp, vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "bd_conv_out_fwd": (i32, [vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "bd_conv_out_bwd": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "bd_upsample2x": (i32, [vp, i64, vp, i64, i32, i32, i32, i32, vp]),
    "bd_upsample2x_bwd": (i32, [vp, i64, vp, i64, i32, i32, i32, i32, vp]),
    "bd_add_f16": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, vp]),
    "bd_attention_fwd_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "bd_attention_bwd_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "bd_attention_fwd": (i32, [vp, i64, vp, vp, i64, vp, i32, i32, i32, i32, f32, i32, vp]),
    "bd_attention_bwd": (i32, [vp, i64, vp, vp, i64, vp, i64, vp, i32, i32, i32, i32, f32, i32, vp]),
    "bd_gradnorm_workspace_floats": (sz, []),
    "bd_grad_norm": (i32, [vp, sz, vp, vp, vp]),
    "bd_adam_step": (i32, [vp, vp, vp, vp, sz, vp, i32, f32, f32, f32, f32, f32, vp, vp, vp]),
    "bd_scaler_update": (i32, [vp, vp, f32, f32, i32, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


class B200BDError(RuntimeError):
    pass


def load():
    """dlopen the library and declare every prototype.  No GPU needed (symbol-export check on CPU)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200BDError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  baddiffusion_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here = header / library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_ready = False


def lib():
    """Library handle for compute calls; requires a CUDA sm_100 device."""
    global _ready
    l = load()
    if not _ready:
        import torch

        if not torch.cuda.is_available():
            raise B200BDError("baddiffusion_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make sure the primary context exists before the library touches it
        if not l.bd_device_supported():
            raise B200BDError("baddiffusion_b200 kernels are built for sm_100a only; this device is not CC 10.x")
        check(l.bd_init())
        _ready = True
    return l


def check(rc: int):
    if rc == BD_OK:
        return
    msg = load().bd_last_error().decode("utf-8", "replace")
    if rc == BD_ERR_INVALID:
        raise ValueError(msg)
    if rc == BD_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise B200BDError(msg)
