"""ctypes binding of the C ABI declared in include/b200cd.h (libb200cd.so, built in-tree by build.py).

There is no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libb200cd.so"

_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_sz = C.c_size_t


class GradSrc(C.Structure):
    """b200cd_grad_src (include/b200cd.h)."""

    _fields_ = [
        ("kind", C.c_int32),
        ("ptr", C.c_void_p),
        ("w", C.c_void_p),
        ("ld", C.c_int64),
        ("n_mod", C.c_int32),
        ("scale_lo", C.c_float),
        ("scale_hi", C.c_float),
    ]


# name -> (restype, argtypes); must list every symbol of include/b200cd.h
SIGNATURES = {
    "b200cd_abi_version": (_i, []),
    "b200cd_last_error": (C.c_char_p, []),
    "b200cd_init": (_i, [_i]),
    "b200cd_device_status": (_i, [_i, _vp]),
    "b200cd_pack_input": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "b200cd_pack_weights": (_i, [_i, _vp, _vp, _i, _i, _i, _vp]),
    "b200cd_pack_job_blocks": (_i, [_i, _i, _i, _i]),
    "b200cd_pack_weights_batched": (_i, [_vp, _i, _i64, _vp]),
    "b200cd_conv_gemm": (_i, [_i, _i, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i, _i, _vp, _i64, _vp, _vp, _vp]),
    "b200cd_conv_gemm_bnbwd": (_i, [_i, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "b200cd_conv_gemm_tiles": (_i, [_i, _i]),
    "b200cd_conv_gemm_stat_rows": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "b200cd_wgrad_gemm": (_i, [_i, _i, _i, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i, _i, _i64, _i64, _i64, _i64, _vp]),
    "b200cd_wgrad_tiles": (_i, [_i, _i, _i]),
    "b200cd_wgrad_ctas_per_split": (_i, [_i, _i, _i, _i]),
    "b200cd_wgrad_reduce": (_i, [_vp, _i, _i64, _i, _i, _i, _i, _vp, _vp]),
    "b200cd_bn_stats": (_i, [_vp, _i, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "b200cd_bn_apply": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "b200cd_bn_bwd": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(GradSrc), _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b200cd_bn_bwd_from_sums": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(GradSrc), _i, _i, _i, _i, _i, _vp, _i, _vp, _vp,
                                     _vp, _vp, _i64, _vp]),
    "b200cd_bn_bwd_ws_floats": (_sz, [_i, _i, _i, _i, _i]),
    "b200cd_head_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp, _i64, _vp, _vp]),
    "b200cd_pad_copy": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _i, _vp]),
    "b200cd_colsum": (_i, [_vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _vp]),
    "b200cd_stat_rowsum": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "b200cd_pj_fwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i64, _i, _vp, _vp, _vp]),
    "b200cd_pj_loss": (_i, [_vp, _vp, _vp]),
    "b200cd_reduce_job_parts": (_i, [_i, _i, _i]),
    "b200cd_reduce_job_blocks": (_i64, [_i, _i, _i, _i]),
    "b200cd_wgrad_reduce_batched": (_i, [_vp, _i, _i64, _vp]),
    "b200cd_confusion_counts": (_i, [_vp, _vp, _i64, _i, _vp, _i, _vp, _vp]),
    "b200cd_adamw_step": (_i, [_vp, _i, _i64, _d, _d, _d, _d, _d, _i64, _vp]),
    "b200cd_pj_bwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i64, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "b200cd_query_workspace": (_i64, [_i, C.POINTER(C.c_int64), _i]),
    "b200cd_comm_load": (_i, [C.c_char_p]),
    "b200cd_comm_version": (_i, []),
    "b200cd_comm_unique_id": (_i, [_vp]),
    "b200cd_comm_init": (_i, [_vp, _i, _i]),
    "b200cd_comm_size": (_i, []),
    "b200cd_allreduce_bucket": (_i, [_vp, _i64, _vp]),
    "b200cd_allreduce_f64": (_i, [_vp, _i64, _vp]),
    "b200cd_comm_destroy": (_i, []),
    "b200cd_graph_instantiate": (_i, [_vp, _i, C.POINTER(C.c_void_p)]),
    "b200cd_graph_launch": (_i, [_vp, _vp]),
    "b200cd_graph_exec_destroy": (_i, [_vp]),
    "b200cd_augment": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "b200cd_bn_eval_affine_batched": (_i, [_vp, _i, _i, _vp]),
    "b200cd_conv_gemm_affine": (_i, [_i, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _vp]),
    # split-bf16 ("precise") mode, ABI version 2
    "b200cd_wgrad_gemm_hp": (_i, [_i, _i, _i, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i, _i, _i64, _i64, _i64, _i64, _vp]),
    "b200cd_pack_input_hp": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "b200cd_pack_job_blocks_hp": (_i, [_i, _i, _i, _i]),
    "b200cd_pack_weights_hp_batched": (_i, [_vp, _i, _i64, _vp]),
    "b200cd_bn_apply_hp": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "b200cd_bn_bwd_hp": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(GradSrc), _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b200cd_head_fwd_hp": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp, _i64, _vp, _vp]),
    "b200cd_colsum_hp": (_i, [_vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None
_inited_devices: set[int] = set()


class B200CDError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libb200cd.so (no build here: __graft_entry__.build() / build.py produce it)."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise B200CDError(
                    f"{LIB_PATH} is missing: build it with `python -m multimodal_siamese_cd_b200.build` "
                    "(there is no CPU or PyTorch fallback for these kernels)")
            lib = C.CDLL(str(LIB_PATH))
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.b200cd_abi_version() != 2:
                raise B200CDError("libb200cd.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200cd_last_error().decode()
        raise B200CDError(f"b200cd error {rc}: {msg}")


def init(device: int) -> None:
    lib = load()
    if device not in _inited_devices:
        check(lib.b200cd_init(device))
        _inited_devices.add(device)


def device_status(device: int, stream: int) -> None:
    check(load().b200cd_device_status(device, stream))
