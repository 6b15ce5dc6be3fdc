"""ctypes binding of libkdfusion_b200.so (the C ABI in include/kdfusion_b200.h).

There is no fallback: if the shared library is missing the import fails loudly,
and every wrapper refuses tensors that are not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_ROOT, "libkdfusion_b200.so")

KDF_F32, KDF_BF16 = 0, 1
REDUCE_MAX, REDUCE_MEAN = 0, 1

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

_SIGNATURES = {
    "kdf_abi_version": (C.c_int, []),
    "kdf_last_error": (C.c_char_p, []),
    "kdf_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "kdf_bev_index": (C.c_int, [_vp, _i, _i64, _i, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp]),
    "kdf_range_index": (C.c_int, [_vp, _i, _i64, _i, _f, _f, _i, _i, _vp, _vp, _vp]),
    "kdf_range_project_fwd": (C.c_int, [_vp, _i, _vp, _i, _i, _i64, _i, _f, _f, _i, _i, _i,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kdf_bev_rasterize": (C.c_int, [_vp, _i, _vp, _i, _i64, _f, _f, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp]),
    "kdf_bev_workspace_bytes": (_sz, [_i, _i64, _i, _i]),
    "kdf_bev_project_fwd": (C.c_int, [_vp, _i, _vp, _i, _i, _i64, _i, _f, _f, _f, _f, _i, _i, _i,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kdf_bev_build_order": (C.c_int, [_vp, _i, _i, _i64, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kdf_bev_reduce": (C.c_int, [_vp, _i, _vp, _vp, _i, _i64, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kdf_bev_project_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _vp]),
    "kdf_rowbn_workspace_bytes": (_sz, [_i]),
    "kdf_rowbn_bwd_workspace_bytes": (_sz, [_i]),
    "kdf_rowbn_stats": (C.c_int, [_vp, _i, _i64, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_rowbn_apply_fwd": (C.c_int, [_vp, _vp, _i, _i64, _i, _vp, _vp, _i, _vp, _vp]),
    "kdf_rowbn_bwd": (C.c_int, [_vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "kdf_bev_reduce_affine": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "kdf_bev_bwd_affine": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "kdf_cls_conv_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "kdf_cls_conv_bwd": (C.c_int, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kdf_stem_conv_fwd": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "kdf_stem_conv_bwd_weight": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "kdf_bev_build_sorted": (C.c_int, [_vp, _i, _i64, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kdf_bev_bwd_share": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kdf_mlp_layer_bwd_share": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_point_moments": (C.c_int, [_vp, _i64, _vp, _vp]),
    "kdf_mlp_layer_fwd": (C.c_int, [_i, _vp, _i64, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "kdf_mlp_eval3_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_dwconv3x3_fwd": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kdf_dwconv3x3_affine_fwd": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "kdf_dwconv3x3_bwd_data": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "kdf_dwconv3x3_bwd_weight": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "kdf_rows_axpb": (C.c_int, [_vp, _vp, _i, _i64, _i, _vp, _vp, _vp]),
    "kdf_fpn_merge_fwd": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "kdf_fpn_up2_bwd": (C.c_int, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "kdf_mlp_layer_bwd": (C.c_int, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "kdf_mlp_l1_stats": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_bn_bwd_coeffs": (C.c_int, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_mlp_l1_bwd": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_bn_finalize": (C.c_int, [_vp, _i64, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdf_pw_conv_fwd": (C.c_int, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "kdf_fusion_weighted_fwd": (C.c_int, [_vp, _vp, _i, _i64, _i] + [_vp] * 8 + [_vp, _vp, _vp]),
    "kdf_fusion_weighted_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i64, _i] + [_vp] * 8 + [_vp] + [_vp] * 7 + [_vp]),
    "kdf_fusion_affine_relu_pair_fwd": (C.c_int, [_vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "kdf_fusion_affine_relu_pair_bwd": (C.c_int, [_vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "kdf_kd_loss_workspace_bytes": (_sz, []),
    "kdf_kd_loss_fwd_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i64, _i, _f, _f, _f, _i64,
                                      _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _f, _vp, _vp, _vp, _vp]),
    "kdf_kd_label_count": (C.c_int, [_vp, _i, _i, _i64, _i64, _vp, _vp]),
    "kdf_kd_loss_fwd_bwd_counted": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i64, _i, _f, _f, _f, _i64,
                                              _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _f, _vp, _vp, _vp, _vp]),
    "kdf_confusion_matrix": (C.c_int, [_vp, _vp, _i, _i, _i64, _i, _i64, _vp, _vp]),
    "kdf_adamw_flat": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _f, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            f"`python {os.path.join(_PKG_ROOT, 'build.py')}` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and this table diverge
        fn.restype, fn.argtypes = res, args
    if lib.kdf_abi_version() != 1:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.kdf_abi_version()} != 1, rebuild it")
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"kdfusion_b200 {what} failed (code {rc}): {lib.kdf_last_error().decode()}")


# kernels launched per ABI call (memsets not counted) -- bench.py reports the total
KERNELS_PER_CALL = {
    "kdf_bev_index": 1, "kdf_range_index": 1, "kdf_range_project_fwd": 4, "kdf_bev_rasterize": 3, "kdf_bev_project_fwd": 4, "kdf_bev_reduce": 1, "kdf_bev_project_bwd": 1,
    "kdf_pw_conv_fwd": 1, "kdf_fusion_weighted_fwd": 1, "kdf_fusion_weighted_bwd": 1,
    "kdf_fusion_affine_relu_pair_fwd": 1, "kdf_fusion_affine_relu_pair_bwd": 1,
    "kdf_kd_loss_fwd_bwd": 2, "kdf_kd_label_count": 1, "kdf_kd_loss_fwd_bwd_counted": 1, "kdf_confusion_matrix": 1, "kdf_adamw_flat": 1,
    "kdf_rowbn_stats": 1, "kdf_rowbn_apply_fwd": 1, "kdf_rowbn_bwd": 2,
    "kdf_mlp_layer_fwd": 1, "kdf_mlp_eval3_fwd": 1, "kdf_mlp_layer_bwd": 1, "kdf_bn_finalize": 1, "kdf_bev_reduce_affine": 1, "kdf_bev_bwd_affine": 1,
    "kdf_bev_build_sorted": 3, "kdf_bev_bwd_share": 1, "kdf_mlp_layer_bwd_share": 1, "kdf_stem_conv_fwd": 1, "kdf_stem_conv_bwd_weight": 1, "kdf_cls_conv_fwd": 1, "kdf_cls_conv_bwd": 1,
    "kdf_point_moments": 1, "kdf_bev_build_order": 3, "kdf_fpn_merge_fwd": 1, "kdf_fpn_up2_bwd": 1, "kdf_rows_axpb": 1, "kdf_mlp_l1_stats": 1, "kdf_bn_bwd_coeffs": 1, "kdf_mlp_l1_bwd": 1, "kdf_dwconv3x3_fwd": 1, "kdf_dwconv3x3_affine_fwd": 1, "kdf_dwconv3x3_bwd_data": 1, "kdf_dwconv3x3_bwd_weight": 1,
}
launch_stats = {"kernels": 0, "calls": 0}


def call(name: str, *args) -> None:
    """Invoke one ABI entry point, raise on failure, account for its kernel launches."""
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"kdfusion_b200 {name} failed (code {rc}): {lib.kdf_last_error().decode()}")
    launch_stats["kernels"] += KERNELS_PER_CALL.get(name, 0)
    launch_stats["calls"] += 1


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return KDF_F32
    if t.dtype == torch.bfloat16:
        return KDF_BF16
    raise TypeError(f"kdfusion_b200 supports float32 and bfloat16 features, got {t.dtype}")


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("kdfusion_b200 kernels run on CUDA tensors only (no CPU fallback); "
                               f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ----------------------------------------------------------------------------- parameter generation
# The kernels update parameters and BatchNorm buffers through raw pointers (kdf_adamw_flat on the flat buffer, the
# running statistics inside kdf_rowbn_stats / kdf_bn_finalize / kdf_mlp_l1_stats -- also inside a replayed CUDA graph,
# where no Python runs at all), which never advances torch's per-tensor ``_version``.  Everything that caches a value
# derived from parameters (eval-mode BatchNorm affine, bf16 weight copies) therefore also keys on this counter, which
# every optimizer step, every training step and every running-statistics update advances.  Modules marked
# ``_kdf_frozen`` (the Trainer marks the frozen teacher) are exempt: nothing writes to them.
_generation = 0


def bump_generation() -> int:
    global _generation
    _generation += 1
    return _generation


def cache_generation(module) -> int:
    """The generation a cache entry of ``module`` is valid for (constant for frozen modules)."""
    return -1 if getattr(module, "_kdf_frozen", False) else _generation


def mark_frozen(module, frozen: bool = True) -> None:
    """Declare that no kernel writes to ``module``'s parameters / buffers (inference-only teacher)."""
    for m in module.modules():
        m._kdf_frozen = frozen


# ----------------------------------------------------------------------------- BatchNorm batch counters
# nn.BatchNorm increments num_batches_tracked once per training forward: 29 one-element kernels per student step.  Inside
# ``deferred_batch_counters()`` (the Trainer's step) the increments are collected and issued as ONE multi-tensor add.
_pending_counters = None


def bump_batch_counter(bn) -> float:
    """Advance bn.num_batches_tracked like nn.BatchNorm does and return the momentum of this update."""
    global _pending_counters
    bump_generation()                                 # the running statistics change behind torch's back
    if bn.momentum is not None and _pending_counters is not None:
        _pending_counters.append(bn.num_batches_tracked)
        return float(bn.momentum)
    bn.num_batches_tracked.add_(1)
    return float(bn.momentum) if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)


class deferred_batch_counters:
    def __enter__(self):
        global _pending_counters
        self._outer = _pending_counters
        _pending_counters = []
        return self

    def __exit__(self, *exc):
        global _pending_counters
        pending, _pending_counters = _pending_counters, self._outer
        if pending:
            import torch
            with torch.no_grad():
                torch._foreach_add_(pending, 1)
        return False
