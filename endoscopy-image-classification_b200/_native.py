"""ctypes binding of libb200ssl.so (include/b200ssl.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing, or a tensor is not a CUDA tensor, the callers raise.  Build the library
with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C
endoscopy-image-classification_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libb200ssl.so"

F32, BF16, F16, I64, I32, U8 = 0, 1, 2, 3, 4, 5
_TORCH2ENUM = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16,
               torch.int64: I64, torch.int32: I32, torch.uint8: U8, torch.bool: U8}

EMA_BLOCK_ELEMS = 4096


class EmaBlock(C.Structure):
    """Mirror of ``b200ssl_ema_block`` (32 bytes)."""
    _fields_ = [("ema", C.c_void_p), ("model", C.c_void_p), ("count", C.c_int32),
                ("dtype", C.c_int32), ("repeat", C.c_int32), ("reserved", C.c_int32)]


assert C.sizeof(EmaBlock) == 32

_vp, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/b200ssl.h declares
class BankShards(C.Structure):
    """``b200ssl_bank_shards`` of include/b200ssl.h (directly addressed rank-sharded bank)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("shard_rows", C.c_int64), ("arenas_host", C.c_void_p),
                ("arenas_dev", C.c_void_p), ("feats_offset", C.c_uint64), ("probs_offset", C.c_uint64),
                ("probs_t_offset", C.c_uint64), ("replicated", C.c_int32), ("reserved", C.c_int32)]


SIGNATURES = {
    "b200ssl_version": (_i32, []),
    "b200ssl_last_error_string": (C.c_char_p, []),
    "b200ssl_debug_set_timing_buffer": (None, [_vp]),
    "b200ssl_workspace_bytes": (_sz, [_i64, _i32, _i64]),
    "b200ssl_debug_smooth_plan": (_i32, [_i64, _i64, _i32, _vp]),
    "b200ssl_debug_set_k3": (None, [_i32, _i32, _i32, _i32]),
    "b200ssl_debug_set_k3_f32_simt": (None, [_i32]),
    "b200ssl_set_head_sm_budget": (_i32, [_i32]),
    "b200ssl_debug_max_active_clusters": (_i32, [_i32]),
    "b200ssl_fixmatch_head_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _i32,
                                             _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_scale_inplace": (_i32, [_vp, _i64, _i32, _vp, _f32, _vp]),
    "b200ssl_labeled_ce_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _sz, _vp]),
    "b200ssl_ce_rows_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _sz, _vp]),
    "b200ssl_scale_rows": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "b200ssl_bad_label_count": (_i32, [_vp, _vp, _i32]),
    "b200ssl_eval_head": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_normalize_views": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _i32, _vp]),
    "b200ssl_comatch_da": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_bank_smooth_partial": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _i32, _i32,
                                           _vp, _vp, _sz, _vp]),
    "b200ssl_comatch_finalize": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _i32, _f32, _f32, _f32, _f32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_comatch_rows_fused": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _vp, _vp,
                                          _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                          _i64, _i32, _vp, _i64, _i32, _vp]),
    "b200ssl_bank_enqueue": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i64, _vp, _i64,
                                    _i64, _i64, _i64, _i64, _vp]),
    "b200ssl_contrast_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _f32, _f32, _vp,
                                    _vp, _sz, _vp]),
    "b200ssl_contrast_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _f32, _vp, _f32, _vp, _vp,
                                    _vp, _i64, _vp, _f32, _vp, _sz, _vp]),
    "b200ssl_ema_multi_tensor": (_i32, [_vp, _i32, _i32, _i32, _f32, _f32, _i32, _vp]),
    "b200ssl_ema_multi_tensor_ctas": (_i32, [_vp, _i32, _i32, _i32, _f32, _f32, _i32, _i32, _vp]),
    "b200ssl_stream_delay": (_i32, [_i64, _vp]),
    "b200ssl_probe_sm_set": (_i32, [_i32, _i32, _vp, _vp, _vp]),
    "b200ssl_ema_multi_tensor_masked": (_i32, [_vp, _i32, _i32, _i32, _f32, _f32, _i32, _vp, _vp, _vp]),
    "b200ssl_opt_ema_multi_tensor": (_i32, [_vp, _i32, _vp, _i32, _f32, _f32, _vp]),
    "b200ssl_opt_ema_multi_tensor_dev": (_i32, [_vp, _i32, _vp, _i32, _f32, _f32, _vp]),
    "b200ssl_peer_control_bytes": (_sz, []),
    "b200ssl_peer_alloc": (_i32, [_sz, _vp, _vp]),
    "b200ssl_peer_open": (_i32, [_vp, _vp]),
    "b200ssl_peer_close": (_i32, [_vp]),
    "b200ssl_peer_free": (_i32, [_vp]),
    "b200ssl_peer_timeouts": (_i32, [_vp, _vp]),
    "b200ssl_bank_enqueue_peer": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200ssl_peer_all_gather": (_i32, [_vp, _sz, _vp, _sz, _vp, _vp, _sz, _sz, _i32, _i32, _i32, _vp]),
    "b200ssl_peer_reduce_scatter_f32": (_i32, [_vp, _vp, _i64, _vp, _sz, _sz, _i32, _i32, _i32, _vp]),
}

_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libb200ssl.so once; raise loudly when it is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = os.environ.get("B200SSL_LIB", str(LIB_PATH))
                if not os.path.exists(path):
                    raise NativeLibraryError(
                        f"{path} not found: the SSL head / EMA kernels are CUDA-only and have no fallback. "
                        "Build them with `python -c 'import __graft_entry__ as g; g.build()'`.")
                l = C.CDLL(path)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)       # AttributeError if the symbol is missing
                    fn.restype, fn.argtypes = res, args
                if l.b200ssl_version() < 100:
                    raise NativeLibraryError("libb200ssl.so is older than this package")
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b200ssl_last_error_string().decode(errors="replace")
        raise RuntimeError(f"libb200ssl {what} failed (code {rc}): {msg}")


def dtype_enum(t: torch.Tensor) -> int:
    try:
        return _TORCH2ENUM[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} for the B200 SSL kernels") from None


def require_cuda(*tensors: torch.Tensor, what: str = "") -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{what}: expected CUDA tensors (the B200 SSL head has no CPU path), got {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"{what}: tensors on different devices ({dev} vs {t.device})")
    return dev


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# One zero-initialised workspace per device, grown on demand.  It serialises the head's
# kernels on one stream per device (the trainer's); pass your own workspace through the C ABI
# to run heads concurrently on several streams.
_workspaces: Dict[int, torch.Tensor] = {}


def workspace(device: torch.device, rows: int, classes: int, bank_rows: int = 0) -> Tuple[int, int]:
    need = int(lib().b200ssl_workspace_bytes(int(rows), int(classes), int(bank_rows)))
    key = device.index if device.index is not None else torch.cuda.current_device()
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("workspace must be allocated before CUDA-graph capture: run the step once eagerly first")
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws.data_ptr(), ws.numel()
