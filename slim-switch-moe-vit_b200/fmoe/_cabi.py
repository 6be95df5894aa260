"""ctypes binding of libmoe_b200.so (C ABI: include/moe_b200.h).

There is no fallback: if the library is missing the import fails, and every compute entry point
raises when handed a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOE_B200_LIB") or os.path.join(_HERE, "libmoe_b200.so")   # env override: kernel experiments only

DTYPE_F32, DTYPE_BF16 = 0, 1
SCORE_TOPK_SOFTMAX, SCORE_FULL_SOFTMAX = 0, 1
AUX_NONE, AUX_SWITCH, AUX_GSHARD = 0, 1, 2
TOKEN_TILE, ROW_ALIGN = 64, 256
GEMM_FC1, GEMM_FC2, GEMM_DGELU, GEMM_DGRAD, GEMM_WGRAD, GEMM_WGRAD_T = range(6)

_p, _i, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/moe_b200.h one to one
SIGNATURES = {
    "moe_last_error": (ctypes.c_char_p, []),
    "moe_version": (_i, []),
    "moe_rows_cap": (_i64, [_i64, _i, _i, _i64]),
    "moe_gate_fwd": (_i, [_p, _i, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "moe_gate_fwd_workspace_bytes": (_sz, [_i, _i]),
    "moe_route_scan": (_i, [_p, _p, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _i, _p, _i, _i64, _i, _p, _p, _i64, _p]),
    "moe_ep_tables": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "moe_ep_repack": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _i, _i, _p]),
    "moe_ep_heap_alloc": (_i, [_sz, ctypes.POINTER(_p), _p]),
    "moe_ep_heap_open": (_i, [_p, ctypes.POINTER(_p)]),
    "moe_ep_heap_close": (_i, [_p]),
    "moe_ep_heap_free": (_i, [_p]),
    "moe_ep_barrier": (_i, [_p, _p, _i, _i, _p, _p]),
    "moe_ep_exchange_counts": (_i, [_p, _p, _p, _p, _i, _i, _i, _i64, _p, _p, _p, _p, _p, _i, _p, _p]),
    "moe_dispatch_fwd_peer": (_i, [_p, _i, _p, _p, _p, _i64, _i, _i, _i, _i64, _p, _i, _i, _i64, _p, _p, _p, _p]),
    "moe_combine_fwd_peer": (_i, [_p, _i, _i, _i64, _p, _p, _i64, _i, _i, _p, _i, _p]),
    "moe_combine_bwd_peer": (_i, [_p, _i, _p, _p, _i, _i, _i64, _p, _p, _p, _p, _i, _i64, _i, _i, _p, _p]),
    "moe_gate_dispatch_bwd_peer": (_i, [_p, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _i, _p]),
    "moe_dispatch_fwd": (_i, [_p, _i, _p, _p, _p, _p, _i64, _i, _i, _i, _i64, _p, _p, _p, _p]),
    "moe_expert_ffn_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _p, _p, _p, _p]),
    "moe_combine_fwd": (_i, [_p, _p, _p, _i64, _i, _i, _p, _i, _p]),
    "moe_combine_bwd": (_i, [_p, _i, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _p, _p, _p]),
    "moe_expert_ffn_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "moe_expert_ffn_bwd_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "moe_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i, _i64]),
    "moe_gate_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _p, _p]),
    "moe_dispatch_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p, _i, _p]),
    "moe_gate_wgrad_workspace_bytes": (_sz, [_i64, _i, _i]),
    "moe_gate_wgrad": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p, _p, _p]),
    "moe_cast_bf16": (_i, [_p, _p, _i64, _p]),
    "moe_cast_bf16_pair": (_i, [_p, _p, _i64, _p, _p, _i64, _p]),
    "moe_segment_colsum_workspace_bytes": (_sz, [_i64, _i]),
    "moe_segment_colsum": (_i, [_p, _p, _i64, _i, _i, _p, _p, _p]),
    "moe_gate_dispatch_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _i, _p]),
    "moe_addln_fwd": (_i, [_p, _p, _i, _p, _p, ctypes.c_float, _i64, _i, _p, _p, _i, _p, _p, _p]),
    "moe_addln_bwd_workspace_bytes": (_sz, [_i64, _i]),
    "moe_addln_bwd": (_i, [_p, _i, _p, _p, _p, _p, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _p]),
    "moe_colsum_workspace_bytes": (_sz, [_i64, _i]),
    "moe_colsum": (_i, [_p, _i, _i64, _i, _p, _p, _p]),
    "moe_wgrad_flags_bytes": (_sz, [_i, _i, _i]),
    "moe_slab_colsum_bytes": (_sz, [_i64, _i]),
    "moe_slab_colsum_final": (_i, [_p, _p, _i, _i, _p, _p]),
    "moe_grouped_gemm": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p]),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C slim-switch-moe-vit_b200/csrc`). There is no CPU / PyTorch fallback for this layer.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class MoeB200Error(RuntimeError):
    pass


# kernels launched by each entry point (for the launch count the bench reports)
KERNELS_PER_CALL = {
    "moe_ep_barrier": 1, "moe_ep_exchange_counts": 1, "moe_dispatch_fwd_peer": 1, "moe_combine_fwd_peer": 1, "moe_combine_bwd_peer": 1,
    "moe_gate_dispatch_bwd_peer": 1,
    "moe_gate_fwd": 2, "moe_route_scan": 2, "moe_ep_tables": 1, "moe_ep_repack": 1, "moe_dispatch_fwd": 1, "moe_expert_ffn_fwd": 2, "moe_combine_fwd": 1,
    "moe_combine_bwd": 1, "moe_expert_ffn_bwd": 8, "moe_gate_bwd": 1, "moe_dispatch_bwd": 1, "moe_gate_dispatch_bwd": 1,
    "moe_gate_wgrad": 2, "moe_addln_fwd": 1, "moe_addln_bwd": 2, "moe_colsum": 2, "moe_cast_bf16": 1, "moe_cast_bf16_pair": 1, "moe_segment_colsum": 2, "moe_grouped_gemm": 1, "moe_slab_colsum_final": 1,
}


class Profiler:
    """Kernel launch counter (always on) + optional per-call CUDA-event timing on the launching
    stream (off by default; bench.py switches it on for the timed region)."""

    def __init__(self):
        self.enabled = False
        self.launches = 0
        self.events = {}   # tag -> list of (start, stop) torch.cuda.Event

    def reset(self):
        self.launches = 0
        self.events = {}

    def summary_ms(self):
        """tag -> (count, mean ms); call after torch.cuda.synchronize()."""
        return {t: (len(ev), sum(a.elapsed_time(b) for a, b in ev) / len(ev)) for t, ev in self.events.items()}


PROF = Profiler()


def call(name: str, *args, tag: str | None = None) -> None:
    PROF.launches += KERNELS_PER_CALL.get(name, 0)
    if PROF.enabled:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        PROF.events.setdefault(tag or name, []).append((a, b))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise MoeB200Error(f"{name} failed: {lib.moe_last_error().decode(errors='replace')}")


def ptr(t: torch.Tensor | None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MoeB200Error("fmoe (B200) has no CPU path: expected a CUDA tensor, got device=" + str(t.device))
    if not t.is_contiguous():
        raise MoeB200Error("internal error: non-contiguous tensor handed to the C ABI")
    return t.data_ptr()


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise MoeB200Error(f"unsupported dtype {t.dtype} (fp32 or bf16 expected)")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_WGRAD_FLAGS: dict = {}
_WGRAD_FLAGS_RETIRED: list = []


def wgrad_flags(E: int, M: int, N: int, dev) -> torch.Tensor:
    """Zero-filled split-K flag workspace of the weight-gradient GEMMs (`aux` of MOE_GEMM_WGRAD / _T), one per device
    and stream; every launch leaves it zero, so it is allocated and cleared once."""
    nbytes = int(lib.moe_wgrad_flags_bytes(E, M, N))
    key = (torch.device(dev).index, stream_ptr())
    t = _WGRAD_FLAGS.get(key)
    if t is None or t.numel() * 4 < nbytes:
        new = torch.zeros(nbytes // 4, dtype=torch.int32, device=dev)
        if torch.cuda.is_current_stream_capturing():
            return new   # lives in the graph's pool (the fill is replayed with the graph): not cached
        if t is not None:
            _WGRAD_FLAGS_RETIRED.append(t)   # a captured graph may still point at it: never freed
        _WGRAD_FLAGS[key] = t = new
    return t


def rows_cap(T: int, k: int, E: int, capacity: int) -> int:
    return int(lib.moe_rows_cap(T, k, E, capacity))
