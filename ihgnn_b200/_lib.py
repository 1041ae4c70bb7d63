"""ctypes binding of libihgnn_b200.so (the C ABI declared in include/ihgnn_b200.h).

There is no CPU path and no fallback: `lib()` raises if the shared library is missing or does
not load, and every call checks the integer status and raises RuntimeError with
`ihg_last_error()`.  Tensors are handed over as raw device pointers + extents + the current
torch CUDA stream, so the calls are asynchronous and CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libihgnn_b200.so")
ABI_VERSION = 4

_lib: Optional[ctypes.CDLL] = None


class IhgCsr(Structure):
    """Mirror of `struct ihg_csr` (include/ihgnn_b200.h)."""
    _fields_ = [
        ("n_rows", c_int64), ("nnz", c_int64),
        ("rowptr", c_void_p), ("col", c_void_p),
        ("chunk_len", c_int32),
        ("n_seg", c_int64), ("n_split", c_int64), ("n_part", c_int64),
        ("seg", c_void_p), ("split_row", c_void_p), ("split_ptr", c_void_p),
    ]


class IhgAdamTensor(Structure):
    """Mirror of `struct ihg_adam_tensor` (include/ihgnn_b200.h)."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("step", c_void_p), ("numel", c_int64)]


P = c_void_p
I32, I64, F32 = c_int32, c_int64, c_float

# name -> (restype, argtypes); must list every symbol include/ihgnn_b200.h declares
SIGNATURES = {
    "ihg_abi_version": (c_int32, []),
    "ihg_last_error": (c_char_p, []),
    "ihg_launch_count": (I64, []),
    "ihg_graph_workspace_bytes": (I64, [I64, I64]),
    "ihg_graph_build": (c_int32, [P, P, P, I64, I64, I64, I64, P, P, P, P, P, P, P, P, I64, P]),
    "ihg_csr_from_keys_workspace_bytes": (I64, [I64, I64]),
    "ihg_csr_from_keys": (c_int32, [P, P, I64, I64, P, P, P, P, P, I64, P]),
    "ihg_segment_plan_workspace_bytes": (I64, [I64]),
    "ihg_segment_plan_build": (c_int32, [P, I64, I32, P, P, P, P, P, I64, P]),
    "ihg_segment_reduce": (c_int32, [POINTER(IhgCsr), P, I64, I32, I64, I64, P, P, I64, P, P, P, P, I64, I32, I32, P]),
    "ihg_two_hop_index_build": (c_int32, [POINTER(IhgCsr), P, I64, I64, P, P, P]),
    "ihg_two_hop_reduce": (c_int32, [POINTER(IhgCsr), P, P, I64, P, F32, F32, F32, P, P, P, I64, I32, P]),
    "ihg_segment_reduce_routed": (c_int32, [POINTER(IhgCsr), P, I64, I32, I64, I64, P, P, P, P, I32, I64, I32, P]),
    "ihg_two_hop_reduce_routed": (c_int32, [POINTER(IhgCsr), P, P, I64, P, F32, F32, F32, P, P, P, I32, I64, I32, P]),
    "ihg_edge_gather_sum": (c_int32, [P, I64, P, F32, P, P, I64, P, I64, I32, P]),
    "ihg_edge_interact_fwd_workspace_bytes": (I64, [I32, I32]),
    "ihg_edge_interact_fwd": (c_int32, [P, I64, P, I64, P, I64, I32, P, I64, P, I64, I32, P, I64, P]),
    "ihg_feature_interact_fwd": (c_int32, [P, I64, P, I64, P, I32, P, I64, P, I64, I32, P, I64, P]),
    "ihg_feature_interact_supported": (c_int32, [I32]),
    "ihg_edge_interact_bwd_workspace_bytes": (I64, [I32, I32]),
    "ihg_edge_interact_bwd": (c_int32, [P, I64, P, I64, P, I64, I32, P, I64, P, P, I32, P, I64, P]),
    "ihg_node_linear": (c_int32, [P, I64, P, I32, I32, I32, I32, P, P, I64, I64, I64, I64, P, I64, P]),
    "ihg_node_linear_wgrad_workspace_bytes": (I64, [I32, I32, I32]),
    "ihg_node_linear_wgrad": (c_int32, [P, I64, P, I64, I64, I64, I64, I32, I32, I32, P, P, P, I64, P]),
    "ihg_copy_rows": (c_int32, [P, I64, P, I64, I64, I32, P]),
    "ihg_gather_rows": (c_int32, [P, I64, P, I64, I64, P, I64, I32, P]),
    "ihg_scatter_add_rows": (c_int32, [P, I64, P, I64, I64, P, I64, I32, P]),
    "ihg_hem_score_fwd": (c_int32, [P, I64, P, I64, P, I64, P, P, F32, I64, I32, P, I32, P, P]),
    "ihg_hem_score_bwd": (c_int32, [P, P, I64, P, I64, P, I64, P, F32, I64, I32, P, P, P, P, I64, P, I64, P, P]),
    "ihg_hem_score_bwd_workspace_bytes": (I64, [I64]),
    "ihg_rank_topk": (c_int32, [P, I64, P, P, I64, I64, P, I64, I64, I64, P, F32, I32, I32, I32, P, P, P]),
    "ihg_sample_batch": (c_int32, [P, P, P, P, I64, I32, I64, ctypes.c_uint64, ctypes.c_uint64, P, P, P, P, P, P, P, P,
                                   P, P, P, I32, P]),
    "ihg_halo_copy": (c_int32, [P, P, P, I32, P, I64, I64, I32, P]),
    "ihg_halo_reduce": (c_int32, [P, I64, P, P, P, I32, I64, P, P, I64, I64, I32, P]),
    "ihg_adam_step": (c_int32, [POINTER(IhgAdamTensor), I32, P, F32, F32, F32, F32, P]),
}

# Optional per-call profiler (bench.py installs one): an object with
# `add(name, tag, algo_bytes, start_event, end_event)`.  None = no instrumentation.
profiler = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; fail loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m ihgnn_b200.build` "
            "(ihgnn_b200 has no CPU or PyTorch fallback)")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = handle.ihg_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libihgnn_b200.so ABI {got} != expected {ABI_VERSION}: rebuild it")
    _lib = handle
    return handle


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().ihg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


class _CurrentStream:
    """Placeholder for "the current torch stream of the device the tensor arguments live on";
    `call` resolves it once that device is known."""

    def __repr__(self):
        return "<current CUDA stream>"


CURRENT_STREAM = _CurrentStream()
_tls = threading.local()


def _take_arg_device() -> Optional[int]:
    dev = getattr(_tls, "device", None)
    _tls.device = None
    return dev


def call(name: str, *args, tag: Optional[str] = None, algo_bytes: int = 0) -> None:
    """Invoke a status-returning entry point and raise on error.  `tag` / `algo_bytes`
    (algorithmic bytes the call must move) only feed the optional profiler.

    Device guard: the launch goes to the device the tensor arguments live on (recorded by `ptr`
    while the argument list was evaluated) and to THAT device's current torch stream, whatever
    torch's current device is -- the reference builds `torch.device('cuda:N')` and never calls
    `set_device` (Main.py:61-64), so the current device is usually 0."""
    fn = getattr(lib(), name)
    dev = _take_arg_device()
    cur = torch.cuda.current_device()
    if dev is None:
        dev = cur
    if any(a is CURRENT_STREAM for a in args):
        st = torch.cuda.current_stream(dev).cuda_stream
        args = tuple(st if a is CURRENT_STREAM else a for a in args)
    if dev != cur:
        with torch.cuda.device(dev):
            _call_on_current_device(fn, name, args, tag, algo_bytes)
    else:
        _call_on_current_device(fn, name, args, tag, algo_bytes)


def _call_on_current_device(fn, name, args, tag, algo_bytes) -> None:
    if profiler is None:
        check(fn(*args), name)
        return
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    start.record()
    status = fn(*args)
    end.record()
    check(status, name)
    profiler.add(name, tag or name, int(algo_bytes), start, end)


def launch_count() -> int:
    """Kernels launched by the library in this process so far."""
    return int(lib().ihg_launch_count())


def stream_ptr() -> _CurrentStream:
    """The stream argument of an entry point: the current torch CUDA stream of the device the
    call's tensors live on (resolved inside `call`)."""
    return CURRENT_STREAM


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL).  Records the tensor's device for the device guard
    of the `call` whose argument list is being evaluated; tensors of two devices in one call raise."""
    if t is None:
        return None
    if t.is_cuda:
        idx = t.device.index
        seen = getattr(_tls, "device", None)
        if seen is None:
            _tls.device = idx
        elif seen != idx:
            _tls.device = None
            raise RuntimeError(f"ihgnn_b200: tensors of one call live on different devices (cuda:{seen} and cuda:{idx})")
    return t.data_ptr()


def require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "ihgnn_b200 runs on CUDA (sm_100a) only: got a tensor on "
                f"{t.device}; there is no CPU fallback")


def rows_f32(t: torch.Tensor) -> torch.Tensor:
    """Return `t` as a float32 2-D tensor whose rows are contiguous (stride(1) == 1) and
    16-byte aligned with a leading dimension that is a multiple of 4; copy only if needed."""
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected float32, got {t.dtype}")
    if t.dim() != 2:
        raise RuntimeError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if (t.shape[0] > 1 and t.stride(0) % 4 != 0) or (t.shape[1] > 1 and t.stride(1) != 1) \
            or t.data_ptr() % 16 != 0 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def ld(t: torch.Tensor) -> int:
    """Leading dimension (row stride in elements) of a 2-D row-major tensor."""
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
