"""Device plumbing: torch is the allocator / stream provider, nothing more."""
import ctypes as C

import numpy as np
import torch


_cuda_ok = False


def require_cuda():
    global _cuda_ok
    if _cuda_ok:                      # (a positive answer cannot change within a process)
        return
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pyxsim_b200 needs a CUDA device (sm_100a): there is no CPU fallback for the hot path"
        )
    _cuda_ok = True


def current_device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(a, dtype=torch.float64, pin=True):
    """numpy / torch (host or device) -> contiguous device tensor of ``dtype``.

    Host arrays go through a pinned staging copy so the H2D transfer is asynchronous.
    """
    dev = current_device()
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != dtype:
            t = t.to(dtype)
        if t.device != dev:
            if t.device.type == "cpu" and pin and not t.is_pinned() and t.numel() > 0:
                t = t.contiguous().pin_memory()
            t = t.to(dev, non_blocking=True)
        return t.contiguous().reshape(-1)
    arr = np.ascontiguousarray(np.ravel(np.asarray(a)), dtype=_np_dtype(dtype))
    t = torch.from_numpy(arr)
    if pin and t.numel() > 0:
        t = t.pin_memory()
    return t.to(dev, non_blocking=True)


def _np_dtype(dt):
    return {torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32,
            torch.uint8: np.uint8}[dt]


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def empty(n, dtype=torch.float64):
    return torch.empty(int(n), dtype=dtype, device=current_device())
