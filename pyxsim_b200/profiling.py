"""Optional per-stage device timing (CUDA events on the current stream) and launch counting.

Disabled by default; ``bench.py`` turns it on to attribute the step time to the individual
libpxb stages and to count the kernels launched inside the timed region."""
import contextlib

import torch

_enabled = False
_records = []       # (name, start_event, stop_event)


def enable(flag=True):
    global _enabled
    _enabled = flag
    reset()


def reset():
    global _records
    _records = []
    from . import _lib

    _lib.launch_count_reset()


def launches():
    """Kernels libpxb launched since ``enable``/``reset`` (counted inside the library at every
    launch site, ``pxb_launch_count``)."""
    from . import _lib

    return _lib.launch_count()


@contextlib.contextmanager
def stage(name):
    if not _enabled:
        yield
        return
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    try:
        yield
    finally:
        b.record()
        _records.append((name, a, b))


def summary():
    """{stage: (calls, total_ms)} -- call after torch.cuda.synchronize()."""
    out = {}
    for name, a, b in _records:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + a.elapsed_time(b))
    return out
