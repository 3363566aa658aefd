"""Optional per-stage device timing (CUDA events on the current stream) and launch counting.

Disabled by default; ``bench.py`` turns it on to attribute the step time to the individual
libpxb stages and to count the kernels launched inside the timed region."""
import contextlib

import torch

_enabled = False
_records = []       # (name, start_event, stop_event)
_launches = 0

# kernels launched by each C-ABI entry point (see pyxsim_b200/csrc/*.cu host functions)
KERNELS_PER_CALL = {
    "pxb_thermal_counts": 2, "pxb_thermal_sample": 3, "pxb_thermal_spectra": 1,
    "pxb_interp_spec": 1, "pxb_scan_counts": 3, "pxb_compact_cells": 1, "pxb_radius2": 1,
    "pxb_powerlaw_counts": 1, "pxb_powerlaw_sample": 1, "pxb_line_counts": 1,
    "pxb_line_sample": 1, "pxb_project": 8, "pxb_doppler_shift": 1, "pxb_pixel_to_cel": 1,
    "pxb_absorb_transmission": 1,
    "pxb_absorb_decide": 1,
    "pxb_column_lookup": 1,
    "pxb_cel_to_pixel": 1, "pxb_event_image": 1, "pxb_event_spectrum": 1,
    "pxb_thermal_flux": 1,
}


def enable(flag=True):
    global _enabled
    _enabled = flag
    reset()


def reset():
    global _records, _launches
    _records = []
    _launches = 0


def count_call(name):
    global _launches
    if _enabled:
        _launches += KERNELS_PER_CALL.get(name, 0)


def launches():
    return _launches


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
