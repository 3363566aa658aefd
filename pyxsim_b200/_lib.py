"""ctypes binding of ``libpxb.so`` (the C ABI declared in ``include/pxb.h``).

The product path has NO CPU fallback: if the CUDA library is missing, importing this
module raises, and every wrapper raises :class:`PxbError` when the library reports a
failure (status < 0, message from ``pxb_last_error``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (PXB_LIBRARY: another build of the same library, e.g. libpxb_dbg.so with the kernels' index
# assertions compiled in -- `python -m pyxsim_b200.build --debug-checks`)
LIB_PATH = os.environ.get("PXB_LIBRARY") or os.path.join(HERE, "libpxb.so")

PXB_MAX_VAR = 64

c_double_p = C.POINTER(C.c_double)
c_void_p = C.c_void_p


class PxbError(RuntimeError):
    pass


class TableDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nbins", C.c_int32), ("nvar", C.c_int32),
        ("xnodes", c_void_p), ("ynodes", c_void_p), ("cosmic", c_void_p),
        ("metal", c_void_p), ("var", c_void_p),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("primary", TableDesc), ("fallback", TableDesc),
        ("log_t", C.c_int32), ("fallback_log_t", C.c_int32),
        ("density_dependent", C.c_int32), ("binscale_log", C.c_int32),
        ("method", C.c_int32), ("reserved", C.c_int32),
        ("K_per_keV", C.c_double),
        ("kT_lo", C.c_double), ("kT_hi", C.c_double), ("nH_lo", C.c_double), ("nH_hi", C.c_double),
        ("bin_edges", c_void_p), ("emid", c_void_p), ("ebins", c_void_p),
    ]


class ThermalCells(C.Structure):
    _fields_ = [
        ("ncell", C.c_int64), ("cell_id0", C.c_int64),
        ("kT", c_void_p), ("em", c_void_p), ("nH", c_void_p), ("zmet", c_void_p),
        ("xh", c_void_p), ("density", c_void_p), ("entropy", c_void_p), ("r2", c_void_p),
        ("elem", c_void_p * PXB_MAX_VAR),
        ("elem_const", C.c_double * PXB_MAX_VAR),
        ("elem_scale", C.c_double * PXB_MAX_VAR),
        ("elem_by_xh", C.c_int32 * PXB_MAX_VAR),
        ("zmet_const", C.c_double), ("zmet_scale", C.c_double),
        ("zmet_by_xh", C.c_int32), ("nvar", C.c_int32),
        ("xh_const", C.c_double), ("spectral_norm", C.c_double),
        ("kT_min", C.c_double), ("kT_max", C.c_double),
        ("max_density", C.c_double), ("min_entropy", C.c_double),
        ("seed", C.c_uint64),
        ("id_block", C.c_int64), ("id_stride", C.c_int64),
        ("generic_kernels", C.c_int32), ("split_factor", C.c_int32),
    ]


class Geometry(C.Structure):
    _fields_ = [
        ("c", C.c_double * 3), ("le", C.c_double * 3), ("re", C.c_double * 3),
        ("dw", C.c_double * 3), ("bulk_v", C.c_double * 3),
        ("periodic", C.c_int32 * 3), ("reserved", C.c_int32),
    ]


class PowerlawCells(C.Structure):
    _fields_ = [
        ("ncell", C.c_int64), ("cell_id0", C.c_int64),
        ("lum", c_void_p), ("alpha", c_void_p), ("r2", c_void_p),
        ("alpha_const", C.c_double), ("e0", C.c_double), ("emin", C.c_double), ("emax", C.c_double),
        ("spectral_norm", C.c_double), ("scale_factor", C.c_double), ("seed", C.c_uint64),
        ("id_block", C.c_int64), ("id_stride", C.c_int64),
    ]


class LineCells(C.Structure):
    _fields_ = [
        ("ncell", C.c_int64), ("cell_id0", C.c_int64),
        ("rate", c_void_p), ("sigma", c_void_p), ("r2", c_void_p),
        ("sigma_const", C.c_double), ("e0", C.c_double),
        ("spectral_norm", C.c_double), ("scale_factor", C.c_double), ("seed", C.c_uint64),
        ("id_block", C.c_int64), ("id_stride", C.c_int64),
    ]


class PhotonList(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64), ("n_photons", C.c_int64), ("n_pairs", C.c_int64),
        ("nph", c_void_p), ("poff", c_void_p), ("qoff", c_void_p),
        ("x", c_void_p), ("y", c_void_p), ("z", c_void_p),
        ("vx", c_void_p), ("vy", c_void_p), ("vz", c_void_p),
        ("dx", c_void_p), ("energy", c_void_p), ("cell_id", c_void_p),
    ]


class ProjectParams(C.Structure):
    _fields_ = [
        ("observer", C.c_int32), ("data_type", C.c_int32), ("kernel", C.c_int32),
        ("no_shifting", C.c_int32), ("vn_axis", C.c_int32), ("sky_mode", C.c_int32),
        ("save_los", C.c_int32), ("absorb_kind", C.c_int32), ("has_sigma_pos", C.c_int32),
        ("events_f32", C.c_int32),
        ("sigma_pos", C.c_double),
        ("x_hat", C.c_double * 3), ("y_hat", C.c_double * 3), ("z_hat", C.c_double * 3),
        ("D_A_kpc", C.c_double), ("sky_center", C.c_double * 2), ("nH", C.c_double),
        ("abs_energy", c_void_p), ("abs_sigma", c_void_p), ("abs_n", C.c_int64),
        ("abs_grid", C.c_int32), ("generic_kernels", C.c_int32),
        ("abs_x0", C.c_double), ("abs_inv_dx", C.c_double),
        ("nH_cell", c_void_p), ("oneplusz", C.c_double),
        ("seed", C.c_uint64), ("cell_id0", C.c_int64),
    ]


ABS_NONE, ABS_TABLE, ABS_WABS = 0, 1, 2
SKY_TAN, SKY_FLAT, SKY_PHYS = 0, 1, 2

# name -> (restype, argtypes); every symbol include/pxb.h declares
_P = C.POINTER
class BandFluxDesc(C.Structure):
    """pxb_band_flux"""
    _fields_ = [("cflux", c_void_p), ("mflux", c_void_p), ("vflux", c_void_p),
                ("fb_cflux", c_void_p), ("fb_mflux", c_void_p), ("fb_vflux", c_void_p)]


SIGNATURES = {
    "pxb_last_error": (C.c_char_p, []),
    "pxb_version": (C.c_int, []),
    "pxb_device_info": (C.c_int, [_P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "pxb_struct_sizes": (C.c_int, [_P(C.c_int64), C.c_int]),
    "pxb_launch_count": (C.c_int64, []),
    "pxb_launch_count_reset": (None, []),
    "pxb_model_create": (C.c_int, [_P(ModelDesc), _P(c_void_p)]),
    "pxb_model_destroy": (C.c_int, [c_void_p]),
    "pxb_model_nonnegative": (C.c_int, [c_void_p]),
    "pxb_model_fusion_pays": (C.c_int, [c_void_p]),
    "pxb_thermal_counts": (C.c_int, [c_void_p, _P(ThermalCells), c_void_p, c_void_p, c_void_p]),
    "pxb_thermal_spectra": (C.c_int, [c_void_p, _P(ThermalCells), c_void_p, c_void_p]),
    "pxb_interp_spec": (C.c_int, [c_void_p, C.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "pxb_thermal_sample_workspace_bytes": (C.c_int, [c_void_p, C.c_int64, C.c_int64, _P(C.c_int64)]),
    "pxb_thermal_sample": (C.c_int, [c_void_p, _P(ThermalCells), C.c_int64, c_void_p, c_void_p,
                                     c_void_p, c_void_p, C.c_int64, C.c_int64, c_void_p, c_void_p,
                                     C.c_int64, c_void_p]),
    "pxb_alias_table_host": (C.c_int, [c_void_p, C.c_int, c_void_p, c_void_p]),
    "pxb_make_events_workspace_bytes": (C.c_int, [c_void_p, C.c_int64, C.c_int64, _P(C.c_int64)]),
    "pxb_make_events": (C.c_int, [c_void_p, _P(ThermalCells), c_void_p, _P(PhotonList),
                                  _P(ProjectParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, C.c_int64, c_void_p]),
    "pxb_last_pipeline_launch": (C.c_int, [_P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "pxb_absorb_decide_words": (C.c_int, [_P(ProjectParams), C.c_int64, c_void_p, c_void_p,
                                          C.c_uint64, C.c_uint64, c_void_p, c_void_p]),
    "pxb_thermal_spectrum_workspace_bytes": (C.c_int, [c_void_p, C.c_int64, _P(C.c_int64)]),
    "pxb_thermal_spectrum": (C.c_int, [c_void_p, _P(ThermalCells), c_void_p, C.c_int64, c_void_p,
                                       c_void_p, c_void_p, C.c_int64, c_void_p]),
    "pxb_thermal_band_workspace_bytes": (C.c_int, [C.c_int64, _P(C.c_int64)]),
    "pxb_thermal_band": (C.c_int, [c_void_p, _P(ThermalCells), C.c_int, C.c_double, C.c_double,
                                   c_void_p, c_void_p, c_void_p, C.c_int64, c_void_p]),
    "pxb_powerlaw_spectrum": (C.c_int, [C.c_int64, c_void_p, c_void_p, c_void_p, C.c_int64, c_void_p,
                                        c_void_p, c_void_p]),
    "pxb_line_spectrum": (C.c_int, [C.c_int64, C.c_double, c_void_p, c_void_p, c_void_p, C.c_int64,
                                    c_void_p, C.c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pxb_shift_factor": (C.c_int, [C.c_int64, c_void_p, c_void_p, c_void_p, _P(C.c_double), c_void_p,
                                   c_void_p]),
    "pxb_powerlaw_norm": (C.c_int, [_P(PowerlawCells), c_void_p, c_void_p, c_void_p]),
    "pxb_scan_workspace_bytes": (C.c_int, [C.c_int64, _P(C.c_int64)]),
    "pxb_scan_counts": (C.c_int, [c_void_p, C.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, C.c_int64, c_void_p]),
    "pxb_compact_cells": (C.c_int, [C.c_int64] + [c_void_p] * 11 + [_P(Geometry)] + [c_void_p] * 12),
    "pxb_radius2": (C.c_int, [C.c_int64, c_void_p, c_void_p, c_void_p, _P(Geometry), c_void_p,
                              c_void_p]),
    "pxb_powerlaw_counts": (C.c_int, [_P(PowerlawCells), c_void_p, c_void_p, c_void_p]),
    "pxb_powerlaw_sample": (C.c_int, [_P(PowerlawCells), C.c_int64, c_void_p, c_void_p, c_void_p,
                                      C.c_int64, c_void_p, c_void_p]),
    "pxb_line_counts": (C.c_int, [_P(LineCells), c_void_p, c_void_p, c_void_p]),
    "pxb_line_sample": (C.c_int, [_P(LineCells), C.c_int64, c_void_p, c_void_p, c_void_p,
                                  C.c_int64, c_void_p, c_void_p]),
    "pxb_rng_poisson": (C.c_int, [C.c_double, C.c_uint64, C.c_int64, c_void_p, c_void_p]),
    "pxb_rng_binomial": (C.c_int, [C.c_int64, C.c_double, C.c_uint64, C.c_int64, c_void_p, c_void_p]),
    "pxb_project_workspace_bytes": (C.c_int, [C.c_int64, C.c_int64, _P(C.c_int64)]),
    "pxb_project": (C.c_int, [_P(PhotonList), _P(ProjectParams), c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, C.c_int64, c_void_p]),
    "pxb_doppler_shift": (C.c_int, [C.c_int64, c_void_p, c_void_p, c_void_p, C.c_int64, c_void_p,
                                    c_void_p]),
    "pxb_pixel_to_cel": (C.c_int, [C.c_int64, c_void_p, c_void_p, C.c_double, C.c_double,
                                   c_void_p]),
    "pxb_absorb_transmission": (C.c_int, [_P(ProjectParams), C.c_int64, c_void_p, c_void_p,
                                          c_void_p]),
    "pxb_thermal_flux": (C.c_int, [c_void_p, _P(ThermalCells), _P(BandFluxDesc), c_void_p, c_void_p,
                                   c_void_p]),
    "pxb_cel_to_pixel": (C.c_int, [C.c_int64, c_void_p, c_void_p, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double, c_void_p, c_void_p, c_void_p]),
    "pxb_event_image": (C.c_int, [C.c_int64, c_void_p, c_void_p, c_void_p, C.c_double, C.c_double,
                                  C.c_double, C.c_double, C.c_int, C.c_double, c_void_p, c_void_p]),
    "pxb_event_spectrum": (C.c_int, [C.c_int64, c_void_p, C.c_int, c_void_p, c_void_p, c_void_p]),
    "pxb_column_lookup": (C.c_int, [C.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int64,
                                    C.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pxb_absorb_decide": (C.c_int, [_P(ProjectParams), C.c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
}

ABI_STRUCTS = [TableDesc, ModelDesc, ThermalCells, Geometry, PowerlawCells, LineCells,
               PhotonList, ProjectParams, BandFluxDesc]

_lib = None


def load():
    """Load libpxb.so (building is a separate, explicit step: ``python -m pyxsim_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library with `python -m pyxsim_b200.build` "
            "(pyxsim_b200 has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    sizes = (C.c_int64 * len(ABI_STRUCTS))()
    n = lib.pxb_struct_sizes(sizes, len(ABI_STRUCTS))
    got = [int(v) for v in sizes[:max(n, 0)]]
    want = [C.sizeof(t) for t in ABI_STRUCTS]
    if got != want:
        raise ImportError(f"libpxb.so ABI mismatch: C sizes {got} vs ctypes sizes {want}")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().pxb_last_error()
        raise PxbError(f"libpxb error {status}: {msg.decode() if msg else '?'}")


def call(name, *args):
    check(getattr(load(), name)(*args))


def last_pipeline_launch():
    """(mode, spec, staged) of the photon kernel this thread launched last (see pxb.h)."""
    m, s, g = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    load().pxb_last_pipeline_launch(C.byref(m), C.byref(s), C.byref(g))
    return m.value, s.value, g.value


def launch_count():
    """Kernels launched by libpxb in this process since the last reset (counted in the library)."""
    return int(load().pxb_launch_count())


def launch_count_reset():
    load().pxb_launch_count_reset()
