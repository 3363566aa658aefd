"""CPU: the C-ABI library loads, exports every symbol include/pxb.h declares, and its struct
layout matches the ctypes binding.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pyxsim_b200 import build, _lib

    build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pxb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pxb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pxb.h but not exported by libpxb.so"


def test_binding_covers_header(lib):
    from pyxsim_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_struct_layout_matches(lib):
    from pyxsim_b200 import _lib

    nst = len(_lib.ABI_STRUCTS)
    sizes = (C.c_int64 * nst)()
    n = lib.pxb_struct_sizes(sizes, nst)
    assert n == nst
    assert [int(v) for v in sizes] == [C.sizeof(t) for t in _lib.ABI_STRUCTS]
    assert lib.pxb_struct_sizes(sizes, 2) < 0
    assert b"room" in lib.pxb_last_error()


def test_no_device_is_a_loud_error(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    assert lib.pxb_device_info(None, None, None) == -4
    assert b"no CPU fallback" in lib.pxb_last_error()
    import pyxsim_b200 as px
    import numpy as np

    src = px.ArrayDataSource({("gas", "kT"): np.ones(4), ("gas", "emission_measure"): np.ones(4),
                              ("index", "x"): np.zeros(4), ("index", "y"): np.zeros(4),
                              ("index", "z"): np.zeros(4)})
    sm = px.ThermalSpectralModel(np.linspace(0.1, 1, 11), np.array([0.5, 2.0]),
                                 np.ones((2, 10)), np.ones((2, 10)))
    model = px.ThermalSourceModel(sm, 0.1, 1.0, 10, 0.3, temperature_field=("gas", "kT"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        px.make_photons("mem:x", src, 0.05, 100.0, 100.0, model)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyxsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_wabs_fast_index_table_invariants():
    """project.cu's fp32 wabs piece lookup indexes a 128-entry table by (float bits >> 19) over
    [2^-4, 2^4) keV; it is only exact if no sub-interval holds two edges and no (float) edge is a
    sub-interval boundary.  Restated here so a change of the edge list cannot silently break it."""
    edges = np.array([0.1, 0.284, 0.4, 0.532, 0.707, 0.867, 1.303, 1.84, 2.471, 3.21, 4.038, 7.111,
                      8.331]).astype(np.float32)
    base = (127 - 4) << 4
    idx = (edges.view(np.int32) >> 19) - base
    assert idx.min() >= 0 and idx.max() < 128
    assert np.unique(idx).size == edges.size                      # at most one edge per sub-interval
    lo = ((idx + base) << 19).astype(np.int32).view(np.float32)
    assert np.all(edges > lo)                                     # never on a boundary
    src = open(os.path.join(ROOT, "pyxsim_b200", "csrc", "projection.cuh")).read()
    assert "WABS_LUT_BASE = (127 - 4) << 4" in src and "WABS_LUT_N = 128" in src
    for v in ("0.284", "0.532", "0.707", "0.867", "1.303", "2.471", "4.038", "7.111", "8.331"):
        assert src.count(v) >= 2                                  # immediate chain and table builder agree


def test_alias_tables_reproduce_the_row_distribution(lib):
    """The sampler's per-row Walker alias tables (built on the host, pxb_alias_table_host): the
    implied bin probabilities (thr_j + sum_{k: other_k = j} (2^32 - thr_k)) / (2^32 nb) must equal
    the normalised row to the table's resolution 2^-32 / nb, bins of zero probability must be
    unreachable, and every alternative must be a valid bin."""
    from pyxsim_b200 import synthetic

    kT, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=4000)
    rng = np.random.RandomState(5)
    rows = [cosmic[3], metal[20], metal[35], rng.uniform(0, 1, 7), np.array([0.0, 0.0, 5.0, 0.0]),
            np.array([1.0]), np.where(rng.uniform(size=513) < 0.7, 0.0, rng.lognormal(0, 3, 513)),
            np.full(64, 0.25)]
    for row in rows:
        row = np.ascontiguousarray(row, "float64")
        nb = row.size
        thr = np.zeros(nb, dtype=np.uint32)
        oth = np.zeros(nb, dtype=np.uint32)
        rc = lib.pxb_alias_table_host(row.ctypes.data_as(C.c_void_p), nb, thr.ctypes.data_as(C.c_void_p),
                                      oth.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert oth.max() < nb
        # the kernel takes bin j when top32 < thr_j, else other_j; thr = 2^32 - 1 with other = j is "always j"
        direct = thr.astype(np.float64)
        spill = 4294967296.0 - direct
        P = direct.copy()
        np.add.at(P, oth, spill)
        P /= 4294967296.0 * nb
        want = row / row.sum()
        assert abs(P.sum() - 1.0) < 1e-12
        # every table entry is rounded to 2^-32: a bin collects one rounding per entry pointing at it
        fan_in = 1.0 + np.bincount(oth, minlength=nb).max()
        assert np.max(np.abs(P - want)) <= fan_in * 2.0 ** -32 / nb
        assert np.all(P[want == 0.0] == 0.0)
