"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Import the reference's own Python modules (``/root/reference/pyxsim/...``) in a container
that has none of their third-party dependencies, by installing minimal stand-ins for
``yt``, ``soxs``, ``unyt``, ``h5py``, ``astropy`` and ``more_itertools`` in ``sys.modules``
and pointing ``pyxsim.lib`` at the natives compiled into ``oracle/_ref``.

The stand-ins carry NO physics: arrays are plain ndarrays that remember a unit string,
"conversions" assume the caller already supplied the requested unit (plus the three
constants the reference needs at import time), ``h5py.File`` is an in-memory dict,
``Orientation`` is the published yt rule restated in ``ref_glue.orientation``.  Everything
that is actually computed -- ``ThermalSourceModel.process_data``,
``PowerLawSourceModel.process_data``, ``LineSourceModel.process_data``,
``SpectralInterpolator1D/2D``, ``PionSpectralModel.get_spectrum``,
``AbsorptionModel.absorb_photons`` and ``_project_photons`` -- is the reference's unmodified
code, executed where it lies.  It is used to generate ``tests/golden/reference_golden.npz``
(``tests/golden/make_reference_golden.py``), which pins the NumPy restatement in
``ref_glue.py`` bit for bit (same RandomState, same draw order).
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("PYXSIM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

_UNIT_FACTORS = {
    ("kpc**2", "cm**2"): 3.0856775809623245e21 ** 2,
    ("km/s", "km/s"): 1.0,
}


class UArray(np.ndarray):
    """ndarray that remembers a unit string (stand-in for unyt_array / YTArray)."""

    def __new__(cls, value, units=""):
        obj = np.asarray(value, dtype="float64").view(cls)
        obj.units = units
        return obj

    def __array_finalize__(self, obj):
        self.units = getattr(obj, "units", "")

    @property
    def d(self):
        return self.view(np.ndarray)

    v = d
    value = d

    @property
    def uq(self):
        return 1.0

    def to_value(self, units=None, equivalence=None):
        f = _UNIT_FACTORS.get((str(self.units), units), 1.0)
        out = self.view(np.ndarray) * f if f != 1.0 else self.view(np.ndarray)
        return out if out.ndim else float(out)

    def to(self, units, equivalence=None):
        return UArray(self.to_value(units), units)

    in_units = to

    def in_cgs(self):
        return self

    def convert_to_units(self, units):
        self.units = units


def UQuantity(value, units=""):
    return UArray(float(value), units)


class _FakeDataset(dict):
    """h5py dataset / attribute holder: numpy array with the few extras the reference uses."""


class FakeH5Dataset:
    def __init__(self, data):
        self._a = np.array(data)

    def __getitem__(self, k):
        if isinstance(k, tuple) and k == ():
            return self._a[()] if self._a.ndim == 0 else self._a
        return self._a[k]

    def __setitem__(self, k, v):
        self._a[k] = v

    @property
    def size(self):
        return self._a.size

    @property
    def shape(self):
        return self._a.shape

    def resize(self, shape):
        n = shape[0]
        a = np.zeros(n, dtype=self._a.dtype)
        m = min(n, self._a.shape[0])
        a[:m] = self._a[:m]
        self._a = a

    def asstr(self):
        outer = self

        class _S:
            def __getitem__(self, k):
                v = outer._a[()]
                return v.decode() if isinstance(v, bytes) else str(v)

        return _S()


class FakeH5Group(dict):
    def __init__(self):
        super().__init__()
        self.attrs = {}

    def create_group(self, name):
        g = FakeH5Group()
        self[name] = g
        return g

    def create_dataset(self, name, data=None, shape=None, dtype=None, **kw):
        if data is None and shape is not None:       # (h5py: allocate, filled by slices later)
            data = np.zeros(shape, dtype=dtype)
        ds = FakeH5Dataset(data)
        self[name] = ds
        return ds


FILES = {}   # in-memory "disk": file name -> FakeH5Group


class FakeH5File(FakeH5Group):
    def __new__(cls, name, mode="r"):
        if mode == "r":
            if name not in FILES:
                raise FileNotFoundError(name)
            return FILES[name]
        obj = super().__new__(cls)
        return obj

    def __init__(self, name, mode="r"):
        if mode != "r":
            super().__init__()
            FILES[name] = self

    def flush(self):
        pass

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


FakeH5Group.flush = lambda self: None
FakeH5Group.close = lambda self: None


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Install the stand-ins and a synthetic ``pyxsim`` package rooted at the reference tree."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REF_ROOT, "pyxsim")):
        raise ImportError(f"reference tree {REF_ROOT} not found")
    from . import build_ref, ref_glue

    if not build_ref.build(verbose=False):
        raise ImportError("oracle/_ref could not be built")

    # --- unyt / astropy / more_itertools ---------------------------------------------
    class UnitConversionError(Exception):
        pass

    clight = UArray(299792.458, "km/s")
    _mod("unyt", clight=clight, unyt_array=UArray, unyt_quantity=UArray)
    _mod("unyt.array", unyt_array=UArray, unyt_quantity=UArray)
    _mod("unyt.exceptions", UnitConversionError=UnitConversionError)
    _mod("astropy")
    _mod("astropy.units", Quantity=type("Quantity", (), {}))

    def chunked(it, n):
        it = list(it)
        for s in range(0, len(it), n):
            yield it[s:s + n]

    def always_iterable(x):
        return x if isinstance(x, (list, tuple)) else (x,)

    _mod("more_itertools", chunked=chunked, always_iterable=always_iterable)

    # --- soxs ----------------------------------------------------------------------------
    def parse_prng(prng):
        if isinstance(prng, np.random.RandomState):
            return prng
        if isinstance(prng, int):
            return np.random.RandomState(prng)
        return np.random

    def _absent(*a, **k):
        raise RuntimeError("soxs is not installed: this stand-in only provides names")

    _mod("soxs", __version__="stub")
    _mod("soxs.constants", erg_per_keV=ref_glue.erg_per_keV, K_per_keV=ref_glue.K_per_keV,
         atomic_weights=np.ones(31), elem_names=[""] * 31, metal_elem=list(range(3, 31)),
         abund_tables={"angr": np.ones(31)})
    _mod("soxs.utils", parse_prng=parse_prng, regrid_spectrum=_absent)
    _mod("soxs.spectra", CIEGenerator=_absent, CloudyCIEGenerator=_absent,
         CloudyPionGenerator=_absent, MekalGenerator=_absent, get_tbabs_absorb=_absent,
         get_wabs_absorb=_absent, CountRateSpectrum=object, Spectrum=object)

    # --- yt --------------------------------------------------------------------------------
    class _Comm:
        size, rank = 1, 0

        def mpi_allreduce(self, x, op="sum"):
            return x

        def barrier(self):
            pass

    class _CommSystem:
        communicators = [_Comm()]

    class Orientation:
        def __init__(self, normal_vector, north_vector=None, steady_north=False):
            uv = ref_glue.orientation(normal_vector, north_vector)
            self.unit_vectors = uv
            self.north_vector = uv[1]
            self.normal_vector = uv[2]

    class YTFieldNotFound(Exception):
        pass

    _mod("yt", __version__="stub")
    _mod("yt.units")
    _mod("yt.units.yt_array", YTArray=UArray, YTQuantity=UArray)
    _mod("yt.utilities")
    _mod("yt.utilities.cosmology", Cosmology=object)
    _mod("yt.utilities.parallel_tools")
    _mod("yt.utilities.parallel_tools.parallel_analysis_interface",
         communication_system=_CommSystem(), parallel_objects=lambda it, *a, **k: it,
         parallel_capable=False)
    _mod("yt.utilities.orientation", Orientation=Orientation)
    _mod("yt.utilities.exceptions", YTFieldNotFound=YTFieldNotFound)
    _mod("yt.data_objects")
    _mod("yt.data_objects.static_output", Dataset=type("Dataset", (), {}))
    _mod("yt.funcs", ensure_numpy_array=lambda x: np.asarray(x, dtype="float64"))

    # --- h5py ------------------------------------------------------------------------------
    _mod("h5py", File=FakeH5File)

    # --- pyxsim package skeleton: submodules load from the reference tree, natives from _ref --
    pkg = types.ModuleType("pyxsim")
    pkg.__path__ = [os.path.join(REF_ROOT, "pyxsim")]
    pkg.__version__ = "reference"
    sys.modules["pyxsim"] = pkg
    lib = types.ModuleType("pyxsim.lib")
    lib.__path__ = [os.path.join(HERE, "_ref")]
    sys.modules["pyxsim.lib"] = lib
    _installed = True


def modules():
    """(photon_list, spectral_models, thermal base, power_law_sources, line_sources) of the reference."""
    install()
    import importlib

    names = ["pyxsim.utils", "pyxsim.spectral_models", "pyxsim.source_models.sources",
             "pyxsim.source_models.thermal_sources.base", "pyxsim.source_models.power_law_sources",
             "pyxsim.source_models.line_sources", "pyxsim.photon_list"]
    return {n.split(".")[-1] if n != "pyxsim.source_models.thermal_sources.base" else "thermal":
            importlib.import_module(n) for n in names}


class FakeChunk(dict):
    """dict-backed chunk: ``chunk[field]`` -> UArray; ``.ds.field_info`` for check_num_cells."""

    def __init__(self, fields, n):
        super().__init__(fields)
        self[("gas", "ones")] = UArray(np.ones(n), "")
        self.ds = types.SimpleNamespace(field_info={("gas", "ones"): None})


class NoBar:
    def update(self, *a, **k):
        pass

    def close(self):
        pass
