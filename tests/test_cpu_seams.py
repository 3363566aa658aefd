"""CPU: the yt / soxs seams of the host layer, exercised WITHOUT yt or soxs.

yt and soxs are third-party front ends that are not installed here; the branches that talk to
them (``chunk[field].to_value(unit)``, ``ds._get_field_info``, ``ds.known_filters``,
``CIEGenerator._get_table`` ...) are driven with the unit-tagged arrays and dict-backed chunks of
``oracle/ref_import.py`` -- the same stand-ins the reference's own modules are imported with --
and with fake generator classes installed in ``sys.modules`` for the duration of a test.
Reference: pyxsim/photon_list.py:41-108,316-343,401-456; pyxsim/spectral_models.py:203-263.
"""
import sys
import types

import numpy as np
import pytest

import pyxsim_b200 as px
from pyxsim_b200 import photon_list as PL
from pyxsim_b200 import source_models as SM
from pyxsim_b200.data import is_yt_object

from oracle.ref_import import UArray, FakeChunk


class FakeFieldInfo:
    def __init__(self, name, units):
        self.name, self.units = name, units


class FakeYTDataset:
    """The handful of attributes the host layer asks a yt Dataset for."""

    def __init__(self, particle=False, filters=None):
        self.particle_types = ("PartType0",) if particle else ()
        self.known_filters = filters or {}
        self._sph_ptypes = ("PartType0",)
        self.index = sys.modules["yt.geometry.particle_geometry_handler"].ParticleIndex() if particle else object()
        self.domain_left_edge = UArray([-500.0] * 3, "kpc")
        self.domain_right_edge = UArray([500.0] * 3, "kpc")
        self.domain_width = UArray([1000.0] * 3, "kpc")
        self.domain_center = UArray([0.0] * 3, "kpc")
        self.periodicity = (True, True, True)
        self.units = {("gas", "metallicity"): "Zsun", ("gas", "O_fraction"): "dimensionless"}

    def _get_field_info(self, field):
        if isinstance(field, str):
            field = ("gas", field)
        return FakeFieldInfo(field, self.units.get(field, "dimensionless"))

    def __str__(self):
        return "FakeYT"


class FakeYTObject:
    """A yt data container: ``.ds``, ``.chunks(fields, kind)``, geometry attributes."""

    def __init__(self, ds, chunks, **attrs):
        self.ds = ds
        self._chunks = chunks
        for k, v in attrs.items():
            setattr(self, k, v)

    def chunks(self, fields, kind):
        assert fields == [] and kind == "io"
        return iter(self._chunks)

    def get_field_parameter(self, name):
        return UArray([1.0, 2.0, 3.0], "kpc")


@pytest.fixture()
def fake_yt(monkeypatch):
    ParticleIndex = type("ParticleIndex", (), {})
    for name, attrs in (("yt", {}), ("yt.geometry", {}),
                        ("yt.geometry.particle_geometry_handler", {"ParticleIndex": ParticleIndex})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
    return ParticleIndex


def test_is_yt_object_and_field_value_on_unit_arrays(fake_yt):
    ds = FakeYTDataset()
    n = 7
    chunk = FakeChunk({("gas", "temperature"): UArray(np.full(n, 3.0), "keV"),
                       ("index", "x"): UArray(np.arange(n), "kpc")}, n)
    obj = FakeYTObject(ds, [chunk])
    assert is_yt_object(obj)
    assert not is_yt_object(px.ArrayDataSource({("index", "x"): np.zeros(2), ("index", "y"): np.zeros(2),
                                                ("index", "z"): np.zeros(2)}))
    # yt chunk: the value comes back through .to_value(unit[, equivalence]) as a plain ndarray
    v = SM.field_value(chunk, ("gas", "temperature"), "keV", "thermal")
    assert type(v) is np.ndarray and np.all(v == 3.0)
    assert np.array_equal(SM.field_value(chunk, ("index", "x"), "kpc"), np.arange(n))
    assert type(SM.field_value(chunk, ("index", "x"))) is np.ndarray          # no unit asked: .d
    assert SM.field_units(chunk, ("gas", "temperature")) == "keV"


def test_determine_fields_follows_the_reference_rules(fake_yt):
    """photon_list.py:41-71: grid data, SPH particles, 'gas' on a particle index, a particle
    filter on gas, point sources."""
    grid = FakeYTObject(FakeYTDataset(), [])
    p, v, w, dt = PL.determine_fields(grid, "gas", False)
    assert p == [("index", ax) for ax in "xyz"] and v == [("gas", f"velocity_{ax}") for ax in "xyz"]
    assert w == ("index", "dx") and dt == "cells"
    assert PL.determine_fields(grid, "gas", True)[2] is None                       # point sources
    sph = FakeYTObject(FakeYTDataset(particle=True), [])
    p, v, w, dt = PL.determine_fields(sph, "PartType0", False)
    assert p[0] == ("PartType0", "particle_position_x") and v[2] == ("PartType0", "particle_velocity_z")
    assert w == ("PartType0", "smoothing_length") and dt == "particles"
    p, v, w, dt = PL.determine_fields(sph, "gas", False)                           # 'gas' -> the SPH type
    assert p[0] == ("PartType0", "particle_position_x") and dt == "particles"
    filt = types.SimpleNamespace(filtered_type="gas")
    hot = FakeYTObject(FakeYTDataset(particle=True, filters={"hot_gas": filt}), [])
    p, v, w, dt = PL.determine_fields(hot, "hot_gas", False)
    assert p == [("hot_gas", ax) for ax in "xyz"] and v[0] == ("hot_gas", "velocity_x") and w is None


def test_find_object_bounds_for_yt_containers(fake_yt):
    """photon_list.py:74-108: boxes, spheres, and the fallback to the domain."""
    ds = FakeYTDataset()
    box = FakeYTObject(ds, [], left_edge=UArray([-10.0, -20.0, -30.0], "kpc"),
                       right_edge=UArray([10.0, 20.0, 30.0], "kpc"))
    le, re = PL.find_object_bounds(box)
    assert np.array_equal(le, [-10.0, -20.0, -30.0]) and np.array_equal(re, [10.0, 20.0, 30.0])
    sphere = FakeYTObject(ds, [], radius=UArray(5.0, "kpc"), center=UArray([1.0, 2.0, 3.0], "kpc"))
    le, re = PL.find_object_bounds(sphere)
    assert np.array_equal(le, [-4.0, -3.0, -2.0]) and np.array_equal(re, [6.0, 7.0, 8.0])
    cut = FakeYTObject(ds, [], base_object=box)                       # a cut region of the box
    assert np.array_equal(PL.find_object_bounds(cut)[1], [10.0, 20.0, 30.0])
    disk = FakeYTObject(ds, [], radius=UArray(5.0, "kpc"), height=UArray(1.0, "kpc"))
    le, re = PL.find_object_bounds(disk)                                # unsupported shape: the domain
    assert np.array_equal(le, [-500.0] * 3) and np.array_equal(re, [500.0] * 3)


def test_yt_chunks_are_dealt_round_robin(fake_yt, monkeypatch):
    """photon_list.py:401: parallel_objects deals the io chunks over the ranks."""
    chunks = [FakeChunk({}, 1) for _ in range(5)]
    obj = FakeYTObject(FakeYTDataset(), chunks)
    from pyxsim_b200 import parallel

    monkeypatch.setattr(parallel, "size", lambda: 2)
    monkeypatch.setattr(parallel, "rank", lambda: 1)
    mine = list(PL._chunks(obj))
    assert mine == [chunks[1], chunks[3]]


def test_thermal_setup_model_resolves_yt_field_info(fake_yt):
    """base.py:146-229: field names through ds._get_field_info, metallicity / element units ->
    Zconvert / mconvert (mass fractions are divided by the solar value and X_H in the kernel)."""
    from pyxsim_b200 import constants as K, synthetic

    kT, ebins, c, m, v = synthetic.apec_like_tables(nT=6, nbins=20, nvar=1)
    sm = px.ThermalSpectralModel(ebins, kT, c, m, v, var_elem_names=("O",))
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 20, ("gas", "metallicity"), var_elem={"O": ("gas", "O_fraction")},
                                  temperature_field="temperature", emission_measure_field="emission_measure")
    obj = FakeYTObject(FakeYTDataset(), [])
    model.setup_model("photons", obj, 0.05)
    assert model.temperature_field == ("gas", "temperature") and model.ftype == "gas"
    assert model.emission_measure_field == ("gas", "emission_measure")
    assert model.Zconvert == 1.0 and model._z_units == "Zsun"
    n_O = K.elem_names.index("O")
    assert model.mconvert["O"] == pytest.approx(K.atomic_weights[1] / (model.atable[n_O] * K.atomic_weights[n_O]))
    assert model.density_field == ("gas", "density") and model.entropy_field == ("gas", "entropy")
    obj.ds.units[("gas", "metallicity")] = "furlongs"
    with pytest.raises(RuntimeError):
        model.setup_model("photons", obj, 0.05)


class FakeCIEGenerator:
    """soxs.spectra.CIEGenerator as TableCIEModel uses it (spectral_models.py:203-263): node
    temperatures, bin edges, and ``_get_table(rows, redshift, velocity)``."""
    calls = []

    def __init__(self, model, emin, emax, nbins, binscale="linear", var_elem=None, model_root=None,
                 model_vers=None, broadening=True, nolines=False, abund_table="angr", nei=False):
        self.Tvals = 0.1 * 2.0 ** np.arange(10)            # 0.1 .. 51.2 keV
        self.nT = self.Tvals.size
        self.ebins = np.linspace(emin, emax, nbins + 1)
        self.nbins = nbins
        self.var_elem_names = list(var_elem or [])
        self.var_ion_names = []
        self.model_vers = model_vers or "3.0.9"
        self.kw = dict(model=model, binscale=binscale, broadening=broadening, nolines=nolines,
                       abund_table=abund_table, nei=nei)

    def _get_table(self, rows, redshift, velocity):
        FakeCIEGenerator.calls.append((tuple(rows), redshift, velocity))
        rows = np.asarray(rows)
        base = (rows[:, None] + 1.0) * np.ones((1, self.nbins)) / (1.0 + redshift)
        var = np.stack([0.5 * base] * len(self.var_elem_names)) if self.var_elem_names else None
        return base, 2.0 * base, var


def test_table_cie_model_through_a_fake_soxs(monkeypatch):
    mod = types.ModuleType("soxs.spectra")
    mod.CIEGenerator = FakeCIEGenerator
    monkeypatch.setitem(sys.modules, "soxs", types.ModuleType("soxs"))
    monkeypatch.setitem(sys.modules, "soxs.spectra", mod)
    FakeCIEGenerator.calls.clear()
    sm = px.TableCIEModel("apec", 0.2, 8.0, 50, 0.5, 7.0, var_elem=["O", "Fe"], thermal_broad=False,
                          abund_table="aspl")
    # rows bracketing [kT_min, kT_max] (spectral_models.py:245-247): 0.4 <= 0.5 ... 7.0 <= 12.8
    assert sm.idx_min == 2 and sm.idx_max == 8
    assert np.array_equal(sm.Tvals, FakeCIEGenerator(None, 0.2, 8.0, 50).Tvals[2:8])
    assert sm.Tvals[0] <= 0.5 and sm.Tvals[-1] >= 7.0
    assert sm.nbins == 50 and sm.var_elem_names == ["O", "Fe"] and sm.model_vers == "3.0.9"
    assert sm.cgen.kw["broadening"] is False and sm.cgen.kw["abund_table"] == "aspl"
    sm.prepare_spectrum(0.25)
    assert FakeCIEGenerator.calls == [((2, 3, 4, 5, 6, 7), 0.25, 0.0)]
    assert sm.cosmic_spec.shape == (6, 50) and sm.var_spec.shape == (2, 6, 50)
    assert np.all(sm.metal_spec == 2.0 * sm.cosmic_spec) and sm.cosmic_spec[0, 0] == 3.0 / 1.25
    # and the source model builds on it like the reference's CIESourceModel (collisional.py:7-218)
    model = px.CIESourceModel("apec", 0.2, 8.0, 50, 0.3, var_elem={"O": 0.5, "Fe": 0.4}, kT_min=0.5, kT_max=7.0)
    assert isinstance(model.spectral_model, px.TableCIEModel) and model.num_var_elem == 2
    with pytest.raises(ImportError):
        monkeypatch.delitem(sys.modules, "soxs.spectra")
        monkeypatch.setitem(sys.modules, "soxs.spectra", None)
        px.TableCIEModel("apec", 0.2, 8.0, 50, 0.5, 7.0)


def test_tbabs_and_pion_need_their_tables_or_soxs(monkeypatch):
    monkeypatch.setitem(sys.modules, "soxs", None)
    monkeypatch.setitem(sys.modules, "soxs.spectra", None)     # (also a stand-in another test installed)
    with pytest.raises(ImportError):
        px.TBabsModel(0.02)
    with pytest.raises(ImportError):
        px.PionSourceModel(0.3, 1.4, 100, 0.3)
    # with a fake soxs the tbabs table is tabulated from get_tbabs_absorb (spectral_models.py:586-609)
    spectra = types.ModuleType("soxs.spectra")
    spectra.get_tbabs_absorb = lambda e, nH, abund_table="angr": np.exp(-2.0e-22 * e ** -2.5 * nH * 1.0e22)
    monkeypatch.setitem(sys.modules, "soxs", types.ModuleType("soxs"))
    monkeypatch.setitem(sys.modules, "soxs.spectra", spectra)
    t = px.TBabsModel(0.05, abund_table="wilm")
    assert t._name == "tbabs" and t.abund_table == "wilm" and t.emid.size == 40001
    np.testing.assert_allclose(t.sigma, 2.0 * t.emid ** -2.5, rtol=1e-7)       # 1e-22 cm^2 units (-log(exp(-tau)) round trip)
