"""GPU: a yt-like data object (unit-tagged arrays, io chunks, ds attributes -- the stand-ins of
oracle/ref_import.py; yt itself is not installed) driven through the yt branches of
``make_photons`` / ``process_data`` must give exactly what the array-backed front end gives.
Reference: pyxsim/photon_list.py:232-343,401-456; thermal_sources/base.py:322-354."""
import numpy as np
import pytest

import pyxsim_b200 as px
from pyxsim_b200 import synthetic

from oracle.ref_import import UArray, FakeChunk
from test_cpu_seams import FakeYTDataset, FakeYTObject, fake_yt  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
TF = ("gas", "temperature")

UNITS = {("gas", "temperature"): "keV", ("gas", "emission_measure"): "cm**-3", ("index", "dx"): "kpc",
         ("gas", "metallicity"): "Zsun"}


def _yt_chunks(fields, bounds):
    out = []
    for a, b in bounds:
        d = {}
        for k, v in fields.items():
            unit = UNITS.get(k, "kpc" if k[0] == "index" else "km/s")
            d[k] = UArray(v[a:b], unit)
        out.append(FakeChunk(d, b - a))
    return out


def test_yt_like_object_equals_the_array_front_end(cuda, fake_yt):
    f = {k: v.cpu().numpy() for k, v in synthetic.beta_model_fields(20, kT_range=(1.0, 9.0), v3d_sigma_kms=250.0,
                                                                     metallicity=(0.1, 0.5)).items()}
    f[TF] = f.pop(("gas", "kT"))
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=500)
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)

    def model():
        return px.ThermalSourceModel(sm, 0.1, 10.0, 500, ("gas", "metallicity"), temperature_field=TF, prng=21)

    n = f[TF].size
    bounds = [(0, 3000), (3000, 3000), (3000, n)]            # (with an empty io chunk in the middle)
    ds = FakeYTDataset()
    ds.domain_left_edge, ds.domain_right_edge = UArray([-1000.0] * 3, "kpc"), UArray([1000.0] * 3, "kpc")
    ds.domain_width, ds.domain_center = UArray([2000.0] * 3, "kpc"), UArray([0.0] * 3, "kpc")
    ds.periodicity = (False, False, False)
    ds.cosmology = px.Cosmology()
    obj = FakeYTObject(ds, _yt_chunks(f, bounds), left_edge=ds.domain_left_edge, right_edge=ds.domain_right_edge)
    units = {TF: "keV"}
    src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3, units=units)

    # the plugin call of the reference's chunk loop (photon_list.py:402-405) on a yt-like chunk
    m1, m2 = model(), model()
    m1.setup_model("photons", obj, 0.05)
    m2.setup_model("photons", src, 0.05)
    r1 = m1.process_data("photons", obj._chunks[0], 2.0e-46)
    r2 = m2.process_data("photons", px.data.ArrayChunk(src, 0, 3000), 2.0e-46)
    assert r1[0] == r2[0] > 0
    for a, b in zip(r1[1:], r2[1:]):
        assert np.array_equal(a, b)

    # the whole of make_photons through the yt branches: centre, bounds, io chunks, running cell ids
    n1 = px.make_photons("mem:yt_ph", obj, 0.05, 3.0e4, 1.0e5, model(), center="c")
    src.chunk_size = None
    n2 = px.make_photons("mem:ar_ph", src, 0.05, 3.0e4, 1.0e5, model())
    assert n1 == n2 and n1[0] > 1.0e5
    a, b = px.load_list("mem:yt_ph"), px.load_list("mem:ar_ph")
    for k in ("x", "y", "z", "vx", "vy", "vz", "dx", "num_photons", "energy", "cell_id"):
        assert np.array_equal(a.numpy(k), b.numpy(k)), k
    assert a.parameters["data_type"] == "cells" and np.array_equal(a.parameters["center"], [0.0] * 3)
    assert a.parameters["velocity_fields"].shape == (3, 2)
