"""CPU: pin the oracle.  The native routines ARE the reference (compiled from its Cython);
here they are checked against independent known answers and the committed golden vectors."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_pixel_to_cel_known_answer(oracle):
    """pyxsim/tests/test_pixel_to_cel.py, with the analytic TAN inverse in place of astropy."""
    prng = np.random.RandomState(24)
    n_evt = 100000
    rr = 100.0 * prng.uniform(size=n_evt)
    theta = 2.0 * np.pi * prng.uniform(size=n_evt)
    xx = rr * np.cos(theta) / 1.0e5
    yy = rr * np.sin(theta) / 1.0e5
    x1, y1 = xx.copy(), yy.copy()
    oracle.ref()["sky_functions"].pixel_to_cel(x1, y1, np.array([30.0, 45.0]))
    ra, dec = oracle.tan_known_answer(xx, yy, 30.0, 45.0)
    np.testing.assert_allclose(x1, ra, rtol=1e-7)
    np.testing.assert_allclose(y1, dec, rtol=1e-7)
    np.testing.assert_allclose(x1, ra, rtol=1e-12)


def test_interp1d_against_numpy(oracle):
    rng = np.random.RandomState(0)
    nodes = np.sort(rng.uniform(0.1, 10, 12))
    c, m = rng.uniform(size=(12, 50)), rng.uniform(size=(12, 50))
    v = rng.uniform(size=(3, 12, 50))
    si = oracle.SpectralInterpolator1D(nodes, c, m, v)
    x = rng.uniform(0.05, 12, 40)
    oc, om, ov = si(x)
    i = np.clip(np.searchsorted(nodes, x, side="right") - 1, 0, 10)
    w = (x - nodes[i]) / (nodes[i + 1] - nodes[i])
    np.testing.assert_allclose(oc, c[i] * (1 - w)[:, None] + c[i + 1] * w[:, None], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ov[2], v[2][i] * (1 - w)[:, None] + v[2][i + 1] * w[:, None], rtol=1e-12, atol=1e-14)


def test_interp2d_against_scipy(oracle):
    from scipy.interpolate import RegularGridInterpolator

    rng = np.random.RandomState(1)
    tn, dn = np.linspace(4, 7, 9), np.linspace(-7, -3, 6)
    tab = rng.uniform(size=(6, 9, 20))       # [y, x, e]
    flat = tab.reshape(54, 20)
    si = oracle.SpectralInterpolator2D(tn, dn, flat, flat.copy(), None)
    x, y = rng.uniform(4, 7, 30), rng.uniform(-7, -3, 30)
    oc, _, _ = si(x, y)
    rgi = RegularGridInterpolator((dn, tn), tab)
    np.testing.assert_allclose(oc, rgi(np.stack([y, x], 1)), rtol=1e-12)


def test_doppler_closed_form(oracle):
    rng = np.random.RandomState(2)
    nph = rng.poisson(5, 100).astype(np.int64)
    vn, v2 = rng.normal(0, 1000, 100), rng.uniform(0, 3e6, 100)
    e = rng.uniform(0.1, 10, nph.sum())
    e0 = e.copy()
    oracle.ref()["sky_functions"].doppler_shift(vn * oracle.scale_shift, v2 * oracle.scale_shift2, nph, e)
    fac = (1 + vn / 299792.458) / np.sqrt(1 - v2 / 299792.458 ** 2)
    np.testing.assert_allclose(e, e0 * np.repeat(fac, nph), rtol=1e-14)


def test_thermal_oracle_self_consistency(oracle):
    """thermal_process_data (RandomState draw order of base.py:466,480) against the exact
    expectation from thermal_spectra."""
    import helpers as H

    t = H.make_tables(nbins=200)
    om = H.oracle_model(oracle, t)
    kT, em = H.random_cells(300, seed=1)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=1e-44)
    cut, tot, lam = oracle.thermal_spectra(om, cfg, dict(kT=kT, em=em))
    nc, nph, idx, e = oracle.thermal_process_data(om, cfg, dict(kT=kT, em=em), np.random.RandomState(0))
    assert abs(nph.sum() - lam.sum()) < 5 * np.sqrt(lam.sum())
    assert e.size == nph.sum() and nc == idx.size
    assert e.min() >= 0.1 and e.max() <= 10.0


def test_golden_vectors(oracle):
    """Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the
    reference natives): the oracle must keep reproducing them bit for bit."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fresh = mg.compute(oracle)
    stored = np.load(os.path.join(HERE, "golden", "hotpath_golden.npz"))
    assert sorted(fresh) == sorted(stored.files)
    for k in stored.files:
        assert np.array_equal(fresh[k], stored[k]), k
