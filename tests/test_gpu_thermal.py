"""GPU parity: thermal photon generation (libpxb through the C ABI) against the oracle.

Tolerances (BASELINE.json north_star): expected counts and per-cell spectra 1e-12 relative in
fp64; sampled distributions KS / chi^2 at p > 0.01; fixed seed => bit-exact with itself.
"""
import numpy as np
import pytest
import torch

import pyxsim_b200 as px
from pyxsim_b200 import synthetic

import helpers as H

pytestmark = pytest.mark.gpu

TF = ("gas", "kT")


def _chunk(src):
    return next(src.chunks())


# ---------------------------------------------------------------------------------------
# table lookup: bit-exact against the reference Cython
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("nvar", [0, 3])
def test_interp1d_spec_bit_exact(cuda, oracle, nvar):
    t = H.make_tables(nbins=257, nvar=nvar)
    kT, _ = H.random_cells(200, kT_range=(0.01, 90.0))   # includes clamped extrapolation
    sm = H.product_model(t)
    sm.prepare_spectrum(0.0)
    c, m, v = sm.get_spectrum(kT)
    rc, rm, rv = H.oracle_model(oracle, t).get_spectrum(kT)
    assert np.array_equal(c, rc) and np.array_equal(m, rm)
    if nvar:
        assert np.array_equal(v, rv)
    else:
        assert v is None and rv is None


def _pion(oracle=None, nvar=0, nbins=300):
    Tv, Dv, ebins, c2, m2, v2, cie = synthetic.pion_like_tables(nT=12, nD=7, nbins=nbins, nvar=nvar)
    cie_T, cc, cm, cv = cie
    cie_prod = px.ThermalSpectralModel(ebins, cie_T, cc, cm, cv, logT=True, binscale="log")
    prod = px.PionSpectralModel(ebins, Tv, Dv, c2, m2, v2, cie_model=cie_prod, binscale="log")
    orc = None
    if oracle is not None:
        ocie = oracle.TableModel(ebins, cie_T, cc, cm, cv, logT=True, binscale="log")
        orc = oracle.PionModel(ebins, Tv, Dv, c2, m2, v2, ocie, binscale="log")
    return prod, orc


@pytest.mark.parametrize("nvar", [0, 2])
def test_interp2d_spec_and_fallback(cuda, oracle, nvar):
    prod, orc = _pion(oracle, nvar)
    prod.prepare_spectrum(0.0)
    rng = np.random.RandomState(5)
    kT = np.exp(rng.uniform(np.log(5e-4), np.log(5.0), 300))      # below, inside and above the table
    nH = np.exp(rng.uniform(np.log(3e-8), np.log(3e-3), 300))
    c, m, v = prod.get_spectrum(kT, nH)
    rc, rm, rv = orc.get_spectrum(kT, nH)
    # the 2-D blend is bit-exact; the host-side division by nH is the same NumPy expression
    assert np.array_equal(c, rc) and np.array_equal(m, rm)
    if nvar:
        assert np.array_equal(v, rv)


# ---------------------------------------------------------------------------------------
# expected counts + per-cell spectra: 1e-12
# ---------------------------------------------------------------------------------------
def _thermal_pair(oracle, t, *, Zmet=0.3, var_elem=None, extra=None, units=None, n=600, seed=3,
                  kT_range=(0.03, 60.0), **model_kw):
    kT, em = H.random_cells(n, seed=seed, kT_range=kT_range)
    src = H.cells_source(kT, em, extra=extra)
    if units:
        src.units.update(units)
    sm = H.product_model(t)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, t["ebins"].size - 1, Zmet, temperature_field=TF,
                                  var_elem=var_elem, prng=1234, binscale=t["binscale"], **model_kw)
    model.setup_model("photons", src, 0.0)
    return src, model, kT, em


def test_lambda_and_spectra_const_Z(cuda, oracle):
    t = H.make_tables(nbins=500)
    src, model, kT, em = _thermal_pair(oracle, t, kT_min=0.05, kT_max=40.0)
    norm = 3.0e-44
    lam, counts = model.expected_counts(_chunk(src), norm)
    spec = model.cell_spectra(_chunk(src)).cpu().numpy()
    cfg = dict(kT_min=0.05, kT_max=40.0, Zmet=0.3, spectral_norm=norm)
    cut, tot, rlam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, dict(kT=kT, em=em))
    lam = lam.cpu().numpy()
    assert cut.sum() > 100 and (~cut).sum() > 10
    assert np.all(lam[~cut] == 0) and np.all(counts.cpu().numpy()[~cut] == 0)
    np.testing.assert_allclose(lam[cut], rlam, rtol=1e-12, atol=0)
    np.testing.assert_allclose(spec[cut], tot, rtol=1e-12, atol=1e-300)
    assert np.all(spec[~cut] == 0)


def test_lambda_fields_var_elem_massfraction(cuda, oracle):
    """Metallicity + two element fields given as mass fractions with an X_H field:
    base.py:395-421 conversions (Zconvert/mconvert divided by X_H)."""
    t = H.make_tables(nbins=300, nvar=3)
    n = 500
    rng = np.random.RandomState(8)
    extra = {("gas", "metallicity"): rng.uniform(0.0, 0.02, n),
             ("gas", "O_fraction"): rng.uniform(0.0, 0.01, n),
             ("gas", "Fe_fraction"): rng.uniform(0.0, 0.002, n),
             ("gas", "h_fraction"): rng.uniform(0.7, 0.76, n),
             ("gas", "density"): 10 ** rng.uniform(-28, -25, n),
             ("gas", "entropy"): 10 ** rng.uniform(0, 3, n)}
    units = {("gas", "metallicity"): "dimensionless", ("gas", "O_fraction"): "dimensionless",
             ("gas", "Fe_fraction"): "Zsun"}
    var_elem = {"O": ("gas", "O_fraction"), "Ne": 0.7, "Fe": ("gas", "Fe_fraction")}
    src, model, kT, em = _thermal_pair(oracle, t, Zmet=("gas", "metallicity"), var_elem=var_elem,
                                       extra=extra, units=units, n=n, h_fraction=("gas", "h_fraction"),
                                       max_density=3e-26, min_entropy=5.0)
    norm = 1.0e-44
    lam, _ = model.expected_counts(_chunk(src), norm)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet="Z", Zconvert=model.Zconvert, z_by_xh=True,
               h_fraction="xh", spectral_norm=norm, max_density=3e-26, min_entropy=5.0,
               var_elem=[("O", model.mconvert["O"], True), (0.7, 1.0, False),
                         ("Fe", model.mconvert["Fe"], False)])
    a = dict(kT=kT, em=em, Z=extra[("gas", "metallicity")], O=extra[("gas", "O_fraction")],
             Fe=extra[("gas", "Fe_fraction")], xh=extra[("gas", "h_fraction")],
             density=extra[("gas", "density")], entropy=extra[("gas", "entropy")])
    cut, tot, rlam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, a)
    assert 50 < cut.sum() < n
    lam = lam.cpu().numpy()
    np.testing.assert_allclose(lam[cut], rlam, rtol=1e-12)
    assert np.all(lam[~cut] == 0)
    spec = model.cell_spectra(_chunk(src)).cpu().numpy()
    np.testing.assert_allclose(spec[cut], tot, rtol=1e-12, atol=1e-300)


def test_lambda_negative_tables_general_path(cuda, oracle):
    """Tables with negative entries: the clip of base.py:460 matters, the O(1) row-sum path
    is illegal and every cell goes through the general kernel."""
    t = H.make_tables(nbins=400, nvar=2, negative=True)
    var_elem = {"O": 0.5, "Fe": 1.5}
    src, model, kT, em = _thermal_pair(oracle, t, Zmet=0.4, var_elem=var_elem)
    assert not model.spectral_model.device_model().nonnegative
    norm = 2.0e-44
    lam, _ = model.expected_counts(_chunk(src), norm)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.4, spectral_norm=norm,
               var_elem=[(0.5, 1.0, False), (1.5, 1.0, False)])
    cut, tot, rlam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, dict(kT=kT, em=em))
    np.testing.assert_allclose(lam.cpu().numpy()[cut], rlam, rtol=1e-12)
    spec = model.cell_spectra(_chunk(src)).cpu().numpy()
    assert (tot == 0).sum() > 100           # the clip really fired
    np.testing.assert_allclose(spec[cut], tot, rtol=1e-12, atol=1e-300)


def test_lambda_extrapolated_cells(cuda, oracle):
    """kT outside the table nodes: clamped index + linear extrapolation (interpolate.pyx:36-41),
    where the blended spectrum can go negative and must be clipped."""
    t = H.make_tables(nbins=300)
    src, model, kT, em = _thermal_pair(oracle, t, kT_range=(0.005, 200.0), kT_min=0.001,
                                       kT_max=1000.0)
    norm = 1.0e-44
    lam, _ = model.expected_counts(_chunk(src), norm)
    cfg = dict(kT_min=0.001, kT_max=1000.0, Zmet=0.3, spectral_norm=norm)
    cut, tot, rlam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, dict(kT=kT, em=em))
    assert ((kT < t["Tvals"][0]) | (kT > t["Tvals"][-1])).sum() > 50
    np.testing.assert_allclose(lam.cpu().numpy()[cut], rlam, rtol=1e-12)


def test_lambda_pion_2d(cuda, oracle):
    prod, orc = _pion(oracle, nvar=2)
    n = 400
    rng = np.random.RandomState(21)
    kT = np.exp(rng.uniform(np.log(2e-3), np.log(4.0), n))
    nH = np.exp(rng.uniform(np.log(5e-8), np.log(2e-3), n))
    em = 1e60 * rng.lognormal(0, 1, n)
    src = H.cells_source(kT, em, extra={("gas", "H_nuclei_density"): nH})
    model = px.PionSourceModel(0.3, 1.4, prod.nbins, 0.3, binscale="log", temperature_field=TF,
                               var_elem={"O": 0.4, "Ne": 1.1}, spectral_model=prod, prng=3,
                               kT_min=1e-3)
    model.setup_model("photons", src, 0.0)
    norm = 1.0e-43
    lam, _ = model.expected_counts(_chunk(src), norm)
    cfg = dict(kT_min=1e-3, kT_max=64.0, Zmet=0.3, spectral_norm=norm,
               var_elem=[(0.4, 1.0, False), (1.1, 1.0, False)])
    cut, tot, rlam = oracle.thermal_spectra(orc, cfg, dict(kT=kT, em=em, nH=nH))
    use_igm = ((kT >= prod.min_table_kT) & (kT <= prod.max_table_kT) &
               (nH >= prod.min_table_nH) & (nH <= prod.max_table_nH))
    assert use_igm.sum() > 50 and (~use_igm).sum() > 50
    np.testing.assert_allclose(lam.cpu().numpy()[cut], rlam, rtol=1e-12)
    spec = model.cell_spectra(_chunk(src)).cpu().numpy()
    np.testing.assert_allclose(spec[cut], tot, rtol=1e-12, atol=1e-300)


def test_internal_observer_r2(cuda, oracle):
    t = H.make_tables(nbins=200)
    src, model, kT, em = _thermal_pair(oracle, t)
    pf = src.position_fields
    model.set_pv(pf, src.velocity_fields, le=np.array([-400.0] * 3), re=np.array([400.0] * 3),
                 dw=np.array([1000.0] * 3), c=np.array([10.0, -20.0, 30.0]),
                 periodicity=(True, False, True), observer="internal")
    r2 = model.compute_radius(_chunk(src)).cpu().numpy()
    pos = [src.fields[f] for f in pf]
    rr2 = oracle.compute_radius(pos, [-400.0] * 3, [400.0] * 3, [1000.0] * 3, [10.0, -20.0, 30.0],
                                (True, False, True))
    np.testing.assert_allclose(r2, rr2, rtol=1e-14)
    norm = 1.0e-12
    lam, _ = model.expected_counts(_chunk(src), norm)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=norm)
    cut, tot, rlam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, dict(kT=kT, em=em, r2=rr2))
    np.testing.assert_allclose(lam.cpu().numpy()[cut], rlam, rtol=1e-12)


# ---------------------------------------------------------------------------------------
# Poisson draws
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("lam0", [0.05, 0.9, 4.0, 9.99, 10.0, 37.5, 1234.5, 2.0e6])
def test_poisson_distribution(cuda, lam0):
    """Counts at a fixed lambda: mean/variance and a chi^2 on the pmf (line model with a
    constant rate gives every cell the same expectation)."""
    from scipy import stats

    n = 400000
    src = px.ArrayDataSource({("gas", "line"): np.full(n, lam0), ("index", "x"): np.zeros(n),
                              ("index", "y"): np.zeros(n), ("index", "z"): np.zeros(n)})
    model = px.LineSourceModel(3.0, ("gas", "line"), 0.01, prng=77)
    model.setup_model("photons", src, 0.0)
    lam, counts = model.expected_counts(_chunk(src), 1.0)
    assert np.allclose(lam.cpu().numpy(), lam0, rtol=1e-15)
    k = counts.cpu().numpy()
    assert k.min() >= 0
    se_mean = np.sqrt(lam0 / n)
    assert abs(k.mean() - lam0) < 5 * se_mean
    var_se = lam0 * np.sqrt(2.0 / n + 1.0 / (n * lam0))
    assert abs(k.var() - lam0) < 6 * var_se
    if lam0 < 2000:
        lo = int(max(0, stats.poisson.ppf(1e-4, lam0)))
        hi = int(stats.poisson.ppf(1 - 1e-4, lam0)) + 1
        edges = np.arange(lo, hi + 1)
        obs = np.array([(k == v).sum() for v in edges[:-1]], dtype=float)
        exp = n * stats.poisson.pmf(edges[:-1], lam0)
        keep = exp > 20
        obs_k, exp_k = obs[keep], exp[keep]
        rest_o, rest_e = n - obs_k.sum(), n - exp_k.sum()
        if rest_e > 20:
            obs_k, exp_k = np.append(obs_k, rest_o), np.append(exp_k, rest_e)
        chi2 = ((obs_k - exp_k) ** 2 / exp_k).sum()
        p = stats.chi2.sf(chi2, obs_k.size - 1)
        assert p > 1e-3, (chi2, obs_k.size)


# ---------------------------------------------------------------------------------------
# energy sampling
# ---------------------------------------------------------------------------------------
def _generate(model, src, norm):
    out = model.generate_chunk(_chunk(src), norm)
    return out


@pytest.mark.parametrize("binscale,negative", [("linear", False), ("log", False), ("linear", True)])
def test_energy_distribution_vs_oracle(cuda, oracle, binscale, negative):
    """Per-cell sampled energies: KS against the oracle's own samples (reference algorithm,
    RandomState) and chi^2 against the exact per-cell spectrum."""
    from scipy import stats

    t = H.make_tables(nbins=300, binscale=binscale, negative=negative)
    kT = np.array([0.3, 1.1, 6.0, 23.0, 0.8])
    em = np.array([4e62, 2e62, 3e62, 2e62, 1e59])
    src = H.cells_source(kT, em)
    sm = H.product_model(t)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 300, 0.3, temperature_field=TF, prng=5,
                                  binscale=binscale)
    model.setup_model("photons", src, 0.0)
    norm = 1.0e-43
    out = _generate(model, src, norm)
    nph = out.nph.cpu().numpy()
    idx = out.idx.cpu().numpy()
    e = out.energy.cpu().numpy()
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=norm)
    om = H.oracle_model(oracle, t)
    ncell, onph, oidx, oe = oracle.thermal_process_data(om, cfg, dict(kT=kT, em=em),
                                                        np.random.RandomState(11))
    cut, tot, lam = oracle.thermal_spectra(om, cfg, dict(kT=kT, em=em))
    assert e.size == nph.sum() and np.all(nph > 0)
    assert e.min() >= t["ebins"][0] and e.max() <= t["ebins"][-1]
    off = np.concatenate([[0], np.cumsum(nph)])
    ooff = np.concatenate([[0], np.cumsum(onph)])
    for k, i in enumerate(idx):
        mine = e[off[k]:off[k + 1]]
        assert abs(mine.size - lam[i]) < 6 * np.sqrt(lam[i]) + 5
        ko = np.where(oidx == i)[0]
        if ko.size and mine.size > 500:
            theirs = oe[ooff[ko[0]]:ooff[ko[0] + 1]]
            assert stats.ks_2samp(mine, theirs).pvalue > 0.01
        if mine.size > 5000:
            # chi^2 of bin counts against the exact spectrum, merged into ~30 groups
            h, _ = np.histogram(mine, bins=t["ebins"])
            p = tot[i] / tot[i].sum()
            grp = np.minimum((np.cumsum(p) * 30).astype(int), 29)
            ho = np.bincount(grp, weights=h, minlength=30)
            he = np.bincount(grp, weights=p * mine.size, minlength=30)
            ok = he > 10
            chi2 = ((ho[ok] - he[ok]) ** 2 / he[ok]).sum()
            assert stats.chi2.sf(chi2, ok.sum() - 1) > 1e-3
            # uniform inside a bin: positions within the most populated bin
            j = np.argmax(h)
            inb = mine[(mine >= t["ebins"][j]) & (mine < t["ebins"][j + 1])]
            if binscale == "linear":
                frac = (inb - t["ebins"][j]) / (t["ebins"][j + 1] - t["ebins"][j])
            else:
                frac = (np.log10(inb) - np.log10(t["ebins"][j])) / np.log10(t["ebins"][j + 1] / t["ebins"][j])
            assert stats.kstest(frac, "uniform").pvalue > 1e-3


def test_accept_reject_bin_centres(cuda):
    t = H.make_tables(nbins=200)
    src = H.cells_source(np.array([2.0, 5.0]), np.array([3e62, 3e62]))
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 200, 0.3, temperature_field=TF,
                                  prng=5, method="accept_reject")
    model.setup_model("photons", src, 0.0)
    out = _generate(model, src, 1e-43)
    e = out.energy.cpu().numpy()
    emid = 0.5 * (t["ebins"][1:] + t["ebins"][:-1])
    j = np.searchsorted(t["ebins"], e) - 1
    assert np.array_equal(e, emid[j])


def test_determinism_and_chunk_independence(cuda):
    """Fixed seed: bit-identical reruns; splitting the dataset into chunks does not change any
    photon (RNG counters are keyed by global cell id)."""
    t = H.make_tables(nbins=256)
    kT, em = H.random_cells(5000, seed=9, em_scale=1.0e63)

    def run(chunk_size):
        src = H.cells_source(kT, em, chunk_size=chunk_size)
        model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 256, 0.3, temperature_field=TF,
                                      prng=2024)
        n_ph, n_c = px.make_photons("mem:det", src, 0.05, 1000.0, 1.0e5, model)
        lst = px.load_list("mem:det")
        return n_ph, n_c, {k: lst.numpy(k) for k in lst.data}

    a = run(None)
    b = run(None)
    c = run(777)
    assert a[0] > 10000
    for other in (b, c):
        assert a[0] == other[0] and a[1] == other[1]
        for k in a[2]:
            assert np.array_equal(a[2][k], other[2][k]), k


def test_scan_and_compact(cuda):
    rng = np.random.RandomState(4)
    for n in (1, 7, 2048, 2049, 100003):
        counts = rng.poisson(0.7, n).astype(np.int64) * rng.randint(0, 3, n)
        src = px.ArrayDataSource({("index", "x"): rng.uniform(0, 1, n), ("index", "y"): rng.uniform(0, 1, n),
                                  ("index", "z"): rng.uniform(0, 1, n), ("index", "dx"): np.full(n, 0.1),
                                  ("gas", "velocity_x"): rng.normal(size=n),
                                  ("gas", "velocity_y"): rng.normal(size=n),
                                  ("gas", "velocity_z"): rng.normal(size=n)},
                                 domain_left_edge=[0, 0, 0], domain_right_edge=[1, 1, 1],
                                 periodicity=(True, True, False), left_edge=[0.2, 0.0, 0.0],
                                 right_edge=[0.8, 1.0, 1.0])
        m = px.SourceModel(prng=1)
        c = np.array([0.5, 0.5, 0.5])
        m.set_pv(src.position_fields, src.velocity_fields, le=src.left_edge, re=src.right_edge,
                 dw=src.domain_width, c=c, periodicity=src.periodicity, observer="external")
        gather = dict(p_fields=src.position_fields, v_fields=src.velocity_fields,
                      w_field=src.width_field, bulk_velocity=np.array([1.0, 2.0, 3.0]))
        out = m._scan_and_compact(torch.from_numpy(counts).cuda(), _chunk(src), gather)
        act = np.where(counts > 0)[0]
        assert out.n_photons == counts.sum() and out.n_cells == act.size
        if out.n_photons == 0:
            continue
        assert np.array_equal(out.idx.cpu().numpy(), act)
        assert np.array_equal(out.nph.cpu().numpy(), counts[act])
        assert np.array_equal(out.poff.cpu().numpy(), np.concatenate([[0], np.cumsum(counts[act])]))
        from oracle import ref_glue

        pos = [src.fields[f] for f in src.position_fields]
        vel = [src.fields[f] for f in src.velocity_fields]
        g = ref_glue.gather_active(pos, vel, src.fields[("index", "dx")], act, src.left_edge,
                                   src.right_edge, src.domain_width, c, [1.0, 2.0, 3.0],
                                   src.periodicity)
        for k in ("x", "y", "z", "vx", "vy", "vz", "dx"):
            assert np.array_equal(getattr(out, k).cpu().numpy(), g[k]), k


def test_energy_distribution_with_variable_elements(cuda, oracle):
    """Mixture sampler with per-cell abundance fields (metallicity + 2 elements): the component
    pick has to weight the var tables per cell."""
    from scipy import stats

    t = H.make_tables(nbins=300, nvar=2)
    n = 400
    rng = np.random.RandomState(14)
    kT = np.exp(rng.uniform(np.log(0.5), np.log(10.0), n))
    em = 3.0e61 * rng.lognormal(0, 0.5, n)
    extra = {("gas", "Z"): rng.uniform(0.0, 1.0, n), ("gas", "O"): rng.uniform(0.0, 3.0, n),
             ("gas", "Fe"): rng.uniform(0.0, 3.0, n)}
    src = H.cells_source(kT, em, extra=extra)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet="Z", Zconvert=1.0, z_by_xh=False, spectral_norm=1e-43,
               var_elem=[("O", 1.0, False), ("Fe", 1.0, False)])
    a = dict(kT=kT, em=em, Z=extra[("gas", "Z")], O=extra[("gas", "O")], Fe=extra[("gas", "Fe")])
    om = H.oracle_model(oracle, t)

    def run(seed_mine, seed_ref):
        model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 300, ("gas", "Z"),
                                      var_elem={"O": ("gas", "O"), "Fe": ("gas", "Fe")},
                                      temperature_field=TF, prng=seed_mine)
        model.setup_model("photons", src, 0.0)
        out = model.generate_chunk(_chunk(src), 1e-43)
        e = out.energy.cpu().numpy()
        nc, nph, idx, oe = oracle.thermal_process_data(om, cfg, a, np.random.RandomState(seed_ref))
        assert abs(e.size - oe.size) < 6 * np.sqrt(oe.size) and e.size > 100000
        # hard and soft halves of the cell list separately (different mixture weights)
        off = np.concatenate([[0], np.cumsum(out.nph.cpu().numpy())])
        ooff = np.concatenate([[0], np.cumsum(nph)])
        mi, oi = out.idx.cpu().numpy(), idx
        hot_m = np.concatenate([e[off[k]:off[k + 1]] for k in np.where(kT[mi] > 3.0)[0]])
        hot_o = np.concatenate([oe[ooff[k]:ooff[k + 1]] for k in np.where(kT[oi] > 3.0)[0]])
        return {"all": stats.ks_2samp(e, oe).pvalue, "hot": stats.ks_2samp(hot_m, hot_o).pvalue}

    H.assert_distributions_agree(run, seeds=((3, 4), (13, 14)))


def test_component_split_fractions_with_disjoint_bands(cuda, oracle, monkeypatch):
    """Multinomial split of a cell's photons over its mixture components (4 rows-x-tables for a
    plain table, rows x (2 + nvar) with variable elements).  The tables here emit in DISJOINT
    energy bands, so every photon's table can be read off its energy: per table, the counts
    summed over cells must match sum_cells N_cell * q_cell,table with q from the oracle's blended
    spectrum -- for bright cells in split mode (threshold lowered to the cells' size), for the same
    cells and for cells with a handful of photons under the per-photon choice from the stored
    cumulative probabilities, and under the per-photon choice that re-derives the weights."""
    nT, nbins, nvar = 12, 400, 2
    rng = np.random.RandomState(21)
    Tvals = np.logspace(np.log10(0.3), np.log10(20.0), nT)
    ebins = np.linspace(0.1, 10.0, nbins + 1)
    band = nbins // (2 + nvar)
    tabs = np.zeros((2 + nvar, nT, nbins))
    for tb in range(2 + nvar):
        shape = rng.uniform(0.2, 1.0, band) * 1e-16
        for r in range(nT):
            tabs[tb, r, tb * band:(tb + 1) * band] = shape * (1.0 + 0.3 * tb) * (r + 1.0) ** (0.5 * (tb + 1))
    t = dict(Tvals=Tvals, ebins=ebins, cosmic=tabs[0], metal=tabs[1], var=tabs[2:], binscale="linear")
    om = H.oracle_model(oracle, t)
    for target, label, sf in ((100.0, "bright", "3"), (100.0, "bright", None), (3.0, "dim", None),
                              (3.0, "dim", "-1")):
        if sf is None:
            monkeypatch.delenv("PXB_SPLIT_FACTOR", raising=False)
            monkeypatch.delenv("PXB_NO_GENERAL_SPLIT", raising=False)
        elif sf == "-1":
            monkeypatch.setenv("PXB_NO_GENERAL_SPLIT", "1")
        else:
            monkeypatch.setenv("PXB_SPLIT_FACTOR", sf)
        n = 3000
        kT = np.exp(rng.uniform(np.log(0.4), np.log(15.0), n))
        em = 1.0e60 * rng.lognormal(0, 0.3, n)
        extra = {("gas", "Z"): rng.uniform(0.0, 1.0, n), ("gas", "O"): rng.uniform(0.0, 2.0, n),
                 ("gas", "Fe"): rng.uniform(0.0, 2.0, n)}
        extra[("gas", "O")][::5] = 0.0          # empty components
        cfg = dict(kT_min=0.025, kT_max=64.0, Zmet="Z", Zconvert=1.0, z_by_xh=False, spectral_norm=1.0,
                   var_elem=[("O", 1.0, False), ("Fe", 1.0, False)])
        a = dict(kT=kT, em=em, Z=extra[("gas", "Z")], O=extra[("gas", "O")], Fe=extra[("gas", "Fe")])
        cut, tot, lam = oracle.thermal_spectra(om, cfg, a)
        assert cut.all()
        norm = target / np.median(lam)
        src = H.cells_source(kT, em, extra=extra)
        model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, nbins, ("gas", "Z"),
                                      var_elem={"O": ("gas", "O"), "Fe": ("gas", "Fe")},
                                      temperature_field=TF, prng=31)
        model.setup_model("photons", src, 0.0)
        out = model.generate_chunk(_chunk(src), norm)
        e = out.energy.cpu().numpy()
        nph = out.nph.cpu().numpy()
        idx = out.idx.cpu().numpy()
        assert (np.median(nph) > 40) if label == "bright" else (np.median(nph) < 8)
        q = np.stack([tot[:, tb * band:(tb + 1) * band].sum(axis=1) for tb in range(2 + nvar)], axis=1)
        q /= q.sum(axis=1, keepdims=True)
        comp = np.minimum(((e - 0.1) / (9.9 / nbins)).astype(int) // band, 1 + nvar)
        cell_of = np.repeat(np.arange(idx.size), nph)
        obs = np.zeros((idx.size, 2 + nvar))
        np.add.at(obs, (cell_of, comp), 1.0)
        N = nph[:, None].astype(float)
        qa = q[idx]
        assert np.all(obs[qa == 0.0] == 0.0)                       # empty components stay empty
        z = (obs.sum(0) - (N * qa).sum(0)) / np.sqrt((N * qa * (1 - qa)).sum(0))
        assert np.all(np.abs(z) < 4.5), (label, z)
        # per-cell: chi^2 over cells with enough photons is chi^2-distributed with (ncomp-1) dof each
        if label == "bright":
            from scipy import stats

            ok = (qa > 0).all(axis=1)
            chi2 = (((obs - N * qa) ** 2 / np.where(qa > 0, N * qa, 1.0))[ok]).sum()
            dof = ok.sum() * (1 + nvar)
            assert stats.chi2.sf(chi2, dof) > 1e-4 and stats.chi2.cdf(chi2, dof) > 1e-4, (chi2, dof)


def test_large_nbins_and_single_bright_cell(cuda, oracle):
    """nbins = 20000 (the reference's own test binning, tests/test_beta_model.py) does not fit
    the staged rows: the sampler gathers from L2 instead, and the general kernel holds a 160 KB
    CDF.  One very bright cell exercises a tile map with thousands of tiles per cell."""
    from scipy import stats

    t = H.make_tables(nbins=20000)
    kT = np.array([4.0, 0.02, 7.5])           # the middle cell extrapolates -> general kernel
    em = np.array([6.0e64, 2.0e67, 1.0e62])
    src = H.cells_source(kT, em)
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 20000, 0.3, temperature_field=TF,
                                  prng=9, kT_min=0.001)
    model.setup_model("photons", src, 0.0)
    out = model.generate_chunk(_chunk(src), 1e-43)
    nph = out.nph.cpu().numpy()
    assert nph[0] > 2_000_000
    cfg = dict(kT_min=0.001, kT_max=64.0, Zmet=0.3, spectral_norm=1e-43)
    cut, tot, lam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, dict(kT=kT, em=em))
    assert np.all(np.abs(nph - lam[out.idx.cpu().numpy()]) < 6 * np.sqrt(lam[out.idx.cpu().numpy()]) + 5)
    e = out.energy.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(nph)])
    for k, i in enumerate(out.idx.cpu().numpy()):
        mine = e[off[k]:off[k + 1]]
        if mine.size < 2000:
            continue
        h, _ = np.histogram(mine, bins=t["ebins"])
        p = tot[i] / tot[i].sum()
        grp = np.minimum((np.cumsum(p) * 40).astype(int), 39)
        ho = np.bincount(grp, weights=h, minlength=40)
        he = np.bincount(grp, weights=p * mine.size, minlength=40)
        ok = he > 10
        chi2 = ((ho[ok] - he[ok]) ** 2 / he[ok]).sum()
        assert stats.chi2.sf(chi2, ok.sum() - 1) > 1e-3, (i, chi2)


def test_empty_and_tiny_inputs(cuda):
    t = H.make_tables(nbins=128)
    sm = H.product_model(t)
    # no emission at all -> (0, 0) and an empty list that projects to zero events
    src = H.cells_source(np.full(50, 5.0), np.zeros(50))
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 128, 0.3, temperature_field=TF, prng=1)
    assert px.make_photons("mem:empty", src, 0.05, 1.0, 1.0, model) == (0, 0)
    assert px.load_list("mem:empty").data["energy"].numel() == 0
    assert px.project_photons("mem:empty", "mem:empty_ev", "z", [0.0, 0.0]) == 0
    # every cell cut by kT_min/kT_max
    src2 = H.cells_source(np.full(10, 1000.0), np.full(10, 1e65))
    assert px.make_photons("mem:cut", src2, 0.05, 1e3, 1e5, model) == (0, 0)
    # one cell, a handful of photons
    src3 = H.cells_source(np.array([3.0]), np.array([2.0e62]))
    model3 = px.ThermalSourceModel(sm, 0.1, 10.0, 128, 0.3, temperature_field=TF, prng=4)
    n_ph, n_c = px.make_photons("mem:one", src3, 0.05, 1.0e3, 1.0e5, model3)
    assert n_c == 1 and n_ph == px.load_list("mem:one").data["energy"].numel() > 0
    assert px.project_photons("mem:one", "mem:one_ev", [0, 1.0, 1.0], [0.0, 0.0], no_shifting=True) == n_ph


@pytest.mark.parametrize("n,p", [(1, 0.3), (7, 0.5), (40, 0.01), (1000, 0.004), (1000, 0.02), (50, 0.37),
                                 (100000, 0.3), (100000, 0.9993), (3000000, 0.5), (12, 0.97),
                                 (10**9, 1e-9), (10**9, 0.25)])
def test_binomial_sampler(cuda, n, p):
    """The device binomial sampler behind the multinomial split of a cell's photons over its
    table rows: mean/variance and a chi^2 on the pmf against scipy.stats.binom."""
    import ctypes as C

    from scipy import stats

    from pyxsim_b200 import _lib
    from pyxsim_b200.device import empty, ptr, stream_ptr

    count = 400000
    out = empty(count, torch.int64)
    _lib.call("pxb_rng_binomial", n, C.c_double(p), C.c_uint64(99), count, ptr(out), stream_ptr())
    k = out.cpu().numpy()
    assert k.min() >= 0 and k.max() <= n
    mu, var = n * p, n * p * (1 - p)
    assert abs(k.mean() - mu) < 5 * np.sqrt(var / count) + 1e-12
    assert abs(k.var() - var) < 8 * var * np.sqrt(2.0 / count) + 6 * np.sqrt(var / count) + 1e-12
    lo = int(stats.binom.ppf(1e-5, n, p))
    hi = int(stats.binom.ppf(1 - 1e-5, n, p)) + 1
    if hi - lo < 5000:
        vals = np.arange(lo, hi + 1)
        obs = np.array([(k == v).sum() for v in vals], dtype=float)
        exp = count * stats.binom.pmf(vals, n, p)
        keep = exp > 20
        o, e = obs[keep], exp[keep]
        ro, re = count - o.sum(), count - e.sum()
        if re > 20:
            o, e = np.append(o, ro), np.append(e, re)
        if o.size > 1:
            chi2 = ((o - e) ** 2 / e).sum()
            assert stats.chi2.sf(chi2, o.size - 1) > 1e-3, (chi2, o.size)
