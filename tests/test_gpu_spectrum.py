"""GPU parity for the spectrum / band modes ("next" row f1): libpxb against the reference's
compiled native routines (pyxsim/lib/spectra.pyx: shift_spectrum, make_band,
power_law_spectrum, line_spectrum) and against the reference's own Python output
(tests/golden/reference_golden.npz, key t1_spec)."""
import os

import numpy as np
import pytest
import torch

import pyxsim_b200 as px
from pyxsim_b200 import synthetic

import helpers as H

pytestmark = pytest.mark.gpu
TF = ("gas", "kT")


def _setup(oracle, nbins=300, nvar=2, n=700, negative=False):
    t = H.make_tables(nbins=nbins, nvar=nvar, negative=negative)
    kT, em = H.random_cells(n, seed=4, kT_range=(0.02, 70.0))
    src = H.cells_source(kT, em)
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, nbins, 0.3, temperature_field=TF,
                                  var_elem={"O": 0.5, "Fe": 1.5} if nvar else None, prng=1)
    model.setup_model("spectrum", src, 0.0)
    model.v_fields = src.velocity_fields
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=2e-44,
               var_elem=[(0.5, 1.0, False), (1.5, 1.0, False)] if nvar else None)
    return t, kT, em, src, model, cfg


@pytest.mark.parametrize("negative", [False, True])
def test_thermal_spectrum_and_band_vs_reference_cython(cuda, oracle, negative):
    t, kT, em, src, model, cfg = _setup(oracle, negative=negative)
    chunk = next(src.chunks())
    om = H.oracle_model(oracle, t)
    ebnew = np.linspace(0.25, 8.0, 201)
    # no shifting
    spec = model.process_data("spectrum", chunk, 2e-44, ebins=ebnew, shifting=False)
    ref = oracle.thermal_spectrum_mode(om, cfg, dict(kT=kT, em=em), ebnew)
    np.testing.assert_allclose(spec, ref, rtol=1e-11, atol=ref.max() * 1e-14)
    # Doppler shifting along an off-axis direction
    normal = np.array([0.3, -0.5, 0.8])
    model._spec_normal = normal
    spec_s = model.process_data("spectrum", chunk, 2e-44, ebins=ebnew, shifting=True)
    v = np.stack([src.fields[f] for f in src.velocity_fields]) / 299792.458
    nh = normal / np.linalg.norm(normal)
    shift = np.sqrt(1.0 - (v ** 2).sum(0)) / (1.0 - nh @ v)
    cut, tot, lam = oracle.thermal_spectra(om, cfg, dict(kT=kT, em=em))
    ref_s = oracle.thermal_spectrum_mode(om, cfg, dict(kT=kT, em=em), ebnew, shift=shift[cut])
    np.testing.assert_allclose(spec_s, ref_s, rtol=1e-11, atol=ref_s.max() * 1e-14)
    assert not np.allclose(spec_s, spec, rtol=1e-6)
    # make_band, both flavours, with and without shifting
    emid = 0.5 * (t["ebins"][1:] + t["ebins"][:-1])
    cnm = em[cut] * 2e-44
    mb = oracle.ref()["spectra"].make_band
    for use_energy in (0, 1):
        for shifting in (False, True):
            out = model.band_per_cell(chunk, 2e-44, 0.5, 7.0, energy=bool(use_energy), shifting=shifting)
            sh = shift[cut] if shifting else np.ones(cut.sum())
            refb = mb(use_energy, 0.5, 7.0, t["ebins"], emid, np.ascontiguousarray(tot), np.ascontiguousarray(sh)) * cnm
            got = out.cpu().numpy()
            np.testing.assert_allclose(got[cut], refb, rtol=1e-12, atol=refb.max() * 1e-15)
            assert np.all(got[~cut] == 0)


def test_thermal_spectrum_vs_reference_python(cuda):
    """process_data(mode="spectrum") against the reference's own Python run (fixture t1_spec)."""
    import importlib.util

    here = os.path.dirname(os.path.abspath(__file__))
    spec_ = importlib.util.spec_from_file_location("mrg", os.path.join(here, "golden", "make_reference_golden.py"))
    mrg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mrg)
    a = mrg.inputs()["t1"]
    R = np.load(os.path.join(here, "golden", "reference_golden.npz"))
    t = a["tables"]
    src = H.cells_source(a["kT"], a["em"], extra={("gas", "Zf"): a["Z"], ("gas", "Fef"): a["Fe"],
                                                  ("gas", "density"): a["dens"]})
    src.units[("gas", "Zf")] = "dimensionless"
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 64, ("gas", "Zf"), temperature_field=TF,
                                  var_elem={"O": 0.5, "Fe": ("gas", "Fef")}, h_fraction=a["h_fraction"],
                                  max_density=3e-26, prng=1)
    model.setup_model("spectrum", src, 0.0)
    model.Zconvert = a["Zconvert"]            # the fixture was generated with this conversion factor
    spec = model.process_data("spectrum", next(src.chunks()), a["norm"], ebins=np.linspace(0.3, 7.0, 41))
    np.testing.assert_allclose(spec, R["t1_spec"], rtol=1e-11, atol=R["t1_spec"].max() * 1e-14)


def test_powerlaw_and_line_spectrum_vs_reference_cython(cuda, oracle):
    sp = oracle.ref()["spectra"]
    rng = np.random.RandomState(6)
    n = 3000
    alpha = rng.uniform(0.5, 2.5, n)
    alpha[:50] = 1.0
    alpha[50:100] = 2.0
    lum = 1e40 * rng.lognormal(0, 1, n)
    f = {("gas", "lum"): lum, ("gas", "alpha"): alpha, ("gas", "line"): 1e38 * rng.lognormal(0, 1, n),
         ("gas", "sig"): rng.uniform(100.0, 2000.0, n), ("index", "x"): np.zeros(n), ("index", "y"): np.zeros(n),
         ("index", "z"): np.zeros(n), ("gas", "velocity_x"): rng.normal(0, 800, n),
         ("gas", "velocity_y"): rng.normal(0, 800, n), ("gas", "velocity_z"): rng.normal(0, 800, n)}
    src = px.ArrayDataSource(f, units={("gas", "lum"): "keV/s"})
    z = 0.3
    sf = 1.0 / (1.0 + z)
    normal = np.array([0.0, 0.6, 0.8])
    v = np.stack([f[("gas", f"velocity_{ax}")] for ax in "xyz"]) / 299792.458
    shift = np.sqrt(1.0 - (v ** 2).sum(0)) / (1.0 - normal @ v)

    class Bar:
        def update(self, *a):
            pass

    # power law
    pl = px.PowerLawSourceModel(1.0, 0.5, 40.0, ("gas", "lum"), ("gas", "alpha"), prng=1)
    eb, spec = pl.make_spectrum(src, 0.4, 30.0, 120, redshift=z, normal=normal)
    K_fac = 40.0 ** (2.0 - alpha) - 0.5 ** (2.0 - alpha)
    K_fac[alpha == 2] = np.log(40.0 / 0.5)
    K_fac[alpha != 2] /= 2.0 - alpha[alpha != 2]
    Kc = lum / K_fac
    emid = 0.5 * (eb[1:] + eb[:-1]) * (1.0 / sf) / 1.0
    ref = sp.power_law_spectrum(n, emid, alpha, Kc, shift, Bar())
    d_l = px.Cosmology().angular_diameter_distance_mpc(0.0, z) * (1 + z) ** 2 * 3.0856775809623245e24
    ref = ref * (1 + z) ** 2 / (4 * np.pi * d_l ** 2)
    np.testing.assert_allclose(spec, ref, rtol=1e-11)
    # line
    ln = px.LineSourceModel(3.5, ("gas", "line"), ("gas", "sig"), prng=1)
    eb, spec = ln.make_spectrum(src, 3.0, 4.0, 150, redshift=z, normal=normal)
    from scipy.stats import norm

    gx = np.linspace(-7, 7, 10000)
    ee = 0.5 * (eb[1:] + eb[:-1]) / sf
    sig_keV = f[("gas", "sig")] * 3.5 / 299792.458
    ref = sp.line_spectrum(n, 3.5, ee, sig_keV, gx, norm.pdf(gx), f[("gas", "line")], shift, Bar())
    ref = ref * (1 + z) ** 2 / (4 * np.pi * d_l ** 2)
    np.testing.assert_allclose(spec, ref, rtol=1e-10, atol=ref.max() * 1e-13)


def test_make_spectrum_thermal_counts_consistency(cuda):
    """The source-frame count-rate spectrum integrates to the total expected photon rate used by
    make_photons (sum of lambda at unit area*time, 4 pi D^2 removed)."""
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=800)
    src = synthetic.beta_model_source(20, kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda")
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 800, 0.3, temperature_field=TF, prng=1)
    eb, spec = model.make_spectrum(src, 0.1, 10.0, 800)
    rate = (spec * np.diff(eb)).sum()
    model.setup_model("photons", src, 0.0)
    lam, _ = model.expected_counts(next(src.chunks()), 1.0)
    np.testing.assert_allclose(rate, float(lam.sum()), rtol=1e-10)
    with pytest.raises(RuntimeError):
        model.make_spectrum(src, 0.1, 10.0, 10, normal="z")


# ---------------------------------------------------------------------------------------
# row f2: band flux functions (make_fluxf) and the per-cell field modes
# ---------------------------------------------------------------------------------------
def _golden():
    import importlib.util

    here = os.path.dirname(os.path.abspath(__file__))
    spec_ = importlib.util.spec_from_file_location("mrg", os.path.join(here, "golden", "make_reference_golden.py"))
    mrg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mrg)
    return mrg.inputs(), np.load(os.path.join(here, "golden", "reference_golden.npz"))


def test_band_flux_modes_vs_reference_python(cuda):
    """make_fluxf + process_data("luminosity" / "photon_rate") against the reference's own Python
    run (fixtures t1_lum, t1_rate, t1_*_cf/mf/vf; pion_lum, pion_rate): 1-D table with a
    mass-fraction metallicity field, one constant and one field element; 2-D table with the CIE
    fallback."""
    I, R = _golden()
    a = I["t1"]
    t = a["tables"]
    src = H.cells_source(a["kT"], a["em"], extra={("gas", "Zf"): a["Z"], ("gas", "Fef"): a["Fe"],
                                                  ("gas", "density"): a["dens"]})
    src.units[("gas", "Zf")] = "dimensionless"
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 64, ("gas", "Zf"), temperature_field=TF,
                                  var_elem={"O": 0.5, "Fe": ("gas", "Fef")}, h_fraction=a["h_fraction"],
                                  max_density=3e-26, prng=1)
    model.setup_model("fields", src, 0.0)
    model.Zconvert = a["Zconvert"]
    chunk = next(src.chunks())
    for tag, energy, mode in (("lum", True, "luminosity"), ("rate", False, "photon_rate")):
        fl = model.make_fluxf(0.5, 2.0, energy=energy)
        got = model.process_data(mode, chunk, a["norm"], emin=0.5, emax=2.0, fluxf=fl)
        ref = R[f"t1_{tag}"]
        assert got.shape == ref.shape and np.array_equal(got == 0.0, ref == 0.0)
        np.testing.assert_allclose(got, ref, rtol=2e-15, atol=0.0)
        c_, m_, v_ = fl(np.array([0.05, 0.3, 1.0, 7.7, 40.0]))
        np.testing.assert_array_equal(c_, R[f"t1_{tag}_cf"])        # same operations, same order
        np.testing.assert_array_equal(m_, R[f"t1_{tag}_mf"])
        np.testing.assert_array_equal(v_, R[f"t1_{tag}_vf"])
        # without fluxf= the model tabulates the band itself
        got2 = model.process_data(mode, chunk, a["norm"], emin=0.5, emax=2.0)
        np.testing.assert_array_equal(got2, got)
    # outside the table scipy's interp1d raises; so does the GPU path
    with pytest.raises(ValueError):
        model.make_fluxf(0.5, 2.0)(np.array([1.0, 500.0]))

    a = I["pion"]
    cie_T, cc, cm, cv = a["cie"]
    cie_prod = px.ThermalSpectralModel(a["ebins"], cie_T, cc, cm, cv, logT=True, binscale="log")
    prod = px.PionSpectralModel(a["ebins"], a["Tv"], a["Dv"], a["c2"], a["m2"], a["v2"], cie_model=cie_prod,
                                var_elem_names=["O", "Ne"], binscale="log")
    src = H.cells_source(a["kT"], a["em"], extra={("gas", "H_nuclei_density"): a["nH"]})
    model = px.IGMSourceModel(0.3, 1.4, 40, 0.3, binscale="log", temperature_field=TF,
                              var_elem={"O": 0.4, "Ne": 1.1}, kT_min=0.1, prng=1, spectral_model=prod)
    model.setup_model("fields", src, 0.0)
    chunk = next(src.chunks())
    for tag, energy, mode in (("lum", True, "luminosity"), ("rate", False, "photon_rate")):
        got = model.process_data(mode, chunk, a["norm"], emin=0.4, emax=1.2,
                                 fluxf=model.make_fluxf(0.4, 1.2, energy=energy))
        ref = R[f"pion_{tag}"]
        assert np.array_equal(got == 0.0, ref == 0.0) and (ref > 0).sum() > 50
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0.0)   # log10 differs in the last ulp


def test_source_and_intensity_fields(cuda, oracle):
    """make_source_fields / make_intensity_fields on an array source: field names and units of the
    reference (sources.py:166-308,351-500); luminosity = emissivity * volume; the Doppler-free
    intensity equals the shifted one for a gas at rest up to the band-edge convention
    (make_fluxf keeps bins strictly inside the band, make_band bins touching it as well)."""
    t = H.make_tables(nbins=400, nvar=0)
    kT, em = H.random_cells(3000, seed=9, kT_range=(0.3, 20.0))
    src = H.cells_source(kT, em)
    for f in src.velocity_fields:
        src.fields[f] = np.zeros_like(src.fields[f])
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 400, 0.3, temperature_field=TF, prng=1)
    names = model.make_source_fields(src, 0.5, 2.0)
    assert [n[1] for n in names] == ["xray_emissivity_0.5_2.0_keV", "xray_luminosity_0.5_2.0_keV",
                                    "xray_photon_emissivity_0.5_2.0_keV", "xray_count_rate_0.5_2.0_keV"]
    lum, emis = src.fields[names[1]].cpu().numpy(), src.fields[names[0]].cpu().numpy()
    vol = (src.fields[("index", "dx")] * 3.0856775809623245e21) ** 3
    np.testing.assert_allclose(emis * vol, lum, rtol=1e-13)
    om = H.oracle_model(oracle, t)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=1.0)
    ref = oracle.thermal_field_mode(om, cfg, dict(kT=kT, em=em), "luminosity", 0.5, 2.0,
                                    fluxf=om.make_fluxf(0.5, 2.0, energy=True)) * 1.602176634e-9
    np.testing.assert_allclose(lum, ref, rtol=1e-13)
    with pytest.raises(RuntimeError):
        model.make_source_fields(src, 0.5, 2.0)                    # exists, no force_override
    with pytest.raises(ValueError):
        model.make_intensity_fields(src, 0.5, 2.0)                 # needs redshift or dist
    # band edges placed mid-bin so both conventions select the same bins
    eb = t["ebins"]
    lo, hi = 0.5 * (eb[40] + eb[41]) , 0.5 * (eb[150] + eb[151])
    z = 0.05
    n1 = model.make_intensity_fields(src, lo / (1 + z), hi / (1 + z), redshift=z, no_doppler=True, band_name="b")
    i_nd = {k: src.fields[k].cpu().numpy() for k in n1}
    n2 = model.make_intensity_fields(src, lo / (1 + z), hi / (1 + z), redshift=z, no_doppler=False, band_name="b")
    assert n1 == n2 and [n[1] for n in n1] == ["xray_intensity_b", "xray_photon_intensity_b"]
    for k in n1:
        np.testing.assert_allclose(src.fields[k].cpu().numpy(), i_nd[k], rtol=1e-12)
        assert (i_nd[k] > 0).all()


def test_unshifted_spectrum_by_linearity_equals_per_cell_path(cuda, oracle, monkeypatch):
    """Without a Doppler shift the spectrum is assembled from per-row weights (O(1) per cell);
    cells whose clip matters (here: temperatures outside the table, which extrapolate to negative
    spectra) are listed and take the per-cell kernel.  Both routes against the reference Cython,
    and against each other."""
    t = H.make_tables(nT=12, nbins=300, nvar=2)
    t["Tvals"] = np.logspace(np.log10(0.8), np.log10(9.0), 12)       # narrow table: many cells extrapolate
    kT, em = H.random_cells(5000, seed=12, kT_range=(0.1, 40.0))
    extra = {("gas", "Z"): np.random.RandomState(1).uniform(0, 1, kT.size)}
    src = H.cells_source(kT, em, extra=extra)
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 300, ("gas", "Z"), temperature_field=TF,
                                  var_elem={"O": 0.5, "Fe": 1.5}, prng=1)
    model.setup_model("spectrum", src, 0.0)
    chunk = next(src.chunks())
    ebnew = np.linspace(0.3, 8.0, 151)
    lin = model.process_data("spectrum", chunk, 3e-44, ebins=ebnew, shifting=False)
    monkeypatch.setenv("PXB_SPECTRUM_PER_CELL", "1")
    per = model.process_data("spectrum", chunk, 3e-44, ebins=ebnew, shifting=False)
    monkeypatch.delenv("PXB_SPECTRUM_PER_CELL")
    np.testing.assert_allclose(lin, per, rtol=1e-11, atol=per.max() * 1e-14)
    om = H.oracle_model(oracle, t)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet="Z", Zconvert=1.0, z_by_xh=False, spectral_norm=3e-44,
               var_elem=[(0.5, 1.0, False), (1.5, 1.0, False)])
    ref = oracle.thermal_spectrum_mode(om, cfg, dict(kT=kT, em=em, Z=extra[("gas", "Z")]), ebnew)
    np.testing.assert_allclose(lin, ref, rtol=1e-11, atol=ref.max() * 1e-14)
    inside = (kT >= t["Tvals"][0]) & (kT <= t["Tvals"][-1])
    assert 500 < inside.sum() < 4500          # both routes carried a real share of the cells
    # the same split for make_band: O(1) prefix-sum differences for the clip-free cells, the
    # warp-per-cell sum for the others; shifted and unshifted, photon and energy bands
    model.v_fields = src.velocity_fields
    model._spec_normal = np.array([0.3, -0.5, 0.8])
    v = np.stack([src.fields[f] for f in src.velocity_fields]) / 299792.458
    nh = model._spec_normal / np.linalg.norm(model._spec_normal)
    shift = np.sqrt(1.0 - (v ** 2).sum(0)) / (1.0 - nh @ v)
    emid = 0.5 * (t["ebins"][1:] + t["ebins"][:-1])
    cut, tot, lam = oracle.thermal_spectra(om, cfg, dict(kT=kT, em=em, Z=extra[("gas", "Z")]))
    cnm = em[cut] * 3e-44
    mb = oracle.ref()["spectra"].make_band
    for use_energy in (0, 1):
        for shifting in (False, True):
            for band in ((0.5, 7.0), (1.0, 1.2), (0.05, 20.0), (3.0, 3.001)):
                fast = model.band_per_cell(chunk, 3e-44, band[0], band[1], energy=bool(use_energy),
                                           shifting=shifting).cpu().numpy()
                monkeypatch.setenv("PXB_BAND_PER_CELL", "1")
                slow = model.band_per_cell(chunk, 3e-44, band[0], band[1], energy=bool(use_energy),
                                           shifting=shifting).cpu().numpy()
                monkeypatch.delenv("PXB_BAND_PER_CELL")
                sh = shift[cut] if shifting else np.ones(cut.sum())
                refb = mb(use_energy, band[0], band[1], t["ebins"], emid, np.ascontiguousarray(tot),
                          np.ascontiguousarray(sh)) * cnm
                scale = np.abs(refb).max() + 1e-300
                np.testing.assert_allclose(slow[cut], refb, rtol=1e-12, atol=scale * 1e-15)
                # prefix differences cancel: absolute accuracy is that of the row totals
                np.testing.assert_allclose(fast[cut], refb, rtol=1e-10, atol=scale * 1e-13)
                assert np.array_equal(fast[~cut], np.zeros((~cut).sum()))
                assert np.array_equal(fast == 0.0, slow == 0.0) or band == (3.0, 3.001)


@pytest.mark.parametrize("wide", [False, True])
def test_shifted_spectrum_bucketed_by_row_vs_reference_cython(cuda, oracle, monkeypatch, wide):
    """Doppler-shifted spectra without variable elements: cells whose clip is a no-op are
    bucketed by table row and regridded from the rows' prefix sums (one lane per cell); the
    others (temperatures outside a narrow table) take the per-cell kernel.  Against the
    reference's compiled shift_spectrum and against the per-cell route; ``wide``: new edges
    beyond both ends of the model's range (np.interp holds the end values there)."""
    t = H.make_tables(nT=12, nbins=300, nvar=0)
    t["Tvals"] = np.logspace(np.log10(0.8), np.log10(9.0), 12)
    kT, em = H.random_cells(6000, seed=21, kT_range=(0.1, 40.0))
    extra = {("gas", "Z"): np.random.RandomState(2).uniform(0, 1, kT.size)}
    src = H.cells_source(kT, em, extra=extra)
    model = px.ThermalSourceModel(H.product_model(t), 0.1, 10.0, 300, ("gas", "Z"), temperature_field=TF,
                                  prng=1)
    model.setup_model("spectrum", src, 0.0)
    model.v_fields = src.velocity_fields
    model._spec_normal = np.array([0.3, -0.5, 0.8])
    chunk = next(src.chunks())
    ebnew = np.linspace(0.02, 14.0, 333) if wide else np.linspace(0.3, 8.0, 151)
    got = model.process_data("spectrum", chunk, 3e-44, ebins=ebnew, shifting=True)
    monkeypatch.setenv("PXB_SPECTRUM_PER_CELL", "1")
    per = model.process_data("spectrum", chunk, 3e-44, ebins=ebnew, shifting=True)
    monkeypatch.delenv("PXB_SPECTRUM_PER_CELL")
    np.testing.assert_allclose(got, per, rtol=1e-11, atol=per.max() * 1e-14)
    v = np.stack([src.fields[f] for f in src.velocity_fields]) / 299792.458
    nh = model._spec_normal / np.linalg.norm(model._spec_normal)
    shift = np.sqrt(1.0 - (v ** 2).sum(0)) / (1.0 - nh @ v)
    om = H.oracle_model(oracle, t)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet="Z", Zconvert=1.0, z_by_xh=False, spectral_norm=3e-44)
    cut, tot, lam = oracle.thermal_spectra(om, cfg, dict(kT=kT, em=em, Z=extra[("gas", "Z")]))
    ref = oracle.thermal_spectrum_mode(om, cfg, dict(kT=kT, em=em, Z=extra[("gas", "Z")]), ebnew,
                                       shift=shift[cut])
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=ref.max() * 1e-14)
    inside = (kT >= t["Tvals"][0]) & (kT <= t["Tvals"][-1])
    assert 500 < inside.sum() < 5500          # both routes carried a real share of the cells
    assert np.abs(shift - 1.0).max() > 1e-3
