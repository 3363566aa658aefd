"""GPU: whole-path tests through the public API (make_photons -> project_photons), including
size-independent properties at larger sizes and file round trips."""
import numpy as np
import pytest
import torch
from scipy import stats

import pyxsim_b200 as px
from pyxsim_b200 import synthetic

import helpers as H

pytestmark = pytest.mark.gpu
TF = ("gas", "kT")


def _cluster(n=32, nbins=1000, **kw):
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=nbins)
    src = synthetic.beta_model_source(n, v3d_sigma_kms=300.0, **kw)
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    return src, sm, (kT_nodes, ebins, cosmic, metal)


def test_beta_model_end_to_end_vs_oracle(cuda, oracle):
    """configs[0]-shaped run (beta-model cluster, thermal source, tabulated absorption, project
    along z): photon and event counts within Poisson/binomial error of the oracle's, energy and
    sky distributions KS-compatible -- the reference's acceptance style (tests/test_beta_model.py)."""
    src, sm, (kT_nodes, ebins, cosmic, metal) = _cluster(32, kT_range=(2.0, 9.0))
    en, sg = synthetic.tbabs_like_sigma()
    area, texp, z = 3000.0, 1.0e5, 0.05

    def run(seed_mine, seed_ref):
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 1000, 0.3, temperature_field=TF, prng=seed_mine)
        n_ph, n_c = px.make_photons("mem:e2e_ph", src, z, area, texp, model)
        n_ev = px.project_photons("mem:e2e_ph", "mem:e2e_ev", "z", [30.0, 45.0],
                                  absorb_model=px.TBabsModel(0.02, energy=en, cross_section=sg),
                                  nH=0.02, prng=seed_mine + 1)
        ph, ev = px.load_list("mem:e2e_ph"), px.load_list("mem:e2e_ev")
        # oracle on identical inputs
        f = {k: v.cpu().numpy() for k, v in src.fields.items()}
        D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, z)
        norm = area * texp / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * (1 + z) ** 2)
        om = oracle.TableModel(ebins, kT_nodes, cosmic, metal)
        cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=norm)
        a = dict(kT=f[TF], em=f[("gas", "emission_measure")])
        nc, nph, idx, e = oracle.thermal_process_data(om, cfg, a, np.random.RandomState(seed_ref))
        cut, tot, lam = oracle.thermal_spectra(om, cfg, a)
        assert abs(n_ph - lam.sum()) < 5 * np.sqrt(lam.sum())
        assert abs(n_ph - nph.sum()) < 5 * np.sqrt(2 * lam.sum())
        g = oracle.gather_active([f[("index", ax)] for ax in "xyz"],
                                 [f[("gas", f"velocity_{ax}")] for ax in "xyz"], f[("index", "dx")],
                                 idx, [-1000.0] * 3, [1000.0] * 3, [2000.0] * 3, [0.0] * 3, [0.0] * 3,
                                 (False,) * 3)
        d = dict(g, num_photons=nph, energy=e)
        params = dict(data_type="cells", observer="external", fid_d_a=D_A, fid_redshift=z)
        ref = oracle.project_photons(d, params, "z", [30.0, 45.0], np.random.RandomState(seed_ref + 1),
                                     absorb_model=oracle.TableAbsorb(0.02, en, sg * 1.0e22), nH=0.02)
        frac_m, frac_r = n_ev / n_ph, ref["n_events"] / nph.sum()
        assert abs(frac_m - frac_r) < 5 * np.sqrt(2 * frac_r * (1 - frac_r) / nph.sum())
        return {"energy": stats.ks_2samp(ph.numpy("energy"), e).pvalue,
                "eobs": stats.ks_2samp(ev.numpy("eobs"), ref["eobs"]).pvalue,
                "xsky": stats.ks_2samp(ev.numpy("xsky"), ref["xsky"]).pvalue,
                "ysky": stats.ks_2samp(ev.numpy("ysky"), ref["ysky"]).pvalue}

    H.assert_distributions_agree(run, seeds=((50, 7), (60, 9)))


def test_plugin_contract_and_file_roundtrip(cuda, tmp_path):
    """process_data returns the reference tuple; lists written to disk (npz fallback of the HDF5
    schema) project to the same events as the in-memory list."""
    src, sm, _ = _cluster(16, nbins=500)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 500, 0.3, temperature_field=TF, prng=3)
    model.setup_model("photons", src, 0.05)
    ncells, nph, idxs, energies = model.process_data("photons", next(src.chunks()), 1.0e-45)
    assert isinstance(nph, np.ndarray) and nph.dtype == np.int64 and energies.dtype == np.float64
    assert ncells == nph.size == idxs.size and energies.size == nph.sum() and np.all(nph > 0)
    assert np.all(np.diff(idxs) > 0)
    with pytest.raises(NotImplementedError):
        model.process_data("energy_field", next(src.chunks()), 1.0)
    prefix = str(tmp_path / "photons")
    model2 = px.ThermalSourceModel(sm, 0.1, 10.0, 500, 0.3, temperature_field=TF, prng=3)
    n1 = px.make_photons(prefix + ".h5", src, 0.05, 5.0e4, 1.0e5, model2)
    model3 = px.ThermalSourceModel(sm, 0.1, 10.0, 500, 0.3, temperature_field=TF, prng=3)
    n2 = px.make_photons("mem:rt", src, 0.05, 5.0e4, 1.0e5, model3)
    assert n1 == n2
    e1 = px.project_photons(prefix, str(tmp_path / "events"), [0.1, 0.2, 0.9], [12.0, -3.0],
                            absorb_model="wabs", nH=0.03, prng=8)
    e2 = px.project_photons("mem:rt", "mem:rt_ev", [0.1, 0.2, 0.9], [12.0, -3.0],
                            absorb_model="wabs", nH=0.03, prng=8)
    assert e1 == e2
    a, b = px.load_list(str(tmp_path / "events")), px.load_list("mem:rt_ev")
    for k in ("xsky", "ysky", "eobs"):
        assert np.array_equal(a.numpy(k), b.numpy(k))
    assert a.parameters["absoption_model"] == "wabs" and float(a.parameters["flux"]) == b.parameters["flux"]
    with pytest.raises(ValueError):
        px.make_photons("mem:bad", src, 0.0, 1.0, 1.0, model3)     # redshift <= 0 without dist
    n3 = px.make_photons("mem:near", src, 0.0, 5.0e4, 1.0e5, model3, dist=(200.0, "Mpc"))
    assert n3[0] > 0


def test_reference_style_host_model_drives_make_photons(cuda):
    """A third-party (host) source model that only honours the reference plugin contract can be
    driven by pyxsim_b200.make_photons."""
    src, sm, _ = _cluster(8, nbins=100)

    class HostModel:
        ftype = "gas"

        def setup_model(self, mode, data_source, redshift):
            pass

        def set_pv(self, p_fields, v_fields, **kw):
            self.p_fields, self.v_fields = p_fields, v_fields
            for k, v in kw.items():
                setattr(self, k, v)

        def process_data(self, mode, chunk, spectral_norm):
            n = chunk.size
            nph = np.zeros(n, dtype=np.int64)
            nph[::7] = 3
            act = nph > 0
            return int(act.sum()), nph[act], act, np.full(int(nph.sum()), 1.5)

        def cleanup_model(self, mode):
            pass

    n_ph, n_c = px.make_photons("mem:host", src, 0.05, 1.0, 1.0, HostModel())
    lst = px.load_list("mem:host")
    assert n_ph == 3 * n_c and np.all(lst.numpy("energy") == 1.5)
    assert np.array_equal(lst.numpy("x"), src.fields[("index", "x")].cpu().numpy()[::7])


@pytest.mark.parametrize("n,photons", [(128, 2.0e8)])
def test_large_properties(cuda, n, photons):
    """Size-independent properties at a size the oracle cannot reach: counts add up, offsets are
    consistent, energies stay inside the table range, no-shift/no-absorb projection is the
    identity on energies, reruns are bit-identical, Doppler-only events are a permutation-free
    rescaling of the photons."""
    src, sm, (kT_nodes, ebins, cosmic, metal) = _cluster(n, nbins=4000, device="cuda")
    em_sum = float(src.fields[("gas", "emission_measure")].sum())
    ss = float(np.interp(6.0, kT_nodes, cosmic.sum(1) + 0.3 * metal.sum(1)))
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, 0.05)
    at = photons / (em_sum * ss / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * 1.05 ** 2))
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, 0.3, temperature_field=TF, prng=77)
    n_ph, n_c = px.make_photons("mem:big", src, 0.05, at / 1e5, 1e5, model)
    assert abs(n_ph - photons) < 6 * np.sqrt(photons)
    lst = px.load_list("mem:big")
    nph, e = lst.data["num_photons"], lst.data["energy"]
    assert int(nph.sum()) == n_ph == e.numel() and int(nph.min()) >= 1 and nph.numel() == n_c
    assert float(e.min()) >= ebins[0] and float(e.max()) <= ebins[-1]
    checksum = float(e.sum())
    n_all = px.project_photons("mem:big", "mem:big_id", "x", [1.0, 2.0], no_shifting=True, prng=5)
    ev = px.load_list("mem:big_id")
    assert n_all == n_ph and torch.equal(ev.data["eobs"], e)
    n1 = px.project_photons("mem:big", "mem:big_a", "z", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=9)
    a = {k: v.clone() for k, v in px.load_list("mem:big_a").data.items()}
    n2 = px.project_photons("mem:big", "mem:big_a", "z", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=9)
    b = px.load_list("mem:big_a").data
    assert n1 == n2 and all(torch.equal(a[k], b[k]) for k in a)
    assert 0.5 * n_ph < n1 < n_ph
    # sky positions stay inside the projected box (+ half a cell), flux bookkeeping is exact
    half = (1000.0 + 2000.0 / n) / (D_A * 1e3) * 180 / np.pi
    assert float((a["ysky"] - 45.0).abs().max()) <= half * 1.001
    model_b = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, 0.3, temperature_field=TF, prng=77)
    px.make_photons("mem:big2", src, 0.05, at / 1e5, 1e5, model_b)
    assert float(px.load_list("mem:big2").data["energy"].sum()) == checksum
    for k in ("mem:big", "mem:big2", "mem:big_id", "mem:big_a"):
        px.drop_list(k)


def test_more_than_2_31_photons(cuda):
    """Offsets, tile indices and output ranks are 64-bit throughout: a photon list with more than
    2^31 photons ("int64 offsets mandatory") generates and projects with
    consistent bookkeeping, stays bit-reproducible, and detects the same fraction as a list
    twenty times smaller."""
    free, _ = torch.cuda.mem_get_info()
    if free < 110 * 2 ** 30:
        pytest.skip("needs ~100 GB of free HBM")
    n = 128
    src, sm, (kT_nodes, ebins, cosmic, metal) = _cluster(n, nbins=4000, device="cuda")
    em_sum = float(src.fields[("gas", "emission_measure")].sum())
    ss = float(np.interp(6.0, kT_nodes, cosmic.sum(1) + 0.3 * metal.sum(1)))
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, 0.05)
    per_photon = em_sum * ss / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * 1.05 ** 2)
    frac = {}
    for tag, photons in (("small", 1.2e8), ("big", 2.4e9)):
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, 0.3, temperature_field=TF, prng=41)
        n_ph, n_c = px.make_photons("mem:i64", src, 0.05, photons / per_photon / 1e5, 1e5, model)
        assert abs(n_ph - photons) < 6 * np.sqrt(photons)
        lst = px.load_list("mem:i64")
        nph, e = lst.data["num_photons"], lst.data["energy"]
        assert e.numel() == n_ph and int(nph.sum()) == n_ph and nph.numel() == n_c
        if tag == "big":
            assert n_ph > 2 ** 31
        # both ends of the energy array were written (the last tiles are the ones whose indices
        # overflow 32 bits)
        assert float(e[:1000].min()) >= ebins[0] and float(e[-1000:].min()) >= ebins[0]
        assert float(e[-1000:].max()) <= ebins[-1]
        assert float(e.min()) >= ebins[0] and float(e.max()) <= ebins[-1]
        n1 = px.project_photons("mem:i64", "mem:i64_ev", "z", [30.0, 45.0], absorb_model="wabs", nH=0.3, prng=9)
        ev = px.load_list("mem:i64_ev")
        assert ev.data["eobs"].numel() == n1
        ck = (float(ev.data["eobs"].sum()), float(ev.data["xsky"].sum()), float(ev.data["ysky"][-1]))
        n2 = px.project_photons("mem:i64", "mem:i64_ev", "z", [30.0, 45.0], absorb_model="wabs", nH=0.3, prng=9)
        ev2 = px.load_list("mem:i64_ev")
        assert n1 == n2 and ck == (float(ev2.data["eobs"].sum()), float(ev2.data["xsky"].sum()),
                                    float(ev2.data["ysky"][-1]))
        assert float(ev2.data["eobs"].min()) == ev2.parameters["emin"]
        frac[tag] = (n1 / n_ph, n_ph)
        px.drop_list("mem:i64")
        px.drop_list("mem:i64_ev")
        del lst, ev, ev2, nph, e
        torch.cuda.empty_cache()
    (f_s, n_s), (f_b, n_b) = frac["small"], frac["big"]
    assert abs(f_s - f_b) < 6 * np.sqrt(f_s * (1 - f_s) * (1 / n_s + 1 / n_b))


def test_sph_particles_igm_model_end_to_end(cuda, oracle):
    """configs[3]-shaped run at test size: SPH-like particles, density-dependent (kT, nH) table
    with the CIE fallback, log energy binning, SPH-kernel position smoothing."""
    n = 40000
    f = synthetic.particle_beta_model_fields(n, device="cuda", seed=35)
    Tv, Dv, ebins, c2, m2, v2, cie = synthetic.pion_like_tables(nT=12, nD=7, nbins=600)
    cie_T, cc, cm, cv = cie
    cie_prod = px.ThermalSpectralModel(ebins, cie_T, cc, cm, cv, logT=True, binscale="log")
    prod = px.PionSpectralModel(ebins, Tv, Dv, c2, m2, v2, cie_model=cie_prod, binscale="log")
    src = px.ArrayDataSource(f, data_type="particles", domain_left_edge=[-1000.0] * 3,
                             domain_right_edge=[1000.0] * 3)
    hf = {k: v.cpu().numpy() for k, v in f.items()}
    ocie = oracle.TableModel(ebins, cie_T, cc, cm, cv, logT=True, binscale="log")
    orc = oracle.PionModel(ebins, Tv, Dv, c2, m2, v2, ocie, binscale="log")
    area, texp, z = 2.0e3, 5.0e5, 0.03
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, z)
    norm = area * texp / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * (1 + z) ** 2)
    cfg = dict(kT_min=0.00431, kT_max=64.0, Zmet=0.3, spectral_norm=norm)
    a = dict(kT=hf[TF], em=hf[("gas", "emission_measure")], nH=hf[("gas", "H_nuclei_density")])
    cut, tot, lam = oracle.thermal_spectra(orc, cfg, a)
    boost = 6.0e4 / lam.sum()                     # scale the emission measure to ~6e4 photons
    f[("gas", "emission_measure")] = f[("gas", "emission_measure")] * boost
    src.fields[("gas", "emission_measure")] = f[("gas", "emission_measure")]
    a["em"] = a["em"] * boost
    lam = lam * boost

    def run(seed_mine, seed_ref):
        model = px.IGMSourceModel(0.3, 1.4, 600, 0.3, binscale="log", temperature_field=TF,
                                  spectral_model=prod, prng=seed_mine)
        n_ph, n_c = px.make_photons("mem:sph_ph", src, z, area, texp, model)
        lst = px.load_list("mem:sph_ph")
        assert str(lst.parameters["data_type"]) == "particles"
        assert abs(n_ph - lam.sum()) < 5 * np.sqrt(lam.sum()) and n_ph > 20000
        nc, nph, idx, e = oracle.thermal_process_data(orc, cfg, a, np.random.RandomState(seed_ref))
        out = {"energy": stats.ks_2samp(lst.numpy("energy"), e).pvalue}
        for kern in ("gaussian", "top_hat"):
            n_ev = px.project_photons("mem:sph_ph", "mem:sph_ev", [0.3, 0.1, 0.9], [45.0, 30.0],
                                      absorb_model="wabs", nH=0.01, kernel=kern, prng=seed_mine + 3)
            ev = px.load_list("mem:sph_ev")
            g = oracle.gather_active([hf[("gas", f"particle_position_{ax}")] for ax in "xyz"],
                                     [hf[("gas", f"particle_velocity_{ax}")] for ax in "xyz"],
                                     hf[("gas", "smoothing_length")], idx, [-1000.0] * 3, [1000.0] * 3,
                                     [2000.0] * 3, [0.0] * 3, [0.0] * 3, (False,) * 3)
            d = dict(g, num_photons=nph, energy=e)
            params = dict(data_type="particles", observer="external", fid_d_a=D_A, fid_redshift=z)
            ref = oracle.project_photons(d, params, [0.3, 0.1, 0.9], [45.0, 30.0],
                                         np.random.RandomState(seed_ref + 3),
                                         absorb_model=oracle.WabsAbsorb(0.01), nH=0.01, kernel=kern)
            for k in ("xsky", "ysky", "eobs"):
                out[f"{kern}:{k}"] = stats.ks_2samp(ev.numpy(k), ref[k]).pvalue
        return out

    H.assert_distributions_agree(run, seeds=((5, 6), (15, 16)))


def test_powerlaw_plus_line_with_redshift_end_to_end(cuda, oracle):
    """configs[4]-shaped run at test size: power-law (alpha field incl. exactly 1 and 2) and a
    line source on a grid at z = 0.5 through make_photons/project_photons."""
    n = 24
    f = synthetic.beta_model_fields(n, v3d_sigma_kms=300.0, device="cuda")
    N = n ** 3
    rng = np.random.RandomState(4)
    alpha = rng.uniform(1.0, 2.5, N)
    alpha[::5] = 1.0
    alpha[1::5] = 2.0
    em = f[("gas", "emission_measure")].cpu().numpy()
    f[("gas", "alpha")] = torch.from_numpy(alpha).cuda()
    z, area, texp = 0.5, 2.0e3, 1.0e5
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, z)
    norm = area * texp / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * (1 + z) ** 2)
    pcfg = dict(e0=1.0, emin=0.5, emax=40.0, spectral_norm=norm, scale_factor=1 / (1 + z))
    lcfg = dict(e0=3.5, spectral_norm=norm, scale_factor=1 / (1 + z))
    lum0, _, _ = oracle.powerlaw_lambda(pcfg, em, alpha)
    lum = em * (5.0e4 / lum0.sum())                                # keV/s, ~5e4 photons
    rate = em * (5.0e4 / (em.sum() * norm / (1 + z)))              # photons/s, ~5e4 photons
    f[("gas", "xray_lum")] = torch.from_numpy(lum).cuda()
    f[("gas", "line_emission")] = torch.from_numpy(rate).cuda()
    src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3,
                             units={("gas", "xray_lum"): "keV/s"})

    def run(seed_mine, seed_ref):
        out = {}
        pl = px.PowerLawSourceModel(1.0, 0.5, 40.0, ("gas", "xray_lum"), ("gas", "alpha"), prng=seed_mine)
        n_ph, _ = px.make_photons("mem:pl", src, z, area, texp, pl)
        rlam, _, _ = oracle.powerlaw_lambda(pcfg, lum, alpha)
        assert abs(n_ph - rlam.sum()) < 5 * np.sqrt(rlam.sum()) and n_ph > 10000
        _, _, _, oe = oracle.powerlaw_process_data(pcfg, lum, alpha, np.random.RandomState(seed_ref))
        out["pl"] = stats.ks_2samp(px.load_list("mem:pl").numpy("energy"), oe).pvalue
        ln = px.LineSourceModel(3.5, ("gas", "line_emission"), (1000.0, "km/s"), prng=seed_mine + 1)
        n_l, _ = px.make_photons("mem:ln", src, z, area, texp, ln)
        _, _, _, le, F = oracle.line_process_data(lcfg, rate, 1000.0 * 3.5 / 299792.458,
                                                  np.random.RandomState(seed_ref + 1))
        assert abs(n_l - F.sum()) < 5 * np.sqrt(F.sum()) and n_l > 10000
        out["line"] = stats.ks_2samp(px.load_list("mem:ln").numpy("energy"), le).pvalue
        n_ev = px.project_photons("mem:pl", "mem:pl_ev", "y", [10.0, 10.0], absorb_model="wabs",
                                  nH=0.05, prng=seed_mine + 2)
        assert 0 < n_ev <= n_ph
        assert float(px.load_list("mem:pl_ev").parameters["redshift"]) == z
        return out

    H.assert_distributions_agree(run, seeds=((21, 22), (31, 32)))


def test_internal_observer_end_to_end(cuda, oracle):
    """observer="internal": photons weighted by 1/r^2 from the chosen centre, redshift forced to
    zero, all-sky projection (project_photons_allsky)."""
    src, sm, (kT_nodes, ebins, cosmic, metal) = _cluster(24, nbins=500)
    c = np.array([700.0, -650.0, 800.0])          # observer well off-centre, inside the box
    f = {k: v.cpu().numpy() for k, v in src.fields.items()}
    pos = [f[("index", ax)] for ax in "xyz"]
    r2 = oracle.compute_radius(pos, [-1000.0] * 3, [1000.0] * 3, [2000.0] * 3, c, (False,) * 3)
    texp = 1.0e5
    a = dict(kT=f[TF], em=f[("gas", "emission_measure")], r2=r2)
    om = oracle.TableModel(ebins, kT_nodes, cosmic, metal)
    cfg1 = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=texp / (4 * np.pi))
    area = 3.0e4 / oracle.thermal_spectra(om, cfg1, a)[2].sum()        # ~3e4 photons
    norm = area * texp / (4 * np.pi)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=norm)
    cut, tot, lam = oracle.thermal_spectra(om, cfg, a)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 500, 0.3, temperature_field=TF, prng=8)
    n_ph, n_c = px.make_photons("mem:int_ph", src, 0.3, area, texp, model, observer="internal", center=c)
    assert lam.sum() > 2.0e4 and abs(n_ph - lam.sum()) < 5 * np.sqrt(lam.sum())
    lst = px.load_list("mem:int_ph")
    assert float(lst.parameters["fid_redshift"]) == 0.0 and str(lst.parameters["observer"]) == "internal"
    idx_x = lst.numpy("x")
    assert np.allclose(idx_x.min(), pos[0].min() - c[0]) or idx_x.min() > pos[0].min() - c[0]
    with pytest.raises(RuntimeError):
        px.project_photons("mem:int_ph", "mem:int_ev", "z", [0.0, 0.0])
    n_ev = px.project_photons_allsky("mem:int_ph", "mem:int_ev", [0.0, 0.0, 1.0], save_los=True, prng=2)
    ev = px.load_list("mem:int_ev")
    assert n_ev == n_ph
    d = {k: lst.numpy(k) for k in ("x", "y", "z", "vx", "vy", "vz", "dx", "num_photons", "energy")}
    params = dict(data_type="cells", observer="internal", fid_d_a=0.0, fid_redshift=0.0)
    ref = oracle.project_photons(d, params, [0.0, 0.0, 1.0], [0.0, 0.0], np.random.RandomState(3), save_los=True)
    for k in ("xsky", "ysky", "los"):
        assert stats.ks_2samp(ev.numpy(k), ref[k]).pvalue > 0.01, k
    np.testing.assert_allclose(np.sort(ev.numpy("eobs")), np.sort(ref["eobs"]), rtol=1e-14)


def test_make_photons_options(cuda):
    """point_sources, fields_to_keep, bulk_velocity, explicit velocity_fields, center."""
    src, sm, _ = _cluster(12, nbins=200)
    src.fields[("gas", "tracer")] = torch.arange(12 ** 3, dtype=torch.float64)
    src.fields[("gas", "wind_x")] = torch.full((12 ** 3,), 7.0, dtype=torch.float64)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 200, 0.3, temperature_field=TF, prng=3)
    vf = [("gas", "wind_x"), ("gas", "velocity_y"), ("gas", "velocity_z")]
    n_ph, n_c = px.make_photons("mem:opt", src, 0.05, 3.0e4, 1.0e5, model, point_sources=True,
                                fields_to_keep=[("gas", "tracer")], bulk_velocity=[1.0, 2.0, 3.0],
                                velocity_fields=vf, center=[10.0, 20.0, 30.0])
    lst = px.load_list("mem:opt")
    idx = lst.numpy("tracer").astype(int)
    assert n_c == idx.size and np.all(np.diff(idx) > 0)
    assert np.all(lst.numpy("dx") == 0.0)                                   # point sources
    assert np.all(lst.numpy("vx") == 6.0)                                   # wind_x - bulk
    x = src.fields[("index", "x")].cpu().numpy()
    assert np.array_equal(lst.numpy("x"), x[idx] - 10.0)
    vy = src.fields[("gas", "velocity_y")].cpu().numpy()
    assert np.array_equal(lst.numpy("vy"), vy[idx] - 2.0)
    # point sources project onto their cell centres exactly (no scatter width)
    n_ev = px.project_photons("mem:opt", "mem:opt_ev", "z", [0.0, 0.0], phys_coord=True, no_shifting=True)
    ev = px.load_list("mem:opt_ev")
    nph = lst.numpy("num_photons")
    assert n_ev == n_ph and np.array_equal(ev.numpy("xsky"), np.repeat(lst.numpy("x"), nph))


def test_specialised_kernels_are_bit_identical_to_the_generic_ones(cuda, monkeypatch):
    """The sampler, detect and emit kernels compiled for the common option set (plain 1-D table,
    linear bins, wabs foreground column, cells + TAN sky) must reproduce the option-generic kernels
    bit for bit: same photons, same events, same flux -- on- and off-axis."""
    src, sm, (kT_nodes, ebins, cosmic, metal) = _cluster(48, nbins=1000, device="cuda")
    em_sum = float(src.fields[("gas", "emission_measure")].sum())
    ss = float(np.interp(6.0, kT_nodes, cosmic.sum(1) + 0.3 * metal.sum(1)))
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, 0.05)
    area = 3.0e6 / (em_sum * ss / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * 1.05 ** 2)) / 1.0e5

    def run(tag):
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 1000, 0.3, temperature_field=TF, prng=5)
        n_ph, n_c = px.make_photons(f"mem:spec_{tag}", src, 0.05, area, 1.0e5, model)
        assert n_ph < 1.0e7
        out = {"n_ph": n_ph, "energy": px.load_list(f"mem:spec_{tag}").data["energy"].clone()}
        for nm, normal in (("on", "z"), ("off", [0.3, -0.5, 0.8])):
            n_ev = px.project_photons(f"mem:spec_{tag}", f"mem:spec_{tag}_ev", normal, [30.0, 45.0],
                                      absorb_model="wabs", nH=0.05, save_los=True, prng=6)
            ev = px.load_list(f"mem:spec_{tag}_ev")
            out[nm] = (n_ev, {k: v.clone() for k, v in ev.data.items()}, ev.parameters["flux"],
                       ev.parameters["emin"], ev.parameters["emax"])
        return out

    a = run("fast")
    monkeypatch.setenv("PXB_GENERIC_KERNELS", "1")
    b = run("generic")
    assert a["n_ph"] == b["n_ph"] > 1e6 and torch.equal(a["energy"], b["energy"])
    for nm in ("on", "off"):
        assert a[nm][0] == b[nm][0] > 0 and a[nm][2:] == b[nm][2:]
        for k in a[nm][1]:
            assert torch.equal(a[nm][1][k], b[nm][1][k]), (nm, k)


# ---------------------------------------------------------------------------------------
# fused make_events == make_photons + project_photons
# ---------------------------------------------------------------------------------------
def _two_stage(tag, src, z, area, texp, model, normal, center, **kw):
    n_ph, n_c = px.make_photons(f"mem:{tag}_ph", src, z, area, texp, model)
    n_ev = px.project_photons(f"mem:{tag}_ph", f"mem:{tag}_ev2", normal, center, **kw)
    return (n_ph, n_c, n_ev), px.load_list(f"mem:{tag}_ev2")


def _same_events(a, b, keys=("xsky", "ysky", "eobs"), flux_rtol=0.0):
    for k in keys:
        assert np.array_equal(a.numpy(k), b.numpy(k)), k
    for k in ("emin", "emax"):
        assert float(a.parameters[k]) == float(b.parameters[k]), k
    # the flux is a sum over per-tile partials: identical when both runs tile the photons alike
    assert abs(float(a.parameters["flux"]) - float(b.parameters["flux"])) <= flux_rtol * float(b.parameters["flux"])


@pytest.mark.parametrize("case", ["common_on", "common_off", "generic_opts", "var_elem", "log_ar", "chunks"])
def test_fused_make_events_equals_the_two_stages(cuda, case, monkeypatch):
    """``make_events`` (sample -> shift -> absorb -> scatter in one kernel, no photon list) must
    give the event list of ``make_photons`` + ``project_photons`` bit for bit: same survivors, same
    order, same sky positions, same flux -- for the kernels compiled for the common option set and
    for the option-generic ones."""
    nbins, nvar, binscale, method = 800, 0, "linear", "invert_cdf"
    kw = dict(absorb_model="wabs", nH=0.05, prng=9)
    normal = "z"
    src_kw = dict(kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda")
    if case == "common_off":
        normal = [0.3, -0.5, 0.8]
    elif case == "generic_opts":
        en, sg = synthetic.tbabs_like_sigma()
        kw = dict(absorb_model=px.TBabsModel(0.03, energy=en, cross_section=sg), nH=0.03, prng=9,
                  sigma_pos=0.4, flat_sky=True, save_los=True, north_vector=[0.0, 1.0, 0.2])
        normal = [0.1, 0.9, -0.3]
    elif case == "var_elem":
        nvar = 3
        src_kw["metallicity"] = (0.1, 0.6)
    elif case == "log_ar":
        binscale, method = "log", "accept_reject"
    kT_nodes, ebins, cosmic, metal, var = synthetic.apec_like_tables(nT=36, nbins=nbins, nvar=nvar,
                                                                      binscale=binscale)
    src = synthetic.beta_model_source(40, chunk_size=(17000 if case == "chunks" else None), **src_kw)
    names = tuple(f"El{k}" for k in range(nvar))
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal, var, var_elem_names=names, binscale=binscale)

    def model():
        ve = {nm: 0.2 + 0.3 * k for k, nm in enumerate(names)} or None
        zm = ("gas", "metallicity") if case == "var_elem" else 0.3
        return px.ThermalSourceModel(sm, 0.1, 10.0, nbins, zm, binscale=binscale, method=method,
                                     temperature_field=TF, var_elem=ve, prng=8)

    area, texp, z = 6.0e4, 1.0e5, 0.05
    tot2, ev2 = _two_stage(case, src, z, area, texp, model(), normal, [30.0, 45.0], **kw)
    from pyxsim_b200 import _lib
    if case in ("var_elem", "log_ar"):
        # option-generic sampler: make_events takes the two stages by itself (they are faster) ...
        tot0 = px.make_events(f"mem:{case}_ev0", src, z, area, texp, model(), normal, [30.0, 45.0], **kw)
        assert tot0 == tot2 and _lib.last_pipeline_launch()[0] == 1
        _same_events(px.load_list(f"mem:{case}_ev0"), ev2)
        # ... and the fused kernel, asked for, still gives the same list
        monkeypatch.setenv("PXB_FUSE_GENERIC", "1")
    tot1 = px.make_events(f"mem:{case}_ev1", src, z, area, texp, model(), normal, [30.0, 45.0], **kw)
    assert tot1 == tot2 and tot1[2] > 1.0e5 and _lib.last_pipeline_launch()[0] == 2
    ev1 = px.load_list(f"mem:{case}_ev1")
    # (with several chunks the fused path sums the flux chunk by chunk, the two-stage path over the
    # concatenated list: same events, the sum differs in the last bits)
    _same_events(ev1, ev2, ("xsky", "ysky", "eobs") + (("los",) if kw.get("save_los") else ()),
                 flux_rtol=1e-13 if case == "chunks" else 0.0)
    assert ev1.parameters["absoption_model"] == ev2.parameters["absoption_model"]
    if case == "common_on":
        # the option-generic fused kernel gives the same list as well
        monkeypatch.setenv("PXB_GENERIC_KERNELS", "1")
        tot3 = px.make_events("mem:common_on_ev3", src, z, area, texp, model(), normal, [30.0, 45.0], **kw)
        assert tot3 == tot1
        _same_events(px.load_list("mem:common_on_ev3"), ev2)
        # ... and so does the variant with one big CTA per SM and the table rows staged in shared memory
        monkeypatch.delenv("PXB_GENERIC_KERNELS")
        monkeypatch.setenv("PXB_FUSED_BIG_CTA", "1")
        tot4 = px.make_events("mem:common_on_ev4", src, z, area, texp, model(), normal, [30.0, 45.0], **kw)
        assert tot4 == tot1 and _lib.last_pipeline_launch() == (2, 1, 1)
        _same_events(px.load_list("mem:common_on_ev4"), ev2)


def test_make_events_falls_back_to_two_stages(cuda):
    """Sources the fused kernel does not take -- tables with negative entries (the clip of
    base.py:460 matters), cells extrapolating beyond the table, power-law sources -- run the two
    stages behind the same call and give the two-stage result."""
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=300, negative=True)
    src = synthetic.beta_model_source(16, kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda")
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    kw = dict(absorb_model="wabs", nH=0.05, prng=9)

    def model():
        return px.ThermalSourceModel(sm, 0.1, 10.0, 300, 0.3, temperature_field=TF, prng=8)

    tot2, ev2 = _two_stage("fbneg", src, 0.05, 6.0e4, 1.0e5, model(), "z", [30.0, 45.0], **kw)
    tot1 = px.make_events("mem:fbneg_ev1", src, 0.05, 6.0e4, 1.0e5, model(), "z", [30.0, 45.0], **kw)
    assert tot1 == tot2 and tot1[2] > 0
    _same_events(px.load_list("mem:fbneg_ev1"), ev2)
    # cells hotter than the last table row are declined by the mixture sampler
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=20, nbins=300)
    src = synthetic.beta_model_source(16, kT_range=(0.5, 20.0), v3d_sigma_kms=300.0, device="cuda")
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    assert float(src.fields[TF].max()) > kT_nodes[-1]
    tot2, ev2 = _two_stage("fbext", src, 0.05, 6.0e4, 1.0e5, model(), "z", [30.0, 45.0], **kw)
    tot1 = px.make_events("mem:fbext_ev1", src, 0.05, 6.0e4, 1.0e5, model(), "z", [30.0, 45.0], **kw)
    assert tot1 == tot2 and tot1[2] > 0
    _same_events(px.load_list("mem:fbext_ev1"), ev2)
    pl = px.PowerLawSourceModel(1.0, 0.5, 8.0, ("gas", "emission_measure"), 1.7, prng=3)
    src.units[("gas", "emission_measure")] = "keV/s"
    tot = px.make_events("mem:fbpl_ev", src, 0.05, 1.0e-20, 1.0, pl, "z", [30.0, 45.0], prng=4)
    assert len(tot) == 3
