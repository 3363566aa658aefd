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


def _replay_reference_cases(O, I, cases):
    """Recompute everything tests/golden/make_reference_golden.py stored, with the NumPy
    restatement (oracle/ref_glue.py) instead of the reference's own Python."""
    out = {}
    a = I["t1"]
    t = a["tables"]
    om = O.TableModel(t["ebins"], t["Tvals"], t["cosmic"], t["metal"], t["var"], binscale=t["binscale"])
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet="Z", Zconvert=a["Zconvert"], z_by_xh=True,
               h_fraction=a["h_fraction"], spectral_norm=a["norm"], max_density=3e-26,
               var_elem=[(0.5, 1.0, False), ("Fe", 1.0, False)])
    nc, nph, idx, e = O.thermal_process_data(
        om, cfg, dict(kT=a["kT"], em=a["em"], Z=a["Z"], Fe=a["Fe"], density=a["dens"]), np.random.RandomState(11))
    out.update(t1_nc=nc, t1_nph=nph, t1_idx=idx, t1_e=e)
    out.update(t1_spec=O.thermal_spectrum_mode(
        om, cfg, dict(kT=a["kT"], em=a["em"], Z=a["Z"], Fe=a["Fe"], density=a["dens"]), np.linspace(0.3, 7.0, 41)))

    cells = dict(kT=a["kT"], em=a["em"], Z=a["Z"], Fe=a["Fe"], density=a["dens"])
    for tag, energy, mode in (("lum", True, "luminosity"), ("rate", False, "photon_rate")):
        fl = om.make_fluxf(0.5, 2.0, energy=energy)
        out[f"t1_{tag}"] = O.thermal_field_mode(om, cfg, cells, mode, 0.5, 2.0, fluxf=fl)
        c_, m_, v_ = fl(np.array([0.05, 0.3, 1.0, 7.7, 40.0]))
        out.update({f"t1_{tag}_cf": c_, f"t1_{tag}_mf": m_, f"t1_{tag}_vf": v_})

    a = I["t2"]
    t = a["tables"]
    om = O.TableModel(t["ebins"], t["Tvals"], t["cosmic"], t["metal"], t["var"], binscale=t["binscale"])
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=a["norm"])
    for meth, seed in (("invert_cdf", 12), ("accept_reject", 13)):
        nc, nph, idx, e = O.thermal_process_data(om, cfg, dict(kT=a["kT"], em=a["em"]),
                                                 np.random.RandomState(seed), method=meth, literal_q3=True)
        out.update({f"t2_{meth}_nph": nph, f"t2_{meth}_idx": idx, f"t2_{meth}_e": e})

    a = I["pion"]
    cie_T, cc, cm, cv = a["cie"]
    ocie = O.TableModel(a["ebins"], cie_T, cc, cm, cv, logT=True, binscale="log")
    pm = O.PionModel(a["ebins"], a["Tv"], a["Dv"], a["c2"], a["m2"], a["v2"], ocie, binscale="log")
    c, m, v = pm.get_spectrum(a["kT"][:40], a["nH"][:40])
    out.update(pion_c=c, pion_m=m, pion_v=v)
    cfg = dict(kT_min=1e-3, kT_max=64.0, Zmet=0.3, spectral_norm=a["norm"],
               var_elem=[(0.4, 1.0, False), (1.1, 1.0, False)])
    nc, nph, idx, e = O.thermal_process_data(pm, cfg, dict(kT=a["kT"], em=a["em"], nH=a["nH"]),
                                             np.random.RandomState(14))
    out.update(pion_nph=nph, pion_idx=idx, pion_e=e)
    cfg["kT_min"] = 0.1
    for tag, energy, mode in (("lum", True, "luminosity"), ("rate", False, "photon_rate")):
        fl = pm.make_fluxf(0.4, 1.2, energy=energy)
        out[f"pion_{tag}"] = O.thermal_field_mode(pm, cfg, dict(kT=a["kT"], em=a["em"], nH=a["nH"]), mode,
                                                  0.4, 1.2, fluxf=fl)

    a = I["pl"]
    cfg = dict(e0=1.0, emin=0.5, emax=40.0, spectral_norm=a["norm"], scale_factor=1.0 / (1.0 + a["z"]))
    nc, nph, act, e = O.powerlaw_process_data(cfg, a["lum"], a["alpha"], np.random.RandomState(15))
    out.update(pl_nph=nph, pl_act=act, pl_e=e)
    nc, nph, act, e = O.powerlaw_process_data(cfg, a["lum"], 1.7, np.random.RandomState(16))
    out.update(plc_nph=nph, plc_act=act, plc_e=e)

    a = I["ln"]
    cfg = dict(e0=3.5, spectral_norm=a["norm"], scale_factor=1.0 / (1.0 + a["z"]))
    nc, nph, act, e, F = O.line_process_data(cfg, a["rate"], a["sig_kms"] * 3.5 / 299792.458,
                                             np.random.RandomState(17))
    out.update(ln_nph=nph, ln_act=act, ln_e=e)

    ab = I["abs"]
    am = O.TableAbsorb(0.15, ab["energy"], ab["sigma"])
    ee = I["plist"]["energy"][:500]
    out.update(abs_trans=am.get_absorb(ee), abs_det=am.absorb_photons(ee, np.random.RandomState(18)))
    d = I["plist"]
    for name, case in cases.items():
        ext = case["obs"] == "external"
        params = dict(data_type=case["data_type"], observer=case["obs"], fid_d_a=200.0 if ext else 0.0,
                      fid_redshift=0.05 if ext else 0.0)
        akw = dict(absorb_model=O.TableAbsorb(0.15, ab["energy"], ab["sigma"]), nH=0.15) \
            if case.get("absorb", True) else {}
        kw = dict(case["kw"])
        if kw.pop("column_file", None):
            col = I["col"]
            uv = O.orientation(case["normal"], kw.get("north_vector"))
            kw["nH_int"] = O.column_lookup(d["x"], d["y"], d["z"], uv, col["wbins"], col["dbins"], col["nH"])
        ref = O.project_photons({k: v.copy() for k, v in d.items()}, params, case["normal"], [30.0, 45.0],
                                np.random.RandomState(case["seed"]), **akw, **kw)
        out[f"pj_{name}_n"] = ref["n_events"]
        for k in ("xsky", "ysky", "eobs"):
            out[f"pj_{name}_{k}"] = ref[k]
        if case["kw"].get("save_los"):
            out[f"pj_{name}_los"] = ref["los"]
        flux = ref["flux_sum"]
        flux *= O.erg_per_keV / 1.0e5 / 1000.0
        out[f"pj_{name}_flux"], out[f"pj_{name}_emin"], out[f"pj_{name}_emax"] = flux, ref["emin"], ref["emax"]
    return out


def _load_make_reference_golden():
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "make_reference_golden", os.path.join(HERE, "golden", "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_restatement_matches_reference_python_bit_for_bit(oracle):
    """tests/golden/reference_golden.npz holds outputs of the reference's OWN Python
    (ThermalSourceModel.process_data, PionSpectralModel.get_spectrum, PowerLaw/Line
    process_data, AbsorptionModel.absorb_photons, _project_photons) run with the same
    RandomState seeds.  The NumPy restatement must reproduce every array exactly: photon
    counts, cell indices, energies, absorption masks, sky positions, fluxes."""
    mrg = _load_make_reference_golden()
    stored = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))
    got = _replay_reference_cases(oracle, mrg.inputs(), mrg.PROJ_CASES)
    assert sorted(got) == sorted(stored.files)
    for k in stored.files:
        a, b = np.asarray(got[k]), stored[k]
        assert a.shape == b.shape, k
        assert np.array_equal(a, b), (k, float(np.max(np.abs(a.astype(float) - b.astype(float)))))


def test_reference_golden_is_reproducible_from_the_reference_tree():
    """When the reference tree is present (build container) re-run its code and compare with
    the committed fixtures; skipped on the GPU box where /root/reference does not exist."""
    if not os.path.isdir("/root/reference/pyxsim"):
        pytest.skip("reference tree not available")
    mrg = _load_make_reference_golden()
    fresh = mrg.run_reference(mrg.inputs())
    stored = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))
    assert sorted(fresh) == sorted(stored.files)
    for k in stored.files:
        assert np.array_equal(np.asarray(fresh[k]), stored[k]), k


def test_list_files_interoperate_with_the_reference(oracle, monkeypatch):
    """The HDF5 branch of save_list/load_list (h5py is absent here, so it is driven through the
    in-memory h5py stand-in of oracle/ref_import.py) writes a photon list the reference's own
    _project_photons reads, and reads back the event list the reference writes: same /info,
    /parameters, /data layout (photon_list.py:345-399,605-646).  The reference's events equal the
    oracle's for the same RandomState."""
    if not os.path.isdir("/root/reference/pyxsim"):
        pytest.skip("reference tree not available")
    import types

    from oracle import ref_import as R
    from pyxsim_b200 import photon_list as PL

    M = R.modules()
    monkeypatch.setattr(PL, "h5py", types.SimpleNamespace(File=R.FakeH5File))
    real_exists = os.path.exists
    monkeypatch.setattr(PL.os.path, "exists", lambda p: p in R.FILES or real_exists(p))

    rng = np.random.RandomState(77)
    nc = 300
    nph = rng.poisson(15.0 * rng.lognormal(0, 1.0, nc)).astype(np.int64)
    d = dict(x=rng.uniform(-300, 300, nc), y=rng.uniform(-300, 300, nc), z=rng.uniform(-300, 300, nc),
             vx=rng.normal(0, 500, nc), vy=rng.normal(0, 500, nc), vz=rng.normal(100, 500, nc),
             dx=rng.uniform(2.0, 30.0, nc), num_photons=nph,
             energy=np.exp(rng.uniform(np.log(0.1), np.log(10.0), int(nph.sum()))))
    params = dict(fid_exp_time=1.0e5, fid_area=1000.0, fid_redshift=0.05, fid_d_a=200.0, hubble=0.71,
                  omega_matter=0.27, omega_lambda=0.73, data_type="cells", observer="external",
                  center=np.zeros(3), bulk_velocity=np.zeros(3),
                  velocity_fields=np.array(["('gas', 'velocity_x')", "('gas', 'velocity_y')", "('gas', 'velocity_z')"]))
    info = dict(pyxsim_version="0.1.0", yt_version="n/a", soxs_version="n/a", dataset="test", data_source="all",
                source_model="test", filenames=["interop_ph.h5"])
    fn = PL.save_list("interop_ph", PL.DeviceList(info, params, d))
    assert fn == "interop_ph.h5" and fn in R.FILES
    back = PL.load_list("interop_ph")
    assert back.parameters["data_type"] == "cells" and np.array_equal(back.numpy("energy"), d["energy"])

    # the reference projects OUR photon list ...
    n_ev = M["photon_list"]._project_photons("external", "interop_ph", "interop_ev", [0.3, -0.4, 0.85],
                                             [30.0, 45.0], prng=np.random.RandomState(5), save_los=True)
    # ... and we read ITS event list
    ev = PL.load_list("interop_ev")
    assert ev.numpy("eobs").size == n_ev == d["energy"].size
    assert str(ev.parameters["absoption_model"]) == "none" and str(ev.parameters["observer"]) == "external"
    ref = oracle.project_photons({k: v.copy() for k, v in d.items()},
                                 dict(data_type="cells", observer="external", fid_d_a=200.0, fid_redshift=0.05),
                                 [0.3, -0.4, 0.85], [30.0, 45.0], np.random.RandomState(5), save_los=True)
    for k in ("xsky", "ysky", "eobs", "los"):
        assert np.array_equal(ev.numpy(k), ref[k]), k
    # EventList / merge_files work on files the reference wrote
    from pyxsim_b200 import event_list as E

    monkeypatch.setattr(E, "load_list", PL.load_list)
    lst = PL.load_list("interop_ev")
    assert float(lst.parameters["exp_time"]) == 1.0e5 and np.allclose(lst.parameters["sky_center"], [30.0, 45.0])
