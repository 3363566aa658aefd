"""GPU: replay the committed golden vectors (tests/golden/hotpath_golden.npz, produced from the
compiled reference natives) through libpxb.  Needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np
import pytest
import torch

import pyxsim_b200 as px
from pyxsim_b200 import _lib, synthetic
from pyxsim_b200.device import ptr, stream_ptr, to_device

import helpers as H

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.npz"))


def test_golden_interp_and_thermal(cuda):
    t = H.make_tables(nbins=64, nvar=2)
    sm = H.product_model(t)
    sm.prepare_spectrum(0.0)
    c, m, v = sm.get_spectrum(G["i1_kT"])
    assert np.array_equal(c, G["i1_c"]) and np.array_equal(m, G["i1_m"]) and np.array_equal(v, G["i1_v"])
    src = H.cells_source(G["i1_kT"], G["th_em"])
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 64, 0.3, temperature_field=("gas", "kT"),
                                  var_elem={"O": 0.5, "Fe": 1.5}, prng=1)
    model.setup_model("photons", src, 0.0)
    lam, _ = model.expected_counts(next(src.chunks()), 2e-44)
    cut = G["th_cut"]
    np.testing.assert_allclose(lam.cpu().numpy()[cut], G["th_lam"], rtol=1e-12)
    spec = model.cell_spectra(next(src.chunks())).cpu().numpy()
    np.testing.assert_allclose(spec[cut], G["th_tot"], rtol=1e-12, atol=1e-300)


def test_golden_doppler_tan_absorb_powerlaw(cuda):
    nph = G["dp_nph"]
    poff = np.concatenate([[0], np.cumsum(nph)]).astype(np.int64)
    e = to_device(G["dp_e"].copy())
    bn, b2 = to_device(G["dp_vn"] * (-1.0 / 299792.458)), to_device(G["dp_v2"] * (1.0 / 299792.458) ** 2)
    po = to_device(poff, torch.int64)
    _lib.call("pxb_doppler_shift", nph.size, ptr(bn), ptr(b2), ptr(po), e.numel(), ptr(e), stream_ptr())
    np.testing.assert_allclose(e.cpu().numpy(), G["dp_out"], rtol=4e-16, atol=0)
    x, y = to_device(G["tan_x"].copy()), to_device(G["tan_y"].copy())
    _lib.call("pxb_pixel_to_cel", x.numel(), ptr(x), ptr(y), 30.0, 45.0, stream_ptr())
    np.testing.assert_allclose(x.cpu().numpy(), G["tan_ra"], rtol=1e-13)
    np.testing.assert_allclose(y.cpu().numpy(), G["tan_dec"], rtol=1e-13)
    en, sg = synthetic.tbabs_like_sigma(n=500)
    np.testing.assert_allclose(px.TBabsModel(0.1, energy=en, cross_section=sg).get_absorb(G["ab_e"]),
                               G["ab_tab"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(px.WabsModel(0.1).get_absorb(G["ab_e"]), G["ab_wabs"], rtol=1e-12,
                               atol=1e-300)
    n = G["pl_lum"].size
    src = px.ArrayDataSource({("gas", "lum"): G["pl_lum"], ("gas", "alpha"): G["pl_alpha"],
                              ("index", "x"): np.zeros(n), ("index", "y"): np.zeros(n),
                              ("index", "z"): np.zeros(n)}, units={("gas", "lum"): "keV/s"})
    model = px.PowerLawSourceModel(1.0, 0.5, 40.0, ("gas", "lum"), ("gas", "alpha"), prng=1)
    model.setup_model("photons", src, 0.5)
    lam, _ = model.expected_counts(next(src.chunks()), 3e-44)
    np.testing.assert_allclose(lam.cpu().numpy(), G["pl_lam"], rtol=1e-12)


def _reference_inputs():
    import importlib.util

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("mrg", os.path.join(here, "golden", "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.inputs(), np.load(os.path.join(here, "golden", "reference_golden.npz"))


def test_gpu_against_reference_python_outputs(cuda):
    """Deterministic quantities against tests/golden/reference_golden.npz, i.e. against outputs
    of the reference's own Python code: Pion (kT, nH) spectra bit for bit, table transmission
    1e-12, Doppler-shifted energies and LOS of an unabsorbed projection, photon statistics of
    the reference's Poisson draws against the GPU's expected counts."""
    from pyxsim_b200.photon_list import DeviceList, save_list, load_list

    I, R = _reference_inputs()
    # --- Pion spectra ------------------------------------------------------------------
    a = I["pion"]
    cie_T, cc, cm, cv = a["cie"]
    cie = px.ThermalSpectralModel(a["ebins"], cie_T, cc, cm, cv, logT=True, binscale="log")
    pm = px.PionSpectralModel(a["ebins"], a["Tv"], a["Dv"], a["c2"], a["m2"], a["v2"], cie_model=cie,
                              binscale="log")
    pm.prepare_spectrum(0.0)
    c, m, v = pm.get_spectrum(a["kT"][:40], a["nH"][:40])
    assert np.array_equal(c, R["pion_c"]) and np.array_equal(m, R["pion_m"]) and np.array_equal(v, R["pion_v"])
    # --- expected counts vs the reference's Poisson draws ------------------------------------
    src = H.cells_source(a["kT"], a["em"], extra={("gas", "H_nuclei_density"): a["nH"]})
    model = px.PionSourceModel(0.3, 1.4, pm.nbins, 0.3, binscale="log", temperature_field=("gas", "kT"),
                               var_elem={"O": 0.4, "Ne": 1.1}, spectral_model=pm, prng=3, kT_min=1e-3)
    model.setup_model("photons", src, 0.0)
    lam, _ = model.expected_counts(next(src.chunks()), a["norm"])
    lam = lam.cpu().numpy()
    ref_counts = np.zeros(lam.size)
    ref_counts[R["pion_idx"]] = R["pion_nph"]
    chi2 = ((ref_counts - lam) ** 2 / np.maximum(lam, 1e-300))[lam > 5].sum()
    ndf = (lam > 5).sum()
    assert ndf > 30 and abs(chi2 - ndf) < 5 * np.sqrt(2 * ndf)
    assert abs(ref_counts.sum() - lam.sum()) < 5 * np.sqrt(lam.sum())
    # --- transmission -----------------------------------------------------------------------------
    ab = I["abs"]
    am = px.AbsorptionModel(0.15, ab["energy"], ab["sigma"])
    np.testing.assert_allclose(am.get_absorb(I["plist"]["energy"][:500]), R["abs_trans"], rtol=1e-12)
    # --- unabsorbed projections: Doppler energies (and LOS) are deterministic --------------------
    d = I["plist"]
    for name, obs, da, z in (("doppler_only", "external", 200.0, 0.05), ("doppler_internal", "internal", 0.0, 0.0)):
        params = dict(fid_exp_time=1.0e5, fid_area=1000.0, fid_redshift=z, fid_d_a=da, data_type="cells",
                      observer=obs)
        save_list("mem:refgold", DeviceList({}, params, {k: torch.from_numpy(v).cuda() for k, v in d.items()}))
        if obs == "external":
            n = px.project_photons("mem:refgold", "mem:refgold_ev", [0.3, -0.5, 0.8], [30.0, 45.0],
                                   save_los=True, prng=1)
        else:
            n = px.project_photons_allsky("mem:refgold", "mem:refgold_ev", [0.3, -0.5, 0.8], prng=1)
        ev = load_list("mem:refgold_ev")
        assert n == int(R[f"pj_{name}_n"]) == d["energy"].size
        np.testing.assert_allclose(ev.numpy("eobs"), R[f"pj_{name}_eobs"], rtol=4e-16, atol=0)
        np.testing.assert_allclose(ev.parameters["flux"], float(R[f"pj_{name}_flux"]), rtol=1e-13)
        assert ev.parameters["emin"] == float(R[f"pj_{name}_emin"])
        if obs == "external":
            np.testing.assert_allclose(ev.numpy("los"), R[f"pj_{name}_los"], rtol=1e-12, atol=1e-11)  # FMA vs mul+add
