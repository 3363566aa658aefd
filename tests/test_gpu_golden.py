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
