"""GPU parity: projection (Doppler, absorption, scatter, sky transform, compaction) and the
power-law / line generators against the oracle (reference Cython + restated glue)."""
import numpy as np
import pytest
import torch
from scipy import stats

import pyxsim_b200 as px
from pyxsim_b200 import _lib, synthetic
from pyxsim_b200.device import empty, ptr, stream_ptr, to_device
from pyxsim_b200.photon_list import DeviceList, save_list, load_list

import helpers as H

pytestmark = pytest.mark.gpu


def make_plist(n_cells=3000, mean_ph=40.0, data_type="cells", seed=13, observer="external",
               zero_cells=False, redshift=0.05):
    rng = np.random.RandomState(seed)
    nph = rng.poisson(mean_ph * rng.lognormal(0, 1.0, n_cells)).astype(np.int64)
    if not zero_cells:
        nph = np.maximum(nph, 1)
    d = dict(x=rng.uniform(-300, 300, n_cells), y=rng.uniform(-300, 300, n_cells),
             z=rng.uniform(-300, 300, n_cells), vx=rng.normal(0, 500, n_cells),
             vy=rng.normal(0, 500, n_cells), vz=rng.normal(100, 500, n_cells),
             dx=rng.uniform(2.0, 30.0, n_cells), num_photons=nph,
             energy=np.exp(rng.uniform(np.log(0.1), np.log(10.0), int(nph.sum()))))
    params = dict(fid_exp_time=1.0e5, fid_area=1000.0, fid_redshift=redshift, fid_d_a=200.0,
                  data_type=data_type, observer=observer)
    return d, params


def put(name, d, params):
    save_list("mem:" + name, DeviceList({}, params, {k: torch.from_numpy(v).cuda() for k, v in d.items()}))
    return "mem:" + name


def events(name):
    lst = load_list("mem:" + name)
    return {k: lst.numpy(k) for k in lst.data}, lst.parameters


# ---------------------------------------------------------------------------------------
# deterministic natives
# ---------------------------------------------------------------------------------------
def test_doppler_shift_matches_reference(cuda, oracle):
    d, _ = make_plist(zero_cells=True)
    sf = oracle.ref()["sky_functions"]
    bn = d["vz"] * oracle.scale_shift
    b2 = (d["vx"] ** 2 + d["vy"] ** 2 + d["vz"] ** 2) * oracle.scale_shift2
    e_ref = d["energy"].copy()
    sf.doppler_shift(bn, b2, d["num_photons"], e_ref)
    poff = np.concatenate([[0], np.cumsum(d["num_photons"])]).astype(np.int64)
    e = to_device(d["energy"].copy())
    dbn, db2, dpoff = to_device(bn), to_device(b2), to_device(poff, torch.int64)
    _lib.call("pxb_doppler_shift", d["num_photons"].size, ptr(dbn), ptr(db2), ptr(dpoff),
              e.numel(), ptr(e), stream_ptr())
    np.testing.assert_allclose(e.cpu().numpy(), e_ref, rtol=2e-16, atol=0)


def test_pixel_to_cel_known_answer(cuda, oracle):
    """pyxsim/tests/test_pixel_to_cel.py with the analytic TAN inverse as the known answer
    (astropy is not installed) and the reference Cython beside it."""
    prng = np.random.RandomState(24)
    n_evt = 100000
    rr = 100.0 * prng.uniform(size=n_evt)
    theta = 2.0 * np.pi * prng.uniform(size=n_evt)
    xx = rr * np.cos(theta) / 1.0e5
    yy = rr * np.sin(theta) / 1.0e5
    xs, ys = to_device(xx.copy()), to_device(yy.copy())
    _lib.call("pxb_pixel_to_cel", n_evt, ptr(xs), ptr(ys), 30.0, 45.0, stream_ptr())
    ra, dec = oracle.tan_known_answer(xx, yy, 30.0, 45.0)
    np.testing.assert_allclose(xs.cpu().numpy(), ra, rtol=1e-7)       # the reference test's bar
    np.testing.assert_allclose(ys.cpu().numpy(), dec, rtol=1e-7)
    x1, y1 = xx.copy(), yy.copy()
    oracle.ref()["sky_functions"].pixel_to_cel(x1, y1, np.array([30.0, 45.0]))
    np.testing.assert_allclose(xs.cpu().numpy(), x1, rtol=1e-12)
    np.testing.assert_allclose(ys.cpu().numpy(), y1, rtol=1e-12)


@pytest.mark.parametrize("kind", ["table_log", "table_lin", "table_irregular", "wabs"])
def test_absorb_transmission(cuda, oracle, kind):
    rng = np.random.RandomState(2)
    e = np.exp(rng.uniform(np.log(0.02), np.log(20.0), 50000))
    if kind == "wabs":
        prod, orc = px.WabsModel(0.07), oracle.WabsAbsorb(0.07)
    else:
        en, sg = synthetic.tbabs_like_sigma(n=3000)
        if kind == "table_lin":
            en2 = np.linspace(en[0], en[-1], 5000)
            sg = np.interp(en2, en, sg)
            en = en2
        elif kind == "table_irregular":
            keep = np.sort(rng.choice(en.size, 1500, replace=False))
            en, sg = en[keep], sg[keep]
        prod, orc = px.TBabsModel(0.07, energy=en, cross_section=sg), oracle.TableAbsorb(0.07, en, sg * 1.0e22)
    np.testing.assert_allclose(prod.get_absorb(e), orc.get_absorb(e), rtol=1e-12, atol=1e-300)


def test_absorb_decision_at_the_threshold(cuda, oracle):
    """The kernels decide u < exp(-sigma nH) in fp32 first and fall back to fp64 inside a
    rigorous error band: the decision must be the fp64 one for u arbitrarily close to the
    threshold and for energies on / one ulp (float or double) either side of every wabs edge."""
    rng = np.random.RandomState(4)
    edges = np.array([0.03, 0.1, 0.284, 0.4, 0.532, 0.707, 0.867, 1.303, 1.84, 2.471, 3.21, 4.038,
                      7.111, 8.331, 10.0])
    ef = edges.astype(np.float32)
    near = [edges, np.nextafter(edges, 0), np.nextafter(edges, 100),
            ef.astype(np.float64), np.nextafter(ef, np.float32(0)).astype(np.float64),
            np.nextafter(ef, np.float32(100)).astype(np.float64),
            edges * (1 - 2.0 ** -25), edges * (1 + 2.0 ** -25), edges * (1 - 2.0 ** -24), edges * (1 + 2.0 ** -24)]
    e = np.concatenate(near + [np.exp(rng.uniform(np.log(0.02), np.log(40.0), 200000))])
    for nH in (0.02, 0.5, 30.0):
        orc = oracle.WabsAbsorb(nH)
        prod = px.WabsModel(nH)
        p = orc.get_absorb(e)
        for delta in (0.5, 1e-2, 1e-4, 1e-5, 1e-6, 1e-7, 1e-9, 1e-12):
            for sign in (-1.0, 1.0):
                u = p * (1.0 + sign * delta)
                ok = (u < 1.0) & (u > 0.0) & (p > 1e-300)

                class Fixed:
                    def uniform(self, size=None):
                        return u

                got = prod.absorb_photons(e, prng=Fixed())
                np.testing.assert_array_equal(got[ok], (u < p)[ok], err_msg=f"nH={nH} delta={sign * delta}")
    # a table model takes the fp64 route throughout
    en, sg = synthetic.tbabs_like_sigma(n=3000)
    prod, orc = px.TBabsModel(0.3, energy=en, cross_section=sg), oracle.TableAbsorb(0.3, en, sg * 1.0e22)
    e = np.exp(rng.uniform(np.log(0.05), np.log(15.0), 100000))
    p = orc.get_absorb(e)
    for delta in (-1e-9, 1e-9, -0.3, 0.3):
        u = np.clip(p * (1 + delta), 0.0, 1.0 - 1e-16)

        class Fixed2:
            def uniform(self, size=None):
                return u

        np.testing.assert_array_equal(prod.absorb_photons(e, prng=Fixed2()), u < p)


def test_absorb_photons_follows_reference_draw_order(cuda, oracle):
    """Same RandomState, same energies -> the mask of the reference's absorb_photons."""
    rng = np.random.RandomState(8)
    e = np.exp(rng.uniform(np.log(0.1), np.log(10.0), 100000))
    got = px.WabsModel(0.1).absorb_photons(e, prng=np.random.RandomState(77))
    ref = np.random.RandomState(77).uniform(size=e.size) < oracle.WabsAbsorb(0.1).get_absorb(e)
    np.testing.assert_array_equal(got, ref)


# ---------------------------------------------------------------------------------------
# projection invariants and distributions
# ---------------------------------------------------------------------------------------
def test_no_shift_no_absorb_identity(cuda):
    """pyxsim/tests/test_sloshing.py:96-117: without shifting and absorption every photon
    becomes an event and the event energies ARE the photon energies (same order)."""
    d, p = make_plist(zero_cells=True)
    put("id_ph", d, p)
    n = px.project_photons("mem:id_ph", "mem:id_ev", "z", [30.0, 45.0], no_shifting=True, prng=3)
    ev, par = events("id_ev")
    assert n == d["energy"].size == ev["eobs"].size
    assert np.array_equal(ev["eobs"], d["energy"])
    assert par["emin"] == d["energy"].min() and par["emax"] == d["energy"].max()
    flux = d["energy"].sum() * 1.602176634e-9 / p["fid_exp_time"] / p["fid_area"]
    np.testing.assert_allclose(par["flux"], flux, rtol=1e-12)


@pytest.mark.parametrize("normal", ["x", "y", "z", [0.3, -0.5, 0.8]])
def test_doppler_in_projection(cuda, oracle, normal):
    d, p = make_plist()
    put("dop_ph", d, p)
    n = px.project_photons("mem:dop_ph", "mem:dop_ev", normal, [30.0, 45.0], prng=3)
    ev, _ = events("dop_ev")
    ref = oracle.project_photons(d, p, normal, [30.0, 45.0], np.random.RandomState(1))
    assert n == ref["n_events"] == d["energy"].size
    np.testing.assert_allclose(ev["eobs"], ref["eobs"], rtol=4e-16, atol=0)


def _ks_ok(a, b, pmin=0.01):
    return stats.ks_2samp(a, b).pvalue > pmin


@pytest.mark.parametrize("case", [
    dict(normal="z", data_type="cells"),
    dict(normal="x", data_type="cells", flat_sky=True),
    dict(normal=[1.0, -2.0, 5.0], data_type="cells"),
    dict(normal=[1.0, -2.0, 5.0], data_type="cells", north_vector=[0.0, 0.0, 1.0], sigma_pos=0.5),
    dict(normal="y", data_type="particles", kernel="top_hat"),
    dict(normal=[0.2, 0.9, -0.4], data_type="particles", kernel="gaussian"),
    dict(normal="z", data_type="cells", phys_coord=True),
])
def test_scatter_and_absorb_distributions(cuda, oracle, case):
    """Sky positions, line-of-sight and energies of the events against the oracle's: KS on
    every coordinate (both are random draws of the same distributions), absorbed fraction
    within binomial error, LOS coordinates exactly equal as multisets."""
    case = dict(case)
    normal = case.pop("normal")
    data_type = case.pop("data_type")
    d, p = make_plist(n_cells=2000, mean_ph=60.0, data_type=data_type)
    en, sg = synthetic.tbabs_like_sigma()
    put("sc_ph", d, p)
    ntot = d["energy"].size

    def run(seed_mine, seed_ref):
        n = px.project_photons("mem:sc_ph", "mem:sc_ev", normal, [30.0, 45.0],
                               absorb_model=px.TBabsModel(0.15, energy=en, cross_section=sg),
                               nH=0.15, save_los=True, prng=seed_mine, **case)
        ev, par = events("sc_ev")
        ref = oracle.project_photons(d, p, normal, [30.0, 45.0], np.random.RandomState(seed_ref),
                                     absorb_model=oracle.TableAbsorb(0.15, en, sg * 1.0e22), nH=0.15,
                                     save_los=True, **case)
        assert 0.2 * ntot < n < 0.98 * ntot
        pbar = 0.5 * (n + ref["n_events"]) / ntot
        assert abs(n - ref["n_events"]) < 5 * np.sqrt(2 * ntot * pbar * (1 - pbar))
        assert ev["eobs"].min() == par["emin"] and ev["eobs"].max() == par["emax"]
        np.testing.assert_allclose(par["flux"], ev["eobs"].sum() * 1.602176634e-9 / 1e5 / 1e3,
                                   rtol=1e-12)
        return {k: stats.ks_2samp(ev[k], ref[k]).pvalue for k in ("xsky", "ysky", "eobs", "los")}

    H.assert_distributions_agree(run)


@pytest.mark.parametrize("data_type,kernel", [("cells", "top_hat"), ("particles", "top_hat"),
                                               ("particles", "gaussian")])
def test_scatter_kernel_shape(cuda, oracle, data_type, kernel):
    """One cell, many photons, no absorption: the scatter offsets themselves (physical
    coordinates) must follow the reference kernel -- including r ~ U(0,1) for the top-hat."""
    n = 200000
    d = dict(x=np.array([10.0]), y=np.array([-20.0]), z=np.array([5.0]), vx=np.zeros(1),
             vy=np.zeros(1), vz=np.zeros(1), dx=np.array([8.0]),
             num_photons=np.array([n], dtype=np.int64), energy=np.full(n, 1.0))
    p = dict(fid_exp_time=1.0, fid_area=1.0, fid_redshift=0.0, fid_d_a=100.0, data_type=data_type,
             observer="external")
    put("k_ph", d, p)
    for normal in ("z", [0.6, 0.0, 0.8]):
        px.project_photons("mem:k_ph", "mem:k_ev", normal, [0.0, 0.0], kernel=kernel,
                           phys_coord=True, no_shifting=True, prng=8)
        ev, _ = events("k_ev")
        ref = oracle.project_photons(d, p, normal, [0.0, 0.0], np.random.RandomState(6),
                                     kernel=kernel, phys_coord=True, no_shifting=True)
        for k in ("xsky", "ysky"):
            assert _ks_ok(ev[k], ref[k]), (normal, k)
        r_mine = np.hypot(ev["xsky"] - np.median(ref["xsky"]), ev["ysky"] - np.median(ref["ysky"]))
        r_ref = np.hypot(ref["xsky"] - np.median(ref["xsky"]), ref["ysky"] - np.median(ref["ysky"]))
        assert _ks_ok(r_mine, r_ref), normal


@pytest.mark.parametrize("data_type,kernel", [("cells", "top_hat"), ("particles", "top_hat"),
                                               ("particles", "gaussian")])
def test_allsky_internal_observer(cuda, oracle, data_type, kernel):
    d, p = make_plist(n_cells=1500, mean_ph=50.0, data_type=data_type, observer="internal",
                      redshift=0.0)
    p["fid_d_a"] = 0.0
    put("as_ph", d, p)
    normal = [0.1, 0.3, 0.9]
    n = px.project_photons_allsky("mem:as_ph", "mem:as_ev", normal, absorb_model="wabs", nH=0.05,
                                  kernel=kernel, save_los=True, center_vector=[1.0, 0.0, 0.0],
                                  prng=4)
    ev, _ = events("as_ev")
    ref = oracle.project_photons(d, p, normal, [0.0, 0.0], np.random.RandomState(9),
                                 absorb_model=oracle.WabsAbsorb(0.05), nH=0.05, kernel=kernel,
                                 save_los=True, north_vector=[1.0, 0.0, 0.0])
    assert abs(n - ref["n_events"]) < 6 * np.sqrt(ref["n_events"])
    for k in ("xsky", "ysky", "eobs", "los"):
        assert _ks_ok(ev[k], ref[k]), k
    assert ev["xsky"].min() >= 0 and ev["xsky"].max() <= 360 and np.abs(ev["ysky"]).max() <= 90


def test_internal_absorption_per_cell(cuda, oracle):
    d, p = make_plist(n_cells=800, mean_ph=80.0)
    rng = np.random.RandomState(3)
    nhc = rng.uniform(0.0, 0.5, 800)
    en, sg = synthetic.tbabs_like_sigma()
    put("ia_ph", d, p)
    n = px.project_photons("mem:ia_ph", "mem:ia_ev", "z", [30.0, 45.0],
                           absorb_model=px.TBabsModel(0.02, energy=en, cross_section=sg), nH=0.02,
                           nH_cell=nhc, prng=12)
    ev, _ = events("ia_ev")
    ref = oracle.project_photons(d, p, "z", [30.0, 45.0], np.random.RandomState(2),
                                 absorb_model=oracle.TableAbsorb(0.02, en, sg * 1.0e22), nH=0.02, nH_int=nhc)
    assert abs(n - ref["n_events"]) < 6 * np.sqrt(ref["n_events"])
    assert _ks_ok(ev["eobs"], ref["eobs"])


def test_column_file_internal_absorption(cuda, oracle, tmp_path):
    """column_file= (photon_list.py:524-546,672-680,715-723): the nH cube is interpolated at the
    rotated cell centres -- scipy's interpn as the reference calls it, to 1e-13; cells
    outside the cube get 0 -- and each cell's photons are absorbed at their source-frame energy
    with that column.  Detected fractions per energy band against the oracle projection."""
    rng = np.random.RandomState(5)
    d, params = make_plist(n_cells=4000, mean_ph=60.0, seed=31)
    normal, north = [0.3, -0.4, 0.85], [0.1, 0.9, 0.2]
    L, north_out, uv = px.utils.get_normal_and_north(normal, north_vector=north)
    nw, nd = 24, 16
    wbins = np.linspace(-250.0, 250.0, nw + 1)       # smaller than the +-300 kpc box: some cells fall outside
    dbins = np.linspace(-280.0, 280.0, nd + 1)
    cube = 0.3 * rng.lognormal(0.0, 1.0, (nw, nw, nd))
    fn = px.write_column_file(str(tmp_path / "columns.h5"), L, north_out, wbins, dbins, cube)
    col = px.read_column_file(fn)
    np.testing.assert_array_equal(col["nH"], cube)
    # the lookup alone: exact nodes, cube faces and outside points included
    x, y, z = d["x"].copy(), d["y"].copy(), d["z"].copy()
    wmid = 0.5 * (wbins[1:] + wbins[:-1])
    inv = np.linalg.inv(uv)
    x[:3], y[:3], z[:3] = inv @ np.array([[wmid[0], wmid[5], wmid[-1]], [wmid[3], wmid[3], wmid[0]],
                                          [0.5 * (dbins[0] + dbins[1]), 0.0, 0.5 * (dbins[-1] + dbins[-2])]])
    got = px.column_density_at_cells(x, y, z, uv, wbins, dbins, cube).cpu().numpy()
    ref = oracle.column_lookup(x, y, z, uv, wbins, dbins, cube)
    assert (ref == 0.0).sum() > 50 and (ref > 0).sum() > 1000
    # (not bit-for-bit: the rotated position comes from a BLAS dot product in the reference)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0.0)
    assert np.array_equal(got == 0.0, ref == 0.0)
    # through project_photons
    put("colf", d, params)
    n_ev = px.project_photons("mem:colf", "mem:colf_ev", normal, [30.0, 45.0], absorb_model="wabs",
                              nH=0.01, north_vector=north, column_file=fn, prng=77)
    ev, _ = events("colf_ev")
    nH_int = oracle.column_lookup(d["x"], d["y"], d["z"], uv, wbins, dbins, cube)

    def run(seed_mine, seed_ref):
        px.project_photons("mem:colf", "mem:colf_ev", normal, [30.0, 45.0], absorb_model="wabs",
                           nH=0.01, north_vector=north, column_file=fn, prng=seed_mine)
        e1, _ = events("colf_ev")
        o = oracle.project_photons(d, params, normal, [30.0, 45.0], np.random.RandomState(seed_ref),
                                   absorb_model=oracle.WabsAbsorb(0.01), nH=0.01, north_vector=north,
                                   nH_int=nH_int)
        assert abs(e1["eobs"].size - o["eobs"].size) < 6 * np.sqrt(o["eobs"].size)
        return {"eobs": stats.ks_2samp(e1["eobs"], o["eobs"]).pvalue,
                "xsky": stats.ks_2samp(e1["xsky"], o["xsky"]).pvalue}

    H.assert_distributions_agree(run, seeds=((78, 79), (80, 81)))
    assert n_ev == ev["eobs"].size
    # error behaviour of the reference (photon_list.py:526-540)
    with pytest.raises(ValueError):
        px.project_photons("mem:colf", "mem:colf_ev", normal, [30.0, 45.0], column_file=fn)
    with pytest.raises(ValueError):
        px.project_photons("mem:colf", "mem:colf_ev", [0.0, 0.0, 1.0], [30.0, 45.0], absorb_model="wabs",
                           nH=0.01, column_file=fn)


def test_projection_bit_reproducible_and_stable_order(cuda):
    d, p = make_plist(n_cells=5000, mean_ph=30.0)
    put("rp_ph", d, p)
    runs = []
    for _ in range(3):
        px.project_photons("mem:rp_ph", "mem:rp_ev", [0.2, 0.3, 0.9], [10.0, -30.0],
                           absorb_model="wabs", nH=0.1, save_los=True, prng=99)
        ev, par = events("rp_ev")
        runs.append((ev, par))
    for ev, par in runs[1:]:
        for k in ev:
            assert np.array_equal(ev[k], runs[0][0][k]), k
        assert par["flux"] == runs[0][1]["flux"]
    # stable compaction: events are a subsequence of the Doppler-shifted photon list
    px.project_photons("mem:rp_ph", "mem:rp_all", [0.2, 0.3, 0.9], [10.0, -30.0], prng=99)
    allev, _ = events("rp_all")
    e_all, e_det = allev["eobs"], runs[0][0]["eobs"]
    ok = True
    idx = 0
    for v in e_det[:20000]:
        while idx < e_all.size and e_all[idx] != v:
            idx += 1
        if idx == e_all.size:
            ok = False
            break
        idx += 1
    assert ok


def test_photon_list_with_empty_cells_and_errors(cuda):
    d, p = make_plist(n_cells=4000, mean_ph=0.3, zero_cells=True)
    assert (d["num_photons"] == 0).sum() > 1000
    put("z_ph", d, p)
    n = px.project_photons("mem:z_ph", "mem:z_ev", "z", [0.0, 0.0], no_shifting=True, prng=1)
    assert n == d["energy"].size
    with pytest.raises(KeyError):
        px.project_photons("mem:z_ph", "mem:z_ev", "z", [0.0, 0.0], absorb_model="nope", nH=0.1)
    with pytest.raises(RuntimeError):
        px.project_photons_allsky("mem:z_ph", "mem:z_ev", [0, 0, 1.0])
    p2 = dict(p, data_type="particles")
    put("z2_ph", d, p2)
    with pytest.raises(RuntimeError):
        px.project_photons("mem:z2_ph", "mem:z_ev", "z", [0.0, 0.0], sigma_pos=1.0)


@pytest.mark.parametrize("field_kpc", [150.0, 3000.0, 40000.0])
def test_in_kernel_tan_transform_is_the_reference_pixel_to_cel(cuda, oracle, field_kpc):
    """Zero-width cells are not scattered, so the sky position of every event is the TAN
    transform of its cell centre: the projection kernels' rearranged inverse gnomonic projection
    (four-term series within 0.9 deg, seven-term within 3.6 deg, general formula beyond) against
    the reference's compiled pixel_to_cel on the same plane coordinates.  D_A = 200 Mpc: the three
    fields reach 0.04, 0.9 and 11 degrees."""
    rng = np.random.RandomState(23)
    n_cells = 20000
    nph = rng.randint(1, 4, n_cells).astype(np.int64)
    d = dict(x=rng.uniform(-field_kpc, field_kpc, n_cells), y=rng.uniform(-field_kpc, field_kpc, n_cells),
             z=rng.uniform(-300, 300, n_cells), vx=np.zeros(n_cells), vy=np.zeros(n_cells), vz=np.zeros(n_cells),
             dx=np.zeros(n_cells), num_photons=nph, energy=rng.uniform(0.5, 5.0, int(nph.sum())))
    params = dict(fid_exp_time=1.0e5, fid_area=1000.0, fid_redshift=0.05, fid_d_a=200.0,
                  data_type="cells", observer="external")
    put("tan_ph", d, params)
    for center in ([30.0, 45.0], [210.0, -72.5], [0.0, 0.0]):
        n = px.project_photons("mem:tan_ph", "mem:tan_ev", "z", center, no_shifting=True, prng=3)
        ev, _ = events("tan_ev")
        assert n == d["energy"].size
        xs = np.repeat(d["x"], nph) / 200.0e3           # on-axis "z": (xsky, ysky) <- (x, y) / D_A
        ys = np.repeat(d["y"], nph) / 200.0e3
        oracle.ref()["sky_functions"].pixel_to_cel(xs, ys, np.asarray(center, "float64"))
        dra = (ev["xsky"] - xs + 180.0) % 360.0 - 180.0
        assert np.max(np.abs(dra) * np.cos(np.deg2rad(ys))) < 2e-13 * max(1.0, np.abs(xs).max() / 50.0) + 1e-12
        np.testing.assert_allclose(ev["ysky"], ys, rtol=0, atol=3e-13)


def test_tiles_spanning_hundreds_of_empty_cells(cuda, oracle):
    """A photon list whose cells are mostly empty (long runs of zero-photon cells between the
    active ones): the pair tiles then hold empty cells (shared-memory bisection instead of the
    ballot look-up) or span more cells than are staged (global bisection).  Doppler-only projection
    is deterministic: every event must carry exactly its own cell's shift and line-of-sight, in
    photon order."""
    rng = np.random.RandomState(17)
    n_cells = 60000
    nph = np.zeros(n_cells, dtype=np.int64)
    active = np.sort(rng.choice(n_cells, 900, replace=False))
    nph[active] = rng.randint(1, 6, active.size)           # few photons per active cell
    nph[30000:31000] = 0                                    # a run of 1000 empty cells
    nph[31000] = 700                                        # then a bright one
    nph[active[:5]] = [300, 1, 257, 256, 2]
    d = dict(x=rng.uniform(-300, 300, n_cells), y=rng.uniform(-300, 300, n_cells),
             z=rng.uniform(-300, 300, n_cells), vx=rng.normal(0, 500, n_cells),
             vy=rng.normal(0, 500, n_cells), vz=rng.normal(100, 500, n_cells),
             dx=rng.uniform(2.0, 30.0, n_cells), num_photons=nph,
             energy=np.exp(rng.uniform(np.log(0.1), np.log(10.0), int(nph.sum()))))
    params = dict(fid_exp_time=1.0e5, fid_area=1000.0, fid_redshift=0.05, fid_d_a=200.0,
                  data_type="cells", observer="external")
    put("sparse_ph", d, params)
    normal = [0.3, -0.5, 0.8]
    n = px.project_photons("mem:sparse_ph", "mem:sparse_ev", normal, [30.0, 45.0], save_los=True, prng=3)
    ev, _ = events("sparse_ev")
    assert n == d["energy"].size
    ref = oracle.project_photons({k: v.copy() for k, v in d.items()}, params, normal, [30.0, 45.0],
                                 np.random.RandomState(4), save_los=True)
    np.testing.assert_allclose(ev["eobs"], ref["eobs"], rtol=4e-16, atol=0)      # per-cell Doppler factor
    np.testing.assert_allclose(ev["los"], ref["los"], rtol=1e-12, atol=1e-11)    # per-cell line of sight
    # with absorption the survivors still carry their own cell's data
    n2 = px.project_photons("mem:sparse_ph", "mem:sparse_ev", normal, [30.0, 45.0], absorb_model="wabs",
                            nH=0.3, save_los=True, prng=3)
    ev2, _ = events("sparse_ev")
    assert 0 < n2 < n
    pairs = {(e, l) for e, l in zip(np.round(ref["eobs"], 12), np.round(ref["los"], 7))}
    assert all((e, l) in pairs for e, l in zip(np.round(ev2["eobs"], 12), np.round(ev2["los"], 7)))


# ---------------------------------------------------------------------------------------
# power law / line
# ---------------------------------------------------------------------------------------
def test_powerlaw_counts_and_energies(cuda, oracle):
    n = 4000
    rng = np.random.RandomState(33)
    alpha = rng.uniform(0.5, 2.5, n)
    alpha[:200] = 1.0
    alpha[200:400] = 2.0
    lum = 1.0e45 * rng.lognormal(0, 1, n)
    src = px.ArrayDataSource({("gas", "lum"): lum, ("gas", "alpha"): alpha,
                              ("index", "x"): np.zeros(n), ("index", "y"): np.zeros(n),
                              ("index", "z"): np.zeros(n)}, units={("gas", "lum"): "keV/s"})
    model = px.PowerLawSourceModel(1.0, 0.5, 40.0, ("gas", "lum"), ("gas", "alpha"), prng=33)
    z = 0.5
    model.setup_model("photons", src, z)
    norm = 3.0e-44
    lam, counts = model.expected_counts(next(src.chunks()), norm)
    cfg = dict(e0=1.0, emin=0.5, emax=40.0, spectral_norm=norm, scale_factor=1 / (1 + z))
    rlam, _, _ = oracle.powerlaw_lambda(cfg, lum, alpha)
    np.testing.assert_allclose(lam.cpu().numpy(), rlam, rtol=1e-12)
    out = model.generate_chunk(next(src.chunks()), norm)
    nc, onph, oact, oe = oracle.powerlaw_process_data(cfg, lum, alpha, np.random.RandomState(1))
    e = out.energy.cpu().numpy()
    assert abs(e.size - rlam.sum()) < 6 * np.sqrt(rlam.sum())
    assert e.min() >= 0.5 / 1.5 * (1 - 1e-12) and e.max() <= 40.0 / 1.5 * (1 + 1e-12)
    assert _ks_ok(e, oe)
    # alpha == 1 and alpha == 2 cells separately
    idx, nph = out.idx.cpu().numpy(), out.nph.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(nph)])
    ooff = np.concatenate([[0], np.cumsum(onph)])
    oidx = np.where(oact)[0]
    for lo, hi in ((0, 200), (200, 400)):
        mine = np.concatenate([e[off[k]:off[k + 1]] for k in np.where((idx >= lo) & (idx < hi))[0]])
        theirs = np.concatenate([oe[ooff[k]:ooff[k + 1]] for k in np.where((oidx >= lo) & (oidx < hi))[0]])
        assert _ks_ok(mine, theirs)
    # the reference returns a bool mask
    res = model.process_data("photons", next(src.chunks()), norm)
    assert res[2].dtype == bool and res[2].sum() == res[0]


def test_line_counts_and_energies(cuda, oracle):
    n = 3000
    rng = np.random.RandomState(3)
    rate = 1.0e42 * rng.lognormal(0, 1, n)
    sig_kms = rng.uniform(100.0, 2000.0, n)
    src = px.ArrayDataSource({("gas", "line"): rate, ("gas", "sig"): sig_kms,
                              ("index", "x"): np.zeros(n), ("index", "y"): np.zeros(n),
                              ("index", "z"): np.zeros(n)})
    z = 0.2
    norm = 5.0e-41
    cfg = dict(e0=3.5, spectral_norm=norm, scale_factor=1 / (1 + z))
    for sigma, sig_keV in ((("gas", "sig")), sig_kms * 3.5 / 299792.458), ((1000.0, "km/s"), 1000.0 * 3.5 / 299792.458), (0.02, 0.02):
        model = px.LineSourceModel(3.5, ("gas", "line"), sigma, prng=5)
        model.setup_model("photons", src, z)
        lam, _ = model.expected_counts(next(src.chunks()), norm)
        nc, onph, oact, oe, F = oracle.line_process_data(cfg, rate, sig_keV, np.random.RandomState(4))
        np.testing.assert_allclose(lam.cpu().numpy(), F, rtol=1e-14)
        out = model.generate_chunk(next(src.chunks()), norm)
        e = out.energy.cpu().numpy()
        assert abs(e.size - F.sum()) < 1.645 * 4 * np.sqrt(F.sum())
        assert _ks_ok(e, oe)
        assert abs(e.mean() - 3.5 / 1.2) < 5 * e.std() / np.sqrt(e.size)


# ---------------------------------------------------------------------------------------
# decisions from random words; float32 events
# ---------------------------------------------------------------------------------------
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011) on uint64-held 32-bit lanes; returns the four words."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & np.uint64(MASK) for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & MASK, int(k1) & MASK
    for _ in range(10):
        p0 = c0 * np.uint64(M0)
        p1 = c2 * np.uint64(M1)
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def test_philox_known_answer():
    """Random123's published known-answer vectors for Philox4x32-10."""
    out = _philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = _philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = _philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(v) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("kind", ["wabs", "table"])
def test_absorb_decision_from_random_words(cuda, oracle, kind):
    """The photon kernels decide absorption from the TOP 32-bit word of a 64-bit uniform and draw
    the low word (refinement stream, Philox counter (draw, cell, stream 10)) only when the top
    word alone does not decide.  Against U = (w + w2 / 2^32) / 2^32 < exp(-sigma nH) evaluated on
    the host with the same Philox words: top words chosen ON the threshold (so the low word
    decides), one either side of it, and at random."""
    from pyxsim_b200 import _lib
    from pyxsim_b200.device import ptr, stream_ptr, to_device
    import ctypes as C

    rng = np.random.RandomState(11)
    n = 60000
    e = np.exp(rng.uniform(np.log(0.05), np.log(12.0), n))
    seed, gid, d0 = 0x1234567890ABCDEF, 987654321012, 5_000_000_000
    for nH in (0.02, 1.0, 20.0):
        if kind == "wabs":
            prod, orc = px.WabsModel(nH), oracle.WabsAbsorb(nH)
        else:
            en, sg = synthetic.tbabs_like_sigma(n=3000)
            prod, orc = px.TBabsModel(nH, energy=en, cross_section=sg), oracle.TableAbsorb(nH, en, sg * 1.0e22)
        p = orc.get_absorb(e)
        d = d0 + np.arange(n, dtype=np.uint64)
        c3 = ((np.uint64(gid) >> np.uint64(32)) << np.uint64(16)) | np.uint64(10)
        w2 = _philox4x32_10(d & np.uint64(0xFFFFFFFF), d >> np.uint64(32), np.uint64(gid) & np.uint64(0xFFFFFFFF),
                            c3, seed & 0xFFFFFFFF, seed >> 32)[0].astype(np.float64)
        on = np.clip(np.floor(p * 4294967296.0), 0, 4294967295)
        for w in (on, np.clip(on - 1, 0, None), np.clip(on + 1, None, 4294967295),
                  rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.float64)):
            U = (w + w2 / 4294967296.0) / 4294967296.0
            want = U < p
            sure = np.abs(U - p) > 4e-16 * np.maximum(p, 1e-300) + 1e-300      # (the host's p is fp64 too)
            P = _lib.ProjectParams()
            keep = prod._fill_params(P)
            P.nH = nH
            P.seed = seed
            de = to_device(e)
            dw = torch.from_numpy(w.astype(np.uint32).view(np.int32)).cuda()
            out = torch.empty(n, dtype=torch.uint8, device="cuda")
            _lib.call("pxb_absorb_decide_words", C.byref(P), n, ptr(de), ptr(dw), C.c_uint64(gid), C.c_uint64(d0),
                      ptr(out), stream_ptr())
            got = out.cpu().numpy().astype(bool)
            del keep
            assert np.array_equal(got[sure], want[sure]), (kind, nH)
        # on the threshold both outcomes occur: the low word really decides
        U = (on + w2 / 4294967296.0) / 4294967296.0
        mid = (p > 1e-6) & (p < 1 - 1e-6)
        assert 0.2 < (U < p)[mid].mean() < 0.8


def test_float32_events(cuda):
    """event_dtype="float32": the same survivors in the same order, energies rounded to single
    precision, sky positions as single-precision offsets from the fp64 centre (so they keep
    ~1e-7 deg at any RA); flux / emin / emax stay fp64; EventList gives absolute positions back."""
    d, p = make_plist(n_cells=4000, mean_ph=40.0)
    put("f32_ph", d, p)
    center = [310.0, -67.5]
    kw = dict(absorb_model="wabs", nH=0.05, save_los=True, prng=5)
    n64 = px.project_photons("mem:f32_ph", "mem:f32_ev64", [0.3, 0.2, 0.9], center, **kw)
    n32 = px.project_photons("mem:f32_ph", "mem:f32_ev32", [0.3, 0.2, 0.9], center, event_dtype="float32", **kw)
    a, b = px.load_list("mem:f32_ev64"), px.load_list("mem:f32_ev32")
    assert n64 == n32 > 1000 and b.data["eobs"].dtype == torch.float32
    assert b.parameters["event_dtype"] == "float32" and b.parameters["sky_offsets"] == 1
    assert np.array_equal(b.numpy("eobs"), a.numpy("eobs").astype(np.float32))
    assert np.array_equal(b.numpy("los"), a.numpy("los").astype(np.float32))
    assert np.array_equal(b.numpy("xsky"), (a.numpy("xsky") - center[0]).astype(np.float32))
    assert np.array_equal(b.numpy("ysky"), (a.numpy("ysky") - center[1]).astype(np.float32))
    for k in ("flux", "emin", "emax"):
        assert a.parameters[k] == b.parameters[k]
    ev = px.EventList("mem:f32_ev32")
    np.testing.assert_allclose(ev["xsky"], a.numpy("xsky"), rtol=0, atol=1e-7)
    np.testing.assert_allclose(ev["ysky"], a.numpy("ysky"), rtol=0, atol=1e-7)
    np.testing.assert_allclose(ev["eobs"], a.numpy("eobs"), rtol=1e-7)
    img64 = px.EventList("mem:f32_ev64").image(20.0, 128).cpu().numpy()
    img32 = ev.image(20.0, 128).cpu().numpy()
    assert img64.sum() == img32.sum() and np.abs(img64 - img32).sum() <= 1e-4 * img64.sum()
    # physical coordinates have no centre to be relative to
    px.project_photons("mem:f32_ph", "mem:f32_evp", "x", center, phys_coord=True, event_dtype="float32", prng=5)
    c = px.load_list("mem:f32_evp")
    assert c.parameters["sky_offsets"] == 0
    px.project_photons("mem:f32_ph", "mem:f32_evp64", "x", center, phys_coord=True, prng=5)
    assert np.array_equal(c.numpy("xsky"), px.load_list("mem:f32_evp64").numpy("xsky").astype(np.float32))
