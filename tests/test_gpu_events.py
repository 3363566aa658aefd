"""GPU parity for the event-list consumers ("next" row f4): EventList binning
(pyxsim/event_list.py:106-134,254-318,344-349) on the device, against NumPy's histograms, the
published TAN projection and the reference's compiled pixel_to_cel."""
import numpy as np
import pytest
import torch

import pyxsim_b200 as px
from pyxsim_b200.event_list import read_fits
from pyxsim_b200.photon_list import DeviceList, save_list

pytestmark = pytest.mark.gpu


def _events(n=400000, seed=5, center=(30.0, 45.0), spread=0.2, name="ev_f4", exp_time=1.0e5):
    rng = np.random.RandomState(seed)
    d = dict(xsky=center[0] + rng.normal(0, spread, n) / np.cos(np.deg2rad(center[1])),
             ysky=center[1] + rng.normal(0, spread, n),
             eobs=np.exp(rng.uniform(np.log(0.1), np.log(10.0), n)))
    params = dict(exp_time=exp_time, area=1000.0, sky_center=np.array(center), observer="external",
                  flux=1.0e-12, emin=0.1, emax=10.0)
    save_list("mem:" + name, DeviceList({"dataset": "test"}, params, {k: torch.from_numpy(v).cuda() for k, v in d.items()}))
    return d, params


def test_cel_to_pixel_inverts_the_reference_pixel_to_cel(cuda, oracle):
    d, p = _events()
    nx, fov = 512, 60.0
    dtheta = fov / 60.0 / nx
    ev = px.EventList("mem:ev_f4")
    xp, yp = [t.cpu().numpy() for t in ev.pixel_coordinates(fov, nx)]
    rx, ry = oracle.world2pix_tan(d["xsky"], d["ysky"], p["sky_center"], nx, dtheta)
    np.testing.assert_allclose(xp, rx, rtol=0, atol=1e-9)
    np.testing.assert_allclose(yp, ry, rtol=0, atol=1e-9)
    # pixel -> intermediate world coordinates in radians (what scatter_events hands pixel_to_cel,
    # sign convention of sky_functions.pyx:167-168) -> the reference's own inverse projection
    crpix = 0.5 * (nx + 1)
    xs = np.deg2rad((xp - crpix) * dtheta)              # pixel_to_cel negates x itself
    ys = np.deg2rad((yp - crpix) * dtheta)
    oracle.ref()["sky_functions"].pixel_to_cel(xs, ys, np.asarray(p["sky_center"], "float64"))
    np.testing.assert_allclose(xs, d["xsky"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(ys, d["ysky"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("band", [(None, None), (0.5, 2.0)])
def test_image_is_numpy_histogram2d(cuda, oracle, band):
    d, p = _events()
    nx, fov = 256, 40.0          # smaller than the cloud: events outside the image are dropped
    ev = px.EventList("mem:ev_f4")
    img = ev.image(fov, nx, emin=band[0], emax=band[1]).cpu().numpy()
    xp, yp = [t.cpu().numpy() for t in ev.pixel_coordinates(fov, nx)]
    # binning semantics in isolation: the GPU's own pixel coordinates through np.histogram2d
    ref = oracle.event_image(d["xsky"], d["ysky"], d["eobs"], p["sky_center"], fov, nx, band[0], band[1],
                             pix=(xp, yp))
    assert img.sum() > 50000 and img.sum() < d["eobs"].size
    np.testing.assert_array_equal(img, ref.astype(np.int64))
    # and end to end against the published projection: only events within 1e-9 pixel of an edge
    # may land in the neighbouring pixel
    ref2 = oracle.event_image(d["xsky"], d["ysky"], d["eobs"], p["sky_center"], fov, nx, band[0], band[1])
    assert np.abs(img - ref2).sum() <= 4


def test_image_edges_follow_numpy(cuda, oracle):
    """Events exactly on pixel edges, on the outer frame and just outside it."""
    nx, fov = 8, 8.0
    dtheta = fov / 60.0 / nx
    # pixel coordinates wanted -> sky through the reference's pixel_to_cel
    want = np.array([0.5, 0.5 - 1e-9, 1.5, 2.5, 8.5, 8.5 + 1e-9, 4.0, 4.5, 7.999999])
    gx, gy = np.meshgrid(want, want)
    crpix = 0.5 * (nx + 1)
    xs = np.deg2rad((gx.ravel() - crpix) * dtheta)
    ys = np.deg2rad((gy.ravel() - crpix) * dtheta)
    oracle.ref()["sky_functions"].pixel_to_cel(xs, ys, np.array([30.0, 45.0]))
    save_list("mem:ev_edge", DeviceList({}, dict(exp_time=1.0, area=1.0, sky_center=np.array([30.0, 45.0])),
                                        {"xsky": torch.from_numpy(xs).cuda(), "ysky": torch.from_numpy(ys).cuda(),
                                         "eobs": torch.ones(xs.size, dtype=torch.float64).cuda()}))
    ev = px.EventList("mem:ev_edge")
    img = ev.image(fov, nx).cpu().numpy()
    xp, yp = [t.cpu().numpy() for t in ev.pixel_coordinates(fov, nx)]
    ref = oracle.event_image(xs, ys, np.ones(xs.size), [30.0, 45.0], fov, nx, pix=(xp, yp))
    np.testing.assert_array_equal(img, ref.astype(np.int64))


def test_spectrum_is_numpy_histogram(cuda, oracle):
    d, p = _events()
    e = d["eobs"].copy()
    ebins = np.linspace(0.3, 7.0, 1025)
    e[:1025] = ebins                       # every edge exactly, first and last included
    e[1025:1030] = [0.3 - 1e-12, 7.0 + 1e-12, np.nextafter(7.0, 0), np.nextafter(0.3, 1), 6.99999]
    save_list("mem:ev_spec", DeviceList({}, p, {"xsky": torch.from_numpy(d["xsky"]).cuda(),
                                               "ysky": torch.from_numpy(d["ysky"]).cuda(),
                                               "eobs": torch.from_numpy(e).cuda()}))
    ev = px.EventList("mem:ev_spec")
    for nchan in (1024, 10, 20000):         # 20000 > the shared-memory histogram: global atomics
        eb, cnt = ev.spectrum(0.3, 7.0, nchan)
        rb, rc = oracle.event_spectrum(e, 0.3, 7.0, nchan)
        np.testing.assert_array_equal(eb, rb)
        np.testing.assert_array_equal(cnt.cpu().numpy(), rc)


def test_fits_writers_and_merge(cuda, oracle, tmp_path):
    d, p = _events(n=50000, name="ev_a", seed=1)
    d2, _ = _events(n=30000, name="ev_b", seed=2)
    ev = px.EventList("mem:ev_a")
    assert ev.tot_num_events == 50000 and ev["eobs"].size == 50000
    nx, fov = 64, 60.0
    f_img, f_spec, f_evt = [str(tmp_path / n) for n in ("img.fits", "spec.fits", "evt.fits")]
    ev.write_fits_image(f_img, fov, nx, emin=0.5, emax=7.0)
    (h, img), = read_fits(f_img)
    assert h["CTYPE1"] == "RA---TAN" and h["CRPIX1"] == 32.5 and h["CDELT1"] == -(fov / 60.0 / nx)
    assert h["CRVAL2"] == 45.0 and h["EXPOSURE"] == 1.0e5
    np.testing.assert_array_equal(img, ev.image(fov, nx, 0.5, 7.0).cpu().numpy())
    ev.write_spectrum(f_spec, 0.1, 10.0, 500)
    hs = read_fits(f_spec)
    assert hs[1][0]["EXTNAME"] == "SPECTRUM" and hs[1][0]["DETCHANS"] == 500 and hs[1][0]["HDUCLAS1"] == "SPECTRUM"
    rb, rc = oracle.event_spectrum(d["eobs"], 0.1, 10.0, 500)
    np.testing.assert_array_equal(hs[1][1]["COUNTS"], rc)
    np.testing.assert_allclose(hs[1][1]["COUNT_RATE"], rc / 1.0e5, rtol=1e-15)
    np.testing.assert_array_equal(hs[1][1]["CHANNEL"], np.arange(500) + 1)
    ev.write_fits_file(f_evt, fov, nx)
    he = read_fits(f_evt)
    xp, yp = oracle.world2pix_tan(d["xsky"], d["ysky"], p["sky_center"], nx, fov / 60.0 / nx)
    keep = (xp >= 0.5) & (xp <= nx + 0.5) & (yp >= 0.5) & (yp <= nx + 0.5)
    assert he[1][0]["NAXIS2"] == keep.sum() and he[1][0]["TCTYP2"] == "RA---TAN"
    np.testing.assert_allclose(he[1][1]["ENERGY"], (d["eobs"][keep] * 1000.0).astype(np.float32), rtol=1e-6)
    np.testing.assert_allclose(he[1][1]["X"], xp[keep], atol=1e-9)
    with pytest.raises(OSError):
        ev.write_fits_image(f_img, fov, nx)
    # merge_files: in HBM and on disk
    px.merge_files(["mem:ev_a", "mem:ev_b"], "mem:ev_ab")
    m = px.EventList("mem:ev_ab")
    assert m.tot_num_events == 80000 and m.parameters["exp_time"] == 1.0e5
    np.testing.assert_array_equal(m["eobs"], np.concatenate([d["eobs"], d2["eobs"]]))
    px.merge_files(["mem:ev_a", "mem:ev_b"], "mem:ev_ab2", add_exposure_times=True)
    assert px.EventList("mem:ev_ab2").parameters["exp_time"] == 2.0e5
    out = px.merge_files(["mem:ev_a", "mem:ev_b"], str(tmp_path / "merged.h5"))
    md = px.EventList(str(tmp_path / "merged.h5"))
    assert md.tot_num_events == 80000
    np.testing.assert_array_equal(md.image(fov, nx).cpu().numpy(), m.image(fov, nx).cpu().numpy())
    with pytest.raises(OSError):
        px.merge_files(["mem:ev_a", "mem:ev_b"], str(tmp_path / "merged.h5"))
    _events(n=10, name="ev_c", seed=3, exp_time=5.0e4)
    with pytest.raises(RuntimeError):
        px.merge_files(["mem:ev_a", "mem:ev_c"], "mem:ev_ac")          # exposure times differ
    px.merge_files(["mem:ev_a", "mem:ev_c"], "mem:ev_ac", add_exposure_times=True)
    assert px.EventList("mem:ev_ac").parameters["exp_time"] == 1.5e5


def test_merge_real_photon_and_event_lists(cuda, tmp_path):
    """merge two lists as make_photons / project_photons actually write them --
    per-list flux/emin/emax differ (recomputed for the merged list), velocity_fields is a
    byte-string array."""
    from pyxsim_b200 import synthetic

    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=300)
    src = synthetic.beta_model_source(12, kT_range=(2.0, 8.0), v3d_sigma_kms=200.0)
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    names = []
    for j, mem in enumerate((True, True, False, False)):
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 300, 0.3, temperature_field=("gas", "kT"), prng=40 + j)
        ph = f"mem:mg_ph{j}" if mem else str(tmp_path / f"ph{j}")
        ev = f"mem:mg_ev{j}" if mem else str(tmp_path / f"ev{j}")
        px.make_photons(ph, src, 0.05, 3.0e4, 1.0e5, model)
        px.project_photons(ph, ev, "z", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=60 + j)
        names.append((ph, ev))
    for a, b, out in ((0, 1, "mem:mg"), (2, 3, str(tmp_path / "mg"))):
        px.merge_files([names[a][0], names[b][0]], out + "_ph")
        px.merge_files([names[a][1], names[b][1]], out + "_ev")
        pa, pb, pm = (px.load_list(n) for n in (names[a][0], names[b][0], out + "_ph"))
        assert pm.numpy("energy").size == pa.numpy("energy").size + pb.numpy("energy").size
        assert np.array_equal(pm.numpy("num_photons"), np.concatenate([pa.numpy("num_photons"), pb.numpy("num_photons")]))
        ea, eb, em = (px.load_list(n) for n in (names[a][1], names[b][1], out + "_ev"))
        assert float(ea.parameters["flux"]) != float(eb.parameters["flux"])
        e = np.concatenate([ea.numpy("eobs"), eb.numpy("eobs")])
        assert np.array_equal(em.numpy("eobs"), e)
        assert float(em.parameters["emin"]) == e.min() and float(em.parameters["emax"]) == e.max()
        np.testing.assert_allclose(float(em.parameters["flux"]), e.sum() * 1.602176634e-9 / 1.0e5 / 3.0e4, rtol=1e-12)
    # a projected list does not record a foreground column it did not use, and names its photon file
    px.project_photons(names[2][0], str(tmp_path / "ev_nonh"), "z", [30.0, 45.0], prng=1)
    lst = px.load_list(str(tmp_path / "ev_nonh"))
    assert "nH" not in lst.parameters and str(lst.info["photon_file"]).endswith("ph2.h5")
    w = px.WabsModel(0.3)
    px.project_photons(names[0][0], "mem:mg_inst", "z", [30.0, 45.0], absorb_model=w, nH=0.02, prng=60)
    assert w.nH == 0.3                                    # a user's instance is not modified
    assert np.array_equal(px.load_list("mem:mg_inst").numpy("eobs"), px.load_list(names[0][1]).numpy("eobs"))


def test_h5py_roundtrip_when_available(cuda, tmp_path):
    """With a real h5py the lists are HDF5 files in the reference's schema (velocity_fields as a
    byte-string array, photon_list.py:371)."""
    h5py = pytest.importorskip("h5py")
    from pyxsim_b200 import synthetic

    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=300)
    src = synthetic.beta_model_source(8, kT_range=(2.0, 8.0))
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 300, 0.3, temperature_field=("gas", "kT"), prng=4)
    px.make_photons(str(tmp_path / "p.h5"), src, 0.05, 3.0e4, 1.0e5, model)
    with h5py.File(tmp_path / "p.h5", "r") as f:
        assert f["parameters"]["velocity_fields"].shape == (3, 2)
        assert f["parameters"]["velocity_fields"].dtype.kind == "S"
    n = px.project_photons(str(tmp_path / "p"), str(tmp_path / "e"), "z", [30.0, 45.0], prng=5)
    assert n == px.load_list(str(tmp_path / "p")).numpy("energy").size


def test_lists_stream_from_the_device_to_the_file(cuda, tmp_path, monkeypatch):
    """File-backed lists are written slab by slab from HBM through two pinned buffers (the download
    of one slab overlapping the write of the previous): same file contents as the in-memory list,
    whatever the slab size; the file-backed two-stage run gives the in-memory run's events."""
    import pyxsim_b200.photon_list as PL
    from pyxsim_b200 import synthetic

    monkeypatch.setattr(PL, "STREAM_CHUNK_BYTES", 40000)        # many slabs, ragged last one
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=300)
    src = synthetic.beta_model_source(12, kT_range=(2.0, 8.0), device="cuda")
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)

    def model():
        return px.ThermalSourceModel(sm, 0.1, 10.0, 300, 0.3, temperature_field=("gas", "kT"), prng=4)

    n_mem = px.make_photons("mem:st_ph", src, 0.05, 3.0e4, 1.0e5, model())
    n_fil = px.make_photons(str(tmp_path / "st_ph"), src, 0.05, 3.0e4, 1.0e5, model())
    assert n_mem == n_fil and n_mem[0] > 20 * 40000 // 8
    a, b = px.load_list("mem:st_ph"), px.load_list(str(tmp_path / "st_ph"))
    assert set(a.data) == set(b.data)
    for k in a.data:
        assert np.array_equal(a.numpy(k), b.numpy(k)), k
    e_mem = px.project_photons("mem:st_ph", "mem:st_ev", "z", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=5)
    e_fil = px.project_photons(str(tmp_path / "st_ph"), str(tmp_path / "st_ev"), "z", [30.0, 45.0],
                               absorb_model="wabs", nH=0.02, prng=5)
    assert e_mem == e_fil > 0
    a, b = px.load_list("mem:st_ev"), px.load_list(str(tmp_path / "st_ev"))
    for k in ("xsky", "ysky", "eobs"):
        assert np.array_equal(a.numpy(k), b.numpy(k)), k
    # background writer: both stages return while their files are still being written, stage 2
    # takes the pending photon list from HBM, and the files end up identical to the synchronous ones
    px.set_async_writes(True)
    try:
        n_bg = px.make_photons(str(tmp_path / "bg_ph"), src, 0.05, 3.0e4, 1.0e5, model())
        e_bg = px.project_photons(str(tmp_path / "bg_ph"), str(tmp_path / "bg_ev"), "z", [30.0, 45.0],
                                  absorb_model="wabs", nH=0.02, prng=5)
        px.wait_for_writes()
    finally:
        px.set_async_writes(False)
    assert n_bg == n_mem and e_bg == e_mem
    for name, ref, keys in (("bg_ph", "st_ph", ("energy", "num_photons", "x", "cell_id")),
                            ("bg_ev", "st_ev", ("xsky", "ysky", "eobs"))):
        a, b = px.load_list(str(tmp_path / name)), px.load_list(str(tmp_path / ref))
        assert isinstance(a.data[keys[0]], np.ndarray)          # read back from the file
        for k in keys:
            assert np.array_equal(a.numpy(k), b.numpy(k)), (name, k)
    # the chunk generator on a device tensor: order, sizes, values
    t = torch.arange(100001, dtype=torch.float64, device="cuda")
    got = np.concatenate([c.copy() for c in PL._host_chunks(t, 8 * 1000)])
    assert np.array_equal(got, np.arange(100001.0))


def test_region_filters(cuda):
    d, p = _events(n=60000, name="ev_flt", seed=8)
    ev = px.EventList("mem:ev_flt")
    (name,) = ev.filter_circle("circ", [30.0, 45.0], 0.15)
    assert name == "mem:circ_ev_flt"
    c = px.EventList(name)
    keep = (d["xsky"] - 30.0) ** 2 + (d["ysky"] - 45.0) ** 2 <= 0.15 ** 2
    np.testing.assert_array_equal(c["eobs"], d["eobs"][keep])
    np.testing.assert_array_equal(c["xsky"], d["xsky"][keep])
    assert c.parameters["emin"] == d["eobs"][keep].min() and c.parameters["emax"] == d["eobs"][keep].max()
    np.testing.assert_allclose(c.parameters["flux"], d["eobs"][keep].sum() * 1.602176634e-9 / 1.0e5 / 1000.0, rtol=1e-12)
    (name,) = ev.filter_rectangle("rect", [29.9, 44.9], [30.1, 45.2])
    r = px.EventList(name)
    keep = (d["xsky"] >= 29.9) & (d["xsky"] < 30.1) & (d["ysky"] >= 44.9) & (d["ysky"] < 45.2)
    np.testing.assert_array_equal(r["ysky"], d["ysky"][keep])
    (name,) = ev.filter_circle("none", [0.0, 0.0], 1e-3)
    lst = px.load_list(name)
    assert lst.data["eobs"].numel() == 0 and lst.parameters["flux"] == 0.0 and np.isnan(lst.parameters["emin"])
