"""CPU: host-side logic (no kernels): orientation, cosmology, units, chunking, list I/O,
argument validation, and the N>1 sharding / count reduction over gloo (world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pyxsim_b200 as px
from pyxsim_b200 import parallel, utils
from pyxsim_b200.photon_list import DeviceList, load_list, save_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_orientation_matches_oracle_and_is_orthonormal(oracle):
    rng = np.random.RandomState(0)
    for _ in range(20):
        L = rng.normal(size=3)
        north = rng.normal(size=3) if rng.uniform() < 0.5 else None
        uv = utils.orientation(L, north)
        np.testing.assert_allclose(uv @ uv.T, np.eye(3), atol=1e-14)
        np.testing.assert_allclose(uv[2], L / np.linalg.norm(L), atol=1e-15)
        assert np.array_equal(uv, oracle.orientation(L, north))
    np.testing.assert_allclose(utils.orientation([0, 0, 1.0]), [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)


def test_cosmology_angular_diameter_distance():
    c = px.Cosmology()
    d = c.angular_diameter_distance_mpc(0.0, 0.05)
    # independent trapezoid integration
    z = np.linspace(0, 0.05, 200001)
    e = np.sqrt(0.27 * (1 + z) ** 3 + 0.73)
    dc = 299792.458 / 71.0 * np.trapezoid(1 / e, z)
    np.testing.assert_allclose(d, dc / 1.05, rtol=1e-9)
    assert 195 < d < 205


def test_parse_value_and_seed():
    assert utils.parse_value((100.0, "ks"), "s") == 1.0e5
    assert utils.parse_value(3.0, "keV") == 3.0
    assert utils.parse_value((1.0, "Mpc"), "kpc") == 1000.0
    with pytest.raises(ValueError):
        utils.parse_value((1.0, "kpc"), "s")
    assert utils.prng_seed(42) == 42
    a = utils.prng_seed(np.random.RandomState(1))
    assert a == utils.prng_seed(np.random.RandomState(1)) and a != utils.prng_seed(np.random.RandomState(2))


def test_array_data_source_chunks_and_units():
    n = 10
    f = {("gas", "temperature"): np.full(n, 1.160451812e7), ("gas", "emission_measure"): np.ones(n),
         ("index", "x"): np.arange(n, dtype=float), ("index", "y"): np.zeros(n), ("index", "z"): np.zeros(n)}
    src = px.ArrayDataSource(f, chunk_size=4)
    ch = list(src.chunks())
    assert [(c.start, c.stop) for c in ch] == [(0, 4), (4, 8), (8, 10)]
    np.testing.assert_allclose(ch[0].to_value(("gas", "temperature"), "keV"), 1.0, rtol=1e-12)
    assert [(c.start, c.stop) for c in src.chunks(rank=1, size=2)] == [(4, 8)]
    assert src.width_field is None
    with pytest.raises(ValueError):
        px.ArrayDataSource({"a": np.ones(3), "b": np.ones(4)})


def test_list_roundtrip_npz(tmp_path):
    lst = DeviceList({"pyxsim_version": "x", "filenames": ["a.h5"]},
                     {"fid_area": 1000.0, "data_type": "cells", "center": np.zeros(3), "absoption_model": "wabs"},
                     {"x": np.arange(5.0), "num_photons": np.arange(5)})
    prefix = str(tmp_path / "plist")
    save_list(prefix, lst)
    back = load_list(prefix)
    assert back.parameters["data_type"] == "cells" and back.parameters["absoption_model"] == "wabs"
    assert float(back.parameters["fid_area"]) == 1000.0
    assert np.array_equal(back.data["x"], np.arange(5.0))
    with pytest.raises(FileNotFoundError):
        load_list(str(tmp_path / "nothing"))


def test_make_photons_argument_errors():
    import torch

    if torch.cuda.is_available():
        pytest.skip("argument checks below rely on the no-GPU guard order")
    # redshift <= 0 without dist must raise ValueError (photon_list.py:233-238) -- after the
    # CUDA guard on a CPU-only box, so check the guard message instead
    src = px.ArrayDataSource({("index", "x"): np.zeros(2), ("index", "y"): np.zeros(2), ("index", "z"): np.zeros(2)})
    with pytest.raises(RuntimeError):
        px.make_photons("mem:x", src, 0.0, 1.0, 1.0, px.SourceModel())


def test_shard_bounds_cover_everything():
    n = 100000
    seen = np.zeros(n, dtype=int)
    for r in range(4):
        for a, b in parallel.shard_bounds(n, block=4096, r=r, s=4):
            seen[a:b] += 1
    assert np.all(seen == 1)


@pytest.mark.parametrize("n,block,size", [(100000, 4096, 4), (10, 4, 3), (4096 * 6, 4096, 2), (5, 8, 2)])
def test_sharded_chunk_is_the_block_cyclic_share(n, block, size):
    """ArrayDataSource.chunks(rank, size) without a chunk_size: ONE chunk per rank holding the
    blocks parallel.shard_bounds assigns to it, with the global-id map the kernels use
    (cell_id0 + (i // id_block) * id_stride + i % id_block)."""
    import torch

    x = np.arange(n, dtype="float64")
    src = px.ArrayDataSource({("index", "x"): x, ("index", "y"): torch.arange(n, dtype=torch.float64),
                              ("index", "z"): x}, shard_block=block, global_offset=1000)
    seen = []
    for r in range(size):
        want = np.concatenate([np.arange(a, b) for a, b in parallel.shard_bounds(n, block=block, r=r, s=size)]
                              or [np.array([], dtype=int)])
        chunks = list(src.chunks(rank=r, size=size))
        if want.size == 0:
            assert chunks == []
            continue
        (ch,) = chunks
        assert ch.size == want.size
        np.testing.assert_array_equal(ch[("index", "x")], want.astype("float64"))
        np.testing.assert_array_equal(ch[("index", "y")].numpy(), want.astype("float64"))
        id0, blk, stride = ch.id_map()
        i = np.arange(ch.size)
        np.testing.assert_array_equal(id0 + (i // blk) * stride + i % blk, want + 1000)
        np.testing.assert_array_equal(ch.global_ids(torch.arange(ch.size)).numpy(), want + 1000)
        seen.append(want)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(n))
    # with a chunk_size the chunks are dealt round-robin (blocks of chunk_size cells)
    src.chunk_size = block
    got = [(c.start, c.stop) for c in src.chunks(rank=1, size=size)]
    assert got == parallel.shard_bounds(n, block=block, r=1, s=size)
    assert all(c.id_map() == (c.start + 1000, 0, 0) for c in src.chunks(rank=1, size=size))


def test_merge_validation_handles_string_arrays_and_derived_keys():
    """velocity_fields is a byte-string array; flux/emin/emax differ between any two
    real event lists and must not block a merge (they are recomputed)."""
    from pyxsim_b200 import event_list as E

    vf = np.array([("gas", "velocity_x"), ("gas", "velocity_y"), ("gas", "velocity_z")]).astype("S")
    a = dict(fid_exp_time=1.0, velocity_fields=vf, data_type="cells")
    b = dict(fid_exp_time=1.0, velocity_fields=vf.astype("U"), data_type=b"cells")
    E.validate_parameters(a, b)
    with pytest.raises(RuntimeError):
        E.validate_parameters(a, dict(b, velocity_fields=vf[::-1].copy()))


def test_absorb_model_registry():
    assert set(px.absorb_models) == {"wabs", "tbabs"}
    m = px.WabsModel(0.1)
    assert m._name == "wabs" and m.nH == 0.1
    m.nH = 0.2
    assert m.nH == 0.2
    with pytest.raises(ImportError):
        px.TBabsModel(0.1)          # needs the soxs table or explicit arrays


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import torch.distributed as dist
from pyxsim_b200 import parallel
import pyxsim_b200 as px
parallel.init_from_env(backend="gloo")
r, s = parallel.rank(), parallel.size()
assert s == 2
tot = parallel.allreduce_sum(10 + r)
assert tot == 21, tot
g = parallel.allgather_counts([r, 100 * r, 7])
assert g.tolist() == [[0, 0, 7], [1, 100, 7]], g
# chunk dealing mirrors parallel_objects round-robin; per-rank file names
src = px.ArrayDataSource({{("index","x"): np.arange(10.0), ("index","y"): np.zeros(10), ("index","z"): np.zeros(10)}}, chunk_size=3)
mine = [(c.start, c.stop) for c in src.chunks(rank=r, size=s)]
assert mine == ([(0, 3), (6, 9)] if r == 0 else [(3, 6), (9, 10)]), mine
from pyxsim_b200.photon_list import _file_name, _file_names, DeviceList, save_list, load_list
assert _file_name("pre") == f"pre.{{r:04d}}"
assert _file_names("pre") == ["pre.0000.h5", "pre.0001.h5"]
save_list({tmp!r} + "/ev", DeviceList({{}}, {{"n": r}}, {{"e": np.full(3, float(r))}}))
parallel.barrier()
other = np.load({tmp!r} + f"/ev.{{1 - r:04d}}.npz")
assert other["data/e"][0] == 1 - r
# optional concatenation of per-rank shards (variable length) onto rank 0
mine = {{"e": torch.arange(3 + 2 * r, dtype=torch.float64) + 100 * r, "x": torch.full((3 + 2 * r,), float(r), dtype=torch.float64)}}
cat = parallel.gather_arrays(mine, dst=0)
if r == 0:
    assert cat["e"].tolist() == [0.0, 1.0, 2.0, 100.0, 101.0, 102.0, 103.0, 104.0], cat["e"]
    assert cat["x"].tolist() == [0.0] * 3 + [1.0] * 5
else:
    assert cat is None
parallel.barrier()
sys.stdout.write(f"rank{{r}}ok\n"); sys.stdout.flush()
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, tmp=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "rank0ok" in p.stdout and "rank1ok" in p.stdout


def test_fits_writer_roundtrip_and_merge_validation(tmp_path):
    """The self-contained FITS writer (event_list.py; astropy is absent) produces standard
    2880-byte blocks that its reader parses back; merge_files' parameter validation
    (pyxsim/utils.py:63-89) on host lists."""
    from pyxsim_b200 import event_list as E
    from pyxsim_b200.photon_list import DeviceList, save_list, load_list

    img = np.arange(35.0).reshape(5, 7)
    f1 = str(tmp_path / "a.fits")
    E._write_fits(f1, [E._primary_hdu(img, [("CRPIX1", 4.0), ("CTYPE1", "RA---TAN"), ("CDELT1", -1.0 / 3.0)])], False)
    assert os.path.getsize(f1) % 2880 == 0
    (h, data), = E.read_fits(f1)
    assert h["NAXIS1"] == 7 and h["NAXIS2"] == 5 and h["CTYPE1"] == "RA---TAN" and h["CDELT1"] == -1.0 / 3.0
    np.testing.assert_array_equal(data, img)
    cols = [("CHANNEL", "J", None, np.arange(9, dtype="int32")), ("ENERGY", "D", "keV", np.linspace(0, 1, 9)),
            ("RATE", "E", None, np.linspace(0, 2, 9))]
    f2 = str(tmp_path / "b.fits")
    E._write_fits(f2, [E._primary_hdu(), E._bintable_hdu("SPECTRUM", cols, [("POISSERR", True), ("TOTCTS", 12.0)])], False)
    hd = E.read_fits(f2)
    assert hd[0][0]["NAXIS"] == 0 and hd[1][0]["TFIELDS"] == 3 and hd[1][0]["TUNIT2"] == "keV" and hd[1][0]["POISSERR"] is True
    np.testing.assert_array_equal(hd[1][1]["ENERGY"], np.linspace(0, 1, 9))
    with pytest.raises(OSError):
        E._write_fits(f2, [E._primary_hdu()], False)

    p = dict(exp_time=10.0, area=2.0, sky_center=np.array([1.0, 2.0]), observer="external")
    a = DeviceList({"x": 1}, p, {"eobs": np.arange(3.0), "xsky": np.zeros(3), "ysky": np.zeros(3)})
    b = DeviceList({"x": 2}, dict(p, exp_time=30.0), {"eobs": np.arange(4.0), "xsky": np.ones(4), "ysky": np.ones(4)})
    fa, fb = str(tmp_path / "la"), str(tmp_path / "lb")
    save_list(fa, a)
    save_list(fb, b)
    with pytest.raises(RuntimeError):
        E.merge_files([fa, fb], str(tmp_path / "lm"))
    E.merge_files([fa, fb], str(tmp_path / "lm"), add_exposure_times=True)
    m = load_list(str(tmp_path / "lm"))
    assert float(m.parameters["exp_time"]) == 40.0 and m.numpy("eobs").size == 7
    np.testing.assert_array_equal(m.numpy("xsky"), np.concatenate([np.zeros(3), np.ones(4)]))
    with pytest.raises(OSError):
        E.merge_files([fa, fb], str(tmp_path / "lm"), add_exposure_times=True)
    with pytest.raises(RuntimeError):
        E.validate_parameters(p, dict(p, area=3.0))
    with pytest.raises(RuntimeError):
        E.validate_parameters(p, dict(p, extra=1))
    E.validate_parameters(p, dict(p, area=2.0 + 1e-12))


def test_save_list_streams_arrays_in_slabs(tmp_path, monkeypatch):
    """save_list writes every data array slab by slab (``_host_chunks``) into an uncompressed
    ``.npz`` member -- no whole-list host copy -- and np.load / load_list read it back unchanged:
    several slabs per array, a ragged last slab, an empty array, int64 / float32 / float64, string
    and byte-string parameters."""
    import torch

    import pyxsim_b200.photon_list as PL

    monkeypatch.setattr(PL, "STREAM_CHUNK_BYTES", 1000)
    monkeypatch.setattr(PL, "h5py", None)
    rng = np.random.RandomState(3)
    data = {"energy": torch.from_numpy(rng.uniform(0.1, 10.0, 1237)),
            "num_photons": rng.randint(0, 50, 300).astype(np.int64),
            "x": torch.from_numpy(rng.normal(size=300).astype(np.float32)),
            "empty": torch.empty(0, dtype=torch.float64)}
    lst = PL.DeviceList({"data_source": "test", "version": 2},
                        {"fid_exp_time": 1.0e5, "observer": "external",
                         "velocity_fields": np.array([["gas", "velocity_x"], ["gas", "velocity_y"]]).astype("S")},
                        data)
    fn = PL.save_list(str(tmp_path / "lst"), lst)
    assert fn.endswith(".npz")
    z = np.load(fn, allow_pickle=False)
    for k, v in data.items():
        want = v.numpy() if isinstance(v, torch.Tensor) else v
        got = z[f"data/{k}"]
        assert got.dtype == want.dtype and np.array_equal(got, want), k
    back = PL.load_list(str(tmp_path / "lst"))
    assert back.parameters["observer"] == "external" and float(back.parameters["fid_exp_time"]) == 1.0e5
    assert back.parameters["velocity_fields"].shape == (2, 2) and back.info["data_source"] == "test"
    # chunk generator: sizes and order
    sizes = [c.size for c in PL._host_chunks(data["energy"], 1000)]
    assert sizes == [125] * 9 + [112] and sum(sizes) == 1237


def test_background_list_writer(tmp_path, monkeypatch):
    """set_async_writes(True): save_list returns at once, a writer thread streams the list to the
    file, load_list of the same prefix hands back the in-memory list while the file is pending,
    wait_for_writes() joins (and re-raises a writer's error), and the file then holds the list."""
    import torch

    import pyxsim_b200 as px
    import pyxsim_b200.photon_list as PL

    monkeypatch.setattr(PL, "h5py", None)
    monkeypatch.setattr(PL, "STREAM_CHUNK_BYTES", 4096)
    e = torch.arange(50000, dtype=torch.float64)
    lst = PL.DeviceList({"k": 1}, {"fid_exp_time": 2.0}, {"energy": e, "num_photons": np.arange(10)})
    px.set_async_writes(True)
    try:
        fn = PL.save_list(str(tmp_path / "bg"), lst)
        assert fn.endswith("bg.npz")
        assert PL.load_list(str(tmp_path / "bg")) is lst or os.path.exists(fn)   # pending -> the same object
        px.wait_for_writes()
        assert not PL._PENDING
        back = PL.load_list(str(tmp_path / "bg"))
        assert back is not lst and np.array_equal(back.data["energy"], e.numpy())
        # an error in the writer (unwritable path) surfaces at wait_for_writes
        PL.save_list(str(tmp_path / "no_such_dir" / "x"), lst)
        with pytest.raises(OSError):
            px.wait_for_writes()
    finally:
        px.set_async_writes(False)
    assert not PL.ASYNC_WRITES
