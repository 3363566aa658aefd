"""GPU: an N-rank run of ONE dataset equals the 1-rank run, bit for bit.

The reference deals yt chunks over MPI ranks (pyxsim/photon_list.py:401), reduces the counts
(:471-472,826) and writes one list per rank (:214-216).  Here the cells of one ArrayDataSource are
split block-cyclically over the ranks (``parallel.shard_bounds`` / ``ShardedChunk``); every random
draw is keyed by the cell's GLOBAL id, so the union of the per-rank photon and event lists must be
exactly the single-rank lists.  Both ranks run on cuda:0 (the box has one GPU), scalar reductions
go through the gloo backend -- the same host code path torchrun + NCCL takes on 8 GPUs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import pyxsim_b200 as px
from pyxsim_b200 import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TF = ("gas", "kT")

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from pyxsim_b200 import parallel, synthetic
import pyxsim_b200 as px
import logging
logging.getLogger("pyxsim_b200").setLevel("WARNING")
parallel.init_from_env(backend="gloo")
mode = {mode!r}
kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=600)
src = synthetic.beta_model_source(24, kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda")
if mode == "contiguous":
    src.chunk_size = (src.size + parallel.size() - 1) // parallel.size()
else:
    src.shard_block = 512
sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
model = px.ThermalSourceModel(sm, 0.1, 10.0, 600, 0.3, temperature_field=("gas", "kT"), prng=77)
out = {out!r}
n_ph, n_c = px.make_photons(out + "/ph", src, 0.05, 4.0e4, 1.0e5, model)
n_ev = px.project_photons(out + "/ph", out + "/ev", [0.2, -0.1, 0.95], [30.0, 45.0], absorb_model="wabs",
                          nH=0.05, prng=78, save_los=True)
if parallel.rank() == 0:
    open(out + "/totals.txt", "w").write(f"{{n_ph}} {{n_c}} {{n_ev}}")
parallel.barrier()
"""


def _single_rank(tmp):
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=600)
    src = synthetic.beta_model_source(24, kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda")
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, 600, 0.3, temperature_field=TF, prng=77)
    n_ph, n_c = px.make_photons("mem:mr_ph", src, 0.05, 4.0e4, 1.0e5, model)
    n_ev = px.project_photons("mem:mr_ph", "mem:mr_ev", [0.2, -0.1, 0.95], [30.0, 45.0],
                              absorb_model="wabs", nH=0.05, prng=78, save_los=True)
    return (n_ph, n_c, n_ev), px.load_list("mem:mr_ph").to_host(), px.load_list("mem:mr_ev").to_host()


def _run_ranks(tmp_path, mode, nranks, port):
    out = tmp_path / f"{mode}{nranks}"
    out.mkdir()
    script = out / "worker.py"
    script.write_text(WORKER.format(root=ROOT, mode=mode, out=str(out)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    totals = tuple(int(v) for v in (out / "totals.txt").read_text().split())
    ph = [np.load(out / f"ph.{r:04d}.npz") for r in range(nranks)]
    ev = [np.load(out / f"ev.{r:04d}.npz") for r in range(nranks)]
    return totals, ph, ev


@pytest.mark.parametrize("mode,nranks,port", [("contiguous", 2, 29741), ("cyclic", 2, 29742),
                                              ("cyclic", 3, 29743)])
def test_n_rank_run_equals_single_rank(cuda, tmp_path, mode, nranks, port):
    tot1, ph1, ev1 = _single_rank(tmp_path)
    totN, phN, evN = _run_ranks(tmp_path, mode, nranks, port)
    assert totN == tot1                                      # the reduced counts (photon_list.py:471-472,826)
    # every rank got work and no cell is owned twice
    ids = np.concatenate([p["data/cell_id"] for p in phN])
    assert all(p["data/cell_id"].size > 0 for p in phN)
    assert np.unique(ids).size == ids.size == ph1.data["cell_id"].size
    cell_keys = ("x", "y", "z", "vx", "vy", "vz", "dx", "num_photons", "cell_id")
    if mode == "contiguous":
        # rank r owns one contiguous range of cells: the concatenation in rank order IS the list
        for k in cell_keys + ("energy",):
            assert np.array_equal(np.concatenate([p[f"data/{k}"] for p in phN]), ph1.data[k]), k
        for k in ("xsky", "ysky", "eobs", "los"):
            assert np.array_equal(np.concatenate([e[f"data/{k}"] for e in evN]), ev1.data[k]), k
    else:
        # block-cyclic: the per-rank lists interleave at block granularity.  Cells are compared
        # after sorting by global id, photons by moving every cell's energies back to its place,
        # events as multisets of (eobs, xsky, ysky, los) rows.
        order = np.argsort(ids, kind="stable")
        for k in cell_keys:
            assert np.array_equal(np.concatenate([p[f"data/{k}"] for p in phN])[order], ph1.data[k]), k
        nph = np.concatenate([p["data/num_photons"] for p in phN])
        en = np.concatenate([p["data/energy"] for p in phN])
        starts = np.concatenate([[0], np.cumsum(nph)])
        en_sorted = np.concatenate([en[starts[c]:starts[c + 1]] for c in order])
        assert np.array_equal(en_sorted, ph1.data["energy"])

        def rows(get):
            a = np.stack([get(k) for k in ("eobs", "xsky", "ysky", "los")], axis=1)
            return a[np.lexsort(a.T[::-1])]

        got = rows(lambda k: np.concatenate([e[f"data/{k}"] for e in evN]))
        want = rows(lambda k: ev1.data[k])
        assert np.array_equal(got, want)
    # flux is a sum over the events of every rank's shard
    fl = sum(float(e["parameters/flux"]) for e in evN)
    assert abs(fl - float(ev1.parameters["flux"])) <= 1e-12 * fl
    assert min(float(e["parameters/emin"]) for e in evN) == float(ev1.parameters["emin"])
    assert max(float(e["parameters/emax"]) for e in evN) == float(ev1.parameters["emax"])


def test_reference_style_list_gets_distinct_keys_per_rank(cuda):
    """A photon list without the ``cell_id`` dataset (one written by the reference) falls back to
    cell_id0 + local index; with cell_id0 = 0 that is the single-rank behaviour."""
    _, ph1, ev1 = _single_rank(None)
    from pyxsim_b200.photon_list import DeviceList, save_list

    d = {k: v for k, v in ph1.data.items() if k != "cell_id"}
    save_list("mem:mr_noid", DeviceList(ph1.info, ph1.parameters, d))
    n = px.project_photons("mem:mr_noid", "mem:mr_noid_ev", [0.2, -0.1, 0.95], [30.0, 45.0],
                           absorb_model="wabs", nH=0.05, prng=78)
    # a different key stream, the same physics: event fraction within binomial error
    n1 = ev1.data["eobs"].size
    npht = ph1.data["energy"].size
    p = n1 / npht
    assert abs(n - n1) < 6 * np.sqrt(2 * npht * p * (1 - p))
    # and with ids 0..n-1 supplied explicitly it reproduces that run exactly
    d2 = dict(d, cell_id=np.arange(ph1.data["num_photons"].size, dtype=np.int64))
    save_list("mem:mr_id", DeviceList(ph1.info, ph1.parameters, d2))
    n2 = px.project_photons("mem:mr_id", "mem:mr_id_ev", [0.2, -0.1, 0.95], [30.0, 45.0],
                            absorb_model="wabs", nH=0.05, prng=78)
    assert n2 == n
    a, b = px.load_list("mem:mr_noid_ev"), px.load_list("mem:mr_id_ev")
    for k in ("xsky", "ysky", "eobs"):
        assert np.array_equal(a.numpy(k), b.numpy(k))
