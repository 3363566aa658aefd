"""CPU: the index-addressed synthetic datasets `bench.py --config amr|sph|plline` and the sharded
default workload are built from (`pyxsim_b200/synthetic.py`).  Every rank materialises only the
cells it owns, so the fields must be pure functions of the GLOBAL cell index, the block-cyclic
index sets of the ranks must partition the dataset, and the nested-AMR numbering must tile the
box exactly once."""
import numpy as np
import pytest
import torch

from pyxsim_b200 import parallel, synthetic


@pytest.mark.parametrize("n,block,size", [(1000, 64, 3), (4096, 256, 8), (777, 1000, 2), (5, 1, 4)])
def test_shard_indices_partition_the_dataset_like_shard_bounds(n, block, size):
    seen = []
    for rank in range(size):
        idx = synthetic.shard_indices(n, block, rank, size).numpy()
        assert (np.diff(idx) > 0).all()                       # ascending, no repeats
        want = np.concatenate([np.arange(a, b) for a, b in parallel.shard_bounds(n, block, rank, size)]
                              or [np.empty(0, np.int64)])
        assert np.array_equal(idx, want)
        seen.append(idx)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(n))


def test_hash_fields_are_functions_of_the_global_index_only():
    n = 24
    idx = torch.arange(n ** 3, dtype=torch.int64)
    full = synthetic.cluster_chain_fields(idx, n, n_clusters=1, kT_range=(1.0, 9.0))
    for rank in range(3):
        mine = synthetic.shard_indices(n ** 3, 64, rank, 3)
        part = synthetic.cluster_chain_fields(mine, n, n_clusters=1, kT_range=(1.0, 9.0))
        for k in full:
            assert torch.equal(part[k], full[k][mine]), k
    u = synthetic.hash_uniform(torch.arange(200000, dtype=torch.int64), 7).numpy()
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 3e-3
    g = synthetic.hash_normal(torch.arange(200000, dtype=torch.int64), 3).numpy()
    assert abs(g.mean()) < 1e-2 and abs(g.std() - 1.0) < 1e-2
    # different salts decorrelate
    u2 = synthetic.hash_uniform(torch.arange(200000, dtype=torch.int64), 8).numpy()
    assert abs(np.corrcoef(u, u2)[0, 1]) < 1e-2


def test_cluster_chain_is_the_single_cluster_repeated():
    n = 12
    one = synthetic.cluster_chain_fields(torch.arange(n ** 3, dtype=torch.int64), n, n_clusters=1)
    two = synthetic.cluster_chain_fields(torch.arange(2 * n ** 3, dtype=torch.int64), n, n_clusters=2)
    em1 = one[("gas", "emission_measure")]
    em2 = two[("gas", "emission_measure")]
    # x-major numbering: block by block (positions are offset by the chain's half length, so the
    # radii agree to rounding, not bit for bit)
    np.testing.assert_allclose(em2[: n ** 3].numpy(), em1.numpy(), rtol=1e-11)
    np.testing.assert_allclose(em2[n ** 3:].numpy(), em1.numpy(), rtol=1e-11)
    assert float(em1.sum()) > 0.0
    # the two clusters sit side by side along x, box-width apart
    x2 = two[("index", "x")]
    np.testing.assert_allclose((x2[n ** 3:] - x2[: n ** 3]).numpy(), 2000.0, rtol=0, atol=1e-9)


def test_amr_leaves_tile_the_box_exactly_once():
    n, box = 16, 2000.0
    nleaf = synthetic.amr_leaf_count(n)
    assert nleaf == 3 * (n ** 3 - (n // 2) ** 3) + n ** 3
    f = synthetic.amr_cluster_fields(torch.arange(nleaf, dtype=torch.int64), n=n, box_kpc=box, nvar=2)
    x, y, z, dx = (f[("index", k)].numpy() for k in ("x", "y", "z", "dx"))
    # volumes add up to the box; four cell sizes, halving per level
    np.testing.assert_allclose((dx ** 3).sum(), box ** 3, rtol=1e-12)
    sizes = np.unique(dx)
    np.testing.assert_allclose(sizes, box / n / 2.0 ** np.arange(3, -1, -1), rtol=1e-14)
    # no two leaves overlap: sample points of the finest lattice are covered exactly once
    fine = box / n / 8.0
    m = int(round(box / fine))
    cover = np.zeros((m, m, m), dtype=np.int32)
    lo = np.rint((np.stack([x, y, z], 1) - 0.5 * dx[:, None] + 0.5 * box) / fine).astype(int)
    w = np.rint(dx / fine).astype(int)
    for (i, j, k), ww in zip(lo, w):
        cover[i:i + ww, j:j + ww, k:k + ww] += 1
    assert cover.min() == 1 and cover.max() == 1
    # a shard sees the same leaves
    mine = synthetic.shard_indices(nleaf, 128, 1, 4)
    part = synthetic.amr_cluster_fields(mine, n=n, box_kpc=box, nvar=2)
    for k in f:
        assert torch.equal(part[k], f[k][mine]), k
    assert {("gas", "El0_fraction"), ("gas", "El1_fraction"), ("gas", "metallicity")} <= set(f)


def test_sph_particles_follow_the_beta_profile_and_shard():
    n = 200000
    f = synthetic.sph_igm_fields(torch.arange(n, dtype=torch.int64))
    r = np.sqrt(sum(f[("gas", f"particle_position_{a}")].numpy() ** 2 for a in "xyz"))
    assert r.max() <= 1000.0 * (1 + 1e-9)
    # radial CDF of rho ~ (1 + (r/rc)^2)^-1: M(<x) ~ x - atan(x)
    xs = np.array([0.5, 1.0, 3.0, 10.0])
    want = (xs - np.arctan(xs)) / (20.0 - np.arctan(20.0))
    got = np.array([(r / 50.0 < x).mean() for x in xs])
    np.testing.assert_allclose(got, want, atol=4e-3)
    mine = synthetic.shard_indices(n, 4096, 2, 8)
    part = synthetic.sph_igm_fields(mine)
    for k in f:
        assert torch.equal(part[k], f[k][mine]), k
