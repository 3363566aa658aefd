"""Shared builders: the same synthetic inputs for the product (pyxsim_b200) and the oracle."""
import numpy as np
import torch

import pyxsim_b200 as px
from pyxsim_b200 import synthetic


def np_fields(fields):
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in fields.items()}


def make_tables(nT=36, nbins=400, nvar=0, negative=False, binscale="linear", seed=7):
    kT, ebins, c, m, v = synthetic.apec_like_tables(nT=nT, nbins=nbins, nvar=nvar, negative=negative,
                                                    binscale=binscale, seed=seed)
    return dict(Tvals=kT, ebins=ebins, cosmic=c, metal=m, var=v, binscale=binscale)


def product_model(t, var_names=()):
    return px.ThermalSpectralModel(t["ebins"], t["Tvals"], t["cosmic"], t["metal"], t["var"],
                                   var_elem_names=var_names, binscale=t["binscale"])


def oracle_model(oracle, t):
    return oracle.TableModel(t["ebins"], t["Tvals"], t["cosmic"], t["metal"], t["var"],
                             binscale=t["binscale"])


def random_cells(n, seed=3, kT_range=(0.03, 60.0), em_scale=1.0e60):
    rng = np.random.RandomState(seed)
    kT = np.exp(rng.uniform(np.log(kT_range[0]), np.log(kT_range[1]), n))
    em = em_scale * rng.lognormal(0.0, 1.5, n)
    return kT, em


def cells_source(kT, em, extra=None, chunk_size=None, data_type="cells"):
    n = kT.size
    rng = np.random.RandomState(99)
    f = {("gas", "kT"): kT, ("gas", "emission_measure"): em,
         ("index", "x"): rng.uniform(-500, 500, n), ("index", "y"): rng.uniform(-500, 500, n),
         ("index", "z"): rng.uniform(-500, 500, n), ("index", "dx"): rng.uniform(1, 20, n),
         ("gas", "velocity_x"): rng.normal(0, 300, n), ("gas", "velocity_y"): rng.normal(0, 300, n),
         ("gas", "velocity_z"): rng.normal(0, 300, n)}
    if extra:
        f.update(extra)
    return px.ArrayDataSource(f, data_type="cells", domain_left_edge=[-500] * 3,
                              domain_right_edge=[500] * 3, chunk_size=chunk_size)


def ks_2samp_p(a, b):
    from scipy.stats import ks_2samp

    return ks_2samp(a, b).pvalue


def assert_distributions_agree(run, seeds=((17, 5),), pmin=0.01):
    """``run(seed_mine, seed_ref) -> {name: p-value}``: ONE run, one gate.  The family of m
    comparisons a test makes must hold at the north_star's level p > 0.01 jointly, i.e. every
    single comparison at the Bonferroni threshold 0.01 / m -- no retries, no second draw.  Power
    comes from the sample sizes the callers choose (>= 1e5 photons per comparison), not from
    repeated looks."""
    ps = run(*seeds[0])
    thr = pmin / max(len(ps), 1)
    bad = {k: v for k, v in ps.items() if not v > thr}
    assert not bad, f"distributions disagree (Bonferroni threshold {thr:.2e} over {len(ps)} comparisons): {bad}"
