"""GPU: parity on the configuration that bench.py measures.

The exact bench inputs -- APEC-shaped table 36 x 4000 bins, isothermal 6 keV beta-model cluster
with hashed 3-D velocities, wabs nH = 0.02, projection along z onto a TAN sky -- on the central
28^3 block of the 256^3 grid, at an exposure the oracle can follow.  The kernels compiled for the
common option set, with the table rows staged in shared memory, must be the ones that run, and
their output must agree with the oracle: expected counts to 1e-12, the sampled spectrum by
chi^2 against the oracle's exact per-cell spectrum, events by KS and binomial counts.
Reference acceptance style: pyxsim/tests/test_beta_model.py:28-122.
"""
import numpy as np
import pytest
import torch
from scipy import stats

import pyxsim_b200 as px
from pyxsim_b200 import _lib, synthetic

import helpers as H

pytestmark = pytest.mark.gpu
TF = ("gas", "kT")
NB, Z, NH = 4000, 0.05, 0.02


def _bench_block(n_sub=28, n=256):
    lo = (n - n_sub) // 2
    ax = torch.arange(lo, lo + n_sub, dtype=torch.int64)
    idx = (ax[:, None, None] * n * n + ax[None, :, None] * n + ax[None, None, :]).reshape(-1)
    return synthetic.cluster_chain_fields(idx.cuda(), n, n_clusters=1, kT_keV=6.0, seed=32)


def test_bench_configuration_against_the_oracle(cuda, oracle):
    f = _bench_block()
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=NB, emin=0.1, emax=10.0, seed=7)
    src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3)
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, Z)
    dfac = 1.0 / (4 * np.pi * (D_A * 3.0856775809623245e24) ** 2 * (1 + Z) ** 2)
    fh = {k: v.cpu().numpy() for k, v in f.items()}
    om = oracle.TableModel(ebins, kT_nodes, cosmic, metal)
    a = dict(kT=fh[TF], em=fh[("gas", "emission_measure")])
    cut, tot, lam1 = oracle.thermal_spectra(om, dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=1.0), a)
    area_time = 2.5e6 / (lam1.sum() * dfac)          # ~2.5e6 photons from the block
    texp, area = 1.0e5, area_time / 1.0e5
    norm = area_time * dfac
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=norm)

    # expected counts: 1e-12 (BASELINE north_star, fp64)
    m0 = px.ThermalSourceModel(sm, 0.1, 10.0, NB, 0.3, temperature_field=TF, prng=50)
    m0.setup_model("photons", src, Z)
    lam, _ = m0.expected_counts(next(src.chunks()), norm)
    cut, tot, rlam = oracle.thermal_spectra(om, cfg, a)
    np.testing.assert_allclose(lam.cpu().numpy()[cut], rlam, rtol=1e-12, atol=0)

    def run(seed_mine, seed_ref):
        model = px.ThermalSourceModel(sm, 0.1, 10.0, NB, 0.3, temperature_field=TF, prng=seed_mine)
        n_ph, n_c = px.make_photons("mem:bs_ph", src, Z, area, texp, model)
        assert _lib.last_pipeline_launch() == (0, 1, 1)       # sampler, common option set, staged rows
        n_ev = px.project_photons("mem:bs_ph", "mem:bs_ev", "z", [30.0, 45.0], absorb_model="wabs", nH=NH,
                                  prng=seed_mine + 1)
        assert _lib.last_pipeline_launch() == (1, 1, 0)       # projection, common option set, on-axis
        ph, ev = px.load_list("mem:bs_ph"), px.load_list("mem:bs_ev")
        model = px.ThermalSourceModel(sm, 0.1, 10.0, NB, 0.3, temperature_field=TF, prng=seed_mine)
        tot1 = px.make_events("mem:bs_ev1", src, Z, area, texp, model, "z", [30.0, 45.0], absorb_model="wabs",
                              nH=NH, prng=seed_mine + 1)
        assert _lib.last_pipeline_launch() == (2, 1, 0)       # fused, common option set (small CTAs, rows via L1)
        ev1 = px.load_list("mem:bs_ev1")
        assert tot1 == (n_ph, n_c, n_ev)
        for k in ("xsky", "ysky", "eobs"):
            assert torch.equal(ev1.data[k], ev.data[k]), k

        # counts
        assert abs(n_ph - rlam.sum()) < 5 * np.sqrt(rlam.sum())
        # the sampled spectrum against the oracle's exact one: every cell is at 6 keV, so all
        # photons share the normalised spectrum tot / sum(tot); chi^2 over the 4000 bins (bins with
        # fewer than 25 expected photons merged with their neighbours)
        e = ph.numpy("energy")
        p = tot[0] / tot[0].sum()
        obs = np.histogram(e, bins=ebins)[0].astype(float)
        exp = p * e.size
        grp = np.cumsum(exp) // 25.0
        o = np.bincount(grp.astype(int), weights=obs)
        x = np.bincount(grp.astype(int), weights=exp)
        keep = x > 0
        chi2 = ((o[keep] - x[keep]) ** 2 / x[keep]).sum()
        out = {"spectrum_chi2": stats.chi2.sf(chi2, keep.sum() - 1)}
        # uniform inside a bin (np.interp inversion, base.py:480-482): the position inside the bin
        frac = (e - ebins[np.clip(np.searchsorted(ebins, e, side="right") - 1, 0, NB - 1)]) / np.diff(ebins)[0]
        out["within_bin"] = stats.kstest(frac[:400000], "uniform").pvalue
        # the oracle's own photons and events
        nc, nph, idx, oe = oracle.thermal_process_data(om, cfg, a, np.random.RandomState(seed_ref))
        out["energy_ks"] = stats.ks_2samp(e[:1000000], oe[:1000000]).pvalue
        g = oracle.gather_active([fh[("index", ax)] for ax in "xyz"], [fh[("gas", f"velocity_{ax}")] for ax in "xyz"],
                                 fh[("index", "dx")], idx, [-1000.0] * 3, [1000.0] * 3, [2000.0] * 3, [0.0] * 3,
                                 [0.0] * 3, (False,) * 3)
        d = dict(g, num_photons=nph, energy=oe)
        params = dict(data_type="cells", observer="external", fid_d_a=D_A, fid_redshift=Z)
        ref = oracle.project_photons(d, params, "z", [30.0, 45.0], np.random.RandomState(seed_ref + 1),
                                     absorb_model=oracle.WabsAbsorb(NH), nH=NH)
        fm, fr = n_ev / n_ph, ref["n_events"] / nph.sum()
        assert abs(fm - fr) < 5 * np.sqrt(2 * fr * (1 - fr) / nph.sum())
        for k in ("eobs", "xsky", "ysky"):
            out[f"{k}_ks"] = stats.ks_2samp(ev.numpy(k)[:1000000], ref[k][:1000000]).pvalue
        return out

    H.assert_distributions_agree(run, seeds=((50, 7),))
