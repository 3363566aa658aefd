"""Generate tests/golden/hotpath_golden.npz: small input/output vectors of the reference's
native routines (compiled from /root/reference into oracle/_ref) and of the restated glue on
seeded synthetic inputs.  Run in the build container:  python tests/golden/make_golden.py

The GPU parity tests replay these on the B200 box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def compute(oracle):
    import helpers as H

    out = {}
    # --- interp1d_spec / interp2d_spec -------------------------------------------------
    t = H.make_tables(nbins=64, nvar=2)
    kT, em = H.random_cells(24, seed=101, kT_range=(0.01, 90.0))
    c, m, v = H.oracle_model(oracle, t).get_spectrum(kT)
    out.update(i1_kT=kT, i1_c=c, i1_m=m, i1_v=v)
    # --- thermal lambda + spectra ---------------------------------------------------------
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=2e-44,
               var_elem=[(0.5, 1.0, False), (1.5, 1.0, False)])
    cut, tot, lam = oracle.thermal_spectra(H.oracle_model(oracle, t), cfg, dict(kT=kT, em=em))
    out.update(th_em=em, th_cut=cut, th_tot=tot, th_lam=lam)
    # --- doppler + pixel_to_cel ---------------------------------------------------------------
    rng = np.random.RandomState(5)
    nph = rng.poisson(4, 50).astype(np.int64)
    vn, v2 = rng.normal(0, 800, 50), rng.uniform(0, 2e6, 50)
    e = rng.uniform(0.1, 10, int(nph.sum()))
    e1 = e.copy()
    sf = oracle.ref()["sky_functions"]
    sf.doppler_shift(vn * oracle.scale_shift, v2 * oracle.scale_shift2, nph, e1)
    out.update(dp_nph=nph, dp_vn=vn, dp_v2=v2, dp_e=e, dp_out=e1)
    x, y = rng.uniform(-5e-3, 5e-3, 200), rng.uniform(-5e-3, 5e-3, 200)
    x1, y1 = x.copy(), y.copy()
    sf.pixel_to_cel(x1, y1, np.array([30.0, 45.0]))
    out.update(tan_x=x, tan_y=y, tan_ra=x1, tan_dec=y1)
    # --- absorption ---------------------------------------------------------------------------
    from pyxsim_b200 import synthetic

    en, sg = synthetic.tbabs_like_sigma(n=500)
    ee = np.exp(rng.uniform(np.log(0.02), np.log(20), 300))
    out.update(ab_e=ee, ab_tab=oracle.TableAbsorb(0.1, en, sg * 1.0e22).get_absorb(ee),
               ab_wabs=oracle.WabsAbsorb(0.1).get_absorb(ee))
    # --- power law / line expected counts ------------------------------------------------------
    alpha = np.concatenate([[1.0, 2.0], rng.uniform(0.5, 2.5, 30)])
    lum = 1e45 * rng.lognormal(0, 1, 32)
    pl, _, _ = oracle.powerlaw_lambda(dict(e0=1.0, emin=0.5, emax=40.0, spectral_norm=3e-44,
                                           scale_factor=1 / 1.5), lum, alpha)
    out.update(pl_alpha=alpha, pl_lum=lum, pl_lam=pl)
    return out


if __name__ == "__main__":
    from oracle import build_ref, ref_glue

    assert build_ref.build(), "needs /root/reference (or a prebuilt oracle/_ref)"
    np.savez(os.path.join(HERE, "hotpath_golden.npz"), **compute(ref_glue))
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"))
