"""Timing of the spectrum mode on one GPU: make_spectrum of a thermal source
with and without Doppler shifts (4000 model bins -> 2000 output bins), device time per pass from
CUDA events, and the two routes of the shifted case (cells bucketed by table row / one CTA per
cell) against each other.  Prints one JSON line per case.

    python scripts/bench_spectrum.py [--n 128] [--reps 5]
"""
import argparse
import json
import logging
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyxsim_b200 as px  # noqa: E402
from pyxsim_b200 import synthetic  # noqa: E402

logging.getLogger("pyxsim_b200").setLevel("WARNING")
TF = ("gas", "kT")


def timed(fn, reps):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.n
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=4000)
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    results = {}
    for label, fkw, Zmet in (("kT 1-10 keV, constant Z", dict(kT_range=(1.0, 10.0)), 0.3),
                             ("isothermal 6 keV, constant Z", dict(kT_keV=6.0), 0.3)):
        f = synthetic.beta_model_fields(n, v3d_sigma_kms=300.0, device="cuda", **fkw)
        src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3)
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, Zmet, temperature_field=TF, prng=1)
        for shifting, normal, env in ((False, None, None), (True, [0.3, -0.5, 0.8], None),
                                      (True, [0.3, -0.5, 0.8], "PXB_SPECTRUM_PER_CELL")):
            if env:
                os.environ[env] = "1"
            try:
                (eb, spec), ms = timed(lambda: model.make_spectrum(src, 0.2, 9.0, 2000, redshift=0.05,
                                                                   normal=normal), args.reps)
            finally:
                if env:
                    os.environ.pop(env, None)
            route = "per cell" if env else ("bucketed by table row" if shifting else "row weights")
            key = (label, shifting)
            line = {"config": f"make_spectrum thermal {n}^3 cells ({label}), 4000 -> 2000 bins",
                    "shifting": shifting, "route": route, "cells": n ** 3, "ms": round(ms, 3),
                    "cells_per_s": n ** 3 / (ms * 1e-3), "timing": "CUDA events around the public call"}
            if key in results:
                ref = results[key]
                line["max_rel_diff_vs_other_route"] = float(np.max(np.abs(spec - ref) / np.max(np.abs(ref))))
            else:
                results[key] = np.asarray(spec)
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
