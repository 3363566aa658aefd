"""Secondary timing of the other BASELINE.json configurations at single-GPU sizes (not the
headline bench: bench.py measures configs[1]).  Prints one JSON line per configuration with
per-stage device times, to catch pathologically slow paths (variable elements, 2-D tables,
SPH kernels, power-law/line samplers, multi-temperature data).

    python scripts/bench_configs.py [--scale 1.0]
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
from pyxsim_b200 import profiling, synthetic  # noqa: E402

logging.getLogger("pyxsim_b200").setLevel("WARNING")
TF = ("gas", "kT")


def timed(name, fn, reps=3):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    profiling.enable(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    st = profiling.summary()
    profiling.enable(False)
    ms = a.elapsed_time(b) / reps
    n_ph = out[0]
    print(json.dumps({"config": name, "photons": n_ph, "events": out[1], "ms_per_step": round(ms, 3),
                      "photons_per_s": n_ph / (ms * 1e-3),
                      "stage_ms": {k: round(v[1] / v[0], 3) for k, v in st.items()}}), flush=True)


def scale_exposure(src, model, target, z=0.05):
    """area*time such that the expected photon count is `target` (one cheap counts pass)."""
    model.setup_model("photons", src, z)
    lam, _ = model.expected_counts(next(src.chunks()), 1.0)
    return target / float(lam.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="", help="comma list of a,b,c,d,f,g (default: all)")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else set("abcdefg")
    s = args.scale
    dev = "cuda"
    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, 0.05) * 3.0856775809623245e24
    norm1 = 1.0 / (4 * np.pi * D_A ** 2 * 1.05 ** 2)

    def section_a():
        # ---- multi-temperature 256^3 (random kT per cell: no row coherence) -------------------
        n = int(256 * s ** (1 / 3))
        kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=4000)
        f = synthetic.beta_model_fields(n, kT_range=(0.5, 12.0), v3d_sigma_kms=300.0, device=dev)
        src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3)
        sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, 0.3, temperature_field=TF, prng=1)
        at = scale_exposure(src, model, 1.0e9 * s) / norm1

        def run_a():
            n_ph, _ = px.make_photons("mem:a", src, 0.05, at / 1e5, 1e5, model)
            n_ev = px.project_photons("mem:a", "mem:a_ev", [0.3, 0.2, 0.9], [30.0, 45.0],
                                      absorb_model="wabs", nH=0.02, prng=2)
            return n_ph, n_ev

        timed("thermal 256^3 random kT 0.5-12 keV, off-axis projection", run_a)

    if "a" in only:
        section_a()

    def section_b():
        # ---- AMR-like cell list with metallicity + 6 variable-element fields --------------------
        kT_nodes, ebins, cosmic, metal, var = synthetic.apec_like_tables(nT=36, nbins=4000, nvar=6)
        ncell = int(2.0e7 * s)
        f = synthetic.beta_model_fields(int(round(ncell ** (1 / 3))), kT_range=(1.0, 10.0),
                                        v3d_sigma_kms=300.0, device=dev, metallicity=(0.1, 0.6))
        names = ["O", "Ne", "Mg", "Si", "S", "Fe"]
        N = f[TF].numel()
        g = torch.Generator(device=dev).manual_seed(3)
        for el in names:
            f[("gas", f"{el}_metallicity")] = 0.1 + 0.8 * torch.rand(N, generator=g, dtype=torch.float64, device=dev)
        src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3)
        sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal, var, var_elem_names=names)
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, ("gas", "metallicity"), temperature_field=TF,
                                      var_elem={el: ("gas", f"{el}_metallicity") for el in names}, prng=4)
        at = scale_exposure(src, model, 5.0e8 * s) / norm1

        def run_b():
            n_ph, _ = px.make_photons("mem:b", src, 0.05, at / 1e5, 1e5, model)
            n_ev = px.project_photons("mem:b", "mem:b_ev", "z", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=5)
            return n_ph, n_ev

        timed(f"thermal {N} cells, Z field + 6 var_elem fields", run_b)

    if "b" in only:
        section_b()

    def section_c():
        # ---- SPH particles, IGM (kT, nH) table + CIE fallback, log bins, gaussian kernel -------
        npart = int(2 ** 24 * s)
        f = synthetic.particle_beta_model_fields(npart, device=dev)
        Tv, Dv, ebins2, c2, m2, v2, cie = synthetic.pion_like_tables(nT=24, nD=12, nbins=4000)
        cie_T, cc, cm, cv = cie
        cie_prod = px.ThermalSpectralModel(ebins2, cie_T, cc, cm, cv, logT=True, binscale="log")
        prod = px.PionSpectralModel(ebins2, Tv, Dv, c2, m2, v2, cie_model=cie_prod, binscale="log")
        src = px.ArrayDataSource(f, data_type="particles", domain_left_edge=[-1000.0] * 3,
                                 domain_right_edge=[1000.0] * 3)
        model = px.IGMSourceModel(0.3, 1.4, 4000, 0.3, binscale="log", temperature_field=TF,
                                  spectral_model=prod, prng=6)
        at = scale_exposure(src, model, 5.0e8 * s) / norm1

        def run_c():
            n_ph, _ = px.make_photons("mem:c", src, 0.05, at / 1e5, 1e5, model)
            n_ev = px.project_photons("mem:c", "mem:c_ev", [0.1, 0.5, 0.8], [30.0, 45.0], absorb_model="wabs",
                                      nH=0.01, kernel="gaussian", prng=7)
            return n_ph, n_ev

        timed(f"IGM model, {npart} SPH particles, gaussian kernel, log bins", run_c)

    if "c" in only:
        section_c()

    def section_d():
        # ---- power law + line on 256^3, z = 0.5 ---------------------------------------------------
        n = int(256 * s ** (1 / 3))
        f = synthetic.beta_model_fields(n, v3d_sigma_kms=300.0, device=dev)
        N = n ** 3
        g = torch.Generator(device=dev).manual_seed(8)
        alpha = 1.0 + 1.5 * torch.rand(N, generator=g, dtype=torch.float64, device=dev)
        alpha[::7] = 1.0
        alpha[1::7] = 2.0
        f[("gas", "alpha")] = alpha
        f[("gas", "lum")] = f[("gas", "emission_measure")].clone()
        f[("gas", "line")] = f[("gas", "emission_measure")].clone()
        src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3,
                                 units={("gas", "lum"): "keV/s"})
        pl = px.PowerLawSourceModel(1.0, 0.5, 40.0, ("gas", "lum"), ("gas", "alpha"), prng=9)
        at_pl = scale_exposure(src, pl, 1.0e9 * s, z=0.5)
        ln = px.LineSourceModel(3.5, ("gas", "line"), (1000.0, "km/s"), prng=10)
        at_ln = scale_exposure(src, ln, 5.0e8 * s, z=0.5)
        D5 = px.Cosmology().angular_diameter_distance_mpc(0.0, 0.5) * 3.0856775809623245e24
        norm5 = 1.0 / (4 * np.pi * D5 ** 2 * 1.5 ** 2)

        def run_d():
            n_ph, _ = px.make_photons("mem:d", src, 0.5, at_pl / norm5 / 1e5, 1e5, pl)
            n_ev = px.project_photons("mem:d", "mem:d_ev", "x", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=11)
            return n_ph, n_ev

        def run_e():
            n_ph, _ = px.make_photons("mem:e", src, 0.5, at_ln / norm5 / 1e5, 1e5, ln)
            n_ev = px.project_photons("mem:e", "mem:e_ev", "x", [30.0, 45.0], absorb_model="wabs", nH=0.02, prng=12)
            return n_ph, n_ev

        timed("power law (alpha field incl. 1 and 2) 256^3, z=0.5", run_d)
        timed("line 3.5 keV sigma 1000 km/s 256^3, z=0.5", run_e)

    if "d" in only:
        section_d()

    def section_f():
        # ---- spectrum mode (row f1): Doppler-regridded thermal spectrum of a whole dataset --------
        import time

        n = int(128 * s ** (1 / 3))
        kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=4000)
        f = synthetic.beta_model_fields(n, kT_range=(1.0, 10.0), v3d_sigma_kms=300.0, device=dev)
        src = px.ArrayDataSource(f, domain_left_edge=[-1000.0] * 3, domain_right_edge=[1000.0] * 3)
        sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, 0.3, temperature_field=TF, prng=1)
        for shifting, normal in ((False, None), (True, [0.3, -0.5, 0.8])):
            model.make_spectrum(src, 0.2, 9.0, 2000, redshift=0.05, normal=normal)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                model.make_spectrum(src, 0.2, 9.0, 2000, redshift=0.05, normal=normal)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            ncell = n ** 3
            print(json.dumps({"config": f"make_spectrum thermal {n}^3 cells, 4000 -> 2000 bins, shifting={shifting}",
                              "cells": ncell, "ms": round(ms, 2), "cells_per_s": ncell / (ms * 1e-3)}), flush=True)
        # Doppler-shifted band per cell (make_band, the integrand of make_intensity_fields)
        model._spec_normal = np.array([0.3, -0.5, 0.8])
        model.v_fields = src.velocity_fields
        chunk = next(src.chunks())
        for label, env in (("O(1) prefix sums", None), ("per-cell sum", "1")):
            if env:
                os.environ["PXB_BAND_PER_CELL"] = env
            model.band_per_cell(chunk, 1.0, 0.5, 2.0, energy=True, shifting=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                model.band_per_cell(chunk, 1.0, 0.5, 2.0, energy=True, shifting=True)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            os.environ.pop("PXB_BAND_PER_CELL", None)
            print(json.dumps({"config": f"shifted intensity band per cell, {label}", "cells": n ** 3,
                              "ms": round(ms, 3), "cells_per_s": n ** 3 / (ms * 1e-3)}), flush=True)
        # CPU reference on a sample: compiled reference shift_spectrum + interp1d_spec via the oracle
        try:
            from oracle import ref_glue as O

            m_cpu = 4000
            kT = f[TF][:m_cpu].cpu().numpy()
            em = f[("gas", "emission_measure")][:m_cpu].cpu().numpy()
            om = O.TableModel(ebins, kT_nodes, cosmic, metal)
            cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=1.0)
            t0 = time.perf_counter()
            O.thermal_spectrum_mode(om, cfg, dict(kT=kT, em=em), np.linspace(0.2, 9.0, 2001), shift=np.full(m_cpu, 1.001))
            dt = time.perf_counter() - t0
            print(json.dumps({"config": "make_spectrum CPU reference (oracle, 1 core)", "cells": m_cpu,
                              "ms": round(dt * 1e3, 1), "cells_per_s": m_cpu / dt}), flush=True)
        except Exception as e:  # the oracle is optional for this script
            print(json.dumps({"config": "make_spectrum CPU reference", "error": str(e)}))

    if "f" in only:
        section_f()

    def section_g():
        # ---- event-list consumers (row f4): image + spectrum binning of a large event list ---------
        import time

        n = int(3.0e8 * s)
        g = torch.Generator(device=dev).manual_seed(21)
        ra = 30.0 + 0.3 * torch.randn(n, generator=g, dtype=torch.float64, device=dev)
        dec = 45.0 + 0.2 * torch.randn(n, generator=g, dtype=torch.float64, device=dev)
        e = torch.exp(torch.empty(n, dtype=torch.float64, device=dev).uniform_(np.log(0.1), np.log(10.0), generator=g))
        from pyxsim_b200.photon_list import DeviceList, save_list

        save_list("mem:g_ev", DeviceList({}, dict(exp_time=1e5, area=1000.0, sky_center=np.array([30.0, 45.0])),
                                         {"xsky": ra, "ysky": dec, "eobs": e}))
        ev = px.EventList("mem:g_ev")
        for label, fn in (("image 2048^2, 0.5-7 keV", lambda: ev.image(60.0, 2048, 0.5, 7.0)),
                          ("spectrum 4096 channels", lambda: ev.spectrum(0.1, 10.0, 4096))):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            print(json.dumps({"config": f"EventList {label}", "events": n, "ms": round(ms, 2),
                              "events_per_s": n / (ms * 1e-3)}), flush=True)
        try:   # CPU reference on a sample: what event_list.py does with NumPy (WCS restated)
            from oracle import ref_glue as O

            m = 5_000_000
            xs, ys, es = ra[:m].cpu().numpy(), dec[:m].cpu().numpy(), e[:m].cpu().numpy()
            t0 = time.perf_counter()
            O.event_image(xs, ys, es, [30.0, 45.0], 60.0, 2048, 0.5, 7.0)
            t1 = time.perf_counter()
            O.event_spectrum(es, 0.1, 10.0, 4096)
            t2 = time.perf_counter()
            print(json.dumps({"config": "EventList CPU (NumPy histogram2d / histogram, 1 core)", "events": m,
                              "image_events_per_s": m / (t1 - t0), "spectrum_events_per_s": m / (t2 - t1)}), flush=True)
        except Exception as ex:
            print(json.dumps({"config": "EventList CPU", "error": str(ex)}))

    if "g" in only:
        section_g()

if __name__ == "__main__":
    main()
