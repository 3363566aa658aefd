#!/usr/bin/env python
"""bench.py -- photons generated + projected per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference]

Workload at N=1 (BASELINE.json configs[1]): synthetic uniform grid 256^3 thermal beta-model
cluster, APEC-shaped table 36 x 4000 bins 0.1-10 keV, 3-D Doppler velocities, wabs absorption
nH = 0.02, area*exposure scaled so that the expected photon count is 1e9.  A "step" is one
full pass: make_photons (expected counts -> Poisson -> offsets -> active-cell gather -> energy
sampling) followed by project_photons (Doppler, absorption rejection, scatter, TAN projection,
compaction).  For N>1 every rank owns a block-cyclic shard of the cells (weak scaling: each
rank gets its own 256^3-sized share of an N-times larger photon budget) and the only
collective is the count all-reduce.

Printed JSON (one line, rank 0): see the contract in the task statement -- value (device
resident), e2e (host buffers in, host events out), roofline of the dominant kernel,
cpu_baseline (the oracle on a bounded sample, 1 core), clocks, gpu_launches.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "photons generated+projected per second"
UNIT = "photons/s"
NBINS = 4000
NT = 36
GRID = 256
TARGET_PHOTONS = 1.0e9
REDSHIFT = 0.05
NH = 0.02
SEED_DATA, SEED_GEN, SEED_PROJ = 32, 50, 51


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=GRID)
    ap.add_argument("--photons", type=float, default=TARGET_PHOTONS)
    ap.add_argument("--nbins", type=int, default=NBINS)
    ap.add_argument("--kT-range", type=float, nargs=2, default=None,
                    help="log-uniform kT range instead of the isothermal 6 keV cluster")
    ap.add_argument("--traffic-from", default=None,
                    help="JSON summary of an ncu --set full capture of this same command")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=120000)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": f"synthetic uniform grid {args.grid}^3 thermal beta-model cluster "
                    f"(configs[1]), APEC-shaped table {NT}x{args.nbins} bins 0.1-10 keV, "
                    f"3-D Doppler velocities, wabs nH={NH}, {args.photons:.3g} expected photons "
                    f"per GPU, make_photons+project_photons (TAN sky)",
        "cells_per_gpu": args.grid ** 3,
        "expected_photons_per_gpu": args.photons,
        "kT": "6 keV isothermal" if args.kT_range is None else f"log-uniform {args.kT_range}",
        "sharding": "none" if n_gpus == 1 else f"block-cyclic cells over {n_gpus} ranks, counts all-reduce only",
        "l2": "inputs (1.2 GB cells + 8 GB photon list per step) far exceed the 126 MB L2",
    }


# -----------------------------------------------------------------------------------------
# clocks
# -----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# -----------------------------------------------------------------------------------------
# workload construction
# -----------------------------------------------------------------------------------------
def build_fields(args, device, rank=0):
    from pyxsim_b200 import synthetic

    return synthetic.beta_model_fields(args.grid, kT_keV=6.0, kT_range=args.kT_range,
                                       v3d_sigma_kms=300.0, bulk_kms=(100.0, -50.0, 200.0),
                                       seed=SEED_DATA + rank, device=device)


def build_tables(args):
    from pyxsim_b200 import synthetic

    return synthetic.apec_like_tables(nT=NT, nbins=args.nbins, emin=0.1, emax=10.0, seed=7)


def exposure_for(args, fields_em_sum, spec_sum_at_6keV, D_A_mpc):
    """area*exp_time such that sum(lambda) == args.photons (isothermal estimate; the realised
    count is reported)."""
    from pyxsim_b200 import constants as K

    D_A_cm = D_A_mpc * K.cm_per_mpc
    dist_fac = 1.0 / (4.0 * np.pi) / (D_A_cm ** 2 * (1 + REDSHIFT) ** 2)
    return args.photons / (fields_em_sum * spec_sum_at_6keV * dist_fac)


def mean_spec_sum(tables, kT_values, Z=0.3):
    kT, ebins, c, m, v = tables
    rs = c.sum(1) + Z * m.sum(1)
    return float(np.mean(np.interp(kT_values, kT, rs)))


# -----------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pyxsim_b200 as px
    from pyxsim_b200 import parallel, profiling

    parallel.init_from_env()
    rank, world = parallel.rank(), parallel.size()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())

    tables = build_tables(args)
    kT_nodes, ebins, cosmic, metal, _ = tables
    fields = build_fields(args, dev, rank)
    box = 2000.0
    cosmo = px.Cosmology()
    D_A = cosmo.angular_diameter_distance_mpc(0.0, REDSHIFT)
    if args.kT_range is None:
        ss = mean_spec_sum(tables, np.array([6.0]))
    else:
        ss = mean_spec_sum(tables, fields[("gas", "kT")][:200000].cpu().numpy())
    em_sum = float(fields[("gas", "emission_measure")].sum())
    area_time = exposure_for(args, em_sum, ss, D_A)
    exp_time, area = 1.0e5, area_time / 1.0e5

    def make_source(f):
        return px.ArrayDataSource(f, data_type="cells", domain_left_edge=[-0.5 * box] * 3,
                                  domain_right_edge=[0.5 * box] * 3, name=f"BetaModel{args.grid}^3",
                                  rank_local=True, global_offset=rank * args.grid ** 3)

    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, args.nbins, 0.3, temperature_field=("gas", "kT"),
                                  prng=SEED_GEN)
    src_dev = make_source(fields)
    pfx, efx = f"mem:bench_ph{rank}", f"mem:bench_ev{rank}"

    def step(src):
        n_ph, n_c = px.make_photons(pfx, src, REDSHIFT, area, exp_time, model)
        n_ev = px.project_photons(pfx, efx, "z", [30.0, 45.0], absorb_model="wabs", nH=NH,
                                  prng=SEED_PROJ)
        return n_ph, n_c, n_ev

    import logging

    logging.getLogger("pyxsim_b200").setLevel("WARNING")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    for _ in range(args.warmup):
        counts = step(src_dev)
    sync_all()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    profiling.enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        counts = step(src_dev)
    ev1.record()
    sync_all()
    clocks = sampler.stop()
    ms_local = ev0.elapsed_time(ev1)
    stages = profiling.summary()
    launches = profiling.launches()
    profiling.enable(False)
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    n_ph_all, n_c_all, n_ev_all = counts           # already summed over ranks
    ms_per_step = ms_total / args.steps
    value = n_ph_all / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
    lst = px.load_list(pfx)
    Np = int(lst.data["energy"].shape[0])
    Na = int(lst.data["num_photons"].shape[0])
    Ne_local = int(px.load_list(efx).data["eobs"].shape[0])
    Nc = args.grid ** 3
    w = 8
    alg = {   # algorithmic bytes per launch (DESIGN.md "Algorithmic bytes")
        "thermal_counts": Nc * 2 * w + Nc * 8,
        "scan_counts": Nc * 8 * 2 + Nc * 16,
        "compact_cells": Nc * 8 + Na * (7 * w) * 2 + Na * 24,
        "thermal_sample": Na * (2 * w + 24) + Np * w,
        "project": Na * (7 * w + 8) + Np * w + Ne_local * 3 * w,
    }
    # measured DRAM traffic of the dominant stage: only from an ncu capture of THIS command
    # handed in with --traffic-from (a JSON {"kernels": {name: {"dram_bytes": ...}}} made by
    # scripts/summarise_profile.py); otherwise null -- the ncu numbers live under profiles/
    traffic = None
    if args.traffic_from:
        try:
            kk = json.load(open(args.traffic_from))["kernels"]

            def dram(prefix):      # kernel names carry template arguments: match by prefix
                return sum(v["dram_bytes"] for k, v in kk.items() if k.startswith(prefix))

            traffic = {"project": dram("k_project_detect") + dram("k_project_emit"),
                       "thermal_sample": dram("k_thermal_sample") + dram("k_thermal_cellprep")}
        except Exception:
            traffic = None
    stage_ms = {k: v[1] / v[0] for k, v in stages.items()}
    dominant = max(stage_ms, key=stage_ms.get) if stage_ms else None
    roofline = None
    if dominant:
        achieved = alg[dominant] / (stage_ms[dominant] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": {"thermal_sample": "k_thermal_sample_cta",
                                               "project": "pxb_project: k_project_detect + scan + k_project_emit"}.get(dominant, dominant),
                    "achieved": round(achieved, 2), "peak": peak_gbs, "unit": "GB/s",
                    "frac": round(achieved / peak_gbs, 4),
                    "traffic": (traffic or {}).get(dominant),
                    "peak_source": peak_src, "ms_per_launch": round(stage_ms[dominant], 3),
                    "algorithmic_bytes_per_launch": int(alg[dominant]),
                    "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
                    "stage_frac_of_hbm_peak": {k: round(alg[k] / (v * 1e-3) / 1e9 / peak_gbs, 4)
                                               for k, v in stage_ms.items() if k in alg}}
    B_total = (Nc * 2 * w + Na * (14 * w + 8) + Np * w) + (Na * (7 * w + 8) + Np * w + Ne_local * 3 * w)
    whole = B_total / (ms_per_step * 1e-3) / 1e9

    # ---- end to end: host buffers in, host events out --------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, px, torch, dist, world, dev, fields, make_source, step, efx, sync_all)

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, fields, tables, area * exp_time, D_A, ncores=1)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "photons_per_step": n_ph_all, "active_cells": n_c_all, "events_per_step": n_ev_all,
            "whole_step_hbm_frac": round(whole / peak_gbs, 4),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, px, torch, dist, world, dev, fields, make_source, step, efx, sync_all):
    """Same step through the public API with HOST inputs (pinned NumPy-visible tensors) and the
    event list read back into pinned host memory, copies inside the timed region."""
    import psutil

    host_fields = {k: v.cpu().pin_memory() for k, v in fields.items()}
    src_host = make_source(host_fields)
    h2d = sum(v.numel() * v.element_size() for k, v in host_fields.items()
              if k != ("gas", "density"))
    # pinned landing buffers for the events, sized from a dry run.  When host memory allows (per
    # rank share) the whole event list lands in pinned memory; otherwise it is streamed through
    # a fixed pinned ring (every byte still crosses PCIe, a consumer would drain the ring).
    _ = step(src_host)
    ev = px.load_list(efx)
    n_ev = int(ev.data["eobs"].shape[0])
    full_cap = int(n_ev * 1.02 + 1024)
    budget = 0.25 * psutil.virtual_memory().available / max(world, 1)
    cap = full_cap if 3 * 8 * full_cap <= budget else max(1 << 20, int(budget // 24))
    cap = min(cap, full_cap)
    landing = {k: torch.empty(cap, dtype=torch.float64, pin_memory=True) for k in ("xsky", "ysky", "eobs")}
    n_steps = max(2, min(args.steps, 8))
    # the device->host copies of a step's events run on their own stream, so the next step's
    # upload and kernels overlap them (what a pipeline that consumes the events would do); every
    # step's events have fully landed before the clock stops
    copy_stream = torch.cuda.Stream()
    sync_all()
    t0 = time.perf_counter()
    tot_ph = 0
    d2h = 0
    for _ in range(n_steps):
        n_ph, n_c, n_e = step(src_host)
        ev = px.load_list(efx)
        m = int(ev.data["eobs"].shape[0])
        ready = torch.cuda.Event()
        ready.record()
        copy_stream.wait_event(ready)
        with torch.cuda.stream(copy_stream):
            for k in landing:
                src_t = ev.data[k]
                src_t.record_stream(copy_stream)       # keep the events alive until copied
                for a in range(0, m, cap):
                    b = min(a + cap, m)
                    landing[k][: b - a].copy_(src_t[a:b], non_blocking=True)
        d2h = 3 * 8 * m
        tot_ph = n_ph
    copy_stream.synchronize()
    sync_all()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    return {"value": tot_ph * n_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": n_steps,
            "host_landing": "full event list in pinned memory" if cap == full_cap else
            f"streamed through a {3 * 8 * cap / 2**30:.1f} GiB pinned ring",
            "note": "host cell arrays -> make_photons -> project_photons -> events in pinned host "
                    "memory; the photon list stays in HBM between the two stages (mem: prefix); a "
                    "step's event download overlaps the next step's upload and kernels"}


# -----------------------------------------------------------------------------------------
# CPU arm: the oracle (reference Cython + restated NumPy glue)
# -----------------------------------------------------------------------------------------
def _cpu_worker(job):
    (sel, kT, em, pos, vel, dx, tables, spectral_norm, D_A, seed) = job
    from oracle import ref_glue as O

    kT_nodes, ebins, cosmic, metal, _ = tables
    om = O.TableModel(ebins, kT_nodes, cosmic, metal)
    cfg = dict(kT_min=0.025, kT_max=64.0, Zmet=0.3, spectral_norm=spectral_norm)
    t0 = time.perf_counter()
    res = O.thermal_process_data(om, cfg, dict(kT=kT, em=em), np.random.RandomState(seed))
    t1 = time.perf_counter()
    if res is None:
        return 0, 0, t1 - t0, 0.0
    nc, nph, idx, energies = res
    g = O.gather_active(pos, vel, dx, idx, [-1000.0] * 3, [1000.0] * 3, [2000.0] * 3, [0.0] * 3,
                        [0.0] * 3, (False, False, False))
    d = dict(g, num_photons=nph, energy=energies)
    params = dict(data_type="cells", observer="external", fid_d_a=D_A, fid_redshift=REDSHIFT)
    t2 = time.perf_counter()
    ev = O.project_photons(d, params, "z", [30.0, 45.0], np.random.RandomState(seed + 1),
                           absorb_model=O.WabsAbsorb(NH), nH=NH)
    t3 = time.perf_counter()
    return int(nph.sum()), ev["n_events"], (t1 - t0), (t3 - t2)


def cpu_baseline(args, fields, tables, area_time, D_A, ncores=1, sample_cells=None):
    """Time the oracle on a uniform random subset of the workload's cells (with the photons
    those cells emit): per-cell and per-photon costs are shape independent, so photons/s of
    the sample estimates photons/s of the full workload."""
    from oracle import build_ref
    from pyxsim_b200 import constants as K

    if not build_ref.build(verbose=False):
        raise SystemExit("oracle/_ref is missing: the CPU baseline cannot run (build it in the "
                         "container that has /root/reference)")
    m = sample_cells or args.cpu_sample_cells
    ncell = args.grid ** 3
    rng = np.random.RandomState(12345)
    sel = np.sort(rng.choice(ncell, size=min(m * ncores, ncell), replace=False))

    def host(k):
        v = fields[k]
        return v[sel].cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)[sel]

    kT, em = host(("gas", "kT")), host(("gas", "emission_measure"))
    pos = [host(("index", a)) for a in "xyz"]
    vel = [host(("gas", f"velocity_{a}")) for a in "xyz"]
    dx = host(("index", "dx"))
    D_A_cm = D_A * K.cm_per_mpc
    spectral_norm = area_time / (4 * np.pi * D_A_cm ** 2 * (1 + REDSHIFT) ** 2)
    parts = np.array_split(np.arange(sel.size), ncores)
    jobs = [(sel[p], kT[p], em[p], [a[p] for a in pos], [a[p] for a in vel], dx[p], tables,
             spectral_norm, D_A, 1000 + j) for j, p in enumerate(parts)]
    t0 = time.perf_counter()
    if ncores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(ncores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    nph = sum(r[0] for r in res)
    return {"value": nph / wall, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": f"uniform random {sel.size} of {ncell} cells and the {nph} photons they emit "
                      f"({wall:.1f} s wall; generate {max(r[2] for r in res):.1f} s, project "
                      f"{max(r[3] for r in res):.1f} s); reference Cython natives compiled "
                      f"unmodified (oracle/_ref) + NumPy restatement of the Python glue",
            "cpu": _cpu_name()}


def _cpu_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores
    (chunk-parallel worker processes, like its MPI mode), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch

    import pyxsim_b200 as px
    from pyxsim_b200 import constants as K

    tables = build_tables(args)
    fields = build_fields(args, "cpu", 0)
    cosmo = px.Cosmology()
    D_A = cosmo.angular_diameter_distance_mpc(0.0, REDSHIFT)
    ss = mean_spec_sum(tables, np.array([6.0])) if args.kT_range is None else \
        mean_spec_sum(tables, fields[("gas", "kT")][:200000].numpy())
    area_time = exposure_for(args, float(fields[("gas", "emission_measure")].sum()), ss, D_A)
    cores = os.cpu_count() or 1
    per_core = max(2000, min(args.cpu_sample_cells // 4, 30000))
    vals, walls = [], []
    last = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = cpu_baseline(args, fields, tables, area_time, D_A, ncores=cores, sample_cells=per_core)
        if i >= args.warmup:
            vals.append(r["value"])
            walls.append(time.perf_counter() - t0)
        last = r
    v = float(np.mean(vals))
    last["value"] = v
    last["kind"] = "port"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": workload_config(args, world), "cpu_baseline": last,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def _protect_stdout():
    """Libraries (NCCL's version banner, for one) print to stdout; the driver wants exactly one
    JSON line there.  Point fd 1 at stderr for the run and keep the real stdout for the result."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    a = parse_args()
    _REAL_STDOUT = _protect_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
