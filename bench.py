#!/usr/bin/env python
"""bench.py -- photons generated + projected per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--config NAME]

Default workload (BASELINE.json configs[1], the one the metric is quoted on): synthetic uniform
grid 256^3 thermal beta-model cluster, APEC-shaped table 36 x 4000 bins 0.1-10 keV, 3-D Doppler
velocities, wabs absorption nH = 0.02, area*exposure scaled so that the expected photon count is
1e9.  A "step" is one full pass through the reference's API: make_photons (expected counts ->
Poisson -> offsets -> active-cell gather -> energy sampling) followed by project_photons (Doppler,
absorption rejection, scatter, TAN projection, compaction).  The same JSON line also reports the
step done by the fused `make_events` (no photon list in between; same events bit for bit).

N > 1: ONE dataset is split block-cyclically over the ranks (blocks of 4096 cells -- one row of
cells for the uniform grids --, every rank
materialises only the cells it owns, RNG counters are keyed by global cell ids so the union of
the per-rank lists is the single-rank list); the only collectives are the scalar count
reductions.  For the default workload the dataset grows with N (a chain of N clusters: weak
scaling).  --config amr | sph | plline run BASELINE configs 3 / 4 / 5 at their named sizes.

Printed JSON (one line, rank 0): value (device resident), e2e (host buffers in, host events
out), roofline of the dominant kernel, cpu_baseline (the oracle on a bounded sample, 1 core),
clocks, gpu_launches (counted inside libpxb).
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
SHARD_BLOCK = 4096
SEED_DATA, SEED_GEN, SEED_PROJ = 32, 50, 51


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cluster", choices=["cluster", "amr", "sph", "plline"])
    ap.add_argument("--grid", type=int, default=None,
                    help="cells per side (cluster: 256, plline: 512, amr: 320 per level)")
    ap.add_argument("--particles", type=float, default=2.0 ** 27)
    ap.add_argument("--photons", type=float, default=None,
                    help="expected photons (cluster / plline: per GPU; amr / sph: in total)")
    ap.add_argument("--nbins", type=int, default=NBINS)
    ap.add_argument("--kT-range", type=float, nargs=2, default=None,
                    help="log-uniform kT range instead of the isothermal 6 keV cluster")
    ap.add_argument("--traffic-from", default=None,
                    help="JSON summary of an ncu --set full capture of this same command")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=120000)
    a = ap.parse_args()
    if a.grid is None:
        a.grid = {"cluster": GRID, "plline": 512, "amr": 320, "sph": 0}[a.config]
    if a.photons is None:
        a.photons = {"cluster": TARGET_PHOTONS, "plline": 1.25e9, "amr": 4.0e9, "sph": 4.0e9}[a.config]
    return a


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
# workloads: every rank materialises its block-cyclic share of ONE global dataset
# -----------------------------------------------------------------------------------------
class Workload:
    """fields: this rank's share (device tensors); make_source(fields) -> ArrayDataSource;
    step(src) / step_fused(src) -> (n_photons, n_cells, n_events) summed over ranks."""
    step_fused = None
    step_e2e = None  # make_events where it is not the headline (it runs the two stages itself)
    cpu = None       # (tables, area_time, D_A) for the oracle arm (cluster only)


def _grid_block(n):
    """Shard block of a uniform n^3 grid: one row of cells.  A 4096-cell block is 16 rows of a
    256^3 grid, i.e. 125 kpc slabs dealt out cyclically: with 4 or 8 ranks the slabs through
    the cluster core all land on two ranks (measured: 37.7 ms vs 23.0 ms per step at N = 4)."""
    return min(SHARD_BLOCK, n)


def _source(px, fields, world, rank, name, data_type="cells", box=2000.0, n_clusters=1,
            block=SHARD_BLOCK):
    le = [-0.5 * box * n_clusters, -0.5 * box, -0.5 * box]
    return px.ArrayDataSource(fields, data_type=data_type, domain_left_edge=le,
                              domain_right_edge=[-v for v in le], name=name,
                              shard_of=(block, rank, world) if world > 1 else None)


def _global_sum(dist, t, world):
    v = t.sum().reshape(1).clone()
    if world > 1:
        dist.all_reduce(v)
    return float(v.item())


def build_tables(args, nvar=0):
    from pyxsim_b200 import synthetic

    return synthetic.apec_like_tables(nT=NT, nbins=args.nbins, emin=0.1, emax=10.0, nvar=nvar, seed=7)


def mean_spec_sum(tables, kT_values, Z=0.3):
    kT, ebins, c, m, v = tables
    rs = c.sum(1) + Z * m.sum(1)
    return float(np.mean(np.interp(kT_values, kT, rs)))


def dist_fac(px, z):
    from pyxsim_b200 import constants as K

    D_A = px.Cosmology().angular_diameter_distance_mpc(0.0, z)
    D_A_cm = D_A * K.cm_per_mpc
    return D_A, 1.0 / (4.0 * np.pi) / (D_A_cm ** 2 * (1 + z) ** 2)


def _exposure_from_probe(px, torch, dist, world, model, make_source, fields, z, weight_field, target):
    """area*exp_time such that the expected photon count is ``target``: the model's own expected
    counts on a probe of this rank's cells at spectral_norm = 1, scaled by the global sum of the
    field the counts are proportional to."""
    n = int(next(iter(fields.values())).shape[0])
    stride = max(1, n // 262144)
    probe = make_source({k: v[::stride].contiguous() for k, v in fields.items()})   # a uniform sample of the cells
    probe.shard_of = None
    model.setup_model("photons", probe, z)
    lam, _ = model.expected_counts(next(probe.chunks()), 1.0)
    ratio = torch.nan_to_num(lam).sum().reshape(1) / probe.fields[weight_field].sum().reshape(1)
    if world > 1:
        dist.all_reduce(ratio)
        ratio /= world
    _, dfac = dist_fac(px, z)
    return target / (float(ratio.item()) * _global_sum(dist, fields[weight_field], world) * dfac)


def wl_cluster(args, px, torch, dist, dev, rank, world):
    """configs[1]; N > 1: a chain of N such clusters in one grid, split block-cyclically."""
    from pyxsim_b200 import synthetic

    W = Workload()
    n = args.grid
    ncell = world * n ** 3
    idx = synthetic.shard_indices(ncell, _grid_block(n), rank, world, device=dev)
    W.fields = synthetic.cluster_chain_fields(idx, n, n_clusters=world, kT_keV=6.0, kT_range=args.kT_range,
                                              seed=SEED_DATA)
    del idx
    W.n_cells_global = ncell
    tables = build_tables(args)
    kT_nodes, ebins, cosmic, metal, _ = tables
    D_A, dfac = dist_fac(px, REDSHIFT)
    if args.kT_range is None:
        ss = mean_spec_sum(tables, np.array([6.0]))
    else:
        ss = mean_spec_sum(tables, W.fields[("gas", "kT")][:200000].cpu().numpy())
    em_sum = _global_sum(dist, W.fields[("gas", "emission_measure")], world)
    area_time = args.photons * world / (em_sum * ss * dfac)
    exp_time, area = 1.0e5, area_time / 1.0e5
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    model = px.ThermalSourceModel(sm, 0.1, 10.0, args.nbins, 0.3, temperature_field=("gas", "kT"), prng=SEED_GEN)
    W.make_source = lambda f: _source(px, f, world, rank, f"BetaModel{n}^3 x{world}", n_clusters=world,
                                      block=_grid_block(n))
    pfx, efx = f"mem:bench_ph{rank}", f"mem:bench_ev{rank}"
    W.pfx, W.efx = pfx, efx

    def step(src):
        n_ph, n_c = px.make_photons(pfx, src, REDSHIFT, area, exp_time, model)
        n_ev = px.project_photons(pfx, efx, "z", [30.0, 45.0], absorb_model="wabs", nH=NH, prng=SEED_PROJ)
        return n_ph, n_c, n_ev

    def step_fused(src):
        return px.make_events(efx, src, REDSHIFT, area, exp_time, model, "z", [30.0, 45.0],
                              absorb_model="wabs", nH=NH, prng=SEED_PROJ)

    W.step, W.step_fused = step, step_fused
    W.step_fused_f32 = lambda src: px.make_events(efx, src, REDSHIFT, area, exp_time, model, "z", [30.0, 45.0],
                                                  absorb_model="wabs", nH=NH, prng=SEED_PROJ, event_dtype="float32")
    W.cpu = (tables, area_time / world, D_A)
    W.nZ, W.nD, W.h2d_skip = 0, 0, [("gas", "density")]
    W.desc = (f"synthetic uniform grid {n}^3 thermal beta-model cluster (configs[1]"
              + (f"; a chain of {world} clusters, {world}x{n}x{n}x{n} cells, in ONE grid" if world > 1 else "")
              + f"), APEC-shaped table {NT}x{args.nbins} bins 0.1-10 keV, 3-D Doppler velocities, wabs nH={NH}, "
              f"{args.photons:.3g} expected photons per GPU, make_photons+project_photons (TAN sky)")
    W.scaling = "weak"
    return W


def wl_amr(args, px, torch, dist, dev, rank, world):
    """configs[2]: 4-level nested AMR cluster, ~1.2e8 leaf cells, metallicity field + 6 variable
    element fields; one dataset, sharded."""
    from pyxsim_b200 import synthetic

    W = Workload()
    n, nvar = args.grid, 6
    ncell = synthetic.amr_leaf_count(n)
    idx = synthetic.shard_indices(ncell, SHARD_BLOCK, rank, world, device=dev)
    W.fields = synthetic.amr_cluster_fields(idx, n=n, nvar=nvar)
    del idx
    W.n_cells_global = ncell
    tables = build_tables(args, nvar=nvar)
    kT_nodes, ebins, cosmic, metal, var = tables
    names = ("O", "Ne", "Mg", "Si", "S", "Fe")[:nvar]
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal, var, var_elem_names=names)
    ve = {nm: ("gas", f"El{k}_fraction") for k, nm in enumerate(names)}
    model = px.ThermalSourceModel(sm, 0.1, 10.0, args.nbins, ("gas", "metallicity"), var_elem=ve,
                                  temperature_field=("gas", "kT"), prng=SEED_GEN)
    units = {("gas", f"El{k}_fraction"): "Zsun" for k in range(nvar)}
    units[("gas", "metallicity")] = "Zsun"

    def make_source(f):
        s = _source(px, f, world, rank, f"AMR4x{n}^3")
        s.units.update(units)
        return s

    W.make_source = make_source
    area_time = _exposure_from_probe(px, torch, dist, world, model, make_source, W.fields, REDSHIFT,
                                     ("gas", "emission_measure"), args.photons)
    exp_time, area = 1.0e5, area_time / 1.0e5
    pfx, efx = f"mem:bench_ph{rank}", f"mem:bench_ev{rank}"
    W.pfx, W.efx = pfx, efx
    normal = [0.3, -0.2, 0.93]

    def step(src):
        n_ph, n_c = px.make_photons(pfx, src, REDSHIFT, area, exp_time, model)
        n_ev = px.project_photons(pfx, efx, normal, [30.0, 45.0], absorb_model="wabs", nH=NH, prng=SEED_PROJ)
        return n_ph, n_c, n_ev

    def step_fused(src):
        return px.make_events(efx, src, REDSHIFT, area, exp_time, model, normal, [30.0, 45.0],
                              absorb_model="wabs", nH=NH, prng=SEED_PROJ)

    # (variable elements: option-generic sampler -- make_events itself takes the two stages there,
    # they are faster; the headline is the two-stage step and the e2e leg calls make_events)
    W.step, W.step_e2e = step, step_fused
    W.nZ, W.nD, W.h2d_skip = 1 + nvar, 0, [("gas", "density")]
    W.desc = (f"synthetic 4-level nested AMR beta-model cluster, {ncell:.4g} leaf cells (4 x {n}^3 patches, "
              f"configs[2]), thermal + metallicity field + {nvar} var_elem fields, table {NT}x{args.nbins}x{2 + nvar}, "
              f"random kT 2-9 keV, off-axis normal, wabs nH={NH}, {args.photons:.3g} expected photons in total")
    W.scaling = "strong"
    return W


def wl_sph(args, px, torch, dist, dev, rank, world):
    """configs[3]: SPH / Gadget-like particles, IGM-style 2-D (kT, nH) table with CIE fallback,
    SPH-kernel (gaussian) position smoothing; one dataset, sharded."""
    from pyxsim_b200 import synthetic

    W = Workload()
    npart = int(args.particles)
    idx = synthetic.shard_indices(npart, SHARD_BLOCK, rank, world, device=dev)
    W.fields = synthetic.sph_igm_fields(idx)
    del idx
    W.n_cells_global = npart
    nb = 2000
    Tv, Dv, ebins, c2, m2, v2, cie = synthetic.pion_like_tables(nT=24, nD=12, nbins=nb, emin=0.3, emax=1.4)
    cie_T, cc, cm, cv = cie
    cie_model = px.ThermalSpectralModel(ebins, cie_T, cc, cm, cv, logT=True, binscale="log")
    sm = px.PionSpectralModel(ebins, Tv, Dv, c2, m2, v2, cie_model=cie_model, binscale="log")
    model = px.IGMSourceModel(0.3, 1.4, nb, 0.3, binscale="log", temperature_field=("gas", "kT"),
                              spectral_model=sm, prng=SEED_GEN)
    W.make_source = lambda f: _source(px, f, world, rank, f"SPH{npart}", data_type="particles")
    area_time = _exposure_from_probe(px, torch, dist, world, model, W.make_source, W.fields, REDSHIFT,
                                     ("gas", "emission_measure"), args.photons)
    exp_time, area = 1.0e5, area_time / 1.0e5
    en, sg = synthetic.tbabs_like_sigma()
    pfx, efx = f"mem:bench_ph{rank}", f"mem:bench_ev{rank}"
    W.pfx, W.efx = pfx, efx
    normal = [0.3, -0.2, 0.93]
    absm = px.TBabsModel(NH, energy=en, cross_section=sg)

    def step(src):
        n_ph, n_c = px.make_photons(pfx, src, REDSHIFT, area, exp_time, model)
        n_ev = px.project_photons(pfx, efx, normal, [30.0, 45.0], absorb_model=absm, nH=NH,
                                  kernel="gaussian", prng=SEED_PROJ)
        return n_ph, n_c, n_ev

    def step_fused(src):
        return px.make_events(efx, src, REDSHIFT, area, exp_time, model, normal, [30.0, 45.0],
                              absorb_model=absm, nH=NH, kernel="gaussian", prng=SEED_PROJ)

    # (2-D log-binned table: option-generic sampler, see wl_amr)
    W.step, W.step_e2e = step, step_fused
    W.nZ, W.nD, W.h2d_skip = 0, 1, []
    W.desc = (f"synthetic SPH/Gadget-like {npart} gas particles (2^{int(np.log2(npart))}, configs[3]), IGM-style 2-D "
              f"(kT,nH) table 24x12x{nb} log bins + CIE fallback, gaussian SPH-kernel position smoothing, "
              f"tabulated absorption nH={NH}, off-axis normal, {args.photons:.3g} expected photons in total")
    W.scaling = "strong"
    return W


def wl_plline(args, px, torch, dist, dev, rank, world):
    """configs[4]: power-law + line sources on a 512^3 grid at a cosmological redshift."""
    from pyxsim_b200 import synthetic

    W = Workload()
    n, z = args.grid, 0.5
    ncell = n ** 3
    idx = synthetic.shard_indices(ncell, _grid_block(n), rank, world, device=dev)
    f = synthetic.cluster_chain_fields(idx, n, n_clusters=1, seed=SEED_DATA)
    em = f.pop(("gas", "emission_measure"))
    f.pop(("gas", "density"))
    f.pop(("gas", "kT"))
    f[("gas", "pl_luminosity")] = em                                  # keV/s up to the exposure scale
    f[("gas", "pl_alpha")] = 1.2 + 1.3 * synthetic.hash_uniform(idx, 77)
    f[("gas", "line_emission")] = em.clone()
    f[("gas", "line_sigma")] = 500.0 + 1500.0 * synthetic.hash_uniform(idx, 78)      # km/s
    del idx
    W.fields = f
    W.n_cells_global = ncell
    pl = px.PowerLawSourceModel(1.0, 0.5, 8.0, ("gas", "pl_luminosity"), ("gas", "pl_alpha"), prng=SEED_GEN)
    ln = px.LineSourceModel(6.4, ("gas", "line_emission"), sigma=("gas", "line_sigma"), prng=SEED_GEN + 1)
    units = {("gas", "pl_luminosity"): "keV/s", ("gas", "line_emission"): "photons/s",
             ("gas", "line_sigma"): "km/s"}

    def make_source(ff):
        s = _source(px, ff, world, rank, f"Grid{n}^3", block=_grid_block(n))
        s.units.update(units)
        return s

    W.make_source = make_source
    tot_target = args.photons * world
    exp_time = 1.0e5
    areas = [_exposure_from_probe(px, torch, dist, world, m, make_source, f, z, fld, share * tot_target) / exp_time
             for m, fld, share in ((pl, ("gas", "pl_luminosity"), 0.6), (ln, ("gas", "line_emission"), 0.4))]
    pfx, efx = f"mem:bench_ph{rank}", f"mem:bench_ev{rank}"
    normal = [0.3, -0.2, 0.93]

    def step(src):
        tot = [0, 0, 0]
        for k, (m, a) in enumerate(((pl, areas[0]), (ln, areas[1]))):
            n_ph, n_c = px.make_photons(pfx + f"_{k}", src, z, a, exp_time, m)
            n_ev = px.project_photons(pfx + f"_{k}", efx + f"_{k}", normal, [30.0, 45.0], absorb_model="wabs",
                                      nH=NH, prng=SEED_PROJ + k)
            tot = [tot[0] + n_ph, tot[1] + n_c, tot[2] + n_ev]
        return tuple(tot)

    W.step = step
    W.pfx, W.efx = pfx + "_0", efx + "_0"
    W.extra_lists = [(pfx + "_1", efx + "_1")]
    W.nZ, W.nD, W.h2d_skip = 0, 0, []
    W.desc = (f"PowerLawSourceModel (alpha field 1.2-2.5) + LineSourceModel (6.4 keV, velocity-dispersion field 500-2000 km/s) on a {n}^3 grid "
              f"(configs[4]), redshift {z}, off-axis normal, wabs nH={NH}, {tot_target:.3g} expected photons in total")
    W.scaling = "weak"
    return W


WORKLOADS = {"cluster": wl_cluster, "amr": wl_amr, "sph": wl_sph, "plline": wl_plline}


def workload_config(args, W, n_gpus):
    blk = _grid_block(args.grid) if args.config in ("cluster", "plline") else SHARD_BLOCK
    return {
        "workload": W.desc, "name": args.config,
        "cells_global": int(W.n_cells_global),
        "expected_photons": args.photons * (n_gpus if W.scaling == "weak" else 1),
        "kT": ("6 keV isothermal" if args.kT_range is None else f"log-uniform {args.kT_range}")
        if args.config == "cluster" else "per cell",
        "sharding": "none" if n_gpus == 1 else
                    f"one dataset, block-cyclic over {n_gpus} ranks (blocks of {blk} cells, RNG keyed by "
                    f"global cell id: the union of the per-rank lists is the single-rank list); collectives: "
                    f"scalar count reductions only",
        "l2": "inputs (per GPU: > 1 GB of cells, 8 GB photon list, 18 GB events per step) far exceed the 126 MB L2",
    }


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
    # CPUs (and with them the pinned buffers allocated from here on) next to this rank's GPU
    W_binding = parallel.bind_host_to_gpu(dev.index)
    import logging

    logging.getLogger("pyxsim_b200").setLevel("WARNING")
    W = WORKLOADS[args.config](args, px, torch, dist, dev, rank, world)
    src_dev = W.make_source(W.fields)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            counts = fn(src_dev)
        sync_all()
        profiling.enable(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(steps):
            counts = fn(src_dev)
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        stages = profiling.summary()
        launches = profiling.launches()
        profiling.enable(False)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return counts, float(t.item()) / steps, {k: v[1] / v[0] for k, v in stages.items()}, launches

    # ---- device-resident timing: the two-stage API ---------------------------------------------
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    counts, ms_per_step, stage_ms, launches = timed(W.step, args.steps, args.warmup)
    n_ph_all, n_c_all, n_ev_all = counts           # already summed over ranks
    value = n_ph_all / (ms_per_step * 1e-3)
    # sizes of this rank's lists (for the algorithmic bytes), before the fused run replaces the events
    Np = Na = Ne_local = 0
    for p2, e2 in [(W.pfx, W.efx)] + getattr(W, "extra_lists", []):
        l2 = px.load_list(p2)
        Np += int(l2.data["energy"].shape[0])
        Na += int(l2.data["num_photons"].shape[0])
        Ne_local += int(px.load_list(e2).data["eobs"].shape[0])

    # ---- the same step fused (make_events) -------------------------------------------------------
    fused = None
    if W.step_fused is not None and not args.no_fused:
        fc, fms, fstage, flaunch = timed(W.step_fused, args.steps, max(1, args.warmup - 1))
        fused = {"value": fc[0] / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms,
                 "photons_per_step": fc[0], "events_per_step": fc[2],
                 "same_counts_as_two_stage": list(fc) == list(counts),
                 "stage_ms": {k: round(v, 3) for k, v in fstage.items()}, "gpu_launches": flaunch,
                 "api": "pyxsim_b200.make_events (generate -> project in one kernel, no photon list; "
                        "events bit-identical to make_photons + project_photons)"}
    clocks = sampler.stop()

    # ---- roofline --------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
    Nc = int(next(iter(W.fields.values())).shape[0])
    w = 8
    cellw = (2 + W.nZ + W.nD) * w
    B_G = Nc * cellw + Na * (14 * w + 8) + Np * w
    B_P = Na * (7 * w + 8) + Np * w + Ne_local * 3 * w
    B_F = Nc * cellw + Na * 7 * w + Ne_local * 3 * w
    alg = {   # algorithmic bytes per launch of every timed stage (SURVEY.md §8d split by stage)
        "thermal_counts": Nc * cellw + Nc * 8,
        "scan_counts": Nc * 8 * 2 + Nc * 16,
        "compact_cells": Nc * 8 + Na * (7 * w) * 2 + Na * 24,
        "thermal_sample": Na * (2 * w + 24) + Np * w,
        "project": B_P,
        "make_events": Na * (cellw + 7 * w + 24) + Ne_local * 3 * w,
    }
    traffic = None
    if args.traffic_from:
        # measured DRAM traffic: only from an ncu capture of THIS command (scripts/summarise_profile.py)
        try:
            kk = json.load(open(args.traffic_from))["kernels"]

            def dram(*pats):
                return sum(v["dram_bytes"] for k, v in kk.items() if any(p in k for p in pats))

            traffic = {"thermal_sample": dram("k_pipeline<0", "k_prep_sample"),
                       "project": dram("k_pipeline<1", "k_prep_project<1>"),
                       "make_events": dram("k_pipeline<2", "k_prep_sample", "k_prep_project<0>")}
        except Exception:
            traffic = None
    kernel_of = {"thermal_sample": "pxb_thermal_sample: k_prep_sample + k_pipeline<MODE_SAMPLE>",
                 "project": "pxb_project: k_prep_project + k_pipeline<MODE_PROJECT> + stats",
                 "make_events": "pxb_make_events: k_prep_sample + k_prep_project + k_pipeline<MODE_FUSED> + stats"}

    def roof(stages, pick=None):
        cand = {k: v for k, v in stages.items() if k in alg}
        if not cand:
            return None
        dom = pick if pick in cand else max(cand, key=cand.get)
        ach = alg[dom] / (cand[dom] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel_of.get(dom, dom), "achieved": round(ach, 2), "peak": peak_gbs,
                "unit": "GB/s", "frac": round(ach / peak_gbs, 4), "traffic": (traffic or {}).get(dom),
                "peak_source": peak_src, "ms_per_launch": round(cand[dom], 3),
                "algorithmic_bytes_per_launch": int(alg[dom]),
                "stage_ms": {k: round(v, 3) for k, v in stages.items()},
                "stage_frac_of_hbm_peak": {k: round(alg[k] / (v * 1e-3) / 1e9 / peak_gbs, 4)
                                           for k, v in cand.items()}}

    roofline = roof(stage_ms)
    whole = (B_G + B_P) / (ms_per_step * 1e-3) / 1e9
    if fused is not None:
        fused["roofline"] = roof(dict(fused.pop("stage_ms")), "make_events")
        fused["whole_step_hbm_frac"] = round(B_F / (fused["ms_per_step"] * 1e-3) / 1e9 / peak_gbs, 4)
        fused["algorithmic_bytes_per_step"] = int(B_F)

    # ---- end to end: host buffers in, host events out ------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, px, torch, dist, world, dev, W, sync_all)
        binds = _gather_bindings(dist, world, W_binding)
        e2e["host_binding"] = {"per_rank": binds,
                               "numa_nodes_seen": sorted({str(b.get("numa_node")) for b in binds}),
                               "note": "each rank pinned to NVML's ideal CPUs of its GPU before any pinned "
                                       "allocation (first touch); one NUMA node for all GPUs = the host hides "
                                       "its topology and the binding cannot help"}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and W.cpu is not None:
        tables, area_time, D_A = W.cpu
        cpu = cpu_baseline(args, W.fields, tables, area_time, D_A, ncores=1)

    if rank == 0:
        two_stage = {"value": value, "unit": UNIT, "ms_per_step": ms_per_step,
                     "api": "pyxsim_b200.make_photons + project_photons (the reference's two calls, the photon "
                            "list resident in HBM between them)",
                     "photons_per_step": n_ph_all, "events_per_step": n_ev_all,
                     "whole_step_hbm_frac": round(whole / peak_gbs, 4),
                     "algorithmic_bytes_per_step": int(B_G + B_P), "roofline": roofline,
                     "gpu_launches": launches}
        if fused is not None:
            # headline: the fused step -- it generates and projects the same photons and leaves the
            # same event list bit for bit (checked every run: same counts; tests: same arrays)
            head = dict(value=fused["value"], ms_per_step=fused["ms_per_step"], roofline=fused["roofline"],
                        whole=fused["whole_step_hbm_frac"], bytes=fused["algorithmic_bytes_per_step"],
                        launches=fused["gpu_launches"], api=fused["api"])
        else:
            head = dict(value=value, ms_per_step=ms_per_step, roofline=roofline, whole=round(whole / peak_gbs, 4),
                        bytes=int(B_G + B_P), launches=launches, api=two_stage["api"])
        cfg = workload_config(args, W, world)
        cfg["step_api"] = head["api"]
        out = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": W.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "photons_per_step": n_ph_all, "active_cells": n_c_all, "events_per_step": n_ev_all,
            "whole_step_hbm_frac": head["whole"],
            "algorithmic_bytes_per_step": head["bytes"],
            "roofline": head["roofline"], "two_stage": two_stage,
            "fused_same_counts_as_two_stage": None if fused is None else fused["same_counts_as_two_stage"],
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": head["launches"],
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, px, torch, dist, world, dev, W, sync_all):
    """Same step through the public API with HOST inputs (pinned NumPy-visible tensors) and the
    event list read back into pinned host memory, copies inside the timed region.  The default
    (float64 events, the reference's file format) is the `e2e` value; `f32_events` repeats it with
    event_dtype="float32" (single-precision energies and sky offsets: half the bytes over PCIe),
    `d2h_ceiling` is a plain pinned device-to-host copy of the same number of bytes."""
    import psutil

    host_fields = {k: v.cpu().pin_memory() for k, v in W.fields.items()}
    src_host = W.make_source(host_fields)
    h2d = sum(v.numel() * v.element_size() for k, v in host_fields.items() if k not in W.h2d_skip)
    res = _e2e_leg(args, px, torch, dist, world, dev, W, sync_all, src_host, h2d, W.step_fused or W.step_e2e or W.step,
                   torch.float64, psutil)
    if getattr(W, "step_fused_f32", None) is not None:
        leg = _e2e_leg(args, px, torch, dist, world, dev, W, sync_all, src_host, h2d, W.step_fused_f32,
                       torch.float32, psutil)
        res["f32_events"] = {k: leg[k] for k in ("value", "unit", "d2h_bytes_per_step", "steps", "host_landing")}
    # what the PCIe link gives for a plain copy of one step's events (pinned, one stream)
    nbytes = int(res["d2h_bytes_per_step"])
    if nbytes > 0:
        chunk = min(nbytes, 1 << 30)
        dsrc = torch.empty(chunk, dtype=torch.uint8, device=dev)
        hdst = torch.empty(chunk, dtype=torch.uint8, pin_memory=True)
        hdst.copy_(dsrc, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(max(1, nbytes // chunk)):
            hdst.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = max(1, nbytes // chunk) * chunk / dt / 1e9
        tt = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        res["d2h_ceiling"] = {"gb_per_s_per_gpu": round(float(tt.item()), 2),
                              "photons_per_s_if_only_the_event_copy": round(
                                  res["photons_per_step"] / (nbytes / world / (float(tt.item()) * 1e9)), 1),
                              "note": "slowest rank's pinned D2H rate with all ranks copying at once"}
    return res


def _gather_bindings(dist, world, mine):
    """Host binding of every rank (NVML ideal CPU affinity, NUMA node of its GPU)."""
    if world == 1:
        return [mine]
    out = [None] * world
    dist.all_gather_object(out, mine)
    return out


def _e2e_leg(args, px, torch, dist, world, dev, W, sync_all, src_host, h2d, step, dtype, psutil):
    efxs = [W.efx] + [e for _, e in getattr(W, "extra_lists", [])]
    # pinned landing buffers for the events, sized from a dry run.  When host memory allows (per
    # rank share) the whole event list lands in pinned memory; otherwise it is streamed through
    # a fixed pinned ring (every byte still crosses PCIe, a consumer would drain the ring).
    _ = step(src_host)
    n_ev = max(int(px.load_list(e).data["eobs"].shape[0]) for e in efxs)
    full_cap = int(n_ev * 1.02 + 1024)
    budget = 0.25 * psutil.virtual_memory().available / max(world, 1)
    cap = full_cap if 3 * 8 * full_cap <= budget else max(1 << 20, int(budget // 24))
    cap = min(cap, full_cap)
    esz = 8 if dtype == torch.float64 else 4
    landing = {k: torch.empty(cap, dtype=dtype, pin_memory=True) for k in ("xsky", "ysky", "eobs")}
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
        n_ph = step(src_host)[0]
        d2h = 0
        for e in efxs:
            ev = px.load_list(e)
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
            d2h += 3 * esz * m
        tot_ph = n_ph
    copy_stream.synchronize()
    sync_all()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    return {"value": tot_ph * n_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": n_steps, "photons_per_step": tot_ph,
            "api": "make_events" if (W.step_fused or W.step_e2e) else "make_photons + project_photons",
            "host_landing": "full event list in pinned memory" if cap == full_cap else
            f"streamed through a {3 * esz * cap / 2**30:.1f} GiB pinned ring",
            "note": "host cell arrays -> generate + project -> events in pinned host memory (per rank: its "
                    "own shard in, its own events out); a step's event download overlaps the next step's "
                    "upload and kernels"}


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
    ncell = int(fields[("gas", "kT")].shape[0])
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
    from pyxsim_b200 import synthetic

    if args.config != "cluster":
        emit({"impl": "reference",
              "unavailable": "the CPU arm is defined for the default workload (configs[1]) only"})
        return
    tables = build_tables(args)
    n = args.grid
    fields = synthetic.cluster_chain_fields(torch.arange(n ** 3, dtype=torch.int64), n, n_clusters=1, kT_keV=6.0,
                                            kT_range=args.kT_range, seed=SEED_DATA)
    D_A, dfac = dist_fac(px, REDSHIFT)
    ss = mean_spec_sum(tables, np.array([6.0])) if args.kT_range is None else \
        mean_spec_sum(tables, fields[("gas", "kT")][:200000].numpy())
    area_time = args.photons / (float(fields[("gas", "emission_measure")].sum()) * ss * dfac)
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

    class _W:
        desc = (f"synthetic uniform grid {n}^3 thermal beta-model cluster (configs[1]"
                + (f"; a chain of {world} clusters, {world}x{n}x{n}x{n} cells, in ONE grid" if world > 1 else "")
                + f"), APEC-shaped table {NT}x{args.nbins} bins 0.1-10 keV, 3-D Doppler velocities, wabs nH={NH}, "
                f"{args.photons:.3g} expected photons per GPU, make_photons+project_photons (TAN sky)")
        n_cells_global = world * n ** 3
        scaling = "weak"

    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, _W, world),
           "cpu_baseline": last,
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
