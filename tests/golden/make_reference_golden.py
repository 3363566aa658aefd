"""Generate tests/golden/reference_golden.npz by running the REFERENCE'S OWN Python code
(``/root/reference/pyxsim``: ThermalSourceModel.process_data, PionSpectralModel.get_spectrum,
PowerLawSourceModel.process_data, LineSourceModel.process_data,
AbsorptionModel.absorb_photons, _project_photons) on seeded synthetic inputs.

The reference's third-party imports (yt, soxs, unyt, h5py, astropy, more_itertools) are not
installed; ``oracle/ref_import.py`` supplies stand-ins that carry no physics (unit-tagged
ndarrays, an in-memory h5py, the published yt Orientation rule).  The source models are
created with ``object.__new__`` and given exactly the attributes their constructors /
``setup_model`` would have set, so that no soxs table generator is needed; every statement
that computes a photon count, an energy, an absorption decision or a sky position is the
reference's.

Run in the build container (needs /root/reference):  python tests/golden/make_reference_golden.py
The fixtures pin ``oracle/ref_glue.py`` bit for bit (tests/test_cpu_oracle.py) and, through
it, the GPU path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def inputs():
    """Seeded inputs shared by the reference run and the oracle replay."""
    import helpers as H
    from pyxsim_b200 import synthetic

    rng = np.random.RandomState(2024)
    I = {}
    # thermal, 1-D table, metallicity mass-fraction field, one constant + one field element
    t = H.make_tables(nbins=64, nvar=2)
    n = 230
    I["t1"] = dict(tables=t, kT=np.exp(rng.uniform(np.log(0.01), np.log(90.0), n)),
                   em=1.0e60 * rng.lognormal(0, 1.0, n), Z=rng.uniform(0.0, 0.02, n),
                   Fe=rng.uniform(0.0, 2.0, n), dens=10 ** rng.uniform(-28, -25, n),
                   norm=4.0e-44, Zconvert=36.5, h_fraction=0.74)
    # thermal, log bins, accept_reject
    t2 = H.make_tables(nbins=48, binscale="log")
    I["t2"] = dict(tables=t2, kT=np.exp(rng.uniform(np.log(0.3), np.log(20.0), 120)),
                   em=1.0e60 * rng.lognormal(0, 1.0, 120), norm=3.0e-44)
    # Pion 2-D + CIE fallback, var elements
    Tv, Dv, ebins, c2, m2, v2, cie = synthetic.pion_like_tables(nT=10, nD=6, nbins=40, nvar=2)
    n = 160
    I["pion"] = dict(Tv=Tv, Dv=Dv, ebins=ebins, c2=c2, m2=m2, v2=v2, cie=cie,
                     kT=np.exp(rng.uniform(np.log(2e-3), np.log(4.0), n)),
                     nH=np.exp(rng.uniform(np.log(5e-8), np.log(2e-3), n)),
                     em=1.0e60 * rng.lognormal(0, 1, n), norm=2.0e-43)
    # power law / line
    n = 140
    alpha = rng.uniform(0.5, 2.5, n)
    alpha[:10] = 1.0
    alpha[10:20] = 2.0
    I["pl"] = dict(lum=1.0e45 * rng.lognormal(0, 1, n), alpha=alpha, norm=3.0e-44, z=0.5)
    I["ln"] = dict(rate=1.0e42 * rng.lognormal(0, 1, n), sig_kms=rng.uniform(100.0, 2000.0, n),
                   norm=5.0e-41, z=0.2)
    # absorption table + photon list for the projection cases
    en, sg = synthetic.tbabs_like_sigma(n=300)
    I["abs"] = dict(energy=en, sigma=sg * 1.0e22)        # base class: tau = sigma * nH (literal)
    nc = 400
    nph = rng.poisson(12.0 * rng.lognormal(0, 1.0, nc)).astype(np.int64)
    I["plist"] = dict(x=rng.uniform(-300, 300, nc), y=rng.uniform(-300, 300, nc),
                      z=rng.uniform(-300, 300, nc), vx=rng.normal(0, 500, nc),
                      vy=rng.normal(0, 500, nc), vz=rng.normal(100, 500, nc),
                      dx=rng.uniform(2.0, 30.0, nc), num_photons=nph,
                      energy=np.exp(rng.uniform(np.log(0.1), np.log(10.0), int(nph.sum()))))
    # internal-absorption cube (internal_absorption.py:117-126) smaller than the cell cloud, so that
    # some cells fall outside it (fill value 0)
    nw, nd = 12, 9
    I["col"] = dict(wbins=np.linspace(-260.0, 260.0, nw + 1), dbins=np.linspace(-240.0, 240.0, nd + 1),
                    nH=0.4 * rng.lognormal(0.0, 1.0, (nw, nw, nd)))
    return I


PROJ_CASES = {
    "onaxis_cells": dict(obs="external", data_type="cells", normal="z", kw=dict(sigma_pos=0.5), seed=21),
    "offaxis_gauss": dict(obs="external", data_type="particles", normal=[0.2, 0.9, -0.4],
                          kw=dict(kernel="gaussian", flat_sky=True, north_vector=[0.0, 0.0, 1.0]), seed=22),
    "offaxis_cells_los": dict(obs="external", data_type="cells", normal=[1.0, -2.0, 5.0],
                              kw=dict(save_los=True), seed=23),
    "tophat_phys": dict(obs="external", data_type="particles", normal="x",
                        kw=dict(kernel="top_hat", phys_coord=True, no_shifting=True, save_los=True), seed=24),
    "allsky_tophat": dict(obs="internal", data_type="particles", normal=[0.1, 0.3, 0.9],
                          kw=dict(kernel="top_hat", save_los=True), seed=25),
    "allsky_cells": dict(obs="internal", data_type="cells", normal=[0.1, 0.3, 0.9], kw=dict(save_los=True), seed=26),
    # no absorption: every photon survives, eobs is the deterministic Doppler-shifted energy
    "doppler_only": dict(obs="external", data_type="cells", normal=[0.3, -0.5, 0.8], kw=dict(save_los=True),
                         seed=27, absorb=False),
    # column_file: nH cube interpolated at the rotated cell centres + per-cell absorption at the
    # source-frame energy (photon_list.py:524-546,672-680,715-723), on top of a foreground column
    "column_file": dict(obs="external", data_type="cells", normal=[0.3, -0.4, 0.85],
                        kw=dict(north_vector=[0.1, 0.9, 0.2], column_file="cols.h5"), seed=29),
    "doppler_internal": dict(obs="internal", data_type="cells", normal=[0.3, -0.5, 0.8], kw=dict(),
                             seed=28, absorb=False),
}


def run_reference(I):
    from oracle import ref_glue as O, ref_import as R

    M = R.modules()
    U = R.UArray
    sm_mod, th = M["spectral_models"], M["thermal"]
    out = {}

    def thermal_model(spectral_model, t, **over):
        m = object.__new__(th.ThermalSourceModel)
        d = dict(temperature_field="T", emission_measure_field="EM", density_field="dens",
                 entropy_field="S", max_density=None, min_entropy=None, kT_min=0.025, kT_max=64.0,
                 nh_field=None, h_fraction=0.74, pbar=R.NoBar(), _nei=False, var_elem_keys=None,
                 var_ion_keys=None, Zmet=0.3, Zconvert=1.0, num_var_elem=0, var_elem={}, mconvert={},
                 observer="external", _density_dependence=False, spectral_model=spectral_model,
                 method="invert_cdf", binscale=t["binscale"], nbins=t["ebins"].size - 1,
                 emid=0.5 * (t["ebins"][1:] + t["ebins"][:-1]), ebins=t["ebins"],
                 bin_edges=np.log10(t["ebins"]) if t["binscale"] == "log" else t["ebins"])
        d.update(over)
        m.__dict__.update(d)
        return m

    def table_model(t, logT=False):
        cls = sm_mod.CloudyCIESpectralModel if logT else sm_mod.TableCIEModel
        s = object.__new__(cls)
        s.si = sm_mod.SpectralInterpolator1D(t["Tvals"], t["cosmic"], t["metal"], t["var"])
        s.nbins = t["ebins"].size - 1
        # what make_fluxf reads (spectral_models.py:101-134)
        s.ebins, s.emid = t["ebins"], 0.5 * (t["ebins"][1:] + t["ebins"][:-1])
        s.Tvals, s.cosmic_spec, s.metal_spec, s.var_spec = t["Tvals"], t["cosmic"], t["metal"], t["var"]
        return s

    # ---- t1 --------------------------------------------------------------------------------
    a = I["t1"]
    n = a["kT"].size
    chunk = R.FakeChunk({"T": U(a["kT"], "keV"), "EM": U(a["em"], "cm**-3"), "Zf": U(a["Z"], "dimensionless"),
                         "Fef": U(a["Fe"], "Zsun"), "dens": U(a["dens"], "g/cm**3")}, n)
    m = thermal_model(table_model(a["tables"]), a["tables"], Zmet="Zf", Zconvert=a["Zconvert"],
                      h_fraction=a["h_fraction"], num_var_elem=2, var_elem={"O": 0.5, "Fe": "Fef"},
                      var_elem_keys=["O", "Fe"], mconvert={"Fe": 1.0}, max_density=3e-26,
                      prng=np.random.RandomState(11))
    nc, nph, idx, e = m.process_data("photons", chunk, a["norm"])
    out.update(t1_nc=nc, t1_nph=nph, t1_idx=idx, t1_e=np.asarray(e))

    # spectrum mode of the same model (no Doppler shifting): base.py:494-495 + shift_spectrum
    m.pbar = R.NoBar()
    ebnew = np.linspace(0.3, 7.0, 41)
    out.update(t1_spec=m.process_data("spectrum", chunk, a["norm"], ebins=ebnew, shifting=False))

    # band fluxes of the same model: make_fluxf + the "luminosity" / "photon_rate" branch
    # (spectral_models.py:101-134, base.py:501-511); also the bare flux function
    for tag, energy, mode in (("lum", True, "luminosity"), ("rate", False, "photon_rate")):
        fl = m.make_fluxf(0.5, 2.0, energy=energy)
        out[f"t1_{tag}"] = m.process_data(mode, chunk, a["norm"], emin=0.5, emax=2.0, fluxf=fl)
        c_, m_, v_ = fl(np.array([0.05, 0.3, 1.0, 7.7, 40.0]))
        out.update({f"t1_{tag}_cf": c_, f"t1_{tag}_mf": m_, f"t1_{tag}_vf": v_})

    # ---- t2: log bins, both sampling methods -----------------------------------------------
    a = I["t2"]
    n = a["kT"].size
    for meth, seed in (("invert_cdf", 12), ("accept_reject", 13)):
        chunk = R.FakeChunk({"T": U(a["kT"], "keV"), "EM": U(a["em"], "cm**-3")}, n)
        m = thermal_model(table_model(a["tables"]), a["tables"], method=meth, prng=np.random.RandomState(seed))
        nc, nph, idx, e = m.process_data("photons", chunk, a["norm"])
        out.update({f"t2_{meth}_nph": nph, f"t2_{meth}_idx": idx, f"t2_{meth}_e": np.asarray(e)})

    # ---- pion --------------------------------------------------------------------------------
    a = I["pion"]
    n = a["kT"].size
    cie_T, cc, cm, cv = a["cie"]
    tcie = dict(Tvals=cie_T, cosmic=cc, metal=cm, var=cv, ebins=a["ebins"], binscale="log")
    pm = object.__new__(sm_mod.PionSpectralModel)
    pm.__dict__.update(nbins=a["ebins"].size - 1, var_spec=a["v2"], nvar_elem=2,
                       min_table_kT=10 ** a["Tv"][0] / R.sys.modules["soxs.constants"].K_per_keV,
                       max_table_kT=10 ** a["Tv"][-1] / R.sys.modules["soxs.constants"].K_per_keV,
                       min_table_nH=10 ** a["Dv"][0], max_table_nH=10 ** a["Dv"][-1],
                       si=sm_mod.SpectralInterpolator2D(a["Tv"], a["Dv"], a["c2"], a["m2"], a["v2"]),
                       cie_model=table_model(tcie, logT=True))
    c, mm, v = pm.get_spectrum(a["kT"][:40], a["nH"][:40])
    out.update(pion_c=c, pion_m=mm, pion_v=v)
    chunk = R.FakeChunk({"T": U(a["kT"], "keV"), "EM": U(a["em"], "cm**-3"), "nH": U(a["nH"], "cm**-3")}, n)
    t = dict(ebins=a["ebins"], binscale="log")
    m = thermal_model(pm, t, nh_field="nH", _density_dependence=True, kT_min=1e-3, num_var_elem=2,
                      var_elem={"O": 0.4, "Ne": 1.1}, var_elem_keys=["O", "Ne"], prng=np.random.RandomState(14))
    nc, nph, idx, e = m.process_data("photons", chunk, a["norm"])
    out.update(pion_nph=nph, pion_idx=idx, pion_e=np.asarray(e))
    # band fluxes of the 2-D model (_get_flux_2d + CIE fallback, spectral_models.py:464-537); the
    # cut keeps kT >= 0.1 keV so that every fallback cell lies inside the CIE table (interp1d raises
    # outside it)
    pm.__dict__.update(ebins=a["ebins"], emid=0.5 * (a["ebins"][1:] + a["ebins"][:-1]), Tvals=a["Tv"],
                       Dvals=a["Dv"], dTvals=np.diff(a["Tv"]), dDvals=np.diff(a["Dv"]), n_T=a["Tv"].size,
                       n_D=a["Dv"].size, cosmic_spec=a["c2"], metal_spec=a["m2"])
    m.kT_min = 0.1
    for tag, energy, mode in (("lum", True, "luminosity"), ("rate", False, "photon_rate")):
        fl = m.make_fluxf(0.4, 1.2, energy=energy)
        out[f"pion_{tag}"] = m.process_data(mode, chunk, a["norm"], emin=0.4, emax=1.2, fluxf=fl)

    # ---- power law ---------------------------------------------------------------------------
    a = I["pl"]
    n = a["lum"].size
    pl = object.__new__(M["power_law_sources"].PowerLawSourceModel)
    pl.__dict__.update(ftype="gas", luminosity_field="L", alpha="alpha", e0=R.UQuantity(1.0, "keV"),
                       emin=R.UQuantity(0.5, "keV"), emax=R.UQuantity(40.0, "keV"),
                       scale_factor=1.0 / (1.0 + a["z"]), observer="external", prng=np.random.RandomState(15))
    chunk = R.FakeChunk({"L": U(a["lum"], "keV/s"), "alpha": U(a["alpha"], "")}, n)
    nc, nph, act, e = pl.process_data("photons", chunk, a["norm"])
    out.update(pl_nph=nph, pl_act=act, pl_e=np.asarray(e))
    pl.alpha = 1.7
    pl.prng = np.random.RandomState(16)
    nc, nph, act, e = pl.process_data("photons", chunk, a["norm"])
    out.update(plc_nph=nph, plc_act=act, plc_e=np.asarray(e))

    # ---- line ----------------------------------------------------------------------------------
    a = I["ln"]
    n = a["rate"].size
    ln = object.__new__(M["line_sources"].LineSourceModel)
    ln.__dict__.update(ftype="gas", emission_field="N", sigma="sig", e0=R.UQuantity(3.5, "keV"),
                       scale_factor=1.0 / (1.0 + a["z"]), observer="external", prng=np.random.RandomState(17))
    chunk = R.FakeChunk({"N": U(a["rate"], "photons/s"), "sig": U(a["sig_kms"], "km/s")}, n)
    nc, nph, act, e = ln.process_data("photons", chunk, a["norm"])
    out.update(ln_nph=nph, ln_act=act, ln_e=np.asarray(e))

    # ---- absorption + projection -----------------------------------------------------------------
    ab = I["abs"]

    class TableAbs(sm_mod.AbsorptionModel):
        _name = "tbabs"

        def __init__(self, nH, abund_table="angr"):
            super().__init__(nH, ab["energy"], ab["sigma"])

    am = TableAbs(0.15)
    ee = I["plist"]["energy"][:500]
    out.update(abs_trans=np.asarray(am.get_absorb(ee)),
               abs_det=am.absorb_photons(ee, prng=np.random.RandomState(18)))
    pl_mod = M["photon_list"]
    d = I["plist"]
    col = I["col"]
    cc = PROJ_CASES["column_file"]
    fc = R.FakeH5File("cols.h5", "w")
    pc = fc.create_group("parameters")
    pc.attrs["normal"] = np.array(cc["normal"])
    pc.attrs["north"] = O.orientation(cc["normal"], cc["kw"]["north_vector"])[1]
    gc = fc.create_group("data")
    for k in ("wbins", "dbins", "nH"):
        gc.create_dataset(k, data=col[k])
    for name, case in PROJ_CASES.items():
        fph = R.FakeH5File(f"ph_{name}.h5", "w")
        p = fph.create_group("parameters")
        for k, v in dict(fid_exp_time=1.0e5, fid_area=1000.0, fid_redshift=0.05 if case["obs"] == "external" else 0.0,
                         fid_d_a=200.0 if case["obs"] == "external" else 0.0, data_type=case["data_type"],
                         observer=case["obs"]).items():
            p.create_dataset(k, data=v)
        g = fph.create_group("data")
        for k, v in d.items():
            g.create_dataset(k, data=v.copy())
        akw = dict(absorb_model=TableAbs, nH=0.15) if case.get("absorb", True) else {}
        n_ev = pl_mod._project_photons(case["obs"], f"ph_{name}", f"ev_{name}", case["normal"],
                                       [30.0, 45.0], prng=np.random.RandomState(case["seed"]),
                                       **akw, **case["kw"])
        ev = R.FILES[f"ev_{name}.h5"]
        out[f"pj_{name}_n"] = n_ev
        for k in ("xsky", "ysky", "eobs", "los"):
            if k in ev["data"]:
                out[f"pj_{name}_{k}"] = ev["data"][k][()]
        for k in ("flux", "emin", "emax"):
            out[f"pj_{name}_{k}"] = ev["parameters"][k][()]
    return out


if __name__ == "__main__":
    res = run_reference(inputs())
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **res)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), len(res), "arrays")
