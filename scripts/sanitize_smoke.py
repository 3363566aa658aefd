"""A small pass through every photon kernel for compute-sanitizer (memcheck / racecheck):
make_photons + project_photons + make_events on a 12^3 cluster (a few 1e5 photons), common and
option-generic kernels, with empty cells in the photon list; the variable-element sampler in
both of its modes; the bucketed Doppler-shifted spectrum.

    compute-sanitizer --tool memcheck  python scripts/sanitize_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pyxsim_b200 as px
from pyxsim_b200 import synthetic
from pyxsim_b200.photon_list import DeviceList, save_list


def main():
    kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=600)
    src = synthetic.beta_model_source(12, kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda")
    sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
    for generic in ("", "1"):
        if generic:
            os.environ["PXB_GENERIC_KERNELS"] = "1"
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 600, 0.3, temperature_field=("gas", "kT"), prng=5)
        n_ph, n_c = px.make_photons("mem:san_ph", src, 0.05, 2.0e3, 1.0e5, model)
        n_ev = px.project_photons("mem:san_ph", "mem:san_ev", [0.2, -0.1, 0.95], [30.0, 45.0], absorb_model="wabs",
                                  nH=0.05, prng=6, save_los=True)
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 600, 0.3, temperature_field=("gas", "kT"), prng=5)
        tot = px.make_events("mem:san_ev1", src, 0.05, 2.0e3, 1.0e5, model, [0.2, -0.1, 0.95], [30.0, 45.0],
                             absorb_model="wabs", nH=0.05, prng=6, save_los=True)
        assert tot == (n_ph, n_c, n_ev), (tot, n_ph, n_c, n_ev)
        a, b = px.load_list("mem:san_ev"), px.load_list("mem:san_ev1")
        assert all(np.array_equal(a.numpy(k), b.numpy(k)) for k in ("xsky", "ysky", "eobs", "los"))
        print(f"generic={bool(generic)}: {n_ph} photons, {n_c} cells, {n_ev} events")
    # a list with empty cells (bisection look-ups) and float32 events
    lst = px.load_list("mem:san_ph").to_host()
    nph = lst.data["num_photons"].copy()
    keep = np.ones(nph.size, bool)
    keep[::3] = False
    e = lst.data["energy"]
    off = np.concatenate([[0], np.cumsum(nph)])
    e2 = np.concatenate([e[off[i]:off[i + 1]] for i in range(nph.size) if keep[i]])
    nph[~keep] = 0
    d = dict(lst.data, num_photons=nph, energy=e2)
    save_list("mem:san_sparse", DeviceList(lst.info, lst.parameters, d))
    n = px.project_photons("mem:san_sparse", "mem:san_sparse_ev", "z", [30.0, 45.0], absorb_model="wabs", nH=0.05,
                           prng=7, event_dtype="float32")
    print("sparse list:", int(nph.sum()), "photons ->", n, "events")
    os.environ.pop("PXB_GENERIC_KERNELS", None)
    # variable elements: stored-threshold component choice (few photons per cell) and the general split
    kT_nodes, ebins, cosmic, metal, var = synthetic.apec_like_tables(nT=36, nbins=600, nvar=3)
    names = ("El0", "El1", "El2")
    smv = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal, var, var_elem_names=names)
    for sf in ("", "2"):
        if sf:
            os.environ["PXB_SPLIT_FACTOR"] = sf
        model = px.ThermalSourceModel(smv, 0.1, 10.0, 600, 0.3, temperature_field=("gas", "kT"), prng=5,
                                      var_elem={nm: 0.2 + 0.3 * k for k, nm in enumerate(names)})
        n_ph, n_c = px.make_photons("mem:san_phv", src, 0.05, 2.0e3, 1.0e5, model)
        print(f"var_elem split_factor={sf or 'default'}: {n_ph} photons, {n_c} cells")
    os.environ.pop("PXB_SPLIT_FACTOR", None)
    # Doppler-shifted spectrum: cells bucketed by table row (constant and per-cell metallicity)
    for zmet in (0.3, ("gas", "metallicity")):
        srcz = synthetic.beta_model_source(12, kT_range=(1.0, 9.0), v3d_sigma_kms=300.0, device="cuda",
                                           metallicity=(0.1, 0.6))
        model = px.ThermalSourceModel(sm, 0.1, 10.0, 600, zmet, temperature_field=("gas", "kT"), prng=5)
        eb, spec = model.make_spectrum(srcz, 0.2, 9.0, 300, redshift=0.05, normal=[0.3, -0.5, 0.8])
        print("shifted spectrum:", float(np.sum(spec)))


if __name__ == "__main__":
    main()
