import cProfile, pstats, sys, time, logging
sys.path.insert(0, '.')
import numpy as np, torch
import pyxsim_b200 as px
from pyxsim_b200 import synthetic
logging.getLogger("pyxsim_b200").setLevel("WARNING")
kT_nodes, ebins, cosmic, metal, _ = synthetic.apec_like_tables(nT=36, nbins=4000)
f = synthetic.beta_model_fields(64, v3d_sigma_kms=300.0, device="cuda")
src = px.ArrayDataSource(f, domain_left_edge=[-1000.0]*3, domain_right_edge=[1000.0]*3)
sm = px.ThermalSpectralModel(ebins, kT_nodes, cosmic, metal)
model = px.ThermalSourceModel(sm, 0.1, 10.0, 4000, 0.3, temperature_field=("gas","kT"), prng=50)
def step():
    n = px.make_photons("mem:p", src, 0.05, 1000.0*44, 1.0e5, model)
    e = px.project_photons("mem:p", "mem:e", "z", [30.0,45.0], absorb_model="wabs", nH=0.02, prng=51)
    return n, e
for _ in range(3): print(step())
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize(); print("ms/step", (time.perf_counter()-t0)/20*1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
