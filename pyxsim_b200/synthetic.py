"""Synthetic emissivity tables, absorption cross-sections and datasets of the shapes named in
BASELINE.json / SURVEY.md §8c-d.

No APEC / Cloudy / tbabs data files exist offline (soxs downloads them), and the hot path
treats tables and cross-sections as opaque arrays, so the tests and the benchmark use these
stand-ins.  Everything is seeded and pure NumPy/torch: the oracle and the GPU path see
identical inputs.
"""
import numpy as np
import torch

from . import constants as K
from .data import ArrayDataSource

M_H = 1.6726219e-24       # g
KPC = K.cm_per_kpc


def apec_like_tables(nT=36, nbins=4000, emin=0.1, emax=10.0, nvar=0, binscale="linear",
                     seed=7, negative=False):
    """(Tvals[keV], ebins, cosmic, metal, var|None): APEC-shaped tables.

    36 log-spaced temperature rows 0.0216..68.5 keV (APEC rows 14..49, dlog T = 0.1), a
    bremsstrahlung-like H/He continuum with a few lines in ``cosmic``, a line-rich ``metal``
    table, and per-element line tables.  Units mimic photons cm^3/s per bin (~1e-17).
    ``negative=True`` injects small negative entries so the clip of base.py:460 matters.
    """
    rng = np.random.RandomState(seed)
    kT = 8.617333262e-8 * 10.0 ** (5.4 + 0.1 * np.arange(nT))
    if binscale == "log":
        ebins = np.logspace(np.log10(emin), np.log10(emax), nbins + 1)
    else:
        ebins = np.linspace(emin, emax, nbins + 1)
    emid = 0.5 * (ebins[1:] + ebins[:-1])
    de = np.diff(ebins)
    T = kT[:, None]
    E = emid[None, :]
    brems = 1.0e-15 * np.exp(-E / T) / (E * np.sqrt(T)) * de[None, :]

    def lines(nline, strength, e_lo, e_hi):
        e0 = np.sort(rng.uniform(e_lo, e_hi, nline))
        amp = strength * rng.lognormal(0.0, 1.0, nline)
        out = np.zeros((nT, nbins))
        for el, a in zip(e0, amp):
            sig = np.maximum(el * 4.0e-4 * np.sqrt(kT), 0.6 * de.mean())[:, None]
            # emissivity peaks where kT ~ line energy
            emis = a * np.exp(-el / T) * np.exp(-T / (8.0 * el)) / np.sqrt(T)
            prof = np.exp(-0.5 * ((E - el) / sig) ** 2) / (sig * np.sqrt(2 * np.pi)) * de[None, :]
            out += emis * prof
        return out

    lo, hi = emin * 1.05, emax * 0.95
    cosmic = brems + lines(6, 2.0e-17, lo, min(hi, 2.0))
    metal = 0.3 * brems * (E / (E + 1.0)) + lines(150, 4.0e-17, lo, hi)
    var = None
    if nvar:
        var = np.stack([lines(25, 3.0e-17, lo, hi) + 0.02 * brems for _ in range(nvar)])
    if negative:
        for tab in ([cosmic, metal] + ([var] if nvar else [])):
            mask = rng.uniform(size=tab.shape) < 0.02
            tab[mask] = -np.abs(tab[mask]) * 0.5
    return kT, ebins, cosmic, metal, var


def pion_like_tables(nT=24, nD=12, nbins=2000, emin=0.3, emax=1.4, nvar=0, binscale="log",
                     seed=11):
    """2-D (log10 T[K], log10 nH) tables, shape (nD*nT, nbins) with T fastest, plus a 1-D
    log-T CIE fallback covering higher temperatures.  Values scale ~ nH like the Cloudy
    photoionisation tables (the lookup divides by nH, spectral_models.py:446-449)."""
    rng = np.random.RandomState(seed)
    Tvals = np.linspace(4.0, 7.1, nT)            # log10 K  (kT ~ 0.9 eV .. 1.08 keV)
    Dvals = np.linspace(-7.0, -3.0, nD)          # log10 cm^-3
    kT_cie = 8.617333262e-8 * 10.0 ** np.linspace(6.0, 8.5, 26)
    _, ebins, c1, m1, v1 = apec_like_tables(26, nbins, emin, emax, nvar, binscale, seed)
    # fallback: same generator evaluated on the CIE grid (log-T coordinate)
    cie = (np.log10(kT_cie * K.K_per_keV), c1, m1, v1)
    kT2 = 10.0 ** Tvals / K.K_per_keV
    emid = 0.5 * (ebins[1:] + ebins[:-1])
    de = np.diff(ebins)
    c2 = np.zeros((nD * nT, nbins))
    m2 = np.zeros((nD * nT, nbins))
    v2 = np.zeros((nvar, nD * nT, nbins)) if nvar else None
    e_lines = np.sort(rng.uniform(emin * 1.05, emax * 0.95, 40))
    for iy, ld in enumerate(Dvals):
        nH = 10.0 ** ld
        ion = 1.0 / (1.0 + 1.0e-5 / nH)          # photoionisation suppresses low-density lines
        for ix, t in enumerate(kT2):
            row = nT * iy + ix
            brem = 1.0e-15 * np.exp(-emid / t) / (emid * np.sqrt(t)) * de
            c2[row] = nH * brem
            prof = np.zeros(nbins)
            for el in e_lines:
                sig = max(el * 1.0e-3, 0.6 * de.mean())
                prof += np.exp(-el / t) * np.exp(-0.5 * ((emid - el) / sig) ** 2) / sig * de
            m2[row] = nH * (0.2 * brem + 3.0e-17 * ion * prof)
            for k in range(nvar):
                v2[k, row] = nH * 1.0e-17 * ion * np.roll(prof, 7 * (k + 1))
    return Tvals, Dvals, ebins, c2, m2, v2, cie


def tbabs_like_sigma(n=4000, emin=0.03, emax=15.0):
    """(energy[keV], sigma[cm^2]) on a log grid: E^-2.5 photoabsorption with K/L edges, about
    2.4e-22 cm^2 at 1 keV -- a stand-in for the tbabs table that ships with soxs."""
    e = np.logspace(np.log10(emin), np.log10(emax), n)
    s = 2.4e-22 * e ** -2.5
    for edge, jump in [(0.284, 1.08), (0.40, 1.06), (0.532, 1.45), (0.707, 1.15), (0.867, 1.12),
                       (1.303, 1.06), (1.84, 1.10), (2.47, 1.06), (7.11, 1.8)]:
        s = np.where(e >= edge, s * jump, s)
    return e, s * 0.55


def beta_model_fields(n, box_kpc=2000.0, rc_kpc=50.0, beta=2.0 / 3.0, rho_c=0.04 * M_H,
                      r_max_kpc=1000.0, kT_keV=6.0, kT_range=None, vz_kms=(400.0, 400.0),
                      v3d_sigma_kms=None, bulk_kms=(0.0, 0.0, 0.0), seed=32, device="cpu",
                      metallicity=None):
    """Uniform ``n^3`` grid with the beta-model cluster of pyxsim/tests/utils.py:34-82
    (rho_c = 0.04 m_H g/cm^3, r_c = 50 kpc, beta = 2/3, rho = 0 beyond 1 Mpc, kT = 6 keV,
    v_z ~ N(400, 400) km/s).  ``EM = n_e n_H dV`` with n_H = 0.76 rho/m_H, n_e = 0.88 rho/m_H
    (fully ionised primordial H/He; stated in SURVEY.md §8d).  Returns a dict of 1-D torch
    float64 tensors on ``device`` keyed like a yt dataset."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dxk = box_kpc / n
    ax = (torch.arange(n, dtype=torch.float64, device=device) + 0.5) * dxk - 0.5 * box_kpc
    x = ax.view(n, 1, 1).expand(n, n, n).reshape(-1)
    y = ax.view(1, n, 1).expand(n, n, n).reshape(-1)
    z = ax.view(1, 1, n).expand(n, n, n).reshape(-1)
    r2 = x * x + y * y + z * z
    rho = rho_c * (1.0 + r2 / (rc_kpc * rc_kpc)) ** (-1.5 * beta)
    rho = torch.where(r2 > r_max_kpc * r_max_kpc, torch.zeros_like(rho), rho)
    dV = (dxk * KPC) ** 3
    em = (0.76 * rho / M_H) * (0.88 * rho / M_H) * dV
    N = n ** 3
    if kT_range is None:
        kT = torch.full((N,), float(kT_keV), dtype=torch.float64, device=device)
    else:
        u = torch.rand(N, generator=g, dtype=torch.float64, device=device)
        lo, hi = np.log(kT_range[0]), np.log(kT_range[1])
        kT = torch.exp(lo + (hi - lo) * u)
    if v3d_sigma_kms is None:
        vx = torch.zeros(N, dtype=torch.float64, device=device)
        vy = torch.zeros(N, dtype=torch.float64, device=device)
        vz = vz_kms[0] + vz_kms[1] * torch.randn(N, generator=g, dtype=torch.float64, device=device)
    else:
        vx, vy, vz = (bulk_kms[i] + v3d_sigma_kms * torch.randn(N, generator=g, dtype=torch.float64,
                                                                device=device) for i in range(3))
    fields = {
        ("gas", "kT"): kT, ("gas", "emission_measure"): em, ("gas", "density"): rho,
        ("index", "x"): x.contiguous(), ("index", "y"): y.contiguous(), ("index", "z"): z.contiguous(),
        ("index", "dx"): torch.full((N,), dxk, dtype=torch.float64, device=device),
        ("gas", "velocity_x"): vx, ("gas", "velocity_y"): vy, ("gas", "velocity_z"): vz,
    }
    if metallicity is not None:
        u = torch.rand(N, generator=g, dtype=torch.float64, device=device)
        fields[("gas", "metallicity")] = metallicity[0] + (metallicity[1] - metallicity[0]) * u
    return fields


def beta_model_source(n, chunk_size=None, **kw):
    """ArrayDataSource of the beta-model grid (temperature stored as kT in keV)."""
    box = kw.get("box_kpc", 2000.0)
    f = beta_model_fields(n, **kw)
    return ArrayDataSource(f, data_type="cells", domain_left_edge=[-0.5 * box] * 3,
                           domain_right_edge=[0.5 * box] * 3, chunk_size=chunk_size,
                           name=f"BetaModel{n}^3")


def particle_beta_model_fields(n, rc_kpc=50.0, r_max_kpc=1000.0, hsml_kpc=10.0, seed=35,
                               device="cpu", kT_range=(0.01, 2.0), nH_range=(1e-7, 1e-3)):
    """SPH-like particle realisation of the same cluster (pyxsim/tests/utils.py:85-138 shape:
    beta-model radial sampling, fixed smoothing length) with log-uniform kT and nH so both the
    2-D table and its CIE fallback are exercised."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def rand(m):
        return torch.rand(m, generator=g, dtype=torch.float64, device=device)

    # beta = 2/3 profile: M(<x) ~ x - atan(x); invert by bisection-free rejection on x
    xmax = r_max_kpc / rc_kpc
    u = rand(n) * (xmax - np.arctan(xmax))
    xr = u + 1.0                       # Newton iterations on x - atan(x) = u
    for _ in range(40):
        f = xr - torch.atan(xr) - u
        xr = torch.clamp(xr - f * (1.0 + xr * xr) / (xr * xr + 1e-30), min=1e-9)
    r = xr * rc_kpc
    ct = 2.0 * rand(n) - 1.0
    st = torch.sqrt(torch.clamp(1.0 - ct * ct, min=0.0))
    ph = 2.0 * np.pi * rand(n)
    x, y, z = r * st * torch.cos(ph), r * st * torch.sin(ph), r * ct
    lk = np.log(kT_range[0]) + (np.log(kT_range[1]) - np.log(kT_range[0])) * rand(n)
    ln = np.log(nH_range[0]) + (np.log(nH_range[1]) - np.log(nH_range[0])) * rand(n)
    nH = torch.exp(ln)
    vol = 4.0 / 3.0 * np.pi * (hsml_kpc * KPC) ** 3 / 32.0
    em = nH * nH * 1.16 * vol
    fields = {
        ("gas", "kT"): torch.exp(lk), ("gas", "emission_measure"): em,
        ("gas", "H_nuclei_density"): nH,
        ("gas", "particle_position_x"): x, ("gas", "particle_position_y"): y,
        ("gas", "particle_position_z"): z,
        ("gas", "smoothing_length"): torch.full((n,), hsml_kpc, dtype=torch.float64, device=device),
    }
    for a in "xyz":
        fields[("gas", f"particle_velocity_{a}")] = 300.0 * torch.randn(n, generator=g, dtype=torch.float64, device=device)
    return fields


# ---------------------------------------------------------------------------------------
# global datasets addressed by cell index: every rank materialises only the cells it owns
# ---------------------------------------------------------------------------------------
def _i64(v):
    """Python int -> two's-complement int64 value (for 64-bit hash constants)."""
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def hash_u64(idx, salt):
    """splitmix64 of ``idx + salt * golden`` on int64 tensors (wrap-around arithmetic, logical
    shifts emulated with masks): a counter-based hash, so field values of a cell depend on its
    GLOBAL index only."""
    x = idx + _i64(0x9E3779B97F4A7C15 * (salt + 1))
    x = (x ^ ((x >> 30) & 0x3FFFFFFFF)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ ((x >> 27) & 0x1FFFFFFFFF)) * _i64(0x94D049BB133111EB)
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def hash_uniform(idx, salt):
    """Uniform in (0, 1) from the top 52 bits of the hash."""
    h = (hash_u64(idx, salt) >> 12) & 0xFFFFFFFFFFFFF
    return (h.to(torch.float64) + 0.5) * (1.0 / 4503599627370496.0)


def hash_normal(idx, salt):
    u1, u2 = hash_uniform(idx, 2 * salt + 101), hash_uniform(idx, 2 * salt + 102)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * np.pi * u2)


def shard_indices(n, block, rank, size, device="cpu"):
    """Global indices of the cells rank ``rank`` of ``size`` owns under the block-cyclic split of
    ``parallel.shard_bounds`` (blocks rank, rank + size, ... of ``block`` cells), ascending."""
    if size == 1:
        return torch.arange(n, dtype=torch.int64, device=device)
    nblocks = (n + block - 1) // block
    mine = torch.arange(rank, nblocks, size, dtype=torch.int64, device=device)
    idx = (mine[:, None] * block + torch.arange(block, dtype=torch.int64, device=device)[None, :]).reshape(-1)
    return idx[idx < n]


def _beta_em(r2, dV, rc_kpc, beta, rho_c, r_max_kpc):
    rho = rho_c * (1.0 + r2 / (rc_kpc * rc_kpc)) ** (-1.5 * beta)
    rho = torch.where(r2 > r_max_kpc * r_max_kpc, torch.zeros_like(rho), rho)
    return (0.76 * rho / M_H) * (0.88 * rho / M_H) * dV, rho


def cluster_chain_fields(idx, n, n_clusters=1, box_kpc=2000.0, kT_keV=6.0, kT_range=None,
                         v3d_sigma_kms=300.0, bulk_kms=(100.0, -50.0, 200.0), seed=32, **beta_kw):
    """Cells ``idx`` (global indices) of a uniform grid of ``(n_clusters * n) x n x n`` cubic cells
    holding a chain of ``n_clusters`` identical beta-model clusters along x, one per ``n^3`` block
    (``n_clusters = 1``: the cluster of ``beta_model_fields``).  Velocities and temperatures are
    hashed from the global index."""
    bk = dict(rc_kpc=50.0, beta=2.0 / 3.0, rho_c=0.04 * M_H, r_max_kpc=1000.0)
    bk.update(beta_kw)
    dxk = box_kpc / n
    ix = idx // (n * n)
    iy = (idx // n) % n
    iz = idx % n
    x = (ix.to(torch.float64) + 0.5) * dxk - 0.5 * box_kpc * n_clusters
    y = (iy.to(torch.float64) + 0.5) * dxk - 0.5 * box_kpc
    z = (iz.to(torch.float64) + 0.5) * dxk - 0.5 * box_kpc
    xc = ((ix // n).to(torch.float64) + 0.5) * box_kpc - 0.5 * box_kpc * n_clusters   # own cluster's centre
    r2 = (x - xc) ** 2 + y * y + z * z
    em, rho = _beta_em(r2, (dxk * KPC) ** 3, **bk)
    if kT_range is None:
        kT = torch.full_like(x, float(kT_keV))
    else:
        lo, hi = np.log(kT_range[0]), np.log(kT_range[1])
        kT = torch.exp(lo + (hi - lo) * hash_uniform(idx, seed))
    f = {("gas", "kT"): kT, ("gas", "emission_measure"): em, ("gas", "density"): rho,
         ("index", "x"): x, ("index", "y"): y, ("index", "z"): z,
         ("index", "dx"): torch.full_like(x, dxk)}
    for a, ax in enumerate("xyz"):
        f[("gas", f"velocity_{ax}")] = bulk_kms[a] + v3d_sigma_kms * hash_normal(idx, seed + 1 + a)
    return f


AMR_LEVELS = 4


def amr_leaf_count(n):
    """Leaf cells of the nested AMR cluster: levels 0..2 keep the shell outside their central
    half, level 3 is complete."""
    return (AMR_LEVELS - 1) * (n ** 3 - (n // 2) ** 3) + n ** 3


def amr_cluster_fields(idx, n=320, box_kpc=2000.0, nvar=6, seed=41, kT_range=(2.0, 9.0), **beta_kw):
    """Leaf cells ``idx`` of a 4-level nested-refinement beta-model cluster: level l is an ``n^3``
    patch of cell size ``box / (n 2^l)`` centred on the cluster; on levels 0..2 the central half is
    covered by the next level, so only the shell cells are leaves (``n = 320``: 1.19e8 leaves,
    BASELINE config 3).  Leaves are numbered level by level, x-slab by x-slab.  Per-cell kT,
    metallicity and ``nvar`` element abundances are hashed from the global index."""
    bk = dict(rc_kpc=50.0, beta=2.0 / 3.0, rho_c=0.04 * M_H, r_max_kpc=1000.0)
    bk.update(beta_kw)
    h, lo = n // 2, n // 4
    hi = lo + h
    shell = n ** 3 - h ** 3
    level = torch.clamp(idx // shell, max=AMR_LEVELS - 1)
    k = idx - level * shell
    full = level == AMR_LEVELS - 1
    # complete level: plain row-major
    fx, fy, fz = k // (n * n), (k // n) % n, k % n
    # shell levels: x-slabs below / inside / above the covered block
    slab_in = n * n - h * h
    a_lo = lo * n * n
    a_mid = a_lo + h * slab_in
    in_lo, in_mid = k < a_lo, (k >= a_lo) & (k < a_mid)
    k2 = k - a_lo
    k3 = k - a_mid
    sx = torch.where(in_lo, k // (n * n), torch.where(in_mid, lo + k2 // slab_in, hi + k3 // (n * n)))
    rem_full = torch.where(in_lo, k % (n * n), k3 % (n * n))
    rem = k2 % slab_in                        # inside a slab that crosses the covered block
    b_lo = lo * n
    b_mid = b_lo + h * (n - h)
    r_lo, r_mid = rem < b_lo, (rem >= b_lo) & (rem < b_mid)
    r2_ = rem - b_lo
    r3_ = rem - b_mid
    t = r2_ % (n - h)
    my = torch.where(r_lo, rem // n, torch.where(r_mid, lo + r2_ // (n - h), hi + r3_ // n))
    mz = torch.where(r_lo, rem % n, torch.where(r_mid, torch.where(t < lo, t, t + h), r3_ % n))
    sy = torch.where(in_mid, my, rem_full // n)
    sz = torch.where(in_mid, mz, rem_full % n)
    ix, iy, iz = (torch.where(full, a, b) for a, b in ((fx, sx), (fy, sy), (fz, sz)))
    dxk = (box_kpc / n) * torch.pow(torch.tensor(0.5, dtype=torch.float64, device=idx.device), level.to(torch.float64))
    half = 0.5 * n * dxk
    x = (ix.to(torch.float64) + 0.5) * dxk - half
    y = (iy.to(torch.float64) + 0.5) * dxk - half
    z = (iz.to(torch.float64) + 0.5) * dxk - half
    em, rho = _beta_em(x * x + y * y + z * z, (dxk * KPC) ** 3, **bk)
    lk0, lk1 = np.log(kT_range[0]), np.log(kT_range[1])
    f = {("gas", "kT"): torch.exp(lk0 + (lk1 - lk0) * hash_uniform(idx, seed)),
         ("gas", "emission_measure"): em, ("gas", "density"): rho,
         ("index", "x"): x, ("index", "y"): y, ("index", "z"): z, ("index", "dx"): dxk,
         ("gas", "metallicity"): 0.1 + 0.5 * hash_uniform(idx, seed + 1)}
    for a, ax in enumerate("xyz"):
        f[("gas", f"velocity_{ax}")] = 300.0 * hash_normal(idx, seed + 10 + a)
    for kk in range(nvar):
        f[("gas", f"El{kk}_fraction")] = 0.05 + 0.6 * hash_uniform(idx, seed + 20 + kk)
    return f


def sph_igm_fields(idx, rc_kpc=50.0, r_max_kpc=1000.0, hsml_kpc=10.0, seed=35,
                   kT_range=(0.01, 2.0), nH_range=(1e-7, 1e-3)):
    """Particles ``idx`` of an SPH-like realisation of the cluster (the shape of
    ``particle_beta_model_fields``, BASELINE config 4): beta-model radial sampling, fixed
    smoothing length, log-uniform kT and nH so both the 2-D (kT, nH) table and its CIE fallback
    are exercised.  Everything is hashed from the global particle id."""
    xmax = r_max_kpc / rc_kpc
    u = hash_uniform(idx, seed) * (xmax - np.arctan(xmax))
    xr = u + 1.0                       # Newton iterations on x - atan(x) = u
    for _ in range(40):
        fv = xr - torch.atan(xr) - u
        xr = torch.clamp(xr - fv * (1.0 + xr * xr) / (xr * xr + 1e-30), min=1e-9)
    r = xr * rc_kpc
    ct = 2.0 * hash_uniform(idx, seed + 1) - 1.0
    st = torch.sqrt(torch.clamp(1.0 - ct * ct, min=0.0))
    ph = 2.0 * np.pi * hash_uniform(idx, seed + 2)
    lk = np.log(kT_range[0]) + (np.log(kT_range[1]) - np.log(kT_range[0])) * hash_uniform(idx, seed + 3)
    ln = np.log(nH_range[0]) + (np.log(nH_range[1]) - np.log(nH_range[0])) * hash_uniform(idx, seed + 4)
    nH = torch.exp(ln)
    vol = 4.0 / 3.0 * np.pi * (hsml_kpc * KPC) ** 3 / 32.0
    f = {("gas", "kT"): torch.exp(lk), ("gas", "emission_measure"): nH * nH * 1.16 * vol,
         ("gas", "H_nuclei_density"): nH,
         ("gas", "particle_position_x"): r * st * torch.cos(ph),
         ("gas", "particle_position_y"): r * st * torch.sin(ph),
         ("gas", "particle_position_z"): r * ct,
         ("gas", "smoothing_length"): torch.full_like(r, hsml_kpc)}
    for a, ax in enumerate("xyz"):
        f[("gas", f"particle_velocity_{ax}")] = 300.0 * hash_normal(idx, seed + 10 + a)
    return f
