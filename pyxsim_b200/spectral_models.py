"""Spectral-table and absorption models (host side of the seam described in SURVEY.md §8b).

Mirrors ``pyxsim/spectral_models.py``: the *construction* of emissivity tables stays with
soxs (``TableCIEModel.prepare_spectrum`` :253-263, ``PionSpectralModel.prepare_spectrum``
:413-427) -- here any object that provides the table arrays will do -- while the *lookup*
(``SpectralInterpolator1D/2D`` :22-82, ``interp1d_spec``/``interp2d_spec``) and the absorption
rejection (``AbsorptionModel.absorb_photons`` :563-583) run on the GPU through libpxb.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, constants as K
from .device import ptr, stream_ptr, to_device, empty, require_cuda


def _f64(a):
    return np.ascontiguousarray(a, dtype="float64")


class _DeviceModel:
    """Owns a ``pxb_model*`` (tables, row sums, bin edges in HBM)."""

    def __init__(self, desc, keepalive):
        require_cuda()
        self.handle = C.c_void_p()
        self._keep = keepalive
        _lib.call("pxb_model_create", C.byref(desc), C.byref(self.handle))
        self.nonnegative = bool(_lib.load().pxb_model_nonnegative(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _lib.load().pxb_model_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def _table_desc(tvals, dvals, cosmic, metal, var):
    d = _lib.TableDesc()
    keep = []
    if cosmic is None:
        return d, keep
    tv = _f64(tvals)
    c = _f64(cosmic)
    m = _f64(metal)
    keep += [tv, c, m]
    d.nx = tv.size
    d.ny = 0
    d.nbins = c.shape[-1]
    d.xnodes = tv.ctypes.data
    d.cosmic = c.ctypes.data
    d.metal = m.ctypes.data
    if dvals is not None:
        dv = _f64(dvals)
        keep.append(dv)
        d.ny = dv.size
        d.ynodes = dv.ctypes.data
    nrows = d.nx * max(d.ny, 1)
    if c.shape != (nrows, d.nbins) or m.shape != c.shape:
        raise ValueError(f"table shape {c.shape} does not match ({nrows}, {d.nbins})")
    if var is not None:
        v = _f64(var)
        if v.shape[1:] != c.shape:
            raise ValueError("var_spec must have shape (nvar, nrows, nbins)")
        keep.append(v)
        d.nvar = v.shape[0]
        d.var = v.ctypes.data
    return d, keep


class ThermalSpectralModel:
    """Array-backed thermal spectral model with the attribute surface the source models use
    (``ebins, emid, de, nbins, Tvals, var_elem_names, var_ion_names, model_vers,
    prepare_spectrum, get_spectrum``; base.py:107-112,227,453-455)."""

    _logT = False
    _density_dependence = False

    def __init__(self, ebins, Tvals, cosmic_spec=None, metal_spec=None, var_spec=None,
                 var_elem_names=(), var_ion_names=(), binscale="linear", logT=None,
                 model_vers="array", table_fn=None):
        self.ebins = _f64(ebins)
        self.emid = 0.5 * (self.ebins[1:] + self.ebins[:-1])
        self.de = np.diff(self.ebins)
        self.nbins = self.emid.size
        self.binscale = binscale
        self.Tvals = _f64(Tvals)
        self.var_elem_names = list(var_elem_names)
        self.var_ion_names = list(var_ion_names)
        self.model_vers = model_vers
        if logT is not None:
            self._logT = bool(logT)
        self._table_fn = table_fn
        self.cosmic_spec = None if cosmic_spec is None else _f64(cosmic_spec)
        self.metal_spec = None if metal_spec is None else _f64(metal_spec)
        self.var_spec = None if var_spec is None else _f64(var_spec)
        self._dev = {}

    # -- reference API ------------------------------------------------------------------
    def prepare_spectrum(self, zobs):
        """Set ``cosmic_spec/metal_spec/var_spec`` for observer redshift ``zobs``.  With a
        ``table_fn(zobs) -> (cosmic, metal, var)`` the tables are rebuilt, otherwise the
        arrays given at construction are taken to be in the observer frame already."""
        if self._table_fn is not None:
            c, m, v = self._table_fn(zobs)
            self.cosmic_spec, self.metal_spec = _f64(c), _f64(m)
            self.var_spec = None if v is None else _f64(v)
            self._dev = {}      # tables changed: the device copy is rebuilt on next use
        if self.cosmic_spec is None:
            raise RuntimeError("spectral model has no tables")

    def _Tconv(self, kT):
        return np.log10(kT * K.K_per_keV) if self._logT else kT

    def _indices(self, vals, nodes):
        # spectral_models.py:35-37
        i = (np.digitize(vals, nodes) - 1).astype("int32")
        return np.minimum(np.maximum(i, 0), len(nodes) - 2).astype("int32")

    def get_spectrum(self, kT):
        """(cosmic, metal, var) spectra for temperatures ``kT`` [keV] -- GPU replacement of
        ``interp1d_spec`` behind ``ThermalSpectralModel.get_spectrum`` (spectral_models.py:94-99)."""
        x = np.atleast_1d(self._Tconv(np.asarray(kT, dtype="float64")))
        return self._interp(x, self._indices(x, self.Tvals), None, None)

    def _interp(self, x, xi, y, yi):
        dm = self.device_model()
        n = x.size
        nv = 0 if self.var_spec is None else self.var_spec.shape[0]
        dx, dxi = to_device(x), to_device(xi, dtype=_torch().int32)
        dy = dyi = None
        if y is not None:
            dy, dyi = to_device(y), to_device(yi, dtype=_torch().int32)
        c = empty(n * self.nbins)
        m = empty(n * self.nbins)
        v = empty(nv * n * self.nbins) if nv else None
        _lib.call("pxb_interp_spec", dm.handle, n, ptr(dx), ptr(dxi), ptr(dy), ptr(dyi), ptr(c),
                  ptr(m), ptr(v), stream_ptr())
        cn = c.cpu().numpy().reshape(n, self.nbins)
        mn = m.cpu().numpy().reshape(n, self.nbins)
        vn = v.cpu().numpy().reshape(nv, n, self.nbins) if nv else None
        return cn, mn, vn

    def _band_row_fluxes(self, emin, emax, energy):
        """Per-row band sums exactly as make_fluxf tabulates them (spectral_models.py:102-117):
        bins strictly inside (emin, emax); photon or energy (x emid) flux."""
        eidxs = (self.ebins[:-1] > emin) & (self.ebins[1:] < emax)
        emid = self.emid[eidxs]
        if energy:
            cf = (self.cosmic_spec[:, eidxs] * emid).sum(axis=-1)
            mf = (self.metal_spec[:, eidxs] * emid).sum(axis=-1)
        else:
            cf = self.cosmic_spec[:, eidxs].sum(axis=-1)
            mf = self.metal_spec[:, eidxs].sum(axis=-1)
        vf = None
        if self.var_spec is not None:
            vf = (self.var_spec[:, :, eidxs] * emid).sum(axis=-1) if energy else self.var_spec[:, :, eidxs].sum(axis=-1)
        return _f64(cf), _f64(mf), None if vf is None else _f64(vf)

    def make_fluxf(self, emin, emax, energy=False):
        """Band flux function (spectral_models.py:101-134): the row fluxes are tabulated on the
        host from the tables (a one-off of nT x nbins adds), their interpolation at the cells'
        temperatures runs on the GPU.  Returns a :class:`BandFlux`: callable like the reference's
        ``_fluxf(kT)`` and accepted by ``ThermalSourceModel.process_data(fluxf=...)``."""
        if self.cosmic_spec is None:
            raise RuntimeError("call prepare_spectrum(zobs) before make_fluxf")
        return BandFlux(self, float(emin), float(emax), bool(energy))

    # -- device ---------------------------------------------------------------------------
    def _bin_edges(self):
        return np.log10(self.ebins) if self.binscale == "log" else self.ebins

    def _fill_desc(self, desc, keep):
        d, k = _table_desc(self.Tvals, None, self.cosmic_spec, self.metal_spec, self.var_spec)
        desc.primary = d
        keep += k
        desc.log_t = int(self._logT)
        desc.density_dependent = 0

    def device_model(self, method="invert_cdf"):
        if self.cosmic_spec is None:
            raise RuntimeError("call prepare_spectrum(zobs) before using the model")
        dm = self._dev.get(method)
        if dm is None:
            desc = _lib.ModelDesc()
            keep = []
            self._fill_desc(desc, keep)
            edges = _f64(self._bin_edges())
            emid = _f64(self.emid)
            lin = _f64(self.ebins)
            keep += [edges, emid, lin]
            desc.bin_edges = edges.ctypes.data
            desc.emid = emid.ctypes.data
            desc.ebins = lin.ctypes.data
            desc.binscale_log = int(self.binscale == "log")
            desc.method = {"invert_cdf": 0, "accept_reject": 1}[method]
            desc.K_per_keV = K.K_per_keV
            dm = _DeviceModel(desc, keep)
            self._dev[method] = dm
        return dm


def _torch():
    import torch

    return torch


class BandFlux:
    """What ``make_fluxf`` returns: the tabulated row fluxes of one energy band, resident on the
    device, plus the reference's call signature ``fluxf(kT[, nH]) -> (cflux, mflux, vflux)``."""

    def __init__(self, model, emin, emax, energy):
        self.model, self.emin, self.emax, self.energy = model, emin, emax, energy
        self.cf, self.mf, self.vf = model._band_row_fluxes(emin, emax, energy)
        cie = getattr(model, "cie_model", None) if model._density_dependence else None
        self.fb = cie._band_row_fluxes(emin, emax, energy) if cie is not None else None
        self._dev = None

    def desc(self):
        """(pxb_band_flux, tensors to keep alive)"""
        if self._dev is None:
            arrs = [self.cf, self.mf, self.vf] + (list(self.fb) if self.fb is not None else [None] * 3)
            self._dev = [None if a is None else to_device(a.reshape(-1)) for a in arrs]
        B = _lib.BandFluxDesc()
        for name, t in zip(("cflux", "mflux", "vflux", "fb_cflux", "fb_mflux", "fb_vflux"), self._dev):
            setattr(B, name, None if t is None else t.data_ptr())
        return B, self._dev

    def evaluate(self, cells_struct, keep):
        """out[i] = cell_nrm * (cf + Z mf + sum e_k vf_k) for a filled ``pxb_thermal_cells``;
        raises ValueError like scipy's interp1d when a selected cell lies outside the table."""
        dm = self.model.device_model()
        B, tensors = self.desc()
        out = empty(cells_struct.ncell)
        bad = _torch().zeros(1, dtype=_torch().int64, device=out.device)
        _lib.call("pxb_thermal_flux", dm.handle, C.byref(cells_struct), C.byref(B), ptr(out),
                  bad.data_ptr(), stream_ptr())
        nbad = int(bad.cpu()[0])
        del tensors, keep
        if nbad:
            raise ValueError(f"A value in x_new is outside the interpolation range "
                             f"({nbad} cells outside the temperature table).")
        return out

    def __call__(self, kT, nH=None):
        """Reference signature: per-table fluxes at temperatures ``kT`` [keV] (and ``nH`` for a
        density-dependent model) as NumPy arrays.  One launch per table: the kernel's "cosmic"
        slot is pointed at the table wanted and every abundance is zero."""
        kT = np.atleast_1d(np.asarray(kT, "float64"))
        n = kT.size
        nv = 0 if self.vf is None else self.vf.shape[0]
        dkT = to_device(kT)
        dem = _torch().ones(n, dtype=_torch().float64, device=dkT.device)
        dnH = None if nH is None else to_device(np.atleast_1d(np.asarray(nH, "float64")))
        dm = self.model.device_model()
        B0, tensors = self.desc()
        nrows, nrows_fb = self.cf.size, (0 if self.fb is None else self.fb[0].size)
        bad = _torch().zeros(1, dtype=_torch().int64, device=dkT.device)
        zeros = _torch().zeros(max(nrows, nrows_fb), dtype=_torch().float64, device=dkT.device)
        outs = []
        for which in ["c", "m"] + list(range(nv)):
            B = _lib.BandFluxDesc()
            for name in ("cflux", "mflux", "vflux", "fb_cflux", "fb_mflux", "fb_vflux"):
                setattr(B, name, getattr(B0, name))
            if which == "m":
                B.cflux, B.fb_cflux = B0.mflux, B0.fb_mflux
            elif which != "c":
                # through the element slot itself (scipy evaluates an N-D y with a different
                # but equivalent expression, and so does the kernel): zero cosmic, weight 1
                B.cflux = zeros.data_ptr()
                B.fb_cflux = zeros.data_ptr() if B0.fb_cflux else None
            Cs = _lib.ThermalCells()
            Cs.ncell, Cs.nvar = n, nv
            Cs.kT, Cs.em = dkT.data_ptr(), dem.data_ptr()
            Cs.nH = None if dnH is None else dnH.data_ptr()
            Cs.kT_min, Cs.kT_max = -1.0e300, 1.0e300
            Cs.max_density, Cs.min_entropy = 1.0e300, -1.0e300
            Cs.spectral_norm, Cs.xh_const, Cs.zmet_const = 1.0, 1.0, 0.0
            if which not in ("c", "m"):
                Cs.elem_const[which] = 1.0
            out = empty(n)
            _lib.call("pxb_thermal_flux", dm.handle, C.byref(Cs), C.byref(B), ptr(out), bad.data_ptr(),
                      stream_ptr())
            outs.append(out.cpu().numpy())
        del tensors
        if int(bad.cpu()[0]):
            raise ValueError("A value in x_new is outside the interpolation range.")
        return outs[0], outs[1], (np.stack(outs[2:]) if nv else None)

class TableCIEModel(ThermalSpectralModel):
    """APEC/SPEX table model.  Same constructor as the reference (spectral_models.py:203-251);
    the tables themselves are produced by soxs' ``CIEGenerator`` exactly as in the reference,
    so this class needs soxs.  Without soxs use :class:`ThermalSpectralModel` with arrays."""

    def __init__(self, model, emin, emax, nbins, kT_min, kT_max, binscale="linear",
                 var_elem=None, model_root=None, model_vers=None, thermal_broad=True,
                 nolines=False, abund_table="angr", nei=False):
        try:
            from soxs.spectra import CIEGenerator  # type: ignore
        except Exception as e:  # pragma: no cover
            raise ImportError(
                "TableCIEModel builds its tables with soxs, which is not installed; pass "
                "precomputed tables to pyxsim_b200.ThermalSpectralModel instead") from e
        cgen = CIEGenerator(model, emin, emax, nbins, binscale=binscale, var_elem=var_elem,
                            model_root=model_root, model_vers=model_vers, broadening=thermal_broad,
                            nolines=nolines, abund_table=abund_table, nei=nei)
        self.cgen = cgen
        # row selection bracketing [kT_min, kT_max]: spectral_models.py:245-247
        self.idx_min = max(np.searchsorted(cgen.Tvals, kT_min) - 1, 0)
        self.idx_max = min(np.searchsorted(cgen.Tvals, kT_max) + 1, cgen.nT - 1)
        super().__init__(cgen.ebins, cgen.Tvals[self.idx_min:self.idx_max], binscale=binscale,
                         var_elem_names=getattr(cgen, "var_elem_names", ()),
                         var_ion_names=getattr(cgen, "var_ion_names", ()),
                         model_vers=cgen.model_vers)
        self.nei = nei

    def prepare_spectrum(self, zobs):
        c, m, v = self.cgen._get_table(list(range(self.idx_min, self.idx_max)), zobs, 0.0)
        self.cosmic_spec, self.metal_spec = _f64(c), _f64(m)
        self.var_spec = None if v is None else _f64(v)
        self._dev = {}


class PionSpectralModel(ThermalSpectralModel):
    """Density-dependent (kT, nH) table with a 1-D CIE fallback outside the table
    (spectral_models.py:329-462).  Array-backed: ``Tvals``/``Dvals`` are log10(T[K]) and
    log10(nH[cm^-3]) nodes, tables have shape (n_D*n_T, nbins) with T fastest
    (interpolate.pyx:112), ``cie_model`` is a log-T :class:`ThermalSpectralModel`."""

    _logT = True
    _density_dependence = True

    def __init__(self, ebins, Tvals, Dvals, cosmic_spec, metal_spec, var_spec=None, cie_model=None,
                 var_elem_names=(), binscale="linear", model_vers="array"):
        super().__init__(ebins, Tvals, cosmic_spec, metal_spec, var_spec,
                         var_elem_names=var_elem_names, binscale=binscale, logT=True,
                         model_vers=model_vers)
        self.Dvals = _f64(Dvals)
        self.n_T, self.n_D = self.Tvals.size, self.Dvals.size
        self.min_table_kT = 10 ** self.Tvals[0] / K.K_per_keV
        self.max_table_kT = 10 ** self.Tvals[-1] / K.K_per_keV
        self.min_table_nH = 10 ** self.Dvals[0]
        self.max_table_nH = 10 ** self.Dvals[-1]
        self.cie_model = cie_model
        self.nvar_elem = 0 if var_spec is None else self.var_spec.shape[0]

    def prepare_spectrum(self, zobs):
        super().prepare_spectrum(zobs)
        if self.cie_model is not None:
            self.cie_model.prepare_spectrum(zobs)

    def get_spectrum(self, kT, nH):
        """spectral_models.py:429-456 evaluated on the GPU (2-D part divided by nH, CIE
        fallback outside the table bounds)."""
        import torch

        kT = np.atleast_1d(np.asarray(kT, "float64"))
        nH = np.atleast_1d(np.asarray(nH, "float64"))
        use_igm = (kT >= self.min_table_kT) & (kT <= self.max_table_kT)
        use_igm &= (nH >= self.min_table_nH) & (nH <= self.max_table_nH)
        nv = self.nvar_elem
        cspec = np.zeros((kT.size, self.nbins))
        mspec = np.zeros((kT.size, self.nbins))
        vspec = np.zeros((nv, kT.size, self.nbins)) if nv else None
        if use_igm.any():
            x = np.log10(kT[use_igm] * K.K_per_keV)
            y = np.log10(nH[use_igm])
            c1, m1, v1 = self._interp(x, self._indices(x, self.Tvals), y,
                                      self._indices(y, self.Dvals))
            nHi = nH[use_igm]
            cspec[use_igm] = c1 / nHi[:, None]
            mspec[use_igm] = m1 / nHi[:, None]
            if nv:
                vspec[:, use_igm] = v1 / nHi[None, :, None]
        if (~use_igm).any():
            if self.cie_model is None:
                raise RuntimeError("cells fall outside the 2-D table and no cie_model was given")
            c2, m2, v2 = self.cie_model.get_spectrum(kT[~use_igm])
            cspec[~use_igm] = c2
            mspec[~use_igm] = m2
            if nv:
                vspec[:, ~use_igm] = v2
        return cspec, mspec, vspec

    def _fill_desc(self, desc, keep):
        d, k = _table_desc(self.Tvals, self.Dvals, self.cosmic_spec, self.metal_spec, self.var_spec)
        desc.primary = d
        keep += k
        desc.log_t = 1
        desc.density_dependent = 1
        desc.kT_lo, desc.kT_hi = self.min_table_kT, self.max_table_kT
        desc.nH_lo, desc.nH_hi = self.min_table_nH, self.max_table_nH
        if self.cie_model is not None:
            cm = self.cie_model
            d2, k2 = _table_desc(cm.Tvals, None, cm.cosmic_spec, cm.metal_spec, cm.var_spec)
            desc.fallback = d2
            keep += k2
            desc.fallback_log_t = int(cm._logT)


# ---------------------------------------------------------------------------------------
# absorption
# ---------------------------------------------------------------------------------------
def _grid_kind(e):
    """0 arbitrary, 1 uniform in E, 2 uniform in log10 E (lets the kernel index directly)."""
    if e.size < 3:
        return 0, 0.0, 0.0
    d = np.diff(e)
    if np.allclose(d, d[0], rtol=1e-9, atol=0.0):
        return 1, float(e[0]), float(1.0 / d[0])
    if e[0] > 0:
        le = np.log10(e)
        dl = np.diff(le)
        if np.allclose(dl, dl[0], rtol=1e-9, atol=0.0):
            return 2, float(le[0]), float(1.0 / dl[0])
    return 0, 0.0, 0.0


class AbsorptionModel:
    """``AbsorptionModel(nH, energy, cross_section)`` (spectral_models.py:540-583): tabulated
    cross-section on an ascending energy grid [keV].  As in the reference the optical depth is
    literally ``cross_section * nH`` (spectral_models.py:560-561): with nH in 1e22 cm^-2 the
    table is in units of 1e-22 cm^2."""

    _name = ""
    _kind = _lib.ABS_TABLE

    def __init__(self, nH, energy, cross_section):
        self._nH = nH
        self.emid = _f64(energy)
        self.sigma = _f64(cross_section)
        self._dev = None

    @property
    def nH(self):
        return self._nH

    @nH.setter
    def nH(self, value):
        self._nH = value

    def _fill_params(self, P):
        """Write the absorption part of a ``pxb_project_params``; returns tensors to keep alive."""
        P.absorb_kind = self._kind
        P.nH = 0.0 if self._nH is None else float(self._nH)
        if self._kind != _lib.ABS_TABLE:
            return []
        if self._dev is None:
            self._dev = (to_device(self.emid), to_device(self.sigma))
        P.abs_energy = self._dev[0].data_ptr()
        P.abs_sigma = self._dev[1].data_ptr()
        P.abs_n = self.emid.size
        P.abs_grid, P.abs_x0, P.abs_inv_dx = _grid_kind(self.emid)
        return list(self._dev)

    def get_absorb(self, e):
        """Transmission exp(-sigma(E) nH) evaluated on the GPU."""
        e = np.atleast_1d(np.asarray(e, "float64"))
        P = _lib.ProjectParams()
        keep = self._fill_params(P)
        de = to_device(e)
        out = empty(e.size)
        _lib.call("pxb_absorb_transmission", C.byref(P), e.size, ptr(de), ptr(out), stream_ptr())
        del keep
        return out.cpu().numpy()

    def absorb_photons(self, eobs, prng=None):
        """Boolean survival mask (spectral_models.py:563-583): the uniforms come from ``prng`` in
        the reference's order (one per photon), the decision ``u < exp(-sigma(E) nH)`` is the
        projection kernels' own (``pxb_absorb_decide``).  ``project_photons`` does not call this
        -- it decides in flight with its counter-based generator."""
        from .utils import parse_prng

        prng = parse_prng(prng)
        eobs = np.ascontiguousarray(eobs, "float64")
        if eobs.size == 0:
            return np.array([], dtype="bool")
        u = prng.uniform(size=eobs.size)
        P = _lib.ProjectParams()
        keep = self._fill_params(P)
        de, du = to_device(eobs), to_device(u)
        out = torch.empty(eobs.size, dtype=torch.uint8, device=de.device)
        _lib.call("pxb_absorb_decide", C.byref(P), eobs.size, ptr(de), ptr(du), out.data_ptr(),
                  stream_ptr())
        del keep
        return out.cpu().numpy().astype(bool)


class TBabsModel(AbsorptionModel):
    """Tuebingen-Boulder absorption (spectral_models.py:586-609).  The cross-section table is
    third-party data (soxs ``get_tbabs_absorb``): it is taken from soxs when importable, else
    it must be supplied as ``energy=`` / ``cross_section=``."""

    _name = "tbabs"

    def __init__(self, nH, abund_table="angr", energy=None, cross_section=None):
        if energy is None:
            try:
                from soxs.spectra import get_tbabs_absorb  # type: ignore
            except Exception as e:  # pragma: no cover
                raise ImportError(
                    "TBabsModel needs the tbabs cross-section table shipped with soxs; pass "
                    "energy=/cross_section= arrays when soxs is not installed") from e
            energy = np.logspace(np.log10(0.01), np.log10(100.0), 40001)
            nh_probe = 1.0e-3
            cross_section = -np.log(get_tbabs_absorb(energy, nh_probe, abund_table=abund_table))
            cross_section /= nh_probe * 1.0e22
        # cross_section is per H atom in cm^2; nH is in 1e22 cm^-2
        super().__init__(nH, energy, np.asarray(cross_section, dtype="float64") * 1.0e22)
        self.abund_table = abund_table


class WabsModel(AbsorptionModel):
    """Wisconsin absorption, Morrison & McCammon 1983 piecewise polynomial
    (spectral_models.py:612-635 -> soxs ``get_wabs_absorb``), evaluated in-kernel."""

    _name = "wabs"
    _kind = _lib.ABS_WABS

    def __init__(self, nH, abund_table="angr"):
        super().__init__(nH, [], [])
        self.abund_table = abund_table


absorb_models = {"wabs": WabsModel, "tbabs": TBabsModel}
