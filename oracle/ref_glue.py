"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's photon-generation / projection path.  The six native
routines are NOT restated: the reference's own Cython sources are compiled unmodified into
``oracle/_ref`` (see build_ref.py) and called from here -- interp1d_spec, interp2d_spec,
doppler_shift, scatter_events, scatter_events_allsky, pixel_to_cel.  The Python glue around
them cannot be imported (it needs yt / soxs / unyt / h5py, none installed), so it is restated
below in NumPy, statement by statement, each function citing the reference lines it follows
(paths relative to /root/reference).

Parity status: PINNED.
  * The native routines are the reference itself (compiled from its Cython, oracle/_ref);
    pixel_to_cel is additionally checked against the analytic inverse-gnomonic known answer of
    pyxsim/tests/test_pixel_to_cel.py.
  * The NumPy glue in this file is pinned bit for bit against outputs of the reference's OWN
    Python code (ThermalSourceModel.process_data, PionSpectralModel.get_spectrum,
    PowerLaw/LineSourceModel.process_data, AbsorptionModel.absorb_photons, _project_photons
    for eight projection configurations) executed in the build container through
    oracle/ref_import.py (stand-ins for the absent yt/soxs/unyt/h5py that carry no physics);
    fixtures: tests/golden/reference_golden.npz, generator: tests/golden/make_reference_golden.py,
    check: tests/test_cpu_oracle.py::test_restatement_matches_reference_python_bit_for_bit.
  * NOT pinned (third-party arithmetic absent from the reference tree and from this image):
    soxs' APEC/Cloudy table builders and tbabs/wabs cross-sections, yt's Orientation.  Tables
    and sigma(E) are inputs; the wabs polynomial and the Orientation rule restate the published
    algorithms.  The reference's golden HDF5 answers (pyxsim/tests/test_sloshing.py) need
    network data.
"""
import numpy as np

from . import build_ref

K_per_keV = 1.160451812e7
erg_per_keV = 1.602176634e-9
clight_kms = 299792.458
cm_per_kpc = 3.0856775809623245e21
cm2_per_kpc2 = cm_per_kpc * cm_per_kpc
scale_shift = -1.0 / clight_kms           # pyxsim/utils.py:10
scale_shift2 = scale_shift * scale_shift  # pyxsim/utils.py:11

_mods = None


def ref():
    """The compiled reference modules {interpolate, sky_functions, spectra}."""
    global _mods
    if _mods is None:
        _mods = build_ref.load()
    return _mods


# ---------------------------------------------------------------------------------------
# spectral models: pyxsim/spectral_models.py
# ---------------------------------------------------------------------------------------
class SpectralInterpolator1D:
    """spectral_models.py:22-47"""

    def __init__(self, tbins, cosmic_spec, metal_spec, var_spec):
        self.tbins = tbins.astype("float64")
        self.cosmic_spec = cosmic_spec
        self.metal_spec = metal_spec
        if var_spec is None:
            self.var_spec = np.zeros((1, 1, 1))
            self.do_var = False
        else:
            self.var_spec = var_spec
            self.do_var = True

    def __call__(self, t_vals):
        x_i = (np.digitize(t_vals, self.tbins) - 1).astype("int32")
        if np.any((x_i == -1) | (x_i == len(self.tbins) - 1)):
            x_i = np.minimum(np.maximum(x_i, 0), len(self.tbins) - 2)
        return ref()["interpolate"].interp1d_spec(
            self.cosmic_spec, self.metal_spec, self.var_spec, t_vals, self.tbins, x_i, self.do_var)


class SpectralInterpolator2D:
    """spectral_models.py:50-82"""

    def __init__(self, tbins, dbins, cosmic_spec, metal_spec, var_spec):
        self.tbins = tbins.astype("float64")
        self.dbins = dbins.astype("float64")
        self.cosmic_spec = cosmic_spec
        self.metal_spec = metal_spec
        if var_spec is None:
            self.var_spec = np.zeros((1, 1, 1))
            self.do_var = False
        else:
            self.var_spec = var_spec
            self.do_var = True

    def __call__(self, t_vals, d_vals):
        x_i = (np.digitize(t_vals, self.tbins) - 1).astype("int32")
        if np.any((x_i == -1) | (x_i == len(self.tbins) - 1)):
            x_i = np.minimum(np.maximum(x_i, 0), len(self.tbins) - 2)
        y_i = (np.digitize(d_vals, self.dbins) - 1).astype("int32")
        if np.any((y_i == -1) | (y_i == len(self.dbins) - 1)):
            y_i = np.minimum(np.maximum(y_i, 0), len(self.dbins) - 2)
        return ref()["interpolate"].interp2d_spec(
            self.cosmic_spec, self.metal_spec, self.var_spec, t_vals, self.tbins, x_i, d_vals,
            self.dbins, y_i, self.do_var)


class TableModel:
    """A prepared 1-D thermal spectral model (ThermalSpectralModel.get_spectrum,
    spectral_models.py:85-99, with the tables prepare_spectrum would have set)."""

    _density_dependence = False

    def __init__(self, ebins, Tvals, cosmic, metal, var=None, logT=False, binscale="linear"):
        self.ebins = np.asarray(ebins, "float64")
        self.emid = 0.5 * (self.ebins[1:] + self.ebins[:-1])
        self.nbins = self.emid.size
        self.binscale = binscale
        self._logT = logT
        self.Tvals = np.asarray(Tvals, "float64")
        self.cosmic_spec, self.metal_spec, self.var_spec = cosmic, metal, var
        self.si = SpectralInterpolator1D(self.Tvals, cosmic, metal, var)

    def _Tconv(self, kT):
        return np.log10(kT * K_per_keV) if self._logT else kT

    def get_spectrum(self, kT):
        kT = np.atleast_1d(self._Tconv(kT))
        return self.si(kT)

    def make_fluxf(self, emin, emax, energy=False):
        """ThermalSpectralModel.make_fluxf, spectral_models.py:101-134 (scipy interp1d, linear,
        bounds_error as scipy defaults it: out-of-table temperatures raise ValueError)."""
        from scipy.interpolate import interp1d

        eidxs = (self.ebins[:-1] > emin) & (self.ebins[1:] < emax)
        emid = self.emid[eidxs]
        if energy:
            cosmic_flux = (self.cosmic_spec[:, eidxs] * emid).sum(axis=-1)
            metal_flux = (self.metal_spec[:, eidxs] * emid).sum(axis=-1)
        else:
            cosmic_flux = self.cosmic_spec[:, eidxs].sum(axis=-1)
            metal_flux = self.metal_spec[:, eidxs].sum(axis=-1)
        cf = interp1d(self.Tvals, cosmic_flux, fill_value=0.0, assume_sorted=True, copy=False)
        mf = interp1d(self.Tvals, metal_flux, fill_value=0.0, assume_sorted=True, copy=False)
        if self.var_spec is not None:
            if energy:
                var_flux = (self.var_spec[:, :, eidxs] * emid).sum(axis=-1)
            else:
                var_flux = self.var_spec[:, :, eidxs].sum(axis=-1)
            vf = interp1d(self.Tvals, var_flux, axis=1, fill_value=0.0, assume_sorted=True, copy=False)
        else:
            def vf(kT):
                pass

        def _fluxf(kT):
            kT = self._Tconv(kT)
            return cf(kT), mf(kT), vf(kT)

        return _fluxf


class PionModel:
    """PionSpectralModel.get_spectrum / _get_spectrum_2d, spectral_models.py:429-462"""

    _density_dependence = True

    def __init__(self, ebins, Tvals, Dvals, cosmic, metal, var, cie_model, binscale="linear"):
        self.ebins = np.asarray(ebins, "float64")
        self.emid = 0.5 * (self.ebins[1:] + self.ebins[:-1])
        self.nbins = self.emid.size
        self.binscale = binscale
        self.Tvals, self.Dvals = np.asarray(Tvals, "float64"), np.asarray(Dvals, "float64")
        self.cosmic_spec, self.metal_spec = cosmic, metal
        self.var_spec = var
        self.nvar_elem = 0 if var is None else var.shape[0]
        self.n_T, self.n_D = self.Tvals.size, self.Dvals.size
        self.dTvals, self.dDvals = np.diff(self.Tvals), np.diff(self.Dvals)
        self.min_table_kT = 10 ** self.Tvals[0] / K_per_keV
        self.max_table_kT = 10 ** self.Tvals[-1] / K_per_keV
        self.min_table_nH = 10 ** self.Dvals[0]
        self.max_table_nH = 10 ** self.Dvals[-1]
        self.cie_model = cie_model
        self.si = SpectralInterpolator2D(self.Tvals, self.Dvals, cosmic, metal, var)

    def get_spectrum(self, kT, nH):
        kT = np.atleast_1d(kT)
        nH = np.atleast_1d(nH)
        use_igm = (kT >= self.min_table_kT) & (kT <= self.max_table_kT)
        use_igm &= (nH >= self.min_table_nH) & (nH <= self.max_table_nH)
        use_cie = ~use_igm
        cspec = np.zeros((kT.size, self.nbins))
        mspec = np.zeros((kT.size, self.nbins))
        vspec = np.zeros((self.nvar_elem, kT.size, self.nbins)) if self.var_spec is not None else None
        if use_igm.sum() > 0:
            nHi = nH[use_igm]
            lkT = np.atleast_1d(np.log10(kT[use_igm] * K_per_keV))
            lnH = np.atleast_1d(np.log10(nHi))
            c1, m1, v1 = self.si(lkT, lnH)
            cspec[use_igm, :] = c1 / nHi[:, np.newaxis]
            mspec[use_igm, :] = m1 / nHi[:, np.newaxis]
            if self.var_spec is not None:
                vspec[:, use_igm, :] = v1 / nHi[np.newaxis, :, np.newaxis]
        if use_cie.sum() > 0:
            c2, m2, v2 = self.cie_model.get_spectrum(kT[use_cie])
            cspec[use_cie, :] = c2
            mspec[use_cie, :] = m2
            if self.var_spec is not None:
                vspec[:, use_cie, :] = v2
        return cspec, mspec, vspec


def _pion_make_fluxf(self, emin, emax, energy=False):
    """PionSpectralModel.make_fluxf + _get_flux_2d, spectral_models.py:464-537"""
    eidxs = (self.ebins[:-1] > emin) & (self.ebins[1:] < emax)
    emid = self.emid[eidxs]
    if energy:
        cosmic_flux = (self.cosmic_spec[:, eidxs] * emid).sum(axis=-1)
        metal_flux = (self.metal_spec[:, eidxs] * emid).sum(axis=-1)
    else:
        cosmic_flux = self.cosmic_spec[:, eidxs].sum(axis=-1)
        metal_flux = self.metal_spec[:, eidxs].sum(axis=-1)
    if self.var_spec is not None:
        var_flux = (self.var_spec[:, :, eidxs] * emid).sum(axis=-1) if energy else \
            self.var_spec[:, :, eidxs].sum(axis=-1)
    else:
        var_flux = None
    cie_fluxf = self.cie_model.make_fluxf(emin, emax, energy=energy)

    def get_flux_2d(kT, nH, cf, mf, vf):
        lkT = np.atleast_1d(np.log10(kT * K_per_keV))
        lnH = np.atleast_1d(np.log10(nH))
        tidxs = np.searchsorted(self.Tvals, lkT) - 1
        didxs = np.searchsorted(self.Dvals, lnH) - 1
        dT = (lkT - self.Tvals[tidxs]) / self.dTvals[tidxs]
        dn = (lnH - self.Dvals[didxs]) / self.dDvals[didxs]
        idx1 = np.ravel_multi_index((didxs + 1, tidxs + 1), (self.n_D, self.n_T))
        idx2 = np.ravel_multi_index((didxs + 1, tidxs), (self.n_D, self.n_T))
        idx3 = np.ravel_multi_index((didxs, tidxs + 1), (self.n_D, self.n_T))
        idx4 = np.ravel_multi_index((didxs, tidxs), (self.n_D, self.n_T))
        dx1 = dT * dn
        dx2 = dn - dx1
        dx3 = dT - dx1
        dx4 = 1.0 + dx1 - dT - dn
        cflux = dx1 * cf[idx1] + dx2 * cf[idx2] + dx3 * cf[idx3] + dx4 * cf[idx4]
        mflux = dx1 * mf[idx1] + dx2 * mf[idx2] + dx3 * mf[idx3] + dx4 * mf[idx4]
        if vf is not None:
            vflux = dx1[np.newaxis, :] * vf[:, idx1]
            vflux += dx2[np.newaxis, :] * vf[:, idx2]
            vflux += dx3[np.newaxis, :] * vf[:, idx3]
            vflux += dx4[np.newaxis, :] * vf[:, idx4]
        else:
            vflux = None
        return cflux, mflux, vflux

    def _fluxf(kT, nH):
        kT = np.atleast_1d(kT)
        nH = np.atleast_1d(nH)
        use_igm = (kT >= self.min_table_kT) & (kT <= self.max_table_kT)
        use_igm &= (nH >= self.min_table_nH) & (nH <= self.max_table_nH)
        use_cie = ~use_igm
        cflux = np.zeros(kT.size)
        mflux = np.zeros(kT.size)
        vflux = np.zeros((self.nvar_elem, kT.size)) if self.var_spec is not None else None
        if use_igm.sum() > 0:
            nHi = nH[use_igm]
            c1, m1, v1 = get_flux_2d(kT[use_igm], nHi, cosmic_flux, metal_flux, var_flux)
            cflux[use_igm] = c1 / nHi
            mflux[use_igm] = m1 / nHi
            if self.var_spec is not None:
                vflux[:, use_igm] = v1 / nHi[np.newaxis, :]
        if use_cie.sum() > 0:
            c2, m2, v2 = cie_fluxf(kT[use_cie])
            cflux[use_cie] = c2
            mflux[use_cie] = m2
            if self.var_spec is not None:
                vflux[:, use_cie] = v2
        return cflux, mflux, vflux

    return _fluxf


PionModel.make_fluxf = _pion_make_fluxf


def thermal_field_mode(model, cfg, a, mode, emin, emax, fluxf=None, shift=None):
    """The per-cell field modes of ThermalSourceModel.process_data (base.py:304-332,497-511,526):
    "luminosity" / "photon_rate" (and the unshifted intensities) through ``fluxf``; shifted
    "intensity" / "photon_intensity" through the compiled reference ``make_band`` on the clipped
    spectrum.  Returns the array over ALL cells (0 where the cut removes a cell).  ``shift``: per
    selected cell; unlike the reference (which hands make_band the whole chunk's shift array for
    every 100-cell tile, base.py:499) each tile gets its own slice."""
    cut, kT, cell_nrm, nH, metalZ, elemZ = thermal_select(cfg, a)
    n = kT.size
    ret = np.zeros(cut.size)
    idxs = np.where(cut)[0]
    shifted = mode.endswith("intensity") and shift is not None
    for b, e in _chunked(n):
        if shifted:
            if model._density_dependence:
                cspec, mspec, vspec = model.get_spectrum(kT[b:e], nH[b:e])
            else:
                cspec, mspec, vspec = model.get_spectrum(kT[b:e])
            t = cspec
            t += metalZ[b:e, np.newaxis] * mspec
            if elemZ is not None:
                t += np.sum(elemZ[:, b:e, np.newaxis] * vspec, axis=0)
            np.clip(t, 0.0, None, out=t)
            I = ref()["spectra"].make_band(int(mode == "intensity"), emin, emax, model.ebins, model.emid, t,
                                           np.ascontiguousarray(shift[b:e]))
            ret[idxs[b:e]] = I * cell_nrm[b:e]
        else:
            if model._density_dependence:
                cflux, mflux, vflux = fluxf(kT[b:e], nH[b:e])
            else:
                cflux, mflux, vflux = fluxf(kT[b:e])
            tot_flux = cflux
            tot_flux += metalZ[b:e] * mflux
            if elemZ is not None:
                tot_flux += np.sum(elemZ[:, b:e] * vflux, axis=0)
            ret[idxs[b:e]] = tot_flux * cell_nrm[b:e]
    return ret


# ---------------------------------------------------------------------------------------
# thermal source: pyxsim/source_models/thermal_sources/base.py:304-524 (photons mode)
# ---------------------------------------------------------------------------------------
def _chunked(n, size=100):
    for s in range(0, n, size):
        yield s, min(s + size, n)


def thermal_select(cfg, a):
    """Cut mask, cell_nrm, abundances: base.py:335-425.  ``cfg`` keys: kT_min, kT_max,
    max_density, min_entropy, Zmet (number or array key), Zconvert, z_by_xh, h_fraction
    (number or key), var_elem (list of (value|key, mconvert, by_xh)), nei, r2 (array|None)."""
    kT = np.ravel(a["kT"])
    cut = np.ones(kT.size, dtype=bool)
    if cfg.get("max_density") is not None:
        cut &= np.ravel(a["density"]) < cfg["max_density"]
    if cfg.get("min_entropy") is not None:
        cut &= np.ravel(a["entropy"]) > cfg["min_entropy"]
    cut &= (kT >= cfg["kT_min"]) & (kT <= cfg["kT_max"])
    cell_nrm = np.ravel(a["em"] * cfg["spectral_norm"])
    nH = np.ravel(a["nH"]) if a.get("nH") is not None else None
    hf = cfg.get("h_fraction", 0.7)
    X_H = np.ravel(a[hf]) if isinstance(hf, str) else hf
    num_cells = int(cut.sum())
    kT = kT[cut]
    cell_nrm = cell_nrm[cut]
    if nH is not None:
        nH = nH[cut]
    X_Hc = X_H if np.isscalar(X_H) else X_H[cut]
    if cfg.get("nei"):
        metalZ = np.zeros(num_cells)
    elif not isinstance(cfg["Zmet"], str):
        metalZ = cfg["Zmet"] * np.ones(num_cells)
    else:
        mZ = np.ravel(a[cfg["Zmet"]])
        fac = cfg.get("Zconvert", 1.0)
        if cfg.get("z_by_xh"):
            fac = fac / X_H            # q6: per-cell X_H against the uncut field
        metalZ = (mZ * fac)[cut]
    elemZ = None
    var = cfg.get("var_elem") or []
    if len(var) > 0:
        elemZ = np.zeros((len(var), num_cells))
        for j, (value, mconv, by_xh) in enumerate(var):
            if not isinstance(value, str):
                elemZ[j, :] = value
            else:
                eZ = np.ravel(a[value])
                fac = mconv
                if by_xh:
                    fac = fac / X_H
                elemZ[j, :] = (eZ * fac)[cut]
    if a.get("r2") is not None:
        cell_nrm = cell_nrm / np.ravel(a["r2"])[cut]     # base.py:423-425
    return cut, kT, cell_nrm, nH, metalZ, elemZ


def thermal_spectra(model, cfg, a):
    """Per selected cell: clipped tot_spec (base.py:453-460) and lambda = spec_sum*cnm
    (:463-464).  Returns (cut, tot_spec[nsel, nbins], lam[nsel])."""
    cut, kT, cell_nrm, nH, metalZ, elemZ = thermal_select(cfg, a)
    n = kT.size
    tot = np.zeros((n, model.nbins))
    for b, e in _chunked(n):
        if model._density_dependence:
            cspec, mspec, vspec = model.get_spectrum(kT[b:e], nH[b:e])
        else:
            cspec, mspec, vspec = model.get_spectrum(kT[b:e])
        t = cspec
        t += metalZ[b:e, np.newaxis] * mspec
        if elemZ is not None:
            t += np.sum(elemZ[:, b:e, np.newaxis] * vspec, axis=0)
        np.clip(t, 0.0, None, out=t)
        tot[b:e] = t
    lam = tot.sum(axis=-1) * cell_nrm
    return cut, tot, lam


def thermal_spectrum_mode(model, cfg, a, ebins_new, shift=None):
    """mode="spectrum" of ThermalSourceModel.process_data (base.py:436-460,494-495): per 100-cell
    tile ``spec += shift_spectrum(self.ebins, ebins, tot_spec, shift, cnm)`` with the compiled
    reference routine (spectra.pyx:67-93).  ``shift``: per selected cell, default ones."""
    cut, kT, cell_nrm, nH, metalZ, elemZ = thermal_select(cfg, a)
    n = kT.size
    spec = np.zeros(ebins_new.size - 1)
    sh = np.ones(n) if shift is None else np.ascontiguousarray(shift, dtype="float64")
    fn = ref()["spectra"].shift_spectrum
    for b, e in _chunked(n):
        if model._density_dependence:
            cspec, mspec, vspec = model.get_spectrum(kT[b:e], nH[b:e])
        else:
            cspec, mspec, vspec = model.get_spectrum(kT[b:e])
        t = cspec
        t += metalZ[b:e, np.newaxis] * mspec
        if elemZ is not None:
            t += np.sum(elemZ[:, b:e, np.newaxis] * vspec, axis=0)
        np.clip(t, 0.0, None, out=t)
        spec += fn(model.ebins, ebins_new, t, np.ascontiguousarray(sh[b:e]), np.ascontiguousarray(cell_nrm[b:e]))
    return spec


def thermal_process_data(model, cfg, a, prng, method="invert_cdf", literal_q3=False):
    """The photons branch, base.py:427-524: returns (ncells, number_of_photons[active],
    idxs[active], energies) exactly like the reference (RandomState draw order included)."""
    cut, kT, cell_nrm, nH, metalZ, elemZ = thermal_select(cfg, a)
    num_cells = kT.size
    if num_cells == 0:
        return None
    bin_edges = np.log10(model.ebins) if model.binscale == "log" else model.ebins
    number_of_photons = np.zeros(num_cells, dtype="int64")
    energies = []
    idxs = np.where(cut)[0]
    for b, e in _chunked(num_cells):
        cnm = cell_nrm[b:e]
        if model._density_dependence:
            cspec, mspec, vspec = model.get_spectrum(kT[b:e], nH[b:e])
        else:
            cspec, mspec, vspec = model.get_spectrum(kT[b:e])
        tot_spec = cspec
        tot_spec += metalZ[b:e, np.newaxis] * mspec
        if elemZ is not None:
            tot_spec += np.sum(elemZ[:, b:e, np.newaxis] * vspec, axis=0)
        np.clip(tot_spec, 0.0, None, out=tot_spec)
        spec_sum = tot_spec.sum(axis=-1)
        cell_norm = spec_sum * cnm
        cell_n = np.atleast_1d(prng.poisson(lam=cell_norm))
        number_of_photons[b:e] = cell_n
        with np.errstate(divide="ignore", invalid="ignore"):
            norm_factor = 1.0 / spec_sum
            p = norm_factor[:, np.newaxis] * tot_spec
        cp = np.insert(np.cumsum(p, axis=-1), 0, 0.0, axis=1)
        for icell in range(e - b):
            cn = cell_n[icell]
            if cn == 0:
                continue
            if method == "invert_cdf":
                randvec = prng.uniform(size=cn)
                randvec.sort()
                cell_e = np.interp(randvec, cp[icell, :], bin_edges)
            else:
                eidxs = prng.choice(model.nbins, size=cn, p=p[icell, :])
                cell_e = model.emid[eidxs]
            energies.append(cell_e)
    active = number_of_photons > 0
    ee = np.concatenate(energies) if energies else np.zeros(0)
    # base.py:522-523 applies 10** for log binning whatever the method; with accept_reject the
    # energies are bin centres (not logarithms), so that is a latent bug.
    # literal_q3=True reproduces it (used to pin this restatement against the reference's
    # output); the default returns the bin centres unchanged, like the GPU path.
    if model.binscale == "log" and (method == "invert_cdf" or literal_q3):
        ee = 10 ** ee
    return int(active.sum()), number_of_photons[active], idxs[active], ee


# ---------------------------------------------------------------------------------------
# power law / line: power_law_sources.py:186-250, line_sources.py:193-234
# ---------------------------------------------------------------------------------------
def powerlaw_lambda(cfg, lum_keV_s, alpha):
    """Expected counts Nph (power_law_sources.py:196-222) and norm_fac (:214)."""
    alpha = np.asarray(alpha, "float64") * np.ones_like(lum_keV_s)
    # 0-d arrays like the reference's ``self.emin.v``: NumPy's array power loop (SIMD) and the
    # scalar libm pow can differ in the last bit
    e0, emin, emax = (np.asarray(cfg[k], dtype="float64") for k in ("e0", "emin", "emax"))
    ei, ef = emin, emax
    etoalpha = e0 ** alpha
    K_fac = emax ** (2.0 - alpha) - emin ** (2.0 - alpha)
    K_fac[alpha == 2] = np.log(emax / emin)
    K_fac *= etoalpha
    if np.any(alpha != 2):
        K_fac[alpha != 2] /= 2.0 - alpha[alpha != 2]
    Kn = lum_keV_s / K_fac
    Nph = ef ** (1.0 - alpha) - ei ** (1.0 - alpha)
    Nph[alpha == 1] = np.log(ef / ei)
    Nph *= Kn * etoalpha
    with np.errstate(divide="ignore", invalid="ignore"):
        norm_fac = Nph / Kn
    if np.any(alpha != 1):
        Nph[alpha != 1] /= 1.0 - alpha[alpha != 1]
    Nph *= cfg["spectral_norm"] * cfg["scale_factor"]
    if cfg.get("r2") is not None:
        Nph /= cfg["r2"]
    return Nph, norm_fac, alpha


def powerlaw_process_data(cfg, lum_keV_s, alpha, prng):
    Nph, norm_fac, alpha = powerlaw_lambda(cfg, lum_keV_s, alpha)
    ei, ef = np.asarray(cfg["emin"], dtype="float64"), np.asarray(cfg["emax"], dtype="float64")
    number_of_photons = prng.poisson(lam=Nph)
    energies = np.zeros(number_of_photons.sum())
    start_e = 0
    end_e = 0
    for i in range(Nph.size):
        if number_of_photons[i] > 0:
            end_e = start_e + number_of_photons[i]
            u = prng.uniform(size=number_of_photons[i])
            if alpha[i] == 1:
                e = ei * (ef / ei) ** u
            else:
                e = ei ** (1.0 - alpha[i]) + u * norm_fac[i]
                e **= 1.0 / (1.0 - alpha[i])
            energies[start_e:end_e] = e * cfg["scale_factor"]
            start_e = end_e
    active = number_of_photons > 0
    return int(active.sum()), number_of_photons[active], active, energies[:end_e].copy()


def line_process_data(cfg, rate, sigma_keV, prng):
    """line_sources.py:204-234 (sigma already in keV)."""
    F = rate * cfg["spectral_norm"] * cfg["scale_factor"]
    if cfg.get("r2") is not None:
        F = F / cfg["r2"]
    sigma = np.asarray(sigma_keV, "float64") * np.ones_like(rate)
    number_of_photons = prng.poisson(lam=F)
    energies = cfg["e0"] * np.ones(number_of_photons.sum())
    start_e = 0
    for i in range(rate.size):
        if number_of_photons[i] > 0:
            end_e = start_e + number_of_photons[i]
            energies[start_e:end_e] += prng.normal(loc=0.0, scale=sigma[i], size=number_of_photons[i])
            start_e = end_e
    energies = energies * cfg["scale_factor"]
    active = number_of_photons > 0
    return int(active.sum()), number_of_photons[active], active, energies, F


# ---------------------------------------------------------------------------------------
# geometry helpers
# ---------------------------------------------------------------------------------------
def compute_radius(pos, le, re, dw, c, periodicity):
    """sources.py:77-87 (pos is (3, n) in kpc); note pos[:, tfl] shifts all three rows."""
    pos = np.array(pos, dtype="float64")
    for i in range(3):
        if periodicity[i]:
            tfl = pos[i] < le[i]
            tfr = pos[i] > re[i]
            pos[:, tfl] += dw[i]
            pos[:, tfr] -= dw[i]
    return np.sum((pos - np.asarray(c)[:, np.newaxis]) ** 2, axis=0) * cm2_per_kpc2


def gather_active(pos, vel, dx, idxs, le, re, dw, c, bulk_velocity, periodicity):
    """photon_list.py:423-451: recentred/unwrapped positions and bulk-subtracted velocities
    of the active cells.  pos/vel: lists of three arrays; dx array or None."""
    out = {}
    dxa = 0.0 if dx is None else dx[idxs]
    for i, ax in enumerate("xyz"):
        p = pos[i][idxs].copy()
        if periodicity[i]:
            tfl = p + dxa < le[i]
            tfr = p - dxa > re[i]
            p[tfl] += dw[i]
            p[tfr] -= dw[i]
        out[ax] = p - c[i]
        out[f"v{ax}"] = vel[i][idxs] - bulk_velocity[i]
    out["dx"] = np.zeros(idxs.size) if dx is None else dx[idxs]
    return out


def orientation(normal, north_vector=None):
    """yt.utilities.orientation.Orientation unit vectors [east, north, normal] (third-party,
    restated from its published source; used via pyxsim/utils.py:261-270)."""
    L = np.array(normal, dtype="float64")
    L /= np.sqrt(np.dot(L, L))
    if north_vector is None:
        vecs = np.identity(3)
        t = np.cross(L, vecs).sum(axis=1)
        ax = t.argmax()
        east = np.cross(vecs[ax, :], L).ravel()
        north = np.cross(L, east).ravel()
    else:
        north = np.array(north_vector, dtype="float64")
        if np.dot(north, L) != 0.0:
            north = north - np.dot(north, L) * L
        east = np.cross(north, L).ravel()
    north /= np.sqrt(np.dot(north, north))
    east /= np.sqrt(np.dot(east, east))
    return np.array([east, north, L])


# ---------------------------------------------------------------------------------------
# absorption: spectral_models.py:540-635
# ---------------------------------------------------------------------------------------
class TableAbsorb:
    """AbsorptionModel with a tabulated cross-section (spectral_models.py:540-583).  As in the
    reference the optical depth is literally ``sigma * nH``: with nH in 1e22 cm^-2 the table
    must be given in units of 1e-22 cm^2."""

    def __init__(self, nH, energy, cross_section):
        self._nH = nH
        self.emid = np.asarray(energy, "float64")
        self.sigma = np.asarray(cross_section, "float64")

    def get_absorb(self, e):
        sigma = np.interp(e, self.emid, self.sigma, left=0.0, right=0.0)
        return np.exp(-sigma * self._nH)

    def absorb_photons(self, eobs, prng):
        n_events = eobs.size
        if n_events == 0:
            return np.array([], dtype="bool")
        absorb = self.get_absorb(eobs)
        randvec = prng.uniform(size=n_events)
        return randvec < absorb


_WABS_EDGE = np.array([0.03, 0.1, 0.284, 0.4, 0.532, 0.707, 0.867, 1.303, 1.84, 2.471, 3.21,
                       4.038, 7.111, 8.331, 10.0])
_WABS_C0 = np.array([17.3, 34.6, 78.1, 71.4, 95.5, 308.9, 120.6, 141.3, 202.7, 342.7, 352.2,
                     433.9, 629.0, 701.2])
_WABS_C1 = np.array([608.1, 267.9, 18.8, 66.8, 145.8, -380.6, 169.3, 146.8, 104.7, 18.7, 18.7,
                     -2.4, 30.9, 25.2])
_WABS_C2 = np.array([-2150.0, -476.1, 4.3, -51.4, -61.1, 294.0, -47.7, -31.5, -17.0, 0.0, 0.0,
                     0.75, 0.0, 0.0])


class WabsAbsorb(TableAbsorb):
    """WabsModel (spectral_models.py:612-635) -> soxs get_wabs_absorb: Morrison & McCammon
    (1983, ApJ 270, 119) Table 2 piecewise polynomial, sigma = (c0+c1 E+c2 E^2)/E^3 1e-24 cm^2.
    soxs is third-party and absent: coefficients are from the paper, UNVERIFIED against soxs."""

    def __init__(self, nH):
        self._nH = nH

    def get_absorb(self, e):
        e = np.asarray(e, "float64")
        k = np.clip(np.searchsorted(_WABS_EDGE, e, side="right") - 1, 0, 13)
        sigma = (_WABS_C0[k] + _WABS_C1[k] * e + _WABS_C2[k] * e * e) / e ** 3 * 1.0e-24
        return np.exp(-sigma * self._nH * 1.0e22)


def column_lookup(x, y, z, unit_vectors, wbins, dbins, nH_grid):
    """Internal-absorption column of every cell (photon_list.py:542-546,672-680): the cube is
    sampled at ``pos = unit_vectors . (x, y, z)`` with scipy's ``interpn`` exactly as the
    reference calls it (linear, no bounds error, fill value 0)."""
    from scipy.interpolate import interpn

    wmid = 0.5 * (wbins[1:] + wbins[:-1])
    dmid = 0.5 * (dbins[1:] + dbins[:-1])
    pos = np.dot(unit_vectors, np.array([x, y, z]))
    return interpn((wmid, wmid, dmid), nH_grid, pos.T, bounds_error=False, fill_value=0.0)


# ---------------------------------------------------------------------------------------
# projection: pyxsim/photon_list.py:661-808
# ---------------------------------------------------------------------------------------
def project_photons(d, params, normal, sky_center, prng, absorb_model=None, nH=None,
                    no_shifting=False, north_vector=None, sigma_pos=None, flat_sky=False,
                    kernel="top_hat", save_los=False, phys_coord=False, nH_int=None,
                    cell_chunk=100000):
    """``d``: photon-list arrays {x,y,z,vx,vy,vz,dx,num_photons,energy}; ``params``:
    {data_type, observer, fid_d_a[Mpc], fid_redshift}.  Returns dict(xsky, ysky, eobs, los,
    n_events, flux_sum, emin, emax).  Statement-by-statement restatement of the tile loop."""
    sf = ref()["sky_functions"]
    observer = params.get("observer", "external")
    data_type = params["data_type"]
    D_A = params["fid_d_a"] * 1.0e3
    oneplusz = 1.0 + params["fid_redshift"]
    if isinstance(normal, str):
        L = np.zeros(3)
        L["xyz".index(normal)] = 1.0
        norm = "xyz".index(normal)
    else:
        L = np.array(normal, dtype="float64")
        norm = L
    uv = orientation(L, north_vector)
    x_hat, y_hat, z_hat = uv[0].copy(), uv[1].copy(), uv[2].copy()
    sky_center = np.asarray(sky_center, "float64")
    n_cells = d["num_photons"].size
    out = {"xsky": [], "ysky": [], "eobs": [], "los": []}
    flux, emin, emax = 0.0, 1.0e99, -1.0e99
    n_events = 0
    start_e = 0
    for start_c in range(0, n_cells, cell_chunk):
        end_c = min(start_c + cell_chunk, n_cells)
        n_ph = np.ascontiguousarray(d["num_photons"][start_c:end_c])
        x = np.ascontiguousarray(d["x"][start_c:end_c])
        y = np.ascontiguousarray(d["y"][start_c:end_c])
        z = np.ascontiguousarray(d["z"][start_c:end_c])
        dx = np.array(d["dx"][start_c:end_c])
        end_e = start_e + n_ph.sum()
        eobs = np.array(d["energy"][start_e:end_e])
        r = np.sqrt(x * x + y * y + z * z) if observer == "internal" else None
        if not no_shifting:
            vx, vy, vz = (d[k][start_c:end_c] for k in ("vx", "vy", "vz"))
            if observer == "internal":
                vn = -(vx * x + vy * y + vz * z) / r
            elif isinstance(normal, str):
                vn = d[f"v{normal}"][start_c:end_c]
            else:
                vn = vx * z_hat[0] + vy * z_hat[1] + vz * z_hat[2]
            v2 = vx * vx + vy * vy + vz * vz
            sf.doppler_shift(np.ascontiguousarray(vn * scale_shift),
                             np.ascontiguousarray(v2 * scale_shift2), n_ph, eobs)
        det = np.ones(eobs.size, dtype="bool")
        if absorb_model is not None:
            if nH_int is not None:
                e_begin = 0
                nhc = nH_int[start_c:end_c]
                for i in range(n_ph.size):
                    e_end = e_begin + n_ph[i]
                    absorb_model._nH = nhc[i]
                    det[e_begin:e_end] &= absorb_model.absorb_photons(
                        eobs[e_begin:e_end] * oneplusz, prng)
                    e_begin += n_ph[i]
            if nH is not None:
                absorb_model._nH = nH
                det &= absorb_model.absorb_photons(eobs, prng)
        num_det = np.int64(det.sum())
        if num_det > 0:
            if observer == "external":
                if data_type == "particles":
                    dx *= 0.5
                xsky, ysky, los = sf.scatter_events(norm, prng, kernel, data_type, num_det, det,
                                                    n_ph, x, y, z, dx, x_hat, y_hat, z_hat)
                if data_type == "cells" and sigma_pos is not None:
                    sigma = sigma_pos * np.repeat(dx, n_ph)[det]
                    xsky += sigma * prng.normal(loc=0.0, scale=1.0, size=num_det)
                    ysky += sigma * prng.normal(loc=0.0, scale=1.0, size=num_det)
                if not phys_coord:
                    xsky /= D_A
                    ysky /= D_A
                    if flat_sky:
                        xsky = sky_center[0] - np.rad2deg(xsky)
                        ysky = sky_center[1] + np.rad2deg(ysky)
                    else:
                        sf.pixel_to_cel(xsky, ysky, sky_center)
            else:
                xsky, ysky, los = sf.scatter_events_allsky(data_type, kernel, prng, num_det, det,
                                                          n_ph, x, y, z, dx, x_hat, y_hat, z_hat)
            out["xsky"].append(xsky)
            out["ysky"].append(ysky)
            out["eobs"].append(eobs[det])
            out["los"].append(los)
            flux += eobs[det].sum()
            emin = min(eobs[det].min(), emin)
            emax = max(eobs[det].max(), emax)
            n_events += int(num_det)
        start_e = end_e
    res = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in out.items()}
    res.update(n_events=n_events, flux_sum=flux, emin=emin, emax=emax)
    return res


def tan_known_answer(x, y, ra0_deg, dec0_deg):
    """Analytic inverse gnomonic (TAN) projection with xi = -x, eta = y: the known answer the
    reference pins pixel_to_cel to through astropy WCS in pyxsim/tests/test_pixel_to_cel.py."""
    a0, d0 = np.deg2rad(ra0_deg), np.deg2rad(dec0_deg)
    xi, eta = -x, y
    den = np.cos(d0) - eta * np.sin(d0)
    ra = a0 + np.arctan2(xi, den)
    dec = np.arctan2(np.sin(d0) + eta * np.cos(d0), np.hypot(xi, den))
    return np.rad2deg(ra), np.rad2deg(dec)


# ---------------------------------------------------------------------------------------
# event-list consumers: pyxsim/event_list.py:106-134,254-318,344-349
# ---------------------------------------------------------------------------------------
def world2pix_tan(ra_deg, dec_deg, sky_center, nx, dtheta):
    """``wcs.wcs_world2pix(ra, dec, 1)`` for the header EventList builds (event_list.py:106-112):
    CTYPE RA---TAN/DEC--TAN, CRVAL = sky_center, CRPIX = (nx+1)/2, CDELT = (-dtheta, dtheta).
    astropy/wcslib is third-party and absent (pyproject.toml: astropy unpinned): this is the
    published gnomonic projection (Calabretta & Greisen 2002, eqs. 5, 54-55) -- the inverse of
    the reference's own pixel_to_cel, which tests/ use to pin it."""
    a0, d0 = np.deg2rad(sky_center[0]), np.deg2rad(sky_center[1])
    a, d = np.deg2rad(ra_deg), np.deg2rad(dec_deg)
    D = np.sin(d0) * np.sin(d) + np.cos(d0) * np.cos(d) * np.cos(a - a0)
    x = np.rad2deg(np.cos(d) * np.sin(a - a0) / D)
    y = np.rad2deg((np.cos(d0) * np.sin(d) - np.sin(d0) * np.cos(d) * np.cos(a - a0)) / D)
    crpix = 0.5 * (nx + 1)
    return crpix + x / (-dtheta), crpix + y / dtheta


def event_image(xsky, ysky, eobs, sky_center, fov_arcmin, nx, emin=None, emax=None, pix=None):
    """EventList.write_fits_image's array (event_list.py:273-295): returns H.T.  ``pix``: use
    these pixel coordinates instead of world2pix_tan's (to test the binning in isolation)."""
    emin = -1.0 if emin is None else emin
    emax = 1.0e10 if emax is None else emax
    dtheta = fov_arcmin / 60.0 / nx
    xbins = np.linspace(0.5, float(nx) + 0.5, nx + 1, endpoint=True)
    ybins = np.linspace(0.5, float(nx) + 0.5, nx + 1, endpoint=True)
    mask = np.logical_and(eobs >= emin, eobs <= emax)
    if pix is None:
        xx, yy = world2pix_tan(xsky[mask], ysky[mask], sky_center, nx, dtheta)
    else:
        xx, yy = pix[0][mask], pix[1][mask]
    H = np.histogram2d(xx, yy, bins=[xbins, ybins])[0]
    return H.T


def event_spectrum(eobs, emin, emax, nchan):
    """EventList.write_spectrum's counts (event_list.py:344-349)."""
    ebins = np.linspace(emin, emax, nchan + 1, endpoint=True)
    return ebins, np.histogram(eobs, bins=ebins)[0]
