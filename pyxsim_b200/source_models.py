"""Source models: the plugin interface ``make_photons`` drives
(``setup_model / set_pv / process_data / cleanup_model``; pyxsim/photon_list.py:316-343,402,469).

Host mirror of ``pyxsim/source_models``: constructor signatures, attribute names and the
``process_data("photons", chunk, spectral_norm)`` return tuple are the reference's; the work
(table lookup, expected counts, Poisson draws, offsets, energy sampling, active-cell gather)
runs in libpxb on the GPU.  Only ``mode="photons"`` is implemented (the hot path); the
spectrum / intensity / field modes are "next" rows.
"""
import ctypes as C
import os
from numbers import Number, Real

import numpy as np
import torch

from . import _lib, constants as K
from .data import ArrayChunk, ArrayDataSource
from .device import empty, ptr, stream_ptr, to_device, require_cuda
from .profiling import stage
from .utils import mylog, parse_prng, parse_value, prng_seed


def field_value(chunk, field, unit=None, equiv=None):
    """``chunk[field]`` in ``unit``: ArrayChunk, or a yt chunk whose fields carry units."""
    if isinstance(chunk, ArrayChunk):
        return chunk.to_value(field, unit)
    v = chunk[field]
    if hasattr(v, "to_value") and unit is not None:
        return v.to_value(unit, equiv) if equiv else v.to_value(unit)
    return getattr(v, "d", v)


def field_units(chunk, field):
    if isinstance(chunk, ArrayChunk):
        return chunk.units(field)
    return str(chunk[field].units)


class ChunkPhotons:
    """Device-resident result of one chunk: the photon list rows of its active cells."""

    __slots__ = ("n_cells", "n_photons", "idx", "nph", "poff", "energy", "x", "y", "z", "vx",
                 "vy", "vz", "dx", "counts", "gid", "_id0", "_chunk_cells", "qoff", "n_pairs")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


class SourceModel:
    """pyxsim/source_models/sources.py:20-99."""

    def __init__(self, prng=None):
        self.spectral_norm = None
        self.redshift = None
        self.prng = parse_prng(prng)
        self._seed = prng_seed(prng)
        self.observer = "external"
        self.ftype = "gas"
        self.p_fields = self.v_fields = None
        self.le = self.re = self.dw = self.c = self.periodicity = None
        self._cell_cursor = 0
        self._scan_ws = None

    # -- plugin interface -----------------------------------------------------------------
    def setup_model(self, mode, data_source, redshift):
        self.redshift = redshift
        self._cell_cursor = 0

    def set_pv(self, p_fields, v_fields, le=None, re=None, dw=None, c=None, periodicity=None,
               observer=None):
        self.p_fields = p_fields
        self.v_fields = v_fields
        self.le = le
        self.re = re
        self.dw = dw
        self.c = c
        self.periodicity = periodicity
        self.observer = observer

    def cleanup_model(self, mode):
        self._scan_ws = None

    def process_data(self, mode, chunk, spectral_norm, ebins=None, emin=None, emax=None, fluxf=None,
                     shifting=False):
        """Reference contract (photon_list.py:402-405): ``None`` or ``(ncells_active,
        number_of_photons[int64], idxs, energies[float64 keV])`` as NumPy arrays for
        ``mode="photons"``; ``mode="spectrum"`` returns the chunk's spectrum on ``ebins``."""
        if mode == "spectrum":
            return self._spectrum_chunk(chunk, spectral_norm, np.asarray(ebins, "float64"), shifting)
        if mode in ("luminosity", "photon_rate", "intensity", "photon_intensity"):
            # per-cell band quantities of make_source_fields / make_intensity_fields
            # (sources.py:236-306,447-500; base.py:497-511): NumPy array shaped like the chunk
            out = self._field_chunk(mode, chunk, spectral_norm, emin, emax, fluxf, shifting)
            return out.cpu().numpy()
        if mode != "photons":
            raise NotImplementedError(
                f"pyxsim_b200 implements the modes 'photons', 'spectrum', 'luminosity', 'photon_rate', "
                f"'intensity' and 'photon_intensity' (got {mode!r})")
        out = self.generate_chunk(chunk, spectral_norm)
        if out is None:
            return None
        return (out.n_cells, out.nph.cpu().numpy(), out.idx.cpu().numpy(),
                out.energy.cpu().numpy())

    # -- helpers shared by all models -----------------------------------------------------
    def _geometry(self, bulk_velocity=None):
        G = _lib.Geometry()
        for i in range(3):
            G.c[i] = 0.0 if self.c is None else float(self.c[i])
            G.le[i] = -1e300 if self.le is None else float(self.le[i])
            G.re[i] = 1e300 if self.re is None else float(self.re[i])
            G.dw[i] = 0.0 if self.dw is None else float(self.dw[i])
            G.periodic[i] = int(bool(self.periodicity[i])) if self.periodicity is not None else 0
            G.bulk_v[i] = 0.0 if bulk_velocity is None else float(bulk_velocity[i])
        return G

    def _chunk_size(self, chunk, field):
        v = chunk[field]
        return int(np.prod(v.shape)) if hasattr(v, "shape") else len(v)

    def _cell_ids(self, Cs, chunk, n):
        """Fill the global-id triple of a cells struct (cell_id0, id_block, id_stride): the RNG
        counters of cell i are keyed by its GLOBAL id so results do not depend on chunking or on
        how the dataset is sharded over ranks.  yt chunks get a running cursor."""
        if isinstance(chunk, ArrayChunk):
            Cs.cell_id0, Cs.id_block, Cs.id_stride = chunk.id_map()
        else:
            Cs.cell_id0, Cs.id_block, Cs.id_stride = self._cell_cursor, 0, 0
            self._cell_cursor += n
        self._last_id_map = (Cs.cell_id0, Cs.id_block, Cs.id_stride)

    def _global_ids(self, idx):
        """Global ids of the chunk-local active-cell indices ``idx`` of the chunk last processed."""
        id0, blk, stride = self._last_id_map
        if blk > 0:
            return id0 + (idx // blk) * stride + idx % blk
        return idx + id0

    def compute_radius(self, chunk):
        """r^2 in cm^2 on the device (sources.py:77-87)."""
        pos = [to_device(field_value(chunk, self.p_fields[i], "kpc")) for i in range(3)]
        n = pos[0].numel()
        r2 = empty(n)
        G = self._geometry()
        _lib.call("pxb_radius2", n, ptr(pos[0]), ptr(pos[1]), ptr(pos[2]), C.byref(G), ptr(r2),
                  stream_ptr())
        return r2

    def _scan_and_compact(self, counts, chunk, gather):
        """counts -> ChunkPhotons without energies (offsets, active cells, gathered geometry)."""
        n = counts.numel()
        need = C.c_int64()
        _lib.call("pxb_scan_workspace_bytes", n, C.byref(need))
        if self._scan_ws is None or self._scan_ws.numel() < need.value:
            self._scan_ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=counts.device)
        offsets = empty(n + 1, torch.int64)
        arank = empty(n + 1, torch.int64)
        pairoff = empty(n + 1, torch.int64)
        totals = empty(3, torch.int64)
        with stage("scan_counts"):
            _lib.call("pxb_scan_counts", ptr(counts), n, ptr(offsets), ptr(arank), ptr(pairoff),
                      ptr(totals), ptr(self._scan_ws), self._scan_ws.numel(), stream_ptr())
        n_photons, n_active, n_pairs = (int(v) for v in totals.cpu())  # the one sync of the chunk
        out = ChunkPhotons()
        out.n_cells, out.n_photons, out.n_pairs = n_active, n_photons, n_pairs
        out.counts = counts
        if n_photons == 0:
            return out
        out.idx = empty(n_active, torch.int64)
        out.nph = empty(n_active, torch.int64)
        out.poff = empty(n_active + 1, torch.int64)
        out.qoff = empty(n_active + 1, torch.int64)
        x = y = z = vx = vy = vz = dx = None
        G = self._geometry()
        if gather is not None:
            G = self._geometry(gather.get("bulk_velocity"))
            pf, vf, wf = gather["p_fields"], gather["v_fields"], gather["w_field"]
            x, y, z = (to_device(field_value(chunk, f, "kpc")) for f in pf)
            vx, vy, vz = (to_device(field_value(chunk, f, "km/s")) for f in vf)
            if wf is not None:
                dx = to_device(field_value(chunk, wf, "kpc"))
            for name in ("x", "y", "z", "vx", "vy", "vz", "dx"):
                setattr(out, name, empty(n_active))
        with stage("compact_cells"):
            _lib.call("pxb_compact_cells", n, ptr(counts), ptr(offsets), ptr(arank), ptr(pairoff),
                      ptr(x), ptr(y), ptr(z), ptr(vx), ptr(vy), ptr(vz), ptr(dx), C.byref(G),
                      ptr(out.idx), ptr(out.nph), ptr(out.poff), ptr(out.qoff), ptr(out.x), ptr(out.y),
                      ptr(out.z), ptr(out.vx), ptr(out.vy), ptr(out.vz), ptr(out.dx), stream_ptr())
        return out

    def generate_chunk(self, chunk, spectral_norm, gather=None):
        raise NotImplementedError

    # -- spectrum mode ("next" row f1) --------------------------------------------------------
    def _spectrum_chunk(self, chunk, spectral_norm, ebins, shifting):
        raise NotImplementedError

    def _field_chunk(self, mode, chunk, spectral_norm, emin, emax, fluxf, shifting):
        raise NotImplementedError(f"{type(self).__name__} has no per-cell field modes on the GPU yet")

    def make_fluxf(self, emin, emax, energy=False):
        raise NotImplementedError

    def make_source_fields(self, data_source, emin, emax, force_override=False, band_name=None):
        """Source-frame band fields of an :class:`ArrayDataSource` (sources.py:166-308):
        ``xray_luminosity_*`` [erg/s], ``xray_emissivity_*`` [erg/cm^3/s], ``xray_count_rate_*``
        [photons/s], ``xray_photon_emissivity_*`` [photons/cm^3/s] as device tensors added to
        ``data_source.fields``; returns the field names in the reference's order.  (For yt
        datasets the reference registers lazy derived fields with ``ds.add_field``; that
        registration is yt plumbing, the per-chunk evaluation is ``process_data``.)"""
        from .photon_list import _chunks

        require_cuda()
        emin, emax = parse_value(emin, "keV"), parse_value(emax, "keV")
        self.setup_model("fields", data_source, 0.0)
        ftype = self.ftype
        if band_name is None:
            band_name = f"{emin}_{emax}_keV"
        emiss_name = (ftype, f"xray_emissivity_{band_name}")
        lum_name = (ftype, f"xray_luminosity_{band_name}")
        phot_emiss_name = (ftype, f"xray_photon_emissivity_{band_name}")
        count_rate_name = (ftype, f"xray_count_rate_{band_name}")
        names = [emiss_name, lum_name, phot_emiss_name, count_rate_name]
        if not force_override:
            clash = [n for n in names if n in data_source.fields]
            if clash:
                raise RuntimeError(f"fields {clash} exist already (use force_override=True)")
        efluxf = self.make_fluxf(emin, emax, energy=True)
        pfluxf = self.make_fluxf(emin, emax, energy=False)
        lum, rate, idv = [], [], []
        for chunk in _chunks(data_source):
            lum.append(self._field_chunk("luminosity", chunk, 1.0, emin, emax, efluxf, False))
            rate.append(self._field_chunk("photon_rate", chunk, 1.0, emin, emax, pfluxf, False))
            idv.append(_inverse_volume(chunk, ftype))
        lum = torch.cat(lum) * K.erg_per_keV      # keV/s -> erg/s
        rate, idv = torch.cat(rate), torch.cat(idv)
        data_source.fields[lum_name] = lum
        data_source.fields[emiss_name] = lum * idv
        data_source.fields[count_rate_name] = rate
        data_source.fields[phot_emiss_name] = rate * idv
        data_source.units.update({lum_name: "erg/s", emiss_name: "erg/cm**3/s",
                                  count_rate_name: "photons/s", phot_emiss_name: "photons/cm**3/s"})
        self.cleanup_model("fields")
        return names

    def make_intensity_fields(self, data_source, emin, emax, redshift=0.0, dist=None, cosmology=None,
                              no_doppler=False, force_override=True, band_name=None, normal="z"):
        """Observer-frame band intensity fields (sources.py:351-500): ``xray_intensity_*``
        [erg/cm^3/s/arcsec^2] and ``xray_photon_intensity_*`` [photons/cm^3/s/arcsec^2] -- the
        integrands of a projection along ``normal`` (the reference takes the axis from the
        projection's field parameter; an array source has no such mechanism, so it is an
        argument).  ``emin``/``emax`` are observer-frame energies."""
        from .data import Cosmology
        from .photon_list import _chunks

        require_cuda()
        if redshift == 0.0 and dist is None:
            raise ValueError("Either 'redshift' must be > 0.0 or 'dist' must not be None!")
        emin, emax = parse_value(emin, "keV"), parse_value(emax, "keV")
        self.setup_model("fields", data_source, 0.0)
        ftype = self.ftype
        # _make_dist_fac(per_sa=True), sources.py:107-130
        if dist is None:
            cosmo = cosmology or getattr(data_source, "cosmology", None) or Cosmology()
            d_a = cosmo.angular_diameter_distance_mpc(0.0, redshift) * K.cm_per_mpc
            d_l = d_a * (1.0 + redshift) ** 2
            angular_scale = d_a                      # cm per radian
        else:
            redshift = 0.0
            d_l = parse_value(dist, "kpc") * K.cm_per_kpc
            angular_scale = d_l
        rad2_per_arcsec2 = (np.pi / 180.0 / 3600.0) ** 2
        dist_fac = 1.0 / (4.0 * np.pi * d_l * d_l) * angular_scale ** 2 * rad2_per_arcsec2
        emin_src, emax_src = emin * (1.0 + redshift), emax * (1.0 + redshift)
        if band_name is None:
            band_name = f"{emin}_{emax}_keV"
        ei_name = (ftype, f"xray_intensity_{band_name}")
        i_name = (ftype, f"xray_photon_intensity_{band_name}")
        if not force_override and (ei_name in data_source.fields or i_name in data_source.fields):
            raise RuntimeError(f"fields {ei_name}, {i_name} exist already (use force_override=True)")
        efluxf = self.make_fluxf(emin_src, emax_src, energy=True) if no_doppler else None
        pfluxf = self.make_fluxf(emin_src, emax_src, energy=False) if no_doppler else None
        if isinstance(normal, str):
            nv = np.zeros(3)
            nv["xyz".index(normal)] = 1.0
            normal = nv
        self._spec_normal = np.asarray(normal, "float64")
        if isinstance(data_source, ArrayDataSource):
            self.v_fields = data_source.velocity_fields
        ei, pi_, idv = [], [], []
        for chunk in _chunks(data_source):
            ei.append(self._field_chunk("intensity", chunk, 1.0, emin_src, emax_src, efluxf, not no_doppler))
            pi_.append(self._field_chunk("photon_intensity", chunk, 1.0, emin_src, emax_src, pfluxf,
                                         not no_doppler))
            idv.append(_inverse_volume(chunk, ftype))
        idv = torch.cat(idv)
        data_source.fields[ei_name] = torch.cat(ei) * idv * (dist_fac * K.erg_per_keV)
        data_source.fields[i_name] = torch.cat(pi_) * idv * ((1.0 + redshift) * dist_fac)
        data_source.units.update({ei_name: "erg/cm**3/s/arcsec**2", i_name: "photons/cm**3/s/arcsec**2"})
        self.cleanup_model("fields")
        return [ei_name, i_name]

    def compute_shift(self, chunk):
        """Per-cell Doppler factor sqrt(1-beta^2)/(1-beta_n) along ``self._spec_normal``
        (sources.py:89-95), device tensor."""
        vf = self.v_fields or [(self.ftype, f"velocity_{ax}") for ax in "xyz"]
        v = [to_device(field_value(chunk, f, "km/s")) for f in vf]
        out = empty(v[0].numel())
        nrm = (C.c_double * 3)(*[float(x) for x in self._spec_normal])
        _lib.call("pxb_shift_factor", v[0].numel(), ptr(v[0]), ptr(v[1]), ptr(v[2]), nrm, ptr(out),
                  stream_ptr())
        return out

    def make_spectrum(self, data_source, emin, emax, nbins, redshift=0.0, dist=None, cosmology=None,
                      normal=None):
        """Spectrum of all the data in ``data_source`` (thermal base.py:231-299, power law
        :93-158, line sources alike): returns ``(ebins, spec)`` -- a count-rate spectrum
        [photons/s/keV] in the source frame, or, with ``redshift``/``dist``, the observer-frame
        flux spectrum [photons/s/cm^2/keV].  ``normal`` Doppler-shifts along that direction."""
        from .data import Cosmology
        from .photon_list import _chunks

        require_cuda()
        if isinstance(normal, str):
            normal = "xyz".index(normal)
        if isinstance(normal, (int, np.integer)):
            nv = np.zeros(3)
            nv[int(normal)] = 1.0
            normal = nv
        if normal is not None:
            if redshift == 0.0 and dist is None:
                raise RuntimeError("Cannot use a normal vector for the line-of-sight when a redshift "
                                   "or the distance is not specified!")
            self._spec_normal = np.asarray(normal, "float64")
        shifting = normal is not None
        emin, emax = parse_value(emin, "keV"), parse_value(emax, "keV")
        self.setup_model("spectrum", data_source, redshift)
        if isinstance(data_source, ArrayDataSource):
            self.v_fields = data_source.velocity_fields
        ebins = np.linspace(emin, emax, nbins + 1)
        spec = np.zeros(nbins)
        for chunk in _chunks(data_source):
            out = self.process_data("spectrum", chunk, 1.0, ebins=ebins, shifting=shifting)
            if out is not None:
                spec += out
        spec = self._finish_spectrum(ebins, spec)
        self.cleanup_model("spectrum")
        if redshift > 0.0 or dist is not None:
            if dist is None:
                cosmo = cosmology or getattr(data_source, "cosmology", None) or Cosmology()
                d_l = cosmo.angular_diameter_distance_mpc(0.0, redshift) * (1.0 + redshift) ** 2 * K.cm_per_mpc
            else:
                redshift = 0.0
                d_l = parse_value(dist, "kpc") * K.cm_per_kpc
            spec = spec * (1.0 + redshift) ** 2 / (4.0 * np.pi * d_l * d_l)   # sources.py:132-135
        return ebins, spec

    def _finish_spectrum(self, ebins, spec):
        return spec


def _inverse_volume(chunk, ftype):
    """1 / cell volume [cm^-3] (sources.py:141-164): ``cell_volume`` if present, else dx^3 for
    cells, else density / mass."""
    for f in (("index", "cell_volume"), (ftype, "cell_volume")):
        if f in chunk:
            return 1.0 / to_device(field_value(chunk, f, "cm**3"))
    if ("index", "dx") in chunk:
        dx = to_device(field_value(chunk, ("index", "dx"), "kpc")) * K.cm_per_kpc
        return 1.0 / (dx * dx * dx)
    if (ftype, "density") in chunk and (ftype, "mass") in chunk:
        return to_device(field_value(chunk, (ftype, "density"), "g/cm**3")) / \
            to_device(field_value(chunk, (ftype, "mass"), "g"))
    raise RuntimeError("No way to compute inverse volume")


# =======================================================================================
# thermal
# =======================================================================================
def _parse_abund_table(abund_table):
    if isinstance(abund_table, str):
        return K.abund_tables[abund_table]
    at = np.concatenate([[0.0], np.asarray(abund_table, dtype="float64")])
    if at.size != 31:
        raise RuntimeError("abund_table must have 30 entries")
    return at


class ThermalSourceModel(SourceModel):
    """pyxsim/source_models/thermal_sources/base.py:23-229 (constructor, setup) and :304-524
    (photons branch of process_data, here ``generate_chunk``)."""

    _density_dependence = False
    _nei = False

    def __init__(self, spectral_model, emin, emax, nbins, Zmet, binscale="linear", kT_min=0.025,
                 kT_max=64.0, var_elem=None, max_density=None, min_entropy=None,
                 method="invert_cdf", abund_table="angr", prng=None, temperature_field=None,
                 emission_measure_field=None, h_fraction=None, nH_min=None, nH_max=None):
        super().__init__(prng=prng)
        self.spectral_model = spectral_model
        self.emin = parse_value(emin, "keV")
        self.emax = parse_value(emax, "keV")
        self.Zmet = Zmet
        if var_elem is None:
            var_elem, var_elem_keys, self.num_var_elem = {}, None, 0
        else:
            var_elem_keys = list(var_elem.keys())
            self.num_var_elem = len(var_elem_keys)
        if self.num_var_elem > _lib.PXB_MAX_VAR:
            raise ValueError(f"at most {_lib.PXB_MAX_VAR} variable elements are supported")
        self.var_elem = dict(var_elem)
        self.var_elem_keys = var_elem_keys
        self.var_ion_keys = var_elem_keys
        self.max_density = None if max_density is None else parse_value(max_density, "g/cm**3")
        self.min_entropy = None if min_entropy is None else parse_value(min_entropy, "keV*cm**2")
        self.temperature_field = temperature_field or ("gas", "temperature")
        self.emission_measure_field = emission_measure_field or ("gas", "emission_measure")
        self.density_field = None
        self.entropy_field = None
        self.nh_field = None
        self.binscale = binscale
        self.abund_table = abund_table
        self.method = method
        if method not in ("invert_cdf", "accept_reject"):
            raise ValueError(f"unknown method {method!r}")
        self.kT_min = kT_min
        self.kT_max = kT_max
        self.nH_min, self.nH_max = nH_min, nH_max
        self.Zconvert = 1.0
        self.mconvert = {}
        self.atable = _parse_abund_table(abund_table)
        if h_fraction is None:
            h_fraction = K.compute_H_abund(self.atable)
        self.h_fraction = h_fraction
        self.ebins = self.spectral_model.ebins
        self.de = self.spectral_model.de
        self.emid = self.spectral_model.emid
        self.bin_edges = np.log10(self.ebins) if self.binscale == "log" else self.ebins
        self.nbins = self.emid.size
        self.model_vers = self.spectral_model.model_vers
        self._z_units = "Zsun"
        self._e_units = {}

    def __repr__(self):
        keys = ["emin", "emax", "nbins", "Zmet", "binscale", "temperature_field",
                "emission_measure_field", "kT_min", "kT_max", "method", "model_vers",
                "max_density", "min_entropy", "abund_table", "h_fraction", "var_elem"]
        body = "".join(f"    {k}={getattr(self, k)}\n" for k in keys)
        return f"{self.__class__.__name__}(\n{body})\n"

    # -- setup (base.py:146-229) ----------------------------------------------------------
    def _units_of(self, data_source, field):
        if isinstance(data_source, ArrayDataSource):
            return data_source.units.get(field, "Zsun")
        ds = getattr(data_source, "ds", data_source)
        return str(ds._get_field_info(field).units)

    def setup_model(self, mode, data_source, redshift):
        super().setup_model(mode, data_source, redshift)
        if isinstance(data_source, ArrayDataSource):
            self.ftype = data_source.ftype
        else:  # yt
            ds = getattr(data_source, "ds", data_source)
            self.emission_measure_field = ds._get_field_info(self.emission_measure_field).name
            self.temperature_field = ds._get_field_info(self.temperature_field).name
            self.ftype = self.emission_measure_field[0]
        metals = slice(3, 31)
        if not self._nei and not isinstance(self.Zmet, Number):
            zu = self._units_of(data_source, self.Zmet)
            self._z_units = zu
            if zu in ("dimensionless", "", "code_metallicity"):
                self.Zconvert = K.atomic_weights[1] / (self.atable * K.atomic_weights)[metals].sum()
            elif zu == "Zsun":
                self.Zconvert = 1.0
            else:
                raise RuntimeError(f"I don't understand metallicity units of {zu}!")
        for key, value in self.var_elem.items():
            if not isinstance(value, Number):
                elem = key.split("^")[0]
                n_elem = K.elem_names.index(elem)
                mu = self._units_of(data_source, value)
                self._e_units[key] = mu
                if mu in ("dimensionless", "", "code_metallicity"):
                    self.mconvert[key] = K.atomic_weights[1] / (self.atable[n_elem] * K.atomic_weights[n_elem])
                elif mu == "Zsun":
                    self.mconvert[key] = 1.0
                else:
                    raise RuntimeError(f"I don't understand units of {mu} for element {key}!")
        self.density_field = (self.ftype, "density")
        self.entropy_field = (self.ftype, "entropy")
        mylog.info("Using emission measure field '%s'.", self.emission_measure_field)
        mylog.info("Using temperature field '%s'.", self.temperature_field)
        self.spectral_model.prepare_spectrum(redshift)
        nv_table = len(getattr(self.spectral_model, "var_elem_names", []) or [])
        if self._nei:
            nv_table = len(getattr(self.spectral_model, "var_ion_names", []) or [])
        if self.spectral_model.var_spec is not None:
            nv_table = self.spectral_model.var_spec.shape[0]
        if nv_table != self.num_var_elem:
            raise RuntimeError(
                f"var_elem has {self.num_var_elem} entries but the spectral table has {nv_table}")

    # -- the hot function -----------------------------------------------------------------
    def _cells_struct(self, chunk, spectral_norm, keep):
        """Fill a ``pxb_thermal_cells`` from the chunk (base.py:335-425)."""
        Cs = _lib.ThermalCells()
        kT = to_device(field_value(chunk, self.temperature_field, "keV", "thermal"))
        em = to_device(field_value(chunk, self.emission_measure_field, "cm**-3"))
        keep += [kT, em]
        Cs.ncell = kT.numel()
        self._cell_ids(Cs, chunk, Cs.ncell)
        Cs.kT, Cs.em = kT.data_ptr(), em.data_ptr()
        if self.nh_field is not None:
            nH = to_device(field_value(chunk, self.nh_field, "cm**-3"))
            keep.append(nH)
            Cs.nH = nH.data_ptr()
        if isinstance(self.h_fraction, Number):
            Cs.xh_const = float(self.h_fraction)
        else:
            xh = to_device(field_value(chunk, self.h_fraction))
            keep.append(xh)
            Cs.xh = xh.data_ptr()
        if self._nei:
            Cs.zmet_const = 0.0
        elif isinstance(self.Zmet, Number):
            Cs.zmet_const = float(self.Zmet)
        else:
            z = to_device(field_value(chunk, self.Zmet))
            keep.append(z)
            Cs.zmet = z.data_ptr()
            Cs.zmet_scale = float(self.Zconvert)
            Cs.zmet_by_xh = int(self._z_units != "Zsun")
        Cs.nvar = self.num_var_elem
        for j, key in enumerate(self.var_elem_keys or []):
            value = self.var_elem[key]
            if isinstance(value, Number):
                Cs.elem_const[j] = float(value)
            else:
                e = to_device(field_value(chunk, value))
                keep.append(e)
                Cs.elem[j] = e.data_ptr()
                Cs.elem_scale[j] = float(self.mconvert[key])
                Cs.elem_by_xh[j] = int(self._e_units.get(key, "Zsun") != "Zsun")
        if self.max_density is not None:
            d = to_device(field_value(chunk, self.density_field, "g/cm**3"))
            keep.append(d)
            Cs.density = d.data_ptr()
            Cs.max_density = float(self.max_density)
        if self.min_entropy is not None:
            s = to_device(field_value(chunk, self.entropy_field, "keV*cm**2"))
            keep.append(s)
            Cs.entropy = s.data_ptr()
            Cs.min_entropy = float(self.min_entropy)
        if self.observer == "internal":
            r2 = self.compute_radius(chunk)
            keep.append(r2)
            Cs.r2 = r2.data_ptr()
        Cs.spectral_norm = float(spectral_norm)
        Cs.kT_min, Cs.kT_max = float(self.kT_min), float(self.kT_max)
        Cs.seed = self._seed
        # diagnostic switches (read once per chunk on the host, never inside the library)
        # (bit 0: option-generic kernels; bit 1: fused mode with one big CTA per SM and staged rows)
        Cs.generic_kernels = int(bool(os.environ.get("PXB_GENERIC_KERNELS"))) | \
            (2 if os.environ.get("PXB_FUSED_BIG_CTA") else 0)
        if os.environ.get("PXB_NO_GENERAL_SPLIT"):
            Cs.split_factor = -1
        elif os.environ.get("PXB_SPLIT_FACTOR"):
            Cs.split_factor = max(int(os.environ["PXB_SPLIT_FACTOR"]), 0)
        return Cs

    def expected_counts(self, chunk, spectral_norm):
        """(lambda, counts) per cell as device tensors -- the deterministic half of the path."""
        require_cuda()
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        dm = self.spectral_model.device_model(self.method)
        lam = empty(Cs.ncell)
        counts = empty(Cs.ncell, torch.int64)
        _lib.call("pxb_thermal_counts", dm.handle, C.byref(Cs), ptr(lam), ptr(counts), stream_ptr())
        return lam, counts

    def cell_spectra(self, chunk, spectral_norm=1.0):
        """Clipped per-cell total spectra [ncell, nbins] (diagnostic; base.py:453-460)."""
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        dm = self.spectral_model.device_model(self.method)
        spec = empty(Cs.ncell * self.nbins)
        _lib.call("pxb_thermal_spectra", dm.handle, C.byref(Cs), ptr(spec), stream_ptr())
        return spec.reshape(Cs.ncell, self.nbins)

    def _spectrum_chunk(self, chunk, spectral_norm, ebins, shifting):
        """base.py:453-460,494-495 + shift_spectrum (spectra.pyx:67-93) in one kernel."""
        require_cuda()
        if self._chunk_size(chunk, self.temperature_field) == 0:
            return None
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        dm = self.spectral_model.device_model(self.method)
        shift = self.compute_shift(chunk) if shifting else None
        eb = to_device(ebins)
        spec = torch.zeros(ebins.size - 1, dtype=torch.float64, device=eb.device)
        need = C.c_int64()
        _lib.call("pxb_thermal_spectrum_workspace_bytes", dm.handle, Cs.ncell, C.byref(need))
        # (diagnostic switch: without a workspace the library takes the per-cell path everywhere)
        ws = None if os.environ.get("PXB_SPECTRUM_PER_CELL") else \
            torch.empty(need.value, dtype=torch.uint8, device=eb.device)
        with stage("thermal_spectrum"):
            _lib.call("pxb_thermal_spectrum", dm.handle, C.byref(Cs), ptr(shift), ebins.size - 1,
                      ptr(eb), ptr(spec), ptr(ws), need.value if ws is not None else 0, stream_ptr())
        del ws
        return spec.cpu().numpy()

    def _finish_spectrum(self, ebins, spec):
        return spec / np.diff(ebins)      # base.py:297

    def make_fluxf(self, emin, emax, energy=False):
        """base.py:301-302"""
        return self.spectral_model.make_fluxf(emin, emax, energy=energy)

    def _field_chunk(self, mode, chunk, spectral_norm, emin, emax, fluxf, shifting):
        """base.py:497-511: shifted intensities integrate the cell's Doppler-shifted spectrum
        (make_band); everything else interpolates the tabulated band fluxes (fluxf)."""
        require_cuda()
        if self._chunk_size(chunk, self.temperature_field) == 0:
            return empty(0)
        if mode.endswith("intensity") and shifting:
            return self.band_per_cell(chunk, spectral_norm, emin, emax, energy=(mode == "intensity"),
                                      shifting=True)
        if fluxf is None:
            fluxf = self.make_fluxf(emin, emax, energy=mode in ("luminosity", "intensity"))
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        return fluxf.evaluate(Cs, keep)

    def band_per_cell(self, chunk, spectral_norm, emin, emax, energy=False, shifting=False):
        """make_band (spectra.pyx:96-126) times cell_nrm (base.py:497-500): per-cell band photon
        rate (or energy rate) of the clipped spectrum, device tensor."""
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        dm = self.spectral_model.device_model(self.method)
        shift = self.compute_shift(chunk) if shifting else None
        out = empty(Cs.ncell)
        need = C.c_int64()
        _lib.call("pxb_thermal_band_workspace_bytes", Cs.ncell, C.byref(need))
        ws = None if os.environ.get("PXB_BAND_PER_CELL") else \
            torch.empty(need.value, dtype=torch.uint8, device=out.device)
        _lib.call("pxb_thermal_band", dm.handle, C.byref(Cs), int(bool(energy)), C.c_double(emin),
                  C.c_double(emax), ptr(shift), ptr(out), ptr(ws), need.value if ws is not None else 0,
                  stream_ptr())
        del ws
        return out

    def generate_chunk(self, chunk, spectral_norm, gather=None):
        require_cuda()
        if self._chunk_size(chunk, self.temperature_field) == 0:
            return None
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        dm = self.spectral_model.device_model(self.method)
        counts = empty(Cs.ncell, torch.int64)
        with stage("thermal_counts"):
            _lib.call("pxb_thermal_counts", dm.handle, C.byref(Cs), None, ptr(counts), stream_ptr())
        out = self._scan_and_compact(counts, chunk, gather)
        if out.n_photons == 0:
            return None
        out.energy = empty(out.n_photons)
        need = C.c_int64()
        _lib.call("pxb_thermal_sample_workspace_bytes", dm.handle, out.n_cells, out.n_pairs, C.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=out.energy.device)
        with stage("thermal_sample"):
            _lib.call("pxb_thermal_sample", dm.handle, C.byref(Cs), out.n_cells, ptr(out.idx),
                      ptr(out.nph), ptr(out.poff), ptr(out.qoff), out.n_photons, out.n_pairs,
                      ptr(out.energy), ptr(ws), need.value, stream_ptr())
        return out

    def count_chunk(self, chunk, spectral_norm, gather=None):
        """First half of ``generate_chunk``: expected counts, Poisson draws, offsets and the
        gather of the active cells -- everything except the energies.  Returns ``(out, Cs, keep)``
        for the fused ``make_events`` path (``keep`` holds the tensors ``Cs`` points into)."""
        require_cuda()
        if self._chunk_size(chunk, self.temperature_field) == 0:
            return None, None, None
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        dm = self.spectral_model.device_model(self.method)
        counts = empty(Cs.ncell, torch.int64)
        with stage("thermal_counts"):
            _lib.call("pxb_thermal_counts", dm.handle, C.byref(Cs), None, ptr(counts), stream_ptr())
        out = self._scan_and_compact(counts, chunk, gather)
        return out, Cs, keep


class CIESourceModel(ThermalSourceModel):
    """pyxsim/source_models/thermal_sources/collisional.py:7-218.  ``model`` may be a model
    name understood by soxs ("apec", "spex", ...; needs soxs) or a ready
    :class:`~pyxsim_b200.spectral_models.ThermalSpectralModel` carrying the tables."""

    def __init__(self, model, emin, emax, nbins, Zmet, binscale="linear",
                 temperature_field=("gas", "temperature"),
                 emission_measure_field=("gas", "emission_measure"), h_fraction=None,
                 kT_min=0.025, kT_max=64.0, max_density=None, min_entropy=None, var_elem=None,
                 method="invert_cdf", thermal_broad=True, model_root=None, model_vers=None,
                 nolines=False, abund_table="angr", trace_abund=1.0, prng=None):
        from .spectral_models import TableCIEModel, ThermalSpectralModel

        var_elem_keys = list(var_elem.keys()) if var_elem else None
        if isinstance(model, ThermalSpectralModel):
            spectral_model = model
        else:
            spectral_model = TableCIEModel(model, emin, emax, nbins, kT_min, kT_max,
                                           binscale=binscale, var_elem=var_elem_keys,
                                           thermal_broad=thermal_broad, model_root=model_root,
                                           model_vers=model_vers, nolines=nolines,
                                           abund_table=abund_table, nei=self._nei)
        self.trace_abund = trace_abund
        super().__init__(spectral_model, emin, emax, nbins, Zmet, binscale=binscale,
                         kT_min=kT_min, kT_max=kT_max, var_elem=var_elem, max_density=max_density,
                         min_entropy=min_entropy, method=method, abund_table=abund_table,
                         prng=prng, temperature_field=temperature_field,
                         emission_measure_field=emission_measure_field, h_fraction=h_fraction)


class NEISourceModel(CIESourceModel):
    """collisional.py:220-378: ion fractions enter through ``var_elem`` (keys like "O^7"),
    the metal table is switched off (``metalZ = 0``, base.py:395-397)."""

    _nei = True

    def __init__(self, model, emin, emax, nbins, var_elem, binscale="linear",
                 temperature_field=("gas", "temperature"),
                 emission_measure_field=("gas", "emission_measure"), h_fraction=None,
                 kT_min=0.025, kT_max=64.0, max_density=None, min_entropy=None,
                 method="invert_cdf", thermal_broad=True, model_root=None, model_vers=None,
                 nolines=False, abund_table="angr", prng=None):
        super().__init__(model, emin, emax, nbins, 0.0, binscale=binscale,
                         temperature_field=temperature_field,
                         emission_measure_field=emission_measure_field, h_fraction=h_fraction,
                         kT_min=kT_min, kT_max=kT_max, max_density=max_density,
                         min_entropy=min_entropy, var_elem=var_elem, method=method,
                         thermal_broad=thermal_broad, model_root=model_root,
                         model_vers=model_vers, nolines=nolines, abund_table=abund_table,
                         prng=prng)


class PionSourceModel(ThermalSourceModel):
    """pyxsim/source_models/thermal_sources/photoionization.py:6-171.  Needs a
    :class:`~pyxsim_b200.spectral_models.PionSpectralModel` (pass it as ``spectral_model``;
    building one from soxs' Cloudy tables requires soxs)."""

    _density_dependence = True

    def __init__(self, emin, emax, nbins, Zmet, binscale="linear", resonant_scattering=False,
                 cxb_factor=0.5, nh_field=("gas", "H_nuclei_density"),
                 temperature_field=("gas", "temperature"),
                 emission_measure_field=("gas", "emission_measure"), h_fraction=None,
                 kT_min=0.00431, kT_max=64.0, max_density=None, min_entropy=None, var_elem=None,
                 method="invert_cdf", model_vers="4_lo", prng=None, spectral_model=None):
        if spectral_model is None:
            raise ImportError(
                "PionSourceModel needs the Cloudy photoionization tables that soxs downloads; "
                "pass spectral_model=PionSpectralModel(...) built from table arrays")
        # the reference fixes the Feldman (1992) table here; it ships with soxs
        abund = "feld" if "feld" in K.abund_tables else "angr"
        super().__init__(spectral_model, emin, emax, nbins, Zmet, binscale=binscale,
                         kT_min=kT_min, kT_max=kT_max, var_elem=var_elem, max_density=max_density,
                         min_entropy=min_entropy, method=method, abund_table=abund, prng=prng,
                         temperature_field=temperature_field,
                         emission_measure_field=emission_measure_field, h_fraction=h_fraction)
        self.nh_field = nh_field
        self.resonant_scattering = resonant_scattering
        self.cxb_factor = cxb_factor


IGMSourceModel = PionSourceModel


# =======================================================================================
# power law / line
# =======================================================================================
class PowerLawSourceModel(SourceModel):
    """pyxsim/source_models/power_law_sources.py:13-250."""

    def __init__(self, e0, emin, emax, luminosity_field, alpha, prng=None):
        super().__init__(prng=prng)
        self.e0 = parse_value(e0, "keV")
        self.emin = parse_value(emin, "keV")
        self.emax = parse_value(emax, "keV")
        self.luminosity_field = luminosity_field
        self.alpha = alpha
        self.scale_factor = 1.0
        if self.e0 != 1.0:
            mylog.warning("PowerLawSourceModel: the reference's inverse CDF is only exact for "
                          "e0 = 1 keV (power_law_sources.py:214); reproducing it literally.")

    def __repr__(self):
        return (f"PowerLawSourceModel(\n    e0={self.e0}\n    emin={self.emin}\n    emax={self.emax}\n"
                f"    luminosity_field={self.luminosity_field}\n    alpha={self.alpha}\n)\n")

    def setup_model(self, mode, data_source, redshift):
        super().setup_model(mode, data_source, redshift)
        self.scale_factor = 1.0 / (1.0 + redshift)
        self.ftype = self.luminosity_field[0] if isinstance(self.luminosity_field, tuple) else "gas"

    def _cells_struct(self, chunk, spectral_norm, keep):
        Cs = _lib.PowerlawCells()
        lum = to_device(field_value(chunk, self.luminosity_field, "keV/s"))
        keep.append(lum)
        Cs.ncell = lum.numel()
        self._cell_ids(Cs, chunk, Cs.ncell)
        Cs.lum = lum.data_ptr()
        if isinstance(self.alpha, Real):   # any real number (the reference mixes Number/float)
            Cs.alpha_const = float(self.alpha)
        else:
            a = to_device(field_value(chunk, self.alpha))
            keep.append(a)
            Cs.alpha = a.data_ptr()
        if self.observer == "internal":
            r2 = self.compute_radius(chunk)
            keep.append(r2)
            Cs.r2 = r2.data_ptr()
        Cs.e0, Cs.emin, Cs.emax = self.e0, self.emin, self.emax
        Cs.spectral_norm = float(spectral_norm)
        Cs.scale_factor = float(self.scale_factor)
        Cs.seed = self._seed
        return Cs

    def expected_counts(self, chunk, spectral_norm):
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        lam, counts = empty(Cs.ncell), empty(Cs.ncell, torch.int64)
        _lib.call("pxb_powerlaw_counts", C.byref(Cs), ptr(lam), ptr(counts), stream_ptr())
        return lam, counts

    def generate_chunk(self, chunk, spectral_norm, gather=None):
        require_cuda()
        if self._chunk_size(chunk, self.luminosity_field) == 0:
            return None
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        counts = empty(Cs.ncell, torch.int64)
        with stage("powerlaw_counts"):
            _lib.call("pxb_powerlaw_counts", C.byref(Cs), None, ptr(counts), stream_ptr())
        out = self._scan_and_compact(counts, chunk, gather)
        if out.n_photons == 0:
            return None
        out.energy = empty(out.n_photons)
        with stage("powerlaw_sample"):
            _lib.call("pxb_powerlaw_sample", C.byref(Cs), out.n_cells, ptr(out.idx), ptr(out.nph),
                      ptr(out.poff), out.n_photons, ptr(out.energy), stream_ptr())
        return out

    def _spectrum_chunk(self, chunk, spectral_norm, ebins, shifting):
        """power_law_sources.py:264-270 + power_law_spectrum (spectra.pyx:11-32)."""
        require_cuda()
        if self._chunk_size(chunk, self.luminosity_field) == 0:
            return None
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        Kc, alpha = empty(Cs.ncell), empty(Cs.ncell)
        _lib.call("pxb_powerlaw_norm", C.byref(Cs), ptr(Kc), ptr(alpha), stream_ptr())
        shift = self.compute_shift(chunk) if shifting else None
        emid = 0.5 * (ebins[1:] + ebins[:-1]) * (1.0 / self.scale_factor) / self.e0
        em = to_device(emid)
        spec = torch.zeros(emid.size, dtype=torch.float64, device=em.device)
        _lib.call("pxb_powerlaw_spectrum", Cs.ncell, ptr(alpha), ptr(Kc), ptr(shift), emid.size, ptr(em),
                  ptr(spec), stream_ptr())
        return spec.cpu().numpy()

    def process_data(self, mode, chunk, spectral_norm, **kwargs):
        res = super().process_data(mode, chunk, spectral_norm, **kwargs)
        if res is None or mode != "photons":
            return res
        # the reference returns a boolean mask here (power_law_sources.py:242-250)
        ncells, nph, idx, energies = res
        mask = np.zeros(self._chunk_size(chunk, self.luminosity_field), dtype=bool)
        mask[idx] = True
        return ncells, nph, mask, energies


class LineSourceModel(SourceModel):
    """pyxsim/source_models/line_sources.py:19-234.  ``sigma``: a number / (value, unit)
    in keV or km/s, or a field of velocity dispersions in km/s."""

    def __init__(self, e0, emission_field, sigma, prng=None):
        super().__init__(prng=prng)
        self.e0 = parse_value(e0, "keV")
        self.emission_field = emission_field
        self.sigma_field = None
        is_vu = (isinstance(sigma, (tuple, list)) and len(sigma) == 2
                 and isinstance(sigma[0], Real) and isinstance(sigma[1], str))
        if is_vu and sigma[1] in ("km/s", "cm/s"):
            v = float(sigma[0]) * (1.0 if sigma[1] == "km/s" else 1e-5)
            self.sigma = v * self.e0 / K.clight_kms       # line_sources.py:60
        elif isinstance(sigma, Real) or is_vu:
            self.sigma = parse_value(sigma, "keV")
        else:
            self.sigma = None
            self.sigma_field = sigma
        self.scale_factor = 1.0

    def __repr__(self):
        return (f"LineSourceModel(\n    e0={self.e0}\n    emission_field={self.emission_field}\n"
                f"    sigma={self.sigma if self.sigma_field is None else self.sigma_field}\n)\n")

    def setup_model(self, mode, data_source, redshift):
        super().setup_model(mode, data_source, redshift)
        self.scale_factor = 1.0 / (1.0 + redshift)
        self.ftype = self.emission_field[0] if isinstance(self.emission_field, tuple) else "gas"

    def _cells_struct(self, chunk, spectral_norm, keep):
        Cs = _lib.LineCells()
        rate = to_device(field_value(chunk, self.emission_field, "photons/s"))
        keep.append(rate)
        Cs.ncell = rate.numel()
        self._cell_ids(Cs, chunk, Cs.ncell)
        Cs.rate = rate.data_ptr()
        if self.sigma_field is None:
            Cs.sigma_const = float(self.sigma)
        else:
            # (v * e0 / c) in keV, line_sources.py:202
            s = to_device(field_value(chunk, self.sigma_field, "km/s")) * (self.e0 / K.clight_kms)
            keep.append(s)
            Cs.sigma = s.data_ptr()
        if self.observer == "internal":
            r2 = self.compute_radius(chunk)
            keep.append(r2)
            Cs.r2 = r2.data_ptr()
        Cs.e0 = self.e0
        Cs.spectral_norm = float(spectral_norm)
        Cs.scale_factor = float(self.scale_factor)
        Cs.seed = self._seed
        return Cs

    def expected_counts(self, chunk, spectral_norm):
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        lam, counts = empty(Cs.ncell), empty(Cs.ncell, torch.int64)
        _lib.call("pxb_line_counts", C.byref(Cs), ptr(lam), ptr(counts), stream_ptr())
        return lam, counts

    def generate_chunk(self, chunk, spectral_norm, gather=None):
        require_cuda()
        if self._chunk_size(chunk, self.emission_field) == 0:
            return None
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        counts = empty(Cs.ncell, torch.int64)
        with stage("line_counts"):
            _lib.call("pxb_line_counts", C.byref(Cs), None, ptr(counts), stream_ptr())
        out = self._scan_and_compact(counts, chunk, gather)
        if out.n_photons == 0:
            return None
        out.energy = empty(out.n_photons)
        with stage("line_sample"):
            _lib.call("pxb_line_sample", C.byref(Cs), out.n_cells, ptr(out.idx), ptr(out.nph),
                      ptr(out.poff), out.n_photons, ptr(out.energy), stream_ptr())
        return out

    _gauss = None

    def _spectrum_chunk(self, chunk, spectral_norm, ebins, shifting):
        """line_sources.py:236-252 + line_spectrum (spectra.pyx:35-64); the unit Gaussian is the
        reference's module-level table gx = linspace(-7, 7, 10000), gpdf = norm.pdf(gx)."""
        require_cuda()
        if self._chunk_size(chunk, self.emission_field) == 0:
            return None
        from scipy.stats import norm

        if LineSourceModel._gauss is None:
            gx = np.linspace(-7, 7, 10000)
            LineSourceModel._gauss = (gx, norm.pdf(gx))
        gx, gpdf = LineSourceModel._gauss
        keep = []
        Cs = self._cells_struct(chunk, spectral_norm, keep)
        n = Cs.ncell
        rate = keep[0]
        if self.sigma_field is None:
            sigma = torch.full((n,), float(self.sigma), dtype=torch.float64, device=rate.device)
        else:
            sigma = keep[1]
        shift = self.compute_shift(chunk) if shifting else None
        ee = 0.5 * (ebins[1:] + ebins[:-1]) / self.scale_factor
        dee, dgx, dgp = to_device(ee), to_device(gx), to_device(gpdf)
        spec = torch.zeros(ee.size, dtype=torch.float64, device=rate.device)
        _lib.call("pxb_line_spectrum", n, C.c_double(self.e0), ptr(sigma), ptr(rate), ptr(shift), ee.size,
                  ptr(dee), gx.size, ptr(dgx), ptr(dgp), ptr(spec), stream_ptr())
        return spec.cpu().numpy()

    def process_data(self, mode, chunk, spectral_norm, **kwargs):
        res = super().process_data(mode, chunk, spectral_norm, **kwargs)
        if res is None or mode != "photons":
            return res
        ncells, nph, idx, energies = res
        mask = np.zeros(self._chunk_size(chunk, self.emission_field), dtype=bool)
        mask[idx] = True
        return ncells, nph, mask, energies
