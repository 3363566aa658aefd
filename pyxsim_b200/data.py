"""Minimal host front end for array-backed datasets.

In the reference the front end is yt (dataset I/O, chunk iteration, units; e.g.
pyxsim/photon_list.py:401,426-451).  yt is not part of the hot path and is not required
here: :class:`ArrayDataSource` exposes the handful of things ``make_photons`` and the source
models ask of a yt data object -- chunk iteration, ``chunk[field]`` access in known units,
domain geometry -- for plain NumPy / torch arrays (host or device).  When yt *is* installed,
yt data objects are accepted as well (see :func:`wrap_data_source`).

Unit conventions of an ArrayDataSource (override per field with ``units={field: unit}``):
  temperature  K            (or "keV")
  emission measure  cm**-3
  density  g/cm**3, entropy keV*cm**2, H nuclei density cm**-3
  metallicity / element fields  Zsun  (or "dimensionless" mass fraction)
  positions, widths  kpc;  velocities  km/s
  luminosity (power law)  keV/s  (or "erg/s");  line emission  photons/s
"""
import numpy as np

from . import constants as K


class Cosmology:
    """Flat/non-flat LCDM angular diameter distance; defaults follow yt's Cosmology()
    (h=0.71, Om=0.27, OL=0.73), which make_photons falls back to (photon_list.py:225-231)."""

    def __init__(self, hubble_constant=0.71, omega_matter=0.27, omega_lambda=0.73,
                 omega_curvature=0.0, omega_radiation=0.0):
        self.hubble_constant = hubble_constant
        self.omega_matter = omega_matter
        self.omega_lambda = omega_lambda
        self.omega_curvature = omega_curvature
        self.omega_radiation = omega_radiation

    def _e(self, z):
        zp1 = 1.0 + z
        return np.sqrt(self.omega_matter * zp1**3 + self.omega_curvature * zp1**2
                       + self.omega_radiation * zp1**4 + self.omega_lambda)

    def comoving_radial_distance_mpc(self, z1, z2):
        from scipy.integrate import quad

        c_over_h0 = K.clight_kms / (100.0 * self.hubble_constant)  # Mpc
        val, _ = quad(lambda z: 1.0 / self._e(z), z1, z2, epsabs=0.0, epsrel=1e-12)
        return c_over_h0 * val

    def angular_diameter_distance_mpc(self, z1, z2):
        dc = self.comoving_radial_distance_mpc(z1, z2)
        c_over_h0 = K.clight_kms / (100.0 * self.hubble_constant)
        ok = self.omega_curvature
        if ok > 0:
            dm = c_over_h0 / np.sqrt(ok) * np.sinh(np.sqrt(ok) * dc / c_over_h0)
        elif ok < 0:
            dm = c_over_h0 / np.sqrt(-ok) * np.sin(np.sqrt(-ok) * dc / c_over_h0)
        else:
            dm = dc
        return dm / (1.0 + z2)


_UNIT_FACTORS = {
    ("K", "keV"): 1.0 / K.K_per_keV, ("keV", "K"): K.K_per_keV,
    ("erg/s", "keV/s"): 1.0 / K.erg_per_keV, ("keV/s", "erg/s"): K.erg_per_keV,
    ("cm", "kpc"): 1.0 / K.cm_per_kpc, ("kpc", "cm"): K.cm_per_kpc,
    ("Mpc", "kpc"): 1.0e3, ("kpc", "Mpc"): 1.0e-3,
    ("cm/s", "km/s"): 1.0e-5, ("km/s", "cm/s"): 1.0e5,
}

_DEFAULT_UNITS = {
    "temperature": "K", "kT": "keV", "emission_measure": "cm**-3", "density": "g/cm**3",
    "entropy": "keV*cm**2", "H_nuclei_density": "cm**-3", "metallicity": "Zsun",
}


def convert(values, unit_from, unit_to):
    if unit_from == unit_to or unit_to is None or unit_from is None:
        return values
    try:
        f = _UNIT_FACTORS[(unit_from, unit_to)]
    except KeyError:
        raise ValueError(f"pyxsim_b200 cannot convert '{unit_from}' to '{unit_to}'") from None
    return values * f


class ArrayChunk:
    """One contiguous slab of an ArrayDataSource (the analogue of a yt io chunk)."""

    def __init__(self, src, start, stop):
        self.src = src
        self.start = start
        self.stop = stop
        self.ds = src

    @property
    def size(self):
        return self.stop - self.start

    def __contains__(self, field):
        return field in self.src.fields

    def __getitem__(self, field):
        return self.src.fields[field][self.start:self.stop]

    def units(self, field):
        u = self.src.units.get(field)
        if u is None:
            name = field[1] if isinstance(field, tuple) else field
            u = _DEFAULT_UNITS.get(name)
        return u

    def to_value(self, field, unit=None):
        return convert(self[field], self.units(field), unit)

    def id_map(self):
        """(cell_id0, id_block, id_stride) of ``pxb_thermal_cells``: global id of local cell i."""
        return self.start + self.src.global_offset, 0, 0

    def global_ids(self, idx):
        """Global cell ids of the chunk-local indices ``idx`` (torch tensor)."""
        return idx + (self.start + self.src.global_offset)


class ShardedChunk(ArrayChunk):
    """One rank's block-cyclic share of an ArrayDataSource gathered into a single chunk: the rank
    owns blocks ``rank, rank + size, ...`` of ``block`` cells (``parallel.shard_bounds``), so the
    photon-rich core of a cluster is spread over all ranks.  Cells keep their global ids (RNG
    counters are keyed by them), which is what makes an N-rank run reproduce the 1-rank run."""

    def __init__(self, src, block, rank, size):
        super().__init__(src, 0, src.size)
        self.block, self.rank, self.nranks = int(block), int(rank), int(size)
        n, b, s, r = src.size, self.block, self.nranks, self.rank
        self._nfull = n // (b * s)                       # complete rounds of s blocks
        t0 = self._nfull * b * s + r * b                 # this rank's block of the last, partial round
        self._tail = (t0, min(t0 + b, n)) if t0 < n else None
        self._cache = {}

    @property
    def size(self):
        return self._nfull * self.block + (self._tail[1] - self._tail[0] if self._tail else 0)

    def __getitem__(self, field):
        if field in self._cache:
            return self._cache[field]
        v = self.src.fields[field]
        b, s, r, nf = self.block, self.nranks, self.rank, self._nfull
        head = v[: nf * b * s].reshape(nf, s, b)[:, r, :].reshape(-1)
        if self._tail is not None:
            tail = v[self._tail[0]:self._tail[1]]
            if not isinstance(v, np.ndarray) and hasattr(v, "device"):     # torch tensor
                import torch

                head = torch.cat([head, tail])
            else:
                head = np.concatenate([head, tail])
        self._cache[field] = head
        return head

    def id_map(self):
        return self.src.global_offset + self.rank * self.block, self.block, self.block * self.nranks

    def global_ids(self, idx):
        id0, blk, stride = self.id_map()
        return id0 + (idx // blk) * stride + idx % blk


class LocalShardChunk(ArrayChunk):
    """A slab of an ArrayDataSource that holds ONE RANK'S block-cyclic share of a larger dataset
    (``shard_of``): local cell i is global cell
    ``(rank + (i // block) * size) * block + i % block``."""

    def id_map(self):
        block, rank, size = self.src.shard_of
        return (self.src.global_offset + rank * block + (self.start // block) * block * size, block,
                block * size)

    def global_ids(self, idx):
        id0, blk, stride = self.id_map()
        return id0 + (idx // blk) * stride + idx % blk


class ArrayDataSource:
    """Array-backed data object (cells or particles).

    fields : dict mapping a field key (any hashable; yt-style tuples recommended) to a 1-D
        array (NumPy, or torch on host/device).  The geometry fields are looked up by
        ``position_fields`` / ``velocity_fields`` / ``width_field``.
    """

    def __init__(self, fields, *, data_type="cells", domain_left_edge=None,
                 domain_right_edge=None, periodicity=(False, False, False), left_edge=None,
                 right_edge=None, center=None, position_fields=None, velocity_fields=None,
                 width_field="default", units=None, chunk_size=None, cosmology=None,
                 name="ArrayDataSource", ftype="gas", rank_local=False, global_offset=0,
                 shard_block=256, shard_of=None):
        self.fields = dict(fields)
        self.units = dict(units or {})
        self.data_type = data_type
        self.name = name
        self.ftype = ftype
        any_field = next(iter(self.fields.values()))
        self.size = int(any_field.shape[0])
        for k, v in self.fields.items():
            if v.ndim != 1 or v.shape[0] != self.size:
                raise ValueError(f"field {k} must be 1-D with {self.size} entries")
        if data_type == "cells":
            pf = [("index", ax) for ax in "xyz"]
            vf = [(ftype, f"velocity_{ax}") for ax in "xyz"]
            wf = ("index", "dx")
        elif data_type == "particles":
            pf = [(ftype, f"particle_position_{ax}") for ax in "xyz"]
            vf = [(ftype, f"particle_velocity_{ax}") for ax in "xyz"]
            wf = (ftype, "smoothing_length")
        else:
            raise ValueError("data_type must be 'cells' or 'particles'")
        self.position_fields = list(position_fields or pf)
        self.velocity_fields = list(velocity_fields or vf)
        self.width_field = wf if width_field == "default" else width_field
        if self.width_field is not None and self.width_field not in self.fields:
            self.width_field = None
        self.periodicity = tuple(bool(p) for p in periodicity)
        self.domain_left_edge = None if domain_left_edge is None else np.asarray(domain_left_edge, "float64")
        self.domain_right_edge = None if domain_right_edge is None else np.asarray(domain_right_edge, "float64")
        if self.domain_left_edge is None:
            lo, hi = [], []
            for f in self.position_fields:
                v = self.fields[f]
                lo.append(float(v.min()))
                hi.append(float(v.max()))
            self.domain_left_edge = np.array(lo)
            self.domain_right_edge = np.array(hi)
        self.left_edge = self.domain_left_edge if left_edge is None else np.asarray(left_edge, "float64")
        self.right_edge = self.domain_right_edge if right_edge is None else np.asarray(right_edge, "float64")
        self.domain_width = self.domain_right_edge - self.domain_left_edge
        self.center = 0.5 * (self.left_edge + self.right_edge) if center is None else np.asarray(center, "float64")
        self.chunk_size = chunk_size
        self.cosmology = cosmology
        # rank_local: this object already holds only this rank's shard (chunks are not dealt
        # over ranks again); global_offset: global id of its first cell (RNG counters)
        self.rank_local = rank_local
        self.global_offset = int(global_offset)
        # cells per block of the block-cyclic split used when several ranks share this object
        self.shard_block = int(shard_block)
        # (block, rank, size): the arrays ARE rank `rank`'s block-cyclic share of a larger dataset
        # (every rank materialised only the cells it owns); implies rank_local
        self.shard_of = None if shard_of is None else tuple(int(v) for v in shard_of)
        if self.shard_of is not None:
            self.rank_local = True
            if self.chunk_size and self.chunk_size % self.shard_of[0]:
                raise ValueError("chunk_size of a sharded source must be a multiple of the shard block")

    def chunks(self, rank=0, size=1):
        """Yield this rank's chunks.  With ``size`` > 1 the cells are split block-cyclically over
        the ranks: with a ``chunk_size`` the chunks are dealt round-robin like yt's
        ``parallel_objects`` (photon_list.py:401), i.e. blocks of ``chunk_size`` cells; without one
        the rank's blocks of ``shard_block`` cells are gathered into ONE chunk so the kernels
        still see a large batch.  ``rank_local`` objects already hold only this rank's share."""
        if self.rank_local:
            rank, size = 0, 1
        if size > 1 and not self.chunk_size:
            ch = ShardedChunk(self, self.shard_block, rank, size)
            if ch.size > 0:
                yield ch
            return
        cs = self.chunk_size or self.size
        starts = list(range(0, self.size, cs)) or [0]
        kind = LocalShardChunk if self.shard_of is not None else ArrayChunk
        for j, s in enumerate(starts):
            if j % size == rank:
                yield kind(self, s, min(s + cs, self.size))

    def __repr__(self):
        return f"{self.name}({self.data_type}, n={self.size})"


def is_yt_object(obj):
    return hasattr(obj, "ds") and hasattr(obj, "chunks") and not isinstance(obj, ArrayDataSource)
