"""Internal absorption columns (row f3 of the scope table): the consumer side of
``pyxsim/internal_absorption.py`` + ``photon_list.py:524-546,672-680,715-723``.

The reference builds a cube of neutral-hydrogen column density with yt projections
(``make_column_density_map``, internal_absorption.py:11-126 -- stays on yt, it is a yt
operation) and ``project_photons(column_file=...)`` interpolates it trilinearly at every
cell's rotated position, then absorbs that cell's photons at their source-frame energy.  Here:

* :func:`read_column_file` / :func:`write_column_file` -- the file layout of
  internal_absorption.py:119-126 (``/parameters`` attrs ``normal``, ``north``; ``/data``
  ``wbins``, ``dbins``, ``nH`` in 1e22 cm^-2), HDF5 when h5py is present, else ``.npz`` with
  the same keys;
* :func:`column_density_at_cells` -- the interpolation on the GPU (``pxb_column_lookup``),
  same node search, weights and summation order as ``scipy.interpolate.interpn`` as the
  reference calls it (agreement 1e-13: the rotated position is a BLAS product there);
* the per-cell absorption itself is fused into the projection kernel (``nH_cell``).
"""
import os

import numpy as np

from . import _lib
from .device import empty, ptr, stream_ptr, to_device

try:  # optional, like the reference's own I/O dependency
    import h5py
except ImportError:  # pragma: no cover
    h5py = None


def write_column_file(outfile, normal, north_vector, wbins, dbins, nH):
    """Write a column-density cube (``nH[nw, nw, nd]`` in 1e22 cm^-2) in the layout
    ``make_column_density_map`` produces (internal_absorption.py:119-126)."""
    wbins, dbins = np.asarray(wbins, "float64"), np.asarray(dbins, "float64")
    nH = np.asarray(nH, "float64")
    if nH.shape != (wbins.size - 1, wbins.size - 1, dbins.size - 1):
        raise ValueError(f"nH has shape {nH.shape}, expected {(wbins.size - 1,) * 2 + (dbins.size - 1,)}")
    normal, north_vector = np.asarray(normal, "float64"), np.asarray(north_vector, "float64")
    if h5py is not None and not outfile.endswith(".npz"):
        with h5py.File(outfile, "w") as f:
            p = f.create_group("parameters")
            p.attrs["normal"] = normal
            p.attrs["north"] = north_vector
            d = f.create_group("data")
            d.create_dataset("wbins", data=wbins)
            d.create_dataset("dbins", data=dbins)
            d.create_dataset("nH", data=nH)
        return outfile
    fn = outfile if outfile.endswith(".npz") else outfile + ".npz"
    np.savez(fn, **{"parameters/normal": normal, "parameters/north": north_vector,
                    "data/wbins": wbins, "data/dbins": dbins, "data/nH": nH})
    return fn


def read_column_file(column_file):
    """-> dict(normal, north, wbins, dbins, nH) (photon_list.py:530-546)."""
    if h5py is not None and os.path.exists(column_file) and not column_file.endswith(".npz"):
        with h5py.File(column_file, "r") as fa:
            return dict(normal=np.asarray(fa["parameters"].attrs["normal"], "float64"),
                        north=np.asarray(fa["parameters"].attrs["north"], "float64"),
                        wbins=fa["data"]["wbins"][()], dbins=fa["data"]["dbins"][()],
                        nH=fa["data"]["nH"][()])
    fn = column_file if column_file.endswith(".npz") else column_file + ".npz"
    if not os.path.exists(fn):
        raise FileNotFoundError(f"no column-density file at {column_file} (or {fn})")
    z = np.load(fn, allow_pickle=False)
    return dict(normal=z["parameters/normal"], north=z["parameters/north"], wbins=z["data/wbins"],
                dbins=z["data/dbins"], nH=z["data/nH"])


def column_density_at_cells(x, y, z, unit_vectors, wbins, dbins, nH_grid):
    """nH (1e22 cm^-2) in front of every cell: ``interpn((wmid, wmid, dmid), nH_grid,
    (unit_vectors @ [x, y, z]).T, bounds_error=False, fill_value=0.0)`` (photon_list.py:672-680).
    ``x, y, z``: device tensors or arrays (kpc, centred); returns a device tensor."""
    wbins, dbins = np.asarray(wbins, "float64"), np.asarray(dbins, "float64")
    wmid = 0.5 * (wbins[1:] + wbins[:-1])
    dmid = 0.5 * (dbins[1:] + dbins[:-1])
    grid = np.ascontiguousarray(nH_grid, "float64")
    if grid.shape != (wmid.size, wmid.size, dmid.size):
        raise ValueError(f"nH cube has shape {grid.shape}, expected {(wmid.size, wmid.size, dmid.size)}")
    dx, dy, dz = to_device(x), to_device(y), to_device(z)
    uv = to_device(np.ascontiguousarray(unit_vectors, "float64").reshape(9))
    dw, dd, dg = to_device(wmid), to_device(dmid), to_device(grid.reshape(-1))
    out = empty(dx.numel())
    _lib.call("pxb_column_lookup", dx.numel(), ptr(dx), ptr(dy), ptr(dz), ptr(uv), wmid.size,
              dmid.size, ptr(dw), ptr(dd), ptr(dg), ptr(out), stream_ptr())
    return out


def make_column_density_map(*args, **kwargs):
    """The cube itself is made of yt projections (internal_absorption.py:66-115); that part is
    a yt operation and is not reimplemented.  Use yt + the reference function to make the file,
    or :func:`write_column_file` for a cube you already have."""
    raise NotImplementedError(make_column_density_map.__doc__)
