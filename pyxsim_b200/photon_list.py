"""``make_photons`` / ``project_photons`` / ``project_photons_allsky`` -- the drop-in API.

Host orchestration mirrors pyxsim/photon_list.py (make_photons :111-478, _project_photons
:481-830, project_photons :833-962, project_photons_allsky :964-1060); every per-cell and
per-photon operation runs in libpxb on the GPU.

Photon and event lists keep the reference's on-disk schema (groups ``info`` / ``parameters``
/ ``data`` and their dataset names, including the ``absoption_model`` key).  They are
written as HDF5 when h5py is importable and as ``.npz`` archives with the same keys
otherwise; a prefix that starts with ``"mem:"`` keeps the list resident in HBM, which is how
stage 1 hands over to stage 2 without leaving the device.
"""
import atexit
import ctypes as C
import os
import threading

import numpy as np
import torch

from . import _lib, constants as K, parallel
from .data import ArrayChunk, ArrayDataSource, Cosmology
from .device import empty, ptr, stream_ptr, to_device, require_cuda
from .internal_absorption import column_density_at_cells, read_column_file
from .profiling import stage
from .spectral_models import absorb_models
from .utils import get_normal_and_north, mylog, parse_value, prng_seed

__version__ = "0.1.0"

try:  # pragma: no cover - not installed in the build image
    import h5py
except Exception:
    h5py = None

_MEMORY = {}   # "mem:" lists: prefix -> DeviceList

_CELL_FIELDS = ["x", "y", "z", "vx", "vy", "vz", "num_photons", "dx"]


class DeviceList:
    """A photon or event list: ``info`` attrs, ``parameters`` scalars, ``data`` arrays
    (torch tensors on the device, or NumPy arrays after loading from disk)."""

    def __init__(self, info=None, parameters=None, data=None):
        self.info = dict(info or {})
        self.parameters = dict(parameters or {})
        self.data = dict(data or {})

    def numpy(self, key):
        v = self.data[key]
        return v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)

    def to_host(self):
        return DeviceList(self.info, self.parameters, {k: self.numpy(k) for k in self.data})


def _strip(prefix):
    return prefix[:-3] if prefix.endswith(".h5") else prefix


def _file_name(prefix):
    if parallel.size() > 1:
        return f"{prefix}.{parallel.rank():04d}"
    return prefix


def _file_names(prefix):
    if parallel.size() > 1:
        return [f"{prefix}.{i:04d}.h5" for i in range(parallel.size())]
    return [f"{prefix}.h5"]


STREAM_CHUNK_BYTES = 64 << 20     # slab of the chunked device -> host -> file copy


def _host_chunks(v, chunk_bytes=None, ready=None):
    """The flat array ``v`` as a sequence of contiguous host chunks (NumPy views).  A device
    tensor is copied slab by slab through two pinned buffers on a side stream: slab k+1 crosses
    PCIe while the caller writes slab k (the reference writes its lists tile by tile as well,
    photon_list.py:785-799) -- no pageable host copy of the whole list is ever made."""
    chunk_bytes = chunk_bytes or STREAM_CHUNK_BYTES
    if not isinstance(v, torch.Tensor) or not v.is_cuda:
        a = v.numpy() if isinstance(v, torch.Tensor) else np.ascontiguousarray(v)
        a = a.reshape(-1)
        per = max(1, chunk_bytes // max(a.itemsize, 1))
        for o in range(0, a.size, per):
            yield a[o:o + per]
        return
    t = v.reshape(-1)
    n = t.numel()
    per = max(1, chunk_bytes // t.element_size())
    nchunk = (n + per - 1) // per
    if nchunk == 0:
        return
    bufs = [torch.empty(min(per, n), dtype=t.dtype, pin_memory=True) for _ in range(min(2, nchunk))]
    evs = [torch.cuda.Event() for _ in bufs]
    side = torch.cuda.Stream(device=t.device)
    if ready is not None:        # (a background writer: the event its producer recorded)
        side.wait_event(ready)
    else:
        side.wait_stream(torch.cuda.current_stream(t.device))

    def issue(c):
        o = c * per
        m = min(per, n - o)
        with torch.cuda.stream(side):
            bufs[c % 2][:m].copy_(t[o:o + m], non_blocking=True)
            evs[c % 2].record(side)
        return m

    m_next = issue(0)
    for c in range(nchunk):
        m = m_next
        if c + 1 < nchunk:
            m_next = issue(c + 1)      # its buffer was consumed (written out) one iteration ago
        evs[c % 2].synchronize()
        yield bufs[c % 2][:m].numpy()
    t.record_stream(side)


def _np_dtype(v):
    if isinstance(v, torch.Tensor):
        return np.dtype(str(v.dtype).replace("torch.", ""))
    return np.asarray(v).dtype


def _write_npy_member(zf, name, v, ready=None):
    """One uncompressed ``.npy`` member of an ``.npz`` archive, streamed chunk by chunk."""
    if isinstance(v, torch.Tensor) and v.dim() != 1:
        v = v.cpu().numpy()
    if isinstance(v, torch.Tensor) or (isinstance(v, np.ndarray) and v.ndim == 1 and v.dtype.kind in "fiu"):
        n = int(v.numel() if isinstance(v, torch.Tensor) else v.size)
        hdr = {"descr": np.lib.format.dtype_to_descr(_np_dtype(v)), "fortran_order": False, "shape": (n,)}
        with zf.open(name + ".npy", "w", force_zip64=True) as f:
            np.lib.format.write_array_header_2_0(f, hdr)
            for chunk in _host_chunks(v, ready=ready):
                f.write(memoryview(chunk).cast("B"))
    else:
        with zf.open(name + ".npy", "w", force_zip64=True) as f:
            np.lib.format.write_array(f, np.asarray(v), allow_pickle=False)


_PENDING = {}          # file base -> _BackgroundWrite of a list still on its way to the disk
ASYNC_WRITES = bool(os.environ.get("PXB_ASYNC_WRITES"))


class _BackgroundWrite(threading.Thread):
    """Writer thread of one list: streams the arrays from HBM to the file while the caller goes
    on (the next stage reads the list from HBM, see load_list).  Errors surface in wait()."""

    def __init__(self, base, lst, ready):
        super().__init__(daemon=False, name=f"pxb-writer:{base}")
        self.base, self.lst, self.ready, self.error, self.file = base, lst, ready, None, None
        self.device = torch.cuda.current_device() if torch.cuda.is_available() else None

    def run(self):
        try:
            if self.device is not None:
                torch.cuda.set_device(self.device)      # (a new thread starts on device 0)
            self.file = _write_file(self.base, self.lst, self.ready)
        except BaseException as e:  # noqa: BLE001 - re-raised in wait()
            self.error = e

    def wait(self):
        self.join()
        if self.error is not None:
            raise self.error
        return self.file


def set_async_writes(flag=True):
    """Write file-backed lists in the background: ``make_photons`` / ``project_photons`` return
    as soon as the list is complete in HBM, a writer thread streams it to the file, and a later
    stage given the same prefix reads the list from HBM instead of from the file.
    ``wait_for_writes()`` (also run at interpreter exit) joins the writers and re-raises their
    errors.  Off by default: the reference returns with its files complete."""
    global ASYNC_WRITES
    ASYNC_WRITES = bool(flag)
    if not flag:
        wait_for_writes()


def wait_for_writes(prefix=None):
    """Block until the background write of ``prefix`` (default: all of them) is on the disk."""
    keys = list(_PENDING) if prefix is None else [k for k in (_file_name(_strip(prefix)),) if k in _PENDING]
    err = None
    for k in keys:
        w = _PENDING.pop(k)
        try:
            w.wait()
        except BaseException as e:  # noqa: BLE001
            err = err or e
    if err is not None:
        raise err


atexit.register(wait_for_writes)


def save_list(prefix, lst):
    """Write a list under ``prefix`` (``mem:`` -> HBM registry, else ``.h5`` / ``.npz``).  The
    data arrays go from the device to the file in slabs (``_host_chunks``), the download of one
    overlapping the write of the previous; with ``set_async_writes`` a writer thread does it
    while the caller continues."""
    if prefix.startswith("mem:"):
        _MEMORY[_file_name(prefix)] = lst
        return _file_name(prefix)
    base = _file_name(prefix)
    wait_for_writes(prefix)                 # (the same name written twice: in order)
    if not ASYNC_WRITES:
        return _write_file(base, lst, None)
    ready = None
    if torch.cuda.is_available() and any(isinstance(v, torch.Tensor) and v.is_cuda for v in lst.data.values()):
        ready = torch.cuda.Event()
        ready.record()
    w = _BackgroundWrite(base, lst, ready)
    _PENDING[base] = w
    w.start()
    return base + (".h5" if h5py is not None else ".npz")


def _write_file(base, lst, ready):
    if h5py is not None:
        fn = base + ".h5"
        with h5py.File(fn, "w") as f:
            g = f.create_group("info")
            for k, v in lst.info.items():
                g.attrs[k] = v
            p = f.create_group("parameters")
            for k, v in lst.parameters.items():
                if isinstance(v, np.ndarray) and v.dtype.kind == "U":
                    v = v.astype("S")      # h5py cannot store NumPy unicode arrays
                p.create_dataset(k, data=v)
            d = f.create_group("data")
            for k, v in lst.data.items():
                if (v.dim() if isinstance(v, torch.Tensor) else np.ndim(v)) != 1:
                    d.create_dataset(k, data=v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                    continue
                n = int(v.numel() if isinstance(v, torch.Tensor) else np.asarray(v).size)
                ds = d.create_dataset(k, shape=(n,), dtype=_np_dtype(v))
                o = 0
                for chunk in _host_chunks(v, ready=ready):
                    ds[o:o + chunk.size] = chunk
                    o += chunk.size
        return fn
    import zipfile

    fn = base + ".npz"
    with zipfile.ZipFile(fn, "w", zipfile.ZIP_STORED, allowZip64=True) as zf:
        for k, v in lst.info.items():
            _write_npy_member(zf, f"info/{k}", np.asarray(v))
        for k, v in lst.parameters.items():
            _write_npy_member(zf, f"parameters/{k}", np.asarray(v))
        for k, v in lst.data.items():
            _write_npy_member(zf, f"data/{k}", v, ready)
    return fn


def load_list(prefix):
    if prefix.startswith("mem:"):
        key = _file_name(prefix)
        if key not in _MEMORY:
            raise FileNotFoundError(f"no in-memory list named {key!r}")
        return _MEMORY[key]
    base = _file_name(prefix)
    if base in _PENDING:                    # still being written: the list is in HBM
        return _PENDING[base].lst
    if h5py is not None and os.path.exists(base + ".h5"):
        with h5py.File(base + ".h5", "r") as f:
            info = dict(f["info"].attrs)
            params = {}
            for k, v in f["parameters"].items():
                val = v[()]
                params[k] = val.decode() if isinstance(val, bytes) else val
            data = {k: v[()] for k, v in f["data"].items()}
        return DeviceList(info, params, data)
    if os.path.exists(base + ".npz"):
        z = np.load(base + ".npz", allow_pickle=False)
        lst = DeviceList()
        for k in z.files:
            grp, name = k.split("/", 1)
            v = z[k]
            if grp == "data":
                lst.data[name] = v
            else:
                v = v[()] if v.ndim == 0 else v
                v = str(v) if isinstance(v, np.str_) else v
                (lst.info if grp == "info" else lst.parameters)[name] = v
        return lst
    raise FileNotFoundError(f"no photon/event list at {base}.h5 or {base}.npz")


def drop_list(prefix):
    _MEMORY.pop(_file_name(prefix), None)


# =======================================================================================
# stage 1: make_photons
# =======================================================================================
def determine_fields(data_source, source_type, point_sources):
    """photon_list.py:41-71 for ArrayDataSource / yt objects."""
    if isinstance(data_source, ArrayDataSource):
        p, v, w = (data_source.position_fields, data_source.velocity_fields,
                   data_source.width_field)
        if point_sources:
            w = None
        return p, v, w, data_source.data_type
    ds = data_source.ds
    from yt.geometry.particle_geometry_handler import ParticleIndex  # pragma: no cover

    ptype = ((source_type in ds.particle_types) | (source_type in ds.known_filters)
             | ((source_type == "gas") & isinstance(ds.index, ParticleIndex)))
    if ptype:
        ppos = [f"particle_position_{ax}" for ax in "xyz"]
        pvel = [f"particle_velocity_{ax}" for ax in "xyz"]
        if source_type in ds.known_filters:
            if ds.known_filters[source_type].filtered_type == "gas":
                ppos = ["x", "y", "z"]
                pvel = [f"velocity_{ax}" for ax in "xyz"]
        elif source_type == "gas":
            source_type = ds._sph_ptypes[0]
        p = [(source_type, ppos[i]) for i in range(3)]
        v = [(source_type, pvel[i]) for i in range(3)]
        w = (source_type, "smoothing_length") if source_type in ds._sph_ptypes else None
        dt = "particles"
    else:
        p = [("index", ax) for ax in "xyz"]
        v = [(source_type, f"velocity_{ax}") for ax in "xyz"]
        w = ("index", "dx")
        dt = "cells"
    if point_sources:
        w = None
    return p, v, w, dt


def find_object_bounds(data_source):
    """photon_list.py:74-108, in kpc."""
    if isinstance(data_source, ArrayDataSource):
        return np.asarray(data_source.left_edge, "float64"), np.asarray(data_source.right_edge, "float64")
    src = getattr(data_source, "base_object", data_source)  # pragma: no cover
    if hasattr(src, "left_edge"):
        c = 0.5 * (src.left_edge + src.right_edge)
        w = src.right_edge - src.left_edge
        le, re = -0.5 * w + c, 0.5 * w + c
    elif hasattr(src, "radius") and not hasattr(src, "height"):
        le, re = -src.radius + src.center, src.radius + src.center
    else:
        mylog.warning("You are using a region that is not currently supported for straddling "
                      "periodic boundaries. Check to make sure that your region does not do so.")
        le, re = data_source.ds.domain_left_edge, data_source.ds.domain_right_edge
    return le.to_value("kpc"), re.to_value("kpc")


def _chunks(data_source):
    if isinstance(data_source, ArrayDataSource):
        return data_source.chunks(rank=parallel.rank(), size=parallel.size())
    citer = data_source.chunks([], "io")  # pragma: no cover
    return (c for j, c in enumerate(citer) if j % parallel.size() == parallel.rank())


class _Generation:
    """What make_photons derives from its arguments before it touches a chunk
    (photon_list.py:214-343): distance / redshift rules, centre, the ``parameters`` group of the
    photon list, ``spectral_norm``, the fields to gather; the source model is set up."""


def _setup_generation(data_source, redshift, area, exp_time, source_model, point_sources,
                      parameters, center, dist, cosmology, velocity_fields, bulk_velocity, observer):
    is_array = isinstance(data_source, ArrayDataSource)
    ds = data_source if is_array else data_source.ds
    parameters = dict(parameters or {})
    if cosmology is None:
        cosmo = getattr(ds, "cosmology", None) or Cosmology()
    else:
        cosmo = cosmology

    # distance / redshift rules, photon_list.py:232-249
    if observer == "external":
        if dist is None:
            if redshift <= 0.0:
                msg = ("If redshift <= 0.0, you must specify a distance to the source using the "
                       "'dist' argument!")
                mylog.error(msg)
                raise ValueError(msg)
            if hasattr(cosmo, "angular_diameter_distance_mpc"):
                D_A_mpc = cosmo.angular_diameter_distance_mpc(0.0, redshift)
            else:  # yt Cosmology
                D_A_mpc = float(cosmo.angular_diameter_distance(0.0, redshift).to_value("Mpc"))
        else:
            D_A_mpc = parse_value(dist, "kpc") * 1.0e-3
            if redshift > 0.0:
                mylog.warning("Redshift must be zero for nearby sources. Resetting redshift to 0.0.")
                redshift = 0.0
    else:
        D_A_mpc = 0.0
        if redshift > 0.0:
            mylog.warning("Redshift must be zero for internal observers. Resetting redshift to 0.0.")
            redshift = 0.0

    if is_array:
        if center is None or (isinstance(center, str) and center in ("c", "center")):
            c = np.asarray(data_source.center, "float64")
        elif isinstance(center, str):
            raise ValueError("ArrayDataSource supports center=None, 'c' or an explicit [x,y,z] in kpc")
        else:
            c = np.asarray(center, "float64")
        dw = np.asarray(data_source.domain_width, "float64")
        periodicity = data_source.periodicity
    else:  # pragma: no cover - needs yt
        if isinstance(center, str):
            cc = ds.domain_center if center in ("center", "c") else ds.find_max("density")[-1]
        elif center is None:
            cc = (0.5 * (data_source.left_edge + data_source.right_edge)
                  if hasattr(data_source, "left_edge") else data_source.get_field_parameter("center"))
        else:
            cc = ds.arr(center, "code_length") if not hasattr(center, "units") else center
        c = cc.to_value("kpc")
        dw = ds.domain_width.to_value("kpc")
        periodicity = ds.periodicity
    if bulk_velocity is None:
        bulk_velocity = np.zeros(3)
    elif hasattr(bulk_velocity, "to_value"):
        bulk_velocity = bulk_velocity.to_value("km/s")
    bulk_velocity = np.asarray(bulk_velocity, "float64")

    exp_time_s = parse_value(exp_time, "s")
    area_cm2 = parse_value(area, "cm**2")
    parameters.update(
        fid_exp_time=exp_time_s, fid_area=area_cm2, fid_redshift=redshift, observer=observer,
        hubble=cosmo.hubble_constant, omega_matter=cosmo.omega_matter,
        omega_lambda=cosmo.omega_lambda, center=c, bulk_velocity=bulk_velocity,
        fid_d_a=D_A_mpc)

    # spectral_norm, photon_list.py:304-310
    D_A_cm = D_A_mpc * K.cm_per_mpc
    dist_fac = 1.0 / (4.0 * np.pi)
    if observer == "external":
        dist_fac /= D_A_cm * D_A_cm * (1.0 + redshift) ** 2
    spectral_norm = area_cm2 * exp_time_s * dist_fac

    le, re = find_object_bounds(data_source)
    source_model.setup_model("photons", data_source, redshift)
    p_fields, v_fields, w_field, data_type = determine_fields(data_source, source_model.ftype,
                                                              point_sources)
    if velocity_fields is not None:
        v_fields = velocity_fields
    parameters["data_type"] = data_type
    # stored like the reference does (photon_list.py:371): a byte-string array, shape (3, 2) for
    # (ftype, name) tuples -- h5py has no conversion path for NumPy unicode arrays
    parameters["velocity_fields"] = np.array(v_fields).astype("S")
    source_model.set_pv(p_fields, v_fields, le=le, re=re, dw=dw, c=c, periodicity=periodicity,
                        observer=observer)

    G = _Generation()
    G.ds, G.parameters, G.spectral_norm, G.data_type = ds, parameters, spectral_norm, data_type
    G.redshift, G.D_A_mpc, G.observer = redshift, D_A_mpc, observer
    G.exp_time, G.area = exp_time_s, area_cm2
    G.gather = dict(p_fields=p_fields, v_fields=v_fields, w_field=w_field, bulk_velocity=bulk_velocity)
    return G


def make_photons(photon_prefix, data_source, redshift, area, exp_time, source_model,
                 point_sources=False, parameters=None, center=None, dist=None, cosmology=None,
                 velocity_fields=None, bulk_velocity=None, observer="external",
                 fields_to_keep=None):
    """Generate a photon list; signature and return value of pyxsim.make_photons
    (photon_list.py:111-127,478): returns ``(n_photons, n_cells)`` summed over ranks."""
    require_cuda()
    photon_prefix = _strip(photon_prefix)
    G = _setup_generation(data_source, redshift, area, exp_time, source_model, point_sources,
                          parameters, center, dist, cosmology, velocity_fields, bulk_velocity, observer)
    ds, parameters, spectral_norm, gather = G.ds, G.parameters, G.spectral_norm, G.gather
    native = hasattr(source_model, "generate_chunk")
    pieces = []
    n_photons = n_cells = 0
    host_cursor = 0          # running global cell id for host (reference-style) models on yt chunks
    for chunk in _chunks(data_source):
        if native:
            out = source_model.generate_chunk(chunk, spectral_norm, gather=gather)
        else:
            out = _generate_with_reference_model(source_model, chunk, spectral_norm, gather)
            if not isinstance(chunk, ArrayChunk) and out is not None:
                out._id0 = host_cursor
                host_cursor += out._chunk_cells
        if out is None or out.n_photons == 0:
            continue
        # global ids of the active cells: the projection keys its RNG counters by them, so the
        # events of a cell do not depend on chunking or on the number of ranks
        if native:
            out.gid = source_model._global_ids(out.idx)
        elif isinstance(chunk, ArrayChunk):
            out.gid = chunk.global_ids(out.idx)
        else:
            out.gid = out.idx + out._id0
        if fields_to_keep:
            from .source_models import field_value

            out_extra = {}
            for f in fields_to_keep:
                v = to_device(field_value(chunk, f))
                out_extra[f[1] if isinstance(f, tuple) else f] = v[out.idx]
            pieces.append((out, out_extra))
        else:
            pieces.append((out, {}))
        n_photons += out.n_photons
        n_cells += out.n_cells
    source_model.cleanup_model("photons")

    def cat(name):
        ts = [getattr(p, name) for p, _ in pieces]
        if not ts:
            return empty(0, torch.int64 if name in ("nph", "gid") else torch.float64)
        return ts[0] if len(ts) == 1 else torch.cat(ts)

    data = {k: cat(k) for k in ("x", "y", "z", "vx", "vy", "vz", "dx")}
    data["num_photons"] = cat("nph")
    data["energy"] = cat("energy")
    data["cell_id"] = cat("gid") if pieces else empty(0, torch.int64)
    if fields_to_keep:
        for name in pieces[0][1] if pieces else []:
            data[name] = torch.cat([e[name] for _, e in pieces])
    info = dict(pyxsim_version=__version__, yt_version="n/a", soxs_version="n/a",
                dataset=str(ds), data_source=str(data_source), source_model=repr(source_model),
                filenames=_file_names(photon_prefix))
    save_list(photon_prefix, DeviceList(info, parameters, data))

    all_nphotons = parallel.allreduce_sum(n_photons)
    all_ncells = parallel.allreduce_sum(n_cells)
    mylog.info("Finished generating photons.")
    mylog.info("Number of photons generated: %d", all_nphotons)
    mylog.info("Number of cells with photons: %d", all_ncells)
    return all_nphotons, all_ncells


def _generate_with_reference_model(source_model, chunk, spectral_norm, gather):
    """Drive a reference-style (host) source model through ``process_data`` and gather the
    active cells on the device -- keeps ``make_photons`` usable with third-party plugins."""
    from .source_models import ChunkPhotons, field_value

    res = source_model.process_data("photons", chunk, spectral_norm)
    n_chunk = int(np.prod(np.shape(field_value(chunk, gather["p_fields"][0], "kpc"))))
    if res is None:
        out = ChunkPhotons()
        out.n_cells = out.n_photons = 0
        out._chunk_cells = n_chunk
        return out
    ncells, nph, idxs, energies = res
    idxs = np.asarray(idxs)
    if idxs.dtype == bool:
        idxs = np.where(idxs)[0]
    out = ChunkPhotons()
    out._chunk_cells = n_chunk
    out.n_cells, out.n_photons = int(ncells), int(np.sum(nph))
    if out.n_photons == 0:
        return out
    out.idx = to_device(idxs, torch.int64)
    out.nph = to_device(nph, torch.int64)
    out.energy = to_device(energies)
    n = field_value(chunk, gather["p_fields"][0], "kpc").shape[0]
    counts = torch.zeros(n, dtype=torch.int64, device=out.idx.device)
    counts[out.idx] = out.nph
    tmp = source_model if hasattr(source_model, "_scan_and_compact") else None
    if tmp is None:
        from .source_models import SourceModel

        tmp = SourceModel()
        tmp.set_pv(source_model.p_fields, source_model.v_fields, le=source_model.le,
                   re=source_model.re, dw=source_model.dw, c=source_model.c,
                   periodicity=source_model.periodicity, observer=source_model.observer)
    full = tmp._scan_and_compact(counts, chunk, gather)
    for name in ("x", "y", "z", "vx", "vy", "vz", "dx", "poff"):
        setattr(out, name, getattr(full, name))
    return out


# =======================================================================================
# stage 2: project_photons
# =======================================================================================
def _exclusive_offsets(counts):
    """``(poff, qoff, n_pairs)``: int64[n+1] exclusive prefixes of ``counts`` and of
    ``ceil(counts / 2)`` (photon pairs, the photon kernels' unit of work) via libpxb's scan."""
    n = counts.numel()
    need = C.c_int64()
    _lib.call("pxb_scan_workspace_bytes", n, C.byref(need))
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=counts.device)
    offsets, arank, pairoff = (empty(n + 1, torch.int64) for _ in range(3))
    totals = empty(3, torch.int64)
    _lib.call("pxb_scan_counts", ptr(counts), n, ptr(offsets), ptr(arank), ptr(pairoff), ptr(totals),
              ptr(ws), ws.numel(), stream_ptr())
    return offsets, pairoff, int(totals.cpu()[2])


def _as_device_f64(v):
    return to_device(v, torch.float64)


class _Projection:
    """Everything ``_project_photons`` derives from its arguments (photon_list.py:481-570,
    596-640): validation, the absorption model, the column file, the ``parameters`` group of the
    event list and the ``pxb_project_params`` block.  Shared by ``project_photons`` and the fused
    ``make_events``."""

    def __init__(self, obs, normal, sky_center, absorb_model=None, nH=None, abund_table="angr",
                 no_shifting=False, north_vector=None, flat_sky=False, sigma_pos=None,
                 kernel="top_hat", save_los=False, phys_coord=False, prng=None, column_file=None,
                 nH_cell=None, event_dtype="float64"):
        self.obs, self.normal, self.nH, self.abund_table = obs, normal, nH, abund_table
        if str(event_dtype) not in ("float64", "float32"):
            raise ValueError("event_dtype must be 'float64' or 'float32'")
        # float32 events (an extension; the reference writes float64): energies rounded to single
        # precision and sky positions stored as single-precision OFFSETS from the fp64 sky centre
        # (the kernels compute in fp64 either way); half the bytes per event
        self.f32 = str(event_dtype) == "float32"
        self.tdtype = torch.float32 if self.f32 else torch.float64
        self.no_shifting, self.north_vector, self.flat_sky = no_shifting, north_vector, flat_sky
        self.sigma_pos, self.kernel, self.save_los, self.phys_coord = sigma_pos, kernel, save_los, phys_coord
        self.nH_cell = nH_cell
        self.seed = prng_seed(prng)
        L, self.north_out, self.uv = get_normal_and_north(normal, north_vector=north_vector)
        self.col = None
        if column_file is not None:
            if absorb_model is None:
                raise ValueError("You must specify an absorption model if you are using an absorption file!")
            if obs == "internal":
                raise ValueError("You cannot use an absorption file with an internal observer!")
            if nH_cell is not None:
                raise ValueError("Give either column_file or nH_cell, not both!")
            self.col = read_column_file(column_file)
            if not np.isclose(self.col["normal"], L).all():
                raise ValueError("The value of the normal vector in the absorption file does not match "
                                 "the value in the call to project_photons!")
            if not np.isclose(self.col["north"], self.north_out).all():
                raise ValueError("The value of the north vector in the absorption file does not match "
                                 "the value in the call to project_photons!")
        self.sky_center = np.asarray(sky_center, dtype="float64")
        if isinstance(absorb_model, str):
            if absorb_model not in absorb_models:
                raise KeyError(f"{absorb_model} is not a known absorption model!")
            absorb_model = absorb_models[absorb_model]
        if absorb_model is not None and isinstance(absorb_model, type):
            absorb_model = absorb_model(nH, abund_table=abund_table)
        # (a ready-made instance is accepted too; it is never modified: ``nH`` travels in the
        # parameter block)
        if absorb_model is not None and nH is not None and parallel.rank() == 0:
            mylog.info("Foreground galactic absorption: using the %s model and nH = %g.",
                       absorb_model._name, nH)
        self.absorb_model = absorb_model
        if kernel not in ("top_hat", "gaussian"):
            raise ValueError(f"unknown kernel {kernel!r}")
        if nH_cell is not None and absorb_model is None:
            raise ValueError("You must specify an absorption model to use per-cell columns!")

    def check_list(self, observer, data_type):
        """photon_list.py:581-594"""
        if self.obs != observer:
            which = {"external": "", "internal": "_allsky"}
            raise RuntimeError(f"The function 'project_photons{which[self.obs]}' does not work with "
                               f"'{observer}' photon lists!")
        if observer == "internal" and isinstance(self.normal, str):
            raise RuntimeError("Must specify a vector for 'normal' if you are doing an 'internal' observation!")
        if self.sigma_pos is not None and data_type == "particles":
            raise RuntimeError("The 'smooth_positions' argument should not be used with particle-based datasets!")

    def event_params(self, exp_time, area, observer, redshift):
        """The ``parameters`` group of the event list (photon_list.py:614-640)."""
        normal = self.normal
        params = dict(exp_time=exp_time, area=area, sky_center=self.sky_center, observer=observer,
                      no_shifting=int(self.no_shifting), flat_sky=int(self.flat_sky),
                      phys_coord=int(self.phys_coord),
                      normal=normal if isinstance(normal, str) else np.asarray(normal, "float64"))
        if self.north_vector is not None:
            params["north_vector"] = np.asarray(self.north_vector, "float64")
        # (sic) key spelled as in photon_list.py:626
        params["absoption_model"] = self.absorb_model._name if self.absorb_model else "none"
        if self.absorb_model is not None:
            if self.nH is not None:      # (a column_file-only run has no foreground column to record)
                params["nH"] = self.nH
            params["abund_table"] = self.abund_table
        if self.sigma_pos is not None:
            params["sigma_pos"] = self.sigma_pos
        params["kernel"] = self.kernel
        params["redshift"] = redshift
        if self.f32:
            params["event_dtype"] = "float32"
            params["sky_offsets"] = int(observer == "external" and not self.phys_coord)
        return params

    def fill(self, observer, data_type, redshift, D_A_kpc, xyz, cell_id0=0):
        """``(pxb_project_params, tensors to keep alive)``; ``xyz``: device positions of the active
        cells (needed for a column file)."""
        normal = self.normal
        x_hat, y_hat, z_hat = self.uv[0], self.uv[1], self.uv[2]
        P = _lib.ProjectParams()
        P.observer = int(observer == "internal")
        P.data_type = int(data_type == "particles")
        P.kernel = int(self.kernel == "gaussian")
        P.no_shifting = int(bool(self.no_shifting))
        if isinstance(normal, str) and observer == "external":
            ax = "xyz".index(normal)
            P.vn_axis = ax
            # on-axis scatter_events: (xsky, ysky, los) <- axes (ax+1, ax+2, ax)
            # (sky_functions.pyx:60-70,92-104)
            eye = np.identity(3)
            x_hat, y_hat, z_hat = eye[(ax + 1) % 3], eye[(ax + 2) % 3], eye[ax]
        else:
            P.vn_axis = -1
        P.sky_mode = _lib.SKY_PHYS if self.phys_coord else (_lib.SKY_FLAT if self.flat_sky else _lib.SKY_TAN)
        P.save_los = int(bool(self.save_los))
        P.has_sigma_pos = int(self.sigma_pos is not None)
        P.sigma_pos = 0.0 if self.sigma_pos is None else float(self.sigma_pos)
        for i in range(3):
            P.x_hat[i], P.y_hat[i], P.z_hat[i] = x_hat[i], y_hat[i], z_hat[i]
        P.D_A_kpc = D_A_kpc
        P.sky_center[0], P.sky_center[1] = float(self.sky_center[0]), float(self.sky_center[1])
        keep = []
        if self.absorb_model is not None:
            keep = self.absorb_model._fill_params(P)
            P.nH = 0.0 if self.nH is None else float(self.nH)
        nH_cell = self.nH_cell
        if self.col is not None:
            # photon_list.py:672-680: the cube is sampled at the cells' positions in the
            # (east, north, normal) frame of the Orientation, also for an on-axis normal
            nH_cell = column_density_at_cells(xyz[0], xyz[1], xyz[2], self.uv, self.col["wbins"],
                                              self.col["dbins"], self.col["nH"])
        if nH_cell is not None:
            nhc = to_device(nH_cell)
            keep.append(nhc)
            P.nH_cell = nhc.data_ptr()
        P.oneplusz = 1.0 + redshift
        P.seed = self.seed
        P.cell_id0 = cell_id0
        P.generic_kernels = int(bool(os.environ.get("PXB_GENERIC_KERNELS")))
        P.events_f32 = int(self.f32)
        return P, keep


def _event_info(photon_file, event_prefix):
    return dict(pyxsim_version=__version__, yt_version="n/a", soxs_version="n/a",
                photon_file=photon_file, filenames=_file_names(event_prefix))


def _project_photons(obs, photon_prefix, event_prefix, normal, sky_center, **kw):
    require_cuda()
    photon_prefix = _strip(photon_prefix)
    event_prefix = _strip(event_prefix)
    pr = _Projection(obs, normal, sky_center, **kw)

    plist = load_list(photon_prefix)
    p = plist.parameters
    redshift = float(p["fid_redshift"])
    data_type = _as_str(p["data_type"])
    observer = _as_str(p.get("observer", "external"))
    pr.check_list(observer, data_type)
    D_A = float(p["fid_d_a"]) * 1.0e3  # Mpc -> kpc, photon_list.py:596

    d = plist.data
    n_photons = int(d["energy"].shape[0])
    n_cells = int(d["num_photons"].shape[0])
    exp_time = float(p["fid_exp_time"])
    area = float(p["fid_area"])
    params = pr.event_params(exp_time, area, observer, redshift)
    info = _event_info(_file_name(photon_prefix) if photon_prefix.startswith("mem:")
                       else _file_name(photon_prefix) + ".h5", event_prefix)

    # RNG counters are keyed by the cells' GLOBAL ids.  Lists written by this package carry them;
    # for any other list (the reference's) the rank's exclusive prefix of the active cell counts
    # makes the keys distinct across ranks (the event-count all-gather of the reference's MPI
    # mode, photon_list.py:826, applied to the cells).
    cell_id0 = 0
    if "cell_id" not in d and parallel.size() > 1:      # (collective: every rank, photons or not)
        per_rank = parallel.allgather_counts([n_cells])[:, 0].tolist()
        cell_id0 = int(sum(per_rank[:parallel.rank()]))
    n_events = 0
    if n_photons == 0:
        mylog.warning("No photons are in file %s, so I am done.", photon_prefix)
    else:
        dev = {k: (_as_device_f64(d[k]) if k != "num_photons" else to_device(d[k], torch.int64))
               for k in ("x", "y", "z", "vx", "vy", "vz", "dx", "num_photons", "energy")}
        poff, qoff, n_pairs = _exclusive_offsets(dev["num_photons"])
        PL = _lib.PhotonList()
        PL.n_cells, PL.n_photons, PL.n_pairs = n_cells, n_photons, n_pairs
        PL.nph, PL.poff, PL.qoff = dev["num_photons"].data_ptr(), poff.data_ptr(), qoff.data_ptr()
        for k in ("x", "y", "z", "vx", "vy", "vz", "dx"):
            setattr(PL, k, dev[k].data_ptr())
        PL.energy = dev["energy"].data_ptr()
        cid = None
        if "cell_id" in d:
            cid = to_device(d["cell_id"], torch.int64)
            PL.cell_id = cid.data_ptr()
        P, keep = pr.fill(observer, data_type, redshift, D_A, (dev["x"], dev["y"], dev["z"]), cell_id0)

        need = C.c_int64()
        _lib.call("pxb_project_workspace_bytes", n_cells, n_pairs, C.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev["energy"].device)
        xsky, ysky, eobs = (empty(n_photons, pr.tdtype) for _ in range(3))
        los = empty(n_photons, pr.tdtype) if pr.save_los else None
        nev = empty(1, torch.int64)
        stats = empty(3)
        with stage("project"):
            _lib.call("pxb_project", C.byref(PL), C.byref(P), ptr(xsky), ptr(ysky), ptr(eobs),
                      ptr(los), ptr(nev), ptr(stats), ptr(ws), need.value, stream_ptr())
        n_events = int(nev.cpu()[0])
        st = stats.cpu().numpy()
        del keep, ws, cid
        data = {"xsky": xsky[:n_events], "ysky": ysky[:n_events], "eobs": eobs[:n_events]}
        if pr.save_los:
            data["los"] = los[:n_events]
        params["flux"] = float(st[0]) * K.erg_per_keV / exp_time / area   # photon_list.py:816
        params["emin"] = float(st[1])
        params["emax"] = float(st[2])
        save_list(event_prefix, DeviceList(info, params, data))

    all_nevents = parallel.allreduce_sum(n_events)
    mylog.info("Detected %d events.", all_nevents)
    return all_nevents


def _as_str(v):
    return v.decode() if isinstance(v, bytes) else str(v)


# =======================================================================================
# both stages in one pass: make_events
# =======================================================================================
def make_events(event_prefix, data_source, redshift, area, exp_time, source_model, normal,
                sky_center, absorb_model=None, nH=None, abund_table="angr", no_shifting=False,
                north_vector=None, sigma_pos=None, flat_sky=False, kernel="top_hat", save_los=False,
                phys_coord=False, prng=None, point_sources=False, parameters=None, center=None,
                dist=None, cosmology=None, velocity_fields=None, bulk_velocity=None,
                event_dtype="float64"):
    """``make_photons`` followed by ``project_photons`` (external observer) in ONE pass over the
    cells: energies are sampled, shifted, absorbed and -- for the survivors only -- scattered onto
    the sky without ever writing a photon list (the two reference stages,
    thermal_sources/base.py:462-492 and photon_list.py:686-799, back to back).  The arguments are
    those of the two reference calls; ``source_model.prng`` seeds the generation, ``prng`` the
    projection.  The event list is bit for bit the one ``make_photons`` + ``project_photons``
    produce with the same seeds.  Returns ``(n_photons, n_cells, n_events)`` summed over ranks.

    Thermal sources with non-negative tables and the common sampler (linear bins, invert_cdf,
    no variable elements, 1-D table) run fused; anything else (power-law / line sources, tables
    with negative entries, cells extrapolating beyond the table, reference-style host models, and
    the option-generic sampler, for which the two stages are faster) takes the two stages through
    an in-memory photon list."""
    require_cuda()
    event_prefix = _strip(event_prefix)
    pr = _Projection("external", normal, sky_center, absorb_model=absorb_model, nH=nH,
                     abund_table=abund_table, no_shifting=no_shifting, north_vector=north_vector,
                     flat_sky=flat_sky, sigma_pos=sigma_pos, kernel=kernel, save_los=save_los,
                     phys_coord=phys_coord, prng=prng, event_dtype=event_dtype)
    fusable = hasattr(source_model, "count_chunk") and hasattr(source_model, "spectral_model")
    if fusable:
        G = _setup_generation(data_source, redshift, area, exp_time, source_model, point_sources,
                              parameters, center, dist, cosmology, velocity_fields, bulk_velocity,
                              "external")
        pr.check_list("external", G.data_type)
        dm = source_model.spectral_model.device_model(source_model.method)
        fusable = bool(_lib.load().pxb_model_nonnegative(dm.handle))
        # with the option-generic sampler the two stages are the faster route (same events);
        # PXB_FUSE_GENERIC=1 keeps the fused kernel on (diagnostics, parity tests)
        if fusable and not _lib.load().pxb_model_fusion_pays(dm.handle):
            fusable = bool(os.environ.get("PXB_FUSE_GENERIC"))
    if not fusable:
        return _make_events_two_stage(event_prefix, data_source, redshift, area, exp_time, source_model,
                                      pr, point_sources, parameters, center, dist, cosmology,
                                      velocity_fields, bulk_velocity)
    D_A_kpc = G.D_A_mpc * 1.0e3
    pieces = []
    n_photons = n_cells = n_events = 0
    flux_sum, emin, emax = 0.0, 1.0e99, -1.0e99
    declined = False
    for chunk in _chunks(data_source):
        out, Cs, keep_c = source_model.count_chunk(chunk, G.spectral_norm, gather=G.gather)
        if out is None or out.n_photons == 0:
            continue
        gid = source_model._global_ids(out.idx)
        PL = _lib.PhotonList()
        PL.n_cells, PL.n_photons, PL.n_pairs = out.n_cells, out.n_photons, out.n_pairs
        PL.nph, PL.poff, PL.qoff = out.nph.data_ptr(), out.poff.data_ptr(), out.qoff.data_ptr()
        for k in ("x", "y", "z", "vx", "vy", "vz"):
            setattr(PL, k, getattr(out, k).data_ptr())
        PL.dx = out.dx.data_ptr() if out.dx is not None else None
        PL.cell_id = gid.data_ptr()
        P, keep_p = pr.fill("external", G.data_type, G.redshift, D_A_kpc, (out.x, out.y, out.z))
        need = C.c_int64()
        _lib.call("pxb_make_events_workspace_bytes", dm.handle, out.n_cells, out.n_pairs, C.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=out.idx.device)
        xsky, ysky, eobs = (empty(out.n_photons, pr.tdtype) for _ in range(3))
        los = empty(out.n_photons, pr.tdtype) if pr.save_los else None
        nev = empty(2, torch.int64)          # [events, cells the mixture sampler declined]
        stats = empty(3)
        with stage("make_events"):
            _lib.call("pxb_make_events", dm.handle, C.byref(Cs), ptr(out.idx), C.byref(PL), C.byref(P),
                      ptr(xsky), ptr(ysky), ptr(eobs), ptr(los), ptr(nev), ptr(stats),
                      C.c_void_p(nev.data_ptr() + 8), ptr(ws), need.value, stream_ptr())
        ne, nd = (int(v) for v in nev.cpu())
        del ws, keep_p, keep_c
        if nd > 0:
            declined = True
            break
        st = stats.cpu().numpy()
        pieces.append((xsky[:ne], ysky[:ne], eobs[:ne], los[:ne] if pr.save_los else None))
        n_photons += out.n_photons
        n_cells += out.n_cells
        n_events += ne
        flux_sum += float(st[0])
        emin, emax = min(emin, float(st[1])), max(emax, float(st[2]))
    source_model.cleanup_model("photons")
    # (collective: every rank must agree before any of them falls back)
    if parallel.allreduce_sum(int(declined)) > 0:
        mylog.info("Some cells lie outside the spectral table: taking the two-stage path.")
        return _make_events_two_stage(event_prefix, data_source, redshift, area, exp_time, source_model,
                                      pr, point_sources, parameters, center, dist, cosmology,
                                      velocity_fields, bulk_velocity)

    def cat(j):
        ts = [pc[j] for pc in pieces]
        return empty(0, pr.tdtype) if not ts else (ts[0] if len(ts) == 1 else torch.cat(ts))

    data = {"xsky": cat(0), "ysky": cat(1), "eobs": cat(2)}
    if pr.save_los:
        data["los"] = cat(3)
    params = pr.event_params(G.exp_time, G.area, "external", G.redshift)
    params["flux"] = flux_sum * K.erg_per_keV / G.exp_time / G.area
    params["emin"], params["emax"] = emin, emax
    save_list(event_prefix, DeviceList(_event_info("none (fused make_events)", event_prefix), params, data))
    tot = [parallel.allreduce_sum(v) for v in (n_photons, n_cells, n_events)]
    mylog.info("Generated %d photons in %d cells; detected %d events.", *tot)
    return tuple(tot)


def _make_events_two_stage(event_prefix, data_source, redshift, area, exp_time, source_model, pr,
                           point_sources, parameters, center, dist, cosmology, velocity_fields,
                           bulk_velocity):
    tmp = f"mem:__make_events_{id(pr):x}"
    n_ph, n_c = make_photons(tmp, data_source, redshift, area, exp_time, source_model,
                             point_sources=point_sources, parameters=parameters, center=center,
                             dist=dist, cosmology=cosmology, velocity_fields=velocity_fields,
                             bulk_velocity=bulk_velocity)
    try:
        n_ev = _project_photons("external", tmp, event_prefix, pr.normal, pr.sky_center,
                                absorb_model=pr.absorb_model, nH=pr.nH, abund_table=pr.abund_table,
                                no_shifting=pr.no_shifting, north_vector=pr.north_vector,
                                flat_sky=pr.flat_sky, sigma_pos=pr.sigma_pos, kernel=pr.kernel,
                                save_los=pr.save_los, phys_coord=pr.phys_coord, prng=pr.seed,
                                event_dtype="float32" if pr.f32 else "float64")
    finally:
        drop_list(tmp)
    return n_ph, n_c, n_ev


def project_photons(photon_prefix, event_prefix, normal, sky_center, absorb_model=None, nH=None,
                    abund_table="angr", no_shifting=False, north_vector=None, sigma_pos=None,
                    flat_sky=False, kernel="top_hat", save_los=False, column_file=None,
                    phys_coord=False, prng=None, nH_cell=None, event_dtype="float64"):
    """pyxsim.project_photons (photon_list.py:833-962); returns the number of events.
    ``nH_cell`` (per-cell internal columns) and ``event_dtype`` ("float32": single-precision
    energies and sky offsets from ``sky_center``, half the bytes) are extensions."""
    return _project_photons("external", photon_prefix, event_prefix, normal, sky_center,
                            absorb_model=absorb_model, nH=nH, abund_table=abund_table,
                            no_shifting=no_shifting, north_vector=north_vector,
                            sigma_pos=sigma_pos, flat_sky=flat_sky, kernel=kernel,
                            save_los=save_los, phys_coord=phys_coord, prng=prng,
                            column_file=column_file, nH_cell=nH_cell, event_dtype=event_dtype)


def project_photons_allsky(photon_prefix, event_prefix, normal, absorb_model=None, nH=None,
                           abund_table="angr", no_shifting=False, center_vector=None,
                           kernel="top_hat", save_los=False, prng=None, event_dtype="float64"):
    """pyxsim.project_photons_allsky (photon_list.py:964-1060)."""
    return _project_photons("internal", photon_prefix, event_prefix, normal, [0.0, 0.0],
                            absorb_model=absorb_model, nH=nH, abund_table=abund_table,
                            no_shifting=no_shifting, kernel=kernel, save_los=save_los,
                            north_vector=center_vector, prng=prng, event_dtype=event_dtype)
