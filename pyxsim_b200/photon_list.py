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
import ctypes as C
import os

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


def save_list(prefix, lst):
    """Write a list under ``prefix`` (``mem:`` -> HBM registry, else ``.h5`` / ``.npz``)."""
    if prefix.startswith("mem:"):
        _MEMORY[_file_name(prefix)] = lst
        return _file_name(prefix)
    host = lst.to_host()
    base = _file_name(prefix)
    if h5py is not None:
        fn = base + ".h5"
        with h5py.File(fn, "w") as f:
            g = f.create_group("info")
            for k, v in host.info.items():
                g.attrs[k] = v
            p = f.create_group("parameters")
            for k, v in host.parameters.items():
                if isinstance(v, np.ndarray) and v.dtype.kind == "U":
                    v = v.astype("S")      # h5py cannot store NumPy unicode arrays
                p.create_dataset(k, data=v)
            d = f.create_group("data")
            for k, v in host.data.items():
                d.create_dataset(k, data=v)
        return fn
    fn = base + ".npz"
    flat = {}
    for k, v in host.info.items():
        flat[f"info/{k}"] = np.asarray(v)
    for k, v in host.parameters.items():
        flat[f"parameters/{k}"] = np.asarray(v)
    for k, v in host.data.items():
        flat[f"data/{k}"] = v
    np.savez(fn, **flat)
    return fn


def load_list(prefix):
    if prefix.startswith("mem:"):
        key = _file_name(prefix)
        if key not in _MEMORY:
            raise FileNotFoundError(f"no in-memory list named {key!r}")
        return _MEMORY[key]
    base = _file_name(prefix)
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


def make_photons(photon_prefix, data_source, redshift, area, exp_time, source_model,
                 point_sources=False, parameters=None, center=None, dist=None, cosmology=None,
                 velocity_fields=None, bulk_velocity=None, observer="external",
                 fields_to_keep=None):
    """Generate a photon list; signature and return value of pyxsim.make_photons
    (photon_list.py:111-127,478): returns ``(n_photons, n_cells)`` summed over ranks."""
    require_cuda()
    photon_prefix = _strip(photon_prefix)
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

    gather = dict(p_fields=p_fields, v_fields=v_fields, w_field=w_field, bulk_velocity=bulk_velocity)
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
    """int64[n+1] exclusive prefix of ``counts`` via libpxb's scan."""
    n = counts.numel()
    need = C.c_int64()
    _lib.call("pxb_scan_workspace_bytes", n, C.byref(need))
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=counts.device)
    offsets, arank, totals = empty(n + 1, torch.int64), empty(n + 1, torch.int64), empty(2, torch.int64)
    _lib.call("pxb_scan_counts", ptr(counts), n, ptr(offsets), ptr(arank), ptr(totals), ptr(ws),
              ws.numel(), stream_ptr())
    return offsets


def _as_device_f64(v):
    return to_device(v, torch.float64)


def _project_photons(obs, photon_prefix, event_prefix, normal, sky_center, absorb_model=None,
                     nH=None, abund_table="angr", no_shifting=False, north_vector=None,
                     flat_sky=False, sigma_pos=None, kernel="top_hat", save_los=False,
                     phys_coord=False, prng=None, column_file=None, nH_cell=None):
    require_cuda()
    photon_prefix = _strip(photon_prefix)
    event_prefix = _strip(event_prefix)
    seed = prng_seed(prng)

    L, north_vector_out, uv = get_normal_and_north(normal, north_vector=north_vector)
    x_hat, y_hat, z_hat = uv[0], uv[1], uv[2]

    if column_file is not None:
        if absorb_model is None:
            raise ValueError("You must specify an absorption model if you are using an absorption file!")
        if obs == "internal":
            raise ValueError("You cannot use an absorption file with an internal observer!")
        if nH_cell is not None:
            raise ValueError("Give either column_file or nH_cell, not both!")
        col = read_column_file(column_file)
        if not np.isclose(col["normal"], L).all():
            raise ValueError("The value of the normal vector in the absorption file does not match "
                             "the value in the call to project_photons!")
        if not np.isclose(col["north"], north_vector_out).all():
            raise ValueError("The value of the north vector in the absorption file does not match "
                             "the value in the call to project_photons!")

    sky_center = np.asarray(sky_center, dtype="float64")
    if isinstance(absorb_model, str):
        if absorb_model not in absorb_models:
            raise KeyError(f"{absorb_model} is not a known absorption model!")
        absorb_model = absorb_models[absorb_model]
    if absorb_model is not None and isinstance(absorb_model, type):
        absorb_model = absorb_model(nH, abund_table=abund_table)
    # (a ready-made instance is accepted too; it is never modified: ``nH`` travels in the
    # parameter block)
    if absorb_model is not None and nH is not None:
        if parallel.rank() == 0:
            mylog.info("Foreground galactic absorption: using the %s model and nH = %g.",
                       absorb_model._name, nH)
    abs_model_name = absorb_model._name if absorb_model else "none"

    plist = load_list(photon_prefix)
    p = plist.parameters
    redshift = float(p["fid_redshift"])
    data_type = str(p["data_type"])
    observer = str(p.get("observer", "external"))
    if obs != observer:
        which = {"external": "", "internal": "_allsky"}
        raise RuntimeError(
            f"The function 'project_photons{which[obs]}' does not work with '{observer}' photon lists!")
    if observer == "internal" and isinstance(normal, str):
        raise RuntimeError("Must specify a vector for 'normal' if you are doing an 'internal' observation!")
    if sigma_pos is not None and data_type == "particles":
        raise RuntimeError("The 'smooth_positions' argument should not be used with particle-based datasets!")
    if kernel not in ("top_hat", "gaussian"):
        raise ValueError(f"unknown kernel {kernel!r}")
    D_A = float(p["fid_d_a"]) * 1.0e3  # Mpc -> kpc, photon_list.py:596

    d = plist.data
    n_photons = int(d["energy"].shape[0])
    n_cells = int(d["num_photons"].shape[0])
    exp_time = float(p["fid_exp_time"])
    area = float(p["fid_area"])
    params = dict(exp_time=exp_time, area=area, sky_center=sky_center, observer=observer,
                  no_shifting=int(no_shifting), flat_sky=int(flat_sky), phys_coord=int(phys_coord),
                  normal=normal if isinstance(normal, str) else np.asarray(normal, "float64"))
    if north_vector is not None:
        params["north_vector"] = np.asarray(north_vector, "float64")
    params["absoption_model"] = abs_model_name  # (sic) key spelled as in photon_list.py:626
    if absorb_model is not None:
        if nH is not None:           # (a column_file-only run has no foreground column to record)
            params["nH"] = nH
        params["abund_table"] = abund_table
    if sigma_pos is not None:
        params["sigma_pos"] = sigma_pos
    params["kernel"] = kernel
    params["redshift"] = redshift
    info = dict(pyxsim_version=__version__, yt_version="n/a", soxs_version="n/a",
                photon_file=(_file_name(photon_prefix) if photon_prefix.startswith("mem:")
                             else _file_name(photon_prefix) + ".h5"),
                filenames=_file_names(event_prefix))

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
        poff = _exclusive_offsets(dev["num_photons"])
        PL = _lib.PhotonList()
        PL.n_cells, PL.n_photons = n_cells, n_photons
        PL.nph, PL.poff = dev["num_photons"].data_ptr(), poff.data_ptr()
        for k in ("x", "y", "z", "vx", "vy", "vz", "dx"):
            setattr(PL, k, dev[k].data_ptr())
        PL.energy = dev["energy"].data_ptr()
        # RNG counters are keyed by the cells' GLOBAL ids.  Lists written by this package carry
        # them; for any other list (the reference's) the rank's exclusive prefix of the active
        # cell counts makes the keys distinct across ranks (the event-count all-gather of the
        # reference's MPI mode, photon_list.py:826, applied to the cells).
        cid = None
        if "cell_id" in d:
            cid = to_device(d["cell_id"], torch.int64)
            PL.cell_id = cid.data_ptr()

        P = _lib.ProjectParams()
        P.observer = int(observer == "internal")
        P.data_type = int(data_type == "particles")
        P.kernel = int(kernel == "gaussian")
        P.no_shifting = int(bool(no_shifting))
        if isinstance(normal, str) and observer == "external":
            ax = "xyz".index(normal)
            P.vn_axis = ax
            # on-axis scatter_events: (xsky, ysky, los) <- axes (ax+1, ax+2, ax)
            # (sky_functions.pyx:60-70,92-104)
            eye = np.identity(3)
            x_hat, y_hat, z_hat = eye[(ax + 1) % 3], eye[(ax + 2) % 3], eye[ax]
        else:
            P.vn_axis = -1
        P.sky_mode = _lib.SKY_PHYS if phys_coord else (_lib.SKY_FLAT if flat_sky else _lib.SKY_TAN)
        P.save_los = int(bool(save_los))
        P.has_sigma_pos = int(sigma_pos is not None)
        P.sigma_pos = 0.0 if sigma_pos is None else float(sigma_pos)
        for i in range(3):
            P.x_hat[i], P.y_hat[i], P.z_hat[i] = x_hat[i], y_hat[i], z_hat[i]
        P.D_A_kpc = D_A
        P.sky_center[0], P.sky_center[1] = float(sky_center[0]), float(sky_center[1])
        keep = []
        if absorb_model is not None:
            keep = absorb_model._fill_params(P)
            P.nH = 0.0 if nH is None else float(nH)
        if column_file is not None:
            # photon_list.py:672-680: the cube is sampled at the cells' positions in the
            # (east, north, normal) frame of the Orientation, also for an on-axis normal
            nH_cell = column_density_at_cells(dev["x"], dev["y"], dev["z"], uv, col["wbins"],
                                              col["dbins"], col["nH"])
        if nH_cell is not None:
            if absorb_model is None:
                raise ValueError("You must specify an absorption model to use per-cell columns!")
            nhc = to_device(nH_cell)
            keep.append(nhc)
            P.nH_cell = nhc.data_ptr()
        P.oneplusz = 1.0 + redshift
        P.seed = seed
        P.cell_id0 = cell_id0
        P.single_pass = int(os.environ.get("PXB_PROJECT_MODE", "") == "fused")
        P.generic_kernels = int(bool(os.environ.get("PXB_GENERIC_KERNELS")))

        need = C.c_int64()
        _lib.call("pxb_project_workspace_bytes", n_cells, n_photons, C.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev["energy"].device)
        xsky, ysky, eobs = empty(n_photons), empty(n_photons), empty(n_photons)
        los = empty(n_photons) if save_los else None
        nev = empty(1, torch.int64)
        stats = empty(3)
        with stage("project"):
            _lib.call("pxb_project", C.byref(PL), C.byref(P), ptr(xsky), ptr(ysky), ptr(eobs),
                      ptr(los), ptr(nev), ptr(stats), ptr(ws), need.value, stream_ptr())
        n_events = int(nev.cpu()[0])
        st = stats.cpu().numpy()
        del keep, ws, cid
        data = {"xsky": xsky[:n_events], "ysky": ysky[:n_events], "eobs": eobs[:n_events]}
        if save_los:
            data["los"] = los[:n_events]
        params["flux"] = float(st[0]) * K.erg_per_keV / exp_time / area   # photon_list.py:816
        params["emin"] = float(st[1])
        params["emax"] = float(st[2])
        save_list(event_prefix, DeviceList(info, params, data))

    all_nevents = parallel.allreduce_sum(n_events)
    mylog.info("Detected %d events.", all_nevents)
    return all_nevents


def project_photons(photon_prefix, event_prefix, normal, sky_center, absorb_model=None, nH=None,
                    abund_table="angr", no_shifting=False, north_vector=None, sigma_pos=None,
                    flat_sky=False, kernel="top_hat", save_los=False, column_file=None,
                    phys_coord=False, prng=None, nH_cell=None):
    """pyxsim.project_photons (photon_list.py:833-962); returns the number of events."""
    return _project_photons("external", photon_prefix, event_prefix, normal, sky_center,
                            absorb_model=absorb_model, nH=nH, abund_table=abund_table,
                            no_shifting=no_shifting, north_vector=north_vector,
                            sigma_pos=sigma_pos, flat_sky=flat_sky, kernel=kernel,
                            save_los=save_los, phys_coord=phys_coord, prng=prng,
                            column_file=column_file, nH_cell=nH_cell)


def project_photons_allsky(photon_prefix, event_prefix, normal, absorb_model=None, nH=None,
                           abund_table="angr", no_shifting=False, center_vector=None,
                           kernel="top_hat", save_los=False, prng=None):
    """pyxsim.project_photons_allsky (photon_list.py:964-1060)."""
    return _project_photons("internal", photon_prefix, event_prefix, normal, [0.0, 0.0],
                            absorb_model=absorb_model, nH=nH, abund_table=abund_table,
                            no_shifting=no_shifting, kernel=kernel, save_los=save_los,
                            north_vector=center_vector, prng=prng)
