"""Event-list consumers ("next" row f4): ``EventList`` and ``merge_files``.

Mirrors ``pyxsim/event_list.py:17-77,181-318,320-385`` and ``pyxsim/utils.py:71-173``.  The
binning the reference does on the host (astropy WCS ``wcs_world2pix`` + ``np.histogram2d`` /
``np.histogram``) runs on the GPU where a ``"mem:"`` event list already lives
(``pxb_cel_to_pixel``, ``pxb_event_image``, ``pxb_event_spectrum``); FITS output goes through a
small self-contained writer (primary image HDU and binary tables with the reference's header
keywords -- astropy is not required).  SIMPUT export is soxs' and is not reimplemented.
"""
import os
from glob import glob

import numpy as np
import torch

from . import _lib
from .device import ptr, stream_ptr, to_device
from .photon_list import DeviceList, load_list, save_list, _strip
from .utils import mylog, parse_value


# ---------------------------------------------------------------------------------------
# minimal FITS writer (FITS standard 4.0: 80-char cards, 2880-byte blocks, big-endian data)
# ---------------------------------------------------------------------------------------
def _card(key, value=None, comment=""):
    if value is None:
        return f"{key:<80}"[:80]
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        r = repr(float(value)).upper()       # shortest string that round-trips (free format
        if "E" not in r and "." not in r and "N" not in r:    # when longer than 20 characters)
            r += ".0"
        v = f"{r:>20}"
    else:
        sv = "'" + str(value).replace("'", "''").ljust(8) + "'"
        v = f"{sv:<20}"
    c = f"{key:<8}= {v}"
    if comment:
        c += f" / {comment}"
    return f"{c:<80}"[:80]


def _pad(b, fill=b"\0"):
    r = (-len(b)) % 2880
    return b + fill * r


def _header(cards):
    text = "".join(cards) + f"{'END':<80}"
    return _pad(text.encode("ascii"), b" ")


def _primary_hdu(image=None, extra=()):
    cards = [_card("SIMPLE", True, "conforms to FITS standard")]
    if image is None:
        cards += [_card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True)]
        data = b""
    else:
        a = np.ascontiguousarray(image, dtype=">f8")
        cards += [_card("BITPIX", -64), _card("NAXIS", a.ndim)]
        cards += [_card(f"NAXIS{i + 1}", n) for i, n in enumerate(a.shape[::-1])]
        cards += [_card("EXTEND", True)]
        data = _pad(a.tobytes())
    cards += [_card(k, v) for k, v in extra]
    return _header(cards) + data


_TFORM = {"E": ">f4", "D": ">f8", "J": ">i4"}


def _bintable_hdu(name, columns, extra=()):
    """columns: list of (name, format letter in E/D/J, unit or None, array)."""
    nrow = len(columns[0][3]) if columns else 0
    dt = np.dtype([(c[0], _TFORM[c[1]]) for c in columns])
    rec = np.zeros(nrow, dtype=dt)
    for cname, _, _, arr in columns:
        rec[cname] = arr
    cards = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8),
             _card("NAXIS", 2), _card("NAXIS1", dt.itemsize), _card("NAXIS2", nrow), _card("PCOUNT", 0),
             _card("GCOUNT", 1), _card("TFIELDS", len(columns))]
    for i, (cname, fmt, unit, _) in enumerate(columns, start=1):
        cards += [_card(f"TTYPE{i}", cname), _card(f"TFORM{i}", fmt)]
        if unit:
            cards.append(_card(f"TUNIT{i}", unit))
    cards.append(_card("EXTNAME", name))
    cards += [_card(k, v) for k, v in extra]
    return _header(cards) + _pad(rec.tobytes())


def _write_fits(path, hdus, overwrite):
    if os.path.exists(path) and not overwrite:
        raise OSError(f"File {path} already exists. If you mean to replace it then use overwrite=True.")
    with open(path, "wb") as f:
        for h in hdus:
            f.write(h)


def read_fits(path):
    """Parse a FITS file written by this module (or any simple image/bintable file) into a list
    of ``(header dict, data)``; used by the tests and handy without astropy."""
    raw = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(raw):
        hdr, end = {}, False
        while not end:
            block = raw[pos:pos + 2880].decode("ascii")
            pos += 2880
            for i in range(0, 2880, 80):
                card = block[i:i + 80]
                key = card[:8].strip()
                if key == "END":
                    end = True
                    break
                if card[8:10] != "= ":
                    continue
                val = card[10:].split(" / ")[0].strip()
                if val.startswith("'"):
                    hdr[key] = val[1:val.rindex("'")].rstrip().replace("''", "'")
                elif val in ("T", "F"):
                    hdr[key] = val == "T"
                else:
                    try:
                        hdr[key] = int(val)
                    except ValueError:
                        hdr[key] = float(val)
        naxis = hdr.get("NAXIS", 0)
        shape = [hdr[f"NAXIS{i}"] for i in range(naxis, 0, -1)]
        nbytes = abs(hdr["BITPIX"]) // 8 * int(np.prod(shape)) if naxis else 0
        data = None
        if nbytes:
            buf = raw[pos:pos + nbytes]
            if hdr.get("XTENSION") == "BINTABLE":
                dt = np.dtype([(hdr[f"TTYPE{i}"], _TFORM[hdr[f"TFORM{i}"].lstrip("1")])
                               for i in range(1, hdr["TFIELDS"] + 1)])
                data = np.frombuffer(buf, dtype=dt)
            else:
                data = np.frombuffer(buf, dtype={-64: ">f8", -32: ">f4", 32: ">i4", 64: ">i8"}[hdr["BITPIX"]]
                                     ).reshape(shape)
            pos += nbytes + (-nbytes) % 2880
        out.append((hdr, data))
    return out


# ---------------------------------------------------------------------------------------
# EventList
# ---------------------------------------------------------------------------------------
class EventList:
    """Read event lists so they can be exported to other formats (event_list.py:17-77).

    ``filespec``: a ``"mem:"`` name, a file name / prefix (``.h5`` or the ``.npz`` fallback), a
    globbable path, or a list of those.  Lists that sit in HBM are binned in place."""

    def __init__(self, filespec):
        if isinstance(filespec, (list, tuple)):
            names = list(filespec)
        elif filespec.startswith("mem:"):
            names = [filespec]
        else:
            base = _strip(filespec[:-4] if filespec.endswith(".npz") else filespec)
            hits = sorted(glob(base + ".h5") + glob(base + ".npz") + glob(base + ".[0-9][0-9][0-9][0-9].h5")
                          + glob(base + ".[0-9][0-9][0-9][0-9].npz"))
            if not hits:
                raise RuntimeError(f"Invalid EventList file spec: {filespec}")
            names = sorted({h.rsplit(".", 1)[0] for h in hits})
        self.parameters, self.info, self.num_events, self.filenames, self._lists = {}, {}, [], [], []
        for nm in sorted(names):
            lst = load_list(nm if nm.startswith("mem:") else _strip(nm))
            n = int(lst.data["xsky"].shape[0]) if "xsky" in lst.data else 0
            if n == 0:
                continue
            if not self.parameters:
                self.parameters, self.info = dict(lst.parameters), dict(lst.info)
            self.num_events.append(n)
            self.filenames.append(nm)
            self._lists.append(lst)
        self.num_files = len(self.filenames)
        self.tot_num_events = int(np.sum(self.num_events)) if self.num_events else 0
        self.observer = str(self.parameters.get("observer", "external"))
        self._data = {}

    def __getitem__(self, item):
        if item not in self._data:
            self._data[item] = np.concatenate([self._dev(lst, item).cpu().numpy() for lst in self._lists]) \
                if self._lists else np.array([])
        return self._data[item]

    def _dev(self, lst, key):
        """Column ``key`` of one list as a float64 device tensor.  Lists written with
        ``event_dtype="float32"`` are widened, and their sky positions -- stored as offsets from
        the fp64 sky centre -- are made absolute again."""
        t = to_device(lst.data[key])
        if key in ("xsky", "ysky") and int(lst.parameters.get("sky_offsets", 0)):
            t = t + float(np.asarray(lst.parameters["sky_center"], "float64")[0 if key == "xsky" else 1])
        return t

    # -- binning on the device ----------------------------------------------------------
    def pixel_coordinates(self, fov, nx):
        """(x, y) FITS pixel coordinates of every event for an ``nx`` x ``nx`` TAN image of field
        of view ``fov`` (arcmin) centred on the list's ``sky_center`` -- ``wcs_world2pix(...,
        1)`` of event_list.py:123."""
        dtheta = parse_value(fov, "arcmin") / 60.0 / nx
        sc = np.asarray(self.parameters["sky_center"], "float64")
        xs, ys = [], []
        for lst in self._lists:
            ra, dec = self._dev(lst, "xsky"), self._dev(lst, "ysky")
            px, py = torch.empty_like(ra), torch.empty_like(ra)
            _lib.call("pxb_cel_to_pixel", ra.numel(), ptr(ra), ptr(dec), float(sc[0]), float(sc[1]),
                      0.5 * (nx + 1), -dtheta, dtheta, ptr(px), ptr(py), stream_ptr())
            xs.append(px)
            ys.append(py)
        return torch.cat(xs), torch.cat(ys)

    def image(self, fov, nx, emin=None, emax=None):
        """Counts image [nx, nx] (row = y pixel: the ``H.T`` of event_list.py:295) as a device
        int64 tensor."""
        emin = -1.0 if emin is None else float(emin)          # event_list.py:273-276
        emax = 1.0e10 if emax is None else float(emax)
        dtheta = parse_value(fov, "arcmin") / 60.0 / nx
        sc = np.asarray(self.parameters["sky_center"], "float64")
        img = torch.zeros(nx * nx, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
        for lst in self._lists:
            ra, dec, e = self._dev(lst, "xsky"), self._dev(lst, "ysky"), self._dev(lst, "eobs")
            _lib.call("pxb_event_image", ra.numel(), ptr(ra), ptr(dec), ptr(e), emin, emax, float(sc[0]),
                      float(sc[1]), int(nx), dtheta, img.data_ptr(), stream_ptr())
        return img.reshape(nx, nx)

    def spectrum(self, emin, emax, nchan):
        """(ebins, counts[nchan]) with ``np.histogram(eobs, bins=np.linspace(emin, emax, nchan+1))``
        semantics (event_list.py:344-349); counts as a device int64 tensor."""
        ebins = np.linspace(emin, emax, nchan + 1, endpoint=True)
        cnt = torch.zeros(nchan, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
        de = to_device(ebins)
        for lst in self._lists:
            e = self._dev(lst, "eobs")
            _lib.call("pxb_event_spectrum", e.numel(), ptr(e), int(nchan), ptr(de), cnt.data_ptr(), stream_ptr())
        return ebins, cnt

    # -- writers -------------------------------------------------------------------------
    def write_fits_image(self, imagefile, fov, nx, emin=None, emax=None, overwrite=False):
        """event_list.py:254-318"""
        dtheta = parse_value(fov, "arcmin") / 60.0 / nx
        H = self.image(fov, nx, emin=emin, emax=emax).cpu().numpy().astype("float64")
        sc = self.parameters["sky_center"]
        hdr = [("MTYPE1", "EQPOS"), ("MFORM1", "RA,DEC"), ("CTYPE1", "RA---TAN"), ("CTYPE2", "DEC--TAN"),
               ("CRPIX1", 0.5 * (nx + 1)), ("CRPIX2", 0.5 * (nx + 1)), ("CRVAL1", float(sc[0])),
               ("CRVAL2", float(sc[1])), ("CUNIT1", "deg"), ("CUNIT2", "deg"), ("CDELT1", -dtheta),
               ("CDELT2", dtheta), ("EXPOSURE", float(self.parameters["exp_time"]))]
        _write_fits(imagefile, [_primary_hdu(H, hdr)], overwrite)

    def write_spectrum(self, specfile, emin, emax, nchan, overwrite=False):
        """event_list.py:320-385 (unconvolved spectrum, OGIP type-I PHA layout)"""
        ebins, cnt = self.spectrum(emin, emax, nchan)
        spec = cnt.cpu().numpy().astype("float64")
        emid = 0.5 * (ebins[1:] + ebins[:-1])
        t = float(self.parameters["exp_time"])
        cols = [("CHANNEL", "J", None, np.arange(nchan).astype("int32") + 1), ("ENERGY", "D", None, emid),
                ("COUNTS", "J", None, spec.astype("int32")), ("COUNT_RATE", "D", None, spec / t)]
        hdr = [("DETCHANS", int(nchan)), ("TOTCTS", float(spec.sum())), ("EXPOSURE", t), ("LIVETIME", t),
               ("CONTENT", "pi"), ("HDUCLASS", "OGIP"), ("HDUCLAS1", "SPECTRUM"), ("HDUCLAS2", "TOTAL"),
               ("HDUCLAS3", "TYPE:I"), ("HDUCLAS4", "COUNT"), ("HDUVERS", "1.1.0"), ("HDUVERS1", "1.1.0"),
               ("CHANTYPE", "pi"), ("BACKFILE", "none"), ("CORRFILE", "none"), ("POISSERR", True),
               ("RESPFILE", "none"), ("ANCRFILE", "none"), ("MISSION", "none"), ("TELESCOP", "none"),
               ("INSTRUME", "none"), ("AREASCAL", 1.0), ("CORRSCAL", 0.0), ("BACKSCAL", 1.0)]
        _write_fits(specfile, [_primary_hdu(), _bintable_hdu("SPECTRUM", cols, hdr)], overwrite)

    def write_fits_file(self, fitsfile, fov, nx, overwrite=False):
        """event_list.py:79-179: events inside the field of view as an OGIP EVENTS table
        (ENERGY [eV], X, Y [pixel])."""
        dtheta = parse_value(fov, "arcmin") / 60.0 / nx
        px, py = self.pixel_coordinates(fov, nx)
        keep = (px >= 0.5) & (px <= float(nx) + 0.5) & (py >= 0.5) & (py <= float(nx) + 0.5)
        e = torch.cat([self._dev(lst, "eobs") for lst in self._lists])
        n_keep = int(keep.sum())
        mylog.info("Threw out %d events because they fell outside the field of view.",
                   self.tot_num_events - n_keep)
        sc = self.parameters["sky_center"]
        t = float(self.parameters["exp_time"])
        cols = [("ENERGY", "E", "eV", (e[keep] * 1000.0).cpu().numpy()), ("X", "D", "pixel", px[keep].cpu().numpy()),
                ("Y", "D", "pixel", py[keep].cpu().numpy())]
        hdr = [("MTYPE1", "sky"), ("MFORM1", "x,y"), ("MTYPE2", "EQPOS"), ("MFORM2", "RA,DEC"),
               ("TCTYP2", "RA---TAN"), ("TCTYP3", "DEC--TAN"), ("TCRVL2", float(sc[0])), ("TCRVL3", float(sc[1])),
               ("TCDLT2", -dtheta), ("TCDLT3", dtheta), ("TCRPX2", 0.5 * (nx + 1)), ("TCRPX3", 0.5 * (nx + 1)),
               ("TLMIN2", 0.5), ("TLMIN3", 0.5), ("TLMAX2", float(nx) + 0.5), ("TLMAX3", float(nx) + 0.5),
               ("EXPOSURE", t), ("TSTART", 0.0), ("TSTOP", t), ("AREA", float(self.parameters["area"])),
               ("HDUVERS", "1.1.0"), ("RADECSYS", "FK5"), ("EQUINOX", 2000.0), ("HDUCLASS", "OGIP"),
               ("HDUCLAS1", "EVENTS"), ("HDUCLAS2", "ACCEPTED")]
        _write_fits(fitsfile, [_primary_hdu(), _bintable_hdu("EVENTS", cols, hdr)], overwrite)

    # -- region filters (event_list.py:387-432) --------------------------------------------
    def _filter(self, prefix, keep_condition):
        """Write, per input list, a list named ``f"{prefix}_{name}"`` holding the events for which
        ``keep_condition(xsky, ysky)`` (device tensors) is true, with flux / emin / emax recomputed
        (the reference assigns emax to the ``emin`` key, event_list.py:413-414 -- a typo; both keys
        are set properly here).  Returns the new names."""
        new_names = []
        for nm, lst in zip(self.filenames, self._lists):
            x, y = self._dev(lst, "xsky"), self._dev(lst, "ysky")
            keep = keep_condition(x, y)
            data = {k: to_device(v, dtype=v.dtype if isinstance(v, torch.Tensor) else torch.float64)[keep]
                    for k, v in lst.data.items()}
            params = dict(lst.parameters)
            e = data["eobs"]
            if e.numel() == 0:
                params.update(flux=0.0, emin=float("nan"), emax=float("nan"))
            else:
                params.update(flux=float(e.sum()) * 1.602176634e-9 / float(params["exp_time"]) / float(params["area"]),
                              emin=float(e.min()), emax=float(e.max()))
            if nm.startswith("mem:"):
                new = "mem:" + prefix + "_" + nm[4:]
            else:
                new = os.path.join(os.path.dirname(nm), prefix + "_" + os.path.basename(nm))
            info = dict(lst.info)
            save_list(new, DeviceList(info, params, data))
            new_names.append(new)
        return new_names

    def filter_circle(self, prefix, center, radius):
        """Events inside a circle of ``radius`` around ``center`` (both in the units of xsky/ysky,
        plain Euclidean distance as in event_list.py:417-423)."""
        return self._filter(prefix, lambda x, y: (x - center[0]) ** 2 + (y - center[1]) ** 2 <= radius * radius)

    def filter_rectangle(self, prefix, left_edge, right_edge):
        """Events with left_edge <= (xsky, ysky) < right_edge (event_list.py:425-431)."""
        return self._filter(prefix, lambda x, y: (x >= left_edge[0]) & (x < right_edge[0])
                            & (y >= left_edge[1]) & (y < right_edge[1]))

    def write_to_simput(self, prefix, overwrite=False):
        raise NotImplementedError("SIMPUT export is soxs' SimputCatalog (event_list.py:214-252); "
                                  "install soxs and use the reference EventList for it")


# ---------------------------------------------------------------------------------------
# merge_files (pyxsim/utils.py:71-173)
# ---------------------------------------------------------------------------------------
def validate_parameters(first, second, skip=()):
    """utils.py:63-89: same keys, strings equal, numbers equal to 1e-10 absolute.  (The reference
    reads both values from ``first`` -- ``v2 = first[k2][()]``, utils.py:75 -- so its value check
    can never fail; the intended comparison is implemented here.)"""
    k1, k2 = sorted(first), sorted(second)
    if k1 != k2:
        raise RuntimeError("The two inputs do not have the same parameters!")
    for k in k1:
        if k in skip:
            continue
        v1, v2 = first[k], second[k]
        if isinstance(v1, (str, bytes)) or isinstance(v2, (str, bytes)):
            ok = _as_str(v1) == _as_str(v2)
        elif np.shape(v1) != np.shape(v2):
            ok = False
        elif np.asarray(v1).dtype.kind in "SUO" or np.asarray(v2).dtype.kind in "SUO":
            # string arrays (velocity_fields): utils.py:79-83 compares them with np.char.equal
            ok = np.array_equal(np.asarray(v1).astype("U"), np.asarray(v2).astype("U"))
        else:
            ok = np.allclose(np.array(v1, "float64"), np.array(v2, "float64"), rtol=0.0, atol=1.0e-10)
        if not ok:
            raise RuntimeError(f"The values for the parameter '{k}' in the two inputs are not identical "
                               f"({v1} vs. {v2})!")


def _as_str(v):
    return v.decode() if isinstance(v, bytes) else str(v)


# per-list summaries of an event list (photon_list.py:816-820): they differ between any two
# real lists, so they are not compared but recomputed for the merged list.  (In the reference the
# value check can never fire -- see validate_parameters -- so merging such lists works there and
# the merged file silently keeps the first list's flux/emin/emax.)
_DERIVED = ("flux", "emin", "emax")


def merge_files(input_files, output_file, overwrite=False, add_exposure_times=False):
    """Merge photon or event lists (utils.py:92-173): every parameter except the exposure time
    (and an event list's derived flux/emin/emax, which are recomputed) must agree; exposure times
    are added (``add_exposure_times``) or their maximum is kept; the ``/data`` arrays are
    concatenated in input order (on the device for ``"mem:"`` lists)."""
    out_base = _strip(output_file)
    if not out_base.startswith("mem:") and not overwrite and \
            (os.path.exists(out_base + ".h5") or os.path.exists(out_base + ".npz")):
        raise OSError(f"Cannot overwrite existing file {output_file}. If you want to do this, "
                      "set overwrite=True.")
    lists = [load_list(f if f.startswith("mem:") else _strip(f)) for f in input_files]
    first = lists[0]
    exp_time_key = next((k for k in first.parameters if k.endswith("exp_time")), "")
    skip = ([exp_time_key] if add_exposure_times else []) + list(_DERIVED)
    for lst in lists[1:]:
        validate_parameters(first.parameters, lst.parameters, skip=skip)
    params = {k: v for k, v in first.parameters.items() if k != exp_time_key}
    tot = 0.0
    info = {}
    for i, lst in enumerate(lists):
        t = float(lst.parameters[exp_time_key])
        tot = tot + t if add_exposure_times else max(tot, t)
        for k, v in lst.info.items():
            info[f"{k}_{i}"] = v
    info["original_files"] = list(input_files)
    params[exp_time_key] = tot
    data = {}
    for k in first.data:
        parts = [lst.data[k] for lst in lists]
        if all(isinstance(p, torch.Tensor) for p in parts):
            data[k] = torch.cat(parts)
        else:
            data[k] = np.concatenate([p.cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p)
                                      for p in parts])
    if "eobs" in data and all(k in params for k in _DERIVED):
        e = data["eobs"]
        n = int(e.shape[0])
        if n == 0:
            params.update(flux=0.0, emin=1.0e99, emax=-1.0e99)
        else:
            area = float(params["area"])
            params.update(flux=float(e.sum()) * 1.602176634e-9 / tot / area, emin=float(e.min()),
                          emax=float(e.max()))
    return save_list(out_base, DeviceList(info, params, data))
