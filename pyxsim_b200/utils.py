"""Host-side helpers used by the hot path (mirrors the used parts of pyxsim/utils.py)."""
import logging
import numbers

import numpy as np

from . import constants as K

# pyxsim/utils.py:10-11
scale_shift = -1.0 / K.clight_kms
scale_shift2 = scale_shift * scale_shift

mylog = logging.getLogger("pyxsim_b200")
if not mylog.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(name)-3s : [%(levelname)-9s] %(asctime)s %(message)s"))
    mylog.addHandler(_h)
    mylog.setLevel("INFO")
    mylog.propagate = False


def parse_prng(prng):
    """int / None / RandomState -> RandomState-like (soxs.utils.parse_prng, used at
    photon_list.py:502 and sources.py:24)."""
    if isinstance(prng, np.random.RandomState):
        return prng
    if isinstance(prng, numbers.Integral):
        return np.random.RandomState(int(prng))
    if prng is None:
        return np.random
    return prng


def prng_seed(prng):
    """64-bit Philox key for the device generators.

    The reference threads one Mersenne-Twister stream through data-dependent draws, which
    cannot be reproduced on a GPU; instead the user's ``prng`` argument only
    *seeds* the counter-based generator: an int is used as is, a RandomState contributes one
    draw (so repeated calls with the same state object give different photons, like the
    reference), None draws from OS entropy.
    """
    if isinstance(prng, numbers.Integral):
        return int(prng) & 0xFFFFFFFFFFFFFFFF
    if prng is None or prng is np.random:
        return int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
    lo = int(prng.randint(0, 2**32, dtype=np.uint64))
    hi = int(prng.randint(0, 2**32, dtype=np.uint64))
    return (hi << 32) | lo


# lengths in kpc so that kpc <-> Mpc conversions are exact
_LENGTH = {"cm": 1.0 / K.cm_per_kpc, "m": 1e2 / K.cm_per_kpc, "km": 1e5 / K.cm_per_kpc,
           "pc": 1e-3, "kpc": 1.0, "Mpc": 1.0e3}
_TIME = {"s": 1.0, "ks": 1e3, "Ms": 1e6, "yr": 3.15576e7, "kyr": 3.15576e10}
_AREA = {"cm**2": 1.0, "cm^2": 1.0, "cm2": 1.0, "m**2": 1e4, "m^2": 1e4}
_ENERGY = {"keV": 1.0, "eV": 1e-3, "MeV": 1e3}
_ANGLE = {"arcsec": 1.0 / 60.0, "arcmin": 1.0, "deg": 60.0, "rad": 60.0 * 180.0 / np.pi}
_FAMILIES = [_LENGTH, _TIME, _AREA, _ENERGY, _ANGLE]


def parse_value(value, default_units):
    """Number | (value, unit) | object with ``to_value`` -> float in ``default_units``
    (pyxsim/utils.py:30-50, without the unyt dependency)."""
    if hasattr(value, "to_value"):
        return float(value.to_value(default_units))
    if isinstance(value, (tuple, list)) and len(value) == 2 and isinstance(value[1], str):
        v, u = value
        for fam in _FAMILIES:
            if u in fam and default_units in fam:
                return float(v) * fam[u] / fam[default_units]
        raise ValueError(f"cannot convert '{u}' to '{default_units}'")
    return float(value)


def orientation(normal, north_vector=None):
    """Unit vectors [east, north, normal] of yt's ``Orientation`` (used through
    pyxsim/utils.py:261-270).  yt is third-party and not vendored in the reference tree: this
    restates its published rule -- normalise L; without a north vector pick the identity axis
    maximising sum(cross(L, I)), east = cross(I[ax], L), north = cross(L, east); with one,
    remove its component along L and set east = cross(north, L)."""
    L = np.array(normal, dtype="float64")
    L = L / np.sqrt(np.dot(L, L))
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
    north = north / np.sqrt(np.dot(north, north))
    east = east / np.sqrt(np.dot(east, east))
    return np.array([east, north, L])


def get_normal_and_north(normal, north_vector=None):
    """(L, north_vector, unit_vectors) as pyxsim/utils.py:261-270."""
    if not isinstance(normal, str):
        L = np.array(normal, dtype="float64")
    else:
        L = np.zeros(3)
        L["xyz".index(normal)] = 1.0
    uv = orientation(L, north_vector=north_vector)
    return L, uv[1], uv
