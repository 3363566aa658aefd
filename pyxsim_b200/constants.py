"""Physical constants and abundance data used on the host side of the hot path.

In the reference these come from third-party packages that are not vendored in its tree:
``soxs.constants`` (``K_per_keV``, ``erg_per_keV``, ``atomic_weights``, ``elem_names``,
``abund_tables``; pyxsim/utils.py:6, spectral_models.py:7, photon_list.py:9) and ``unyt``
(``clight``, kpc; pyxsim/utils.py:7-11, source_models/sources.py:15).  When soxs is
importable its values are used; otherwise the CODATA-2018 values below (soxs version is
unpinned in the reference -- pyproject.toml:30 -- so last-digit agreement is unverified).
"""
import numpy as np

K_per_keV = 1.160451812e7          # 1 keV / k_B in K
erg_per_keV = 1.602176634e-9
clight_kms = 299792.458
cm_per_kpc = 3.0856775809623245e21
cm_per_mpc = cm_per_kpc * 1.0e3
cm2_per_kpc2 = cm_per_kpc * cm_per_kpc

elem_names = ["", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al",
              "Si", "P", "S", "Cl", "Ar", "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co",
              "Ni", "Cu", "Zn"]

# standard atomic weights (u), index = atomic number
atomic_weights = np.array([
    0.0, 1.00794, 4.00262, 6.941, 9.012182, 10.811, 12.0107, 14.0067, 15.9994, 18.9984,
    20.1797, 22.9898, 24.3050, 26.9815, 28.0855, 30.9738, 32.0650, 35.4530, 39.9480,
    39.0983, 40.0780, 44.9559, 47.8670, 50.9415, 51.9961, 54.9380, 55.8450, 58.9332,
    58.6934, 63.5460, 65.3800])

# Anders & Grevesse (1989) number abundances relative to H, index = atomic number
_angr = np.array([
    0.0, 1.00e+00, 9.77e-02, 1.45e-11, 1.41e-11, 3.98e-10, 3.63e-04, 1.12e-04, 8.51e-04,
    3.63e-08, 1.23e-04, 2.14e-06, 3.80e-05, 2.95e-06, 3.55e-05, 2.82e-07, 1.62e-05,
    3.16e-07, 3.63e-06, 1.32e-07, 2.29e-06, 1.26e-09, 9.77e-08, 1.00e-08, 4.68e-07,
    2.45e-07, 4.68e-05, 8.32e-08, 1.78e-06, 1.62e-08, 3.98e-08])
abund_tables = {"angr": _angr}

try:  # prefer the reference's own third-party source when it exists
    from soxs import constants as _sc  # type: ignore

    K_per_keV = float(_sc.K_per_keV)
    erg_per_keV = float(_sc.erg_per_keV)
    atomic_weights = np.asarray(_sc.atomic_weights, dtype="float64")
    elem_names = list(_sc.elem_names)
    abund_tables = dict(_sc.abund_tables)
except Exception:  # pragma: no cover - soxs is not installed in the build image
    pass

metal_elem = [z for z in range(3, 31)]


def compute_H_abund(abund_table):
    """X_H = A_H / sum(A * atable)  (pyxsim/utils.py:236-238)."""
    atable = abund_tables[abund_table] if isinstance(abund_table, str) else np.asarray(abund_table)
    return atomic_weights[1] / (atomic_weights * atable).sum()


def compute_zsolar(abund_table):
    """Solar metal mass fraction: the inverse of Zconvert in base.py:173-175 times X_H."""
    atable = abund_tables[abund_table] if isinstance(abund_table, str) else np.asarray(abund_table)
    if atable[1] != 1.0:
        atable = atable / atable[1]
    return (atomic_weights[3:] * atable[3:]).sum() / (atomic_weights * atable).sum()
