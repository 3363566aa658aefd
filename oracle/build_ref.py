"""Build the reference's own native layer into ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

The three Cython modules of the reference (``pyxsim/lib/interpolate.pyx``,
``sky_functions.pyx``, ``spectra.pyx``; build recipe mirrored from the reference's
``setup.py:14-35``: language C, libm, numpy include dir) are cythonized *where they
lie* under ``/root/reference`` and compiled with gcc.  Nothing from the reference tree
is copied into the repository: generated ``.c`` files and the ``.so`` files land in
``oracle/_ref/`` which is git-ignored (but travels to the GPU box with the snapshot).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
use the result; the product package (``pyxsim_b200``) never imports it.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_LIB = os.environ.get("PYXSIM_REFERENCE", "/root/reference") + "/pyxsim/lib"
MODULES = ("interpolate", "sky_functions", "spectra")


def so_path(name):
    return os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))


def have_ref():
    return all(os.path.exists(so_path(m)) for m in MODULES)


def build(force=False, verbose=True):
    """Compile the reference Cython sources. Returns True when all modules exist."""
    if have_ref() and not force:
        return True
    if not os.path.isdir(REF_LIB):
        if verbose:
            print(f"[oracle] reference tree {REF_LIB} absent; using prebuilt files only")
        return have_ref()
    import numpy as np

    os.makedirs(OUT, exist_ok=True)
    inc_py = sysconfig.get_paths()["include"]
    inc_np = np.get_include()
    for m in MODULES:
        pyx = os.path.join(REF_LIB, m + ".pyx")
        c_file = os.path.join(OUT, m + ".c")
        subprocess.check_call(
            [sys.executable, "-m", "cython", "-3", pyx, "-o", c_file],
            cwd=OUT,
        )
        cmd = [
            "gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-w",
            "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
            f"-I{inc_py}", f"-I{inc_np}", c_file, "-o", so_path(m), "-lm",
        ]
        subprocess.check_call(cmd)
        if verbose:
            print(f"[oracle] built {so_path(m)}")
    return have_ref()


def load():
    """Import the compiled reference modules; returns a dict name -> module."""
    import importlib.util

    if not have_ref():
        raise ImportError(
            "oracle/_ref is not built: run `python oracle/build_ref.py` in the container "
            "that has /root/reference"
        )
    mods = {}
    for m in MODULES:
        spec = importlib.util.spec_from_file_location(m, so_path(m))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[m] = mod
    return mods


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref ready" if ok else "oracle/_ref NOT available")
    sys.exit(0 if ok else 1)
