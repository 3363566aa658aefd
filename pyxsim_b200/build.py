"""Build libpxb.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

``python -m pyxsim_b200.build`` compiles every ``csrc/*.cu`` for
``-gencode arch=compute_100a,code=sm_100a`` (cross-compiles without a GPU) and links
``pyxsim_b200/libpxb.so`` against the static CUDA runtime, so the library only needs the
driver at run time.  Objects are cached by source mtime under ``pyxsim_b200/csrc/_obj``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpxb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs.append(os.path.join(HERE, "..", "include", "pxb.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose, objdir=OBJ, extra=()):
    obj = os.path.join(objdir, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    newest = max(os.path.getmtime(spath), _headers_mtime())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [NVCC, *ARCH, *FLAGS, *extra, "-c", spath, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr


def build(force=False, verbose=False, debug_checks=False, variant=None, defines=()):
    """debug_checks: the same library with the kernels' index assertions compiled in
    (-DPXB_DEBUG_CHECKS), as pyxsim_b200/libpxb_dbg.so; load it with PXB_LIBRARY=<path>.
    variant / defines: an A/B build of the same sources with extra -D flags, as
    pyxsim_b200/libpxb_<variant>.so (``--variant r2 -DPXB_PROJ_ROUNDS=2``)."""
    if debug_checks:
        variant, defines = "dbg", ("-DPXB_DEBUG_CHECKS",) + tuple(defines)
    objdir = OBJ + "_" + variant if variant else OBJ
    lib = LIB.replace("libpxb.so", f"libpxb_{variant}.so") if variant else LIB
    extra = tuple(defines)
    os.makedirs(objdir, exist_ok=True)
    if force:
        for f in os.listdir(objdir):
            os.remove(os.path.join(objdir, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose, objdir, extra), srcs))
    objs = [r[0] for r in results]
    log = "".join(r[1] for r in results)
    if log:
        with open(os.path.join(objdir, "ptxas.log"), "a") as f:
            f.write(log)
        if verbose:
            print(log)
    if (not os.path.exists(lib)) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", lib, *objs, "-cudart", "static"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return lib


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                debug_checks="--debug-checks" in sys.argv, variant=_variant,
                defines=[a for a in sys.argv if a.startswith("-D")]))
