"""Per-launch instruction shares of an ncu report (`ncu --set full --import-source on`), by
source line and by enclosing function / lambda ("region").

    python scripts/prof_regions.py REPORT [--lines N] [--only SUBSTRING]

For every launch in the report: executed warp instructions, thread instructions per unit of work
when `--per N` gives the units (e.g. photons), the regions ranked by instructions, and the top
source lines.  Regions come from the repository's own sources: a line belongs to the last
`__device__` / `__global__` function or `auto name = [&]` lambda that starts before it.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pyxsim_b200", "csrc")
START = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;]*?\b(\w+)\s*\($")
START2 = re.compile(r"(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(")
LAMBDA = re.compile(r"^\s*auto\s+(\w+)\s*=\s*\[&\]")


def ncu(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def region_map(fname):
    """line number -> name of the enclosing function / lambda of a csrc file"""
    path = os.path.join(CSRC, fname)
    if not os.path.exists(path):
        return {}
    out, func, lam, lam_indent, want_name = {}, "(file scope)", None, None, False
    for n, line in enumerate(open(path), 1):
        stripped = line.strip()
        if want_name:                      # kernel name on the line after __launch_bounds__(...)
            m = re.match(r"^(\w+)\s*\(", stripped)
            if m:
                func, want_name = m.group(1), False
        m = LAMBDA.match(line)
        if m:
            lam, lam_indent = m.group(1), len(line) - len(line.lstrip())
        elif lam is not None and line.rstrip() == " " * lam_indent + "};":
            out[n] = f"{func}::{lam}"
            lam = None
            continue
        elif ("__device__" in line or "__global__" in line) and not stripped.startswith("//"):
            m = START2.search(line)
            if m:
                if m.group(1) == "__launch_bounds__":
                    want_name = True
                else:
                    func = m.group(1)
                lam = None
        out[n] = f"{func}::{lam}" if lam else func
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--lines", type=int, default=14)
    ap.add_argument("--only", default="")
    ap.add_argument("--per", type=float, default=0.0, help="units of work per launch (e.g. photons)")
    args = ap.parse_args()
    raw = ncu(args.report, "raw")
    hdr = raw[0]
    names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in raw[2:]]
    maps = {}
    for j, name in enumerate(names):
        if args.only and args.only not in name:
            continue
        src = ncu(args.report, "source", ["--print-source", "cuda,sass", "--launch-skip", str(j),
                                          "--launch-count", "1"])
        fname, si, ii, ti_ = None, 0, 0, 0
        lines = collections.defaultdict(lambda: [0, 0, 0, ""])
        for r in src:
            if not r:
                continue
            if r[0] in ("File Path", "File Name"):
                fname = r[1].split("/")[-1]
                continue
            if r[0] == "Line No":
                si, ii = r.index("# Samples"), r.index("Instructions Executed")
                ti_ = r.index("Thread Instructions Executed")
                continue
            try:
                ln, s, n, t = int(r[0]), int(float(r[si] or 0)), int(float(r[ii] or 0)), int(float(r[ti_] or 0))
            except (ValueError, IndexError):
                continue
            a = lines[(fname, ln)]
            a[0] += s
            a[1] += n
            a[2] += t
            a[3] = r[1].strip()[:96]
        tot_i = sum(v[1] for v in lines.values()) or 1
        tot_s = sum(v[0] for v in lines.values()) or 1
        tot_t = sum(v[2] for v in lines.values())
        print(f"== launch {j}: {name}: {tot_i:.3e} warp instructions, {tot_t:.3e} thread instructions"
              + (f" = {tot_t / args.per:.0f} per unit" if args.per else ""))
        regions = collections.defaultdict(lambda: [0, 0, 0])
        for (f, ln), (s, n, t, _) in lines.items():
            if f not in maps:
                maps[f] = region_map(f)
            reg = f"{f}:{maps[f].get(ln, '?')}"
            regions[reg][0] += s
            regions[reg][1] += n
            regions[reg][2] += t
        print("  -- regions (enclosing function / lambda)")
        for reg, (s, n, t) in sorted(regions.items(), key=lambda kv: -kv[1][1])[:18]:
            per = f" {t / args.per:6.1f}/unit" if args.per else ""
            print(f"  {reg:52s} inst {100 * n / tot_i:5.1f}%{per}  samples {100 * s / tot_s:5.1f}%")
        print("  -- source lines")
        for (f, ln), (s, n, t, text) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:args.lines]:
            print(f"  {f:14s}:{ln:4d} inst {100 * n / tot_i:5.1f}% samples {100 * s / tot_s:5.1f}%  {text}")
        print()


if __name__ == "__main__":
    main()
