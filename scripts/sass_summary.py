"""Static instruction mix of libpxb's kernels from `cuobjdump -sass` (no GPU needed), and full
listings of the four kernels of the headline path.

    python scripts/sass_summary.py            # writes profiles/r2_sass_summary.txt + r2_sass_<kernel>.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyxsim_b200", "libpxb.so")
PROF = os.path.join(ROOT, "profiles")
KEEP = ("ATOMG", "ATOMS", "BAR", "DADD", "DFMA", "DMUL", "IMAD", "LDG", "LDL", "LDS", "MUFU", "NANOSLEEP", "REDUX",
        "SHFL", "STG", "STL", "STS", "SYNCS", "UBLKCP", "UTMALDG", "VOTE", "VOTEU", "HMMA", "UTCHMMA", "F2I", "I2F")
LISTINGS = {"k_pipeline<0, 1, 32>": "sample_common", "k_pipeline<1, 1, 8>": "project_common",
            "k_pipeline<2, 1, 8>": "fused_common", "k_prep_sample": "prep_sample",
            "k_spec_bucket_run<1, 2>": "spectrum_bucket_run"}


def demangle(names):
    out = subprocess.run(["cu++filt", *names], capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"^void ", "", o.replace("(int)", "").replace("(bool)", "").split("(")[0]) for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif line.startswith("Fatbin ") or line.startswith("\t.section"):
            cur = None
        elif cur is not None:
            funcs[cur].append(line)
    names = dict(zip(funcs, demangle(list(funcs))))
    rows = []
    for f, lines in funcs.items():
        ops = collections.Counter()
        n = 0
        for ln in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
            if m:
                n += 1
                ops[m.group(1)] += 1
        rows.append((names[f], n, ops, f, lines))
    rows.sort(key=lambda r: r[0])
    with open(os.path.join(PROF, "r2_sass_summary.txt"), "w") as o:
        o.write("# python scripts/sass_summary.py: cuobjdump -sass pyxsim_b200/libpxb.so, static instruction mix per kernel\n"
                "# UBLKCP = cp.async.bulk (TMA bulk copy of the alias rows), SYNCS = mbarrier ops, REDUX = warp reductions\n"
                "# (reduce_or / reduce_add / min / max), LDL/STL = register spills; no UTMALDG (1-D bulk copies need no\n"
                "# tensor map) and no UTC*MMA / HMMA: the path is gather / sample / project, not a contraction.\n\n")
        for name, n, ops, _, _ in rows:
            mix = ", ".join(f"{k} {ops[k]}" for k in KEEP if ops.get(k))
            o.write(f"{name}\n    {n} SASS instructions; {mix}\n")
    for name, n, ops, f, lines in rows:
        tag = LISTINGS.get(name)
        if tag:
            with open(os.path.join(PROF, f"r2_sass_{tag}.txt"), "w") as o:
                o.write(f"# cuobjdump -sass pyxsim_b200/libpxb.so, function {name} (sm_100a)\n{f}\n")
                for ln in lines:          # instruction lines only, without the encodings
                    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                    if m:
                        o.write(f"/*{m.group(1)}*/ {m.group(2).strip()} ;\n")
                    elif ln.strip().startswith(".") and ":" in ln:
                        o.write(ln.strip() + "\n")
    print(len(rows), "kernels")


if __name__ == "__main__":
    main()
