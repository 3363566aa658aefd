"""Turn the ncu artefacts a gpurun call brought back into the tracked summaries under profiles/.

    python scripts/summarise_profile.py --tag r1_final \
        --launches gpurun_out/launches.csv --report gpurun_out/prof.ncu-rep \
        --command "python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"

Writes profiles/<tag>_launches.csv (raw launch list), <tag>_launch_shares.txt (share of libpxb
kernel time per kernel), <tag>_traffic.json (measured DRAM bytes per launch, read by bench.py
for `roofline.traffic`) and <tag>_ncu_summary.txt (per-kernel metrics, stall samples and
source-level hot spots).  Needs the `ncu` CLI (reads reports only; no GPU).
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_sample_count"]


def ncu_csv(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], capture_output=True,
                         text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * scale[unit]


def to_ms(value, unit):
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
             "second": 1e3}
    return float(value) * scale[unit]


def launch_shares(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or r[ki] == "Kernel Name":
            continue
        name = r[ki].split("(")[0]
        if not name.replace("void ", "").startswith("k_"):
            continue   # torch kernels of the benchmark's input synthesis: not part of the step
        try:
            us = to_ms(r[vi].replace(",", ""), r[ui]) * 1e3
        except (ValueError, KeyError):
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    return agg


def source_hot_spots(report, index, topn):
    """Source lines of the index-th launch of the report by executed instructions."""
    rows = ncu_csv(report, "source", ["--print-source", "cuda,sass", "--launch-skip", str(index),
                                      "--launch-count", "1"])
    cur, hdr, si, ii = None, None, 0, 0
    agg = collections.defaultdict(lambda: [0, 0, ""])
    tot_i = tot_s = 0
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= ii:
            continue
        try:
            ln = int(r[0])
            s, n = int(float(r[si] or 0)), int(float(r[ii] or 0))
        except ValueError:
            continue
        a = agg[(cur, ln)]
        a[0] += s
        a[1] += n
        a[2] = r[1].strip()[:95]
        tot_i += n
        tot_s += s
    lines = []
    for (f, ln), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        lines.append(f"{f:12s}:{ln:4d} inst {100 * n / max(tot_i, 1):5.1f}% samp {100 * s / max(tot_s, 1):5.1f}%  {src}")
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--report")
    ap.add_argument("--command", default="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline")
    ap.add_argument("--workload", default='{"grid": 256, "photons": 1e9, "nbins": 4000}')
    ap.add_argument("--top", type=int, default=16)
    ap.add_argument("--outdir", default=os.path.join(ROOT, "profiles"),
                    help="where the summaries go (gpurun_out/ when run on the GPU box)")
    args = ap.parse_args()
    prof = args.outdir
    if args.launches:
        shutil.copy(args.launches, os.path.join(prof, f"{args.tag}_launches.csv"))
        agg = launch_shares(args.launches)
        total = sum(v[1] for v in agg.values())
        with open(os.path.join(prof, f"{args.tag}_launch_shares.txt"), "w") as f:
            f.write(f"# launch list of `{args.command}`\n"
                    "# ncu --metrics gpu__time_duration.sum --clock-control none "
                    f"(raw: profiles/{args.tag}_launches.csv)\n"
                    "# share of libpxb kernel time per kernel (cold-cache, serialised timings: compare "
                    "shares, not absolutes)\n")
            for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{us / n:12.1f} us/launch  {100 * us / total:6.2f}%  x{n:3d}  {name}\n")
    if args.report:
        rows = ncu_csv(args.report, "raw")
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued")]
        traffic = {"workload": json.loads(args.workload), "command": args.command, "kernels": {},
                   "source": f"ncu --set full --clock-control none, one launch each; summary in "
                             f"profiles/{args.tag}_ncu_summary.txt"}
        with open(os.path.join(prof, f"{args.tag}_ncu_summary.txt"), "w") as f:
            f.write(f"# {args.command}\n# ncu --set full --clock-control none --import-source on (one launch per kernel)\n\n")
            names = []
            for r in rows[2:]:
                name = r[ki].split("(")[0].replace("void ", "")
                names.append(name)
                f.write(f"kernel: {r[ki][:70]}\n")
                for w in WANT + stalls:
                    if w in hdr:
                        i = hdr.index(w)
                        if r[i] not in ("", "0"):
                            f.write(f"  {w:70s} {r[i][:30]:>22s} {units[i]}\n")
                f.write("\n")
                rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
                wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
                ti = hdr.index("gpu__time_duration.sum")
                traffic["kernels"][name] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                                            "ms": to_ms(r[ti], units[ti])}
            for j, name in enumerate(names):
                f.write(f"## source-level hot spots, {name}\n")
                f.write("\n".join(source_hot_spots(args.report, j, args.top)) + "\n\n")
        with open(os.path.join(prof, f"{args.tag}_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print(json.dumps(traffic["kernels"], indent=1))


if __name__ == "__main__":
    main()
