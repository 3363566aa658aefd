"""Per-kernel summary + per-source-line / per-region instruction shares of an ncu report
(`ncu --set full --import-source on`).  Usage: python scripts/prof_regions.py REPORT [--lines]"""
import collections
import csv
import io
import subprocess
import sys


def ncu(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct']


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr = rows[0]
    names = []
    for r in rows[2:]:
        nm = r[hdr.index('Kernel Name')].split('(')[0]
        names.append(nm)
        print("kernel:", nm)
        for w in WANT:
            if w in hdr:
                print(f"  {w:62s} {r[hdr.index(w)][:24]:>24s} {rows[1][hdr.index(w)]}")
        st = {h: int(float(r[hdr.index(h)] or 0)) for h in hdr
              if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('_not_issued')}
        tot = sum(st.values()) or 1
        print('  stalls:', ', '.join(f"{k.split('stalled_')[1]}={100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    src = ncu(rep, "source", ["--print-source", "cuda,sass"])
    kern, cur, fname = [], None, None
    for r in src:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            if fname == 'common.cuh':
                cur = collections.defaultdict(lambda: [0, 0, ''])
                kern.append(cur)
            continue
        if r[0] == 'Line No':
            si, ii = r.index('# Samples'), r.index('Instructions Executed')
            continue
        try:
            ln, s, n = int(r[0]), int(float(r[si] or 0)), int(float(r[ii] or 0))
        except (ValueError, IndexError):
            continue
        a = cur[(fname, ln)]
        a[0] += s
        a[1] += n
        a[2] = r[1].strip()[:100]
    for nm, k in zip(names, kern):
        ti = sum(v[1] for v in k.values()) or 1
        ts = sum(v[0] for v in k.values()) or 1
        print(f"== {nm}: {ti:.3e} warp-instr")
        top = 200 if "--lines" in sys.argv else 24
        for (f, ln), (s, n, srcl) in sorted(k.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"  {f:14s}:{ln:4d} inst {100 * n / ti:5.1f}% ({n * 32 / 1e9:5.1f}) samp {100 * s / ts:5.1f}%  {srcl}")


if __name__ == "__main__":
    main()
