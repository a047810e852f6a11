#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.

  python profiles/summarize.py launches gpurun_out/<launch list>.csv  > profiles/<name>_launches.md
  python profiles/summarize.py rep gpurun_out/<report>.ncu-rep        > profiles/<name>_kernels.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__cycles_active.avg", "sm__cycles_active.avg"]


def read_csv(text):
    lines = text.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def launches(path):
    rows = read_csv(open(path).read())
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0][-70:]
        v = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# launch list: {path}\n\n{len(rows)} launches, {tot / 1e6:.3f} ms total (ncu per-launch times are cold-cache and serialised: read SHARES)\n")
    print("| kernel | launches | total us | avg us | share | grid | block |\n|---|---|---|---|---|---|---|")
    for n, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1e3:.1f} | {t / 1e3 / c:.1f} | {100 * t / tot:.1f}% | {g} | {b} |")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = read_csv(out)
    print(f"# ncu --set full: {path}\n")
    units = rows[0] if rows and rows[0].get('ID', '') == '' else {}
    for r in rows:
        if r.get('ID', '') == '':
            continue
        print(f"## {r['Kernel Name'].split('(')[0]}  (id {r['ID']}, grid {r.get('Grid Size')}, block {r.get('Block Size')})\n")
        print("| metric | value |\n|---|---|")
        for k in KEYS:
            if k in r and r[k] != "":
                print(f"| {k} | {r[k]} {units.get(k, '')} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2])
