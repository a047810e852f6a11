#!/usr/bin/env python
"""Per-instruction view of one kernel of an ncu --set full report: executed-instruction mix by opcode, warp-stall samples
by reason, and the most-sampled SASS instructions with their dominant stall reason.

    ncu -i gpurun_out/<report>.ncu-rep --page source --csv --kernel-name regex:<kernel> > /tmp/k.csv
    python profiles/hotspots.py /tmp/k.csv > profiles/<name>_hotspots.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_e = tot_s = 0
    byop, samp = collections.Counter(), collections.Counter()
    reasons = collections.Counter()
    lines = []
    for i, r in enumerate(data):
        try:
            e, sm = int(r[col["Instructions Executed"]]), int(r[col["# Samples"]])
        except (ValueError, IndexError):
            continue
        src = r[col["Source"]].strip()
        toks = src.split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += e
        samp[op] += sm
        tot_e += e
        tot_s += sm
        st = {}
        for h in stall_cols:
            try:
                st[h] = int(r[col[h]])
            except ValueError:
                st[h] = 0
            reasons[h] += st[h]
        lines.append((sm, i, src, e, max(st, key=st.get) if sm else "-"))
    print(f"# {kernel[:110]}\n")
    print(f"{tot_e} warp-instructions executed, {tot_s} warp-stall samples\n")
    print("## executed instructions by opcode (share of executed, share of samples)\n```")
    for op, c in byop.most_common(22):
        print(f"{op:12s} {100 * c / tot_e:5.1f}%   {100 * samp[op] / tot_s:5.1f}%")
    print("```\n## stall samples by reason\n```")
    tot_r = sum(reasons.values())
    for h, c in reasons.most_common(12):
        print(f"{h:24s} {100 * c / tot_r:5.1f}%")
    print("```\n## most-sampled instructions (share of samples, executions, dominant reason)\n```")
    lines.sort(reverse=True)
    for sm, i, src, e, why in lines[:30]:
        print(f"{100 * sm / tot_s:4.1f}%  {e:10d}  {why:20s} {src[:100]}")
    print("```")


if __name__ == "__main__":
    main(sys.argv[1])
