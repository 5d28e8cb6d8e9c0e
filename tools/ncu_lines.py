#!/usr/bin/env python
"""Per-source-line share of warp-stall samples and executed instructions of one kernel in an .ncu-rep captured with
--import-source on.  Usage: python tools/ncu_lines.py rep.ncu-rep kernel_substring [top_n]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg, cur_file, hdr, active = {}, None, None, False
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            active = kern in r[1]
            continue
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            continue
        if active and hdr and len(r) == len(hdr) and r[0]:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            s = int(r[hdr.index("# Samples")] or 0)
            i = int(r[hdr.index("Instructions Executed")] or 0)
            k = (cur_file, ln)
            a = agg.get(k, (0, 0, r[1][:100]))
            agg[k] = (a[0] + s, a[1] + i, a[2])
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print("kernel %s: %d samples, %d warp instructions" % (kern, ts, ti))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-18s %4d  samples %5.1f%%  inst %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / ts, 100 * v[1] / ti, v[2]))


if __name__ == "__main__":
    main()
