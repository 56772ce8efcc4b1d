#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples from an .ncu-rep (compiled with -lineinfo).
usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    fname = ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue   # keep only the per-source-line aggregate rows (Address == "-")
        d = dict(zip(hdr, r))
        try:
            inst = int(d["Instructions Executed"])
            samp = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        lines.append((inst, samp, fname, r[0], r[1].strip()[:110], d.get("L1 Wavefronts Shared Excessive", "0")))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for inst, samp, f, ln, src, exc in sorted(lines, key=lambda l: -l[0])[:top]:
        print("%5.1f%% inst %5.1f%% stall  %s:%s  %s   [smem excess wavefronts %s]" % (100.0 * inst / tot_i, 100.0 * samp / tot_s, f, ln, src, exc))


if __name__ == "__main__":
    main()
