#!/usr/bin/env python3
"""Top stall-sample locations of an `ncu --page source --csv --print-source sass` export. usage: ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
tot = sum(int(r[ismp]) for r in rows[2:])
top = sorted(((int(r[ismp]), int(r[ia], 16) - base, r[isrc].strip()[:70], r[iex]) for r in rows[2:]), reverse=True)
for s, off, src, ex in top[:N]:
    print("%5.1f%%  %05x  exec %9s  %s" % (100.0 * s / tot, off, ex, src))
