#!/usr/bin/env python3
"""Sum executed warp instructions of an `ncu --page source --csv --print-source sass` export between barriers / address marks.
usage: ncu_phase.py file.csv [hexoffset ...]  - splits the kernel at BAR.SYNC and at the given offsets (from the first address)"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
marks = sorted(int(x, 16) for x in sys.argv[2:])
phase, tot, smp, ops = 0, collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in rows[2:]:
    off = int(r[ia], 16) - base
    while marks and off >= marks[0]:
        marks.pop(0); phase += 1
    n = int(r[iex]); s = int(r[ismp])
    op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].rstrip(";")
    tot[phase] += n; smp[phase] += s; ops[phase][op.split(".")[0]] += n
    if op.startswith("BAR"): phase += 1
T = sum(tot.values())
for p in sorted(tot):
    top = ", ".join("%s %.1f%%" % (k, 100.0 * v / tot[p]) for k, v in ops[p].most_common(9))
    print("phase %d: %6.2f M instr (%4.1f %%)  samples %5d (%4.1f %%) | %s" % (p, tot[p] / 1e6, 100.0 * tot[p] / T, smp[p], 100.0 * smp[p] / max(1, sum(smp.values())), top))
print("total %.2f M" % (T / 1e6))
