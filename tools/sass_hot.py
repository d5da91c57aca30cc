#!/usr/bin/env python
"""Summarise the source page of an ncu report (ncu -i X.ncu-rep --page source --csv > f.csv):
instruction-weighted regions of the SASS, with stall samples.   python tools/sass_hot.py f.csv [bucket]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
ix = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
body = rows[hdr_i + 1:]
tot_i = sum(float(r[ix["Instructions Executed"]] or 0) for r in body)
tot_s = sum(float(r[ix["# Samples"]] or 0) for r in body)
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print("total inst %.3g samples %d lines %d" % (tot_i, tot_s, len(body)))
for b0 in range(0, len(body), bucket):
    blk = body[b0:b0 + bucket]
    ins = sum(float(r[ix["Instructions Executed"]] or 0) for r in blk)
    thr = sum(float(r[ix["Thread Instructions Executed"]] or 0) for r in blk)
    smp = sum(float(r[ix["# Samples"]] or 0) for r in blk)
    if ins / tot_i > 0.004 or smp / tot_s > 0.004:
        first = blk[0][ix["Source"]].strip()[:40]
        print("%5d-%5d inst %5.1f%% samples %5.1f%% thr/inst %4.1f  %s" % (b0, b0 + len(blk), 100 * ins / tot_i, 100 * smp / tot_s, thr / max(ins, 1), first))
