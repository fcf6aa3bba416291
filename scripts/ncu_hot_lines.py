#!/usr/bin/env python
"""Per-source-line instruction/sample shares from `ncu --page source --csv --print-source sass,cuda`.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > x.csv; python scripts/ncu_hot_lines.py x.csv [N]"""
import collections
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
kern = cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; hdr = None; continue
    if r[0] == "Function Name":
        kern = r[1][:48]; continue
    if r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
        si = hdr.index("# Samples") if "# Samples" in hdr else None
        continue
    if hdr is None or ii is None:
        continue
    try:
        ln = int(r[0]); ie = int(r[ii] or 0); sm = int(r[si] or 0)
    except Exception:
        continue
    agg[(kern, cur, ln)][0] += ie
    agg[(kern, cur, ln)][1] += sm
cache = {}
def line(path, ln):
    if path not in cache:
        try: cache[path] = open(path).read().split("\n")
        except Exception: cache[path] = []
    L = cache[path]
    return L[ln - 1].strip()[:88] if 0 < ln <= len(L) else ""
for K in sorted(set(k[0] for k in agg)):
    items = [(k, v) for k, v in agg.items() if k[0] == K]
    tot = sum(v[0] for k, v in items) or 1
    ts = sum(v[1] for k, v in items) or 1
    print("==", K, "warp-inst", tot, "samples", ts)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print("  %5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / ts, os.path.basename(k[1]), k[2], line(k[1], k[2])))
