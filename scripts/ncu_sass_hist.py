#!/usr/bin/env python
"""Dynamic SASS opcode histogram (warp instructions executed, stall samples) of one kernel from
`ncu -i X.ncu-rep --page source --print-source sass --csv --kernel-name regex:K > file.csv`.

    python scripts/ncu_sass_hist.py file.csv [top_n]
"""
import csv
import sys
from collections import Counter


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    iS, iI, iN, iT = (hdr.index(k) for k in ("Source", "Instructions Executed", "# Samples", "Thread Instructions Executed"))
    tot = sum(int(r[iI]) for r in data) or 1
    tot_s = sum(int(r[iN]) for r in data) or 1
    c, s, t = Counter(), Counter(), Counter()
    for r in data:
        parts = r[iS].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        key = op.split(".")[0].rstrip(";")
        c[key] += int(r[iI]); s[key] += int(r[iN]); t[key] += int(r[iT])
    print("warp instructions %d, samples %d, static SASS %d" % (tot, tot_s, len(data)))
    for k, v in c.most_common(top):
        print("%-10s %6.2f%% inst  %6.2f%% samples  thr/inst %4.1f" % (k, 100 * v / tot, 100 * s[k] / tot_s, t[k] / max(v, 1)))


if __name__ == "__main__":
    main()
