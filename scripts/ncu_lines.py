#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K`.

    python scripts/ncu_lines.py file.csv [top_n]
"""
import csv
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    fpath, hdr, out = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":  # a CUDA source line (aggregated)
            d = dict(zip(hdr, r))
            try:
                out.append((int(d["Instructions Executed"]), int(d["# Samples"]), int(d["Thread Instructions Executed"]),
                            fpath, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    tot_i = sum(o[0] for o in out) or 1
    tot_s = sum(o[1] for o in out) or 1
    print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
    for o in sorted(out, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%d  %s" % (100 * o[0] / tot_i, 100 * o[1] / tot_s,
                                                                    o[2] / max(o[0], 1), o[3], o[4], o[5]))


if __name__ == "__main__":
    main()
