#!/usr/bin/env python
"""profiles/k3_traffic.json from an `ncu --set full` capture of one chunk's K3 stage kernels.

    python scripts/make_k3_traffic.py gpurun_out/prof_k3_TAG.ncu-rep HYPOTHESES_PER_CHUNK "source note"

Per stage kernel (first captured launch of each): DRAM read + write bytes, issue-slot utilisation, warp
instructions, active lanes per instruction, duration.  bench.py reads `dram_bytes_per_chunk` for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys


def main(path, per_chunk, note):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per, issue, inst, lanes, dur = {}, {}, {}, {}, {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("rb2::", "")
        if not name.startswith("ransac_") or name in per:
            continue
        unit_r, unit_w = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per[name] = int(float(r[col["dram__bytes_read.sum"]]) * scale[unit_r] + float(r[col["dram__bytes_write.sum"]]) * scale[unit_w])
        issue[name] = round(float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), 1)
        inst[name] = int(float(r[col["smsp__inst_executed.sum"]]))
        lanes[name] = round(float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]), 1)
        dur[name] = round(float(r[col["gpu__time_duration.sum"]]), 1)
    doc = {"dram_bytes_per_chunk": sum(per.values()), "hypotheses_per_chunk": int(per_chunk), "per_kernel": per,
           "issue_slots_busy_pct": issue, "warp_instructions": inst, "active_lanes_per_instruction": lanes,
           "duration_us_under_ncu": dur, "source": note}
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k3_traffic.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), sys.argv[3])
