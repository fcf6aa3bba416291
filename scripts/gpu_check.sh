#!/bin/bash
# usage (on the GPU box): bash scripts/gpu_check.sh TAG   -> gpurun_out/pytest_TAG.log, bench_TAG.json
TAG=${1:-x}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_exit=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g  ms/step %.3f  kernel_ms %.3f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]))
print(d["counters"]); print(d.get("stage_ms"))
PY
