#!/bin/bash
# Quick A/B on the GPU box: the GPU tests, then the K3-only bench under each "NAME:ENV=VAL,ENV=VAL" variant given.
# usage: bash scripts/gpu_ab.sh TAG [variant ...]
TAG=${1:-ab}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?
grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_$TAG.log | tail -15
B="python bench.py --steps 5 --warmup 3 --no-cpu --spectra 0"
for v in default "$@"; do
  name=${v%%:*}; envs=${v#*:}; [ "$v" = default ] && envs=RB2_X=0
  env $(echo $envs | tr ',' ' ') timeout 600 $B > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo $name rc=$?
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s value %.4g  e2e %.4g  ms/step %.3f  kernel_ms %.3f frac %.4f e2e_ms %.3f" % (f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
