#!/bin/bash
# Two-GPU check of a tree: rank-count independence of the sharded fit, then the bench line at N = 2.  usage: bash scripts/gpu_n2.sh TAG
TAG=${1:-n2}
mkdir -p gpurun_out
T0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR scripts/check_dist.py > gpurun_out/check_dist_$TAG.log 2>&1; echo check_dist rc=$? t=$((SECONDS-T0)); tail -4 gpurun_out/check_dist_$TAG.log
timeout 300 $TR bench.py --gpus 2 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo bench rc=$? t=$((SECONDS-T0))
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
print("n2 value %.4g e2e %.4g ms/step %.3f kernel_ms %.3f e2e_ms %.3f seq %.3f spectra/s %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["ms_per_step"], d["sequential"]["ms_per_step"], d.get("spectra_per_sec", {}).get("value")))
PY
