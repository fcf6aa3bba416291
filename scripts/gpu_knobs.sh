#!/bin/bash
# smoke(), then the K3-only bench under chunk / lane knobs at a rank's share (1.25e7), a quarter (2.5e7) and the full step (1e8).
# usage: bash scripts/gpu_knobs.sh TAG
TAG=${1:-knobs}
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo smoke rc=$? t=$((SECONDS-T0)); tail -2 gpurun_out/smoke_$TAG.log
one() {  # name hyp env...
  local name=$1 hyp=$2; shift; shift
  env RB2_X=0 "$@" timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu --spectra 0 --e2e-steps 2 --hyp $hyp 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-28s hyp %-8s K3 %.3f ms  step %.3f  seq %.3f  e2e %.3f  | %s' % ('$name', '$hyp', d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['sequential']['ms_per_step'], d['e2e']['ms_per_step'], d['config']['parallelism']))"
}
one default 1.25e7
one floor22 1.25e7 RB2_CHUNK_FLOOR_LOG2=22
one floor22_lanes2 1.25e7 RB2_CHUNK_FLOOR_LOG2=22 RB2_PIPE_LANES=2
one floor21_2perlane 1.25e7 RB2_CHUNK_FLOOR_LOG2=21 RB2_CHUNKS_PER_LANE=2
one onechunk 1.25e7 RB2_CHUNK_FLOOR_LOG2=24
one floor22_lanes4 1.25e7 RB2_CHUNK_FLOOR_LOG2=21 RB2_PIPE_LANES=4
one default 1.25e7
echo t=$((SECONDS-T0))
one default 2.5e7
one floor22_lanes4 2.5e7 RB2_CHUNK_FLOOR_LOG2=22 RB2_PIPE_LANES=4
one floor22_2perlane 2.5e7 RB2_CHUNK_FLOOR_LOG2=22 RB2_CHUNKS_PER_LANE=2
echo t=$((SECONDS-T0))
one default 1e8
one lanes4 1e8 RB2_PIPE_LANES=4
one 2perlane 1e8 RB2_CHUNKS_PER_LANE=2
one a4 1e8 RB2_A_CTAS_PER_SM=4
one c1_6 1e8 RB2_C1_CTAS_PER_SM=6
one c1_3 1e8 RB2_C1_CTAS_PER_SM=3
one default 1e8
echo total t=$((SECONDS-T0))
