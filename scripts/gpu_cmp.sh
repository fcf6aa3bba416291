#!/bin/bash
# K3-only bench of the working tree against a variant library, alternating.  usage: bash scripts/gpu_cmp.sh VARIANT_SO [reps] [bench args...]
V=$1; R=${2:-2}; shift; shift
one() { env "$1" python bench.py --no-cpu --spectra 0 "${@:3}" | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', '%.4g' % d['value'], 'kernel_ms %.3f' % d['roofline']['kernel_ms_per_step'], 'ms/step %.3f' % d['ms_per_step'], 'e2e_ms %.3f' % d['e2e']['ms_per_step'])"; }
for i in $(seq $R); do one RB2_X=0 new "$@"; one RB2_LIB=$PWD/$V base "$@"; done
