# usage: bash scripts/gpu_check.sh TAG   -> tests + 1-GPU bench into gpurun_out/
TAG=${1:-x}
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value %.4g ms/step %.3f e2e %.4g (%.2f ms) frac %.4f kernel_ms %.3f cpu %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d.get('cpu_baseline',{}).get('value',0)))"
tail -3 gpurun_out/bench_$TAG.err
