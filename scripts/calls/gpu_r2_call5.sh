# Round 2, GPU call 5 (one box): the whole GPU suite on the new defaults, the default bench with both CPU baselines, the
# reference arm, ncu evidence of the shipped build, BASELINE config 5 at one GPU (4 - 64 M), config 4 as a workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "^PASSED" | tail -60 > gpurun_out/r2c5_gpu_tests.txt; tail -4 gpurun_out/r2c5_gpu_tests.txt
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); c=d['config']; print('bench', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in c['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), 'sync', d['e2e'].get('synchronous_value'), 'roofline', d['roofline']['frac'], 'step frac', c['step_hbm_frac'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline_ieee_build',{}).get('value'), c['timed_region'])"
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cut -c1-600 gpurun_out/r2_bench_reference_arm.json; tail -3 gpurun_out/r2_bench_reference_arm.err
TAG=r2_final bash scripts/gpu_profile.sh
python scripts/bench_cfg4.py > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; cut -c1-900 gpurun_out/r2_bench_cfg4.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --preset y --preroll 300 > gpurun_out/r2_bench_goo.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_goo.json')); print('goo', round(d['value']/1e9,3), 'G', {k: round(x*1e3,1) for k,x in d['config']['stage_ms'].items()})"
NGPU=1 TOTALS="${TOTALS:-4 8 16 32 64}" bash scripts/sweep_cfg5.sh
