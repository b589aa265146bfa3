# Round 2, GPU call 7 (one box): the shipped build (SPH_TRIM back on) against the untrimmed loop; ncu evidence; bench record
mkdir -p gpurun_out
VARIANTS="notrim" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c7_variants.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2c7_gpu_tests.txt
TAG=r2_final bash scripts/gpu_profile.sh
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); c=d['config']; print('bench', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in c['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), 'roofline', d['roofline'], 'step frac', c['step_hbm_frac'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline_ieee_build',{}).get('value'))"
python scripts/bench_cfg4.py > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg4.json')); print('cfg4', round(d['value']/1e9,3), 'G', [(p['preset'], round(p['particle_steps_per_s']/1e9,3)) for p in d['phases']])"
