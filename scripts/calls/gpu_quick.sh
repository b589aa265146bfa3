mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 60 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value']/1e9, 'G p-steps/s', d['ms_per_step'], 'ms', {k: round(v*1e3,1) for k,v in d['config']['stage_ms'].items()}, 'nbrs', round(d['config']['mean_neighbours_per_particle'],1), 'e2e', d['e2e']['value']/1e9, 'b2b', d['config']['l2_resident_ms_per_step'], d['clocks'], d.get('cpu_baseline'))"; tail -3 gpurun_out/bench_quick.err
