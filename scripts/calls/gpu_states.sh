mkdir -p gpurun_out
for P in 0 4000; do
python bench.py --steps 100 --warmup 5 --preroll $P --no-cpu-baseline > gpurun_out/bench_state_$P.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_state_$P.json')); print('preroll', $P, round(d['value']/1e9,3), 'G p-steps/s', round(d['ms_per_step'],4), 'ms', {k: round(v*1e3,1) for k,v in d['config']['stage_ms'].items()}, 'nbrs', round(d['config']['mean_neighbours_per_particle'],1), 'max_bucket', d['config']['max_bucket'], 'e2e', round(d['e2e']['value']/1e9,3), 'hbm_frac', round(d['config']['step_hbm_frac'],3))"
done
