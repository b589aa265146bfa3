# Round 2 (second half): the bench with its end-to-end leg in blocks from the restored state -- default run (N = 1) and
# the driver's shape (--steps 20 --warmup 5)
mkdir -p gpurun_out
python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_n1.json')); c=d['config']; print('bench', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', 'e2e', round(d['e2e']['value']/1e9,3), 'sync', round(d['e2e'].get('synchronous_value',0)/1e9,3), 'l2res', round(c['l2_resident_value']/1e9,3), d['e2e'].get('pipelined_error'))" || tail -5 gpurun_out/r2b_bench_n1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_n1_s20.json 2> gpurun_out/r2b_bench_n1_s20.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_n1_s20.json')); c=d['config']; print('steps20', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', 'e2e', round(d['e2e']['value']/1e9,3), 'sync', round(d['e2e'].get('synchronous_value',0)/1e9,3), d['e2e'].get('pipelined_error'))" || tail -5 gpurun_out/r2b_bench_n1_s20.err
