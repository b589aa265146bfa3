# Round 2 (second half): the default bench on N GPUs exactly as the driver runs it (no tests)
mkdir -p gpurun_out
N=${NGPU:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2b_bench_${N}gpu.json') if l.startswith('{')][-1]); c=d['config']; print(d['n_gpus'], 'GPUs', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', 'e2e', round(d['e2e']['value']/1e9,3), 'sync', round(d['e2e'].get('synchronous_value',0)/1e9,3), d['e2e'].get('pipelined_error'), 'parity', (c.get('slab_parity') or {}).get('result'), c['integrity'])" || tail -5 gpurun_out/r2b_bench_${N}gpu.err
