# Round 2, final multi-GPU records on N GPUs of one box (NGPU=8 / 4 / 2): the default bench exactly as the driver runs
# it, the reference-shaped configuration beside it, the slab tests of this world size, BASELINE config 5 points
mkdir -p gpurun_out/cfg5
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python -c "
import json; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); c=d['config']; print('$2', d['n_gpus'], 'GPUs', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in c['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), 'parity', (c.get('slab_parity') or {}).get('result'), 'spread', c['timed_region']['spread_rel'], 'blocks', c['timed_region']['blocks']); print('   per slab', [r[:1] + r[2:] for r in c[[k for k in c if k.startswith('per_slab')][0]]]); print('   integrity', c['integrity'], d['e2e'].get('pipelined_error')); print('   cfg3', c.get('cfg3_16m'))" || tail -5 ${1%.json}.err; }
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; show gpurun_out/r2_bench_${N}gpu.json default
$TR bench.py --gpus $N --steps 20 --warmup 5 --exchanges-per-step 2 --balance count --no-cfg3 > gpurun_out/r2_bench_${N}gpu_reference_shape.json 2> gpurun_out/r2_bench_${N}gpu_reference_shape.err; show gpurun_out/r2_bench_${N}gpu_reference_shape.json reference-shape
if [ -n "$TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_zy_gpu_stabilised_and_feed.py tests/test_zzz_gpu_round2_first_contact.py -m gpu -q -k "$TESTS" 2>&1 | tail -4 | tee gpurun_out/r2_slab_tests_${N}gpu.txt; fi
NGPU=$N TOTALS="$TOTALS" bash scripts/sweep_cfg5.sh
