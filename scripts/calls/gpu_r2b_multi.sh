# Round 2 (second half), N-GPU check of the shipped build (NGPU=2 by default): the slab tests that skip on one GPU and the
# default bench exactly as the driver runs it
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python -c "
import json; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); c=d['config']; print('$2', d['n_gpus'], 'GPUs', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in c['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), 'parity', (c.get('slab_parity') or {}).get('result'), 'spread', c['timed_region']['spread_rel'], 'blocks', c['timed_region']['blocks']); print('   per slab', [r[:1] + r[2:] for r in c[[k for k in c if k.startswith('per_slab')][0]]]); print('   integrity', c['integrity'], d['e2e'].get('pipelined_error'))" || tail -5 ${1%.json}.err; }
timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_zy_gpu_stabilised_and_feed.py tests/test_zzz_gpu_round2_first_contact.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2b_slab_tests_${N}gpu.txt
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; show gpurun_out/r2b_bench_${N}gpu.json default
