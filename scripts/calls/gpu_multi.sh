mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L | wc -l
[ -n "$SKIP_TESTS" ] || python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.json') if l.startswith('{')][-1]); print(d['n_gpus'], 'GPUs', d['value']/1e9, 'G p-steps/s', d['ms_per_step'], 'ms', {k: round(v*1e3,1) for k,v in d['config']['stage_ms'].items()}, 'nbrs', round(d['config']['mean_neighbours_per_particle'],1), 'e2e', d['e2e']['value']/1e9, 'b2b', d['config']['l2_resident_ms_per_step'], d['clocks']); print(d['config'].get('per_slab_[n_local,n_ghost,neighbours,gather_us,sort_us]'))"; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_${N}gpu.err | tail -5
