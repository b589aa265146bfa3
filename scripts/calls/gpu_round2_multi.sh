# Multi-GPU A/B of round 2: the shipped two-exchange build against the one-exchange build (ghosts relaxed
# redundantly, neighbours meet once per step), each with the reference's count-based edge policy and with the
# cost-based one.  Build the variant first:  python -m sph_b200.build --variant onex -DSPH_ONE_EXCHANGE=1
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'NGPU=8 bash scripts/calls/gpu_round2_multi.sh'
mkdir -p gpurun_out
N=${NGPU:-2}
cp sph_b200/libsph_b200.so /tmp/base.so
for v in base ${VARIANTS:-onex}; do
  if [ "$v" = base ]; then cp /tmp/base.so sph_b200/libsph_b200.so; else cp sph_b200/variants/$v.so sph_b200/libsph_b200.so; fi
  t=$(python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -1)
  for pol in count cost; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --balance $pol ${BENCH_ARGS} > gpurun_out/bench_${N}gpu_${v}_${pol}.json 2> gpurun_out/bench_${N}gpu_${v}_${pol}.err
    python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${N}gpu_${v}_${pol}.json') if l.startswith('{')][-1]); print('$v', '$pol', d['n_gpus'], 'GPUs', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in d['config']['stage_ms'].items()}, '| slab tests:', '''$t''')"
  done
done
cp /tmp/base.so sph_b200/libsph_b200.so
