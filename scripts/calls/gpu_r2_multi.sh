# Round 2, multi-GPU A/B: the two-exchange build against the one-exchange build with exchange periods 1, 2, 4
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'NGPU=2 bash scripts/calls/gpu_r2_multi.sh'
mkdir -p gpurun_out
N=${NGPU:-2}
cp sph_b200/libsph_b200.so /tmp/base.so
run() {   # name, lib, extra bench args
  cp $2 sph_b200/libsph_b200.so
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline $3 ${BENCH_ARGS} > gpurun_out/bench_${N}gpu_$1.json 2> gpurun_out/bench_${N}gpu_$1.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${N}gpu_$1.json') if l.startswith('{')][-1]); print('$1', d['n_gpus'], 'GPUs', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in d['config']['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3))" || tail -5 gpurun_out/bench_${N}gpu_$1.err
}
if [ -z "$SKIP_TESTS" ]; then
  timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_zy_gpu_stabilised_and_feed.py -m gpu -q -k "slab" 2>&1 | tail -5 | tee gpurun_out/r2_multi_tests_${N}gpu.txt
fi
cp /tmp/base.so /tmp/one.so
python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_ref.json')); print('single', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us')"
run base /tmp/base.so ""
run base_cost /tmp/base.so "--balance cost"
run onex_p1 sph_b200/variants/onex.so "--exchange-period 1"
run onex_p2 sph_b200/variants/onex.so "--exchange-period 2"
run onex_p4 sph_b200/variants/onex.so "--exchange-period 4"
run onex_p4_cost sph_b200/variants/onex.so "--exchange-period 4 --balance cost"
cp /tmp/base.so sph_b200/libsph_b200.so
