# Round 2, 8-GPU A/B: two-exchange build vs one-exchange build with exchange periods 1/2/4 (bench with state restore)
mkdir -p gpurun_out
N=${NGPU:-8}
cp sph_b200/libsph_b200.so /tmp/base.so
run() {   # name, lib, extra bench args
  cp $2 sph_b200/libsph_b200.so
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline --min-timed-ms 250 $3 ${BENCH_ARGS} > gpurun_out/bench_${N}gpu_$1.json 2> gpurun_out/bench_${N}gpu_$1.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${N}gpu_$1.json') if l.startswith('{')][-1]); c=d['config']; print('$1', d['n_gpus'], 'GPUs', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in c['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), 'parity', (c.get('slab_parity') or {}).get('result'), 'spread', c['timed_region']['spread_rel'], 'blocks', c['timed_region']['blocks']); print('   per slab', [r[2:] for r in c[[k for k in c if k.startswith('per_slab')][0]]]); print('   integrity', c['integrity'], d['e2e'].get('pipelined_error'), c.get('cfg3_16m'))" || tail -5 gpurun_out/bench_${N}gpu_$1.err
}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-ms 250 > gpurun_out/bench_1gpu_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_ref.json')); print('single', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', d['config']['timed_region'], 'e2e', round(d['e2e']['value']/1e9,3), d['config']['mean_neighbours_per_particle'])"
run base /tmp/base.so "--no-cfg3"
run onex_p1 sph_b200/variants/onex.so "--exchange-period 1 --no-cfg3"
run onex_p2 sph_b200/variants/onex.so "--exchange-period 2 --no-cfg3"
run onex_p4 sph_b200/variants/onex.so "--exchange-period 4"
run onex_p4_cost sph_b200/variants/onex.so "--exchange-period 4 --balance cost --no-cfg3"
cp /tmp/base.so sph_b200/libsph_b200.so
