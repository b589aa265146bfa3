# A/B runs of tuning variants built by `python -m sph_b200.build --variant NAME -D...` (sph_b200/variants/NAME.so).
# Each variant replaces the library in this scratch copy of the repo, runs the tight parity tests and a short bench.
mkdir -p gpurun_out
cp sph_b200/libsph_b200.so /tmp/base.so
for v in base $VARIANTS; do
  if [ "$v" = base ]; then cp /tmp/base.so sph_b200/libsph_b200.so; else cp sph_b200/variants/$v.so sph_b200/libsph_b200.so; fi
  t=$(python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1)
  python bench.py --steps 60 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in d['config']['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), '| tests:', '''$t''')"
done
cp /tmp/base.so sph_b200/libsph_b200.so
