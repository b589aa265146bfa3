# Round 2 (second half), GPU call 10 (one box)
mkdir -p gpurun_out
cp sph_b200/libsph_b200.so /tmp/base.so
for v in base t4 sg3 sg5 gr6 gr12; do
  if [ "$v" = base ]; then cp /tmp/base.so sph_b200/libsph_b200.so; else cp sph_b200/variants/$v.so sph_b200/libsph_b200.so; fi
  timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --min-timed-ms 300 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in d['config']['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3))" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2b_c10_variants.txt
cp /tmp/base.so sph_b200/libsph_b200.so
