# Round 2 (second half), GPU call 3 (one box): A/B on top of n1; per-kernel times of n1 from an ncu launch list
mkdir -p gpurun_out
cp sph_b200/libsph_b200.so /tmp/base.so
for v in n1 n1_pv4 n1_scan4 n1_scan16 n1_sg4 n1_sg6 n1_sg12 n1_rb5 n1_ab5 n1_gr4 n1_gr16 n1_ga4 n1_ga16; do
  cp sph_b200/variants/$v.so sph_b200/libsph_b200.so
  t=-
  case $v in n1_pv4) t=$(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1);; esac
  timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --min-timed-ms 300 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us', {k: round(x*1e3,1) for k,x in d['config']['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), '| tests:', '''$t''')" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2b_c3_variants.txt
cp sph_b200/variants/n1.so sph_b200/libsph_b200.so
SPH_PROFILE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b_n1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-200
cp /tmp/base.so sph_b200/libsph_b200.so
