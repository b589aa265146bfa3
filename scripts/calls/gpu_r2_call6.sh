# Round 2, GPU call 6 (one box): the sort without k_scan_totals (tile totals from the producers' atomics) against the
# four-kernel sort; the GPU suite; config 4 as a workload; ncu evidence of the shipped build
mkdir -p gpurun_out
VARIANTS="notile" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c6_variants.txt
timeout 1200 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "^PASSED" | tail -40 > gpurun_out/r2c6_gpu_tests.txt; tail -4 gpurun_out/r2c6_gpu_tests.txt
python scripts/bench_cfg4.py > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; cut -c1-1200 gpurun_out/r2_bench_cfg4.json; tail -2 gpurun_out/r2_bench_cfg4.err
TAG=r2_final bash scripts/gpu_profile.sh
