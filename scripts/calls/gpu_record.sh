# the round's record runs: default bench (as the driver runs it) and the reference arm
mkdir -p gpurun_out
python bench.py > gpurun_out/BENCH_local_n1.json 2> gpurun_out/BENCH_local_n1.err; tail -c 2600 gpurun_out/BENCH_local_n1.json
python bench.py --impl reference > gpurun_out/BENCH_local_ref.json 2>/dev/null; cat gpurun_out/BENCH_local_ref.json | cut -c1-400
