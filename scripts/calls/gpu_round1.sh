mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 60 --warmup 5 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 90 --csv --log-file gpurun_out/launches_r1_a.csv python bench.py --steps 3 --warmup 3 --preroll 60 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1; tail -3 gpurun_out/ncu_a.log
