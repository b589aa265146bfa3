mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_advect|k_density|k_relax" -s 3009 -c 3 -f -o gpurun_out/prof_r1_gather python bench.py --steps 3 --warmup 3 --preroll 1000 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1; tail -3 gpurun_out/ncu_gather.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 9027 -c 60 --csv --log-file gpurun_out/launches_r1_b.csv python bench.py --steps 3 --warmup 3 --preroll 1000 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
