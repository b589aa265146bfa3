mkdir -p gpurun_out
TAG=${TAG:-r1}
ncu --set full --clock-control none --import-source on -k regex:"k_advect|k_density|k_relax" -s 3009 -c 3 -f -o gpurun_out/prof_${TAG}_gather python bench.py --steps 3 --warmup 3 --preroll 1000 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1; tail -2 gpurun_out/ncu_gather.log | cut -c1-300
