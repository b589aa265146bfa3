# ncu evidence for the current kernels: launch list of two timed steps + one full capture per kernel,
# instrumenting only the timed region (cudaProfilerStart/Stop in bench.py).  State = bench state.
# VARIANT=packed: profile sph_b200/variants/packed.so instead of the default build (restored afterwards)
mkdir -p gpurun_out
TAG=${TAG:-r1_final}
if [ -n "$VARIANT" ]; then cp sph_b200/libsph_b200.so /tmp/base_profile.so; cp sph_b200/variants/$VARIANT.so sph_b200/libsph_b200.so; TAG=${TAG}_$VARIANT; fi
SPH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-200
SPH_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_advect|k_density|k_relax|k_scan_totals|k_scan_apply|k_scatter|k_reorder" -c 7 -f -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log | cut -c1-200
if [ -n "$VARIANT" ]; then cp /tmp/base_profile.so sph_b200/libsph_b200.so; fi
