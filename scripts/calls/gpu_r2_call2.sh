# Round 2, GPU call 2 (one box): second A/B + ncu captures (launch list, full set with source) of the default and the all1 build
mkdir -p gpurun_out
VARIANTS="packed2 trim packed_trim g16s4 g16s2 g12s3 pdl2 pdl2_s4 all1 all2" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c2_variants.txt
TAG=r2c2 bash scripts/gpu_profile.sh
TAG=r2c2 VARIANT=all1 bash scripts/gpu_profile.sh
ls -la gpurun_out | tail -20
