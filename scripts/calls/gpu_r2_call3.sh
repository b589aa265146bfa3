# Round 2, GPU call 3 (one box): A/B of SPH_PIPE (key-based rows, L1 prefetch of the next particle, deferred slot store)
mkdir -p gpurun_out
VARIANTS="nopipe pt pt_nopipe pt_g" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c3_variants.txt
TAG=r2c3 VARIANT=pt bash scripts/gpu_profile.sh
