# Round 2, GPU call 4: branch-free relax loop (SPH_RELAX_BF), SPH_PIPE with the slot store really deferred
mkdir -p gpurun_out
VARIANTS="pt_nopipe pt_pipe pt_bf pt_bf_pipe pt_bf_pd4" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c4_variants.txt
TAG=r2c4 VARIANT=pt_bf bash scripts/gpu_profile.sh
