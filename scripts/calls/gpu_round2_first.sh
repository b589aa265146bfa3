# First GPU call of round 2 (one box, ~15 min): everything that was written after round 1's GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/calls/gpu_round2_first.sh'
# 1. the GPU suite (first hardware run of k_coupling / k_advect<true> and of the stabilised-viscosity tests)
# 2. A/B of the packed-FP32 builds against the default (tight parity tests + short bench each)
# 3. cost of the stabilised viscosity pass on the goo preset
# Build the variants BEFORE calling gpurun (they travel as .so files): bash scripts/calls/build_variants.sh, i.e.
#   python -m sph_b200.build --variant packed -DSPH_PACKED=1
#   python -m sph_b200.build --variant packed_relax -DSPH_PACKED=1 -DSPH_PACKED_RELAX=1
#   python -m sph_b200.build --variant packed_b3 -DSPH_PACKED=1 -DSPH_BLOCKS_ADVECT=3     # no spills, fewer warps
#   python -m sph_b200.build --variant pd4 -DSPH_RELAX_PD4=1                               # k_relax: one 16-byte record per neighbour
#   python -m sph_b200.build --variant relax_b3 -DSPH_BLOCKS_RELAX=3                       # k_relax at 85 registers (param reloads gone)
#   python -m sph_b200.build --variant pdl -DSPH_PDL=1                                     # programmatic dependent launch of all 11 kernels of a step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2_gpu_tests.txt
VARIANTS="${VARIANTS:-packed packed_relax packed_b3 pd4 relax_b3 pdl packed_pdl}" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2_variants.txt
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --preset y --visc-stab 0.5 --preroll 300 > gpurun_out/bench_goo_stab.json 2> gpurun_out/bench_goo_stab.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_goo_stab.json"))
print("goo stabilised", round(d["value"] / 1e9, 3), "G", {k: round(v * 1e3, 1) for k, v in d["config"]["stage_ms"].items()})
PY
