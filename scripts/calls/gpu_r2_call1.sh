# Round 2, GPU call 1 (one box): full GPU suite WITHOUT -x (every failure visible), then the A/B of all build variants.
#   bash scripts/calls/build_variants.sh && /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/calls/gpu_r2_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt
timeout 1000 python -m pytest tests -m gpu -q -rA 2>&1 | tail -120 > gpurun_out/r2c1_gpu_tests.txt
tail -5 gpurun_out/r2c1_gpu_tests.txt
VARIANTS="${VARIANTS:-packed packed_relax packed_b3 pd4 relax_b3 pdl packed_pdl flat m4 m16 packed_flat packed_pd4_pdl_flat}" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c1_variants.txt
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --preset y --preroll 300 > gpurun_out/bench_goo_default.json 2> gpurun_out/bench_goo_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_goo_default.json"))
print("goo (default = stabilised)", round(d["value"] / 1e9, 3), "G", {k: round(v * 1e3, 1) for k, v in d["config"]["stage_ms"].items()})
PY
