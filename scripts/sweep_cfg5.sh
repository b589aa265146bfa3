# BASELINE.json config 5: strong / weak scaling sweep, 4 M - 64 M particles at 1/2/4/8 B200.  One call per GPU count
# (gpurun hands out N GPUs of one box):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'NGPU=8 TOTALS="16 32 64" bash scripts/sweep_cfg5.sh'
# TOTALS: total particle counts in millions (the same totals at every N = the strong-scaling rows; the diagonal
# total = 4 M x N and 8 M x N = the weak-scaling rows).  One JSON line per point in gpurun_out/cfg5/.
mkdir -p gpurun_out/cfg5
N=${NGPU:-1}
for T in ${TOTALS:-4 8 16 32 64}; do
  PER=$(( T * 1000000 / N ))
  OUT=gpurun_out/cfg5/cfg5_${T}m_${N}gpu.json
  if [ "$N" = 1 ]; then
    python bench.py --particles $PER --steps 10 --warmup 3 --preroll ${PREROLL:-1000} --min-timed-ms 200 --no-cpu-baseline > $OUT 2> $OUT.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --particles $PER --steps 10 --warmup 3 --preroll ${PREROLL:-1000} --min-timed-ms 200 --no-cpu-baseline --no-cfg3 --no-parity-check > $OUT 2> $OUT.err
  fi
  python - <<PY || tail -3 $OUT.err
import json
d = json.loads([l for l in open("$OUT") if l.startswith("{")][-1]); c = d["config"]
print("cfg5 total ${T} M on $N GPU(s):", round(d["value"] / 1e9, 3), "G particle-steps/s,", round(d["ms_per_step"], 3), "ms/step, step HBM frac", round(c["step_hbm_frac"], 4),
      "dominant-kernel frac", round(d["roofline"]["frac"], 4), "neighbours", round(c["mean_neighbours_per_particle"], 1), "e2e", round(d["e2e"]["value"] / 1e9, 3), c["integrity"])
PY
done
