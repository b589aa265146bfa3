# What the reference-named route costs (verdict item 9): the reference's UNMODIFIED start_simulation() on the GPU
# path (oracle/_ref/sph_ref_gpu_drive, BASELINE config 1: 1508 particles), as one rank and as 2 / 3 ranks whose slab
# messages go through the host's MPI_Sendrecv (sph_exchange_via_host, two exchanges per step); beside it the pure CPU
# reference through the same harness and the handle API's graph step at the same size.  All ranks share device 0.
# FRAMES frames of 4 steps; the "timing:" line leaves out start-up and the first fifth of the frames.
F=${FRAMES:-400}
O=gpurun_out/r2d_ref_route.txt
mkdir -p gpurun_out
: > $O
run() { echo "== $*" >> $O; "$@" 2>&1 | grep -E "^timing:|sph_ref_api:|error" >> $O; }
run oracle/_ref/sph_ref_gpu_drive --frames $F --out /tmp/g1.bin
run env SPH_REF_MIRROR_EVERY=4 oracle/_ref/sph_ref_gpu_drive --frames $F --out /tmp/g1m.bin
run env SPH_REF_MIRROR_EVERY=4 oracle/_ref/sph_ref_gpu_drive --ranks 2 --frames $F --out /tmp/g2.bin
run env SPH_REF_MIRROR_EVERY=4 oracle/_ref/sph_ref_gpu_drive --ranks 3 --frames $F --out /tmp/g3.bin
run oracle/_ref/sph_ref_cpu_drive --frames $F --out /tmp/c1.bin
run oracle/_ref/sph_ref_cpu_drive --ranks 3 --frames $F --out /tmp/c3.bin
echo "== handle API, same problem (graph step, no host mirror)" >> $O
${PYTHON:-python} - >> $O 2>&1 <<'PY'
import time, sys
sys.path.insert(0, ".")
import os
if os.environ.get("SPH_EMU_LIB"):
    import ctypes as C, sph_b200
    sph_b200._lib = sph_b200._bind(C.CDLL(os.environ["SPH_EMU_LIB"]))
import sph_b200
prob = sph_b200.make_problem(1500)
t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], "x")
c = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], prob["n_global"] + 64)
c.set_params(t); c.init_lattice(prob)
steps = int(os.environ.get("FRAMES", "400")) * 4
c.step(steps // 5); c.synchronize()
t0 = time.perf_counter(); c.step(steps - steps // 5); c.synchronize(); dt = time.perf_counter() - t0
print(f"timing: handle API: {steps - steps // 5} steps of {prob['n_global']} particles in {dt:.6f} s = {1e6 * dt / (steps - steps // 5):.2f} us per step")
PY
cat $O
