mkdir -p gpurun_out
N=${NGPU:-2}
export SPH_SPIN_TIMEOUT_MS=${SPIN_MS:-10000}
for NTOT in 200000; do
timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/slab_gpu_worker.py /tmp/dbgslab $NTOT ${STEPS:-120} p2p > gpurun_out/dbg_worker.log 2>&1; echo worker $NTOT rc=$?
grep -E "^rank|edges|Error" gpurun_out/dbg_worker.log | tail -5 | cut -c1-250
done
