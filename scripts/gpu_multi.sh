mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -30
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 2500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
