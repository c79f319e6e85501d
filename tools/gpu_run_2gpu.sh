#!/bin/bash
# 2-GPU weak-scaling check: the driver's own launch line + the world-size-2 search test
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit=$?" >> gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
echo "exit=$?" >> gpurun_out/bench_2gpu_ref.err
echo done
