#!/bin/bash
# one 8-GPU line of the shipped bench, launched the way the driver launches it
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err
echo rc=$?; tail -c 1500 gpurun_out/bench_r02_n8.json; tail -3 gpurun_out/bench_r02_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > gpurun_out/bench_r02_n8_reference.json 2> gpurun_out/bench_r02_n8_reference.err
echo rc=$?; tail -c 600 gpurun_out/bench_r02_n8_reference.json
