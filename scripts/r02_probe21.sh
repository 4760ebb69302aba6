#!/bin/bash
# 2-GPU line launched as the driver launches it: stdout must hold exactly the one JSON line
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "rc=$? stdout_lines=$(wc -l < gpurun_out/bench_r02_n2.json) nccl_banner_in_stderr=$(grep -c 'NCCL version' gpurun_out/bench_r02_n2.err)"
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n2.json')); print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['kernel_ms_per_rank'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_r02_n2_reference.json 2>/dev/null
echo "reference rc=$? stdout_lines=$(wc -l < gpurun_out/bench_r02_n2_reference.json)"
