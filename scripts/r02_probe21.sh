#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "rc=$? lines=$(wc -l < gpurun_out/bench_r02_n2.json)"; head -c 150 gpurun_out/bench_r02_n2.json; echo
grep -c "NCCL version" gpurun_out/bench_r02_n2.err
timeout 200 python tests/tools/dropin_breakdown.py 2>&1 | tail -6
