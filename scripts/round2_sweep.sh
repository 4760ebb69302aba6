#!/bin/bash
# Prepared for the next round (no GPU minutes were left to run it in round 1).
# Step 1, HERE (CPU, ~1 min each, in parallel):   scripts/round2_sweep.sh build
# Step 2, one gpurun call (~12 s per variant):     gpurun --timeout 600 -- 'scripts/round2_sweep.sh run'
# Why these: with the shared-memory cap gone (17.8 KB / CTA) the register budget alone sets occupancy, so the
# occupancy/width points measured earlier under the 9-CTA cap are worth re-measuring; --streams 2 overlaps the
# ~13 ms tail of one step's persistent kernel with the head of the next (steady-state throughput).
set -u
cd "$(dirname "$0")/.."
variants=(
  "mb11|-DDIB_MINBLOCKS4=11"
  "mb12|-DDIB_MINBLOCKS4=12"
  "t32x20|-DDIB_THREADS=32 -DDIB_MINBLOCKS4=20 -DDIB_MINBLOCKS6=16"
  "t32x24|-DDIB_THREADS=32 -DDIB_MINBLOCKS4=24 -DDIB_MINBLOCKS6=16"
  "t128x5|-DDIB_THREADS=128 -DDIB_MINBLOCKS4=5 -DDIB_MINBLOCKS6=4"
  "gps3|-DDIB_GPS=3"
  "ilp2|-DDIB_EXACT_ILP=2 -DDIB_MINBLOCKS4=8"
  "six9|-DDIB_MINBLOCKS6=9 -DDIB_RING=64 -DDIB_GPS=1"
  "six10|-DDIB_MINBLOCKS6=10 -DDIB_RING=64 -DDIB_GPS=1"
)
case "${1:-}" in
  build)
    for v in "${variants[@]}"; do
      name="${v%%|*}"; flags="${v#*|}"
      DIB_NVCC_EXTRA="$flags" python -m deepi2p_b200.build --out "$name" > /dev/null 2>&1 &
    done
    wait; ls -la deepi2p_b200/lib/variants/ ;;
  run)
    names=(default); for v in "${variants[@]}"; do names+=("${v%%|*}"); done; names+=(default)
    scripts/ab_prebuilt.sh "${names[@]}"
    echo "--- config 2 (4096 x 1 init): CTA width"
    BENCH_ARGS="--workload single_init" SWEEP_SAMPLES=4096 scripts/ab_prebuilt.sh default "default|DIB_WIDE_BELOW=1000000" t128x5 t32x20
    echo "--- the in-library 128-thread build (frustum_solver_wide.cu) through the whole parity suite"
    DIB_WIDE_BELOW=1000000000 python -m pytest tests/test_frustum_gpu.py -m gpu -q 2>&1 | tail -3
    echo "--- 6-DoF occupancy (96 registers, smaller rings: 9-10 CTAs/SM instead of 8)"
    BENCH_ARGS="--is-3d" SWEEP_SAMPLES=256 scripts/ab_prebuilt.sh default six9 six10 default
    echo "--- overlapped steps"
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --streams 2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams=2 value', d['value'], 'e2e', d['e2e']['value'])"
    ;;
  *) echo "usage: $0 build|run" ;;
esac
