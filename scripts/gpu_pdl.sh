#!/bin/bash
mkdir -p gpurun_out
echo "=== parity subset"
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for pdl in 0 1; do
  for f in "l2t 1" "l2t 5" "nano2rnn 1"; do
    echo "=== pdl=$pdl profile_step $f"
    ND_PDL=$pdl timeout 300 python scripts/profile_step.py $f 2>&1 | head -4
  done
done
for smp in none nvml smi; do
  echo "=== bench sampler=$smp"
  ND_BENCH_SAMPLER=$smp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-260
done
