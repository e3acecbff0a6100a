#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streams or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary.txt
for ns in 1 2 4 8; do
  echo "=== streams $ns" | tee -a gpurun_out/summary.txt
  ND_STREAMS=$ns timeout 300 python scripts/profile_step.py l2t 1 2>&1 | head -3 | tee -a gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tee -a gpurun_out/summary.txt
