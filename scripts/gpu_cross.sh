#!/bin/bash
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -12
for m in 1 0; do
  echo "=== cross_mode=$m"
  ND_CROSS=$m timeout 300 python scripts/profile_step.py l2t 1 2>&1
done
echo "=== t2t cross_mode=1"; ND_CROSS=1 timeout 300 python scripts/profile_step.py t2t 1 2>&1 | head -4
