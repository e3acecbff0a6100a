#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  echo "=== variant $v"
  timeout -k 5 120 python scripts/lstm_check.py $v 2>&1 | tail -15
  echo "exit $?"
done
