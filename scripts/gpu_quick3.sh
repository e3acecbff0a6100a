#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
echo "=== pytest -m gpu" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -12 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary.txt
for f in "l2t 1" "l2t 5" "t2t 1" "nano2rnn 1" "cnn2cnn 1"; do
  echo "=== profile_step $f" | tee -a gpurun_out/summary.txt
  timeout 300 python scripts/profile_step.py $f 2>&1 | tee -a gpurun_out/summary.txt
done
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/bench.err | tee -a gpurun_out/summary.txt
