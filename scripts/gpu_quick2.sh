#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
echo "=== pytest -m gpu" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -12 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary.txt
timeout 120 python scripts/gemm_timeline.py 2>&1 | tee -a gpurun_out/summary.txt
for f in "l2t 1" "l2t 5" "t2t 1"; do
  echo "=== profile_step $f" | tee -a gpurun_out/summary.txt
  timeout 300 python scripts/profile_step.py $f 2>&1 | tee -a gpurun_out/summary.txt
done
echo "=== bench_gemm" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/bench_gemm.py 2>&1 | grep -v simt | tee -a gpurun_out/summary.txt
