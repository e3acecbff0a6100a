#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
echo "=== pytest -m gpu" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary.txt
for f in "l2t 1" "l2t 5" "t2t 1" "nano2rnn 1"; do
  echo "=== profile_step $f" | tee -a gpurun_out/summary.txt
  timeout 300 python scripts/profile_step.py $f 2>&1 | tee -a gpurun_out/summary.txt
done
echo "=== ncu full gemm" | tee -a gpurun_out/summary.txt
python scripts/profile_step.py l2t 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 0 -c 14 -o gpurun_out/prof_gemm -f python scripts/profile_step.py l2t 1 > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_kernel|self_attn" -s 0 -c 5 -o gpurun_out/prof_lstm_self -f python scripts/profile_step.py l2t 1 > gpurun_out/ncu_lstm.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
