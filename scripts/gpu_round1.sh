#!/bin/bash
# First GPU bring-up: every group in its own process under a timeout so one hang cannot block the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $to python -m pytest "$@" -q -s -m gpu > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -4 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run gemm_simt 300 tests/test_gpu_kernels.py -k "gemm and simt"
run frontend 300 tests/test_gpu_kernels.py -k "frontend"
run nano_simt 300 tests/test_gpu_kernels.py -k "nano and simt"
run gemm_3xtf32 200 tests/test_gpu_kernels.py -k "gemm_plain and 3xtf32"
run gemm_tf32 200 tests/test_gpu_kernels.py -k "gemm and tf32 and not 3xtf32"
run gemm_ln_3x 200 tests/test_gpu_kernels.py -k "layernorm and 3xtf32"
run nano_tc 300 tests/test_gpu_kernels.py -k "nano and 3xtf32"
run parity_simt 600 tests/test_gpu_parity.py -k "greedy_matches and simt"
run parity_tc 900 tests/test_gpu_parity.py -k "greedy_matches and 3xtf32"
run beam 600 tests/test_gpu_parity.py -k "beam_matches"
run ragged 600 tests/test_gpu_parity.py -k "ragged or min_length or translator_api"
run full 600 tests/test_gpu_parity.py -k "full_batch"
cat gpurun_out/summary.txt
