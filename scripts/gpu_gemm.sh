#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -8
timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "524288|5120" | grep -v simt
timeout 300 python scripts/profile_step.py l2t 1 2>&1 | head -3
timeout 300 python scripts/profile_step.py t2t 1 2>&1 | head -3
