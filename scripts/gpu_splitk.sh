#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -4
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "beam or full_batch or streams" 2>&1 | tail -3
timeout 300 python scripts/profile_step.py l2t 5 2>&1 | tail -9
timeout 300 python scripts/profile_step.py l2t 1 2>&1 | sed -n 3,3p
