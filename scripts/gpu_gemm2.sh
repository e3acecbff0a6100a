#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -4
timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "524288" | grep -v simt
