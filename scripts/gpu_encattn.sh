#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "tensor_core_vs_ffma" 2>&1 | grep -E "T=|passed|failed|Error|error" | head -20
timeout 300 python scripts/profile_step.py t2t 1 2>&1 | tail -8
