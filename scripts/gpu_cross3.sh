#!/bin/bash
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cross_attention_formulations" 2>&1 | tail -3
ND_CROSS=1 timeout 300 python scripts/profile_step.py l2t 1 2>&1
