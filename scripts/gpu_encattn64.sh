#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "tensor_core_vs_ffma" 2>&1 | grep -E "T=|passed|failed|Error|error" | head -20
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "d512" 2>&1 | tail -2
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 1024 2>&1 | tail -8
