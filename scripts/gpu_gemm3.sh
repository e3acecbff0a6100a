#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm_persistent" 2>&1 | grep -E "assert|Error|passed|failed" | head -8
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=8, max_src_len=64, max_tgt_len=4, gemm_mode="3xtf32")
for pm in (1, 2, 0):
    eng.set_option("gemm_persistent", pm)
    for (M, N, K) in [(524288, 1024, 256), (524288, 512, 256)]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
        eng.test_gemm("3xtf32", A, W, bias=b); torch.cuda.synchronize()
        eng.profile_enable(["gemm"])
        for _ in range(3): eng.test_gemm("3xtf32", A, W, bias=b)
        ms, n = eng.profile_read()["gemm"]; eng.profile_enable([])
        us = 1e3 * ms / n
        print("persist=%d M=%d N=%d K=%d  %.1f us  %.1f TFLOP/s" % (pm, M, N, K, us, 2.0 * M * N * K / us / 1e6))
PY
