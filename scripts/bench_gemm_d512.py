"""Per-shape timing of the decode-step GEMMs at d=512 (M = 1024 rows) and the beam shapes (M = 5120)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=8, max_src_len=64, max_tgt_len=4, gemm_mode="3xtf32")
shapes = [(1024, 1536, 512, True), (1024, 512, 512, False), (1024, 2048, 512, True), (1024, 512, 2048, False)]
for (M, N, K, ln) in shapes:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    g = torch.ones(K, device="cuda"); be = torch.zeros(K, device="cuda")
    eng.test_gemm("3xtf32", A, W, bias=b, ln=(g, be) if ln else None); torch.cuda.synchronize()
    eng.profile_enable(["gemm"])
    for _ in range(20): eng.test_gemm("3xtf32", A, W, bias=b, ln=(g, be) if ln else None)
    ms, n = eng.profile_read()["gemm"]; eng.profile_enable([])
    us = 1e3 * ms / n
    print("M=%d N=%d K=%d ln=%d  %.1f us  %.1f TFLOP/s fp32-equiv" % (M, N, K, ln, us, 2.0 * M * N * K / us / 1e6), flush=True)
