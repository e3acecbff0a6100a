"""Per-shape timing of the tcgen05 GEMM through the C ABI (hot: same kernel back to back)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine

cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=8, max_src_len=64, max_tgt_len=4, gemm_mode="3xtf32")
import os
shapes = [(1024, 768, 256, True), (1024, 256, 256, False), (1024, 2048, 256, True), (1024, 256, 2048, False),
          (5120, 768, 256, True), (5120, 256, 256, False), (5120, 2048, 256, True), (5120, 256, 2048, False),
          (524288, 1024, 256, False), (524288, 512, 256, False), (524288, 256, 256, False)]
if os.environ.get("ND_ATM"):
    eng.set_option("gemm_a_tmem", int(os.environ["ND_ATM"]))
if os.environ.get("ND_PERSIST"):
    eng.set_option("gemm_persistent", int(os.environ["ND_PERSIST"]))
if os.environ.get("ND_M"):      # e.g. ND_M=5120: only the decode shapes at that row count
    shapes = [(int(os.environ["ND_M"]), n, k, l) for (n, k, l) in ((768, 256, True), (256, 256, False), (2048, 256, True), (256, 2048, False))]
for (M, N, K, ln) in shapes:
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    g = torch.ones(K, device="cuda"); be = torch.zeros(K, device="cuda")
    for mode in ("3xtf32", "tf32", "simt"):
        if mode == "simt" and M > 10000:
            continue
        reps = 20 if M < 10000 else 3
        eng.test_gemm(mode, A, W, bias=b, ln=(g, be) if ln else None)
        torch.cuda.synchronize()
        eng.profile_enable(["gemm"])
        for _ in range(reps):
            eng.test_gemm(mode, A, W, bias=b, ln=(g, be) if ln else None)
        pr = eng.profile_read()
        eng.profile_enable([])
        ms, n = pr.get("gemm", (0, 1))
        us = 1e3 * ms / max(n, 1)
        tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
        print("M=%7d N=%5d K=%5d ln=%d %-7s %9.1f us  %7.2f TFLOP/s(fp32-equiv)" % (M, N, K, ln, mode, us, tf), flush=True)
