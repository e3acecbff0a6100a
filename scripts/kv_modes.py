#!/usr/bin/env python
"""Storage formats of the decoder's memory keys / values against each other (GPU box): cross-attention time per
launch (CUDA events around every launch) and logit error against the fp32-storage run, d = 256 and d = 512, B = 1024.

    python scripts/kv_modes.py [modes, default 0,1,2,3,4,5]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine

NAMES = {0: "f32", 1: "q24 (I2F)", 2: "q16 (I2F)", 3: "q23 (magic)", 4: "q15 (magic)", 5: "fp24"}
modes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3,4,5").split(",")]
fasts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1").split(",")]       # cross_packed_fast variants
B, T, L = 1024, 512, 24
chunks, lengths = synth.make_chunks(B, T=T, seed=1234, ragged=True, read_len=16)
order = torch.argsort(lengths, descending=True, stable=True)
src, lens = chunks[order].cuda(), lengths[order].cuda()
for label, cfg in (("d=256 (l2t 3+3)", ModelConfig.family("l2t")),
                   ("d=512 (t2t enc1 dec3)", ModelConfig.family("t2t", d_model=512, enc_layers=1, dec_layers=3))):
    sd = synth.make_state_dict(cfg)
    eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=L)
    ref = None
    print(label)
    for m, fast in [(m, f) for m in modes for f in (fasts if m >= 3 else fasts[:1])]:
        eng.set_option("kv_mode", m)
        eng.set_option("cross_packed_fast", fast)
        eng.encode(src, lens)
        out = eng.decode_greedy(L, return_logits=True)
        torch.cuda.synchronize()
        lg, ids = out["logits"].cpu(), out["ids"].cpu()
        if ref is None:
            ref = (lg, ids)
        err = float((lg - ref[0]).abs().max() / ref[0].abs().max())
        same = int(ids.eq(ref[1]).all(1).sum())
        eng.profile_enable(["cross_attn"])
        eng.decode_greedy(L)
        ms, n = eng.profile_read()["cross_attn"]
        eng.profile_enable([])
        print("  kv_mode %d fast %d %-12s cross attention %7.1f us/launch (%d launches)   logits rel err vs f32 storage %.2e   "
              "%d/%d chunks identical" % (m, fast, NAMES[m], 1e3 * ms / n, n, err, same, B), flush=True)
    eng.close()
