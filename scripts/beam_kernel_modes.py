import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
cfg = ModelConfig.family("l2t")
sd = synth.make_state_dict(cfg, seed=3)
K, L = 5, 4
for B in (6, 300):
  for T in (96, 128, 160, 256, 512):
    chunks, lengths = synth.make_chunks(B, T=T, seed=5, ragged=True, read_len=7)
    res = {}
    for mode, grp in ((0, 2), (2, 2), (2, 1), (1, 2)):
        eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=L, max_beam=K)
        eng.set_option("cross_beam_kernel", mode); eng.set_option("cross_ring_groups", grp)
        eng.encode(chunks.cuda(), lengths.cuda())
        out = eng.decode_beam(K, K, L, L - 1)
        torch.cuda.synchronize()
        res[(mode, grp)] = out["scores"].cpu().numpy()
    ref = res[(0, 2)]
    print("B", B, "T", T, " ".join("mode%d/g%d maxdiff %.2e" % (m, g, np.abs(v - ref).max()) for (m, g), v in res.items() if m), flush=True)
