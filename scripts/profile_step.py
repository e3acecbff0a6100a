"""Two steps of the bench workload (L2T greedy, B=1024) for ncu: the first warms up, the second is
the one to read in the launch list.  Usage: python scripts/profile_step.py [family] [beam]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine

family = sys.argv[1] if len(sys.argv) > 1 else "l2t"
beam = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
kw = eval(os.environ["ND_KW"]) if os.environ.get("ND_KW") else {}     # e.g. ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)"
cfg = ModelConfig.family(family, **kw)
sd = synth.make_state_dict(cfg)
eng = Engine(cfg, sd, max_batch=B, max_src_len=512, max_tgt_len=100, max_beam=beam)
if os.environ.get("ND_CROSS"):
    eng.set_option("cross_mode", int(os.environ["ND_CROSS"]))
if os.environ.get("ND_PDL"):
    eng.set_option("pdl", int(os.environ["ND_PDL"]))
if os.environ.get("ND_STREAMS"):
    eng.set_option("decode_streams", int(os.environ["ND_STREAMS"]))
for kv in filter(None, os.environ.get("ND_OPTS", "").split(",")):      # e.g. ND_OPTS="cross_beam_kernel=1,pdl=0"
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
MINLEN = int(os.environ.get("ND_MINLEN", "0"))
chunks, lengths = synth.make_chunks(B, T=512, seed=1234, ragged=True, read_len=16)
order = torch.argsort(lengths, descending=True, stable=True)
src, lens = chunks[order].cuda(), lengths[order].cuda()
cats = ["gemm", "lstm", "cross_attn", "self_attn", "enc_attn", "mlp_attn", "generator", "beam", "other"]
for it in range(4):
    torch.cuda.synchronize()
    eng.reset_launch_count()
    if it == 3:
        eng.profile_enable(cats)
    t0 = time.perf_counter()
    eng.encode(src, lens)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = eng.decode_beam(beam, 1, 100, MINLEN) if beam > 1 else eng.decode_greedy(100)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("step %d (%s): encode %.2f ms, decode %.2f ms, %d launches" % (
        it, ["eager", "graph capture", "graph replay", "profiled, 1 stream"][it], 1e3 * (t1 - t0), 1e3 * (t2 - t1),
        eng.launch_count))
prof = eng.profile_read()
tot = sum(v[0] for v in prof.values())
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("  %-11s %8.2f ms  %5d launches  %7.1f us/launch  %5.1f %%" % (k, ms, n, 1e3 * ms / n, 100 * ms / tot))
