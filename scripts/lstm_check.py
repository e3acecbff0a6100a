"""LSTM tensor-core recurrence: parity vs the oracle for one variant + timing at B=1024.
Usage: python scripts/lstm_check.py <variant 0|1> [B]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanodecoder_b200 import synth, _lib
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
from oracle.model import OracleModel

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = ModelConfig.family("l2t", d_model=256, d_ff=128, enc_layers=3, dec_layers=1)
sd = synth.make_state_dict(cfg)
B, T = 70, 160
chunks, lengths = synth.make_chunks(B, T=T, seed=3, ragged=True, read_len=2)
lengths[0] = T
lengths[5] = 1
chunks[5, 1:] = 0
order = torch.argsort(lengths, descending=True, stable=True)
chunks, lengths = chunks[order], lengths[order]
eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=4)
eng.set_option("lstm_variant", variant)
eng.encode(chunks.cuda(), lengths.cuda())
mb, lens = eng.memory_bank()
torch.cuda.synchronize()
om = OracleModel(sd, cfg)
with torch.no_grad():
    _, want, wl = om.encoder(chunks.t().contiguous().unsqueeze(2), lengths)
err = float((mb.cpu().double() - want.double()).abs().max() / want.double().abs().max())
print("variant %d: nano encoder d=256 B=%d rel err %.3e  %s" % (variant, B, err, "OK" if err < 1e-3 else "FAIL"))
del eng

B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = ModelConfig.family("l2t", enc_layers=2, dec_layers=1)
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=B, max_src_len=512, max_tgt_len=4)
eng.set_option("lstm_variant", variant)
lib = _lib.load()
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
chunks, lengths = synth.make_chunks(B, T=512, seed=1, ragged=False)
src, lens = chunks.cuda(), lengths.cuda()
for rep in range(2):
    buf.zero_()
    lib.nd_debug_gemm_timeline(C.c_void_p(buf.data_ptr()))
    eng.encode(src, lens)
    torch.cuda.synchronize()
    lib.nd_debug_gemm_timeline(C.c_void_p(0))
t = buf.cpu().tolist()[16:]
names = ["step start", "mma issued+commit", "pre loads issued", "mma_done seen", "G exchanged", "pointwise+stores", "proxy fence", "cluster arrive"]
for i, n in enumerate(names):
    print("%8d cyc  %s" % (t[i] - t[0], n))
eng.profile_enable(["lstm"])
eng.encode(src, lens)
print(eng.profile_read())
