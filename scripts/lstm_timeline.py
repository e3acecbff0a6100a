import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanodecoder_b200 import synth, _lib
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
cfg = ModelConfig.family("l2t", enc_layers=1, dec_layers=1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=B, max_src_len=512, max_tgt_len=4)
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
names = ["step start", "mma issued+commit", "xin loads issued", "mma_done seen", "G exchanged", "pointwise+stores", "proxy fence", "cluster.sync done"]
for i, n in enumerate(names):
    print("%8d cyc  %s" % (t[i] - t[0], n))
eng.profile_enable(["lstm"])
eng.encode(src, lens)
print(eng.profile_read())
