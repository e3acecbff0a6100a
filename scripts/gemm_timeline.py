"""clock64() timeline of CTA 0 of the tcgen05 GEMM for the decode shapes (tuning aid)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanodecoder_b200 import synth, _lib
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=8, max_src_len=64, max_tgt_len=4, gemm_mode="3xtf32")
lib = _lib.load()
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
names = {0: "entry", 1: "setup done", 2: "tma: first issued", 3: "tma: all issued", 4: "mma: first conv_full",
         5: "mma: last commit", 6: "conv: first raw_full", 7: "conv: first arrive", 8: "conv: all done",
         9: "epi: tmem_full", 10: "epi: stored/parked", 11: "cluster sync 1", 12: "reduce+store done",
         13: "cluster sync 2", 14: "dealloc"}
for (M, N, K, ln) in [(1024, 256, 256, False), (1024, 768, 256, True), (1024, 2048, 256, True), (1024, 256, 2048, False)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    g = torch.ones(K, device="cuda"); be = torch.zeros(K, device="cuda")
    for rep in range(3):
        buf.zero_()
        lib.nd_debug_gemm_timeline(C.c_void_p(buf.data_ptr()))
        eng.test_gemm("3xtf32", A, W, bias=b, ln=(g, be) if ln else None)
        torch.cuda.synchronize()
        lib.nd_debug_gemm_timeline(C.c_void_p(0))
    t = buf.cpu().tolist()
    print("M=%d N=%d K=%d ln=%d" % (M, N, K, ln))
    for i in sorted(names, key=lambda i: t[i] if t[i] else 1 << 62):
        if t[i]:
            print("   %8d cyc  %s" % (t[i] - t[0], names[i]))
