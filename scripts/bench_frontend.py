"""Front-end throughput: synthetic reads of int16 N ~ U(2000, 200000), values ~ N(500, 80) clipped to
[0, 2047]; nd_frontend_stats (exact median + MAD) and nd_frontend_chunks (normalise + chunk gather) timed with CUDA
events; algorithmic bytes = 2 N read (x2: median and MAD both need the data) + 4 N len/stride written."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
from nanodecoder_b200.inputters.nano_dataset import chunk_table

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
eng = Engine(cfg, synth.make_state_dict(cfg), max_batch=8, max_src_len=64, max_tgt_len=4)
rng = np.random.RandomState(0)
lens = rng.randint(2000, 200001, size=n_reads).astype(np.int64)
N = int(lens.sum())
sig = torch.from_numpy(np.clip(np.round(rng.normal(500, 80, size=N)), 0, 2047).astype(np.int16)).cuda()
off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).cuda()
for (fast, L, S) in ((1, 512, 512), (1, 300, 60), (1, 304, 64), (0, 512, 512)):
    eng.set_option("frontend_fast", fast)
    print("frontend_fast =", fast)
    cr, cs = chunk_table(lens, L, S)
    crd, csd = torch.from_numpy(cr).cuda(), torch.from_numpy(cs).cuda()
    for rep in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        center, scale = eng.frontend_stats(sig, off, "median")
        e[1].record()
        chunks, clen = eng.frontend_chunks(sig, off, center, scale, crd, csd, L)
        e[2].record()
        torch.cuda.synchronize()
    t_stats, t_chunks = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    b_stats = 2.0 * N * 2                      # int16 read for the median and again for the MAD
    b_chunks = 2.0 * N * (L / S) + 4.0 * len(cr) * L
    print("reads %d, samples %.1f M, chunk %d/%d -> %d chunks" % (n_reads, N / 1e6, L, S, len(cr)))
    print("  stats : %7.3f ms  %7.1f GB/s algorithmic (%.0f M samples/s)" % (t_stats, b_stats / t_stats / 1e6, N / t_stats / 1e3))
    print("  chunks: %7.3f ms  %7.1f GB/s algorithmic" % (t_chunks, b_chunks / t_chunks / 1e6))
# CPU oracle (numpy float64) on a bounded sample, one thread
from oracle import frontend as ofe
t0 = time.perf_counter()
done = 0
raw = sig.cpu().numpy()
offs = off.cpu().numpy()
for r in range(min(n_reads, 64)):
    ofe.frontend(raw[offs[r]:offs[r + 1]], "median", 512, 512)
    done += int(lens[r])
dt = time.perf_counter() - t0
print("CPU oracle (numpy, 1 thread): %.1f M samples/s on %d reads" % (done / dt / 1e6, min(n_reads, 64)))
