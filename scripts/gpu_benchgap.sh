#!/bin/bash
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig
from nanodecoder_b200.engine import Engine
cfg = ModelConfig.family("l2t"); sd = synth.make_state_dict(cfg)
B = 1024
eng = Engine(cfg, sd, max_batch=B, max_src_len=512, max_tgt_len=100)
chunks, lengths = synth.make_chunks(B, T=512, seed=1234, ragged=True, read_len=16)
order = torch.argsort(lengths, descending=True, stable=True)
src, lens = chunks[order].cuda(), lengths[order].cuda()
for _ in range(3):
    eng.encode(src, lens); eng.decode_greedy(100)
torch.cuda.synchronize()
N = 12
evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * N + 1)]
t0 = time.perf_counter()
evs[0].record()
for i in range(N):
    eng.encode(src, lens)
    evs[2 * i + 1].record()
    eng.decode_greedy(100)
    evs[2 * i + 2].record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.1f ms, total %.1f ms for %d steps" % (1e3 * (t1 - t0), 1e3 * (t2 - t0), N))
for i in range(N):
    print("step %2d: encode %.2f ms decode %.2f ms" % (i, evs[2 * i].elapsed_time(evs[2 * i + 1]), evs[2 * i + 1].elapsed_time(evs[2 * i + 2])))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'], d['e2e']['value'])"
