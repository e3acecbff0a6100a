#!/bin/bash
# round 2, 2-GPU call: bench.py under torchrun (read-sharded e2e with the NCCL record gather) and translate.py under
# torchrun against a single-process run on the same reads
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02g_bench_n2.json 2> $O/r02g_bench_n2.err; echo "bench n2 exit $?"; tail -3 $O/r02g_bench_n2.err; cut -c1-1800 $O/r02g_bench_n2.json
timeout 600 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r02g_bench_ref.json 2> $O/r02g_bench_ref.err; echo "bench ref exit $?"; cut -c1-500 $O/r02g_bench_ref.json
D=/tmp/cli2; rm -rf $D; mkdir -p $D/reads
python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from nanodecoder_b200 import checkpoint, synth
from nanodecoder_b200.config import ModelConfig
cfg = ModelConfig.family("l2t")
checkpoint.save_checkpoint(synth.make_checkpoint(cfg, seed=2025), "/tmp/cli2/m.pt")
rng = np.random.RandomState(0)
for i in range(48):
    n = int(rng.randint(20000, 60000))
    raw = np.clip(np.round(rng.normal(500, 80, size=n)), 0, 2047).astype(np.int16)
    open("/tmp/cli2/reads/read%03d.signal" % i, "w").write(" ".join(map(str, raw.tolist())))
open("/tmp/cli2/reads/read900.signal", "w").write("12 zz 7")          # corrupt: reported, skipped
open("/tmp/cli2/reads/read901.signal", "w").write(" ".join("%.3f" % v for v in rng.normal(90, 12, size=30000)))   # float-valued
PY
ARGS="-model $D/m.pt -src_dir $D/reads -src_seq_length 512 -src_seq_stride 256 -beam_size 1 -max_length 100 -batch_size 1024 -thread 4"
SECONDS=0; python translate.py $ARGS -save_data $D/out1 -gpu 0 2>&1 | tail -2; echo "1 process: ${SECONDS}s"
SECONDS=0; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 translate.py $ARGS -save_data $D/out2 2>&1 | tail -2; echo "2 ranks: ${SECONDS}s"
python - <<'PY'
import glob, os
a = sorted(glob.glob("/tmp/cli2/out1/result/*.fasta")); b = sorted(glob.glob("/tmp/cli2/out2/result/*.fasta"))
print("fasta files", len(a), len(b))
same = sum(open(x).read() == open(x.replace("out1", "out2")).read() for x in a if os.path.exists(x.replace("out1", "out2")))
print("identical fasta", same)
s1 = [l.split("\t")[0::2] for l in open("/tmp/cli2/out1/speed.txt")]; s2 = [l.split("\t")[0::2] for l in open("/tmp/cli2/out2/speed.txt")]
print("speed.txt rows", len(s1), len(s2), "same reads and base counts:", s1 == s2)
PY
