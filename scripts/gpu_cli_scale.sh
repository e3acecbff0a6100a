#!/bin/bash
# end-to-end CLI at scale: 64 synthetic reads x ~60k samples, overlapping windows (stride 256 of 512)
echo start
D=/tmp/cli_scale; rm -rf $D; mkdir -p $D/reads
python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from nanodecoder_b200 import checkpoint, synth
from nanodecoder_b200.config import ModelConfig
cfg = ModelConfig.family("l2t")
checkpoint.save_checkpoint(synth.make_checkpoint(cfg, seed=2025), "/tmp/cli_scale/m.pt")
rng = np.random.RandomState(0)
for i in range(64):
    n = int(rng.randint(30000, 90000))
    raw = np.clip(np.round(rng.normal(500, 80, size=n)), 0, 2047).astype(np.int16)
    open("/tmp/cli_scale/reads/read%03d.signal" % i, "w").write(" ".join(map(str, raw.tolist())))
PY
SECONDS=0; python translate.py -model $D/m.pt -src_dir $D/reads -save_data $D/out -src_seq_length 512 -src_seq_stride 256 -beam_size 1 -max_length 100 -batch_size 1024 -gpu 0 -thread 8 2>&1 | tail -3; echo "translate.py wall: ${SECONDS}s"

head -3 $D/out/speed.txt; wc -l $D/out/speed.txt; ls $D/out/result | wc -l
python - <<'PY'
import glob
tot = sum(len(open(f).read().split("\n")[1]) for f in glob.glob("/tmp/cli_scale/out/result/*.fasta"))
print("total bases", tot)
PY
