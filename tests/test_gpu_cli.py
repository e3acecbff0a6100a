"""GPU: the translate.py command line end to end (raw int16 reads -> front end -> encoder -> greedy decode ->
result/<read>.fasta, segment/<read>.txt, speed.txt), checked against the oracle chain on the same reads."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from nanodecoder_b200 import checkpoint, synth
from nanodecoder_b200.config import ModelConfig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("stride", [128, 96])
def test_translate_cli_matches_oracle_chain(tmp_path, stride):
    from oracle import decode as od
    from oracle import frontend as ofe
    from oracle.model import OracleModel
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=2, dec_layers=2)
    ckpt = str(tmp_path / "m.pt")
    checkpoint.save_checkpoint(synth.make_checkpoint(cfg, seed=5), ckpt)
    src_dir, out_dir = tmp_path / "reads", tmp_path / "out"
    src_dir.mkdir()
    rng = np.random.RandomState(3)
    reads = {}
    for i, n in enumerate([700, 129, 128, 1000]):
        raw = np.clip(np.round(rng.normal(500, 80, size=n)), 0, 2047).astype(np.int16)
        reads["read%d" % i] = raw
        (src_dir / ("read%d.signal" % i)).write_text(" ".join(str(int(v)) for v in raw))
    L = 24
    cmd = [sys.executable, os.path.join(ROOT, "translate.py"), "-model", ckpt, "-src_dir", str(src_dir), "-save_data",
           str(out_dir), "-src_seq_length", "128", "-src_seq_stride", str(stride), "-beam_size", "1", "-max_length",
           str(L), "-batch_size", "5", "-gpu", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]

    from nanodecoder_b200.utils.labelop import index2base, simple_assembly
    om = OracleModel(synth.make_state_dict(cfg, seed=5), cfg)
    speed = (out_dir / "speed.txt").read_text().strip().splitlines()
    assert len(speed) == len(reads)
    for name, raw in reads.items():
        # the reference translates read by read, in consecutive batches of batch_size chunks (translate.py:113-120)
        chunks = ofe.chunk(ofe.normalise(raw, "median"), 128, stride)
        want_preds = [None] * len(chunks)
        for src, lens, idx in ofe.make_batches(chunks, 5):
            o = od.greedy(om, torch.from_numpy(src), torch.from_numpy(lens), max_length=L)
            for j, i in enumerate(idx):
                want_preds[int(i)] = [" ".join(od.build_target_tokens(o["predictions"][j], cfg.vocab))]
        seg = (out_dir / "segment" / (name + ".txt")).read_text().splitlines()
        assert seg == [p[0] for p in want_preds], name
        if stride < 128:
            want = index2base(np.argmax(simple_assembly(want_preds), axis=0))
        else:
            want = simple_assembly(want_preds, flag_intersection=False)
        fasta = (out_dir / "result" / (name + ".fasta")).read_text()
        assert fasta == ">%s\n%s" % (name, want), name
        row = [ln for ln in speed if ln.split("\t")[0] == name][0].split("\t")
        assert int(row[2]) == len(want)
