"""GPU: translate.py fed with .fast5 files (single-read gzip, single-read VBZ, one multi-read VBZ file) writes the same
result / segment files as the same reads fed as .signal text: the library's own HDF5 / inflate / zstd / streamvbyte
readers in front of the CUDA path (utils/labelop.py:199-217 of the reference reads both kinds into the same samples)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from nanodecoder_b200 import checkpoint, synth
from nanodecoder_b200.config import ModelConfig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(ckpt, src_dir, out_dir):
    cmd = [sys.executable, os.path.join(ROOT, "translate.py"), "-model", ckpt, "-src_dir", str(src_dir), "-save_data",
           str(out_dir), "-src_seq_length", "128", "-src_seq_stride", "96", "-beam_size", "1", "-max_length", "24",
           "-batch_size", "5", "-gpu", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_fast5_inputs_give_the_same_files_as_signal_text(tmp_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import h5_writer
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=2, dec_layers=2)
    ckpt = str(tmp_path / "m.pt")
    checkpoint.save_checkpoint(synth.make_checkpoint(cfg, seed=5), ckpt)
    rng = np.random.RandomState(7)
    reads = {"r%d" % i: np.clip(np.round(rng.normal(500, 80, size=n)), 0, 2047).astype(np.int16)
             for i, n in enumerate([700, 129, 128, 1000, 333])}
    as_text, as_fast5 = tmp_path / "text", tmp_path / "fast5"
    as_text.mkdir()
    as_fast5.mkdir()
    for name, raw in reads.items():
        (as_text / (name + ".signal")).write_text(" ".join(str(int(v)) for v in raw))
    (as_fast5 / "r0.fast5").write_bytes(h5_writer.make_fast5(reads["r0"], read_name="Read_10", chunk=256, filters=(2, 1)))
    (as_fast5 / "r1.fast5").write_bytes(h5_writer.make_fast5(reads["r1"], read_name="Read_11", chunk=64, filters=(32020,),
                                                             kw_vbz_version=1))
    (as_fast5 / "batch_0.fast5").write_bytes(h5_writer.make_multi_fast5({k: reads[k] for k in ("r2", "r3", "r4")}, chunk=200))
    _run(ckpt, as_text, tmp_path / "out_text")
    _run(ckpt, as_fast5, tmp_path / "out_fast5")
    for name in reads:
        for sub, ext in (("result", ".fasta"), ("segment", ".txt")):
            a = (tmp_path / "out_text" / sub / (name + ext)).read_text()
            b = (tmp_path / "out_fast5" / sub / (name + ext)).read_text()
            assert a == b and len(a) > 0, (name, sub)
    rows = sorted(ln.split("\t")[0] for ln in (tmp_path / "out_fast5" / "speed.txt").read_text().strip().splitlines())
    assert rows == sorted(reads)
