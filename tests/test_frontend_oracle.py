"""CPU: front-end oracle vs the golden output of the reference's extract_fast5_raw, and the
restated statsmodels MAD vs scipy's independent implementation."""
import numpy as np
import scipy.stats

from helpers import GOLDEN
from oracle import frontend as ofe
from nanodecoder_b200.inputters.nano_dataset import pooled_batches, reference_pad_lengths, batch_order, chunk_table, parse_segments


def _golden():
    return np.load(GOLDEN + "/frontend.npz")


def test_frontend_oracle_matches_reference_golden():
    g = _golden()
    for ri in range(int(g["n_reads"])):
        raw = g["raw_%d" % ri]
        for norm in ("median", "mean"):
            for (L, S) in ((512, 512), (300, 60)):
                key = "r%d_%s_%d_%d" % (ri, norm, L, S)
                chunks = ofe.chunk(ofe.normalise(raw.astype(np.float64), norm), L, S)
                np.testing.assert_array_equal([len(c) for c in chunks], g[key + "_lens"])
                flat = np.concatenate(chunks)
                assert abs(flat.sum() - float(g[key + "_sum"])) <= 1e-9 * max(1.0, np.abs(flat).sum())
                f32 = flat.astype(np.float32)
                want = g[key + "_flat"]
                got = f32 if f32.size < 6000 else f32[::7]
                np.testing.assert_array_equal(got, want)             # bit exact
                # the host-side chunk table reproduces the same segmentation
                cr, cs = chunk_table([raw.size], L, S)
                np.testing.assert_array_equal(np.minimum(cs + L, raw.size) - cs, g[key + "_lens"])


def test_mad_restatement_against_scipy():
    rng = np.random.default_rng(0)
    for n in (5, 6, 1001, 4096):
        x = np.round(rng.normal(500, 80, n)).astype(np.int16).astype(np.float64)
        a = ofe.mad(x)
        b = scipy.stats.median_abs_deviation(x, scale="normal")
        assert abs(a - b) <= 4e-16 * abs(b) + 1e-300 or abs(a - b) / abs(b) < 1e-15


def test_batch_order_and_parse_segments():
    lens = np.array([512, 512, 100, 512, 7, 512])
    order = batch_order(lens, 4)
    assert [o.tolist() for o in order] == [[0, 1, 3, 2], [5, 4]]
    chunks, l = parse_segments(["0.5 1.0 -2.25", "3.0"])
    assert chunks.shape == (2, 3) and l.tolist() == [3, 1] and float(chunks[1, 1]) == 0.0
    src, lengths, idx = ofe.make_batches([np.zeros(5), np.ones(9), np.ones(5)], 3)[0]
    assert idx.tolist() == [1, 0, 2] and src.shape == (9, 3, 1)


def test_pooled_batches_keep_reference_padding_width():
    """Chunks pooled over reads are padded exactly as in the reference's read-by-read batches."""
    read_a = [128, 128, 128, 128, 128, 60]           # batch_size 5 -> groups [128 x5] and [60]
    read_b = [128, 77]
    pad = np.concatenate([reference_pad_lengths(read_a, 5), reference_pad_lengths(read_b, 5)])
    assert pad.tolist() == [128] * 5 + [60] + [128, 128]
    lens = np.array(read_a + read_b)
    plan = pooled_batches(lens, pad, 4)
    assert [(i.tolist(), w) for i, w in plan] == [([0, 1, 2, 3], 128), ([4, 6, 7], 128), ([5], 60)]
    for src, l, idx in ofe.make_batches([np.zeros(n) for n in read_a], 5):
        assert all(src.shape[0] == pad[i] for i in idx)
