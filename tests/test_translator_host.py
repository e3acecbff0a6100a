"""Host logic of the Translator mirror (translate/translator.py:181-369 of the reference) without a GPU: batch plan,
padding widths, result order, hypothesis cutting and string building, driven through a recording stub in place of the
engine.  The stub is NOT an arithmetic fallback: it returns a deterministic function of what it was given, so the test
can tell which chunk went where with which padded width."""
import numpy as np
import pytest
import torch

from nanodecoder_b200.inputters.nano_dataset import reference_pad_lengths
from nanodecoder_b200.opts import default_translate_opt
from nanodecoder_b200.translate.translator import Translator, _Field
from nanodecoder_b200.config import ModelConfig

VOCAB = ["<unk>", "<blank>", "<s>", "</s>", "A", "C", "G", "T"]
EOS = 3


class _Vocab(object):
    def __init__(self, itos):
        self.itos = list(itos)
        self.stoi = {t: i for i, t in enumerate(itos)}


def _tokens(chunk: np.ndarray, length: int, width: int, L: int) -> np.ndarray:
    """What the stub 'decodes' for one chunk: depends on its samples, its length AND the padded width it arrived with."""
    key = int(np.abs(chunk[:length]).sum() * 64) + 7 * length + 13 * width
    n = 3 + key % (L - 3)
    ids = 4 + (key // (1 + np.arange(L))) % 4
    ids[n:] = EOS
    return ids.astype(np.int64)


class RecordingEngine(object):
    device = torch.device("cpu")

    def __init__(self, L):
        self.L = L
        self.calls = []

    def encode(self, src, lengths):
        assert src.dim() == 2 and src.is_contiguous() and lengths.shape == (src.shape[0],)
        assert (lengths[:-1] >= lengths[1:]).all(), "batches must be sorted by length, longest first"
        self.src, self.lengths = src.numpy().copy(), lengths.numpy().copy()
        self.calls.append(tuple(src.shape))

    def _ids(self):
        B, T = self.src.shape
        return np.stack([_tokens(self.src[i], int(self.lengths[i]), T, self.L) for i in range(B)])

    def decode_greedy(self, max_length, min_length=0, return_attn=False, return_logits=False):
        ids = self._ids()
        return {"ids": torch.from_numpy(ids), "scores": torch.from_numpy(-ids.sum(1).astype(np.float32)), "attn": None}

    def decode_beam(self, K, n_best, max_length, min_length=0, alpha=0.0, return_attn=False):
        ids = self._ids()
        B, L = ids.shape
        out = np.full((B, n_best, L), -1, dtype=np.int64)
        lens = np.zeros((B, n_best), dtype=np.int32)
        for i in range(B):
            n = int((ids[i] != EOS).sum())
            for k in range(n_best):                      # hypothesis k = the greedy one shortened by k tokens, then </s>
                m = max(n - k, 0)
                out[i, k, :m] = ids[i, :m]
                out[i, k, m] = EOS
                lens[i, k] = m + 1
        scores = -np.arange(n_best, dtype=np.float32)[None, :] - lens
        return {"ids": torch.from_numpy(out), "lens": torch.from_numpy(lens), "scores": torch.from_numpy(scores)}


def _expected(chunks, lengths, widths, L):
    out = []
    for i in range(len(lengths)):
        ids = _tokens(chunks[i].numpy(), int(lengths[i]), int(widths[i]), L)
        out.append(" ".join(VOCAB[t] for t in ids[: int((ids != EOS).sum())]))
    return out


def _make(n, T, seed):
    rng = np.random.default_rng(seed)
    lengths = rng.integers(5, T + 1, size=n)
    lengths[0] = T
    chunks = np.zeros((n, T), dtype=np.float32)
    for i in range(n):
        chunks[i, : lengths[i]] = np.round(rng.standard_normal(lengths[i]) * 64) / 64
    return torch.from_numpy(chunks), torch.from_numpy(lengths.astype(np.int64))


def _translator(L, beam=1, n_best=1, batch_size=8):
    opt = default_translate_opt(beam_size=beam, fast=beam > 1, n_best=n_best, batch_size=batch_size, max_length=L,
                                min_length=0, src_seq_length=64, gpu=0)
    eng = RecordingEngine(L)
    return Translator(eng, {"tgt": _Field(_Vocab(VOCAB))}, opt, ModelConfig.family("l2t")), eng


def test_results_come_back_in_input_order_over_several_batches():
    L = 20
    tr, eng = _translator(L, batch_size=8)
    chunks, lengths = _make(21, 48, seed=1)
    scores, preds = tr.translate(src=(chunks, lengths), batch_size=8)
    # read-by-read semantics: consecutive groups of 8, each padded to ITS longest chunk (inputter.py:86-95)
    widths = reference_pad_lengths(lengths.numpy(), 8)
    assert [p[0] for p in preds] == _expected(chunks, lengths, widths, L)
    assert eng.calls == [(8, int(widths[0])), (8, int(widths[8])), (5, int(widths[16]))]
    assert all(len(s) == 1 for s in scores)


def test_pooled_chunks_keep_the_padding_width_of_their_own_read():
    L = 16
    tr, eng = _translator(L, batch_size=16)
    reads = [_make(n, 40, seed=10 + n) for n in (5, 9, 3)]
    reads[1][1][0] = 33                                   # this read's longest chunk is shorter: width 33, not 40
    reads[1][0][:, 33:] = 0
    reads[1][1].clamp_(max=33)
    chunks = torch.cat([r[0] for r in reads])
    lengths = torch.cat([r[1] for r in reads])
    pad_to = np.concatenate([reference_pad_lengths(r[1].numpy(), 16) for r in reads])
    _, preds = tr.translate(src=(chunks, lengths, pad_to), batch_size=16)
    assert [p[0] for p in preds] == _expected(chunks, lengths, pad_to, L)
    assert sorted(c[1] for c in eng.calls) == sorted(set(int(w) for w in pad_to))


def test_reference_wire_format_strings_are_accepted():
    L = 12
    tr, _ = _translator(L, batch_size=4)
    chunks, lengths = _make(6, 30, seed=3)
    src = [" ".join(repr(float(v)) for v in chunks[i, : lengths[i]]) for i in range(6)]
    _, preds = tr.translate(src=src, batch_size=4)
    widths = reference_pad_lengths(lengths.numpy(), 4)
    assert [p[0] for p in preds] == _expected(chunks, lengths, widths, L)


def test_beam_n_best_hypotheses_are_cut_at_eos_and_padding():
    L = 14
    tr, _ = _translator(L, beam=5, n_best=3, batch_size=8)
    chunks, lengths = _make(11, 32, seed=4)
    scores, preds = tr.translate(src=(chunks, lengths), batch_size=8)
    widths = reference_pad_lengths(lengths.numpy(), 8)
    best = _expected(chunks, lengths, widths, L)
    for i in range(11):
        toks = best[i].split()
        assert preds[i] == [" ".join(toks[: max(len(toks) - k, 0)]) for k in range(3)]
        assert [float(s) for s in scores[i]] == [-(k + max(len(toks) - k, 0) + 1) for k in range(3)]


def test_gold_scoring_and_missing_batch_size_are_refused():
    tr, _ = _translator(8)
    chunks, lengths = _make(2, 16, seed=5)
    with pytest.raises(ValueError):
        tr.translate(src=(chunks, lengths))
    with pytest.raises(ValueError):
        tr.translate(src=(chunks, lengths), tgt=["A C"], batch_size=2)


def test_object_beam_flags_reach_the_engine_and_are_refused_where_the_reference_asserts():
    """-block_ngram_repeat / -ignore_when_blocking / -coverage_penalty / -beta: passed to the object beam (translator.py:
    836-848, beam.py:181-199); the reference asserts block_ngram_repeat == 0 in its greedy and --fast paths (:411, :633)."""
    base = dict(n_best=1, batch_size=8, max_length=8, min_length=0, src_seq_length=64, gpu=0)
    for kw in (dict(beam_size=1, fast=False), dict(beam_size=4, fast=True)):
        opt = default_translate_opt(block_ngram_repeat=3, **kw, **base)
        with pytest.raises(ValueError, match="object beam"):
            Translator(RecordingEngine(8), {"tgt": _Field(_Vocab(VOCAB))}, opt, ModelConfig.family("l2t"))
    opt = default_translate_opt(beam_size=4, fast=False, block_ngram_repeat=3, ignore_when_blocking=["A", "T"],
                                coverage_penalty="wu", beta=0.25, **base)
    tr = Translator(RecordingEngine(8), {"tgt": _Field(_Vocab(VOCAB))}, opt, ModelConfig.family("l2t"))
    assert tr._object_beam_extras() == dict(block_ngram_repeat=3, exclude_ids=[VOCAB.index(t) for t in ("A", "T")] if
                                            list(tr.ignore_when_blocking) == ["A", "T"] else
                                            [VOCAB.index(t) for t in tr.ignore_when_blocking],
                                            coverage_penalty="wu", beta=0.25, stepwise_penalty=False)


class _Log(object):
    def __init__(self):
        self.lines = []

    def info(self, msg):
        self.lines.append(msg)


@pytest.mark.parametrize("beam,n_best", [(1, 1), (5, 2)])
def test_array_path_equals_the_translation_builder_path(beam, n_best):
    """translate() builds the strings with array operations; with -verbose it goes through translate_batch and the
    TranslationBuilder mirror (translate/translation.py:27-41,62-119) one object per chunk.  Same predictions, scores."""
    L = 15
    chunks, lengths = _make(13, 36, seed=7)
    tr, _ = _translator(L, beam=beam, n_best=n_best, batch_size=8)
    s_fast, p_fast = tr.translate(src=(chunks, lengths), batch_size=8)
    tr2, _ = _translator(L, beam=beam, n_best=n_best, batch_size=8)
    tr2.verbose, tr2.logger = True, _Log()
    s_slow, p_slow = tr2.translate(src=(chunks, lengths), batch_size=8)
    assert p_fast == p_slow
    assert [[float(x) for x in s] for s in s_fast] == [[float(x) for x in s] for s in s_slow]
    assert len(tr2.logger.lines) == 13 and "PRED" in tr2.logger.lines[0]


def test_attention_dump_has_the_reference_layout():
    """-attn_debug block of one chunk (reference translate/translator.py:284-335): "{:>8.7} " cells for the source
    samples and the predicted tokens (both headers on one line, as the reference writes them), "{:>8.5f} " weights."""
    from nanodecoder_b200.translate.translator import format_attention
    txt = format_attention([0.5, -1.234375, 12.0], ["A", "C"], [[0.25, 0.75, 0.0], [1.0, 0.0, 0.0]])
    assert txt == ("       >      0.5  -1.2343     12.0 " "       |        A        C     </s> \n"
                   " 0.25000  0.75000  0.00000 \n" " 1.00000  0.00000  0.00000 \n")


def test_flags_outside_the_path_are_refused_with_the_reason():
    """flags the reference would act on but this engine does not implement must not be ignored silently"""
    for kw, word in ((dict(fft=True), "-fft"), (dict(dump_beam="beams.json"), "dump_beam"), (dict(replace_unk=True), "replace_unk"),
                     (dict(random_sampling_topk=5), "random sampling")):
        opt = default_translate_opt(beam_size=1, max_length=8, src_seq_length=64, gpu=0, **kw)
        with pytest.raises(ValueError, match=word):
            Translator(RecordingEngine(8), {"tgt": _Field(_Vocab(VOCAB))}, opt, ModelConfig.family("l2t"))
