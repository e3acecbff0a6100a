"""Host post-processing: the vectorised token join equals the per-hypothesis join of the reference's
TranslationBuilder (translate/translation.py:27-41, translator.py:271-273)."""
import numpy as np

from nanodecoder_b200.translate.translator import join_tokens

ITOS = ["<unk>", "<blank>", "<s>", "</s>", "A", "C", "G", "T"]


def _slow(ids, cut):
    return [[" ".join(ITOS[t] for t in ids[j, n, : cut[j, n]]) for n in range(ids.shape[1])] for j in range(ids.shape[0])]


def test_single_character_tokens():
    rng = np.random.default_rng(0)
    ids = rng.integers(4, 8, size=(37, 3, 25))
    cut = rng.integers(0, 26, size=(37, 3))
    ids[np.arange(25)[None, None, :] >= cut[:, :, None]] = -1          # padding beyond the hypothesis
    assert join_tokens(ids, cut, ITOS) == _slow(ids, cut)


def test_special_tokens_inside_a_hypothesis_take_the_general_path():
    rng = np.random.default_rng(1)
    ids = rng.integers(0, 8, size=(9, 2, 12))
    cut = rng.integers(0, 13, size=(9, 2))
    assert join_tokens(ids, cut, ITOS) == _slow(ids, cut)


def test_empty_hypotheses_and_zero_length():
    ids = np.full((4, 1, 6), 4)
    cut = np.array([[0], [1], [6], [0]])
    assert join_tokens(ids, cut, ITOS) == [[""], ["A"], ["A A A A A A"], [""]]
    assert join_tokens(np.zeros((2, 1, 0), dtype=np.int64), np.zeros((2, 1), dtype=np.int64), ITOS) == [[""], [""]]
