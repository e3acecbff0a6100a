"""CPU: read assembly (utils/labelop.py:320-352) — the libnanodec longest-block search against CPython's difflib and
the vectorised simple_assembly against golden outputs of the reference's own simple_assembly / index2base."""
import difflib

import numpy as np

from helpers import GOLDEN
from nanodecoder_b200.utils.labelop import index2base, longest_match, simple_assembly


def _difflib_longest(a, b):
    blocks = difflib.SequenceMatcher(None, a, b).get_matching_blocks()
    return tuple(max(blocks, key=lambda x: x[2]))


def test_longest_match_equals_difflib_including_ties_and_autojunk():
    rng = np.random.RandomState(1)
    cases = [("", ""), ("A", ""), ("", "A"), ("ACGT", "ACGT"), ("AAAA", "AA"), ("ACAC", "CACA"), ("ACGT", "TGCA"),
             ("ACGTACGTAC" * 30, "GTACGTACGT" * 25), ("A" * 250, "A" * 250), ("ACGT" * 60, "ACGT" * 60)]
    for _ in range(300):
        na, nb = int(rng.randint(0, 120)), int(rng.randint(0, 120))
        alpha = "ACGT" if rng.rand() < 0.8 else "AC"
        cases.append(("".join(rng.choice(list(alpha), size=na)), "".join(rng.choice(list(alpha), size=nb))))
    for _ in range(30):                                    # len(b) >= 200: popular elements leave the index
        a = "".join(rng.choice(list("ACGT"), size=int(rng.randint(150, 320))))
        cut = int(rng.randint(0, 60))
        b = a[cut:] + "".join(rng.choice(list("ACGT"), size=cut))
        cases.append((a, b))
    for a, b in cases:
        want = _difflib_longest(a, b)
        got = longest_match(a, b)
        assert got == want, (a[:40], b[:40], got, want)


def test_simple_assembly_matches_reference_golden():
    g = np.load(GOLDEN + "/assembly.npz")
    for ci in range(int(g["n_cases"])):
        bpreads = [[x] for x in str(g["case%d_input" % ci]).split("\n")]
        votes = simple_assembly(bpreads)
        np.testing.assert_array_equal(votes.astype(np.int32), g["case%d_votes" % ci])
        assert index2base(np.argmax(votes, axis=0)) == str(g["case%d_fasta" % ci])
        assert simple_assembly(bpreads, flag_intersection=False) == str(g["case%d_concat" % ci])


def _reference_restatement(bpreads):
    """simple_assembly + add_count of the reference (utils/labelop.py:311-352) with CPython's difflib, line for line in
    behaviour (including what numpy does when a chunk does not fit: IndexError; unknown base: KeyError)."""
    from nanodecoder_b200.utils.labelop import base_dict, base_keys
    valid = [x[0].replace(" ", "") for x in bpreads if x[0] != ""]
    conc = np.zeros([len(base_keys), 1000])
    pos, length, census_len = 0, 0, 1000

    def add_count(conc, start, seg):
        if start < 0:
            seg = seg[-start:]
            start = 0
        for i, base in enumerate(seg):
            conc[base_dict[base.upper()]][start + i] += 1

    for indx, bp in enumerate(valid):
        if indx == 0:
            add_count(conc, 0, bp)
            continue
        blocks = difflib.SequenceMatcher(None, valid[indx - 1], bp).get_matching_blocks()
        mb = max(blocks, key=lambda x: x[2])
        disp = mb[0] - mb[1]
        if disp + pos + len(bp) > census_len:
            conc = np.pad(conc, ((0, 0), (0, 1000)), mode="constant", constant_values=0)
            census_len += 1000
        add_count(conc, pos + disp, bp)
        pos += disp
        length = max(length, pos + len(bp))
    return conc[:, :length]


def test_simple_assembly_in_libnanodec_equals_the_difflib_restatement_on_noisy_reads():
    rng = np.random.RandomState(7)
    for case in range(12):
        n = int(rng.randint(1, 90))
        genome = "".join(rng.choice(list("ACGT"), size=n * 60 + 200))
        width, stride = (100, 60) if case % 3 else (300, 60)        # authors' setting: stride 60 of 300
        reads = []
        for i in range(n):
            s = np.array(list(genome[i * stride: i * stride + width]))
            m = rng.rand(len(s)) < 0.08
            s[m] = rng.choice(list("ACGT"), size=int(m.sum()))
            if rng.rand() < 0.1:
                s = s[: int(rng.randint(0, 20))]                     # short / empty chunk
            if rng.rand() < 0.05:
                s = np.char.lower(s)                                 # base.upper() in add_count
            reads.append([" ".join(s)])
        got, want = simple_assembly(reads), _reference_restatement(reads)
        assert got.dtype == want.dtype == np.float64
        np.testing.assert_array_equal(got, want)


def test_simple_assembly_edge_cases_behave_like_the_reference():
    assert simple_assembly([]).shape == (5, 0)
    assert simple_assembly([["A C G T"]]).shape == (5, 0)             # `length` only advances from the 2nd chunk on
    two = simple_assembly([["A C G T"], ["G T A A"]])
    np.testing.assert_array_equal(two, _reference_restatement([["A C G T"], ["G T A A"]]))
    neg = [["G T A C"], ["A A A A G T A C"]]                          # negative start: the head of the chunk is trimmed
    np.testing.assert_array_equal(simple_assembly(neg), _reference_restatement(neg))
    import pytest
    with pytest.raises(KeyError):
        simple_assembly([["A C"], ["A X"]])
    long_first = [[" ".join("ACGT" * 300)], ["A C G T"]]              # 1200 bases do not fit the initial 1000 columns
    with pytest.raises(IndexError):
        _reference_restatement(long_first)
    with pytest.raises(IndexError):
        simple_assembly(long_first)
