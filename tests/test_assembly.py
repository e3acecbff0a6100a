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
