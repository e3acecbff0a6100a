"""Shared helpers for the parity tests (test infrastructure)."""
import ast
import os

import numpy as np
import torch

from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["l2t_d256", "t2t_d256", "nano2rnn_d256", "brnn2rnn_d256", "cnn2cnn_d256", "l2t_d64", "t2t_d64",
                "t2t_d512_6x6", "nano2rnn_general_d64", "brnn2rnn_dot_d64"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = ModelConfig(**ast.literal_eval(str(g["cfg_json"])))
    sd = synth.make_state_dict(cfg, seed=int(g["weight_seed"]))
    src = torch.from_numpy(g["src"])            # [B,T] chunk-major, already in iterator order
    lengths = torch.from_numpy(g["lengths"])
    return g, cfg, sd, src, lengths


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


BEAM_GOLDEN_CASES = ["beam_l2t_d256_min99", "beam_l2t_d256_min20", "beam_l2t_d256_min20_alpha", "beam_t2t_d256_min99",
                     "beam_t2t_d256_min20", "beam_nano2rnn_d256_min99", "beam_nano2rnn_d256_min20",
                     "beam_brnn2rnn_d256_min99", "beam_brnn2rnn_d256_min20", "beam_cnn2cnn_d256_min99",
                     "beam_cnn2cnn_d256_min20", "beam_t2t_d512_6x6_min20"]


def check_beam_against_golden(g, ids, lens, scores, atol=5e-3):
    """ids [B,n_best,L], lens [B,n_best], scores [B,n_best] (numpy) vs a beam_* golden of the reference's --fast beam."""
    B, NB = g["beam_ids"].shape[:2]
    for i in range(B):
        for n in range(NB):
            want = g["beam_ids"][i, n]
            want = want[want >= 0]
            assert len(want) > 1                                  # the point of these cases: non-degenerate hypotheses
            np.testing.assert_array_equal(ids[i, n, : int(lens[i, n])], want, err_msg="chunk %d hyp %d" % (i, n))
    np.testing.assert_allclose(scores[:, :NB], g["beam_scores"], atol=atol)
