"""Shared helpers for the parity tests (test infrastructure)."""
import ast
import os

import numpy as np
import torch

from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["l2t_d256", "t2t_d256", "nano2rnn_d256", "brnn2rnn_d256", "cnn2cnn_d256", "l2t_d64", "t2t_d64",
                "t2t_d512_6x6", "nano2rnn_general_d64", "brnn2rnn_dot_d64", "t2t_pe_d64", "nano2rnn_pe_d64",
                "cnn2cnn_pe_d64", "brnn2rnn_std_d256", "brnn2rnn_std_general_d64",
               "rnn2rnn_d256", "rnn2rnn_d64", "nano2rnn_gru_d64", "brnn2rnn_gru_d256", "l2t_gru_d64",
               "rnn2rnn_gru_std_d64", "resnet2t_d256", "resnet2rnn_d256", "resnet2t_d64", "crnn2t_d64", "crnn2rnn_d64",
                "ctrans2t_d64", "brnn2rnn_bridge_d64", "rnn2rnn_gru_bridge_d64", "t2t_avg_d64", "l2t_avg_d256"]


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


def check_beam_against_golden(g, ids, lens, scores, atol=5e-3, tie=1e-3):
    """ids [B,n_best,L], lens [B,n_best], scores [B,n_best] (numpy) vs a beam_* golden of the reference's --fast beam.
    north_star: "beam outputs match ... any divergence explained by a logit tie": a hypothesis may differ from the
    reference's only if its cumulative score equals the reference's to within `tie` = 1e-3: a 100-token score of ~ -130
    carries ~2e-4 of accumulated fp32 rounding (per-step log-probs agree to ~2e-6), and e.g. the reference's own 2nd and
    3rd hypotheses of beam_nano2rnn_d256_min99 chunk 2 are -131.399216 and -131.399231 -- ONE ulp apart, they swap places
    under any reassociation of the sums -- while distinct hypotheses are >= 0.05 apart.  -> number of such tie swaps."""
    B, NB = g["beam_ids"].shape[:2]
    swaps = 0
    for i in range(B):
        for n in range(NB):
            want = g["beam_ids"][i, n]
            want = want[want >= 0]
            assert len(want) > 1                                  # the point of these cases: non-degenerate hypotheses
            got = ids[i, n, : int(lens[i, n])]
            if len(got) == len(want) and (got == want).all():
                continue
            gap = abs(float(scores[i, n]) - float(g["beam_scores"][i, n]))
            assert gap <= tie, "chunk %d hyp %d differs and is no tie (score gap %.3g):\n%s\n%s" % (i, n, gap, got, want)
            swaps += 1
    np.testing.assert_allclose(scores[:, :NB], g["beam_scores"], atol=atol)
    assert swaps <= max(1, (B * NB) // 10), "%d of %d hypotheses are tie swaps" % (swaps, B * NB)
    return swaps


OBJ_EXTRA_CASES = ["objx_l2t_d64_ngram3", "objx_l2t_d64_ngram3_ignoreA", "objx_l2t_d64_covwu", "objx_nano2rnn_d64_covsummary", "objx_nano2rnn_d64_covwu", "objx_l2t_d64_stepwise_wu",
                                  "objx_nano2rnn_d64_stepwise_summary"]


def load_case_npz(name):
    """goldens that carry their own inputs (beam_attn_*, objx_*): -> (g, cfg, sd, src [B,T], lengths)"""
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = ModelConfig(**ast.literal_eval(str(g["cfg_json"])))
    sd = synth.make_state_dict(cfg, seed=int(g["weight_seed"]))
    return g, cfg, sd, torch.from_numpy(g["src"]), torch.from_numpy(g["lengths"])
