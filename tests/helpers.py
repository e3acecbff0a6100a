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
