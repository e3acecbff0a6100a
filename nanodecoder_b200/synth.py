"""Seeded synthetic checkpoints and synthetic signal for tests and benchmarks.

There is no network in the build / GPU environment, so every number in this repo is measured on
random-init weights of the named architecture and on synthetic raw signal.
The state dict uses the reference's exact key names and shapes (models/model_builder.py:65-214,
onmt/models/model_saver.py:105-115) so the same dict loads into the reference's own modules
(`oracle/make_golden.py` does that with ``load_state_dict``).

Values come from numpy PCG64 streams keyed by (seed, parameter name) so they do not depend on
generation order, torch version, or device.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np
import torch

from .config import ModelConfig


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


class _Builder:
    def __init__(self, seed: int):
        self.seed = seed
        self.sd: Dict[str, torch.Tensor] = {}

    def normal(self, key, shape, std):
        x = _rng(self.seed, key).standard_normal(shape).astype(np.float32) * np.float32(std)
        self.sd[key] = torch.from_numpy(x)
        return self.sd[key]

    def matrix(self, key, out_f, in_f, gain=1.0, extra=()):
        return self.normal(key, (out_f, in_f) + tuple(extra), gain / np.sqrt(in_f * max(1, int(np.prod(extra)))))

    def bias(self, key, n, std=0.05):
        return self.normal(key, (n,), std)

    def affine(self, prefix, n):
        """LayerNorm / BatchNorm affine: scale around 1, small shift."""
        w = 1.0 + 0.1 * _rng(self.seed, prefix + ".weight").standard_normal(n)
        b = 0.05 * _rng(self.seed, prefix + ".bias").standard_normal(n)
        self.sd[prefix + ".weight"] = torch.from_numpy(w.astype(np.float32))
        self.sd[prefix + ".bias"] = torch.from_numpy(b.astype(np.float32))

    def linear(self, prefix, out_f, in_f, bias=True, gain=1.0):
        self.matrix(prefix + ".weight", out_f, in_f, gain)
        if bias:
            self.bias(prefix + ".bias", out_f)

    def lstm(self, prefix, suffix, in_f, hidden, gain_ih=1.0, gates=4):
        """nn.LSTM (gates 4) / nn.GRU (gates 3) parameters of one direction of one layer"""
        self.matrix("%s.weight_ih%s" % (prefix, suffix), gates * hidden, in_f, gain_ih)
        self.matrix("%s.weight_hh%s" % (prefix, suffix), gates * hidden, hidden, 1.0)
        self.bias("%s.bias_ih%s" % (prefix, suffix), gates * hidden)
        self.bias("%s.bias_hh%s" % (prefix, suffix), gates * hidden)

    def mha(self, prefix, d, qk_gain=1.0, out_gain=1.0):
        self.linear(prefix + ".linear_keys", d, d, gain=qk_gain)
        self.linear(prefix + ".linear_values", d, d)
        self.linear(prefix + ".linear_query", d, d, gain=qk_gain)
        self.linear(prefix + ".final_linear", d, d, gain=out_gain)

    def ffn(self, prefix, d, ff, out_gain=1.0):
        self.linear(prefix + ".w_1", ff, d)
        self.linear(prefix + ".w_2", d, ff, gain=out_gain)
        self.affine(prefix + ".layer_norm", d)

    def batchnorm(self, prefix, n, scale=1.0):
        """eval-mode BatchNorm buffers + affine"""
        w = scale * (1.0 + 0.1 * _rng(self.seed, prefix + ".weight").standard_normal(n))
        self.sd[prefix + ".weight"] = torch.from_numpy(w.astype(np.float32))
        self.bias(prefix + ".bias", n)
        self.normal(prefix + ".running_mean", (n,), 0.05)
        rv = 0.8 + 0.2 * np.abs(_rng(self.seed, prefix + ".running_var").standard_normal(n))
        self.sd[prefix + ".running_var"] = torch.from_numpy(rv.astype(np.float32))
        self.sd[prefix + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)

    def resnet(self, prefix, num_classes, fc_gain=1.0):
        """encoder/resnet_encoder.py:119-152 (ResNet with BasicBlock, layers [2,2,2,2]): (5,3) kernels of which only row
        2 meets the height-1 signal; He-scaled for the 3 taps that act, second BatchNorm of a block damped so the
        residual stream keeps its scale over the 8 blocks."""
        self.normal(prefix + ".conv1.weight", (64, 1, 5, 3), np.sqrt(2.0 / 3))
        self.batchnorm(prefix + ".bn1", 64)
        inplanes = 64
        for li, planes in enumerate((64, 128, 256, 512)):
            for bi in range(2):
                bp = "%s.layer%d.%d" % (prefix, li + 1, bi)
                cin = inplanes if bi == 0 else planes
                self.normal(bp + ".conv1.weight", (planes, cin, 5, 3), np.sqrt(2.0 / (3 * cin)))
                self.bias(bp + ".conv1.bias", planes)
                self.batchnorm(bp + ".bn1", planes)
                self.normal(bp + ".conv2.weight", (planes, planes, 5, 3), np.sqrt(2.0 / (3 * planes)))
                self.bias(bp + ".conv2.bias", planes)
                self.batchnorm(bp + ".bn2", planes, scale=0.5)
                if bi == 0 and inplanes != planes:
                    self.normal(bp + ".downsample.0.weight", (planes, inplanes, 1, 1), np.sqrt(1.0 / inplanes))
                    self.batchnorm(bp + ".downsample.1", planes)
            inplanes = planes
        self.linear(prefix + ".fc", num_classes, 512, gain=fc_gain)

    def wnconv(self, prefix, d, width):
        """WeightNormConv2d (onmt/modules/weight_norm.py:101-169): weight==V, bias==b aliases and
        the Polyak buffers equal to the live values, as in a trained checkpoint, so the
        eval-time buffer update is exactly a no-op."""
        v = self.matrix(prefix + ".V", 2 * d, d, 1.0, extra=(width, 1))
        g = 1.0 + 0.1 * _rng(self.seed, prefix + ".g").standard_normal(2 * d)
        g = torch.from_numpy(g.astype(np.float32))
        b = self.bias(prefix + ".b", 2 * d)
        self.sd[prefix + ".g"] = g
        self.sd[prefix + ".weight"] = v.clone()
        self.sd[prefix + ".bias"] = b.clone()
        self.sd[prefix + ".V_avg"] = v.clone()
        self.sd[prefix + ".g_avg"] = g.clone()
        self.sd[prefix + ".b_avg"] = b.clone()


# Gains that make a random-init model signal-sensitive: sharp cross-attention and a weak
# previous-token embedding, so different chunks decode to different, non-constant sequences
# (plain 1/sqrt(fan_in) init collapses to one repeated token; SURVEY.md §7 "hard parts").
DEFAULT_GAINS = dict(emb_std=0.5, attn_qk=2.0, ctx_out=2.0, self_qk=1.0, self_out=1.0, ffn_out=1.0,
                     gen=3.0, eos_bias=-1.5)


# per-family gains found by a small random search (entropy of the greedy token histogram >= 1.4 bits
# and (nearly) all of 16 chunks decoding to different sequences at d=256, 3+3 layers, seed 2025)
FAMILY_GAINS = {
    ("nano", "transformer"): dict(emb_std=0.3, attn_qk=2.0, ctx_out=3.0, self_qk=2.0, self_out=4.0,
                                  ffn_out=2.0, gen=1.0, eos_bias=-0.07),
    ("transformer", "transformer"): dict(emb_std=0.3, attn_qk=1.0, ctx_out=1.0, self_qk=2.0,
                                         self_out=0.5, ffn_out=1.0, gen=2.0, eos_bias=-1.0),
    ("nano", "rnn"): dict(emb_std=0.1, attn_qk=2.5, ctx_out=6.0, gen=1.0, eos_bias=-0.5),
    ("brnn", "rnn"): dict(emb_std=1.0, attn_qk=1.5, ctx_out=2.0, gen=3.0, eos_bias=-1.0),
    ("rnn", "rnn"): dict(emb_std=1.0, attn_qk=1.5, ctx_out=2.0, gen=3.0, eos_bias=-1.0),
    ("cnn", "cnn"): dict(emb_std=1.0, attn_qk=2.5, ctx_out=2.0, gen=2.0, eos_bias=-0.5),
    # ResNet-stem encoders: random search over the same gains + the stem's output gain (fc), 16 chunks
    ("resnet", "transformer"): dict(stem_gain=2.95, emb_std=0.1, attn_qk=2.9, ctx_out=3.3, self_qk=2.14, self_out=2.24,
                                    ffn_out=1.53, gen=2.57, eos_bias=-0.88),
    ("resnet", "rnn"): dict(stem_gain=0.92, emb_std=0.3, attn_qk=2.41, ctx_out=5.72, gen=1.76, eos_bias=-0.86),
    ("crnn", "transformer"): dict(stem_gain=0.4, emb_std=1.0, attn_qk=1.4, ctx_out=1.45, self_qk=1.87, self_out=1.55,
                                  ffn_out=1.67, gen=1.4, eos_bias=-0.09),
    ("crnn", "rnn"): dict(stem_gain=0.1, emb_std=0.3, attn_qk=2.55, ctx_out=5.89, gen=2.35, eos_bias=-1.21),
    ("ctransformer", "transformer"): dict(stem_gain=2.41, emb_std=0.1, attn_qk=2.63, ctx_out=1.01, self_qk=2.29,
                                          self_out=0.62, ffn_out=1.73, gen=1.35, eos_bias=-0.21),
}


def rnn_decoder_keys(layer: int, input_feed) -> tuple:
    """(weight_ih, weight_hh, bias_ih, bias_hh) state-dict keys of decoder layer `layer`."""
    if input_feed:
        p = "decoder.rnn.layers.%d" % layer
        return (p + ".weight_ih", p + ".weight_hh", p + ".bias_ih", p + ".bias_hh")
    return tuple("decoder.rnn.%s_l%d" % (n, layer) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))


def positional_encoding_table(dim: int, max_len: int = 5000) -> torch.Tensor:
    """The registered buffer ``pe`` of the reference's PositionalEncoding (onmt/modules/embeddings.py:21-32), built
    with the same fp32 torch ops, so a synthetic checkpoint carries what a trained one would: [max_len, 1, dim]."""
    import math
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp((torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim)))
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe.unsqueeze(1)


def make_state_dict(cfg: ModelConfig, seed: int = 2025, gains: dict = None) -> Dict[str, torch.Tensor]:
    """Flat state dict: reference ``model.state_dict()`` keys incl. ``generator.0.*``."""
    b = _Builder(seed)
    G = dict(DEFAULT_GAINS)
    G.update(FAMILY_GAINS.get((cfg.encoder_type, cfg.decoder_type), {}))
    G.update(gains or {})
    d, V = cfg.d_model, cfg.vocab_size
    h = d // 2
    ng = 3 if cfg.rnn_type == "GRU" else 4       # gate rows per hidden unit

    # ---------------- encoder
    if cfg.encoder_type == "nano":              # encoder/nano_encoder.py:26-77
        b.linear("encoder.W", d, d, bias=False)
        for l in range(cfg.enc_layers):
            in_f = 1 if l == 0 else d
            for sfx in ("_l0", "_l0_reverse"):
                b.lstm("encoder.rnn_%d" % l, sfx, in_f, h, gain_ih=1.0, gates=ng)
            p = "encoder.batchnorm_%d" % l
            b.affine(p, d)
            b.normal(p + ".running_mean", (d,), 0.05)
            rv = 0.1 + 0.05 * np.abs(_rng(seed, p + ".running_var").standard_normal(d))
            b.sd[p + ".running_var"] = torch.from_numpy(rv.astype(np.float32))
            b.sd[p + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)
    elif cfg.encoder_type in ("rnn", "brnn"):   # encoder/rnn_encoder.py:23-62
        dirs = 2 if cfg.encoder_type == "brnn" else 1
        hh = d // dirs
        for l in range(cfg.enc_layers):
            in_f = 1 if l == 0 else d
            b.lstm("encoder.rnn", "_l%d" % l, in_f, hh, gates=ng)
            if dirs == 2:
                b.lstm("encoder.rnn", "_l%d_reverse" % l, in_f, hh, gates=ng)
        if cfg.bridge:                          # encoder/rnn_encoder.py:86-99: one Linear per state (h, c)
            tot = hh * cfg.enc_layers
            for i in range(2 if cfg.rnn_type == "LSTM" else 1):
                b.linear("encoder.bridge.%d" % i, tot, tot, gain=2.0)
    elif cfg.encoder_type == "transformer":     # encoder/transformer.py:87-104
        b.linear("encoder.linear", d, 1)
        for l in range(cfg.enc_layers):
            p = "encoder.transformer.%d" % l
            b.mha(p + ".self_attn", d)
            b.ffn(p + ".feed_forward", d, cfg.d_ff)
            b.affine(p + ".layer_norm", d)
        b.affine("encoder.layer_norm", d)
    elif cfg.encoder_type == "cnn":             # encoder/cnn_encoder.py:18-27
        b.linear("encoder.linear", d, 1)
        for l in range(cfg.enc_layers):
            b.wnconv("encoder.cnn.layers.%d.conv" % l, d, cfg.cnn_kernel_width)
    elif cfg.encoder_type == "resnet":          # encoder/resnet_encoder.py:214-216
        b.resnet("encoder.cnn", d, G.get("stem_gain", 1.0))
    elif cfg.encoder_type == "crnn":            # encoder/crnn_encoder.py:27-84: stem + nano stack, first layer d wide
        b.resnet("encoder.cnn", d, G.get("stem_gain", 1.0))
        b.linear("encoder.W", d, d, bias=False)
        for l in range(cfg.enc_layers):
            for sfx in ("_l0", "_l0_reverse"):
                b.lstm("encoder.rnn_%d" % l, sfx, d, h, gain_ih=1.0, gates=ng)
            b.batchnorm("encoder.batchnorm_%d" % l, d)
    elif cfg.encoder_type == "ctransformer":    # encoder/ctransformer.py:45-64
        b.resnet("encoder.cnn", d, G.get("stem_gain", 1.0))
        for l in range(cfg.enc_layers):
            p = "encoder.transformer.%d" % l
            b.mha(p + ".self_attn", d)
            b.ffn(p + ".feed_forward", d, cfg.d_ff)
            b.affine(p + ".layer_norm", d)
        b.affine("encoder.layer_norm", d)

    # ---------------- decoder
    b.normal("decoder.embeddings.make_embedding.emb_luts.0.weight", (V, d), G["emb_std"])
    if cfg.position_encoding:
        b.sd["decoder.embeddings.make_embedding.pe.pe"] = positional_encoding_table(d)
    if cfg.decoder_type == "transformer":       # decoder/transformer.py:145-171
        for l in range(cfg.dec_layers):
            p = "decoder.transformer_layers.%d" % l
            if cfg.self_attn_type == "average":  # onmt/modules/average_attn.py:22-30
                b.ffn(p + ".self_attn.average_layer", d, d, out_gain=G["self_out"])
                b.linear(p + ".self_attn.gating_layer", 2 * d, 2 * d)
            else:
                b.mha(p + ".self_attn", d, qk_gain=G["self_qk"], out_gain=G["self_out"])
            b.mha(p + ".context_attn", d, qk_gain=G["attn_qk"], out_gain=G["ctx_out"])
            b.ffn(p + ".feed_forward", d, cfg.d_ff, out_gain=G["ffn_out"])
            b.affine(p + ".layer_norm_1", d)
            b.affine(p + ".layer_norm_2", d)
        b.affine("decoder.layer_norm", d)
    elif cfg.decoder_type == "rnn":             # onmt/decoders/decoder.py:57-106,352-366
        for l in range(cfg.dec_layers):
            in_f = (2 * d if cfg.input_feed else d) if l == 0 else d
            # InputFeedRNNDecoder: StackedLSTM of LSTMCells (stacked_rnn.py:15-20); StdRNNDecoder (-input_feed 0):
            # one multi-layer nn.LSTM (decoder.py:264-266 via rnn_factory) -- same arithmetic, other parameter names
            names = rnn_decoder_keys(l, cfg.input_feed)
            b.matrix(names[0], ng * d, in_f)
            b.matrix(names[1], ng * d, d)
            b.bias(names[2], ng * d)
            b.bias(names[3], ng * d)
        # onmt/modules/global_attention.py:71-93
        if cfg.global_attention == "mlp":
            b.linear("decoder.attn.linear_context", d, d, bias=False)
            b.linear("decoder.attn.linear_query", d, d)
            b.linear("decoder.attn.v", 1, d, bias=False, gain=4.0 * G["attn_qk"])
            b.linear("decoder.attn.linear_out", d, 2 * d, bias=True, gain=G["ctx_out"])
        else:
            if cfg.global_attention == "general":
                b.linear("decoder.attn.linear_in", d, d, bias=False)
            b.linear("decoder.attn.linear_out", d, 2 * d, bias=False, gain=G["ctx_out"])
    elif cfg.decoder_type == "cnn":             # onmt/decoders/cnn_decoder.py:20-49
        b.linear("decoder.linear", d, d)
        for l in range(cfg.dec_layers):
            b.wnconv("decoder.conv_layers.%d.conv" % l, d, cfg.cnn_kernel_width)
            b.linear("decoder.attn_layers.%d.linear_in" % l, d, d, gain=G["attn_qk"])

    # ---------------- generator (models/model_builder.py:331-334)
    b.matrix("generator.0.weight", V, d, G["gen"])
    gb = b.bias("generator.0.bias", V, 0.2)
    # make the specials (<unk>, <blank>, <s>) unattractive so sequences look like bases + </s>
    gb[0:3] -= 6.0
    gb[3] += G["eos_bias"]
    return b.sd


def make_checkpoint(cfg: ModelConfig, seed: int = 2025) -> dict:
    """Checkpoint dict in the reference layout (onmt/models/model_saver.py:105-115)."""
    from .checkpoint import make_vocab_entry
    sd = make_state_dict(cfg, seed)
    return {
        "model": {k: v for k, v in sd.items() if not k.startswith("generator.")},
        "generator": {k[len("generator."):]: v for k, v in sd.items() if k.startswith("generator.")},
        "vocab": make_vocab_entry(cfg.vocab),
        "opt": cfg.to_opt(),
        "optim": None,
    }


# --------------------------------------------------------------------------- synthetic signal
def make_chunks(n_chunks: int, T: int = 512, seed: int = 1234, ragged: bool = True,
                read_len: int = 16):
    """Normalised signal chunks as the front end would emit them.

    -> (src [n_chunks, T] fp32 zero padded, lengths [n_chunks] int64).
    Values are N(0,1) quantised to 1/64 with 1 % of samples exactly 0.0 and 0.1 % exactly 1.0,
    so the value-equality attention masks of the reference are exercised
    (encoder/transformer.py:120-121, decoder/transformer.py:220-221).  With ``ragged`` the last
    chunk of each synthetic read of ``read_len`` chunks is short (64..T-1 samples).
    """
    rng = np.random.default_rng([seed, 7])
    x = np.round(rng.standard_normal((n_chunks, T)) * 64.0) / 64.0
    u = rng.random((n_chunks, T))
    x[u < 0.01] = 0.0
    x[u > 0.999] = 1.0
    lengths = np.full((n_chunks,), T, dtype=np.int64)
    if ragged:
        last = np.arange(read_len - 1, n_chunks, read_len)
        lengths[last] = rng.integers(64, T, size=last.shape[0])
        for i in last:
            x[i, lengths[i]:] = 0.0
    return torch.from_numpy(x.astype(np.float32)), torch.from_numpy(lengths)


def make_raw_reads(n_reads: int, seed: int = 99, min_len: int = 2000, max_len: int = 200000):
    """Synthetic int16 raw reads: N ~ U(min_len,max_len), values ~ N(500,80) clipped to [0,2047]."""
    rng = np.random.default_rng([seed, 11])
    lens = rng.integers(min_len, max_len + 1, size=n_reads)
    reads = []
    for n in lens:
        # slow level changes + noise, like a squiggle
        levels = rng.normal(500.0, 80.0, size=int(n) // 9 + 2)
        sig = np.repeat(levels, 9)[: int(n)] + rng.normal(0.0, 12.0, size=int(n))
        reads.append(np.clip(np.round(sig), 0, 2047).astype(np.int16))
    return reads
