"""Model configuration for the translate path.

Mirrors the model flags the reference stores in ``checkpoint['opt']`` (reference
``models/opts.py:15-200``) and the dispatch in ``models/model_builder.py:65-214``.
Only the flags that change the inference arithmetic are kept.
"""
from __future__ import annotations

import argparse
from dataclasses import dataclass, field, asdict
from typing import List

# vocabulary order produced by the reference's field builder
# (inputters/inputter.py:315-320, onmt/inputters/dataset_base.py:10-13)
SPECIALS = ["<unk>", "<blank>", "<s>", "</s>"]
UNK, PAD, BOS, EOS = 0, 1, 2, 3

ENCODER_TYPES = ("nano", "transformer", "cnn", "rnn", "brnn", "resnet", "crnn", "ctransformer")
DECODER_TYPES = ("transformer", "rnn", "cnn")

# named model families of BASELINE.json / SURVEY.md Appendix A
FAMILIES = {
    "l2t": dict(encoder_type="nano", decoder_type="transformer"),
    "t2t": dict(encoder_type="transformer", decoder_type="transformer"),
    "nano2rnn": dict(encoder_type="nano", decoder_type="rnn"),
    "brnn2rnn": dict(encoder_type="brnn", decoder_type="rnn"),
    "rnn2rnn": dict(encoder_type="rnn", decoder_type="rnn"),         # unidirectional encoder (encoder/rnn_encoder.py)
    "cnn2cnn": dict(encoder_type="cnn", decoder_type="cnn"),
    # ResNet stem (encoder/resnet_encoder.py) alone, in front of the nano stack, in front of transformer layers;
    # the authors' pipeline-train.sh trains resnet -> transformer and resnet -> rnn
    "resnet2t": dict(encoder_type="resnet", decoder_type="transformer"),
    "resnet2rnn": dict(encoder_type="resnet", decoder_type="rnn"),
    "crnn2t": dict(encoder_type="crnn", decoder_type="transformer"),
    "crnn2rnn": dict(encoder_type="crnn", decoder_type="rnn"),
    "ctrans2t": dict(encoder_type="ctransformer", decoder_type="transformer"),
}


@dataclass
class ModelConfig:
    encoder_type: str = "nano"
    decoder_type: str = "transformer"
    enc_layers: int = 3
    dec_layers: int = 3
    d_model: int = 256              # enc_rnn_size == dec_rnn_size == tgt_word_vec_size
    heads: int = 8                  # models/opts.py:140
    d_ff: int = 2048                # models/opts.py:142
    vocab: List[str] = field(default_factory=lambda: SPECIALS + ["A", "C", "G", "T"])
    cnn_kernel_width: int = 3
    enc_pooling: List[int] = field(default_factory=lambda: [1])   # -audio_enc_pooling
    rnn_type: str = "LSTM"
    input_feed: int = 1
    global_attention: str = "mlp"
    position_encoding: bool = False
    self_attn_type: str = "scaled-dot"   # Transformer decoder: "average" = AverageAttention (onmt/modules/average_attn.py)
    bridge: bool = False            # -bridge: Linear + ReLU on the rnn / brnn encoder's final states (rnn_encoder.py:82-118)

    def __post_init__(self):
        if self.encoder_type not in ENCODER_TYPES:
            raise ValueError("unsupported encoder_type %r (supported: %s)" %
                             (self.encoder_type, ", ".join(ENCODER_TYPES)))
        if self.decoder_type not in DECODER_TYPES:
            raise ValueError("unsupported decoder_type %r" % (self.decoder_type,))
        if len(self.enc_pooling) == 1:
            self.enc_pooling = list(self.enc_pooling) * self.enc_layers
        if len(self.enc_pooling) != self.enc_layers:
            raise ValueError("audio_enc_pooling must have 1 or enc_layers entries")
        if self.rnn_type not in ("LSTM", "GRU"):
            raise ValueError("rnn_type must be LSTM or GRU (SRU is outside the supported path; got %r)" % self.rnn_type)
        if self.d_model % self.heads:
            raise ValueError("d_model must be divisible by heads")
        if self.self_attn_type not in ("scaled-dot", "average"):
            raise ValueError("self_attn_type must be scaled-dot or average")
        if self.encoder_type in ("resnet", "ctransformer") and self.decoder_type == "cnn":
            # model_builder.py:133-140 builds ResNetEncoder for the cnn decoder, whose outputs are laid out [d,B,T]
            # with a [1,1,B,T] "embedding" (resnet_encoder.py:195-199): not a configuration anyone can decode with
            raise ValueError("%s encoder with the cnn decoder is outside the supported translate path" % self.encoder_type)

    @property
    def vocab_size(self) -> int:
        return len(self.vocab)

    @property
    def brnn(self) -> bool:
        # models/model_builder.py:190-191: nano/crnn encoders force a bidirectional bridge
        return self.encoder_type in ("brnn", "nano", "crnn")

    @classmethod
    def family(cls, name: str, **kw) -> "ModelConfig":
        base = dict(FAMILIES[name])
        base.update(kw)
        return cls(**base)

    # ---- conversion to / from the reference's checkpoint['opt'] Namespace ------------------
    def to_opt(self) -> argparse.Namespace:
        return argparse.Namespace(
            encoder_type=self.encoder_type, decoder_type=self.decoder_type,
            enc_layers=self.enc_layers, dec_layers=self.dec_layers, layers=-1,
            rnn_size=-1, enc_rnn_size=self.d_model, dec_rnn_size=self.d_model,
            src_word_vec_size=self.d_model, tgt_word_vec_size=self.d_model, word_vec_size=-1,
            heads=self.heads, transformer_ff=self.d_ff, cnn_kernel_width=self.cnn_kernel_width,
            audio_enc_pooling=",".join(str(p) for p in self.enc_pooling),
            rnn_type=self.rnn_type, input_feed=self.input_feed, bridge=self.bridge,
            brnn=self.encoder_type == "brnn",
            global_attention=self.global_attention, global_attention_function="softmax",
            self_attn_type=self.self_attn_type, position_encoding=self.position_encoding,
            copy_attn=False, coverage_attn=False, context_gate=None, reuse_copy_attn=False,
            generator_function="softmax", dropout=0.0, feat_merge="concat",
            feat_vec_exponent=0.7, feat_vec_size=-1, optim="sgd", model_type="nano",
            sample_rate=4000, window_size=0.075, param_init=0.0, param_init_glorot=False,
        )

    @classmethod
    def from_opt(cls, opt, vocab_itos) -> "ModelConfig":
        g = lambda k, dflt=None: getattr(opt, k, dflt)
        d_enc, d_dec = g("enc_rnn_size", 500), g("dec_rnn_size", 500)
        if g("rnn_size", -1) not in (-1, None):           # models/model_builder.py:248-250
            d_enc = d_dec = opt.rnn_size
        if d_enc != d_dec:
            raise ValueError("enc_rnn_size != dec_rnn_size is not supported")
        for flag in ("copy_attn", "coverage_attn"):
            if g(flag, False):
                raise ValueError("-%s is outside the supported translate path" % flag)
        if g("context_gate") is not None:
            raise ValueError("-context_gate is outside the supported translate path")
        if g("self_attn_type", "scaled-dot") not in ("scaled-dot", "average"):
            raise ValueError("unknown -self_attn_type %r" % (g("self_attn_type"),))
        # flags that change the arithmetic and would otherwise load and decode to wrong bases without an error
        for flag in ("global_attention_function", "generator_function"):
            if g(flag, "softmax") not in ("softmax", None):
                raise ValueError("-%s %s is outside the supported translate path (softmax only)" % (flag, g(flag)))
        if list(vocab_itos[:4]) != SPECIALS:
            raise ValueError("target vocabulary must start with %s (the engine's <s> / </s> / <blank> ids), got %s"
                             % (SPECIALS, list(vocab_itos[:4])))
        enc = g("encoder_type", "rnn")
        if enc == "rnn" and g("brnn", False):
            enc = "brnn"
        return cls(
            encoder_type=enc, decoder_type=g("decoder_type", "rnn"),
            enc_layers=g("enc_layers", 2), dec_layers=g("dec_layers", 2), d_model=d_dec,
            heads=g("heads", 8), d_ff=g("transformer_ff", 2048), vocab=list(vocab_itos),
            cnn_kernel_width=g("cnn_kernel_width", 3),
            enc_pooling=[int(p) for p in str(g("audio_enc_pooling", "1")).split(",")],
            rnn_type=g("rnn_type", "LSTM"), input_feed=g("input_feed", 1),
            global_attention=g("global_attention", "general"),
            position_encoding=bool(g("position_encoding", False)),
            # models/model_builder.py:78-83: only RNNEncoder takes the flag; the other encoders never see it
            bridge=bool(g("bridge", False)) and enc in ("rnn", "brnn"),
            # decoder/transformer.py:33-38: only the Transformer decoder's self attention has the two types
            self_attn_type=g("self_attn_type", "scaled-dot") if g("decoder_type", "rnn") == "transformer" else "scaled-dot",
        )

    def asdict(self):
        return asdict(self)
