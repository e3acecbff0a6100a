"""ORACLE (test infrastructure): fp32 PyTorch restatement of the reference's encoders, decoders
and generator, written functionally over a flat state dict with the reference's key names.

Citations are relative to /root/reference.  Nothing here is used by the product path.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

SCALE_WEIGHT = 0.5 ** 0.5          # onmt/utils/cnn_factory.py:10


def _lin(sd, prefix, x, bias=True):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"] if bias else None)


def _ln(sd, prefix, x):
    # nn.LayerNorm(d, eps=1e-6): encoder/transformer.py:32, decoder/transformer.py:42-43,
    # onmt/modules/position_ffn.py:22
    return F.layer_norm(x, (x.size(-1),), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-6)


# ======================================================================================= LSTM
def _make_lstm(sd, prefix, in_f, hidden, layers, bidirectional, rnn_type="LSTM"):
    """torch.nn.LSTM / torch.nn.GRU carrying the checkpoint weights: the same third-party op the reference calls
    (onmt/utils/rnn_factory.py:8-17: getattr(nn, rnn_type))."""
    m = getattr(torch.nn, rnn_type)(input_size=in_f, hidden_size=hidden, num_layers=layers,
                                    bidirectional=bidirectional)
    own = m.state_dict()
    m.load_state_dict({k: sd[prefix + "." + k] for k in own})
    return m.eval()


def lstm_direction_explicit(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse):
    """Explicit-cell LSTM over padded ``x [T,B,in]`` with per-sequence lengths: packed-sequence
    semantics (state starts at zero at each sequence's own first/last valid step, outputs beyond
    the length are zero).  Gate order i,f,g,o (torch.nn.LSTM).  Cross-check for the LSTM kernels.
    -> (out [T,B,h], h_n [B,h], c_n [B,h])"""
    T, B, _ = x.shape
    hdim = w_hh.size(1)
    h = x.new_zeros(B, hdim)
    c = x.new_zeros(B, hdim)
    out = x.new_zeros(T, B, hdim)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = F.linear(x[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        valid = (lengths > t).unsqueeze(1)
        c = torch.where(valid, c_new, c)
        h = torch.where(valid, h_new, h)
        out[t] = torch.where(valid, h_new, torch.zeros_like(h_new))
    return out, h, c


# ======================================================================================= encoders
def nano_encoder(sd, cfg, src, lengths):
    """encoder/nano_encoder.py:79-124.  src [T,B,1], lengths [B] (descending, as the reference's
    iterator sorts them) -> (enc_final (h,c) zeros, memory_bank [T',B,d], lengths')."""
    d = cfg.d_model
    hdim = d // 2
    B = src.size(1)
    lens = [int(v) for v in lengths.view(-1).tolist()]              # :90
    memory_bank = None
    for l in range(cfg.enc_layers):
        stride = cfg.enc_pooling[l]
        rnn = _make_lstm(sd, "encoder.rnn_%d" % l, 1 if l == 0 else d, hdim, 1, True, cfg.rnn_type)
        packed = pack_padded_sequence(src, lens, enforce_sorted=False)   # :97
        memory_bank = pad_packed_sequence(rnn(packed)[0])[0]         # :98-99  [t,B,2h]
        memory_bank = memory_bank.transpose(0, 2)                   # :101
        memory_bank = F.max_pool1d(memory_bank, stride)             # :102
        lens = [int(math.floor((n - stride) / stride + 1)) for n in lens]   # :103-104
        memory_bank = memory_bank.transpose(0, 2)                   # :105
        t, _, nf = memory_bank.shape
        p = "encoder.batchnorm_%d" % l
        src = F.batch_norm(memory_bank.contiguous().view(-1, nf), sd[p + ".running_mean"],
                           sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                           False, 0.1, 1e-5).view(t, -1, nf)        # :108-109 (eval mode)
    # NB: the PRE-batchnorm pooled output of the last layer goes through W      (:113-115)
    mb = F.linear(memory_bank.contiguous().view(-1, memory_bank.size(2)), sd["encoder.W.weight"])
    mb = mb.view(-1, B, d)
    state = mb.new_zeros(cfg.dec_layers * 2, B, hdim)               # :117-121
    return ((state, state) if cfg.rnn_type == "LSTM" else state), mb, lengths.new_tensor(lens)


def multi_head_attention(sd, prefix, key, value, query, heads, mask=None, cache=None, kind=None):
    """onmt/modules/multi_headed_attn.py:69-192.  key/value/query [B,len,d]; mask [B,1,klen] bool.
    -> (output [B,qlen,d], head-0 attention [B,qlen,klen])"""
    B = key.size(0)
    d = query.size(-1)
    dh = d // heads

    def shape(x):
        return x.view(B, -1, heads, dh).transpose(1, 2)             # :116-118

    if cache is not None and kind == "self":                        # :126-141
        q = _lin(sd, prefix + ".linear_query", query)
        k = shape(_lin(sd, prefix + ".linear_keys", query))
        v = shape(_lin(sd, prefix + ".linear_values", query))
        if cache.get("self_keys") is not None:
            k = torch.cat((cache["self_keys"], k), dim=2)
            v = torch.cat((cache["self_values"], v), dim=2)
        cache["self_keys"], cache["self_values"] = k, v
    elif cache is not None and kind == "context":                   # :142-153
        q = _lin(sd, prefix + ".linear_query", query)
        if cache.get("memory_keys") is None:
            cache["memory_keys"] = shape(_lin(sd, prefix + ".linear_keys", key))
            cache["memory_values"] = shape(_lin(sd, prefix + ".linear_values", value))
        k, v = cache["memory_keys"], cache["memory_values"]
    else:                                                           # :154-159
        k = shape(_lin(sd, prefix + ".linear_keys", key))
        v = shape(_lin(sd, prefix + ".linear_values", value))
        q = _lin(sd, prefix + ".linear_query", query)
    q = shape(q)
    q = q / math.sqrt(dh)                                           # :167  scale BEFORE QK^T
    scores = torch.matmul(q, k.transpose(2, 3))                     # :168
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(1), -1e18)       # :170-172
    attn = torch.softmax(scores, dim=-1)                            # :175
    ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, -1, d)   # :177
    out = _lin(sd, prefix + ".final_linear", ctx)                   # :179
    return out, attn[:, 0, :, :].contiguous()                       # :187-190


def feed_forward(sd, prefix, x):
    """onmt/modules/position_ffn.py:38-40."""
    inter = torch.relu(_lin(sd, prefix + ".w_1", _ln(sd, prefix + ".layer_norm", x)))
    return _lin(sd, prefix + ".w_2", inter) + x


def transformer_encoder(sd, cfg, src, lengths):
    """encoder/transformer.py:106-127.  -> (emb [T,B,d], memory_bank [T,B,d], lengths)."""
    emb = _lin(sd, "encoder.linear", src)                           # :113  Linear(1,d), no PE
    out = emb.transpose(0, 1).contiguous()
    mask = src[:, :, 0].transpose(0, 1).eq(0.0).unsqueeze(1)        # :117-121  value == 0.0 !
    for l in range(cfg.enc_layers):
        p = "encoder.transformer.%d" % l
        xn = _ln(sd, p + ".layer_norm", out)                        # :50
        ctx, _ = multi_head_attention(sd, p + ".self_attn", xn, xn, xn, cfg.heads, mask=mask)
        out = feed_forward(sd, p + ".feed_forward", ctx + out)      # :53-54
    out = _ln(sd, "encoder.layer_norm", out)                        # :125
    return emb, out.transpose(0, 1).contiguous(), lengths


def _wn_weight(sd, prefix):
    """Eval-mode WeightNormConv2d weights: Polyak buffers, w = g/||V|| * V
    (onmt/modules/weight_norm.py:8-19,153-165)."""
    v, g, b = sd[prefix + ".V_avg"], sd[prefix + ".g_avg"], sd[prefix + ".b_avg"]
    scalar = g / torch.norm(v.view(v.size(0), -1), 2, 1)
    return scalar.view(-1, 1, 1, 1) * v, b


def gated_conv(sd, prefix, x, width, nopad):
    """onmt/utils/cnn_factory.py:18-35.  x [B,d,len,1]."""
    w, b = _wn_weight(sd, prefix + ".conv")
    y = F.conv2d(x, w, b, stride=(1, 1), padding=(width // 2 * (1 - int(nopad)), 0))
    out, gate = y.split(y.size(1) // 2, 1)
    return out * torch.sigmoid(gate)


def cnn_encoder(sd, cfg, src, lengths):
    """encoder/cnn_encoder.py:29-44.  NB: returns tensors laid out [d,B,T] (the reference squeezes
    and transposes the conv layout, :43-44), not [T,B,d]."""
    emb = src.transpose(0, 1).contiguous()                          # [B,T,1]
    remap = _lin(sd, "encoder.linear", emb.view(emb.size(0) * emb.size(1), -1))
    remap = remap.view(emb.size(0), emb.size(1), -1).transpose(1, 2).unsqueeze(3)   # [B,d,T,1]
    x = remap
    for l in range(cfg.enc_layers):                                 # cnn_factory.py:50-54
        x = (x + gated_conv(sd, "encoder.cnn.layers.%d" % l, x, cfg.cnn_kernel_width, False)) \
            * SCALE_WEIGHT
    return (remap.squeeze(3).transpose(0, 1).contiguous(),
            x.squeeze(3).transpose(0, 1).contiguous(), lengths)


def rnn_encoder(sd, cfg, src, lengths):
    """encoder/rnn_encoder.py:64-84 (no bridge).  -> ((h_n,c_n) [Le*dirs,B,hh], memory_bank, lengths)"""
    bi = cfg.encoder_type == "brnn"
    hh = cfg.d_model // (2 if bi else 1)
    rnn = _make_lstm(sd, "encoder.rnn", 1, hh, cfg.enc_layers, bi, cfg.rnn_type)
    packed = pack_padded_sequence(src, lengths.view(-1).tolist(), enforce_sorted=False)
    mb, final = rnn(packed)
    if getattr(cfg, "bridge", False):                               # :82-83, 101-118
        tot = hh * cfg.enc_layers

        def bottle(i, states):      # NB: view(-1, hh * layers) of [layers*dirs, B, hh] mixes `layers` batch neighbours per row
            return torch.relu(_lin(sd, "encoder.bridge.%d" % i, states.contiguous().view(-1, tot))).view(states.size())

        final = tuple(bottle(i, st) for i, st in enumerate(final)) if isinstance(final, tuple) else bottle(0, final)
    return final, pad_packed_sequence(mb)[0], lengths


# --------------------------------------------------------------------------- ResNet stem and the encoders built on it
RESNET_PLANES = (64, 128, 256, 512)      # encoder/resnet_encoder.py:121-130 (the active ResNet.__init__)
RESNET_BLOCKS = (2, 2, 2, 2)             # models/model_builder.py:87,120,137,145: layers = [2, 2, 2, 2]


def _bn2d(sd, prefix, x):
    """nn.BatchNorm2d in eval mode (eps 1e-5)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.1, 1e-5)


def resnet_stem(sd, prefix, src):
    """encoder/resnet_encoder.py:82-170 (class ResNet with BasicBlock, :19-47) on the signal as the callers shape it
    (:192, crnn_encoder.py:97, ctransformer.py:75): src [T,B,1] -> image [B, 1, 1, T].  The kernels are (5,3) with
    padding (2,1) and the strides act on the height-1 axis only, so every layer is a width-3 convolution along time
    (kernel row 2 is the only one that meets data) and T is never shortened.  -> [T, B, num_classes]."""
    T, B, nf = src.shape
    x = src.transpose(0, 1).transpose(1, 2).contiguous().view(B, nf, -1, T)
    x = F.conv2d(x, sd[prefix + ".conv1.weight"], None, stride=(2, 1), padding=(2, 1))          # :123-124 (no bias)
    x = torch.relu(_bn2d(sd, prefix + ".bn1", x))
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip(RESNET_PLANES, RESNET_BLOCKS)):
        for bi in range(blocks):
            bp = "%s.layer%d.%d" % (prefix, li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            out = F.conv2d(x, sd[bp + ".conv1.weight"], sd[bp + ".conv1.bias"], stride=(stride, 1), padding=(2, 1))
            out = torch.relu(_bn2d(sd, bp + ".bn1", out))
            out = F.conv2d(out, sd[bp + ".conv2.weight"], sd[bp + ".conv2.bias"], stride=(1, 1), padding=(2, 1))
            out = _bn2d(sd, bp + ".bn2", out)
            residual = x
            if bi == 0 and (stride != 1 or inplanes != planes):                                 # :137-144
                residual = F.conv2d(x, sd[bp + ".downsample.0.weight"], None, stride=(stride, 1))
                residual = _bn2d(sd, bp + ".downsample.1", residual)
            x = torch.relu(out + residual)                                                      # :41-45
        inplanes = planes
    x = x.squeeze(2).transpose(0, 1).transpose(0, 2).contiguous()                               # :165  [T,B,512]
    return _lin(sd, prefix + ".fc", x)


def resnet_encoder(sd, cfg, src, lengths):
    """encoder/resnet_encoder.py:202-251 (ResNetForRNNEncoder: what model_builder.py:141-150 builds for the rnn and
    transformer decoders).  -> (zero state, memory_bank [T,B,d], lengths)."""
    B = src.size(1)
    mb = resnet_stem(sd, "encoder.cnn", src).view(-1, B, cfg.d_model)
    state = mb.new_zeros(cfg.dec_layers, B, cfg.d_model)
    return ((state, state) if cfg.rnn_type == "LSTM" else state), mb, lengths


def crnn_encoder(sd, cfg, src, lengths):
    """encoder/crnn_encoder.py:86-133: ResNet stem, then the NanoEncoder stack with a d-wide first layer."""
    d = cfg.d_model
    hdim = d // 2
    B = src.size(1)
    src = resnet_stem(sd, "encoder.cnn", src)                       # :97-98
    lens = [int(v) for v in lengths.view(-1).tolist()]
    memory_bank = None
    for l in range(cfg.enc_layers):
        stride = cfg.enc_pooling[l]
        rnn = _make_lstm(sd, "encoder.rnn_%d" % l, d, hdim, 1, True, cfg.rnn_type)
        packed = pack_padded_sequence(src, lens, enforce_sorted=False)
        memory_bank = pad_packed_sequence(rnn(packed)[0])[0]
        memory_bank = F.max_pool1d(memory_bank.transpose(0, 2), stride)
        lens = [int(math.floor((n - stride) / stride + 1)) for n in lens]
        memory_bank = memory_bank.transpose(0, 2)
        t, _, nf = memory_bank.shape
        p = "encoder.batchnorm_%d" % l
        src = F.batch_norm(memory_bank.contiguous().view(-1, nf), sd[p + ".running_mean"],
                           sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                           False, 0.1, 1e-5).view(t, -1, nf)
    mb = F.linear(memory_bank.contiguous().view(-1, memory_bank.size(2)), sd["encoder.W.weight"]).view(-1, B, d)
    state = mb.new_zeros(cfg.dec_layers * 2, B, hdim)
    return ((state, state) if cfg.rnn_type == "LSTM" else state), mb, lengths.new_tensor(lens)


def ctransformer_encoder(sd, cfg, src, lengths):
    """encoder/ctransformer.py:66-91: ResNet stem instead of Linear(1,d); the key mask is taken from channel 0 of the
    STEM OUTPUT (== 0.0, :80-84), i.e. practically never set."""
    emb = resnet_stem(sd, "encoder.cnn", src)                       # :75-77  [T,B,d]
    out = emb.transpose(0, 1).contiguous()
    mask = emb[:, :, 0].transpose(0, 1).eq(0).unsqueeze(1)
    for l in range(cfg.enc_layers):
        p = "encoder.transformer.%d" % l
        xn = _ln(sd, p + ".layer_norm", out)
        ctx, _ = multi_head_attention(sd, p + ".self_attn", xn, xn, xn, cfg.heads, mask=mask)
        out = feed_forward(sd, p + ".feed_forward", ctx + out)
    out = _ln(sd, "encoder.layer_norm", out)
    return emb, out.transpose(0, 1).contiguous(), lengths


ENCODERS = {"nano": nano_encoder, "transformer": transformer_encoder, "cnn": cnn_encoder,
            "rnn": rnn_encoder, "brnn": rnn_encoder, "resnet": resnet_encoder, "crnn": crnn_encoder,
            "ctransformer": ctransformer_encoder}


# ======================================================================================= decoders
class TransformerDecoder(object):
    """decoder/transformer.py:120-266 (one token per call, ``step`` given)."""

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.state = {}

    def init_state(self, src, memory_bank, enc_final):              # :173-176
        self.state = {"src": src, "cache": None}

    def map_state(self, fn):                                        # :178-189
        self.state["src"] = fn(self.state["src"], 1)
        if self.state["cache"] is not None:
            for lc in self.state["cache"]:
                for k, v in lc.items():
                    if v is not None:
                        lc[k] = fn(v, 0)

    def __call__(self, tgt, memory_bank, memory_lengths=None, step=None):
        """tgt [1,B',1] int64 -> (dec_out [1,B',d], attn [1,B',T])"""
        sd, cfg = self.sd, self.cfg
        if step == 0:                                               # :198-199,248-266
            self.state["cache"] = [dict(memory_keys=None, memory_values=None, self_keys=None,
                                        self_values=None) for _ in range(cfg.dec_layers)]
            if cfg.self_attn_type == "average":                     # :261-262: prev_g [B',1,d] zeros
                for lc in self.state["cache"]:
                    lc["prev_g"] = memory_bank.new_zeros(memory_bank.size(1), 1, memory_bank.size(-1))
        emb = F.embedding(tgt[:, :, 0], sd["decoder.embeddings.make_embedding.emb_luts.0.weight"])
        if cfg.position_encoding:                                   # onmt/modules/embeddings.py:36-43
            emb = emb * math.sqrt(cfg.d_model) + _positional_encoding(cfg.d_model, step, emb, sd.get("decoder.embeddings.make_embedding.pe.pe"))
        out = emb.transpose(0, 1).contiguous()                      # [B',1,d]
        mem = memory_bank.transpose(0, 1).contiguous()              # [B',T,d]
        # pad_idx is the TARGET <blank> id (1) compared against the raw signal value   :219-221
        src_pad_mask = self.state["src"][:, :, 0].transpose(0, 1).eq(1).unsqueeze(1)
        attn = None
        for l in range(cfg.dec_layers):
            p = "decoder.transformer_layers.%d" % l
            lc = self.state["cache"][l]
            xn = _ln(sd, p + ".layer_norm_1", out)                  # :74
            if cfg.self_attn_type == "average":                     # :82-84, onmt/modules/average_attn.py:52-106
                g = (xn + step * lc["prev_g"]) / (step + 1)         # cumulative average (:71-75)
                lc["prev_g"] = g
                a = feed_forward(sd, p + ".self_attn.average_layer", g)
                gate = _lin(sd, p + ".self_attn.gating_layer", torch.cat((xn, a), -1))
                ig, fg = torch.chunk(gate, 2, dim=2)
                q = torch.sigmoid(ig) * xn + torch.sigmoid(fg) * a
            else:
                q, _ = multi_head_attention(sd, p + ".self_attn", xn, xn, xn, cfg.heads,
                                            mask=None, cache=lc, kind="self")       # :76-80
            q = q + out                                             # :85
            qn = _ln(sd, p + ".layer_norm_2", q)                    # :87
            mid, attn = multi_head_attention(sd, p + ".context_attn", mem, mem, qn, cfg.heads,
                                             mask=src_pad_mask, cache=lc, kind="context")
            out = feed_forward(sd, p + ".feed_forward", mid + q)    # :92
        out = _ln(sd, "decoder.layer_norm", out)                    # :235
        return out.transpose(0, 1).contiguous(), attn.transpose(0, 1).contiguous()


def _positional_encoding(dim, step, like, table=None):
    """onmt/modules/embeddings.py:21-32,36-41: row `step` of the registered buffer when the checkpoint carries it."""
    if table is not None:
        return table[step].to(like)
    pos = torch.tensor([[float(step)]])
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim))
    pe = torch.zeros(1, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(like)


def global_attention_mlp(sd, prefix, h_t, memory, memory_lengths, attn_type="mlp"):
    """onmt/modules/global_attention.py:95-227, one step.  h_t [B,d], memory [B,T,d]
    -> (attn_h [B,d], align [B,T])"""
    B, T, d = memory.shape
    if attn_type == "mlp":                                          # :123-136
        wq = _lin(sd, prefix + ".linear_query", h_t).view(B, 1, d)
        uh = _lin(sd, prefix + ".linear_context", memory.contiguous().view(-1, d), bias=False)
        align = F.linear(torch.tanh(wq + uh.view(B, T, d)).view(-1, d),
                         sd[prefix + ".v.weight"]).view(B, T)
    else:                                                           # :114-122
        q = _lin(sd, prefix + ".linear_in", h_t, bias=False) if attn_type == "general" else h_t
        align = torch.bmm(q.unsqueeze(1), memory.transpose(1, 2)).squeeze(1)
    if memory_lengths is not None:                                  # :180-183
        mask = torch.arange(T).unsqueeze(0) < memory_lengths.unsqueeze(1)
        align = align.masked_fill(~mask, -float("inf"))
    align = torch.softmax(align, -1)                                # :187
    c = torch.bmm(align.unsqueeze(1), memory).squeeze(1)            # :194
    attn_h = _lin(sd, prefix + ".linear_out", torch.cat([c, h_t], 1), bias=(attn_type == "mlp"))
    if attn_type != "mlp":
        attn_h = torch.tanh(attn_h)                                 # :199-200
    return attn_h, align


class InputFeedRNNDecoder(object):
    """onmt/decoders/decoder.py:108-184,303-366 + onmt/models/stacked_rnn.py:22-36 (LSTM)."""

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.state = {}

    def init_state(self, src, memory_bank, enc_final):              # decoder.py:108-129
        def fix(hid):
            if self.cfg.brnn:
                hid = torch.cat([hid[0:hid.size(0):2], hid[1:hid.size(0):2]], 2)
            return hid
        # LSTM: (h, c); GRU: the final hidden alone, wrapped in a 1-tuple (decoder.py:118-122)
        hidden = tuple(fix(e) for e in enc_final) if isinstance(enc_final, tuple) else (fix(enc_final),)
        B = hidden[0].size(1)
        self.state = {"hidden": hidden,
                      "input_feed": hidden[0].new_zeros(1, B, self.cfg.d_model)}

    def map_state(self, fn):                                        # :131-134
        self.state["hidden"] = tuple(fn(x, 1) for x in self.state["hidden"])
        self.state["input_feed"] = fn(self.state["input_feed"], 1)

    def __call__(self, tgt, memory_bank, memory_lengths=None, step=None):
        sd, cfg = self.sd, self.cfg
        emb_t = F.embedding(tgt[0, :, 0], sd["decoder.embeddings.make_embedding.emb_luts.0.weight"])
        if cfg.position_encoding:       # decoder.py:323 calls the embeddings without a step: a one-token input gets pe[0]
            emb_t = emb_t * math.sqrt(cfg.d_model) + _positional_encoding(
                cfg.d_model, 0, emb_t, sd.get("decoder.embeddings.make_embedding.pe.pe")).view(1, -1)
        x = torch.cat([emb_t, self.state["input_feed"].squeeze(0)], 1) if cfg.input_feed else emb_t
        gru = cfg.rnn_type == "GRU"
        h0 = self.state["hidden"][0]
        c0 = None if gru else self.state["hidden"][1]
        h1, c1 = [], []
        for l in range(cfg.dec_layers):                             # stacked_rnn.py:25-31 (LSTM), :56-65 (GRU)
            # InputFeedRNNDecoder: StackedLSTM of LSTMCells; StdRNNDecoder (decoder.py:203-262, -input_feed 0): one
            # multi-layer nn.LSTM fed one token -- the same cell arithmetic under nn.LSTM's parameter names
            if cfg.input_feed:
                p = "decoder.rnn.layers.%d" % l
                wi, wh, bi, bh = p + ".weight_ih", p + ".weight_hh", p + ".bias_ih", p + ".bias_hh"
            else:
                wi, wh, bi, bh = ("decoder.rnn.%s_l%d" % (n, l) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
            if gru:                                                 # nn.GRUCell: gates r, z, n
                gi, gh = F.linear(x, sd[wi], sd[bi]), F.linear(h0[l], sd[wh], sd[bh])
                i_r, i_z, i_n = gi.chunk(3, 1)
                h_r, h_z, h_n = gh.chunk(3, 1)
                r = torch.sigmoid(i_r + h_r)
                z = torch.sigmoid(i_z + h_z)
                n = torch.tanh(i_n + r * h_n)
                h = n + z * (h0[l] - n)
                h1.append(h)
                x = h
                continue
            g = F.linear(x, sd[wi], sd[bi]) + F.linear(h0[l], sd[wh], sd[bh])
            i, f, gg, o = g.chunk(4, 1)
            c = torch.sigmoid(f) * c0[l] + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            h1.append(h)
            c1.append(c)
            x = h
        out, align = global_attention_mlp(sd, "decoder.attn", x, memory_bank.transpose(0, 1),
                                          memory_lengths, cfg.global_attention)   # :336-339
        self.state["hidden"] = (torch.stack(h1),) if gru else (torch.stack(h1), torch.stack(c1))
        self.state["input_feed"] = out.unsqueeze(0)                 # :347, :171
        return out.unsqueeze(0), align.unsqueeze(0)


class CNNDecoder(object):
    """onmt/decoders/cnn_decoder.py:59-132 + onmt/modules/conv_multi_step_attention.py:38-82.
    Like the reference it re-runs the whole prefix each step."""

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.state = {}

    def init_state(self, src, memory_bank, enc_hidden):             # :59-64  ([d,B,T] layout)
        self.state = {"src": (memory_bank + enc_hidden) * SCALE_WEIGHT, "previous_input": None}

    def map_state(self, fn):                                        # :66-69
        self.state["src"] = fn(self.state["src"], 1)
        if self.state["previous_input"] is not None:
            self.state["previous_input"] = fn(self.state["previous_input"], 1)

    def __call__(self, tgt, memory_bank, memory_lengths=None, step=None):
        sd, cfg = self.sd, self.cfg
        prev = self.state["previous_input"]
        if prev is not None:
            tgt = torch.cat([prev, tgt], 0)                         # :79-80
        emb = F.embedding(tgt[:, :, 0], sd["decoder.embeddings.make_embedding.emb_luts.0.weight"])
        if cfg.position_encoding:       # cnn_decoder.py:89: the whole prefix is embedded, position = index
            emb = emb * math.sqrt(cfg.d_model) + torch.stack([_positional_encoding(
                cfg.d_model, t, emb, sd.get("decoder.embeddings.make_embedding.pe.pe")).view(1, -1)
                for t in range(emb.size(0))])
        tgt_emb = emb.transpose(0, 1).contiguous()                  # [B',t,d]
        enc_top = memory_bank.transpose(0, 1).contiguous()          # [B',d,T]
        enc_comb = self.state["src"].transpose(0, 1).contiguous()   # [B',d,T]
        x = _lin(sd, "decoder.linear", tgt_emb.view(-1, tgt_emb.size(2)))
        x = x.view(tgt_emb.size(0), tgt_emb.size(1), -1).transpose(1, 2).unsqueeze(3)   # [B',d,t,1]
        pad = x.new_zeros(x.size(0), x.size(1), cfg.cnn_kernel_width - 1, 1)            # :105-108
        base = x
        attn = None
        for l in range(cfg.dec_layers):                             # :111-116
            out = gated_conv(sd, "decoder.conv_layers.%d" % l, torch.cat([pad, x], 2),
                             cfg.cnn_kernel_width, True)
            # conv_multi_step_attention.py:66-82
            B_, d_, t_, _ = out.shape
            pre = _lin(sd, "decoder.attn_layers.%d.linear_in" % l,
                       out.transpose(1, 2).contiguous().view(B_ * t_, d_))
            pre = pre.view(B_, t_, d_, 1).transpose(1, 2)
            target = ((base + pre) * SCALE_WEIGHT).squeeze(3).transpose(1, 2)           # [B',t,d]
            attn = torch.softmax(torch.bmm(target, enc_top), dim=2)                      # [B',t,T]
            c = torch.bmm(attn, enc_comb.transpose(1, 2)).unsqueeze(3).transpose(1, 2)   # [B',d,t,1]
            x = (x + (c + out) * SCALE_WEIGHT) * SCALE_WEIGHT
        dec_outs = x.squeeze(3).transpose(1, 2).transpose(0, 1).contiguous()            # [t,B',d]
        attn_out = attn.transpose(0, 1)                             # [t,B',T] (normalised layout)
        if prev is not None:                                        # :122-125
            dec_outs = dec_outs[prev.size(0):]
            attn_out = attn_out[prev.size(0):]
        self.state["previous_input"] = tgt                          # :131
        return dec_outs, attn_out.contiguous()


DECODERS = {"transformer": TransformerDecoder, "rnn": InputFeedRNNDecoder, "cnn": CNNDecoder}


def generator(sd, x):
    """models/model_builder.py:331-334: Linear(d,V) + LogSoftmax."""
    return torch.log_softmax(F.linear(x, sd["generator.0.weight"], sd["generator.0.bias"]), dim=-1)


class OracleModel(object):
    """The ``model.encoder / model.decoder / model.generator`` protocol the reference's
    Translator drives (translate/translator.py:419-421,550-551,584-591)."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg):
        self.sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
        self.cfg = cfg
        self.decoder = DECODERS[cfg.decoder_type](self.sd, cfg)

    def encoder(self, src, lengths):
        return ENCODERS[self.cfg.encoder_type](self.sd, self.cfg, src, lengths)

    def generator(self, x):
        return generator(self.sd, x)
