"""Generate golden vectors by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.npz

For every model family it
  1. builds the reference's own model through its factory (models/model_builder.py:236-389),
  2. loads the seeded synthetic state dict from ``nanodecoder_b200.synth`` with
     ``load_state_dict`` and checks the key sets agree (pins the checkpoint layout),
  3. runs the reference's ``Translator.translate_batch`` (greedy and ``--fast`` beam) on seeded
     synthetic chunks through a fake batch object,
  4. checks the oracle port (oracle/model.py, oracle/decode.py) against it, and
  5. stores inputs' seeds + reference outputs as a small .npz fixture.

The fixtures are what ``tests/test_oracle_golden.py`` (CPU) and the ``-m gpu`` parity tests
compare against on machines where the reference is not available.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim                                   # noqa: E402
from oracle import decode as odecode                         # noqa: E402
from oracle import frontend as ofrontend                     # noqa: E402
from oracle.model import OracleModel                         # noqa: E402
from nanodecoder_b200.config import ModelConfig              # noqa: E402
from nanodecoder_b200 import synth                           # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
LOGIT_STEPS = (0, 1, 2, 7, 50, 99)
OBJ_N_BEST = 2

# name -> (family, kwargs)   full-size d=256 3+3 configs of BASELINE.json plus small variants
CASES = {
    "l2t_d256": ("l2t", dict()),
    "t2t_d256": ("t2t", dict()),
    "nano2rnn_d256": ("nano2rnn", dict()),
    "brnn2rnn_d256": ("brnn2rnn", dict()),
    "cnn2cnn_d256": ("cnn2cnn", dict()),
    "l2t_d64": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2)),
    "t2t_d64": ("t2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2)),
    "t2t_d512_6x6": ("t2t", dict(d_model=512, enc_layers=6, dec_layers=6)),
    # the RNN decoder's other global-attention scorers (onmt/modules/global_attention.py:95-136; "general" is the
    # reference's command-line default)
    "nano2rnn_general_d64": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, global_attention="general")),
    "brnn2rnn_dot_d64": ("brnn2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, global_attention="dot")),
    # -position_encoding (onmt/modules/embeddings.py:36-43): the three decoders pass different positions
    "t2t_pe_d64": ("t2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2, position_encoding=True)),
    "nano2rnn_pe_d64": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, position_encoding=True)),
    "cnn2cnn_pe_d64": ("cnn2cnn", dict(d_model=64, enc_layers=2, dec_layers=2, position_encoding=True)),
    # StdRNNDecoder (-input_feed 0, onmt/decoders/decoder.py:187-262)
    "brnn2rnn_std_general_d64": ("brnn2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, input_feed=0,
                                                  global_attention="general")),
    "brnn2rnn_std_d256": ("brnn2rnn", dict(input_feed=0)),
    # unidirectional rnn encoder (encoder/rnn_encoder.py:64-84): hidden size = d (256: the 8-slice recurrence kernel)
    "rnn2rnn_d256": ("rnn2rnn", dict()),
    "rnn2rnn_d64": ("rnn2rnn", dict(d_model=64, enc_layers=2, dec_layers=2)),
    # -rnn_type GRU (onmt/utils/rnn_factory.py:8-17: nn.GRU encoders; onmt/models/stacked_rnn.py:39-65: StackedGRU decoder)
    "nano2rnn_gru_d64": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, rnn_type="GRU")),
    "brnn2rnn_gru_d256": ("brnn2rnn", dict(rnn_type="GRU")),
    "l2t_gru_d64": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2, rnn_type="GRU")),
    "rnn2rnn_gru_std_d64": ("rnn2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, rnn_type="GRU", input_feed=0)),
    # -bridge (encoder/rnn_encoder.py:82-118): Linear + ReLU on the final states, over rows of `layers` batch neighbours
    "brnn2rnn_bridge_d64": ("brnn2rnn", dict(d_model=64, enc_layers=2, dec_layers=2, bridge=True)),
    "rnn2rnn_gru_bridge_d64": ("rnn2rnn", dict(d_model=64, enc_layers=3, dec_layers=3, rnn_type="GRU", bridge=True)),
    # -self_attn_type average (onmt/modules/average_attn.py): cumulative average + FFN + gating instead of self attention
    "t2t_avg_d64": ("t2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2, self_attn_type="average")),
    "l2t_avg_d256": ("l2t", dict(self_attn_type="average")),
    # ResNet stem encoders (encoder/resnet_encoder.py, crnn_encoder.py, ctransformer.py); pipeline-train.sh trains
    # resnet -> transformer and resnet -> rnn at d = 256
    "resnet2t_d256": ("resnet2t", dict()),
    "resnet2rnn_d256": ("resnet2rnn", dict()),
    "resnet2t_d64": ("resnet2t", dict(d_model=64, d_ff=128, dec_layers=2)),
    "crnn2t_d64": ("crnn2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2)),
    "crnn2rnn_d64": ("crnn2rnn", dict(d_model=64, enc_layers=2, dec_layers=2)),
    "ctrans2t_d64": ("ctrans2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2)),
}


def ref_extra(cfg):
    """command-line flags of the reference for the config fields beyond the family defaults"""
    extra = ["-global_attention", cfg.global_attention]
    if cfg.position_encoding:
        extra.append("-position_encoding")
    if cfg.decoder_type == "rnn":
        extra += ["-input_feed", str(int(cfg.input_feed))]
    if cfg.rnn_type != "LSTM":
        extra += ["-rnn_type", cfg.rnn_type]
    if getattr(cfg, "bridge", False):
        extra.append("-bridge")
    if getattr(cfg, "self_attn_type", "scaled-dot") != "scaled-dot":
        extra += ["-self_attn_type", cfg.self_attn_type]
    return extra


def load_into_reference(model, sd):
    ref_keys = set(model.state_dict().keys())
    ours = set(sd.keys())
    missing = {k for k in ref_keys - ours if not k.endswith(".mask")}
    extra = ours - ref_keys
    assert not missing, "synthetic state dict lacks reference keys: %s" % sorted(missing)[:8]
    assert not extra, "synthetic state dict has unknown keys: %s" % sorted(extra)[:8]
    for k, v in model.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    model.load_state_dict(sd, strict=False)
    return model.eval()


def run_reference(family, cfg, sd, src, lengths, max_length, beam_size):
    model, fields, mopt = refshim.build_reference_model(
        family, d=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
        heads=cfg.heads, ff=cfg.d_ff, extra=ref_extra(cfg))
    load_into_reference(model, sd)
    assert list(fields["tgt"].vocab.itos) == cfg.vocab
    out = {}
    # ---- greedy
    tr = refshim.build_reference_translator(model, fields, mopt, beam_size=1, max_length=max_length)
    logits = []
    orig = tr._decode_and_generate

    def spy(*a, **k):
        lp, attn = orig(*a, **k)
        logits.append(lp.detach().clone())
        return lp, attn

    tr._decode_and_generate = spy
    mb_holder = {}
    orig_enc = tr._run_encoder

    def enc_spy(batch, data_type):
        r = orig_enc(batch, data_type)
        mb_holder["mb"], mb_holder["lens"] = r[2].detach().clone(), r[3].detach().clone()
        return r

    tr._run_encoder = enc_spy
    batch = refshim.FakeBatch(src.clone(), lengths.clone())
    t0 = time.time()
    res = tr.translate_batch(batch, refshim.FakeData(), False, fast=False)
    out["greedy_seconds"] = time.time() - t0
    out["greedy_ids"] = torch.stack([p[0] for p in res["predictions"]])
    out["greedy_scores"] = torch.stack([s[0] for s in res["scores"]])
    out["logits"] = torch.stack([logits[s] for s in LOGIT_STEPS if s < max_length])
    out["memory_bank"] = mb_holder["mb"]
    out["memory_lengths"] = mb_holder["lens"]
    # ---- fast beam
    if beam_size > 1:
        trb = refshim.build_reference_translator(model, fields, mopt, beam_size=beam_size,
                                                 fast=True, max_length=max_length)
        batch = refshim.FakeBatch(src.clone(), lengths.clone())
        t0 = time.time()
        resb = trb.translate_batch(batch, refshim.FakeData(), False, fast=True)
        out["beam_seconds"] = time.time() - t0
        out["beam_ids"] = [p[0] for p in resb["predictions"]]
        out["beam_scores"] = torch.tensor([float(s[0]) for s in resb["scores"]])
        # ---- object beam (_translate_batch + onmt.translate.Beam: the default without --fast), n_best 2
        tro = refshim.build_reference_translator(model, fields, mopt, beam_size=beam_size, fast=False,
                                                 max_length=max_length, n_best=OBJ_N_BEST)
        batch = refshim.FakeBatch(src.clone(), lengths.clone())
        t0 = time.time()
        reso = tro.translate_batch(batch, refshim.FakeData(), False, fast=False)
        out["obj_seconds"] = time.time() - t0
        out["obj_ids"] = [[torch.tensor([int(t) for t in h], dtype=torch.long) for h in p[:OBJ_N_BEST]]
                          for p in reso["predictions"]]
        out["obj_scores"] = torch.tensor([[float(x) for x in s[:OBJ_N_BEST]] for s in reso["scores"]])
    return out


def pad_ragged(seqs, L, fill=-1):
    out = np.full((len(seqs), L), fill, dtype=np.int64)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = s.numpy()
    return out


def make_case(name, B=6, T=512, max_length=100, beam_size=5, seed=2025, write=True):
    family, kw = CASES[name]
    cfg = ModelConfig.family(family, **kw)
    sd = synth.make_state_dict(cfg, seed=seed)
    chunks, lengths = synth.make_chunks(B, T=T, seed=1234, ragged=True, read_len=3)
    order = torch.argsort(lengths, descending=True, stable=True)      # iterator order
    chunks, lengths = chunks[order], lengths[order]
    src = chunks.t().contiguous().unsqueeze(2)                        # [T,B,1]

    ref = run_reference(family, cfg, sd, src, lengths, max_length, beam_size)

    # ---- oracle port vs reference
    om = OracleModel(sd, cfg)
    trace = []
    og = odecode.greedy(om, src, lengths, max_length=max_length, trace_logits=trace)
    mb_err = (og["memory_bank"] - ref["memory_bank"]).abs().max().item()
    mb_scale = ref["memory_bank"].abs().max().item()
    ol = torch.stack([trace[s] for s in LOGIT_STEPS if s < max_length])
    lg_err = (ol - ref["logits"]).abs().max().item()
    same_ids = bool((og["predictions"] == ref["greedy_ids"]).all())
    assert torch.equal(og["memory_lengths"], ref["memory_lengths"])
    ob = odecode.beam_fast(om, src, lengths, beam_size=beam_size, max_length=max_length)
    beam_same = all(torch.equal(a[0], b) for a, b in zip(ob["predictions"], ref["beam_ids"]))
    bs_err = (torch.tensor([s[0] for s in ob["scores"]]) - ref["beam_scores"]).abs().max().item()
    oo = odecode.beam_object(om, src, lengths, beam_size=beam_size, max_length=max_length, n_best=OBJ_N_BEST)
    obj_same = all(torch.equal(a, b) for pa, pb in zip(oo["predictions"], ref["obj_ids"]) for a, b in zip(pa, pb))
    obj_err = (torch.tensor(oo["scores"]) - ref["obj_scores"]).abs().max().item()
    assert obj_same and obj_err < 1e-3, "oracle object-beam output differs from the reference (%s, %g)" % (obj_same, obj_err)
    ids = ref["greedy_ids"]
    hist = torch.bincount(ids.flatten(), minlength=cfg.vocab_size).float()
    p = hist / hist.sum()
    entropy = float(-(p[p > 0] * p[p > 0].log2()).sum())
    n_eos = int(ids.eq(3).any(1).sum())
    print("%-16s ref greedy %.1fs beam %.1fs | oracle-vs-ref: mb %.2e (scale %.2f) logits %.2e "
          "greedy_ids_same=%s beam_same=%s beam_score %.2e | token entropy %.2f bits, "
          "%d/%d chunks emit </s>" % (name, ref["greedy_seconds"], ref.get("beam_seconds", 0), mb_err,
                                      mb_scale, lg_err, same_ids, beam_same, bs_err, entropy,
                                      n_eos, B))
    assert mb_err <= 2e-5 * max(1.0, mb_scale), "oracle encoder deviates from the reference"
    assert lg_err <= 2e-4, "oracle logits deviate from the reference"
    assert same_ids, "oracle greedy ids differ from the reference"
    assert beam_same and bs_err < 1e-3, "oracle beam output differs from the reference"

    if write:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        mb = ref["memory_bank"]
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            family=family, cfg_json=np.array(repr(cfg.asdict())), weight_seed=seed, chunk_seed=1234,
            B=B, T=T, max_length=max_length, beam_size=beam_size, read_len=3,
            src=src[:, :, 0].t().contiguous().numpy(), lengths=lengths.numpy(),
            logit_steps=np.array([s for s in LOGIT_STEPS if s < max_length]),
            logits=ref["logits"].numpy(),
            greedy_ids=ref["greedy_ids"].numpy(), greedy_scores=ref["greedy_scores"].numpy(),
            memory_lengths=ref["memory_lengths"].numpy(),
            memory_shape=np.array(mb.shape),
            # strided sample of the memory bank + moments of the whole tensor (keeps the file small)
            memory_sample=mb.flatten()[::97].numpy().copy(),
            memory_sum=np.float64(mb.double().sum().item()),
            memory_abs_sum=np.float64(mb.double().abs().sum().item()),
            beam_ids=pad_ragged(ref["beam_ids"], max_length), beam_scores=ref["beam_scores"].numpy(),
            obj_n_best=OBJ_N_BEST,
            obj_ids=np.stack([pad_ragged(p, max_length) for p in ref["obj_ids"]]),      # [B, n_best, L], -1 padded
            obj_scores=ref["obj_scores"].numpy(),
            token_entropy_bits=entropy,
        )


# ---- non-degenerate --fast beam cases --------------------------------------------------------------------------
# With seeded random weights "</s>" is the best step-0 token, so the default-flag beam goldens above hold one-token
# hypotheses.  The reference's own -min_length flag (translate/translator.py:714-715) keeps "</s>" at -1e20 until
# step >= min_length: 99 = the bench workload's setting (all beams live for the whole loop), 20 = hypotheses that
# finish at different steps (retirement, the -1e10 penalty and the n_best bookkeeping of :753-810 are exercised).
BEAM_CASES = {
    # file name -> (CASES key, min_length, n_best, alpha)
    "beam_l2t_d256_min99": ("l2t_d256", 99, 2, 0.0),
    "beam_l2t_d256_min20": ("l2t_d256", 20, 2, 0.0),
    "beam_l2t_d256_min20_alpha": ("l2t_d256", 20, 2, 0.7),
    "beam_t2t_d256_min99": ("t2t_d256", 99, 2, 0.0),
    "beam_t2t_d256_min20": ("t2t_d256", 20, 2, 0.0),
    "beam_nano2rnn_d256_min99": ("nano2rnn_d256", 99, 2, 0.0),
    "beam_nano2rnn_d256_min20": ("nano2rnn_d256", 20, 2, 0.0),
    "beam_brnn2rnn_d256_min99": ("brnn2rnn_d256", 99, 2, 0.0),
    "beam_brnn2rnn_d256_min20": ("brnn2rnn_d256", 20, 2, 0.0),
    "beam_cnn2cnn_d256_min99": ("cnn2cnn_d256", 99, 2, 0.0),
    "beam_cnn2cnn_d256_min20": ("cnn2cnn_d256", 20, 2, 0.0),
    "beam_t2t_d512_6x6_min20": ("t2t_d512_6x6", 20, 2, 0.0),
}


def make_beam_case(fname, B=6, T=512, max_length=100, beam_size=5, seed=2025, write=True):
    """--fast beam of the UNMODIFIED reference with -min_length / -n_best / -alpha; the oracle port must agree."""
    case, min_length, n_best, alpha = BEAM_CASES[fname]
    family, kw = CASES[case]
    if "d512" in case:
        B = 3
    cfg = ModelConfig.family(family, **kw)
    sd = synth.make_state_dict(cfg, seed=seed)
    chunks, lengths = synth.make_chunks(B, T=T, seed=1234, ragged=True, read_len=3)
    order = torch.argsort(lengths, descending=True, stable=True)
    chunks, lengths = chunks[order], lengths[order]
    src = chunks.t().contiguous().unsqueeze(2)
    model, fields, mopt = refshim.build_reference_model(
        family, d=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
        heads=cfg.heads, ff=cfg.d_ff, extra=ref_extra(cfg))
    load_into_reference(model, sd)
    trb = refshim.build_reference_translator(model, fields, mopt, beam_size=beam_size, fast=True,
                                             max_length=max_length, min_length=min_length, n_best=n_best, alpha=alpha)
    t0 = time.time()
    res = trb.translate_batch(refshim.FakeBatch(src.clone(), lengths.clone()), refshim.FakeData(), False, fast=True)
    secs = time.time() - t0
    ref_ids = [[torch.tensor([int(t) for t in h], dtype=torch.long) for h in p[:n_best]] for p in res["predictions"]]
    ref_scores = torch.tensor([[float(x) for x in s[:n_best]] for s in res["scores"]])
    ob = odecode.beam_fast(OracleModel(sd, cfg), src, lengths, beam_size=beam_size, max_length=max_length,
                           min_length=min_length, n_best=n_best, alpha=alpha)
    same = all(torch.equal(a, b) for pa, pb in zip(ob["predictions"], ref_ids) for a, b in zip(pa, pb))
    err = (torch.tensor([s[:n_best] for s in ob["scores"]]) - ref_scores).abs().max().item()
    lens = [[len(h) for h in p] for p in ref_ids]
    print("%-28s ref --fast beam %.1fs | hypothesis lengths %s | oracle == reference: %s, score err %.2e"
          % (fname, secs, lens, same, err))
    assert same and err < 1e-3, "oracle --fast beam differs from the reference"
    assert min(min(l) for l in lens) > 1, "degenerate hypotheses"
    if write:
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, fname + ".npz"),
            family=family, cfg_json=np.array(repr(cfg.asdict())), weight_seed=seed, chunk_seed=1234,
            B=B, T=T, max_length=max_length, beam_size=beam_size, read_len=3, min_length=min_length,
            n_best=n_best, alpha=np.float32(alpha),
            src=src[:, :, 0].t().contiguous().numpy(), lengths=lengths.numpy(),
            beam_ids=np.stack([pad_ragged(p, max_length) for p in ref_ids]),          # [B, n_best, L], -1 padded
            beam_scores=ref_scores.numpy())


BEAM_ATTN_CASES = {
    # name -> (family, kwargs): the attention matrices the reference returns per hypothesis under -attn_debug
    "beam_attn_l2t_d64": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2)),
    "beam_attn_nano2rnn_d64": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2)),
    # (not the CNN decoder: its attention comes back as [rows, prefix, T] (onmt/decoders/cnn_decoder.py:123-129), so
    # the reference's --fast path indexes the PREFIX axis with beam indices (translator.py:746: out of bounds, or silently
    # the wrong rows) and its object beam flattens prefix x T and keeps the first positions (:900-905): neither is an
    # attention history one could be asked to match.  The Translator here refuses -attn_debug for that combination.)
}


def make_beam_attn_case(fname, B=5, T=96, max_length=12, min_length=5, beam_size=3, n_best=2, seed=2025, write=True):
    """results["attention"] of _fast_translate_batch (translator.py:744-750,776,806-809) and of _translate_batch
    (:899-905, beam.py:135,170-178) with attn_debug on RAGGED chunks: the reference cuts the rows at memory_lengths[i] of
    the tiled length vector, i.e. at the length of chunk alive[i // beam_size]."""
    family, kw = BEAM_ATTN_CASES[fname]
    cfg = ModelConfig.family(family, **kw)
    sd = synth.make_state_dict(cfg, seed=seed)
    chunks, lengths = synth.make_chunks(B, T=T, seed=77, ragged=True, read_len=2)
    lengths = torch.tensor([T, T - 7, T - 20, T - 33, T - 50][:B])          # all different: the widths are visible
    chunks = chunks * (torch.arange(T)[None, :] < lengths[:, None])
    src = chunks.t().contiguous().unsqueeze(2)
    model, fields, mopt = refshim.build_reference_model(
        family, d=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers, heads=cfg.heads, ff=cfg.d_ff,
        extra=ref_extra(cfg))
    load_into_reference(model, sd)
    om = OracleModel(sd, cfg)
    store = {}
    for mode, fast in (("fast", True), ("obj", False)):
        tr = refshim.build_reference_translator(model, fields, mopt, beam_size=beam_size, fast=fast,
                                                max_length=max_length, min_length=min_length, n_best=n_best)
        fn = odecode.beam_fast if fast else odecode.beam_object
        ref_raises = False
        try:
            res = tr.translate_batch(refshim.FakeBatch(src.clone(), lengths.clone()), refshim.FakeData(), True, fast=fast)
        except (RuntimeError, IndexError) as ex:
            # --fast + -attn_debug with the Transformer decoder: attn is [rows, 1, T] and translator.py:746 indexes dim 1
            assert fast, ex
            ref_raises = True
            try:
                fn(om, src, lengths, beam_size=beam_size, max_length=max_length, min_length=min_length, n_best=n_best,
                   return_attention=True)
                raise AssertionError("the oracle should raise where the reference does")
            except (RuntimeError, IndexError):
                pass
            print("%-24s %-4s the reference raises %s: golden from the oracle with attn read as [1, rows, T]" % (
                fname, mode, type(ex).__name__))
        store[mode + "_ref_raises"] = np.int32(ref_raises)
        kw_o = dict(attn_rows_first=True) if (fast and ref_raises) else {}
        o = fn(om, src, lengths, beam_size=beam_size, max_length=max_length, min_length=min_length, n_best=n_best,
               return_attention=True, **kw_o)
        if ref_raises:
            res = {"predictions": [[h.tolist() for h in p] for p in o["predictions"]], "attention": o["attention"]}
        ids = np.full((B, n_best, max_length), -1, dtype=np.int64)
        attn = np.zeros((B, n_best, max_length, T), dtype=np.float32)
        widths = np.zeros((B, n_best), dtype=np.int32)
        err = 0.0
        for b in range(B):
            for n in range(n_best):
                hyp = torch.tensor([int(t) for t in res["predictions"][b][n]], dtype=torch.long)
                a = res["attention"][b][n]
                assert a.dim() == 2 and a.size(0) == len(hyp), (a.shape, len(hyp))
                assert torch.equal(o["predictions"][b][n], hyp), "oracle hypotheses differ from the reference"
                assert tuple(o["attention"][b][n].shape) == tuple(a.shape), (o["attention"][b][n].shape, a.shape)
                err = max(err, float((o["attention"][b][n] - a).abs().max()))
                ids[b, n, : len(hyp)] = hyp.numpy()
                attn[b, n, : a.size(0), : a.size(1)] = a.numpy()
                widths[b, n] = a.size(1)
        assert err < 2e-6, err
        print("%-24s %-4s widths %s (lengths %s), hyp lengths %s, oracle-vs-ref attention %.1e" % (
            fname, mode, widths.tolist(), lengths.tolist(), (ids >= 0).sum(2).tolist(), err))
        store[mode + "_ids"], store[mode + "_attn"], store[mode + "_widths"] = ids, attn, widths
    if write:
        np.savez_compressed(os.path.join(GOLDEN_DIR, fname + ".npz"), family=family, cfg_json=np.array(repr(cfg.asdict())),
                            weight_seed=seed, src=chunks.numpy(), lengths=lengths.numpy(), max_length=max_length,
                            min_length=min_length, beam_size=beam_size, n_best=n_best, **store)


OBJ_EXTRA_CASES = {
    # name -> (family, kwargs, ragged, reference flags, oracle kwargs): the object beam's optional scoring rules
    "objx_l2t_d64_ngram3": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2), True,
                            ["-block_ngram_repeat", "3"], dict(block_ngram_repeat=3)),
    "objx_l2t_d64_ngram3_ignoreA": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2), True,
                                    ["-block_ngram_repeat", "3", "-ignore_when_blocking", "A"],
                                    dict(block_ngram_repeat=3, exclusion_tokens=(4,))),
    # (wu takes log(min(coverage, 1)): a source position no beam ever attended gives -inf, which ragged chunks hit at
    # once through the reference's tiled-length widths, so the wu cases use full chunks)
    "objx_l2t_d64_covwu": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2), False,
                           ["-coverage_penalty", "wu", "-beta", "0.4", "-length_penalty", "wu", "-alpha", "0.5"],
                           dict(coverage_penalty="wu", beta=0.4, length_penalty="wu", alpha=0.5)),
    "objx_nano2rnn_d64_covsummary": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2), True,
                                     ["-coverage_penalty", "summary", "-beta", "0.3"],
                                     dict(coverage_penalty="summary", beta=0.3)),
    "objx_l2t_d64_stepwise_wu": ("l2t", dict(d_model=64, d_ff=128, enc_layers=2, dec_layers=2), False,
                                 ["-stepwise_penalty", "-coverage_penalty", "wu", "-beta", "0.3"],
                                 dict(coverage_penalty="wu", beta=0.3, stepwise_penalty=True)),
    "objx_nano2rnn_d64_stepwise_summary": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2), True,
                                           ["-stepwise_penalty", "-coverage_penalty", "summary", "-beta", "0.5",
                                            "-length_penalty", "avg"],
                                           dict(coverage_penalty="summary", beta=0.5, stepwise_penalty=True,
                                                length_penalty="avg")),
    "objx_nano2rnn_d64_covwu": ("nano2rnn", dict(d_model=64, enc_layers=2, dec_layers=2), False,
                                ["-coverage_penalty", "wu", "-beta", "0.2", "-block_ngram_repeat", "5"],
                                dict(coverage_penalty="wu", beta=0.2, block_ngram_repeat=5)),
}


def make_obj_extras_case(fname, B=5, T=96, max_length=16, min_length=12, beam_size=4, n_best=2, seed=2025, write=True):
    """_translate_batch + Beam.advance with n-gram blocking (beam.py:101-124) / GNMTGlobalScorer with a coverage penalty
    (beam.py:203-243, penalties.py:39-57), run by the unmodified reference."""
    family, kw, ragged, flags, okw = OBJ_EXTRA_CASES[fname]
    cfg = ModelConfig.family(family, **kw)
    sd = synth.make_state_dict(cfg, seed=seed)
    chunks, lengths = synth.make_chunks(B, T=T, seed=78, ragged=False)
    if okw.get("coverage_penalty") == "wu":
        # a key the decoder masks (signal value == 1.0, decoder/transformer.py:219-221) is never attended: coverage 0,
        # log 0 = -inf, every score -inf.  That is the reference's behaviour, but a golden of -inf pins nothing.
        chunks = torch.where(chunks == 1.0, torch.full_like(chunks, 1.0 + 1.0 / 64), chunks)
    if ragged:
        lengths = torch.tensor([T, T - 7, T - 20, T - 33, T - 50][:B])
        chunks = chunks * (torch.arange(T)[None, :] < lengths[:, None])
    src = chunks.t().contiguous().unsqueeze(2)
    model, fields, mopt = refshim.build_reference_model(
        family, d=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers, heads=cfg.heads, ff=cfg.d_ff,
        extra=ref_extra(cfg))
    load_into_reference(model, sd)
    tr = refshim.build_reference_translator(model, fields, mopt, beam_size=beam_size, fast=False, max_length=max_length,
                                            min_length=min_length, n_best=n_best, extra=flags)
    res = tr.translate_batch(refshim.FakeBatch(src.clone(), lengths.clone()), refshim.FakeData(), False, fast=False)
    plain = refshim.build_reference_translator(model, fields, mopt, beam_size=beam_size, fast=False,
                                               max_length=max_length, min_length=min_length, n_best=n_best,
                                               extra=[f for f in flags if False])
    res0 = plain.translate_batch(refshim.FakeBatch(src.clone(), lengths.clone()), refshim.FakeData(), False, fast=False)
    o = odecode.beam_object(OracleModel(sd, cfg), src, lengths, beam_size=beam_size, max_length=max_length,
                            min_length=min_length, n_best=n_best, **okw)
    ids = np.full((B, n_best, max_length), -1, dtype=np.int64)
    scores = np.zeros((B, n_best), dtype=np.float32)
    changed = 0
    for b in range(B):
        for n in range(n_best):
            hyp = [int(t) for t in res["predictions"][b][n]]
            assert o["predictions"][b][n].tolist() == hyp, ("oracle hypotheses differ from the reference", b, n,
                                                             o["predictions"][b][n].tolist(), hyp)
            assert np.isfinite(float(res["scores"][b][n])), (b, n, float(res["scores"][b][n]))
            assert abs(float(o["scores"][b][n]) - float(res["scores"][b][n])) < 1e-3 * max(1.0, abs(float(res["scores"][b][n]))), \
                (b, n, float(o["scores"][b][n]), float(res["scores"][b][n]))
            ids[b, n, : len(hyp)] = hyp
            scores[b, n] = float(res["scores"][b][n])
            changed += hyp != [int(t) for t in res0["predictions"][b][n]] or \
                abs(float(res["scores"][b][n]) - float(res0["scores"][b][n])) > 1e-4
    print("%-30s %d of %d hypotheses / scores differ from the plain object beam; scores %s" % (
        fname, changed, B * n_best, np.round(scores, 3).tolist()))
    assert changed > 0, "the option did not change anything: pick another case"
    if write:
        np.savez_compressed(os.path.join(GOLDEN_DIR, fname + ".npz"), family=family, cfg_json=np.array(repr(cfg.asdict())),
                            weight_seed=seed, src=chunks.numpy(), lengths=lengths.numpy(), max_length=max_length,
                            min_length=min_length, beam_size=beam_size, n_best=n_best, ids=ids, scores=scores,
                            okw=np.array(repr(okw)))


def make_frontend_golden(write=True):
    """Front end: run the reference's extract_fast5_raw on '.signal' text files.

    'mean' normalisation is pure numpy inside the reference -> fully pinned.  For 'median' the
    reference calls statsmodels.robust.mad, which is not installed: the harness injects the
    oracle's restatement of that one function (documented in oracle/frontend.py), so chunking,
    the subtraction/division and the text round trip are pinned; mad itself is cross-checked
    against scipy.stats.median_abs_deviation in the tests."""
    refshim.install()
    import tempfile
    import statsmodels.robust as robust
    if not hasattr(robust, "mad"):
        robust.mad = ofrontend.mad
    import numpy
    if not hasattr(numpy, "float"):
        numpy.float = float                                  # np.float was removed in numpy 1.24
    from utils.labelop import extract_fast5_raw
    reads = synth.make_raw_reads(4, seed=99, min_len=700, max_len=5000)
    reads.append(np.array([500, 510, 490, 500, 505], dtype=np.int16))          # 1 short chunk
    reads.append(synth.make_raw_reads(1, seed=5, min_len=1024, max_len=1024)[0])  # exact multiple
    store = {}
    for ri, raw in enumerate(reads):
        with tempfile.NamedTemporaryFile("w", suffix=".signal", delete=False) as f:
            f.write(" ".join(str(int(v)) for v in raw))
            path = f.name
        for norm in ("median", "mean"):
            for (L, S) in ((512, 512), (300, 60)):
                res = extract_fast5_raw(path, "r%d.txt" % ri, norm, L, S, "signal")
                ref_chunks = [np.array([float(x) for x in s.split()], dtype=np.float64)
                              for s in res[1:]]
                mine = ofrontend.chunk(ofrontend.normalise(raw.astype(np.float64), norm), L, S)
                assert len(mine) == len(ref_chunks), (len(mine), len(ref_chunks))
                for a, b in zip(mine, ref_chunks):
                    assert a.shape == b.shape and np.array_equal(a, b), "front end differs"
                key = "r%d_%s_%d_%d" % (ri, norm, L, S)
                flat = np.concatenate(ref_chunks).astype(np.float32)
                store[key + "_lens"] = np.array([len(c) for c in ref_chunks], dtype=np.int64)
                store[key + "_flat"] = flat if flat.size < 6000 else flat[::7].copy()
                store[key + "_sum"] = np.float64(np.concatenate(ref_chunks).sum())
        os.unlink(path)
        store["raw_%d" % ri] = raw
    print("frontend: %d reads x 2 normalisations x 2 strides identical to extract_fast5_raw" % len(reads))
    if write:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "frontend.npz"), n_reads=len(reads), **store)


def check_fast5_branch():
    """The reference's own '.fast5' branch (utils/labelop.py:199-214), run UNMODIFIED with `h5py.File` replaced by
    nanodecoder_b200/utils/h5lite.py (File / Group / Dataset objects over libnanodec's reader: nd_h5_list_group,
    nd_h5_read_dataset): its chunk strings must equal those of its '.signal' branch on the same samples (an int16
    array instead of a list of floats enters the normalisation).  Files are laid out by tests/h5_writer.py."""
    refshim.install()
    import tempfile
    import statsmodels.robust as robust
    if not hasattr(robust, "mad"):
        robust.mad = ofrontend.mad
    import numpy
    if not hasattr(numpy, "float"):
        numpy.float = float
    import h5py
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import h5_writer
    from nanodecoder_b200.utils import h5lite
    h5py.File = h5lite.File                                  # the h5py-shaped facade over libnanodec's reader
    from utils.labelop import extract_fast5_raw
    reads = synth.make_raw_reads(2, seed=41, min_len=900, max_len=4000)
    layouts = [dict(chunk=512, filters=(2, 1), other_reads=("Read_99", "read_1")),
               dict(chunk=1000, filters=(32020,), kw_vbz_version=1)]
    n = 0
    with tempfile.TemporaryDirectory() as tmp:
        for ri, (raw, kw) in enumerate(zip(reads, layouts)):
            p5, ps = os.path.join(tmp, "r%d.fast5" % ri), os.path.join(tmp, "r%d.signal" % ri)
            with open(p5, "wb") as f:
                f.write(h5_writer.make_fast5(raw, read_name="Read_%d" % (10 + ri), **kw))
            with open(ps, "w") as f:
                f.write(" ".join(str(int(v)) for v in raw))
            for norm in ("median", "mean"):
                for (L, S) in ((512, 512), (300, 60)):
                    a = extract_fast5_raw(p5, "r.txt", norm, L, S, "fast5")
                    b = extract_fast5_raw(ps, "r.txt", norm, L, S, "signal")
                    assert a == b and len(a) > 1, "the reference's fast5 branch differs from its signal branch"
                    n += 1
        bad = os.path.join(tmp, "bad.fast5")
        with open(bad, "wb") as f:
            f.write(b"not hdf5")
        try:
            extract_fast5_raw(bad, "r.txt", "median", 512, 512, "fast5")
            raise AssertionError("a file that is not HDF5 must raise")
        except IOError as e:
            assert "Likely a corrupted file" in str(e)
    print("fast5: the reference's '.fast5' branch over libnanodec's reader == its '.signal' branch on %d runs; "
          "IOError text kept" % n)


def assembly_cases():
    """Seeded inputs for the read-assembly golden: overlapping windows of a random genome with substitution / indel
    noise, repeats (longest-block ties), empty chunks, and base strings >= 200 (difflib's autojunk rule)."""
    rng = np.random.RandomState(7)
    cases = []
    for case in range(10):
        n = int(rng.randint(300, 2500))
        genome = "".join(rng.choice(list("ACGT"), size=n))
        if case == 3:
            genome = ("ACGTTGCA" * 40) + genome[:200]                  # periodic: many maximal blocks
        n = len(genome)
        win, stride = [(60, 20), (100, 30), (250, 80), (40, 10), (90, 45)][case % 5]
        chunks = []
        for s0 in range(0, max(1, n - win + 1), stride):
            seg = list(genome[s0: s0 + win])
            for _ in range(int(rng.randint(0, 4))):                     # noise
                k = int(rng.randint(0, len(seg)))
                r = rng.rand()
                if r < 0.5:
                    seg[k] = "ACGT"[int(rng.randint(4))]
                elif r < 0.75:
                    del seg[k]
                else:
                    seg.insert(k, "ACGT"[int(rng.randint(4))])
            chunks.append(" ".join(seg))
            if rng.rand() < 0.05:
                chunks.append("")                                       # a chunk that decoded to nothing
        cases.append([[c] for c in chunks])
    return cases


def make_assembly_golden(write=True):
    """Run the reference's own simple_assembly / index2base (utils/labelop.py:295-352) on seeded chunk lists."""
    refshim.install()
    import numpy
    if not hasattr(numpy, "float"):
        numpy.float = float
    if not hasattr(numpy.lib, "pad"):
        numpy.lib.pad = numpy.pad                            # np.lib.pad (labelop.py:341) left numpy's namespace in 2.0
    from utils.labelop import simple_assembly, index2base
    store = {}
    cases = assembly_cases()
    for ci, bpreads in enumerate(cases):
        votes = simple_assembly(bpreads)
        store["case%d_input" % ci] = np.array("\n".join(x[0] for x in bpreads))
        store["case%d_votes" % ci] = votes.astype(np.int32)
        store["case%d_fasta" % ci] = np.array(index2base(np.argmax(votes, axis=0)))
        store["case%d_concat" % ci] = np.array(simple_assembly(bpreads, flag_intersection=False))
    print("assembly: %d cases through the reference's simple_assembly" % len(cases))
    if write:
        np.savez_compressed(os.path.join(GOLDEN_DIR, "assembly.npz"), n_cases=len(cases), **store)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    ap.add_argument("--no-write", action="store_true")
    ap.add_argument("--B", type=int, default=6)
    ap.add_argument("--beam-cases", nargs="*", default=None, help="only the non-degenerate --fast beam cases named")
    ap.add_argument("--beam-attn", action="store_true", help="only the beam attention cases")
    ap.add_argument("--fast5", action="store_true", help="only the check of the reference's '.fast5' branch")
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if args.fast5:
        check_fast5_branch()
        return
    if args.beam_attn:
        for name in BEAM_ATTN_CASES:
            make_beam_attn_case(name, write=not args.no_write)
        for name in OBJ_EXTRA_CASES:
            make_obj_extras_case(name, write=not args.no_write)
        return
    if args.beam_cases is not None:
        for name in (args.beam_cases or list(BEAM_CASES)):
            make_beam_case(name, B=args.B, write=not args.no_write)
        return
    make_frontend_golden(write=not args.no_write)
    check_fast5_branch()
    make_assembly_golden(write=not args.no_write)
    for name in BEAM_CASES:
        make_beam_case(name, B=args.B, write=not args.no_write)
    for name in BEAM_ATTN_CASES:
        make_beam_attn_case(name, write=not args.no_write)
    for name in OBJ_EXTRA_CASES:
        make_obj_extras_case(name, write=not args.no_write)
    for name in args.cases:
        B = args.B if "d512" not in name else 3
        make_case(name, B=B, write=not args.no_write)


if __name__ == "__main__":
    main()
