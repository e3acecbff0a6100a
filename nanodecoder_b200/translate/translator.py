"""Translator: same surface as the reference's translate/translator.py (build_translator :65-90,
Translator.translate :181-369, translate_batch :505-540, setAttnFile :178-179), backed by the CUDA
engine instead of PyTorch modules.

Dispatch (translator.py:521-540): beam_size == 1 -> greedy; ``fast`` -> batched beam search;
otherwise the object-per-chunk beam (`_translate_batch` + onmt.translate.Beam), which the engine runs as a
mode of the same on-device beam kernels (nd_decode_beam_object).
"""
from __future__ import annotations

import math
import os
from itertools import count
from typing import List

import numpy as np
import torch

from ..checkpoint import load_checkpoint
from ..config import EOS
from ..engine import Engine
from ..inputters.nano_dataset import batch_order, parse_segments, pooled_batches
from .translation import TranslationBuilder


KV_MODES = {"f32": 0, "q23": 3, "q15": 4}        # nd_set_int("kv_mode"), include/nanodec.h


class _Field(object):
    def __init__(self, vocab):
        self.vocab = vocab
        self.init_token, self.eos_token, self.pad_token, self.unk_token = "<s>", "</s>", "<blank>", "<unk>"


class _Batch(object):
    """What the reference's torchtext Batch exposes to the translator (translator.py:413,543,549;
    translation.py:53-64)."""

    def __init__(self, src, src_lengths, indices):
        self.src = src                    # [T,B,1] fp32 (device)
        self.src_lengths = src_lengths    # [B] int64
        self.indices = indices            # [B] int64 positions in the original order
        self.batch_size = src.size(1)


class _Data(object):
    data_type = "nano"
    examples = None


class GNMTGlobalScorer(object):
    """alpha / beta / penalty names (onmt/translate/beam.py:181-242).  --fast only uses alpha; the object beam
    ranks finished hypotheses with the length penalty (none | wu | avg) and subtracts the coverage penalty (none | wu |
    summary, weight beta)."""

    def __init__(self, opt):
        self.alpha = opt.alpha
        self.beta = opt.beta
        self.length_penalty = getattr(opt, "length_penalty", "none")
        self.coverage_penalty = getattr(opt, "coverage_penalty", "none")


def join_tokens(ids: np.ndarray, cut: np.ndarray, itos) -> List[List[str]]:
    """ids [B,n_best,L] token ids, cut [B,n_best] hypothesis lengths -> [[" ".join(tokens)] * n_best] * B, the strings
    ``TranslationBuilder`` produces (translation.py:27-41 + translator.py:271-273).  When every emitted token is a
    single character (the bases) the whole batch is laid out as ONE byte buffer "A C G T ..." and sliced; any other
    token inside a hypothesis (<unk>, <s> ...) takes the per-hypothesis join."""
    B, N, L = ids.shape
    codes = np.array([ord(t) if len(t) == 1 and ord(t) < 128 else 0 for t in itos], dtype=np.uint8)
    c = codes[ids]
    live = np.arange(L)[None, None, :] < cut[:, :, None]
    if L == 0 or bool(((c == 0) & live).any()):
        itos_a = np.array(list(itos), dtype=object)
        return [[" ".join(itos_a[ids[j, n, : cut[j, n]]]) for n in range(N)] for j in range(B)]
    buf = np.full((B, N, 2 * L), 32, dtype=np.uint8)
    buf[:, :, 0::2] = np.where(live, c, 32)
    raw = buf.tobytes().decode("ascii")
    W = 2 * L
    cl = cut.reshape(-1).tolist()
    flat = [raw[i * W: i * W + max(2 * cl[i] - 1, 0)] for i in range(B * N)]
    return [flat[j * N: (j + 1) * N] for j in range(B)]


def format_attention(src_raw, pred_tokens, attn_rows) -> str:
    """The text block the reference writes per chunk under -attn_debug (translate/translator.py:284-335): a header of the
    source samples and of the predicted tokens (+ "</s>") on ONE line, then one line of attention weights per decode
    step.  src_raw: the chunk's samples (1-D), attn_rows: [steps][src_len] floats."""
    preds = list(pred_tokens) + ["</s>"]
    srcs = [str(x) for x in np.asarray(src_raw, dtype=np.float32).reshape(-1)]
    row_format = "{:>8.5f} " * len(srcs)
    out = ("{:>8.7} " + "{:>8.7} " * len(srcs)).format(">", *srcs)
    out += ("{:>8.7} " + "{:>8.7} " * len(preds)).format("|", *preds) + "\n"
    for row in attn_rows:
        out += row_format.format(*row) + "\n"
    return out


def build_translator(opt, report_score=False, logger=None, out_file=None):
    """translate/translator.py:65-90.  ``opt`` comes from nanodecoder_b200.opts.translate_opts."""
    if len(opt.models) != 1:
        raise ValueError("ensemble decoding (several -model files) is outside the supported translate path")
    if logger:
        logger.info("Loading model...")
    cfg, sd, vocab = load_checkpoint(opt.models[0])
    return Translator.from_state(cfg, sd, vocab, opt, report_score=report_score, logger=logger)


class Translator(object):
    def __init__(self, model: Engine, fields, opt, model_opt, global_scorer=None, report_score=False, logger=None):
        self.model = model
        self.fields = fields
        self.gpu = opt.gpu
        self.cuda = True
        self.n_best = opt.n_best
        self.max_length = opt.max_length
        self.beam_size = opt.beam_size
        self.min_length = opt.min_length
        self.fast = opt.fast
        self.data_type = getattr(opt, "data_type", "nano")
        self.verbose = opt.verbose
        self.global_scorer = global_scorer if global_scorer is not None else GNMTGlobalScorer(opt)
        self.report_score = report_score
        self.logger = logger
        self.out_file_attn = None
        self.model_opt = model_opt
        if opt.random_sampling_topk != 1:
            raise ValueError("random sampling (topk != 1) is outside the supported translate path")
        if opt.dump_beam or opt.replace_unk:
            raise ValueError("dump_beam / replace_unk are outside the supported translate path")
        if getattr(opt, "fft", False):
            # translate/translator.py:143,227 -> inputters/nano_dataset.py:64-77: librosa STFT magnitudes as encoder input
            raise ValueError("-fft (spectrogram input through librosa's STFT) is outside the supported translate path")
        self.block_ngram_repeat = int(opt.block_ngram_repeat)
        self.ignore_when_blocking = set(getattr(opt, "ignore_when_blocking", []) or [])
        if self.block_ngram_repeat != 0 and (self.beam_size == 1 or self.fast):
            # the reference asserts this in its greedy and --fast paths (translate/translator.py:411, 633)
            raise ValueError("block_ngram_repeat is only implemented by the object beam search (no -fast, beam_size > 1)")
        if self.global_scorer.coverage_penalty not in ("none", "wu", "summary"):
            raise ValueError("unknown coverage penalty %r" % (self.global_scorer.coverage_penalty,))
        self.stepwise_penalty = bool(getattr(opt, "stepwise_penalty", False))

    @classmethod
    def from_state(cls, cfg, state_dict, vocab, opt, report_score=False, logger=None):
        engine = Engine(cfg, state_dict, max_batch=opt.batch_size, max_src_len=opt.src_seq_length,
                        max_tgt_len=opt.max_length, max_beam=max(1, opt.beam_size),
                        gemm_mode=getattr(opt, "gemm_mode", "3xtf32"), device=max(0, opt.gpu))
        engine.set_option("kv_mode", KV_MODES[getattr(opt, "kv_mode", "q23")])
        return cls(engine, {"tgt": _Field(vocab)}, opt, cfg, report_score=report_score, logger=logger)

    def setAttnFile(self, out_file_attn):
        self.out_file_attn = out_file_attn

    # --------------------------------------------------------------------------------------
    def translate(self, src, tgt=None, src_dir=None, batch_size=None, attn_debug=False):
        """src: list of space separated float strings (the reference's format: the chunks of ONE read), or a
        ``(chunks [n,T] fp32, lengths [n] int64[, pad_to [n] int64])`` tuple of host or device tensors, possibly
        pooled over many reads; ``pad_to`` = the width the reference would pad each chunk to
        (inputters.nano_dataset.reference_pad_lengths) keeps pooled results identical to read-by-read calls.
        -> (all_scores, all_predictions) in input order; all_predictions[i] is a list of n_best
        space-joined token strings (translator.py:271-273)."""
        assert src is not None
        if batch_size is None:
            raise ValueError("batch_size must be set")
        if tgt is not None:
            raise ValueError("gold scoring (tgt) is outside the supported translate path")
        pad_to = None
        if isinstance(src, tuple):
            chunks, lengths = src[0], src[1]
            pad_to = np.asarray(src[2], dtype=np.int64) if len(src) > 2 and src[2] is not None else None
        else:
            chunks, lengths = parse_segments(src)
        n = chunks.size(0)
        dev = self.model.device
        host_lengths = lengths.cpu().numpy()
        if not chunks.is_cuda and torch.device(dev).type == "cuda":       # (a cpu `dev` only exists in the stub tests)
            chunks = chunks.pin_memory().to(dev, non_blocking=True)      # ONE host->device copy per read
        lengths_d = lengths.to(dev, non_blocking=True)
        builder = TranslationBuilder(_Data(), self.fields, self.n_best)
        all_scores: List = [None] * n
        all_predictions: List = [None] * n
        counter = count(1)
        pred_score_total, pred_words_total = 0.0, 0
        if pad_to is None:
            plan = [(idx, int(host_lengths[idx].max())) for idx in batch_order(host_lengths, batch_size)]
        else:
            plan = pooled_batches(host_lengths, pad_to, batch_size)
        fast_host = not (self.verbose or attn_debug)
        itos = np.array(self.fields["tgt"].vocab.itos, dtype=object)
        eos_id = list(itos).index(self.fields["tgt"].eos_token)
        totals = [0.0, 0]

        def finish(job):
            # same results as the TranslationBuilder path below, without one Python object per chunk / token:
            # ids come back as ONE array, hypotheses are cut at the first </s> with array ops
            idx, (ids, lens, scores) = job[0], self._collect_arrays(job[1])
            is_eos = ids == eos_id
            pos = np.arange(ids.shape[2])[None, None, :]
            is_end = is_eos | (pos >= lens[:, :, None])
            cut = np.where(is_end.any(2), is_end.argmax(2), ids.shape[2])
            strings = join_tokens(ids[:, : self.n_best], cut[:, : self.n_best], itos)
            score_rows = scores[:, : self.n_best].unbind(0)         # per chunk: n_best 0-d tensors when iterated
            for j, i in enumerate(idx.tolist()):
                all_scores[i] = score_rows[j]
                all_predictions[i] = strings[j]
            totals[0] += float(scores[:, 0].sum())
            totals[1] += int(cut[:, 0].sum())

        pending = None
        chunks_host = None
        for k, (idx, T) in enumerate(plan):
            idx_t = torch.from_numpy(idx).to(dev)
            if fast_host:
                # the device work of batch k and the copy of its results are only ENQUEUED here; the host side of
                # batch k-1 (strings) runs while the GPU works on batch k
                job = (idx, self._launch_arrays(chunks.index_select(0, idx_t)[:, :T].contiguous(),
                                                lengths_d.index_select(0, idx_t), slot=k & 1))
                if pending is not None:
                    finish(pending)
                pending = job
                continue
            b_src = chunks.index_select(0, idx_t)[:, :T].t().contiguous().unsqueeze(2)       # [T,B,1]
            batch = _Batch(b_src, lengths_d.index_select(0, idx_t), torch.arange(len(idx)))
            batch_data = self.translate_batch(batch, _Data(), attn_debug, fast=self.fast)
            for j, trans in enumerate(builder.from_batch(batch_data)):
                i = int(idx[j])
                all_scores[i] = trans.pred_scores[: self.n_best]
                pred_score_total += float(trans.pred_scores[0])
                pred_words_total += len(trans.pred_sents[0])
                all_predictions[i] = [" ".join(p) for p in trans.pred_sents[: self.n_best]]
                if self.verbose:
                    out = trans.log(next(counter))
                    (self.logger.info(out) if self.logger else os.write(1, out.encode("utf-8")))
                if attn_debug:
                    if self.out_file_attn is not None:
                        if chunks_host is None:
                            chunks_host = chunks.cpu().numpy()
                        self.out_file_attn.write(format_attention(chunks_host[i, : int(host_lengths[i])],
                                                                  trans.pred_sents[0], trans.attns[0].tolist()))
        if pending is not None:
            finish(pending)
        pred_score_total += totals[0]
        pred_words_total += totals[1]
        if self.report_score:
            msg = self._report_score("PRED", pred_score_total, pred_words_total)
            (self.logger.info(msg) if self.logger else print(msg))
        return all_scores, all_predictions

    # --------------------------------------------------------------------------------------
    def _object_beam_extras(self):
        """translator.py:836-848: n-gram blocking with its exclusion ids, GNMTGlobalScorer's coverage penalty"""
        vocab = self.fields["tgt"].vocab
        return dict(block_ngram_repeat=self.block_ngram_repeat,
                    exclude_ids=[vocab.stoi[t] for t in self.ignore_when_blocking],
                    coverage_penalty=self.global_scorer.coverage_penalty, beta=self.global_scorer.beta,
                    stepwise_penalty=self.stepwise_penalty)

    def translate_batch(self, batch, data, attn_debug, fast=False):
        """translator.py:505-540.  batch.src [T,B,1] (any device), batch.src_lengths [B].
        -> {"predictions": [[LongTensor]*n_best]*B, "scores", "attention", "batch", "gold_score"}"""
        eng = self.model
        src = batch.src.to(eng.device)[:, :, 0].t().contiguous()           # chunk-major [B,T]
        lengths = batch.src_lengths.to(eng.device, dtype=torch.int64)
        B = src.size(0)
        eng.encode(src, lengths)
        results = {"batch": batch, "gold_score": [0] * B}
        if self.beam_size == 1:
            out = eng.decode_greedy(self.max_length, self.min_length, return_attn=attn_debug)
            ids = out["ids"].cpu()                                        # the one device->host read
            scores = out["scores"].cpu()
            attn = out["attn"].cpu() if out["attn"] is not None else None
            mlen = eng.memory_bank()[1].cpu() if attn is not None else None
            results["predictions"] = [[ids[i]] for i in range(B)]
            results["scores"] = [[scores[i]] for i in range(B)]
            results["attention"] = [[attn[:, i, : int(mlen[i])]] if attn is not None else [[]] for i in range(B)]
            return results
        if attn_debug and getattr(self.model_opt, "decoder_type", None) == "cnn":
            raise ValueError("-attn_debug with beam search and the CNN decoder: the reference indexes the decoder's "
                             "[rows, prefix, T] attention with beam indices (translate/translator.py:746,900-905) and "
                             "returns no usable history; use -beam_size 1")
        if not fast:
            # object beam (translator.py:827-926): ranking by the GNMT global score (length penalty only)
            out = eng.decode_beam_object(self.beam_size, self.n_best, self.max_length, self.min_length,
                                         self.global_scorer.length_penalty, self.global_scorer.alpha,
                                         return_attn=attn_debug, **self._object_beam_extras())
        else:
            out = eng.decode_beam(self.beam_size, self.n_best, self.max_length, self.min_length,
                                  self.global_scorer.alpha, return_attn=attn_debug)
        ids, lens, scores = out["ids"].cpu(), out["lens"].cpu(), out["scores"].cpu()
        results["predictions"] = [[ids[i, n, : int(lens[i, n])] for n in range(self.n_best)] for i in range(B)]
        results["scores"] = [[scores[i, n] for n in range(self.n_best)] for i in range(B)]
        if out.get("attn") is not None:
            # translator.py:776,806-812 / :899-905: one [len, memory_length] attention matrix per returned hypothesis
            # (their width is memory_lengths[i] of the reference's TILED length vector: include/nanodec.h)
            attn, widths = out["attn"].cpu(), out["attn_widths"].cpu()
            results["attention"] = [[attn[i, n, : int(lens[i, n]), : int(widths[i, n])] for n in range(self.n_best)]
                                    for i in range(B)]
        else:
            results["attention"] = [[[] for _ in range(self.n_best)] for _ in range(B)]
        return results

    def _launch_arrays(self, src, lengths, slot=0):
        """src [B,T] fp32 chunk-major on the device: enqueue encode + decode (same dispatch as translate_batch) and the
        device->host copy of the results into pinned buffers of `slot` (two slots: the copy of batch k may still be
        in flight while batch k+1 is enqueued).  Nothing here waits for the GPU.  -> handle for _collect_arrays."""
        eng = self.model
        eng.encode(src, lengths)
        if self.beam_size == 1:
            out = eng.decode_greedy(self.max_length, self.min_length)
            dev_t = {"ids": out["ids"], "scores": out["scores"]}
        else:
            if self.fast:
                out = eng.decode_beam(self.beam_size, self.n_best, self.max_length, self.min_length,
                                      self.global_scorer.alpha)
            else:
                out = eng.decode_beam_object(self.beam_size, self.n_best, self.max_length, self.min_length,
                                             self.global_scorer.length_penalty, self.global_scorer.alpha,
                                             **self._object_beam_extras())
            dev_t = {"ids": out["ids"], "lens": out["lens"], "scores": out["scores"]}
        if not dev_t["ids"].is_cuda:
            # only reachable with the recording stub of tests/test_translator_host.py (host-logic tests without a GPU);
            # Engine itself refuses to exist without CUDA and always returns device tensors — there is no CPU decode
            return dev_t, None
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        host = {}
        for name, t in dev_t.items():
            key = (slot, name, tuple(t.shape), t.dtype)
            buf = self._pinned.get(key)
            if buf is None:
                buf = self._pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host[name] = buf.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev_t["ids"].device))
        return host, ev

    def _collect_arrays(self, handle):
        """-> (ids [B,n_best,L] int64 numpy (-1 padded), lens [B,n_best] int64 numpy, scores [B,n_best] torch cpu);
        waits for the batch's copy only.  The arrays are copies: the pinned buffers are reused two batches later."""
        host, ev = handle
        if ev is not None:
            ev.synchronize()
        ids = host["ids"].numpy()
        scores = host["scores"].clone() if ev is not None else host["scores"]
        if "lens" not in host:                                   # greedy: one hypothesis of max_length tokens
            ids = ids[:, None, :]
            return ids, np.full((ids.shape[0], 1), ids.shape[2], dtype=np.int64), scores[:, None]
        return ids, host["lens"].numpy().astype(np.int64), scores

    def _decode_arrays(self, src, lengths):
        """one batch, synchronously (see _launch_arrays / _collect_arrays)"""
        return self._collect_arrays(self._launch_arrays(src, lengths))

    def _report_score(self, name, score_total, words_total):
        if words_total == 0:
            return "%s No words predicted" % (name,)
        return "%s AVG SCORE: %.4f, %s PPL: %.4f" % (name, score_total / words_total, name,
                                                     math.exp(-score_total / words_total))


def count_bases(ids: torch.Tensor) -> int:
    """bases = tokens emitted before the first </s> of each row (what translate.py:85-95 reports as
    len(c_bpread) when stride == length)."""
    is_eos = ids.eq(EOS)
    first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full_like(ids[:, 0], ids.size(1)))
    return int(first.sum())
