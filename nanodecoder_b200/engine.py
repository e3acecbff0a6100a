"""Python handle on the CUDA engine: PyTorch owns the buffers and the stream, libnanodec does the work.

Mirrors the model protocol the reference's Translator drives
(``model.encoder`` / ``model.decoder`` / ``model.generator``, translate/translator.py:419-421,
550-551, 584-591) at batch granularity: ``encode`` then ``decode_greedy`` / ``decode_beam``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .config import ModelConfig


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine(object):
    def __init__(self, cfg: ModelConfig, state_dict: Dict[str, torch.Tensor], max_batch: int,
                 max_src_len: int = 512, max_tgt_len: int = 100, max_beam: int = 1,
                 gemm_mode: str = "3xtf32", device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("nanodecoder_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.max_batch, self.max_src_len, self.max_tgt_len, self.max_beam = max_batch, max_src_len, max_tgt_len, max_beam
        self.gemm_mode = gemm_mode
        c = _lib.NdConfig()
        c.api_version = _lib.ND_API_VERSION
        c.device = device
        c.encoder_type = _lib.ENC[cfg.encoder_type]
        c.decoder_type = _lib.DEC[cfg.decoder_type]
        c.enc_layers, c.dec_layers = cfg.enc_layers, cfg.dec_layers
        c.d_model, c.heads, c.d_ff, c.vocab_size = cfg.d_model, cfg.heads, cfg.d_ff, cfg.vocab_size
        c.cnn_kernel_width = cfg.cnn_kernel_width
        for i in range(8):
            c.enc_pooling[i] = cfg.enc_pooling[i] if i < len(cfg.enc_pooling) else 1
        c.input_feed = int(cfg.input_feed)
        c.attn_type = _lib.ATTN[cfg.global_attention]
        c.position_encoding = int(cfg.position_encoding)
        c.rnn_type = _lib.RNN[cfg.rnn_type]
        c.self_attn_average = int(getattr(cfg, "self_attn_type", "scaled-dot") == "average" and cfg.decoder_type == "transformer")
        c.bridge = int(bool(getattr(cfg, "bridge", False)) and cfg.encoder_type in ("rnn", "brnn"))
        c.max_batch, c.max_src_len, c.max_tgt_len, c.max_beam = max_batch, max_src_len, max_tgt_len, max_beam
        c.gemm_mode = _lib.GEMM[gemm_mode]
        self._h = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            rc = self.lib.nd_create(C.byref(c), C.byref(self._h))
        if rc != 0:
            raise _lib.NanodecError(rc, self.lib.nd_last_error(None).decode())
        for name, t in state_dict.items():
            if not torch.is_floating_point(t):
                continue
            t = t.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
            self._check(self.lib.nd_load_weight(self._h, name.encode(), _ptr(t), shape, t.dim(), _lib.DTYPE_F32))
        self._check(self.lib.nd_finalize_weights(self._h))
        self._B = 0
        self._Tp = 0

    # ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise _lib.NanodecError(rc, self.lib.nd_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.nd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.nd_launch_count(self._h))

    def reset_launch_count(self):
        self.lib.nd_reset_launch_count(self._h)

    def set_option(self, name: str, value: int):
        """e.g. set_option("decode_streams", 1)"""
        self._check(self.lib.nd_set_int(self._h, name.encode(), int(value)))

    def profile_enable(self, categories=()):
        """Bracket launches of the named kernel categories with CUDA events (see _lib.PROF_CATS)."""
        mask = 0
        for c in categories:
            mask |= 1 << _lib.PROF_CATS.index(c)
        self._check(self.lib.nd_profile_enable(self._h, mask))

    def profile_read(self):
        """-> {category: (total_ms, launches)} since the last read; synchronises the device."""
        n = len(_lib.PROF_CATS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._check(self.lib.nd_profile_read(self._h, ms, cnt))
        return {c: (ms[i], int(cnt[i])) for i, c in enumerate(_lib.PROF_CATS) if cnt[i]}

    # ------------------------------------------------------------------------------------------
    def encode(self, src: torch.Tensor, lengths: torch.Tensor) -> None:
        """src [B,T] fp32 chunk-major zero padded, lengths [B] int64 — both on this device."""
        assert src.is_cuda and lengths.is_cuda and src.dtype == torch.float32 and lengths.dtype == torch.int64
        src = src.contiguous()
        B, T = src.shape
        self._check(self.lib.nd_encode(self._h, _ptr(src), _ptr(lengths), B, T, self._stream()))
        self._B = B

    def memory_bank(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (memory_bank [T',B,d] in the reference layout ([d,B,T] for the CNN encoder), memory lengths [B])."""
        tp = C.c_int32(0)
        self._check(self.lib.nd_get_memory_bank(self._h, C.c_void_p(0), C.c_void_p(0), C.byref(tp), self._stream()))
        shape = (self.cfg.d_model, self._B, tp.value) if self.cfg.encoder_type == "cnn" else \
            (tp.value, self._B, self.cfg.d_model)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        lens = torch.empty((self._B,), dtype=torch.int64, device=self.device)
        self._check(self.lib.nd_get_memory_bank(self._h, _ptr(out), _ptr(lens), C.byref(tp), self._stream()))
        return out, lens

    def decode_greedy(self, max_len: int = 100, min_len: int = 0, return_attn: bool = False,
                      return_logits: bool = False):
        """-> dict(ids [B,L] int64, scores [B], attn [L,B,T'] | None, logits [L,B,V] | None)"""
        B = self._B
        ids = torch.empty((B, max_len), dtype=torch.int64, device=self.device)
        scores = torch.empty((B,), dtype=torch.float32, device=self.device)
        attn = logits = None
        if return_attn:
            tp = C.c_int32(0)
            self._check(self.lib.nd_get_memory_bank(self._h, C.c_void_p(0), C.c_void_p(0), C.byref(tp), self._stream()))
            attn = torch.empty((max_len, B, tp.value), dtype=torch.float32, device=self.device)
        if return_logits:
            logits = torch.empty((max_len, B, self.cfg.vocab_size), dtype=torch.float32, device=self.device)
        self._check(self.lib.nd_decode_greedy(self._h, max_len, min_len, _ptr(ids), _ptr(scores), _ptr(attn),
                                              _ptr(logits), self._stream()))
        return {"ids": ids, "scores": scores, "attn": attn, "logits": logits}

    def _want_beam_attention(self, on: bool):
        """the option drops the captured graphs, so it is only touched when it changes"""
        if bool(on) != getattr(self, "_beam_attn", False):
            self.set_option("beam_attention", int(bool(on)))
            self._beam_attn = bool(on)

    def _beam_attention(self, n_best: int, max_len: int):
        """-> ([B, n_best, max_len, T'] attention rows of the hypotheses of the last beam decode (zero past their end),
        [B, n_best] int32 widths: how many source positions the reference returns of them, see include/nanodec.h)"""
        tp = C.c_int32(0)
        self._check(self.lib.nd_get_memory_bank(self._h, C.c_void_p(0), C.c_void_p(0), C.byref(tp), self._stream()))
        attn = torch.empty((self._B, n_best, max_len, tp.value), dtype=torch.float32, device=self.device)
        widths = torch.empty((self._B, n_best), dtype=torch.int32, device=self.device)
        self._check(self.lib.nd_beam_attention(self._h, n_best, max_len, _ptr(attn), _ptr(widths), self._stream()))
        return attn, widths

    def decode_beam(self, beam_size: int = 5, n_best: int = 1, max_len: int = 100, min_len: int = 0,
                    alpha: float = 0.0, return_attn: bool = False):
        """-> dict(ids [B,n_best,L] int64 (-1 padded), lens [B,n_best] int32, scores [B,n_best],
        attn None or [B,n_best,L,T'] (translator.py:806-812: the attention history of each returned hypothesis),
        attn_widths None or [B,n_best])"""
        B = self._B
        ids = torch.empty((B, n_best, max_len), dtype=torch.int64, device=self.device)
        lens = torch.empty((B, n_best), dtype=torch.int32, device=self.device)
        scores = torch.empty((B, n_best), dtype=torch.float32, device=self.device)
        self._want_beam_attention(return_attn)
        self._check(self.lib.nd_decode_beam(self._h, beam_size, n_best, max_len, min_len, float(alpha),
                                            _ptr(ids), _ptr(lens), _ptr(scores), self._stream()))
        attn, widths = self._beam_attention(n_best, max_len) if return_attn else (None, None)
        return {"ids": ids, "lens": lens, "scores": scores, "attn": attn, "attn_widths": widths}

    LENGTH_PENALTY = {"none": 0, "wu": 1, "avg": 2}

    COVERAGE_PENALTY = {"none": 0, "wu": 1, "summary": 2}

    def _object_beam_options(self, block_ngram_repeat, exclude_ids, coverage_penalty, beta, stepwise_penalty=False):
        """engine options of the object beam's extras; only touched when they change (they drop the captured graphs)"""
        mask = 0
        for v in exclude_ids:
            mask |= 1 << int(v)
        want = (int(block_ngram_repeat), mask, self.COVERAGE_PENALTY[coverage_penalty], float(beta), int(bool(stepwise_penalty)))
        if want != getattr(self, "_obj_opts", (0, 0, 0, 0.0, 0)):
            self.set_option("block_ngram_repeat", want[0])
            self.set_option("block_ngram_exclude", want[1])
            self.set_option("coverage_penalty", want[2])
            self._check(self.lib.nd_set_float(self._h, b"beta", want[3]))
            self.set_option("stepwise_penalty", want[4])
            self._obj_opts = want

    def decode_beam_object(self, beam_size: int = 5, n_best: int = 1, max_len: int = 100, min_len: int = 0,
                           length_penalty: str = "none", alpha: float = 0.0, return_attn: bool = False,
                           block_ngram_repeat: int = 0, exclude_ids=(), coverage_penalty: str = "none",
                           beta: float = 0.0, stepwise_penalty: bool = False):
        """Object beam search (the reference's default without --fast).  Same outputs as decode_beam.
        block_ngram_repeat / exclude_ids: onmt/translate/beam.py:101-124; coverage_penalty (none | wu | summary) with
        weight beta: penalties.py:39-57, beam.py:203-243."""
        self._want_beam_attention(return_attn)
        self._object_beam_options(block_ngram_repeat, exclude_ids, coverage_penalty, beta, stepwise_penalty)
        B = self._B
        ids = torch.empty((B, n_best, max_len), dtype=torch.int64, device=self.device)
        lens = torch.empty((B, n_best), dtype=torch.int32, device=self.device)
        scores = torch.empty((B, n_best), dtype=torch.float32, device=self.device)
        self._check(self.lib.nd_decode_beam_object(self._h, beam_size, n_best, max_len, min_len,
                                                   self.LENGTH_PENALTY[length_penalty], float(alpha), _ptr(ids),
                                                   _ptr(lens), _ptr(scores), self._stream()))
        attn, widths = self._beam_attention(n_best, max_len) if return_attn else (None, None)
        return {"ids": ids, "lens": lens, "scores": scores, "attn": attn, "attn_widths": widths}

    # ------------------------------------------------------------------------------------------
    def frontend_stats(self, signal: torch.Tensor, offsets: torch.Tensor, normalization: str = "median"):
        """signal int16 (raw DAC samples) or float64 (float-valued `.signal` files) [N_total] (device), offsets int64
        [n_reads+1] (device) -> (center, scale) fp64"""
        assert signal.dtype in (torch.int16, torch.float64) and offsets.dtype == torch.int64
        assert signal.is_cuda and offsets.is_cuda
        n = offsets.numel() - 1
        center = torch.empty((n,), dtype=torch.float64, device=self.device)
        scale = torch.empty((n,), dtype=torch.float64, device=self.device)
        fn = self.lib.nd_frontend_stats if signal.dtype == torch.int16 else self.lib.nd_frontend_stats_f64
        self._check(fn(self._h, _ptr(signal), _ptr(offsets), n, _lib.NORM[normalization], _ptr(center), _ptr(scale),
                       self._stream()))
        return center, scale

    def frontend_chunks(self, signal, offsets, center, scale, chunk_read, chunk_start, chunk_len: int):
        n = chunk_read.numel()
        out = torch.empty((n, chunk_len), dtype=torch.float32, device=self.device)
        lens = torch.empty((n,), dtype=torch.int64, device=self.device)
        fn = self.lib.nd_frontend_chunks if signal.dtype == torch.int16 else self.lib.nd_frontend_chunks_f64
        self._check(fn(self._h, _ptr(signal), _ptr(offsets), _ptr(center), _ptr(scale), _ptr(chunk_read),
                       _ptr(chunk_start), n, chunk_len, _ptr(out), _ptr(lens), self._stream()))
        return out, lens

    def test_gemm(self, mode: str, A, W, bias=None, residual=None, ln=None, relu=0):
        M, K = A.shape
        N = W.shape[0]
        Cout = torch.empty((M, N), dtype=torch.float32, device=self.device)
        g, b = (ln if ln is not None else (None, None))
        self._check(self.lib.nd_test_gemm(self._h, _lib.GEMM[mode], _ptr(A), _ptr(W), _ptr(bias), _ptr(residual),
                                          _ptr(g), _ptr(b), _ptr(Cout), M, N, K, relu, self._stream()))
        return Cout
