"""Host side of the signal front end: what the reference does in
utils/labelop.py:194-243 (extract_fast5_raw), inputters/nano_dataset.py:42-83,120-132 and
inputters/inputter.py:86-95,469-487, minus the text round trip.

The arithmetic (median / MAD / std, normalisation, fp64->fp32 rounding, chunk gather) runs in the
CUDA kernels of csrc/frontend.cu; this module only builds the chunk table (pure integer
bookkeeping that mirrors utils/labelop.py:225-233) and the batch order.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch


def chunk_table(read_lengths: Sequence[int], max_length: int, stride: int) -> Tuple[np.ndarray, np.ndarray]:
    """(chunk_read int32 [n], chunk_start int64 [n]) for reads of the given lengths.
    utils/labelop.py:225-233: for ind in range(ceil(N/stride)): [ind*stride, min(ind*stride+max_length, N));
    stop after the first chunk that reaches the end of the read."""
    reads, starts = [], []
    for r, n in enumerate(read_lengths):
        n = int(n)
        for ind in range(0, math.ceil(n / stride)):
            s = ind * stride
            reads.append(r)
            starts.append(s)
            if s + max_length >= n:
                break
    return np.asarray(reads, dtype=np.int32), np.asarray(starts, dtype=np.int64)


def batch_order(lengths: np.ndarray, batch_size: int) -> List[np.ndarray]:
    """Indices of each batch: consecutive groups of batch_size chunks, sorted by length descending,
    stable (OrderedIterator.create_batches + torchtext sort_within_batch, inputter.py:469-487)."""
    out = []
    n = len(lengths)
    for b0 in range(0, n, batch_size):
        idx = np.arange(b0, min(b0 + batch_size, n))
        out.append(idx[np.argsort(-lengths[idx], kind="stable")])
    return out


def reference_pad_lengths(chunk_lengths: Sequence[int], batch_size: int) -> np.ndarray:
    """Width every chunk of ONE read is zero-padded to by the reference, which translates read by read
    (translate.py:113-120) in consecutive groups of ``batch_size`` chunks padded to the group's longest
    (inputter.py:86-95).  The width matters: the Transformer decoder's cross attention also attends the padded
    positions (decoder/transformer.py:219-221 masks by signal VALUE, not by length).  Pooling chunks of many
    reads into large GPU batches keeps results identical as long as each chunk is still padded to this width."""
    lens = np.asarray(chunk_lengths, dtype=np.int64)
    out = np.empty_like(lens)
    for b0 in range(0, len(lens), batch_size):
        out[b0: b0 + batch_size] = lens[b0: b0 + batch_size].max()
    return out


def pooled_batches(lengths: np.ndarray, pad_to: np.ndarray, batch_size: int) -> List[Tuple[np.ndarray, int]]:
    """Batches for chunks pooled over many reads: chunks are grouped by their reference padding width (stable),
    each group cut into batches of ``batch_size`` sorted by length descending.  -> [(indices, width)]"""
    out = []
    for w in sorted(set(int(v) for v in pad_to), reverse=True):
        idx = np.nonzero(pad_to == w)[0]
        for b0 in range(0, len(idx), batch_size):
            part = idx[b0: b0 + batch_size]
            out.append((part[np.argsort(-lengths[part], kind="stable")], w))
    return out


def parse_segments(src: Sequence[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's wire format between extract_fast5_raw and the translator: one string of
    space separated floats per chunk (labelop.py:231; nano_dataset.py:49-58,81).
    -> (chunks [n, T] fp32 zero padded, lengths [n] int64) on the host."""
    arrs = [np.array(s.split(), dtype=np.float64).astype(np.float32) for s in src]
    T = max((a.size for a in arrs), default=0)
    out = np.zeros((len(arrs), T), dtype=np.float32)
    for i, a in enumerate(arrs):
        out[i, : a.size] = a
    return torch.from_numpy(out), torch.tensor([a.size for a in arrs], dtype=torch.int64)


class SignalFrontend(object):
    """Raw int16 reads -> normalised fp32 chunks on the device."""

    def __init__(self, engine, normalization: str = "median", max_length: int = 512, stride: int = 512):
        self.engine = engine
        self.normalization = normalization
        self.max_length = max_length
        self.stride = stride

    def __call__(self, reads: Sequence[np.ndarray]):
        """reads: list of int16 arrays -> (chunks [n,T] fp32 cuda, lengths [n] int64 cuda,
        chunk_read int32 numpy [n])"""
        lens = np.array([r.size for r in reads], dtype=np.int64)
        # raw DAC samples are int16; a float-valued `.signal` file arrives as float64 and takes the fp64 kernels (int16
        # reads pooled with it are widened: exact, and the statistics of both paths agree on integer samples)
        dt = np.float64 if any(np.asarray(r).dtype != np.int16 for r in reads) else np.int16
        flat = np.concatenate([np.ascontiguousarray(r, dtype=dt) for r in reads]) if len(reads) else np.zeros((0,), dt)
        return self.from_flat(torch.from_numpy(flat), lens)

    def from_flat(self, flat: torch.Tensor, read_lengths: Sequence[int]):
        """flat: the int16 samples of all reads back to back (host tensor, ideally pinned: ONE host->device copy, or
        already on the device); read_lengths: samples per read.  Same outputs as __call__."""
        dev = self.engine.device
        lens = np.asarray(read_lengths, dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        assert flat.dtype in (torch.int16, torch.float64) and flat.numel() == int(offsets[-1])
        sig = flat.to(dev, non_blocking=True)
        off = torch.from_numpy(offsets).to(dev, non_blocking=True)
        center, scale = self.engine.frontend_stats(sig, off, self.normalization)
        cr, cs = chunk_table(lens, self.max_length, self.stride)
        chunks, clen = self.engine.frontend_chunks(sig, off, center, scale, torch.from_numpy(cr).to(dev),
                                                   torch.from_numpy(cs).to(dev), self.max_length)
        return chunks, clen, cr
