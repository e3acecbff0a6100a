"""ORACLE (test infrastructure): numpy float64 restatement of the reference's signal front end.

Follows utils/labelop.py:194-243 (whole-read normalisation + fixed-stride chunking),
inputters/nano_dataset.py:49-58,81 (text -> float64 -> FloatTensor: one rounding to fp32),
inputters/inputter.py:86-95 (zero padding to the batch maximum, [T,B,1] layout) and the
ordering of ``OrderedIterator`` + torchtext ``sort_within_batch`` (inputters/inputter.py:469-487).

Third-party arithmetic restated here (absent from /root/reference and from this image):
``statsmodels.robust.mad`` (statsmodels 0.9-0.14: ``median(fabs(a - median(a)) / c)`` with
``c = norm.ppf(0.75) = 0.6744897501960817`` — NOTE the division happens before the median).
"""
from __future__ import annotations

import math
from typing import List

import numpy as np

MAD_C = 0.6744897501960817          # scipy.stats.norm.ppf(3/4.)


def mad(a: np.ndarray) -> float:
    center = np.median(a)
    return float(np.median(np.fabs(a - center) / MAD_C))


def normalise(raw, normalization: str = "median") -> np.ndarray:
    """utils/labelop.py:219-223.  ``raw``: int16 Signal array (fast5) or floats (.signal)."""
    raw = np.array(raw)
    if normalization == "mean":
        return (raw - np.median(raw)) / float(np.std(raw))     # sic: subtracts the MEDIAN
    if normalization == "median":
        return (raw - np.median(raw)) / float(mad(raw))
    return raw.astype(np.float64)


def chunk(norm: np.ndarray, max_length: int, stride: int) -> List[np.ndarray]:
    """utils/labelop.py:225-233."""
    out = []
    n = norm.size
    for ind in range(0, math.ceil(n / stride)):
        start = ind * stride
        end = min(start + max_length, n)
        out.append(norm[start:end])
        if end >= n:
            break
    return out


def to_float32(seg: np.ndarray) -> np.ndarray:
    """str(float64) -> float() -> torch.FloatTensor: exact repr round trip then ONE rounding."""
    return np.asarray(seg, dtype=np.float64).astype(np.float32)


def make_batches(chunks: List[np.ndarray], batch_size: int):
    """-> list of (src [T,B,1] fp32 zero padded, lengths [B], indices [B]) per batch:
    consecutive groups of ``batch_size`` chunks, each sorted by length descending, stable."""
    batches = []
    for b0 in range(0, len(chunks), batch_size):
        idx = list(range(b0, min(b0 + batch_size, len(chunks))))
        idx.sort(key=lambda i: len(chunks[i]))                 # create_batches (ascending, stable)
        idx.sort(key=lambda i: len(chunks[i]), reverse=True)   # sort_within_batch (stable)
        t = max(len(chunks[i]) for i in idx)
        src = np.zeros((t, len(idx), 1), dtype=np.float32)
        for j, i in enumerate(idx):
            src[: len(chunks[i]), j, 0] = to_float32(chunks[i])
        lengths = np.array([len(chunks[i]) for i in idx], dtype=np.int64)
        batches.append((src, lengths, np.array(idx, dtype=np.int64)))
    return batches


def frontend(raw, normalization="median", max_length=512, stride=512):
    """raw read -> list of fp32 chunks (what the encoder finally sees, before padding)."""
    return [to_float32(c) for c in chunk(normalise(raw, normalization), max_length, stride)]
