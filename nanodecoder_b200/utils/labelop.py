"""Post-processing helpers with the reference's names (utils/labelop.py:295-352).

Host-side Python, same algorithm as the reference (difflib longest matching block + per-column
vote); SURVEY.md §8f ranks a device version of this as the next row after the decode path.
"""
from __future__ import annotations

import difflib

import numpy as np

base_keys = ["A", "C", "G", "T", "M"]
base_dict = {"A": 0, "C": 1, "G": 2, "T": 3, "M": 4}


def index2base(read):
    """utils/labelop.py:295-309."""
    return "".join(base_keys[x] for x in read)


def add_count(concensus, start_indx, segment):
    """utils/labelop.py:311-318."""
    if start_indx < 0:
        segment = segment[-start_indx:]
        start_indx = 0
    for i, base in enumerate(segment):
        concensus[base_dict[base.upper()]][start_indx + i] += 1


def simple_assembly(bpreads, flag_intersection=True):
    """utils/labelop.py:320-352.  ``bpreads``: list (per chunk) of n_best lists of space-joined tokens."""
    valid = [x[0].replace(" ", "") for x in bpreads if x[0] != ""]
    if not flag_intersection:
        return "".join(valid)
    concensus = np.zeros([len(base_keys), 1000])
    pos, length, census_len = 0, 0, 1000
    for indx, bpread in enumerate(valid):
        if indx == 0:
            add_count(concensus, 0, bpread)
            continue
        d = difflib.SequenceMatcher(None, valid[indx - 1], bpread)
        match_block = max(d.get_matching_blocks(), key=lambda x: x[2])
        disp = match_block[0] - match_block[1]
        if disp + pos + len(bpread) > census_len:
            concensus = np.pad(concensus, ((0, 0), (0, 1000)), mode="constant", constant_values=0)
            census_len += 1000
        add_count(concensus, pos + disp, bpread)
        pos += disp
        length = max(length, pos + len(bpread))
    return concensus[:, :length]


def read_raw_signal(path, suffix):
    """Raw int16 samples of one read (utils/labelop.py:199-219 without the normalisation)."""
    if suffix == "fast5":
        try:
            import h5py
        except ImportError as e:       # pragma: no cover - h5py is not part of this image
            raise ImportError("reading .fast5 files needs h5py (not installed here); "
                              "export reads as .signal text files instead") from e
        with h5py.File(path, "r") as f:
            raw = list(f["/Raw/Reads/"].values())[0]["Signal"][()]
        return np.asarray(raw, dtype=np.int16)
    vals = np.array(open(path, "r").read().split(), dtype=np.float64)
    ints = np.round(vals)
    if vals.size and (np.abs(vals - ints).max() > 0 or ints.min() < -32768 or ints.max() > 32767):
        raise ValueError("%s: the GPU front end takes raw int16 DAC samples; got non-integer values" % path)
    return ints.astype(np.int16)
