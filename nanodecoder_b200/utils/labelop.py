"""Post-processing helpers with the reference's names (utils/labelop.py:295-352).

Same algorithm as the reference (difflib longest matching block + per-column vote); the
longest-block search runs in libnanodec (host C++), the vote is a numpy scatter.
"""
from __future__ import annotations

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


def longest_match(a: str, b: str):
    """(i, j, size) of difflib.SequenceMatcher(None, a, b).find_longest_match(0, len(a), 0, len(b)) -- the block
    max(get_matching_blocks(), key=size) selects -- computed by libnanodec (nd_longest_match, host C++)."""
    import ctypes as C
    from .. import _lib
    out = (C.c_int32 * 3)()
    ea, eb = a.encode("latin-1"), b.encode("latin-1")
    rc = _lib.load().nd_longest_match(ea, len(ea), eb, len(eb), out)
    if rc != 0:
        raise RuntimeError("nd_longest_match failed (%d)" % rc)
    return out[0], out[1], out[2]


def assembly_offsets(segments):
    """disp[i] between consecutive base strings (disp[0] = 0): one C call per read (nd_assembly_offsets)."""
    import ctypes as C
    from .. import _lib
    enc = [x.encode("latin-1") for x in segments]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(x) for x in enc])
    disp = np.zeros(max(1, len(enc)), dtype=np.int32)
    rc = _lib.load().nd_assembly_offsets(b"".join(enc) + b"\0", offs.ctypes.data_as(C.POINTER(C.c_int64)), len(enc),
                                         disp.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise RuntimeError("nd_assembly_offsets failed (%d)" % rc)
    return disp[: len(enc)]


def simple_assembly(bpreads, flag_intersection=True):
    """utils/labelop.py:320-352.  ``bpreads``: list (per chunk) of n_best lists of space-joined tokens.
    Same result as the reference (vote matrix [5, length]); the difflib longest-block search of every chunk pair
    runs in libnanodec and the per-base Python loop of add_count is a numpy scatter."""
    valid = [x[0].replace(" ", "") for x in bpreads if x[0] != ""]
    if not flag_intersection:
        return "".join(valid)
    lut = np.full(256, -1, dtype=np.int64)
    for k, v in base_dict.items():
        lut[ord(k)] = v
        lut[ord(k.lower())] = v
    concensus = np.zeros([len(base_keys), 1000])
    pos, length, census_len = 0, 0, 1000
    disps = assembly_offsets(valid) if len(valid) > 1 else np.zeros(len(valid), dtype=np.int32)
    for indx, bpread in enumerate(valid):
        disp = 0 if indx == 0 else int(disps[indx])
        if indx > 0 and disp + pos + len(bpread) > census_len:
            concensus = np.pad(concensus, ((0, 0), (0, 1000)), mode="constant", constant_values=0)
            census_len += 1000
        # add_count (labelop.py:311-318): a negative start trims the head of the segment
        start = pos + disp
        seg = bpread
        if start < 0:
            seg = seg[-start:]
            start = 0
        codes = lut[np.frombuffer(seg.encode("latin-1"), dtype=np.uint8)]
        if (codes < 0).any():
            raise KeyError(seg[int(np.argmax(codes < 0))])            # base_dict[base.upper()] in the reference
        if start + len(seg) > census_len:
            raise IndexError("index %d is out of bounds for axis 1 with size %d" % (start + len(seg) - 1, census_len))
        np.add.at(concensus, (codes, start + np.arange(len(seg))), 1)
        if indx > 0:
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
