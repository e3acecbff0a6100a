"""Post-processing helpers with the reference's names (utils/labelop.py:295-352).

Same algorithm as the reference (difflib longest matching block + per-column vote); both run in
libnanodec (host C++, nd_simple_assembly), as does the parsing of `.signal` text files (nd_parse_signal_text).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .. import _lib

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
    Same result as the reference (vote matrix [5, length], float64): the difflib longest-block search of every chunk
    pair AND the votes of add_count (:311-318) run in libnanodec (nd_simple_assembly, host C++), one call per read."""
    import ctypes as C
    from .. import _lib
    valid = [x[0].replace(" ", "") for x in bpreads if x[0] != ""]
    if not flag_intersection:
        return "".join(valid)
    lut = np.full(256, -1, dtype=np.int8)
    for k, v in base_dict.items():
        lut[ord(k)] = v
        lut[ord(k.lower())] = v
    enc = [x.encode("latin-1") for x in valid]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        offs[1:] = np.cumsum([len(x) for x in enc])
    cap = int(offs[-1]) + 2000
    counts = np.zeros((len(base_keys), cap), dtype=np.int32)
    length, err, args = C.c_int64(0), C.c_int32(0), (C.c_int64 * 2)()
    rc = _lib.load().nd_simple_assembly(b"".join(enc) + b"\0", offs.ctypes.data_as(C.POINTER(C.c_int64)), len(enc),
                                        lut.ctypes.data_as(C.POINTER(C.c_int8)),
                                        counts.ctypes.data_as(C.POINTER(C.c_int32)), cap, C.byref(length), C.byref(err),
                                        args)
    if rc != 0:
        raise RuntimeError("nd_simple_assembly failed (%d)" % rc)
    if err.value == 2:
        raise KeyError(chr(args[0]))                                      # base_dict[base.upper()] in the reference
    if err.value == 1:
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (args[0], args[1]))
    return counts[:, : length.value].astype(np.float64)


_MAP_ABOVE = 8 << 20           # larger files (multi-read) are memory mapped: only the pages a read needs are touched


def _fast5_bytes(path, always_map=False):
    """(object passed as the `file` argument, size, keep-alive): the file's bytes, or a read-only memory map of them"""
    size = os.path.getsize(path)
    if size <= _MAP_ABOVE and not (always_map and size > 0):
        raw = open(path, "rb").read()
        return raw, len(raw), raw
    m = np.memmap(path, dtype=np.uint8, mode="r")
    return C.cast(m.ctypes.data, C.c_char_p), m.size, m


def _fast5_error(why):
    if "signature not found" in why or "superblock" in why:
        return IOError("Error opening file. Likely a corrupted file. (%s)" % why)
    return RuntimeError("Raw data is not stored in Raw/Reads/Read_[read#] so new segments cannot be identified. (%s)" % why)


def list_fast5_reads(path):
    """-> (layout, [read names]): layout 1 = single-read file (members of /Raw/Reads in h5py's name order; the reference
    decodes the first), 2 = multi-read file (/read_<uuid>/Raw/Signal), which the reference cannot read and the CLI here
    expands into one read per member."""
    buf, n, keep = _fast5_bytes(path, always_map=True)       # listing touches the root group's pages only
    lib = _lib.load()
    need, count, layout, err = C.c_int64(0), C.c_int32(0), C.c_int32(0), C.create_string_buffer(512)
    rc = lib.nd_fast5_list_reads(buf, n, None, 0, C.byref(need), C.byref(count), C.byref(layout), err, 512)
    names = C.create_string_buffer(max(1, need.value))
    if rc == 0:
        rc = lib.nd_fast5_list_reads(buf, n, names, need.value, C.byref(need), C.byref(count), C.byref(layout), err, 512)
    if rc != 0:
        raise _fast5_error(err.value.decode("latin-1"))
    return layout.value, [x.decode("latin-1") for x in names.raw[: need.value].split(b"\0")[: count.value]]


def read_fast5_signal(path, read_name=None):
    """(read name, int16 samples) of a .fast5 file: the reference's
    `list(h5py.File(path)['/Raw/Reads/'].values())[0]['Signal'].value` (utils/labelop.py:199-214), read by libnanodec's own
    HDF5 reader (nd_fast5_read_signal, csrc/fast5.cu: h5py / libhdf5 are not needed); with `read_name`, that member of a
    single- or multi-read file.  Errors keep the reference's types and texts: a file that is not HDF5 ->
    IOError('Error opening file. Likely a corrupted file.') (:203-204), anything wrong below the root ->
    RuntimeError('Raw data is not stored in Raw/Reads/Read_[read#] ...') (:236-239), each with the reader's reason
    appended."""
    buf, n, keep = _fast5_bytes(path)
    lib = _lib.load()
    count, name, err = C.c_int64(0), C.create_string_buffer(256), C.create_string_buffer(512)

    def call(out, cap):
        if read_name is None:
            return lib.nd_fast5_read_signal(buf, n, out, cap, C.byref(count), name, 256, err, 512)
        return lib.nd_fast5_read_signal_of(buf, n, read_name.encode("latin-1"), out, cap, C.byref(count), err, 512)

    rc = call(None, 0)
    out = np.empty(max(1, count.value), dtype=np.int16)
    if rc == 0:
        rc = call(out.ctypes.data_as(C.POINTER(C.c_int16)), out.size)
    if rc != 0:
        raise _fast5_error(err.value.decode("latin-1"))
    return (read_name if read_name is not None else name.value.decode("latin-1")), out[: count.value]


def read_raw_signal(path, suffix):
    """Raw samples of one read (utils/labelop.py:199-219 without the normalisation): int16 for DAC values (fast5 `Signal`
    datasets, integer `.signal` files), float64 for a `.signal` file with non-integer values."""
    if suffix.startswith("fast5"):                               # "fast5" or, for one read of a multi-read file, "fast5:<name>"
        return read_fast5_signal(path, suffix[6:] or None)[1]
    raw = open(path, "rb").read()
    out = np.empty(len(raw) // 2 + 1, dtype=np.int16)           # every sample takes at least a digit and a separator
    count, status = C.c_int64(0), C.c_int32(0)
    rc = _lib.load().nd_parse_signal_text(raw, len(raw), out.ctypes.data_as(C.POINTER(C.c_int16)), out.size,
                                          C.byref(count), C.byref(status))
    if rc == 0 and status.value == 0:                           # plain integers: parsed in libnanodec (17x numpy)
        return out[: count.value].copy()
    # floats / exponents: the reference's own parse, [float(x) for x in text.split()] (utils/labelop.py:216-217)
    vals = np.array([float(x) for x in raw.decode("latin-1").split()], dtype=np.float64)
    ints = np.round(vals)
    if vals.size and (np.abs(vals - ints).max() > 0 or ints.min() < -32768 or ints.max() > 32767):
        return vals                                             # float-valued read: the fp64 front-end kernels
    return ints.astype(np.int16)
