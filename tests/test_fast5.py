"""CPU: .fast5 ingestion (utils/labelop.py:199-214: h5py.File(path)['/Raw/Reads/'] -> first member -> ['Signal'].value)
through libnanodec's own HDF5 reader (nd_fast5_read_signal / nd_h5_read_dataset, csrc/fast5.cu; host code, no GPU).

Pinning: (1) a file written by libhdf5 itself (MATLAB 7.3 = HDF5 with a 512-byte user block; scipy ships it as a test
fixture, copied to tests/golden/libhdf5_matlab73.mat) pins superblock, symbol-table group, B-tree, SNOD, local heap,
object header, dataspace, datatype and contiguous layout; (2) single-read fast5 skeletons laid out by tests/h5_writer.py
from the format specification pin the chunk B-tree and the filter pipeline; (3) the inflate implementation is checked
against Python's zlib over every deflate block type."""
import ctypes as C
import os
import sys
import zlib

import numpy as np
import pytest

from nanodecoder_b200 import _lib
from nanodecoder_b200.utils import labelop
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import h5_writer as hw  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def read_dataset(raw, path):
    lib = _lib.load()
    info, err = (C.c_int64 * 8)(), C.create_string_buffer(512)
    rc = lib.nd_h5_read_dataset(raw, len(raw), path.encode(), None, 0, info, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    buf = (C.c_uint8 * max(1, info[5]))()
    rc = lib.nd_h5_read_dataset(raw, len(raw), path.encode(), buf, info[5], info, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return list(info), bytes(buf)[: info[5]]


def read_signal(raw, tmp_path, name="r.fast5"):
    p = tmp_path / name
    p.write_bytes(raw)
    return labelop.read_fast5_signal(str(p))


def dac(n, seed=0):
    """DAC-like samples: a slow level sequence plus noise, so deflate finds matches and literals"""
    rng = np.random.default_rng(seed)
    levels = np.repeat(rng.integers(350, 700, n // 9 + 1), 9)[:n]
    return (levels + rng.integers(-12, 13, n)).astype(np.int16)


def test_file_written_by_libhdf5():
    raw = open(os.path.join(GOLDEN, "libhdf5_matlab73.mat"), "rb").read()
    assert raw[512:520] == b"\x89HDF\r\n\x1a\n"                    # superblock behind MATLAB's 512-byte user block
    info, data = read_dataset(raw, "/testdouble")
    assert info[:5] == [1, 8, 1, 0, 2] and info[6:8] == [9, 1]     # float, 8 bytes, little endian, rank 2, 9 x 1
    got = np.frombuffer(data, "<f8")
    assert np.array_equal(got, np.arange(0, 2 * np.pi + 1e-9, np.pi / 4))   # scipy's `testdouble` vector, bit for bit
    with pytest.raises(RuntimeError, match="no object named 'nope'"):
        read_dataset(raw, "/nope")
    with pytest.raises(RuntimeError, match="Raw"):                # a valid HDF5 file without /Raw/Reads
        p = os.path.join(GOLDEN, "libhdf5_matlab73.mat")
        labelop.read_fast5_signal(p)


@pytest.mark.parametrize("flavour,chunk,filters", [
    ("old", 4096, (2, 1)),            # MinKNOW's layout: chunked, shuffle + gzip
    ("old", 1000, (1,)),              # gzip only, ragged last chunk
    ("old", 777, (3, 2, 1)),          # fletcher32 below shuffle + gzip
    ("old", 512, (1, 3)),             # fletcher32 on top of gzip
    ("old", 300, ()),                 # chunked, no filter
    ("old", "contiguous", ()),
    ("old", "compact", ()),
    ("old_sb1", 2048, (2, 1)),        # superblock version 1
    ("new", None, (2, 1)),            # libver latest: superblock 2, OHDR v2, link messages, single-chunk index
    ("new", None, ()),
    ("new", "implicit", ()),
    ("new", 640, (2, 1)),             # new-style groups around a v3 chunked layout
])
def test_signal_round_trip(tmp_path, flavour, chunk, filters):
    n = 3000 if chunk == "compact" else 20011
    sig = dac(n, seed=len(filters) + n)
    raw = hw.make_fast5(sig, read_name="Read_271", chunk=chunk, filters=filters, flavour=flavour)
    name, got = read_signal(raw, tmp_path)
    assert name == "Read_271"
    assert got.dtype == np.int16 and np.array_equal(got, sig)


def test_first_read_in_name_order_and_user_block(tmp_path):
    sig = dac(5000, 3)
    for flavour in ("old", "new"):
        # h5py's values() iterates by name: "Read_1000" < "Read_999" < "read_5" bytewise
        raw = hw.make_fast5(sig, read_name="Read_1000", chunk=1024 if flavour == "old" else None, flavour=flavour,
                            other_reads=("Read_999", "read_5", "Read_10000"), userblock=1024)
        name, got = read_signal(raw, tmp_path)
        assert name == "Read_1000" and np.array_equal(got, sig)


def test_wide_groups_and_deep_trees(tmp_path):
    sig = dac(70000, 5)
    decoys = tuple("Read_%05d" % i for i in range(20000, 20060))      # 61 members: 8 SNODs under a 3-level B-tree
    raw = hw.make_fast5(sig, read_name="Read_00042", chunk=256, filters=(2, 1), other_reads=decoys, group_levels=3,
                        fan=5, cache_root=False)                       # 274 chunks under a 4-level chunk B-tree
    name, got = read_signal(raw, tmp_path)
    assert name == "Read_00042" and np.array_equal(got, sig)


def test_filter_mask_missing_chunks_continuation_and_types(tmp_path):
    sig = dac(10000, 7)
    w = hw.H5Writer()
    raw = hw.make_fast5(sig, chunk=1000, filters=(2, 1), skip_filter_on=(2, 9), missing=(4,),
                        extra_messages=[w.msg_attribute_stub(), w.msg_attribute_stub()], split_at=3)
    want = sig.copy()
    want[4000:5000] = 0                                                # a chunk that was never written reads as fill value 0
    assert np.array_equal(read_signal(raw, tmp_path)[1], want)
    # other integer types h5py would hand back are accepted while they fit the int16 DAC range
    for dt in ("<u2", ">i2", "<i4", "<u1", ">i8"):
        vals = (np.abs(sig[:3000]) % (200 if dt == "<u1" else 30000)).astype(dt)
        got = read_signal(hw.make_fast5(vals, chunk=700, filters=(2, 1)), tmp_path)[1]
        assert np.array_equal(got, vals.astype(np.int16))
    big = np.array([1, 40000, 2], "<i4")
    with pytest.raises(RuntimeError, match="outside the int16 DAC range"):
        read_signal(hw.make_fast5(big, chunk="contiguous"), tmp_path)
    assert read_signal(hw.make_fast5(np.zeros(0, np.int16), chunk="contiguous"), tmp_path)[1].size == 0
    v1 = hw.make_fast5(sig, chunk="contiguous", layout_version=1)      # layout message version 1 (HDF5 1.4 / 1.6 files)
    assert np.array_equal(read_signal(v1, tmp_path)[1], sig)
    v2 = hw.make_fast5(sig, chunk=999, filters=(1,), layout_version=2)
    assert np.array_equal(read_signal(v2, tmp_path)[1], sig)


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_inflate_against_zlib(tmp_path, level, strategy):
    """stored, fixed-Huffman and dynamic-Huffman blocks; long matches, maximal distances, incompressible input"""
    rng = np.random.default_rng(level * 10 + strategy)
    cases = [dac(50000, 1),
             rng.integers(-32768, 32768, 40000).astype(np.int16),                  # incompressible: stored / literal blocks
             np.zeros(70000, np.int16),                                             # length-258 matches at distance 1
             np.tile(rng.integers(0, 900, 16500).astype(np.int16), 3),              # distances beyond 32 000 bytes
             np.array([5], np.int16),
             np.arange(40000, dtype=np.int16)]
    for sig in cases:
        raw = hw.make_fast5(sig, chunk=len(sig), filters=(1,), level=level, strategy=strategy)
        assert np.array_equal(read_signal(raw, tmp_path)[1], sig)


def zstd_decompress(blob, cap):
    lib = _lib.load()
    out, count, err = (C.c_uint8 * max(1, cap))(), C.c_int64(0), C.create_string_buffer(512)
    rc = lib.nd_zstd_decompress(blob, len(blob), out, cap, C.byref(count), err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return bytes(out)[: count.value]


def zstd_corpus(level):
    rng = np.random.default_rng(level)
    text = b" ".join(str(int(x)).encode() for x in dac(60000, 2))                    # the .signal text of a read: ~230 KB
    r = rng.integers(0, 256, 400000, dtype=np.uint8)
    first = bytes(r[:131072])
    return [b"", b"a", b"abc" * 5, bytes(1000), bytes(300000), b"\x07" * 131073,
            dac(3000, 1).tobytes(), dac(200000, 4).tobytes(), text, text[:700], text[:5000],
            rng.integers(0, 256, 100000, dtype=np.uint8).tobytes(),                  # incompressible: raw blocks
            rng.integers(0, 4, 70000, dtype=np.uint8).tobytes(),                     # 2-bit entropy: Huffman heavy, few matches
            rng.integers(0, 2, 50000, dtype=np.uint8).tobytes() + rng.integers(0, 200, 50000, dtype=np.uint8).tobytes(),
            np.tile(rng.integers(0, 256, 150000, dtype=np.uint8), 2).tobytes(),      # a match 150 000 bytes back
            bytes(rng.integers(0, 256, 40, dtype=np.uint8)) * 3000,                  # period-40 repeats: repeat offsets
            b"".join(bytes([i % 251]) * (i % 37 + 1) for i in range(20000)),
            bytes(rng.choice([65, 66, 67, 68], 200, p=[.7, .15, .1, .05]).astype(np.uint8)),   # one Huffman stream
            b"".join(bytes([b]) + b"ABCDEFGH" for b in r[:40000]),                   # identical sequences: RLE code tables
            first + b"".join(b"x" + first[i * 100:i * 100 + 60] for i in range(1300)),   # a block whose literals are all 'x'
            b"".join(bytes([b]) + b"ABC" for b in r[:100000]),                       # ~30 000 sequences per block
            b"".join(bytes([b]) + b"ABCD" for b in r[:100000])]


def test_zstd_corpus_covers_the_format():
    """the libzstd-written frames the decoder is checked on use every construct the decoder implements, except the
    3-byte sequence count (>= 32 512 sequences in one block, which libzstd does not emit for 128 KiB blocks)"""
    from collections import Counter
    found = Counter()
    for level in (1, 3, 9, 19):
        for raw in zstd_corpus(level):
            hw.zstd_frame_features(hw.zstd_compress(raw, level), found)
    need = ["block_raw", "block_rle", "block_compressed", "literals_raw", "literals_rle", "literals_huffman",
            "literals_treeless", "streams_1", "streams_4", "weights_direct", "weights_fse", "no_sequences", "nseq_1_byte",
            "nseq_2_byte"] + [t + m for t in ("ll_", "of_", "ml_") for m in ("predefined", "rle", "fse", "repeat")]
    assert [k for k in need if not found[k]] == [], found


@pytest.mark.parametrize("level", [1, 3, 9, 19])
def test_zstd_against_libzstd(level):
    """frames written by libzstd (pyarrow's bundled copy): raw / RLE / compressed blocks, raw / RLE / Huffman / treeless
    literals in one and four streams, direct and FSE-compressed weights, predefined / RLE / FSE / repeat sequence tables,
    repeat offsets, frames of several blocks (> 128 KiB), long offsets"""
    rng = np.random.default_rng(level)
    cases = zstd_corpus(level)
    for raw in cases:
        blob = hw.zstd_compress(raw, level)
        assert zstd_decompress(blob, len(raw)) == raw
        assert zstd_decompress(blob + blob, 2 * len(raw)) == raw + raw               # concatenated frames
        if len(raw) > 10:
            with pytest.raises(RuntimeError, match="longer than the space"):
                zstd_decompress(blob, len(raw) - 1)
            with pytest.raises(RuntimeError):
                zstd_decompress(blob[:-3], len(raw))
    # bit flips never crash and are (almost always) caught by the stream-end / size checks
    raw = dac(30000, 8).tobytes()
    blob = hw.zstd_compress(raw, level)
    caught = 0
    for _ in range(300):
        bad = bytearray(blob)
        bad[int(rng.integers(4, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        try:
            caught += zstd_decompress(bytes(bad), len(raw)) != raw
        except RuntimeError:
            caught += 1
    assert caught >= 295, caught


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("flavour,chunk,level", [("old", 4096, 1), ("old", 1000, 3), ("old", 70000, 1), ("new", None, 1),
                                                 ("old", 512, 0)])
def test_vbz_compressed_signal(tmp_path, flavour, chunk, level, version):
    """MinKNOW's default since 2019: filter 32020, client values {vbz version, 2, 1, zstd level}; version 1 ("svb16") is
    what ont_fast5_api writes today, version 0 its `vbz_legacy_v0`"""
    sig = dac(50021, 13)
    sig[1000:1010] = [-32768, 32767, -32768, 0, 32767, 32767, -1, 1, -32768, -32768]   # 3-byte zig-zag deltas
    raw = hw.make_fast5(sig, read_name="Read_9", chunk=chunk if flavour == "old" else 2048, filters=(32020,), level=level,
                        flavour=flavour, kw_vbz_version=version)
    name, got = read_signal(raw, tmp_path)
    assert name == "Read_9" and np.array_equal(got, sig)
    assert len(raw) < sig.nbytes * (0.75 if level else 1.2)                            # it does compress
    if level and flavour == "old" and chunk == 70000:
        # a constant chunk is a couple of zstd RLE blocks: far beyond deflate's 1032 x expansion, still a valid read
        flat = np.full(300000, 517, np.int16)
        raw = hw.make_fast5(flat, chunk=150000, filters=(32020,), level=level, kw_vbz_version=version)
        assert len(raw) < 20000 and np.array_equal(read_signal(raw, tmp_path)[1], flat)


def test_errors_keep_the_reference_types(tmp_path):
    sig = dac(4000, 9)
    good = hw.make_fast5(sig, chunk=1000, filters=(2, 1))
    with pytest.raises(IOError, match="Likely a corrupted file"):                   # labelop.py:203-204
        read_signal(b"not an hdf5 file" * 100, tmp_path)
    with pytest.raises(IOError, match="Likely a corrupted file"):
        read_signal(b"", tmp_path)
    # unknown VBZ versions and unknown filters are named in the message
    with pytest.raises(RuntimeError, match="VBZ version 2"):
        read_signal(hw.make_fast5(sig, chunk=1000, filters=(32020,), kw_vbz_version=2), tmp_path)
    with pytest.raises(RuntimeError, match="unsupported HDF5 filter id 307"):
        read_signal(hw.make_fast5(sig, chunk=1000, filters=(307,)), tmp_path)
    # truncation and bit flips anywhere never crash: they either read (flip in padding) or raise one of the two errors
    for cut in (len(good) // 7, len(good) // 2, len(good) - 9):
        with pytest.raises((RuntimeError, IOError)):
            read_signal(good[:cut], tmp_path)
    rng = np.random.default_rng(0)
    outcomes = {"ok": 0, "same": 0, "raised": 0}
    for _ in range(400):
        bad = bytearray(good)
        for pos in rng.integers(0, len(bad), 3):
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            got = read_signal(bytes(bad), tmp_path)[1]
            outcomes["same" if np.array_equal(got, sig) else "ok"] += 1
        except (RuntimeError, IOError):
            outcomes["raised"] += 1
    assert sum(outcomes.values()) == 400 and outcomes["raised"] > 50, outcomes
    # a corrupted deflate stream is caught by the adler32 check or the length check, not returned as samples
    start = good.index(zlib.compress(hw.H5Writer.encode_chunk(sig[:1000].tobytes(), 2, (2,)), 1)[:8])
    bad = bytearray(good)
    bad[start + 40] ^= 0x10
    with pytest.raises(RuntimeError):
        read_signal(bytes(bad), tmp_path)


def test_read_raw_signal_and_cli_loader(tmp_path):
    """the reference's worker: suffix 'fast5' -> the Signal dataset, the same samples a .signal export of the read holds"""
    sig = dac(12345, 11)
    (tmp_path / "a.fast5").write_bytes(hw.make_fast5(sig, chunk=4096))
    (tmp_path / "a.signal").write_text(" ".join(str(int(x)) for x in sig))
    a = labelop.read_raw_signal(str(tmp_path / "a.fast5"), "fast5")
    b = labelop.read_raw_signal(str(tmp_path / "a.signal"), "signal")
    assert a.dtype == b.dtype == np.int16 and np.array_equal(a, b)


def test_multi_read_files(tmp_path, monkeypatch):
    """beyond the reference (its reader needs /Raw/Reads): /read_<uuid>/Raw/Signal members, listed in name order and read
    by name; files above the mapping threshold go through a read-only memory map"""
    rng = np.random.default_rng(21)
    ids = ["%08x-0000-4000-8000-%012x" % (int(rng.integers(0, 2 ** 32)), i) for i in range(23)]
    reads = {rid: dac(int(rng.integers(1, 9000)), 100 + i) for i, rid in enumerate(ids)}
    for mapped in (False, True):
        if mapped:
            monkeypatch.setattr(labelop, "_MAP_ABOVE", 1 << 16)
        p = tmp_path / ("multi%d.fast5" % mapped)
        p.write_bytes(hw.make_multi_fast5(reads, chunk=1024, vbz_version=int(mapped), pad_to=(1 << 18) if mapped else 0))
        layout, names = labelop.list_fast5_reads(str(p))
        assert layout == 2 and names == sorted("read_" + r for r in ids)
        for rid in ids:
            name, got = labelop.read_fast5_signal(str(p), "read_" + rid)
            assert name == "read_" + rid and np.array_equal(got, reads[rid])
            assert np.array_equal(labelop.read_raw_signal(str(p), "fast5:read_" + rid), reads[rid])
        with pytest.raises(RuntimeError, match="no object named 'read_nope'"):
            labelop.read_fast5_signal(str(p), "read_nope")
        with pytest.raises(RuntimeError, match="Raw"):                      # the reference's call on a multi-read file
            labelop.read_fast5_signal(str(p))
    # a single-read file lists the members of /Raw/Reads; the first is what the reference decodes
    sig = dac(3000, 1)
    q = tmp_path / "single.fast5"
    q.write_bytes(hw.make_fast5(sig, read_name="Read_12", chunk=700, other_reads=("Read_7",)))
    assert labelop.list_fast5_reads(str(q)) == (1, ["Read_12", "Read_7"])
    assert np.array_equal(labelop.read_fast5_signal(str(q), "Read_12")[1], sig)
    assert np.array_equal(labelop.read_fast5_signal(str(q), "Read_7")[1], np.full(3, -7, np.int16))


def test_generic_dataset_reader(tmp_path):
    """nd_h5_read_dataset on layouts other than the Signal's: 2-D contiguous doubles (as in the libhdf5-written file),
    compact and chunked + deflate datasets in nested old-style groups, big-endian integers"""
    w = hw.H5Writer(userblock=512)
    a = np.arange(12, dtype="<f8").reshape(3, 4) / 7.0
    b = np.arange(-5, 6, dtype="<i2")
    c = dac(5000, 3)
    e = (np.arange(40) * 1000 - 7).astype(">i4")
    inner = w.group_old({"a": w.dataset_contiguous(a, w.msg_float64()), "b": w.dataset_compact(b, w.msg_fixed(2)),
                         "c": w.dataset_chunked(c, w.msg_fixed(2), 512, filters=(2, 1)),
                         "e": w.dataset_contiguous(e, w.msg_fixed(4, big_endian=True))})[0]
    root, btree, heap = w.group_old({"grp": w.group_old({"inner": inner})[0]})
    raw = w.finish(root, btree, heap)
    info, data = read_dataset(raw, "/grp/inner/a")
    assert info[:5] == [1, 8, 1, 0, 2] and info[6:8] == [3, 4] and np.array_equal(np.frombuffer(data, "<f8").reshape(3, 4), a)
    info, data = read_dataset(raw, "grp/inner/b/")
    assert info[:5] == [0, 2, 1, 0, 1] and np.array_equal(np.frombuffer(data, "<i2"), b)
    info, data = read_dataset(raw, "/grp/inner/c")
    assert info[5] == c.nbytes and np.array_equal(np.frombuffer(data, "<i2"), c)
    info, data = read_dataset(raw, "/grp/inner/e")
    assert info[:4] == [0, 4, 1, 1] and np.array_equal(np.frombuffer(data, ">i4"), e)     # bytes as stored
    with pytest.raises(RuntimeError, match="not a dataset"):
        read_dataset(raw, "/grp/inner")
    with pytest.raises(RuntimeError, match="not a group"):
        read_dataset(raw, "/grp/inner/a/x")


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_reference_fast5_branch_over_this_reader():
    """the UNMODIFIED reference's `extract_fast5_raw(..., 'fast5')` (utils/labelop.py:199-214) with `h5py.File` served by
    libnanodec's reader gives the chunk strings of its '.signal' branch (oracle/make_golden.py::check_fast5_branch; in a
    subprocess: the harness patches torch for the torch-1.0 reference)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "make_golden.py"), "--fast5"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "== its '.signal' branch on 8 runs" in r.stdout


def test_h5py_shaped_facade(tmp_path):
    """nanodecoder_b200/utils/h5lite.py: the File / Group / Dataset surface the reference's reader touches
    (utils/labelop.py:199-214), over libnanodec's reader"""
    from nanodecoder_b200.utils import h5lite
    sig = dac(7000, 17)
    p = tmp_path / "a.fast5"
    p.write_bytes(hw.make_fast5(sig, read_name="Read_33", chunk=1024, filters=(32020,), kw_vbz_version=1,
                                other_reads=("Read_4", "Read_9")))
    f = h5lite.File(str(p), "r")
    reads = f["/Raw/Reads/"]
    assert list(reads.keys()) == ["Read_33", "Read_4", "Read_9"] and len(reads) == 3 and "Read_4" in reads
    ds = list(reads.values())[0]["Signal"]                                    # the reference's expression
    assert ds.shape == (7000,) and ds.dtype == np.dtype("<i2") and len(ds) == 7000
    assert np.array_equal(ds.value, sig) and np.array_equal(ds[()], sig) and np.array_equal(ds[10:20], sig[10:20])
    assert np.array_equal(np.asarray(ds), sig)
    assert sorted(f.keys()) == ["Analyses", "PreviousReadInfo", "Raw", "UniqueGlobalKey"]
    assert np.array_equal(f["Raw"]["Reads"]["Read_9/Signal"][...], np.full(3, -7, np.int16))
    assert [k for k, _ in f["/UniqueGlobalKey"].items()] == ["channel_id", "context_tags", "tracking_id"]
    with pytest.raises(KeyError):
        f["/Raw/Reads/Read_5"]
    assert "nope" not in f
    f.close()
    with h5lite.File(os.path.join(GOLDEN, "libhdf5_matlab73.mat")) as m:      # the libhdf5-written file
        d = m["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
        assert np.array_equal(d[()][:, 0], np.arange(0, 2 * np.pi + 1e-9, np.pi / 4))
    bad = tmp_path / "bad.fast5"
    bad.write_bytes(b"junk" * 50)
    with pytest.raises(IOError):                                              # what utils/labelop.py:203 catches
        h5lite.File(str(bad), "r")
    with pytest.raises(ValueError):
        h5lite.File(str(p), "w")


def test_committed_fast5_fixtures():
    """tests/golden/fast5_*.fast5: three small files laid out by tests/h5_writer.py (single-read gzip + shuffle, single-read
    VBZ v1, multi-read VBZ v1) with their samples in fast5_samples.npz.  They are committed so that anyone with h5py and
    the VBZ plugin can check the WRITER (whole B-tree nodes, heap free list, end-of-file address as libhdf5 lays them out)
    against libhdf5 (`h5py.File(f)['/Raw/Reads/Read_101/Signal'][()]`) — the one
    cross-check this image cannot run — and they pin the reader against accidental format drift of the writer."""
    want = np.load(os.path.join(GOLDEN, "fast5_samples.npz"))
    name, got = labelop.read_fast5_signal(os.path.join(GOLDEN, "fast5_single_gzip.fast5"))
    assert name == "Read_101" and np.array_equal(got, want["single_gzip"])
    name, got = labelop.read_fast5_signal(os.path.join(GOLDEN, "fast5_single_vbz.fast5"))
    assert name == "Read_102" and np.array_equal(got, want["single_vbz"])
    multi = os.path.join(GOLDEN, "fast5_multi_vbz.fast5")
    layout, names = labelop.list_fast5_reads(multi)
    assert layout == 2 and len(names) == 3
    for nm in names:
        assert np.array_equal(labelop.read_fast5_signal(multi, nm)[1], want["multi_" + nm[5:].replace("-", "_")])
