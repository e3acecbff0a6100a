"""Test infrastructure: a small HDF5 *writer* that lays out single-read .fast5 files the way libhdf5 1.8 does
(superblock v0, old-style groups = symbol-table message -> v1 B-tree -> SNOD nodes + local heap, object headers v1,
chunked datasets through the v1 chunk B-tree, filter pipeline v1 with shuffle / deflate / fletcher32), plus the
`libver='latest'` flavour (superblock v2, object headers v2 with compact link messages, layout v4 single chunk).

It follows the published "HDF5 File Format Specification Version 3.0" independently of the reader in
nanodecoder_b200/csrc/fast5.cu.  The reader's group / heap / B-tree / object-header / contiguous-layout code is ALSO
pinned against a file written by libhdf5 itself (scipy's testhdf5_7.4_GLNX86.mat, tests/golden/libhdf5_matlab73.mat);
the chunk B-tree and the filter pipeline are pinned by this writer and the specification only, because no HDF5 library
exists in this image to write such a file (stated in DESIGN.md §2).

Checksums of v2 structures (Jenkins lookup3) are written as zero: the reader does not verify them.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class H5Writer:
    def __init__(self, userblock=0, superblock=0, leaf_k=4):
        self.buf = bytearray(userblock)
        self.base = userblock
        self.sb_version = superblock
        self.leaf_k = leaf_k                       # at most 2*leaf_k symbols per SNOD
        self.sb_size = {0: 96, 1: 104}.get(superblock, 48)
        self.buf += b"\0" * self.sb_size

    # ---- raw space
    def alloc(self, data, align=8):
        self.buf += b"\0" * (-(len(self.buf) - self.base) % align)
        addr = len(self.buf) - self.base
        self.buf += data
        return addr

    # ---- header messages
    @staticmethod
    def msg_dataspace(dims, maxdims=None):
        flags = 1 if maxdims is not None else 0
        b = struct.pack("<BBBB4x", 1, len(dims), flags, 0)
        b += b"".join(struct.pack("<Q", d) for d in dims)
        if maxdims is not None:
            b += b"".join(struct.pack("<Q", d) for d in maxdims)
        return (0x01, b)

    @staticmethod
    def msg_fixed(size, signed=True, big_endian=False):
        bits0 = (1 if big_endian else 0) | (8 if signed else 0)
        return (0x03, struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, size) + struct.pack("<HH", 0, 8 * size))

    @staticmethod
    def msg_float64():
        # IEEE double, little endian: class 1, bit field (0x20, 0x3f, 0x00), then offset, precision, exponent location / size,
        # mantissa location / size, exponent bias
        return (0x03, struct.pack("<BBBBI", 0x10 | 1, 0x20, 0x3F, 0x00, 8) +
                struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023))

    @staticmethod
    def msg_pipeline(filters):
        """filters: list of (id, name, client values) in the order they are applied when WRITING"""
        b = struct.pack("<BB2x4x", 1, len(filters))
        for fid, name, cd in filters:
            nm = _pad8(name.encode() + b"\0") if name else b""
            b += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + nm
            b += b"".join(struct.pack("<I", v) for v in cd)
            if len(cd) & 1:
                b += b"\0" * 4
        return (0x0B, b)

    @staticmethod
    def msg_attribute_stub():
        """an attribute message the reader has to step over (v1: name 'x', scalar int8 value)"""
        name = _pad8(b"x\0")
        dtype = _pad8(struct.pack("<BBBBI", 0x10, 8, 0, 0, 1) + struct.pack("<HH", 0, 8))
        space = _pad8(struct.pack("<BBBB4x", 1, 0, 0, 0))
        return (0x0C, struct.pack("<BxHHH", 1, 2, len(dtype), len(space)) + name + dtype + space + _pad8(b"\x07"))

    def object_header_v1(self, messages, split_at=None):
        """messages: list of (type, body).  split_at: put messages[split_at:] into a continuation block."""
        def pack(ms):
            out = b""
            for t, body in ms:
                body = _pad8(body)
                out += struct.pack("<HHB3x", t, len(body), 0) + body
            return out
        n = len(messages)
        if split_at is None:
            block = pack(messages)
        else:
            cont = pack(messages[split_at:])
            cont_addr = self.alloc(cont)
            block = pack(messages[:split_at] + [(0x10, struct.pack("<QQ", cont_addr, len(cont)))])
            n += 1
        hdr = struct.pack("<BxHII4x", 1, n, 1, len(block))
        return self.alloc(hdr + block)

    def object_header_v2(self, messages, split_at=None, times=False):
        def pack(ms):
            return b"".join(struct.pack("<BHB", t, len(body), 0) + body for t, body in ms)
        flags = 0x02 | (0x20 if times else 0)                       # 4-byte chunk-0 size
        if split_at is not None:
            cont = b"OCHK" + pack(messages[split_at:]) + b"\0\0\0\0"
            cont_addr = self.alloc(cont)
            messages = messages[:split_at] + [(0x10, struct.pack("<QQ", cont_addr, len(cont)))]
        block = pack(messages) + b"\0\0"                              # a 2-byte gap (smaller than a message header)
        hdr = b"OHDR" + struct.pack("<BB", 2, flags) + (b"\0" * 16 if times else b"") + struct.pack("<I", len(block))
        return self.alloc(hdr + block + b"\0\0\0\0")

    # ---- groups
    def group_old(self, members, levels=1):
        """members: {name: object header address}.  Returns the group's object header address."""
        names = sorted(members, key=lambda s: s.encode())
        heap_data = bytearray(8)                                       # offset 0: the empty string
        offs = {}
        for nm in names:
            offs[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b"\0")
        heap_data += struct.pack("<QQ", 1, 16)                         # the one free block: next = 1 (none), 16 bytes
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), len(heap_data) - 16, data_addr))
        per = 2 * self.leaf_k
        groups = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        children = []                                                  # (address, heap offset of the largest name below)
        for g in groups:
            ent = b""
            for nm in g:
                ent += struct.pack("<QQII16x", offs[nm], members[nm], 0, 0)
            ent += b"\0" * (40 * (per - len(g)))
            children.append((self.alloc(b"SNOD" + struct.pack("<BxH", 1, len(g)) + ent), offs[g[-1]] if g else 0))
        level = 0
        while True:
            fan = 2 if levels > 1 else len(children)                  # levels > 1: a binary tree over the SNODs
            nodes = []
            for i in range(0, len(children), fan):
                part = children[i:i + fan]
                body = struct.pack("<Q", 0)                            # key 0: empty string sorts before everything
                for addr, key in part:
                    body += struct.pack("<QQ", addr, key)
                node = b"TREE" + struct.pack("<BBHQQ", 0, level, len(part), UNDEF, UNDEF) + body
                node += b"\0" * (24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 - len(node))   # libhdf5 reads whole nodes (K = 16)
                nodes.append((self.alloc(node), part[-1][1]))
            children = nodes
            level += 1
            if len(children) == 1:
                break
        btree = children[0][0]
        return self.object_header_v1([(0x11, struct.pack("<QQ", btree, heap))]), btree, heap

    def group_new(self, members, order=None):
        """compact link messages, in the given (creation) order"""
        msgs = [(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)),      # link info: no fractal heap = compact storage
                (0x0A, struct.pack("<BBHH", 0, 0, 8, 6))]              # group info
        for nm in (order or list(members)):
            b = nm.encode()
            msgs.append((0x06, struct.pack("<BBB", 1, 0, len(b)) + b + struct.pack("<Q", members[nm])))
        return self.object_header_v2(msgs, times=True)

    # ---- datasets
    def dataset_contiguous(self, arr, type_msg, v2=False, layout_version=3):
        raw = arr.tobytes()
        addr = self.alloc(raw)
        if layout_version == 3:
            lay = struct.pack("<BBQQ", 3, 1, addr, len(raw))
        else:                                   # v1 / v2: dimensionality (rank + 1), class, 5 reserved, address, dims + element
            lay = (struct.pack("<BBB5xQ", layout_version, arr.ndim + 1, 1, addr) +   # size, as libhdf5 writes them (the
                   b"".join(struct.pack("<I", d) for d in arr.shape + (arr.dtype.itemsize,)))  # MATLAB file's layout message)
        msgs = [self.msg_dataspace(arr.shape), type_msg, (0x08, lay)]
        return (self.object_header_v2 if v2 else self.object_header_v1)(msgs)

    def dataset_compact(self, arr, type_msg):
        raw = arr.tobytes()
        msgs = [self.msg_dataspace(arr.shape), type_msg, (0x08, struct.pack("<BBH", 3, 0, len(raw)) + raw)]
        return self.object_header_v1(msgs)

    @staticmethod
    def encode_chunk(raw, elem, filters, level=1, strategy=zlib.Z_DEFAULT_STRATEGY, vbz_version=0):
        for fid in filters:
            if fid == 2:                                               # shuffle: byte j of every element together
                a = np.frombuffer(raw, np.uint8)
                ne = len(a) // elem
                raw = a[:ne * elem].reshape(ne, elem).T.tobytes() + a[ne * elem:].tobytes()
            elif fid == 1:
                c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
                raw = c.compress(raw) + c.flush()
            elif fid == 3:
                raw = raw + b"\xde\xad\xbe\xef"                        # the reader strips, does not verify
            elif fid == 32020:
                raw = vbz_encode(raw, elem, level, version=vbz_version)
            else:                                                      # unknown filter: stored as is, the reader must refuse
                pass
        return raw

    def dataset_chunked(self, arr, type_msg, chunk, filters=(2, 1), level=1, strategy=zlib.Z_DEFAULT_STRATEGY, fan=64,
                        skip_filter_on=(), missing=(), extra_messages=(), split_at=None, layout_version=3,
                        kw_vbz_version=0):
        """1-D chunked dataset.  skip_filter_on: chunk numbers stored with every filter skipped (filter mask set);
        missing: chunk numbers never written (read back as zeros)."""
        n, elem = arr.shape[0], arr.dtype.itemsize
        entries = []                                                   # (stored size, mask, element offset, address)
        for c in range((n + chunk - 1) // chunk):
            if c in missing:
                continue
            part = np.zeros(chunk, arr.dtype)
            seg = arr[c * chunk:(c + 1) * chunk]
            part[:len(seg)] = seg
            if c in skip_filter_on:
                raw, mask = part.tobytes(), (1 << len(filters)) - 1
            else:
                raw, mask = self.encode_chunk(part.tobytes(), elem, filters, level, strategy, kw_vbz_version), 0
            entries.append((len(raw), mask, c * chunk, self.alloc(raw)))

        def key(e):
            return struct.pack("<IIQQ", e[0], e[1], e[2], 0)
        level_no, children = 0, entries
        while True:
            nodes = []
            for i in range(0, max(len(children), 1), fan):
                part = children[i:i + fan]
                body = b""
                for e in part:
                    body += key(e) + struct.pack("<Q", e[3])
                body += struct.pack("<IIQQ", 0, 0, (part[-1][2] + chunk) if part else 0, 0)   # closing key
                node = b"TREE" + struct.pack("<BBHQQ", 1, level_no, len(part), UNDEF, UNDEF) + body
                node += b"\0" * max(0, 24 + (2 * 32 + 1) * 24 + 2 * 32 * 8 - len(node))   # whole node at the default K = 32
                first = part[0] if part else (0, 0, 0, 0)
                nodes.append((first[0], first[1], first[2], self.alloc(node)))
            children = nodes
            level_no += 1
            if len(children) == 1:
                break
        btree = children[0][3] if entries else UNDEF
        names = {1: "deflate", 2: "shuffle", 3: "fletcher32", 32020: "vbz"}
        cds = {1: [level], 2: [elem], 3: [], 32020: [kw_vbz_version, elem, 1, level]}
        if layout_version == 3:
            lay = struct.pack("<BBBQ", 3, 2, 2, btree) + struct.pack("<II", chunk, elem)
        else:                                                          # v1 / v2: dimensionality, class, reserved, address, dims
            lay = struct.pack("<BBB5xQ", layout_version, 2, 2, btree) + struct.pack("<II", chunk, elem)
        msgs = [self.msg_dataspace((n,), (UNDEF,)), type_msg] + list(extra_messages) + [(0x08, lay)]
        if filters:
            msgs.append(self.msg_pipeline([(f, names.get(f, ""), cds.get(f, [])) for f in filters]))
        return self.object_header_v1(msgs, split_at=split_at)

    def dataset_single_chunk_v4(self, arr, type_msg, filters=(2, 1), level=1, implicit_chunk=None):
        n, elem = arr.shape[0], arr.dtype.itemsize
        if implicit_chunk:                                             # implicit index: all chunks back to back, no filters
            nch = (n + implicit_chunk - 1) // implicit_chunk
            part = np.zeros(nch * implicit_chunk, arr.dtype)
            part[:n] = arr
            addr = self.alloc(part.tobytes())
            lay = struct.pack("<BBBBB", 4, 2, 0, 2, 4) + struct.pack("<II", implicit_chunk, elem) + struct.pack("<BQ", 2, addr)
            msgs = [self.msg_dataspace((n,)), type_msg, (0x08, lay)]
            return self.object_header_v2(msgs)
        raw = self.encode_chunk(arr.tobytes(), elem, filters, level)
        addr = self.alloc(raw)
        lay = struct.pack("<BBBBB", 4, 2, 0x02 if filters else 0, 2, 4) + struct.pack("<II", n, elem) + struct.pack("<B", 1)
        if filters:
            lay += struct.pack("<QI", len(raw), 0)
        lay += struct.pack("<Q", addr)
        msgs = [(0x01, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", n)), type_msg, (0x08, lay)]
        if filters:
            b = struct.pack("<BB", 2, len(filters))
            for f in filters:
                cd = {1: [level], 2: [elem], 3: []}[f]
                b += struct.pack("<HHH", f, 1, len(cd)) + b"".join(struct.pack("<I", v) for v in cd)
            msgs.append((0x0B, b))
        return self.object_header_v2(msgs, split_at=2)

    # ---- finish
    def finish(self, root, btree=UNDEF, heap=UNDEF, cache_root=True):
        eof = len(self.buf)            # stored as in the libhdf5-written MATLAB file: base address + relative end (= file size)
        if self.sb_version in (0, 1):
            sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", self.sb_version, 0, 0, 0, 8, 8, self.leaf_k, 16, 0)
            if self.sb_version == 1:
                sb += struct.pack("<H2x", 32)
            sb += struct.pack("<QQQQ", self.base, UNDEF, eof, UNDEF)
            if cache_root and btree != UNDEF:
                sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
            else:
                sb += struct.pack("<QQII16x", 0, root, 0, 0)
        else:
            sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", self.sb_version, 8, 8, 0)
            sb += struct.pack("<QQQQI", self.base, UNDEF, eof, root, 0)
        assert len(sb) <= self.sb_size, (len(sb), self.sb_size)
        self.buf[self.base:self.base + len(sb)] = sb
        return bytes(self.buf)


def zstd_compress(raw, level=1):
    """libzstd itself (the copy bundled with pyarrow)"""
    import pyarrow as pa
    return pa.Codec("zstd", compression_level=level).compress(raw, asbytes=True)


def zstd_frame_features(blob, found=None):
    """Which parts of the format one libzstd-written frame uses (block types, literal section types, stream counts,
    Huffman weight encodings, sequence table modes): the decoder test asserts its corpus covers all of them."""
    from collections import Counter
    found = Counter() if found is None else found
    o = 4
    fhd = blob[o]
    single, fcs = (fhd >> 5) & 1, fhd >> 6
    o += 1 + (0 if single else 1) + [0, 1, 2, 4][fhd & 3] + ((1 if single else 0) if fcs == 0 else [0, 2, 4, 8][fcs])
    while True:
        bh = int.from_bytes(blob[o:o + 3], "little")
        o += 3
        last, kind, size = bh & 1, (bh >> 1) & 3, bh >> 3
        found["block_" + ("raw", "rle", "compressed")[kind]] += 1
        if kind == 2:
            p = blob[o:o + size]
            lt, fmt = p[0] & 3, (p[0] >> 2) & 3
            found["literals_" + ("raw", "rle", "huffman", "treeless")[lt]] += 1
            if lt < 2:
                regen, q = ((p[0] >> 3, 1) if fmt & 1 == 0 else ((p[0] >> 4) | (p[1] << 4), 2) if fmt == 1 else
                            ((p[0] >> 4) | (p[1] << 4) | (p[2] << 12), 3))
                q += regen if lt == 0 else 1
            else:
                h = int.from_bytes(p[:5], "little")
                hdr, comp = ((3, (h >> 14) & 0x3FF) if fmt < 2 else (4, (h >> 18) & 0x3FFF) if fmt == 2 else
                             (5, (h >> 22) & 0x3FFFF))
                found["streams_%d" % (1 if fmt == 0 else 4)] += 1
                if lt == 2:
                    found["weights_fse" if p[hdr] < 128 else "weights_direct"] += 1
                q = hdr + comp
            ns = p[q]
            if ns == 0:
                found["no_sequences"] += 1
            else:
                q += 1 if ns < 128 else 2 if ns < 255 else 3
                found["nseq_%d_byte" % (1 if ns < 128 else 2 if ns < 255 else 3)] += 1
                for name, sh in (("ll", 6), ("of", 4), ("ml", 2)):
                    found[name + "_" + ("predefined", "rle", "fse", "repeat")[(p[q] >> sh) & 3]] += 1
        o += 1 if kind == 1 else size
        if last:
            return found


def vbz_encode(raw, elem, level=1, zigzag=True, version=0):
    """ONT's VBZ chunk: uint32 size | zstd(streamvbyte(zig-zag(delta(x)))).
    version 0: values widened to 32 bits, Lemire's streamvbyte: one 2-bit key per value (data bytes - 1, four values per
    key byte, low bits first), all keys, then all data bytes little endian.
    version 1 with 16-bit samples ("svb16"): 16-bit wrapping delta and zig-zag, ONE key bit per value (0 = one data byte,
    1 = two), eight values per key byte, low bit first."""
    x = np.frombuffer(raw, {1: "<i1", 2: "<i2", 4: "<i4"}[elem]).astype(np.int64)
    bits = 16 if (version == 1 and elem == 2) else 32
    if zigzag:
        d = np.diff(x, prepend=0)
        d = ((d + 2 ** (bits - 1)) % 2 ** bits - 2 ** (bits - 1)).astype(np.int64)   # the filter's wrapping arithmetic
        u = ((d << 1) ^ (d >> (bits - 1))) & (2 ** bits - 1)
    else:
        u = x & (2 ** bits - 1)
    if bits == 16:
        nbytes = np.where(u < 1 << 8, 1, 2)
        keys = np.zeros((len(u) + 7) // 8 * 8, np.uint8)
        keys[:len(u)] = nbytes - 1
        key_bytes = np.packbits(keys.reshape(-1, 8), axis=1, bitorder="little").tobytes()
    else:
        nbytes = np.where(u < 1 << 8, 1, np.where(u < 1 << 16, 2, np.where(u < 1 << 24, 3, 4)))
        keys = np.zeros((len(u) + 3) // 4 * 4, np.uint8)
        keys[:len(u)] = nbytes - 1
        keys = keys.reshape(-1, 4)
        key_bytes = (keys[:, 0] | (keys[:, 1] << 2) | (keys[:, 2] << 4) | (keys[:, 3] << 6)).astype(np.uint8).tobytes()
    data = b"".join(int(v).to_bytes(int(k), "little") for v, k in zip(u, nbytes))
    body = key_bytes + data
    if level:
        body = zstd_compress(body, level)
    return struct.pack("<I", len(raw)) + body


def make_fast5(signal, read_name="Read_17", chunk=None, filters=(2, 1), level=1, flavour="old", userblock=0,
               other_reads=(), **kw):
    """A single-read fast5 skeleton: /Raw/Reads/<read_name>/Signal plus the sibling groups MinKNOW writes
    (/UniqueGlobalKey/{channel_id,context_tags,tracking_id}, /Analyses, /PreviousReadInfo)."""
    signal = np.asarray(signal)
    cache_root = kw.pop("cache_root", True)
    group_levels = kw.pop("group_levels", 1)
    sb = 2 if flavour == "new" else (1 if flavour == "old_sb1" else 0)
    w = H5Writer(userblock=userblock, superblock=sb)
    t = w.msg_fixed(signal.dtype.itemsize, signed=signal.dtype.kind == "i", big_endian=signal.dtype.byteorder == ">")
    if flavour == "new":
        grp = lambda m, **k: w.group_new(m, **k)                      # noqa: E731
        if chunk is None:
            ds = w.dataset_single_chunk_v4(signal, t, filters=filters, level=level)
        elif chunk == "implicit":
            ds = w.dataset_single_chunk_v4(signal, t, implicit_chunk=kw.pop("implicit_chunk", 100))
        else:
            ds = w.dataset_chunked(signal, t, chunk, filters=filters, level=level, **kw)
    else:
        grp = lambda m, **k: w.group_old(m)[0]                        # noqa: E731
        if chunk == "contiguous":
            ds = w.dataset_contiguous(signal, t, layout_version=kw.pop("layout_version", 3))
        elif chunk == "compact":
            ds = w.dataset_compact(signal, t)
        else:
            ds = w.dataset_chunked(signal, t, chunk or max(1, len(signal)), filters=filters, level=level, **kw)
    reads = {read_name: grp({"Signal": ds})}
    for nm in other_reads:                                             # decoys: h5py's values()[0] is the first in NAME order
        decoy = w.dataset_contiguous(np.full(3, -7, np.int16), w.msg_fixed(2))
        reads[nm] = grp({"Signal": decoy})
    if flavour == "new":
        reads_g = w.group_new(reads, order=list(reversed(sorted(reads))))   # creation order != name order
    else:
        reads_g = w.group_old(reads, levels=group_levels)[0]
    raw_g = grp({"Reads": reads_g})
    empty = lambda: grp({})                                             # noqa: E731
    ugk = grp({"channel_id": empty(), "context_tags": empty(), "tracking_id": empty()})
    members = {"Analyses": empty(), "PreviousReadInfo": empty(), "Raw": raw_g, "UniqueGlobalKey": ugk}
    if flavour == "new":
        return w.finish(w.group_new(members, order=["UniqueGlobalKey", "Raw", "PreviousReadInfo", "Analyses"]))
    root, btree, heap = w.group_old(members)
    return w.finish(root, btree, heap, cache_root=cache_root)


def make_multi_fast5(reads, chunk=4096, filters=(32020,), level=1, vbz_version=1, userblock=0, pad_to=0):
    """A multi-read fast5 skeleton as ont_fast5_api writes it (default libver: old-style groups):
    /read_<id>/Raw/Signal per read plus the per-read metadata groups.  reads: {read id: int16 samples}."""
    w = H5Writer(userblock=userblock)
    members = {}
    for rid, signal in reads.items():
        signal = np.asarray(signal)
        t = w.msg_fixed(signal.dtype.itemsize, signed=signal.dtype.kind == "i")
        ds = w.dataset_chunked(signal, t, min(chunk, max(1, len(signal))), filters=filters, level=level,
                               kw_vbz_version=vbz_version)
        raw_g = w.group_old({"Signal": ds})[0]
        members["read_" + rid] = w.group_old({"Raw": raw_g, "channel_id": w.group_old({})[0],
                                              "context_tags": w.group_old({})[0], "tracking_id": w.group_old({})[0]})[0]
    if pad_to:
        w.alloc(bytes(max(0, pad_to - len(w.buf))))
    root, btree, heap = w.group_old(members, levels=3)
    return w.finish(root, btree, heap)
