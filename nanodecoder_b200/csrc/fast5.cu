// .fast5 ingestion (host code).  The reference reads a read's raw DAC samples with
//     fast5_data = h5py.File(path, 'r'); list(fast5_data['/Raw/Reads/'].values())[0]['Signal'].value
// (utils/labelop.py:199-214).  h5py / libhdf5 are third-party dependencies that are not part of this image, so the subset
// of the published HDF5 file format ("HDF5 File Format Specification Version 3.0") that single-read fast5 files use is
// restated here as a reader over an in-memory copy of the file:
//   superblock versions 0-3 (user block allowed), object headers v1 and v2 (continuation blocks), old-style groups
//   (symbol-table message -> v1 B-tree + SNOD nodes + local heap) and new-style groups with compact link messages,
//   dataspace v1/v2, fixed-point / floating-point datatypes, data layout v1-v3 (compact, contiguous, chunked through the
//   v1 chunk B-tree) and v4 (single-chunk / implicit index), filter pipeline v1/v2 with deflate (own inflate, RFC 1950 /
//   1951),
//   shuffle, fletcher32 and ONT's VBZ (id 32020: zstd over streamvbyte, own decoders in vbz.cu).
// Refused with the reason in the message: dense link storage (fractal heap), the v4 fixed / extensible array and v2
// B-tree chunk indexes.
// h5py iterates a group's members in NAME order (H5_INDEX_NAME, increasing), which is what "the first read" means above:
// the B-tree of an old-style group is already sorted by name; link messages are sorted here.
// Every offset read from the file is bounds-checked: a corrupt file gives an error, never a wild read.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nanodec.h"
#include "host_io.h"

namespace {

typedef ndhost::Error H5Error;
using ndhost::fail;

// ---------------------------------------------------------------- inflate (RFC 1951) + zlib container (RFC 1950)
// Bits are consumed from the low end of a 64-bit window that is topped up a byte at a time.
struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;
  uint64_t acc = 0;
  int cnt = 0;
  BitReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  inline void refill() {
    while (cnt <= 56 && pos < n) {
      acc |= (uint64_t)p[pos++] << cnt;
      cnt += 8;
    }
  }
  inline uint32_t bits(int need) {                    // need <= 16
    if (cnt < need) {
      refill();
      if (cnt < need) fail("deflate stream ends early");
    }
    const uint32_t v = (uint32_t)(acc & ((1ull << need) - 1ull));
    acc >>= need;
    cnt -= need;
    return v;
  }
  inline void drop(int k) { acc >>= k; cnt -= k; }
  void align() { drop(cnt & 7); }                      // to the next byte boundary; whole bytes stay in the window
  size_t byte_pos() const { return pos - (size_t)(cnt >> 3); }
};

// canonical Huffman code: count[len] codes of each length, symbols ordered by (length, value); codes of up to kFastBits
// bits are decoded with one lookup in `fast` (entry = length << 9 | symbol, 0 = longer code: canonical walk).
const int kFastBits = 10;
struct Huffman {
  uint16_t count[16];
  uint16_t symbol[288];
  uint16_t fast[1 << kFastBits];
  bool build(const uint8_t* lens, int n) {
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;                     // over-subscribed
    }
    uint16_t offs[16];
    uint32_t next_code[16];
    offs[1] = 0;
    next_code[1] = 0;
    for (int l = 1; l < 15; ++l) {
      offs[l + 1] = offs[l] + count[l];
      next_code[l + 1] = (next_code[l] + count[l]) << 1;
    }
    memset(fast, 0, sizeof(fast));
    for (int i = 0; i < n; ++i) {
      const int l = lens[i];
      if (!l) continue;
      symbol[offs[l]++] = (uint16_t)i;
      const uint32_t code = next_code[l]++;
      if (l <= kFastBits) {
        uint32_t rev = 0;                             // codes enter the stream most significant bit first
        for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1u) << (l - 1 - b);
        for (uint32_t idx = rev; idx < (1u << kFastBits); idx += 1u << l) fast[idx] = (uint16_t)((l << 9) | i);
      }
    }
    return true;
  }
  inline int decode(BitReader& br) const {
    if (br.cnt < 15) br.refill();
    const uint16_t e = fast[br.acc & ((1u << kFastBits) - 1u)];
    if (e) {
      const int l = e >> 9;
      if (l > br.cnt) fail("deflate stream ends early");
      br.drop(l);
      return e & 511;
    }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
      code |= (int)br.bits(1);
      const int c = count[l];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first = (first + c) << 1;
      code <<= 1;
    }
    fail("invalid Huffman code in deflate stream");
  }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                               131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
                                2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12,
                                13, 13};

// `out` is sized to `limit` up front; `have` counts the bytes produced so far.
void inflate_codes(BitReader& br, const Huffman& lit, const Huffman& dist, uint8_t* out, size_t& have, size_t limit) {
  for (;;) {
    const int sym = lit.decode(br);
    if (sym < 256) {
      if (have >= limit) fail("deflate stream is longer than the chunk it fills");
      out[have++] = (uint8_t)sym;
    } else if (sym == 256) {
      return;
    } else {
      if (sym > 285) fail("invalid length symbol in deflate stream");
      const size_t len = kLenBase[sym - 257] + br.bits(kLenExtra[sym - 257]);
      const int ds = dist.decode(br);
      if (ds > 29) fail("invalid distance symbol in deflate stream");
      const size_t d = kDistBase[ds] + br.bits(kDistExtra[ds]);
      if (d > have) fail("deflate distance reaches before the start of the output");
      if (have + len > limit) fail("deflate stream is longer than the chunk it fills");
      uint8_t* dst = out + have;
      const uint8_t* src = dst - d;
      for (size_t i = 0; i < len; ++i) dst[i] = src[i];                   // overlapping copies repeat, byte by byte
      have += len;
    }
  }
}

void inflate_raw(BitReader& br, uint8_t* out, size_t& have, size_t limit) {
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  Huffman lit, dist;
  for (;;) {
    const uint32_t last = br.bits(1);
    const uint32_t type = br.bits(2);
    if (type == 0) {
      br.align();
      const uint32_t len = br.bits(16), nlen = br.bits(16);
      if ((len ^ 0xffffu) != nlen) fail("stored deflate block with inconsistent length");
      if (have + len > limit) fail("deflate stream is longer than the chunk it fills");
      uint32_t todo = len;
      while (todo && br.cnt) { out[have++] = (uint8_t)br.bits(8); --todo; }   // bytes already in the window
      if (br.pos + todo > br.n) fail("deflate stream ends early");
      memcpy(out + have, br.p + br.pos, todo);
      have += todo;
      br.pos += todo;
    } else if (type == 1) {
      uint8_t lens[288];
      for (int i = 0; i < 144; ++i) lens[i] = 8;
      for (int i = 144; i < 256; ++i) lens[i] = 9;
      for (int i = 256; i < 280; ++i) lens[i] = 7;
      for (int i = 280; i < 288; ++i) lens[i] = 8;
      lit.build(lens, 288);
      uint8_t dl[30];
      for (int i = 0; i < 30; ++i) dl[i] = 5;
      dist.build(dl, 30);
      inflate_codes(br, lit, dist, out, have, limit);
    } else if (type == 2) {
      const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
      if (nlen > 286 || ndist > 30) fail("deflate block with too many codes");
      uint8_t lens[320];
      memset(lens, 0, sizeof(lens));
      for (int i = 0; i < ncode; ++i) lens[order[i]] = (uint8_t)br.bits(3);
      Huffman cl;
      if (!cl.build(lens, 19)) fail("deflate code-length code is over-subscribed");
      uint8_t all[320];
      int idx = 0;
      while (idx < nlen + ndist) {
        const int sym = cl.decode(br);
        if (sym < 16) {
          all[idx++] = (uint8_t)sym;
        } else {
          uint8_t val = 0;
          int rep;
          if (sym == 16) {
            if (idx == 0) fail("deflate repeat code without a previous length");
            val = all[idx - 1];
            rep = 3 + (int)br.bits(2);
          } else if (sym == 17) {
            rep = 3 + (int)br.bits(3);
          } else {
            rep = 11 + (int)br.bits(7);
          }
          if (idx + rep > nlen + ndist) fail("deflate code lengths overrun");
          while (rep--) all[idx++] = val;
        }
      }
      if (all[256] == 0) fail("deflate block without an end code");
      if (!lit.build(all, nlen)) fail("deflate literal code is over-subscribed");
      if (!dist.build(all + nlen, ndist)) fail("deflate distance code is over-subscribed");
      inflate_codes(br, lit, dist, out, have, limit);
    } else {
      fail("reserved deflate block type");
    }
    if (last) return;
  }
}

uint32_t adler32(const uint8_t* p, size_t n) {
  uint32_t a = 1, b = 0;
  while (n) {
    const size_t k = n < 5552 ? n : 5552;            // largest run that cannot overflow 32 bits
    for (size_t i = 0; i < k; ++i) { a += p[i]; b += a; }
    a %= 65521u; b %= 65521u;
    p += k; n -= k;
  }
  return (b << 16) | a;
}

void zlib_decompress(const uint8_t* p, size_t n, std::vector<uint8_t>& out, size_t limit) {
  if (n < 6) fail("zlib stream shorter than its header and checksum");
  const uint32_t cmf = p[0], flg = p[1];
  if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u) fail("not a zlib (deflate) stream");
  if (flg & 0x20u) fail("zlib stream with a preset dictionary");
  BitReader br(p + 2, n - 2);
  out.resize(limit);
  size_t have = 0;
  inflate_raw(br, out.data(), have, limit);
  out.resize(have);
  br.align();
  const size_t at = br.byte_pos();
  if (at + 4 > br.n) fail("zlib stream without its adler32 checksum");
  const uint8_t* c = br.p + at;
  const uint32_t want = ((uint32_t)c[0] << 24) | ((uint32_t)c[1] << 16) | ((uint32_t)c[2] << 8) | c[3];
  if (adler32(out.data(), out.size()) != want) fail("zlib adler32 checksum mismatch");
}

// ---------------------------------------------------------------- HDF5 structures
const uint64_t kUndef = ~0ull;

struct Filter {
  int id;
  std::vector<uint32_t> client;
};

struct Dataset {
  int type_class = -1;          // 0 fixed point, 1 floating point
  uint32_t elem = 0;            // bytes per element
  bool is_signed = false, big_endian = false;
  std::vector<uint64_t> dims;
  bool have_space = false, have_type = false, have_layout = false;
  int layout_class = -1;        // 0 compact, 1 contiguous, 2 chunked (v1 B-tree), 3 chunked (v4 single chunk / implicit)
  uint64_t addr = kUndef, size = 0;                 // contiguous / compact
  std::vector<uint64_t> chunk;                      // chunk dims incl. the trailing element-size "dimension" (v1-v3)
  uint64_t single_filtered_size = 0;                // v4 single chunk with filters
  uint32_t single_filter_mask = 0;
  int v4_index = 0;
  std::vector<Filter> filters;
};

struct Link {
  std::string name;
  uint64_t addr;
};

struct Message {
  int type;
  uint64_t off, size;
  int flags;                    // bit 1: the body is a reference to a shared message in another object
};

class H5File {
 public:
  H5File(const uint8_t* p, uint64_t n) : p_(p), n_(n) { open(); }

  uint64_t root() const { return root_; }

  std::vector<Link> links(uint64_t header_addr, bool sorted = true) {
    std::vector<Link> out;
    bool is_group = false;
    uint64_t cached_btree = kUndef, cached_heap = kUndef;
    if (header_addr == root_ && root_btree_ != kUndef) { cached_btree = root_btree_; cached_heap = root_heap_; }
    for (const Message& m : messages(header_addr)) {
      if (m.type == 0x11) {                                   // symbol table message: old-style group
        need(m.off, 2ull * so_);
        group_btree(get(m.off, so_), get(m.off + so_, so_), out, 0);
        is_group = true;
        cached_btree = kUndef;
      } else if (m.type == 0x06) {                            // link message (compact new-style group)
        Link l;
        if (link_message(m, l)) out.push_back(l);
        is_group = true;
      } else if (m.type == 0x02) {                            // link info: dense storage lives in a fractal heap
        need(m.off, 2);
        const int flags = p_[m.off + 1];
        uint64_t o = m.off + 2 + ((flags & 1) ? 8 : 0);
        need(o, so_);
        if (get(o, so_) != undef_) fail("group with dense link storage (fractal heap) is not supported");
        is_group = true;
      }
    }
    if (cached_btree != kUndef && out.empty()) { group_btree(cached_btree, cached_heap, out, 0); is_group = true; }
    if (!is_group) fail("object is not a group");
    if (sorted) std::stable_sort(out.begin(), out.end(), [](const Link& a, const Link& b) { return a.name < b.name; });
    return out;
  }

  // "/a/b/c" from the root group; an empty component is skipped (h5py accepts a trailing '/')
  uint64_t resolve(const std::string& path) { return resolve(path, root_); }
  uint64_t resolve(const std::string& path, uint64_t cur) {
    size_t i = 0;
    while (i < path.size()) {
      size_t j = path.find('/', i);
      if (j == std::string::npos) j = path.size();
      if (j > i) {
        const std::string name = path.substr(i, j - i);
        bool found = false;
        for (const Link& l : links(cur, false))
          if (l.name == name) { cur = l.addr; found = true; break; }
        if (!found) fail("no object named '" + name + "' on the path '" + path + "'");
      }
      i = j + 1;
    }
    return cur;
  }

  Dataset dataset(uint64_t header_addr) {
    Dataset d;
    for (const Message& m : messages(header_addr)) {
      if ((m.flags & 2) && (m.type == 0x01 || m.type == 0x03 || m.type == 0x0B))
        fail("dataset with a shared (committed) dataspace / datatype / filter message is not supported");
      if (m.type == 0x01) dataspace(m, d);
      else if (m.type == 0x03) datatype(m, d);
      else if (m.type == 0x08) layout(m, d);
      else if (m.type == 0x0B) pipeline(m, d);
    }
    if (!d.have_space || !d.have_type || !d.have_layout) fail("object is not a dataset (dataspace, datatype or layout missing)");
    return d;
  }

  static bool has_vbz(const Dataset& d) {
    for (const Filter& f : d.filters)
      if (f.id == 32020) return true;
    return false;
  }

  uint64_t count(const Dataset& d) const {
    uint64_t c = 1;
    for (uint64_t x : d.dims) {
      if (x && c > (1ull << 40) / x) fail("dataset too large");
      c *= x;
    }
    // deflate expands at most ~1032 x: a dataspace far beyond what the file could hold is a corrupt header, not a read
    // (zstd has no such bound - a constant chunk is one RLE block - so VBZ datasets get an absolute cap instead)
    if (has_vbz(d) ? c * d.elem > (1ull << 31) : c * d.elem > n_ * 1100 + (1ull << 20))
      fail("dataspace larger than the file can hold");
    return c;
  }

  void read(const Dataset& d, std::vector<uint8_t>& out) {
    const uint64_t total = count(d) * d.elem;
    out.assign(total, 0);                                                   // unwritten chunks read as the fill value 0
    if (total == 0) return;
    if (d.layout_class == 0 || d.layout_class == 1) {
      if (d.addr == kUndef) return;                                         // never written
      if (!d.filters.empty()) fail("filters on a dataset that is not chunked");
      if (d.size < total) fail("dataset storage smaller than its dataspace");
      need_abs(d.addr, total);
      memcpy(out.data(), p_ + d.addr, total);
    } else if (d.layout_class == 2) {
      if (d.addr == kUndef) return;
      if (d.dims.size() != 1 || d.chunk.size() != 2) fail("chunked datasets are read for rank 1 only (Signal is 1-D)");
      if (d.chunk[0] == 0 || d.chunk[1] != d.elem) fail("inconsistent chunk shape");
      chunk_btree(d, base_ + d.addr, out, 0);
    } else {
      if (d.addr == kUndef) return;
      if (d.dims.size() != 1 || d.chunk.size() != 1 || d.chunk[0] == 0) fail("chunked datasets are read for rank 1 only (Signal is 1-D)");
      const uint64_t cbytes = d.chunk[0] * d.elem;
      if (d.v4_index == 1) {                                                // single chunk
        const uint64_t stored = d.filters.empty() ? cbytes : d.single_filtered_size;
        place_chunk(d, base_ + d.addr, stored, d.single_filter_mask, 0, out);
      } else {                                                              // implicit: chunks back to back, no filters
        if (!d.filters.empty()) fail("implicit chunk index with filters");
        const uint64_t nchunks = (d.dims[0] + d.chunk[0] - 1) / d.chunk[0];
        for (uint64_t c = 0; c < nchunks; ++c) place_chunk(d, base_ + d.addr + c * cbytes, cbytes, 0, c * d.chunk[0], out);
      }
    }
  }

 private:
  const uint8_t* p_;
  uint64_t n_;
  int so_ = 8, sl_ = 8;                    // size of offsets / lengths
  uint64_t base_ = 0, root_ = 0, undef_ = kUndef;
  uint64_t root_btree_ = kUndef, root_heap_ = kUndef;
  uint64_t visits_ = 0;                    // B-tree nodes walked: a corrupt child pointer must not turn into a long loop
  void visit() {
    if (++visits_ > n_ / 16 + 1024) fail("B-tree walk does not end (corrupt child pointers)");
  }

  void need_abs(uint64_t off, uint64_t len) const {
    if (off > n_ || len > n_ - off) fail("structure at offset " + std::to_string(off) + " (+" + std::to_string(len) + ") reaches past the end of the file");
  }
  void need(uint64_t off, uint64_t len) const { need_abs(off, len); }
  uint64_t get(uint64_t off, int nbytes) const {
    need_abs(off, (uint64_t)nbytes);
    uint64_t v = 0;
    for (int i = 0; i < nbytes; ++i) v |= (uint64_t)p_[off + i] << (8 * i);
    return v;
  }
  bool sig(uint64_t off, const char* s) const { return off + 4 <= n_ && memcmp(p_ + off, s, 4) == 0; }

  void open() {
    static const uint8_t magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    uint64_t sb = kUndef;
    for (uint64_t off = 0; off + 8 <= n_; off = off ? off * 2 : 512)         // 0, 512, 1024, 2048, ...
      if (memcmp(p_ + off, magic, 8) == 0) { sb = off; break; }
    if (sb == kUndef) fail("HDF5 signature not found");
    const int version = (int)get(sb + 8, 1);
    if (version == 0 || version == 1) {
      so_ = (int)get(sb + 13, 1);
      sl_ = (int)get(sb + 14, 1);
      check_sizes();
      uint64_t o = sb + 24 + (version == 1 ? 4 : 0);
      base_ = get(o, so_);
      o += 4ull * so_;                                                      // base, free-space info, end of file, driver info
      // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch pad
      root_ = get(o + so_, so_);
      const uint32_t cache = (uint32_t)get(o + 2ull * so_, 4);
      if (cache == 1) {
        root_btree_ = get(o + 2ull * so_ + 8, so_);
        root_heap_ = get(o + 3ull * so_ + 8, so_);
      }
    } else if (version == 2 || version == 3) {
      so_ = (int)get(sb + 9, 1);
      sl_ = (int)get(sb + 10, 1);
      check_sizes();
      base_ = get(sb + 12, so_);
      root_ = get(sb + 12 + 3ull * so_, so_);
    } else {
      fail("unknown HDF5 superblock version " + std::to_string(version));
    }
    undef_ = so_ == 8 ? ~0ull : ((1ull << (8 * so_)) - 1);
    if (base_ == undef_) base_ = 0;
    if (base_ == 0 && sb != 0) base_ = sb;                                  // relative to the superblock when not recorded
    if (root_ == undef_) fail("file without a root group");
  }
  void check_sizes() const {
    if (!((so_ == 2 || so_ == 4 || so_ == 8) && (sl_ == 2 || sl_ == 4 || sl_ == 8))) fail("unsupported size of offsets / lengths");
  }
  uint64_t addr_field(uint64_t off) const {                                  // file address -> kUndef or relative address
    const uint64_t a = get(off, so_);
    return a == undef_ ? kUndef : a;
  }

  // All header messages of an object, continuation blocks followed (absolute offsets of the message bodies).
  std::vector<Message> messages(uint64_t rel_addr) {
    std::vector<Message> out;
    const uint64_t a = base_ + rel_addr;
    need(a, 16);
    if (sig(a, "OHDR")) {
      if (p_[a + 4] != 2) fail("unknown object header version");
      const int flags = p_[a + 5];
      uint64_t o = a + 6;
      if (flags & 0x20) o += 16;
      if (flags & 0x10) o += 4;
      const int szb = 1 << (flags & 3);
      const uint64_t chunk0 = get(o, szb);
      o += szb;
      std::vector<std::pair<uint64_t, uint64_t>> blocks;                      // [start, end) of message areas
      need(o, chunk0 + 4);
      blocks.push_back({o, o + chunk0});
      const uint64_t hdr = 4ull + ((flags & 0x04) ? 2 : 0);
      for (size_t b = 0; b < blocks.size(); ++b) {
        if (blocks.size() > 4096) fail("object header continuation loop");
        uint64_t q = blocks[b].first;
        const uint64_t end = blocks[b].second;
        while (q + hdr <= end) {
          const int type = p_[q];
          const uint64_t size = get(q + 1, 2);
          const uint64_t body = q + hdr;
          if (body + size > end) fail("object header message overruns its block");
          if (type == 0x10) {
            const uint64_t ca = get(body, so_), cl = get(body + so_, sl_);
            const uint64_t s = base_ + ca;
            need(s, cl);
            if (cl < 8 || !sig(s, "OCHK")) fail("object header continuation without its signature");
            blocks.push_back({s + 4, s + cl - 4});
          } else if (type != 0) {
            out.push_back({type, body, size, (int)p_[q + 3]});
          }
          q = body + size;
        }
      }
      return out;
    }
    if (p_[a] != 1) fail("object header not found at offset " + std::to_string(a));
    const uint64_t nmsg = get(a + 2, 2);
    const uint64_t hsize = get(a + 8, 4);
    std::vector<std::pair<uint64_t, uint64_t>> blocks;
    need(a + 16, hsize);
    blocks.push_back({a + 16, a + 16 + hsize});
    uint64_t seen = 0;
    for (size_t b = 0; b < blocks.size(); ++b) {
      if (blocks.size() > 4096) fail("object header continuation loop");
      uint64_t q = blocks[b].first;
      const uint64_t end = blocks[b].second;
      while (q + 8 <= end && seen < nmsg) {
        const int type = (int)get(q, 2);
        const uint64_t size = get(q + 2, 2);
        const uint64_t body = q + 8;
        if (body + size > end) fail("object header message overruns its block");
        ++seen;
        if (type == 0x10) {
          const uint64_t ca = get(body, so_), cl = get(body + so_, sl_);
          need(base_ + ca, cl);
          blocks.push_back({base_ + ca, base_ + ca + cl});
        } else if (type != 0) {
          out.push_back({type, body, size, (int)p_[q + 4]});
        }
        q = body + size;                                                    // v1 message sizes are multiples of 8
      }
    }
    return out;
  }

  std::string heap_string(uint64_t heap_rel, uint64_t off) const {
    const uint64_t h = base_ + heap_rel;
    need(h, 8 + 2ull * sl_ + so_);
    if (!sig(h, "HEAP")) fail("local heap signature missing");
    const uint64_t dsize = get(h + 8, sl_);
    const uint64_t daddr = base_ + get(h + 8 + 2ull * sl_, so_);
    need(daddr, dsize);
    if (off >= dsize) fail("link name outside the local heap");
    const uint8_t* s = p_ + daddr + off;
    const void* z = memchr(s, 0, dsize - off);
    if (!z) fail("unterminated link name in the local heap");
    return std::string((const char*)s, (const uint8_t*)z - s);
  }

  void group_btree(uint64_t node_rel, uint64_t heap_rel, std::vector<Link>& out, int depth) {
    if (depth > 32) fail("group B-tree too deep");
    visit();
    const uint64_t a = base_ + node_rel;
    need(a, 8 + 2ull * so_);
    if (sig(a, "SNOD")) {
      const uint64_t nsym = get(a + 6, 2);
      const uint64_t esz = 2ull * so_ + 24;
      need(a + 8, nsym * esz);
      for (uint64_t i = 0; i < nsym; ++i) {
        const uint64_t e = a + 8 + i * esz;
        out.push_back({heap_string(heap_rel, get(e, so_)), get(e + so_, so_)});
      }
      return;
    }
    if (!sig(a, "TREE")) fail("group B-tree node signature missing");
    if (p_[a + 4] != 0) fail("group B-tree node of the wrong type");
    const uint64_t used = get(a + 6, 2);
    uint64_t q = a + 8 + 2ull * so_;
    need(q, used * (sl_ + so_) + sl_);
    for (uint64_t i = 0; i < used; ++i) {
      q += sl_;                                                             // key i
      group_btree(get(q, so_), heap_rel, out, depth + 1);
      q += so_;
    }
  }

  bool link_message(const Message& m, Link& l) const {
    uint64_t o = m.off;
    const uint64_t end = m.off + m.size;
    if (m.size < 4 || p_[o] != 1) fail("unknown link message version");
    const int flags = p_[o + 1];
    o += 2;
    int type = 0;
    if (flags & 0x08) type = p_[o++];
    if (flags & 0x04) o += 8;
    if (flags & 0x10) o += 1;
    const int lb = 1 << (flags & 3);
    if (o + lb > end) fail("truncated link message");
    const uint64_t len = get(o, lb);
    o += lb;
    if (o + len > end) fail("truncated link message");
    l.name.assign((const char*)p_ + o, len);
    o += len;
    if (type != 0) return false;                                            // soft / external links are not followed
    if (o + so_ > end) fail("truncated link message");
    l.addr = get(o, so_);
    return true;
  }

  void dataspace(const Message& m, Dataset& d) const {
    need(m.off, 4);
    const int version = p_[m.off], rank = p_[m.off + 1];
    uint64_t o;
    if (version == 1) o = m.off + 8;
    else if (version == 2) {
      o = m.off + 4;
      if (p_[m.off + 3] == 2) fail("dataset with a null dataspace");
    } else fail("unknown dataspace version");
    if (rank > 32) fail("dataspace rank too large");
    need(o, (uint64_t)rank * sl_);
    d.dims.clear();
    for (int i = 0; i < rank; ++i) d.dims.push_back(get(o + (uint64_t)i * sl_, sl_));
    d.have_space = true;
  }

  void datatype(const Message& m, Dataset& d) const {
    need(m.off, 8);
    d.type_class = p_[m.off] & 15;
    const int bits0 = p_[m.off + 1];
    d.elem = (uint32_t)get(m.off + 4, 4);
    d.big_endian = bits0 & 1;
    if (d.type_class == 0) d.is_signed = (bits0 & 8) != 0;
    else if (d.type_class == 1) d.is_signed = true;
    else fail("datatype class " + std::to_string(d.type_class) + " is neither fixed nor floating point");
    if (d.elem == 0 || d.elem > 16 || (d.type_class == 0 && d.elem > 8)) fail("unsupported element size");
    d.have_type = true;
  }

  void layout(const Message& m, Dataset& d) const {
    need(m.off, 2);
    const int version = p_[m.off];
    if (version == 1 || version == 2) {
      need(m.off, 8);
      const int ndim = p_[m.off + 1], cls = p_[m.off + 2];
      uint64_t o = m.off + 8;
      if (cls != 0) { d.addr = addr_field(o); o += so_; }
      need(o, 4ull * ndim + 8);
      std::vector<uint64_t> dims;
      for (int i = 0; i < ndim; ++i) dims.push_back(get(o + 4ull * i, 4));
      o += 4ull * ndim;
      if (cls == 2) {
        d.chunk = dims;                                                     // rank + 1 entries: the last is the element size
        d.layout_class = 2;
      } else if (cls == 1) {
        d.layout_class = 1;
        d.size = kUndef >> 1;                                               // v1/v2: the size is the dataspace's
      } else if (cls == 0) {
        d.size = get(o, 4);
        d.addr = o + 4;
        need(d.addr, d.size);
        d.layout_class = 0;
      } else fail("unknown data layout class");
      if (d.layout_class == 1 && d.addr != kUndef) d.addr += base_;
    } else if (version == 3 || version == 4) {
      const int cls = p_[m.off + 1];
      if (cls == 0) {
        d.size = get(m.off + 2, 2);
        d.addr = m.off + 4;
        need(d.addr, d.size);
        d.layout_class = 0;
      } else if (cls == 1) {
        d.addr = addr_field(m.off + 2);
        if (d.addr != kUndef) d.addr += base_;
        d.size = get(m.off + 2 + so_, sl_);
        d.layout_class = 1;
      } else if (cls == 2 && version == 3) {
        const int ndim = (int)get(m.off + 2, 1);
        d.addr = addr_field(m.off + 3);
        uint64_t o = m.off + 3 + so_;
        need(o, 4ull * ndim);
        d.chunk.clear();
        for (int i = 0; i < ndim; ++i) d.chunk.push_back(get(o + 4ull * i, 4));
        d.layout_class = 2;
      } else if (cls == 2) {
        const int flags = (int)get(m.off + 2, 1), ndim = (int)get(m.off + 3, 1), enc = (int)get(m.off + 4, 1);
        if (enc < 1 || enc > 8 || ndim < 1) fail("bad chunk dimension encoding");
        uint64_t o = m.off + 5;
        need(o, (uint64_t)ndim * enc + 1);
        d.chunk.clear();
        for (int i = 0; i + 1 < ndim; ++i) d.chunk.push_back(get(o + (uint64_t)i * enc, enc));   // last = element size
        o += (uint64_t)ndim * enc;
        d.v4_index = (int)get(o, 1);
        o += 1;
        if (d.v4_index == 1) {
          if (flags & 2) {
            d.single_filtered_size = get(o, sl_);
            d.single_filter_mask = (uint32_t)get(o + sl_, 4);
            o += sl_ + 4;
          }
        } else if (d.v4_index == 2) {
        } else {
          static const char* names[] = {"", "", "", "fixed array", "extensible array", "v2 B-tree"};
          fail(std::string("chunk index '") + (d.v4_index >= 3 && d.v4_index <= 5 ? names[d.v4_index] : "unknown") +
               "' (files written with libver='latest') is not supported");
        }
        d.addr = addr_field(o);
        d.layout_class = 3;
      } else fail("unknown data layout class (virtual datasets are not supported)");
    } else fail("unknown data layout version");
    d.have_layout = true;
  }

  void pipeline(const Message& m, Dataset& d) const {
    need(m.off, 2);
    const int version = p_[m.off], nf = p_[m.off + 1];
    uint64_t o = m.off + (version == 1 ? 8 : 2);
    const uint64_t end = m.off + m.size;
    if (version != 1 && version != 2) fail("unknown filter pipeline version");
    d.filters.clear();
    for (int f = 0; f < nf; ++f) {
      if (o >= end) fail("truncated filter pipeline message");
      Filter flt;
      flt.id = (int)get(o, 2);
      o += 2;
      uint64_t name_len = 0;
      if (version == 1 || flt.id >= 256) { name_len = get(o, 2); o += 2; }
      o += 2;                                                               // flags (optional filter bit)
      const uint64_t ncd = get(o, 2);
      o += 2;
      if (version == 1) name_len = (name_len + 7) & ~7ull;
      o += name_len;
      need(o, 4 * ncd);
      for (uint64_t i = 0; i < ncd; ++i) flt.client.push_back((uint32_t)get(o + 4 * i, 4));
      o += 4 * ncd;
      if (version == 1 && (ncd & 1)) o += 4;
      d.filters.push_back(flt);
    }
  }

  // One stored chunk -> elements [elem_off, elem_off + chunk) of the dataset (edge chunks are stored whole).
  void place_chunk(const Dataset& d, uint64_t abs_addr, uint64_t stored, uint32_t mask, uint64_t elem_off, std::vector<uint8_t>& out) {
    const uint64_t cbytes = d.chunk[0] * d.elem;
    need_abs(abs_addr, stored);
    // deflate expands at most ~1032 x: a larger chunk shape is a corrupt layout message, not something to allocate
    if (cbytes > (1ull << 32) || (has_vbz(d) ? cbytes > (1ull << 28) : cbytes > stored * 1100 + 65536))
      fail("chunk shape larger than its stored bytes can fill");
    std::vector<uint8_t> a(p_ + abs_addr, p_ + abs_addr + stored), b;
    for (int f = (int)d.filters.size() - 1; f >= 0; --f) {                  // undo the pipeline back to front
      if (mask & (1u << f)) continue;                                       // the writer skipped this filter for this chunk
      const Filter& flt = d.filters[f];
      if (flt.id == 1) {
        zlib_decompress(a.data(), a.size(), b, cbytes + 4);                  // + a fletcher32 checksum below it
        a.swap(b);
      } else if (flt.id == 2) {
        const uint64_t es = flt.client.empty() ? d.elem : flt.client[0];
        if (es > 1 && a.size() >= es) {
          const uint64_t ne = a.size() / es;
          b.assign(a.size(), 0);
          for (uint64_t j = 0; j < es; ++j)
            for (uint64_t i = 0; i < ne; ++i) b[i * es + j] = a[j * ne + i];
          for (uint64_t i = ne * es; i < a.size(); ++i) b[i] = a[i];
          a.swap(b);
        }
      } else if (flt.id == 3) {
        if (a.size() < 4) fail("chunk shorter than its fletcher32 checksum");
        a.resize(a.size() - 4);
      } else if (flt.id == 32020) {                                          // ONT's VBZ: zstd over streamvbyte (vbz.cu)
        ndhost::vbz_decompress(a.data(), a.size(), flt.client.data(), (int)flt.client.size(), b, cbytes + 4);
        a.swap(b);
      } else {
        fail("unsupported HDF5 filter id " + std::to_string(flt.id));
      }
    }
    if (a.size() != cbytes) fail("chunk decodes to " + std::to_string(a.size()) + " bytes, expected " + std::to_string(cbytes));
    if (elem_off >= d.dims[0]) return;
    const uint64_t ncopy = std::min<uint64_t>(d.chunk[0], d.dims[0] - elem_off) * d.elem;
    memcpy(out.data() + elem_off * d.elem, a.data(), ncopy);
  }

  void chunk_btree(const Dataset& d, uint64_t abs_node, std::vector<uint8_t>& out, int depth) {
    if (depth > 32) fail("chunk B-tree too deep");
    visit();
    need(abs_node, 8 + 2ull * so_);
    if (!sig(abs_node, "TREE") || p_[abs_node + 4] != 1) fail("chunk B-tree node signature missing");
    const int level = p_[abs_node + 5];
    const uint64_t used = get(abs_node + 6, 2);
    const uint64_t ksz = 8 + 8ull * d.chunk.size();
    uint64_t q = abs_node + 8 + 2ull * so_;
    need(q, used * (ksz + so_) + ksz);
    for (uint64_t i = 0; i < used; ++i) {
      const uint64_t stored = get(q, 4);
      const uint32_t mask = (uint32_t)get(q + 4, 4);
      const uint64_t off0 = get(q + 8, 8);
      const uint64_t child = get(q + ksz, so_);
      if (level == 0) place_chunk(d, base_ + child, stored, mask, off0, out);
      else chunk_btree(d, base_ + child, out, depth + 1);
      q += ksz + so_;
    }
  }
};

void put_error(char* err, int32_t errcap, const std::string& m) {
  if (!err || errcap <= 0) return;
  const size_t k = std::min<size_t>(m.size(), (size_t)errcap - 1);
  memcpy(err, m.data(), k);
  err[k] = 0;
}

// element i of a fixed-point dataset as a signed 64-bit value
int64_t fixed_value(const uint8_t* p, const Dataset& d) {
  uint64_t v = 0;
  for (uint32_t b = 0; b < d.elem && b < 8; ++b) {
    const uint8_t byte = d.big_endian ? p[d.elem - 1 - b] : p[b];
    v |= (uint64_t)byte << (8 * b);
  }
  if (d.is_signed && d.elem < 8 && (v >> (8 * d.elem - 1))) v |= ~0ull << (8 * d.elem);
  return (int64_t)v;
}

// the Signal dataset below `group` -> int16 samples; *count = samples, written when cap allows (cap < count: size query)
void read_signal_of(H5File& f, uint64_t group, const std::string& where, int16_t* out, int64_t cap, int64_t* count) {
  uint64_t sig = kUndef;
  for (const Link& l : f.links(group, false))
    if (l.name == "Signal") sig = l.addr;
  if (sig == kUndef) fail(where + " has no Signal dataset");
  const Dataset d = f.dataset(sig);
  if (d.type_class != 0) fail("Signal is not an integer dataset");
  const uint64_t n = f.count(d);
  *count = (int64_t)n;
  if ((uint64_t)cap < n) return;
  std::vector<uint8_t> bytes;
  f.read(d, bytes);
  for (uint64_t i = 0; i < n; ++i) {
    const int64_t v = fixed_value(bytes.data() + i * d.elem, d);
    if (v < -32768 || v > 32767) fail("Signal sample " + std::to_string(v) + " is outside the int16 DAC range");
    out[i] = (int16_t)v;
  }
}

}  // namespace

extern "C" {

int nd_h5_read_dataset(const uint8_t* file, int64_t nbytes, const char* path, uint8_t* out, int64_t cap, int64_t* info,
                       char* err, int32_t errcap) {
  if (!file || nbytes < 0 || !path || !info || cap < 0 || (cap > 0 && !out)) return ND_ERR_INVALID;
  try {
    H5File f(file, (uint64_t)nbytes);
    const Dataset d = f.dataset(f.resolve(path));
    std::vector<uint8_t> bytes;
    const uint64_t total = f.count(d) * d.elem;
    info[0] = d.type_class; info[1] = d.elem; info[2] = d.is_signed; info[3] = d.big_endian;
    info[4] = (int64_t)d.dims.size(); info[5] = (int64_t)total;
    info[6] = d.dims.size() > 0 ? (int64_t)d.dims[0] : 1;
    info[7] = d.dims.size() > 1 ? (int64_t)d.dims[1] : 1;
    if ((uint64_t)cap < total) return ND_OK;                                 // size query: info[5] bytes are needed
    f.read(d, bytes);
    if (total) memcpy(out, bytes.data(), total);
    return ND_OK;
  } catch (const H5Error& e) {
    put_error(err, errcap, e.msg);
    return ND_ERR_INVALID;
  } catch (const std::exception& e) {
    put_error(err, errcap, e.what());
    return ND_ERR_NOMEM;
  }
}

int nd_fast5_read_signal(const uint8_t* file, int64_t nbytes, int16_t* out, int64_t cap, int64_t* count, char* read_name,
                         int32_t name_cap, char* err, int32_t errcap) {
  if (!file || nbytes < 0 || !count || cap < 0 || (cap > 0 && !out)) return ND_ERR_INVALID;
  *count = 0;
  try {
    H5File f(file, (uint64_t)nbytes);
    const uint64_t reads = f.resolve("/Raw/Reads/");
    const std::vector<Link> members = f.links(reads);
    if (members.empty()) fail("/Raw/Reads has no members");                 // list(...)[0] raises IndexError
    put_error(read_name, name_cap, members[0].name);
    read_signal_of(f, members[0].addr, "/Raw/Reads/" + members[0].name, out, cap, count);
    return ND_OK;
  } catch (const H5Error& e) {
    put_error(err, errcap, e.msg);
    return ND_ERR_INVALID;
  } catch (const std::exception& e) {
    put_error(err, errcap, e.what());
    return ND_ERR_NOMEM;
  }
}

int nd_fast5_list_reads(const uint8_t* file, int64_t nbytes, char* names, int64_t names_cap, int64_t* names_bytes,
                        int32_t* n_reads, int32_t* layout, char* err, int32_t errcap) {
  if (!file || nbytes < 0 || !names_bytes || !n_reads || !layout || names_cap < 0 || (names_cap > 0 && !names)) return ND_ERR_INVALID;
  *names_bytes = 0; *n_reads = 0; *layout = 0;
  try {
    H5File f(file, (uint64_t)nbytes);
    std::vector<Link> members;
    bool single = false;
    for (const Link& l : f.links(f.root())) single = single || l.name == "Raw";
    if (single) {
      members = f.links(f.resolve("/Raw/Reads/"));
      *layout = 1;
    } else {
      for (const Link& l : f.links(f.root()))
        if (l.name.compare(0, 5, "read_") == 0) members.push_back(l);
      *layout = 2;
      if (members.empty()) fail("neither /Raw/Reads (single-read) nor /read_* groups (multi-read) in this file");
    }
    int64_t at = 0;
    for (const Link& l : members) {
      const int64_t len = (int64_t)l.name.size() + 1;
      if (at + len <= names_cap) memcpy(names + at, l.name.c_str(), (size_t)len);
      at += len;
    }
    *names_bytes = at;
    *n_reads = (int32_t)members.size();
    return ND_OK;
  } catch (const H5Error& e) {
    put_error(err, errcap, e.msg);
    return ND_ERR_INVALID;
  } catch (const std::exception& e) {
    put_error(err, errcap, e.what());
    return ND_ERR_NOMEM;
  }
}

int nd_h5_list_group(const uint8_t* file, int64_t nbytes, const char* path, char* names, int64_t names_cap, int64_t* names_bytes,
                     int32_t* n_members, char* err, int32_t errcap) {
  if (!file || nbytes < 0 || !path || !names_bytes || !n_members || names_cap < 0 || (names_cap > 0 && !names)) return ND_ERR_INVALID;
  *names_bytes = 0; *n_members = 0;
  try {
    H5File f(file, (uint64_t)nbytes);
    const std::vector<Link> members = f.links(f.resolve(path));             // name order, like h5py's iteration
    int64_t at = 0;
    for (const Link& l : members) {
      const int64_t len = (int64_t)l.name.size() + 1;
      if (at + len <= names_cap) memcpy(names + at, l.name.c_str(), (size_t)len);
      at += len;
    }
    *names_bytes = at;
    *n_members = (int32_t)members.size();
    return ND_OK;
  } catch (const H5Error& e) {
    put_error(err, errcap, e.msg);
    return ND_ERR_INVALID;
  } catch (const std::exception& e) {
    put_error(err, errcap, e.what());
    return ND_ERR_NOMEM;
  }
}

int nd_fast5_read_signal_of(const uint8_t* file, int64_t nbytes, const char* read_name, int16_t* out, int64_t cap,
                            int64_t* count, char* err, int32_t errcap) {
  if (!file || nbytes < 0 || !read_name || !count || cap < 0 || (cap > 0 && !out)) return ND_ERR_INVALID;
  *count = 0;
  try {
    H5File f(file, (uint64_t)nbytes);
    const std::string name(read_name);
    uint64_t raw = kUndef, member = kUndef;                                  // one walk over the root group (4000 members
    for (const Link& l : f.links(f.root(), false)) {                         // in a multi-read file)
      if (l.name == "Raw") raw = l.addr;
      if (l.name == name) member = l.addr;
    }
    if (raw != kUndef) read_signal_of(f, f.resolve("Reads/" + name, raw), "/Raw/Reads/" + name, out, cap, count);
    else if (member != kUndef) read_signal_of(f, f.resolve("Raw", member), "/" + name + "/Raw", out, cap, count);
    else fail("no object named '" + name + "' in the root group");
    return ND_OK;
  } catch (const H5Error& e) {
    put_error(err, errcap, e.msg);
    return ND_ERR_INVALID;
  } catch (const std::exception& e) {
    put_error(err, errcap, e.what());
    return ND_ERR_NOMEM;
  }
}

}  // extern "C"
