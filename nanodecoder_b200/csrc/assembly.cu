// Read assembly helpers (host code).  The reference stitches the base strings of overlapping
// chunks with difflib.SequenceMatcher(None, prev, cur).get_matching_blocks() and takes the LONGEST block
// (utils/labelop.py:320-352).  The longest block of get_matching_blocks() is what
// SequenceMatcher.find_longest_match(0, len(a), 0, len(b)) returns (earliest in a, then earliest in b, among the
// maximal ones; the sentinel (len(a), len(b), 0) when nothing matches), so that routine is restated here, including CPython difflib's "autojunk" rule: when len(b) >= 200,
// elements occurring more than len(b)//100 + 1 times are dropped from the b2j index (they still extend a match).
// Pure-Python difflib costs 1-3 ms per chunk pair, i.e. seconds per GPU batch; this is ~10 us.
#include <stdint.h>

#include <vector>

#include "../../include/nanodec.h"

namespace {

// scratch arrays reused over the chunk pairs of a read (the allocations cost as much as the search at 100 bases)
struct Workspace {
  std::vector<int> next, j2len, newj2len, touched, touched_new;
  void reserve(int nb) {
    if ((int)next.size() < nb + 1) { next.assign(nb + 1, -1); j2len.assign(nb + 2, 0); newj2len.assign(nb + 2, 0); }
  }
};

void longest_match(const char* a, int na, const char* b, int nb, int* out, Workspace& ws) {
  // b2j: positions of every byte value in b, ascending; popular bytes removed when len(b) >= 200
  ws.reserve(nb);
  int head[256], cnt[256];
  for (int c = 0; c < 256; ++c) { head[c] = -1; cnt[c] = 0; }
  std::vector<int>& next = ws.next;
  std::vector<int>& j2len = ws.j2len;            // all zero between calls (restored below)
  std::vector<int>& newj2len = ws.newj2len;
  std::vector<int>& touched = ws.touched;
  std::vector<int>& touched_new = ws.touched_new;
  touched.clear();
  for (int j = nb - 1; j >= 0; --j) {
    const unsigned char c = (unsigned char)b[j];
    next[j] = head[c];
    head[c] = j;
    ++cnt[c];
  }
  if (nb >= 200) {
    const int ntest = nb / 100 + 1;
    for (int c = 0; c < 256; ++c)
      if (cnt[c] > ntest) head[c] = -1;
  }
  int besti = 0, bestj = 0, bestsize = 0;                     // j2len / newj2len: index j + 1 (so j - 1 = -1 is slot 0)
  for (int i = 0; i < na; ++i) {
    touched_new.clear();
    for (int j = head[(unsigned char)a[i]]; j >= 0; j = next[j]) {
      const int k = j2len[j] + 1;                            // j2len.get(j - 1, 0) + 1
      newj2len[j + 1] = k;
      touched_new.push_back(j + 1);
      if (k > bestsize) { besti = i - k + 1; bestj = j - k + 1; bestsize = k; }
    }
    for (int t : touched) j2len[t] = 0;
    for (int t : touched_new) { j2len[t] = newj2len[t]; newj2len[t] = 0; }
    touched.swap(touched_new);
  }
  for (int t : touched) j2len[t] = 0;                          // leave the workspace clean
  // extension with non-junk elements (nothing is junk: isjunk is None), then the junk loops (no-ops here)
  while (besti > 0 && bestj > 0 && a[besti - 1] == b[bestj - 1]) { --besti; --bestj; ++bestsize; }
  while (besti + bestsize < na && bestj + bestsize < nb && a[besti + bestsize] == b[bestj + bestsize]) ++bestsize;
  // no common element at all: get_matching_blocks() is just its sentinel (len(a), len(b), 0), and that is what
  // max(..., key=size) hands to simple_assembly (disp = len(a) - len(b))
  if (bestsize == 0) { besti = na; bestj = nb; }
  out[0] = besti; out[1] = bestj; out[2] = bestsize;
}

}  // namespace

extern "C" {

int nd_longest_match(const char* a, int32_t na, const char* b, int32_t nb, int32_t* out3) {
  if ((!a && na > 0) || (!b && nb > 0) || !out3 || na < 0 || nb < 0) return ND_ERR_INVALID;
  int o[3];
  Workspace ws;
  longest_match(a, na, b, nb, o, ws);
  out3[0] = o[0]; out3[1] = o[1]; out3[2] = o[2];
  return ND_OK;
}

int nd_assembly_offsets(const char* text, const int64_t* offsets, int32_t n, int32_t* disp) {
  if (!text || !offsets || !disp || n < 0) return ND_ERR_INVALID;
  Workspace ws;
  for (int i = 0; i < n; ++i) {
    disp[i] = 0;
    if (i == 0) continue;
    int o[3];
    longest_match(text + offsets[i - 1], (int)(offsets[i] - offsets[i - 1]), text + offsets[i],
                  (int)(offsets[i + 1] - offsets[i]), o, ws);
    disp[i] = o[0] - o[1];
  }
  return ND_OK;
}

// simple_assembly (utils/labelop.py:320-352) + add_count (:311-318) for one read: displacement of every chunk from the
// longest matching block with its predecessor, then one vote per base into counts[code][column].  Reference quirks
// kept: the vote matrix starts 1000 columns wide and grows by 1000 at most ONCE per chunk (a chunk that still does not
// fit is numpy's IndexError), a negative start trims the head of the chunk, `length` only advances from the second
// chunk on.  counts: [n_codes][cap] zero-filled by the caller; lut: byte -> row or -1.
// *err: 0 ok | 1 column out of bounds (args = index, matrix width) | 2 byte without a row (args = byte, chunk).
int nd_simple_assembly(const char* text, const int64_t* offsets, int32_t n, const int8_t* lut, int32_t* counts,
                       int64_t cap, int64_t* length_out, int32_t* err, int64_t* err_args) {
  if (!text || !offsets || !lut || !counts || !length_out || !err || !err_args || n < 0 || cap < 1000)
    return ND_ERR_INVALID;
  int64_t pos = 0, length = 0, census_len = 1000;
  *err = 0;
  Workspace ws;
  for (int i = 0; i < n; ++i) {
    const char* seg = text + offsets[i];
    int64_t len = offsets[i + 1] - offsets[i];
    int64_t disp = 0;
    if (i > 0) {
      int o[3];
      longest_match(text + offsets[i - 1], (int)(offsets[i] - offsets[i - 1]), seg, (int)len, o, ws);
      disp = o[0] - o[1];
      if (disp + pos + len > census_len) census_len += 1000;
    }
    if (census_len > cap) return ND_ERR_INVALID;
    int64_t start = pos + disp, slen = len;
    if (start < 0) {                               // seg[-start:]
      const int64_t skip = -start < slen ? -start : slen;
      seg += skip;
      slen -= skip;
      start = 0;
    }
    for (int64_t k = 0; k < slen; ++k)
      if (lut[(unsigned char)seg[k]] < 0) {
        *err = 2; err_args[0] = (unsigned char)seg[k]; err_args[1] = i;
        return ND_OK;
      }
    if (start + slen > census_len) {
      *err = 1; err_args[0] = start + slen - 1; err_args[1] = census_len;
      return ND_OK;
    }
    for (int64_t k = 0; k < slen; ++k) ++counts[(int64_t)lut[(unsigned char)seg[k]] * cap + start + k];
    if (i > 0) {
      pos += disp;
      if (pos + len > length) length = pos + len;
    }
  }
  *length_out = length;
  return ND_OK;
}

// `.signal` text files (one read: whitespace separated integer DAC samples, the text form of the fast5 `Signal` dataset
// that utils/labelop.py:199-219 reads): straight to int16.  status: 0 ok (*count samples written), 1 = a token is not a
// plain integer (caller falls back to a float parser and its own checks), 2 = value outside int16, 3 = more than cap.
int nd_parse_signal_text(const char* text, int64_t nbytes, int16_t* out, int64_t cap, int64_t* count, int32_t* status) {
  if ((!text && nbytes > 0) || !out || !count || !status || nbytes < 0 || cap < 0) return ND_ERR_INVALID;
  int64_t n = 0, i = 0;
  *status = 0;
  while (i < nbytes) {
    const char c = text[i];
    if (c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v') { ++i; continue; }
    bool neg = false;
    if (c == '-' || c == '+') { neg = c == '-'; ++i; }
    if (i >= nbytes || text[i] < '0' || text[i] > '9') { *status = 1; *count = n; return ND_OK; }
    int64_t v = 0;
    while (i < nbytes && text[i] >= '0' && text[i] <= '9') {
      v = v * 10 + (text[i] - '0');
      if (v > 1000000) v = 1000000;                // saturate: range is checked below
      ++i;
    }
    if (i < nbytes) {
      const char e = text[i];
      if (!(e == ' ' || e == '\n' || e == '\t' || e == '\r' || e == '\f' || e == '\v')) {   // "1.0", "1e3", "12a"
        *status = 1; *count = n; return ND_OK;
      }
    }
    if (neg) v = -v;
    if (v < -32768 || v > 32767) { *status = 2; *count = n; return ND_OK; }
    if (n >= cap) { *status = 3; *count = n; return ND_OK; }
    out[n++] = (int16_t)v;
  }
  *count = n;
  return ND_OK;
}

}  // extern "C"
