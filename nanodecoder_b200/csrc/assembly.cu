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

void longest_match(const char* a, int na, const char* b, int nb, int* out) {
  // b2j: positions of every byte value in b, ascending; popular bytes removed when len(b) >= 200
  std::vector<int> head(256, -1), next(nb > 0 ? nb : 1, -1), cnt(256, 0);
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
  int besti = 0, bestj = 0, bestsize = 0;
  std::vector<int> j2len(nb + 1, 0), newj2len(nb + 1, 0);     // index j + 1 (so j - 1 = -1 is slot 0)
  std::vector<int> touched, touched_new;
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
  longest_match(a, na, b, nb, o);
  out3[0] = o[0]; out3[1] = o[1]; out3[2] = o[2];
  return ND_OK;
}

int nd_assembly_offsets(const char* text, const int64_t* offsets, int32_t n, int32_t* disp) {
  if (!text || !offsets || !disp || n < 0) return ND_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    disp[i] = 0;
    if (i == 0) continue;
    int o[3];
    longest_match(text + offsets[i - 1], (int)(offsets[i] - offsets[i - 1]), text + offsets[i],
                  (int)(offsets[i + 1] - offsets[i]), o);
    disp[i] = o[0] - o[1];
  }
  return ND_OK;
}

}  // extern "C"
