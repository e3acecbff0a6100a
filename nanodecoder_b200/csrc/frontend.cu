// Signal front end (reference: utils/labelop.py:219-233 + inputters/nano_dataset.py:49-58,81).
//
// Pass 1 — per-read statistics, exact:
//   center = np.median(x)                      two middle order statistics by 2-level radix select
//   scale  = statsmodels.robust.mad(x)         = median(|x - center| / 0.6744897501960817)
//          | np.std(x)  ('mean' mode)          numpy's pairwise summation order in fp64
// The samples are int16, so 2*center and 2*|x - center| are integers: the selects run on integer
// keys and only the final two divisions happen in fp64 — bit-identical to numpy's float64 path.
// Pass 2 — out[c][j] = float( (double(x) - center) / scale ), zero padded chunks, one rounding.
#include "kernels.cuh"

namespace nd {

namespace {

constexpr int kStatThreads = 512;
constexpr double kMadC = 0.6744897501960817;      // scipy.stats.norm.ppf(0.75)

// Select the elements of rank k1 <= k2 (0-based) among keys key(i), i < n, keys < nb1*256.
// Two passes over the data: histogram of key>>8, then of key&255 inside the selected bins.
template <class KeyFn>
__device__ void select2(KeyFn key, int64_t n, int nb1, int64_t k1, int64_t k2, unsigned* hist /*[1024]*/,
                        int* sel /*[8]*/, unsigned& out1, unsigned& out2) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = tid; i < n; i += blockDim.x) atomicAdd(&hist[key(i) >> 8], 1u);
  __syncthreads();
  if (tid == 0) {
    int64_t cum = 0;
    int b1 = -1, b2 = -1;
    int64_t r1 = 0, r2 = 0;
    for (int b = 0; b < nb1; ++b) {
      const int64_t c = hist[b];
      if (b1 < 0 && k1 < cum + c) { b1 = b; r1 = k1 - cum; }
      if (b2 < 0 && k2 < cum + c) { b2 = b; r2 = k2 - cum; }
      cum += c;
    }
    sel[0] = b1; sel[1] = b2; sel[2] = (int)r1; sel[3] = (int)r2;
  }
  __syncthreads();
  const int b1 = sel[0], b2 = sel[1];
  for (int i = tid; i < 512; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = tid; i < n; i += blockDim.x) {
    const unsigned k = key(i);
    const int hb = (int)(k >> 8);
    if (hb == b1) atomicAdd(&hist[k & 255], 1u);
    if (hb == b2 && b2 != b1) atomicAdd(&hist[256 + (k & 255)], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    int64_t cum = 0;
    int v1 = 0, v2 = 0;
    bool f1 = false, f2 = false;
    const int r1 = sel[2], r2 = sel[3];
    for (int b = 0; b < 256; ++b) {
      const int64_t c = hist[b];
      if (!f1 && r1 < cum + c) { v1 = b; f1 = true; }
      if (b2 == b1 && !f2 && r2 < cum + c) { v2 = b; f2 = true; }
      cum += c;
    }
    if (b2 != b1) {
      cum = 0;
      for (int b = 0; b < 256; ++b) {
        const int64_t c = hist[256 + b];
        if (!f2 && r2 < cum + c) { v2 = b; f2 = true; }
        cum += c;
      }
    }
    sel[4] = (b1 << 8) | v1;
    sel[5] = (b2 << 8) | v2;
  }
  __syncthreads();
  out1 = (unsigned)sel[4];
  out2 = (unsigned)sel[5];
  __syncthreads();
}

// numpy's summation order for a contiguous float64 array (numpy/core/src/umath/loops_utils.h.src, pairwise_sum): blocks
// of <= 128 elements with 8 interleaved accumulators, halves split at multiples of 8.  Run by ONE thread so that np.std
// is reproduced to the bit (checked against numpy on the host for n = 1 ... 200 000).
template <class F>
__device__ double pairwise_sum(F f, int64_t lo, int64_t n) {
  if (n < 8) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res += f(lo + i);
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += f(lo + i + j);
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += f(lo + i);
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return pairwise_sum(f, lo, n2) + pairwise_sum(f, lo + n2, n - n2);
}

// todo != nullptr: only the reads the histogram kernel left (todo[r] != 0)
__global__ void __launch_bounds__(kStatThreads) stats_kernel(const int16_t* __restrict__ signal,
                                                             const int64_t* __restrict__ offsets, int mode,
                                                             double* __restrict__ center,
                                                             double* __restrict__ scale,
                                                             const int* __restrict__ todo) {
  __shared__ unsigned hist[1024];
  __shared__ int sel[8];
  const int r = blockIdx.x;
  if (todo && !todo[r]) return;
  const int16_t* x = signal + offsets[r];
  const int64_t n = offsets[r + 1] - offsets[r];
  if (n <= 0) {
    if (threadIdx.x == 0) { center[r] = 0.0; scale[r] = 1.0; }
    return;
  }
  if (mode == 2) {                                  // normalization 'None'
    if (threadIdx.x == 0) { center[r] = 0.0; scale[r] = 1.0; }
    return;
  }
  const int64_t k1 = (n - 1) / 2, k2 = n / 2;
  unsigned u1, u2;
  select2([&](int64_t i) { return (unsigned)((int)x[i] + 32768); }, n, 256, k1, k2, hist, sel, u1, u2);
  const int med2 = ((int)u1 - 32768) + ((int)u2 - 32768);          // 2 * median, exact
  const double c = 0.5 * (double)med2;
  double s;
  if (mode == 0) {
    unsigned d1, d2;
    select2([&](int64_t i) { const int v = 2 * (int)x[i] - med2; return (unsigned)(v < 0 ? -v : v); }, n, 512, k1,
            k2, hist, sel, d1, d2);
    const double m1 = (0.5 * (double)d1) / kMadC;
    const double m2 = (0.5 * (double)d2) / kMadC;
    s = (k1 == k2) ? m1 : (m1 + m2) / 2.0;
  } else {
    // np.std of the float64 array the reference builds from the text file (utils/labelop.py:216-221): numpy's own
    // summation order, so the scale is the same double (an exact-integer variance differs from it in the last ulps)
    s = 1.0;
    if (threadIdx.x == 0) {
      const double mean = pairwise_sum([&](int64_t i) { return (double)x[i]; }, 0, n) / (double)n;
      const double ss = pairwise_sum([&](int64_t i) { const double t = (double)x[i] - mean; return t * t; }, 0, n);
      s = sqrt(ss / (double)n);
    }
  }
  if (threadIdx.x == 0) { center[r] = c; scale[r] = s; }
}


// ---------------------------------------------------------------------------------------------
// Fast path of the statistics: ONE histogram per read.  Raw DAC samples span a few thousand distinct values, so the
// whole read is histogrammed into shared memory (kHistBins bins) with 128-bit loads in one pass, and both the median
// and the MAD come out of that histogram: the MAD's order statistics of |2x - 2*median| are found by walking outwards
// from the median over the bins.  One vectorised pass instead of four scalar ones; reads with a wider value range
// take stats_kernel.
constexpr int kHistBins = 16384;

// f(v) for every sample of x[0, n): 8 samples per 128-bit load on the 16-byte aligned body, scalars at both ends
template <class F>
__device__ __forceinline__ void for_each_sample(const int16_t* x, int64_t n, F f) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(x) & 15)) & 15) >> 1);
  if (head > n) head = n;
  if (tid < head) f((int)x[tid]);
  const int64_t nvec = (n - head) >> 3;
  const int4* xv = reinterpret_cast<const int4*>(x + head);
  for (int64_t i = tid; i < nvec; i += nt) {
    const int4 q = __ldg(xv + i);
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f((int)(short)(w[k] & 0xffff));
      f(w[k] >> 16);
    }
  }
  const int64_t done = head + (nvec << 3);
  if (tid < n - done) f((int)x[done + tid]);
}

__global__ void __launch_bounds__(kStatThreads) stats_hist_kernel(const int16_t* __restrict__ signal,
                                                                  const int64_t* __restrict__ offsets, int mode,
                                                                  double* __restrict__ center,
                                                                  double* __restrict__ scale, int* __restrict__ todo) {
  extern __shared__ unsigned hist_s[];                // [kHistBins]
  __shared__ int s_red[2][kStatThreads / 32];
  __shared__ unsigned long long s_scan[kStatThreads];
  __shared__ int s_val[2];
  const int r = blockIdx.x, tid = threadIdx.x;
  const int16_t* x = signal + offsets[r];
  const int64_t n = offsets[r + 1] - offsets[r];
  if (tid == 0) todo[r] = 0;
  if (n <= 0 || mode == 2) {
    if (tid == 0) { center[r] = 0.0; scale[r] = 1.0; }
    return;
  }
  // ---- optimistic single pass: a window of kHistBins values centred on the first sample holds every sample of a real
  // read (a few thousand DAC levels); any sample outside sets the overflow flag and the read is redone with its true
  // minimum as the origin (or left to the radix-select kernel when its range exceeds the histogram)
  __shared__ int s_over;
  int lo = (int)x[0] - kHistBins / 2, hi = lo + kHistBins - 1;
  if (tid == 0) s_over = 0;
  for (int i = tid; i < kHistBins; i += kStatThreads) hist_s[i] = 0u;
  __syncthreads();
  {
    bool over = false;
    for_each_sample(x, n, [&](int v) {
      const unsigned b = (unsigned)(v - lo);
      if (b < (unsigned)kHistBins) atomicAdd(&hist_s[b], 1u); else over = true;
    });
    if (over) s_over = 1;
  }
  __syncthreads();
  if (s_over) {
    lo = 32767; hi = -32768;
    for_each_sample(x, n, [&](int v) { lo = min(lo, v); hi = max(hi, v); });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(ND_FULL, lo, o));
      hi = max(hi, __shfl_xor_sync(ND_FULL, hi, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = lo; s_red[1][tid >> 5] = hi; }
    __syncthreads();
    lo = s_red[0][0]; hi = s_red[1][0];
#pragma unroll
    for (int w = 1; w < kStatThreads / 32; ++w) { lo = min(lo, s_red[0][w]); hi = max(hi, s_red[1][w]); }
    if (hi - lo + 1 > kHistBins) {                    // wide value range: the radix-select kernel takes this read
      if (tid == 0) todo[r] = 1;
      return;
    }
    for (int i = tid; i < kHistBins; i += kStatThreads) hist_s[i] = 0u;
    __syncthreads();
    for_each_sample(x, n, [&](int v) { atomicAdd(&hist_s[v - lo], 1u); });
    __syncthreads();
  }
  const int range = min(hi, 32767) - lo + 1 > kHistBins ? kHistBins : (s_over ? hi - lo + 1 : kHistBins);
  // ---- the two middle order statistics: per-thread bin groups, block scan of the group counts, local walk
  const int64_t k1 = (n - 1) / 2, k2 = n / 2;
  const int bpt = (range + kStatThreads - 1) / kStatThreads;
  const int b0 = tid * bpt, b1 = min(range, b0 + bpt);
  unsigned long long mine = 0;
  for (int b = b0; b < b1; ++b) mine += hist_s[b];
  s_scan[tid] = mine;
  __syncthreads();
  for (int o = 1; o < kStatThreads; o <<= 1) {        // inclusive Hillis-Steele scan (512 values)
    const unsigned long long add = tid >= o ? s_scan[tid - o] : 0ull;
    __syncthreads();
    s_scan[tid] += add;
    __syncthreads();
  }
  const unsigned long long before = s_scan[tid] - mine;
  for (int q = 0; q < 2; ++q) {
    const unsigned long long k = (unsigned long long)(q ? k2 : k1);
    if (k >= before && k < before + mine) {
      unsigned long long cum = before;
      for (int b = b0; b < b1; ++b) {
        cum += hist_s[b];
        if (k < cum) { s_val[q] = lo + b; break; }
      }
    }
  }
  __syncthreads();
  const int med2 = s_val[0] + s_val[1];               // 2 * median, exact
  if (tid != 0) return;
  const double c = 0.5 * (double)med2;
  double s = 1.0;
  if (mode == 0) {
    // order statistics k1, k2 of d = |2x - med2|: merge the bins below the median (descending x) and above it
    // (ascending x) by increasing d
    int a = (int)floor(0.5 * (double)med2);           // largest x with 2x <= med2
    int b = a + 1;                                    // smallest x with 2x > med2
    long long cum = 0;
    int d1 = 0, d2 = 0;
    bool f1 = false, f2 = false;
    while (!(f1 && f2)) {
      const int da = a >= lo ? med2 - 2 * a : 0x7fffffff;
      const int db = b <= hi ? 2 * b - med2 : 0x7fffffff;
      int d;
      long long cnt;
      if (da <= db) { d = da; cnt = hist_s[a - lo]; --a; if (da == db) { cnt += hist_s[b - lo]; ++b; } }
      else { d = db; cnt = hist_s[b - lo]; ++b; }
      if (!f1 && k1 < cum + cnt) { d1 = d; f1 = true; }
      if (!f2 && k2 < cum + cnt) { d2 = d; f2 = true; }
      cum += cnt;
    }
    const double m1 = (0.5 * (double)d1) / kMadC;
    const double m2 = (0.5 * (double)d2) / kMadC;
    s = (k1 == k2) ? m1 : (m1 + m2) / 2.0;
  } else {
    const double mean = pairwise_sum([&](int64_t i) { return (double)x[i]; }, 0, n) / (double)n;
    const double ss = pairwise_sum([&](int64_t i) { const double t = (double)x[i] - mean; return t * t; }, 0, n);
    s = sqrt(ss / (double)n);
  }
  center[r] = c;
  scale[r] = s;
}

// Chunk gather, one warp per chunk and 8 chunks per CTA: the chunk's samples are staged in shared memory with 128-bit
// loads of the aligned segments that cover them (reads start at arbitrary sample offsets), normalised in fp64 and stored
// as 128-bit vectors.  (x - center) * (1 / scale) replaces the fp64 division (20+ instructions per sample, the kernel's former
// bound) wherever that is provably the same float: the true quotient and its fp64 rounding both lie within 2^-52
// relative of the product, so if the product's neighbours at +-2^-50 round to the same float, so does the reference's
// float(float64 quotient); otherwise (about one sample in 10^8) the division is done.
constexpr int kChunkWarps = 8;                       // chunks per CTA: one warp each (a CTA per 1 KB chunk is launch bound)

__global__ void __launch_bounds__(kChunkWarps * 32) chunks_vec_kernel(const int16_t* __restrict__ signal,
                                                                      const int64_t* __restrict__ offsets,
                                                                      const double* __restrict__ center,
                                                                      const double* __restrict__ scale,
                                                                      const int32_t* __restrict__ chunk_read,
                                                                      const int64_t* __restrict__ chunk_start,
                                                                      int n_chunks, int chunk_len,
                                                                      float* __restrict__ out,
                                                                      int64_t* __restrict__ out_len) {
  extern __shared__ __align__(16) int16_t stage_all[]; // [kChunkWarps][chunk_len + 16 rounded up to 8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kChunkWarps + warp;
  if (c >= n_chunks) return;
  int16_t* stage = stage_all + (size_t)warp * ((chunk_len + 23) & ~7);       // 16-byte aligned pitch
  const int r = chunk_read[c];
  const int64_t start = chunk_start[c];
  const int64_t n = offsets[r + 1] - offsets[r];
  int64_t len64 = n - start;
  if (len64 > chunk_len) len64 = chunk_len;
  if (len64 < 0) len64 = 0;
  const int len = (int)len64;
  if (lane == 0) out_len[c] = len;
  const int16_t* x = signal + offsets[r] + start;
  const int mis = (int)((reinterpret_cast<uintptr_t>(x) & 15) >> 1);         // samples before x in its 16-byte segment
  const int4* seg = reinterpret_cast<const int4*>(x - mis);
  const int nseg = (mis + len + 7) >> 3;
  // the last segment may reach past the read (and, for the last read, past the buffer): guard it with scalar loads
  const int16_t* buf_end = signal + offsets[r + 1];
  for (int i = lane; i < nseg; i += 32) {
    const int16_t* p = reinterpret_cast<const int16_t*>(seg + i);
    if (p + 8 <= buf_end) {                           // (p >= signal: the buffer base is 16-byte aligned)
      reinterpret_cast<int4*>(stage)[i] = __ldg(seg + i);
    } else {
      for (int k = 0; k < 8; ++k) stage[i * 8 + k] = (p + k < buf_end) ? p[k] : (int16_t)0;
    }
  }
  __syncwarp();
  const double ctr = center[r], scl = scale[r];
  const double rcp = 1.0 / scl;
  float* o = out + (int64_t)c * chunk_len;
  for (int j0 = lane * 4; j0 < chunk_len; j0 += 128) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i;
      float f = 0.f;
      if (j < len) {
        const double num = (double)stage[mis + j] - ctr;
        const double q = num * rcp;
        f = (float)q;
        if ((float)(q * (1.0 - 0x1p-50)) != (float)(q * (1.0 + 0x1p-50))) f = (float)(num / scl);
      }
      v[i] = f;
    }
    *reinterpret_cast<float4*>(o + j0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// general form (any chunk_len / alignment): 4 samples per thread, 2-byte loads, fp64 division
__global__ void __launch_bounds__(128) chunks_kernel(const int16_t* __restrict__ signal,
                                                     const int64_t* __restrict__ offsets,
                                                     const double* __restrict__ center,
                                                     const double* __restrict__ scale,
                                                     const int32_t* __restrict__ chunk_read,
                                                     const int64_t* __restrict__ chunk_start, int chunk_len,
                                                     float* __restrict__ out, int64_t* __restrict__ out_len) {
  const int c = blockIdx.x;
  const int r = chunk_read[c];
  const int64_t start = chunk_start[c];
  const int64_t n = offsets[r + 1] - offsets[r];
  int64_t len = n - start;
  if (len > chunk_len) len = chunk_len;
  if (len < 0) len = 0;
  if (threadIdx.x == 0) out_len[c] = len;
  const int16_t* x = signal + offsets[r] + start;
  const double ctr = center[r], scl = scale[r];
  float* o = out + (int64_t)c * chunk_len;
  const bool vec = (chunk_len & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int j0 = threadIdx.x * 4; j0 < chunk_len; j0 += blockDim.x * 4) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i;
      v[i] = (j < len) ? (float)(((double)x[j] - ctr) / scl) : 0.f;     // one rounding fp64 -> fp32
    }
    if (vec) {
      *reinterpret_cast<float4*>(o + j0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int i = 0; i < 4 && j0 + i < chunk_len; ++i) o[j0 + i] = v[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Float-valued reads (`.signal` text files holding non-integer samples: the reference parses every token with float(),
// utils/labelop.py:216-217, so it takes them): the same statistics on fp64 samples.
//   median / MAD: order statistics by an 8-pass byte-wise radix select over order-preserving 64-bit keys;
//   std:          numpy's own arithmetic -- pairwise summation with 8 interleaved accumulators per block of 128
//                 (numpy/core/src/umath/loops_utils.h: pairwise_sum), mean = sum / n, sum((x - mean)^2) / n, sqrt -- by
//                 ONE thread in numpy's order, so the scale is the double np.std returns, not a value one ulp beside it.
__device__ __forceinline__ unsigned long long f64_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_key(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// element of rank k (0-based) among key(i), i < n: 8 passes, most significant byte first
template <class KeyFn>
__device__ unsigned long long select_u64(KeyFn key, int64_t n, int64_t k, unsigned* hist /*[256]*/,
                                         unsigned long long* state /*[2]*/) {
  const int tid = threadIdx.x;
  unsigned long long prefix = 0, mask = 0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += blockDim.x) {
      const unsigned long long kk = key(i);
      if ((kk & mask) == prefix) atomicAdd(&hist[(kk >> shift) & 255], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int64_t cum = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (k < cum + (int64_t)hist[b]) break;
        cum += hist[b];
      }
      state[0] = prefix | ((unsigned long long)b << shift);
      state[1] = (unsigned long long)(k - cum);
    }
    __syncthreads();
    prefix = state[0];
    k = (int64_t)state[1];
    mask |= 0xffull << shift;
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(kStatThreads) stats_f64_kernel(const double* __restrict__ signal,
                                                                 const int64_t* __restrict__ offsets, int mode,
                                                                 double* __restrict__ center,
                                                                 double* __restrict__ scale) {
  __shared__ unsigned hist[256];
  __shared__ unsigned long long state[2];
  const int r = blockIdx.x;
  const double* x = signal + offsets[r];
  const int64_t n = offsets[r + 1] - offsets[r];
  if (n <= 0 || mode == 2) {
    if (threadIdx.x == 0) { center[r] = 0.0; scale[r] = 1.0; }
    return;
  }
  const int64_t k1 = (n - 1) / 2, k2 = n / 2;
  const double v1 = f64_from_key(select_u64([&](int64_t i) { return f64_key(x[i]); }, n, k1, hist, state));
  const double v2 = k2 == k1 ? v1 : f64_from_key(select_u64([&](int64_t i) { return f64_key(x[i]); }, n, k2, hist, state));
  const double c = k2 == k1 ? v1 : (v1 + v2) / 2.0;                      // np.median: mean of the two middle elements
  double s = 1.0;
  if (mode == 0) {
    auto dkey = [&](int64_t i) { return (unsigned long long)__double_as_longlong(fabs(x[i] - c)); };   // >= 0: bits order
    const double d1 = __longlong_as_double((long long)select_u64(dkey, n, k1, hist, state));
    const double d2 = k2 == k1 ? d1 : __longlong_as_double((long long)select_u64(dkey, n, k2, hist, state));
    const double m1 = d1 / kMadC, m2 = d2 / kMadC;
    s = (k1 == k2) ? m1 : (m1 + m2) / 2.0;
  } else if (threadIdx.x == 0) {
    const double mean = pairwise_sum([&](int64_t i) { return x[i]; }, 0, n) / (double)n;
    const double ss = pairwise_sum([&](int64_t i) { const double t = x[i] - mean; return t * t; }, 0, n);
    s = sqrt(ss / (double)n);
  }
  if (threadIdx.x == 0) { center[r] = c; scale[r] = s; }
}

__global__ void __launch_bounds__(128) chunks_f64_kernel(const double* __restrict__ signal,
                                                         const int64_t* __restrict__ offsets,
                                                         const double* __restrict__ center,
                                                         const double* __restrict__ scale,
                                                         const int32_t* __restrict__ chunk_read,
                                                         const int64_t* __restrict__ chunk_start, int chunk_len,
                                                         float* __restrict__ out, int64_t* __restrict__ out_len) {
  const int c = blockIdx.x;
  const int r = chunk_read[c];
  const int64_t start = chunk_start[c];
  const int64_t n = offsets[r + 1] - offsets[r];
  int64_t len = n - start;
  if (len > chunk_len) len = chunk_len;
  if (len < 0) len = 0;
  if (threadIdx.x == 0) out_len[c] = len;
  const double* x = signal + offsets[r] + start;
  const double ctr = center[r], scl = scale[r];
  float* o = out + (int64_t)c * chunk_len;
  for (int j = threadIdx.x; j < chunk_len; j += blockDim.x)
    o[j] = (j < len) ? (float)((x[j] - ctr) / scl) : 0.f;                  // one rounding fp64 -> fp32
}

}  // namespace

static void ensure_stack() {                        // pairwise_sum recurses (depth ~ log2(n / 128))
  static PerDeviceFlag stack_set;
  bool& set = stack_set.cur();
  if (!set) {
    size_t cur = 0;
    cudaDeviceGetLimit(&cur, cudaLimitStackSize);
    if (cur < 4096) cudaDeviceSetLimit(cudaLimitStackSize, 4096);
    set = true;
  }
}

cudaError_t frontend_stats_f64(const double* signal, const int64_t* offsets, int n_reads, int mode, double* center,
                               double* scale, cudaStream_t stream) {
  if (n_reads <= 0) return cudaSuccess;
  ensure_stack();
  stats_f64_kernel<<<n_reads, kStatThreads, 0, stream>>>(signal, offsets, mode, center, scale);
  return cudaGetLastError();
}

cudaError_t frontend_chunks_f64(const double* signal, const int64_t* offsets, const double* center, const double* scale,
                                const int32_t* chunk_read, const int64_t* chunk_start, int n_chunks, int chunk_len,
                                float* out, int64_t* out_len, cudaStream_t stream) {
  if (n_chunks <= 0) return cudaSuccess;
  chunks_f64_kernel<<<n_chunks, 128, 0, stream>>>(signal, offsets, center, scale, chunk_read, chunk_start, chunk_len,
                                                 out, out_len);
  return cudaGetLastError();
}

int g_frontend_fast = 1;
void frontend_set_fast(int on) { g_frontend_fast = on; }

cudaError_t frontend_stats(const int16_t* signal, const int64_t* offsets, int n_reads, int mode, double* center,
                           double* scale, int* todo, cudaStream_t stream) {
  if (n_reads <= 0) return cudaSuccess;
  ensure_stack();
  if (todo && g_frontend_fast) {
    static PerDeviceFlag attr_set;
    bool& set = attr_set.cur();
    if (!set) {
      cudaError_t err = cudaFuncSetAttribute(stats_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kHistBins * (int)sizeof(unsigned));
      if (err != cudaSuccess) return err;
      set = true;
    }
    stats_hist_kernel<<<n_reads, kStatThreads, kHistBins * sizeof(unsigned), stream>>>(signal, offsets, mode, center, scale,
                                                                                      todo);
    stats_kernel<<<n_reads, kStatThreads, 0, stream>>>(signal, offsets, mode, center, scale, todo);   // wide-range reads
    return cudaGetLastError();
  }
  stats_kernel<<<n_reads, kStatThreads, 0, stream>>>(signal, offsets, mode, center, scale, nullptr);
  return cudaGetLastError();
}

cudaError_t frontend_chunks(const int16_t* signal, const int64_t* offsets, const double* center, const double* scale,
                            const int32_t* chunk_read, const int64_t* chunk_start, int n_chunks, int chunk_len,
                            float* out, int64_t* out_len, cudaStream_t stream) {
  if (n_chunks <= 0) return cudaSuccess;
  if (g_frontend_fast && (chunk_len & 3) == 0 && chunk_len <= 2048 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(signal) & 15) == 0) {
    const size_t smem = (size_t)kChunkWarps * ((chunk_len + 23) & ~7) * sizeof(int16_t);
    chunks_vec_kernel<<<cdiv(n_chunks, kChunkWarps), kChunkWarps * 32, smem, stream>>>(
        signal, offsets, center, scale, chunk_read, chunk_start, n_chunks, chunk_len, out, out_len);
    return cudaGetLastError();
  }
  chunks_kernel<<<n_chunks, 128, 0, stream>>>(signal, offsets, center, scale, chunk_read, chunk_start, chunk_len, out,
                                             out_len);
  return cudaGetLastError();
}

}  // namespace nd
