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

__global__ void __launch_bounds__(kStatThreads) stats_kernel(const int16_t* __restrict__ signal,
                                                             const int64_t* __restrict__ offsets, int mode,
                                                             double* __restrict__ center,
                                                             double* __restrict__ scale) {
  __shared__ unsigned hist[1024];
  __shared__ int sel[8];
  const int r = blockIdx.x;
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

// 4 samples per thread: coalesced 2-byte loads, one 128-bit store
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

cudaError_t frontend_stats(const int16_t* signal, const int64_t* offsets, int n_reads, int mode, double* center,
                           double* scale, cudaStream_t stream) {
  if (n_reads <= 0) return cudaSuccess;
  ensure_stack();
  stats_kernel<<<n_reads, kStatThreads, 0, stream>>>(signal, offsets, mode, center, scale);
  return cudaGetLastError();
}

cudaError_t frontend_chunks(const int16_t* signal, const int64_t* offsets, const double* center, const double* scale,
                            const int32_t* chunk_read, const int64_t* chunk_start, int n_chunks, int chunk_len,
                            float* out, int64_t* out_len, cudaStream_t stream) {
  if (n_chunks <= 0) return cudaSuccess;
  chunks_kernel<<<n_chunks, 128, 0, stream>>>(signal, offsets, center, scale, chunk_read, chunk_start, chunk_len, out,
                                             out_len);
  return cudaGetLastError();
}

}  // namespace nd
