// Signal front end (reference: utils/labelop.py:219-233 + inputters/nano_dataset.py:49-58,81).
//
// Pass 1 — per-read statistics, exact:
//   center = np.median(x)                      two middle order statistics by 2-level radix select
//   scale  = statsmodels.robust.mad(x)         = median(|x - center| / 0.6744897501960817)
//          | np.std(x)  ('mean' mode)          exact integer sums, one correctly rounded sqrt
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

__global__ void __launch_bounds__(kStatThreads) stats_kernel(const int16_t* __restrict__ signal,
                                                             const int64_t* __restrict__ offsets, int mode,
                                                             double* __restrict__ center,
                                                             double* __restrict__ scale) {
  __shared__ unsigned hist[1024];
  __shared__ int sel[8];
  __shared__ long long s_sum[kStatThreads / 32], s_sq[kStatThreads / 32];
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
    long long ls = 0, lq = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const long long v = x[i]; ls += v; lq += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ls += __shfl_xor_sync(ND_FULL, ls, o);
      lq += __shfl_xor_sync(ND_FULL, lq, o);
    }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = ls; s_sq[threadIdx.x >> 5] = lq; }
    __syncthreads();
    s = 1.0;
    if (threadIdx.x == 0) {
      long long S = 0, Q = 0;
      for (int w = 0; w < kStatThreads / 32; ++w) { S += s_sum[w]; Q += s_sq[w]; }
      // population variance (np.std): (N*Q - S^2) / N^2 with an exact 128-bit numerator
      const __int128 num = (__int128)n * (__int128)Q - (__int128)S * (__int128)S;
      const double var = (double)num / ((double)n * (double)n);
      s = sqrt(var);
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

}  // namespace

cudaError_t frontend_stats(const int16_t* signal, const int64_t* offsets, int n_reads, int mode, double* center,
                           double* scale, cudaStream_t stream) {
  if (n_reads <= 0) return cudaSuccess;
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
