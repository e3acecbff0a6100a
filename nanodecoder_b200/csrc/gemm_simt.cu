// fp32 FFMA GEMM (bring-up / cross-check path for the tcgen05 kernel in gemm_tc.cu).
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread, fused prologue (LayerNorm / column affine)
// and epilogue (bias, q-scaling, ReLU, residual).
#include "gemm.cuh"

namespace nd {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_mean[BM], s_rstd[BM];

  const int tid = threadIdx.x;
  if (p.alive && *p.alive == 0) return;          // uniform over the grid (see GemmParams::alive)
  // 1-D grid, n-tile fastest: CTAs that share an A row block run in the same wave (L2 reuse)
  const int ntn = (p.N + BN - 1) / BN;
  const int m0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
  const int warp = tid >> 5, lane = tid & 31;

  if (p.prologue == PRO_LAYERNORM) {
    // one warp per row, two-pass mean / variance over K (population variance, like nn.LayerNorm)
    for (int r = warp; r < BM; r += NT / 32) {
      const int m = m0 + r;
      float mean = 0.f, rstd = 1.f;
      if (m < p.M) {
        const float* a = p.A + (int64_t)m * p.lda;
        float s = 0.f;
        for (int k = lane; k < p.K; k += 32) s += a[k];
        mean = warp_sum(s) / (float)p.K;
        float v = 0.f;
        for (int k = lane; k < p.K; k += 32) { float d = a[k] - mean; v += d * d; }
        rstd = 1.0f / sqrtf(warp_sum(v) / (float)p.K + p.eps);
      }
      if (lane == 0) { s_mean[r] = mean; s_rstd[r] = rstd; }
    }
    __syncthreads();
  }

  const int ty = tid >> 4, tx = tid & 15;          // 16x16 thread grid, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread (row = tid/4, k = (tid%4)*4..+3)
  const int lr = tid >> 2, lk = (tid & 3) * 4;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    {
      const int m = m0 + lr;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < p.M) {
        const float* a = p.A + (int64_t)m * p.lda + k0 + lk;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = k0 + lk + i;
          if (k < p.K) {
            float x = a[i];
            if (p.prologue == PRO_LAYERNORM) x = (x - s_mean[lr]) * s_rstd[lr] * p.pg[k] + p.pb[k];
            else if (p.prologue == PRO_AFFINE) x = x * p.pg[k] + p.pb[k];
            v[i] = x;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[lk + i][lr] = v[i];
    }
    {
      const int n = n0 + lr;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < p.N) {
        const float* w = p.W + (int64_t)n * p.ldw + k0 + lk;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + lk + i < p.K) v[i] = w[i] + (p.W_lo ? p.W_lo[(int64_t)n * p.ldw + k0 + lk + i] : 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[lk + i][lr] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (n < p.div_ncols) v = v / p.div_by;
      if (p.relu == 1) v = fmaxf(v, 0.f); else if (p.relu == 2) v = tanhf(v);
      if (p.residual) v += p.residual[(int64_t)m * p.ldr + n];
      if (p.relu == 3) v = fmaxf(v, 0.f);
      p.C[(int64_t)m * p.ldc + n] = v;
    }
  }
}

}  // namespace

cudaError_t gemm_simt(const GemmParams& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  const int64_t tiles = (int64_t)cdiv(p.N, BN) * cdiv(p.M, BM);
  gemm_simt_kernel<<<(unsigned)tiles, NT, 0, stream>>>(p);
  return cudaGetLastError();
}

void split_tf32_host(const float* w, float* hi, float* lo, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    union { float f; uint32_t u; } a, h, l;
    a.f = w[i];
    h.u = (a.u + 0x1000u) & 0xffffe000u;      // round to nearest tf32 (unbiased residual)
    hi[i] = h.f;
    l.f = a.f - h.f;                          // exact in fp32
    l.u = (l.u + 0x1000u) & 0xffffe000u;      // and the residual itself rounded to tf32
    lo[i] = l.f;
  }
}

}  // namespace nd
