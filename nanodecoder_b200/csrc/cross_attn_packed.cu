// Decode-step cross attention over FIXED-POINT memory keys / values (kernels.cuh: KV_Q24 / KV_Q16) and the packer.
//
// Reference rows: onmt/modules/multi_headed_attn.py:142-190 through decoder/transformer.py:87-91 (context attention
// of the Transformer decoder).  The fp32 kernel (attention.cu, cross_attn_kernel) already runs at 98 % of the measured
// HBM bandwidth and is 55 % of the translate step, so the only way down is fewer bytes: the projected keys / values
// of a chunk are written ONCE per batch and then re-read dec_layers * max_length (300 ... 600) times, so they are
// stored as 24-bit fixed point with one power-of-two step per row part (3 bytes per element, the same absolute
// rounding error fp32 has on the part's largest element) or, as the reduced-precision mode, as 16-bit fixed point.
//
// Kernel structure = cross_attn_kernel: one CTA per chunk, lane l owns columns [l*VPL, (l+1)*VPL) of every row, the
// 32/H lanes of a head combine their partial dot products with xor-shuffles; phases scores / softmax / context.
// The integers are rebuilt with one PRMT per element (int16 pair register + uint8 quad register -> sign-extended int32),
// converted with I2F and used UNSCALED: the row's step multiplies the finished score (keys) or the probability
// (values) once per row instead of once per element.
#include <float.h>

#include "kernels.cuh"
#include "kv_fixed.cuh"

namespace nd {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- loads: NB contiguous bytes per lane (32 / 16 / 8 / 4 / 2), streaming (read once per step)
template <int NB>
__device__ __forceinline__ void load_bytes(const uint8_t* p, uint32_t* r) {
  if constexpr (NB == 32) {
    float f[8];
    ldg_stream8(reinterpret_cast<const float*>(p), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(f[i]);
  } else if constexpr (NB == 16) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
  } else if constexpr (NB == 8) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(p));
  } else if constexpr (NB == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r[0]) : "l"(p));
  } else {
    uint16_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    r[0] = v;
  }
}

template <int VPL, int FMT>
struct RowRegs {
  static constexpr int NHI = VPL >= 2 ? VPL / 2 : 1;
  static constexpr int NLO = VPL >= 4 ? VPL / 4 : 1;
  uint32_t hi[NHI];
  uint32_t lo[fmt_has_lo(FMT) ? NLO : 1];
  __device__ __forceinline__ void load(const uint8_t* hi_p, const uint8_t* lo_p) {
    load_bytes<2 * VPL>(hi_p, hi);
    if constexpr (fmt_has_lo(FMT)) load_bytes<VPL>(lo_p, lo);
  }
  __device__ __forceinline__ float get(int j) const { return unpack_elem<FMT>(hi, lo, j); }
};

// =============================================================================================
template <int VPL, int NQMAX, int FMT>
__global__ void __launch_bounds__(kThreads) cross_attn_packed_kernel(CrossAttnParams p) {
  constexpr int R = VPL >= 16 ? 4 : 8;             // rows in flight per warp iteration (6-12 registers per row)
  extern __shared__ __align__(16) float smem_f[];
  const int chunk = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (p.retired && p.retired[chunk]) return;
  const int d = 32 * VPL, T = p.T, H = p.H, NQ = p.NQ;
  const int TS = T + 1;
  const int LPH = 32 / H;
  float* q_s = smem_f;                            // [NQ][d]
  float* sc = q_s + NQ * d;                       // [NQ*H][TS]   (later reused as red[warps][NQ*d])
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < NQ * d; i += kThreads) {
    const int qi = i / d, c = i - qi * d;
    q_s[i] = p.q[((int64_t)chunk * NQ + qi) * p.q_ld + c] / p.q_div;
  }
  __syncthreads();

  const int64_t row0 = (int64_t)chunk * T;
  const uint8_t* hiK = reinterpret_cast<const uint8_t*>(p.kv_hi) + row0 * (4 * d) + lane * (2 * VPL);
  const uint8_t* loK = reinterpret_cast<const uint8_t*>(p.kv_lo) + row0 * (2 * d) + lane * VPL;
  const float* stp = p.kv_scale + row0 * 2;
  const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
  const int head = lane / LPH;

  // ---------------- phase 1: scores
  for (int t0 = warp * R; t0 < T; t0 += kWarps * R) {
    RowRegs<VPL, FMT> kv[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r < T) kv[r].load(hiK + (int64_t)(t0 + r) * (4 * d), loK + (int64_t)(t0 + r) * (2 * d));
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < T) {                                 // warp-uniform
        const bool masked = srow && (srow[t] == p.mask_value);
        const float step = fmt_scaled(FMT) ? stp[2 * t] : 1.0f;
        float kf[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) kf[i] = kv[r].get(i);
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float* qq = q_s + qi * d + lane * VPL;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) s = fmaf(qq[i], kf[i], s);
            for (int o = LPH >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(ND_FULL, s, o);
            if ((lane % LPH) == 0) sc[(qi * H + head) * TS + t] = masked ? -1e18f : s * step;
          }
        }
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: softmax rows (torch.softmax: exp(x - max) / sum); the value step of row t is folded
  // into the probability AFTER the normalisation (and after the optional attention output)
  for (int row = warp; row < NQ * H; row += kWarps) {
    float* s = sc + row * TS;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s[t] - m); s[t] = e; sum += e; }
    sum = warp_sum(sum);
    float* a = (p.attn && (row % H) == 0) ? p.attn + ((int64_t)chunk * NQ + row / H) * T : nullptr;
    for (int t = lane; t < T; t += 32) {
      const float pr = s[t] / sum;
      if (a) a[t] = pr;
      s[t] = fmt_scaled(FMT) ? pr * stp[2 * t + 1] : pr;
    }
  }
  __syncthreads();

  // ---------------- phase 3: context
  float acc[NQMAX][VPL];
#pragma unroll
  for (int qi = 0; qi < NQMAX; ++qi)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[qi][i] = 0.f;
  const uint8_t* hiV = hiK + 2 * d;                // V = columns [d, 2d) of the planes
  const uint8_t* loV = loK + d;
  for (int t0 = warp * R; t0 < T; t0 += kWarps * R) {
    RowRegs<VPL, FMT> vv[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r < T) vv[r].load(hiV + (int64_t)(t0 + r) * (4 * d), loV + (int64_t)(t0 + r) * (2 * d));
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < T) {
        float vf[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) vf[i] = vv[r].get(i);
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float pr = sc[(qi * H + head) * TS + t];
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc[qi][i] = fmaf(pr, vf[i], acc[qi][i]);
          }
        }
      }
    }
  }
  __syncthreads();                                 // scores no longer needed: reuse as reduction buffer
  float* red = sc;                                 // [warps][NQ*d]
#pragma unroll
  for (int qi = 0; qi < NQMAX; ++qi) {
    if (qi < NQ) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) red[(warp * NQ + qi) * d + lane * VPL + i] = acc[qi][i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * d; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * NQ * d + i];
    const int qi = i / d, c = i - qi * d;
    p.ctx[((int64_t)chunk * NQ + qi) * p.ctx_ld + c] = s;
  }
}


// =============================================================================================
// Fast path: one query per chunk (greedy), 256-column slices.  The generic kernel above spends ~12 instructions per
// element (run-time lanes-per-head modulo and shuffle loop, 64-bit address arithmetic and predicates per row, the
// query re-read from shared memory per row): at 3 bytes per element it is issue-bound (58 % issue-active at 32 %
// occupancy, profiles/r02_kv_formats.md) and no faster than the fp32 kernel.  Here
//   * a CTA owns ONE 256-column slice of a chunk (d = 256: the chunk; d = 512: heads 0-3 or 4-7 -- heads are
//     independent, so the slices need no exchange and the grid doubles),
//   * lane l owns columns [8l, 8l+8) of the slice, LPH = 4 (d = 256) or 8 (d = 512) lanes share a head; all of that is
//     compile-time,
//   * the query lives in 8 registers, the row steps and the key mask (sign of the step) in shared memory,
//   * rows go in blocks of LPH: the LPH x LPH (lane, row) partial sums are reduce-scattered with LPH - 1 shuffles, lane j
//     finishes row j and every lane stores one score (no predicates),
//   * 6 KB of rows per warp are requested before the first is consumed (8 rows of a 256-column slice, 16 of a
//     128-column one); ~120 registers -> 2 CTAs (16 warps, 96 KB of loads in flight) per SM; an 80-register build
//     (3 CTAs) spills row registers and measured slower (155 vs 148 us at d = 256).
// one stage of a reduce-scatter over lanes W apart: lanes whose bit W is set keep the upper N values, the others the lower N
template <int N, int W>
__device__ __forceinline__ void rs_stage(float* v, int lane) {
  const bool up = (lane & W) != 0;
#pragma unroll
  for (int r = 0; r < N; ++r) {
    const float keep = up ? v[r + N] : v[r];
    const float send = up ? v[r] : v[r + N];
    v[r] = keep + __shfl_xor_sync(ND_FULL, send, W);
  }
}

// CHAIN (the 128-column tail CTAs of d = 256, VPL = 4, LPH = 8): the lane pair (2k, 2k+1) holds the 8 columns ONE lane
// of the whole-chunk form (VPL = 8, LPH = 4) holds; the even lane's 4-product fma chain is handed to the odd lane, which
// continues it, and the four pair sums are then combined over the odd lanes in the order of the 4-lane reduce-scatter.
// With 8 rows per iteration and the same row -> warp assignment every sum is formed in the same order as in the
// whole-chunk CTA: a chunk gets the same bits whether it is decoded by one CTA or by two tail CTAs.
template <int VPL, int LPH, int FMT, bool CHAIN>
__device__ __forceinline__ void fast_body(const CrossAttnParams& p, const int chunk, const int part) {
  constexpr int DS = 32 * VPL, HP = 32 / LPH, RB = CHAIN ? 8 : 64 / VPL;   // slice width, heads per slice, rows per iteration
  static_assert(!CHAIN || (VPL == 4 && LPH == 8), "pair chaining is the 128-column form of d = 256");
  extern __shared__ __align__(16) float smem_f[];
  if (p.retired && p.retired[chunk]) return;
  const int T = p.T, d = p.d;
  const int Tup = (T + RB - 1) & ~(RB - 1);
  const int TS = Tup + LPH;                      // score pitch: lanes of a block hit 32 distinct banks, 16-byte multiple
  float* sc = smem_f;                            // [HP][TS]   (later red[warps][DS])
  float* kstep = sc + (HP * TS > kWarps * DS ? HP * TS : kWarps * DS);   // [Tup] key step, negative = masked key
  float* vstep = kstep + Tup;                    // [Tup]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = lane / LPH, j = lane % LPH;
  const int64_t row0 = (int64_t)chunk * T;

  float qr[VPL];
  {
    const float* qp = p.q + (int64_t)chunk * p.q_ld + part * DS + lane * VPL;
#pragma unroll
    for (int i = 0; i < VPL; ++i) qr[i] = qp[i] / p.q_div;
  }
  {
    const float* stp = p.kv_scale + row0 * 2;
    const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
    for (int t = threadIdx.x; t < Tup; t += kThreads) {
      float ks = 1.0f, vs = 1.0f;
      if (t < T) {
        if (fmt_scaled(FMT)) { ks = stp[2 * t]; vs = stp[2 * t + 1]; }
        if (srow && srow[t] == p.mask_value) ks = -ks;
      }
      kstep[t] = ks;
      vstep[t] = vs;
    }
  }
  __syncthreads();

  const uint32_t hi_pitch = 4u * d, lo_pitch = 2u * d;          // bytes per row of the planes ([K | V] columns)
  const uint8_t* hiK = reinterpret_cast<const uint8_t*>(p.kv_hi) + row0 * hi_pitch + part * (2 * DS) + lane * (2 * VPL);
  const uint8_t* loK = reinterpret_cast<const uint8_t*>(p.kv_lo) + row0 * lo_pitch + part * DS + lane * VPL;
  const int nit = Tup / RB;

  // ---------------- phase 1: scores.  UNR iterations' rows are requested together (the tail CTAs own 4 columns per
  // lane: two 8-row iterations = the 6 KB per warp the 8-column form has in flight); they are consumed in order.
  constexpr int UNR = CHAIN ? 2 : 1;
  for (int it = warp; it < nit; it += UNR * kWarps) {
    RowRegs<VPL, FMT> rr[UNR][RB];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int t0 = (it + u * kWarps) * RB;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t0 + r < T) {
          rr[u][r].load(hiK + (uint32_t)(t0 + r) * hi_pitch, loK + (uint32_t)(t0 + r) * lo_pitch);
        } else {
#pragma unroll
          for (int i = 0; i < RowRegs<VPL, FMT>::NHI; ++i) rr[u][r].hi[i] = 0u;
#pragma unroll
          for (int i = 0; i < (fmt_has_lo(FMT) ? RowRegs<VPL, FMT>::NLO : 1); ++i) rr[u][r].lo[i] = 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int t0 = (it + u * kWarps) * RB;
      if (t0 >= Tup) break;                                          // warp-uniform
      if constexpr (CHAIN) {
#pragma unroll
        for (int b0 = 0; b0 < RB; b0 += 4) {
          float v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float kf[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) kf[i] = rr[u][b0 + r].get(i);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) s = fmaf(qr[i], kf[i], s);
            s = __shfl_up_sync(ND_FULL, s, 1);                       // odd lanes: the even partner's chain
#pragma unroll
            for (int i = 0; i < VPL; ++i) s = fmaf(qr[i], kf[i], s);
            v[r] = s;                                                // complete 8-column chain in the odd lanes
          }
          rs_stage<2, 4>(v, lane);
          rs_stage<1, 2>(v, lane);
          const int t = t0 + b0 + ((lane >> 1) & 3);
          const float ks = kstep[t];
          if (lane & 1) sc[head * TS + t] = ks < 0.f ? -1e18f : v[0] * ks;
        }
      } else {
#pragma unroll
        for (int b0 = 0; b0 < RB; b0 += LPH) {
          float v[LPH];
#pragma unroll
          for (int r = 0; r < LPH; ++r) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) s = fmaf(qr[i], rr[u][b0 + r].get(i), s);
            v[r] = s;
          }
          // reduce-scatter over the LPH lanes of the head: lane j ends with the complete sum of row j
          if constexpr (LPH >= 8) rs_stage<4, 4>(v, lane);
          rs_stage<2, 2>(v, lane);
          rs_stage<1, 1>(v, lane);
          const int t = t0 + b0 + j;
          const float ks = kstep[t];
          sc[head * TS + t] = ks < 0.f ? -1e18f : v[0] * ks;     // rows >= T: finite junk, never read by the softmax
        }
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: softmax rows (torch.softmax: exp(x - max) / sum); afterwards the row holds
  // probability * value step, zero beyond T
  for (int row = warp; row < HP; row += kWarps) {
    float* s = sc + row * TS;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s[t] - m); s[t] = e; sum += e; }
    sum = warp_sum(sum);
    float* a = (p.attn && part == 0 && row == 0) ? p.attn + (int64_t)chunk * T : nullptr;
    for (int t = lane; t < Tup; t += 32) {
      float pr = 0.f;
      if (t < T) {
        pr = s[t] / sum;
        if (a) a[t] = pr;
        pr *= vstep[t];
      }
      s[t] = pr;
    }
  }
  __syncthreads();

  // ---------------- phase 3: context
  float acc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) acc[i] = 0.f;
  const uint8_t* hiV = hiK + 2 * d;                // V = columns [d, 2d) of the planes
  const uint8_t* loV = loK + d;
  for (int it = warp; it < nit; it += UNR * kWarps) {
    RowRegs<VPL, FMT> rr[UNR][RB];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int t0 = (it + u * kWarps) * RB;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (t0 + r < T) {
          rr[u][r].load(hiV + (uint32_t)(t0 + r) * hi_pitch, loV + (uint32_t)(t0 + r) * lo_pitch);
        } else {
#pragma unroll
          for (int i = 0; i < RowRegs<VPL, FMT>::NHI; ++i) rr[u][r].hi[i] = 0u;
#pragma unroll
          for (int i = 0; i < (fmt_has_lo(FMT) ? RowRegs<VPL, FMT>::NLO : 1); ++i) rr[u][r].lo[i] = 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int t0 = (it + u * kWarps) * RB;
      if (t0 >= Tup) break;
#pragma unroll
      for (int r4 = 0; r4 < RB; r4 += 4) {
        const float4 pq = *reinterpret_cast<const float4*>(sc + head * TS + t0 + r4);
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          acc[i] = fmaf(pq.x, rr[u][r4].get(i), fmaf(pq.y, rr[u][r4 + 1].get(i), fmaf(pq.z, rr[u][r4 + 2].get(i),
                   fmaf(pq.w, rr[u][r4 + 3].get(i), acc[i]))));
      }
    }
  }
  __syncthreads();                                 // scores no longer needed: reuse as reduction buffer
  float* red = sc;                                 // [warps][DS]
#pragma unroll
  for (int i = 0; i < VPL; ++i) red[warp * DS + lane * VPL + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < DS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * DS + threadIdx.x];
    p.ctx[(int64_t)chunk * p.ctx_ld + part * DS + threadIdx.x] = s;
  }
}

template <int VPL, int LPH, int FMT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) cross_attn_packed_fast_kernel(CrossAttnParams p, int split) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunk = blockIdx.x / split;
  fast_body<VPL, LPH, FMT, false>(p, chunk, blockIdx.x - chunk * split);
}

// d = 256 with a split tail.  One CTA per chunk leaves the last round of CTAs partly empty (1024 chunks on 148 x 2
// slots = 3.46 rounds: the 4th is 46 % full); two 128-column CTAs per chunk fill the rounds but run 7 % slower overall
// (8-byte loads).  So the first n_full chunks (whole rounds) take one CTA each and the rest -- dispatched last -- two
// 128-column CTAs each, which need ~0.6 of a whole-chunk CTA's time: 3 + 0.6 rounds instead of 4.  The tail CTAs form
// every sum in the order of the whole-chunk CTA (see CHAIN above), so a chunk's result does not depend on its position.
template <int FMT>
__global__ void __launch_bounds__(kThreads, 2) cross_attn_packed_tail_kernel(CrossAttnParams p, int n_full) {
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.x < n_full) {
    fast_body<8, 4, FMT, false>(p, blockIdx.x, 0);
  } else {
    const int i = blockIdx.x - n_full;
    fast_body<4, 8, FMT, true>(p, n_full + (i >> 1), i & 1);
  }
}

template <int FMT>
cudaError_t launch_tail(const CrossAttnParams& p, int n_full, cudaStream_t stream) {
  const int Tup = (p.T + 7) & ~7;
  const size_t a = (size_t)8 * (Tup + 4), b = (size_t)4 * (Tup + 8), red_f = (size_t)kWarps * 256;
  size_t sc_f = a > b ? a : b;
  if (red_f > sc_f) sc_f = red_f;
  const size_t smem = (sc_f + 2 * (size_t)Tup) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static PerDeviceFlag attr_set;
  bool& set = attr_set.cur();
  if (!set) {
    cudaError_t err = cudaFuncSetAttribute(cross_attn_packed_tail_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024);
    if (err != cudaSuccess) return err;
    set = true;
  }
  launch_k_heavy(cross_attn_packed_tail_kernel<FMT>, dim3(n_full + 2 * (p.n_chunks - n_full)), dim3(kThreads), smem, stream,
                 p, n_full);
  return cudaGetLastError();
}

template <int VPL, int LPH, int FMT, int MINB>
cudaError_t launch_fast(const CrossAttnParams& p, cudaStream_t stream) {
  constexpr int DS = 32 * VPL, RB = 64 / VPL;
  const int split = p.d / DS;
  const int Tup = (p.T + RB - 1) & ~(RB - 1), HP = 32 / LPH;
  const size_t sc_f = (size_t)HP * (Tup + LPH), red_f = (size_t)kWarps * DS;
  const size_t smem = ((sc_f > red_f ? sc_f : red_f) + 2 * (size_t)Tup) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static PerDeviceFlag attr_set;
  bool& set = attr_set.cur();
  if (!set) {
    cudaError_t err = cudaFuncSetAttribute(cross_attn_packed_fast_kernel<VPL, LPH, FMT, MINB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (err != cudaSuccess) return err;
    set = true;
  }
  launch_k_heavy(cross_attn_packed_fast_kernel<VPL, LPH, FMT, MINB>, dim3(p.n_chunks * split), dim3(kThreads), smem, stream, p, split);
  return cudaGetLastError();
}


// =============================================================================================
// Several queries per chunk (beam search: the K beams of a chunk share its keys / values).  Same slice layout as
// fast_body; the queries stay in shared memory (NQ x 8 registers per lane plus 48 row registers would spill at two CTAs
// per SM), each row is decoded ONCE and used by all NQ queries, scores and probabilities are [NQ][heads][T].
template <int VPL, int LPH, int FMT, int NQT>
__device__ __forceinline__ void fast_body_mq(const CrossAttnParams& p, const int chunk, const int part) {
  constexpr int DS = 32 * VPL, HP = 32 / LPH, RB = 64 / VPL;
  extern __shared__ __align__(16) float smem_f[];
  if (p.retired && p.retired[chunk]) return;
  const int T = p.T, d = p.d, NQ = p.NQ;
  const int Tup = (T + RB - 1) & ~(RB - 1);
  const int TS = Tup + LPH;
  const int sc_f = NQ * HP * TS, red_f = kWarps * NQ * DS;
  float* sc = smem_f;                            // [NQ][HP][TS]   (later red[warps][NQ][DS])
  float* kstep = sc + (sc_f > red_f ? sc_f : red_f);
  float* vstep = kstep + Tup;
  float* q_s = vstep + Tup;                      // [NQ][DS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = lane / LPH, j = lane % LPH;
  const int64_t row0 = (int64_t)chunk * T;

  for (int i = threadIdx.x; i < NQ * DS; i += kThreads) {
    const int qi = i / DS, c = i - qi * DS;
    q_s[i] = p.q[((int64_t)chunk * NQ + qi) * p.q_ld + part * DS + c] / p.q_div;
  }
  {
    const float* stp = p.kv_scale + row0 * 2;
    const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
    for (int t = threadIdx.x; t < Tup; t += kThreads) {
      float ks = 1.0f, vs = 1.0f;
      if (t < T) {
        if (fmt_scaled(FMT)) { ks = stp[2 * t]; vs = stp[2 * t + 1]; }
        if (srow && srow[t] == p.mask_value) ks = -ks;
      }
      kstep[t] = ks;
      vstep[t] = vs;
    }
  }
  __syncthreads();

  const uint32_t hi_pitch = 4u * d, lo_pitch = 2u * d;
  const uint8_t* hiK = reinterpret_cast<const uint8_t*>(p.kv_hi) + row0 * hi_pitch + part * (2 * DS) + lane * (2 * VPL);
  const uint8_t* loK = reinterpret_cast<const uint8_t*>(p.kv_lo) + row0 * lo_pitch + part * DS + lane * VPL;
  const int nit = Tup / RB;

  // ---------------- phase 1: scores
  for (int it = warp; it < nit; it += kWarps) {
    const int t0 = it * RB;
    RowRegs<VPL, FMT> rr[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (t0 + r < T) {
        rr[r].load(hiK + (uint32_t)(t0 + r) * hi_pitch, loK + (uint32_t)(t0 + r) * lo_pitch);
      } else {
#pragma unroll
        for (int i = 0; i < RowRegs<VPL, FMT>::NHI; ++i) rr[r].hi[i] = 0u;
#pragma unroll
        for (int i = 0; i < (fmt_has_lo(FMT) ? RowRegs<VPL, FMT>::NLO : 1); ++i) rr[r].lo[i] = 0u;
      }
    }
#pragma unroll
    for (int b0 = 0; b0 < RB; b0 += LPH) {
      float v[NQT][LPH];
#pragma unroll
      for (int r = 0; r < LPH; ++r) {
        float kf[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) kf[i] = rr[b0 + r].get(i);
#pragma unroll
        for (int qi = 0; qi < NQT; ++qi) {
          float s = 0.f;
          if (qi < NQ) {
            const float4 qa = *reinterpret_cast<const float4*>(q_s + qi * DS + lane * VPL);
            const float4 qb = *reinterpret_cast<const float4*>(q_s + qi * DS + lane * VPL + 4);
            s = fmaf(qa.x, kf[0], fmaf(qa.y, kf[1], fmaf(qa.z, kf[2], fmaf(qa.w, kf[3],
                fmaf(qb.x, kf[4], fmaf(qb.y, kf[5], fmaf(qb.z, kf[6], qb.w * kf[7])))))));
          }
          v[qi][r] = s;
        }
      }
      const int t = t0 + b0 + j;
      const float ks = kstep[t];
#pragma unroll
      for (int qi = 0; qi < NQT; ++qi) {
        if (qi < NQ) {                             // warp-uniform
          if constexpr (LPH >= 8) rs_stage<4, 4>(v[qi], lane);
          rs_stage<2, 2>(v[qi], lane);
          rs_stage<1, 1>(v[qi], lane);
          sc[(qi * HP + head) * TS + t] = ks < 0.f ? -1e18f : v[qi][0] * ks;
        }
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: softmax rows; afterwards probability * value step, zero beyond T
  for (int row = warp; row < NQ * HP; row += kWarps) {
    float* s = sc + row * TS;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s[t] - m); s[t] = e; sum += e; }
    sum = warp_sum(sum);
    float* a = (p.attn && part == 0 && (row % HP) == 0) ? p.attn + ((int64_t)chunk * NQ + row / HP) * T : nullptr;
    for (int t = lane; t < Tup; t += 32) {
      float pr = 0.f;
      if (t < T) {
        pr = s[t] / sum;
        if (a) a[t] = pr;
        pr *= vstep[t];
      }
      s[t] = pr;
    }
  }
  __syncthreads();

  // ---------------- phase 3: context
  float acc[NQT][VPL];
#pragma unroll
  for (int qi = 0; qi < NQT; ++qi)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[qi][i] = 0.f;
  const uint8_t* hiV = hiK + 2 * d;
  const uint8_t* loV = loK + d;
  for (int it = warp; it < nit; it += kWarps) {
    const int t0 = it * RB;
    RowRegs<VPL, FMT> rr[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (t0 + r < T) {
        rr[r].load(hiV + (uint32_t)(t0 + r) * hi_pitch, loV + (uint32_t)(t0 + r) * lo_pitch);
      } else {
#pragma unroll
        for (int i = 0; i < RowRegs<VPL, FMT>::NHI; ++i) rr[r].hi[i] = 0u;
#pragma unroll
        for (int i = 0; i < (fmt_has_lo(FMT) ? RowRegs<VPL, FMT>::NLO : 1); ++i) rr[r].lo[i] = 0u;
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float vf[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) vf[i] = rr[r].get(i);
#pragma unroll
      for (int qi = 0; qi < NQT; ++qi) {
        if (qi < NQ) {
          const float pr = sc[(qi * HP + head) * TS + t0 + r];
#pragma unroll
          for (int i = 0; i < VPL; ++i) acc[qi][i] = fmaf(pr, vf[i], acc[qi][i]);
        }
      }
    }
  }
  __syncthreads();
  float* red = sc;                                 // [warps][NQ][DS]
#pragma unroll
  for (int qi = 0; qi < NQT; ++qi)
    if (qi < NQ) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) red[(warp * NQ + qi) * DS + lane * VPL + i] = acc[qi][i];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * DS; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * NQ * DS + i];
    const int qi = i / DS, c = i - qi * DS;
    p.ctx[((int64_t)chunk * NQ + qi) * p.ctx_ld + part * DS + c] = s;
  }
}

template <int LPH, int FMT, int NQT>
__global__ void __launch_bounds__(kThreads, 2) cross_attn_packed_mq_kernel(CrossAttnParams p, int split) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunk = blockIdx.x / split;
  fast_body_mq<8, LPH, FMT, NQT>(p, chunk, blockIdx.x - chunk * split);
}

template <int LPH, int FMT, int NQT>
cudaError_t launch_mq(const CrossAttnParams& p, cudaStream_t stream) {
  constexpr int DS = 256, RB = 8, HP = 32 / LPH;
  const int split = p.d / DS;
  const int Tup = (p.T + RB - 1) & ~(RB - 1);
  const size_t sc_f = (size_t)p.NQ * HP * (Tup + LPH), red_f = (size_t)kWarps * p.NQ * DS;
  const size_t smem = ((sc_f > red_f ? sc_f : red_f) + 2 * (size_t)Tup + (size_t)p.NQ * DS) * sizeof(float);
  if (smem > 110 * 1024) return cudaErrorInvalidValue;         // two CTAs per SM
  static PerDeviceFlag attr_set;
  bool& set = attr_set.cur();
  if (!set) {
    cudaError_t err = cudaFuncSetAttribute(cross_attn_packed_mq_kernel<LPH, FMT, NQT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (err != cudaSuccess) return err;
    set = true;
  }
  launch_k_heavy(cross_attn_packed_mq_kernel<LPH, FMT, NQT>, dim3(p.n_chunks * split), dim3(kThreads), smem, stream, p, split);
  return cudaGetLastError();
}

template <int LPH>
cudaError_t launch_mq_any(const CrossAttnParams& p, cudaStream_t stream) {
  if (p.kv_fmt == KV_Q23M) return p.NQ <= 5 ? launch_mq<LPH, KV_Q23M, 5>(p, stream) : launch_mq<LPH, KV_Q23M, 8>(p, stream);
  if (p.kv_fmt == KV_Q15M) return p.NQ <= 5 ? launch_mq<LPH, KV_Q15M, 5>(p, stream) : launch_mq<LPH, KV_Q15M, 8>(p, stream);
  return cudaErrorInvalidValue;
}

// one query per chunk, 8 heads, d = 256 or 512, planes 32-byte aligned
bool fast_supported(const CrossAttnParams& p) {
  return p.NQ == 1 && p.H == 8 && (p.d == 256 || p.d == 512) && p.kv_fmt >= KV_Q23M && p.kv_fmt <= KV_FP24 &&
         (reinterpret_cast<uintptr_t>(p.kv_hi) & 31) == 0 && (reinterpret_cast<uintptr_t>(p.kv_lo) & 15) == 0 &&
         (int64_t)p.T * 4 * p.d < (int64_t)1 << 31;
}

int g_packed_fast = 1;     // 0: generic kernel; 1 / 2 / 3: slice kernels, see launch_fast_any

template <int VPL, int LPH, int MINB>
cudaError_t launch_fast_fmt(const CrossAttnParams& p, cudaStream_t stream) {
  switch (p.kv_fmt) {
    case KV_Q23M: return launch_fast<VPL, LPH, KV_Q23M, MINB>(p, stream);
    case KV_Q15M: return launch_fast<VPL, LPH, KV_Q15M, MINB>(p, stream);
    default: return launch_fast<VPL, LPH, KV_FP24, MINB>(p, stream);
  }
}

// slice choice.  d = 512 (dh = 64): 256-column slices of 4 heads (8 lanes x 8 columns per head), 2048 CTAs for 1024 chunks.
// d = 256 (dh = 32): g_packed_fast 1 (default) = one CTA per chunk with a split tail (cross_attn_packed_tail_kernel),
// 3 = one CTA per chunk throughout (4 lanes x 8 columns per head; 147.7 us at B = 1024),
// 2 = two 128-column CTAs per chunk throughout (8 lanes x 4 columns per head; 158.3 us).
cudaError_t launch_fast_any(const CrossAttnParams& p, cudaStream_t stream) {
  if (p.d == 512) return launch_fast_fmt<8, 8, 2>(p, stream);
  if (g_packed_fast == 3) return launch_fast_fmt<8, 4, 2>(p, stream);
  if (g_packed_fast == 2) return launch_fast_fmt<4, 8, 2>(p, stream);
  // default: whole rounds of one-CTA chunks, the remainder as two 128-column CTAs per chunk when that fits one round
  static int slots[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!slots[dev & 63]) {
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    slots[dev & 63] = 2 * n_sm;
  }
  const int sl = slots[dev & 63], rem = p.n_chunks % sl;
  const int n_full = (rem > 0 && 2 * rem <= sl && p.n_chunks > sl) ? p.n_chunks - rem : p.n_chunks;
  switch (p.kv_fmt) {
    case KV_Q23M: return launch_tail<KV_Q23M>(p, n_full, stream);
    case KV_Q15M: return launch_tail<KV_Q15M>(p, n_full, stream);
    default: return launch_tail<KV_FP24>(p, n_full, stream);
  }
}

template <int VPL, int NQMAX, int FMT>
cudaError_t launch_one(const CrossAttnParams& p, size_t smem, cudaStream_t stream) {
  static PerDeviceFlag attr_set;
  bool& set = attr_set.cur();
  if (!set) {
    cudaError_t err = cudaFuncSetAttribute(cross_attn_packed_kernel<VPL, NQMAX, FMT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (err != cudaSuccess) return err;
    set = true;
  }
  launch_k_heavy(cross_attn_packed_kernel<VPL, NQMAX, FMT>, dim3(p.n_chunks), dim3(kThreads), smem, stream, p);
  return cudaGetLastError();
}

template <int VPL>
cudaError_t launch_packed(const CrossAttnParams& p, cudaStream_t stream) {
  const int d = 32 * VPL;
  const size_t sc_f = (size_t)p.NQ * p.H * (p.T + 1);
  const size_t red_f = (size_t)kWarps * p.NQ * d;
  const size_t smem = ((size_t)p.NQ * d + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  switch (p.kv_fmt) {
    case KV_Q24: return p.NQ == 1 ? launch_one<VPL, 1, KV_Q24>(p, smem, stream) : launch_one<VPL, 8, KV_Q24>(p, smem, stream);
    case KV_Q16: return p.NQ == 1 ? launch_one<VPL, 1, KV_Q16>(p, smem, stream) : launch_one<VPL, 8, KV_Q16>(p, smem, stream);
    case KV_Q23M: return p.NQ == 1 ? launch_one<VPL, 1, KV_Q23M>(p, smem, stream) : launch_one<VPL, 8, KV_Q23M>(p, smem, stream);
    case KV_Q15M: return p.NQ == 1 ? launch_one<VPL, 1, KV_Q15M>(p, smem, stream) : launch_one<VPL, 8, KV_Q15M>(p, smem, stream);
    case KV_FP24: return p.NQ == 1 ? launch_one<VPL, 1, KV_FP24>(p, smem, stream) : launch_one<VPL, 8, KV_FP24>(p, smem, stream);
    default: return cudaErrorInvalidValue;
  }
}

// =============================================================================================
// Packer: one warp per (row, part); lane l quantises columns [l*VPL, (l+1)*VPL) of the part.  128-bit loads, the
// integers are packed in registers and leave as one or two vector stores per plane (per-element 2-byte / 1-byte stores
// made this kernel 5 ms per layer at d = 512: more than the projection GEMM in front of it).
template <int NB>
__device__ __forceinline__ void store_bytes(uint8_t* p, const uint32_t* r) {
  if constexpr (NB >= 16) {
#pragma unroll
    for (int i = 0; i < NB / 16; ++i)
      *reinterpret_cast<uint4*>(p + 16 * i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  } else if constexpr (NB == 8) {
    *reinterpret_cast<uint2*>(p) = make_uint2(r[0], r[1]);
  } else if constexpr (NB == 4) {
    *reinterpret_cast<uint32_t*>(p) = r[0];
  } else {
    *reinterpret_cast<uint16_t*>(p) = (uint16_t)r[0];
  }
}

template <int VPL>
__global__ void __launch_bounds__(256) kv_pack_kernel(const float* __restrict__ kmat, int64_t ldk,
                                                      const float* __restrict__ vmat, int64_t ldv, int64_t rows, int fmt,
                                                      int16_t* __restrict__ hi, uint8_t* __restrict__ lo,
                                                      float* __restrict__ scale) {
  constexpr int d = 32 * VPL;
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);       // row * 2 + part
  if (item >= rows * 2) return;
  const int64_t row = item >> 1;
  const int part = (int)(item & 1);
  const float* src = (part ? vmat + row * ldv : kmat + row * ldk) + lane * VPL;
  float x[VPL];
  if constexpr (VPL % 4 == 0) {
#pragma unroll
    for (int i = 0; i < VPL; i += 4) {
      const float4 t = ldg_stream4(src + i);
      x[i] = t.x; x[i + 1] = t.y; x[i + 2] = t.z; x[i + 3] = t.w;
    }
  } else {
    const float2 t = ldg_stream2(src);
    x[0] = t.x; x[1] = t.y;
  }
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) amax = fmaxf(amax, fabsf(x[i]));
  amax = warp_max(amax);
  // 2^e > amax (amax = f * 2^e, f in [0.5, 1)); clamp the exponent so the step stays a normal fp32 number
  int e = 0;
  if (amax > 0.f) frexpf(amax, &e);
  e = max(e, -90);
  const int bits = fmt == KV_Q24 ? 23 : (fmt == KV_Q23M ? 22 : 15);
  const float step = ldexpf(1.0f, e - bits);
  const float inv = ldexpf(1.0f, bits - e);
  const float lim = ldexpf(1.0f, bits) - 1.0f;
  if (lane == 0) scale[row * 2 + part] = fmt == KV_FP24 ? 1.0f : step;
  uint32_t hw[VPL >= 2 ? VPL / 2 : 1], lw[VPL >= 4 ? VPL / 4 : 1];
#pragma unroll
  for (int i = 0; i < (VPL >= 2 ? VPL / 2 : 1); ++i) hw[i] = 0u;
#pragma unroll
  for (int i = 0; i < (VPL >= 4 ? VPL / 4 : 1); ++i) lw[i] = 0u;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    uint32_t h16, l8 = 0u;
    if (fmt == KV_FP24) {
      // fp32 rounded to nearest-even at 16 significant bits: the top 24 bits of the pattern
      const uint32_t b = __float_as_uint(x[i]);
      const uint32_t r = (b + 0x7fu + ((b >> 8) & 1u)) >> 8;
      h16 = r >> 8;
      l8 = r & 255u;
    } else {
      const int m = (int)fminf(fmaxf(rintf(x[i] * inv), -lim), lim);         // x * 2^k is exact; rint ties to even
      if (fmt == KV_Q24 || fmt == KV_Q23M) {
        h16 = (uint32_t)(m >> 8) & 0xffffu;
        l8 = (uint32_t)m & 255u;
      } else {
        h16 = (uint32_t)m & 0xffffu;
      }
    }
    hw[i >> 1] |= h16 << (16 * (i & 1));
    lw[i >> 2] |= l8 << (8 * (i & 3));
  }
  store_bytes<2 * VPL>(reinterpret_cast<uint8_t*>(hi + row * (2 * d) + part * d + lane * VPL), hw);
  if (lo) store_bytes<VPL>(lo + row * (2 * d) + part * d + lane * VPL, lw);
}

}  // namespace

void cross_attention_packed_set_fast(int on) { g_packed_fast = on; }

bool cross_attention_packed_beams_ok(int NQ, int d, int H, int T, int fmt) {
  if (!(NQ > 1 && NQ <= 8 && H == 8 && (d == 256 || d == 512) && (fmt == KV_Q23M || fmt == KV_Q15M) && g_packed_fast)) return false;
  // d = 256: only through the shared-memory-ring kernel (the slice kernel below is arithmetic-bound there and loses to
  // the ring over fp32 rows: 369 vs 215 us on C3); d = 512: the slice kernel (no ring kernel at that width)
  if (d == 256) return g_cross_beam_kernel == 2 && cross_attention_ring_shape_ok(NQ, d, H, T, fmt);
  const int LPH = d == 256 ? 4 : 8, HP = 32 / LPH, Tup = (T + 7) & ~7;
  const size_t sc_f = (size_t)NQ * HP * (Tup + LPH), red_f = (size_t)kWarps * NQ * 256;
  return ((sc_f > red_f ? sc_f : red_f) + 2 * (size_t)Tup + (size_t)NQ * 256) * sizeof(float) <= 110 * 1024 &&
         (int64_t)T * 4 * d < (int64_t)1 << 31;
}

// several queries per chunk over fixed-point planes: the slice kernel with the queries in shared memory
bool cross_attention_packed_mq_supported(const CrossAttnParams& p) {
  if (!(p.NQ > 1 && p.NQ <= 8 && p.H == 8 && (p.d == 256 || p.d == 512) && (p.kv_fmt == KV_Q23M || p.kv_fmt == KV_Q15M) &&
        (reinterpret_cast<uintptr_t>(p.kv_hi) & 31) == 0 && (reinterpret_cast<uintptr_t>(p.kv_lo) & 15) == 0 &&
        (int64_t)p.T * 4 * p.d < (int64_t)1 << 31))
    return false;
  const int LPH = p.d == 256 ? 4 : 8, HP = 32 / LPH, Tup = (p.T + 7) & ~7;
  const size_t sc_f = (size_t)p.NQ * HP * (Tup + LPH), red_f = (size_t)kWarps * p.NQ * 256;
  return ((sc_f > red_f ? sc_f : red_f) + 2 * (size_t)Tup + (size_t)p.NQ * 256) * sizeof(float) <= 110 * 1024;
}

bool kv_pack_supported(int d) { return d == 64 || d == 128 || d == 256 || d == 512; }

cudaError_t kv_pack2(const float* k, int64_t ldk, const float* v, int64_t ldv, int64_t rows, int d, int fmt, int16_t* hi,
                     uint8_t* lo, float* scale, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (!kv_pack_supported(d) || fmt < KV_Q24 || fmt > KV_FP24 || (fmt_has_lo(fmt) && !lo)) return cudaErrorInvalidValue;
  if ((ldk & 3) || (ldv & 3) || (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15))
    return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)cdiv64(rows * 2, 8);
  switch (d / 32) {
    case 2: kv_pack_kernel<2><<<grid, 256, 0, stream>>>(k, ldk, v, ldv, rows, fmt, hi, lo, scale); break;
    case 4: kv_pack_kernel<4><<<grid, 256, 0, stream>>>(k, ldk, v, ldv, rows, fmt, hi, lo, scale); break;
    case 8: kv_pack_kernel<8><<<grid, 256, 0, stream>>>(k, ldk, v, ldv, rows, fmt, hi, lo, scale); break;
    default: kv_pack_kernel<16><<<grid, 256, 0, stream>>>(k, ldk, v, ldv, rows, fmt, hi, lo, scale); break;
  }
  return cudaGetLastError();
}

cudaError_t kv_pack(const float* kv, int64_t rows, int d, int fmt, int16_t* hi, uint8_t* lo, float* scale,
                    cudaStream_t stream) {
  return kv_pack2(kv, 2 * (int64_t)d, kv + d, 2 * (int64_t)d, rows, d, fmt, hi, lo, scale, stream);
}

// =============================================================================================
// Global attention of the RNN decoder (onmt/modules/global_attention.py:95-227; mlp: score = v . tanh(wq + uh),
// general / dot: score = q . H) and the conv attention of the CNN decoder (conv_multi_step_attention.py:38-82) over
// fixed-point planes: the same kernel shape as mlp_attn_kernel (attention.cu), rows decoded from int16 + uint8 planes.
// The key step cannot leave the tanh, but it costs nothing: tanh(fma(m, step, wq)) replaces tanh(wq + uh), and
// m * step is exact (power-of-two step).  The value step multiplies the probability once per row.
__device__ __forceinline__ float tanh_fast_p(float x) {          // 1 - 2 / (e^(2x) + 1), as attention.cu's tanh_fast
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

// (one query: 4 rows in flight and 64 registers, 4 CTAs per SM -- the first version, 8 rows / 106 registers / 2 CTAs, ran
// at 44 % issue-active with 23 % of the warp slots filled and LOST to fp32 rows: 207 vs 181 us)
template <int VPL, int NQMAX, int FMT>
__global__ void __launch_bounds__(kThreads, (NQMAX == 1 && VPL <= 8) ? 4 : 1) mlp_attn_packed_kernel(MlpAttnParams p) {
  constexpr int R = 4;                             // rows in flight per warp iteration
  extern __shared__ __align__(16) float smem_f[];
  const int chunk = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (p.retired && p.retired[chunk]) return;
  constexpr int d = 32 * VPL;
  const int T = p.T, NQ = p.NQ;
  float* q_s = smem_f;                            // [NQ][d]
  float* sc = q_s + NQ * d;                       // [NQ][T]  (later red[warps][NQ*d])
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NQ * d; i += kThreads) q_s[i] = p.wq[((int64_t)chunk * NQ + i / d) * d + (i % d)];
  float vreg[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) vreg[i] = p.dot ? 0.f : p.v[lane * VPL + i];
  __syncthreads();
  const int len = p.lengths ? (int)p.lengths[chunk] : T;
  const int64_t row0 = (int64_t)chunk * T;
  const uint32_t hi_pitch = 4u * d, lo_pitch = 2u * d;
  const uint8_t* hiK = reinterpret_cast<const uint8_t*>(p.kv_hi) + row0 * hi_pitch + lane * (2 * VPL);
  const uint8_t* loK = reinterpret_cast<const uint8_t*>(p.kv_lo) + row0 * lo_pitch + lane * VPL;
  const float* stp = p.kv_scale + row0 * 2;

  // ---------------- scores
  for (int t0 = warp * R; t0 < T; t0 += kWarps * R) {
    RowRegs<VPL, FMT> rr[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r < T) rr[r].load(hiK + (uint32_t)(t0 + r) * hi_pitch, loK + (uint32_t)(t0 + r) * lo_pitch);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < T) {
        const float ks = __ldg(stp + 2 * t);
        float u[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) u[i] = rr[r].get(i);
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float* qq = q_s + qi * d + lane * VPL;
            float s = 0.f;
            if (p.dot) {
#pragma unroll
              for (int i = 0; i < VPL; ++i) s = fmaf(qq[i], u[i], s);
              s *= ks;
            } else {
#pragma unroll
              for (int i = 0; i < VPL; ++i) s = fmaf(vreg[i], tanh_fast_p(fmaf(u[i], ks, qq[i])), s);
            }
            s = warp_sum(s);
            if (lane == 0) sc[qi * T + t] = (t < len) ? s : -INFINITY;     // sequence_mask, -inf
          }
        }
      }
    }
  }
  __syncthreads();
  // ---------------- softmax (the probability leaves with the value step folded in)
  for (int row = warp; row < NQ; row += kWarps) {
    float* s = sc + row * T;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s[t] - m); s[t] = e; sum += e; }
    sum = warp_sum(sum);
    float* a = p.attn ? p.attn + ((int64_t)chunk * NQ + row) * T : nullptr;
    for (int t = lane; t < T; t += 32) {
      const float pr = s[t] / sum;
      if (a) a[t] = pr;
      s[t] = pr * __ldg(stp + 2 * t + 1);
    }
  }
  __syncthreads();
  // ---------------- context
  float acc[NQMAX][VPL];
#pragma unroll
  for (int qi = 0; qi < NQMAX; ++qi)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[qi][i] = 0.f;
  const int tmax = len < T ? len : T;             // weights beyond the length are exactly 0
  const uint8_t* hiV = hiK + 2 * d;
  const uint8_t* loV = loK + d;
  for (int t0 = warp * R; t0 < tmax; t0 += kWarps * R) {
    RowRegs<VPL, FMT> rr[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r < tmax) rr[r].load(hiV + (uint32_t)(t0 + r) * hi_pitch, loV + (uint32_t)(t0 + r) * lo_pitch);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < tmax) {
        float vf[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) vf[i] = rr[r].get(i);
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float pr = sc[qi * T + t];
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc[qi][i] = fmaf(pr, vf[i], acc[qi][i]);
          }
        }
      }
    }
  }
  __syncthreads();
  float* red = sc;
#pragma unroll
  for (int qi = 0; qi < NQMAX; ++qi) {
    if (qi < NQ) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) red[(warp * NQ + qi) * d + lane * VPL + i] = acc[qi][i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * d; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * NQ * d + i];
    p.ctx[((int64_t)chunk * NQ + i / d) * p.ctx_ld + (i % d)] = s;
  }
}

template <int VPL, int NQMAX, int FMT>
cudaError_t launch_mlp_packed_one(const MlpAttnParams& p, cudaStream_t stream) {
  constexpr int d = 32 * VPL;
  const size_t sc_f = (size_t)p.NQ * p.T, red_f = (size_t)kWarps * p.NQ * d;
  const size_t smem = ((size_t)p.NQ * d + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static PerDeviceFlag attr_set;
  bool& set = attr_set.cur();
  if (!set) {
    cudaError_t err = cudaFuncSetAttribute(mlp_attn_packed_kernel<VPL, NQMAX, FMT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (err != cudaSuccess) return err;
    set = true;
  }
  launch_k_heavy(mlp_attn_packed_kernel<VPL, NQMAX, FMT>, dim3(p.n_chunks), dim3(kThreads), smem, stream, p);
  return cudaGetLastError();
}

template <int VPL>
cudaError_t launch_mlp_packed(const MlpAttnParams& p, cudaStream_t stream) {
  if (p.kv_fmt == KV_Q23M)
    return p.NQ == 1 ? launch_mlp_packed_one<VPL, 1, KV_Q23M>(p, stream) : launch_mlp_packed_one<VPL, 8, KV_Q23M>(p, stream);
  return p.NQ == 1 ? launch_mlp_packed_one<VPL, 1, KV_Q15M>(p, stream) : launch_mlp_packed_one<VPL, 8, KV_Q15M>(p, stream);
}

cudaError_t mlp_attention_packed(const MlpAttnParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  if (!kv_pack_supported(p.d) || p.NQ > 8 || p.NQ < 1 || !p.kv_hi || !p.kv_scale || (p.kv_fmt != KV_Q23M && p.kv_fmt != KV_Q15M) ||
      (p.kv_fmt == KV_Q23M && !p.kv_lo) || (int64_t)p.T * 4 * p.d >= (int64_t)1 << 31)
    return cudaErrorInvalidValue;
  switch (p.d / 32) {
    case 2: return launch_mlp_packed<2>(p, stream);
    case 4: return launch_mlp_packed<4>(p, stream);
    case 8: return launch_mlp_packed<8>(p, stream);
    default: return launch_mlp_packed<16>(p, stream);
  }
}

cudaError_t cross_attention_packed(const CrossAttnParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  if (!kv_pack_supported(p.d) || 32 % p.H || p.NQ > 8 || p.NQ < 1 || (p.d / p.H) % (p.d / 32) || !p.kv_hi ||
      !p.kv_scale || (fmt_has_lo(p.kv_fmt) && !p.kv_lo))
    return cudaErrorInvalidValue;
  if (g_packed_fast && fast_supported(p)) return launch_fast_any(p, stream);
  if (g_packed_fast && g_cross_beam_kernel == 2 && cross_attention_ring_supported(p)) return cross_attention_ring(p, stream);
  if (g_packed_fast && cross_attention_packed_mq_supported(p)) return p.d == 256 ? launch_mq_any<4>(p, stream) : launch_mq_any<8>(p, stream);
  switch (p.d / 32) {
    case 2: return launch_packed<2>(p, stream);
    case 4: return launch_packed<4>(p, stream);
    case 8: return launch_packed<8>(p, stream);
    default: return launch_packed<16>(p, stream);
  }
}

}  // namespace nd
