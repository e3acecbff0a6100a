// Memory-bound attention kernels of the decode loop and the encoder self-attention (fp32).
#include <float.h>

#include "kernels.cuh"

namespace nd {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;

template <int VPL>
__device__ __forceinline__ void load_slice(const float* p, float (&v)[VPL]) {
  if constexpr (VPL % 8 == 0) {
#pragma unroll
    for (int i = 0; i < VPL; i += 8) ldg_stream8(p + i, &v[i]);
  } else if constexpr (VPL % 4 == 0) {
#pragma unroll
    for (int i = 0; i < VPL; i += 4) {
      const float4 t = ldg_stream4(p + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else if constexpr (VPL % 2 == 0) {
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
      const float2 t = ldg_stream2(p + i);
      v[i] = t.x; v[i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = __ldg(p + i);
  }
}

// =============================================================================================
// Cross attention, one CTA per chunk.  d = 32*VPL; lane l owns columns [l*VPL, (l+1)*VPL) of every
// K / V row, so one warp reads one full row (coalesced 128-bit loads) and the 32/H lanes that share
// a head reduce their partial dot products with xor-shuffles.
//   phase 1: scores[q][h][t] = q_h . K[t]_h  (masked keys -> -1e18)     -> shared memory
//   phase 2: softmax over t per (q, h)                                   (warp per row)
//   phase 3: ctx[q] = sum_t p[q][h][t] * V[t]                            (register accumulators)
// K and V are each read exactly once per chunk per step: 2*T*d*4 bytes — the roofline of the decode.
template <int VPL, int NQMAX>
__global__ void __launch_bounds__(kAttnThreads) cross_attn_kernel(CrossAttnParams p) {
  // four rows of K / V in flight per warp iteration.  (Two rows and a minimum-blocks launch bound were tried for
  // d = 512: 485 -> 469 us there, but any second __launch_bounds__ argument changes the register allocation of the
  // d = 256 instance (64 -> 80/86 registers, 171 -> 177/214 us), so the plain form stays.)
  constexpr int R = 4;
  extern __shared__ __align__(16) float smem_f[];
  const int chunk = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (p.retired && p.retired[chunk]) return;
  const int d = 32 * VPL, T = p.T, H = p.H, NQ = p.NQ;
  const int TS = T + 1;                           // odd stride: the H score rows of a query hit distinct banks
  const int LPH = 32 / H;                         // lanes per head
  float* q_s = smem_f;                            // [NQ][d]
  float* sc = q_s + NQ * d;                       // [NQ*H][TS]   (later reused as red[warps][NQ*d])
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < NQ * d; i += kAttnThreads) {
    const int qi = i / d, c = i - qi * d;
    q_s[i] = p.q[((int64_t)chunk * NQ + qi) * p.q_ld + c] / p.q_div;
  }
  __syncthreads();

  const float* Kb = p.K + (int64_t)chunk * T * p.kv_ld + lane * VPL;
  const float* Vb = p.V + (int64_t)chunk * T * p.kv_ld + lane * VPL;
  const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
  const int head = lane / LPH;

  // ---------------- phase 1
  for (int t0 = warp * R; t0 < T; t0 += kAttnWarps * R) {
    float kv[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r < T) load_slice<VPL>(Kb + (int64_t)(t0 + r) * p.kv_ld, kv[r]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < T) {                                 // warp-uniform
        const bool masked = srow && (srow[t] == p.mask_value);
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float* qq = q_s + qi * d + lane * VPL;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) s = fmaf(qq[i], kv[r][i], s);
            for (int o = LPH >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(ND_FULL, s, o);
            if ((lane % LPH) == 0) sc[(qi * H + head) * TS + t] = masked ? -1e18f : s;
          }
        }
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: softmax rows (torch.softmax: exp(x - max) / sum)
  for (int row = warp; row < NQ * H; row += kAttnWarps) {
    float* s = sc + row * TS;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s[t] - m); s[t] = e; sum += e; }
    sum = warp_sum(sum);
    for (int t = lane; t < T; t += 32) s[t] = s[t] / sum;
    if (p.attn && (row % H) == 0) {
      float* a = p.attn + ((int64_t)chunk * NQ + row / H) * T;
      for (int t = lane; t < T; t += 32) a[t] = s[t];
    }
  }
  __syncthreads();

  // ---------------- phase 3
  float acc[NQMAX][VPL];
#pragma unroll
  for (int qi = 0; qi < NQMAX; ++qi)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[qi][i] = 0.f;
  for (int t0 = warp * R; t0 < T; t0 += kAttnWarps * R) {
    float vv[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r < T) load_slice<VPL>(Vb + (int64_t)(t0 + r) * p.kv_ld, vv[r]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < T) {
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float pr = sc[(qi * H + head) * TS + t];
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc[qi][i] = fmaf(pr, vv[r][i], acc[qi][i]);
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
  for (int i = threadIdx.x; i < NQ * d; i += kAttnThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) s += red[w * NQ * d + i];
    const int qi = i / d, c = i - qi * d;
    p.ctx[((int64_t)chunk * NQ + qi) * p.ctx_ld + c] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Beam-search variant for d = 256, H = 8 (the K beams of a chunk share ONE pass over its K / V).  The generic kernel
// above re-reads the queries from shared memory for every row and all-reduces every (query, row) score over the 4
// lanes of a head (606 us at K = 5 vs 171 us for one query).  Here a warp works on blocks of 4 rows:
//   * the next block's K (or V) rows are requested before the current block is consumed (register double buffer),
//   * per query the 4 x 4 (lane, row) partial sums are reduce-scattered with 3 shuffles: lane j of a head group ends
//     with the complete score of row j and writes it (every lane stores, no predication),
//   * the probabilities of a block come back as one LDS.128 per query.
template <int NQT>
__global__ void __launch_bounds__(kAttnThreads) cross_attn_beam_kernel(CrossAttnParams p) {
  constexpr int VPL = 8, d = 256, H = 8, LPH = 4;
  extern __shared__ __align__(16) float smem_f[];
  const int chunk = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (p.retired && p.retired[chunk]) return;
  const int T = p.T, NQ = p.NQ;
  const int TS = (T + 7) & ~3;                     // row stride of the score matrix: multiple of 4 (LDS.128), >= T + 4
  float* q_s = smem_f;                             // [NQ][d]
  float* sc = q_s + NQ * d;                        // [NQ*H][TS]   (later reused as red[warps][NQ*d])
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = lane / LPH, j4 = lane & 3;

  for (int i = threadIdx.x; i < NQ * d; i += kAttnThreads) {
    const int qi = i / d, c = i - qi * d;
    q_s[i] = p.q[((int64_t)chunk * NQ + qi) * p.q_ld + c] / p.q_div;
  }
  __syncthreads();
  const float* Kb = p.K + (int64_t)chunk * T * p.kv_ld + lane * VPL;
  const float* Vb = p.V + (int64_t)chunk * T * p.kv_ld + lane * VPL;
  const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
  const int nblk = (T + 3) >> 2;                   // blocks of 4 rows; block b belongs to warp b % 8

  auto load_block = [&](const float* base, int blk, float (&r)[4][VPL]) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int t = blk * 4 + rr;
      if (t < T) {
        ldg_stream8(base + (int64_t)t * p.kv_ld, r[rr]);
      } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) r[rr][i] = 0.f;
      }
    }
  };

  // ---------------- phase 1: scores
  {
    float cur[4][VPL], nxt[4][VPL];
    int blk = warp;
    if (blk < nblk) load_block(Kb, blk, cur);
    for (; blk < nblk; blk += kAttnWarps) {
      const int nb = blk + kAttnWarps;
      if (nb < nblk) load_block(Kb, nb, nxt);
      const int t = blk * 4 + j4;                  // the row whose complete scores this lane ends up with
      const bool masked = srow && t < T && (srow[t] == p.mask_value);
#pragma unroll
      for (int qi = 0; qi < NQT; ++qi) {
        if (qi < NQ) {
          const float4 qa = *reinterpret_cast<const float4*>(q_s + qi * d + lane * VPL);
          const float4 qb = *reinterpret_cast<const float4*>(q_s + qi * d + lane * VPL + 4);
          float v[4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
            v[rr] = fmaf(qa.x, cur[rr][0], fmaf(qa.y, cur[rr][1], fmaf(qa.z, cur[rr][2], fmaf(qa.w, cur[rr][3],
                    fmaf(qb.x, cur[rr][4], fmaf(qb.y, cur[rr][5], fmaf(qb.z, cur[rr][6], qb.w * cur[rr][7])))))));
          // reduce-scatter over the 4 lanes of the head: lane j4 keeps row j4
          const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
          const float a0 = (up2 ? v[2] : v[0]) + __shfl_xor_sync(ND_FULL, up2 ? v[0] : v[2], 2);
          const float a1 = (up2 ? v[3] : v[1]) + __shfl_xor_sync(ND_FULL, up2 ? v[1] : v[3], 2);
          const float sres = (up1 ? a1 : a0) + __shfl_xor_sync(ND_FULL, up1 ? a0 : a1, 1);
          if (t < T) sc[(qi * H + head) * TS + t] = masked ? -1e18f : sres;
        }
      }
      if (nb < nblk) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int i = 0; i < VPL; ++i) cur[rr][i] = nxt[rr][i];
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: softmax rows (torch.softmax: exp(x - max) / sum); pad the tail of each row with zeros
  for (int row = warp; row < NQ * H; row += kAttnWarps) {
    float* srw = sc + row * TS;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, srw[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(srw[t] - m); srw[t] = e; sum += e; }
    sum = warp_sum(sum);
    for (int t = lane; t < T; t += 32) srw[t] = srw[t] / sum;
    for (int t = T + lane; t < TS; t += 32) srw[t] = 0.f;
    if (p.attn && (row % H) == 0) {
      float* a = p.attn + ((int64_t)chunk * NQ + row / H) * T;
      for (int t = lane; t < T; t += 32) a[t] = srw[t];
    }
  }
  __syncthreads();

  // ---------------- phase 3: context
  float acc[NQT][VPL];
#pragma unroll
  for (int qi = 0; qi < NQT; ++qi)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[qi][i] = 0.f;
  {
    float cur[4][VPL], nxt[4][VPL];
    int blk = warp;
    if (blk < nblk) load_block(Vb, blk, cur);
    for (; blk < nblk; blk += kAttnWarps) {
      const int nb = blk + kAttnWarps;
      if (nb < nblk) load_block(Vb, nb, nxt);
#pragma unroll
      for (int qi = 0; qi < NQT; ++qi) {
        if (qi < NQ) {
          const float4 pr = *reinterpret_cast<const float4*>(sc + (qi * H + head) * TS + blk * 4);
#pragma unroll
          for (int i = 0; i < VPL; ++i)
            acc[qi][i] = fmaf(pr.x, cur[0][i], fmaf(pr.y, cur[1][i], fmaf(pr.z, cur[2][i], fmaf(pr.w, cur[3][i], acc[qi][i]))));
        }
      }
      if (nb < nblk) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int i = 0; i < VPL; ++i) cur[rr][i] = nxt[rr][i];
      }
    }
  }
  __syncthreads();                                 // scores no longer needed: reuse as reduction buffer
  float* red = sc;                                 // [warps][NQ*d]
#pragma unroll
  for (int qi = 0; qi < NQT; ++qi) {
    if (qi < NQ) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) red[(warp * NQ + qi) * d + lane * VPL + i] = acc[qi][i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * d; i += kAttnThreads) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) sum += red[w * NQ * d + i];
    const int qi = i / d, c = i - qi * d;
    p.ctx[((int64_t)chunk * NQ + qi) * p.ctx_ld + c] = sum;
  }
}

template <int NQT>
cudaError_t launch_cross_beam(const CrossAttnParams& p, cudaStream_t stream) {
  const int TS = (p.T + 7) & ~3;
  const size_t sc_f = (size_t)p.NQ * 8 * TS;
  const size_t red_f = (size_t)kAttnWarps * p.NQ * 256;
  const size_t smem = ((size_t)p.NQ * 256 + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  cudaFuncSetAttribute(cross_attn_beam_kernel<NQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_k_heavy(cross_attn_beam_kernel<NQT>, dim3(p.n_chunks), dim3(kAttnThreads), smem, stream, p);
  return cudaGetLastError();
}

template <int VPL>
cudaError_t launch_cross(const CrossAttnParams& p, cudaStream_t stream) {
  const int d = 32 * VPL;
  const size_t sc_f = (size_t)p.NQ * p.H * (p.T + 1);
  const size_t red_f = (size_t)kAttnWarps * p.NQ * d;
  const size_t smem = ((size_t)p.NQ * d + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  if (p.NQ == 1) {
    cudaFuncSetAttribute(cross_attn_kernel<VPL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k_heavy(cross_attn_kernel<VPL, 1>, dim3(p.n_chunks), dim3(kAttnThreads), smem, stream, p);
  } else {
    cudaFuncSetAttribute(cross_attn_kernel<VPL, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k_heavy(cross_attn_kernel<VPL, 8>, dim3(p.n_chunks), dim3(kAttnThreads), smem, stream, p);
  }
  return cudaGetLastError();
}

}  // namespace

// several queries per chunk at d = 256: 2 = persistent shared-memory-ring kernel (cross_attn_ring.cu), 1 = register
// prefetch kernel above, 0 = the generic kernel (cross-checks)
int g_cross_beam_kernel = 2;
void cross_attention_set_beam_kernel(int on) { g_cross_beam_kernel = on; }

cudaError_t cross_attention(const CrossAttnParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  if (p.kv_fmt != KV_F32) return cross_attention_packed(p, stream);
  if (p.d % 32 || 32 % p.H || p.NQ > 8 || p.NQ < 1 || (p.d / p.H) % (p.d / 32)) return cudaErrorInvalidValue;
  if (g_cross_beam_kernel == 2 && cross_attention_ring_supported(p)) return cross_attention_ring(p, stream);
  if (g_cross_beam_kernel && p.NQ > 1 && p.d == 256 && p.H == 8 && (p.kv_ld % 8) == 0 &&
      (reinterpret_cast<uintptr_t>(p.K) & 31) == 0 && (reinterpret_cast<uintptr_t>(p.V) & 31) == 0) {
    if (p.NQ <= 4) return launch_cross_beam<4>(p, stream);
    if (p.NQ == 5) return launch_cross_beam<5>(p, stream);
    return launch_cross_beam<8>(p, stream);
  }
  switch (p.d / 32) {
    case 1: return launch_cross<1>(p, stream);
    case 2: return launch_cross<2>(p, stream);
    case 4: return launch_cross<4>(p, stream);
    case 8: return launch_cross<8>(p, stream);
    case 16: return launch_cross<16>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

// =============================================================================================
// Cross attention in memory-bank space (see CrossMbParams).  One CTA (256 threads) per chunk streams the
// chunk's memory bank through shared memory ONCE in tiles of 32 positions (one 32 KB cp.async.bulk per tile,
// mbarrier transaction counts, NST stages in flight) and keeps an online softmax per head:
//   phase A  S[h][t] = mb[t] . qt_h      lane = position, warp g = column slice [g d/8, (g+1) d/8): 128-bit row
//            reads rotated by the lane index (conflict-free without padding); 8 partial sums per (h, t) combined
//            through shared memory
//   softmax  warp h: mask, running max / sum, p = exp(S - m), rescale factor for the accumulators
//   phase B  acc[h][j] += p[h][t] mb[t][j]  thread = 4 columns x all heads, the 1024/d thread groups split the rows
// fp32 FFMA throughout (4 T d H flops per chunk): at d = 256 the FFMA time is about the HBM time of the
// T*d*4-byte read, so the kernel stays on the HBM roofline at half the bytes of the K/V formulation.
namespace {

constexpr int kMbTT = 32;                          // positions per tile

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int H>
__global__ void __launch_bounds__(256) cross_attn_mb_kernel(CrossMbParams p, int nst) {
  extern __shared__ __align__(16) float smem_f[];
  pdl_launch_dependents();
  pdl_wait();
  const int chunk = blockIdx.x;
  const int d = p.d, T = p.T;
  const int pitch = d;                             // a tile is a contiguous [32][d] block: ONE bulk copy
  const int tile_f = kMbTT * pitch;
  float* tiles = smem_f;                           // [nst][32][pitch]
  float* qt_s = tiles + (size_t)nst * tile_f;      // [H][d]
  float* part = qt_s + H * d;                      // [8][H][32] phase-A partial sums
  float* prob = part + 8 * H * kMbTT;              // [32][H]
  float* alpha_s = prob + kMbTT * H;               // [H] accumulator rescale of this tile, then 1/l at the end
  uint64_t* full = reinterpret_cast<uint64_t*>(alpha_s + H);   // [nst]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (T + kMbTT - 1) / kMbTT;
  const float* mb = p.mb + (int64_t)chunk * T * d;
  const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // producer = one thread: a tile of 32 positions is rows*d*4 contiguous bytes of the memory bank
  auto issue_tile = [&](int i) {
    const int s = i % nst;
    const int rows = min(kMbTT, T - i * kMbTT);
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(rows * d * 4);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(&full[s])), "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr_u32(tiles + (size_t)s * tile_f)),
                   "l"(mb + (int64_t)i * kMbTT * d), "r"(bytes), "r"(smem_addr_u32(&full[s]))
                   : "memory");
    }
  };
  if (warp == 0)
    for (int i = 0; i < nst && i < ntiles; ++i) issue_tile(i);
  for (int i = tid; i < H * d; i += 256) qt_s[i] = p.qt[(int64_t)chunk * H * d + i];

  const int JW = d / 8;                            // phase-A columns per warp
  const int TPR = d / 4;                           // phase-B threads per row
  const int RG = 256 / TPR;                        // phase-B row groups
  const int RPG = kMbTT / RG;                      // rows per group and tile
  const int rg = tid / TPR, j4 = (tid % TPR) * 4;
  float acc[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h) { acc[h][0] = 0.f; acc[h][1] = 0.f; acc[h][2] = 0.f; acc[h][3] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;            // running max / sum of head `warp` (warp-uniform)
  __syncthreads();

  for (int i = 0; i < ntiles; ++i) {
    const int s = i % nst;
    const float* tile = tiles + (size_t)s * tile_f;
    const int rows = min(kMbTT, T - i * kMbTT);
    {
      const uint32_t parity = (uint32_t)((i / nst) & 1);
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra DONE_%=;\n"
          "bra WAIT_%=;\n"
          "DONE_%=:\n"
          "}\n" ::"r"(smem_addr_u32(&full[s])),
          "r"(parity)
          : "memory");
    }
    // ---- phase A
    {
      float sa[H];
#pragma unroll
      for (int h = 0; h < H; ++h) sa[h] = 0.f;
      if (lane < rows) {
        // rows are d*4 bytes apart (a multiple of 128): lane r starts at 16-byte unit r of its slice and
        // wraps, so the 8 lanes of a shared-memory wavefront touch 8 different bank groups
        const float* mrow = tile + lane * pitch + warp * JW;
        const float* qcol = qt_s + warp * JW;
        const int nch = JW >> 2;
        for (int ci = 0; ci < nch; ++ci) {
          const int c = ((ci + lane) & (nch - 1)) << 2;
          const float4 mv = *reinterpret_cast<const float4*>(mrow + c);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float4 qv = *reinterpret_cast<const float4*>(qcol + h * d + c);
            sa[h] = fmaf(mv.x, qv.x, sa[h]);
            sa[h] = fmaf(mv.y, qv.y, sa[h]);
            sa[h] = fmaf(mv.z, qv.z, sa[h]);
            sa[h] = fmaf(mv.w, qv.w, sa[h]);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < H; ++h) part[(warp * H + h) * kMbTT + lane] = sa[h];
    }
    __syncthreads();
    // ---- online softmax: warp h owns head h (H <= 8 warps)
    if (warp < H) {
      float sc = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sc += part[(g * H + warp) * kMbTT + lane];
      const int t = i * kMbTT + lane;
      if (t >= T) sc = -INFINITY;
      else if (srow && srow[t] == p.mask_value) sc = -1e18f;          // masked_fill(mask, -1e18)
      const float m_new = fmaxf(m_run, warp_max(sc));
      const float al = expf(m_run - m_new);                           // 0 on the first tile
      const float pr = expf(sc - m_new);
      l_run = l_run * al + warp_sum(pr);
      m_run = m_new;
      prob[lane * H + warp] = pr;
      if (lane == 0) alpha_s[warp] = al;
    }
    __syncthreads();
    // ---- phase B
    {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float al = alpha_s[h];
        acc[h][0] *= al; acc[h][1] *= al; acc[h][2] *= al; acc[h][3] *= al;
      }
      const int r1 = min(rows, (rg + 1) * RPG);
      for (int r = rg * RPG; r < r1; ++r) {
        const float4 mv = *reinterpret_cast<const float4*>(tile + r * pitch + j4);
        float pw[H];
#pragma unroll
        for (int h = 0; h < H; h += 4) {
          const float4 pv = *reinterpret_cast<const float4*>(prob + r * H + h);
          pw[h] = pv.x; pw[h + 1] = pv.y; pw[h + 2] = pv.z; pw[h + 3] = pv.w;
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
          acc[h][0] = fmaf(pw[h], mv.x, acc[h][0]);
          acc[h][1] = fmaf(pw[h], mv.y, acc[h][1]);
          acc[h][2] = fmaf(pw[h], mv.z, acc[h][2]);
          acc[h][3] = fmaf(pw[h], mv.w, acc[h][3]);
        }
      }
    }
    __syncthreads();                               // stage s, part and prob are free again
    if (warp == 0 && i + nst < ntiles) issue_tile(i + nst);
  }

  // ---- combine the row groups, normalise by the softmax sum, write ctxt[h][j]
  if (warp < H && lane == 0) alpha_s[warp] = 1.0f / l_run;
  float* red = tiles;                              // [RG][H][d] = 1024 H floats <= one stage
#pragma unroll
  for (int h = 0; h < H; ++h)
    *reinterpret_cast<float4*>(red + ((size_t)rg * H + h) * d + j4) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
  __syncthreads();
  float* out = p.ctxt + (int64_t)chunk * H * d;
  for (int i = tid; i < H * d; i += 256) {
    float sum = 0.f;
    for (int g = 0; g < RG; ++g) sum += red[(size_t)g * H * d + i];
    out[i] = sum * alpha_s[i / d];
  }
}

// ---------------------------------------------------------------------------------------------
// v2 for d = 256, H = 8: register-resident query and accumulators, packed fp32 FMAs (FFMA2).
//   lane  = column slice [8 lane, 8 lane + 8)      warp w = positions 4w .. 4w+3 of every 32-position tile
//   qt[8 heads][8 cols] and acc[8][8] live in registers for the whole chunk; the 4 x 8 tile values a thread
//   loads (two LDS.128 per row, conflict free) feed both the scores and the context, so each memory-bank
//   element crosses shared memory once.  Per tile: 256 FFMA2 per thread, one 32-value butterfly reduce-scatter
//   (lane L ends with the score of position L>>3, head L&7), ONE CTA barrier; every warp recomputes the
//   per-head running max / sum of the online softmax from the 8 x 32 score tile in shared memory.
constexpr int kMb2D = 256, kMb2H = 8;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__global__ void __launch_bounds__(256, 1) cross_attn_mb2_kernel(CrossMbParams p, int nst) {
  extern __shared__ __align__(16) float smem_f[];
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = kMb2D, H = kMb2H;
  const int chunk = blockIdx.x, T = p.T;
  const int tile_f = kMbTT * D;
  float* tiles = smem_f;                              // [nst][32][256]
  float* sbuf = tiles + (size_t)nst * tile_f;         // [2][H][32] scores of the current / next tile
  float* l_s = sbuf + 2 * H * kMbTT;                  // [H] 1 / softmax sum
  uint64_t* full = reinterpret_cast<uint64_t*>(l_s + H);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (T + kMbTT - 1) / kMbTT;
  const float* mb = p.mb + (int64_t)chunk * T * D;
  const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_tile = [&](int i) {                      // one thread: rows*1 KB contiguous bytes of the memory bank
    const int s = i % nst;
    const uint32_t bytes = (uint32_t)(min(kMbTT, T - i * kMbTT) * D * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(&full[s])), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_u32(tiles + (size_t)s * tile_f)),
                 "l"(mb + (int64_t)i * kMbTT * D), "r"(bytes), "r"(smem_addr_u32(&full[s]))
                 : "memory");
  };
  if (tid == 0)
    for (int i = 0; i < nst && i < ntiles; ++i) issue_tile(i);

  // query (already projected into memory-bank space) of this thread's 8 columns, all heads
  float2 qt[H][4], acc[H][4];
  {
    const float* q = p.qt + (int64_t)chunk * H * D + 8 * lane;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float4 a = *reinterpret_cast<const float4*>(q + h * D), b = *reinterpret_cast<const float4*>(q + h * D + 4);
      qt[h][0] = make_float2(a.x, a.y); qt[h][1] = make_float2(a.z, a.w);
      qt[h][2] = make_float2(b.x, b.y); qt[h][3] = make_float2(b.z, b.w);
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[h][c] = make_float2(0.f, 0.f);
    }
  }
  const int hh = lane & 7, rq = lane >> 3;            // this lane's (head, position-in-warp / row quarter) after the reduce-scatter
  float m_run = -INFINITY, l_run = 0.f;               // running max / sum of head hh (replicated in every warp)

  for (int i = 0; i < ntiles; ++i) {
    const int s = i % nst, buf = i & 1;
    const float* tile = tiles + (size_t)s * tile_f;
    const int rows = min(kMbTT, T - i * kMbTT);
    // mask of this lane's position (prefetched before the wait)
    const int t_own = i * kMbTT + 4 * warp + rq;
    const float src_own = (srow && t_own < T) ? srow[t_own] : 0.f;
    {
      const uint32_t parity = (uint32_t)((i / nst) & 1);
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra DONE_%=;\n"
          "bra WAIT_%=;\n"
          "DONE_%=:\n"
          "}\n" ::"r"(smem_addr_u32(&full[s])),
          "r"(parity)
          : "memory");
    }
    // ---- this thread's 4 positions x 8 columns (zero beyond the end of the chunk)
    float2 mv[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 4 * warp + r;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (row < rows) {
        a = *reinterpret_cast<const float4*>(tile + row * D + 8 * lane);
        b = *reinterpret_cast<const float4*>(tile + row * D + 8 * lane + 4);
      }
      mv[r][0] = make_float2(a.x, a.y); mv[r][1] = make_float2(a.z, a.w);
      mv[r][2] = make_float2(b.x, b.y); mv[r][3] = make_float2(b.z, b.w);
    }
    // ---- partial scores of (position r, head h) over this lane's columns
    float v[32];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) a2 = ffma2(mv[r][c], qt[h][c], a2);
        v[r * 8 + h] = a2.x + a2.y;
      }
    // ---- butterfly reduce-scatter over the 32 lanes: lane L ends with the full sum of value L
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int k = 0; k < off; ++k) {
        const float keep = up ? v[k + off] : v[k];
        const float send = up ? v[k] : v[k + off];
        v[k] = keep + __shfl_xor_sync(ND_FULL, send, off);
      }
    }
    float sc = v[0];                                  // position 4 warp + rq, head hh
    if (t_own >= T) sc = -INFINITY;
    else if (srow && src_own == p.mask_value) sc = -1e18f;          // masked_fill(mask, -1e18)
    sbuf[(buf * H + hh) * kMbTT + 4 * warp + rq] = sc;
    __syncthreads();                                  // scores of the tile complete; stage s fully read
    if (tid == 0 && i + nst < ntiles) issue_tile(i + nst);
    // ---- online softmax of head hh (every warp, redundantly): lanes (hh, rq) cover positions 8 rq .. 8 rq + 7
    float pr_own;
    {
      const float4 s0 = *reinterpret_cast<const float4*>(sbuf + (buf * H + hh) * kMbTT + 8 * rq);
      const float4 s1 = *reinterpret_cast<const float4*>(sbuf + (buf * H + hh) * kMbTT + 8 * rq + 4);
      float tm = fmaxf(fmaxf(fmaxf(s0.x, s0.y), fmaxf(s0.z, s0.w)), fmaxf(fmaxf(s1.x, s1.y), fmaxf(s1.z, s1.w)));
      tm = fmaxf(tm, __shfl_xor_sync(ND_FULL, tm, 8));
      tm = fmaxf(tm, __shfl_xor_sync(ND_FULL, tm, 16));
      const float m_new = fmaxf(m_run, tm);
      float ts = expf(s0.x - m_new) + expf(s0.y - m_new) + expf(s0.z - m_new) + expf(s0.w - m_new) +
                 expf(s1.x - m_new) + expf(s1.y - m_new) + expf(s1.z - m_new) + expf(s1.w - m_new);
      ts += __shfl_xor_sync(ND_FULL, ts, 8);
      ts += __shfl_xor_sync(ND_FULL, ts, 16);
      const float al = expf(m_run - m_new);           // 0 on the first tile
      l_run = l_run * al + ts;
      m_run = m_new;
      pr_own = expf(sc - m_new);
      // rescale the accumulators when a running max moved (rare after the first tiles)
      if (__any_sync(ND_FULL, al != 1.0f)) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float a = __shfl_sync(ND_FULL, al, h);
          const float2 a2 = make_float2(a, a);
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[h][c] = make_float2(acc[h][c].x * a2.x, acc[h][c].y * a2.y);
        }
      }
    }
    // ---- context: acc[h][cols] += p[position r][h] * mb[position r][cols]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float pw = __shfl_sync(ND_FULL, pr_own, r * 8 + h);
        const float2 p2 = make_float2(pw, pw);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[h][c] = ffma2(p2, mv[r][c], acc[h][c]);
      }
  }

  // ---- combine the 8 warps (each saw a quarter... an eighth of the positions), normalise, write ctxt[h][col]
  if (warp == 0 && lane < H) l_s[lane] = 1.0f / l_run;           // lane = hh for lanes 0..7
  __syncthreads();                                    // every warp is done with the tiles
  float* red = tiles;                                 // [8 warps][H][256] = 64 KB
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float* o = red + ((size_t)warp * H + h) * D + 8 * lane;
    *reinterpret_cast<float4*>(o) = make_float4(acc[h][0].x, acc[h][0].y, acc[h][1].x, acc[h][1].y);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[h][2].x, acc[h][2].y, acc[h][3].x, acc[h][3].y);
  }
  __syncthreads();
  float* out = p.ctxt + (int64_t)chunk * H * D;
  for (int i = tid; i < H * D; i += 256) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[(size_t)w * H * D + i];
    out[i] = sum * l_s[i / D];
  }
}

}  // namespace

bool cross_attention_mb_supported(int d, int H) {
  return H == 8 && (d == 32 || d == 64 || d == 128 || d == 256 || d == 512 || d == 1024);
}

int g_cross_mb_version = 2;        // 2: register-resident v2 when d = 256; 1: always the generic v1 (experiments)
void cross_attention_mb_set_version(int v) { g_cross_mb_version = v; }

cudaError_t cross_attention_mb(const CrossMbParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  if (!cross_attention_mb_supported(p.d, p.H)) return cudaErrorInvalidValue;
  if (g_cross_mb_version == 2 && p.d == kMb2D && p.H == kMb2H) {
    const int nst = 5;                                 // 5 x 32 KB tiles in flight per SM (one CTA per SM)
    const size_t smem = (size_t)nst * kMbTT * kMb2D * 4 + (2 * kMb2H * kMbTT + kMb2H) * 4 + 64;
    static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(cross_attn_mb2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    return launch_k_heavy(cross_attn_mb2_kernel, dim3(p.n_chunks), dim3(256), smem, stream, p, nst);
  }
  constexpr int H = 8;
  const size_t tile_b = (size_t)kMbTT * p.d * sizeof(float);
  const size_t fixed = ((size_t)H * p.d + 8 * H * kMbTT + kMbTT * H + H) * sizeof(float) + 64;
  int nst = 4;
  while (nst > 2 && nst * tile_b + fixed > 100 * 1024) --nst;        // two CTAs per SM when the tiles allow it
  const size_t smem = nst * tile_b + fixed;
  static size_t attr_smem_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  size_t& attr_smem = attr_smem_dev[dev_ & 63];
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(cross_attn_mb_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  return launch_k_heavy(cross_attn_mb_kernel<H>, dim3(p.n_chunks), dim3(256), smem, stream, p, nst);
}

// =============================================================================================
// Decode-step self attention.  One CTA per row, one warp per head (looping when H > warps).
namespace {

// Cache layout: [slot][head][pos][dh] so that one head's keys / values of consecutive positions are
// contiguous: dh/4 lanes cover one position with 128-bit loads, 32/(dh/4) positions per warp access.
// U = independent 128-bit loads in flight per lane (the kernel is latency bound); DH = head size: with run-time lanes per
// position the shuffle reductions were loops and 2/3 of the executed instructions were address and control work
// (ncu source page, profiles/r02l_*: IMAD / BRA / LEA / ISETP 48 %, FFMA 6 %)
template <int U, int DH>
__global__ void __launch_bounds__(256) self_attn_kernel(SelfAttnParams p) {
  extern __shared__ __align__(16) float smem_f[];
  const int row = p.row0 + blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (p.retired && p.retired[row / p.rows_per_chunk]) return;
  constexpr int dh = DH;
  const int d = p.d, H = p.H, L = p.step + 1;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int LPP = DH >> 2;                     // lanes per position (dh = 8/16/32/64 -> 2/4/8/16)
  constexpr int PPI = 32 / LPP;                    // positions per warp iteration
  const int sub = lane % LPP, grp = lane / LPP;
  float* q_s = smem_f;                             // [d]
  float* p_s = smem_f + d;                         // [nwarps][Lmax]
  const float* qkv = p.qkv + (int64_t)row * 3 * d;
  // append this step's k, v to the cache (own slot = row) and stage q
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const int h = i / dh, e = i - h * dh;
    const int64_t o = (((int64_t)row * H + h) * p.Lmax + p.step) * dh + e;
    q_s[i] = qkv[i] / p.q_div;
    p.Kc[o] = qkv[d + i];
    p.Vc[o] = qkv[2 * d + i];
  }
  __syncthreads();
  const int* anc = p.anc ? p.anc + (int64_t)row * p.anc_ld : nullptr;
  for (int h = warp; h < H; h += nwarps) {
    float* ps = p_s + warp * p.Lmax;
    const float4 qq = *reinterpret_cast<const float4*>(q_s + h * dh + sub * 4);
    const int64_t own = (((int64_t)row * H + h) * p.Lmax) * dh + sub * 4;       // this row's own cache slot (greedy)
    // ---- scores: U independent 128-bit loads in flight per lane (the kernel is latency bound)
    float m = -FLT_MAX;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < L; j0 += U * PPI) {
      float4 kk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * PPI + grp;
        kk[u] = zero4;
        if (j < L) {
          const float* kr = (j == p.step) ? qkv + d + h * dh + sub * 4
                            : anc     ? p.Kc + (((int64_t)anc[j] * H + h) * p.Lmax + j) * dh + sub * 4
                                      : p.Kc + own + j * dh;
          kk[u] = *reinterpret_cast<const float4*>(kr);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * PPI + grp;
        float sc = fmaf(qq.x, kk[u].x, fmaf(qq.y, kk[u].y, fmaf(qq.z, kk[u].z, qq.w * kk[u].w)));
#pragma unroll
        for (int o = LPP >> 1; o > 0; o >>= 1) sc += __shfl_xor_sync(ND_FULL, sc, o);
        if (j < L) {
          if (sub == 0) ps[j] = sc;
          m = fmaxf(m, sc);
        }
      }
    }
    m = warp_max(m);
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = expf(ps[j] - m); ps[j] = e; sum += e; }
    sum = warp_sum(sum);
    for (int j = lane; j < L; j += 32) ps[j] = ps[j] / sum;
    __syncwarp();
    // ---- context: each lane accumulates 4 features over its share of the positions
    float4 acc = zero4;
    for (int j0 = 0; j0 < L; j0 += U * PPI) {
      float4 vv[U];
      float pj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * PPI + grp;
        vv[u] = zero4;
        pj[u] = 0.f;
        if (j < L) {
          const float* vr = (j == p.step) ? qkv + 2 * d + h * dh + sub * 4
                            : anc     ? p.Vc + (((int64_t)anc[j] * H + h) * p.Lmax + j) * dh + sub * 4
                                      : p.Vc + own + j * dh;
          vv[u] = *reinterpret_cast<const float4*>(vr);
          pj[u] = ps[j];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x = fmaf(pj[u], vv[u].x, acc.x); acc.y = fmaf(pj[u], vv[u].y, acc.y);
        acc.z = fmaf(pj[u], vv[u].z, acc.z); acc.w = fmaf(pj[u], vv[u].w, acc.w);
      }
    }
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(ND_FULL, acc.x, o); acc.y += __shfl_xor_sync(ND_FULL, acc.y, o);
      acc.z += __shfl_xor_sync(ND_FULL, acc.z, o); acc.w += __shfl_xor_sync(ND_FULL, acc.w, o);
    }
    if (grp == 0) *reinterpret_cast<float4*>(p.ctx + (int64_t)row * d + h * dh + sub * 4) = acc;
    __syncwarp();
  }
}

}  // namespace

cudaError_t self_attention_step(const SelfAttnParams& p, cudaStream_t stream) {
  if (p.rows <= 0) return cudaSuccess;
  const int dh = p.d / p.H;
  if (dh != 8 && dh != 16 && dh != 32 && dh != 64) return cudaErrorInvalidValue;
  const int nw = p.H < 8 ? p.H : 8;
  const size_t smem = ((size_t)p.d + (size_t)nw * p.Lmax) * sizeof(float);
  // U = 8 at head size 64 measured slower (64.7 vs 56.1 us per launch at d = 512, B = 1024, call r02f): 4 everywhere
  switch (dh) {
    case 8: launch_k(self_attn_kernel<4, 8>, dim3(p.rows), dim3(nw * 32), smem, stream, p); break;
    case 16: launch_k(self_attn_kernel<4, 16>, dim3(p.rows), dim3(nw * 32), smem, stream, p); break;
    case 32: launch_k(self_attn_kernel<4, 32>, dim3(p.rows), dim3(nw * 32), smem, stream, p); break;
    default: launch_k(self_attn_kernel<4, 64>, dim3(p.rows), dim3(nw * 32), smem, stream, p); break;
  }
  return cudaGetLastError();
}

// =============================================================================================
// Encoder self attention (flash style, fp32 FFMA): CTA = (chunk, head, 128 queries); thread = query.
// K/V tiles of 64 keys are staged in shared memory and broadcast-read; online softmax per 8 keys.
namespace {

template <int DH>
__global__ void __launch_bounds__(128) enc_attn_kernel(EncAttnParams p) {
  constexpr int KT = 64;
  __shared__ __align__(16) float Ks[KT][DH];
  __shared__ __align__(16) float Vs[KT][DH];
  __shared__ float Ms[KT];
  const int T = p.T, d = p.d;
  const int qtiles = (T + 127) / 128;
  const int qt = blockIdx.x % qtiles;
  const int h = (blockIdx.x / qtiles) % p.H;
  const int b = blockIdx.x / (qtiles * p.H);
  const int tq = qt * 128 + threadIdx.x;
  const bool qok = tq < T;
  const float* base = p.qkv + (int64_t)b * T * 3 * d;

  float q[DH], o[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) { q[e] = qok ? base[(int64_t)tq * 3 * d + h * DH + e] / p.q_div : 0.f; o[e] = 0.f; }
  float m = -FLT_MAX, l = 0.f;

  for (int k0 = 0; k0 < T; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * DH; i += 128) {
      const int j = i / DH, e = i - j * DH;
      const int t = k0 + j;
      Ks[j][e] = t < T ? base[(int64_t)t * 3 * d + d + h * DH + e] : 0.f;
      Vs[j][e] = t < T ? base[(int64_t)t * 3 * d + 2 * d + h * DH + e] : 0.f;
    }
    for (int j = threadIdx.x; j < KT; j += 128) {
      const int t = k0 + j;
      // 0: valid key, 1: masked (src == 0.0), 2: beyond T (does not exist)
      Ms[j] = t < T ? (p.src[(int64_t)b * T + t] == 0.0f ? 1.f : 0.f) : 2.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int j0 = 0; j0 < KT; j0 += 8) {
      float s[8];
      float mt = m;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < DH; ++e) a = fmaf(q[e], Ks[j0 + jj][e], a);
        const float flag = Ms[j0 + jj];
        a = flag == 1.f ? -1e18f : a;              // masked_fill(mask, -1e18)
        s[jj] = a;
        if (flag != 2.f) mt = fmaxf(mt, a);
      }
      const float corr = expf(m - mt);             // m == -FLT_MAX on the first block -> 0
      l *= corr;
#pragma unroll
      for (int e = 0; e < DH; ++e) o[e] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (Ms[j0 + jj] == 2.f) continue;
        const float pj = expf(s[jj] - mt);
        l += pj;
#pragma unroll
        for (int e = 0; e < DH; ++e) o[e] = fmaf(pj, Vs[j0 + jj][e], o[e]);
      }
      m = mt;
    }
  }
  if (qok) {
    float* out = p.ctx + ((int64_t)b * T + tq) * d + h * DH;
#pragma unroll
    for (int e = 0; e < DH; ++e) out[e] = o[e] / l;
  }
}

}  // namespace

cudaError_t encoder_attention(const EncAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  const int dh = p.d / p.H;
  const int64_t grid = (int64_t)p.B * p.H * ((p.T + 127) / 128);
  switch (dh) {
    case 8: enc_attn_kernel<8><<<(unsigned)grid, 128, 0, stream>>>(p); break;
    case 16: enc_attn_kernel<16><<<(unsigned)grid, 128, 0, stream>>>(p); break;
    case 32: enc_attn_kernel<32><<<(unsigned)grid, 128, 0, stream>>>(p); break;
    case 64: enc_attn_kernel<64><<<(unsigned)grid, 128, 0, stream>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// =============================================================================================
// mlp (Bahdanau) / dot global attention, one CTA per chunk; same three phases as cross_attn_kernel
// with a single "head": score[t] = sum_c v[c] * tanh(wq[c] + uh[t][c])   (or q . mem[t]).
namespace {

// tanh(x) = 1 - 2 / (exp(2x) + 1) with the MUFU exponential and reciprocal: absolute error <= ~2e-7 (the terms are summed
// with |v| weights, so absolute error is what matters), 7 instructions instead of tanhf's ~25 — with 8 tanh per row and
// lane the accurate version made the kernel issue-bound (104 us of issue time against 165 us of HBM time per launch).
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));      // exp(2x) = 2^(2x log2 e)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

template <int VPL, int NQMAX>
__global__ void __launch_bounds__(kAttnThreads) mlp_attn_kernel(MlpAttnParams p) {
  extern __shared__ __align__(16) float smem_f[];
  const int chunk = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (p.retired && p.retired[chunk]) return;
  const int d = 32 * VPL, T = p.T, NQ = p.NQ;
  float* q_s = smem_f;                            // [NQ][d]
  float* v_s = q_s + NQ * d;                      // [d]
  float* sc = v_s + d;                            // [NQ][T]  (later red[warps][NQ*d])
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NQ * d; i += kAttnThreads)
    q_s[i] = p.wq[((int64_t)chunk * NQ + i / d) * d + (i % d)];
  for (int i = threadIdx.x; i < d; i += kAttnThreads) v_s[i] = p.dot ? 0.f : p.v[i];
  __syncthreads();
  const int len = p.lengths ? (int)p.lengths[chunk] : T;
  const float* Ub = p.uh + (int64_t)chunk * T * d + lane * VPL;    // dot mode: uh = key matrix
  const float* Mb = p.mem + (int64_t)chunk * T * d + lane * VPL;

  for (int t0 = warp * 4; t0 < T; t0 += kAttnWarps * 4) {
    float u[4][VPL];
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (t0 + r < T) load_slice<VPL>(Ub + (int64_t)(t0 + r) * d, u[r]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int t = t0 + r;
      if (t < T) {
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float* qq = q_s + qi * d + lane * VPL;
            float s = 0.f;
            if (p.dot) {
#pragma unroll
              for (int i = 0; i < VPL; ++i) s = fmaf(qq[i], u[r][i], s);
            } else {
#pragma unroll
              for (int i = 0; i < VPL; ++i) s = fmaf(v_s[lane * VPL + i], tanh_fast(qq[i] + u[r][i]), s);
            }
            s = warp_sum(s);
            if (lane == 0) sc[qi * T + t] = (t < len) ? s : -INFINITY;     // sequence_mask, -inf
          }
        }
      }
    }
  }
  __syncthreads();
  for (int row = warp; row < NQ; row += kAttnWarps) {
    float* s = sc + row * T;
    float m = -FLT_MAX;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
    m = warp_max(m);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s[t] - m); s[t] = e; sum += e; }
    sum = warp_sum(sum);
    for (int t = lane; t < T; t += 32) s[t] = s[t] / sum;
    if (p.attn) {
      float* a = p.attn + ((int64_t)chunk * NQ + row) * T;
      for (int t = lane; t < T; t += 32) a[t] = s[t];
    }
  }
  __syncthreads();
  float acc[NQMAX][VPL];
#pragma unroll
  for (int qi = 0; qi < NQMAX; ++qi)
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[qi][i] = 0.f;
  const int tmax = len < T ? len : T;             // weights beyond the length are exactly 0
  for (int t0 = warp * 4; t0 < tmax; t0 += kAttnWarps * 4) {
    float vv[4][VPL];
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (t0 + r < tmax) load_slice<VPL>(Mb + (int64_t)(t0 + r) * d, vv[r]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int t = t0 + r;
      if (t < tmax) {
#pragma unroll
        for (int qi = 0; qi < NQMAX; ++qi) {
          if (qi < NQ) {
            const float pr = sc[qi * T + t];
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc[qi][i] = fmaf(pr, vv[r][i], acc[qi][i]);
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
  for (int i = threadIdx.x; i < NQ * d; i += kAttnThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) s += red[w * NQ * d + i];
    p.ctx[((int64_t)chunk * NQ + i / d) * p.ctx_ld + (i % d)] = s;
  }
}

template <int VPL>
cudaError_t launch_mlp(const MlpAttnParams& p, cudaStream_t stream) {
  const int d = 32 * VPL;
  const size_t sc_f = (size_t)p.NQ * p.T;
  const size_t red_f = (size_t)kAttnWarps * p.NQ * d;
  const size_t smem = ((size_t)p.NQ * d + d + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  if (p.NQ == 1) {
    cudaFuncSetAttribute(mlp_attn_kernel<VPL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k_heavy(mlp_attn_kernel<VPL, 1>, dim3(p.n_chunks), dim3(kAttnThreads), smem, stream, p);
  } else {
    cudaFuncSetAttribute(mlp_attn_kernel<VPL, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k_heavy(mlp_attn_kernel<VPL, 8>, dim3(p.n_chunks), dim3(kAttnThreads), smem, stream, p);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t mlp_attention(const MlpAttnParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  if (p.kv_fmt != KV_F32) return mlp_attention_packed(p, stream);
  if (p.d % 32 || p.NQ > 8 || p.NQ < 1) return cudaErrorInvalidValue;
  switch (p.d / 32) {
    case 1: return launch_mlp<1>(p, stream);
    case 2: return launch_mlp<2>(p, stream);
    case 4: return launch_mlp<4>(p, stream);
    case 8: return launch_mlp<8>(p, stream);
    case 16: return launch_mlp<16>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nd
