// Beam-search cross attention (several queries per chunk), d = 256, H = 8: persistent CTAs fed by a shared-memory ring.
//
// Reference rows: onmt/modules/multi_headed_attn.py:142-199 (context attention of the Transformer decoder) with the
// K beams of a chunk attending the SAME keys / values (translator.py:619-825 tiles the memory bank beam_size times;
// here the beams share one pass over it).
//
// Why a second kernel: with 5-8 queries the per-row arithmetic needs ~130 registers, two CTAs per SM hold only 16
// warps x 4-8 KB of loads in flight, and the three phases (scores / softmax / context) drain the memory pipe twice per
// chunk: the register-prefetch kernel in attention.cu reaches 4.0 TB/s (62 % of the measured HBM peak).  Here
//   * one CTA per SM walks the chunk list (no wave quantisation: 1024 chunks on 148 SMs),
//   * a producer thread streams K rows, then V rows, of chunk after chunk into a ring of 32-row stages with TMA
//     (cp.async.bulk.tensor.2d, one 32 x 1 KB box per stage, completion counted in bytes on the stage's mbarrier); it runs ahead of the
//     consumer warps (two groups of 8 taking alternate stages) through the softmax and the final reduction, so
//     64-128 KB per SM stay in flight at all times,
//   * consumers keep the queries in registers and use packed fp32 FMAs (FFMA2: each half is an IEEE fp32 fma); lane l
//     owns the 8 contiguous columns [8l, 8l+8) = a quarter of head l/4,
//   * per query the 4 row partial sums of a 4-row block are reduce-scattered over the 4 lanes of a head with 3 shuffles;
//     lane j stores the finished score of row j & 3.
// Algorithmic bytes per launch: n_chunks * (2*T*d*4 + NQ*2*d*4) — K and V are read exactly once.
#include <cfloat>

#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace nd {
namespace {

constexpr int kD = 256, kH = 8;
constexpr int kStageWarps = 8;                     // warps that share one stage (4 rows each)
constexpr int kStageRows = 32;
constexpr int kRowBytes = kD * 4;
constexpr int kStageBytes = kStageRows * kRowBytes;
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Release a ring stage once the rows are in registers.  An arrive issued right behind the LDS instructions does NOT
// wait for their data (a refill served from L2 overtook them: 1e-3 errors with few chunks), and ptxas deletes a dead
// `mov` of the registers, so one word of every load feeds a comparison that guards a (practically never executed) store:
// the warp waits on the scoreboard for all eight loads, then arrives; the arithmetic of the stage overlaps the refill.
__device__ __forceinline__ void release_stage(uint64_t* bar, const float2 (&rows)[4][4], int lane, uint32_t* sink) {
  uint32_t x = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) x ^= __float_as_uint(rows[r][1].y) ^ __float_as_uint(rows[r][3].y);
  if (x == 0x7fedcba9u) *sink = x;
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// G groups of 8 consumer warps: group g takes the stages whose running number is g mod G (more warps per SM to hide the
// shuffle / LDS latencies of the per-row arithmetic; the ring itself is unchanged)
template <int NQT, int G>
__global__ void __launch_bounds__(G * 256 + 32, 1)
cross_attn_ring_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, CrossAttnParams p,
                       int n_stages) {
  constexpr int kConsWarps = G * kStageWarps, kConsThreads = kConsWarps * 32;
  auto cons_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); };
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  const int T = p.T, NQ = p.NQ;
  const int nit = (T + kStageRows - 1) / kStageRows;           // stages per K (or V) pass of one chunk
  const int TS = nit * kStageRows + 4;                         // score row pitch (16-byte multiple, != 0 mod 32 banks)
  uint8_t* ring = smem_raw;                                    // [n_stages][32 rows][1 KB]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)n_stages * kStageBytes);
  uint64_t* empty = full + kMaxStages;
  uint32_t* sink = reinterpret_cast<uint32_t*>(empty + kMaxStages);       // see release_stage (16 bytes reserved)
  float* msk_s = reinterpret_cast<float*>(empty + kMaxStages) + 4;        // [TS] 1.0 = masked key of this chunk
  float* q_s = msk_s + TS;                                                // [NQ][d]
  float* sc = q_s + NQ * kD;                                   // [NQ*H][TS], later red[8 warps][NQ*d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kStageWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kConsWarps) {
    // ===================================================================== producer
    // one 32-row x 1 KB box per stage (a per-row cp.async.bulk loop costs ~50 cycles of serial uniform-datapath work
    // per row and caps the SM at half its share of the HBM bandwidth).  The box of a chunk's last stage may run into
    // the next chunk's rows (or past the tensor: zero filled); consumers never use rows >= T.
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
      uint32_t cnt = 0;
      for (int chunk = blockIdx.x; chunk < p.n_chunks; chunk += gridDim.x) {
        if (p.retired && p.retired[chunk]) continue;
        for (int pass = 0; pass < 2; ++pass) {
          for (int it = 0; it < nit; ++it, ++cnt) {
            const int s = (int)(cnt % (uint32_t)n_stages);
            const uint32_t round = cnt / (uint32_t)n_stages;
            mbar_wait(&empty[s], (round & 1) ^ 1);
            mbar_expect_tx(&full[s], kStageBytes);
            tma_load_2d(ring + (size_t)s * kStageBytes, pass ? &tmV : &tmK, &full[s], 0, chunk * T + it * kStageRows);
          }
        }
      }
    }
    return;
  }

  // ======================================================================= consumers (G x 8 warps)
  // lane l owns the 8 contiguous columns [8l, 8l+8) = a quarter of head l/4 (two LDS.128 per row; the 32-byte lane
  // pitch makes them 2-way bank conflicted, which the shared-memory pipe has room for)
  const int wg = warp / kStageWarps, ws = warp % kStageWarps;   // stage group, warp slot inside a stage
  const int head = lane >> 2, j4 = lane & 3;
  const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
  const uint32_t smask = (uint32_t)n_stages - 1;   // n_stages is a power of two
  const int sshift = n_stages == 4 ? 2 : 1;
  uint32_t cnt = 0;                                // running stage number at the start of the current pass
  for (int chunk = blockIdx.x; chunk < p.n_chunks; chunk += gridDim.x) {
    if (p.retired && p.retired[chunk]) continue;
    for (int i = threadIdx.x; i < NQ * kD; i += kConsThreads) {
      const int qi = i / kD, c = i - qi * kD;
      q_s[i] = p.q[((int64_t)chunk * NQ + qi) * p.q_ld + c] / p.q_div;
    }
    {
      // key mask of the chunk staged once: a global load per stage sat on the critical path of phase 1 (228 -> 207 us).
      // (Requesting the NEXT chunk's queries and mask during the V pass was tried on top: no gain, 207 -> 213 us.)
      const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
      for (int i = threadIdx.x; i < T; i += kConsThreads) msk_s[i] = (srow && srow[i] == p.mask_value) ? 1.f : 0.f;
    }
    cons_sync();

    // ---------------- phase 1: scores
    {
      float2 q2[NQT][4];
#pragma unroll
      for (int qi = 0; qi < NQT; ++qi) {
        const int qq = qi < NQ ? qi : 0;
        const float4 a = *reinterpret_cast<const float4*>(q_s + qq * kD + 8 * lane);
        const float4 b = *reinterpret_cast<const float4*>(q_s + qq * kD + 8 * lane + 4);
        q2[qi][0] = make_float2(a.x, a.y); q2[qi][1] = make_float2(a.z, a.w);
        q2[qi][2] = make_float2(b.x, b.y); q2[qi][3] = make_float2(b.z, b.w);
      }
      float* my_sc = sc + head * TS + ws * 4 + j4;  // + qi * kH * TS + it * 32
      for (int it = (int)((wg + G - cnt % G) % G); it < nit; it += G) {
        const uint32_t c = cnt + (uint32_t)it;
        const int s = (int)(c & smask);
        const int t = it * kStageRows + ws * 4 + j4;     // the row whose finished scores this lane stores
        const bool masked = msk_s[t] != 0.f;             // t < TS always; rows >= T are never stored
        mbar_wait(&full[s], (c >> sshift) & 1);
        const uint8_t* rows = ring + (size_t)s * kStageBytes + (size_t)(ws * 4) * kRowBytes + 32 * lane;
        float2 k2[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(rows + r * kRowBytes);
          const float4 b = *reinterpret_cast<const float4*>(rows + r * kRowBytes + 16);
          k2[r][0] = make_float2(a.x, a.y); k2[r][1] = make_float2(a.z, a.w);
          k2[r][2] = make_float2(b.x, b.y); k2[r][3] = make_float2(b.z, b.w);
        }
        release_stage(&empty[s], k2, lane, sink);
#pragma unroll
        for (int qi = 0; qi < NQT; ++qi) {
          if (qi < NQ) {
            float v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              float2 a2 = __fmul2_rn(q2[qi][3], k2[r][3]);
              a2 = __ffma2_rn(q2[qi][2], k2[r][2], a2);
              a2 = __ffma2_rn(q2[qi][1], k2[r][1], a2);
              a2 = __ffma2_rn(q2[qi][0], k2[r][0], a2);
              v[r] = a2.x + a2.y;
            }
            // reduce-scatter over the 4 lanes of the head: lane j4 keeps row j4
            const float a0 = (up2 ? v[2] : v[0]) + __shfl_xor_sync(ND_FULL, up2 ? v[0] : v[2], 2);
            const float a1 = (up2 ? v[3] : v[1]) + __shfl_xor_sync(ND_FULL, up2 ? v[1] : v[3], 2);
            const float w = (up1 ? a1 : a0) + __shfl_xor_sync(ND_FULL, up1 ? a0 : a1, 1);
            if (t < T) my_sc[qi * kH * TS + it * kStageRows] = masked ? -1e18f : w;
          }
        }
      }
      cnt += (uint32_t)nit;
    }
    cons_sync();

    // ---------------- phase 2: softmax rows, exp(x - max) / sum; zero the tail of each row.  The memory pipe idles
    // for part of this phase (the ring fills and stalls), so it uses the MUFU exponential (2 ulp) and one reciprocal
    // per row instead of expf / a division per element; the row lives in registers when T <= 512.
    for (int row = warp; row < NQ * kH; row += kConsWarps) {
      float* srw = sc + row * TS;
      if (T <= 512) {
        float e[16];
        float m = -FLT_MAX;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int t = lane + 32 * i;
          e[i] = t < T ? srw[t] : -FLT_MAX;
          m = fmaxf(m, e[i]);
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          e[i] = lane + 32 * i < T ? __expf(e[i] - m) : 0.f;
          sum += e[i];
        }
        const float inv = 1.0f / warp_sum(sum);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int t = lane + 32 * i;
          if (t < T) srw[t] = e[i] * inv;
        }
      } else {
        float m = -FLT_MAX;
        for (int t = lane; t < T; t += 32) m = fmaxf(m, srw[t]);
        m = warp_max(m);
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) { const float e = __expf(srw[t] - m); srw[t] = e; sum += e; }
        const float inv = 1.0f / warp_sum(sum);
        for (int t = lane; t < T; t += 32) srw[t] = srw[t] * inv;
      }
      __syncwarp();
      for (int t = T + lane; t < TS; t += 32) srw[t] = 0.f;
      if (p.attn && (row % kH) == 0) {
        float* a = p.attn + ((int64_t)chunk * NQ + row / kH) * T;
        for (int t = lane; t < T; t += 32) a[t] = srw[t];
      }
    }
    cons_sync();

    // ---------------- phase 3: context
    float2 acc2[NQT][4];
#pragma unroll
    for (int qi = 0; qi < NQT; ++qi)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc2[qi][i] = make_float2(0.f, 0.f);
    {
      const float* my_p = sc + head * TS + ws * 4;  // + qi * kH * TS + it * 32
      for (int it = (int)((wg + G - cnt % G) % G); it < nit; it += G) {
        const uint32_t c = cnt + (uint32_t)it;
        const int s = (int)(c & smask);
        mbar_wait(&full[s], (c >> sshift) & 1);
        const uint8_t* rows = ring + (size_t)s * kStageBytes + (size_t)(ws * 4) * kRowBytes + 32 * lane;
        const int t0 = it * kStageRows + ws * 4;
        float2 v2[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (t0 + r < T) {                        // rows past T belong to the next chunk (or are stale)
            a = *reinterpret_cast<const float4*>(rows + r * kRowBytes);
            b = *reinterpret_cast<const float4*>(rows + r * kRowBytes + 16);
          }
          v2[r][0] = make_float2(a.x, a.y); v2[r][1] = make_float2(a.z, a.w);
          v2[r][2] = make_float2(b.x, b.y); v2[r][3] = make_float2(b.z, b.w);
        }
        release_stage(&empty[s], v2, lane, sink);
#pragma unroll
        for (int qi = 0; qi < NQT; ++qi) {
          if (qi < NQ) {
            const float4 pr = *reinterpret_cast<const float4*>(my_p + qi * kH * TS + it * kStageRows);
            const float pv[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float2 pp = make_float2(pv[r], pv[r]);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc2[qi][i] = __ffma2_rn(pp, v2[r][i], acc2[qi][i]);
            }
          }
        }
      }
      cnt += (uint32_t)nit;
    }
    cons_sync();                                   // probabilities no longer needed: reuse as the reduction buffer
    float* red = sc;                               // [warps][NQ*d]
#pragma unroll
    for (int qi = 0; qi < NQT; ++qi) {
      if (qi < NQ) {
        float* r0 = red + (warp * NQ + qi) * kD + 8 * lane;
        *reinterpret_cast<float4*>(r0) = make_float4(acc2[qi][0].x, acc2[qi][0].y, acc2[qi][1].x, acc2[qi][1].y);
        *reinterpret_cast<float4*>(r0 + 4) = make_float4(acc2[qi][2].x, acc2[qi][2].y, acc2[qi][3].x, acc2[qi][3].y);
      }
    }
    cons_sync();
    for (int i = threadIdx.x; i < NQ * kD; i += kConsThreads) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kConsWarps; ++w) sum += red[w * NQ * kD + i];
      const int qi = i / kD, c = i - qi * kD;
      p.ctx[((int64_t)chunk * NQ + qi) * p.ctx_ld + c] = sum;
    }
    cons_sync();                                   // red / q_s are rewritten by the next chunk
  }
}

struct RingPlan {
  int stages = 0;
  size_t smem = 0;
};

RingPlan plan(const CrossAttnParams& p, int groups) {
  RingPlan r;
  const int kConsWarps = groups * kStageWarps;
  const int nit = (p.T + kStageRows - 1) / kStageRows;
  const size_t TS = (size_t)nit * kStageRows + 4;
  const size_t sc_f = (size_t)p.NQ * kH * TS, red_f = (size_t)kConsWarps * p.NQ * kD;
  const size_t fixed = 2 * kMaxStages * sizeof(uint64_t) + 16 + TS * sizeof(float) + ((size_t)p.NQ * kD + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  const size_t budget = 227 * 1024 - 128;
  if (fixed + 2 * (size_t)kStageBytes > budget) return r;
  const int st = (int)((budget - fixed) / kStageBytes);
  r.stages = st >= 4 ? 4 : 2;                      // power of two: stage index and phase by mask / shift
  r.smem = (size_t)r.stages * kStageBytes + fixed;
  return r;
}

int g_sm_count = 0;

template <int NQT, int G>
cudaError_t launch_ring(const CrossAttnParams& p, const RingPlan& pl, cudaStream_t stream) {
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cross_attn_ring_kernel<NQT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = p.n_chunks < g_sm_count ? p.n_chunks : g_sm_count;
  CUtensorMap tmK, tmV;
  const int64_t rows = (int64_t)p.n_chunks * p.T;
  if (!make_plain_map(&tmK, p.K, rows, kD, p.kv_ld, kStageRows, kD)) return cudaErrorInvalidValue;
  if (!make_plain_map(&tmV, p.V, rows, kD, p.kv_ld, kStageRows, kD)) return cudaErrorInvalidValue;
  launch_k_heavy(cross_attn_ring_kernel<NQT, G>, dim3(grid), dim3(G * 256 + 32), pl.smem, stream, tmK, tmV, p, pl.stages);
  return cudaGetLastError();
}

}  // namespace

bool cross_attention_ring_supported(const CrossAttnParams& p) {
  if (p.NQ < 2 || p.NQ > 8 || p.d != kD || p.H != kH || p.T < 1) return false;
  if ((p.kv_ld & 3) || (reinterpret_cast<uintptr_t>(p.K) & 15) || (reinterpret_cast<uintptr_t>(p.V) & 15)) return false;
  if ((int64_t)p.n_chunks * p.T > 0x7fffffffLL) return false;           // tensor-map row coordinate
  return plan(p, 1).stages >= 2;
}

int g_ring_groups = 2;
void cross_attention_ring_set_groups(int g) { g_ring_groups = g; }

cudaError_t cross_attention_ring(const CrossAttnParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  if (!cross_attention_ring_supported(p)) return cudaErrorInvalidValue;
  // two groups of 8 warps while the queries fit 120 registers per thread and the wider reduction buffer fits
  const RingPlan pl2 = plan(p, 2);
  if (g_ring_groups == 2 && p.NQ <= 5 && pl2.stages >= 2)
    return p.NQ <= 4 ? launch_ring<4, 2>(p, pl2, stream) : launch_ring<5, 2>(p, pl2, stream);
  const RingPlan pl = plan(p, 1);
  if (p.NQ <= 4) return launch_ring<4, 1>(p, pl, stream);
  if (p.NQ == 5) return launch_ring<5, 1>(p, pl, stream);
  return launch_ring<8, 1>(p, pl, stream);
}

}  // namespace nd
