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
//
// FMT (kernels.cuh KV_*): KV_F32 reads fp32 rows; KV_Q23M / KV_Q15M read the fixed-point planes of DESIGN.md 4.5 (a stage
// = 32 rows of the int16 plane + 32 rows of the uint8 plane, two TMA boxes on one mbarrier; 24 / 16 KB instead of 32),
// decode each element once for all NQ queries (PRMT + IADD + FADD, kv_fixed.cuh) and fold the per-row steps into the
// finished score / probability.  2*T*d*(3|2) bytes per chunk instead of 2*T*d*4.
#include <cfloat>

#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "kv_fixed.cuh"

namespace nd {
namespace {

constexpr int kD = 256, kH = 8;
constexpr int kStageWarps = 8;                     // warps that share one stage (4 rows each)
constexpr int kStageRows = 32;
constexpr int kMaxStages = 8;

template <int FMT>
struct RingFmt {
  static constexpr bool packed = FMT != KV_F32;
  static constexpr int hi_row = packed ? kD * 2 : kD * 4;                  // bytes of one row in the stage's first box
  static constexpr int lo_row = (packed && fmt_has_lo(FMT)) ? kD : 0;      // ... second box (uint8 plane)
  static constexpr int lo_off = kStageRows * hi_row;
  static constexpr int stage_bytes = kStageRows * (hi_row + lo_row);
  static constexpr int lane_hi = hi_row / 32, lane_lo = lo_row / 32;        // bytes of a lane's 8 columns
};

// the four rows a warp takes from a stage, as loaded (fp32 words, or int16 pairs + uint8 quads)
template <int FMT>
struct RawRows {
  using F = RingFmt<FMT>;
  uint32_t hi[4][F::lane_hi / 4];
  uint32_t lo[4][F::lane_lo ? F::lane_lo / 4 : 1];
  // row `row` of the stage into slot r
  __device__ __forceinline__ void load(const uint8_t* stage, int row, int lane, int r, bool valid) {
    constexpr int NH = F::lane_hi / 4;
    if (valid) {
      const uint8_t* h = stage + (size_t)row * F::hi_row + F::lane_hi * lane;
#pragma unroll
      for (int i = 0; i < NH; i += 4) {
        const uint4 a = *reinterpret_cast<const uint4*>(h + 4 * i);
        hi[r][i] = a.x; hi[r][i + 1] = a.y; hi[r][i + 2] = a.z; hi[r][i + 3] = a.w;
      }
      if constexpr (F::lane_lo != 0) {
        const uint2 b = *reinterpret_cast<const uint2*>(stage + F::lo_off + (size_t)row * F::lo_row + F::lane_lo * lane);
        lo[r][0] = b.x; lo[r][1] = b.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NH; ++i) hi[r][i] = 0u;
      lo[r][0] = 0u;
      if constexpr (F::lane_lo != 0) lo[r][1] = 0u;
    }
  }
  // elements 2i, 2i+1 of row r
  __device__ __forceinline__ float2 get2(int r, int i) const {
    if constexpr (!F::packed) return make_float2(__uint_as_float(hi[r][2 * i]), __uint_as_float(hi[r][2 * i + 1]));
    else return unpack_pair_magic<FMT>(hi[r], lo[r], 2 * i);
  }
  // one word of every load, for release_stage
  __device__ __forceinline__ uint32_t witness() const {
    uint32_t x = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      constexpr int NH = F::lane_hi / 4;
#pragma unroll
      for (int i = 3; i < NH; i += 4) x ^= hi[r][i];
      if constexpr (F::lane_lo != 0) x ^= lo[r][1];
    }
    return x;
  }
};

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
// 2^x on the MUFU (results below 2^-126 flush to zero: probabilities that small add nothing to an fp32 sum)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Release a ring stage once the rows are in registers.  An arrive issued right behind the LDS instructions does NOT
// wait for their data (a refill served from L2 overtook them: 1e-3 errors with few chunks), and ptxas deletes a dead
// `mov` of the registers, so one word of every load feeds a comparison that guards a (practically never executed) store:
// the warp waits on the scoreboard for all eight loads, then arrives; the arithmetic of the stage overlaps the refill.
__device__ __forceinline__ void release_stage(uint64_t* bar, uint32_t x, int lane, uint32_t* sink) {
  if (x == 0x7fedcba9u) *sink = x;
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// G groups of 8 consumer warps: group g takes the stages whose running number is g mod G (more warps per SM to hide the
// shuffle / LDS latencies of the per-row arithmetic; the ring itself is unchanged)
template <int NQT, int G, int FMT>
__global__ void __launch_bounds__(G * 256 + 32, 1)
cross_attn_ring_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmKlo, const __grid_constant__ CUtensorMap tmVlo,
                       CrossAttnParams p, int n_stages) {
  using F = RingFmt<FMT>;
  constexpr int kStageBytes = F::stage_bytes;
  constexpr bool kExact = NQT == 5;                // the launcher picks NQT = 5 only for NQ = 5: no per-query guards
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
  float* kst_s = reinterpret_cast<float*>(empty + kMaxStages) + 4;        // [TS] key step of the row (1 for fp32 rows), negative = masked key
  float* vst_s = kst_s + TS;                                              // [TS] value step
  float* q_s = vst_s + TS;                                                // [NQ][d]
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
      int s = 0;                                   // stage of the running box number, and the parity of its round
      uint32_t round = 0;
      for (int chunk = blockIdx.x; chunk < p.n_chunks; chunk += gridDim.x) {
        if (p.retired && p.retired[chunk]) continue;
        for (int pass = 0; pass < 2; ++pass) {
          for (int it = 0; it < nit; ++it) {
            mbar_wait(&empty[s], (round & 1) ^ 1);
            mbar_expect_tx(&full[s], kStageBytes);
            tma_load_2d(ring + (size_t)s * kStageBytes, pass ? &tmV : &tmK, &full[s], 0, chunk * T + it * kStageRows);
            if constexpr (F::lo_row != 0)
              tma_load_2d(ring + (size_t)s * kStageBytes + F::lo_off, pass ? &tmVlo : &tmKlo, &full[s], 0,
                          chunk * T + it * kStageRows);
            if (++s == n_stages) { s = 0; ++round; }
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
  // stage / round parity of a running box number c: c % n_stages, (c / n_stages) & 1, kept incrementally (any stage
  // count: the ring takes whatever shared memory the scores leave)
  uint32_t cnt = 0;                                // running stage number at the start of the current pass
  for (int chunk = blockIdx.x; chunk < p.n_chunks; chunk += gridDim.x) {
    if (p.retired && p.retired[chunk]) continue;
    // queries in shared memory as [query][half][lane][4]: lane l reads its columns [8l, 8l+8) with two conflict-free
    // 128-bit loads per query and visit (kept in registers they left the row arithmetic short of registers:
    // 96 per thread with 17 warps)
    for (int i = threadIdx.x; i < NQ * kD; i += kConsThreads) {
      const int qi = i / kD, c = i - qi * kD;
      q_s[qi * kD + ((c >> 2) & 1) * 128 + (c >> 3) * 4 + (c & 3)] = p.q[((int64_t)chunk * NQ + qi) * p.q_ld + c] / p.q_div;
    }
    {
      // key mask of the chunk staged once: a global load per stage sat on the critical path of phase 1 (228 -> 207 us).
      // (Requesting the NEXT chunk's queries and mask during the V pass was tried on top: no gain, 207 -> 213 us.)
      const float* srow = p.src ? p.src + (int64_t)chunk * p.src_ld : nullptr;
      const float* stp = F::packed ? p.kv_scale + (int64_t)chunk * T * 2 : nullptr;
      for (int i = threadIdx.x; i < nit * kStageRows; i += kConsThreads) {
        float ks = 1.f, vs = 0.f;                  // rows >= T of the last stage: probability * 0
        if (i < T) {
          vs = 1.f;
          if constexpr (F::packed) { ks = stp[2 * i]; vs = stp[2 * i + 1]; }
          if (srow && srow[i] == p.mask_value) ks = -ks;
        }
        kst_s[i] = ks;
        vst_s[i] = vs;
      }
    }
    cons_sync();

    // ---------------- phase 1: scores
    {
      const float* q_lane = q_s + 4 * lane;       // + qi * kD (+ 128 for the second half)
      const int q_pitch = kH * TS;
      float* my_sc = sc + head * TS + ws * 4 + j4;  // + qi * q_pitch + it * 32
      const int it0 = (int)((wg + G - cnt % G) % G);
      int s = (int)((cnt + (uint32_t)it0) % (uint32_t)n_stages);
      uint32_t par = ((cnt + (uint32_t)it0) / (uint32_t)n_stages) & 1;
      for (int it = it0; it < nit; it += G) {
        const int t = it * kStageRows + ws * 4 + j4;     // the row whose finished scores this lane stores
        const float ks = kst_s[t];                       // t < nit * 32 always
        const float off = t < T ? -1e18f : -FLT_MAX;     // masked key / row past the chunk (exp -> 0 in phase 2)
        float* dst = my_sc + it * kStageRows;
        mbar_wait(&full[s], par);
        float2 k2[4][4];
        {
          RawRows<FMT> raw;
#pragma unroll
          for (int r = 0; r < 4; ++r)              // slot r holds row r ^ j4: the reduce-scatter below needs no selects
            raw.load(ring + (size_t)s * kStageBytes, ws * 4 + (r ^ j4), lane, r, true);
          release_stage(&empty[s], raw.witness(), lane, sink);
          s += G;                                  // next box of this group
          if (s >= n_stages) { s -= n_stages; par ^= 1; }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) k2[r][i] = raw.get2(r, i);
        }
#pragma unroll
        for (int qi = 0; qi < NQT; ++qi) {
          if (kExact || qi < NQ) {
            const float4 qa = *reinterpret_cast<const float4*>(q_lane + qi * kD);
            const float4 qb = *reinterpret_cast<const float4*>(q_lane + qi * kD + 128);
            const float2 q0 = make_float2(qa.x, qa.y), q1 = make_float2(qa.z, qa.w);
            const float2 q2 = make_float2(qb.x, qb.y), q3 = make_float2(qb.z, qb.w);
            float v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              float2 a2 = __fmul2_rn(q3, k2[r][3]);
              a2 = __ffma2_rn(q2, k2[r][2], a2);
              a2 = __ffma2_rn(q1, k2[r][1], a2);
              a2 = __ffma2_rn(q0, k2[r][0], a2);
              v[r] = a2.x + a2.y;
            }
            // reduce-scatter over the 4 lanes of the head: lane j4 keeps row j4 = its slot 0.  Slot r of lane j4 is row
            // r ^ j4, so slot 2 / 3 of lane j4 ^ 2 are this lane's rows of slot 0 / 1, and slot 1 of lane j4 ^ 1 is row j4.
            const float a0 = v[0] + __shfl_xor_sync(ND_FULL, v[2], 2);
            const float a1 = v[1] + __shfl_xor_sync(ND_FULL, v[3], 2);
            const float w = a0 + __shfl_xor_sync(ND_FULL, a1, 1);
            dst[qi * q_pitch] = ks < 0.f || t >= T ? off : w * ks;
          }
        }
      }
      cnt += (uint32_t)nit;
    }
    cons_sync();

    // ---------------- phase 2: softmax rows, exp(x - max) / sum, then probability * value step.  The memory pipe idles
    // for part of this phase (the ring fills and stalls), so it uses the MUFU exponential (2 ulp) and one reciprocal
    // per row instead of expf / a division per element; the row lives in registers when T <= 512.  Rows [T, nit * 32)
    // hold -FLT_MAX (exp -> 0) and a value step of 0, so nothing here is predicated on t < T except the attn output.
    for (int row = warp; row < NQ * kH; row += kConsWarps) {
      float* srw = sc + row * TS;
      float* a = (p.attn && (row % kH) == 0) ? p.attn + ((int64_t)chunk * NQ + row / kH) * T : nullptr;
      if (nit <= 16) {
        float e[16];
        float m = -FLT_MAX;
#pragma unroll
        for (int i = 0; i < 16; ++i) {             // e[] is written unconditionally (stays in registers)
          float x = -FLT_MAX;
          if (i < nit) x = srw[lane + 32 * i];
          e[i] = x;
          m = fmaxf(m, x);
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          e[i] = ex2_approx((e[i] - m) * 1.4426950408889634f);
          sum += e[i];
        }
        const float inv = 1.0f / warp_sum(sum);
        if (a) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < nit && lane + 32 * i < T) a[lane + 32 * i] = e[i] * inv;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < nit) srw[lane + 32 * i] = (e[i] * inv) * vst_s[lane + 32 * i];
      } else {
        const int Tup = nit * kStageRows;
        float m = -FLT_MAX;
        for (int t = lane; t < Tup; t += 32) m = fmaxf(m, srw[t]);
        m = warp_max(m);
        float sum = 0.f;
        for (int t = lane; t < Tup; t += 32) {
          const float e = ex2_approx((srw[t] - m) * 1.4426950408889634f);
          srw[t] = e;
          sum += e;
        }
        const float inv = 1.0f / warp_sum(sum);
        for (int t = lane; t < Tup; t += 32) {
          const float pr = srw[t] * inv;
          if (a && t < T) a[t] = pr;
          srw[t] = pr * vst_s[t];
        }
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
      const int it0 = (int)((wg + G - cnt % G) % G);
      int s = (int)((cnt + (uint32_t)it0) % (uint32_t)n_stages);
      uint32_t par = ((cnt + (uint32_t)it0) / (uint32_t)n_stages) & 1;
      for (int it = it0; it < nit; it += G) {
        mbar_wait(&full[s], par);
        const int t0 = it * kStageRows + ws * 4;
        float2 v2[4][4];
        {
          RawRows<FMT> raw;
#pragma unroll
          // rows past T belong to the next chunk (or lie past the tensor: zero filled); their probabilities are 0.
          // Fixed-point rows always decode to finite numbers; fp32 rows could be Inf / NaN, so those are zeroed.
          for (int r = 0; r < 4; ++r)
            raw.load(ring + (size_t)s * kStageBytes, ws * 4 + r, lane, r, F::packed || t0 + r < T);
          release_stage(&empty[s], raw.witness(), lane, sink);
          s += G;                                  // next box of this group
          if (s >= n_stages) { s -= n_stages; par ^= 1; }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) v2[r][i] = raw.get2(r, i);
        }
#pragma unroll
        for (int qi = 0; qi < NQT; ++qi) {
          if (kExact || qi < NQ) {
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

int stage_bytes_of(int fmt) {
  return fmt == KV_F32 ? RingFmt<KV_F32>::stage_bytes : fmt == KV_Q23M ? RingFmt<KV_Q23M>::stage_bytes : RingFmt<KV_Q15M>::stage_bytes;
}

RingPlan plan(const CrossAttnParams& p, int groups) {
  RingPlan r;
  const int kConsWarps = groups * kStageWarps;
  const size_t stage = (size_t)stage_bytes_of(p.kv_fmt);
  const int nit = (p.T + kStageRows - 1) / kStageRows;
  const size_t TS = (size_t)nit * kStageRows + 4;
  const size_t sc_f = (size_t)p.NQ * kH * TS, red_f = (size_t)kConsWarps * p.NQ * kD;
  const size_t fixed = 2 * kMaxStages * sizeof(uint64_t) + 16 + 2 * TS * sizeof(float) +
                       ((size_t)p.NQ * kD + (sc_f > red_f ? sc_f : red_f)) * sizeof(float);
  const size_t budget = 227 * 1024 - 128;
  if (fixed + 2 * stage > budget) return r;
  const int st = (int)((budget - fixed) / stage);
  r.stages = st > kMaxStages ? kMaxStages : st;
  // With two consumer groups the stage count must be even: group g takes the boxes c = g (mod 2), so with an even count
  // a stage is always consumed by the same group, whose warps meet its boxes in order.  With an odd count the groups
  // alternate on a stage, a warp can reach box c + n of a stage before box c has landed, and the parity wait on the
  // stage's `full` barrier passes on the phase of box c - n: it reads a half-filled stage and releases it twice.
  if (groups == 2) r.stages &= ~1;
  r.smem = (size_t)r.stages * stage + fixed;
  return r;
}

int g_sm_count = 0;

template <int NQT, int G, int FMT>
cudaError_t launch_ring_fmt(const CrossAttnParams& p, const RingPlan& pl, cudaStream_t stream) {
  using F = RingFmt<FMT>;
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cross_attn_ring_kernel<NQT, G, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = p.n_chunks < g_sm_count ? p.n_chunks : g_sm_count;
  CUtensorMap tmK, tmV, tmKlo, tmVlo;
  const int64_t rows = (int64_t)p.n_chunks * p.T;
  if constexpr (!F::packed) {
    if (!make_plain_map(&tmK, p.K, rows, kD, p.kv_ld, kStageRows, kD)) return cudaErrorInvalidValue;
    if (!make_plain_map(&tmV, p.V, rows, kD, p.kv_ld, kStageRows, kD)) return cudaErrorInvalidValue;
    tmKlo = tmK;
    tmVlo = tmV;
  } else {
    // the planes as fp32 words (TMA only moves bytes): a row of the int16 plane is [K: d/2 words | V: d/2 words], a row
    // of the uint8 plane [K: d/4 | V: d/4]
    const float* hi = reinterpret_cast<const float*>(p.kv_hi);
    if (!make_plain_map(&tmK, hi, rows, kD / 2, kD, kStageRows, kD / 2)) return cudaErrorInvalidValue;
    if (!make_plain_map(&tmV, hi + kD / 2, rows, kD / 2, kD, kStageRows, kD / 2)) return cudaErrorInvalidValue;
    tmKlo = tmK;
    tmVlo = tmV;
    if constexpr (F::lo_row != 0) {
      const float* lo = reinterpret_cast<const float*>(p.kv_lo);
      if (!make_plain_map(&tmKlo, lo, rows, kD / 4, kD / 2, kStageRows, kD / 4)) return cudaErrorInvalidValue;
      if (!make_plain_map(&tmVlo, lo + kD / 4, rows, kD / 4, kD / 2, kStageRows, kD / 4)) return cudaErrorInvalidValue;
    }
  }
  launch_k_heavy(cross_attn_ring_kernel<NQT, G, FMT>, dim3(grid), dim3(G * 256 + 32), pl.smem, stream, tmK, tmV, tmKlo,
                 tmVlo, p, pl.stages);
  return cudaGetLastError();
}

template <int NQT, int G>
cudaError_t launch_ring(const CrossAttnParams& p, const RingPlan& pl, cudaStream_t stream) {
  if (p.kv_fmt == KV_Q23M) return launch_ring_fmt<NQT, G, KV_Q23M>(p, pl, stream);
  if (p.kv_fmt == KV_Q15M) return launch_ring_fmt<NQT, G, KV_Q15M>(p, pl, stream);
  return launch_ring_fmt<NQT, G, KV_F32>(p, pl, stream);
}

}  // namespace

bool cross_attention_ring_shape_ok(int NQ, int d, int H, int T, int fmt) {
  if (NQ < 2 || NQ > 8 || d != kD || H != kH || T < 1) return false;
  if (fmt != KV_F32 && fmt != KV_Q23M && fmt != KV_Q15M) return false;
  CrossAttnParams q;
  q.NQ = NQ; q.d = d; q.H = H; q.T = T; q.kv_fmt = fmt;
  return plan(q, 1).stages >= 2;
}

bool cross_attention_ring_supported(const CrossAttnParams& p) {
  if (!cross_attention_ring_shape_ok(p.NQ, p.d, p.H, p.T, p.kv_fmt)) return false;
  if (p.kv_fmt == KV_F32) {
    if ((p.kv_ld & 3) || (reinterpret_cast<uintptr_t>(p.K) & 15) || (reinterpret_cast<uintptr_t>(p.V) & 15)) return false;
  } else {
    if (!p.kv_hi || !p.kv_scale || (reinterpret_cast<uintptr_t>(p.kv_hi) & 15)) return false;
    if (p.kv_fmt == KV_Q23M && (!p.kv_lo || (reinterpret_cast<uintptr_t>(p.kv_lo) & 15))) return false;
  }
  if ((int64_t)p.n_chunks * p.T > 0x7fffffffLL) return false;           // tensor-map row coordinate
  return true;
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
