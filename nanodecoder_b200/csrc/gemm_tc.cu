// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epi( pro(A)[M,K] . W[N,K]^T )   (fp32 in/out)
//
// Arithmetic: kind::tf32 UMMA (M=128, N=BN, K=8 per instruction), fp32 accumulation in TMEM.
//   npass == 3  "3xTF32": a = a_hi + a_lo, w = w_hi + w_lo (tf32-exact parts);
//               D += a_lo.w_hi + a_hi.w_lo + a_hi.w_hi   -> ~2^-21 relative error (fp32 parity mode)
//   npass == 1  single TF32 pass (fast mode, ~2^-11)
//
// CTA = 320 threads, one 128 x BN output tile:
//   warp 8     TMA producer: A (raw fp32), W_hi, W_lo tiles of 128B-swizzled K-major rows (BK = 32)
//   warps 0-7  converters: split the A tile in place into tf32 hi/lo (and accumulate the LayerNorm row
//              moments), publish to the async proxy; afterwards the epilogue (tcgen05.ld -> bias /
//              folded LayerNorm / ReLU -> transpose through shared memory -> residual -> coalesced stores)
//   warp 9     TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit to mbarriers
// Pipeline: kStages smem stages, mbarriers raw_full (TMA tx) -> conv_full (128 arrivals) -> empty (commit).
//
// Split-K over a thread-block cluster (template S > 1), for the decode-step projections where
// M = batch rows is small and a 128 x BN tile grid cannot fill 148 SMs: the S CTAs of a cluster
// each run the pipeline over 1/S of K into their own TMEM accumulator, park the partial tile in
// shared memory, and after one cluster barrier every CTA reduces 128/S rows over the S partials
// through distributed shared memory in a FIXED order (deterministic) and applies the epilogue.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cudaTypedefs.h>

#include "gemm.cuh"

namespace nd {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = one 128-byte swizzle row
constexpr int kConvWarps = 8;
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kTmaWarp = kConvWarps, kMmaWarp = kConvWarps + 1;
constexpr int kThreads = kConvThreads + 64;
constexpr int A_TILE_BYTES = BM * 128;

// ATM (3xTF32, Q == 1): tf32 hi / lo parts of the A tile go to tensor memory (64 columns per stage) instead of back
// to shared memory; see gemm_tc_persist_kernel.
template <int BN, int NPASS, int S, int Q = 1, bool ATM = false>
struct Cfg {
  static_assert(!ATM || (NPASS == 3 && Q == 1), "A in tensor memory: 3xTF32, one accumulator");
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int A_BYTES = A_TILE_BYTES * ((NPASS == 3 && !ATM) ? 2 : 1);
  static constexpr int STAGE_BYTES = A_BYTES + B_TILE_BYTES * (NPASS == 3 ? 2 : 1);
  static constexpr int PS = BN + 4;                                   // padded pitch of split-K partial rows
  static constexpr int RECV_BYTES = S > 1 ? BM * PS * 4 : 0;          // peers push their partial rows here
  static constexpr int kStagesFit = (200 * 1024 - RECV_BYTES) / STAGE_BYTES;
  static constexpr int kStages = kStagesFit >= 6 ? 6 : kStagesFit;       // 6 x 64 A columns + BN <= 512 for ATM
  static constexpr int AUX_BYTES = 256 /*barriers*/ + 2 * BM * 4 /*row stats*/ + 2 * BN * 4 /*epilogue vectors*/ +
                                   Q * 6 * BM * 4 /*half-row moments (per serial K slice)*/;
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + RECV_BYTES + 1024 /*align*/ + AUX_BYTES;
  static constexpr int ACC_COLS = Q * BN;                            // one accumulator per serial K slice
  static constexpr int NEED_COLS = ACC_COLS + (ATM ? kStages * 64 : 0);
  static constexpr int TMEM_COLS = NEED_COLS <= 32 ? 32 : (NEED_COLS <= 64 ? 64 : (NEED_COLS <= 128 ? 128 : (NEED_COLS <= 256 ? 256 : 512)));
  static_assert(Q == 1 || S == 1, "serial K slices replace the cluster split");
  static_assert(NEED_COLS <= 512, "tensor memory");
};

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (ignored for swizzled K-major; 1), [32,46) SBO >> 4
//   (8 rows x 128 B = 1024 B between 8-row groups), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column per k element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------ epilogue
// bias / folded LayerNorm, ReLU, residual and store for CH consecutive columns (kept small on purpose:
// these kernels run for ~10 us, cold instruction fetch of a bulky epilogue costs as much as the math)
// [nb, nb+CH) of output row m; v[] holds the fp32 accumulators.  All register indices static.
// cold path: ragged N tail or unaligned C / residual — out of line so it costs no instruction fetch
template <int CH>
__device__ __noinline__ void epilogue_store_slow(const GemmParams& p, float (&v)[CH], int m, int nb, bool fold,
                                                 float ln_mean, float ln_rstd) {
  float* crow = p.C + (int64_t)m * p.ldc + nb;
  const float* rrow = p.residual ? p.residual + (int64_t)m * p.ldr + nb : nullptr;
#pragma unroll 1
  for (int j = 0; j < CH; ++j) {
    const int n = nb + j;
    if (n < p.N) {
      float x = v[j];
      if (fold) x = ln_rstd * (x - ln_mean * __ldg(p.ln_cvec + n)) + __ldg(p.ln_dvec + n);
      else if (p.bias) x += __ldg(p.bias + n);
      if (p.relu == 1) x = fmaxf(x, 0.f);
      if (rrow) x += rrow[j];
      if (p.relu == 3) x = fmaxf(x, 0.f);
      crow[j] = x;
    }
  }
}

// row-wise part of the epilogue on CH accumulators of one output row: bias or folded LayerNorm, ReLU.
// vec0 / vec1: the tile's epilogue vectors staged in shared memory, indexed by tile column `col`.
template <int CH>
__device__ __forceinline__ void epilogue_rowwise(const GemmParams& p, float (&v)[CH], int col, bool fold,
                                                 float ln_mean, float ln_rstd, const float* vec0, const float* vec1) {
  if (fold) {
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      const float4 cv = *reinterpret_cast<const float4*>(vec0 + col + j);
      const float4 dv = *reinterpret_cast<const float4*>(vec1 + col + j);
      v[j] = ln_rstd * (v[j] - ln_mean * cv.x) + dv.x;
      v[j + 1] = ln_rstd * (v[j + 1] - ln_mean * cv.y) + dv.y;
      v[j + 2] = ln_rstd * (v[j + 2] - ln_mean * cv.z) + dv.z;
      v[j + 3] = ln_rstd * (v[j + 3] - ln_mean * cv.w) + dv.w;
    }
  } else if (p.bias) {
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(vec0 + col + j);
      v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
    }
  }
  if (p.relu == 1) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

#define ND_TS(slot) do { if (p.dbg && blockIdx.x == 0) p.dbg[slot] = clock64(); } while (0)

// ------------------------------------------------------------------------------------ kernel
// warps 0-7: converters + epilogue, warp 8: TMA producer, warp 9: TMEM owner + MMA issuer
// Q > 1 ("serial split", S == 1): ONE CTA walks the whole K range but keeps the Q K-slices of the cluster split in Q
// separate accumulators (and Q separate sets of LayerNorm moments) and adds them in the cluster kernel's order, so
// the result is bit-identical to gemm_tc_kernel<*, NPASS, Q> whatever the tile width.  Used when there are enough
// rows (beam search: beam x batch) that the cluster split would run several waves of short CTAs.
template <int BN, int NPASS, int S, int Q = 1, bool ATM = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
               const __grid_constant__ CUtensorMap tmWlo, GemmParams p) {
  using C = Cfg<BN, NPASS, S, Q, ATM>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by SWIZZLE_128B operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* part = reinterpret_cast<float*>(smem + C::kStages * C::STAGE_BYTES);   // split-K receive buffer [S][BM/S][PS]
  uint8_t* aux = smem + C::kStages * C::STAGE_BYTES + C::RECV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* raw_full = bars;
  uint64_t* conv_full = bars + C::kStages;
  uint64_t* empty = bars + 2 * C::kStages;
  uint64_t* tmem_full = bars + 3 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 1);
  float* s_stats = reinterpret_cast<float*>(aux + 256);              // [BM][2] (mean, M2) per (source rank, row)
  float* s_vec0 = s_stats + 2 * BM;                                  // [BN] bias | folded-LN cvec of this tile
  float* s_vec1 = s_vec0 + BN;                                       // [BN] folded-LN dvec
  float* s_mom = s_vec1 + BN;                                        // [2][BM][3] half-row (n, mean, M2)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) ND_TS(0);
  int crank = 0;
  if constexpr (S > 1) crank = (int)cooperative_groups::this_cluster().block_rank();
  const int tile = blockIdx.x / S;
  const int ntn = (p.N + BN - 1) / BN;
  const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN;
  const int KBtot = (p.K + BK - 1) / BK;
  const int kb_per = KBtot / S;                       // launcher guarantees KBtot % S == 0
  const int kb0 = crank * kb_per;
  const int KB = kb_per;
  const int kb_q = KB / Q;                            // k-blocks per serial slice (launcher: KBtot % Q == 0)
  constexpr int PS = C::PS;

  auto a_hi = [&](int s) { return tiles + s * C::STAGE_BYTES; };
  auto a_lo = [&](int s) { return tiles + s * C::STAGE_BYTES + A_TILE_BYTES; };
  auto b_hi = [&](int s) { return tiles + s * C::STAGE_BYTES + C::A_BYTES; };
  auto b_lo = [&](int s) { return b_hi(s) + C::B_TILE_BYTES; };
  auto issue_w = [&](int i) {                  // weight tiles: never written by another kernel of the stream
    const int s = i % C::kStages;
    mbar_expect_tx(&raw_full[s], A_TILE_BYTES + C::B_TILE_BYTES * (NPASS == 3 ? 2 : 1));
    const int kc = (kb0 + i) * BK;
    tma_load_2d(b_hi(s), &tmWhi, &raw_full[s], kc, n0);
    if (NPASS == 3) tma_load_2d(b_lo(s), &tmWlo, &raw_full[s], kc, n0);
  };
  auto issue_a = [&](int i) {                  // activation tile: produced by the previous kernel
    const int s = i % C::kStages;
    tma_load_2d(a_hi(s), &tmA, &raw_full[s], (kb0 + i) * BK, m0);
  };
  auto issue_tma = [&](int i) { issue_w(i); issue_a(i); };
  pdl_launch_dependents();                     // the next kernel's prologue may overlap this kernel
  if (p.alive) {                               // uniform over the grid: nothing has been set up yet
    pdl_wait();
    if (*p.alive == 0) return;
  }

  int tma_issued = 0;
  if (warp == kTmaWarp && lane == 0) {
    // the producer owns the barriers: initialise them and put the first stages in flight before
    // the CTA-wide setup barrier, so the first TMA round trip overlaps TMEM allocation and setup
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], kConvThreads);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    if (NPASS == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
    // programmatic dependent launch: the weight tiles of the first stages are requested while the previous
    // kernel may still be running; everything that kernel produced (A, residual, C) is touched after pdl_wait
    const int first = KB < C::kStages ? KB : C::kStages;
    for (int i = 0; i < first; ++i) issue_w(i);
    pdl_wait();
    for (; tma_issued < first; ++tma_issued) issue_a(tma_issued);
    ND_TS(2);
  } else {
    pdl_wait();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x < BN) {                      // epilogue vectors of this tile's columns -> shared memory
    const int n = n0 + (int)threadIdx.x;
    const bool f = p.ln_cvec != nullptr;
    s_vec0[threadIdx.x] = n < p.N ? (f ? p.ln_cvec[n] : (p.bias ? p.bias[n] : 0.f)) : 0.f;
    s_vec1[threadIdx.x] = (f && n < p.N) ? p.ln_dvec[n] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) ND_TS(1);

  const bool fold = p.ln_cvec != nullptr;
  const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                      (!p.residual || (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)));

  if (warp == kTmaWarp) {
    // ===================================================================== TMA producer (remaining stages)
    if (lane == 0) {
      for (int i = tma_issued; i < KB; ++i) {
        const int s = i % C::kStages;
        const uint32_t it = i / C::kStages;
        mbar_wait(&empty[s], (it & 1) ^ 1);
        issue_tma(i);
      }
      ND_TS(3);
    }
  } else if (warp == kMmaWarp) {
    // ===================================================================== MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
    // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int i = 0; i < KB; ++i) {
      const int s = i % C::kStages;
      const uint32_t it = i / C::kStages;
      mbar_wait(&conv_full[s], it & 1);
      tc_fence_after();
      if (lane == 0 && i == 0) ND_TS(4);
      if (lane == 0) {
        const uint64_t dah = make_desc(smem_u32(a_hi(s)));
        const uint64_t dbh = make_desc(smem_u32(b_hi(s)));
        const uint64_t dal = make_desc(smem_u32(a_lo(s)));
        const uint64_t dbl = make_desc(smem_u32(b_lo(s)));
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);      // 32 bytes per K=8 step inside the swizzle row
          const uint32_t first = ((Q > 1 ? i % kb_q : i) == 0 && k == 0) ? 0u : 1u;
          const uint32_t acc_t = Q > 1 ? tmem_base + (uint32_t)((i / kb_q) * BN) : tmem_base;
          if constexpr (ATM) {
            const uint32_t ah = tmem_base + (uint32_t)(C::ACC_COLS + s * 64 + k * 8), al = ah + 32;
            umma_tf32_ts(acc_t, al, dbh + adv, idesc, first);
            umma_tf32_ts(acc_t, ah, dbl + adv, idesc, 1u);
            umma_tf32_ts(acc_t, ah, dbh + adv, idesc, 1u);
          } else if (NPASS == 3) {
            umma_tf32(acc_t, dal + adv, dbh + adv, idesc, first);
            umma_tf32(acc_t, dah + adv, dbl + adv, idesc, 1u);
            umma_tf32(acc_t, dah + adv, dbh + adv, idesc, 1u);
          } else {
            umma_tf32(acc_t, dah + adv, dbh + adv, idesc, first);
          }
        }
        umma_commit(&empty[s]);                 // frees the smem stage when these MMAs retire
        if (i == KB - 1) { umma_commit(tmem_full); ND_TS(5); }
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== converters (warps 0-7)
    // thread = (row, half): half h converts 16-byte chunks [4h, 4h+4) of the row's 128-byte k-block slice
    const int row = threadIdx.x & (BM - 1);
    const int half = threadIdx.x >> 7;
    const int m = m0 + row;
    const int sw = row & 7;
    float x0 = 0.f, s1 = 0.f, s2 = 0.f;         // shifted one-pass moments for the folded LayerNorm
    int cnt = 0;
    for (int i = 0; i < KB; ++i) {
      const int s = i % C::kStages;
      const uint32_t it = i / C::kStages;
      mbar_wait(&raw_full[s], it & 1);
      if (threadIdx.x == 0 && i == 0) ND_TS(6);
      if (NPASS == 3 || fold) {
        uint8_t* rh = a_hi(s) + row * 128;
        uint8_t* rl = a_lo(s) + row * 128;
        // all four chunks first (loads cannot be hoisted over the in-place stores by the compiler)
        float4 vin[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) vin[c] = *reinterpret_cast<float4*>(rh + (((4 * half + c) ^ sw) << 4));
        if (fold && (Q > 1 ? i % kb_q : i) == 0) x0 = vin[0].x;   // shift = first element of THIS thread's share (race free)
        uint32_t th[16], tl[16];                // ATM: this thread's 16 columns of the hi / lo tile
#pragma unroll
        for (int c = 0; c < 4; ++c) {           // logical 16-byte chunk lc lives at physical chunk lc ^ (row & 7)
          const int lc = 4 * half + c;
          const int pc = (lc ^ sw) << 4;
          const float x[4] = {vin[c].x, vin[c].y, vin[c].z, vin[c].w};
          const int kbase = (kb0 + i) * BK + lc * 4;
          if (fold) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool in = kbase + q < p.K;
              const float dlt = in ? x[q] - x0 : 0.f;
              cnt += in ? 1 : 0;
              s1 += dlt;
              s2 = fmaf(dlt, dlt, s2);
            }
          }
          if (NPASS == 3) {
            float h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              h[q] = __uint_as_float((__float_as_uint(x[q]) + 0x1000u) & 0xffffe000u);     // nearest tf32
              l[q] = __uint_as_float((__float_as_uint(x[q] - h[q]) + 0x1000u) & 0xffffe000u);
            }
            if constexpr (ATM) {
#pragma unroll
              for (int q = 0; q < 4; ++q) { th[4 * c + q] = __float_as_uint(h[q]); tl[4 * c + q] = __float_as_uint(l[q]); }
            } else {
              *reinterpret_cast<float4*>(rh + pc) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(rl + pc) = make_float4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        if constexpr (ATM) {
          // warp w owns tensor-memory lanes 32*(w & 3)..: its rows; columns [16*half, 16*half + 16) of the stage
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(C::ACC_COLS + s * 64 + half * 16);
          tmem_st16(ta, th);
          tmem_st16(ta + 32, tl);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        } else if (NPASS == 3) {
          fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core (async proxy)
        }
      }
      mbar_arrive(&conv_full[s]);
      if (threadIdx.x == 0 && i == 0) ND_TS(7);
      if constexpr (Q > 1) {
        if (fold && (i + 1) % kb_q == 0) {      // end of a serial K slice: park its half-row moments, start afresh
          float mean_h = 0.f, m2_h = 0.f;
          if (cnt > 0) {
            const float ds = s1 / (float)cnt;
            mean_h = x0 + ds;
            m2_h = fmaxf(s2 - s1 * ds, 0.f);
          }
          float* mo = s_mom + (((i / kb_q) * 2 + half) * BM + row) * 3;
          mo[0] = (float)cnt; mo[1] = mean_h; mo[2] = m2_h;
          cnt = 0; s1 = 0.f; s2 = 0.f;
        }
      }
    }
    if (threadIdx.x == 0) ND_TS(8);
    // each thread: (cnt, mean, M2) of its half-row share; combine the two halves (Chan et al.) into the
    // moments of this CTA's K slice
    float mean_s = 0.f, m2_s = 0.f;
    if constexpr (Q > 1) {
      if (fold) {
        asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
        float mean_all = 0.f, m2_all = 0.f;
        int n_all = 0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {            // same two-level combination as the cluster kernel (halves, then slices)
          const float* a = s_mom + ((q * 2 + 0) * BM + row) * 3;
          const float* b = s_mom + ((q * 2 + 1) * BM + row) * 3;
          const float na = a[0], nb = b[0], nt = na + nb;
          float mu = 0.f, mm = 0.f;
          if (nt > 0.f) {
            const float delta = b[1] - a[1];
            mu = a[1] + delta * (nb / nt);
            mm = a[2] + b[2] + delta * delta * (na * nb / nt);
          }
          const int k_lo = q * kb_q * BK;
          const int nn = max(0, min(p.K, k_lo + kb_q * BK) - k_lo);
          if (nn > 0) {
            const float delta = mu - mean_all;
            const int ntot = n_all + nn;
            mean_all += delta * ((float)nn / (float)ntot);
            m2_all += mm + delta * delta * ((float)n_all * (float)nn / (float)ntot);
            n_all = ntot;
          }
        }
        mean_s = mean_all;
        m2_s = m2_all;
      }
    } else if (fold) {
      float mean_h = 0.f, m2_h = 0.f;
      if (cnt > 0) {
        const float ds = s1 / (float)cnt;
        mean_h = x0 + ds;
        m2_h = fmaxf(s2 - s1 * ds, 0.f);
      }
      float* mo = s_mom + (half * BM + row) * 3;
      mo[0] = (float)cnt; mo[1] = mean_h; mo[2] = m2_h;
      asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
      const float* a = s_mom + row * 3;
      const float* b = s_mom + (BM + row) * 3;
      const float na = a[0], nb = b[0], nt = na + nb;
      if (nt > 0.f) {
        const float delta = b[1] - a[1];
        mean_s = a[1] + delta * (nb / nt);
        m2_s = a[2] + b[2] + delta * delta * (na * nb / nt);
      }
    }

    // ===================================================================== epilogue (all 8 warps)
    // warp w reads TMEM lanes 32*(w&3).. (its rows) and the column half (w>>2) of the tile
    const int lg = warp & 3, chalf = warp >> 2;
    const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
    constexpr int HC = BN / 2;                   // columns per warp
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (threadIdx.x == 0) ND_TS(9);
    if constexpr (S == 1) {
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (fold) {
        ln_mean = mean_s;
        ln_rstd = 1.0f / sqrtf(m2_s / (float)p.K + p.eps);
      }
      // per-warp staging tile [32 rows][36 floats] in the (now idle) pipeline stages: accumulators are
      // transposed through it so that global stores / residual loads are 128-byte contiguous per row
      float* stg = reinterpret_cast<float*>(tiles) + warp * (32 * 36);
      const int rsub = lane >> 3, cq = lane & 7;
#pragma unroll 1
      for (int cb = 0; cb < HC; cb += 32) {
        const int c0 = chalf * HC + cb;
        const int nb = n0 + c0;
        float v[32];
        if constexpr (Q == 1) {
          tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);   // warp-collective: all lanes participate
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
          for (int q = 0; q < Q; ++q) {          // slice order of the cluster reduction -> bitwise the same sum
            float t[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(q * BN + c0), t);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += t[j];
          }
        }
        if (vec_ok && nb + 32 <= p.N) {
          // residual rows in flight while the row-wise math runs (4 rows x 128 B per instruction)
          float4 res[8];
          if (p.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mm = m0 + lg * 32 + 4 * i + rsub;
              res[i] = mm < p.M ? *reinterpret_cast<const float4*>(p.residual + (int64_t)mm * p.ldr + nb + cq * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          epilogue_rowwise<32>(p, v, c0, fold, ln_mean, ln_rstd, s_vec0, s_vec1);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            const int mm = m0 + lg * 32 + r;
            float4 o = *reinterpret_cast<const float4*>(stg + r * 36 + cq * 4);
            if (p.residual) { o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w; }
            if (p.relu == 3) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (mm < p.M) *reinterpret_cast<float4*>(p.C + (int64_t)mm * p.ldc + nb + cq * 4) = o;
          }
          __syncwarp();
        } else if (m < p.M && nb < p.N) {
          epilogue_store_slow<32>(p, v, m, nb, fold, ln_mean, ln_rstd);
        }
        __syncwarp();
      }
    } else {
      // push this row's partial sums (and its partial LayerNorm moments) into the receive buffer of the
      // CTA that owns the row: remote shared-memory stores are posted, one cluster barrier publishes them
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      constexpr int RPC = BM / S;
      const int owner = row / RPC, rl = row % RPC;
      float* dst = cluster.map_shared_rank(part + (size_t)(crank * RPC + rl) * PS, owner);
#pragma unroll 1
      for (int cb = 0; cb < HC; cb += 32) {
        const int c0 = chalf * HC + cb;
        float v[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (chalf == 0) {
        float* sdst = cluster.map_shared_rank(s_stats + 2 * (crank * RPC + rl), owner);
        sdst[0] = mean_s;
        sdst[1] = m2_s;
      }
    }
    if (threadIdx.x == 0) ND_TS(10);
  }

  if constexpr (S > 1) {
    namespace cg = cooperative_groups;
    constexpr int RPC = BM / S;                  // rows reduced by this CTA
    constexpr int TPR = kConvThreads / RPC;      // threads per row (2S)
    constexpr int CH = BN / TPR;                 // columns per thread
    static_assert(CH % 4 == 0, "split-K column chunk");
    const int rl = threadIdx.x / TPR;
    const int cbeg = (threadIdx.x % TPR) * CH;
    const int m = m0 + crank * RPC + rl;
    const bool conv = warp < kConvWarps;
    const bool fast = vec_ok && n0 + cbeg + CH <= p.N;
    float4 res[CH / 4];
    if (conv && p.residual && m < p.M && fast) {      // overlaps the barrier
      const float* rrow = p.residual + (int64_t)m * p.ldr + n0 + cbeg;
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) res[j] = *reinterpret_cast<const float4*>(rrow + 4 * j);
    }
    tc_fence_before();
    cg::this_cluster().sync();                   // every partial row has landed in its owner's shared memory
    if (threadIdx.x == 0) ND_TS(11);
    if (conv) {
      float acc[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) acc[j] = 0.f;
      float mean_all = 0.f, m2_all = 0.f;
      int n_all = 0;
#pragma unroll
      for (int s = 0; s < S; ++s) {              // fixed order -> bitwise deterministic
        const float* rp = part + (size_t)(s * RPC + rl) * PS + cbeg;
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(rp + j);
          acc[j] += t.x; acc[j + 1] += t.y; acc[j + 2] += t.z; acc[j + 3] += t.w;
        }
        if (fold) {
          // Chan et al. pairwise update of (n, mean, M2) with K slice s
          const float* sp = s_stats + 2 * (s * RPC + rl);
          const int k_lo = s * kb_per * BK;
          const int nn = max(0, min(p.K, k_lo + kb_per * BK) - k_lo);
          if (nn > 0) {
            const float mu = sp[0], mm = sp[1];
            const float delta = mu - mean_all;
            const int nt = n_all + nn;
            mean_all += delta * ((float)nn / (float)nt);
            m2_all += mm + delta * delta * ((float)n_all * (float)nn / (float)nt);
            n_all = nt;
          }
        }
      }
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (fold) {
        ln_mean = mean_all;
        ln_rstd = 1.0f / sqrtf(m2_all / (float)p.K + p.eps);
      }
      if (m < p.M && n0 + cbeg < p.N) {
        if (fast) {
          epilogue_rowwise<CH>(p, acc, cbeg, fold, ln_mean, ln_rstd, s_vec0, s_vec1);
          float* crow = p.C + (int64_t)m * p.ldc + n0 + cbeg;
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            if (p.residual) { o.x += res[j / 4].x; o.y += res[j / 4].y; o.z += res[j / 4].z; o.w += res[j / 4].w; }
            if (p.relu == 3) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(crow + j) = o;
          }
        } else {
          epilogue_store_slow<CH>(p, acc, m, n0 + cbeg, fold, ln_mean, ln_rstd);
        }
      }
    }
    if (threadIdx.x == 0) ND_TS(12);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
    if (lane == 0) ND_TS(14);
  }
}

// ------------------------------------------------------------------------------------ persistent kernel
// Same arithmetic and epilogue as gemm_tc_kernel<BN, NPASS, 1>, for the large-M projections of the encoder side
// (input projections of the LSTM layers, memory K/V, Transformer encoder): one CTA per SM walks the tile list
// (n fastest, so the CTAs running at any moment share their A rows in L2), the TMA / converter / MMA pipeline
// never drains between tiles, and the accumulator is double buffered in tensor memory so that the epilogue of
// tile i (4 dedicated warps) overlaps the main loop of tile i+1.  The one-tile-per-CTA kernel spends ~7 us per
// tile outside the MMAs (launch, TMEM allocation, pipeline fill, epilogue); here only the MMA time remains.
//   warps 0-7  converters (tf32 hi/lo split in place, LayerNorm row moments)
//   warps 8-11 epilogue (tcgen05.ld -> bias / folded LayerNorm / ReLU -> smem transpose -> residual -> stores)
//   warp 12    TMA producer        warp 13  TMEM owner + MMA issuer
constexpr int kPEpiWarps = 4;
constexpr int kPEpiThreads = kPEpiWarps * 32;
constexpr int kPTmaWarp = kConvWarps + kPEpiWarps, kPMmaWarp = kPTmaWarp + 1;
constexpr int kPThreads = (kPMmaWarp + 1) * 32;

// ATM (3xTF32 only): the converters write the tf32 hi / lo parts of the A tile into TENSOR MEMORY (tcgen05.st, 64
// columns per stage) instead of back to shared memory, and the MMAs take A from there: a third less shared-memory
// traffic per k-block, no generic->async proxy fence, cheaper MMAs (A is not fetched through the shared-memory port).
template <int BN, int NPASS, bool ATM = false>
struct PCfg {
  static_assert(!ATM || NPASS == 3, "A in tensor memory is the 3xTF32 variant");
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int A_BYTES = A_TILE_BYTES * ((NPASS == 3 && !ATM) ? 2 : 1);
  static constexpr int STAGE_BYTES = A_BYTES + B_TILE_BYTES * (NPASS == 3 ? 2 : 1);
  static constexpr int STG_BYTES = kPEpiWarps * 32 * 36 * 4;
  static constexpr int AUX_BYTES = 256 /*barriers*/ + 2 * BM * 2 * 4 /*row stats x2*/ + 2 * 2 * BN * 4 /*vectors x2*/ +
                                   2 * BM * 3 * 4 /*half-row moments*/;
  static constexpr int kStagesFit = (222 * 1024 - STG_BYTES - AUX_BYTES - 1024) / STAGE_BYTES;
  static constexpr int kStagesCap = ATM ? (512 - 2 * BN) / 64 : 6;       // ATM: 64 tensor-memory columns per stage
  static constexpr int kStages = kStagesFit >= kStagesCap ? kStagesCap : kStagesFit;
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + STG_BYTES + AUX_BYTES + 1024 /*align*/;
  static constexpr int ACC_COLS = 2 * BN;                                 // A stages start here
  static constexpr int NEED_COLS = ACC_COLS + (ATM ? kStages * 64 : 0);
  static constexpr int TMEM_COLS = NEED_COLS <= 32 ? 32 : (NEED_COLS <= 64 ? 64 : (NEED_COLS <= 128 ? 128 : (NEED_COLS <= 256 ? 256 : 512)));
  static_assert(kStages >= 2, "pipeline needs two stages");
  static_assert(NEED_COLS <= 512, "tensor memory");
};

// Measured limits (profiles/r01_gemm_persistent.md): at 150 TFLOP/s fp32-equivalent the kernel moves 7 TB/s of
// operand tiles from L2 and ~190 KB of shared-memory traffic per k-block (the three passes re-read the hi / lo tiles).
// Fetching ONE fp32 weight tile and splitting it on chip like the activations was tried: a third less L2 traffic,
// but the extra converter traffic on shared memory made it 15 % slower.
template <int BN, int NPASS, bool ATM = false>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
                       const __grid_constant__ CUtensorMap tmWlo, GemmParams p, int n_tiles) {
  using C = PCfg<BN, NPASS, ATM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* stg_base = reinterpret_cast<float*>(smem + C::kStages * C::STAGE_BYTES);
  uint8_t* aux = smem + C::kStages * C::STAGE_BYTES + C::STG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* raw_full = bars;
  uint64_t* conv_full = bars + C::kStages;
  uint64_t* empty = bars + 2 * C::kStages;
  uint64_t* acc_full = bars + 3 * C::kStages;          // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                  // [2] epilogue -> MMA / converters (stats buffer)
  uint64_t* stat_full = acc_empty + 2;                 // [2] converters -> epilogue (LayerNorm row moments)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_full + 2);
  float* s_stats = reinterpret_cast<float*>(aux + 256);        // [2][BM][2] (mean, M2)
  float* s_vec = s_stats + 2 * BM * 2;                          // [2][2][BN]
  float* s_mom = s_vec + 2 * 2 * BN;                            // [2][BM][3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = (p.N + BN - 1) / BN;
  const int KB = (p.K + BK - 1) / BK;
  const bool fold = p.ln_cvec != nullptr;
  if (p.alive && *p.alive == 0) return;        // uniform over the grid (see GemmParams::alive)

  auto a_hi = [&](int s) { return tiles + s * C::STAGE_BYTES; };
  auto a_lo = [&](int s) { return tiles + s * C::STAGE_BYTES + A_TILE_BYTES; };      // unused with ATM
  auto b_hi = [&](int s) { return tiles + s * C::STAGE_BYTES + C::A_BYTES; };
  auto b_lo = [&](int s) { return b_hi(s) + C::B_TILE_BYTES; };

  if (warp == kPTmaWarp && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], kConvThreads);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kPEpiThreads);
      mbar_init(&stat_full[b], BM);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    if (NPASS == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
  }
  if (warp == kPMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kPTmaWarp) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % C::kStages;
          mbar_wait(&empty[s], ((it / C::kStages) & 1) ^ 1);
          mbar_expect_tx(&raw_full[s], A_TILE_BYTES + C::B_TILE_BYTES * (NPASS == 3 ? 2 : 1));
          tma_load_2d(a_hi(s), &tmA, &raw_full[s], kb * BK, m0);
          tma_load_2d(b_hi(s), &tmWhi, &raw_full[s], kb * BK, n0);
          if (NPASS == 3) tma_load_2d(b_lo(s), &tmWlo, &raw_full[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == kPMmaWarp) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    uint32_t it = 0, j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const uint32_t buf = j & 1;
      mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % C::kStages;
        mbar_wait(&conv_full[s], (it / C::kStages) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t dah = make_desc(smem_u32(a_hi(s)));
          const uint64_t dbh = make_desc(smem_u32(b_hi(s)));
          const uint64_t dal = make_desc(smem_u32(a_lo(s)));
          const uint64_t dbl = make_desc(smem_u32(b_lo(s)));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            if constexpr (ATM) {
              const uint32_t ah = tmem_base + (uint32_t)(C::ACC_COLS + s * 64 + k * 8), al = ah + 32;
              umma_tf32_ts(tmem_d, al, dbh + adv, idesc, first);
              umma_tf32_ts(tmem_d, ah, dbl + adv, idesc, 1u);
              umma_tf32_ts(tmem_d, ah, dbh + adv, idesc, 1u);
            } else if (NPASS == 3) {
              umma_tf32(tmem_d, dal + adv, dbh + adv, idesc, first);
              umma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1u);
              umma_tf32(tmem_d, dah + adv, dbh + adv, idesc, 1u);
            } else {
              umma_tf32(tmem_d, dah + adv, dbh + adv, idesc, first);
            }
          }
          umma_commit(&empty[s]);
          if (kb == KB - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < kConvWarps) {
    // ===================================================================== converters
    // thread = (row, half): half h converts 16-byte chunks [4h, 4h+4) of the row's 128-byte k-block slice
    const int row = threadIdx.x & (BM - 1);
    const int half = threadIdx.x >> 7;
    const int sw = row & 7;
    uint32_t it = 0, j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      float x0 = 0.f, s1 = 0.f, s2 = 0.f;
      int cnt = 0;
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % C::kStages;
        mbar_wait(&raw_full[s], (it / C::kStages) & 1);
        if (NPASS == 3 || fold) {
          uint8_t* rh = a_hi(s) + row * 128;
          uint8_t* rl = a_lo(s) + row * 128;
          float4 vin[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) vin[c] = *reinterpret_cast<float4*>(rh + (((4 * half + c) ^ sw) << 4));
          if (fold && kb == 0) x0 = vin[0].x;     // shift = first element of THIS thread's share
          uint32_t th[16], tl[16];                // ATM: this thread's 16 columns of the hi / lo tile
#pragma unroll
          for (int c = 0; c < 4; ++c) {           // logical 16-byte chunk lc lives at physical chunk lc ^ (row & 7)
            const int lc = 4 * half + c;
            const int pc = (lc ^ sw) << 4;
            const float x[4] = {vin[c].x, vin[c].y, vin[c].z, vin[c].w};
            const int kbase = kb * BK + lc * 4;
            if (fold) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const bool in = kbase + q < p.K;
                const float dlt = in ? x[q] - x0 : 0.f;
                cnt += in ? 1 : 0;
                s1 += dlt;
                s2 = fmaf(dlt, dlt, s2);
              }
            }
            if (NPASS == 3) {
              float h[4], l[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                h[q] = __uint_as_float((__float_as_uint(x[q]) + 0x1000u) & 0xffffe000u);
                l[q] = __uint_as_float((__float_as_uint(x[q] - h[q]) + 0x1000u) & 0xffffe000u);
              }
              if constexpr (ATM) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { th[4 * c + q] = __float_as_uint(h[q]); tl[4 * c + q] = __float_as_uint(l[q]); }
              } else {
                *reinterpret_cast<float4*>(rh + pc) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(rl + pc) = make_float4(l[0], l[1], l[2], l[3]);
              }
            }
          }
          if constexpr (ATM) {
            // warp w owns tensor-memory lanes 32*(w & 3)..: exactly its rows; columns [16*half, 16*half + 16) of the stage
            const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(C::ACC_COLS + s * 64 + half * 16);
            tmem_st16(ta, th);
            tmem_st16(ta + 32, tl);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          } else if (NPASS == 3) {
            fence_proxy_async();
          }
        }
        mbar_arrive(&conv_full[s]);
      }
      if (fold) {
        // (cnt, mean, M2) of this thread's half row; combine the halves (Chan et al.) and hand the row moments
        // of this tile to the epilogue through the double-buffered stats array
        float mean_h = 0.f, m2_h = 0.f;
        if (cnt > 0) {
          const float ds = s1 / (float)cnt;
          mean_h = x0 + ds;
          m2_h = fmaxf(s2 - s1 * ds, 0.f);
        }
        float* mo = s_mom + (half * BM + row) * 3;
        mo[0] = (float)cnt; mo[1] = mean_h; mo[2] = m2_h;
        asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
        if (half == 0) {
          const float* a = s_mom + row * 3;
          const float* b = s_mom + (BM + row) * 3;
          const float na = a[0], nb = b[0], nt = na + nb;
          float mean_s = 0.f, m2_s = 0.f;
          if (nt > 0.f) {
            const float delta = b[1] - a[1];
            mean_s = a[1] + delta * (nb / nt);
            m2_s = a[2] + b[2] + delta * delta * (na * nb / nt);
          }
          const uint32_t buf = j & 1;
          mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);      // the epilogue two tiles back has read its moments
          s_stats[(buf * BM + row) * 2] = mean_s;
          s_stats[(buf * BM + row) * 2 + 1] = m2_s;
          mbar_arrive(&stat_full[buf]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
      }
    }
  } else {
    // ===================================================================== epilogue (warps 8-11)
    const int lg = warp - kConvWarps;                 // == warp % 4: the TMEM lane quadrant this warp may read
    const int et = threadIdx.x - kConvThreads;        // 0..127
    const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.residual || (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)));
    float* stg = stg_base + lg * (32 * 36);
    const int rsub = lane >> 3, cq = lane & 7;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const uint32_t buf = j & 1;
      const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN;
      const int m = m0 + lg * 32 + lane;
      float* v0 = s_vec + (buf * 2) * BN;
      float* v1 = v0 + BN;
      for (int i = et; i < BN; i += kPEpiThreads) {
        const int n = n0 + i;
        v0[i] = n < p.N ? (fold ? p.ln_cvec[n] : (p.bias ? p.bias[n] : 0.f)) : 0.f;
        v1[i] = (fold && n < p.N) ? p.ln_dvec[n] : 0.f;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kPEpiThreads) : "memory");
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (fold) {
        mbar_wait(&stat_full[buf], (j >> 1) & 1);
        ln_mean = s_stats[(buf * BM + lg * 32 + lane) * 2];
        ln_rstd = 1.0f / sqrtf(s_stats[(buf * BM + lg * 32 + lane) * 2 + 1] / (float)p.K + p.eps);
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int nb = n0 + c0;
        float v[32];
        tmem_ld32(tmem_base + lane_base + buf * BN + (uint32_t)c0, v);
        if (vec_ok && nb + 32 <= p.N) {
          float4 res[8];
          if (p.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mm = m0 + lg * 32 + 4 * i + rsub;
              res[i] = mm < p.M ? *reinterpret_cast<const float4*>(p.residual + (int64_t)mm * p.ldr + nb + cq * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          epilogue_rowwise<32>(p, v, c0, fold, ln_mean, ln_rstd, v0, v1);
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4)
            *reinterpret_cast<float4*>(stg + lane * 36 + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            const int mm = m0 + lg * 32 + r;
            float4 o = *reinterpret_cast<const float4*>(stg + r * 36 + cq * 4);
            if (p.residual) { o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w; }
            if (p.relu == 3) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (mm < p.M) *reinterpret_cast<float4*>(p.C + (int64_t)mm * p.ldc + nb + cq * 4) = o;
          }
          __syncwarp();
        } else if (m < p.M && nb < p.N) {
          epilogue_store_slow<32>(p, v, m, nb, fold, ln_mean, ln_rstd);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);                   // accumulator (and its stats / vectors) may be reused
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kPMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------ host side
long long* g_dbg = nullptr;
PFN_cuTensorMapEncodeTiled g_encode = nullptr;
bool g_lookup_done = false;
const char* g_why = "";

bool lookup() {
  if (g_lookup_done) return g_encode != nullptr;
  g_lookup_done = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    g_why = "cuTensorMapEncodeTiled driver entry point not found";
    cudaGetLastError();
    return false;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  return true;
}

bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int NPASS, int S, int Q = 1, bool ATM = false>
cudaError_t launch(const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN, NPASS, S, Q, ATM>;
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS, S, Q, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tmA, tmWhi, tmWlo;
  if (!make_map(&tmA, p.A, p.M, p.K, p.lda, BM)) return cudaErrorInvalidValue;
  if (!make_map(&tmWhi, p.W, p.N, p.K, p.ldw, BN)) return cudaErrorInvalidValue;
  if (!make_map(&tmWlo, NPASS == 3 ? p.W_lo : p.W, p.N, p.K, p.ldw, BN)) return cudaErrorInvalidValue;
  const int64_t tiles = (int64_t)cdiv(p.N, BN) * cdiv(p.M, BM);
  GemmParams pp = p;
  pp.dbg = g_dbg;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * S));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, NPASS, S, Q, ATM>, tmA, tmWhi, tmWlo, pp);
}

int g_persist = 2;       // 0: always one tile per CTA (cross-check), 1: persistent kernel, 2 (default): ... with A in tensor memory
int g_serial_split = 1;  // 0: cluster split-K also for many rows (cross-check: results are bit-identical)
int g_atm = 1;           // one-tile-per-CTA kernels: A operand in tensor memory (0: shared memory; same bits)
int g_wide_wave = 1;     // 128-wide tiles when the 64-wide tiles of a 1024-row problem exceed one wave (see dispatch)
int g_n_sm = 0;

template <int BN, int NPASS, bool ATM = false>
cudaError_t launch_persist(const GemmParams& p, cudaStream_t stream) {
  using C = PCfg<BN, NPASS, ATM>;
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_persist_kernel<BN, NPASS, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  CUtensorMap tmA, tmWhi, tmWlo;
  if (!make_map(&tmA, p.A, p.M, p.K, p.lda, BM)) return cudaErrorInvalidValue;
  if (!make_map(&tmWhi, p.W, p.N, p.K, p.ldw, BN)) return cudaErrorInvalidValue;
  if (!make_map(&tmWlo, NPASS == 3 ? p.W_lo : p.W, p.N, p.K, p.ldw, BN)) return cudaErrorInvalidValue;
  const int64_t tiles = (int64_t)cdiv(p.N, BN) * cdiv(p.M, BM);
  const int grid = (int)(tiles < g_n_sm ? tiles : g_n_sm);
  gemm_tc_persist_kernel<BN, NPASS, ATM><<<grid, kPThreads, C::SMEM_BYTES, stream>>>(tmA, tmWhi, tmWlo, p, (int)tiles);
  return cudaGetLastError();
}

template <int NPASS>
cudaError_t dispatch(const GemmParams& p, cudaStream_t stream) {
  // Tile width follows the amount of work (wide tiles for the encoder-side GEMMs).  The split-K factor
  // fixes the summation order of every output element, so for the decode-step GEMMs (M < 8192 rows) it
  // is a function of (N, K) ONLY: the same chunk then produces bit-identical results whatever batch or
  // stream group it is decoded in.  It is sized so that ~1024 rows fill about one wave of the 148 SMs.
  const int KB = cdiv(p.K, BK);
  const int64_t tiles128 = (int64_t)cdiv(p.N, 128) * cdiv(p.M, BM);
  int split = 1;
  if (p.M < 8192) {
    const int ref_tiles = 8 * cdiv(p.N, 64);           // 64-wide tiles of a 1024-row problem
    while (split < 4 && KB % (split * 2) == 0 && KB / (split * 2) >= 2 && ref_tiles * split * 2 <= 160) split *= 2;
  }
  // many tiles per SM: the persistent kernel (pipeline never drains, epilogue overlapped)
  if (g_persist && split == 1 && p.N > 64 && tiles128 >= 200) {
    if constexpr (NPASS == 3) {
      if (g_persist == 2) return launch_persist<128, 3, true>(p, stream);      // A operand in tensor memory
    }
    return launch_persist<128, NPASS>(p, stream);
  }
  if constexpr (NPASS == 3) {
    // A operand in tensor memory (same bits) for the 64-wide tiles: decode 81.2 -> 80.7 ms.  The one-tile 128-wide
    // kernel (N = 2048 at M = 1024) measured slower with it (26.8 -> 29.0 us) and keeps A in shared memory.
    const bool wide_for_wave = g_wide_wave && split == 1 && p.M < 8192 && 8 * cdiv(p.N, 64) > 148 && p.N >= 256;
    if (g_atm && !(g_serial_split && split > 1 && p.M >= 2048) && !(split == 1 && (tiles128 >= 120 || wide_for_wave) && p.N > 64)) {
      switch (split) {
        case 4: return launch<64, 3, 4, 1, true>(p, stream);
        case 2: return launch<64, 3, 2, 1, true>(p, stream);
        default: return launch<64, 3, 1, 1, true>(p, stream);
      }
    }
  }
  // 64-wide tiles of a 1024-row problem that do not fit one wave of SMs (N = 1536: 192 CTAs, the second wave 30 % full):
  // 128-wide tiles instead (96 CTAs).  A function of N only, and the tile width never changes the summation order.
  const bool wide_for_wave = g_wide_wave && split == 1 && p.M < 8192 && 8 * cdiv(p.N, 64) > 148 && p.N >= 256;
  if (split == 1 && (tiles128 >= 120 || wide_for_wave) && p.N > 64) return launch<128, NPASS, 1>(p, stream);
  // many rows (beam x batch): the cluster split would run several waves of short CTAs; one CTA per 128 x 128 tile
  // with the K slices in separate accumulators gives the same bits in one wave
  if (g_serial_split && split > 1 && p.M >= 2048)
    return split == 4 ? launch<128, NPASS, 1, 4>(p, stream) : launch<128, NPASS, 1, 2>(p, stream);
  switch (split) {
    case 4: return launch<64, NPASS, 4>(p, stream);
    case 2: return launch<64, NPASS, 2>(p, stream);
    default: return launch<64, NPASS, 1>(p, stream);
  }
}

}  // namespace

bool make_plain_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int box_cols) {
  if (!lookup()) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void gemm_tc_set_debug(long long* dev_buf) { g_dbg = dev_buf; }
void gemm_tc_set_persistent(int on) { g_persist = on; }
void gemm_tc_set_serial_split(int on) { g_serial_split = on; }
void gemm_tc_set_a_tmem(int on) { g_atm = on; }
void gemm_tc_set_wide_wave(int on) { g_wide_wave = on; }

bool gemm_tc_available(const char** why) {
  const bool ok = lookup();
  if (why) *why = g_why;
  return ok;
}

cudaError_t gemm_tc(const GemmParams& p, int npass, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  if (!lookup()) return cudaErrorNotSupported;
  // the tensor-core kernel keeps a minimal epilogue; prologues are folded into the weights by the caller
  if (p.prologue != PRO_NONE || p.div_ncols != 0 || p.relu == 2) return cudaErrorInvalidValue;
  // TMA needs 16-byte aligned bases and row pitches
  if ((p.lda & 3) || (p.ldw & 3) || (reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.W) & 15) ||
      (npass == 3 && (p.W_lo == nullptr || (reinterpret_cast<uintptr_t>(p.W_lo) & 15))))
    return cudaErrorInvalidValue;
  return npass == 3 ? dispatch<3>(p, stream) : dispatch<1>(p, stream);
}

}  // namespace nd
