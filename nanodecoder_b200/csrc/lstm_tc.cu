// Persistent tensor-core LSTM recurrence for sm_100a (H = 128 hidden units per direction).
//
// Same contract as lstm_kernel in lstm.cu (packed-sequence nn.LSTM semantics, one layer, forward and
// reverse directions as separate work items; reference: encoder/nano_encoder.py:97-99,
// encoder/rnn_encoder.py:70-78), but the per-step recurrent product
//     gates[4H, BT] = W_hh[4H, H] . h[BT, H]^T
// runs on the 5th-generation tensor cores with fp32-level accuracy from a two-term fp16 split:
//     x = x_hi + x_lo  (x_hi = fp16(x), x_lo = fp16(x - x_hi): 22 significant bits, absolute error
//     <= 2^-25 for |x| <= 1),   W.h ~= W_lo.h_hi + W_hi.h_lo + W_hi.h_hi   (fp32 accumulation in TMEM).
// |h| <= 1 by construction and |W_hh| is checked against the fp16 range by the caller.
//
// One work item = (tile of BT = 32 chunks, direction), executed by a cluster of 2 CTAs; CTA r owns
// hidden units [64r, 64r+64) = 256 gate rows = two M=128 UMMA tiles (tile m: gates i,f,g,o of units
// 64r + 32m + [0,32)).  The W_hh slice of a CTA stays ON CHIP for all T steps: in TENSOR MEMORY as the
// A operand (variant A_TMEM: 256 of the 512 TMEM columns; each MMA then only reads the 1 KB h tile
// from shared memory) or in shared memory (variant A_SMEM, 128 KB).
//   per step:  one thread issues 2 x 24 tcgen05.mma.kind::f16 (K = 16 each) and commits each tile to
//              its own mbarrier; 4 warps per tile read the accumulators (tcgen05.ld), regroup the
//              four gates of each unit through shared memory, do the cell update in fp32 and write the
//              new h (fp16 hi/lo, 16-byte vectors) into the B-operand tiles of BOTH CTAs with st.async
//              (distributed shared memory, 8 KB per step to the peer); the stores complete transaction
//              bytes on an mbarrier of the destination CTA, which is all the MMA issuer waits for: no
//              cluster barrier, no proxy fence and no memory barrier inside the step loop.
//   input-side gate terms (x.W_ih^T + b_ih, a tcgen05 GEMM output) are prefetched one step ahead.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "lstm.cuh"

namespace cg = cooperative_groups;

namespace nd {

namespace {

constexpr int H = 128, C = 2, UPC = H / C, BT = 32;
constexpr int kPwThreads = 256;                 // 8 warps: TMEM readers + pointwise (4 per M tile)
constexpr int kThreads = kPwThreads + 32;       // + 1 MMA / TMEM-owner warp
constexpr int HB_TILE = BT * 128;               // bytes of one [BT rows x 64 fp16] h operand tile
constexpr int HB_KB = 2 * HB_TILE;              // one k-block: the hi tile followed by the lo tile (an N = 2 BT operand)
constexpr int HB_BUF = 2 * HB_KB;               // one buffer: 2 k-blocks
constexpr int HB_TOTAL = 2 * HB_BUF;            // double buffered
constexpr int G_FLOATS = 2 * 4 * BT * 32;       // gate exchange [tile][gate][chunk][unit]
constexpr int W_TILE = 128 * 128;               // A_SMEM: bytes of one [128 rows x 64 fp16] W operand tile
constexpr int W_TOTAL = 8 * W_TILE;             // [tile m][hi, lo][k-block]
constexpr int VEC_FLOATS = 3 * 4 * UPC;         // b_hh, w_ih0, b_ih0 of this CTA's units: [which][gate][unit]
constexpr uint32_t kHBytesPerStep = C * kPwThreads * 32;   // h bytes landing in one CTA's buffer per step
constexpr int TAIL_BYTES = G_FLOATS * 4 + VEC_FLOATS * 4 + 64 + BT * 4;

// The A_TMEM variant allocates all 512 TMEM columns: its shared-memory request is padded past half an
// SM so that two CTAs can never share an SM (the second tcgen05.alloc would wait forever).
template <bool A_TMEM>
constexpr int smem_bytes() {
  return A_TMEM ? (HB_TOTAL + TAIL_BYTES + 1024 > 120 * 1024 ? HB_TOTAL + TAIL_BYTES + 1024 : 120 * 1024)
                : HB_TOTAL + W_TOTAL + TAIL_BYTES + 1024 /*align*/;
}

// TMEM columns: A_TMEM: W tile m hi at m*128, lo at m*128 + 64 (two fp16 per 32-bit column), D at 256 + 64 m.
// D tile m = [W_hi.h_hi + W_lo.h_hi | W_hi.h_lo]: the hi and lo h tiles are adjacent in shared memory, so one
// N = 64 MMA per K step multiplies W_hi with both (the A operand is fetched once), one N = 32 MMA adds W_lo.h_hi
constexpr uint32_t kTmemColsA = 512, kTmemColsS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {     // K-major SWIZZLE_128B, SBO = 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// 16-byte asynchronous store into (possibly remote) shared memory; its bytes are counted on `mbar`, an
// mbarrier of the destination CTA
__device__ __forceinline__ void st_async_v4(uint32_t dst, const uint4& v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
               : "memory");
}

// x = hi + lo with hi, lo in fp16; returned packed as (lo of pair: element 0 in the low half)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
  const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
  hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
  lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}
// branch-free logistic: 1 / (1 + 2^(-x log2 e)) from MUFU.EX2 + MUFU.RCP (2 + 1 ulp); abs error ~2e-7.
// (__frcp_rn / exp2f carry slow-path calls that serialise the eight cells of a thread.)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.0f * sigmoid_fast(2.0f * x) - 1.0f; }
// One LSTM cell update (gate order i, f, g, o).  The four gate activations share ONE reciprocal:
// with a_x = 1 + e^(-x), 1/a_i = a_f a_g a_o / (a_i a_f a_g a_o); arguments are clamped to +-20 (sigmoid
// saturates below fp32 resolution there) so the product stays < 6e34.  7 MUFU per cell instead of 10:
// the cell update is MUFU-bound (16 / clk / SM).
__device__ __forceinline__ void lstm_cell(float xi, float xf, float xg, float xo, float& c, float& h) {
  const float k = -1.4426950408889634f;
  const float ai = 1.0f + ex2_approx(k * fminf(fmaxf(xi, -20.f), 20.f));
  const float af = 1.0f + ex2_approx(k * fminf(fmaxf(xf, -20.f), 20.f));
  const float ag = 1.0f + ex2_approx(k * fminf(fmaxf(2.0f * xg, -20.f), 20.f));
  const float ao = 1.0f + ex2_approx(k * fminf(fmaxf(xo, -20.f), 20.f));
  const float pif = ai * af, pgo = ag * ao;
  const float r = rcp_approx(pif * pgo);
  const float rif = r * pgo, rgo = r * pif;                   // 1/(a_i a_f), 1/(a_g a_o)
  const float si = rif * af, sf = rif * ai, so = rgo * ag;
  const float tg = 2.0f * (rgo * ao) - 1.0f;
  c = sf * c + si * tg;
  h = so * tanh_fast(c);
}

#define LTS(slot) do { if (p.dbg && blockIdx.x == 0 && s == 100) p.dbg[slot] = clock64(); } while (0)

template <bool A_TMEM>
__global__ void __launch_bounds__(kThreads, 1) lstm_tc_kernel(LstmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned (swizzle atoms)
  uint8_t* hb = smem;                                  // [2 buffers][2 k-blocks][hi, lo][BT x 128 B]
  uint8_t* wsm = smem + HB_TOTAL;                      // A_SMEM only: [m][hi, lo][k-block][128 x 128 B]
  float* G = reinterpret_cast<float*>(smem + HB_TOTAL + (A_TMEM ? 0 : W_TOTAL));
  float* s_vec = G + G_FLOATS;                         // [b_hh | w_ih0 | b_ih0][gate][unit of this CTA]
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(s_vec + VEC_FLOATS);   // [2] one per M tile
  uint64_t* hfull = mma_done + 2;                      // [2] one per h buffer: transaction bytes of one step
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 4);
  int* s_len = reinterpret_cast<int*>(mma_done + 8);   // [BT]

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster.block_rank();
  const int item = blockIdx.x / C;
  const int dir = item % p.dirs;
  const int tile = item / p.dirs;
  const int b0 = tile * BT;
  constexpr uint32_t kCols = A_TMEM ? kTmemColsA : kTmemColsS;
  constexpr uint32_t kDCol = A_TMEM ? 256u : 0u;

  // ---- one-time setup
  for (int i = tid; i < HB_TOTAL / 16; i += kThreads) reinterpret_cast<float4*>(hb)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < BT) s_len[tid] = (b0 + tid < p.B) ? (int)p.lengths[b0 + tid] : 0;
  for (int i = tid; i < VEC_FLOATS; i += kThreads) {
    const int which = i / (4 * UPC), g = (i / UPC) & 3, uu = i % UPC;
    const int64_t row = (int64_t)dir * 4 * H + g * H + rank * UPC + uu;
    float v = 0.f;
    if (which == 0) v = p.b_hh[row];
    else if (p.x0) v = (which == 1 ? p.w_ih0 : p.b_ih0)[row];
    s_vec[i] = v;
  }
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&hfull[0], 1);
    mbar_init(&hfull[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // W_hh slice of this CTA -> fp16 hi/lo A operand.  Pointwise thread (warp w, lane) owns operand row
  // 32 (w & 3) + lane of tile m = w >> 2, i.e. gate g = w & 3 of unit 64 rank + 32 m + lane.
  if (warp < 8) {
    const int m = warp >> 2, g = warp & 3;
    const int trow = g * 32 + lane;
    const float* wrow = p.w_hh + ((int64_t)dir * 4 * H + g * H + rank * UPC + m * 32 + lane) * H;
#pragma unroll
    for (int c = 0; c < 4; ++c) {                      // 32 K elements = 16 packed columns per round
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(wrow + c * 32 + q * 4);
        split2(v.x, v.y, hi[2 * q], lo[2 * q]);
        split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
      }
      if constexpr (A_TMEM) {
        const uint32_t t0 = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(m * 128 + c * 16);
        tmem_st16(t0, hi);
        tmem_st16(t0 + 64, lo);
      } else {
        // K-major SWIZZLE_128B: k-block = K / 64, 16-byte unit ((K % 64) / 8) ^ (row & 7)
        const int kb = c >> 1;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int unit16 = (c & 1) * 4 + q4;
          const int off = trow * 128 + ((unit16 ^ (trow & 7)) << 4);
          uint8_t* dh = wsm + ((m * 2 + 0) * 2 + kb) * W_TILE + off;
          uint8_t* dl = wsm + ((m * 2 + 1) * 2 + kb) * W_TILE + off;
          *reinterpret_cast<uint4*>(dh) = make_uint4(hi[4 * q4], hi[4 * q4 + 1], hi[4 * q4 + 2], hi[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(dl) = make_uint4(lo[4 * q4], lo[4 * q4 + 1], lo[4 * q4 + 2], lo[4 * q4 + 3]);
        }
      }
    }
    if constexpr (A_TMEM) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");            // operand tiles written by the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  int maxlen = 0;
#pragma unroll 4
  for (int b = 0; b < BT; ++b) maxlen = max(maxlen, s_len[b]);
  maxlen = min(maxlen, p.T);

  // pointwise ownership: thread t' = tid % 128 of tile m owns chunk b = t' / 4 and the 8 units
  // 32 m + 8 o + [0,8), o = t' % 4, of this CTA
  const int m = (warp >> 2) & 1;
  const int tq = tid & 127;
  const int b = tq >> 2, o = tq & 3;
  const int ul = m * 32 + o * 8;                              // first owned unit (local to the CTA)
  const int ucol = rank * UPC + ul;                           // ... in [0, H)
  const int len_b = s_len[b];
  const int bb = min(b0 + b, p.B - 1);
  const int out_ld = p.dirs * H;
  float c_state[8], h_state[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c_state[i] = 0.f; h_state[i] = 0.f; }
  // instruction descriptors: D = F32, A = B = F16, both K-major, M = 128, N = BT or 2 BT
  const uint32_t idesc32 = (1u << 4) | ((uint32_t)(BT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc64 = (1u << 4) | ((uint32_t)(2 * BT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  // raw input-side terms of step sn for this thread's cells.  Nothing is computed on the loaded values
  // here, so the loads stay in flight for a whole step; make_pre() finishes them at the point of use.
  auto load_raw = [&](int sn, float4 (&raw)[8]) {
    const bool active = sn < len_b;
    const int t = active ? (dir == 0 ? sn : len_b - 1 - sn) : 0;
    if (p.x0) {
      raw[0].x = p.x0[(int64_t)bb * p.T + t];
    } else {
      const float* xr = p.xg + ((int64_t)bb * p.T + t) * p.xg_ld + (int64_t)dir * 4 * H + ucol;
#pragma unroll
      for (int g = 0; g < 4; ++g) { raw[2 * g] = ldg_stream4(xr + g * H); raw[2 * g + 1] = ldg_stream4(xr + g * H + 4); }
    }
  };
  auto make_pre = [&](const float4 (&raw)[8], float (&pre)[4][8]) {       // x.W_ih^T + b_ih + b_hh
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 q0 = *reinterpret_cast<const float4*>(s_vec + g * UPC + ul);
      const float4 q1 = *reinterpret_cast<const float4*>(s_vec + g * UPC + ul + 4);
      const float bh[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      if (p.x0) {
        const float xv = raw[0].x;
#pragma unroll
        for (int i = 0; i < 8; ++i) pre[g][i] = (xv * s_vec[(4 + g) * UPC + ul + i] + s_vec[(8 + g) * UPC + ul + i]) + bh[i];
      } else {
        const float4 v0 = raw[2 * g], v1 = raw[2 * g + 1];
        pre[g][0] = v0.x + bh[0]; pre[g][1] = v0.y + bh[1]; pre[g][2] = v0.z + bh[2]; pre[g][3] = v0.w + bh[3];
        pre[g][4] = v1.x + bh[4]; pre[g][5] = v1.y + bh[5]; pre[g][6] = v1.z + bh[6]; pre[g][7] = v1.w + bh[7];
      }
    }
  };

  float4 raw_next[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) raw_next[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp < 8) load_raw(0, raw_next);

  // destinations of this thread's 16-byte h vectors: k-block `rank` of the operand, row b, 16-byte unit
  // (4 m + o) ^ (b & 7); in this CTA and in the peer, for both buffers; and the mbarriers that count them
  uint32_t dst_own[2], dst_peer[2], bar_own[2], bar_peer[2];
  {
    const int off = rank * HB_KB + b * 128 + ((((m << 2) | o) ^ (b & 7)) << 4);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t a = smem_u32(hb + (size_t)q * HB_BUF + off), mb = smem_u32(&hfull[q]);
      dst_own[q] = mapa_u32(a, (uint32_t)rank);
      dst_peer[q] = mapa_u32(a, (uint32_t)(rank ^ 1));
      bar_own[q] = mapa_u32(mb, (uint32_t)rank);
      bar_peer[q] = mapa_u32(mb, (uint32_t)(rank ^ 1));
    }
  }
  cluster.sync();            // barriers initialised and h buffers zeroed in both CTAs before any st.async arrives

  for (int s = 0; s < maxlen; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    if (tid == 0) LTS(0);
    if (warp == 8) {
      // ================= MMA issuer: gates = W_hh . h_cur^T  (W_hi.[h_hi ; h_lo] + W_lo.h_hi)
      if (lane == 0) {
        if (s > 0) mbar_wait(&hfull[cur], (uint32_t)(((s - 1) >> 1) & 1));   // h_s landed (both CTAs' halves)
        mbar_arrive_expect_tx(&hfull[nxt], kHBytesPerStep);                   // arm the buffer h_{s+1} goes to
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t hcur = smem_u32(hb + (size_t)cur * HB_BUF);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t dcol = tmem_base + kDCol + (uint32_t)(mt * 2 * BT);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const int kb = ks >> 2, k16 = ks & 3;
            const uint64_t db = make_desc(hcur + kb * HB_KB) + (uint64_t)(k16 * 2);        // 32 B per K = 16
            if constexpr (A_TMEM) {
              const uint32_t a_hi = tmem_base + (uint32_t)(mt * 128 + ks * 8);
              umma_f16_ts(dcol, a_hi, db, idesc64, ks ? 1u : 0u);          // W_hi . [h_hi ; h_lo]
              umma_f16_ts(dcol, a_hi + 64, db, idesc32, 1u);               // W_lo . h_hi
            } else {
              const uint64_t da_hi = make_desc(smem_u32(wsm + ((mt * 2 + 0) * 2 + kb) * W_TILE)) + (uint64_t)(k16 * 2);
              const uint64_t da_lo = make_desc(smem_u32(wsm + ((mt * 2 + 1) * 2 + kb) * W_TILE)) + (uint64_t)(k16 * 2);
              umma_f16_ss(dcol, da_hi, db, idesc64, ks ? 1u : 0u);
              umma_f16_ss(dcol, da_lo, db, idesc32, 1u);
            }
          }
          umma_commit(&mma_done[mt]);
        }
        LTS(1);
      }
      __syncwarp();
    } else {
      float4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) raw[i] = raw_next[i];
      if (s + 1 < maxlen) load_raw(s + 1, raw_next);          // in flight during this whole step
      if (tid == 0) LTS(2);
      // ================= accumulators -> shared memory, regrouped as G[tile][gate][chunk][unit]
      mbar_wait(&mma_done[m], (uint32_t)(s & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) LTS(3);
      {
        const int g = warp & 3;                               // TMEM lanes 32 g .. 32 g + 31: gate g, unit = lane
        float v[32], v2[32];
        const uint32_t t0 = tmem_base + ((uint32_t)(g * 32) << 16) + kDCol + (uint32_t)(m * 2 * BT);
        tmem_ld32(t0, v);
        tmem_ld32(t0 + BT, v2);
        float* Gw = G + ((m * 4 + g) * BT) * 32 + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) Gw[i * 32] = v[i] + v2[i];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (m == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid == 0) LTS(4);
      // ================= cell update of chunk b, units ul .. ul+7
      if (s < len_b) {
        float pre[4][8];
        make_pre(raw, pre);
        float gate[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 a0 = *reinterpret_cast<const float4*>(G + ((m * 4 + g) * BT + b) * 32 + o * 8);
          const float4 a1 = *reinterpret_cast<const float4*>(G + ((m * 4 + g) * BT + b) * 32 + o * 8 + 4);
          gate[g][0] = pre[g][0] + a0.x; gate[g][1] = pre[g][1] + a0.y; gate[g][2] = pre[g][2] + a0.z; gate[g][3] = pre[g][3] + a0.w;
          gate[g][4] = pre[g][4] + a1.x; gate[g][5] = pre[g][5] + a1.y; gate[g][6] = pre[g][6] + a1.z; gate[g][7] = pre[g][7] + a1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) lstm_cell(gate[0][i], gate[1][i], gate[2][i], gate[3][i], c_state[i], h_state[i]);
        const int t = dir == 0 ? s : len_b - 1 - s;
        float* op = p.out + ((int64_t)(b0 + b) * p.T + t) * out_ld + dir * H + ucol;
        *reinterpret_cast<float4*>(op) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(h_state[4], h_state[5], h_state[6], h_state[7]);
      }
      // ================= h (frozen once the chunk has ended) -> B operand tiles of both CTAs.  st.async writes
      // through the async proxy (what the tensor cores read) and counts its bytes on the destination CTA's
      // mbarrier: no proxy fence, no cluster barrier, no round trip.
      uint4 hi4, lo4;
      split2(h_state[0], h_state[1], hi4.x, lo4.x);
      split2(h_state[2], h_state[3], hi4.y, lo4.y);
      split2(h_state[4], h_state[5], hi4.z, lo4.z);
      split2(h_state[6], h_state[7], hi4.w, lo4.w);
      const uint32_t dp = nxt ? dst_peer[1] : dst_peer[0], bp = nxt ? bar_peer[1] : bar_peer[0];
      const uint32_t dn = nxt ? dst_own[1] : dst_own[0], bn = nxt ? bar_own[1] : bar_own[0];
      st_async_v4(dp, hi4, bp);
      st_async_v4(dp + HB_TILE, lo4, bp);
      st_async_v4(dn, hi4, bn);
      st_async_v4(dn + HB_TILE, lo4, bn);
      if (tid == 0) LTS(5);
    }
  }
  // every st.async aimed at this CTA has landed before it may exit
  if (tid == 0 && maxlen > 0) mbar_wait(&hfull[maxlen & 1], (uint32_t)(((maxlen - 1) >> 1) & 1));

  if (p.h_n && warp < 8 && b0 + b < p.B) {
    const int64_t oo = ((int64_t)dir * p.B + b0 + b) * H + ucol;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p.h_n[oo + i] = h_state[i]; p.c_n[oo + i] = c_state[i]; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols));
  }
  cluster.sync();                                             // no CTA exits while its peer may still write its smem
}

long long* g_lstm_dbg = nullptr;
int g_variant = 0;                                            // 0: W_hh in tensor memory, 1: W_hh in shared memory

template <bool A_TMEM>
cudaError_t launch(const LstmParams& p, cudaStream_t stream) {
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_kernel<A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<A_TMEM>());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = cdiv(p.B, BT);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * p.dirs * C));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes<A_TMEM>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LstmParams pp = p;
  pp.dbg = g_lstm_dbg;
  return cudaLaunchKernelEx(&cfg, lstm_tc_kernel<A_TMEM>, pp);
}

}  // namespace

void lstm_tc_set_debug(long long* dev_buf) { g_lstm_dbg = dev_buf; }
void lstm_tc_set_variant(int v) { g_variant = v; }

bool lstm_tc_supported(int H_) { return H_ == H; }

cudaError_t lstm_layer_tc(const LstmParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  if (p.H != H) return cudaErrorNotSupported;
  return g_variant == 0 ? launch<true>(p, stream) : launch<false>(p, stream);
}

}  // namespace nd
