// Transformer-encoder self attention on the 5th-generation tensor cores (head size 32 or 64, T <= 512).
// Reference: encoder/transformer.py:36-54 -> onmt/modules/multi_headed_attn.py:69-192 (q / sqrt(dh) before
// Q.K^T, masked_fill(-1e18) of the keys whose SIGNAL VALUE is 0.0, softmax in fp32, context = P.V).
//
// One CTA per (chunk, head).  fp32-level accuracy from two-term fp16 splits (x = x_hi + x_lo, 22 significant
// bits), three tensor-core products per matrix product, fp32 accumulation in tensor memory:
//   setup      K and V of the head (T x 32 fp32 each) are split ONCE into tcgen05 operand tiles in shared memory:
//              K rows [K_hi | K_lo] (one 128-byte swizzle row per key), V transposed to [feature][key] hi / lo
//   per 128-query tile, per 128-key block (flash style, running max / sum per query row in registers):
//     S  = Q.K^T      A = [Q_hi | Q_hi], [Q_lo | .] (shared memory), B = key rows: 6 MMAs (M 128, N 128, K 16)
//     P  = exp(S - m) 8 warps: tcgen05.ld, mask, row max (two half rows exchanged through shared memory),
//                     MUFU.EX2, fp16 hi / lo split, tcgen05.st -> P is the A operand IN TENSOR MEMORY
//     O_j = P.V       B = [V_hi ; V_lo] (N = 64) for P_hi plus V_hi (N = 32) for P_lo: 16 MMAs (K = 128 keys)
//     o  = o * exp(m_old - m_new) + O_j * exp(m_j - m_new)    in registers (16 features per thread)
// The FFMA kernel in attention.cu (one query per thread) stays as the cross-check and for other head sizes.
#include <cuda_fp16.h>
#include <float.h>

#include "kernels.cuh"

namespace nd {

namespace {

constexpr int DH = 32, TQ = 128, TK = 128, TMAX = 512;
constexpr int kSmWarps = 8, kSmThreads = kSmWarps * 32;       // softmax / conversion warps
constexpr int kThreads = kSmThreads + 32;                     // + MMA / TMEM-owner warp
constexpr int KOP_BYTES = TMAX * 128;                          // key rows [K_hi | K_lo]
constexpr int VKB_BYTES = 64 * 128;                            // one 64-key k-block: [V_hi: 32 rows][V_lo: 32 rows] x 128 B
constexpr int VOP_BYTES = (TMAX / 64) * VKB_BYTES;
constexpr int QOP_BYTES = 2 * TQ * 128;                        // A1 = [Q_hi | Q_hi], A2 = [Q_lo | 0]
constexpr int SMEM_BYTES = KOP_BYTES + VOP_BYTES + QOP_BYTES + TMAX * 4 /*flags*/ + 6 * TQ * 4 /*max (x2) and sum exchange*/ + 128 + 1024;
// tensor memory columns
constexpr uint32_t kColS = 0, kColPhi = 128, kColPlo = 192, kColO = 256, kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {     // K-major SWIZZLE_128B, SBO = 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// x = hi + lo in fp16; element 0 in the low half of the packed word
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);                 // one packed conversion per pair
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(kThreads, 1) enc_attn_tc_kernel(EncAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* kop = smem;                                   // [TMAX keys][128 B]
  uint8_t* vop = kop + KOP_BYTES;                        // [8 k-blocks][64 rows][128 B]
  uint8_t* qop = vop + VOP_BYTES;                        // A1, A2: [128 rows][128 B] each
  float* kflag = reinterpret_cast<float*>(qop + QOP_BYTES);        // [TMAX] 0 valid, 1 masked, 2 beyond T
  float* mx = kflag + TMAX;                              // [2 parities][2 halves][128 rows] max, then [2][128] sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(mx + 6 * TQ);
  uint64_t* go = bars;                                   // softmax warps -> MMA: Q operand ready / O_j consumed
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* o_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, d = p.d;
  const float* base = p.qkv + (int64_t)b * T * 3 * d + h * DH;   // + t*3d: q | + d: k | + 2d: v
  const int nkb = (T + TK - 1) / TK;                     // key blocks of 128
  const int nqt = (T + TQ - 1) / TQ;

  if (tid == 0) {
    mbar_init(go, kSmThreads);
    mbar_init(s_full, 1);
    mbar_init(p_full, kSmThreads);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kSmWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- K and V of this head -> operand tiles (zero beyond T)
  if (warp < kSmWarps) {
    for (int i = tid; i < nkb * TK * 4; i += kSmThreads) {       // task = (key, 8-feature chunk); lanes = consecutive keys
      const int c = i / (nkb * TK), t = i - c * (nkb * TK);
      float kf[8], vf[8];
      if (t < T) {
        const float4 k0 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + d + 8 * c);
        const float4 k1 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + d + 8 * c + 4);
        const float4 v0 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + 2 * d + 8 * c);
        const float4 v1 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + 2 * d + 8 * c + 4);
        kf[0] = k0.x; kf[1] = k0.y; kf[2] = k0.z; kf[3] = k0.w; kf[4] = k1.x; kf[5] = k1.y; kf[6] = k1.z; kf[7] = k1.w;
        vf[0] = v0.x; vf[1] = v0.y; vf[2] = v0.z; vf[3] = v0.w; vf[4] = v1.x; vf[5] = v1.y; vf[6] = v1.z; vf[7] = v1.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) { kf[q] = 0.f; vf[q] = 0.f; }
      }
      // key row t: 16-byte unit c holds K_hi[8c .. 8c+7], unit 4 + c holds K_lo
      uint4 hi4, lo4;
      split2(kf[0], kf[1], hi4.x, lo4.x); split2(kf[2], kf[3], hi4.y, lo4.y);
      split2(kf[4], kf[5], hi4.z, lo4.z); split2(kf[6], kf[7], hi4.w, lo4.w);
      uint8_t* krow = kop + t * 128;
      *reinterpret_cast<uint4*>(krow + ((c ^ (t & 7)) << 4)) = hi4;
      *reinterpret_cast<uint4*>(krow + (((4 + c) ^ (t & 7)) << 4)) = lo4;
      // V transposed: k-block t / 64, row = feature (hi) | 32 + feature (lo), column t % 64
      uint8_t* vkb = vop + (t >> 6) * VKB_BYTES;
      const int kk = t & 63;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = 8 * c + q;
        const __half vh = __float2half_rn(vf[q]);
        const __half vl = __float2half_rn(vf[q] - __half2float(vh));
        *reinterpret_cast<__half*>(vkb + e * 128 + (((kk >> 3) ^ (e & 7)) << 4) + ((kk & 7) << 1)) = vh;
        *reinterpret_cast<__half*>(vkb + (32 + e) * 128 + (((kk >> 3) ^ (e & 7)) << 4) + ((kk & 7) << 1)) = vl;
      }
    }
    for (int t = tid; t < TMAX; t += kSmThreads)
      kflag[t] = t < T ? (p.src[(int64_t)b * T + t] == 0.0f ? 1.f : 0.f) : 2.f;
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptors: D = F32, A = B = F16, K-major, M = 128
  const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == kSmWarps) {
    // ===================================================================== MMA issuer
    uint32_t it = 0;
    for (int qt = 0; qt < nqt; ++qt)
      for (int j = 0; j < nkb; ++j, ++it) {
        mbar_wait(go, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          // S = Q.K^T for key block j: [Q_hi | Q_hi] . [K_hi | K_lo] (k16 steps 0..3) + [Q_lo | .] . [K_hi | .] (steps 0..1)
          const uint64_t da1 = make_desc(smem_u32(qop));
          const uint64_t da2 = make_desc(smem_u32(qop + TQ * 128));
          const uint64_t dbk = make_desc(smem_u32(kop + j * TK * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + kColS, da1 + 2 * k, dbk + 2 * k, idesc128, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_ss(tmem_base + kColS, da2 + 2 * k, dbk + 2 * k, idesc128, 1u);
          umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          // O_j = P.V over the 128 keys of the block: 8 k16 steps; [V_hi ; V_lo] with P_hi (N 64), V_hi with P_lo (N 32)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t dbv = make_desc(smem_u32(vop + (2 * j + (ks >> 2)) * VKB_BYTES)) + (uint64_t)(2 * (ks & 3));
            umma_ts(tmem_base + kColO, tmem_base + kColPhi + ks * 8, dbv, idesc64, ks ? 1u : 0u);
            umma_ts(tmem_base + kColO, tmem_base + kColPlo + ks * 8, dbv, idesc32, 1u);
          }
          umma_commit(o_full);
        }
        __syncwarp();
      }
  } else {
    // ===================================================================== softmax warps
    const int lq = warp & 3, hf = warp >> 2;               // TMEM lane quadrant, key / feature half
    const int row = lq * 32 + lane;                        // query row of the tile = TMEM lane
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    const float kLog2e = 1.4426950408889634f;
    uint32_t it = 0;
    for (int qt = 0; qt < nqt; ++qt) {
      // ---- Q tile -> A operands (task = (row, 8-feature chunk): 512 tasks, two per thread)
      for (int i = tid; i < TQ * 4; i += kSmThreads) {
        const int r = i >> 2, c = i & 3;
        const int tq = qt * TQ + r;
        float qf[8];
        if (tq < T) {
          const float4 q0 = *reinterpret_cast<const float4*>(base + (int64_t)tq * 3 * d + 8 * c);
          const float4 q1 = *reinterpret_cast<const float4*>(base + (int64_t)tq * 3 * d + 8 * c + 4);
          qf[0] = q0.x / p.q_div; qf[1] = q0.y / p.q_div; qf[2] = q0.z / p.q_div; qf[3] = q0.w / p.q_div;
          qf[4] = q1.x / p.q_div; qf[5] = q1.y / p.q_div; qf[6] = q1.z / p.q_div; qf[7] = q1.w / p.q_div;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) qf[q] = 0.f;
        }
        uint4 hi4, lo4;
        split2(qf[0], qf[1], hi4.x, lo4.x); split2(qf[2], qf[3], hi4.y, lo4.y);
        split2(qf[4], qf[5], hi4.z, lo4.z); split2(qf[6], qf[7], hi4.w, lo4.w);
        uint8_t* a1 = qop + r * 128;
        uint8_t* a2 = qop + TQ * 128 + r * 128;
        *reinterpret_cast<uint4*>(a1 + ((c ^ (r & 7)) << 4)) = hi4;             // [Q_hi | Q_hi]
        *reinterpret_cast<uint4*>(a1 + (((4 + c) ^ (r & 7)) << 4)) = hi4;
        *reinterpret_cast<uint4*>(a2 + ((c ^ (r & 7)) << 4)) = lo4;             // [Q_lo | unused]
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      float m_run = -INFINITY, l_run = 0.f;
      float o_run[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o_run[i] = 0.f;
      mbar_arrive(go);                                     // Q operand of this tile is in place

      for (int j = 0; j < nkb; ++j, ++it) {
        mbar_wait(s_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- this thread's half row of scores: keys j*128 + 64 hf + [0, 64)
        float s[64];
        {
          float v0[32], v1[32];
          tmem_ld32(tmem_base + lane_base + kColS + 64 * hf, v0);
          tmem_ld32(tmem_base + lane_base + kColS + 64 * hf + 32, v1);
#pragma unroll
          for (int i = 0; i < 32; ++i) { s[i] = v0[i]; s[32 + i] = v1[i]; }
        }
        const float* fl = kflag + j * TK + 64 * hf;
        float mloc = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 f = *reinterpret_cast<const float4*>(fl + i);
          s[i] = f.x == 0.f ? s[i] : (f.x == 1.f ? -1e18f : -INFINITY);         // masked_fill(-1e18) | nonexistent key
          s[i + 1] = f.y == 0.f ? s[i + 1] : (f.y == 1.f ? -1e18f : -INFINITY);
          s[i + 2] = f.z == 0.f ? s[i + 2] : (f.z == 1.f ? -1e18f : -INFINITY);
          s[i + 3] = f.w == 0.f ? s[i + 3] : (f.w == 1.f ? -1e18f : -INFINITY);
          mloc = fmaxf(fmaxf(mloc, fmaxf(s[i], s[i + 1])), fmaxf(s[i + 2], s[i + 3]));
        }
        float* mxb = mx + (it & 1) * 2 * TQ;               // double buffered by block parity: one barrier per block
        mxb[hf * TQ + row] = mloc;
        asm volatile("bar.sync 1, %0;" ::"n"(kSmThreads) : "memory");
        const float m_j = fmaxf(mloc, mxb[(hf ^ 1) * TQ + row]);           // block max of the row (finite: a key exists)
        // ---- P = exp(S - m_j) in (0, 1], fp16 hi / lo, packed pairs -> tensor memory (A operand of P.V)
        float lsum = 0.f;
        uint32_t phi[32], plo[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          // (s - m) first: exact 0 for the row maximum also when both are the -1e18 mask value
          const float p0 = ex2_approx((s[i] - m_j) * kLog2e);
          const float p1 = ex2_approx((s[i + 1] - m_j) * kLog2e);
          lsum += p0 + p1;
          split2(p0, p1, phi[i >> 1], plo[i >> 1]);
        }
        tmem_st32(tmem_base + lane_base + kColPhi + 32 * hf, phi);
        tmem_st32(tmem_base + lane_base + kColPlo + 32 * hf, plo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(p_full);
        // ---- fold block j into the running state (l_run: this thread's HALF of the keys; halves meet at the end)
        const float m_new = fmaxf(m_run, m_j);
        const float a_old = ex2_approx((m_run - m_new) * kLog2e);          // 0 on the first block
        const float a_blk = ex2_approx((m_j - m_new) * kLog2e);
        l_run = l_run * a_old + lsum * a_blk;
        m_run = m_new;
        mbar_wait(o_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          float oa[16], ob[16];                            // features 16 hf + [0,16): P_hi.V_hi + P_lo.V_hi | P_hi.V_lo
          tmem_ld16(tmem_base + lane_base + kColO + 16 * hf, oa);
          tmem_ld16(tmem_base + lane_base + kColO + 32 + 16 * hf, ob);
#pragma unroll
          for (int i = 0; i < 16; ++i) o_run[i] = o_run[i] * a_old + (oa[i] + ob[i]) * a_blk;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (j + 1 < nkb) mbar_arrive(go);                  // S, P and O of this block are consumed
      }
      // the two halves of a row share m_run; their partial sums meet once per tile
      float* lx = mx + 4 * TQ;
      lx[hf * TQ + row] = l_run;
      asm volatile("bar.sync 1, %0;" ::"n"(kSmThreads) : "memory");
      l_run += lx[(hf ^ 1) * TQ + row];
      asm volatile("bar.sync 1, %0;" ::"n"(kSmThreads) : "memory");       // read before the next tile rewrites lx
      const int tq = qt * TQ + row;
      if (tq < T) {
        const float inv = 1.0f / l_run;
        float* out = p.ctx + ((int64_t)b * T + tq) * d + h * DH + 16 * hf;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(out + i) = make_float4(o_run[i] * inv, o_run[i + 1] * inv, o_run[i + 2] * inv, o_run[i + 3] * inv);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kSmWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------
// Head size 64 (d = 512, 8 heads).  Same scheme; the operand tiles of all 512 keys would take 256 KB, so the keys
// are processed in two halves of 256 and the K / V operand tiles of a half (64 + 64 KB) are rebuilt for every
// query tile (8 conversions of 256 keys per CTA instead of 2: ~15 % of the CTA's time).
//   K half:  two planes [K_hi | K_lo], each [256 keys][128 B = 64 fp16]
//   V half:  4 k-blocks of 64 keys, each [V_hi: 64 feature rows][V_lo: 64 rows] x 128 B
//   S = Q_hi.K_hi + Q_hi.K_lo + Q_lo.K_hi (12 MMAs, N 128); O_j = P_hi.[V_hi ; V_lo] (N 128) + P_lo.V_hi (N 64)
constexpr int DH64 = 64, KHALF = 256;
constexpr int K64_PLANE = KHALF * 128;                         // 32 KB
constexpr int V64_KB = 128 * 128;                              // 16 KB per 64-key k-block
constexpr int Q64_BYTES = 2 * TQ * 128;                        // A_hi, A_lo
constexpr int SMEM64_BYTES = 2 * K64_PLANE + (KHALF / 64) * V64_KB + Q64_BYTES + TMAX * 4 + 6 * TQ * 4 + 128 + 1024;
constexpr uint32_t kColO64 = 256;                              // O: 128 columns [hh + lh | hl]

__global__ void __launch_bounds__(kThreads, 1) enc_attn_tc64_kernel(EncAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* kop = smem;                                   // [2 planes][256 keys][128 B]
  uint8_t* vop = kop + 2 * K64_PLANE;                    // [4 k-blocks][128 rows][128 B]
  uint8_t* qop = vop + (KHALF / 64) * V64_KB;            // A_hi, A_lo: [128 rows][128 B]
  float* kflag = reinterpret_cast<float*>(qop + Q64_BYTES);
  float* mx = kflag + TMAX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(mx + 6 * TQ);
  uint64_t* go = bars;
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* o_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, d = p.d;
  const float* base = p.qkv + (int64_t)b * T * 3 * d + h * DH64;
  const int nkb = (T + TK - 1) / TK;
  const int nqt = (T + TQ - 1) / TQ;

  if (tid == 0) {
    mbar_init(go, kSmThreads);
    mbar_init(s_full, 1);
    mbar_init(p_full, kSmThreads);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kSmWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp < kSmWarps)
    for (int t = tid; t < TMAX; t += kSmThreads)
      kflag[t] = t < T ? (p.src[(int64_t)b * T + t] == 0.0f ? 1.f : 0.f) : 2.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == kSmWarps) {
    // ===================================================================== MMA issuer
    uint32_t it = 0;
    for (int qt = 0; qt < nqt; ++qt)
      for (int j = 0; j < nkb; ++j, ++it) {
        const int jb = j & 1;                              // key block inside the resident half
        mbar_wait(go, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint64_t dah = make_desc(smem_u32(qop));
          const uint64_t dal = make_desc(smem_u32(qop + TQ * 128));
          const uint64_t dkh = make_desc(smem_u32(kop + jb * TK * 128));
          const uint64_t dkl = make_desc(smem_u32(kop + K64_PLANE + jb * TK * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + kColS, dah + 2 * k, dkh + 2 * k, idesc128, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + kColS, dah + 2 * k, dkl + 2 * k, idesc128, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + kColS, dal + 2 * k, dkh + 2 * k, idesc128, 1u);
          umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t dbv = make_desc(smem_u32(vop + (2 * jb + (ks >> 2)) * V64_KB)) + (uint64_t)(2 * (ks & 3));
            umma_ts(tmem_base + kColO64, tmem_base + kColPhi + ks * 8, dbv, idesc128, ks ? 1u : 0u);   // [V_hi ; V_lo]
            umma_ts(tmem_base + kColO64, tmem_base + kColPlo + ks * 8, dbv, idesc64, 1u);              // V_hi
          }
          umma_commit(o_full);
        }
        __syncwarp();
      }
  } else {
    // ===================================================================== softmax / conversion warps
    const int lq = warp & 3, hf = warp >> 2;
    const int row = lq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    const float kLog2e = 1.4426950408889634f;
    // K and V of keys [256 kh, 256 kh + 256) -> operand tiles (task = (key, 8-feature chunk), lanes = consecutive keys)
    auto convert_half = [&](int kh) {
      for (int i = tid; i < KHALF * 8; i += kSmThreads) {
        const int c = i / KHALF, tl = i - c * KHALF;
        const int t = kh * KHALF + tl;
        float kf[8], vf[8];
        if (t < T) {
          const float4 k0 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + d + 8 * c);
          const float4 k1 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + d + 8 * c + 4);
          const float4 v0 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + 2 * d + 8 * c);
          const float4 v1 = *reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d + 2 * d + 8 * c + 4);
          kf[0] = k0.x; kf[1] = k0.y; kf[2] = k0.z; kf[3] = k0.w; kf[4] = k1.x; kf[5] = k1.y; kf[6] = k1.z; kf[7] = k1.w;
          vf[0] = v0.x; vf[1] = v0.y; vf[2] = v0.z; vf[3] = v0.w; vf[4] = v1.x; vf[5] = v1.y; vf[6] = v1.z; vf[7] = v1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) { kf[q] = 0.f; vf[q] = 0.f; }
        }
        uint4 hi4, lo4;
        split2(kf[0], kf[1], hi4.x, lo4.x); split2(kf[2], kf[3], hi4.y, lo4.y);
        split2(kf[4], kf[5], hi4.z, lo4.z); split2(kf[6], kf[7], hi4.w, lo4.w);
        const int ko = tl * 128 + ((c ^ (tl & 7)) << 4);
        *reinterpret_cast<uint4*>(kop + ko) = hi4;
        *reinterpret_cast<uint4*>(kop + K64_PLANE + ko) = lo4;
        uint8_t* vkb = vop + (tl >> 6) * V64_KB;
        const int kk = tl & 63;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int e = 8 * c + q;
          const __half vh = __float2half_rn(vf[q]);
          const __half vl = __float2half_rn(vf[q] - __half2float(vh));
          const int vo = (((kk >> 3) ^ (e & 7)) << 4) + ((kk & 7) << 1);
          *reinterpret_cast<__half*>(vkb + e * 128 + vo) = vh;
          *reinterpret_cast<__half*>(vkb + (64 + e) * 128 + vo) = vl;
        }
      }
    };
    uint32_t it = 0;
    int resident = -1;                                     // which key half the operand tiles hold
    for (int qt = 0; qt < nqt; ++qt) {
      // ---- Q tile -> A operands (task = (row, 8-feature chunk): 1024 tasks, four per thread)
      for (int i = tid; i < TQ * 8; i += kSmThreads) {
        const int r = i >> 3, c = i & 7;
        const int tq = qt * TQ + r;
        float qf[8];
        if (tq < T) {
          const float4 q0 = *reinterpret_cast<const float4*>(base + (int64_t)tq * 3 * d + 8 * c);
          const float4 q1 = *reinterpret_cast<const float4*>(base + (int64_t)tq * 3 * d + 8 * c + 4);
          qf[0] = q0.x / p.q_div; qf[1] = q0.y / p.q_div; qf[2] = q0.z / p.q_div; qf[3] = q0.w / p.q_div;
          qf[4] = q1.x / p.q_div; qf[5] = q1.y / p.q_div; qf[6] = q1.z / p.q_div; qf[7] = q1.w / p.q_div;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) qf[q] = 0.f;
        }
        uint4 hi4, lo4;
        split2(qf[0], qf[1], hi4.x, lo4.x); split2(qf[2], qf[3], hi4.y, lo4.y);
        split2(qf[4], qf[5], hi4.z, lo4.z); split2(qf[6], qf[7], hi4.w, lo4.w);
        const int qo = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(qop + qo) = hi4;
        *reinterpret_cast<uint4*>(qop + TQ * 128 + qo) = lo4;
      }
      float m_run = -INFINITY, l_run = 0.f;
      float o_run[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o_run[i] = 0.f;

      for (int j = 0; j < nkb; ++j, ++it) {
        // operands of this block: rebuild the key half when it is not the resident one (every MMA that read the
        // tiles has completed: this thread has seen o_full of the previous block)
        if ((j >> 1) != resident) {
          convert_half(j >> 1);
          resident = j >> 1;
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_arrive(go);                                   // Q / K / V operands in place, S, P, O of the last block consumed
        mbar_wait(s_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float s[64];
        {
          float v0[32], v1[32];
          tmem_ld32(tmem_base + lane_base + kColS + 64 * hf, v0);
          tmem_ld32(tmem_base + lane_base + kColS + 64 * hf + 32, v1);
#pragma unroll
          for (int i = 0; i < 32; ++i) { s[i] = v0[i]; s[32 + i] = v1[i]; }
        }
        const float* fl = kflag + j * TK + 64 * hf;
        float mloc = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 f = *reinterpret_cast<const float4*>(fl + i);
          s[i] = f.x == 0.f ? s[i] : (f.x == 1.f ? -1e18f : -INFINITY);
          s[i + 1] = f.y == 0.f ? s[i + 1] : (f.y == 1.f ? -1e18f : -INFINITY);
          s[i + 2] = f.z == 0.f ? s[i + 2] : (f.z == 1.f ? -1e18f : -INFINITY);
          s[i + 3] = f.w == 0.f ? s[i + 3] : (f.w == 1.f ? -1e18f : -INFINITY);
          mloc = fmaxf(fmaxf(mloc, fmaxf(s[i], s[i + 1])), fmaxf(s[i + 2], s[i + 3]));
        }
        float* mxb = mx + (it & 1) * 2 * TQ;
        mxb[hf * TQ + row] = mloc;
        asm volatile("bar.sync 1, %0;" ::"n"(kSmThreads) : "memory");
        const float m_j = fmaxf(mloc, mxb[(hf ^ 1) * TQ + row]);
        float lsum = 0.f;
        uint32_t phi[32], plo[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          const float p0 = ex2_approx((s[i] - m_j) * kLog2e);
          const float p1 = ex2_approx((s[i + 1] - m_j) * kLog2e);
          lsum += p0 + p1;
          split2(p0, p1, phi[i >> 1], plo[i >> 1]);
        }
        tmem_st32(tmem_base + lane_base + kColPhi + 32 * hf, phi);
        tmem_st32(tmem_base + lane_base + kColPlo + 32 * hf, plo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(p_full);
        const float m_new = fmaxf(m_run, m_j);
        const float a_old = ex2_approx((m_run - m_new) * kLog2e);
        const float a_blk = ex2_approx((m_j - m_new) * kLog2e);
        l_run = l_run * a_old + lsum * a_blk;
        m_run = m_new;
        mbar_wait(o_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          float oa[32], ob[32];                            // features 32 hf + [0,32): P_hi.V_hi + P_lo.V_hi | P_hi.V_lo
          tmem_ld32(tmem_base + lane_base + kColO64 + 32 * hf, oa);
          tmem_ld32(tmem_base + lane_base + kColO64 + 64 + 32 * hf, ob);
#pragma unroll
          for (int i = 0; i < 32; ++i) o_run[i] = o_run[i] * a_old + (oa[i] + ob[i]) * a_blk;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      float* lx = mx + 4 * TQ;
      lx[hf * TQ + row] = l_run;
      asm volatile("bar.sync 1, %0;" ::"n"(kSmThreads) : "memory");
      l_run += lx[(hf ^ 1) * TQ + row];
      asm volatile("bar.sync 1, %0;" ::"n"(kSmThreads) : "memory");
      const int tq = qt * TQ + row;
      if (tq < T) {
        const float inv = 1.0f / l_run;
        float* out = p.ctx + ((int64_t)b * T + tq) * d + h * DH64 + 32 * hf;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(out + i) = make_float4(o_run[i] * inv, o_run[i + 1] * inv, o_run[i + 2] * inv, o_run[i + 3] * inv);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kSmWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

}  // namespace

bool encoder_attention_tc_supported(const EncAttnParams& p) {
  return p.H > 0 && (p.d / p.H == DH || p.d / p.H == DH64) && p.d % p.H == 0 && p.T <= TMAX && (p.d % 4) == 0;
}

cudaError_t encoder_attention_tc(const EncAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  if (!encoder_attention_tc_supported(p)) return cudaErrorInvalidValue;
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(enc_attn_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM64_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (p.d / p.H == DH64) enc_attn_tc64_kernel<<<p.B * p.H, kThreads, SMEM64_BYTES, stream>>>(p);
  else enc_attn_tc_kernel<<<p.B * p.H, kThreads, SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace nd
