// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epi( pro(A)[M,K] . W[N,K]^T )   (fp32 in/out)
//
// Arithmetic: kind::tf32 UMMA (M=128, N=BN, K=8 per instruction), fp32 accumulation in TMEM.
//   npass == 3  "3xTF32": a = a_hi + a_lo, w = w_hi + w_lo (tf32-exact parts);
//               D += a_lo.w_hi + a_hi.w_lo + a_hi.w_hi   -> ~2^-21 relative error (fp32 parity mode)
//   npass == 1  single TF32 pass (fast mode, ~2^-11)
//
// CTA = 192 threads, one 128 x BN output tile:
//   warp 4     TMA producer: A (raw fp32), W_hi, W_lo tiles of 128B-swizzled K-major rows (BK = 32)
//   warps 0-3  converters: apply the prologue (LayerNorm / column affine) to the A tile in place,
//              split it into tf32 hi/lo, publish to the async proxy; afterwards the epilogue
//              (tcgen05.ld -> bias / scale / ReLU / residual -> global)
//   warp 5     TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit to mbarriers
// Pipeline: kStages smem stages, mbarriers raw_full (TMA tx) -> conv_full (128 arrivals) -> empty (commit).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "gemm.cuh"

namespace nd {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = one 128-byte swizzle row
constexpr int kConvThreads = 128;
constexpr int kThreads = 192;
constexpr int A_TILE_BYTES = BM * 128;

template <int BN, int NPASS>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_TILE_BYTES * (NPASS == 3 ? 2 : 1) + B_TILE_BYTES * (NPASS == 3 ? 2 : 1);
  static constexpr int kStages = (200 * 1024) / STAGE_BYTES >= 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BM * 4;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
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

// ------------------------------------------------------------------------------------ kernel
template <int BN, int NPASS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
               const __grid_constant__ CUtensorMap tmWlo, GemmParams p) {
  using C = Cfg<BN, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by SWIZZLE_128B operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::STAGE_BYTES);
  uint64_t* raw_full = bars;
  uint64_t* conv_full = bars + C::kStages;
  uint64_t* empty = bars + 2 * C::kStages;
  uint64_t* tmem_full = bars + 3 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 1);
  float* s_mean = reinterpret_cast<float*>(smem + C::kStages * C::STAGE_BYTES + 256);
  float* s_rstd = s_mean + BM;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = (p.N + BN - 1) / BN;
  const int m0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
  const int KB = (p.K + BK - 1) / BK;

  auto a_hi = [&](int s) { return tiles + s * C::STAGE_BYTES; };
  auto a_lo = [&](int s) { return tiles + s * C::STAGE_BYTES + A_TILE_BYTES; };
  auto b_hi = [&](int s) { return tiles + s * C::STAGE_BYTES + A_TILE_BYTES * (NPASS == 3 ? 2 : 1); };
  auto b_lo = [&](int s) { return b_hi(s) + C::B_TILE_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], kConvThreads);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    if (NPASS == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % C::kStages;
        const uint32_t it = kb / C::kStages;
        mbar_wait(&empty[s], (it & 1) ^ 1);
        mbar_expect_tx(&raw_full[s], A_TILE_BYTES + C::B_TILE_BYTES * (NPASS == 3 ? 2 : 1));
        tma_load_2d(a_hi(s), &tmA, &raw_full[s], kb * BK, m0);
        tma_load_2d(b_hi(s), &tmWhi, &raw_full[s], kb * BK, n0);
        if (NPASS == 3) tma_load_2d(b_lo(s), &tmWlo, &raw_full[s], kb * BK, n0);
      }
    }
  } else if (warp == 5) {
    // ===================================================================== MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
    // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % C::kStages;
      const uint32_t it = kb / C::kStages;
      mbar_wait(&conv_full[s], it & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dah = make_desc(smem_u32(a_hi(s)));
        const uint64_t dbh = make_desc(smem_u32(b_hi(s)));
        const uint64_t dal = make_desc(smem_u32(a_lo(s)));
        const uint64_t dbl = make_desc(smem_u32(b_lo(s)));
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);      // 32 bytes per K=8 step inside the swizzle row
          const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
          if (NPASS == 3) {
            umma_tf32(tmem_base, dal + adv, dbh + adv, idesc, first);
            umma_tf32(tmem_base, dah + adv, dbl + adv, idesc, 1u);
            umma_tf32(tmem_base, dah + adv, dbh + adv, idesc, 1u);
          } else {
            umma_tf32(tmem_base, dah + adv, dbh + adv, idesc, first);
          }
        }
        umma_commit(&empty[s]);                 // frees the smem stage when these MMAs retire
        if (kb == KB - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== converters (warps 0-3)
    const int row = threadIdx.x;                // one A-tile row per thread (and one TMEM lane in the epilogue)
    const int m = m0 + row;
    const bool fold = p.ln_cvec != nullptr;
    const int sw = row & 7;
    float x0 = 0.f, s1 = 0.f, s2 = 0.f;         // shifted one-pass row moments for the folded LayerNorm
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % C::kStages;
      const uint32_t it = kb / C::kStages;
      mbar_wait(&raw_full[s], it & 1);
      if (NPASS == 3 || p.prologue != PRO_NONE || fold) {
        uint8_t* rh = a_hi(s) + row * 128;
        uint8_t* rl = a_lo(s) + row * 128;
        if (fold && kb == 0) x0 = *reinterpret_cast<float*>(rh + (sw << 4));      // logical element k = 0
#pragma unroll
        for (int c = 0; c < 8; ++c) {           // logical 16-byte chunk c lives at physical chunk c ^ (row & 7)
          const int pc = (c ^ sw) << 4;
          float4 v = *reinterpret_cast<float4*>(rh + pc);
          float x[4] = {v.x, v.y, v.z, v.w};
          const int kbase = kb * BK + c * 4;
          if (fold) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float dlt = (kbase + i < p.K) ? x[i] - x0 : 0.f;
              s1 += dlt;
              s2 = fmaf(dlt, dlt, s2);
            }
          }
          if (p.prologue == PRO_AFFINE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = kbase + i;
              x[i] = (k < p.K && m < p.M) ? x[i] * p.pg[k] + p.pb[k] : 0.f;
            }
          }
          if (NPASS == 3) {
            float h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              h[i] = __uint_as_float((__float_as_uint(x[i]) + 0x1000u) & 0xffffe000u);     // nearest tf32
              l[i] = __uint_as_float((__float_as_uint(x[i] - h[i]) + 0x1000u) & 0xffffe000u);
            }
            *reinterpret_cast<float4*>(rh + pc) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(rl + pc) = make_float4(l[0], l[1], l[2], l[3]);
          } else if (p.prologue == PRO_AFFINE) {
            *reinterpret_cast<float4*>(rh + pc) = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
        if (NPASS == 3 || p.prologue == PRO_AFFINE)
          fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core (async proxy)
      }
      mbar_arrive(&conv_full[s]);
    }
    float ln_mean = 0.f, ln_rstd = 1.f;
    if (fold) {
      const float invK = 1.0f / (float)p.K;
      const float ds = s1 * invK;
      ln_mean = x0 + ds;
      const float var = fmaxf(s2 * invK - ds * ds, 0.f);
      ln_rstd = 1.0f / sqrtf(var + p.eps);
    }

    // ===================================================================== epilogue
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool row_ok = m < p.M;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.residual || (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)));
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);     // warp-collective: all lanes participate
      const int nb = n0 + c0;
      if (row_ok && nb < p.N) {
        float* crow = p.C + (int64_t)m * p.ldc + nb;
        const float* rrow = p.residual ? p.residual + (int64_t)m * p.ldr + nb : nullptr;
        if (nb + 32 <= p.N) {
          // ---- full 32-column block: everything with compile-time register indices
          if (fold) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 cv = __ldg(reinterpret_cast<const float4*>(p.ln_cvec + nb + j));
              const float4 dv = __ldg(reinterpret_cast<const float4*>(p.ln_dvec + nb + j));
              v[j] = ln_rstd * (v[j] - ln_mean * cv.x) + dv.x;
              v[j + 1] = ln_rstd * (v[j + 1] - ln_mean * cv.y) + dv.y;
              v[j + 2] = ln_rstd * (v[j + 2] - ln_mean * cv.z) + dv.z;
              v[j + 3] = ln_rstd * (v[j + 3] - ln_mean * cv.w) + dv.w;
            }
          } else if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
              v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
            }
          }
          if (nb + 32 <= p.div_ncols) {                       // q / sqrt(dh): IEEE division like the reference
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fdiv_rn(v[j], p.div_by);
          } else if (nb < p.div_ncols) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (nb + j < p.div_ncols) ? __fdiv_rn(v[j], p.div_by) : v[j];
          }
          if (p.relu == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (p.relu == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
          }
          if (vec_ok) {
            if (rrow) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r = *reinterpret_cast<const float4*>(rrow + j);
                v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) crow[j] = v[j] + (rrow ? rrow[j] : 0.f);
          }
        } else {
          // ---- ragged tail block (N not a multiple of 32): predicated scalar path, still unrolled
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            if (n < p.N) {
              float x = v[j];
              if (fold) x = ln_rstd * (x - ln_mean * __ldg(p.ln_cvec + n)) + __ldg(p.ln_dvec + n);
              else if (p.bias) x += __ldg(p.bias + n);
              if (n < p.div_ncols) x = __fdiv_rn(x, p.div_by);
              if (p.relu == 1) x = fmaxf(x, 0.f); else if (p.relu == 2) x = tanhf(x);
              if (rrow) x += rrow[j];
              crow[j] = x;
            }
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------ host side
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

template <int BN, int NPASS>
cudaError_t launch(const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN, NPASS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tmA, tmWhi, tmWlo;
  if (!make_map(&tmA, p.A, p.M, p.K, p.lda, BM)) return cudaErrorInvalidValue;
  if (!make_map(&tmWhi, p.W, p.N, p.K, p.ldw, BN)) return cudaErrorInvalidValue;
  if (!make_map(&tmWlo, NPASS == 3 ? p.W_lo : p.W, p.N, p.K, p.ldw, BN)) return cudaErrorInvalidValue;
  const int64_t tiles = (int64_t)cdiv(p.N, BN) * cdiv(p.M, BM);
  gemm_tc_kernel<BN, NPASS><<<(unsigned)tiles, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmWhi, tmWlo, p);
  return cudaGetLastError();
}

}  // namespace

bool gemm_tc_available(const char** why) {
  const bool ok = lookup();
  if (why) *why = g_why;
  return ok;
}

cudaError_t gemm_tc(const GemmParams& p, int npass, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  if (!lookup()) return cudaErrorNotSupported;
  // TMA needs 16-byte aligned bases and row pitches
  if ((p.lda & 3) || (p.ldw & 3) || (reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.W) & 15) ||
      (npass == 3 && (p.W_lo == nullptr || (reinterpret_cast<uintptr_t>(p.W_lo) & 15))))
    return cudaErrorInvalidValue;
  // tile width: fill the 148 SMs when M is small (decode), wide tiles when M is large (encoder)
  const int64_t tiles128 = (int64_t)cdiv(p.N, 128) * cdiv(p.M, BM);
  const bool narrow = (p.N <= 64) || (tiles128 < 148 && p.N % 128 != 0) || (tiles128 < 74);
  if (npass == 3) return narrow ? launch<64, 3>(p, stream) : launch<128, 3>(p, stream);
  return narrow ? launch<64, 1>(p, stream) : launch<128, 1>(p, stream);
}

}  // namespace nd
