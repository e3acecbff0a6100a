#pragma once
// Fixed-point memory keys / values: how one element of the int16 "hi" plane (+ the uint8 "lo" plane)
// becomes the float holding its unscaled integer.  Shared by the slice kernels (cross_attn_packed.cu) and the
// shared-memory-ring beam kernel (cross_attn_ring.cu).
#include "common.cuh"
#include "kernels.cuh"

namespace nd {

__device__ __forceinline__ int prmt(uint32_t a, uint32_t b, uint32_t sel) {
  int d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// element j of a lane's slice as a float holding the (unscaled) integer.  hi: int16 pairs, lo: uint8 quads.
// prmt selector nibbles: 0-3 = bytes of a (hi pair register), 4-7 = bytes of b (lo quad register), +8 = replicate
// the sign bit of the selected byte.
// Integer -> float without a conversion instruction: for |m| < 2^22 the bit pattern 0x4B400000 + m is the float
// 1.5 * 2^23 + m exactly (one IADD on the integer pipe), and subtracting 1.5 * 2^23 is exact (one FADD).
__device__ __forceinline__ float magic_i2f(int m) { return __int_as_float(m + 0x4B400000) - 12582912.0f; }

__host__ __device__ constexpr bool fmt_has_lo(int fmt) { return fmt == KV_Q24 || fmt == KV_Q23M || fmt == KV_FP24; }
__host__ __device__ constexpr bool fmt_scaled(int fmt) { return fmt != KV_FP24; }

template <int FMT>
__device__ __forceinline__ float unpack_elem(const uint32_t* hi, const uint32_t* lo, int j) {
  const uint32_t k = 2u * (j & 1);                                 // byte index of the int16 inside its register
  if constexpr (FMT == KV_FP24) {
    // the top 24 bits of an fp32: [hi_b1 | hi_b0 | lo | (byte 0 cleared)]
    const uint32_t sel = (4u + (j & 3)) | ((4u + (j & 3)) << 4) | (k << 8) | ((k + 1) << 12);
    return __uint_as_float((uint32_t)prmt(hi[j >> 1], lo[j >> 2], sel) & 0xffffff00u);
  } else if constexpr (fmt_has_lo(FMT)) {
    const uint32_t sel = (4u + (j & 3)) | (k << 4) | ((k + 1) << 8) | (((k + 1) | 8u) << 12);
    const int m = prmt(hi[j >> 1], lo[j >> 2], sel);
    if constexpr (FMT == KV_Q24) return __int2float_rn(m);
    return magic_i2f(m);
  } else {
    const uint32_t sel = k | ((k + 1) << 4) | (((k + 1) | 8u) << 8) | (((k + 1) | 8u) << 12);
    const int m = prmt(hi[j >> 1], 0u, sel);
    if constexpr (FMT == KV_Q16) return __int2float_rn(m);
    return magic_i2f(m);
  }
}

// elements j, j + 1 (j even) of the magic-number formats as a pair: two PRMT, two integer adds, ONE packed fp32 add
template <int FMT>
__device__ __forceinline__ float2 unpack_pair_magic(const uint32_t* hi, const uint32_t* lo, int j) {
  static_assert(FMT == KV_Q23M || FMT == KV_Q15M, "magic-number formats only");
  int m0, m1;
  if constexpr (FMT == KV_Q23M) {
    m0 = prmt(hi[j >> 1], lo[j >> 2], (4u + (j & 3)) | (0u << 4) | (1u << 8) | (9u << 12));
    m1 = prmt(hi[j >> 1], lo[j >> 2], (4u + ((j + 1) & 3)) | (2u << 4) | (3u << 8) | (11u << 12));
  } else {
    m0 = prmt(hi[j >> 1], 0u, 0u | (1u << 4) | (9u << 8) | (9u << 12));
    m1 = prmt(hi[j >> 1], 0u, 2u | (3u << 4) | (11u << 8) | (11u << 12));
  }
  const float2 f = make_float2(__int_as_float(m0 + 0x4B400000), __int_as_float(m1 + 0x4B400000));
  return __fadd2_rn(f, make_float2(-12582912.0f, -12582912.0f));
}

}  // namespace nd
