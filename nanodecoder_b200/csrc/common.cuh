// Shared helpers for the sm_100a kernels of libnanodec.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ND_WARP 32
#define ND_FULL 0xffffffffu

namespace nd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ND_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(ND_FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ND_FULL, v, o);
  return v;
}

// accurate logistic: 1/(1+exp(-x)) with IEEE division (parity with torch.sigmoid on CPU)
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// streaming 128-bit loads that do not pollute L1 (K/V caches are read once per step)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// 256-bit streaming load (sm_100: LDG.E.256): a lane that owns 8 consecutive floats of a row touches its 32-byte
// sector once instead of twice with two 128-bit loads.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg_stream8(const float* p, float* r) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

// ---- programmatic dependent launch (PDL).  Every kernel of the decode step loop starts with
// pdl_launch_dependents() (a dependent launch that follows in the stream may become resident and run its prologue
// now) and executes pdl_wait() before its first access to memory another kernel may have written or may still
// read; pdl_wait() returns once ALL prerequisite grids have completed and flushed.  Both are no-ops for a kernel
// launched without the attribute / without a dependent behind it.
// Which launches carry the attribute is decided by g_pdl (measured, profiles/r01_pdl_experiment.md):
//   2 (default)  only the tcgen05 GEMMs: they need a whole SM each (200 KB of shared memory), so they trickle onto
//                SMs as the preceding kernel frees them and overlap their launch latency, TMEM allocation and
//                weight-tile prefetch with its tail: decode 85.6 -> 83.3 ms
//   1            every step kernel: the 1024-CTA attention kernels become resident next to a running GEMM on the
//                SMs it leaves free and stay badly balanced: 85.4 -> 112 ms
//   3            the GEMMs and the small step kernels (self attention, generator, embedding, beam step, element-wise),
//                not the chunk-per-CTA cross / MLP attention kernels
//   0            plain stream order
#ifndef ND_PDL_EARLY
#define ND_PDL_EARLY 1
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if ND_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern int g_pdl;

// <<<grid, block, smem, stream>>> replacement for kernels that follow the PDL protocol above
template <class... Params, class... Args>
static inline cudaError_t launch_k(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl == 1 || g_pdl == 3) ? 1 : 0;   // g_pdl == 2: only the tcgen05 GEMM launches are dependent launches
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}
// the big HBM-bound attention kernels (one CTA per chunk): dependent launches only under g_pdl == 1
template <class... Params, class... Args>
static inline cudaError_t launch_k_heavy(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl == 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

// Function attributes (dynamic shared-memory limit) are per DEVICE: a launcher's "already set" flag must be too, or the
// second GPU used by one process launches with the default 48 KB limit.  -> slot of the current device in a static
// per-launcher table (returns a reference the launcher flips after cudaFuncSetAttribute succeeded).
struct PerDeviceFlag {
  bool set[64] = {};
  bool& cur() {
    int dev = 0;
    cudaGetDevice(&dev);
    return set[dev & 63];
  }
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nd
