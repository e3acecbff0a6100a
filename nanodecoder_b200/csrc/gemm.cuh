// Dense projection interface shared by the SIMT and tcgen05 GEMM kernels.
//   C[M,N] = epilogue( prologue(A)[M,K] . W[N,K]^T )
// A, W, C row-major fp32 (W is an nn.Linear weight: K contiguous).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace nd {

enum { PRO_NONE = 0, PRO_LAYERNORM = 1, PRO_AFFINE = 2 };

struct GemmParams {
  const float* A = nullptr;  int64_t lda = 0;
  const float* W = nullptr;  int64_t ldw = 0;   // hi part (or the full fp32 weight)
  const float* W_lo = nullptr;                   // tf32 residual W - tf32(W) (3xTF32 mode only)
  const float* bias = nullptr;                   // [N]
  float* C = nullptr;        int64_t ldc = 0;
  const float* residual = nullptr; int64_t ldr = 0;   // added after the activation
  int M = 0, N = 0, K = 0;
  // prologue applied to A while it is staged:
  //   PRO_LAYERNORM: (a - mean_row) * rstd_row * pg[k] + pb[k]   (nn.LayerNorm, eps)
  //   PRO_AFFINE:    a * pg[k] + pb[k]                          (eval BatchNorm1d folded to alpha/beta)
  int prologue = PRO_NONE;
  const float* pg = nullptr; const float* pb = nullptr; float eps = 1e-6f;
  int relu = 0;                                  // 0 none, 1 ReLU, 2 tanh (SIMT path only), 3 ReLU AFTER the residual add
  // columns [0, div_ncols) are divided by div_by after the bias (q / sqrt(dh), multi_headed_attn.py:167)
  float div_by = 1.0f; int div_ncols = 0;
  // tcgen05 path only — LayerNorm folded around the GEMM instead of applied to A:
  //   LN(a) . W^T + bias = rstd_m * (a . (W*g)^T - mean_m * cvec[n]) + dvec[n]
  //   with W given as W*g (gamma folded into the columns), cvec[n] = sum_k W[n,k] g[k],
  //   dvec[n] = sum_k W[n,k] beta[k] + bias[n].  Row mean / rstd are accumulated from the A tiles
  //   as they stream through shared memory (no extra pass over A).
  const float* ln_cvec = nullptr; const float* ln_dvec = nullptr;
  // optional device counter: when it reads 0 the whole GEMM returns at once (beam search: every chunk of the batch
  // has retired, the remaining steps of the captured loop only cost their launches)
  const int* alive = nullptr;
  long long* dbg = nullptr;      // optional [32] clock64() timeline written by CTA 0 (tuning aid)
};

// both return cudaError_t of the launch
cudaError_t gemm_simt(const GemmParams& p, cudaStream_t stream);
// npass: 3 = 3xTF32 (fp32 parity), 1 = single TF32
cudaError_t gemm_tc(const GemmParams& p, int npass, cudaStream_t stream);
void gemm_tc_set_persistent(int mode);         // large tile counts: 2 (default) persistent kernel with A in tensor memory,
                                               // 1 persistent kernel with A in shared memory, 0 one tile per CTA
void gemm_tc_set_serial_split(int on);         // 1 (default): >= 2048 rows run the split-K sum inside one CTA (same bits)
void gemm_tc_set_wide_wave(int on);            // 1 (default): 128-wide tiles when 64-wide ones would need a second, mostly empty wave
void gemm_tc_set_a_tmem(int on);               // 1 (default): one-tile-per-CTA 3xTF32 kernels keep the A operand in tensor memory
void gemm_tc_set_debug(long long* dev_buf);   // timeline buffer for subsequent gemm_tc launches (nullptr = off)
// one-time driver entry-point lookup for tensor-map encoding; returns false if unavailable
bool gemm_tc_available(const char** why);

// un-swizzled 2-D tensor map over row-major fp32 [rows, cols] (row pitch ld elements) with a [box_rows, box_cols] box:
// rows land densely in shared memory (box_cols * 4 bytes apart).  Used by the ring-fed attention kernel.
bool make_plain_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int box_cols);

// split an fp32 weight into tf32 hi (low 13 mantissa bits cleared) and lo = w - hi
void split_tf32_host(const float* w, float* hi, float* lo, size_t n);

}  // namespace nd
