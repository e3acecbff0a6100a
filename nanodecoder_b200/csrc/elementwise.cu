// Small memory-bound kernels: normalisation, embedding, layout changes, generator + token selection.
#include <float.h>
#include <math.h>

#include "kernels.cuh"

namespace nd {

namespace {

// ---------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                 const float* __restrict__ b, float eps, float* __restrict__ y,
                                 int64_t M, int d) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) s += xr[k];
  const float mean = warp_sum(s) / (float)d;
  float v = 0.f;
  for (int k = lane; k < d; k += 32) { const float t = xr[k] - mean; v += t * t; }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)d + eps);
  float* yr = y + row * d;
  for (int k = lane; k < d; k += 32) yr[k] = (xr[k] - mean) * rstd * g[k] + b[k];
}

__global__ void embed_kernel(const int* __restrict__ tok, const float* __restrict__ emb, float* __restrict__ x,
                             int64_t x_ld, int rows, int d, const float* __restrict__ pe_row, float emb_scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int t = tok[row];
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = emb[(int64_t)t * d + c];
    if (pe_row) v = v * emb_scale + pe_row[c];     // onmt/modules/embeddings.py:36-41: emb * sqrt(d) + pe[step]
    x[(int64_t)row * x_ld + c] = v;
  }
}

// four columns per thread, one 128-bit store (the scalar form wrote 1 GB at 1.4 TB/s)
__global__ void linear_in1_vec_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ y, int64_t n, int d4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d4) return;
  const int64_t r = i / d4;
  const int c = (int)(i - r * d4);
  const float xv = x[r];
  const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + c);
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c);
  reinterpret_cast<float4*>(y)[i] = make_float4(xv * wv.x + bv.x, xv * wv.y + bv.y, xv * wv.z + bv.z, xv * wv.w + bv.w);
}

__global__ void linear_in1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ y, int64_t n, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const int64_t r = i / d;
  const int c = (int)(i - r * d);
  y[i] = x[r] * w[c] + bias[c];
}

__global__ void maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int d,
                               int stride) {
  const int Tp = T / stride;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Tp * d) return;
  const int c = (int)(i % d);
  const int tp = (int)((i / d) % Tp);
  const int b = (int)(i / ((int64_t)d * Tp));
  float m = -FLT_MAX;
  for (int j = 0; j < stride; ++j) m = fmaxf(m, in[((int64_t)b * T + tp * stride + j) * d + c]);
  out[i] = m;
}

__global__ void pool_lengths_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, int B, int stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out[i] = (int64_t)floor((double)(in[i] - stride) / (double)stride + 1.0);      // nano_encoder.py:103-104
}

__global__ void transpose_bt_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int d) {
  // one CTA per (b, t) row of d floats
  const int64_t r = blockIdx.x;
  const int b = (int)(r / T), t = (int)(r % T);
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    out[((int64_t)t * B + b) * d + c] = in[r * d + c];
}

__global__ void gather_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                   const int* __restrict__ parent, int row0, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = row0 + blockIdx.x;
  const int s = parent[r];
  for (int c = threadIdx.x; c < width; c += blockDim.x) dst[(int64_t)r * width + c] = src[(int64_t)s * width + c];
}

__global__ void lstm_cell_kernel(const float* __restrict__ ga, const float* __restrict__ gb,
                                 const float* __restrict__ c_in, float* __restrict__ h_out,
                                 float* __restrict__ c_out, int rows, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * d) return;
  const int64_t r = i / d;
  const int c = (int)(i - r * d);
  const float* a = ga + r * 4 * d;
  const float* b = gb + r * 4 * d;
  const float ig = sigmoid_acc(a[c] + b[c]);
  const float fg = sigmoid_acc(a[d + c] + b[d + c]);
  const float gg = tanhf(a[2 * d + c] + b[2 * d + c]);
  const float og = sigmoid_acc(a[3 * d + c] + b[3 * d + c]);
  const float cy = fg * c_in[i] + ig * gg;
  c_out[i] = cy;
  h_out[i] = og * tanhf(cy);
}

// nn.GRUCell (onmt/models/stacked_rnn.py:39-65): ga = x W_ih^T + b_ih, gb = h W_hh^T + b_hh, rows of [r | z | n]
__global__ void gru_cell_kernel(const float* __restrict__ ga, const float* __restrict__ gb,
                                const float* __restrict__ h_in, float* __restrict__ h_out, int rows, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * d) return;
  const int64_t r = i / d;
  const int c = (int)(i - r * d);
  const float* a = ga + r * 3 * d;
  const float* b = gb + r * 3 * d;
  const float rg = sigmoid_acc(a[c] + b[c]);
  const float zg = sigmoid_acc(a[d + c] + b[d + c]);
  const float ng = tanhf(a[2 * d + c] + rg * b[2 * d + c]);
  h_out[i] = ng + zg * (h_in[i] - ng);
}

__global__ void fill_int_kernel(int* p, int n, int value) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = value;
}

__global__ void glu_residual_kernel(const float* __restrict__ y, const float* __restrict__ x,
                                    float* __restrict__ out, float* __restrict__ glu_out, int64_t rows, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const int64_t r = i / d;
  const int c = (int)(i - r * d);
  const float a = y[r * 2 * d + c], g = y[r * 2 * d + d + c];
  const float glu = a * sigmoid_acc(g);
  if (glu_out) glu_out[i] = glu;
  if (out) out[i] = (x[i] + glu) * 0.70710678118654757f;   // SCALE_WEIGHT = 0.5 ** 0.5 rounded to fp32
}

// all pointers are row-range views starting at row0, except prev (whole buffer: parents are global rows)
__global__ void avg_cumulate_kernel(const float* __restrict__ xn, const float* __restrict__ prev, const int* __restrict__ parent,
                                    int row0, int rows, int d, int step, float* __restrict__ g) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * d) return;
  const int r = (int)(i / d), c = (int)(i - (int64_t)r * d);
  float v = xn[i];
  if (step > 0) {
    const int pr = parent ? parent[row0 + r] : row0 + r;
    v = (v + (float)step * prev[(int64_t)pr * d + c]) / (float)(step + 1);
  }
  g[i] = v;
}

__global__ void avg_gate_kernel(const float* __restrict__ gate, const float* __restrict__ xn, const float* __restrict__ a,
                                const float* __restrict__ x, float* __restrict__ out, int64_t rows, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const int64_t r = i / d;
  const int c = (int)(i - r * d);
  const float ig = gate[r * 2 * d + c], fg = gate[r * 2 * d + d + c];
  const float si = 1.0f / (1.0f + expf(-ig)), sf = 1.0f / (1.0f + expf(-fg));
  out[i] = (si * xn[i] + sf * a[i]) + x[i];
}

// first conv of the ResNet stem: 1 -> 64 channels, 3 taps, folded BatchNorm, ReLU
__global__ void resnet_stem_kernel(const float* __restrict__ src, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ out, int B, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pos = i >> 4;
  if (pos >= (int64_t)B * T) return;
  const int c0 = (int)(i & 15) * 4;
  const int t = (int)(pos % T);
  const float xm = t > 0 ? src[pos - 1] : 0.f, x0 = src[pos], xp = t + 1 < T ? src[pos + 1] : 0.f;
  float4 o;
  float* op = reinterpret_cast<float*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* wc = w + (c0 + j) * 3;
    // conv2d accumulates the three taps in order; the folded bias is added last
    const float acc = fmaf(wc[2], xp, fmaf(wc[1], x0, wc[0] * xm)) + b[c0 + j];
    op[j] = fmaxf(acc, 0.f);
  }
  *reinterpret_cast<float4*>(out + pos * 64 + c0) = o;
}

__global__ void take_column_kernel(const float* __restrict__ x, int ld, int col, float* __restrict__ out, int64_t rows) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) out[r] = x[r * ld + col];
}

__global__ void im2col_kernel(const float* __restrict__ x, float* __restrict__ A, int B, int T, int d, int k,
                              int left) {
  const int64_t r = blockIdx.x;                    // (b, t)
  const int b = (int)(r / T), t = (int)(r % T);
  for (int i = threadIdx.x; i < k * d; i += blockDim.x) {
    const int j = i / d, c = i - j * d;
    const int ts = t + j - left;
    A[r * k * d + i] = (ts >= 0 && ts < T) ? x[((int64_t)b * T + ts) * d + c] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Generator: one warp per row.  VMAX bounds the vocabulary (4 specials + bases <= 16).
template <int VMAX>
__global__ void __launch_bounds__(128) generator_kernel(GenParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  const float* xr = p.x + (int64_t)row * p.x_ld;
  const int d = p.d, V = p.V;
  float mean = 0.f, rstd = 1.f;
  float acc[VMAX];
#pragma unroll
  for (int v = 0; v < VMAX; ++v) acc[v] = 0.f;
  if ((d & 127) == 0 && d <= 1024 && (p.x_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0) {
    // the row lives in registers (d / 128 float4 per lane): ONE round of loads instead of three dependent passes of
    // scalar loads (the kernel is latency bound: 4 warps per CTA, ncu long-scoreboard 17.9 per issue)
    constexpr int MAXQ = 8;
    const int nq = d >> 7;
    float4 xq[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
      xq[i] = i < nq ? *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.ln_g) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXQ; ++i) s += (xq[i].x + xq[i].y) + (xq[i].z + xq[i].w);
      mean = warp_sum(s) / (float)d;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < MAXQ; ++i) {
        if (i < nq) {
          const float a = xq[i].x - mean, b = xq[i].y - mean, c = xq[i].z - mean, e = xq[i].w - mean;
          v += (a * a + b * b) + (c * c + e * e);
        }
      }
      rstd = 1.0f / sqrtf(warp_sum(v) / (float)d + p.eps);
#pragma unroll
      for (int i = 0; i < MAXQ; ++i) {
        if (i < nq) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g) + i * 32 + lane);
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.ln_b) + i * 32 + lane);
          xq[i].x = (xq[i].x - mean) * rstd * g.x + b.x; xq[i].y = (xq[i].y - mean) * rstd * g.y + b.y;
          xq[i].z = (xq[i].z - mean) * rstd * g.z + b.z; xq[i].w = (xq[i].w - mean) * rstd * g.w + b.w;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
      if (v < V) {
        const float4* wr = reinterpret_cast<const float4*>(p.Wg + (int64_t)v * d);
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
          if (i < nq) {
            const float4 w = __ldg(wr + i * 32 + lane);
            acc[v] = fmaf(xq[i].x, w.x, fmaf(xq[i].y, w.y, fmaf(xq[i].z, w.z, fmaf(xq[i].w, w.w, acc[v]))));
          }
        }
      }
    }
  } else {
    if (p.ln_g) {
      float s = 0.f;
      for (int k = lane; k < d; k += 32) s += xr[k];
      mean = warp_sum(s) / (float)d;
      float v = 0.f;
      for (int k = lane; k < d; k += 32) { const float t = xr[k] - mean; v += t * t; }
      rstd = 1.0f / sqrtf(warp_sum(v) / (float)d + p.eps);
    }
    for (int k = lane; k < d; k += 32) {
      float xv = xr[k];
      if (p.ln_g) xv = (xv - mean) * rstd * p.ln_g[k] + p.ln_b[k];
#pragma unroll
      for (int v = 0; v < VMAX; ++v)
        if (v < V) acc[v] = fmaf(xv, __ldg(p.Wg + (int64_t)v * d + k), acc[v]);
    }
  }
  float mx = -FLT_MAX;
#pragma unroll
  for (int v = 0; v < VMAX; ++v) {
    if (v < V) {
      acc[v] = warp_sum(acc[v]) + p.bg[v];
      mx = fmaxf(mx, acc[v]);
    }
  }
  float se = 0.f;
#pragma unroll
  for (int v = 0; v < VMAX; ++v)
    if (v < V) se += expf(acc[v] - mx);
  const float lse = logf(se);
  int best = 0;
  float bestv = -FLT_MAX;
#pragma unroll
  for (int v = 0; v < VMAX; ++v) {
    if (v < V) {
      float lp = acc[v] - mx - lse;                 // log_softmax
      if (lane == 0 && p.trace) p.trace[(int64_t)row * V + v] = lp;
      if (v == p.eos && p.step < p.min_len) lp = -1e20f;      // translator.py:469-470 / 714-715
      if (lane == 0) p.logp[(int64_t)row * V + v] = lp;
      if (lp > bestv) { bestv = lp; best = v; }     // strict >: lowest index wins ties
    }
  }
  if (lane == 0 && p.ids) {
    p.ids[(int64_t)row * p.ids_ld + p.step] = best;
    p.scores[row] = bestv;
    p.next_tok[row] = best;
  }
}

}  // namespace

cudaError_t layernorm_rows(const float* x, const float* g, const float* b, float eps, float* y, int64_t M, int d,
                           cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  layernorm_kernel<<<(unsigned)cdiv64(M, 8), 256, 0, stream>>>(x, g, b, eps, y, M, d);
  return cudaGetLastError();
}

cudaError_t embed_rows(const int* tok, const float* emb, float* x, int64_t x_ld, int rows, int d, const float* pe_row,
                       float emb_scale, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  launch_k(embed_kernel, dim3(rows), dim3(d < 256 ? d : 256), 0, stream, tok, emb, x, x_ld, rows, d, pe_row, emb_scale);
  return cudaGetLastError();
}

cudaError_t linear_in1(const float* x, const float* w, const float* bias, float* y, int64_t n, int d,
                       cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
    linear_in1_vec_kernel<<<(unsigned)cdiv64(n * (d / 4), 256), 256, 0, stream>>>(x, w, bias, y, n, d / 4);
    return cudaGetLastError();
  }
  linear_in1_kernel<<<(unsigned)cdiv64(n * d, 256), 256, 0, stream>>>(x, w, bias, y, n, d);
  return cudaGetLastError();
}

cudaError_t maxpool_time(const float* in, float* out, int B, int T, int d, int stride, cudaStream_t stream) {
  const int64_t n = (int64_t)B * (T / stride) * d;
  if (n <= 0) return cudaSuccess;
  maxpool_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, stream>>>(in, out, B, T, d, stride);
  return cudaGetLastError();
}

cudaError_t pool_lengths(const int64_t* in, int64_t* out, int B, int stride, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  pool_lengths_kernel<<<cdiv(B, 256), 256, 0, stream>>>(in, out, B, stride);
  return cudaGetLastError();
}

cudaError_t transpose_bt(const float* in, float* out, int B, int T, int d, cudaStream_t stream) {
  if ((int64_t)B * T <= 0) return cudaSuccess;
  transpose_bt_kernel<<<(unsigned)((int64_t)B * T), d < 256 ? d : 256, 0, stream>>>(in, out, B, T, d);
  return cudaGetLastError();
}

cudaError_t gather_rows(const float* src, float* dst, const int* parent, int row0, int rows, int width,
                        cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  launch_k(gather_rows_kernel, dim3(rows), dim3(width < 256 ? width : 256), 0, stream, src, dst, parent, row0, width);
  return cudaGetLastError();
}

namespace {
__global__ void cnn_window_kernel(const float* __restrict__ x, float* __restrict__ hist, const int* __restrict__ anc,
                                  int anc_ld, const int* __restrict__ retired, int rows_per_chunk,
                                  float* __restrict__ A, int row0, int t, int k, int d, int Lmax) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = row0 + blockIdx.x;
  if (retired && retired[r / rows_per_chunk]) return;     // retired chunks keep stale ancestor tables
  for (int i = threadIdx.x; i < k * d; i += blockDim.x) {
    const int j = i / d, c = i - j * d;
    const int pos = t - (k - 1) + j;
    float v = 0.f;
    if (pos == t) {
      v = x[(int64_t)r * d + c];
      hist[((int64_t)r * Lmax + t) * d + c] = v;
    } else if (pos >= 0) {
      const int slot = anc ? anc[(int64_t)r * anc_ld + pos] : r;
      v = hist[((int64_t)slot * Lmax + pos) * d + c];
    }
    A[(int64_t)r * k * d + i] = v;
  }
}
__global__ void tanh_kernel(float* __restrict__ x, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = tanhf(x[i]);
}
__global__ void add_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, float s,
                                 float* __restrict__ out, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a[i] + b[i]) * s;
}
__global__ void cnn_combine_kernel(const float* __restrict__ x, const float* __restrict__ c,
                                   const float* __restrict__ o, float s, float* __restrict__ out, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (x[i] + (c[i] + o[i]) * s) * s;
}
}  // namespace

cudaError_t cnn_window(const float* x, float* hist, const int* anc, int anc_ld, const int* retired,
                       int rows_per_chunk, float* A, int row0, int rows, int t, int k, int d, int Lmax,
                       cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  launch_k(cnn_window_kernel, dim3(rows), dim3(256), 0, stream, x, hist, anc, anc_ld, retired, rows_per_chunk, A, row0, t, k, d, Lmax);
  return cudaGetLastError();
}
cudaError_t tanh_inplace(float* x, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  launch_k(tanh_kernel, dim3((unsigned)cdiv64(n, 256)), dim3(256), 0, stream, x, n);
  return cudaGetLastError();
}

cudaError_t add_scale(const float* a, const float* b, float s, float* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  launch_k(add_scale_kernel, dim3((unsigned)cdiv64(n, 256)), dim3(256), 0, stream, a, b, s, out, n);
  return cudaGetLastError();
}
cudaError_t cnn_combine(const float* x, const float* c, const float* o, float s, float* out, int64_t n,
                        cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  launch_k(cnn_combine_kernel, dim3((unsigned)cdiv64(n, 256)), dim3(256), 0, stream, x, c, o, s, out, n);
  return cudaGetLastError();
}

cudaError_t gru_cell_pointwise(const float* gates_a, const float* gates_b, const float* h_in, float* h_out, int rows, int d,
                               cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  launch_k(gru_cell_kernel, dim3((unsigned)cdiv64((int64_t)rows * d, 256)), dim3(256), 0, stream, gates_a, gates_b, h_in, h_out,
           rows, d);
  return cudaGetLastError();
}

cudaError_t lstm_cell_pointwise(const float* gates_a, const float* gates_b, const float* c_in, float* h_out,
                                float* c_out, int rows, int d, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  launch_k(lstm_cell_kernel, dim3((unsigned)cdiv64((int64_t)rows * d, 256)), dim3(256), 0, stream, gates_a, gates_b, c_in, h_out, c_out,
                                                                                rows, d);
  return cudaGetLastError();
}

namespace {
__global__ void transpose_to_dbt_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int d) {
  __shared__ float tile[32][33];
  // in viewed as [B*T, d] -> out [d, B*T]
  const int64_t R = (int64_t)B * T;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < d) ? in[r * d + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (r < R && c < d) out[(int64_t)c * R + r] = tile[threadIdx.x][i];
  }
}
}  // namespace

cudaError_t transpose_to_dbt(const float* in, float* out, int B, int T, int d, cudaStream_t stream) {
  const int64_t R = (int64_t)B * T;
  if (R <= 0) return cudaSuccess;
  dim3 grid((unsigned)cdiv64(R, 32), (unsigned)cdiv(d, 32));
  transpose_to_dbt_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, out, B, T, d);
  return cudaGetLastError();
}

cudaError_t fill_int(int* p, int n, int value, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  launch_k(fill_int_kernel, dim3(cdiv(n, 256)), dim3(256), 0, stream, p, n, value);
  return cudaGetLastError();
}

cudaError_t glu_residual(const float* y, const float* x, float* out, float* glu_out, int64_t rows, int d,
                         cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  launch_k(glu_residual_kernel, dim3((unsigned)cdiv64(rows * d, 256)), dim3(256), 0, stream, y, x, out, glu_out, rows, d);
  return cudaGetLastError();
}

cudaError_t avg_attn_cumulate(const float* xn, const float* prev, const int* parent, int row0, int rows, int d, int step,
                              float* g, cudaStream_t stream) {
  const int64_t n = (int64_t)rows * d;
  if (n <= 0) return cudaSuccess;
  avg_cumulate_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(xn, prev, parent, row0, rows, d, step, g);
  return cudaGetLastError();
}

cudaError_t avg_attn_gate(const float* gate, const float* xn, const float* a, const float* x, float* out, int64_t rows, int d,
                          cudaStream_t stream) {
  const int64_t n = rows * d;
  if (n <= 0) return cudaSuccess;
  avg_gate_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(gate, xn, a, x, out, rows, d);
  return cudaGetLastError();
}

cudaError_t resnet_stem_conv(const float* src, const float* w, const float* b, float* out, int B, int T, cudaStream_t stream) {
  const int64_t n = (int64_t)B * T * 16;            // one thread per (position, 4 channels)
  if (n <= 0) return cudaSuccess;
  resnet_stem_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(src, w, b, out, B, T);
  return cudaGetLastError();
}

cudaError_t take_column(const float* x, int ld, int col, float* out, int64_t rows, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  take_column_kernel<<<(unsigned)cdiv(rows, 256), 256, 0, stream>>>(x, ld, col, out, rows);
  return cudaGetLastError();
}

cudaError_t im2col_time(const float* x, float* A, int B, int T, int d, int k, int left, cudaStream_t stream) {
  if ((int64_t)B * T <= 0) return cudaSuccess;
  im2col_kernel<<<(unsigned)((int64_t)B * T), 256, 0, stream>>>(x, A, B, T, d, k, left);
  return cudaGetLastError();
}

cudaError_t generator_step(const GenParams& p, cudaStream_t stream) {
  if (p.rows <= 0) return cudaSuccess;
  if (p.V > 16) return cudaErrorInvalidValue;
  launch_k(generator_kernel<16>, dim3(cdiv(p.rows, 4)), dim3(128), 0, stream, p);
  return cudaGetLastError();
}

}  // namespace nd
