// Persistent LSTM recurrence for sm_100a (packed-sequence semantics of nn.LSTM, one layer,
// forward and reverse directions concurrently).  Reference: encoder/nano_encoder.py:97-99 and
// encoder/rnn_encoder.py:70-78 (cuDNN / ATen LSTM over a PackedSequence).
//
// Mapping: one work item = (tile of BT chunks, direction).  An item is executed by a thread-block
// cluster of C CTAs; CTA `rank` owns hidden units [rank*H/C, (rank+1)*H/C).  W_hh stays in
// registers for all T steps (weights-stationary; fp32 exact), the hidden state of the tile lives
// in shared memory (double buffered) and is exchanged between the CTAs of the cluster through
// distributed shared memory once per step.
//   per step:  gates[row][b] = xg[b][t][row] + b_hh[row] + sum_k W_hh[row][k] * h[b][k]
//              c = sig(f)*c + sig(i)*tanh(g);  h = sig(o)*tanh(c)          (gate order i,f,g,o)
#include <cooperative_groups.h>

#include "lstm.cuh"

namespace cg = cooperative_groups;

namespace nd {

namespace {

// Thread layout: tid = u*4 + kq.  Thread (u, kq) holds, for hidden unit u of this CTA, the K-quarter
// [kq*H/4, (kq+1)*H/4) of all FOUR gate rows (i,f,g,o) in registers (4 * H/4 = H weights), so every
// hidden-state value fetched from shared memory feeds four FMAs (the shared-memory -> register
// return path, not the FMA pipe, bounds a one-row-per-thread layout).  The four K-quarter partial
// sums are combined with a 2-stage reduce-scatter over xor-shuffles, after which lane kq owns the
// complete gates of chunks [kq*BT/4, (kq+1)*BT/4) of the tile and performs their cell update.
// one stage of the reduce-scatter over the K slices: lanes whose kq bit W is set keep the upper N values
template <int N, int W>
__device__ __forceinline__ void rs_stage(float* v, int kq) {
  const bool up = (kq & W) != 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float keep = up ? v[i + N] : v[i];
    const float send = up ? v[i] : v[i + N];
    v[i] = keep + __shfl_xor_sync(ND_FULL, send, W);
  }
}

// KSPL = K slices per hidden unit (4; 8 for H = 256, where a quarter of W_hh's four gate rows would be 256 registers)
// NG = gate rows per hidden unit: 4 = LSTM (i, f, g, o), 3 = GRU (r, z, n; nn.GRU / nn.GRUCell:
//      r = sig(x_r + h_r), z = sig(x_z + h_z), n = tanh(x_n + r * h_n), h' = n + z * (h - n) with h_* = W_h* h + b_h*)
template <int H, int C, int BT, int KSPL = 4, int NG = 4>
__global__ void __launch_bounds__(KSPL * (H / C), 1) lstm_kernel(LstmParams p) {
  constexpr int UPC = H / C;          // hidden units owned by this CTA
  constexpr int NT = KSPL * UPC;
  constexpr int KQ = H / KSPL;        // K elements per thread
  constexpr int KS = KQ + 4;          // padded slice stride (floats): conflict-free 128-bit slice reads
  constexpr int PPT = BT / KSPL;      // chunks whose cell update this lane performs
  static_assert(BT % KSPL == 0 && KQ % 4 == 0 && (KSPL == 4 || KSPL == 8), "tile shape");

  __shared__ __align__(16) float h_buf[2][BT][KSPL * KS];
  __shared__ int s_len[BT];

  const int tid = threadIdx.x;
  int rank = 0;
  if constexpr (C > 1) rank = (int)cg::this_cluster().block_rank();
  const int item = blockIdx.x / C;
  const int dir = item % p.dirs;
  const int tile = item / p.dirs;
  const int b0 = tile * BT;
  const int u = tid / KSPL, kq = tid % KSPL;
  const int ucol = rank * UPC + u;                       // hidden unit index in [0, H)
  const int64_t gbase = (int64_t)dir * NG * H + ucol;    // + g*H = row of gate g in the [dirs*NG*H] dimension

  float w[NG][KQ];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int k = 0; k < KQ; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(p.w_hh + (gbase + g * H) * H + kq * KQ + k);
      w[g][k] = v.x; w[g][k + 1] = v.y; w[g][k + 2] = v.z; w[g][k + 3] = v.w;
    }
  float bhh[NG], wi0[NG], bi0[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    bhh[g] = p.b_hh[gbase + g * H];
    wi0[g] = p.x0 ? p.w_ih0[gbase + g * H] : 0.f;
    bi0[g] = p.x0 ? p.b_ih0[gbase + g * H] : 0.f;
  }

  if (tid < BT) s_len[tid] = (b0 + tid < p.B) ? (int)p.lengths[b0 + tid] : 0;
  for (int i = tid; i < 2 * BT * KSPL * KS; i += NT) (&h_buf[0][0][0])[i] = 0.f;
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int b = 0; b < BT; ++b) maxlen = max(maxlen, s_len[b]);
  maxlen = min(maxlen, p.T);

  float c_state[PPT], h_state[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) { c_state[j] = 0.f; h_state[j] = 0.f; }

  if constexpr (C > 1) cg::this_cluster().sync();

  const int out_ld = p.dirs * H;
  const int hoff = (ucol / KQ) * KS + (ucol % KQ);       // where unit ucol lives inside a padded h row
  for (int s = 0; s < maxlen; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    // ---- input-side gate terms of the chunks this lane finishes (latency hides behind the FMAs)
    float xin[NG][PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int b = kq * PPT + j;
      const int len = s_len[b];
      const bool active = s < len;
      const int t = active ? (dir == 0 ? s : len - 1 - s) : 0;
      const int bb = min(b0 + b, p.B - 1);
      if (p.x0) {
        const float xv = p.x0[(int64_t)bb * p.T + t];
#pragma unroll
        for (int g = 0; g < NG; ++g) xin[g][j] = xv * wi0[g] + bi0[g];
      } else {
        const float* xr = p.xg + ((int64_t)bb * p.T + t) * p.xg_ld + gbase;
#pragma unroll
        for (int g = 0; g < NG; ++g) xin[g][j] = xr[g * H];
      }
    }
    // ---- partial recurrent products over this thread's K-quarter, all BT chunks
    float acc[NG][BT];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[g][b] = 0.f;
#pragma unroll
    for (int k = 0; k < KQ; k += 4) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(&h_buf[cur][b][kq * KS + k]);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          acc[g][b] = fmaf(w[g][k], hv.x, acc[g][b]);
          acc[g][b] = fmaf(w[g][k + 1], hv.y, acc[g][b]);
          acc[g][b] = fmaf(w[g][k + 2], hv.z, acc[g][b]);
          acc[g][b] = fmaf(w[g][k + 3], hv.w, acc[g][b]);
        }
      }
    }
    // ---- reduce-scatter over the KSPL K slices: lane kq ends up with chunks [kq*PPT, (kq+1)*PPT).  Every stage
    // halves the chunk range a lane keeps (upper half when its kq bit is set) and adds the partner's partial sums.
    float fin[NG][PPT];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if constexpr (KSPL == 8) rs_stage<4 * PPT, 4>(acc[g], kq);
      rs_stage<2 * PPT, 2>(acc[g], kq);
      rs_stage<PPT, 1>(acc[g], kq);
#pragma unroll
      for (int j = 0; j < PPT; ++j) fin[g][j] = acc[g][j];
    }
    // ---- cell update of (unit ucol, chunk kq*PPT + j)
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int b = kq * PPT + j;
      const int len = s_len[b];
      if (s < len) {
        if constexpr (NG == 4) {
          const float ig = sigmoid_acc(xin[0][j] + (fin[0][j] + bhh[0]));
          const float fg = sigmoid_acc(xin[1][j] + (fin[1][j] + bhh[1]));
          const float gg = tanhf(xin[2][j] + (fin[2][j] + bhh[2]));
          const float og = sigmoid_acc(xin[3][j] + (fin[3][j] + bhh[3]));
          c_state[j] = fg * c_state[j] + ig * gg;
          h_state[j] = og * tanhf(c_state[j]);
        } else {
          const float rg = sigmoid_acc(xin[0][j] + (fin[0][j] + bhh[0]));
          const float zg = sigmoid_acc(xin[1][j] + (fin[1][j] + bhh[1]));
          const float ng = tanhf(xin[2][j] + rg * (fin[2][j] + bhh[2]));
          h_state[j] = ng + zg * (h_state[j] - ng);
        }
        const int t = dir == 0 ? s : len - 1 - s;
        p.out[((int64_t)(b0 + b) * p.T + t) * out_ld + dir * H + ucol] = h_state[j];
      }
      float* dst = &h_buf[nxt][b][hoff];
      if constexpr (C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int r = 0; r < C; ++r) *cluster.map_shared_rank(dst, r) = h_state[j];
      } else {
        *dst = h_state[j];
      }
    }
    if constexpr (C > 1) cg::this_cluster().sync(); else __syncthreads();
  }

  if (p.h_n) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int b = kq * PPT + j;
      if (b0 + b < p.B) {
        const int64_t o = ((int64_t)dir * p.B + b0 + b) * H + ucol;
        p.h_n[o] = h_state[j];
        if (p.c_n) p.c_n[o] = c_state[j];
      }
    }
  }
  if constexpr (C > 1) cg::this_cluster().sync();     // no CTA may exit while peers still write its smem
}

template <int H, int C, int BT, int KSPL = 4, int NG = 4>
cudaError_t launch(const LstmParams& p, cudaStream_t stream) {
  const int tiles = cdiv(p.B, BT);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * p.dirs * C));
  cfg.blockDim = dim3(KSPL * (H / C));
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lstm_kernel<H, C, BT, KSPL, NG>, p);
}

template <int H, int C, int NG = 4>
cudaError_t pick_bt(const LstmParams& p, int n_sm, cudaStream_t stream) {
  // cost model: waves x per-step cycles (FMA issue ~ 2.5*H*BT per SMSP pair, ~700 cycles of sync)
  int best = 4;
  double best_cost = 1e30;
  const int cand[4] = {4, 8, 12, 16};
  for (int i = 0; i < 4; ++i) {
    const int bt = cand[i];
    const int64_t ctas = (int64_t)cdiv(p.B, bt) * p.dirs * C;
    const double waves = (double)cdiv64(ctas, n_sm);
    const double cost = waves * (2.5 * H * bt + 700.0);
    if (cost < best_cost) { best_cost = cost; best = bt; }
  }
  switch (best) {
    case 4: return launch<H, C, 4, 4, NG>(p, stream);
    case 8: return launch<H, C, 8, 4, NG>(p, stream);
    case 12: return launch<H, C, 12, 4, NG>(p, stream);
    default: return launch<H, C, 16, 4, NG>(p, stream);
  }
}

}  // namespace

// H = 256 (unidirectional rnn encoder at d = 256): 8 K slices per unit (32 x 4 weight registers per thread), clusters
// of 8 CTAs of 32 units each
cudaError_t launch_h256(const LstmParams& p, int n_sm, cudaStream_t stream) {
  const int64_t c8 = (int64_t)cdiv(p.B, 8) * p.dirs * 8, c16 = (int64_t)cdiv(p.B, 16) * p.dirs * 8;
  const double cost8 = (double)cdiv64(c8, n_sm) * (2.5 * 256 * 8 + 700.0);
  const double cost16 = (double)cdiv64(c16, n_sm) * (2.5 * 256 * 16 + 700.0);
  if (p.cell == 1) return cost8 <= cost16 ? launch<256, 8, 8, 8, 3>(p, stream) : launch<256, 8, 16, 8, 3>(p, stream);
  return cost8 <= cost16 ? launch<256, 8, 8, 8>(p, stream) : launch<256, 8, 16, 8>(p, stream);
}

bool lstm_supported(int H) { return H == 16 || H == 32 || H == 64 || H == 128 || H == 256; }

cudaError_t lstm_layer(const LstmParams& p, int n_sm, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  if (p.cell == 1) {                                   // GRU
    switch (p.H) {
      case 16: return pick_bt<16, 1, 3>(p, n_sm, stream);
      case 32: return pick_bt<32, 1, 3>(p, n_sm, stream);
      case 64: return pick_bt<64, 1, 3>(p, n_sm, stream);
      case 128: return pick_bt<128, 2, 3>(p, n_sm, stream);
      case 256: return launch_h256(p, n_sm, stream);
      default: return cudaErrorNotSupported;
    }
  }
  switch (p.H) {
    case 16: return pick_bt<16, 1>(p, n_sm, stream);
    case 32: return pick_bt<32, 1>(p, n_sm, stream);
    case 64: return pick_bt<64, 1>(p, n_sm, stream);
    case 128: return pick_bt<128, 2>(p, n_sm, stream);
    case 256: return launch_h256(p, n_sm, stream);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace nd
