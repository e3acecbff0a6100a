// Persistent LSTM recurrence for sm_100a (packed-sequence semantics of nn.LSTM, one layer,
// forward and reverse directions concurrently).  Reference: encoder/nano_encoder.py:97-99 and
// encoder/rnn_encoder.py:70-78 (cuDNN / ATen LSTM over a PackedSequence).
//
// Mapping: one work item = (tile of BT chunks, direction).  An item is executed by a thread-block
// cluster of C CTAs; CTA `rank` owns hidden units [rank*H/C, (rank+1)*H/C).  Each thread owns ONE
// gate row of W_hh and keeps it in registers for all T steps (weights-stationary; fp32 exact),
// the hidden state of the tile lives in shared memory (double buffered) and is exchanged between
// the CTAs of the cluster through distributed shared memory once per step.
//   per step:  gates[row][b] = xg[b][t][row] + b_hh[row] + sum_k W_hh[row][k] * h[b][k]
//              c = sig(f)*c + sig(i)*tanh(g);  h = sig(o)*tanh(c)          (gate order i,f,g,o)
#include <cooperative_groups.h>

#include "lstm.cuh"

namespace cg = cooperative_groups;

namespace nd {

namespace {

template <int H, int C, int BT>
__global__ void __launch_bounds__(4 * (H / C), 1) lstm_kernel(LstmParams p) {
  constexpr int UPC = H / C;          // hidden units owned by this CTA
  constexpr int NT = 4 * UPC;         // one thread per gate row
  constexpr int PPT = BT / 4;         // (unit, chunk) pairs per thread in the pointwise phase
  static_assert(BT % 4 == 0, "BT must be a multiple of 4");

  __shared__ __align__(16) float h_buf[2][BT][H];
  __shared__ float gates[BT][NT];
  __shared__ int s_len[BT];

  const int tid = threadIdx.x;
  int rank = 0;
  if constexpr (C > 1) rank = (int)cg::this_cluster().block_rank();
  const int item = blockIdx.x / C;
  const int dir = item % p.dirs;
  const int tile = item / p.dirs;
  const int b0 = tile * BT;

  const int gate = tid / UPC, u = tid % UPC;
  const int grow = gate * H + rank * UPC + u;                 // row in the [4H] gate dimension
  const int64_t wrow = (int64_t)dir * 4 * H + grow;

  float w[H];
#pragma unroll
  for (int k = 0; k < H; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p.w_hh + wrow * H + k);
    w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
  }
  const float bhh = p.b_hh[wrow];
  const float wi0 = p.x0 ? p.w_ih0[wrow] : 0.f;
  const float bi0 = p.x0 ? p.b_ih0[wrow] : 0.f;

  if (tid < BT) s_len[tid] = (b0 + tid < p.B) ? (int)p.lengths[b0 + tid] : 0;
  for (int i = tid; i < 2 * BT * H; i += NT) (&h_buf[0][0][0])[i] = 0.f;
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
  for (int s = 0; s < maxlen; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    // ---- input-side gate term (independent of the recurrence: its latency hides behind the FMAs)
    float xin[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const int len = s_len[b];
      const bool active = s < len;
      const int t = active ? (dir == 0 ? s : len - 1 - s) : 0;
      const int bb = min(b0 + b, p.B - 1);
      if (p.x0) {
        xin[b] = p.x0[(int64_t)bb * p.T + t] * wi0 + bi0;
      } else {
        xin[b] = p.xg[((int64_t)bb * p.T + t) * p.xg_ld + wrow];
      }
    }
    // ---- recurrent matvec: this thread's gate row against the tile's hidden states
    float acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = bhh;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(&h_buf[cur][b][k]);   // warp-broadcast
        acc[b] = fmaf(w[k], hv.x, acc[b]);
        acc[b] = fmaf(w[k + 1], hv.y, acc[b]);
        acc[b] = fmaf(w[k + 2], hv.z, acc[b]);
        acc[b] = fmaf(w[k + 3], hv.w, acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) gates[b][tid] = xin[b] + acc[b];
    __syncthreads();
    // ---- pointwise cell update; thread owns pairs (unit uu, chunk b)
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int pidx = tid + j * NT;
      const int uu = pidx % UPC, b = pidx / UPC;
      const int len = s_len[b];
      if (s < len) {
        const float ig = sigmoid_acc(gates[b][0 * UPC + uu]);
        const float fg = sigmoid_acc(gates[b][1 * UPC + uu]);
        const float gg = tanhf(gates[b][2 * UPC + uu]);
        const float og = sigmoid_acc(gates[b][3 * UPC + uu]);
        c_state[j] = fg * c_state[j] + ig * gg;
        h_state[j] = og * tanhf(c_state[j]);
        const int t = dir == 0 ? s : len - 1 - s;
        p.out[((int64_t)(b0 + b) * p.T + t) * out_ld + dir * H + rank * UPC + uu] = h_state[j];
      }
      float* dst = &h_buf[nxt][b][rank * UPC + uu];
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
      const int pidx = tid + j * NT;
      const int uu = pidx % UPC, b = pidx / UPC;
      if (b0 + b < p.B) {
        const int64_t o = ((int64_t)dir * p.B + b0 + b) * H + rank * UPC + uu;
        p.h_n[o] = h_state[j];
        p.c_n[o] = c_state[j];
      }
    }
  }
  if constexpr (C > 1) cg::this_cluster().sync();     // no CTA may exit while peers still write its smem
}

template <int H, int C, int BT>
cudaError_t launch(const LstmParams& p, cudaStream_t stream) {
  const int tiles = cdiv(p.B, BT);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * p.dirs * C));
  cfg.blockDim = dim3(4 * (H / C));
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lstm_kernel<H, C, BT>, p);
}

template <int H, int C>
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
    case 4: return launch<H, C, 4>(p, stream);
    case 8: return launch<H, C, 8>(p, stream);
    case 12: return launch<H, C, 12>(p, stream);
    default: return launch<H, C, 16>(p, stream);
  }
}

}  // namespace

bool lstm_supported(int H) { return H == 16 || H == 32 || H == 64 || H == 128; }

cudaError_t lstm_layer(const LstmParams& p, int n_sm, cudaStream_t stream) {
  if (p.B <= 0) return cudaSuccess;
  switch (p.H) {
    case 16: return pick_bt<16, 1>(p, n_sm, stream);
    case 32: return pick_bt<32, 1>(p, n_sm, stream);
    case 64: return pick_bt<64, 1>(p, n_sm, stream);
    case 128: return pick_bt<128, 2>(p, n_sm, stream);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace nd
