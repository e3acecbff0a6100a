#pragma once
#include "common.cuh"

namespace nd {

struct LstmParams {
  // input side: either the precomputed projection xg (x.W_ih^T + b_ih for every (chunk, t, dir, gate row))
  // or, for input_size == 1, the raw samples x0 with the W_ih column / b_ih applied in-kernel
  const float* xg = nullptr;      // [B, T, dirs*4H]
  int64_t xg_ld = 0;              // dirs*4H
  const float* x0 = nullptr;      // [B, T]
  const float* w_ih0 = nullptr;   // [dirs*4H]
  const float* b_ih0 = nullptr;   // [dirs*4H]
  const float* w_hh = nullptr;    // [dirs*4H, H]
  const float* b_hh = nullptr;    // [dirs*4H]
  const int64_t* lengths = nullptr;   // [B]
  float* out = nullptr;           // [B, T, dirs*H]; positions t >= length must be pre-zeroed by the caller
  float* h_n = nullptr;           // optional [dirs, B, H]
  float* c_n = nullptr;
  int B = 0, T = 0, dirs = 2, H = 0;
  int cell = 0;                   // 0 LSTM (4 gate rows per unit: xg / w_* are [dirs*4H]), 1 GRU (3 gate rows, no c_n)
  long long* dbg = nullptr;       // optional [32] clock64() stamps of step 100 of CTA 0 (tuning aid)
};

bool lstm_supported(int H);
cudaError_t lstm_layer(const LstmParams& p, int n_sm, cudaStream_t stream);       // fp32 FFMA, weights in registers
bool lstm_tc_supported(int H);
cudaError_t lstm_layer_tc(const LstmParams& p, cudaStream_t stream);   // tcgen05 fp16x2-split, W_hh on chip; needs |W_hh| < 6e4
void lstm_tc_set_debug(long long* dev_buf);
void lstm_tc_set_variant(int v);   // 0: W_hh resident in tensor memory (default), 1: in shared memory

}  // namespace nd
