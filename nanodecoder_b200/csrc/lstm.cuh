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
};

bool lstm_supported(int H);
cudaError_t lstm_layer(const LstmParams& p, int n_sm, cudaStream_t stream);

}  // namespace nd
