// Launchers of the non-GEMM kernels (attention, normalisation, token selection, front end).
#pragma once
#include "common.cuh"

namespace nd {

// ---------------------------------------------------------------------------------------------
// Decode-step cross attention over the per-chunk projected memory (decoder/transformer.py:87-91,
// onmt/modules/multi_headed_attn.py:142-190).  One CTA per chunk; the NQ query rows of a chunk
// (1 for greedy, beam_size for beam search) share one pass over K and V.
struct CrossAttnParams {
  const float* q = nullptr;  int64_t q_ld = 0;    // [n_chunks*NQ, d] raw projection; divided by q_div here
  float q_div = 1.0f;                             // sqrt(dh): q / sqrt(dh) as multi_headed_attn.py:167 (IEEE division)
  const float* K = nullptr;                       // [n_chunks, T, *] row stride kv_ld
  const float* V = nullptr;
  int64_t kv_ld = 0;
  const float* src = nullptr;                     // [n_chunks, src_ld]: key t masked iff src == mask_value
  int64_t src_ld = 0;
  float mask_value = 1.0f;
  const int* retired = nullptr;                   // optional [n_chunks]: non-zero -> skip the chunk
  float* ctx = nullptr;      int64_t ctx_ld = 0;  // [n_chunks*NQ, d]
  float* attn = nullptr;                          // optional [n_chunks*NQ, T]: head-0 probabilities
  int n_chunks = 0, NQ = 1, T = 0, d = 0, H = 8;
  // Packed memory keys / values (kv_fmt != 0; K / V / kv_ld are then unused), see kv_pack below:
  //   hi [n_chunks*T, 2d] int16 and (q24 only) lo [n_chunks*T, 2d] uint8, K in columns [0,d), V in [d,2d);
  //   scale [n_chunks*T, 2] = the power-of-two step of the row's K part and V part
  int kv_fmt = 0;
  const int16_t* kv_hi = nullptr;
  const uint8_t* kv_lo = nullptr;
  const float* kv_scale = nullptr;
};
cudaError_t cross_attention(const CrossAttnParams& p, cudaStream_t stream);

// Fixed-point storage of the projected memory keys / values (multi_headed_attn.py:142-153 computes them once per
// chunk; the decode loop then re-reads them L * Ld times: the bytes that bound the whole translate path).
//   KV_Q24: every row part (the d keys, or the d values, of one memory position) is stored as 24-bit signed integers
//           m = rint(x / step), step = 2^(e-23) with 2^e > max|x| of that part: int16 high plane + uint8 low plane
//           + one fp32 step per part.  |x - m*step| <= step/2 = 2^-24 * 2^e: the absolute rounding error fp32 itself has
//           on the part's largest element; 3 bytes per element instead of 4.
//   KV_Q16: m = rint(x / 2^(e-15)) in one int16 plane: |error| <= 2^-16 * 2^e (a bf16 element has 2^-9 |x|);
//           2 bytes per element.  Reduced-precision mode with a stated, measured bound, never the default.
//   KV_Q23M / KV_Q15M: the same with one bit less and the integers rebuilt as floats by exponent arithmetic instead
//           of a conversion instruction; KV_FP24: the top 24 bits of the fp32 value (16 significant bits, no step).
enum { KV_F32 = 0, KV_Q24 = 1, KV_Q16 = 2, KV_Q23M = 3, KV_Q15M = 4, KV_FP24 = 5 };
bool kv_pack_supported(int d);
// kv [rows, 2d] fp32 (row pitch 2d) -> planes as described in CrossAttnParams
// the two parts of a row from separate matrices: part "K" = k[row*ldk .. +d), part "V" = v[row*ldv .. +d)
cudaError_t kv_pack2(const float* k, int64_t ldk, const float* v, int64_t ldv, int64_t rows, int d, int fmt, int16_t* hi,
                     uint8_t* lo, float* scale, cudaStream_t stream);
cudaError_t kv_pack(const float* kv, int64_t rows, int d, int fmt, int16_t* hi, uint8_t* lo, float* scale,
                    cudaStream_t stream);
cudaError_t cross_attention_packed(const CrossAttnParams& p, cudaStream_t stream);   // kv_fmt != 0 (cross_attn_packed.cu)
bool cross_attention_packed_beams_ok(int NQ, int d, int H, int T, int fmt);   // host-side form of the check below
bool cross_attention_packed_mq_supported(const CrossAttnParams& p);   // several queries per chunk (beam search) at d = 256 / 512
void cross_attention_packed_set_fast(int on);   // 1 (default): 256-column-slice kernel for one query per chunk; 0: generic
// several queries per chunk (beam search) at d = 256, H = 8: 2 (default) persistent CTAs fed by a cp.async.bulk ring
// (cross_attn_ring.cu), 1 register-prefetch kernel, 0 generic kernel
void cross_attention_set_beam_kernel(int mode);
bool cross_attention_ring_supported(const CrossAttnParams& p);
bool cross_attention_ring_shape_ok(int NQ, int d, int H, int T, int fmt);   // fmt: KV_F32 / KV_Q23M / KV_Q15M
extern int g_cross_beam_kernel;
cudaError_t cross_attention_ring(const CrossAttnParams& p, cudaStream_t stream);
void cross_attention_ring_set_groups(int g);    // consumer warp groups of the ring kernel (2 default, 1 for > 5 queries)

// Greedy-decode cross attention in MEMORY-BANK space (one query per chunk).  With K = mb Wk^T + bk and
// V = mb Wv^T + bv (multi_headed_attn.py:142-153) the per-head score and context are, in exact arithmetic,
//     q_h . K_h[t]   = mb[t] . qt_h + const_h        qt_h = Wk_h^T q_h / sqrt(dh)      (const_h cancels in softmax)
//     sum_t a_h[t] V_h[t] = Wv_h (sum_t a_h[t] mb[t]) + bv_h                             (sum_t a_h[t] = 1)
// so the kernel reads the memory bank ONCE (T*d*4 bytes per chunk instead of K and V, 2*T*d*4) and returns
// ctxt[h] = sum_t a_h[t] mb[t]; the two d x d products move into the projections around it
// (QT = LN2(x) P^T + pb with P_h = Wk_h^T Wq_h / sqrt(dh);  out = ctxt Mcat^T + mo with M_h = Wo_h Wv_h).
struct CrossMbParams {
  const float* qt = nullptr;                      // [n_chunks, H*d]
  const float* mb = nullptr;                      // [n_chunks, T, d] memory bank
  const float* src = nullptr; int64_t src_ld = 0; // key t masked iff src == mask_value (decoder/transformer.py:219-221)
  float mask_value = 1.0f;
  float* ctxt = nullptr;                          // [n_chunks, H*d]
  int n_chunks = 0, T = 0, d = 0, H = 8;
};
bool cross_attention_mb_supported(int d, int H);
cudaError_t cross_attention_mb(const CrossMbParams& p, cudaStream_t stream);
void cross_attention_mb_set_version(int v);

// Decode-step self attention with a device-resident KV cache (decoder/transformer.py:76-80,
// multi_headed_attn.py:126-141).  qkv holds this step's [q | k | v]; k and v are appended
// to the cache at position `step`.  With beam search the history of row r at position j lives in
// cache row anc[r*anc_ld + j] (parent-pointer indirection instead of reordering the cache).
struct SelfAttnParams {
  const float* qkv = nullptr;                     // [rows, 3d] raw [q | k | v]
  float q_div = 1.0f;                             // q / sqrt(dh)
  float* Kc = nullptr; float* Vc = nullptr;       // [rows, H, Lmax, dh] (head-major: contiguous per head)
  const int* anc = nullptr; int anc_ld = 0;       // optional [rows, Lmax]
  const int* retired = nullptr; int rows_per_chunk = 1;
  float* ctx = nullptr;                           // [rows, d]
  int row0 = 0;                                   // first row of the range handled by this launch (all bases global)
  int rows = 0, d = 0, H = 8, Lmax = 0, step = 0;
};
cudaError_t self_attention_step(const SelfAttnParams& p, cudaStream_t stream);

// Encoder self attention over a whole chunk (encoder/transformer.py:50-52): qkv [B*T, 3d] with
// q divided by sqrt(dh) on load; keys with src == 0.0 are masked (encoder/transformer.py:117-121).
struct EncAttnParams {
  const float* qkv = nullptr;
  float q_div = 1.0f;                             // q / sqrt(dh)
  const float* src = nullptr;                     // [B, T]
  float* ctx = nullptr;                           // [B*T, d]
  int B = 0, T = 0, d = 0, H = 8;
};
cudaError_t encoder_attention(const EncAttnParams& p, cudaStream_t stream);        // fp32 FFMA, one query per thread
bool encoder_attention_tc_supported(const EncAttnParams& p);                        // head size 32, T <= 512
cudaError_t encoder_attention_tc(const EncAttnParams& p, cudaStream_t stream);     // tcgen05, fp16 two-term split

// Bahdanau / "mlp" global attention, one decode step (onmt/modules/global_attention.py:123-136,
// 180-194): score = v . tanh(wq + uh[t]); length mask; softmax; context = sum_t a[t] * mem[t].
struct MlpAttnParams {
  const float* wq = nullptr;                      // [n_chunks*NQ, d]  (Wq h + b)
  const float* uh = nullptr;                      // [n_chunks, T, d]  (Uk mem, cached once per chunk); dot: keys
  const float* mem = nullptr;                     // [n_chunks, T, d]
  const float* v = nullptr;                       // [d]
  const int64_t* lengths = nullptr;               // [n_chunks] memory lengths
  const int* retired = nullptr;
  float* ctx = nullptr; int64_t ctx_ld = 0;       // [n_chunks*NQ, *] context written at column 0
  float* attn = nullptr;                          // optional [n_chunks*NQ, T]
  int n_chunks = 0, NQ = 1, T = 0, d = 0;
  int dot = 0;                                    // 1: score = q . mem[t] (general / dot attention)
  // Fixed-point planes (kv_fmt != 0; uh / mem are then unused): the "K" part of a row holds uh (dot: the keys), the "V"
  // part mem, packed by kv_pack2 exactly like the Transformer decoder's memory keys / values (CrossAttnParams)
  int kv_fmt = 0;
  const int16_t* kv_hi = nullptr;
  const uint8_t* kv_lo = nullptr;
  const float* kv_scale = nullptr;
};
cudaError_t mlp_attention(const MlpAttnParams& p, cudaStream_t stream);
cudaError_t mlp_attention_packed(const MlpAttnParams& p, cudaStream_t stream);   // kv_fmt KV_Q23M / KV_Q15M, d 64 - 512

// ---------------------------------------------------------------------------------------------
cudaError_t layernorm_rows(const float* x, const float* g, const float* b, float eps, float* y,
                           int64_t M, int d, cudaStream_t stream);
// x[row] = emb[tok[row]] (* emb_scale + pe_row[:] when pe_row: the checkpoint's positional-encoding row of this step)
cudaError_t embed_rows(const int* tok, const float* emb, float* x, int64_t x_ld, int rows, int d,
                       const float* pe_row, float emb_scale, cudaStream_t stream);
// y[b,t,:] = w[:] * x[b,t] + bias[:]   (Linear(1,d): encoder/transformer.py:113, cnn_encoder.py:38)
cudaError_t linear_in1(const float* x, const float* w, const float* bias, float* y, int64_t n, int d,
                       cudaStream_t stream);
// MaxPool1d(stride) over time: in [B,T,d] -> out [B,T/stride,d]   (encoder/nano_encoder.py:101-105)
// ResNet stem, first layer (encoder/resnet_encoder.py:123-126,154-156): width-3 convolution of the one-channel signal
// with zero padding at the chunk ends, eval BatchNorm folded into w / b, ReLU.  src [B,T] -> out [B*T, 64]
cudaError_t resnet_stem_conv(const float* src, const float* w, const float* b, float* out, int B, int T, cudaStream_t stream);
// out[r] = x[r*ld + col]
cudaError_t take_column(const float* x, int ld, int col, float* out, int64_t rows, cudaStream_t stream);
cudaError_t maxpool_time(const float* in, float* out, int B, int T, int d, int stride, cudaStream_t stream);
// lengths after MaxPool1d(stride): out[i] = floor((in[i] - stride) / stride + 1) (in == out allowed)
cudaError_t pool_lengths(const int64_t* in, int64_t* out, int B, int stride, cudaStream_t stream);
cudaError_t fill_int(int* p, int n, int value, cudaStream_t stream);
// out[c,b,t] = in[b,t,c]  (the CNN encoder's reference layout, encoder/cnn_encoder.py:43-44)
cudaError_t transpose_to_dbt(const float* in, float* out, int B, int T, int d, cudaStream_t stream);
// out[t,b,:] = in[b,t,:]  (chunk-major -> reference time-major layout)
cudaError_t transpose_bt(const float* in, float* out, int B, int T, int d, cudaStream_t stream);

// Generator + token selection (models/model_builder.py:331-334, translate/translator.py:371-375,
// 469-477): optional final LayerNorm, Linear(d,V), log_softmax; greedy argmax (lowest index on ties).
struct GenParams {
  const float* x = nullptr; int64_t x_ld = 0;     // [rows, d]
  const float* ln_g = nullptr; const float* ln_b = nullptr; float eps = 1e-6f;   // null -> no LN
  const float* Wg = nullptr; const float* bg = nullptr;                         // [V,d], [V]
  float* logp = nullptr;                          // [rows, V] log-probs (always written)
  int rows = 0, d = 0, V = 0;
  // greedy extras (ids == nullptr -> skipped)
  int64_t* ids = nullptr; int ids_ld = 0;         // ids[row*ids_ld + step]
  float* scores = nullptr;                        // [rows] log-prob of the chosen token
  int* next_tok = nullptr;                        // [rows] feeds the next step's embedding
  float* trace = nullptr;                         // optional [rows, V] copy (per-step logits dump)
  int step = 0, min_len = 0, eos = 3;
};
cudaError_t generator_step(const GenParams& p, cudaStream_t stream);

// Beam step (translate/translator.py:714-762, 781-810): see beam.cu
struct BeamState {
  // all [B] / [B,K] device arrays owned by the engine
  float* topk_log_probs = nullptr;                // [B,K]
  int* alive_seq = nullptr;                       // [2][B*K, Lmax+1] double buffered
  int* anc = nullptr;                             // [2][B*K, Lmax] self-attention ancestor slots
  int* cur_tok = nullptr;                         // [B*K]
  int* parent = nullptr;                          // [B*K] parent row of the current beams
  int* retired = nullptr;                         // [B]
  int* retire_step = nullptr;                     // [B] fast mode: the step at which the chunk left the batch
  int* top_finished = nullptr;                    // [B]
  int* n_hyp = nullptr;                           // [B] finished hypotheses seen so far
  float* hyp_score = nullptr;                     // [B, n_best]
  int* hyp_len = nullptr;                         // [B, n_best]
  int* hyp_seq = nullptr;                         // [B, n_best, Lmax]
  int* hyp_anc = nullptr;                         // [B, n_best, Lmax] row that ran step j of the hypothesis (ancestor table)
  int* hyp_meta = nullptr;                        // [B, n_best] (arrival number << 8) | beam index of the hypothesis
  int* n_alive = nullptr;                         // [1] chunks not yet retired
  int* n_done = nullptr;                          // [1] object mode: chunks whose Beam.done() is true
  int* stop_step = nullptr;                       // [1] object mode: first step that is NOT executed any more
};
struct BeamParams {
  const float* logp = nullptr;                    // [B*K, V]
  BeamState st;
  int B = 0, K = 0, V = 0, Lmax = 0, step = 0, max_len = 0, min_len = 0, n_best = 1, eos = 3;
  int b0 = 0, nb = 0;                             // chunk range [b0, b0+nb) handled by this launch (B = total)
  float alpha = 0.f;
  // mode 1 = object beam (translator.py:827-926 + onmt/translate/beam.py): EOS rows get -1e20 children, every
  // chunk keeps advancing until ALL chunks are done, finished hypotheses are ranked by the global score
  // = beam score / gs_div (length penalty none / wu / avg)
  int mode = 0;
  int lp_mode = 0;                                // 0 none, 1 wu, 2 avg
  // object mode extras (onmt/translate/beam.py:101-124, penalties.py:39-57, beam.py:203-243):
  int block_ngram = 0;                            // n > 0: a beam whose hypothesis repeats an n-gram gets -10e20 children
  unsigned excl_mask = 0;                         // vocabulary ids (bit v) whose n-grams are never blocked
  int cov_mode = 0;                               // coverage penalty: 0 none, 1 wu, 2 summary
  int stepwise = 0;                               // -stepwise_penalty: the penalty steers the search (beam.py:87-88,218-227)
  float beta = 0.f;
  const float* attn_step = nullptr;               // [B*K, Tp] this step's attention per row (cov_mode != 0)
  float* cov = nullptr;                           // [2][B*K, Tp] coverage = sum of the attention along the hypothesis
  float* cov_pen = nullptr;                       // [B*K] coverage penalty of the current beams
  const int64_t* mem_len = nullptr;               // [B]
  int Tp = 0;
};
cudaError_t beam_init(const BeamParams& p, int bos, cudaStream_t stream);
cudaError_t beam_step(const BeamParams& p, cudaStream_t stream);
cudaError_t beam_finalize(const BeamParams& p, int64_t* out_ids, int* out_lens, float* out_scores,
                          cudaStream_t stream);
// attention rows of the finished hypotheses: out[b, n, j, :] = hist[j][hyp_anc[b,n,j]][:] for j < hyp_len[b,n], else 0
// (hist [*, rows, Tp] = the head-0 cross attention every decode step wrote per row)
// widths [B, n_best]: how many source positions the reference keeps of those rows -- memory_lengths[i] of the TILED
// length vector indexed by the chunk's position i in the (fast mode: not yet retired) batch, i.e. the length of chunk
// alive[i / K] (translator.py:776 and :905; equal to the chunk's own length whenever the lengths are equal)
cudaError_t beam_gather_attention(const BeamState& st, const float* hist, const int64_t* mem_len, int B, int K, int n_best,
                                  int Lmax, int max_len, int rows, int Tp, int mode, float* out, int* widths,
                                  cudaStream_t stream);
// reorder per-row recurrent state by parent: dst[r] = src[parent[r]]  (rows of `width` floats)
cudaError_t gather_rows(const float* src, float* dst, const int* parent, int row0, int rows, int width,
                        cudaStream_t stream);

// AverageAttention at one decode step (onmt/modules/average_attn.py:52-75):
//   g[r] = (xn[r] + step * prev[parent ? parent[r] : r]) / (step + 1)      (prev ignored at step 0)
cudaError_t avg_attn_cumulate(const float* xn, const float* prev, const int* parent, int row0, int rows, int d, int step,
                              float* g, cudaStream_t stream);
//   out[r] = sigmoid(gate[r, :d]) * xn[r] + sigmoid(gate[r, d:]) * a[r] + x[r]   (:100-104 + the layer's residual)
cudaError_t avg_attn_gate(const float* gate, const float* xn, const float* a, const float* x, float* out, int64_t rows, int d,
                          cudaStream_t stream);
// x = tanh(x) (accurate tanhf): output stage of the general / dot global attention (global_attention.py:201-203)
cudaError_t tanh_inplace(float* x, int64_t n, cudaStream_t stream);

// LSTMCell pointwise (onmt/models/stacked_rnn.py:25-31): gates [rows,4d] (= x.W_ih^T+b_ih + h.W_hh^T+b_hh)
cudaError_t lstm_cell_pointwise(const float* gates_a, const float* gates_b, const float* c_in,
                                float* h_out, float* c_out, int rows, int d, cudaStream_t stream);
// GRUCell pointwise: gates [rows,3d] as [r | z | n]; h_out = n + z * (h_in - n), n = tanh(a_n + r * b_n)
cudaError_t gru_cell_pointwise(const float* gates_a, const float* gates_b, const float* h_in, float* h_out, int rows,
                               int d, cudaStream_t stream);
// out = (x + a * sigmoid(g)) * sqrt(0.5) with y = [a | g] rows of 2d   (onmt/utils/cnn_factory.py:31-34,52-53)
cudaError_t glu_residual(const float* y, const float* x, float* out, float* glu_out, int64_t rows, int d,
                         cudaStream_t stream);
// im2col over time for a (k x 1) convolution: A[(b,t)][j*d + c] = x[b][t + j - left][c] (0 outside [0,T))
cudaError_t im2col_time(const float* x, float* A, int B, int T, int d, int k, int left, cudaStream_t stream);

// ---- CNN decoder step helpers (onmt/decoders/cnn_decoder.py:105-116, conv_multi_step_attention.py:66-82)
// hist[slot][pos][d] keeps every layer input; row r reads position p from slot anc[r][p] (beam) or r.
// Stores x (this step's layer input) at position t and builds the causal conv window
// A[r][j*d + c] = layer input at position t-(k-1)+j (zeros before the sequence start).
cudaError_t cnn_window(const float* x, float* hist, const int* anc, int anc_ld, const int* retired,
                       int rows_per_chunk, float* A, int row0, int rows, int t, int k, int d, int Lmax,
                       cudaStream_t stream);
// out = (a + b) * s
cudaError_t add_scale(const float* a, const float* b, float s, float* out, int64_t n, cudaStream_t stream);
// out = (x + (c + o) * s) * s
cudaError_t cnn_combine(const float* x, const float* c, const float* o, float s, float* out, int64_t n,
                        cudaStream_t stream);

// Front end (utils/labelop.py:219-233): see frontend.cu
// todo: optional int[n_reads] scratch: with it the histogram kernel runs first and the radix-select kernel only takes the
// reads whose value range exceeds the histogram (todo[r] != 0)
cudaError_t frontend_stats(const int16_t* signal, const int64_t* offsets, int n_reads, int mode,
                           double* center, double* scale, int* todo, cudaStream_t stream);
void frontend_set_fast(int on);     // 1 (default): histogram statistics + vectorised chunk gather; 0: the general kernels
cudaError_t frontend_chunks(const int16_t* signal, const int64_t* offsets, const double* center,
                            const double* scale, const int32_t* chunk_read, const int64_t* chunk_start,
                            int n_chunks, int chunk_len, float* out, int64_t* out_len, cudaStream_t stream);

// the same for float-valued reads (fp64 samples): see frontend.cu
cudaError_t frontend_stats_f64(const double* signal, const int64_t* offsets, int n_reads, int mode,
                               double* center, double* scale, cudaStream_t stream);
cudaError_t frontend_chunks_f64(const double* signal, const int64_t* offsets, const double* center,
                                const double* scale, const int32_t* chunk_read, const int64_t* chunk_start,
                                int n_chunks, int chunk_len, float* out, int64_t* out_len, cudaStream_t stream);

}  // namespace nd
