// libnanodec engine: weight registry + packing, workspace, encoder / decoder orchestration and the
// extern "C" ABI declared in include/nanodec.h.  All device work is enqueued on the caller's stream.
#include <math.h>
#include <string.h>

#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/nanodec.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "lstm.cuh"

using namespace nd;

namespace nd { int g_pdl = 2; }   // programmatic dependent launch for the tcgen05 GEMMs only, see common.cuh

namespace {

std::string g_create_error;

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> f;        // fp32 payload
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

// a dense projection y = x W^T + b in the layouts the GEMM kernels want
struct Lin {
  const float* W = nullptr;    // [N,K] fp32 (SIMT path)
  const float* W_hi = nullptr; // tf32-exact high part (tcgen05 path)
  const float* W_lo = nullptr; // residual
  const float* b = nullptr;    // [N] or null
  int N = 0, K = 0;
  int64_t ld = 0;
  // tcgen05 path with a prologue folded into the weights at pack time (see GemmParams::ln_cvec):
  int folded = 0;              // 0 none, 1 LayerNorm (W_hi/W_lo hold W*g, cvec/dvec set), 2 column affine
  const float* cvec = nullptr; // [N] sum_k W[n,k] g[k]
  const float* dvec = nullptr; // [N] LN: sum_k W[n,k] beta[k] + b[n];  affine: b[n] + sum_k W[n,k] shift[k]
};

struct LnW { const float* g = nullptr; const float* b = nullptr; };

// ResNet stem of the resnet / crnn / ctransformer encoders (encoder/resnet_encoder.py:82-170): every layer is a
// width-3 convolution along time, packed as a [Cout, 3*Cin] GEMM weight over im2col rows (column j*Cin + c), with the
// eval BatchNorm that follows folded in (w * alpha, b * alpha + beta)
struct ResBlockW {
  Lin conv1, conv2, down;        // down: the 1x1 projection of the residual (first block of layers 2-4)
  bool has_down = false;
  int cin = 0, cout = 0;
};
struct ResNetW {
  const float* stem_w = nullptr; // [64][3] (BatchNorm folded)
  const float* stem_b = nullptr; // [64]
  std::vector<ResBlockW> blocks;
  Lin fc;                        // [d, 512]
};

struct EncLayerT {             // transformer encoder layer
  Lin qkv, out, w1, w2;
  LnW ln, ln_ff;
};
struct DecLayerT {             // transformer decoder layer
  Lin qkv, self_out, cq, ckv, ctx_out, w1, w2;
  Lin avg_w1, avg_w2, gate_x, gate_a;   // -self_attn_type average: average_layer (FFN d -> d -> d), gating_layer split by input
  Lin cqt, cm;                 // memory-bank-space cross attention (kernels.cuh CrossMbParams): P [H*d, d], Mcat [d, H*d]
  LnW ln1, ln2, ln_ff, ln_avg;
};
struct LstmW {                 // one bidirectional (or unidirectional) LSTM layer
  Lin ih;                      // [dirs*4H, in] with b_ih (input projection GEMM), in > 1
  const float* w_ih0 = nullptr;   // [dirs*4H] when in == 1
  const float* b_ih0 = nullptr;
  const float* w_hh = nullptr; // [dirs*4H, H]
  const float* b_hh = nullptr; // [dirs*4H]
  const float* bn_alpha = nullptr;   // eval BatchNorm of this layer's output folded to y = x*alpha + beta
  const float* bn_beta = nullptr;
  int in = 1;
  bool tc_ok = false;          // |W_hh| within the fp16 range: the tensor-core recurrence may be used
};
struct ConvW { Lin conv; };    // weight-normalised (k x 1) conv as [2d, k*d] GEMM

}  // namespace

struct nd_engine {
  nd_config cfg;
  std::string err;
  bool sticky = false;
  int n_sm = 148;
  int64_t launches = 0;
  std::map<std::string, HostTensor> raw;
  std::set<std::string> used;          // checkpoint tensors the packer consumed (finalize() rejects leftovers)
  std::vector<void*> allocs;
  bool finalized = false;
  bool oom = false;
  // per-category CUDA-event timing of kernel launches (nd_profile_enable / nd_profile_read)
  uint32_t prof_mask = 0;
  std::vector<cudaEvent_t> prof_pool;
  size_t prof_used = 0;
  std::vector<int> prof_cat;           // category of event pair i (events 2i, 2i+1)

  // ---- packed weights
  std::vector<LstmW> lstm;           // nano / rnn / brnn encoders
  Lin encW;                          // nano: final projection
  Lin enc_lin_in;                    // transformer / cnn encoder Linear(1,d) (W used as vector)
  std::vector<EncLayerT> encT;
  LnW enc_ln;
  std::vector<ConvW> enc_conv, dec_conv;
  ResNetW resnet;
  float* rbuf[3] = {nullptr, nullptr, nullptr};   // [maxB*maxT, 512] activations of the ResNet stem
  float* cmask = nullptr;            // ctransformer: channel 0 of the stem output, the encoder's key-mask source [B,T]
  std::vector<DecLayerT> decT;
  LnW dec_ln;
  const float* emb = nullptr;        // [V,d]
  const float* pe = nullptr;         // [max_tgt_len, d] rows of the checkpoint's PositionalEncoding buffer (or null)
  Lin gen;                           // generator (W, b)
  // rnn decoder
  struct RnnCell { Lin ih_e, ih_f, ih, hh; };
  std::vector<RnnCell> cells;
  Lin attn_ctx, attn_q, attn_in, attn_out_c, attn_out_h;
  const float* attn_v = nullptr;
  // cnn decoder
  Lin dec_lin;
  std::vector<Lin> dec_attn_in;

  // ---- workspace (sized at create)
  float* src = nullptr;              // [maxB, maxT]
  int64_t* lengths = nullptr;        // [maxB]
  int64_t* mem_len = nullptr;        // [maxB]
  float* bufA = nullptr; float* bufB = nullptr;    // [maxB*maxT, d] activations ping-pong
  float* bufC = nullptr;                            // [maxB*maxT, d]
  float* big = nullptr;              // [maxB*maxT, max(4d, 3d, ff, k*d)] xg / qkv / ffn hidden / im2col
  float* big2 = nullptr;             // [maxB*maxT, 2d] conv outputs
  float* mb = nullptr;               // memory bank [B, T', d]
  float* emb_remap = nullptr;        // cnn encoder: Linear(1,d) output (decoder init_state needs it)
  float* enc_hn = nullptr; float* enc_cn = nullptr;   // rnn encoders: [Le*dirs, B, hh]
  float* enc_hb = nullptr; float* enc_cb = nullptr;   // the same after -bridge (Linear + ReLU over rows of hh*Le floats)
  Lin bridge_h, bridge_c;
  int* fe_todo = nullptr; int fe_todo_cap = 0;     // front end: reads left for the radix-select kernel
  int B = 0, T = 0, Tp = 0;          // last encode
  bool encoded = false;

  // decoder workspace
  std::vector<float*> ckv;           // per layer [B*T', 2d]
  std::vector<float*> selfK, selfV;  // per layer [rows, L, d]
  float *x = nullptr, *qkv = nullptr, *sctx = nullptr, *x1 = nullptr, *qc = nullptr, *cctx = nullptr, *x2 = nullptr,
        *ffh = nullptr, *logp = nullptr, *gscore = nullptr;
  float *qt = nullptr, *cctxt = nullptr;   // [maxB, H*d] memory-bank-space cross attention (greedy)
  // 0 (default): K/V cross attention, HBM-bound at 96 % of the measured peak.  1: greedy decode reads the memory
  // bank once per layer-step (half the bytes, 8x the fp32 FMAs): measured 253 us (v2) vs 171 us per launch at d = 256
  // (profiles/r01_cross_mb_experiment.md) -- kept as an option until its FFMA side is as good as its byte count
  const int* gemm_alive = nullptr;         // set while a beam search runs: GEMMs return at once when it reads 0
  int cross_mode = 0;
  // storage of the projected memory keys / values (kernels.cuh): KV_Q23M (default: 3 bytes per element, absolute
  // error <= 2^-23 of the row part's largest element; same greedy sequences and identity rates as fp32 storage),
  // KV_F32, KV_Q15M (reduced precision, half the bytes), others = cross-checks
  int kv_mode = KV_Q23M;
  int beam_n_best = 1, beam_K = 1, beam_mode = 0;   // of the last beam decode (layout of the hypothesis tables)
  // object beam extras (options block_ngram_repeat / block_ngram_exclude / coverage_penalty, nd_set_float "beta")
  int block_ngram = 0; unsigned excl_mask = 0; int cov_mode = 0; float beta = 0.f; int stepwise = 0;
  float* cov = nullptr; float* cov_pen = nullptr; float* cov_attn = nullptr;   // allocated on first use
  bool beam_attn = false;                  // option "beam_attention": beam decodes keep the per-step head-0 cross attention
  float* attn_hist = nullptr;              // [max_tgt_len][max_batch*max_beam][max_src_len], allocated on first use
  int attn_rows = 0, attn_Tp = 0, attn_steps = 0;   // geometry of the history the last beam decode wrote
  int kv_beam_packed = 1;                  // beam search reads the fixed-point planes too where the kernel exists
  bool kv_packed = false;                  // set by decoder_init: this decode reads the packed planes
  int enc_attn_tc = 1;                     // Transformer-encoder self attention on the tensor cores when dh = 32
  int* cur_tok = nullptr;
  // rnn decoder
  float* uh = nullptr;               // [B, T', d]
  int pk_fmt = KV_Q23M;              // format of the planes in pk (the CNN decoder always takes q23)
  float* pk = nullptr;               // RNN / CNN decoders: fixed-point planes of (uh | H) or (enc top | enc combined), sized like fp32 [B*T', 2d]
  std::vector<float*> rh[2], rc[2];  // ping-pong recurrent state [rows,d] per layer
  float* feed[2] = {nullptr, nullptr};
  float *ga = nullptr, *gb = nullptr, *wq = nullptr, *actx = nullptr;
  // cnn decoder
  float* enc_comb = nullptr;         // (mb + emb_remap) * sqrt(.5)  [B,T,d]
  std::vector<float*> chist[2];      // per layer input history [rows, L, d] ping-pong for beams
  float *cbase = nullptr, *cA = nullptr, *cy = nullptr, *cout = nullptr, *cpre = nullptr, *ctgt = nullptr,
        *cctx2 = nullptr;
  // beam state
  BeamState beam;
  int max_rows = 0;
  // decode-loop concurrency: the batch is cut into `decode_streams` contiguous chunk groups that run the
  // step loop on their own streams (chunks are independent), so the latency-bound projections of one
  // group overlap the HBM-bound attention of another
  int decode_streams = 1;
  // CUDA graphs: the whole decode loop of a (mode, B, T, L, ...) configuration is captured once (second
  // call with the same key) and replayed; outputs go to engine-owned buffers and are copied to the caller's
  int use_graphs = 1;
  struct GraphEntry {
    std::vector<int64_t> key;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::vector<GraphEntry> graphs;
  std::vector<int64_t> last_key;
  cudaStream_t main_stream = nullptr;
  cudaEvent_t g_in = nullptr, g_out = nullptr;
  int64_t* o_ids = nullptr; float* o_scores = nullptr; int* o_lens = nullptr;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> join_ev;
  cudaEvent_t fork_ev = nullptr;
};

namespace {

// ------------------------------------------------------------------------------------------ errors
int fail(nd_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg; else g_create_error = msg;
  return code;
}
#define ND_CUDA(e, call)                                                                      \
  do {                                                                                        \
    cudaError_t _err = (call);                                                                \
    if (_err != cudaSuccess) {                                                                \
      (e)->sticky = true;                                                                     \
      return fail((e), ND_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_err));    \
    }                                                                                         \
  } while (0)
// kernel launch through a launcher returning cudaError_t
#define ND_LAUNCH(e, call) ND_LAUNCH_CAT(e, ND_PROF_OTHER, nullptr, call)
// same, attributed to a profiling category and (when that category is enabled) bracketed by CUDA events
#define ND_LAUNCH_CAT(e, cat, st_, call)                                                      \
  do {                                                                                        \
    ++(e)->launches;                                                                          \
    const bool _prof = ((e)->prof_mask >> (cat)) & 1u;                                        \
    if (_prof) prof_begin((e), (cat), (cudaStream_t)(st_));                                   \
    ND_CUDA(e, call);                                                                         \
    if (_prof) prof_end((e), (cudaStream_t)(st_));                                            \
  } while (0)

void prof_begin(nd_engine* e, int cat, cudaStream_t st) {
  if (e->prof_used + 2 > e->prof_pool.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t ev;
      cudaEventCreate(&ev);
      e->prof_pool.push_back(ev);
    }
  }
  e->prof_cat.push_back(cat);
  cudaEventRecord(e->prof_pool[e->prof_used], st);
}
void prof_end(nd_engine* e, cudaStream_t st) {
  cudaEventRecord(e->prof_pool[e->prof_used + 1], st);
  e->prof_used += 2;
}

template <class T>
T* dalloc(nd_engine* e, size_t n) {
  void* p = nullptr;
  if (n == 0) n = 1;
  cudaError_t err = cudaMalloc(&p, n * sizeof(T));
  if (err != cudaSuccess) {
    e->err = std::string("cudaMalloc of ") + std::to_string(n * sizeof(T)) + " bytes failed: " + cudaGetErrorString(err);
    cudaGetLastError();
    return nullptr;
  }
  e->allocs.push_back(p);
  return static_cast<T*>(p);
}

const float* upload(nd_engine* e, const std::vector<float>& v) {
  float* p = dalloc<float>(e, v.size());
  if (!p) { e->oom = true; return nullptr; }
  if (!v.empty()) cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
  return p;
}

bool tc_mode(const nd_engine* e) { return e->cfg.gemm_mode != ND_GEMM_SIMT_FP32; }

// fold: 0 none; 1 LayerNorm(fg = gamma, fb = beta) in front of the projection; 2 column affine
// a*fg[k] + fb[k] (eval BatchNorm) in front of it.  Folding only affects the tcgen05 operands; the
// SIMT cross-check path keeps the original weight and applies the prologue to A like the reference.
Lin make_lin(nd_engine* e, const std::vector<float>& W, int N, int K, const std::vector<float>* b, int fold = 0,
             const std::vector<float>* fg = nullptr, const std::vector<float>* fb = nullptr) {
  Lin l;
  l.N = N; l.K = K; l.ld = K;
  l.W = upload(e, W);
  if (tc_mode(e)) {
    std::vector<float> Wf(W);
    if (fold) {
      std::vector<float> cvec(N), dvec(N);
      for (int n = 0; n < N; ++n) {
        double c = 0.0, dsum = 0.0;
        for (int k = 0; k < K; ++k) {
          const float w = W[(size_t)n * K + k];
          Wf[(size_t)n * K + k] = w * (*fg)[k];
          c += (double)w * (double)(*fg)[k];
          dsum += (double)w * (double)(*fb)[k];
        }
        cvec[n] = (float)c;
        dvec[n] = (float)(dsum + (b ? (double)(*b)[n] : 0.0));
      }
      l.folded = fold;
      l.cvec = upload(e, cvec);
      l.dvec = upload(e, dvec);
    }
    std::vector<float> hi(W.size()), lo(W.size());
    split_tf32_host(Wf.data(), hi.data(), lo.data(), W.size());
    l.W_hi = upload(e, hi);
    l.W_lo = upload(e, lo);
  }
  if (b) l.b = upload(e, *b);
  return l;
}
// view of columns [c0, c0+k) of a packed weight (ld stays the full row pitch)
Lin col_slice(const Lin& l, int c0, int k, bool keep_bias) {
  Lin s = l;
  s.W = l.W + c0;
  if (l.W_hi) { s.W_hi = l.W_hi + c0; s.W_lo = l.W_lo + c0; }
  s.K = k;
  if (!keep_bias) s.b = nullptr;
  return s;
}

struct GemmOpt {
  int prologue = PRO_NONE; const float* pg = nullptr; const float* pb = nullptr; float eps = 1e-6f;
  int act = 0; const float* residual = nullptr; int64_t ldr = 0; float div_by = 1.f; int div_ncols = 0;
};

int run_gemm(nd_engine* e, const Lin& l, const float* A, int64_t lda, float* C, int64_t ldc, int64_t M,
             const GemmOpt& o, cudaStream_t st) {
  GemmParams p;
  p.A = A; p.lda = lda; p.C = C; p.ldc = ldc; p.M = (int)M; p.N = l.N; p.K = l.K; p.ldw = l.ld;
  p.bias = l.b; p.prologue = o.prologue; p.pg = o.pg; p.pb = o.pb; p.eps = o.eps; p.relu = o.act;
  p.residual = o.residual; p.ldr = o.ldr; p.div_by = o.div_by; p.div_ncols = o.div_ncols;
  p.alive = e->gemm_alive;
  const bool tma_ok = (lda % 4 == 0) && (l.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(l.W) & 15) == 0) && l.K >= 8;
  if (tc_mode(e) && tma_ok && l.W_hi) {
    p.W = l.W_hi; p.W_lo = l.W_lo;
    if (o.prologue == PRO_LAYERNORM) {
      if (l.folded != 1) return fail(e, ND_ERR_STATE, "internal: LayerNorm GEMM without folded weights");
      p.prologue = PRO_NONE; p.ln_cvec = l.cvec; p.ln_dvec = l.dvec; p.bias = nullptr;
    } else if (o.prologue == PRO_AFFINE) {
      if (l.folded != 2) return fail(e, ND_ERR_STATE, "internal: affine GEMM without folded weights");
      p.prologue = PRO_NONE; p.bias = l.dvec;
    } else if (l.folded) {
      return fail(e, ND_ERR_STATE, "internal: folded weights used without their prologue");
    }
    ND_LAUNCH_CAT(e, ND_PROF_GEMM, st, gemm_tc(p, e->cfg.gemm_mode == ND_GEMM_TC_3XTF32 ? 3 : 1, st));
  } else {
    p.W = l.W; p.W_lo = nullptr;
    ND_LAUNCH_CAT(e, ND_PROF_GEMM, st, gemm_simt(p, st));
  }
  return ND_OK;
}
#define ND_TRY(expr) do { int _rc = (expr); if (_rc != ND_OK) return _rc; } while (0)

// NVTX range over the host-side enqueue of one phase (encode / memory K,V / decode loop): nsys and ncu --nvtx group the
// kernels of a phase by it; header-only NVTX v3 resolves its injection library lazily, so without a tool it is a no-op
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------------------------------ weights
const HostTensor* find(nd_engine* e, const std::string& k) {
  auto it = e->raw.find(k);
  if (it != e->raw.end()) e->used.insert(k);
  return it == e->raw.end() ? nullptr : &it->second;
}
int need(nd_engine* e, const std::string& k, std::vector<int64_t> shape, const HostTensor** out) {
  const HostTensor* t = find(e, k);
  if (!t) return fail(e, ND_ERR_WEIGHT, "missing checkpoint tensor '" + k + "'");
  if (t->shape != shape) {
    std::string s = "tensor '" + k + "' has shape [";
    for (auto v : t->shape) s += std::to_string(v) + ",";
    s += "] expected [";
    for (auto v : shape) s += std::to_string(v) + ",";
    return fail(e, ND_ERR_WEIGHT, s + "]");
  }
  *out = t;
  return ND_OK;
}
int load_lin(nd_engine* e, const std::string& prefix, int N, int K, bool bias, Lin* out,
             const std::string& ln_prefix = "") {
  const HostTensor *w, *b = nullptr;
  ND_TRY(need(e, prefix + ".weight", {N, K}, &w));
  if (bias) ND_TRY(need(e, prefix + ".bias", {N}, &b));
  if (!ln_prefix.empty()) {
    const HostTensor *g, *bb;
    ND_TRY(need(e, ln_prefix + ".weight", {K}, &g));
    ND_TRY(need(e, ln_prefix + ".bias", {K}, &bb));
    *out = make_lin(e, w->f, N, K, b ? &b->f : nullptr, 1, &g->f, &bb->f);
  } else {
    *out = make_lin(e, w->f, N, K, b ? &b->f : nullptr);
  }
  return ND_OK;
}
int load_cat_lin(nd_engine* e, const std::vector<std::string>& prefixes, int N_each, int K, Lin* out,
                 const std::string& ln_prefix = "") {
  std::vector<float> W, b;
  for (auto& pfx : prefixes) {
    const HostTensor *w, *bb;
    ND_TRY(need(e, pfx + ".weight", {N_each, K}, &w));
    ND_TRY(need(e, pfx + ".bias", {N_each}, &bb));
    W.insert(W.end(), w->f.begin(), w->f.end());
    b.insert(b.end(), bb->f.begin(), bb->f.end());
  }
  if (!ln_prefix.empty()) {
    const HostTensor *g, *bb;
    ND_TRY(need(e, ln_prefix + ".weight", {K}, &g));
    ND_TRY(need(e, ln_prefix + ".bias", {K}, &bb));
    *out = make_lin(e, W, N_each * (int)prefixes.size(), K, &b, 1, &g->f, &bb->f);
  } else {
    *out = make_lin(e, W, N_each * (int)prefixes.size(), K, &b);
  }
  return ND_OK;
}
int load_ln(nd_engine* e, const std::string& prefix, int d, LnW* out) {
  const HostTensor *g, *b;
  ND_TRY(need(e, prefix + ".weight", {d}, &g));
  ND_TRY(need(e, prefix + ".bias", {d}, &b));
  out->g = upload(e, g->f);
  out->b = upload(e, b->f);
  return ND_OK;
}
// one (bi)directional LSTM layer: checkpoint keys <prefix>.weight_ih<sfx>[_reverse] ...
int load_lstm(nd_engine* e, const std::string& prefix, const std::string& sfx, int in, int H, int dirs, LstmW* out,
              const std::vector<float>* in_alpha = nullptr, const std::vector<float>* in_beta = nullptr) {
  const int NG = e->cfg.rnn_type == ND_RNN_GRU ? 3 : 4;     // gate rows per hidden unit (nn.GRU: r, z, n)
  std::vector<float> Wih, bih, Whh, bhh;
  for (int dir = 0; dir < dirs; ++dir) {
    const std::string s = sfx + (dir ? "_reverse" : "");
    const HostTensor *wi, *wh, *bi, *bh;
    ND_TRY(need(e, prefix + ".weight_ih" + s, {NG * H, in}, &wi));
    ND_TRY(need(e, prefix + ".weight_hh" + s, {NG * H, H}, &wh));
    ND_TRY(need(e, prefix + ".bias_ih" + s, {NG * H}, &bi));
    ND_TRY(need(e, prefix + ".bias_hh" + s, {NG * H}, &bh));
    Wih.insert(Wih.end(), wi->f.begin(), wi->f.end());
    Whh.insert(Whh.end(), wh->f.begin(), wh->f.end());
    bih.insert(bih.end(), bi->f.begin(), bi->f.end());
    bhh.insert(bhh.end(), bh->f.begin(), bh->f.end());
  }
  out->in = in;
  if (in == 1) {
    out->w_ih0 = upload(e, Wih);
    out->b_ih0 = upload(e, bih);
  } else {
    out->ih = in_alpha ? make_lin(e, Wih, dirs * NG * H, in, &bih, 2, in_alpha, in_beta)
                       : make_lin(e, Wih, dirs * NG * H, in, &bih);
  }
  out->w_hh = upload(e, Whh);
  out->b_hh = upload(e, bhh);
  float wmax = 0.f;
  for (float v : Whh) wmax = std::max(wmax, fabsf(v));
  out->tc_ok = wmax < 6.0e4f && NG == 4;           // the tensor-core recurrence kernel is the LSTM cell
  return ND_OK;
}
// weight-normalised conv (onmt/modules/weight_norm.py:153-165, eval: Polyak buffers):
// w = g/||V|| * V, packed as a [2d, k*d] GEMM weight with column j*d + c  <->  V[o][c][j]
int load_wnconv(nd_engine* e, const std::string& prefix, int d, int k, Lin* out) {
  const HostTensor *V, *g, *b;
  ND_TRY(need(e, prefix + ".V_avg", {2 * d, d, k, 1}, &V));
  ND_TRY(need(e, prefix + ".g_avg", {2 * d}, &g));
  ND_TRY(need(e, prefix + ".b_avg", {2 * d}, &b));
  std::vector<float> W((size_t)2 * d * k * d);
  for (int o = 0; o < 2 * d; ++o) {
    const float* v = V->f.data() + (size_t)o * d * k;
    // torch.norm(v.view(out,-1), 2, 1) in fp32
    double ss = 0.0;
    for (int i = 0; i < d * k; ++i) ss += (double)v[i] * (double)v[i];
    const float scalar = g->f[o] / (float)sqrt(ss);
    for (int c = 0; c < d; ++c)
      for (int j = 0; j < k; ++j) W[(size_t)o * k * d + (size_t)j * d + c] = scalar * v[c * k + j];
  }
  *out = make_lin(e, W, 2 * d, k * d, &b->f);
  return ND_OK;
}

// Memory-bank-space cross attention weights of one decoder layer (see CrossMbParams), products in double:
//   P[h*d + j][k]   = sum_i Wk[h*dh+i][j] Wq[h*dh+i][k] / sqrt(dh)      pb[h*d + j] = sum_i Wk[h*dh+i][j] bq[h*dh+i] / sqrt(dh)
//   M[n][h*d + j]   = sum_i Wo[n][h*dh+i] Wv[h*dh+i][j]                 mo[n] = bo[n] + sum_i Wo[n][i] bv[i]
// (bk only adds a per-head constant to every score of a row, which the softmax removes.)
int load_cross_mb(nd_engine* e, const std::string& p, const std::string& ln_prefix, int d, int H, Lin* cqt, Lin* cm) {
  const HostTensor *wq, *bq, *wk, *wv, *bv, *wo, *bo, *g, *bb;
  ND_TRY(need(e, p + ".linear_query.weight", {d, d}, &wq));
  ND_TRY(need(e, p + ".linear_query.bias", {d}, &bq));
  ND_TRY(need(e, p + ".linear_keys.weight", {d, d}, &wk));
  ND_TRY(need(e, p + ".linear_values.weight", {d, d}, &wv));
  ND_TRY(need(e, p + ".linear_values.bias", {d}, &bv));
  ND_TRY(need(e, p + ".final_linear.weight", {d, d}, &wo));
  ND_TRY(need(e, p + ".final_linear.bias", {d}, &bo));
  ND_TRY(need(e, ln_prefix + ".weight", {d}, &g));
  ND_TRY(need(e, ln_prefix + ".bias", {d}, &bb));
  const int dh = d / H;
  const double sc = 1.0 / sqrt((double)dh);
  const int64_t HD = (int64_t)H * d;
  std::vector<float> P((size_t)HD * d), pb((size_t)HD), M((size_t)d * HD), mo((size_t)d);
  std::vector<double> row((size_t)d);
  for (int h = 0; h < H; ++h)
    for (int j = 0; j < d; ++j) {
      std::fill(row.begin(), row.end(), 0.0);
      double bsum = 0.0;
      for (int i = 0; i < dh; ++i) {
        const double wkij = wk->f[(size_t)(h * dh + i) * d + j];
        const float* wqr = wq->f.data() + (size_t)(h * dh + i) * d;
        for (int k = 0; k < d; ++k) row[k] += wkij * (double)wqr[k];
        bsum += wkij * (double)bq->f[h * dh + i];
      }
      float* pr = P.data() + ((size_t)h * d + j) * d;
      for (int k = 0; k < d; ++k) pr[k] = (float)(row[k] * sc);
      pb[(size_t)h * d + j] = (float)(bsum * sc);
    }
  for (int n = 0; n < d; ++n) {
    double bsum = bo->f[n];
    for (int i = 0; i < d; ++i) bsum += (double)wo->f[(size_t)n * d + i] * (double)bv->f[i];
    mo[n] = (float)bsum;
    for (int h = 0; h < H; ++h) {
      std::fill(row.begin(), row.end(), 0.0);
      for (int i = 0; i < dh; ++i) {
        const double w = wo->f[(size_t)n * d + h * dh + i];
        const float* wvr = wv->f.data() + (size_t)(h * dh + i) * d;
        for (int j = 0; j < d; ++j) row[j] += w * (double)wvr[j];
      }
      float* mr = M.data() + (size_t)n * HD + (size_t)h * d;
      for (int j = 0; j < d; ++j) mr[j] = (float)row[j];
    }
  }
  *cqt = make_lin(e, P, (int)HD, d, &pb, 1, &g->f, &bb->f);
  *cm = make_lin(e, M, d, (int)HD, &mo);
  return ND_OK;
}

// eval BatchNorm as y = x*alpha + beta (ATen batch_norm_cpu_transform_input)
int load_bn_fold(nd_engine* e, const std::string& bp, int n, std::vector<float>* alpha, std::vector<float>* beta) {
  const HostTensor *w, *b, *rm, *rv;
  ND_TRY(need(e, bp + ".weight", {n}, &w));
  ND_TRY(need(e, bp + ".bias", {n}, &b));
  ND_TRY(need(e, bp + ".running_mean", {n}, &rm));
  ND_TRY(need(e, bp + ".running_var", {n}, &rv));
  alpha->resize(n); beta->resize(n);
  for (int i = 0; i < n; ++i) {
    const float invstd = 1.0f / sqrtf(rv->f[i] + 1e-5f);
    (*alpha)[i] = w->f[i] * invstd;
    (*beta)[i] = b->f[i] - rm->f[i] * (*alpha)[i];
  }
  return ND_OK;
}

// conv (kh x kw) + BatchNorm of the ResNet stem as a GEMM weight.  The image is [B, C, 1, T]: with kh = 5, padding 2
// only kernel row 2 meets data (kh = 1: row 0), and the strides act on the height-1 axis only.
int load_resconv(nd_engine* e, const std::string& conv, const std::string& bn, int cout, int cin, int kh, int kw,
                 bool bias, Lin* out) {
  const HostTensor *w, *b = nullptr;
  ND_TRY(need(e, conv + ".weight", {cout, cin, kh, kw}, &w));
  if (bias) ND_TRY(need(e, conv + ".bias", {cout}, &b));
  std::vector<float> alpha, beta;
  ND_TRY(load_bn_fold(e, bn, cout, &alpha, &beta));
  const int row = kh / 2;
  std::vector<float> W((size_t)cout * kw * cin), bb(cout);
  for (int o = 0; o < cout; ++o) {
    for (int c = 0; c < cin; ++c)
      for (int j = 0; j < kw; ++j)
        W[(size_t)o * kw * cin + (size_t)j * cin + c] = alpha[o] * w->f[(((size_t)o * cin + c) * kh + row) * kw + j];
    bb[o] = (b ? b->f[o] * alpha[o] : 0.f) + beta[o];
  }
  *out = make_lin(e, W, cout, kw * cin, &bb);
  return ND_OK;
}

int load_resnet(nd_engine* e, const std::string& p, int d) {
  ResNetW& R = e->resnet;
  {
    const HostTensor* w;
    ND_TRY(need(e, p + ".conv1.weight", {64, 1, 5, 3}, &w));
    std::vector<float> alpha, beta, sw(64 * 3);
    ND_TRY(load_bn_fold(e, p + ".bn1", 64, &alpha, &beta));
    for (int o = 0; o < 64; ++o)
      for (int j = 0; j < 3; ++j) sw[o * 3 + j] = alpha[o] * w->f[((size_t)o * 5 + 2) * 3 + j];
    R.stem_w = upload(e, sw);
    R.stem_b = upload(e, beta);
  }
  static const int planes[4] = {64, 128, 256, 512};
  int inplanes = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < 2; ++bi) {
      const std::string bp = p + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      ResBlockW B;
      B.cin = bi == 0 ? inplanes : planes[li];
      B.cout = planes[li];
      ND_TRY(load_resconv(e, bp + ".conv1", bp + ".bn1", B.cout, B.cin, 5, 3, true, &B.conv1));
      ND_TRY(load_resconv(e, bp + ".conv2", bp + ".bn2", B.cout, B.cout, 5, 3, true, &B.conv2));
      B.has_down = bi == 0 && B.cin != B.cout;
      if (B.has_down) ND_TRY(load_resconv(e, bp + ".downsample.0", bp + ".downsample.1", B.cout, B.cin, 1, 1, false, &B.down));
      R.blocks.push_back(B);
    }
    inplanes = planes[li];
  }
  ND_TRY(load_lin(e, p + ".fc", d, 512, true, &R.fc));
  return ND_OK;
}

int finalize(nd_engine* e) {
  const nd_config& c = e->cfg;
  const int d = c.d_model, V = c.vocab_size;
  // ---------------- encoder
  if (c.encoder_type == ND_ENC_RESNET || c.encoder_type == ND_ENC_CRNN || c.encoder_type == ND_ENC_CTRANSFORMER)
    ND_TRY(load_resnet(e, "encoder.cnn", d));
  if (c.encoder_type == ND_ENC_RESNET) {
    // encoder/resnet_encoder.py:202-251: the stem is the whole encoder
  } else if (c.encoder_type == ND_ENC_NANO || c.encoder_type == ND_ENC_CRNN) {
    const int H = d / 2;
    const bool crnn = c.encoder_type == ND_ENC_CRNN;       // crnn_encoder.py:62-68: the first layer reads the d-wide stem
    if (!lstm_supported(H)) return fail(e, ND_ERR_INVALID, "LSTM hidden size " + std::to_string(H) + " unsupported (16/32/64/128/256)");
    e->lstm.resize(c.enc_layers);
    std::vector<std::vector<float>> alphas(c.enc_layers), betas(c.enc_layers);
    for (int l = 0; l < c.enc_layers; ++l) {
      // eval BatchNorm1d as y = x*alpha + beta (ATen batch_norm_cpu_transform_input)
      const std::string bp = "encoder.batchnorm_" + std::to_string(l);
      const HostTensor *w, *b, *rm, *rv;
      ND_TRY(need(e, bp + ".weight", {d}, &w));
      ND_TRY(need(e, bp + ".bias", {d}, &b));
      ND_TRY(need(e, bp + ".running_mean", {d}, &rm));
      ND_TRY(need(e, bp + ".running_var", {d}, &rv));
      alphas[l].resize(d); betas[l].resize(d);
      for (int i = 0; i < d; ++i) {
        const float invstd = 1.0f / sqrtf(rv->f[i] + 1e-5f);
        alphas[l][i] = w->f[i] * invstd;
        betas[l][i] = b->f[i] - rm->f[i] * alphas[l][i];
      }
    }
    for (int l = 0; l < c.enc_layers; ++l) {
      // layer l >= 1 reads BatchNorm_{l-1}(output of layer l-1): folded into its input projection
      ND_TRY(load_lstm(e, "encoder.rnn_" + std::to_string(l), "_l0", (l == 0 && !crnn) ? 1 : d, H, 2, &e->lstm[l],
                       l ? &alphas[l - 1] : nullptr, l ? &betas[l - 1] : nullptr));
      e->lstm[l].bn_alpha = upload(e, alphas[l]);
      e->lstm[l].bn_beta = upload(e, betas[l]);
    }
    ND_TRY(load_lin(e, "encoder.W", d, d, false, &e->encW));
  } else if (c.encoder_type == ND_ENC_BRNN || c.encoder_type == ND_ENC_RNN) {
    const int dirs = c.encoder_type == ND_ENC_BRNN ? 2 : 1;
    const int H = d / dirs;
    if (!lstm_supported(H)) return fail(e, ND_ERR_INVALID, "LSTM hidden size " + std::to_string(H) + " unsupported (16/32/64/128/256)");
    e->lstm.resize(c.enc_layers);
    for (int l = 0; l < c.enc_layers; ++l)
      ND_TRY(load_lstm(e, "encoder.rnn", "_l" + std::to_string(l), l == 0 ? 1 : d, H, dirs, &e->lstm[l]));
    if (c.bridge) {
      // encoder/rnn_encoder.py:86-99: nn.Linear(hh * layers, hh * layers) per state (h; c for LSTM)
      const int tot = H * c.enc_layers;
      ND_TRY(load_lin(e, "encoder.bridge.0", tot, tot, true, &e->bridge_h));
      if (c.rnn_type != ND_RNN_GRU) ND_TRY(load_lin(e, "encoder.bridge.1", tot, tot, true, &e->bridge_c));
    }
  } else if (c.encoder_type == ND_ENC_TRANSFORMER || c.encoder_type == ND_ENC_CTRANSFORMER) {
    if (c.encoder_type == ND_ENC_TRANSFORMER) ND_TRY(load_lin(e, "encoder.linear", d, 1, true, &e->enc_lin_in));
    e->encT.resize(c.enc_layers);
    for (int l = 0; l < c.enc_layers; ++l) {
      const std::string p = "encoder.transformer." + std::to_string(l);
      EncLayerT& L = e->encT[l];
      ND_TRY(load_cat_lin(e, {p + ".self_attn.linear_query", p + ".self_attn.linear_keys", p + ".self_attn.linear_values"}, d, d, &L.qkv, p + ".layer_norm"));
      ND_TRY(load_lin(e, p + ".self_attn.final_linear", d, d, true, &L.out));
      ND_TRY(load_lin(e, p + ".feed_forward.w_1", c.d_ff, d, true, &L.w1, p + ".feed_forward.layer_norm"));
      ND_TRY(load_lin(e, p + ".feed_forward.w_2", d, c.d_ff, true, &L.w2));
      ND_TRY(load_ln(e, p + ".layer_norm", d, &L.ln));
      ND_TRY(load_ln(e, p + ".feed_forward.layer_norm", d, &L.ln_ff));
    }
    ND_TRY(load_ln(e, "encoder.layer_norm", d, &e->enc_ln));
  } else if (c.encoder_type == ND_ENC_CNN) {
    ND_TRY(load_lin(e, "encoder.linear", d, 1, true, &e->enc_lin_in));
    e->enc_conv.resize(c.enc_layers);
    for (int l = 0; l < c.enc_layers; ++l)
      ND_TRY(load_wnconv(e, "encoder.cnn.layers." + std::to_string(l) + ".conv", d, c.cnn_kernel_width, &e->enc_conv[l].conv));
  } else {
    return fail(e, ND_ERR_INVALID, "unknown encoder_type");
  }
  // ---------------- decoder
  {
    const HostTensor* em;
    ND_TRY(need(e, "decoder.embeddings.make_embedding.emb_luts.0.weight", {V, d}, &em));
    e->emb = upload(e, em->f);
    if (c.position_encoding) {
      // the sinusoid table is a registered buffer of the reference's PositionalEncoding (onmt/modules/embeddings.py:
      // 21-32) and travels in the checkpoint: use ITS values (no re-derivation with other sin / cos / exp roundings)
      const HostTensor* pe = find(e, "decoder.embeddings.make_embedding.pe.pe");
      if (!pe || pe->shape.size() != 3 || pe->shape[1] != 1 || pe->shape[2] != d)
        return fail(e, ND_ERR_WEIGHT, "position_encoding: checkpoint lacks decoder.embeddings.make_embedding.pe.pe [n,1,d]");
      if (pe->shape[0] < c.max_tgt_len)
        return fail(e, ND_ERR_WEIGHT, "position_encoding table is shorter than max_tgt_len");
      e->pe = upload(e, std::vector<float>(pe->f.begin(), pe->f.begin() + (size_t)c.max_tgt_len * d));
    }
  }
  if (c.decoder_type == ND_DEC_TRANSFORMER) {
    e->decT.resize(c.dec_layers);
    for (int l = 0; l < c.dec_layers; ++l) {
      const std::string p = "decoder.transformer_layers." + std::to_string(l);
      DecLayerT& L = e->decT[l];
      if (c.self_attn_average) {
        // onmt/modules/average_attn.py:22-30: PositionwiseFeedForward(d, d) + Linear(2d, 2d) over cat(inputs, average)
        const std::string a = p + ".self_attn.average_layer";
        ND_TRY(load_lin(e, a + ".w_1", d, d, true, &L.avg_w1, a + ".layer_norm"));
        ND_TRY(load_lin(e, a + ".w_2", d, d, true, &L.avg_w2));
        Lin gate;
        ND_TRY(load_lin(e, p + ".self_attn.gating_layer", 2 * d, 2 * d, true, &gate));
        L.gate_x = col_slice(gate, 0, d, true);
        L.gate_a = col_slice(gate, d, d, false);
      } else {
        ND_TRY(load_cat_lin(e, {p + ".self_attn.linear_query", p + ".self_attn.linear_keys", p + ".self_attn.linear_values"}, d, d, &L.qkv, p + ".layer_norm_1"));
        ND_TRY(load_lin(e, p + ".self_attn.final_linear", d, d, true, &L.self_out));
      }
      ND_TRY(load_lin(e, p + ".context_attn.linear_query", d, d, true, &L.cq, p + ".layer_norm_2"));
      ND_TRY(load_cat_lin(e, {p + ".context_attn.linear_keys", p + ".context_attn.linear_values"}, d, d, &L.ckv));
      ND_TRY(load_lin(e, p + ".context_attn.final_linear", d, d, true, &L.ctx_out));
      if (cross_attention_mb_supported(d, c.heads))
        ND_TRY(load_cross_mb(e, p + ".context_attn", p + ".layer_norm_2", d, c.heads, &L.cqt, &L.cm));
      ND_TRY(load_lin(e, p + ".feed_forward.w_1", c.d_ff, d, true, &L.w1, p + ".feed_forward.layer_norm"));
      ND_TRY(load_lin(e, p + ".feed_forward.w_2", d, c.d_ff, true, &L.w2));
      ND_TRY(load_ln(e, p + ".layer_norm_1", d, &L.ln1));
      ND_TRY(load_ln(e, p + ".layer_norm_2", d, &L.ln2));
      ND_TRY(load_ln(e, p + ".feed_forward.layer_norm", d, &L.ln_ff));
      if (c.self_attn_average) ND_TRY(load_ln(e, p + ".self_attn.average_layer.layer_norm", d, &L.ln_avg));
    }
    ND_TRY(load_ln(e, "decoder.layer_norm", d, &e->dec_ln));
  } else if (c.decoder_type == ND_DEC_RNN) {
    e->cells.resize(c.dec_layers);
    for (int l = 0; l < c.dec_layers; ++l) {
      // InputFeedRNNDecoder: StackedLSTM of LSTMCells "decoder.rnn.layers.<l>.weight_ih" (stacked_rnn.py:15-20);
      // StdRNNDecoder (-input_feed 0, decoder.py:187-262): one multi-layer nn.LSTM "decoder.rnn.weight_ih_l<l>"
      const std::string p = c.input_feed ? "decoder.rnn.layers." + std::to_string(l) + "." : "decoder.rnn.";
      const std::string sfx = c.input_feed ? "" : "_l" + std::to_string(l);
      const int in = l == 0 ? (c.input_feed ? 2 * d : d) : d;
      const int ng = c.rnn_type == ND_RNN_GRU ? 3 : 4;       // StackedGRU / nn.GRU (stacked_rnn.py:39-65)
      const HostTensor *wi, *wh, *bi, *bh;
      ND_TRY(need(e, p + "weight_ih" + sfx, {ng * d, in}, &wi));
      ND_TRY(need(e, p + "weight_hh" + sfx, {ng * d, d}, &wh));
      ND_TRY(need(e, p + "bias_ih" + sfx, {ng * d}, &bi));
      ND_TRY(need(e, p + "bias_hh" + sfx, {ng * d}, &bh));
      e->cells[l].ih = make_lin(e, wi->f, ng * d, in, &bi->f);
      e->cells[l].hh = make_lin(e, wh->f, ng * d, d, &bh->f);
      if (l == 0 && c.input_feed) {
        e->cells[l].ih_e = col_slice(e->cells[l].ih, 0, d, true);      // embedding columns (+ b_ih)
        e->cells[l].ih_f = col_slice(e->cells[l].ih, d, d, false);     // input-feed columns
      }
    }
    if (c.attn_type == ND_ATTN_MLP) {
      ND_TRY(load_lin(e, "decoder.attn.linear_context", d, d, false, &e->attn_ctx));
      ND_TRY(load_lin(e, "decoder.attn.linear_query", d, d, true, &e->attn_q));
      const HostTensor* v;
      ND_TRY(need(e, "decoder.attn.v.weight", {1, d}, &v));
      e->attn_v = upload(e, v->f);
      Lin out;
      ND_TRY(load_lin(e, "decoder.attn.linear_out", d, 2 * d, true, &out));
      e->attn_out_c = col_slice(out, 0, d, true);
      e->attn_out_h = col_slice(out, d, d, false);
    } else {
      if (c.attn_type == ND_ATTN_GENERAL) ND_TRY(load_lin(e, "decoder.attn.linear_in", d, d, false, &e->attn_in));
      Lin out;
      ND_TRY(load_lin(e, "decoder.attn.linear_out", d, 2 * d, false, &out));
      e->attn_out_c = col_slice(out, 0, d, false);
      e->attn_out_h = col_slice(out, d, d, false);
    }
  } else if (c.decoder_type == ND_DEC_CNN) {
    ND_TRY(load_lin(e, "decoder.linear", d, d, true, &e->dec_lin));
    e->dec_conv.resize(c.dec_layers);
    e->dec_attn_in.resize(c.dec_layers);
    for (int l = 0; l < c.dec_layers; ++l) {
      ND_TRY(load_wnconv(e, "decoder.conv_layers." + std::to_string(l) + ".conv", d, c.cnn_kernel_width, &e->dec_conv[l].conv));
      ND_TRY(load_lin(e, "decoder.attn_layers." + std::to_string(l) + ".linear_in", d, d, true, &e->dec_attn_in[l]));
    }
  } else {
    return fail(e, ND_ERR_INVALID, "unknown decoder_type");
  }
  {
    const HostTensor *gw, *gb;
    ND_TRY(need(e, "generator.0.weight", {V, d}, &gw));
    ND_TRY(need(e, "generator.0.bias", {V}, &gb));
    e->gen.W = upload(e, gw->f);
    e->gen.b = upload(e, gb->f);
    e->gen.N = V; e->gen.K = d; e->gen.ld = d;
  }
  if (e->oom) return fail(e, ND_ERR_NOMEM, "device allocation failed while packing weights: " + e->err);
  // A floating-point tensor nothing consumed means the checkpoint holds a module this engine does not run (e.g.
  // encoder.bridge.*): decoding would silently produce other bases than the reference.  Known aliases / dead weights:
  // the weight-norm convs' training copies (eval uses the *_avg buffers, onmt/modules/weight_norm.py:154-156).
  std::string left;
  int n_left = 0;
  for (auto& kv : e->raw) {
    const std::string& k = kv.first;
    if (e->used.count(k)) continue;
    auto ends = [&](const char* sfx) { const size_t n = strlen(sfx); return k.size() >= n && k.compare(k.size() - n, n, sfx) == 0; };
    if (ends(".conv.weight") || ends(".conv.bias") || ends(".conv.V") || ends(".conv.g") || ends(".conv.b")) continue;
    if (n_left++ < 6) left += (left.empty() ? "" : ", ") + k;
  }
  if (n_left)
    return fail(e, ND_ERR_WEIGHT, "checkpoint holds " + std::to_string(n_left) + " tensor(s) this engine does not use (" +
                                      left + (n_left > 6 ? ", ..." : "") + "): unsupported model variant");
  e->raw.clear();
  return ND_OK;
}

// ------------------------------------------------------------------------------------------ workspace
int alloc_workspace(nd_engine* e) {
  const nd_config& c = e->cfg;
  const int64_t B = c.max_batch, T = c.max_src_len, d = c.d_model, L = c.max_tgt_len, K = c.max_beam;
  const int64_t BT = B * T, rows = B * K;
  e->max_rows = (int)rows;
  bool ok = true;
  auto F = [&](float*& p, int64_t n) { p = dalloc<float>(e, (size_t)n); ok = ok && p; };
  F(e->src, BT);
  e->lengths = dalloc<int64_t>(e, B);
  e->mem_len = dalloc<int64_t>(e, B);
  F(e->bufA, BT * d); F(e->bufB, BT * d); F(e->mb, BT * d);
  int64_t wide = 4 * d;                                    // LSTM input projection (2 dirs * 4 * d/2)
  if (c.encoder_type == ND_ENC_NANO || c.encoder_type == ND_ENC_CRNN)
    for (int l = 0; l < c.enc_layers; ++l)
      if (c.enc_pooling[l] > 1 && !e->bufC) F(e->bufC, BT * d);
  if (c.encoder_type == ND_ENC_TRANSFORMER || c.encoder_type == ND_ENC_CTRANSFORMER) {
    wide = std::max<int64_t>(3 * d, c.d_ff);
    F(e->bufC, BT * d);
  }
  if (c.encoder_type == ND_ENC_RESNET || c.encoder_type == ND_ENC_CRNN || c.encoder_type == ND_ENC_CTRANSFORMER) {
    wide = std::max<int64_t>(wide, 3 * 512);               // im2col rows of the widest stem layer
    for (int i = 0; i < 3; ++i) F(e->rbuf[i], BT * 512);
    if (c.encoder_type == ND_ENC_CTRANSFORMER) F(e->cmask, BT);
  }
  if (c.encoder_type == ND_ENC_CNN) { wide = (int64_t)c.cnn_kernel_width * d; F(e->big2, BT * 2 * d); F(e->emb_remap, BT * d); }
  F(e->big, BT * wide);
  if (c.encoder_type == ND_ENC_BRNN || c.encoder_type == ND_ENC_RNN) {
    F(e->enc_hn, (int64_t)c.enc_layers * B * d); F(e->enc_cn, (int64_t)c.enc_layers * B * d);
    if (c.bridge) { F(e->enc_hb, (int64_t)c.enc_layers * B * d); F(e->enc_cb, (int64_t)c.enc_layers * B * d); }
  }
  // decoder
  F(e->logp, rows * c.vocab_size); F(e->gscore, rows);
  e->cur_tok = dalloc<int>(e, rows);
  if (c.decoder_type == ND_DEC_TRANSFORMER) {
    e->ckv.resize(c.dec_layers); e->selfK.resize(c.dec_layers); e->selfV.resize(c.dec_layers);
    for (int l = 0; l < c.dec_layers; ++l) { F(e->ckv[l], BT * 2 * d); F(e->selfK[l], rows * L * d); F(e->selfV[l], rows * L * d); }
    F(e->x, rows * d); F(e->qkv, rows * 3 * d); F(e->sctx, rows * d); F(e->x1, rows * d); F(e->qc, rows * d);
    F(e->cctx, rows * d); F(e->x2, rows * d); F(e->ffh, rows * c.d_ff);
    if (cross_attention_mb_supported((int)d, c.heads)) { F(e->qt, B * c.heads * d); F(e->cctxt, B * c.heads * d); }
  } else if (c.decoder_type == ND_DEC_RNN) {
    F(e->uh, BT * d);
    if (kv_pack_supported((int)d)) F(e->pk, BT * 2 * d);
    for (int s = 0; s < 2; ++s) {
      e->rh[s].resize(c.dec_layers); e->rc[s].resize(c.dec_layers);
      for (int l = 0; l < c.dec_layers; ++l) { F(e->rh[s][l], rows * d); F(e->rc[s][l], rows * d); }
      F(e->feed[s], rows * d);
    }
    F(e->x, rows * d); F(e->ga, rows * 4 * d); F(e->gb, rows * 4 * d); F(e->wq, rows * d); F(e->actx, rows * d);
  } else {
    F(e->enc_comb, BT * d);
    if (kv_pack_supported((int)d)) F(e->pk, BT * 2 * d);
    for (int s = 0; s < 2; ++s) {
      e->chist[s].resize(c.dec_layers);
      for (int l = 0; l < c.dec_layers; ++l) F(e->chist[s][l], rows * L * d);
    }
    F(e->x, rows * d); F(e->cbase, rows * d); F(e->cA, rows * c.cnn_kernel_width * d); F(e->cy, rows * 2 * d);
    F(e->cout, rows * d); F(e->cpre, rows * d); F(e->ctgt, rows * d); F(e->cctx2, rows * d); F(e->x1, rows * d);
  }
  e->o_ids = dalloc<int64_t>(e, (size_t)rows * L);
  e->o_scores = dalloc<float>(e, (size_t)rows);
  e->o_lens = dalloc<int>(e, (size_t)rows);
  ok = ok && e->o_ids && e->o_scores && e->o_lens;
  // beam state
  BeamState& b = e->beam;
  b.topk_log_probs = dalloc<float>(e, rows);
  b.alive_seq = dalloc<int>(e, 2 * rows * (L + 1));
  b.anc = dalloc<int>(e, 2 * rows * L);
  b.cur_tok = e->cur_tok;
  b.parent = dalloc<int>(e, rows);
  b.retired = dalloc<int>(e, B);
  b.retire_step = dalloc<int>(e, B);
  b.top_finished = dalloc<int>(e, B);
  b.n_hyp = dalloc<int>(e, B);
  b.hyp_score = dalloc<float>(e, B * K);
  b.hyp_len = dalloc<int>(e, B * K);
  b.hyp_seq = dalloc<int>(e, B * K * L);
  b.hyp_anc = dalloc<int>(e, B * K * L);
  b.hyp_meta = dalloc<int>(e, B * K);
  b.n_alive = dalloc<int>(e, 1);
  b.n_done = dalloc<int>(e, 1);
  b.stop_step = dalloc<int>(e, 1);
  ok = ok && b.n_done && b.stop_step && e->lengths && e->mem_len && e->cur_tok && b.topk_log_probs && b.alive_seq && b.anc && b.parent &&
       b.retired && b.top_finished && b.n_hyp && b.hyp_score && b.hyp_len && b.hyp_seq && b.hyp_anc && b.hyp_meta && b.n_alive && b.retire_step;
  return ok ? ND_OK : ND_ERR_NOMEM;
}

// ------------------------------------------------------------------------------------------ encoders
// ResNet stem: src [B,T] -> out [B*T, d].  conv(1 -> 64) + BN + ReLU, 8 BasicBlocks
// (relu(bn2(conv2(relu(bn1(conv1 x)))) + residual), encoder/resnet_encoder.py:30-47), Linear(512, d).
int encode_resnet_stem(nd_engine* e, float* out, cudaStream_t st) {
  const ResNetW& R = e->resnet;
  const int B = e->B, T = e->T;
  const int64_t M = (int64_t)B * T;
  float* cur = e->rbuf[0];
  float* t1 = e->rbuf[1];
  float* spare = e->rbuf[2];
  ND_LAUNCH(e, resnet_stem_conv(e->src, R.stem_w, R.stem_b, cur, B, T, st));
  for (const ResBlockW& K : R.blocks) {
    ND_LAUNCH(e, im2col_time(cur, e->big, B, T, K.cin, 3, 1, st));
    GemmOpt o1; o1.act = 1;
    ND_TRY(run_gemm(e, K.conv1, e->big, 3 * (int64_t)K.cin, t1, K.cout, M, o1, st));
    ND_LAUNCH(e, im2col_time(t1, e->big, B, T, K.cout, 3, 1, st));
    GemmOpt o2; o2.act = 3; o2.ldr = K.cout;             // ReLU after the residual add
    if (K.has_down) {
      GemmOpt od;
      ND_TRY(run_gemm(e, K.down, cur, K.cin, spare, K.cout, M, od, st));
      o2.residual = spare;
      ND_TRY(run_gemm(e, K.conv2, e->big, 3 * (int64_t)K.cout, cur, K.cout, M, o2, st));   // cur is free by now
    } else {
      o2.residual = cur;
      ND_TRY(run_gemm(e, K.conv2, e->big, 3 * (int64_t)K.cout, spare, K.cout, M, o2, st));
      std::swap(cur, spare);
    }
  }
  GemmOpt of;
  ND_TRY(run_gemm(e, R.fc, cur, 512, out, e->cfg.d_model, M, of, st));
  return ND_OK;
}

// encoder/resnet_encoder.py:227-251 (ResNetForRNNEncoder): memory bank = the stem output, lengths unchanged, zero state
int encode_resnet(nd_engine* e, cudaStream_t st) {
  ND_TRY(encode_resnet_stem(e, e->mb, st));
  e->Tp = e->T;
  ND_CUDA(e, cudaMemcpyAsync(e->mem_len, e->lengths, (size_t)e->B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return ND_OK;
}

int encode_lstm_stack(nd_engine* e, cudaStream_t st) {
  const nd_config& c = e->cfg;
  const int d = c.d_model, B = e->B;
  const bool nano = c.encoder_type == ND_ENC_NANO || c.encoder_type == ND_ENC_CRNN;   // pool / BatchNorm / W stack
  const int dirs = (c.encoder_type == ND_ENC_RNN) ? 1 : 2;
  const int H = d / dirs;
  int T = e->T;
  // lengths the layers see: the source lengths until a pooling layer shortens them -- on the device (the reference
  // does lengths.tolist() and the arithmetic in Python, nano_encoder.py:90,103-104; here nothing leaves the stream)
  const int64_t* lens_dev = e->lengths;
  const float* in = nullptr;                // previous layer output [B,T,d]
  if (c.encoder_type == ND_ENC_CRNN) {
    // crnn_encoder.py:97-98: the ResNet stem's [B,T,d] output feeds the first LSTM layer.  It is parked in the memory
    // bank buffer, which nothing else touches until the stack's last step writes W . out into it.
    ND_TRY(encode_resnet_stem(e, e->mb, st));
    in = e->mb;
  }
  float* outs[2] = {e->bufA, e->bufB};
  float* last = nullptr;
  for (int l = 0; l < c.enc_layers; ++l) {
    const LstmW& W = e->lstm[l];
    float* out = outs[l & 1];
    LstmParams p;
    p.B = B; p.T = T; p.dirs = dirs; p.H = H; p.cell = c.rnn_type == ND_RNN_GRU ? 1 : 0;
    p.w_hh = W.w_hh; p.b_hh = W.b_hh; p.lengths = lens_dev; p.out = out;
    if (W.in == 1) {
      p.x0 = e->src; p.w_ih0 = W.w_ih0; p.b_ih0 = W.b_ih0;
    } else {
      GemmOpt o;
      if (nano && l > 0) { o.prologue = PRO_AFFINE; o.pg = e->lstm[l - 1].bn_alpha; o.pb = e->lstm[l - 1].bn_beta; }
      const int64_t ng = p.cell ? 3 : 4;
      ND_TRY(run_gemm(e, W.ih, in, d, e->big, dirs * ng * H, (int64_t)B * T, o, st));
      p.xg = e->big; p.xg_ld = dirs * ng * H;
    }
    if (!nano) {
      p.h_n = e->enc_hn + (int64_t)l * dirs * B * H;
      p.c_n = p.cell ? nullptr : e->enc_cn + (int64_t)l * dirs * B * H;
    }
    ND_CUDA(e, cudaMemsetAsync(out, 0, (size_t)B * T * d * sizeof(float), st));
    if (tc_mode(e) && lstm_tc_supported(H) && W.tc_ok) ND_LAUNCH_CAT(e, ND_PROF_LSTM, st, lstm_layer_tc(p, st));
    else ND_LAUNCH_CAT(e, ND_PROF_LSTM, st, lstm_layer(p, e->n_sm, st));
    last = out;
    if (nano && c.enc_pooling[l] > 1) {
      // MaxPool1d over time (nano_encoder.py:101-105); lengths follow floor((len - s)/s + 1)
      const int s = c.enc_pooling[l];
      const int Tn = T / s;
      float* pooled = e->bufC;
      ND_LAUNCH(e, maxpool_time(out, pooled, B, T, d, s, st));
      ND_CUDA(e, cudaMemcpyAsync(out, pooled, (size_t)B * Tn * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
      T = Tn;
      ND_LAUNCH(e, pool_lengths(lens_dev, e->mem_len, B, s, st));       // mem_len doubles as the working copy
      lens_dev = e->mem_len;
    }
    in = out;
  }
  e->Tp = T;
  if (nano) {
    // memory bank = W . (pre-BatchNorm pooled output of the last layer)          nano_encoder.py:113-115
    GemmOpt o;
    ND_TRY(run_gemm(e, e->encW, last, d, e->mb, d, (int64_t)B * T, o, st));
  } else {
    ND_CUDA(e, cudaMemcpyAsync(e->mb, last, (size_t)B * T * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (c.bridge) {
      // encoder/rnn_encoder.py:101-118: relu(Linear(states.view(-1, hh * layers))).view(size) -- the view runs over the
      // flat [layers*dirs, B, hh] array, so one row is `layers` consecutive (layer-dir, chunk) vectors: reproduced as is
      const int tot = H * c.enc_layers;
      const int64_t rows_b = (int64_t)dirs * B;
      GemmOpt ob; ob.act = 1;
      ND_TRY(run_gemm(e, e->bridge_h, e->enc_hn, tot, e->enc_hb, tot, rows_b, ob, st));
      if (c.rnn_type != ND_RNN_GRU) ND_TRY(run_gemm(e, e->bridge_c, e->enc_cn, tot, e->enc_cb, tot, rows_b, ob, st));
    }
  }
  if (lens_dev != e->mem_len)
    ND_CUDA(e, cudaMemcpyAsync(e->mem_len, e->lengths, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return ND_OK;
}

int encode_transformer(nd_engine* e, cudaStream_t st) {
  const nd_config& c = e->cfg;
  const int d = c.d_model, B = e->B, T = e->T;
  const int64_t M = (int64_t)B * T;
  const float sq = sqrtf((float)(d / c.heads));
  float* x = e->bufA;
  float* x1 = e->bufB;
  float* ctx = e->bufC;
  const float* mask_src = e->src;           // key t is masked where this is 0.0 (encoder/transformer.py:117-121)
  if (c.encoder_type == ND_ENC_CTRANSFORMER) {
    // encoder/ctransformer.py:75-84: the stem replaces Linear(1,d), and the mask is read from channel 0 of ITS output
    ND_TRY(encode_resnet_stem(e, x, st));
    ND_LAUNCH(e, take_column(x, d, 0, e->cmask, M, st));
    mask_src = e->cmask;
  } else {
    ND_LAUNCH(e, linear_in1(e->src, e->enc_lin_in.W, e->enc_lin_in.b, x, M, d, st));     // encoder/transformer.py:113
  }
  for (int l = 0; l < c.enc_layers; ++l) {
    const EncLayerT& L = e->encT[l];
    GemmOpt o1; o1.prologue = PRO_LAYERNORM; o1.pg = L.ln.g; o1.pb = L.ln.b;
    ND_TRY(run_gemm(e, L.qkv, x, d, e->big, 3 * d, M, o1, st));
    EncAttnParams a; a.qkv = e->big; a.q_div = sq; a.src = mask_src; a.ctx = ctx; a.B = B; a.T = T; a.d = d; a.H = c.heads;
    if (tc_mode(e) && e->enc_attn_tc && encoder_attention_tc_supported(a))
      ND_LAUNCH_CAT(e, ND_PROF_ENC_ATTN, st, encoder_attention_tc(a, st));
    else
      ND_LAUNCH_CAT(e, ND_PROF_ENC_ATTN, st, encoder_attention(a, st));
    GemmOpt o2; o2.residual = x; o2.ldr = d;
    ND_TRY(run_gemm(e, L.out, ctx, d, x1, d, M, o2, st));
    GemmOpt o3; o3.prologue = PRO_LAYERNORM; o3.pg = L.ln_ff.g; o3.pb = L.ln_ff.b; o3.act = 1;
    ND_TRY(run_gemm(e, L.w1, x1, d, e->big, c.d_ff, M, o3, st));
    GemmOpt o4; o4.residual = x1; o4.ldr = d;
    ND_TRY(run_gemm(e, L.w2, e->big, c.d_ff, x, d, M, o4, st));
  }
  ND_LAUNCH(e, layernorm_rows(x, e->enc_ln.g, e->enc_ln.b, 1e-6f, e->mb, M, d, st));
  e->Tp = T;
  ND_CUDA(e, cudaMemcpyAsync(e->mem_len, e->lengths, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return ND_OK;
}

int encode_cnn(nd_engine* e, cudaStream_t st) {
  const nd_config& c = e->cfg;
  const int d = c.d_model, B = e->B, T = e->T, k = c.cnn_kernel_width;
  const int64_t M = (int64_t)B * T;
  ND_LAUNCH(e, linear_in1(e->src, e->enc_lin_in.W, e->enc_lin_in.b, e->emb_remap, M, d, st));
  const float* x = e->emb_remap;
  float* outs[2] = {e->bufA, e->bufB};
  for (int l = 0; l < c.enc_layers; ++l) {
    ND_LAUNCH(e, im2col_time(x, e->big, B, T, d, k, k / 2, st));
    GemmOpt o;
    ND_TRY(run_gemm(e, e->enc_conv[l].conv, e->big, (int64_t)k * d, e->big2, 2 * d, M, o, st));
    float* out = (l + 1 == c.enc_layers) ? e->mb : outs[l & 1];
    ND_LAUNCH(e, glu_residual(e->big2, x, out, nullptr, M, d, st));
    x = out;
  }
  e->Tp = T;
  ND_CUDA(e, cudaMemcpyAsync(e->mem_len, e->lengths, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return ND_OK;
}

// ------------------------------------------------------------------------------------------ decoders
struct DecodeCtx {
  int K = 1, step = 0, Lmax = 0;
  int c0 = 0, nc = 0;                 // chunk range handled by this call; rows = [c0*K, (c0+nc)*K)
  bool beam = false;
  bool cross_mb = false;              // cross attention in memory-bank space (greedy, no attention output)
  float* attn_out = nullptr;          // optional [rows_total, T'] for this step
};

// planes of layer l inside its K/V buffer (sized for fp32 [maxB*maxT, 2d]); offsets follow the LAST encode's B*T'
struct KvPlanes { int16_t* hi; uint8_t* lo; float* scale; };
KvPlanes kv_planes(const nd_engine* e, int l) {
  const size_t n = (size_t)e->B * e->Tp * 2 * e->cfg.d_model;
  uint8_t* base = reinterpret_cast<uint8_t*>(e->ckv[l]);
  KvPlanes p;
  p.hi = reinterpret_cast<int16_t*>(base);
  p.lo = base + 2 * n;
  p.scale = reinterpret_cast<float*>(base + 3 * n);
  return p;
}

// RNN / CNN decoders: the same plane layout inside e->pk
KvPlanes pk_planes(const nd_engine* e) {
  const size_t n = (size_t)e->B * e->Tp * 2 * e->cfg.d_model;
  uint8_t* base = reinterpret_cast<uint8_t*>(e->pk);
  KvPlanes p;
  p.hi = reinterpret_cast<int16_t*>(base);
  p.lo = base + 2 * n;
  p.scale = reinterpret_cast<float*>(base + 3 * n);
  return p;
}

bool use_cross_mb(const nd_engine* e, int K, bool want_attn) {
  return e->cross_mode == 1 && K == 1 && !want_attn && e->cfg.decoder_type == ND_DEC_TRANSFORMER && e->qt != nullptr;
}

int decoder_init(nd_engine* e, int K, cudaStream_t st, bool cross_mb = false) {
  NvtxRange nvtx("nd:decoder_init (memory keys/values, attention pre-projection, state)");
  const nd_config& c = e->cfg;
  const int d = c.d_model, B = e->B, Tp = e->Tp;
  const int64_t M = (int64_t)B * Tp;
  if (c.decoder_type == ND_DEC_TRANSFORMER) {
    if (Tp != e->T)
      return fail(e, ND_ERR_INVALID, "transformer decoder needs -audio_enc_pooling 1 (the reference builds its cross "
                                     "mask from the un-pooled signal, decoder/transformer.py:201-221)");
    // memory keys / values, projected once per chunk (multi_headed_attn.py:142-153); not needed when the
    // cross attention runs in memory-bank space
    // Packed planes: greedy decode only (the beam kernels of cross_attn_ring.cu read fp32 rows); the GEMM writes the
    // fp32 projection into the encoder's scratch (free once the memory bank exists), the packer re-writes it as
    // [int16 plane | uint8 plane | steps] into the layer's K/V buffer (3/4 or 1/2 of it).
    // (beam search: only where the multi-query slice kernel exists -- d = 256 / 512, the beams' score rows fit two CTAs
    // per SM -- and the option kv_beam_packed is on; otherwise fp32 rows through the ring kernel)
    e->kv_packed = !cross_mb && e->kv_mode != KV_F32 && kv_pack_supported(d) &&
                   (K == 1 || (e->kv_beam_packed && cross_attention_packed_beams_ok(K, d, c.heads, Tp, e->kv_mode)));
    for (int l = 0; l < c.dec_layers && !cross_mb; ++l) {
      GemmOpt o;
      if (!e->kv_packed) {
        ND_TRY(run_gemm(e, e->decT[l].ckv, e->mb, d, e->ckv[l], 2 * d, M, o, st));
        continue;
      }
      ND_TRY(run_gemm(e, e->decT[l].ckv, e->mb, d, e->big, 2 * d, M, o, st));
      KvPlanes pl = kv_planes(e, l);
      ND_LAUNCH(e, kv_pack(e->big, M, d, e->kv_mode, pl.hi, pl.lo, pl.scale, st));
    }
  } else if (c.decoder_type == ND_DEC_RNN) {
    const int rows = B * K;
    if (c.attn_type == ND_ATTN_MLP) {
      GemmOpt o;                     // uh = Uk . H, exact same product the reference recomputes every step
      ND_TRY(run_gemm(e, e->attn_ctx, e->mb, d, e->uh, d, M, o, st));
    }
    // The global attention re-reads uh and H (mlp) or H twice (general / dot) at every step: same fixed-point storage
    // as the Transformer decoder's memory keys / values, part "K" = what the scores read, part "V" = H
    e->kv_packed = e->kv_mode != KV_F32 && e->pk != nullptr && (e->kv_mode == KV_Q23M || e->kv_mode == KV_Q15M);
    e->pk_fmt = e->kv_mode;
    if (e->kv_packed) {
      const KvPlanes pl = pk_planes(e);
      ND_LAUNCH(e, kv_pack2(c.attn_type == ND_ATTN_MLP ? e->uh : e->mb, d, e->mb, d, M, d, e->pk_fmt, pl.hi, pl.lo, pl.scale, st));
    }
    // decoder.py:108-129: hidden from the encoder final state ([fwd;bwd] per layer), input_feed = 0
    for (int l = 0; l < c.dec_layers; ++l) {
      ND_CUDA(e, cudaMemsetAsync(e->rh[0][l], 0, (size_t)rows * d * sizeof(float), st));
      ND_CUDA(e, cudaMemsetAsync(e->rc[0][l], 0, (size_t)rows * d * sizeof(float), st));
    }
    ND_CUDA(e, cudaMemsetAsync(e->feed[0], 0, (size_t)rows * d * sizeof(float), st));
    if (c.encoder_type == ND_ENC_BRNN || c.encoder_type == ND_ENC_RNN) {
      if (c.enc_layers != c.dec_layers) return fail(e, ND_ERR_INVALID, "rnn encoder/decoder layer counts differ");
      const int dirs = c.encoder_type == ND_ENC_BRNN ? 2 : 1;
      const int H = d / dirs;
      for (int l = 0; l < c.dec_layers; ++l)
        for (int dir = 0; dir < dirs; ++dir)
          for (int k = 0; k < K; ++k) {
            // rows are chunk-major: row = b*K + k; copy with a strided 2-D memcpy (width H floats)
            const float* hs = (c.bridge ? e->enc_hb : e->enc_hn) + ((int64_t)l * dirs + dir) * B * H;
            const float* cs = (c.bridge ? e->enc_cb : e->enc_cn) + ((int64_t)l * dirs + dir) * B * H;
            ND_CUDA(e, cudaMemcpy2DAsync(e->rh[0][l] + (int64_t)k * d + dir * H, (size_t)K * d * sizeof(float), hs,
                                         (size_t)H * sizeof(float), (size_t)H * sizeof(float), B,
                                         cudaMemcpyDeviceToDevice, st));
            if (c.rnn_type != ND_RNN_GRU)
              ND_CUDA(e, cudaMemcpy2DAsync(e->rc[0][l] + (int64_t)k * d + dir * H, (size_t)K * d * sizeof(float), cs,
                                           (size_t)H * sizeof(float), (size_t)H * sizeof(float), B,
                                           cudaMemcpyDeviceToDevice, st));
          }
    }
  } else {
    if (c.encoder_type != ND_ENC_CNN)
      return fail(e, ND_ERR_INVALID, "cnn decoder needs the cnn encoder (its init_state adds the encoder's "
                                     "input projection, onmt/decoders/cnn_decoder.py:59-64)");
    // state["src"] = (memory_bank + enc_hidden) * SCALE_WEIGHT                      cnn_decoder.py:63
    ND_LAUNCH(e, add_scale(e->mb, e->emb_remap, 0.70710678118654757f, e->enc_comb, M * d, st));
    // conv attention: keys = encoder top, values = the combined state; read by every layer at every step.  Always the
    // 3-byte format: the conv attention's scores are raw dot products of un-normalised states and its softmax is peaked;
    // 2-byte keys moved one logit in 1024 chunks by 4e-3 (profiles/r02x_identity_rnn_cnn.json), which no stated bound
    // worth having covers, so kv_mode q15 means q23 for this decoder
    e->kv_packed = e->kv_mode != KV_F32 && e->pk != nullptr && (e->kv_mode == KV_Q23M || e->kv_mode == KV_Q15M);
    e->pk_fmt = KV_Q23M;
    if (e->kv_packed) {
      const KvPlanes pl = pk_planes(e);
      ND_LAUNCH(e, kv_pack2(e->mb, d, e->enc_comb, d, M, d, e->pk_fmt, pl.hi, pl.lo, pl.scale, st));
    }
  }
  return ND_OK;
}

// one decoder step for the chunk range of dc: cur_tok -> logp [rows, V] (and greedy selection when gp.ids != null)
int decoder_step(nd_engine* e, const DecodeCtx& dc, GenParams gp, cudaStream_t st) {
  const nd_config& c = e->cfg;
  const int d = c.d_model, Tp = e->Tp, K = dc.K;
  const int r0 = dc.c0 * K, rows = dc.nc * K, B_total = e->B;
  const int* retired = dc.beam ? e->beam.retired : nullptr;
  const int* anc = (dc.beam && dc.step > 0) ? e->beam.anc + (int64_t)(dc.step & 1) * B_total * K * dc.Lmax : nullptr;
  auto R = [&](float* p, int64_t w) { return p + (int64_t)r0 * w; };            // row-range view
  const float s2 = 0.70710678118654757f;                                        // SCALE_WEIGHT = 0.5 ** 0.5
  // position of this step's token in the sinusoid table: the Transformer decoder passes `step` (decoder/transformer.py:
  // 214), the CNN decoder embeds the whole prefix (position = index, cnn_decoder.py:89), the RNN decoders call the
  // embeddings WITHOUT a step on a one-token input, i.e. every token gets row 0 (onmt/decoders/decoder.py:323)
  const int pe_pos = c.decoder_type == ND_DEC_RNN ? 0 : dc.step;
  ND_LAUNCH(e, embed_rows(e->cur_tok + r0, e->emb, R(e->x, d), d, rows, d, e->pe ? e->pe + (int64_t)pe_pos * d : nullptr,
                          (float)sqrt((double)d), st));
  if (c.decoder_type == ND_DEC_TRANSFORMER) {
    const float sq = sqrtf((float)(d / c.heads));
    float* x = R(e->x, d);
    for (int l = 0; l < c.dec_layers; ++l) {
      const DecLayerT& L = e->decT[l];
      if (c.self_attn_average) {
        // AverageAttention (onmt/modules/average_attn.py:77-106, decoder/transformer.py:82-85): cumulative average of the
        // normalised inputs along the hypothesis (state prev_g per row, followed through the beams' parents), FFN, gates
        float* xn = R(e->sctx, d);
        // the state of even / odd steps lives in the (otherwise unused) self-attention K / V cache buffers of the layer
        float* g_cur = (dc.step & 1) ? e->selfV[l] : e->selfK[l];                        // [rows, d]
        const float* g_prev = (dc.step & 1) ? e->selfK[l] : e->selfV[l];
        ND_LAUNCH(e, layernorm_rows(x, L.ln1.g, L.ln1.b, 1e-6f, xn, rows, d, st));
        ND_LAUNCH(e, avg_attn_cumulate(xn, g_prev, dc.beam ? e->beam.parent : nullptr, r0, rows, d, dc.step,
                                       g_cur + (int64_t)r0 * d, st));
        GemmOpt oa1; oa1.prologue = PRO_LAYERNORM; oa1.pg = L.ln_avg.g; oa1.pb = L.ln_avg.b; oa1.act = 1;
        ND_TRY(run_gemm(e, L.avg_w1, g_cur + (int64_t)r0 * d, d, R(e->qc, d), d, rows, oa1, st));
        GemmOpt oa2; oa2.residual = g_cur + (int64_t)r0 * d; oa2.ldr = d;
        ND_TRY(run_gemm(e, L.avg_w2, R(e->qc, d), d, R(e->cctx, d), d, rows, oa2, st));     // a = FFN(g)
        float* gate = R(e->qkv, 3 * d);                                                  // [rows, 2d] view, pitch 2d
        GemmOpt og1;
        ND_TRY(run_gemm(e, L.gate_x, xn, d, gate, 2 * d, rows, og1, st));
        GemmOpt og2; og2.residual = gate; og2.ldr = 2 * d;
        ND_TRY(run_gemm(e, L.gate_a, R(e->cctx, d), d, gate, 2 * d, rows, og2, st));
        ND_LAUNCH(e, avg_attn_gate(gate, xn, R(e->cctx, d), x, R(e->x1, d), rows, d, st));
      } else {
      GemmOpt o1; o1.prologue = PRO_LAYERNORM; o1.pg = L.ln1.g; o1.pb = L.ln1.b;
      ND_TRY(run_gemm(e, L.qkv, x, d, R(e->qkv, 3 * d), 3 * d, rows, o1, st));
      SelfAttnParams sa;
      sa.qkv = e->qkv; sa.q_div = sq; sa.Kc = e->selfK[l]; sa.Vc = e->selfV[l]; sa.ctx = e->sctx; sa.row0 = r0;
      sa.rows = rows; sa.d = d; sa.H = c.heads; sa.Lmax = dc.Lmax; sa.step = dc.step; sa.retired = retired;
      sa.rows_per_chunk = K; sa.anc = anc; sa.anc_ld = dc.Lmax;
      ND_LAUNCH_CAT(e, ND_PROF_SELF_ATTN, st, self_attention_step(sa, st));
      GemmOpt o2; o2.residual = x; o2.ldr = d;
      ND_TRY(run_gemm(e, L.self_out, R(e->sctx, d), d, R(e->x1, d), d, rows, o2, st));
      }
      if (dc.cross_mb) {
        // cross attention in memory-bank space: QT = LN2(x1) P^T + pb -> attention over mb -> x2 = x1 + ctxt Mcat^T + mo
        const int64_t HD = (int64_t)c.heads * d;
        GemmOpt o3; o3.prologue = PRO_LAYERNORM; o3.pg = L.ln2.g; o3.pb = L.ln2.b;
        ND_TRY(run_gemm(e, L.cqt, R(e->x1, d), d, R(e->qt, HD), HD, rows, o3, st));
        CrossMbParams cm;
        cm.qt = R(e->qt, HD); cm.mb = e->mb + (int64_t)dc.c0 * Tp * d;
        cm.src = e->src + (int64_t)dc.c0 * e->T; cm.src_ld = e->T; cm.mask_value = 1.0f;   // decoder/transformer.py:219-221
        cm.ctxt = R(e->cctxt, HD); cm.n_chunks = dc.nc; cm.T = Tp; cm.d = d; cm.H = c.heads;
        ND_LAUNCH_CAT(e, ND_PROF_CROSS_ATTN, st, cross_attention_mb(cm, st));
        GemmOpt o4; o4.residual = R(e->x1, d); o4.ldr = d;
        ND_TRY(run_gemm(e, L.cm, R(e->cctxt, HD), HD, R(e->x2, d), d, rows, o4, st));
      } else {
        GemmOpt o3; o3.prologue = PRO_LAYERNORM; o3.pg = L.ln2.g; o3.pb = L.ln2.b;
        ND_TRY(run_gemm(e, L.cq, R(e->x1, d), d, R(e->qc, d), d, rows, o3, st));
        CrossAttnParams ca;
        ca.q = R(e->qc, d); ca.q_ld = d; ca.q_div = sq;
        ca.K = e->ckv[l] + (int64_t)dc.c0 * Tp * 2 * d; ca.V = ca.K + d; ca.kv_ld = 2 * d;
        if (e->kv_packed) {
          const KvPlanes pl = kv_planes(e, l);
          const int64_t r0k = (int64_t)dc.c0 * Tp;
          ca.kv_fmt = e->kv_mode; ca.kv_hi = pl.hi + r0k * 2 * d; ca.kv_lo = pl.lo + r0k * 2 * d; ca.kv_scale = pl.scale + r0k * 2;
        }
        ca.src = e->src + (int64_t)dc.c0 * e->T; ca.src_ld = e->T; ca.mask_value = 1.0f;   // decoder/transformer.py:219-221
        ca.retired = retired ? retired + dc.c0 : nullptr; ca.ctx = R(e->cctx, d); ca.ctx_ld = d; ca.n_chunks = dc.nc;
        ca.NQ = K; ca.T = Tp; ca.d = d; ca.H = c.heads;
        ca.attn = (l + 1 == c.dec_layers && dc.attn_out) ? dc.attn_out + (int64_t)r0 * Tp : nullptr;
        ND_LAUNCH_CAT(e, ND_PROF_CROSS_ATTN, st, cross_attention(ca, st));
        GemmOpt o4; o4.residual = R(e->x1, d); o4.ldr = d;
        ND_TRY(run_gemm(e, L.ctx_out, R(e->cctx, d), d, R(e->x2, d), d, rows, o4, st));
      }
      GemmOpt o5; o5.prologue = PRO_LAYERNORM; o5.pg = L.ln_ff.g; o5.pb = L.ln_ff.b; o5.act = 1;
      ND_TRY(run_gemm(e, L.w1, R(e->x2, d), d, R(e->ffh, c.d_ff), c.d_ff, rows, o5, st));
      GemmOpt o6; o6.residual = R(e->x2, d); o6.ldr = d;
      ND_TRY(run_gemm(e, L.w2, R(e->ffh, c.d_ff), c.d_ff, x, d, rows, o6, st));
    }
    gp.x = x; gp.x_ld = d; gp.ln_g = e->dec_ln.g; gp.ln_b = e->dec_ln.b;
  } else if (c.decoder_type == ND_DEC_RNN) {
    const int cur = dc.step & 1, nxt = cur ^ 1;
    if (dc.beam && dc.step > 0) {
      // reorder the recurrent state by parent beam (translator.py:820-821 map_state index_select): the state
      // written by step-1 sits in buffers [cur]; gather through the spare [nxt] buffers and copy back
      for (int l = 0; l < c.dec_layers; ++l) {
        ND_LAUNCH(e, gather_rows(e->rh[cur][l], e->rh[nxt][l], e->beam.parent, r0, rows, d, st));
        ND_LAUNCH(e, gather_rows(e->rc[cur][l], e->rc[nxt][l], e->beam.parent, r0, rows, d, st));
        ND_CUDA(e, cudaMemcpyAsync(R(e->rh[cur][l], d), R(e->rh[nxt][l], d), (size_t)rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
        ND_CUDA(e, cudaMemcpyAsync(R(e->rc[cur][l], d), R(e->rc[nxt][l], d), (size_t)rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
      }
      ND_LAUNCH(e, gather_rows(e->feed[cur], e->feed[nxt], e->beam.parent, r0, rows, d, st));
      ND_CUDA(e, cudaMemcpyAsync(R(e->feed[cur], d), R(e->feed[nxt], d), (size_t)rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    const float* below = nullptr;
    const bool gru = c.rnn_type == ND_RNN_GRU;
    const int64_t gw = (gru ? 3 : 4) * (int64_t)d;       // gate row width (buffers are sized for 4d)
    for (int l = 0; l < c.dec_layers; ++l) {
      const nd_engine::RnnCell& Rc = e->cells[l];
      GemmOpt oa;
      if (l == 0 && c.input_feed) {
        ND_TRY(run_gemm(e, Rc.ih_e, R(e->x, d), d, R(e->ga, 4 * d), gw, rows, oa, st));
        GemmOpt of; of.residual = R(e->ga, 4 * d); of.ldr = gw;
        ND_TRY(run_gemm(e, Rc.ih_f, R(e->feed[cur], d), d, R(e->ga, 4 * d), gw, rows, of, st));
      } else {
        ND_TRY(run_gemm(e, Rc.ih, l == 0 ? R(e->x, d) : below, d, R(e->ga, 4 * d), gw, rows, oa, st));
      }
      GemmOpt ob;
      ND_TRY(run_gemm(e, Rc.hh, R(e->rh[cur][l], d), d, R(e->gb, 4 * d), gw, rows, ob, st));
      if (gru)
        ND_LAUNCH(e, gru_cell_pointwise(R(e->ga, 4 * d), R(e->gb, 4 * d), R(e->rh[cur][l], d), R(e->rh[nxt][l], d), rows, d, st));
      else
        ND_LAUNCH(e, lstm_cell_pointwise(R(e->ga, 4 * d), R(e->gb, 4 * d), R(e->rc[cur][l], d), R(e->rh[nxt][l], d),
                                         R(e->rc[nxt][l], d), rows, d, st));
      below = R(e->rh[nxt][l], d);
    }
    MlpAttnParams ma;
    ma.mem = e->mb + (int64_t)dc.c0 * Tp * d; ma.lengths = e->mem_len + dc.c0;
    ma.retired = retired ? retired + dc.c0 : nullptr; ma.ctx = R(e->actx, d); ma.ctx_ld = d;
    ma.n_chunks = dc.nc; ma.NQ = K; ma.T = Tp; ma.d = d;
    ma.attn = dc.attn_out ? dc.attn_out + (int64_t)r0 * Tp : nullptr;
    if (c.attn_type == ND_ATTN_MLP) {
      GemmOpt oq;                    // wq = W_q h + b; uh = U_k H was cached by decoder_init   global_attention.py:123-136
      ND_TRY(run_gemm(e, e->attn_q, below, d, R(e->wq, d), d, rows, oq, st));
      ma.wq = R(e->wq, d); ma.uh = e->uh + (int64_t)dc.c0 * Tp * d; ma.v = e->attn_v;
    } else {
      // "general": score = (W_in h) . H_s; "dot": score = h . H_s                             global_attention.py:113-122
      ma.dot = 1;
      if (c.attn_type == ND_ATTN_GENERAL) {
        GemmOpt oq;
        ND_TRY(run_gemm(e, e->attn_in, below, d, R(e->wq, d), d, rows, oq, st));
        ma.wq = R(e->wq, d);
      } else {
        ma.wq = below;
      }
      ma.uh = e->mb + (int64_t)dc.c0 * Tp * d;
    }
    if (e->kv_packed) {
      const KvPlanes pl = pk_planes(e);
      const int64_t r0k = (int64_t)dc.c0 * Tp;
      ma.kv_fmt = e->pk_fmt; ma.kv_hi = pl.hi + r0k * 2 * d; ma.kv_lo = pl.lo + r0k * 2 * d; ma.kv_scale = pl.scale + r0k * 2;
    }
    ND_LAUNCH_CAT(e, ND_PROF_MLP_ATTN, st, mlp_attention(ma, st));
    // attn_h = W_out [c ; h] (+ b, no tanh, for mlp; tanh, no bias, for general / dot)   global_attention.py:197-203
    GemmOpt oc;
    ND_TRY(run_gemm(e, e->attn_out_c, R(e->actx, d), d, R(e->wq, d), d, rows, oc, st));
    GemmOpt oh; oh.residual = R(e->wq, d); oh.ldr = d;
    ND_TRY(run_gemm(e, e->attn_out_h, below, d, R(e->feed[nxt], d), d, rows, oh, st));
    if (c.attn_type != ND_ATTN_MLP) ND_LAUNCH(e, tanh_inplace(R(e->feed[nxt], d), (int64_t)rows * d, st));
    gp.x = R(e->feed[nxt], d); gp.x_ld = d; gp.ln_g = nullptr; gp.ln_b = nullptr;
  } else {
    // CNN decoder, incremental (the reference re-runs the whole prefix each step, cnn_decoder.py:79-80; the
    // convolutions are causal, so position t only needs the last k inputs of every layer)
    const int k = c.cnn_kernel_width;
    GemmOpt ol;
    ND_TRY(run_gemm(e, e->dec_lin, R(e->x, d), d, R(e->cbase, d), d, rows, ol, st));       // :97-100
    const float* xl = R(e->cbase, d);
    float* xbuf[2] = {R(e->x1, d), R(e->x, d)};
    for (int l = 0; l < c.dec_layers; ++l) {
      ND_LAUNCH(e, cnn_window(xl - (int64_t)r0 * d, e->chist[0][l], anc, dc.Lmax, retired, K, e->cA, r0, rows,
                              dc.step, k, d, dc.Lmax, st));
      GemmOpt ocv;
      ND_TRY(run_gemm(e, e->dec_conv[l].conv, R(e->cA, (int64_t)k * d), (int64_t)k * d, R(e->cy, 2 * d), 2 * d, rows, ocv, st));
      ND_LAUNCH(e, glu_residual(R(e->cy, 2 * d), nullptr, nullptr, R(e->cout, d), rows, d, st));       // out = a*sigmoid(g)
      GemmOpt oi;
      ND_TRY(run_gemm(e, e->dec_attn_in[l], R(e->cout, d), d, R(e->cpre, d), d, rows, oi, st));
      ND_LAUNCH(e, add_scale(R(e->cbase, d), R(e->cpre, d), s2, R(e->ctgt, d), (int64_t)rows * d, st));
      MlpAttnParams ma;
      ma.dot = 1; ma.wq = R(e->ctgt, d); ma.uh = e->mb + (int64_t)dc.c0 * Tp * d;        // keys: encoder top
      ma.mem = e->enc_comb + (int64_t)dc.c0 * Tp * d;                                     // values: combined
      ma.lengths = nullptr; ma.retired = retired ? retired + dc.c0 : nullptr; ma.ctx = R(e->cctx2, d); ma.ctx_ld = d;
      ma.n_chunks = dc.nc; ma.NQ = K; ma.T = Tp; ma.d = d;
      ma.attn = (l + 1 == c.dec_layers && dc.attn_out) ? dc.attn_out + (int64_t)r0 * Tp : nullptr;
      if (e->kv_packed) {
        const KvPlanes pl = pk_planes(e);
        const int64_t r0k = (int64_t)dc.c0 * Tp;
        ma.kv_fmt = e->pk_fmt; ma.kv_hi = pl.hi + r0k * 2 * d; ma.kv_lo = pl.lo + r0k * 2 * d; ma.kv_scale = pl.scale + r0k * 2;
      }
      ND_LAUNCH_CAT(e, ND_PROF_MLP_ATTN, st, mlp_attention(ma, st));
      float* xn = xbuf[l & 1];
      ND_LAUNCH(e, cnn_combine(xl, R(e->cctx2, d), R(e->cout, d), s2, xn, (int64_t)rows * d, st));      // :116
      xl = xn;
    }
    gp.x = xl; gp.x_ld = d; gp.ln_g = nullptr; gp.ln_b = nullptr;
  }
  gp.Wg = e->gen.W; gp.bg = e->gen.b; gp.logp = R(e->logp, c.vocab_size); gp.rows = rows; gp.d = d;
  gp.V = c.vocab_size; gp.step = dc.step;
  ND_LAUNCH_CAT(e, ND_PROF_GENERATOR, st, generator_step(gp, st));
  return ND_OK;
}

// fork the caller's stream into the engine's decode streams / join them back
int fork_streams(nd_engine* e, cudaStream_t st, int n) {
  ND_CUDA(e, cudaEventRecord(e->fork_ev, st));
  for (int g = 0; g < n; ++g) ND_CUDA(e, cudaStreamWaitEvent(e->streams[g], e->fork_ev, 0));
  return ND_OK;
}
int join_streams(nd_engine* e, cudaStream_t st, int n) {
  for (int g = 0; g < n; ++g) {
    ND_CUDA(e, cudaEventRecord(e->join_ev[g], e->streams[g]));
    ND_CUDA(e, cudaStreamWaitEvent(st, e->join_ev[g], 0));
  }
  return ND_OK;
}
int n_groups(const nd_engine* e, int B) {
  int n = e->decode_streams;
  if (e->prof_mask) n = 1;                       // per-kernel event timing wants kernels that run alone
  if (n > (int)e->streams.size()) n = (int)e->streams.size();
  while (n > 1 && B / n < 64) --n;               // tiny groups only add launch overhead
  return n < 1 ? 1 : n;
}

int greedy_body(nd_engine* e, int max_len, int min_len, int64_t* out_ids, float* out_scores, float* out_attn,
                       float* out_logits, cudaStream_t st) {
  const int B = e->B, V = e->cfg.vocab_size;
  const bool cross_mb = use_cross_mb(e, 1, out_attn != nullptr);
  ND_TRY(decoder_init(e, 1, st, cross_mb));
  ND_LAUNCH(e, fill_int(e->cur_tok, B, 2, st));          // <s> for every row (translator.py:451-452)
  NvtxRange nvtx("nd:greedy decode loop");
  const int G = n_groups(e, B);
  if (G > 1) ND_TRY(fork_streams(e, st, G));
  for (int step = 0; step < max_len; ++step) {           // no EOS early exit, like the reference (:455)
    for (int g = 0; g < G; ++g) {
      DecodeCtx dc;
      dc.K = 1; dc.Lmax = e->cfg.max_tgt_len; dc.beam = false; dc.step = step; dc.cross_mb = cross_mb;
      dc.c0 = (int)((int64_t)B * g / G);
      dc.nc = (int)((int64_t)B * (g + 1) / G) - dc.c0;
      dc.attn_out = out_attn ? out_attn + (int64_t)step * B * e->Tp : nullptr;
      GenParams gp;
      gp.ids = out_ids + (int64_t)dc.c0 * max_len; gp.ids_ld = max_len; gp.scores = out_scores + dc.c0;
      gp.next_tok = e->cur_tok + dc.c0;
      gp.trace = out_logits ? out_logits + ((int64_t)step * B + dc.c0) * V : nullptr;
      gp.min_len = min_len;
      ND_TRY(decoder_step(e, dc, gp, G > 1 ? e->streams[g] : st));
    }
  }
  if (G > 1) ND_TRY(join_streams(e, st, G));
  return ND_OK;
}

// mode 0: --fast batched beam (translator.py:619-825); mode 1: object beam (translator.py:827-926 + beam.py)
int beam_body(nd_engine* e, int K, int n_best, int max_len, int min_len, float alpha, int mode, int lp_mode,
              int64_t* out_ids, int32_t* out_lens, float* out_scores, cudaStream_t st) {
  const int B = e->B;
  ND_TRY(decoder_init(e, K, st));
  BeamParams bp;
  bp.logp = e->logp; bp.st = e->beam; bp.B = B; bp.K = K; bp.V = e->cfg.vocab_size; bp.Lmax = e->cfg.max_tgt_len;
  bp.max_len = max_len; bp.min_len = min_len; bp.n_best = n_best; bp.alpha = alpha; bp.mode = mode; bp.lp_mode = lp_mode;
  ND_LAUNCH(e, beam_init(bp, 2, st));
  e->attn_steps = 0;
  if (e->beam_attn) {
    // translator.py:806-812 (alive_attn) / beam.py:135 (Beam.attn): the head-0 cross attention of every step is kept
    // per ROW; a hypothesis' rows are found afterwards through its ancestor table (nd_beam_attention)
    if (!e->attn_hist) {
      e->attn_hist = dalloc<float>(e, (size_t)e->cfg.max_tgt_len * e->max_rows * e->cfg.max_src_len);
      if (!e->attn_hist) return fail(e, ND_ERR_NOMEM, "beam_attention: no memory for the attention history");
    }
    e->attn_rows = B * K; e->attn_Tp = e->Tp; e->attn_steps = max_len;
  }
  const bool cov_on = mode == 1 && e->cov_mode != 0;
  if (mode == 1) { bp.block_ngram = e->block_ngram; bp.excl_mask = e->excl_mask; }
  if (cov_on) {
    const size_t n = (size_t)e->max_rows * e->cfg.max_src_len;
    if (!e->cov) {
      e->cov = dalloc<float>(e, 2 * n);
      e->cov_pen = dalloc<float>(e, (size_t)e->max_rows);
      e->cov_attn = dalloc<float>(e, n);
      if (!e->cov || !e->cov_pen || !e->cov_attn) return fail(e, ND_ERR_NOMEM, "coverage penalty: no memory for the coverage vectors");
    }
    bp.cov_mode = e->cov_mode; bp.beta = e->beta; bp.cov = e->cov; bp.cov_pen = e->cov_pen; bp.stepwise = e->stepwise;
    bp.mem_len = e->mem_len; bp.Tp = e->Tp;
  }
  NvtxRange nvtx(mode == 1 ? "nd:object beam decode loop" : "nd:fast beam decode loop");
  // object mode: stop_step / n_done are shared by all chunks (every Beam advances until ALL are done), so chunk
  // groups on independent streams would see the stop at different steps: one stream
  const int G = mode == 1 ? 1 : n_groups(e, B);
  if (G > 1) ND_TRY(fork_streams(e, st, G));
  // once every chunk has retired (fast) / every Beam is done (object) the remaining steps only cost their launches
  struct AliveScope { nd_engine* e; ~AliveScope() { e->gemm_alive = nullptr; } } alive_scope{e};
  e->gemm_alive = e->beam.n_alive;
  for (int step = 0; step < max_len; ++step) {
    for (int g = 0; g < G; ++g) {
      cudaStream_t gs = G > 1 ? e->streams[g] : st;
      DecodeCtx dc;
      dc.K = K; dc.Lmax = e->cfg.max_tgt_len; dc.beam = true; dc.step = step;
      if (e->beam_attn) dc.attn_out = e->attn_hist + (size_t)step * B * K * e->Tp;
      else if (cov_on) dc.attn_out = e->cov_attn;
      bp.attn_step = dc.attn_out;
      dc.c0 = (int)((int64_t)B * g / G);
      dc.nc = (int)((int64_t)B * (g + 1) / G) - dc.c0;
      GenParams gp;
      // EOS is suppressed while step < min_length (fast, translator.py:714-715) / while len(next_ys) = step + 1
      // < min_length (object, beam.py:88-92)
      gp.min_len = mode == 1 ? (min_len > 0 ? min_len - 1 : 0) : min_len;
      ND_TRY(decoder_step(e, dc, gp, gs));
      bp.step = step; bp.b0 = dc.c0; bp.nb = dc.nc;
      ND_LAUNCH_CAT(e, ND_PROF_BEAM, gs, beam_step(bp, gs));
    }
  }
  if (G > 1) ND_TRY(join_streams(e, st, G));
  ND_LAUNCH(e, beam_finalize(bp, out_ids, out_lens, out_scores, st));
  return ND_OK;
}

// Run `body(stream)` either eagerly on the caller's stream or as a cached CUDA graph on the engine's
// main stream (bracketed by events so it is ordered inside the caller's stream).
template <class Body>
int run_cached(nd_engine* e, const std::vector<int64_t>& key, cudaStream_t st, Body body) {
  nd_engine::GraphEntry* hit = nullptr;
  for (auto& g : e->graphs)
    if (g.key == key) hit = &g;
  if (!hit) {
    if (e->last_key != key) {            // first sighting: run eagerly (also warms up lazy kernel attributes)
      e->last_key = key;
      return body(st);
    }
    // second call with the same configuration: capture
    cudaGraph_t graph = nullptr;
    const int64_t before = e->launches;
    ND_CUDA(e, cudaStreamBeginCapture(e->main_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body(e->main_stream);
    cudaError_t err = cudaStreamEndCapture(e->main_stream, &graph);
    if (rc != ND_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (err != cudaSuccess) { e->sticky = true; return fail(e, ND_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(err)); }
    nd_engine::GraphEntry ge;
    ge.key = key;
    ge.launches = e->launches - before;
    e->launches = before;
    err = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) { e->sticky = true; return fail(e, ND_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(err)); }
    if (e->graphs.size() >= 8) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
    e->graphs.push_back(ge);
    hit = &e->graphs.back();
  }
  ND_CUDA(e, cudaEventRecord(e->g_in, st));
  ND_CUDA(e, cudaStreamWaitEvent(e->main_stream, e->g_in, 0));
  ND_CUDA(e, cudaGraphLaunch(hit->exec, e->main_stream));
  ND_CUDA(e, cudaEventRecord(e->g_out, e->main_stream));
  ND_CUDA(e, cudaStreamWaitEvent(st, e->g_out, 0));
  e->launches += hit->launches;
  return ND_OK;
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

int nd_api_version(void) { return ND_API_VERSION; }

const char* nd_last_error(const nd_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int nd_create(const nd_config* cfg, nd_engine** out) {
  if (!cfg || !out) return fail(nullptr, ND_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->api_version != ND_API_VERSION) return fail(nullptr, ND_ERR_INVALID, "nd_config.api_version mismatch");
  if (cfg->d_model % 32 || cfg->heads <= 0 || 32 % cfg->heads || cfg->d_model % cfg->heads)
    return fail(nullptr, ND_ERR_INVALID, "d_model must be a multiple of 32 and heads a power of two <= 32");
  if (cfg->vocab_size > 16 || cfg->vocab_size < 5) return fail(nullptr, ND_ERR_INVALID, "vocab_size must be in [5,16]");
  if (cfg->max_beam < 1 || cfg->max_beam > 8) return fail(nullptr, ND_ERR_INVALID, "max_beam must be in [1,8]");
  if (cfg->rnn_type != ND_RNN_LSTM && cfg->rnn_type != ND_RNN_GRU) return fail(nullptr, ND_ERR_INVALID, "rnn_type must be ND_RNN_LSTM or ND_RNN_GRU");
  if (cfg->max_batch < 1 || cfg->max_src_len < 1 || cfg->max_tgt_len < 1) return fail(nullptr, ND_ERR_INVALID, "bad max sizes");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, ND_ERR_CUDA, "no CUDA device available (libnanodec has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, ND_ERR_INVALID, "bad device ordinal");
  if ((err = cudaSetDevice(cfg->device)) != cudaSuccess) return fail(nullptr, ND_ERR_CUDA, cudaGetErrorString(err));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10) return fail(nullptr, ND_ERR_CUDA, "libnanodec is built for sm_100a only; device is sm_" +
                                                             std::to_string(prop.major) + std::to_string(prop.minor));
  std::unique_ptr<nd_engine> e(new nd_engine());
  e->cfg = *cfg;
  e->n_sm = prop.multiProcessorCount;
  if (cfg->gemm_mode != ND_GEMM_SIMT_FP32) {
    const char* why = "";
    if (!gemm_tc_available(&why)) return fail(nullptr, ND_ERR_CUDA, std::string("tcgen05 GEMM path unavailable: ") + why);
  }
  int rc = alloc_workspace(e.get());
  if (rc != ND_OK) {
    g_create_error = "workspace allocation failed: " + e->err;
    for (void* p : e->allocs) if (p) cudaFree(p);
    return rc;
  }
  if (cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->g_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->g_out, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->main_stream, cudaStreamNonBlocking) != cudaSuccess)
    return fail(nullptr, ND_ERR_CUDA, "cudaEventCreate / cudaStreamCreate failed");
  nd_engine* raw = e.release();
  rc = nd_set_int(raw, "decode_streams", 4);
  if (rc == ND_OK) raw->decode_streams = 1;   // streams exist, default is the single-stream schedule
  if (rc != ND_OK) { g_create_error = raw->err; nd_destroy(raw); return rc; }
  *out = raw;
  return ND_OK;
}

int nd_destroy(nd_engine* e) {
  if (!e) return ND_OK;
  cudaSetDevice(e->cfg.device);
  for (cudaEvent_t ev : e->prof_pool) cudaEventDestroy(ev);
  for (cudaStream_t s2 : e->streams) cudaStreamDestroy(s2);
  for (cudaEvent_t ev : e->join_ev) cudaEventDestroy(ev);
  if (e->fork_ev) cudaEventDestroy(e->fork_ev);
  for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->g_in) cudaEventDestroy(e->g_in);
  if (e->g_out) cudaEventDestroy(e->g_out);
  if (e->main_stream) cudaStreamDestroy(e->main_stream);
  for (void* p : e->allocs) if (p) cudaFree(p);
  delete e;
  return ND_OK;
}

int nd_load_weight(nd_engine* e, const char* name, const void* data, const int64_t* shape, int32_t ndim, int32_t dtype) {
  if (!e || !name || !data || (ndim > 0 && !shape)) return fail(e, ND_ERR_INVALID, "null argument");
  if (e->finalized) return fail(e, ND_ERR_STATE, "weights already finalized");
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  if (dtype == ND_DTYPE_F32) {
    t.f.resize((size_t)n);
    cudaSetDevice(e->cfg.device);
    ND_CUDA(e, cudaMemcpy(t.f.data(), data, (size_t)n * sizeof(float), cudaMemcpyDefault));
  } else {
    return ND_OK;          // integer buffers (num_batches_tracked, mask) carry no arithmetic
  }
  e->raw[name] = std::move(t);
  return ND_OK;
}

int nd_finalize_weights(nd_engine* e) {
  if (!e) return ND_ERR_INVALID;
  if (e->finalized) return ND_OK;
  cudaSetDevice(e->cfg.device);
  int rc = finalize(e);
  if (rc == ND_OK) e->finalized = true;
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { e->sticky = true; return fail(e, ND_ERR_CUDA, cudaGetErrorString(err)); }
  return rc;
}

int nd_profile_enable(nd_engine* e, uint32_t category_mask) {
  if (!e) return ND_ERR_INVALID;
  e->prof_mask = category_mask;
  e->prof_used = 0;
  e->prof_cat.clear();
  return ND_OK;
}

int nd_profile_read(nd_engine* e, double* out_ms, int64_t* out_count) {
  if (!e || !out_ms || !out_count) return ND_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  for (int c = 0; c < ND_PROF_NCAT; ++c) { out_ms[c] = 0.0; out_count[c] = 0; }
  if (e->prof_used) {
    cudaError_t err = cudaEventSynchronize(e->prof_pool[e->prof_used - 1]);
    if (err != cudaSuccess) { e->sticky = true; return fail(e, ND_ERR_CUDA, cudaGetErrorString(err)); }
  }
  for (size_t i = 0; i < e->prof_cat.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->prof_pool[2 * i], e->prof_pool[2 * i + 1]) == cudaSuccess) {
      out_ms[e->prof_cat[i]] += ms;
      out_count[e->prof_cat[i]] += 1;
    }
  }
  e->prof_used = 0;
  e->prof_cat.clear();
  return ND_OK;
}

int64_t nd_launch_count(const nd_engine* e) { return e ? e->launches : 0; }
int nd_reset_launch_count(nd_engine* e) { if (e) e->launches = 0; return ND_OK; }

static int check_ready(nd_engine* e) {
  if (!e) return ND_ERR_INVALID;
  if (e->sticky) return ND_ERR_CUDA;
  if (!e->finalized) return fail(e, ND_ERR_STATE, "nd_finalize_weights has not been called");
  cudaSetDevice(e->cfg.device);
  return ND_OK;
}

int nd_frontend_stats(nd_engine* e, const int16_t* signal, const int64_t* read_offsets, int32_t n_reads,
                      int32_t normalization, double* out_center, double* out_scale, void* stream) {
  if (!e) return ND_ERR_INVALID;
  if (e->sticky) return ND_ERR_CUDA;
  cudaSetDevice(e->cfg.device);
  if (normalization < 0 || normalization > 2) return fail(e, ND_ERR_INVALID, "unknown normalization");
  if (n_reads > e->fe_todo_cap) {                     // scratch of the histogram path: grows with the largest call seen
    int* p = nullptr;
    ND_CUDA(e, cudaMalloc(&p, (size_t)n_reads * sizeof(int)));
    e->allocs.push_back(p);                           // (the old block is freed with the engine)
    e->fe_todo = p;
    e->fe_todo_cap = n_reads;
  }
  ND_LAUNCH(e, frontend_stats(signal, read_offsets, n_reads, normalization, out_center, out_scale, e->fe_todo,
                              (cudaStream_t)stream));
  return ND_OK;
}

int nd_frontend_chunks(nd_engine* e, const int16_t* signal, const int64_t* read_offsets, const double* center,
                       const double* scale, const int32_t* chunk_read, const int64_t* chunk_start, int32_t n_chunks,
                       int32_t chunk_len, float* out_chunks, int64_t* out_lengths, void* stream) {
  if (!e) return ND_ERR_INVALID;
  if (e->sticky) return ND_ERR_CUDA;
  cudaSetDevice(e->cfg.device);
  ND_LAUNCH(e, frontend_chunks(signal, read_offsets, center, scale, chunk_read, chunk_start, n_chunks, chunk_len,
                               out_chunks, out_lengths, (cudaStream_t)stream));
  return ND_OK;
}

int nd_frontend_stats_f64(nd_engine* e, const double* signal, const int64_t* read_offsets, int32_t n_reads,
                          int32_t normalization, double* out_center, double* out_scale, void* stream) {
  if (!e) return ND_ERR_INVALID;
  if (e->sticky) return ND_ERR_CUDA;
  cudaSetDevice(e->cfg.device);
  if (normalization < 0 || normalization > 2) return fail(e, ND_ERR_INVALID, "unknown normalization");
  ND_LAUNCH(e, frontend_stats_f64(signal, read_offsets, n_reads, normalization, out_center, out_scale, (cudaStream_t)stream));
  return ND_OK;
}

int nd_frontend_chunks_f64(nd_engine* e, const double* signal, const int64_t* read_offsets, const double* center,
                           const double* scale, const int32_t* chunk_read, const int64_t* chunk_start, int32_t n_chunks,
                           int32_t chunk_len, float* out_chunks, int64_t* out_lengths, void* stream) {
  if (!e) return ND_ERR_INVALID;
  if (e->sticky) return ND_ERR_CUDA;
  cudaSetDevice(e->cfg.device);
  ND_LAUNCH(e, frontend_chunks_f64(signal, read_offsets, center, scale, chunk_read, chunk_start, n_chunks, chunk_len,
                                   out_chunks, out_lengths, (cudaStream_t)stream));
  return ND_OK;
}

int nd_encode(nd_engine* e, const float* src, const int64_t* lengths, int32_t B, int32_t T, void* stream) {
  ND_TRY(check_ready(e));
  cudaStream_t st = (cudaStream_t)stream;
  if (B < 1 || B > e->cfg.max_batch || T < 1 || T > e->cfg.max_src_len)
    return fail(e, ND_ERR_INVALID, "nd_encode: B or T outside the sizes given to nd_create");
  e->encoded = false;
  e->B = B; e->T = T;
  ND_CUDA(e, cudaMemcpyAsync(e->src, src, (size_t)B * T * sizeof(float), cudaMemcpyDefault, st));
  ND_CUDA(e, cudaMemcpyAsync(e->lengths, lengths, (size_t)B * sizeof(int64_t), cudaMemcpyDefault, st));
  const nd_config& c = e->cfg;
  int rc;
  NvtxRange nvtx("nd:encode");
  if (c.encoder_type == ND_ENC_TRANSFORMER || c.encoder_type == ND_ENC_CTRANSFORMER) rc = encode_transformer(e, st);
  else if (c.encoder_type == ND_ENC_RESNET) rc = encode_resnet(e, st);
  else if (c.encoder_type == ND_ENC_CNN) rc = encode_cnn(e, st);
  else rc = encode_lstm_stack(e, st);
  if (rc == ND_OK) e->encoded = true;
  return rc;
}

int nd_get_memory_bank(nd_engine* e, float* out, int64_t* out_lengths, int32_t* out_Tp, void* stream) {
  ND_TRY(check_ready(e));
  if (!e->encoded) return fail(e, ND_ERR_STATE, "nd_get_memory_bank before nd_encode");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_Tp) *out_Tp = e->Tp;
  if (out) {
    if (e->cfg.encoder_type == ND_ENC_CNN) ND_LAUNCH(e, transpose_to_dbt(e->mb, out, e->B, e->Tp, e->cfg.d_model, st));
    else ND_LAUNCH(e, transpose_bt(e->mb, out, e->B, e->Tp, e->cfg.d_model, st));
  }
  if (out_lengths)
    ND_CUDA(e, cudaMemcpyAsync(out_lengths, e->mem_len, (size_t)e->B * sizeof(int64_t), cudaMemcpyDefault, st));
  return ND_OK;
}

int nd_decode_greedy(nd_engine* e, int32_t max_len, int32_t min_len, int64_t* out_ids, float* out_scores,
                     float* out_attn, float* out_logits, void* stream) {
  ND_TRY(check_ready(e));
  if (!e->encoded) return fail(e, ND_ERR_STATE, "nd_decode_greedy before nd_encode");
  if (max_len < 1 || max_len > e->cfg.max_tgt_len) return fail(e, ND_ERR_INVALID, "max_len outside nd_create sizes");
  if (!out_ids || !out_scores) return fail(e, ND_ERR_INVALID, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_attn || out_logits || e->prof_mask || !e->use_graphs)
    return greedy_body(e, max_len, min_len, out_ids, out_scores, out_attn, out_logits, st);
  const int B = e->B;
  const std::vector<int64_t> key = {0, B, e->T, e->Tp, max_len, min_len, 1, 1, 0, n_groups(e, B), g_pdl, e->cross_mode, e->kv_mode};
  ND_TRY(run_cached(e, key, st, [&](cudaStream_t s2) {
    return greedy_body(e, max_len, min_len, e->o_ids, e->o_scores, nullptr, nullptr, s2);
  }));
  ND_CUDA(e, cudaMemcpyAsync(out_ids, e->o_ids, (size_t)B * max_len * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  ND_CUDA(e, cudaMemcpyAsync(out_scores, e->o_scores, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return ND_OK;
}

static int decode_beam_any(nd_engine* e, int32_t beam_size, int32_t n_best, int32_t max_len, int32_t min_len,
                           float alpha, int mode, int lp_mode, int64_t* out_ids, int32_t* out_lens, float* out_scores,
                           void* stream) {
  ND_TRY(check_ready(e));
  if (!e->encoded) return fail(e, ND_ERR_STATE, "nd_decode_beam before nd_encode");
  if (beam_size < 1 || beam_size > e->cfg.max_beam) return fail(e, ND_ERR_INVALID, "beam_size outside nd_create sizes");
  if (n_best < 1 || n_best > beam_size) return fail(e, ND_ERR_INVALID, "n_best must be in [1, beam_size]");
  if (max_len < 1 || max_len > e->cfg.max_tgt_len) return fail(e, ND_ERR_INVALID, "max_len outside nd_create sizes");
  if (lp_mode < 0 || lp_mode > 2) return fail(e, ND_ERR_INVALID, "length penalty must be 0 (none), 1 (wu) or 2 (avg)");
  if (mode == 1 && beam_size > e->cfg.vocab_size)
    return fail(e, ND_ERR_INVALID, "object beam: beam_size > vocabulary (the reference's step-0 topk fails too)");
  if (!out_ids || !out_lens || !out_scores) return fail(e, ND_ERR_INVALID, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = e->B, K = beam_size;
  e->beam_n_best = n_best; e->beam_K = beam_size; e->beam_mode = mode;
  if (e->prof_mask || !e->use_graphs || e->beam_attn || (mode == 1 && (e->block_ngram || e->cov_mode)))
    return beam_body(e, K, n_best, max_len, min_len, alpha, mode, lp_mode, out_ids, out_lens, out_scores, st);
  int32_t alpha_bits;
  memcpy(&alpha_bits, &alpha, sizeof(alpha_bits));
  const std::vector<int64_t> key = {1 + mode * 4 + lp_mode, B, e->T, e->Tp, max_len, min_len, K, n_best, alpha_bits,
                                    n_groups(e, B), g_pdl, e->kv_mode, e->kv_beam_packed};
  ND_TRY(run_cached(e, key, st, [&](cudaStream_t s2) {
    return beam_body(e, K, n_best, max_len, min_len, alpha, mode, lp_mode, e->o_ids, e->o_lens, e->o_scores, s2);
  }));
  ND_CUDA(e, cudaMemcpyAsync(out_ids, e->o_ids, (size_t)B * n_best * max_len * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  ND_CUDA(e, cudaMemcpyAsync(out_lens, e->o_lens, (size_t)B * n_best * sizeof(int), cudaMemcpyDeviceToDevice, st));
  ND_CUDA(e, cudaMemcpyAsync(out_scores, e->o_scores, (size_t)B * n_best * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return ND_OK;
}

int nd_beam_attention(nd_engine* e, int32_t n_best, int32_t max_len, float* out, int32_t* out_widths, void* stream) {
  ND_TRY(check_ready(e));
  if (!e->beam_attn || !e->attn_hist || e->attn_steps == 0)
    return fail(e, ND_ERR_STATE, "nd_beam_attention: set the option beam_attention before the beam decode");
  if (n_best != e->beam_n_best || max_len != e->attn_steps || !out)
    return fail(e, ND_ERR_INVALID, "nd_beam_attention: n_best / max_len differ from the last beam decode");
  ND_LAUNCH(e, beam_gather_attention(e->beam, e->attn_hist, e->mem_len, e->B, e->beam_K, n_best, e->cfg.max_tgt_len,
                                     max_len, e->attn_rows, e->attn_Tp, e->beam_mode, out, out_widths,
                                     (cudaStream_t)stream));
  return ND_OK;
}

int nd_decode_beam(nd_engine* e, int32_t beam_size, int32_t n_best, int32_t max_len, int32_t min_len, float alpha,
                   int64_t* out_ids, int32_t* out_lens, float* out_scores, void* stream) {
  return decode_beam_any(e, beam_size, n_best, max_len, min_len, alpha, 0, 0, out_ids, out_lens, out_scores, stream);
}

int nd_decode_beam_object(nd_engine* e, int32_t beam_size, int32_t n_best, int32_t max_len, int32_t min_len,
                          int32_t length_penalty, float alpha, int64_t* out_ids, int32_t* out_lens,
                          float* out_scores, void* stream) {
  return decode_beam_any(e, beam_size, n_best, max_len, min_len, alpha, 1, length_penalty, out_ids, out_lens, out_scores,
                         stream);
}

int nd_set_int(nd_engine* e, const char* name, int64_t value) {
  if (!e || !name) return ND_ERR_INVALID;
  // captured decode loops bake in the kernel choices of the moment (several switches are process-wide and not part
  // of the graph key): any option change drops them, the next two calls run eagerly and re-capture
  for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
  e->last_key.clear();
  if (strcmp(name, "decode_streams") == 0) {
    if (value < 1 || value > 16) return fail(e, ND_ERR_INVALID, "decode_streams must be in [1,16]");
    cudaSetDevice(e->cfg.device);
    while ((int64_t)e->streams.size() < value) {
      cudaStream_t s2;
      cudaEvent_t ev;
      ND_CUDA(e, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      ND_CUDA(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->streams.push_back(s2);
      e->join_ev.push_back(ev);
    }
    e->decode_streams = (int)value;
    return ND_OK;
  }
  if (strcmp(name, "gemm_persistent") == 0) {   // process-wide: large-M GEMMs as a persistent kernel (default 1)
    gemm_tc_set_persistent(value < 0 ? 0 : (value > 2 ? 2 : (int)value));   // 2: A operand in tensor memory
    return ND_OK;
  }
  if (strcmp(name, "cross_mode") == 0) {     // 1: memory-bank-space cross attention for greedy decode, 0: K/V
    e->cross_mode = value != 0;
    return ND_OK;
  }
  if (strcmp(name, "kv_mode") == 0) {        // storage of the memory keys / values: 0 fp32, 1 q24, 2 q16 (kernels.cuh)
    if (value < 0 || value > KV_FP24) return fail(e, ND_ERR_INVALID, "kv_mode must be in [0, 5] (nanodec.h)");
    e->kv_mode = (int)value;
    return ND_OK;
  }
  if (strcmp(name, "kv_beam_packed") == 0) {
    e->kv_beam_packed = value != 0;
    return ND_OK;
  }
  if (strcmp(name, "beam_attention") == 0) {
    e->beam_attn = value != 0;
    return ND_OK;
  }
  if (strcmp(name, "block_ngram_repeat") == 0) {
    if (value < 0 || value > 16) return fail(e, ND_ERR_INVALID, "block_ngram_repeat must be in [0, 16]");
    e->block_ngram = (int)value;
    return ND_OK;
  }
  if (strcmp(name, "block_ngram_exclude") == 0) {
    e->excl_mask = (unsigned)value;
    return ND_OK;
  }
  if (strcmp(name, "stepwise_penalty") == 0) {
    e->stepwise = value != 0;
    return ND_OK;
  }
  if (strcmp(name, "coverage_penalty") == 0) {
    if (value < 0 || value > 2) return fail(e, ND_ERR_INVALID, "coverage_penalty must be 0 (none), 1 (wu) or 2 (summary)");
    e->cov_mode = (int)value;
    return ND_OK;
  }
  if (strcmp(name, "cross_packed_fast") == 0) {   // process-wide
    cross_attention_packed_set_fast((int)value);
    return ND_OK;
  }
  if (strcmp(name, "gemm_a_tmem") == 0) {         // process-wide
    gemm_tc_set_a_tmem(value != 0);
    return ND_OK;
  }
  if (strcmp(name, "frontend_fast") == 0) {       // process-wide
    frontend_set_fast(value != 0);
    return ND_OK;
  }
  if (strcmp(name, "gemm_wide_wave") == 0) {      // process-wide
    gemm_tc_set_wide_wave(value != 0);
    return ND_OK;
  }
  if (strcmp(name, "gemm_serial_split") == 0) {   // process-wide
    gemm_tc_set_serial_split(value != 0);
    return ND_OK;
  }
  if (strcmp(name, "cross_ring_groups") == 0) {   // process-wide: 1 or 2 groups of 8 consumer warps
    cross_attention_ring_set_groups(value == 1 ? 1 : 2);
    return ND_OK;
  }
  if (strcmp(name, "cross_beam_kernel") == 0) {   // process-wide
    cross_attention_set_beam_kernel(value);
    return ND_OK;
  }
  if (strcmp(name, "enc_attn_tc") == 0) {
    e->enc_attn_tc = value != 0;
    return ND_OK;
  }
  if (strcmp(name, "cross_mb_version") == 0) {
    cross_attention_mb_set_version((int)value);
    return ND_OK;
  }
  if (strcmp(name, "pdl") == 0) {            // programmatic dependent launch in the decode loop (process-wide)
    g_pdl = (int)value;
    return ND_OK;
  }
  if (strcmp(name, "lstm_variant") == 0) {
    if (value < 0 || value > 1) return fail(e, ND_ERR_INVALID, "lstm_variant must be 0 (W_hh in TMEM) or 1 (in smem)");
    lstm_tc_set_variant((int)value);
    return ND_OK;
  }
  if (strcmp(name, "use_graphs") == 0) {
    e->use_graphs = value != 0;
    return ND_OK;
  }
  return fail(e, ND_ERR_INVALID, std::string("unknown option '") + name + "'");
}

int nd_set_float(nd_engine* e, const char* name, double value) {
  if (!e || !name) return ND_ERR_INVALID;
  if (strcmp(name, "beta") == 0) {                // GNMTGlobalScorer beta (coverage penalty weight), object beam
    e->beta = (float)value;
    return ND_OK;
  }
  return fail(e, ND_ERR_INVALID, std::string("unknown float option ") + name);
}

int nd_debug_gemm_timeline(int64_t* dev_buf32) {
  gemm_tc_set_debug(reinterpret_cast<long long*>(dev_buf32));
  lstm_tc_set_debug(dev_buf32 ? reinterpret_cast<long long*>(dev_buf32) + 16 : nullptr);   // slots 16..31
  return ND_OK;
}

int nd_test_gemm(nd_engine* e, int32_t mode, const float* A, const float* W, const float* bias, const float* residual,
                 const float* ln_gamma, const float* ln_beta, float* C, int32_t M, int32_t N, int32_t K, int32_t relu,
                 void* stream) {
  if (!e) return ND_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  GemmParams p;
  p.A = A; p.lda = K; p.W = W; p.ldw = K; p.bias = bias; p.C = C; p.ldc = N; p.residual = residual; p.ldr = N;
  p.M = M; p.N = N; p.K = K; p.relu = relu;
  if (ln_gamma) { p.prologue = PRO_LAYERNORM; p.pg = ln_gamma; p.pb = ln_beta; p.eps = 1e-6f; }
  if (mode == ND_GEMM_SIMT_FP32) {
    ND_LAUNCH_CAT(e, ND_PROF_GEMM, st, gemm_simt(p, st));
    return ND_OK;
  }
  const char* why = "";
  if (!gemm_tc_available(&why)) return fail(e, ND_ERR_CUDA, why);
  float *hi = nullptr, *lo = nullptr, *cv = nullptr, *dv = nullptr;
  {
    // host-side weight preparation exactly as the engine's packer does it (test path only)
    std::vector<float> w((size_t)N * K), h((size_t)N * K), l((size_t)N * K);
    ND_CUDA(e, cudaMemcpy(w.data(), W, w.size() * sizeof(float), cudaMemcpyDefault));
    if (ln_gamma) {
      std::vector<float> g(K), b(K), bias_h(N, 0.f), cvec(N), dvec(N);
      ND_CUDA(e, cudaMemcpy(g.data(), ln_gamma, K * sizeof(float), cudaMemcpyDefault));
      ND_CUDA(e, cudaMemcpy(b.data(), ln_beta, K * sizeof(float), cudaMemcpyDefault));
      if (bias) ND_CUDA(e, cudaMemcpy(bias_h.data(), bias, N * sizeof(float), cudaMemcpyDefault));
      for (int n = 0; n < N; ++n) {
        double c = 0.0, dsum = 0.0;
        for (int k = 0; k < K; ++k) {
          const float wv = w[(size_t)n * K + k];
          c += (double)wv * g[k];
          dsum += (double)wv * b[k];
          w[(size_t)n * K + k] = wv * g[k];
        }
        cvec[n] = (float)c;
        dvec[n] = (float)(dsum + bias_h[n]);
      }
      ND_CUDA(e, cudaMalloc(&cv, N * sizeof(float)));
      ND_CUDA(e, cudaMalloc(&dv, N * sizeof(float)));
      ND_CUDA(e, cudaMemcpy(cv, cvec.data(), N * sizeof(float), cudaMemcpyHostToDevice));
      ND_CUDA(e, cudaMemcpy(dv, dvec.data(), N * sizeof(float), cudaMemcpyHostToDevice));
      p.prologue = PRO_NONE; p.ln_cvec = cv; p.ln_dvec = dv; p.bias = nullptr;
    }
    split_tf32_host(w.data(), h.data(), l.data(), w.size());
    if (mode != ND_GEMM_TC_3XTF32) h = w;            // single pass: the tensor core truncates the operand itself
    ND_CUDA(e, cudaMalloc(&hi, w.size() * sizeof(float)));
    ND_CUDA(e, cudaMalloc(&lo, w.size() * sizeof(float)));
    ND_CUDA(e, cudaMemcpy(hi, h.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    ND_CUDA(e, cudaMemcpy(lo, l.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    p.W = hi; p.W_lo = lo;

  }
  ++e->launches;
  const bool prof = (e->prof_mask >> ND_PROF_GEMM) & 1u;
  if (prof) prof_begin(e, ND_PROF_GEMM, st);
  cudaError_t err = gemm_tc(p, mode == ND_GEMM_TC_3XTF32 ? 3 : 1, st);
  if (prof) prof_end(e, st);
  cudaError_t err2 = cudaStreamSynchronize(st);
  if (hi) cudaFree(hi);
  if (lo) cudaFree(lo);
  if (cv) cudaFree(cv);
  if (dv) cudaFree(dv);
  if (err != cudaSuccess) { e->sticky = true; return fail(e, ND_ERR_CUDA, cudaGetErrorString(err)); }
  if (err2 != cudaSuccess) { e->sticky = true; return fail(e, ND_ERR_CUDA, cudaGetErrorString(err2)); }
  return ND_OK;
}

}  // extern "C"
