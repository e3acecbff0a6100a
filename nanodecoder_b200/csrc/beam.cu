// On-device batched beam search step (reference: translate/translator.py:619-825, `--fast`).
// One warp per chunk; nothing round-trips to the host inside the decode loop.
//
// Differences in mechanism (not in results):
//   * finished chunks are not compacted out of the batch; a `retired` flag makes every per-chunk
//     kernel skip them (same saved work, no reallocation);
//   * the memory bank / cross K,V are never tiled or reordered (beams of a chunk share them);
//   * the self-attention cache is never reordered: `anc[row][j]` records which cache row holds
//     position j of the hypothesis now living in `row` (parent pointers);
//   * only the best `n_best` finished hypotheses are kept (stable in arrival order on equal scores),
//     which is what the reference's final `sorted(..., reverse=True)[:n_best]` selects.
#include <float.h>
#include <math.h>

#include "kernels.cuh"

namespace nd {

namespace {

constexpr int kMaxCandPerLane = 4;     // K*V <= 128

__global__ void beam_init_kernel(BeamParams p, int bos) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = p.B * p.K;
  if (i < rows) {
    p.st.topk_log_probs[i] = (i % p.K == 0) ? 0.0f : -INFINITY;     // translator.py:691-693
    p.st.cur_tok[i] = bos;
    p.st.parent[i] = i;
    p.st.alive_seq[(int64_t)i * (p.Lmax + 1)] = bos;
  }
  if (i < p.B) {
    p.st.retired[i] = 0;
    p.st.retire_step[i] = 0x7fffffff;
    p.st.top_finished[i] = 0;
    p.st.n_hyp[i] = 0;
    for (int n = 0; n < p.n_best; ++n) {
      p.st.hyp_score[i * p.n_best + n] = -INFINITY;
      p.st.hyp_len[i * p.n_best + n] = 0;
    }
  }
  if (i == 0) {
    *p.st.n_alive = p.B;
    *p.st.n_done = 0;
    *p.st.stop_step = 0x7fffffff;
  }
}

// length_penalty: fast mode ((5+step+1)/6)^alpha; object mode: divisor of the global score of hypotheses
// finishing at this step (beam.py:200-208)
__global__ void __launch_bounds__(128) beam_step_kernel(BeamParams p, float length_penalty) {
  pdl_launch_dependents();
  pdl_wait();
  const int bl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bl >= p.nb) return;
  const int b = p.b0 + bl;
  if (p.st.retired[b]) return;
  const bool obj = p.mode == 1;
  if (obj && p.step >= *p.st.stop_step) {              // every Beam was done before this step (translator.py:883-884)
    // the loop has ended: from the next step on every per-chunk kernel skips this chunk (its ancestor tables
    // are no longer extended, so the attention kernels must not walk them)
    if (lane == 0) {
      p.st.retired[b] = 1;
      if (bl == 0) *p.st.n_alive = 0;                // the projections of the remaining steps return at once
    }
    return;
  }
  const int K = p.K, V = p.V, NC = K * V, Lp1 = p.Lmax + 1;
  const int cur = p.step & 1, nxt = cur ^ 1;
  const int rows = p.B * K;
  const int* seq_cur = p.st.alive_seq + (int64_t)cur * rows * Lp1;
  int* seq_nxt = p.st.alive_seq + (int64_t)nxt * rows * Lp1;
  const int* anc_cur = p.st.anc + (int64_t)cur * rows * p.Lmax;
  int* anc_nxt = p.st.anc + (int64_t)nxt * rows * p.Lmax;

  // object mode, n-gram blocking (beam.py:101-124): a beam whose hypothesis so far contains the same n-gram twice
  // (n-grams touching an excluded token are neither recorded nor checked) has ALL its children set to -10e20
  unsigned ngram_fail = 0;                             // bit k: beam k is blocked
  if (obj && p.block_ngram > 0 && p.step > 0) {
    const int n = p.block_ngram, len = p.step;         // hypothesis tokens: seq_cur[row][1 .. step]
    for (int k = 0; k < K; ++k) {
      const int* hyp = seq_cur + (int64_t)(b * K + k) * Lp1 + 1;
      bool fail = false;
      for (int i = n - 1 + lane; i < len && !fail; i += 32) {        // n-gram ending at i (full length only can repeat)
        bool excl = false;
        for (int q = 0; q < n; ++q) excl |= ((p.excl_mask >> hyp[i - q]) & 1u) != 0;
        if (excl) continue;
        for (int i1 = n - 1; i1 < i && !fail; ++i1) {
          bool same = true, ex1 = false;
          for (int q = 0; q < n; ++q) {
            same &= hyp[i1 - q] == hyp[i - q];
            ex1 |= ((p.excl_mask >> hyp[i1 - q]) & 1u) != 0;
          }
          fail = same && !ex1;
        }
      }
      if (__any_sync(ND_FULL, fail)) ngram_fail |= 1u << k;
    }
  }

  // object mode, stepwise coverage penalty (GNMTGlobalScorer.update_score, beam.py:218-227, called first thing in
  // Beam.advance): scores += prev_penalty; scores -= penalty(coverage + this step's attention), in the beams' current order
  __shared__ float base_s[4][32];
  float* base = base_s[threadIdx.x >> 5];
  if (lane < K) base[lane] = p.st.topk_log_probs[b * K + lane];
  if (obj && p.stepwise && p.cov_mode != 0 && p.step > 0) {
    const int width = (int)p.mem_len[b / K];
    const float* cov_cur = p.cov + (int64_t)cur * rows * p.Tp;
    for (int k = 0; k < K; ++k) {
      const float* a = p.attn_step + (int64_t)(b * K + k) * p.Tp;
      const float* cc = cov_cur + (int64_t)(b * K + k) * p.Tp;
      float acc = 0.f;
      for (int t = lane; t < width; t += 32) {
        const float c = cc[t] + a[t];
        acc += p.cov_mode == 1 ? -logf(fminf(c, 1.0f)) : fmaxf(c, 1.0f);
      }
      acc = warp_sum(acc);
      if (p.cov_mode == 2) acc -= (float)width;
      if (lane == k) base[k] = (base[k] + p.cov_pen[b * K + k]) - p.beta * acc;
    }
  }
  __syncwarp();

  // candidate scores: (log_probs + beam score) / length_penalty            translator.py:718-725
  float cand[kMaxCandPerLane];
  unsigned used = 0;
#pragma unroll
  for (int i = 0; i < kMaxCandPerLane; ++i) {
    const int c = lane + 32 * i;
    cand[i] = -INFINITY;
    if (c < NC) {
      const int k = c / V;
      const float lpv = p.logp[((int64_t)b * K + k) * V + (c - k * V)] + base[k];
      if (obj) {
        cand[i] = (p.step > 0 && p.st.cur_tok[b * K + k] == p.eos) ? -1e20f : lpv;         // beam.py:94-100
        if ((ngram_fail >> k) & 1u) cand[i] = -10e20f;                                      // beam.py:123-124
      }
      else cand[i] = lpv / length_penalty;
    } else {
      used |= 1u << i;
    }
  }
  // top-K: K rounds of warp arg-max, ties -> lowest flat index
  float sel_score = 0.f;
  int sel_idx = 0;                                     // lane k keeps the k-th selection
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kMaxCandPerLane; ++i) {
      const int c = lane + 32 * i;
      if (!(used & (1u << i)) && (bi == 0x7fffffff || cand[i] > bv)) { bv = cand[i]; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(ND_FULL, bv, o);
      const int oi = __shfl_xor_sync(ND_FULL, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if ((bi & 31) == lane) used |= 1u << (bi >> 5);
    if (lane == k) { sel_score = bv; sel_idx = bi; }
  }
  // lane k (< K) now owns new beam k
  const bool mine = lane < K;
  const int beam = mine ? sel_idx / V : 0;
  const int tok = mine ? sel_idx - beam * V : 0;
  const int prow = b * K + beam;                       // parent row (global)
  const int nrow = b * K + lane;
  bool finished = mine && (tok == p.eos || (!obj && p.step + 1 == p.max_len));  // :753-755 | beam.py:140-141
  float new_lp = obj ? sel_score : sel_score * length_penalty;                 // :729 | beam.py:127
  const unsigned fin_mask = __ballot_sync(ND_FULL, finished);

  // new alive sequences / ancestor tables (warp-cooperative row copies)
  for (int k = 0; k < K; ++k) {
    const int pr = __shfl_sync(ND_FULL, prow, k);
    const int tk = __shfl_sync(ND_FULL, tok, k);
    const int nr = b * K + k;
    for (int j = lane; j <= p.step; j += 32) seq_nxt[(int64_t)nr * Lp1 + j] = seq_cur[(int64_t)pr * Lp1 + j];
    for (int j = lane; j < p.step; j += 32) anc_nxt[(int64_t)nr * p.Lmax + j] = anc_cur[(int64_t)pr * p.Lmax + j];
    if (lane == 0) {
      seq_nxt[(int64_t)nr * Lp1 + p.step + 1] = tk;
      anc_nxt[(int64_t)nr * p.Lmax + p.step] = pr;
    }
  }
  __syncwarp();

  // object mode, coverage penalty (beam.py:229-243 update_global_state; penalties.py:39-57): coverage of new beam k =
  // coverage of its parent + this step's attention of the parent row, over the columns the reference passes to
  // Beam.advance (memory_lengths[j] of the TILED vector: translator.py:905); penalty = beta * (wu: -sum log min(cov, 1),
  // summary: sum max(cov, 1) - width)
  float my_pen = 0.f;                                  // lane k: penalty of new beam k
  if (obj && p.cov_mode != 0) {
    const int width = (int)p.mem_len[b / K];
    const float* cov_cur = p.cov + (int64_t)cur * rows * p.Tp;
    float* cov_nxt = p.cov + (int64_t)nxt * rows * p.Tp;
    for (int k = 0; k < K; ++k) {
      const int pr = __shfl_sync(ND_FULL, prow, k);
      const float* a = p.attn_step + (int64_t)pr * p.Tp;
      const float* cc = cov_cur + (int64_t)pr * p.Tp;
      float* cn = cov_nxt + (int64_t)(b * K + k) * p.Tp;
      float acc = 0.f;
      for (int t = lane; t < width; t += 32) {
        const float c = (p.step > 0 ? cc[t] : 0.f) + a[t];
        cn[t] = c;
        acc += p.cov_mode == 1 ? -logf(fminf(c, 1.0f)) : fmaxf(c, 1.0f);
      }
      acc = warp_sum(acc);
      if (p.cov_mode == 2) acc -= (float)width;
      if (lane == k) my_pen = p.beta * acc;
    }
    // (stepwise mode reads this back as global_state["prev_penalty"], which the FIRST update_global_state sets to zero:
    // beam.py:231-232)
    if (mine) p.cov_pen[nrow] = (p.stepwise && p.step == 0) ? 0.f : my_pen;
  }

  if (fin_mask) {
    if (finished && !obj) new_lp = -1e10f;             // :760 (object mode keeps the score: beam.py:97-100 blocks the children)
    int top_fin = p.st.top_finished[b] | ((fin_mask & 1u) ? 1 : 0);            // :762
    int n_hyp = p.st.n_hyp[b];
    const int nb = p.n_best;
    float* hs = p.st.hyp_score + (int64_t)b * nb;
    int* hl = p.st.hyp_len + (int64_t)b * nb;
    int* hq = p.st.hyp_seq + (int64_t)b * nb * p.Lmax;
    int* ha = p.st.hyp_anc + (int64_t)b * nb * p.Lmax;
    int* hm = p.st.hyp_meta + (int64_t)b * nb;
    const int len = p.step + 1;
    // GNMTGlobalScorer.score with the length penalty "none" and a coverage penalty (beam.py:203-216): length_none
    // returns the beam's score tensor ITSELF and `normalized_probs -= penalty` is in place, so each of the step's
    // finished hypotheses lowers the RUNNING scores of all K beams by their coverage penalties -- and the scores stored
    // in Beam.finished are views of that tensor, so every one of them reads the value after ALL of the step's calls
    const bool alias_scores = obj && p.cov_mode != 0 && p.lp_mode == 0 && !p.stepwise;
    if (alias_scores)
      for (int c = __popc(fin_mask); c > 0; --c) new_lp -= my_pen;
    for (int k = 0; k < K; ++k) {                      // finished beams in beam order (:773-778)
      if (!(fin_mask & (1u << k))) continue;
      float sc = __shfl_sync(ND_FULL, sel_score, k);
      const float pen_k = __shfl_sync(ND_FULL, my_pen, k);
      if (alias_scores) {
        sc = __shfl_sync(ND_FULL, new_lp, k);
      } else if (obj) {
        // global score: length penalty (a new tensor), then coverage -- unless the penalty already steered the search
        sc = p.stepwise ? sc / length_penalty : sc / length_penalty - pen_k;
      }
      ++n_hyp;
      // stable insertion into the best-n_best list (descending score, earlier first on ties)
      int pos = nb;
      for (int i = 0; i < nb; ++i) {
        const float cur_s = hs[i];
        const bool empty = hl[i] == 0;
        if (empty || sc > cur_s) { pos = i; break; }
      }
      __syncwarp();
      if (pos < nb) {
        for (int i = nb - 1; i > pos; --i) {
          for (int j = lane; j < p.Lmax; j += 32) {
            hq[i * p.Lmax + j] = hq[(i - 1) * p.Lmax + j];
            ha[i * p.Lmax + j] = ha[(i - 1) * p.Lmax + j];
          }
          __syncwarp();
          if (lane == 0) { hs[i] = hs[i - 1]; hl[i] = hl[i - 1]; hm[i] = hm[i - 1]; }
          __syncwarp();
        }
        const int nr = b * K + k;
        for (int j = lane; j < len; j += 32) {
          hq[pos * p.Lmax + j] = seq_nxt[(int64_t)nr * Lp1 + 1 + j];
          ha[pos * p.Lmax + j] = anc_nxt[(int64_t)nr * p.Lmax + j];       // rows that ran steps 0 .. step
        }
        if (lane == 0) { hs[pos] = sc; hl[pos] = len; hm[pos] = ((n_hyp - 1) << 8) | k; }
        __syncwarp();
      }
    }
    if (lane == 0) {
      p.st.top_finished[b] = top_fin;
      p.st.n_hyp[b] = n_hyp;
      if (!obj && top_fin && n_hyp >= nb) {            // :781
        p.st.retired[b] = 1;
        p.st.retire_step[b] = p.step;
        atomicSub(p.st.n_alive, 1);
      }
      if (obj && (top_fin & 1) && n_hyp >= nb && !(top_fin & 2)) {      // Beam.done() became true (beam.py:151-152)
        p.st.top_finished[b] = top_fin | 2;
        if (atomicAdd(p.st.n_done, 1) + 1 == p.B) atomicMin(p.st.stop_step, p.step + 1);
      }
    }
  }
  if (mine) {
    p.st.topk_log_probs[nrow] = new_lp;
    p.st.cur_tok[nrow] = tok;
    p.st.parent[nrow] = prow;
  }
}

// Object mode, after the loop: Beam.sort_finished(minimum=n_best) (beam.py:154-168) tops the finished list up
// from the live beam in beam order; they arrive last, so they rank after equal-score finished entries.
__global__ void __launch_bounds__(32) beam_object_fill_kernel(BeamParams p, float gs_none, float alpha) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int K = p.K, Lp1 = p.Lmax + 1, nb = p.n_best;
  int n_hyp = p.st.n_hyp[b];
  if (n_hyp >= nb) return;
  const int steps = min(p.max_len, *p.st.stop_step);   // steps actually executed = len(next_ys) - 1
  const int buf = ((steps - 1) & 1) ^ 1;               // alive_seq buffer written by the last executed step
  const int* seq = p.st.alive_seq + (int64_t)buf * p.B * K * Lp1;
  const int* anc = p.st.anc + (int64_t)buf * p.B * K * p.Lmax;
  int* ha = p.st.hyp_anc + (int64_t)b * nb * p.Lmax;
  float div = 1.0f;                                    // global score divisor at len(next_ys) = steps + 1
  if (p.lp_mode == 1) div = (float)(pow(5.0 + (double)(steps + 1), (double)alpha) / pow(6.0, (double)alpha));
  else if (p.lp_mode == 2) div = (float)(steps + 1);
  float* hs = p.st.hyp_score + (int64_t)b * nb;
  int* hl = p.st.hyp_len + (int64_t)b * nb;
  int* hq = p.st.hyp_seq + (int64_t)b * nb * p.Lmax;
  int* hm = p.st.hyp_meta + (int64_t)b * nb;
  const bool alias_scores = p.cov_mode != 0 && p.lp_mode == 0 && !p.stepwise;
  if (alias_scores) {
    // the in-place subtraction of the coverage penalty (see beam_step_kernel) once per topped-up hypothesis; the views
    // stored for hypotheses that finished at the LAST step point into the same tensor and move with it
    const int m = min(nb - n_hyp, K);
    if (lane < K)
      for (int c = 0; c < m; ++c) p.st.topk_log_probs[b * K + lane] -= p.cov_pen[b * K + lane];
    __syncwarp();
    if (lane == 0)
      for (int q = 0; q < n_hyp; ++q)
        if (hl[q] == steps) hs[q] = p.st.topk_log_probs[b * K + (hm[q] & 255)];
    __syncwarp();
  }
  for (int i = 0; n_hyp < nb && i < K; ++i, ++n_hyp) {
    const float sc = alias_scores ? p.st.topk_log_probs[b * K + i]
                                  : p.st.topk_log_probs[b * K + i] / div -
                                        ((p.cov_mode != 0 && !p.stepwise) ? p.cov_pen[b * K + i] : 0.f);
    const int pos = n_hyp;                             // appended; the list is sorted below
    for (int j = lane; j < steps; j += 32) {
      hq[pos * p.Lmax + j] = seq[(int64_t)(b * K + i) * Lp1 + 1 + j];
      ha[pos * p.Lmax + j] = anc[(int64_t)(b * K + i) * p.Lmax + j];
    }
    if (lane == 0) { hs[pos] = sc; hl[pos] = steps; hm[pos] = (n_hyp << 8) | i; }
    __syncwarp();
  }
  // finished.sort(key=lambda a: -a[0]) (beam.py:165): descending score, stable in arrival order
  for (int pass = 0; pass < n_hyp; ++pass) {
    for (int q = 0; q + 1 < n_hyp - pass; ++q) {
      const bool swap = hs[q + 1] > hs[q] || (hs[q + 1] == hs[q] && (hm[q + 1] >> 8) < (hm[q] >> 8));
      __syncwarp();
      if (swap) {
        for (int j = lane; j < p.Lmax; j += 32) {
          const int t0 = hq[q * p.Lmax + j]; hq[q * p.Lmax + j] = hq[(q + 1) * p.Lmax + j]; hq[(q + 1) * p.Lmax + j] = t0;
          const int t1 = ha[q * p.Lmax + j]; ha[q * p.Lmax + j] = ha[(q + 1) * p.Lmax + j]; ha[(q + 1) * p.Lmax + j] = t1;
        }
        __syncwarp();
        if (lane == 0) {
          const float ts = hs[q]; hs[q] = hs[q + 1]; hs[q + 1] = ts;
          const int tl = hl[q]; hl[q] = hl[q + 1]; hl[q + 1] = tl;
          const int tm = hm[q]; hm[q] = hm[q + 1]; hm[q + 1] = tm;
        }
        __syncwarp();
      }
    }
  }
  (void)gs_none;
}

// one CTA per (chunk, hypothesis, step): copy (or zero) one attention row
__global__ void beam_gather_attention_kernel(BeamState st, const float* __restrict__ hist, int n_best, int Lmax, int max_len,
                                             int rows, int Tp, float* __restrict__ out) {
  const int64_t i = blockIdx.x;                        // (b * n_best + n) * max_len + j
  const int j = (int)(i % max_len);
  const int64_t bn = i / max_len;
  const int len = st.hyp_len[bn];
  float* o = out + i * Tp;
  if (j < len) {
    const int row = st.hyp_anc[bn * Lmax + j];
    const float* h = hist + ((int64_t)j * rows + row) * Tp;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) o[t] = h[t];
  } else {
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) o[t] = 0.f;
  }
}

__global__ void beam_attention_width_kernel(BeamState st, const int64_t* __restrict__ mem_len, int B, int K, int n_best,
                                            int mode, int* __restrict__ widths) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // b * n_best + n
  if (i >= B * n_best) return;
  const int b = i / n_best;
  int src = b / K;                                           // object mode: all chunks stay in the batch (:905)
  if (mode == 0) {
    const int s = st.hyp_len[i] - 1;                         // the step at which the hypothesis finished
    int pos = 0;                                             // position of the chunk among those still in the batch
    for (int c = 0; c < b; ++c) pos += st.retire_step[c] >= s;
    const int want = pos / K;                                // :776 memory_lengths[i] of the tiled vector
    src = b;
    for (int c = 0, seen = 0; c < B; ++c) {
      if (st.retire_step[c] >= s) {
        if (seen == want) { src = c; break; }
        ++seen;
      }
    }
  }
  widths[i] = (int)mem_len[src];
}

__global__ void beam_finalize_kernel(BeamParams p, int64_t* out_ids, int* out_lens, float* out_scores) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x;                            // (b, n)
  const int len = p.st.hyp_len[i];
  if (threadIdx.x == 0) { out_lens[i] = len; out_scores[i] = p.st.hyp_score[i]; }
  for (int j = threadIdx.x; j < p.max_len; j += blockDim.x)
    out_ids[(int64_t)i * p.max_len + j] = j < len ? (int64_t)p.st.hyp_seq[(int64_t)i * p.Lmax + j] : -1;
}

}  // namespace

cudaError_t beam_init(const BeamParams& p, int bos, cudaStream_t stream) {
  const int n = p.B * p.K;
  launch_k(beam_init_kernel, dim3(cdiv(n, 256)), dim3(256), 0, stream, p, bos);
  return cudaGetLastError();
}

cudaError_t beam_step(const BeamParams& p, cudaStream_t stream) {
  if (p.K * p.V > 32 * kMaxCandPerLane || p.K > 32) return cudaErrorInvalidValue;
  // ((5 + step + 1) / 6) ** alpha in double like the Python expression (translator.py:720-721)
  double lp = pow((5.0 + (double)(p.step + 1)) / 6.0, (double)p.alpha);
  if (p.mode == 1) {
    // object mode: divisor of the global score for hypotheses finishing now, len(next_ys) = step + 2
    // (penalties.py:65-88: wu ((5 + n) ** a) / ((5 + 1) ** a), avg n, none 1)
    const double n = (double)(p.step + 2);
    lp = p.lp_mode == 1 ? pow(5.0 + n, (double)p.alpha) / pow(6.0, (double)p.alpha) : (p.lp_mode == 2 ? n : 1.0);
  }
  if (p.nb <= 0) return cudaSuccess;
  launch_k(beam_step_kernel, dim3(cdiv(p.nb, 4)), dim3(128), 0, stream, p, (float)lp);
  return cudaGetLastError();
}

cudaError_t beam_finalize(const BeamParams& p, int64_t* out_ids, int* out_lens, float* out_scores,
                          cudaStream_t stream) {
  if (p.mode == 1) beam_object_fill_kernel<<<p.B, 32, 0, stream>>>(p, 1.0f, p.alpha);
  launch_k(beam_finalize_kernel, dim3(p.B * p.n_best), dim3(128), 0, stream, p, out_ids, out_lens, out_scores);
  return cudaGetLastError();
}

cudaError_t beam_gather_attention(const BeamState& st, const float* hist, const int64_t* mem_len, int B, int K, int n_best,
                                  int Lmax, int max_len, int rows, int Tp, int mode, float* out, int* widths,
                                  cudaStream_t stream) {
  const int64_t n = (int64_t)B * n_best * max_len;
  if (n <= 0) return cudaSuccess;
  beam_gather_attention_kernel<<<(unsigned)n, 128, 0, stream>>>(st, hist, n_best, Lmax, max_len, rows, Tp, out);
  if (widths) beam_attention_width_kernel<<<cdiv(B * n_best, 128), 128, 0, stream>>>(st, mem_len, B, K, n_best, mode, widths);
  return cudaGetLastError();
}

}  // namespace nd
