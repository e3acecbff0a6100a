"""ORACLE (test infrastructure): restatement of the reference's decode loops
(translate/translator.py) over ``oracle.model.OracleModel``.  Citations relative to /root/reference.
"""
from __future__ import annotations

import torch

UNK, PAD, BOS, EOS = 0, 1, 2, 3


def _run_encoder(model, src, src_lengths):
    """translate/translator.py:542-559."""
    enc_states, memory_bank, lengths = model.encoder(src, src_lengths)
    return src, enc_states, memory_bank, lengths


def _decode_and_generate(model, decoder_in, memory_bank, memory_lengths, step):
    """translate/translator.py:561-617 (no copy attention)."""
    dec_out, attn = model.decoder(decoder_in, memory_bank, memory_lengths=memory_lengths, step=step)
    return model.generator(dec_out.squeeze(0)), attn


def greedy(model, src, src_lengths, max_length=100, min_length=0, return_attention=False,
           trace_logits=None):
    """translate/translator.py:396-503 with keep_topk == 1 (argmax; :371-375).
    Runs all ``max_length`` steps — the reference has no EOS early exit (:455).
    -> dict(predictions [B,L] int64, scores [B] (LAST step's log-prob only, :494),
            attention [L,B,T] or None, memory_bank, memory_lengths)"""
    with torch.no_grad():
        B = src.size(1)
        src, enc_states, memory_bank, memory_lengths = _run_encoder(model, src, src_lengths)
        model.decoder.init_state(src, memory_bank, enc_states)
        seq = torch.full([B, 1], BOS, dtype=torch.long)
        attns = []
        topk_scores = None
        for step in range(max_length):
            decoder_input = seq[:, -1].view(1, -1, 1)
            log_probs, attn = _decode_and_generate(model, decoder_input, memory_bank,
                                                   memory_lengths, step)
            if step < min_length:
                log_probs[:, EOS] = -1e20                           # :469-470
            if trace_logits is not None:
                trace_logits.append(log_probs.clone())
            topk_scores, topk_ids = log_probs.topk(1, dim=-1)       # :375
            seq = torch.cat([seq, topk_ids.view(-1, 1)], -1)        # :477
            if return_attention:
                attns.append(attn.reshape(-1, attn.size(-1)))
        return {"predictions": seq[:, 1:].contiguous(), "scores": topk_scores[:, 0].clone(),
                "attention": torch.stack(attns) if attns else None,
                "memory_bank": memory_bank, "memory_lengths": memory_lengths}


def _tile(x, count, dim=0):
    """onmt/utils/misc.py:28-47: each batch entry repeated ``count`` times consecutively."""
    return x.repeat_interleave(count, dim=dim)


def beam_fast(model, src, src_lengths, beam_size=5, max_length=100, min_length=0, n_best=1,
              alpha=0.0, margins=None, return_attention=False, attn_rows_first=False):
    """translate/translator.py:619-825 (``--fast`` batched beam search).
    -> dict(predictions: list[B] of list[n_best] of LongTensor, scores: list[B] of list[float],
            attention: list[B] of list[n_best] of [len, width] tensors when return_attention (:744-750, :776: the width
            is ``memory_lengths[i]`` of the TILED length vector indexed by the chunk's position among the chunks still in
            the batch -- i.e. the length of chunk alive[i // beam_size], the reference's own indexing).
            ``attn_rows_first`` reads a decoder's attention as [1, rows, T] whatever its shape (diagnostics only)))
    ``margins`` (optional list, diagnostics only): filled with one float per chunk = the smallest gap, over all steps
    the chunk was alive, between two neighbouring candidates among the best beam_size + 1 of the step (the pruning
    boundary and the order of the kept beams).  A chunk whose margin is of the order of fp32 rounding (~1e-6) can take
    another search path under any re-association of the sums: that is what "explained by a logit tie" means for a
    beam search."""
    with torch.no_grad():
        B, K = src.size(1), beam_size
        src, enc_states, memory_bank, src_lengths = _run_encoder(model, src, src_lengths)
        model.decoder.init_state(src, memory_bank, enc_states)
        model.decoder.map_state(lambda s, dim: _tile(s, K, dim))    # :667-668
        memory_bank = _tile(memory_bank, K, 1)                      # :673
        memory_lengths = _tile(src_lengths, K)                      # :676
        top_beam_finished = torch.zeros([B], dtype=torch.bool)
        batch_offset = torch.arange(B, dtype=torch.long)
        beam_offset = torch.arange(0, B * K, step=K, dtype=torch.long)
        alive_seq = torch.full([B * K, 1], BOS, dtype=torch.long)
        topk_log_probs = torch.tensor([0.0] + [float("-inf")] * (K - 1)).repeat(B)   # :691-693
        hypotheses = [[] for _ in range(B)]
        results = {"predictions": [[] for _ in range(B)], "scores": [[] for _ in range(B)],
                   "attention": [[] for _ in range(B)]}
        alive_attn = None

        for step in range(max_length):
            decoder_input = alive_seq[:, -1].view(1, -1, 1)
            log_probs, attn = _decode_and_generate(model, decoder_input, memory_bank,
                                                   memory_lengths, step)
            V = log_probs.size(-1)
            if step < min_length:
                log_probs[:, EOS] = -1e20                           # :714-715
            log_probs = log_probs + topk_log_probs.view(-1).unsqueeze(1)        # :718
            length_penalty = ((5.0 + (step + 1)) / 6.0) ** alpha    # :720-721
            curr_scores = (log_probs / length_penalty).reshape(-1, K * V)
            topk_scores, topk_ids = curr_scores.topk(K, dim=-1)     # :726
            if margins is not None:
                if step == 0:
                    margins[:] = [float("inf")] * B
                top = curr_scores.topk(min(K + 1, curr_scores.size(1)), dim=-1)[0]
                gaps = top[:, :-1] - top[:, 1:]
                gaps = torch.where(torch.isfinite(gaps), gaps, torch.full_like(gaps, float("inf"))).min(1)[0]
                for i in range(gaps.size(0)):
                    b = int(batch_offset[i])
                    margins[b] = min(margins[b], float(gaps[i]))
            topk_log_probs = topk_scores * length_penalty           # :729
            topk_beam_index = torch.div(topk_ids, V, rounding_mode="trunc")    # :732
            topk_ids = topk_ids.fmod(V)                             # :733
            batch_index = topk_beam_index + beam_offset[:topk_beam_index.size(0)].unsqueeze(1)
            select_indices = batch_index.view(-1)
            alive_seq = torch.cat([alive_seq.index_select(0, select_indices),
                                   topk_ids.view(-1, 1)], -1)       # :742-744
            if return_attention:                                    # :744-750
                if attn_rows_first:
                    attn = attn.reshape(1, -1, attn.size(-1))
                current_attn = attn.index_select(1, select_indices)
                alive_attn = current_attn if alive_attn is None else \
                    torch.cat([alive_attn.index_select(1, select_indices), current_attn], 0)
            is_finished = topk_ids.eq(EOS)
            if step + 1 == max_length:
                is_finished.fill_(True)                             # :754-755
            if is_finished.any():
                topk_log_probs = topk_log_probs.masked_fill(is_finished, -1e10)   # :760
                top_beam_finished |= is_finished[:, 0]
                predictions = alive_seq.view(-1, K, alive_seq.size(-1))
                attention = alive_attn.view(alive_attn.size(0), -1, K, alive_attn.size(-1)) \
                    if alive_attn is not None else None             # :764-767
                non_finished_batch = []
                for i in range(is_finished.size(0)):
                    b = int(batch_offset[i])
                    for j in is_finished[i].nonzero().view(-1).tolist():
                        hypotheses[b].append((topk_scores[i, j], predictions[i, j, 1:],
                                              attention[:, i, j, :memory_lengths[i]].clone()     # :776 (sic)
                                              if attention is not None else None))
                    if top_beam_finished[i] and len(hypotheses[b]) >= n_best:   # :781
                        best = sorted(hypotheses[b], key=lambda x: x[0], reverse=True)
                        for n, (score, pred, hyp_attn) in enumerate(best):
                            if n >= n_best:
                                break
                            results["scores"][b].append(float(score))
                            results["predictions"][b].append(pred.clone())
                            results["attention"][b].append(hyp_attn if hyp_attn is not None else [])
                    else:
                        non_finished_batch.append(i)
                non_finished = torch.tensor(non_finished_batch, dtype=torch.long)
                if len(non_finished) == 0:
                    break
                top_beam_finished = top_beam_finished.index_select(0, non_finished)
                batch_offset = batch_offset.index_select(0, non_finished)
                topk_log_probs = topk_log_probs.index_select(0, non_finished)
                batch_index = batch_index.index_select(0, non_finished)
                select_indices = batch_index.view(-1)
                alive_seq = predictions.index_select(0, non_finished).view(-1, alive_seq.size(-1))
                if alive_attn is not None:                          # :806-809
                    alive_attn = attention.index_select(1, non_finished).view(alive_attn.size(0), -1,
                                                                              alive_attn.size(-1))
            memory_bank = memory_bank.index_select(1, select_indices)           # :813-817
            memory_lengths = memory_lengths.index_select(0, select_indices)
            model.decoder.map_state(lambda s, dim: s.index_select(dim, select_indices))
        return results


def beam_object(model, src, src_lengths, beam_size=5, max_length=100, min_length=0, n_best=1,
                length_penalty="none", alpha=0.0, return_attention=False, block_ngram_repeat=0,
                exclusion_tokens=(), coverage_penalty="none", beta=0.0, stepwise_penalty=False):
    """translate/translator.py:827-926 (``_translate_batch``, the default when ``--fast`` is absent) with
    onmt/translate/beam.py:74-178 (``Beam.advance / done / sort_finished / get_hyp``) and the
    GNMTGlobalScorer of beam.py:181-208 with coverage penalty "none" (penalties.py:59-63), length penalty
    none / wu / avg (penalties.py:65-88), coverage penalty none / wu / summary weighted by beta (penalties.py:39-57,
    beam.py:203-243), n-gram blocking (beam.py:101-124); no stepwise penalty.
    -> dict(predictions: list[B] of list[n_best] of LongTensor, scores: list[B] of list[n_best] of float)"""
    with torch.no_grad():
        B, K = src.size(1), beam_size
        src, enc_states, memory_bank, src_lengths = _run_encoder(model, src, src_lengths)
        model.decoder.init_state(src, memory_bank, enc_states)
        model.decoder.map_state(lambda s, dim: _tile(s, K, dim))    # :872-873
        memory_bank = _tile(memory_bank, K, 1)                      # :878
        memory_lengths = _tile(src_lengths, K)                      # :879

        coverage = [None] * B                                        # beam.global_state["coverage"] [K, width]

        def cov_pen(cov):                                           # penalties.py:39-57
            if coverage_penalty == "wu":
                return beta * (-torch.min(cov, cov.clone().fill_(1.0)).log().sum(1))
            if coverage_penalty == "summary":
                return beta * (torch.max(cov, cov.clone().fill_(1.0)).sum(1) - cov.size(1))
            return torch.zeros(cov.size(0))

        def global_score(scores, n_ys, b=None):                     # beam.py:200-216, penalties.py:65-94
            if length_penalty == "wu":
                out = scores / (((5 + n_ys) ** alpha) / ((5 + 1) ** alpha))
            elif length_penalty == "avg":
                out = scores / n_ys
            else:
                out = scores                                        # length_none returns ITS ARGUMENT (penalties.py:90-94)
            if b is not None and coverage_penalty != "none" and not stepwise_penalty:     # beam.py:208
                # beam.py:214 `normalized_probs -= penalty` is IN PLACE: with the length penalty "none" it subtracts the
                # coverage penalty from the beam's RUNNING scores, once per call (= once per finished hypothesis, and once
                # per hypothesis topped up by sort_finished), and the search goes on from the lowered scores
                out.sub_(cov_pen(coverage[b]))
            return out
        exclusion = set(int(t) for t in exclusion_tokens)
        track_attn = return_attention or coverage_penalty != "none"
        prev_penalty = [None] * B                                    # beam.global_state["prev_penalty"]

        # per chunk Beam state (beam.py:20-60)
        scores = [torch.zeros(K) for _ in range(B)]
        prev_ks = [[] for _ in range(B)]
        next_ys = [[torch.full((K,), PAD, dtype=torch.long)] for _ in range(B)]
        for b in range(B):
            next_ys[b][0][0] = BOS
        eos_top = [False] * B
        finished = [[] for _ in range(B)]
        beam_attns = [[] for _ in range(B)]                          # Beam.attn (beam.py:135)

        for step in range(max_length):
            if all(eos_top[b] and len(finished[b]) >= n_best for b in range(B)):    # :883-884, beam.py:151-152
                break
            inp = torch.stack([next_ys[b][-1] for b in range(B)]).view(1, -1, 1)
            out, step_attn = _decode_and_generate(model, inp, memory_bank, memory_lengths, step)
            out = out.view(B, K, -1)
            step_attn = step_attn.view(B, K, -1)                    # :900
            select = []
            for b in range(B):
                word_probs = out[b]
                V = word_probs.size(1)
                if stepwise_penalty and prev_penalty[b] is not None:                # beam.py:87-88, 218-227 update_score
                    scores[b].add_(prev_penalty[b])
                    scores[b].sub_(cov_pen(coverage[b] + step_attn[b, :, :memory_lengths[b]])
                                   if coverage_penalty != "none" else torch.zeros(K))
                cur_len = len(next_ys[b])
                if cur_len < min_length:                            # beam.py:89-92
                    word_probs[:, EOS] = -1e20
                if len(prev_ks[b]) > 0:
                    beam_scores = word_probs + scores[b].unsqueeze(1)
                    for i in range(K):                              # "Don't let EOS have children" :97-100
                        if next_ys[b][-1][i] == EOS:
                            beam_scores[i] = -1e20
                    if block_ngram_repeat > 0:                      # :101-124
                        le = len(next_ys[b])
                        for j in range(K):
                            hyp, kk = [], j                         # get_hyp(le - 1, j)
                            for q in range(le - 2, -1, -1):
                                hyp.append(int(next_ys[b][q + 1][kk]))
                                kk = int(prev_ks[b][q][kk])
                            hyp = hyp[::-1]
                            ngrams, fail, gram = set(), False, []
                            for i in range(le - 1):
                                gram = (gram + [hyp[i]])[-block_ngram_repeat:]
                                if set(gram) & exclusion:
                                    continue
                                if tuple(gram) in ngrams:
                                    fail = True
                                ngrams.add(tuple(gram))
                            if fail:
                                beam_scores[j] = -10e20
                else:
                    beam_scores = word_probs[0]
                best_scores, best_id = beam_scores.reshape(-1).topk(K, 0, True, True)
                scores[b] = best_scores
                prev_k = torch.div(best_id, V, rounding_mode="trunc")
                prev_ks[b].append(prev_k)
                next_ys[b].append(best_id - prev_k * V)
                if track_attn:                                      # :904-905 (memory_lengths is the TILED vector, sic)
                    beam_attns[b].append(step_attn[b, :, :memory_lengths[b]].index_select(0, prev_k))
                    if len(prev_ks[b]) == 1:                        # update_global_state, beam.py:229-243
                        coverage[b] = beam_attns[b][-1]
                        prev_penalty[b] = torch.zeros(K)
                    else:
                        coverage[b] = coverage[b].index_select(0, prev_k).add(beam_attns[b][-1])
                        prev_penalty[b] = cov_pen(coverage[b]) if coverage_penalty != "none" else torch.zeros(K)
                for i in range(K):                                  # :140-144
                    if next_ys[b][-1][i] == EOS:
                        s_i = global_score(scores[b], len(next_ys[b]), b)[i]     # a VIEW, like the reference's
                        finished[b].append((s_i, len(next_ys[b]) - 1, i))
                if next_ys[b][-1][0] == EOS:                        # :147-149
                    eos_top[b] = True
                select.append(prev_k + b * K)
            select = torch.cat(select)
            model.decoder.map_state(lambda s, dim: s.index_select(dim, select))      # :913-914

        results = {"predictions": [], "scores": [], "attention": []}
        for b in range(B):
            i = 0
            while len(finished[b]) < n_best:                        # sort_finished(minimum=n_best) :157-163
                s_i = global_score(scores[b], len(next_ys[b]), b)[i]
                finished[b].append((s_i, len(next_ys[b]) - 1, i))
                i += 1
            finished[b] = [(float(sc), t, k) for sc, t, k in finished[b]]      # the views are read here, after all calls
            finished[b].sort(key=lambda a: -a[0])                   # stable
            hyps, atts = [], []
            for (sc, t, k) in finished[b][:n_best]:                 # get_hyp :170-178
                hyp, att = [], []
                for j in range(t - 1, -1, -1):
                    hyp.append(int(next_ys[b][j + 1][k]))
                    if return_attention:
                        att.append(beam_attns[b][j][k])
                    k = int(prev_ks[b][j][k])
                hyps.append(torch.tensor(hyp[::-1], dtype=torch.long))
                atts.append(torch.stack(att[::-1]) if att else [])
            results["predictions"].append(hyps)
            results["attention"].append(atts)
            results["scores"].append([sc for sc, _, _ in finished[b][:n_best]])
        return results


def build_target_tokens(pred, itos):
    """translate/translation.py:31-41: ids -> tokens, cut at the first ``</s>``."""
    tokens = []
    for tok in pred.tolist() if hasattr(pred, "tolist") else pred:
        if itos[tok] == "</s>":
            break
        tokens.append(itos[tok])
    return tokens


def count_bases(pred_ids: torch.Tensor) -> int:
    """Bases = tokens before the first </s> (what translate.py:85-95 counts as len(c_bpread) when
    stride == length; specials other than </s> are counted as the reference's string join keeps
    them as tokens too)."""
    is_eos = pred_ids.eq(EOS)
    first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full_like(pred_ids[:, 0], pred_ids.size(1)))
    return int(first.sum())
