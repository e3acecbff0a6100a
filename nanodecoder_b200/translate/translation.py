"""ids -> tokens (reference: translate/translation.py:7-105, 108-156)."""
from __future__ import annotations


class Translation(object):
    """Container for one translated chunk (translate/translation.py:108-156)."""

    def __init__(self, src, src_raw, pred_sents, attn, pred_scores, tgt_sent, gold_score):
        self.src = src
        self.src_raw = src_raw
        self.pred_sents = pred_sents
        self.attns = attn
        self.pred_scores = pred_scores
        self.gold_sent = tgt_sent
        self.gold_score = gold_score

    def log(self, sent_number):
        out = "\nSENT {}: {}\n".format(sent_number, self.src_raw)
        out += "PRED {}: {}\n".format(sent_number, " ".join(self.pred_sents[0]))
        out += "PRED SCORE: {:.4f}\n".format(self.pred_scores[0])
        if len(self.pred_sents) > 1:
            out += "\nBEST HYP:\n"
            for score, sent in zip(self.pred_scores, self.pred_sents):
                out += "[{:.4f}] {}\n".format(score, sent)
        return out


class TranslationBuilder(object):
    def __init__(self, data, fields, n_best=1, replace_unk=False, has_tgt=False):
        self.data = data
        self.fields = fields
        self.n_best = n_best
        self.replace_unk = replace_unk
        self.has_tgt = has_tgt

    def _build_target_tokens(self, pred):
        # translate/translation.py:31-41: stop at the first </s>
        itos = self.fields["tgt"].vocab.itos
        eos = self.fields["tgt"].eos_token
        tokens = []
        for tok in (pred.tolist() if hasattr(pred, "tolist") else pred):
            if tok < 0:
                break
            t = itos[tok]
            if t == eos:
                break
            tokens.append(t)
        return tokens

    def from_batch(self, translation_batch):
        batch = translation_batch["batch"]
        order = sorted(range(batch.batch_size), key=lambda i: int(batch.indices[i]))     # :55-64
        translations = []
        for i in order:
            preds = translation_batch["predictions"][i]
            pred_sents = [self._build_target_tokens(preds[n]) for n in range(min(self.n_best, len(preds)))]
            translations.append(Translation(None, None, pred_sents, translation_batch["attention"][i],
                                            translation_batch["scores"][i], None,
                                            translation_batch["gold_score"][i]))
        return translations
