"""Import harness for the UNMODIFIED reference (achilles1989/NanoDecoder at /root/reference).

TEST INFRASTRUCTURE ONLY.  This module exists so that `oracle/make_golden.py` can run the
reference's own PyTorch code in the build container (where /root/reference is mounted) and write
golden vectors into `tests/golden/`.  Nothing here is imported by the product path
(`nanodecoder_b200/`), by `bench.py`'s own arm, or by anything that runs on the GPU box
(/root/reference does not exist there).

What it does:
  * puts /root/reference on sys.path;
  * installs `sys.modules` stubs for packages the reference imports at module scope but which are
    not installed here: configargparse (-> argparse), legacy torchtext (Field / Vocab / data.*),
    h5py, statsmodels.robust, matplotlib(.pyplot), librosa;
  * patches three torch-1.0 -> torch-2.x behaviour changes the reference relies on
    (integer Tensor.div truncation: translate/translator.py:732, onmt/translate/beam.py:134;
     `1 - bool_mask`: onmt/modules/global_attention.py:183).
The reference sources are never edited or copied.
"""
from __future__ import annotations

import argparse
import os
import sys
import types
from collections import Counter, defaultdict

REFERENCE_ROOT = os.environ.get("NANODECODER_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "translate"))


# ----------------------------------------------------------------------------- configargparse
def _make_configargparse() -> types.ModuleType:
    mod = types.ModuleType("configargparse")

    def _strip(kwargs):
        kwargs.pop("is_config_file_arg", None)
        kwargs.pop("is_write_out_config_file_arg", None)
        return kwargs

    class _Group(argparse._ArgumentGroup):
        def add(self, *a, **k):
            return self.add_argument(*a, **_strip(k))

    class ArgumentParser(argparse.ArgumentParser):
        def __init__(self, *a, **k):
            k.pop("config_file_parser_class", None)
            super().__init__(*a, **k)

        def add(self, *a, **k):
            return self.add_argument(*a, **_strip(k))

        def add_argument_group(self, *a, **k):
            g = _Group(self, *a, **k)
            self._action_groups.append(g)
            return g

    mod.ArgumentParser = ArgumentParser
    mod.Action = argparse.Action
    mod.HelpFormatter = argparse.HelpFormatter
    mod.ArgumentDefaultsHelpFormatter = argparse.ArgumentDefaultsHelpFormatter
    mod.SUPPRESS = argparse.SUPPRESS
    mod.ArgumentTypeError = argparse.ArgumentTypeError
    mod.YAMLConfigFileParser = object
    return mod


# ----------------------------------------------------------------------------- torchtext (legacy)
def _make_torchtext():
    tt = types.ModuleType("torchtext")
    data = types.ModuleType("torchtext.data")
    vocab = types.ModuleType("torchtext.vocab")

    class Vocab(object):
        """Legacy torchtext Vocab: specials first, then tokens by (-freq, token)."""

        def __init__(self, counter, max_size=None, min_freq=1, specials=("<pad>",), **_):
            self.freqs = counter
            counter = counter.copy()
            self.itos = list(specials)
            for tok in specials:
                del counter[tok]
            words = sorted(counter.items(), key=lambda kv: kv[0])
            words.sort(key=lambda kv: kv[1], reverse=True)
            for w, f in words:
                if f < min_freq or (max_size is not None and len(self.itos) >= max_size + len(specials)):
                    break
                self.itos.append(w)
            self.stoi = defaultdict(lambda: 0)
            self.stoi.update({t: i for i, t in enumerate(self.itos)})

        def __len__(self):
            return len(self.itos)

    class Field(object):
        vocab_cls = Vocab

        def __init__(self, sequential=True, use_vocab=True, init_token=None, eos_token=None,
                     pad_token="<pad>", unk_token="<unk>", dtype=None, postprocessing=None,
                     include_lengths=False, **kw):
            self.sequential = sequential
            self.use_vocab = use_vocab
            self.init_token = init_token
            self.eos_token = eos_token
            self.pad_token = pad_token if sequential else None
            self.unk_token = unk_token
            self.dtype = dtype
            self.postprocessing = postprocessing
            self.include_lengths = include_lengths

        def preprocess(self, x):
            return x

    class Example(object):
        pass

    class Dataset(object):
        def __init__(self, examples, fields, filter_pred=None):
            self.examples = list(examples)
            self.fields = dict(fields)

    class Iterator(object):
        def __init__(self, *a, **k):
            raise NotImplementedError("torchtext.data.Iterator stub: drive translate_batch directly")

    def batch(data, batch_size, batch_size_fn=None):
        mb = []
        for ex in data:
            mb.append(ex)
            if len(mb) == batch_size:
                yield mb
                mb = []
        if mb:
            yield mb

    vocab.Vocab = Vocab
    data.Field = Field
    data.Example = Example
    data.Dataset = Dataset
    data.Iterator = Iterator
    data.batch = batch
    tt.data = data
    tt.vocab = vocab
    return tt, data, vocab


_INSTALLED = False


def install():
    """Install stubs + torch compat patches, put the reference on sys.path. Idempotent."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    import torch

    sys.modules.setdefault("configargparse", _make_configargparse())
    if "torchtext" not in sys.modules:
        tt, data, vocab = _make_torchtext()
        sys.modules["torchtext"] = tt
        sys.modules["torchtext.data"] = data
        sys.modules["torchtext.vocab"] = vocab
    for name in ("h5py", "statsmodels", "statsmodels.robust", "matplotlib", "matplotlib.pyplot",
                 "librosa"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["statsmodels"], "robust"):
        sys.modules["statsmodels"].robust = sys.modules["statsmodels.robust"]
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    # --- torch 1.0 semantics the reference depends on -------------------------------------
    _orig_div = torch.Tensor.div
    _orig_truediv = torch.Tensor.__truediv__
    _orig_rsub = torch.Tensor.__rsub__

    def _is_int(t):
        return isinstance(t, torch.Tensor) and not t.is_floating_point() and t.dtype != torch.bool

    def div(self, other, *a, **k):
        if _is_int(self) and (isinstance(other, int) or _is_int(other)) and "rounding_mode" not in k:
            return _orig_div(self, other, rounding_mode="trunc")
        return _orig_div(self, other, *a, **k)

    def truediv(self, other):
        if _is_int(self) and (isinstance(other, int) or _is_int(other)):
            return _orig_div(self, other, rounding_mode="trunc")
        return _orig_truediv(self, other)

    def rsub(self, other):
        if self.dtype == torch.bool and other == 1:
            return ~self
        return _orig_rsub(self, other)

    torch.Tensor.div = div
    torch.Tensor.__truediv__ = truediv
    torch.Tensor.__rsub__ = rsub

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True


# ----------------------------------------------------------------------------- model / translator
FAMILY_FLAGS = {
    # SURVEY.md Appendix A step 3
    "t2t": ["-encoder_type", "transformer", "-decoder_type", "transformer"],
    "l2t": ["-encoder_type", "nano", "-decoder_type", "transformer", "-audio_enc_pooling", "1"],
    "nano2rnn": ["-encoder_type", "nano", "-decoder_type", "rnn", "-audio_enc_pooling", "1"],
    "brnn2rnn": ["-encoder_type", "brnn", "-decoder_type", "rnn"],
    "rnn2rnn": ["-encoder_type", "rnn", "-decoder_type", "rnn"],
    "cnn2cnn": ["-encoder_type", "cnn", "-decoder_type", "cnn"],
    "resnet2t": ["-encoder_type", "resnet", "-decoder_type", "transformer"],
    "resnet2rnn": ["-encoder_type", "resnet", "-decoder_type", "rnn"],
    "crnn2t": ["-encoder_type", "crnn", "-decoder_type", "transformer", "-audio_enc_pooling", "1"],
    "crnn2rnn": ["-encoder_type", "crnn", "-decoder_type", "rnn", "-audio_enc_pooling", "1"],
    "ctrans2t": ["-encoder_type", "ctransformer", "-decoder_type", "transformer"],
}


def model_flags(family, d=256, enc_layers=3, dec_layers=3, heads=8, ff=2048, extra=()):
    return (["-data", "x"] + FAMILY_FLAGS[family] +
            ["-enc_layers", str(enc_layers), "-dec_layers", str(dec_layers),
             "-enc_rnn_size", str(d), "-dec_rnn_size", str(d), "-tgt_word_vec_size", str(d),
             "-heads", str(heads), "-transformer_ff", str(ff),
             "-dropout", "0", "-global_attention", "mlp", "-param_init", "0"] + list(extra))


def build_reference_model(family, **kw):
    """-> (model, fields, model_opt): the reference's NMTModel built by its own factory."""
    install()
    import configargparse
    import models.opts as opts
    import models.model_builder as mb
    import inputters.inputter as inputters

    parser = configargparse.ArgumentParser()
    opts.model_opts(parser)
    opts.train_opts(parser)
    opt = parser.parse_args(model_flags(family, **kw))
    opt.brnn = opt.encoder_type == "brnn"          # models/train_single.py:61
    opt.model_type = "nano"                        # train.py:120
    fields = inputters.get_fields("nano", 0, 0)
    inputters._build_field_vocab(fields["tgt"], Counter({"A": 4, "C": 3, "G": 2, "T": 1}))
    model = mb.build_base_model(opt, fields, gpu=False, checkpoint=None).eval()
    model.generator.eval()
    return model, fields, opt


def build_reference_translator(model, fields, model_opt, beam_size=1, fast=False, max_length=100,
                               min_length=0, n_best=1, alpha=0.0, extra=()):
    install()
    import configargparse
    import models.opts as opts
    import onmt.translate
    from translate.translator import Translator

    parser = configargparse.ArgumentParser()
    opts.translate_opts(parser)
    argv = ["-model", "m", "-save_data", "s", "-beam_size", str(beam_size),
            "-max_length", str(max_length), "-min_length", str(min_length),
            "-n_best", str(n_best), "-alpha", str(alpha)]
    if fast:
        argv.append("-fast")
    argv += list(extra)
    topt = parser.parse_args(argv)
    topt.data_type = "nano"
    scorer = onmt.translate.GNMTGlobalScorer(topt)
    return Translator(model, fields, topt, model_opt, global_scorer=scorer, logger=None)


class FakeBatch(object):
    """What _run_encoder / from_batch read: translate/translator.py:413,543,549."""

    def __init__(self, src, src_lengths):
        import torch
        self.src = src                      # [T, B, 1] fp32
        self.src_lengths = src_lengths      # [B] int64
        self.batch_size = src.size(1)
        self.indices = torch.arange(self.batch_size)


class FakeData(object):
    data_type = "nano"
