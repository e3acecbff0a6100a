"""Checkpoint layout of the reference, loaded without torchtext / onmt installed.

Reference: ``onmt/models/model_saver.py:105-115`` writes
``{'model', 'generator', 'vocab', 'opt', 'optim'}``; ``models/model_builder.py:217-233`` loads it,
back-fills missing flags with defaults and applies the legacy LayerNorm key fix-up (``:345-353``).
The pickles reference ``torchtext.vocab.Vocab`` and ``onmt.utils.optimizers.*`` classes; this
module maps them onto local stand-ins at unpickle time so the file loads byte-for-byte unchanged.
"""
from __future__ import annotations

import io
import pickle
import re
import sys
import types
from collections import Counter, defaultdict
from typing import Dict, List, Tuple

import torch

from .config import ModelConfig, SPECIALS


class Vocab(object):
    """Stand-in for legacy ``torchtext.vocab.Vocab`` (attributes itos / stoi / freqs)."""

    def __init__(self, itos: List[str] = None, freqs: Counter = None):
        self.itos = list(itos or [])
        self.freqs = freqs if freqs is not None else Counter()
        self.stoi = defaultdict(lambda: 0, {t: i for i, t in enumerate(self.itos)})

    def __len__(self):
        return len(self.itos)

    # same pickle protocol as the reference patches in (inputters/inputter.py:25-35)
    def __getstate__(self):
        return dict(self.__dict__, stoi=dict(self.stoi))

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.stoi = defaultdict(lambda: 0, self.stoi)


class _Opaque(object):
    """Placeholder for pickled training-only objects (the optimizer)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


def _install_pickle_aliases():
    """Make ``torchtext.vocab.Vocab`` resolvable for pickling/unpickling when torchtext is absent."""
    if "torchtext.vocab" in sys.modules and hasattr(sys.modules["torchtext.vocab"], "Vocab"):
        return sys.modules["torchtext.vocab"].Vocab
    tt = sys.modules.setdefault("torchtext", types.ModuleType("torchtext"))
    tv = types.ModuleType("torchtext.vocab")
    tv.Vocab = Vocab
    tt.vocab = tv
    sys.modules["torchtext.vocab"] = tv
    Vocab.__module__ = "torchtext.vocab"
    return Vocab


def make_vocab_entry(itos: List[str]) -> List[Tuple[str, object]]:
    """``checkpoint['vocab']``: list of (field name, Vocab) (inputters/inputter.py:180-190)."""
    cls = _install_pickle_aliases()
    bases = [t for t in itos if t not in SPECIALS]
    freqs = Counter({t: len(bases) - i for i, t in enumerate(bases)})
    if cls is Vocab:
        v = Vocab(itos, freqs)
    else:                                   # a real (or harness-provided) torchtext Vocab
        v = cls(freqs, specials=list(SPECIALS))
        assert list(v.itos) == list(itos), (v.itos, itos)
    return [("tgt", v)]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("torchtext") and name == "Vocab":
            return Vocab
        if module.startswith("onmt.") or module.startswith("torchtext"):
            try:
                return super().find_class(module, name)
            except Exception:
                return _Opaque
        return super().find_class(module, name)


class _PickleModule(object):
    """pickle-module shim for ``torch.load(pickle_module=...)``."""
    __name__ = "pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load())
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    DEFAULT_PROTOCOL = pickle.DEFAULT_PROTOCOL
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def _fix_key(s: str) -> str:
    # models/model_builder.py:345-353
    s = re.sub(r"(.*)\.layer_norm((_\d+)?)\.b_2", r"\1.layer_norm\2.bias", s)
    s = re.sub(r"(.*)\.layer_norm((_\d+)?)\.a_2", r"\1.layer_norm\2.weight", s)
    return s


def save_checkpoint(ckpt: dict, path: str) -> None:
    _install_pickle_aliases()
    torch.save(ckpt, path)


def load_checkpoint(path_or_dict) -> Tuple[ModelConfig, Dict[str, torch.Tensor], Vocab]:
    """-> (config, flat fp32 state dict incl. ``generator.0.*``, target Vocab)."""
    if isinstance(path_or_dict, dict):
        ckpt = path_or_dict
    else:
        ckpt = torch.load(path_or_dict, map_location="cpu", weights_only=False,
                          pickle_module=_PickleModule)
    for k in ("model", "generator", "vocab", "opt"):
        if k not in ckpt:
            raise KeyError("checkpoint is missing %r (not a NanoDecoder/OpenNMT checkpoint)" % k)
    vocab = dict(ckpt["vocab"])["tgt"]
    cfg = ModelConfig.from_opt(ckpt["opt"], list(vocab.itos))
    sd = {}
    for k, v in ckpt["model"].items():
        if "generator" in k:                              # onmt/models/model_saver.py:106-107
            continue
        sd[_fix_key(k)] = v
    for k, v in ckpt["generator"].items():
        sd["generator." + k] = v
    sd = {k: (v.detach().float().contiguous() if v.is_floating_point() else v)
          for k, v in sd.items() if not k.endswith(".mask")}
    return cfg, sd, vocab
