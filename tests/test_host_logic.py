"""CPU: host-side logic of the product package (no CUDA calls)."""
import argparse
import ctypes
import os
import re

import pytest
import torch

import nanodecoder_b200 as pkg
from nanodecoder_b200 import _lib, checkpoint, synth
from nanodecoder_b200.config import FAMILIES, ModelConfig
from nanodecoder_b200.opts import default_translate_opt, translate_opts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_checkpoint_roundtrip_keeps_reference_layout(tmp_path):
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=2, dec_layers=2)
    ck = synth.make_checkpoint(cfg)
    assert set(ck) == {"model", "generator", "vocab", "opt", "optim"}          # model_saver.py:109-115
    assert not any(k.startswith("generator") for k in ck["model"])
    assert set(ck["generator"]) == {"0.weight", "0.bias"}
    path = str(tmp_path / "m.pt")
    checkpoint.save_checkpoint(ck, path)
    cfg2, sd2, vocab = checkpoint.load_checkpoint(path)
    assert cfg2.asdict() == cfg.asdict()
    assert vocab.itos == ["<unk>", "<blank>", "<s>", "</s>", "A", "C", "G", "T"]
    assert vocab.stoi["</s>"] == 3 and vocab.stoi["never-seen"] == 0
    sd = synth.make_state_dict(cfg)
    assert set(sd2) == {k for k in sd if sd[k].is_floating_point() or True} - set()
    for k in sd:
        assert torch.equal(sd2[k], sd[k]) if sd[k].is_floating_point() else True


def test_legacy_layernorm_keys_are_fixed():
    cfg = ModelConfig.family("t2t", d_model=32, d_ff=64, enc_layers=1, dec_layers=1, heads=4)
    ck = synth.make_checkpoint(cfg)
    m = ck["model"]
    m["encoder.layer_norm.a_2"] = m.pop("encoder.layer_norm.weight")
    m["encoder.layer_norm.b_2"] = m.pop("encoder.layer_norm.bias")
    _, sd, _ = checkpoint.load_checkpoint(ck)
    assert "encoder.layer_norm.weight" in sd and "encoder.layer_norm.a_2" not in sd


@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_synthetic_state_dicts_are_seed_deterministic(family):
    cfg = ModelConfig.family(family, d_model=32, d_ff=64, enc_layers=2, dec_layers=2, heads=4)
    a, b = synth.make_state_dict(cfg, seed=7), synth.make_state_dict(cfg, seed=7)
    c = synth.make_state_dict(cfg, seed=8)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert any(not torch.equal(a[k], c[k]) for k in a)


def test_unsupported_configs_fail_loudly():
    with pytest.raises(ValueError):
        ModelConfig(encoder_type="mean")
    with pytest.raises(ValueError):
        ModelConfig(encoder_type="resnet", decoder_type="cnn")          # [d,B,T] outputs + [1,1,B,T] embedding
    with pytest.raises(ValueError):
        ModelConfig(rnn_type="SRU")
    assert ModelConfig(rnn_type="GRU").rnn_type == "GRU"
    opt = ModelConfig.family("l2t").to_opt()
    opt.copy_attn = True
    with pytest.raises(ValueError):
        ModelConfig.from_opt(opt, ["<unk>", "<blank>", "<s>", "</s>", "A", "C", "G", "T"])


def test_translate_flags_match_reference_names():
    p = argparse.ArgumentParser()
    translate_opts(p)
    opt = p.parse_args("-model m.pt -src_dir d -save_data s --src_seq_length 300 --src_seq_stride 60 --fast "
                       "--beam_size 5 --max_length 100 --batch_size 800 --thread 10 -gpu 0".split())
    assert opt.models == ["m.pt"] and opt.fast and opt.beam_size == 5 and opt.src_seq_stride == 60
    d = default_translate_opt()
    assert (d.max_length, d.beam_size, d.batch_size, d.n_best, d.alpha) == (100, 5, 100, 1, 0.0)


def test_cabi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "nanodec.h")).read()
    declared = set(re.findall(r"ND_EXPORT\s+[\w\s\*]+?\b(nd_\w+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        from nanodecoder_b200.build import build
        build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.nd_api_version.restype = ctypes.c_int
    assert lib.nd_api_version() == _lib.ND_API_VERSION


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from nanodecoder_b200.engine import Engine
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=1, dec_layers=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(cfg, synth.make_state_dict(cfg), max_batch=2)


def test_product_package_does_not_import_the_oracle():
    pkg_dir = os.path.dirname(pkg.__file__)
    for root, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_every_documented_option_is_handled_and_vice_versa():
    """The integer options listed above nd_set_int in include/nanodec.h are exactly the names engine.cu compares
    against (an undocumented switch, or a documented one that silently does nothing, is an ABI bug)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "nanodec.h")).read()
    block = header[header.index("integer options."): header.index("nd_set_int(")]
    documented = set(re.findall(r'^\s*\*\s+"([a-z_0-9]+)"', block, flags=re.M))
    src = open(os.path.join(root, "nanodecoder_b200", "csrc", "engine.cu")).read()
    body = src[src.index("int nd_set_int("):]
    body = body[: body.index("\n}\n")]
    handled = set(re.findall(r'strcmp\(name, "([a-z_0-9]+)"\)', body))
    assert documented == handled, (sorted(documented - handled), sorted(handled - documented))


def test_library_sass_is_blackwell_native():
    """cuobjdump -sass of the built library (no GPU needed): the GEMMs, the LSTM and the encoder attention issue
    tcgen05.mma (UTC*MMA) with tensor-memory loads / stores (LDTM / STTM), operands arrive by TMA (UTMALDG), the ring
    cross attention is TMA-fed with packed fp32 FMAs, and no legacy tensor-core path (HMMA) is compiled in."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import sass_inventory as si
    inv = si.inventory()
    names = si.demangle(list(inv))
    by = {}
    for n, c in zip(names, inv.values()):
        for prefix in ("gemm_tc_kernel", "gemm_tc_persist_kernel", "lstm_tc_kernel", "enc_attn_tc_kernel",
                       "enc_attn_tc64_kernel", "cross_attn_ring_kernel"):
            if n.startswith(prefix + "<") or n == prefix:
                by.setdefault(prefix, []).append(c)
    for k in ("gemm_tc_kernel", "gemm_tc_persist_kernel"):
        assert by[k] and all(c["UTC*MMA"] and c["UTMALDG"] and c["LDTM"] for c in by[k]), k
    assert any(c["STTM"] for c in by["gemm_tc_persist_kernel"])           # A operand written to tensor memory
    assert all(c["UTC*MMA"] and c["LDTM"] and c["STAS"] for c in by["lstm_tc_kernel"])
    assert all(c["UTC*MMA"] and c["LDTM"] and c["STTM"] for c in by["enc_attn_tc_kernel"] + by["enc_attn_tc64_kernel"])
    assert all(c["UTMALDG"] and c["FFMA2"] for c in by["cross_attn_ring_kernel"])
    assert sum(c["HMMA"] for c in inv.values()) == 0


def test_model_flags_that_change_the_arithmetic_are_refused():
    """ModelConfig.from_opt (models/model_builder.py:65-214): a checkpoint trained with a module the engine does not run
    must fail at load, not decode to other bases."""
    import pytest
    from nanodecoder_b200.config import ModelConfig, SPECIALS
    vocab = SPECIALS + ["A", "C", "G", "T"]
    base = ModelConfig.family("brnn2rnn").to_opt()
    assert ModelConfig.from_opt(base, vocab).encoder_type == "brnn"
    opt = ModelConfig.family("brnn2rnn").to_opt()
    opt.bridge = True
    assert ModelConfig.from_opt(opt, vocab).bridge                       # -bridge is run (rnn / brnn encoders)
    opt = ModelConfig.family("l2t").to_opt()
    opt.bridge = True
    assert not ModelConfig.from_opt(opt, vocab).bridge                   # model_builder.py: the other encoders never see it
    for flag, value in (("global_attention_function", "sparsemax"), ("generator_function", "sparsemax"),
                        ("copy_attn", True), ("context_gate", "both"), ("self_attn_type", "sideways")):
        opt = ModelConfig.family("brnn2rnn").to_opt()
        setattr(opt, flag, value)
        with pytest.raises(ValueError):
            ModelConfig.from_opt(opt, vocab)
    with pytest.raises(ValueError):
        ModelConfig.from_opt(base, ["<unk>", "<s>", "<blank>", "</s>", "A", "C", "G", "T"])     # specials out of order


def test_nd_config_layout_matches_the_header(tmp_path):
    """The ctypes mirror of nd_config (nanodecoder_b200/_lib.py) against the C struct of include/nanodec.h, field by
    field, as a C compiler lays it out: a field added to one side only would shift every later field silently."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    header = open(os.path.join(ROOT, "include", "nanodec.h")).read()
    body = header[header.index("typedef struct nd_config {"): header.index("} nd_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"int32_t\s+(\w+)\s*(?:\[\d+\])?\s*;", body)
    assert fields == [f[0] for f in _lib.NdConfig._fields_], (fields, [f[0] for f in _lib.NdConfig._fields_])
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nanodec.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(nd_config));\n' +
                   "".join('  printf("%%zu\\n", offsetof(nd_config, %s));\n' % f for f in fields) + "  return 0;\n}\n")
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    out = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.NdConfig)
    assert out[1:] == [getattr(_lib.NdConfig, f).offset for f in fields]


def test_ctypes_signatures_have_the_header_parameter_counts():
    """Every prototype of include/nanodec.h against the ctypes argtypes in _lib.SIGNATURES: same number of parameters
    (a parameter added to one side only corrupts the call frame without any error)."""
    header = open(os.path.join(ROOT, "include", "nanodec.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = dict(re.findall(r"ND_EXPORT\s+[\w\s\*]+?\b(nd_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
