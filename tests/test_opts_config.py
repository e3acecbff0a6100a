"""CPU: `-config FILE.yml` / `-save_config FILE.yml` of the CLI (models/opts.py:8-13 + translate.py:177-185 of the reference:
configargparse with the YAML config-file parser; here argparse plus nanodecoder_b200/opts.py::parse_args)."""
import argparse

import pytest
import yaml

from nanodecoder_b200 import opts


def _parser():
    p = argparse.ArgumentParser(description="translate.py")
    opts.config_opts(p)
    opts.translate_opts(p)
    return p


def test_plain_command_line_is_untouched():
    a = opts.parse_args(_parser(), ["-model", "m.pt", "-save_data", "out", "-batch_size", "7", "--fast"])
    b = _parser().parse_args(["-model", "m.pt", "-save_data", "out", "-batch_size", "7", "--fast"])
    assert vars(a) == vars(b) and a.config is None and a.save_config is None


def test_config_file_sits_between_defaults_and_command_line(tmp_path):
    cfg = tmp_path / "translate.yml"
    cfg.write_text(yaml.safe_dump({"model": ["a.pt"], "save_data": "from_cfg", "batch_size": 800, "beam_size": 3,
                                   "fast": True, "attn_debug": "false", "ignore_when_blocking": ["A", "C"],
                                   "length_penalty": "wu", "alpha": 0.7, "src_dir": "reads"}))
    opt = opts.parse_args(_parser(), ["-config", str(cfg), "-beam_size", "5", "-save_data", "from_cli"])
    assert opt.models == ["a.pt"] and opt.src_dir == "reads"          # required -model satisfied by the file
    assert opt.batch_size == 800 and opt.fast is True and opt.attn_debug is False
    assert opt.ignore_when_blocking == ["A", "C"] and opt.length_penalty == "wu" and opt.alpha == 0.7
    assert opt.beam_size == 5 and opt.save_data == "from_cli"         # the command line wins
    assert opt.max_length == 100                                      # untouched default
    # the long spelling and `=` work too
    opt = opts.parse_args(_parser(), ["--config=%s" % cfg])
    assert opt.beam_size == 3 and opt.save_data == "from_cfg"


def test_config_file_errors(tmp_path, capsys):
    cfg = tmp_path / "bad.yml"
    cfg.write_text("no_such_flag: 3\nmodel: [m.pt]\nsave_data: o\n")
    with pytest.raises(SystemExit):
        opts.parse_args(_parser(), ["-config", str(cfg)])
    assert "unrecognized arguments: --no_such_flag=3" in capsys.readouterr().err
    cfg.write_text("- just\n- a list\n")
    with pytest.raises(SystemExit):
        opts.parse_args(_parser(), ["-config", str(cfg), "-model", "m", "-save_data", "o"])
    assert "YAML mapping" in capsys.readouterr().err
    cfg.write_text("fast: maybe\nmodel: [m.pt]\nsave_data: o\n")
    with pytest.raises(SystemExit):
        opts.parse_args(_parser(), ["-config", str(cfg)])
    assert "Unexpected value for fast" in capsys.readouterr().err
    cfg.write_text("beam_size: [1, 2]\nmodel: [m.pt]\nsave_data: o\n")
    with pytest.raises(SystemExit):
        opts.parse_args(_parser(), ["-config", str(cfg)])
    assert "can't be set to a list" in capsys.readouterr().err
    with pytest.raises(SystemExit):
        opts.parse_args(_parser(), ["-config", str(tmp_path / "missing.yml"), "-model", "m", "-save_data", "o"])
    assert "Unable to open config file" in capsys.readouterr().err


def test_save_config_writes_the_options_and_exits(tmp_path, capsys):
    out = tmp_path / "saved.yml"
    with pytest.raises(SystemExit) as e:
        opts.parse_args(_parser(), ["-model", "m.pt", "-save_data", "o", "-batch_size", "1024", "--fast", "-kv_mode", "q15",
                                    "-save_config", str(out)])
    assert e.value.code == 0 and "Wrote config file to" in capsys.readouterr().err
    saved = yaml.safe_load(out.read_text())
    assert saved["model"] == ["m.pt"] and saved["batch_size"] == 1024 and saved["fast"] is True and saved["kv_mode"] == "q15"
    assert "config" not in saved and "save_config" not in saved
    # the saved file is a config file: it reproduces the options
    a = opts.parse_args(_parser(), ["-config", str(out)])
    b = opts.parse_args(_parser(), ["-model", "m.pt", "-save_data", "o", "-batch_size", "1024", "--fast", "-kv_mode", "q15"])
    va, vb = vars(a), vars(b)
    va.pop("config"), vb.pop("config")
    # -fft is declared `type=bool` as in the reference (models/opts.py): bool("False") is True, from a config file as from
    # the command line
    assert va.pop("fft") is True and vb.pop("fft") is False
    assert va == vb
