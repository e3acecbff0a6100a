"""Command-line flags of the translate path, same names / defaults as the reference
(models/opts.py:504-658 ``translate_opts``; configargparse is replaced by argparse)."""
from __future__ import annotations

import argparse
import sys


def config_opts(parser: argparse.ArgumentParser) -> None:
    """models/opts.py:8-13: `-config FILE` (YAML) and `-save_config FILE`, configargparse's two special arguments."""
    g = parser.add_argument_group("Config")
    g.add_argument("-config", "--config", required=False, help="config file path (YAML mapping: option name -> value)")
    g.add_argument("-save_config", "--save_config", required=False, help="config file save path")


def _config_file_args(parser: argparse.ArgumentParser, path: str):
    """The command-line arguments a YAML config file stands for, as configargparse's YAMLConfigFileParser +
    `convert_item_to_command_line_arg` build them: keys are option names without their dashes; flags take true / false;
    lists feed `nargs` options; everything else becomes `--key=value`.  Unknown keys stay in as `--key=value`, so that
    argparse reports them as unrecognised arguments, like the reference's parser does."""
    import yaml
    try:
        with open(path) as f:
            items = yaml.safe_load(f)
    except OSError as e:
        parser.error("Unable to open config file: %s. Error: %s" % (path, e))
    except yaml.YAMLError as e:
        parser.error("Couldn't parse config file: %s" % e)
    if items is None:
        items = {}
    if not isinstance(items, dict):
        parser.error("The config file doesn't appear to contain 'key: value' pairs (aka. a YAML mapping). "
                     "yaml.load('%s') returned type '%s' instead of 'dict'." % (path, type(items).__name__))
    known = {}
    for action in parser._actions:
        for opt in action.option_strings:
            known.setdefault(opt.lstrip("-"), action)
    args = []
    for key, value in items.items():
        action = known.get(str(key))
        if action is None:
            args.append("--%s=%s" % (key, value))
            continue
        opt = action.option_strings[0]
        if action.nargs == 0:                                   # store_true / store_false / store_const flags
            text = str(value).lower()
            if text in ("true", "yes", "1"):
                args.append(opt)
            elif text not in ("false", "no", "0"):
                parser.error("Unexpected value for %s: '%s'. Expecting 'true', 'false', 'yes', 'no', '1' or '0'"
                             % (key, value))
        elif isinstance(value, list):
            if not value:
                continue
            if action.nargs in ("+", "*") or isinstance(action.nargs, int):
                args += [opt] + [str(v) for v in value]
            else:
                parser.error("%s can't be set to a list '%s' unless its action type is changed to 'append' or nargs "
                             "is set to '*', '+', or > 1" % (key, value))
        else:
            args.append("%s=%s" % (opt, value))
    return args


def parse_args(parser: argparse.ArgumentParser, argv=None) -> argparse.Namespace:
    """parser.parse_args with configargparse's config-file semantics (translate.py:177-185 of the reference): values of
    `-config FILE` sit between the defaults and the command line (the command line wins); `-save_config FILE` writes the
    resulting options as a YAML config file and exits."""
    argv = list(sys.argv[1:] if argv is None else argv)
    pre = argparse.ArgumentParser(add_help=False, allow_abbrev=False)
    pre.add_argument("-config", "--config")
    pre.add_argument("-save_config", "--save_config")
    special, _ = pre.parse_known_args(argv)
    if special.config is None and special.save_config is None:
        return parser.parse_args(argv)
    if special.config is not None:
        argv = _config_file_args(parser, special.config) + argv
    opt = parser.parse_args(argv)
    if special.save_config is not None:
        import yaml
        items = {}
        for action in parser._actions:
            if not action.option_strings or action.dest in ("help", "config", "save_config"):
                continue
            value = getattr(opt, action.dest, None)
            if value is None or value == []:                   # nothing to write for an unset or empty list option
                continue
            long_opts = [o for o in action.option_strings if o.startswith("--")] or action.option_strings
            items[long_opts[0].lstrip("-")] = value
        try:
            with open(special.save_config, "w") as f:
                yaml.safe_dump(items, f, default_flow_style=False, sort_keys=False)
        except OSError as e:
            parser.error("Couldn't open %s for writing: %s" % (special.save_config, e))
        parser.exit(0, "Wrote config file to %s\n" % special.save_config)
    return opt


class _LoggingLevel(argparse.Action):
    """-log_file_level NAME | NUMBER -> the logging level (models/opts.py:730-751)"""
    import logging
    LEVELS = {"CRITICAL": logging.CRITICAL, "ERROR": logging.ERROR, "WARNING": logging.WARNING, "INFO": logging.INFO,
              "DEBUG": logging.DEBUG, "NOTSET": logging.NOTSET}
    CHOICES = list(LEVELS.keys()) + [str(v) for v in LEVELS.values()]

    def __call__(self, parser, namespace, value, option_string=None):
        setattr(namespace, self.dest, self.LEVELS.get(value, value))


class _Deprecated(argparse.Action):
    """a flag that only says what replaced it (models/opts.py:754-765)"""

    def __init__(self, option_strings, dest, help=None, **kwargs):
        super().__init__(option_strings, dest, nargs=0, help=help, **kwargs)

    def __call__(self, parser, namespace, values, flag_name=None):
        parser.error("Flag '%s' is deprecated. %s" % (flag_name, self.help or ""))


def translate_opts(parser: argparse.ArgumentParser) -> None:
    g = parser.add_argument_group("Model")
    g.add_argument("--model", "-model", dest="models", metavar="MODEL", nargs="+", type=str, default=[],
                   required=True, help="Path to model .pt file (ensembles are not supported)")
    g.add_argument("--avg_raw_probs", "-avg_raw_probs", action="store_true",
                   help="ensemble decoding only: average raw probabilities instead of log-probabilities (accepted, "
                        "without effect: a single model is decoded)")
    g = parser.add_argument_group("Data")
    g.add_argument("--thread", "-thread", type=int, default=4, help="reader threads")
    g.add_argument("--normalization_raw", default="median", help="median | mean | None")
    g.add_argument("--src_dir", "-src_dir", default="", help="directory of .fast5 / .signal reads")
    g.add_argument("--src_seq_length", "-src_seq_length", type=int, default=512)
    g.add_argument("--src_seq_stride", "-src_seq_stride", type=int, default=512)
    g.add_argument("--save_data", "-save_data", required=True, help="output folder")
    g = parser.add_argument_group("Random Sampling")
    g.add_argument("--random_sampling_topk", "-random_sampling_topk", default=1, type=int)
    g.add_argument("--random_sampling_temp", "-random_sampling_temp", default=1.0, type=float)
    g = parser.add_argument_group("Beam")
    g.add_argument("--fast", "-fast", action="store_true", help="batched beam search")
    g.add_argument("--beam_size", "-beam_size", type=int, default=5)
    g.add_argument("--min_length", "-min_length", type=int, default=0)
    g.add_argument("--max_length", "-max_length", type=int, default=100)
    g.add_argument("--max_sent_length", "-max_sent_length", action=_Deprecated,
                   help="Deprecated, use `-max_length` instead")
    g.add_argument("--stepwise_penalty", "-stepwise_penalty", action="store_true")
    g.add_argument("--length_penalty", "-length_penalty", default="none", choices=["none", "wu", "avg"])
    g.add_argument("--coverage_penalty", "-coverage_penalty", default="none", choices=["none", "wu", "summary"])
    g.add_argument("--alpha", "-alpha", type=float, default=0.0)
    g.add_argument("--beta", "-beta", type=float, default=-0.0)
    g.add_argument("--block_ngram_repeat", "-block_ngram_repeat", type=int, default=0)
    g.add_argument("--ignore_when_blocking", "-ignore_when_blocking", nargs="+", type=str, default=[])
    g.add_argument("--replace_unk", "-replace_unk", action="store_true")
    g = parser.add_argument_group("Logging")
    g.add_argument("--verbose", "-verbose", action="store_true")
    g.add_argument("--log_file", "-log_file", type=str, default="")
    g.add_argument("--log_file_level", "-log_file_level", type=str, action=_LoggingLevel, choices=_LoggingLevel.CHOICES,
                   default="0")
    g.add_argument("--attn_debug", "-attn_debug", action="store_true")
    g.add_argument("--dump_beam", "-dump_beam", type=str, default="")
    g.add_argument("--n_best", "-n_best", type=int, default=1)
    g = parser.add_argument_group("Efficiency")
    g.add_argument("--batch_size", "-batch_size", type=int, default=100)
    g.add_argument("--gpu", "-gpu", type=int, default=0, help="CUDA device (this engine has no CPU path)")
    g.add_argument("--gemm_mode", "-gemm_mode", default="3xtf32", choices=["3xtf32", "tf32", "simt"],
                   help="arithmetic of the dense projections: 3xtf32 = fp32-parity tcgen05, tf32 = fast")
    g.add_argument("--kv_mode", "-kv_mode", default="q23", choices=["f32", "q23", "q15"],
                   help="storage of what the decoder's attention re-reads at every step (Transformer: memory keys / "
                        "values; RNN: uh | H; CNN: encoder top | combined state): q23 = 3-byte fixed point (parity mode, "
                        "3/4 of the bytes), f32, q15 = reduced precision (half the bytes; the CNN decoder keeps q23)")
    g = parser.add_argument_group("SpeechLike")
    g.add_argument("--fft", "-fft", type=bool, default=False)
    g.add_argument("--sample_rate", "-sample_rate", type=int, default=4000)
    g.add_argument("--window_size", "-window_size", type=float, default=0.075)
    g.add_argument("--window_stride", "-window_stride", type=float, default=0.015)
    g.add_argument("--window", "-window", default="hamming")


def default_translate_opt(**kw) -> argparse.Namespace:
    p = argparse.ArgumentParser()
    translate_opts(p)
    opt = p.parse_args(["-model", "m", "-save_data", "s"])
    opt.data_type = "nano"
    opt.tgt = None
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt
