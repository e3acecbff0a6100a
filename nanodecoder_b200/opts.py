"""Command-line flags of the translate path, same names / defaults as the reference
(models/opts.py:504-658 ``translate_opts``; configargparse is replaced by argparse)."""
from __future__ import annotations

import argparse


def translate_opts(parser: argparse.ArgumentParser) -> None:
    g = parser.add_argument_group("Model")
    g.add_argument("--model", "-model", dest="models", metavar="MODEL", nargs="+", type=str, default=[],
                   required=True, help="Path to model .pt file (ensembles are not supported)")
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
