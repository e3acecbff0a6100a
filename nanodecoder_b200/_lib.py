"""ctypes binding of libnanodec.so (C ABI in include/nanodec.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded the import
of the product path fails with an explicit error (build it with ``python -m nanodecoder_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnanodec.so")

ND_API_VERSION = 1
ND_OK, ND_ERR_INVALID, ND_ERR_CUDA, ND_ERR_STATE, ND_ERR_WEIGHT, ND_ERR_NOMEM = 0, -1, -2, -3, -4, -5
RNN = {"LSTM": 0, "GRU": 1}
ENC = {"nano": 0, "transformer": 1, "cnn": 2, "rnn": 3, "brnn": 4, "resnet": 5, "crnn": 6, "ctransformer": 7}
DEC = {"transformer": 0, "rnn": 1, "cnn": 2}
ATTN = {"mlp": 0, "general": 1, "dot": 2}
GEMM = {"simt": 0, "3xtf32": 1, "tf32": 2}
NORM = {"median": 0, "mean": 1, "none": 2, "None": 2}
DTYPE_F32, DTYPE_I64 = 0, 1
PROF_CATS = ["gemm", "lstm", "cross_attn", "self_attn", "enc_attn", "mlp_attn", "generator", "beam", "other"]


class NdConfig(C.Structure):
    _fields_ = [
        ("api_version", C.c_int32), ("device", C.c_int32), ("encoder_type", C.c_int32),
        ("decoder_type", C.c_int32), ("enc_layers", C.c_int32), ("dec_layers", C.c_int32),
        ("d_model", C.c_int32), ("heads", C.c_int32), ("d_ff", C.c_int32), ("vocab_size", C.c_int32),
        ("cnn_kernel_width", C.c_int32), ("enc_pooling", C.c_int32 * 8), ("input_feed", C.c_int32),
        ("attn_type", C.c_int32), ("position_encoding", C.c_int32), ("max_batch", C.c_int32),
        ("max_src_len", C.c_int32), ("max_tgt_len", C.c_int32), ("max_beam", C.c_int32),
        ("gemm_mode", C.c_int32), ("rnn_type", C.c_int32), ("bridge", C.c_int32), ("self_attn_average", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class NanodecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libnanodec error %d: %s" % (code, msg))
        self.code = code


# every symbol include/nanodec.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "nd_api_version": (C.c_int, []),
    "nd_longest_match": (C.c_int, [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int32)]),
    "nd_assembly_offsets": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32)]),
    "nd_parse_signal_text": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_int16), C.c_int64, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int32)]),
    "nd_fast5_read_signal": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_int16), C.c_int64, C.POINTER(C.c_int64),
                                       C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]),
    "nd_fast5_list_reads": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_char_p, C.c_int32]),
    "nd_fast5_read_signal_of": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.POINTER(C.c_int16), C.c_int64,
                                          C.POINTER(C.c_int64), C.c_char_p, C.c_int32]),
    "nd_h5_list_group": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32), C.c_char_p, C.c_int32]),
    "nd_h5_read_dataset": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.POINTER(C.c_uint8), C.c_int64,
                                     C.POINTER(C.c_int64), C.c_char_p, C.c_int32]),
    "nd_zstd_decompress": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int64),
                                     C.c_char_p, C.c_int32]),
    "nd_simple_assembly": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int8),
                                     C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int64)]),
    "nd_create": (C.c_int, [C.POINTER(NdConfig), C.POINTER(_P)]),
    "nd_destroy": (C.c_int, [_P]),
    "nd_last_error": (C.c_char_p, [_P]),
    "nd_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32, C.c_int32]),
    "nd_finalize_weights": (C.c_int, [_P]),
    "nd_frontend_stats": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "nd_frontend_chunks": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "nd_frontend_stats_f64": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "nd_frontend_chunks_f64": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "nd_encode": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "nd_get_memory_bank": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int32), _P]),
    "nd_decode_greedy": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "nd_decode_beam": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P, _P]),
    "nd_decode_beam_object": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P,
                                        _P, _P]),
    "nd_beam_attention": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P]),
    "nd_set_int": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "nd_set_float": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "nd_profile_enable": (C.c_int, [_P, C.c_uint32]),
    "nd_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "nd_launch_count": (C.c_int64, [_P]),
    "nd_reset_launch_count": (C.c_int, [_P]),
    "nd_debug_gemm_timeline": (C.c_int, [_P]),
    "nd_test_gemm": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, _P]),
}

_lib = None


def load():
    """Load libnanodec.so (after torch, so both share one libcudart / primary context)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "nanodecoder_b200: %s not found. The translate path has no CPU/PyTorch fallback; build the "
            "CUDA extension first:  python -m nanodecoder_b200.build" % LIB_PATH)
    import torch  # noqa: F401  (loads libcudart.so.12 that libnanodec links against)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.nd_api_version() != ND_API_VERSION:
        raise ImportError("libnanodec.so API version %d != binding %d" % (lib.nd_api_version(), ND_API_VERSION))
    _lib = lib
    return lib
