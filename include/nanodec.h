/*
 * nanodec.h — C ABI of the B200-native NanoDecoder translate engine (libnanodec.so).
 *
 * Drop-in boundary for the reference's translate path (achilles1989/NanoDecoder).  The reference
 * is pure Python/PyTorch and has NO native interface of its own, so each entry point cites the
 * reference function whose arithmetic it replaces (paths relative to the reference root):
 *
 *   nd_frontend_*        utils/labelop.py:194-243 (extract_fast5_raw: normalise + chunk) and
 *                        inputters/nano_dataset.py:49-58,81 + inputters/inputter.py:86-95
 *                        (text -> float -> zero-padded fp32 batch)
 *   nd_encode            translate/translator.py:542-559 (_run_encoder) ->
 *                        encoder/nano_encoder.py:79-124, encoder/transformer.py:106-127,
 *                        encoder/cnn_encoder.py:29-44, encoder/rnn_encoder.py:64-84
 *   nd_decode_greedy     translate/translator.py:396-503 (_translate_random_sampling, topk 1) with
 *                        decoder/transformer.py:194-246, onmt/decoders/decoder.py:303-366,
 *                        onmt/decoders/cnn_decoder.py:74-132, models/model_builder.py:331-334
 *   nd_decode_beam       translate/translator.py:619-825 (_fast_translate_batch)
 *   nd_decode_beam_object translate/translator.py:827-926 (_translate_batch) + onmt/translate/beam.py:74-178
 *   nd_load_weight       models/model_builder.py:343-357 (load_state_dict of checkpoint tensors)
 *
 * Conventions
 *   - Every call returns 0 on success or a negative nd_status; nd_last_error() gives the message.
 *     No C++ exception crosses the ABI.  CUDA errors are sticky per engine.
 *   - The caller owns every input/output buffer (plain device pointers unless stated otherwise).
 *     The engine owns only its weights copy, workspace and KV caches, sized at nd_create.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and no call
 *     synchronises the host unless documented.
 *   - Thread-compatible, not re-entrant: one call at a time per engine (the reference drives the
 *     translator from one result-handler thread, translate.py:100-129).
 *   - There is no CPU fallback: without a CUDA device nd_create fails with ND_ERR_CUDA.
 */
#ifndef NANODEC_H_
#define NANODEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ND_API_VERSION 1
#if defined(__GNUC__)
#define ND_EXPORT __attribute__((visibility("default")))
#else
#define ND_EXPORT
#endif

typedef enum {
  ND_OK = 0,
  ND_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  ND_ERR_CUDA = -2,        /* CUDA runtime or driver error (sticky) */
  ND_ERR_STATE = -3,       /* call order violated (e.g. decode before encode) */
  ND_ERR_WEIGHT = -4,      /* missing / mis-shaped checkpoint tensor */
  ND_ERR_NOMEM = -5
} nd_status;

enum { ND_ENC_NANO = 0, ND_ENC_TRANSFORMER = 1, ND_ENC_CNN = 2, ND_ENC_RNN = 3, ND_ENC_BRNN = 4,
       /* ResNet stem (encoder/resnet_encoder.py) alone, before the nano stack, before transformer layers */
       ND_ENC_RESNET = 5, ND_ENC_CRNN = 6, ND_ENC_CTRANSFORMER = 7 };
enum { ND_DEC_TRANSFORMER = 0, ND_DEC_RNN = 1, ND_DEC_CNN = 2 };
enum { ND_ATTN_MLP = 0, ND_ATTN_GENERAL = 1, ND_ATTN_DOT = 2 };
enum { ND_RNN_LSTM = 0, ND_RNN_GRU = 1 };        /* onmt/utils/rnn_factory.py:8-17, onmt/models/stacked_rnn.py */
/* arithmetic of the dense projections */
enum {
  ND_GEMM_SIMT_FP32 = 0,   /* fp32 FFMA tiles (bring-up / cross-check path) */
  ND_GEMM_TC_3XTF32 = 1,   /* tcgen05 kind::tf32, 3-pass split: fp32-parity mode */
  ND_GEMM_TC_TF32 = 2      /* tcgen05 kind::tf32 single pass: fast mode, ~1e-3 relative error */
};
enum { ND_DTYPE_F32 = 0, ND_DTYPE_I64 = 1 };
enum { ND_NORM_MEDIAN = 0, ND_NORM_MEAN = 1, ND_NORM_NONE = 2 };

typedef struct nd_config {
  int32_t api_version;        /* ND_API_VERSION */
  int32_t device;             /* CUDA device ordinal */
  int32_t encoder_type;       /* ND_ENC_* */
  int32_t decoder_type;       /* ND_DEC_* */
  int32_t enc_layers;
  int32_t dec_layers;
  int32_t d_model;            /* enc_rnn_size == dec_rnn_size == tgt_word_vec_size */
  int32_t heads;
  int32_t d_ff;
  int32_t vocab_size;         /* 4 specials + bases */
  int32_t cnn_kernel_width;
  int32_t enc_pooling[8];     /* per-layer MaxPool1d stride of the nano encoder */
  int32_t input_feed;
  int32_t attn_type;          /* ND_ATTN_* (RNN decoder) */
  int32_t position_encoding;
  int32_t max_batch;          /* chunks per nd_encode call */
  int32_t max_src_len;        /* samples per chunk (T) */
  int32_t max_tgt_len;        /* decode steps (L) */
  int32_t max_beam;           /* largest beam_size that will be requested (>=1) */
  int32_t gemm_mode;          /* ND_GEMM_* */
  int32_t rnn_type;           /* ND_RNN_*: cell of the nano / rnn / brnn encoders and of the RNN decoder (took reserved[0]) */
  int32_t bridge;             /* 1: rnn / brnn encoder with -bridge: Linear + ReLU on the final states (encoder/rnn_encoder.py:82-118;
                                 took reserved[1]) */
  int32_t self_attn_average;  /* 1: Transformer decoder with -self_attn_type average (onmt/modules/average_attn.py; took reserved[2]) */
  int32_t reserved[5];
} nd_config;

typedef struct nd_engine nd_engine;

/* lifecycle --------------------------------------------------------------------------------- */
ND_EXPORT int nd_create(const nd_config* cfg, nd_engine** out);
ND_EXPORT int nd_destroy(nd_engine* e);
ND_EXPORT const char* nd_last_error(const nd_engine* e);   /* e may be NULL: last create error */
ND_EXPORT int nd_api_version(void);

/* weights: `name` is the checkpoint key ("encoder.rnn_0.weight_ih_l0", "generator.0.bias", ...);
 * `data` may be a host or device pointer (copied synchronously). */
ND_EXPORT int nd_load_weight(nd_engine* e, const char* name, const void* data, const int64_t* shape,
                   int32_t ndim, int32_t dtype);
ND_EXPORT int nd_finalize_weights(nd_engine* e);           /* checks completeness, packs kernel layouts */

/* signal front end ---------------------------------------------------------------------------
 * signal:       int16 raw samples of n_reads reads, concatenated (device)
 * read_offsets: [n_reads+1] int64 sample offsets (device)
 * Pass 1 (nd_frontend_stats): exact per-read median and scale (MAD/0.6745 or std) in fp64.
 * Pass 2 (nd_frontend_chunks): normalise in fp64, round once to fp32, gather fixed-stride chunks
 *   chunk_read/chunk_start: [n_chunks] which read / first sample of each chunk (device; the chunk
 *   table is built by the host mirror from read lengths exactly as utils/labelop.py:225-233)
 *   out_chunks: [n_chunks, chunk_len] fp32 zero padded; out_lengths: [n_chunks] int64            */
ND_EXPORT int nd_frontend_stats(nd_engine* e, const int16_t* signal, const int64_t* read_offsets,
                      int32_t n_reads, int32_t normalization, double* out_center,
                      double* out_scale, void* stream);
ND_EXPORT int nd_frontend_chunks(nd_engine* e, const int16_t* signal, const int64_t* read_offsets,
                       const double* center, const double* scale, const int32_t* chunk_read,
                       const int64_t* chunk_start, int32_t n_chunks, int32_t chunk_len,
                       float* out_chunks, int64_t* out_lengths, void* stream);

/* the same two passes for float-valued reads: fp64 samples (a `.signal` file with non-integer tokens; the reference
 * parses every token with float(), utils/labelop.py:216-217).  Median / MAD by radix select over the doubles, std with
 * numpy's pairwise summation order (= the double np.std returns).                                  */
ND_EXPORT int nd_frontend_stats_f64(nd_engine* e, const double* signal, const int64_t* read_offsets,
                          int32_t n_reads, int32_t normalization, double* out_center,
                          double* out_scale, void* stream);
ND_EXPORT int nd_frontend_chunks_f64(nd_engine* e, const double* signal, const int64_t* read_offsets,
                           const double* center, const double* scale, const int32_t* chunk_read,
                           const int64_t* chunk_start, int32_t n_chunks, int32_t chunk_len,
                           float* out_chunks, int64_t* out_lengths, void* stream);

/* encoder ------------------------------------------------------------------------------------
 * src: [B, T] fp32 chunk-major zero padded; lengths: [B] int64 (device).                       */
ND_EXPORT int nd_encode(nd_engine* e, const float* src, const int64_t* lengths, int32_t B, int32_t T,
              void* stream);
/* copy out the memory bank of the last nd_encode as [T', B, d] (reference layout; for the CNN
 * encoder [d, B, T]) and its lengths [B]; T' is returned through out_Tp.  For parity tests.      */
ND_EXPORT int nd_get_memory_bank(nd_engine* e, float* out, int64_t* out_lengths, int32_t* out_Tp,
                       void* stream);

/* decode -------------------------------------------------------------------------------------
 * Greedy: runs exactly max_len steps like the reference (no EOS early exit).
 *   out_ids    [B, max_len] int64
 *   out_scores [B] fp32: log-prob of the LAST step's token (translator.py:494)
 *   out_attn   NULL or [max_len, B, T'] fp32 (head-0 cross attention of the last layer)
 *   out_logits NULL or [max_len, B, V] fp32 log-probs of every step (parity tests)               */
ND_EXPORT int nd_decode_greedy(nd_engine* e, int32_t max_len, int32_t min_len, int64_t* out_ids,
                     float* out_scores, float* out_attn, float* out_logits, void* stream);
/* Fast batched beam search (--fast).
 *   out_ids     [B, n_best, max_len] int64, padded with -1 after each hypothesis' last token
 *   out_lens    [B, n_best] int32 hypothesis lengths (including the final </s> if emitted)
 *   out_scores  [B, n_best] fp32                                                                 */
ND_EXPORT int nd_decode_beam(nd_engine* e, int32_t beam_size, int32_t n_best, int32_t max_len,
                   int32_t min_len, float alpha, int64_t* out_ids, int32_t* out_lens,
                   float* out_scores, void* stream);

/* Object beam search (translate/translator.py:827-926 _translate_batch + onmt/translate/beam.py:74-178, the
 * reference's default when --fast is absent): rows ending in </s> stay in the beam for one step with all their
 * children at -1e20; every chunk keeps advancing until EVERY chunk is done (top beam ended and >= n_best
 * finished) or max_len steps ran; finished hypotheses are ranked by the GNMT global score = beam score /
 * length penalty (length_penalty: 0 none, 1 wu ((5+len)^alpha / 6^alpha), 2 avg; penalties.py:65-88), stable in
 * arrival order; chunks with fewer than n_best finished hypotheses are topped up from the live beam
 * (beam.py:154-168).  Coverage penalty and n-gram blocking: integer options "coverage_penalty", "block_ngram_repeat",
 * "block_ngram_exclude", "stepwise_penalty" and the float option "beta".
 * Outputs as nd_decode_beam.                                                                                 */
ND_EXPORT int nd_decode_beam_object(nd_engine* e, int32_t beam_size, int32_t n_best, int32_t max_len,
                          int32_t min_len, int32_t length_penalty, float alpha, int64_t* out_ids,
                          int32_t* out_lens, float* out_scores, void* stream);

/* Attention of the hypotheses the last nd_decode_beam / nd_decode_beam_object returned (the reference's
 * results["attention"] under -attn_debug: translator.py:806-812 alive_attn, beam.py:135,170-178 Beam.attn / get_hyp):
 *   out [B, n_best, max_len, T'] fp32: row j of hypothesis (b, n) = head-0 cross attention of the last decoder layer at
 *   decode step j (global attention for the RNN decoder), zero for j >= the hypothesis length.
 *   out_widths NULL or [B, n_best] int32: the number of source positions the reference keeps of those rows, namely
 *   memory_lengths[i] of its TILED length vector indexed by the chunk's position i in the batch (fast mode: among the
 *   chunks not yet retired when the hypothesis finished): the length of chunk alive[i / beam_size] -- the chunk's own
 *   length when all lengths are equal, a longer one otherwise (translator.py:776, :905).
 * Needs the integer option "beam_attention" set to 1 BEFORE the decode (the per-step attention of every beam row is then kept:
 * max_tgt_len * max_batch * max_beam * max_src_len floats, allocated on first use); n_best and max_len as in that decode. */
ND_EXPORT int nd_beam_attention(nd_engine* e, int32_t n_best, int32_t max_len, float* out, int32_t* out_widths,
                                void* stream);

/* Read assembly helpers (host code, no GPU work; utils/labelop.py:320-352).
 * nd_longest_match: difflib.SequenceMatcher(None, a, b).find_longest_match(0, na, 0, nb), i.e. the longest block of
 *   get_matching_blocks() the reference's simple_assembly() selects (CPython's autojunk rule included; the sentinel
 *   {na, nb, 0} when nothing matches) -> out3 = {i, j, size}.
 * nd_assembly_offsets: chunks i = text[offsets[i] : offsets[i+1]], i < n; disp[i] = i_match - j_match between chunk
 *   i-1 and chunk i (disp[0] = 0): the displacement simple_assembly() accumulates.                              */
ND_EXPORT int nd_longest_match(const char* a, int32_t na, const char* b, int32_t nb, int32_t* out3);
ND_EXPORT int nd_assembly_offsets(const char* text, const int64_t* offsets, int32_t n, int32_t* disp);
/* nd_simple_assembly: the whole of simple_assembly() + add_count() (utils/labelop.py:311-352) for one read: chunk i is
 *   text[offsets[i] : offsets[i+1]]; counts[code][column] (caller-zeroed int32 [n_codes][cap], cap >= total bytes +
 *   2000) receives one vote per base, lut[byte] = row of that base or -1; *length = columns in use (the reference
 *   returns concensus[:, :length]).  *err: 0 ok, 1 = the reference would raise IndexError (err_args = index, matrix
 *   width: it grows by 1000 at most once per chunk), 2 = KeyError (err_args = byte, chunk index).                  */
/* nd_parse_signal_text: one `.signal` file (whitespace separated integer DAC samples = the fast5 `Signal` dataset that
 *   extract_fast5_raw reads, utils/labelop.py:199-219, as text) -> int16 samples ready for nd_frontend_stats.
 *   *status: 0 ok, 1 = some token is not a plain integer (nothing decided: use a float parser), 2 = value outside
 *   int16, 3 = more than cap samples; *count = samples written so far.                                            */
ND_EXPORT int nd_parse_signal_text(const char* text, int64_t nbytes, int16_t* out, int64_t cap, int64_t* count,
                                   int32_t* status);
/* nd_fast5_read_signal: the raw samples of a single-read .fast5 file, i.e. what the reference's
 *   h5py.File(path)['/Raw/Reads/'] -> first member (name order) -> ['Signal'].value reads (utils/labelop.py:199-214),
 *   from an in-memory copy of the file (host pointers; no GPU involved).  *count = samples in the dataset; they are
 *   written to out when cap >= *count (call with cap = 0 to size the buffer).  read_name receives the member's name
 *   ("Read_1234").  Gzip, shuffle, fletcher32 and VBZ (zstd + streamvbyte) chunks are decoded.  A file that is not
 *   HDF5, lacks the path, or uses a feature outside the subset named in csrc/fast5.cu (dense groups, libver='latest'
 *   chunk indexes) returns ND_ERR_INVALID with the reason in err (NUL terminated, cut at errcap).
 * nd_h5_read_dataset: the same reader for any fixed- or floating-point dataset at `path` ("/a/b/c"): raw little/big
 *   endian element bytes as stored.  info[8] = {type class (0 fixed, 1 float), element bytes, signed, big endian, rank,
 *   total bytes, dim 0, dim 1}; bytes are written when cap >= info[5].                                            */
ND_EXPORT int nd_fast5_read_signal(const uint8_t* file, int64_t nbytes, int16_t* out, int64_t cap, int64_t* count,
                                   char* read_name, int32_t name_cap, char* err, int32_t errcap);
/* nd_fast5_list_reads / nd_fast5_read_signal_of: the same by read name, and beyond the reference (whose reader fails on
 *   them): MULTI-read .fast5 files (/read_<uuid>/Raw/Signal, what MinKNOW and ont_fast5_api write since 2019).
 *   list: names NUL separated in `names` (written while they fit names_cap; *names_bytes = bytes needed), name order;
 *   *layout = 1 single-read (members of /Raw/Reads: the reference takes the first), 2 multi-read (root members read_*).
 *   read_signal_of: `read_name` as listed.  `file` may be a read-only memory map: only the pages of the structures
 *   visited are touched.                                                                                              */
ND_EXPORT int nd_fast5_list_reads(const uint8_t* file, int64_t nbytes, char* names, int64_t names_cap, int64_t* names_bytes,
                                  int32_t* n_reads, int32_t* layout, char* err, int32_t errcap);
ND_EXPORT int nd_fast5_read_signal_of(const uint8_t* file, int64_t nbytes, const char* read_name, int16_t* out, int64_t cap,
                                      int64_t* count, char* err, int32_t errcap);
/* nd_h5_list_group: member names of the group at `path` ("/" = root), NUL separated, in name order (h5py's iteration
 *   order); written while they fit names_cap, *names_bytes = bytes needed.  With nd_h5_read_dataset this is what
 *   nanodecoder_b200/utils/h5lite.py builds its h5py-shaped File / Group / Dataset objects from.                     */
ND_EXPORT int nd_h5_list_group(const uint8_t* file, int64_t nbytes, const char* path, char* names, int64_t names_cap,
                               int64_t* names_bytes, int32_t* n_members, char* err, int32_t errcap);
ND_EXPORT int nd_h5_read_dataset(const uint8_t* file, int64_t nbytes, const char* path, uint8_t* out, int64_t cap,
                                 int64_t* info, char* err, int32_t errcap);
/* nd_zstd_decompress: the Zstandard decoder (RFC 8878, no dictionaries) behind the VBZ filter of .fast5 chunks, exposed
 *   so that it can be checked against libzstd-written frames: src -> out (at most cap bytes), *count = bytes written. */
ND_EXPORT int nd_zstd_decompress(const uint8_t* src, int64_t nbytes, uint8_t* out, int64_t cap, int64_t* count, char* err,
                                 int32_t errcap);
ND_EXPORT int nd_simple_assembly(const char* text, const int64_t* offsets, int32_t n, const int8_t* lut,
                                 int32_t* counts, int64_t cap, int64_t* length, int32_t* err, int64_t* err_args);

/* integer options.  Scheduling only (results never depend on them):
 *   "decode_streams" (default 1, 1..16): engine-owned CUDA streams the decode loop spreads contiguous
 *                    chunk groups over (chunks are independent);
 *   "use_graphs"     (default 1): capture the decode loop of a repeated (mode, B, T, L, ...) configuration
 *                    into a CUDA graph on its second call and replay it afterwards;
 *   "pdl"            (default 2, process-wide): 2 = the tcgen05 GEMMs of the step loop are programmatic dependent
 *                    launches, 1 = every step kernel, 3 = the GEMMs and the small step kernels (not the chunk-per-CTA
 *                    attention kernels), 0 = plain stream order.
 * Kernel selection (same arithmetic contract, results agree to fp32 rounding; all are covered by parity tests):
 *   "gemm_persistent" (default 2, process-wide): large-M projections as the persistent tcgen05 kernel with the
 *                    A operand (tf32 hi / lo parts) in tensor memory; 1 = A in shared memory, 0 = one tile per CTA;
 *   "enc_attn_tc"    (default 1): Transformer-encoder self attention on the tensor cores (head size 32), 0 = FFMA;
 *   "lstm_variant"   (default 0, process-wide): tensor-core LSTM keeps W_hh in tensor memory (0) or shared memory (1);
 *   "cross_mode"     (default 0): 1 = greedy decode runs the cross attention in memory-bank space (opt-in, slower);
 *   "cross_mb_version" (default 2, process-wide): implementation used by cross_mode 1;
 *   "cross_beam_kernel" (default 2, process-wide): cross attention with several beams per chunk at d = 256:
 *                    2 = persistent TMA-ring kernel, 1 = register-prefetch kernel, 0 = generic kernel (cross-checks);
 *   "cross_ring_groups" (default 2, process-wide): consumer warp groups of the ring kernel (2 for <= 5 beams);
 *   "gemm_a_tmem"    (default 1, process-wide): the 64-wide one-tile-per-CTA 3xTF32 kernels (decode step) keep the
 *                    A operand in tensor memory like the persistent kernel; 0 = shared memory (same bits);
 *   "frontend_fast"  (default 1, process-wide): per-read statistics from one shared-memory histogram (reads whose value
 *                    range exceeds 16384 fall back to the radix select) and the 8-samples-per-thread chunk gather;
 *                    0 = the general kernels; same bits;
 *   "gemm_wide_wave" (default 1, process-wide): decode-step projections whose 64-wide tiles would not fit one wave of SMs
 *                    (N = 1536 at d = 512) use 128-wide tiles; same bits;
 *   "gemm_serial_split" (default 1, process-wide): projections with >= 2048 rows run the split-K sum inside one CTA
 *                    (bit-identical to the cluster split), 0 = always the cluster split.
 *   "cross_packed_fast" (default 1, process-wide): greedy cross attention over fixed-point keys / values at d = 256 /
 *                    512 as the 256-column-slice kernel (1: 2 CTAs per SM, 2: 3 CTAs per SM), 0 = generic kernel.
 * Storage format (changes the stored precision of one intermediate; bounds in DESIGN.md, measured in profiles/):
 *   "kv_mode"        (default 3): how the Transformer decoder's projected memory keys / values
 *                    (onmt/modules/multi_headed_attn.py:142-153) are kept between the decode steps of a greedy batch:
 *                    0 = fp32 rows; 3 = 23-bit fixed point with one power-of-two step per row part (3 bytes per
 *                    element, absolute error <= 2^-23 of the part's largest element; parity mode: identical greedy
 *                    sequences and identity rates to fp32 storage, profiles/r02_identity_rates.md);
 *                    4 = 15-bit fixed point (2 bytes per element; the reduced-precision mode, never the default);
 *                    5 = the top 24 bits of the fp32 value; 1 / 2 = 24 / 16-bit fixed point decoded with conversion
 *                    instructions (cross-checks).  Modes 3 / 4 also apply to what the other decoders' attention
 *                    re-reads at every step: the RNN decoder's uh | H (global_attention.py:123-136) and the CNN
 *                    decoder's encoder top | combined state (conv_multi_step_attention.py:38-82; always 3 bytes);
 *   "kv_beam_packed" (default 1): beam search reads the fixed-point planes as well (kv_mode 3 / 4; d = 256: the TMA-ring
 *                    kernel over the int16 + uint8 planes, 2 - 8 beams; d = 512: the multi-query slice kernel while the
 *                    beams' score rows fit two CTAs per SM); 0 = fp32 rows.
 *   "block_ngram_repeat" (default 0): object beam: n > 0 gives a beam whose hypothesis repeats an n-gram -10e20
 *                    children (beam.py:101-124);
 *   "block_ngram_exclude" (default 0): bit v set = n-grams containing vocabulary id v are never blocked
 *                    (-ignore_when_blocking);
 *   "coverage_penalty" (default 0): object beam: 1 = wu, 2 = summary (penalties.py:39-57), weighted by the float
 *                    option beta (nd_set_float) and subtracted from the global score of finished hypotheses
 *                    (beam.py:203-216; with the length penalty "none" the reference subtracts it IN PLACE from the
 *                    running scores -- reproduced);
 *   "stepwise_penalty" (default 0): object beam: the coverage penalty is applied to the running scores at every step
 *                    (GNMTGlobalScorer.update_score, beam.py:218-227) instead of to finished hypotheses;
 *   "beam_attention" (default 0): beam decodes keep the per-step attention of every beam row for nd_beam_attention
 *                    (runs the loop without CUDA graphs);
 * Any nd_set_int call drops the engine's captured CUDA graphs (they are re-captured on the following calls).  */
ND_EXPORT int nd_set_int(nd_engine* e, const char* name, int64_t value);
/* float options: "beta" (default 0): weight of the object beam's coverage penalty (GNMTGlobalScorer, beam.py:190-193). */
ND_EXPORT int nd_set_float(nd_engine* e, const char* name, double value);

/* per-kernel-category device timing for bench.py's roofline figures: while a category bit is set,
 * every launch of that category is bracketed by CUDA events on the launching stream;
 * nd_profile_read synchronises on the last event, returns the summed milliseconds and launch
 * counts per category (arrays of ND_PROF_NCAT) and clears the log.                                */
enum {
  ND_PROF_GEMM = 0, ND_PROF_LSTM = 1, ND_PROF_CROSS_ATTN = 2, ND_PROF_SELF_ATTN = 3, ND_PROF_ENC_ATTN = 4,
  ND_PROF_MLP_ATTN = 5, ND_PROF_GENERATOR = 6, ND_PROF_BEAM = 7, ND_PROF_OTHER = 8, ND_PROF_NCAT = 9
};
ND_EXPORT int nd_profile_enable(nd_engine* e, uint32_t category_mask);
ND_EXPORT int nd_profile_read(nd_engine* e, double* out_ms, int64_t* out_count);

/* bookkeeping for bench.py: kernels launched by this engine since creation / last reset.        */
ND_EXPORT int64_t nd_launch_count(const nd_engine* e);
ND_EXPORT int nd_reset_launch_count(nd_engine* e);

/* standalone kernel entry points (unit tests / microbenchmarks) -------------------------------
 * C[M,N] = act(LN?(A)[M,K] . W[N,K]^T + bias) (+ residual); row-major fp32 device pointers.      */
ND_EXPORT int nd_test_gemm(nd_engine* e, int32_t mode, const float* A, const float* W, const float* bias,
                 const float* residual, const float* ln_gamma, const float* ln_beta, float* C,
                 int32_t M, int32_t N, int32_t K, int32_t relu, void* stream);

/* tuning aid: when non-NULL, CTA 0 of every subsequent tcgen05 GEMM launch writes clock64() stamps of
 * its pipeline phases into dev_buf32[0..31] (device memory).  NULL switches it off.                 */
ND_EXPORT int nd_debug_gemm_timeline(int64_t* dev_buf32);

#ifdef __cplusplus
}
#endif
#endif /* NANODEC_H_ */
