"""GPU: unit parity of the individual CUDA kernels, called through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_err
from nanodecoder_b200 import synth
from nanodecoder_b200.config import ModelConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_engine():
    from nanodecoder_b200.engine import Engine
    cfg = ModelConfig.family("l2t", d_model=64, d_ff=128, enc_layers=2, dec_layers=2)
    return Engine(cfg, synth.make_state_dict(cfg), max_batch=8, max_src_len=128, max_tgt_len=8, gemm_mode="simt")


SHAPES = [(128, 128, 256), (1024, 768, 256), (1024, 256, 2048), (300, 200, 96), (77, 64, 64), (4096, 1024, 256),
          (5, 8, 32), (130, 2048, 256)]


@pytest.mark.parametrize("mode,tol", [("simt", 2e-6), ("3xtf32", 1e-5), ("tf32", 3e-3)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain(small_engine, mode, tol, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    C = small_engine.test_gemm(mode, A, W, bias=bias)
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t() + bias.double())
    assert rel_err(C, ref) < tol, (mode, M, N, K, rel_err(C, ref))


DEFAULT_PERSISTENT = 2


@pytest.mark.parametrize("mode,tol", [("3xtf32", 1e-5), ("tf32", 3e-3)])
@pytest.mark.parametrize("M,N,K,ln", [(20000, 1024, 256, False), (19999, 520, 96, True), (40000, 256, 512, True),
                                      (16384, 2048, 256, True)])
def test_gemm_persistent_kernel(small_engine, mode, tol, M, N, K, ln):
    """Large tile counts run as the persistent kernel (double-buffered TMEM accumulator, dedicated epilogue warps):
    ragged M and N, LayerNorm folding, ReLU and residual must match the one-tile-per-CTA kernel and fp64."""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 1.5 + 0.25).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    gam = (1 + 0.1 * torch.randn(K, generator=g)).cuda()
    bet = (0.1 * torch.randn(K, generator=g)).cuda()
    kw = dict(bias=bias, residual=res, ln=(gam, bet) if ln else None, relu=1 if ln else 0)
    C = small_engine.test_gemm(mode, A, W, **kw)
    small_engine.set_option("gemm_persistent", 0)
    try:
        C1 = small_engine.test_gemm(mode, A, W, **kw)
        small_engine.set_option("gemm_persistent", 2)          # A operand (tf32 hi / lo) in tensor memory
        C2 = small_engine.test_gemm(mode, A, W, **kw)
    finally:
        small_engine.set_option("gemm_persistent", DEFAULT_PERSISTENT)
    torch.cuda.synchronize()
    if mode == "3xtf32":
        small_engine.set_option("gemm_persistent", 1)
        try:
            C3 = small_engine.test_gemm(mode, A, W, **kw)
            torch.cuda.synchronize()
        finally:
            small_engine.set_option("gemm_persistent", DEFAULT_PERSISTENT)
        assert torch.equal(C2, C3)                               # same tf32 operands, same MMA order: same bits
    An = torch.nn.functional.layer_norm(A.double(), (K,), gam.double(), bet.double(), 1e-6) if ln else A.double()
    ref = An @ W.double().t() + bias.double()
    ref = (torch.relu(ref) if ln else ref) + res.double()
    assert rel_err(C, ref) < tol, rel_err(C, ref)
    assert rel_err(C, C1.double()) < (5e-6 if mode == "3xtf32" else 1e-5)   # LayerNorm moments are grouped differently


@pytest.mark.parametrize("mode,tol", [("simt", 3e-6), ("3xtf32", 1e-5), ("tf32", 3e-3)])
def test_gemm_layernorm_relu_residual(small_engine, mode, tol):
    g = torch.Generator().manual_seed(5)
    M, N, K = 1000, 512, 256
    A = (torch.randn(M, K, generator=g) * 2 + 0.5).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    gam = (1 + 0.1 * torch.randn(K, generator=g)).cuda()
    bet = (0.1 * torch.randn(K, generator=g)).cuda()
    C = small_engine.test_gemm(mode, A, W, bias=bias, residual=res, ln=(gam, bet), relu=1)
    torch.cuda.synchronize()
    An = torch.nn.functional.layer_norm(A.double(), (K,), gam.double(), bet.double(), 1e-6)
    ref = torch.relu(An @ W.double().t() + bias.double()) + res.double()
    assert rel_err(C, ref) < tol, rel_err(C, ref)
    if mode == "3xtf32":
        # A operand in tensor memory (default) vs in shared memory: same tf32 operands, same MMA order -> same bits,
        # for the cluster split-K shapes of the decode step as well
        for (M2, N2, K2) in ((M, N, K), (1024, 256, 256), (1024, 256, 2048), (300, 768, 256)):
            A2 = (torch.randn(M2, K2, generator=g) * 2 + 0.5).cuda()
            W2 = (torch.randn(N2, K2, generator=g) / K2 ** 0.5).cuda()
            b2 = torch.randn(N2, generator=g).cuda()
            g2 = (1 + 0.1 * torch.randn(K2, generator=g)).cuda()
            be2 = (0.1 * torch.randn(K2, generator=g)).cuda()
            c_t = small_engine.test_gemm(mode, A2, W2, bias=b2, ln=(g2, be2))
            small_engine.set_option("gemm_a_tmem", 0)
            try:
                c_s = small_engine.test_gemm(mode, A2, W2, bias=b2, ln=(g2, be2))
                torch.cuda.synchronize()
            finally:
                small_engine.set_option("gemm_a_tmem", 1)
            assert torch.equal(c_t, c_s), (M2, N2, K2)


def test_frontend_bit_exact_vs_oracle_and_reference_golden(small_engine):
    from oracle import frontend as ofe
    from nanodecoder_b200.inputters.nano_dataset import SignalFrontend
    g = np.load(GOLDEN + "/frontend.npz")
    reads = [g["raw_%d" % i] for i in range(int(g["n_reads"]))]
    reads += synth.make_raw_reads(5, seed=3, min_len=3000, max_len=60000)
    reads.append(np.full(40, 7, dtype=np.int16) + np.arange(40, dtype=np.int16) % 3)
    rng = np.random.default_rng(12)
    reads.append(rng.integers(-30000, 30000, size=5003).astype(np.int16))     # value range > histogram: radix-select path
    reads.append(rng.integers(-32768, 32768, size=9).astype(np.int16))
    reads.append(np.array([5, 5, 5, 9], dtype=np.int16))                       # MAD = 0 -> inf / nan like numpy
    for norm in ("median", "mean"):
        for (L, S) in ((512, 512), (300, 60), (128, 64)):
            fe = SignalFrontend(small_engine, norm, L, S)
            chunks, lens, cread = fe(reads)
            small_engine.set_option("frontend_fast", 0)                       # the general kernels: same bits
            chunks_g, lens_g, _ = fe(reads)
            small_engine.set_option("frontend_fast", 1)
            torch.cuda.synchronize()
            assert torch.equal(lens, lens_g)
            np.testing.assert_array_equal(chunks.cpu().numpy(), chunks_g.cpu().numpy())
            chunks, lens = chunks.cpu().numpy(), lens.cpu().numpy()
            pos = 0
            for ri, raw in enumerate(reads):
                with np.errstate(all="ignore"):
                    want = ofe.frontend(raw, norm, L, S)
                for w in want:
                    assert cread[pos] == ri and lens[pos] == len(w)
                    got = chunks[pos, : len(w)]
                    np.testing.assert_array_equal(got, w)                # bit exact (std: numpy's summation order)
                    assert not chunks[pos, len(w):].any()
                    pos += 1
            assert pos == len(lens)
    # the reference's own golden output for read 0 (median, 512/512)
    fe = SignalFrontend(small_engine, "median", 512, 512)
    chunks, lens, _ = fe([reads[0]])
    flat = np.concatenate([chunks[i, : int(lens[i])].cpu().numpy() for i in range(len(lens))])
    want = g["r0_median_512_512_flat"]
    np.testing.assert_array_equal(flat if flat.size < 6000 else flat[::7], want)


def test_frontend_float_valued_reads_bit_exact_vs_oracle(small_engine):
    """`.signal` files with non-integer samples (the reference parses every token with float(), labelop.py:216-217):
    fp64 kernels -- radix-select median / MAD over doubles, np.std in numpy's summation order -- against the numpy
    oracle, bit for bit, alone and pooled with int16 reads (which are widened and must not change)."""
    from oracle import frontend as ofe
    from nanodecoder_b200.inputters.nano_dataset import SignalFrontend
    rng = np.random.default_rng(5)
    reads = [rng.normal(0.3, 1.7, size=n) for n in (1, 2, 7, 300, 4097, 70001)]
    reads.append(np.round(rng.normal(90.0, 12.0, size=2500), 2))                 # pA-like values with many ties
    reads.append(np.array([1.5, -2.25, 1.5, 1.5, 8.0, -2.25]))
    ints = synth.make_raw_reads(2, seed=8, min_len=900, max_len=5000)
    for norm in ("median", "mean"):
        for (L, S) in ((512, 512), (300, 60)):
            fe = SignalFrontend(small_engine, norm, L, S)
            pool = reads + ints
            chunks, lens, cread = fe(pool)
            torch.cuda.synchronize()
            chunks, lens = chunks.cpu().numpy(), lens.cpu().numpy()
            pos = 0
            for ri, raw in enumerate(pool):
                if raw.size == 1 or (norm == "median" and ri == 1 and False):
                    pass
                want = ofe.frontend(raw, norm, L, S)
                for w in want:
                    assert cread[pos] == ri and lens[pos] == len(w)
                    np.testing.assert_array_equal(chunks[pos, : len(w)], w, err_msg="read %d %s" % (ri, norm))
                    pos += 1
            assert pos == len(lens)


@pytest.mark.parametrize("mode", ["simt", "3xtf32"])
@pytest.mark.parametrize("d", [64, 256])
def test_nano_encoder_vs_oracle(d, mode):
    """LSTM recurrence kernel (+ input projection GEMMs, BN prologue, W) against the fp32 oracle."""
    from nanodecoder_b200.engine import Engine
    from oracle.model import OracleModel
    cfg = ModelConfig.family("l2t", d_model=d, d_ff=128, enc_layers=3, dec_layers=1)
    sd = synth.make_state_dict(cfg)
    B, T = 21, 160
    chunks, lengths = synth.make_chunks(B, T=T, seed=3, ragged=True, read_len=2)
    lengths[0] = T
    lengths[5] = 1
    chunks[5, 1:] = 0
    eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=4, gemm_mode=mode)
    eng.encode(chunks.cuda(), lengths.cuda())
    mb, lens = eng.memory_bank()
    torch.cuda.synchronize()
    om = OracleModel(sd, cfg)
    with torch.no_grad():
        _, want, wl = om.encoder(chunks.t().contiguous().unsqueeze(2), lengths)
    assert torch.equal(lens.cpu(), wl)
    assert mb.shape == want.shape
    err = rel_err(mb.cpu(), want)
    assert err < 1e-3, err
    print("nano encoder d=%d rel err %.2e" % (d, err))


@pytest.mark.parametrize("d", [256, 512])
@pytest.mark.parametrize("T", [512, 333, 128])
def test_transformer_encoder_attention_tensor_core_vs_ffma_and_oracle(T, d):
    """Encoder self attention on tcgen05 (fp16 two-term split, P in tensor memory) against the fp32 FFMA kernel
    and the oracle: ragged T, keys masked by value (src == 0.0), several chunks."""
    from nanodecoder_b200.engine import Engine
    from oracle.model import OracleModel
    cfg = ModelConfig.family("t2t", d_model=d, d_ff=512, enc_layers=2, dec_layers=1)     # head size 32 | 64
    sd = synth.make_state_dict(cfg, seed=4)
    B = 9
    chunks, lengths = synth.make_chunks(B, T=T, seed=8, ragged=True, read_len=3)
    chunks[1, : T // 3] = 0.0                               # a long run of masked keys
    order = torch.argsort(lengths, descending=True, stable=True)
    chunks, lengths = chunks[order], lengths[order]
    outs = []
    for tc in (1, 0):
        eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=4)
        eng.set_option("enc_attn_tc", tc)
        eng.encode(chunks.cuda(), lengths.cuda())
        outs.append(eng.memory_bank()[0].cpu())
    with torch.no_grad():
        _, want, _ = OracleModel(sd, cfg).encoder(chunks.t().contiguous().unsqueeze(2), lengths)
    e_tc, e_ff, e_x = rel_err(outs[0], want), rel_err(outs[1], want), rel_err(outs[0], outs[1].double())
    print("T=%d d=%d: tensor-core vs oracle %.2e, FFMA vs oracle %.2e, tensor-core vs FFMA %.2e" % (T, d, e_tc, e_ff, e_x))
    assert e_tc < 1e-4 and e_ff < 1e-4 and e_x < 1e-4


@pytest.mark.parametrize("N,K,ln", [(256, 256, True), (256, 256, False), (256, 2048, False), (512, 512, True)])
def test_gemm_split_k_many_row_tiles_is_batch_invariant(small_engine, N, K, ln):
    """Split-K shapes with many rows (beam search: M = 5 x batch) run the K slices of the cluster split as separate
    accumulators of ONE CTA per 128 x 128 tile.  The summation order depends on (N, K) only, so the first 1024 rows must
    be BIT-identical to a 1024-row call (cluster split-K kernel) and to the cluster kernel run on all rows."""
    g = torch.Generator().manual_seed(N + K)
    M = 5121
    A = (torch.randn(M, K, generator=g) * 1.3 + 0.2).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    gam = (1 + 0.1 * torch.randn(K, generator=g)).cuda()
    bet = (0.1 * torch.randn(K, generator=g)).cuda()
    lnp = (gam, bet) if ln else None
    full = small_engine.test_gemm("3xtf32", A, W, bias=bias, residual=res, ln=lnp)
    part = small_engine.test_gemm("3xtf32", A[:1024].contiguous(), W, bias=bias, residual=res[:1024].contiguous(), ln=lnp)
    small_engine.set_option("gemm_serial_split", 0)
    try:
        clus = small_engine.test_gemm("3xtf32", A, W, bias=bias, residual=res, ln=lnp)
    finally:
        small_engine.set_option("gemm_serial_split", 1)
    torch.cuda.synchronize()
    An = torch.nn.functional.layer_norm(A.double(), (K,), gam.double(), bet.double(), 1e-6) if ln else A.double()
    ref = An @ W.double().t() + bias.double() + res.double()
    assert rel_err(full, ref) < 1e-5, rel_err(full, ref)
    assert torch.equal(full[:1024], part)
    assert torch.equal(full, clus)
