"""CPU: the oracle port reproduces the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, check_beam_against_golden, load_golden
from oracle import decode as odecode
from oracle.model import OracleModel, lstm_direction_explicit

FAST_CASES = ["l2t_d256", "t2t_d64", "nano2rnn_d256", "brnn2rnn_d256", "cnn2cnn_d256", "l2t_d64", "t2t_pe_d64",
              "nano2rnn_pe_d64", "cnn2cnn_pe_d64", "brnn2rnn_std_d256", "brnn2rnn_std_general_d64",
               "rnn2rnn_d256", "rnn2rnn_d64", "nano2rnn_gru_d64", "brnn2rnn_gru_d256", "l2t_gru_d64",
               "rnn2rnn_gru_std_d64", "resnet2t_d256", "resnet2rnn_d256", "resnet2t_d64", "crnn2t_d64",
               "crnn2rnn_d64", "ctrans2t_d64", "brnn2rnn_bridge_d64", "rnn2rnn_gru_bridge_d64", "t2t_avg_d64", "l2t_avg_d256"]


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_matches_reference_golden(name):
    g, cfg, sd, src, lengths = load_golden(name)
    om = OracleModel(sd, cfg)
    L = int(g["max_length"])
    trace = []
    out = odecode.greedy(om, src.t().contiguous().unsqueeze(2), lengths, max_length=L, trace_logits=trace)
    mb = out["memory_bank"]
    assert list(mb.shape) == list(g["memory_shape"])
    np.testing.assert_allclose(mb.flatten()[::97].numpy(), g["memory_sample"], rtol=0, atol=2e-5)
    assert abs(float(mb.double().sum()) - float(g["memory_sum"])) <= 1e-3 * max(1.0, float(g["memory_abs_sum"]) * 1e-3)
    np.testing.assert_array_equal(out["memory_lengths"].numpy(), g["memory_lengths"])
    logits = torch.stack([trace[int(s)] for s in g["logit_steps"]])
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=2e-4)
    np.testing.assert_array_equal(out["predictions"].numpy(), g["greedy_ids"])
    np.testing.assert_allclose(out["scores"].numpy(), g["greedy_scores"], atol=2e-4)


@pytest.mark.parametrize("name", ["l2t_d256", "nano2rnn_d256", "cnn2cnn_d256", "t2t_d64"])
def test_oracle_beam_matches_reference_golden(name):
    g, cfg, sd, src, lengths = load_golden(name)
    om = OracleModel(sd, cfg)
    L, K = int(g["max_length"]), int(g["beam_size"])
    out = odecode.beam_fast(om, src.t().contiguous().unsqueeze(2), lengths, beam_size=K, max_length=L)
    for i, hyp in enumerate(out["predictions"]):
        want = g["beam_ids"][i]
        want = want[want >= 0]
        np.testing.assert_array_equal(hyp[0].numpy(), want)
    np.testing.assert_allclose([s[0] for s in out["scores"]], g["beam_scores"], atol=1e-3)


@pytest.mark.parametrize("name", ["beam_l2t_d256_min20_alpha", "beam_nano2rnn_d256_min20", "beam_cnn2cnn_d256_min20",
                                  "beam_t2t_d256_min99"])
def test_oracle_fast_beam_matches_nondegenerate_reference_golden(name):
    """--fast beam of the unmodified reference with -min_length 99 / 20, -n_best 2 (and -alpha 0.7): hypotheses of
    21 ... 100 tokens, chunks retiring at different steps (translate/translator.py:714-810)."""
    g, cfg, sd, src, lengths = load_golden(name)
    L, K, NB = int(g["max_length"]), int(g["beam_size"]), int(g["n_best"])
    out = odecode.beam_fast(OracleModel(sd, cfg), src.t().contiguous().unsqueeze(2), lengths, beam_size=K,
                            max_length=L, min_length=int(g["min_length"]), n_best=NB, alpha=float(g["alpha"]))
    ids = np.full((len(out["predictions"]), NB, L), -1, dtype=np.int64)
    lens = np.zeros((ids.shape[0], NB), dtype=np.int64)
    for i, hyps in enumerate(out["predictions"]):
        for n in range(NB):
            lens[i, n] = len(hyps[n])
            ids[i, n, : lens[i, n]] = hyps[n].numpy()
    check_beam_against_golden(g, ids, lens, np.array([s[:NB] for s in out["scores"]]), atol=1e-3)


@pytest.mark.parametrize("name", ["l2t_d256", "brnn2rnn_d256", "cnn2cnn_d256", "t2t_d64", "l2t_d64"])
def test_oracle_object_beam_matches_reference_golden(name):
    """_translate_batch + onmt.translate.Beam (the default without --fast), n_best 2."""
    g, cfg, sd, src, lengths = load_golden(name)
    om = OracleModel(sd, cfg)
    L, K, NB = int(g["max_length"]), int(g["beam_size"]), int(g["obj_n_best"])
    out = odecode.beam_object(om, src.t().contiguous().unsqueeze(2), lengths, beam_size=K, max_length=L, n_best=NB)
    for i, hyps in enumerate(out["predictions"]):
        for n in range(NB):
            want = g["obj_ids"][i, n]
            np.testing.assert_array_equal(hyps[n].numpy(), want[want >= 0])
    np.testing.assert_allclose(np.array(out["scores"]), g["obj_scores"], atol=1e-3)


def test_object_beam_length_penalties_change_ranking_only():
    """wu / avg length penalties rescale the finished scores (penalties.py:65-88); beams themselves are unchanged."""
    g, cfg, sd, src, lengths = load_golden("l2t_d64")
    om = OracleModel(sd, cfg)
    s = src.t().contiguous().unsqueeze(2)
    a = odecode.beam_object(om, s, lengths, beam_size=4, max_length=30, n_best=4)
    b = odecode.beam_object(om, s, lengths, beam_size=4, max_length=30, n_best=4, length_penalty="avg")
    for ha, hb in zip(a["predictions"], b["predictions"]):
        assert sorted(tuple(h.tolist()) for h in ha) == sorted(tuple(h.tolist()) for h in hb)
    assert a["scores"] != b["scores"]


def test_greedy_runs_all_steps_and_golden_is_diverse():
    g, cfg, sd, src, lengths = load_golden("l2t_d256")
    assert g["greedy_ids"].shape == (int(g["B"]), int(g["max_length"]))        # no EOS early exit
    assert float(g["token_entropy_bits"]) > 1.4
    assert len({tuple(r) for r in g["greedy_ids"].tolist()}) >= 4              # signal dependent


def test_explicit_lstm_matches_packed_nn_lstm():
    torch.manual_seed(0)
    T, B, I, H = 37, 5, 3, 16
    x = torch.randn(T, B, I)
    lengths = torch.tensor([37, 30, 30, 11, 1])
    for t in range(B):
        x[lengths[t]:, t] = 0
    m = torch.nn.LSTM(I, H, 1, bidirectional=True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lengths.tolist())
    ref, (hn, cn) = m(packed)
    ref = torch.nn.utils.rnn.pad_packed_sequence(ref)[0]
    with torch.no_grad():
        f, hf, cf = lstm_direction_explicit(x, lengths, m.weight_ih_l0, m.weight_hh_l0, m.bias_ih_l0, m.bias_hh_l0, False)
        r, hr, cr = lstm_direction_explicit(x, lengths, m.weight_ih_l0_reverse, m.weight_hh_l0_reverse,
                                            m.bias_ih_l0_reverse, m.bias_hh_l0_reverse, True)
    torch.testing.assert_close(torch.cat([f, r], 2), ref.detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(hf, hn[0].detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(cr, cn[1].detach(), atol=1e-5, rtol=1e-5)


def test_nano_encoder_ignores_last_batchnorm_and_decoder_masks_value_one():
    # reference quirks restated in SURVEY.md §7 / Appendix A.5
    g, cfg, sd, src, lengths = load_golden("l2t_d64")
    om = OracleModel(sd, cfg)
    s = src.t().contiguous().unsqueeze(2)
    _, mb1, _ = om.encoder(s, lengths)
    sd2 = dict(sd)
    last = "encoder.batchnorm_%d" % (cfg.enc_layers - 1)
    sd2[last + ".weight"] = sd[last + ".weight"] * 3.0
    _, mb2, _ = OracleModel(sd2, cfg).encoder(s, lengths)
    assert torch.equal(mb1, mb2)
    out = odecode.greedy(om, s, lengths, max_length=3, return_attention=True)
    attn = out["attention"]                       # [L,B,T]
    one = s[:, :, 0].t().eq(1.0)                  # [B,T]
    assert one.any()
    assert float(attn[:, one].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["beam_attn_l2t_d64", "beam_attn_nano2rnn_d64"])
def test_oracle_beam_attention_matches_reference_golden(name):
    """results["attention"] of both beam searches under -attn_debug (translator.py:744-750,776,806-809 and :899-905 +
    beam.py:135,170-178), made by the unmodified reference on ragged chunks: one [len, width] matrix per hypothesis, the
    width being memory_lengths[i] of the reference's TILED length vector."""
    import os
    from helpers import GOLDEN
    import ast
    from nanodecoder_b200 import synth
    from nanodecoder_b200.config import ModelConfig
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = ModelConfig(**ast.literal_eval(str(g["cfg_json"])))
    sd = synth.make_state_dict(cfg, seed=int(g["weight_seed"]))
    src = torch.from_numpy(g["src"]).t().contiguous().unsqueeze(2)
    lengths = torch.from_numpy(g["lengths"])
    om = OracleModel(sd, cfg)
    K, NB, L, ML = int(g["beam_size"]), int(g["n_best"]), int(g["max_length"]), int(g["min_length"])
    for mode, fn in (("fast", odecode.beam_fast), ("obj", odecode.beam_object)):
        o = fn(om, src, lengths, beam_size=K, max_length=L, min_length=ML, n_best=NB, return_attention=True)
        for b in range(src.size(1)):
            for n in range(NB):
                want = g[mode + "_ids"][b, n]
                want = want[want >= 0]
                np.testing.assert_array_equal(o["predictions"][b][n].numpy(), want)
                a = o["attention"][b][n]
                w = int(g[mode + "_widths"][b, n])
                assert tuple(a.shape) == (len(want), w)
                np.testing.assert_allclose(a.numpy(), g[mode + "_attn"][b, n, : len(want), :w], atol=2e-6)


def test_oracle_object_beam_extras_match_reference_golden():
    """n-gram blocking (beam.py:101-124, incl. -ignore_when_blocking) and the coverage penalties of GNMTGlobalScorer
    (beam.py:203-243, penalties.py:39-57) -- incl. what the length penalty "none" makes of them: `normalized_probs -=
    penalty` runs IN PLACE on the beam's running scores and Beam.finished keeps views of them."""
    import ast
    from helpers import OBJ_EXTRA_CASES, load_case_npz
    for name in OBJ_EXTRA_CASES:
        g, cfg, sd, src, lengths = load_case_npz(name)
        okw = ast.literal_eval(str(g["okw"]))
        o = odecode.beam_object(OracleModel(sd, cfg), src.t().contiguous().unsqueeze(2), lengths,
                                beam_size=int(g["beam_size"]), max_length=int(g["max_length"]),
                                min_length=int(g["min_length"]), n_best=int(g["n_best"]), **okw)
        for b in range(src.size(0)):
            for n in range(int(g["n_best"])):
                want = g["ids"][b, n]
                np.testing.assert_array_equal(o["predictions"][b][n].numpy(), want[want >= 0], err_msg=name)
        np.testing.assert_allclose(np.array(o["scores"], dtype=np.float32), g["scores"], rtol=1e-3, err_msg=name)


def test_resnet_stem_is_a_stack_of_width3_time_convolutions():
    """What the engine's ResNet stem relies on: on the [B, 1, 1, T] image the reference builds, the (5,3)
    kernels with padding (2,1) only ever meet data through kernel row 2, and the (stride, 1) strides act on the height-1
    axis, so T is never shortened.  Checked by re-running the oracle's conv2d stem with every other kernel row zeroed and
    against a conv1d formulation of the first block."""
    import torch.nn.functional as F
    from nanodecoder_b200 import synth
    from nanodecoder_b200.config import ModelConfig
    from oracle.model import resnet_stem
    cfg = ModelConfig.family("resnet2rnn", d_model=64, dec_layers=2)
    sd = synth.make_state_dict(cfg, seed=5)
    src = torch.randn(37, 3, 1)
    full = resnet_stem(sd, "encoder.cnn", src)
    assert list(full.shape) == [37, 3, 64]                      # T unchanged by the "stride 2" layers
    sd2 = dict(sd)
    for k, v in sd.items():
        if k.startswith("encoder.cnn.") and v.dim() == 4 and v.size(2) == 5:
            w = torch.zeros_like(v)
            w[:, :, 2, :] = v[:, :, 2, :]
            sd2[k] = w
    assert torch.equal(resnet_stem(sd2, "encoder.cnn", src), full)
    # first layer as a plain conv1d over time with kernel row 2
    x = src[:, :, 0].t().unsqueeze(1)                           # [B, 1, T]
    y1 = F.conv1d(x, sd["encoder.cnn.conv1.weight"][:, :, 2, :], None, padding=1)
    y2 = F.conv2d(x.unsqueeze(2), sd["encoder.cnn.conv1.weight"], None, stride=(2, 1), padding=(2, 1)).squeeze(2)
    assert torch.allclose(y1, y2, atol=1e-6)
