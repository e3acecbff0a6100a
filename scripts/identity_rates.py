#!/usr/bin/env python
"""Identity rates of the CUDA path against the oracle port over >= 1000 chunks per model family (GPU box).

    python scripts/identity_rates.py [--n 1024] [--out gpurun_out/identity_rates.json] [--kv-mode q24]

north_star criteria: greedy token sequences identical; --fast beam outputs identical on >= 99.9 % of chunks, any
divergence explained by a logit tie.  For every family the GPU decodes all n chunks in one batch; the oracle port
(oracle/, the CPU restatement pinned bit-for-bit against the unmodified reference, tests/golden) decodes the same
chunks on the host cores in slices.  A greedy mismatch is explained with the oracle's own log-probs at the first
diverging step (gap between the two tokens); a beam mismatch with the score gap between the two hypotheses.
This script is test infrastructure (it imports oracle/); nothing in the product path does.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nanodecoder_b200 import synth  # noqa: E402
from nanodecoder_b200.config import ModelConfig  # noqa: E402
from nanodecoder_b200.engine import Engine  # noqa: E402
from oracle import decode as od  # noqa: E402
from oracle.model import OracleModel  # noqa: E402

T, L = 512, 100
GREEDY = [("l2t", {}), ("t2t", {}), ("nano2rnn", {}), ("brnn2rnn", {}), ("cnn2cnn", {}),
          ("t2t", dict(d_model=512, enc_layers=6, dec_layers=6))]
BEAMS = [("l2t", {}, 99), ("l2t", {}, 20), ("nano2rnn", {}, 20)]


def chunks_for(n, seed):
    chunks, lengths = synth.make_chunks(n, T=T, seed=seed, ragged=True, read_len=16)
    order = torch.argsort(lengths, descending=True, stable=True)
    return chunks[order].contiguous(), lengths[order].contiguous()


def oracle_slices(n, slice_b):
    """The GPU decodes all n chunks (sorted by length, longest first) as ONE batch padded to T; the CPU oracle takes
    them in slices.  A slice must be padded to the same width (the Transformer decoder attends the padding,
    decoder/transformer.py:219-221, and the reference pads a batch to its longest chunk), so every slice is a strided
    pick i, i + k, i + 2k ... of the sorted list: it starts with a full-length chunk and stays sorted."""
    k = (n + slice_b - 1) // slice_b
    return [torch.arange(i, n, k) for i in range(k)]


def greedy_case(family, kw, n, slice_b, opts_list):
    """-> one row per engine option set (the oracle runs once)"""
    cfg = ModelConfig.family(family, **kw)
    sd = synth.make_state_dict(cfg, seed=2025)
    chunks, lengths = chunks_for(n, 4321)
    gpu = []
    for opts in opts_list:
        eng = Engine(cfg, sd, max_batch=n, max_src_len=T, max_tgt_len=L)
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.encode(chunks.cuda(), lengths.cuda())
        got = eng.decode_greedy(L, return_logits=True)
        torch.cuda.synchronize()
        gpu.append((got["ids"].cpu(), got["logits"].cpu()))            # [n,L], [L,n,V]
        eng.close()
    om = OracleModel(sd, cfg)
    acc = [dict(same=0, worst=0.0, explain=[]) for _ in opts_list]
    t0 = time.time()
    for sel in oracle_slices(n, slice_b):
        trace = []
        want = od.greedy(om, chunks[sel].t().contiguous().unsqueeze(2), lengths[sel], max_length=L, trace_logits=trace)
        tr = torch.stack(trace)                                      # [L, b, V]
        for (ids, logits), a in zip(gpu, acc):
            for j, c in enumerate(sel.tolist()):
                if torch.equal(ids[c], want["predictions"][j]):
                    a["same"] += 1
                    rel = float((logits[:, c] - tr[:, j]).abs().max() / tr[:, j].abs().max())
                    a["worst"] = max(a["worst"], rel)
                else:
                    t = int((ids[c] != want["predictions"][j]).nonzero()[0])
                    lp = tr[t, j]
                    x, y = int(want["predictions"][j][t]), int(ids[c, t])
                    a["explain"].append({"chunk": c, "step": t, "oracle_tok": x, "gpu_tok": y,
                                         "oracle_logp_gap": float(lp[x] - lp[y]),
                                         "gpu_logp_gap": float(logits[t, c, x] - logits[t, c, y])})
    secs = round(time.time() - t0, 1)
    return [{"family": family, "cfg": kw, "mode": "greedy", "engine_opts": opts, "chunks": n, "identical": a["same"],
             "rate": a["same"] / n, "max_rel_logit_err_identical_chunks": a["worst"], "n_mismatch": len(a["explain"]),
             "mismatches": a["explain"][:12], "oracle_cpu_seconds": secs} for opts, a in zip(opts_list, acc)]


def beam_case(family, kw, min_len, n, slice_b, opts, K=5, NB=1):
    cfg = ModelConfig.family(family, **kw)
    sd = synth.make_state_dict(cfg, seed=2025)
    chunks, lengths = chunks_for(n, 4322)
    eng = Engine(cfg, sd, max_batch=n, max_src_len=T, max_tgt_len=L, max_beam=K)
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.encode(chunks.cuda(), lengths.cuda())
    got = eng.decode_beam(K, NB, L, min_len=min_len)
    torch.cuda.synchronize()
    ids, lens, sc = got["ids"].cpu(), got["lens"].cpu(), got["scores"].cpu()
    eng.close()
    om = OracleModel(sd, cfg)
    same, explain, hist = 0, [], []
    all_margins = []
    t0 = time.time()
    for sel in oracle_slices(n, slice_b):
        margins = []
        want = od.beam_fast(om, chunks[sel].t().contiguous().unsqueeze(2), lengths[sel], beam_size=K, max_length=L,
                            min_length=min_len, n_best=NB, margins=margins)
        all_margins.extend(margins)
        for j, c in enumerate(sel.tolist()):
            w = want["predictions"][j][0]
            g = ids[c, 0, : int(lens[c, 0])]
            hist.append(len(w))
            if torch.equal(g, w):
                same += 1
            else:
                explain.append({"chunk": c, "tightest_candidate_gap_in_the_oracle_search": margins[j],
                                "oracle_len": len(w), "gpu_len": int(lens[c, 0]),
                                "oracle_score": float(want["scores"][j][0]), "gpu_score": float(sc[c, 0]),
                                "score_gap": float(want["scores"][j][0]) - float(sc[c, 0])})
    return {"family": family, "cfg": kw, "mode": "--fast beam %d, min_length %d" % (K, min_len), "chunks": n,
            "identical": same, "rate": same / n, "hyp_len_min": min(hist), "hyp_len_mean": sum(hist) / len(hist),
            "tightest_gap_quantiles_all_chunks": {q: float(torch.tensor(all_margins).quantile(q)) for q in
                                                  (0.001, 0.01, 0.1, 0.5)},
            "chunks_with_gap_below_1e-5": int((torch.tensor(all_margins) < 1e-5).sum()),
            "mismatches": explain[:20], "oracle_cpu_seconds": round(time.time() - t0, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--n-big", type=int, default=256, help="chunks for the d=512 6+6 model (CPU oracle is slow there)")
    ap.add_argument("--slice", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "identity_rates.json"))
    ap.add_argument("--opts", default="", help="engine options of the beam cases, e.g. cross_beam_kernel=1")
    ap.add_argument("--greedy-opts", default="kv_mode=0;kv_mode=3;kv_mode=4",
                    help="';'-separated engine option sets compared against ONE oracle run per family")
    ap.add_argument("--only", default="", help="greedy | beam")
    ap.add_argument("--families", default="", help="comma-separated family names to keep (default: all)")
    args = ap.parse_args()
    parse = lambda txt: {k: int(v) for k, v in (kv.split("=") for kv in txt.split(",") if kv)}
    opts = parse(args.opts)
    greedy_opts = [parse(x) for x in args.greedy_opts.split(";")]
    torch.set_num_threads(os.cpu_count() or 1)
    rows = []
    if args.only in ("", "greedy"):
        for family, kw in GREEDY:
            if args.families and family not in args.families.split(","):
                continue
            n = args.n_big if kw.get("d_model", 256) > 256 else args.n
            for row in greedy_case(family, kw, n, args.slice, greedy_opts):
                rows.append(row)
                print(json.dumps(row), flush=True)
    if args.only in ("", "beam"):
        for family, kw, ml in BEAMS:
            if args.families and family not in args.families.split(","):
                continue
            rows.append(beam_case(family, kw, ml, args.n, args.slice, opts))
            print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"opts": opts, "host_cores": os.cpu_count(), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
