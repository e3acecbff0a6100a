#!/usr/bin/env python
"""Benchmark of the translate hot path (BASELINE.json metric: basecalled bases/s at 1/2/4/8 B200, chunk 512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of 1024 synthetic chunks of 512 samples:
    encode -> project + pack the memory keys / values -> 100 greedy decode steps -> token ids.
Default workload at every N = the configuration the metric is quoted on, BASELINE.json configs[4] ("C5"):
Transformer2Transformer d=512, 6+6 layers, 8 heads, ff 2048, greedy, read-sharded (each GPU decodes its own reads;
weak scaling, no collective inside the step).  The other BASELINE configurations (C2 LSTM2Transformer greedy, C3
beam 5, C4 RNN2RNN / Conv2Conv) are measured with fewer steps into the same JSON line ("other_workloads", N = 1).

Printed JSON line (rank 0): the task contract, plus
  roofline      dominant kernel (decode cross-attention): achieved HBM GB/s vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/, CPU restatement of the reference) on a bounded sample of the same workload
  e2e           the same metric from RAW int16 reads in pinned host memory through the public API:
                SignalFrontend (H2D, median/MAD, chunking) -> Translator.translate (batch after batch, results
                double-buffered to the host) -> base strings -> shard.gather_records (NCCL at N > 1), all timed
  reduced_precision   the same workload with 16-bit memory keys / values (kv_mode q15), never the headline
``--impl reference`` times the CPU oracle port of the reference's translate path on the host cores (same workload).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "basecalled_bases_per_sec"
UNIT = "bases/s"
WORKLOAD = "t2t512_greedy_b1024"    # BASELINE.json configs[4]: the configuration the metric's 1/2/4/8-GPU clause names
B_PER_GPU, T, L = 1024, 512, 100
READ_CHUNKS = 16                    # chunks per synthetic read (the last one short)
WORKLOADS = {
    # name: (family, config kwargs, beam size, model description, roofline kernel category, BASELINE config)
    "t2t512_greedy_b1024": ("t2t", dict(d_model=512, enc_layers=6, dec_layers=6), 1,
                            "Transformer2Transformer 6+6, d=512, heads=8, ff=2048", "cross_attn", "C5"),
    "l2t_greedy_b1024": ("l2t", {}, 1, "NanoDecoder LSTM2Transformer enc3(biLSTM)+dec3, d=256, heads=8, ff=2048",
                         "cross_attn", "C2"),
    "l2t_beam5_b1024": ("l2t", {}, 5, "NanoDecoder LSTM2Transformer enc3(biLSTM)+dec3, d=256, --fast beam 5",
                        "cross_attn", "C3"),
    "t2t_greedy_b1024": ("t2t", {}, 1, "Transformer2Transformer 3+3, d=256, heads=8, ff=2048", "cross_attn", "C1 on GPU"),
    "nano2rnn_greedy_b1024": ("nano2rnn", {}, 1, "NanoDecoder LSTM enc3 -> InputFeed LSTM dec3, mlp attention, d=256",
                              "mlp_attn", "C4"),
    "brnn2rnn_greedy_b1024": ("brnn2rnn", {}, 1, "brnn enc3 -> InputFeed LSTM dec3, mlp attention, d=256", "mlp_attn", "C4"),
    "cnn2cnn_greedy_b1024": ("cnn2cnn", {}, 1, "Conv2Conv 3+3, kernel width 3, d=256", "mlp_attn", "C4"),
    # not a BASELINE.json config: the other encoder the authors' pipeline-train.sh trains (resnet -> transformer)
    "resnet2t_greedy_b1024": ("resnet2t", {}, 1, "ResNet stem (17 width-3 convs, 64-512 channels) -> Transformer dec3, d=256",
                              "cross_attn", "pipeline-train.sh resnet2transformer"),
}
KV_MODES = {"f32": 0, "q23": 3, "q15": 4}           # nd_set_int("kv_mode"), include/nanodec.h
KV_BYTES = {"f32": 4, "q23": 3, "q15": 2}


class Workload(object):
    def __init__(self, name):
        self.name = name
        self.family, self.kw, self.beam, self.desc, self.roof_cat, self.baseline_cfg = WORKLOADS[name]

    @property
    def min_length(self):
        """Beam workloads: with random-init weights the best hypothesis is "</s>" at step 0 (every further token only
        lowers the cumulative log-prob), every chunk would retire at once and the metric would be 0 bases/s.
        min_length = max_length - 1 (the reference's own -min_length flag) keeps all beams alive for the whole loop;
        tests/golden/beam_*_min99.npz pins exactly this setting against the unmodified reference."""
        return L - 1 if self.beam > 1 else 0

    def model_and_weights(self):
        from nanodecoder_b200 import synth
        from nanodecoder_b200.config import ModelConfig
        cfg = ModelConfig.family(self.family, **self.kw)
        return cfg, synth.make_state_dict(cfg, seed=2025)

    def kv_packed(self, kv):
        """The engine's rule (engine.cu decoder_init): every decoder keeps what its attention re-reads at each step in
        fixed point -- the Transformer decoder's memory keys / values (greedy, and beam search at d = 256 / 512), the RNN
        decoder's uh | H, the CNN decoder's encoder top | combined state."""
        return kv != "f32"        # round 2: the RNN decoder's (uh | H) and the CNN decoder's (keys | values) as well

    def config(self, n_gpus, kv):
        return {"workload": self.name, "baseline_config": self.baseline_cfg, "model": self.desc,
                "decode": "greedy" if self.beam == 1 else "--fast beam %d" % self.beam,
                "chunks_per_step_per_gpu": B_PER_GPU, "chunk_len": T, "max_length": L, "min_length": self.min_length,
                "global_chunks_per_step": B_PER_GPU * n_gpus, "parallelism": "read-sharded x%d" % n_gpus,
                "memory_kv_storage": kv if self.kv_packed(kv) else "f32",
                "l2_policy": "working set (GBs of attention keys / values per step) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock / throttle-reason samples DURING the timed region.  NVML from a thread of this process
    (two cheap queries per sample); falls back to an `nvidia-smi -lms` child when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu, interval_s=float(os.environ.get("ND_BENCH_SAMPLE_S", "0.05"))):
        self.gpu, self.interval = gpu, interval_s
        self.proc = self.thread = None
        self.samples, self.power, self.reason_bits, self.max_mhz = [], [], 0, None
        self._stop = False

    def _nvml_loop(self, nv, h):
        while not self._stop:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))))
                self.power.append((time.perf_counter(), nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(self.interval)

    def start(self):
        if os.environ.get("ND_BENCH_SAMPLER", "nvml") == "none":      # experiments only
            return
        if os.environ.get("ND_BENCH_SAMPLER", "nvml") == "nvml":
            try:
                import threading
                import pynvml as nv
                nv.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
                h = nv.nvmlDeviceGetHandleByIndex(idx)
                self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
                self._nv = nv
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    @staticmethod
    def _summary(sm, mx, reasons, source):
        # samples under load = the upper half (idle samples before/after the region drag the median down)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons), "source": source}

    def stop(self, window=None):
        """window = (t0, t1) perf_counter bounds of the timed region: only samples inside it are reported."""
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            inside = [m for (t, m) in self.samples if window is None or window[0] <= t <= window[1]]
            self.samples = inside if inside else [m for (_, m) in self.samples]
            nv, bits, reasons = self._nv, self.reason_bits, set()
            for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
                alt = attr.replace("ClocksEventReason", "ClocksThrottleReason")
                mask = getattr(nv, attr, getattr(nv, alt, 0))
                if bits & int(mask):
                    reasons.add(name)
            out = self._summary(self.samples, self.max_mhz, reasons, "nvml")
            pw = [w for (t, w) in self.power if window is None or window[0] <= t <= window[1]]
            if pw:
                out["power_w_mean"], out["power_w_max"] = round(sum(pw) / len(pw), 1), round(max(pw), 1)
            if self.samples:
                out["sm_mhz_min"] = min(self.samples)
            return out
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return self._summary(sm, max(mx) if mx else None, reasons, "nvidia-smi")


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(wl, n_chunks, steps, warmup, target_s=None, max_steps=40):
    """Oracle port (CPU restatement of the reference's translate path) on a bounded sample.  With ``target_s`` the
    step count is chosen from the first step's duration so that the sample takes about that many seconds of CPU work
    (never fewer than ``steps``, never more than ``max_steps``)."""
    from nanodecoder_b200 import synth
    from oracle import decode as od
    from oracle.model import OracleModel
    cfg, sd = wl.model_and_weights()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    chunks, lengths = synth.make_chunks(n_chunks, T=T, seed=1234, ragged=True, read_len=READ_CHUNKS)
    order = torch.argsort(lengths, descending=True, stable=True)
    src = chunks[order].t().contiguous().unsqueeze(2)
    lengths = lengths[order]
    om = OracleModel(sd, cfg)
    times, bases = [], 0
    i = 0
    while i < warmup + steps:
        t0 = time.perf_counter()
        if wl.beam == 1:
            out = od.greedy(om, src, lengths, max_length=L)
            n_bases = od.count_bases(out["predictions"])
        else:
            out = od.beam_fast(om, src, lengths, beam_size=wl.beam, max_length=L, min_length=wl.min_length)
            n_bases = sum(int((p[0] != 3).sum()) for p in out["predictions"])
        dt = time.perf_counter() - t0
        if target_s and i == 0:
            steps = max(steps, min(max_steps, int(target_s / max(dt, 1e-3))))
        i += 1
        if i > warmup:
            times.append(dt)
            bases += n_bases
    total = sum(times)
    return {"value": bases / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d chunks x %d steps of %s (oracle/ CPU port of translate_batch, fp32, torch %s), %.1f s"
                      % (n_chunks, steps, wl.name, torch.__version__, total),
            "chunks_per_s": n_chunks * steps / total, "ms_per_step": 1e3 * total / steps}


def cpu_chunks_for(wl, budget_s_per_step):
    """Sample size of the CPU arm: the reference README's batch of 50 chunks when a step of it fits the budget, else
    fewer (>= 8): a 4-chunk probe step gives the seconds per chunk."""
    from nanodecoder_b200 import synth
    from oracle import decode as od
    from oracle.model import OracleModel
    cfg, sd = wl.model_and_weights()
    torch.set_num_threads(os.cpu_count() or 1)
    chunks, lengths = synth.make_chunks(4, T=T, seed=99, ragged=False)
    om = OracleModel(sd, cfg)
    t0 = time.perf_counter()
    if wl.beam == 1:
        od.greedy(om, chunks.t().contiguous().unsqueeze(2), lengths, max_length=L)
    else:
        od.beam_fast(om, chunks.t().contiguous().unsqueeze(2), lengths, beam_size=wl.beam, max_length=L,
                     min_length=wl.min_length)
    per_chunk = (time.perf_counter() - t0) / 4.0            # pessimistic: small batches use the cores worse
    return int(max(8, min(50, budget_s_per_step / max(per_chunk, 1e-4))))


def run_reference(args):
    wl = Workload(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1)
    budget = float(os.environ.get("ND_REF_BUDGET_S", "200"))          # whole run: a few minutes
    n_chunks = cpu_chunks_for(wl, budget / (args.steps + warm))
    r = cpu_sample(wl, n_chunks, args.steps, warm)
    cfg = wl.config(1, "f32")
    cfg["chunks_per_step_per_gpu"] = n_chunks
    cfg["global_chunks_per_step"] = n_chunks
    cfg["parallelism"] = "cpu x%d threads" % r["cores"]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "chunks_per_s": r["chunks_per_s"], "token_steps_per_s": r["chunks_per_s"] * L,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def count_bases_dev(ids):
    is_eos = ids.eq(3) | ids.lt(0)
    first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full_like(ids[:, 0], ids.size(1)))
    return first.sum()


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def roofline_entry(wl, cfg, kv, prof, prof_ms, n_prof_steps):
    """Dominant kernel of the decode loop: algorithmic bytes per launch / CUDA-event time per launch."""
    peak, peak_src = peak_hbm()
    ms_cat, n_cat = prof.get(wl.roof_cat, (0.0, 0))
    d, B, K = cfg.d_model, B_PER_GPU, wl.beam
    packed = wl.kv_packed(kv)
    bpe = KV_BYTES[kv] if packed else 4
    # K and V (or uh and H, or keys and values of the conv attention) of every chunk once + (packed) the two steps per
    # memory position + query in and context out per row; the beams of a chunk share the K/V pass
    bytes_per_launch = B * (2 * T * d * bpe + (2 * T * 4 if packed else 0) + 2 * K * d * 4)
    achieved = bytes_per_launch / (ms_cat / max(n_cat, 1) * 1e-3) / 1e9 if n_cat else None
    if wl.roof_cat == "mlp_attn" and packed:
        kname = "mlp_attn_packed_kernel<%d,%d,%s> (decode global / conv attention over fixed-point planes)" % (
            d // 32, 1 if K == 1 else 8, kv)
    elif wl.roof_cat == "mlp_attn":
        kname = "mlp_attn_kernel<%d,1> (decode global / conv attention, fp32)" % (d // 32)
    elif packed and K > 1:
        kname = ("cross_attn_ring_kernel<%d,%d,%s>" % ((4, 2, kv) if K <= 4 else (5, 2, kv) if K == 5 else (8, 1, kv))
                 if d == 256 else "cross_attn_packed_mq_kernel<%s>" % kv) + \
                " (decode cross-attention, the beams of a chunk share one pass over its fixed-point K/V planes)"
    elif packed:
        kname = "cross_attn_packed_fast_kernel<%s> (decode cross-attention, fixed-point K/V, %d CTAs per chunk)" % (kv, 2)
    elif K == 1:
        kname = "cross_attn_kernel<%d,1> (decode cross-attention, fp32 K/V)" % (d // 32)
    else:
        kname = ("cross_attn_ring_kernel<%d,2>" % (4 if K <= 4 else 5) if d == 256 and K <= 5 else
                 "cross_attn_ring_kernel<8,1>" if d == 256 else "cross_attn_kernel<%d,8>" % (d // 32)) + \
                " (decode cross-attention, beams share the fp32 K/V pass)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "cross_attn_traffic.json"))).get(
            "%s/%s" % (wl.name, kv if packed else "f32"), {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    return {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_per_launch, "launches_timed": n_cat,
            "avg_launch_us": 1e3 * ms_cat / max(n_cat, 1), "share_of_step": ms_cat / prof_ms if prof_ms else None,
            "timing": "CUDA events around each launch on its stream, profiled pass of %d steps (no graph)" % n_prof_steps}


def device_leg(wl, args, dev, local, rank, world, kv, steps, warmup, sampler=None, barrier=None, want_profile=True):
    """value leg: inputs resident in HBM.  -> dict(ms (this rank), bases, launches, roofline, engine, cfg, sd, ...)"""
    from nanodecoder_b200 import synth
    from nanodecoder_b200.engine import Engine
    cfg, sd = wl.model_and_weights()
    B = B_PER_GPU
    eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=L, max_beam=wl.beam, gemm_mode=args.gemm_mode,
                 device=local)
    eng.set_option("kv_mode", KV_MODES[kv])
    chunks, lengths = synth.make_chunks(B, T=T, seed=1234 + rank, ragged=True, read_len=READ_CHUNKS)   # this rank's reads
    order = torch.argsort(lengths, descending=True, stable=True)
    chunks, lengths = chunks[order].contiguous(), lengths[order].contiguous()
    src_d, len_d = chunks.to(dev), lengths.to(dev)

    def step_device():
        eng.encode(src_d, len_d)
        if wl.beam == 1:
            return eng.decode_greedy(L)["ids"]
        return eng.decode_beam(wl.beam, 1, L, min_len=wl.min_length)["ids"][:, 0, :]   # best hypothesis, -1 padded

    bases_t = torch.zeros((), dtype=torch.int64, device=dev)
    for _ in range(warmup):
        bases_t += count_bases_dev(step_device())   # also loads torch's lazily-loaded kernels before the timed region
    torch.cuda.synchronize()
    eng.reset_launch_count()
    if barrier:
        barrier()
    t_region0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bases_t.zero_()
    ev0.record()
    for _ in range(steps):
        bases_t += count_bases_dev(step_device())
    ev1.record()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    t_region1 = time.perf_counter()
    out = {"ms": ev0.elapsed_time(ev1), "bases": int(bases_t.item()), "launches": eng.launch_count, "cfg": cfg, "sd": sd,
           "engine": eng, "chunks": chunks, "lengths": lengths, "window": (t_region0, t_region1)}
    if sampler is not None:
        out["clocks"] = sampler.stop((t_region0, t_region1))
    if want_profile:
        # per-kernel pass: the same steps with CUDA-event brackets around every launch of the dominant kernel
        n_prof = max(1, min(steps, 3))
        eng.profile_enable([wl.roof_cat])
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record()
        for _ in range(n_prof):
            step_device()
        ev3.record()
        torch.cuda.synchronize()
        prof = eng.profile_read()
        eng.profile_enable([])
        out["roofline"] = roofline_entry(wl, cfg, kv, prof, ev2.elapsed_time(ev3), n_prof)
    return out


def make_raw_pool(n_reads, seed):
    """Synthetic raw reads of READ_CHUNKS chunks each (the last chunk 64 .. T-1 samples long) as ONE pinned
    int16 buffer + the read lengths: what a file reader hands to the front end."""
    rng = np.random.default_rng([seed, 23])
    lens = (READ_CHUNKS - 1) * T + rng.integers(64, T, size=n_reads)
    total = int(lens.sum())
    levels = rng.normal(500.0, 80.0, size=total // 9 + 2)
    sig = np.repeat(levels, 9)[:total] + rng.normal(0.0, 12.0, size=total)
    flat = torch.from_numpy(np.clip(np.round(sig), 0, 2047).astype(np.int16)).pin_memory()
    return flat, lens


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gemm-mode", default="3xtf32", choices=["3xtf32", "tf32", "simt"])
    ap.add_argument("--kv-mode", default="q23", choices=sorted(KV_MODES),
                    help="storage of the decoder's memory keys / values (greedy): q23 = parity mode (default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE workloads / the q16 line")
    ap.add_argument("--workload", default=WORKLOAD, choices=sorted(WORKLOADS),
                    help="default = BASELINE.json configs[4] (the metric's configuration)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    wl = Workload(args.workload)

    import torch.distributed as dist
    from nanodecoder_b200 import shard
    from nanodecoder_b200.checkpoint import Vocab
    from nanodecoder_b200.inputters.nano_dataset import SignalFrontend
    from nanodecoder_b200.opts import default_translate_opt
    from nanodecoder_b200.translate.translator import Translator, _Field

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler runs from before the warm-up (no idle gap in front of the timed region: the GPU would drop
    # its clocks and the first timed step would pay the ramp); only samples taken inside the timed region are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    B = B_PER_GPU
    leg = device_leg(wl, args, dev, local, rank, world, args.kv_mode, args.steps, args.warmup,
                     sampler=sampler if rank == 0 else None, barrier=barrier)
    eng, cfg, sd = leg["engine"], leg["cfg"], leg["sd"]
    t = torch.tensor([leg["ms"]], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(leg["bases"]), float(leg["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # slowest rank defines the step time
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)      # per-rank counts gathered over NCCL
    ms_max = float(t.item())
    value = float(tot[0].item()) / (ms_max / 1e3)

    # ---------------- end to end: RAW reads in pinned host memory -> front end -> Translator.translate -> strings
    # -> per-read records gathered on rank 0.  The job is a pool of world * steps * 64 reads (= steps batches of 1024
    # chunks per GPU) sharded over the ranks with shard.partition_reads; every rank streams its reads through its
    # engine batch after batch (results double-buffered to the host, strings built behind the GPU).
    opt = default_translate_opt(beam_size=wl.beam, fast=wl.beam > 1, batch_size=B, max_length=L,
                                min_length=wl.min_length, src_seq_length=T, gpu=local, gemm_mode=args.gemm_mode)
    tr = Translator(eng, {"tgt": _Field(Vocab(cfg.vocab))}, opt, cfg)
    fe = SignalFrontend(eng, "median", T, T)
    reads_per_step = B // READ_CHUNKS
    n_reads_job = world * args.steps * reads_per_step
    sizes = (READ_CHUNKS - 1) * T + np.random.default_rng([77, 23]).integers(64, T, size=n_reads_job)
    mine = shard.partition_reads(sizes.tolist(), world)[rank]           # same table on every rank, no communication
    flat, lens = make_raw_pool(len(mine), 500 + rank)

    def e2e_once(flat_t, lens_a, read_ids):
        chunks_d, clen_d, chunk_read = fe.from_flat(flat_t, lens_a)
        _, preds = tr.translate(src=(chunks_d, clen_d), batch_size=B)
        nb = np.fromiter(((p[0].count(" ") + 1) if p[0] else 0 for p in preds), dtype=np.int64, count=len(preds))
        per_read = np.bincount(chunk_read, weights=nb, minlength=len(lens_a)).astype(np.int64)
        records = [(int(read_ids[i]), int(per_read[i])) for i in range(len(lens_a))]
        merged = shard.gather_records(records, dst=0)                   # NCCL at N > 1 (sizes, then payload)
        return int(nb.sum()), merged

    w_flat, w_lens = make_raw_pool(reads_per_step, 400 + rank)
    e2e_once(w_flat, w_lens, list(range(reads_per_step)))                # warm (pinned result buffers, graph reuse)
    barrier()
    t0 = time.perf_counter()
    e2e_bases, merged = e2e_once(flat, lens, mine)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    be = torch.tensor([float(e2e_bases)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(be, op=dist.ReduceOp.SUM)
    e2e_value = float(be.item()) / float(te.item())
    n_chunks_rank = len(mine) * READ_CHUNKS
    if rank == 0:
        assert merged is not None and len(merged) == n_reads_job and sum(r[1] for r in merged) == int(be.item())

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": ("f32 (tcgen05 3xTF32 GEMMs, fp32 attention arithmetic, memory K/V stored as %s)" % args.kv_mode)
                if args.gemm_mode == "3xtf32" else "f32/" + args.gemm_mode, "data": "synthetic",
                "config": wl.config(world, args.kv_mode), "clocks": leg.get("clocks"),
                "chunks_per_s": B * world * args.steps / (ms_max / 1e3),
                "token_steps_per_s": B * world * L * args.steps / (ms_max / 1e3),
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": int(flat.numel() * 2 / max(args.steps, 1) + n_chunks_rank * 12 / max(args.steps, 1)),
                        "d2h_bytes_per_step": int(B * L * 8 + B * 4 + B * 8),
                        "chunks_per_s": n_chunks_rank * world / float(te.item()), "seconds": float(te.item()),
                        "reads": n_reads_job, "chunks": n_chunks_rank * world,
                        "api": "SignalFrontend.from_flat(pinned int16 reads) -> Translator.translate(batch_size=1024) "
                               "-> base strings -> shard.gather_records (rank 0)",
                        "limiter": "GPU decode (host work of batch k overlaps the GPU work of batch k+1); see DESIGN.md 6"},
                "n_model_params": int(sum(v.numel() for v in sd.values() if torch.is_floating_point(v))),
                "gpu_launches": int(tot[1].item()), "roofline": leg["roofline"]}
    eng.close()
    del tr, fe, eng, leg
    torch.cuda.empty_cache()

    if world == 1 and not args.no_extras:
        # the reduced-precision line and the other BASELINE configurations, fewer steps, same harness
        k2, w2 = max(1, min(args.steps, 5)), 3
        other = {}
        for name, kv in [(wl.name, "q15")] + [(n, args.kv_mode) for n in WORKLOADS if n != wl.name]:
            w = Workload(name)
            try:
                r = device_leg(w, args, dev, local, rank, world, kv, k2, w2)
                r["engine"].close()
                entry = {"value": r["bases"] / (r["ms"] / 1e3), "unit": UNIT, "ms_per_step": r["ms"] / k2, "steps": k2,
                         "chunks_per_s": B * k2 / (r["ms"] / 1e3), "gpu_launches": int(r["launches"]),
                         "config": w.config(1, kv), "roofline": r["roofline"]}
            except Exception as e:                                      # noqa: BLE001 - recorded, not hidden
                entry = {"error": "%s: %s" % (type(e).__name__, e)}
            del r
            torch.cuda.empty_cache()
            if name == wl.name:
                entry["note"] = ("2-byte fixed-point memory keys / values: reduced-precision mode, bound and identity "
                                 "rate in DESIGN.md 4.5 / profiles/; never the headline")
                line["reduced_precision"] = entry
            else:
                other[name] = entry
        line["other_workloads"] = other

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = cpu_chunks_for(wl, 8.0)
            line["cpu_baseline"] = {k: v for k, v in cpu_sample(wl, n_cpu, 2, 1, target_s=20.0, max_steps=6).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
