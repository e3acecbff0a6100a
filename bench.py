#!/usr/bin/env python
"""Benchmark of the translate hot path (BASELINE.json metric: basecalled bases/s, chunk 512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic chunks:
    encode (LSTM stack) -> project memory K/V -> 100 greedy decode steps -> token ids.
Workload at every N: BASELINE config[1] — NanoDecoder LSTM2Transformer (3+3 layers, d=256, 8 heads,
ff 2048), greedy, batch 1024 chunks of 512 samples, max_length 100 — per GPU (weak scaling: reads
shard across GPUs with no collective inside the step; NCCL only gathers counts and timings).

Printed JSON line (rank 0): see the task contract; additionally
  roofline      dominant kernel (decode cross-attention) achieved HBM GB/s vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/, CPU restatement of the reference) on a bounded sample
  e2e           same metric through Translator.translate() with pinned HOST buffers
``--impl reference`` times the CPU oracle port of the reference's translate path on the host cores.
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

import torch  # noqa: E402

METRIC = "basecalled_bases_per_sec"
UNIT = "bases/s"
WORKLOAD = "l2t_greedy_b1024"       # BASELINE.json configs[1]: the default and the contract's bench line
B_PER_GPU, T, L = 1024, 512, 100
# the other BASELINE.json configs (parity-test cases; measured with --workload for the record, profiles/)
WORKLOADS = {
    # name: (family, config kwargs, beam size, model description, roofline kernel category)
    "l2t_greedy_b1024": ("l2t", {}, 1, "NanoDecoder LSTM2Transformer enc3(biLSTM)+dec3, d=256, heads=8, ff=2048", "cross_attn"),
    "l2t_beam5_b1024": ("l2t", {}, 5, "NanoDecoder LSTM2Transformer enc3(biLSTM)+dec3, d=256, --fast beam 5", "cross_attn"),
    "t2t_greedy_b1024": ("t2t", {}, 1, "Transformer2Transformer 3+3, d=256, heads=8, ff=2048", "cross_attn"),
    "t2t512_greedy_b1024": ("t2t", dict(d_model=512, enc_layers=6, dec_layers=6), 1,
                            "Transformer2Transformer 6+6, d=512, heads=8, ff=2048", "cross_attn"),
    "nano2rnn_greedy_b1024": ("nano2rnn", {}, 1, "NanoDecoder LSTM enc3 -> InputFeed LSTM dec3, mlp attention, d=256", "mlp_attn"),
    "brnn2rnn_greedy_b1024": ("brnn2rnn", {}, 1, "brnn enc3 -> InputFeed LSTM dec3, mlp attention, d=256", "mlp_attn"),
    "cnn2cnn_greedy_b1024": ("cnn2cnn", {}, 1, "Conv2Conv 3+3, kernel width 3, d=256", "mlp_attn"),
}
FAMILY, FAMILY_KW, BEAM, MODEL_DESC, ROOF_CAT = WORKLOADS[WORKLOAD]


def min_length():
    """Beam workloads: with random-init weights the best hypothesis is "</s>" at step 0 (every further token only
    lowers the cumulative log-prob), every chunk would retire at once and the metric would be 0 bases/s.
    min_length = max_length - 1 (the reference's own -min_length flag) keeps all beams alive for the whole loop."""
    return L - 1 if BEAM > 1 else 0


def select_workload(name):
    global WORKLOAD, FAMILY, FAMILY_KW, BEAM, MODEL_DESC, ROOF_CAT
    WORKLOAD = name
    FAMILY, FAMILY_KW, BEAM, MODEL_DESC, ROOF_CAT = WORKLOADS[name]


def model_and_weights():
    from nanodecoder_b200 import synth
    from nanodecoder_b200.config import ModelConfig
    cfg = ModelConfig.family(FAMILY, **FAMILY_KW)
    return cfg, synth.make_state_dict(cfg, seed=2025)


def base_config(n_gpus):
    return {"workload": WORKLOAD, "model": MODEL_DESC,
            "decode": "greedy" if BEAM == 1 else "--fast beam %d" % BEAM, "chunks_per_step_per_gpu": B_PER_GPU,
            "chunk_len": T, "max_length": L, "min_length": min_length(),
            "global_chunks_per_step": B_PER_GPU * n_gpus, "parallelism": "read-sharded x%d" % n_gpus,
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
def cpu_sample(n_chunks, steps, warmup, target_s=None):
    """Oracle port (CPU restatement of the reference's translate path) on a bounded sample.  With
    ``target_s`` the step count is chosen from the warm-up step's duration so the sample takes about that
    many seconds of CPU work (never fewer than ``steps``, never more than 40)."""
    from nanodecoder_b200 import synth
    from oracle import decode as od
    from oracle.model import OracleModel
    cfg, sd = model_and_weights()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    chunks, lengths = synth.make_chunks(n_chunks, T=T, seed=1234, ragged=True, read_len=16)
    order = torch.argsort(lengths, descending=True, stable=True)
    src = chunks[order].t().contiguous().unsqueeze(2)
    lengths = lengths[order]
    om = OracleModel(sd, cfg)
    times, bases = [], 0
    i = 0
    while i < warmup + steps:
        t0 = time.perf_counter()
        if BEAM == 1:
            out = od.greedy(om, src, lengths, max_length=L)
            n_bases = od.count_bases(out["predictions"])
        else:
            out = od.beam_fast(om, src, lengths, beam_size=BEAM, max_length=L, min_length=min_length())
            n_bases = sum(int((p[0] != 3).sum()) for p in out["predictions"])
        dt = time.perf_counter() - t0
        if target_s and i == 0:
            steps = max(steps, min(40, int(target_s / max(dt, 1e-3))))
        i += 1
        if i > warmup:
            times.append(dt)
            bases += n_bases
    total = sum(times)
    return {"value": bases / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d chunks x %d steps of the same workload (oracle/ CPU port of translate_batch, fp32, "
                      "torch %s), %.1f s" % (n_chunks, steps, torch.__version__, total),
            "chunks_per_s": n_chunks * steps / total, "ms_per_step": 1e3 * total / steps}


def run_reference(args):
    select_workload(getattr(args, "workload", WORKLOAD))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_chunks = 50                                  # the reference README's batch size (README.md:26)
    r = cpu_sample(n_chunks, args.steps, min(args.warmup, 1))
    cfg = base_config(1)
    cfg["chunks_per_step_per_gpu"] = n_chunks
    cfg["global_chunks_per_step"] = n_chunks
    cfg["parallelism"] = "cpu x%d threads" % r["cores"]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gemm-mode", default="3xtf32", choices=["3xtf32", "tf32", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=WORKLOAD, choices=sorted(WORKLOADS),
                    help="default = BASELINE.json configs[1]; the others are the remaining BASELINE configs")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from nanodecoder_b200 import synth
    from nanodecoder_b200.engine import Engine
    from nanodecoder_b200.opts import default_translate_opt
    from nanodecoder_b200.translate.translator import Translator, _Field, count_bases
    from nanodecoder_b200.checkpoint import Vocab

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

    cfg, sd = model_and_weights()
    B = B_PER_GPU
    eng = Engine(cfg, sd, max_batch=B, max_src_len=T, max_tgt_len=L, max_beam=BEAM, gemm_mode=args.gemm_mode,
                 device=local)
    # every rank decodes its own shard of reads (different seed -> different chunks)
    chunks, lengths = synth.make_chunks(B, T=T, seed=1234 + rank, ragged=True, read_len=16)
    order = torch.argsort(lengths, descending=True, stable=True)
    chunks, lengths = chunks[order].contiguous(), lengths[order].contiguous()
    src_d, len_d = chunks.to(dev), lengths.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        eng.encode(src_d, len_d)
        if BEAM == 1:
            return eng.decode_greedy(L)["ids"]
        return eng.decode_beam(BEAM, 1, L, min_len=min_length())["ids"][:, 0, :]   # best hypothesis, -1 padded after </s>

    # the clock sampler runs from before the warm-up (no idle gap in front of the timed region: the GPU
    # would drop its clocks and the first timed step would pay the ramp); only samples taken inside the
    # timed region are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def count_bases_dev(ids):
        is_eos = ids.eq(3) | ids.lt(0)
        first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full_like(ids[:, 0], L))
        return first.sum()

    bases_t = torch.zeros((), dtype=torch.int64, device=dev)
    for _ in range(args.warmup):
        ids = step_device()
        bases_t += count_bases_dev(ids)       # also loads torch's lazily-loaded kernels before the timed region
    torch.cuda.synchronize()

    # ---------------- timed region 1: inputs resident in HBM
    eng.reset_launch_count()
    barrier()
    t_region0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bases_t.zero_()
    trace = [] if os.environ.get("ND_BENCH_TRACE") else None
    ev0.record()
    for _ in range(args.steps):
        ids = step_device()
        if trace is not None:
            trace.append(torch.cuda.Event(enable_timing=True))
            trace[-1].record()
        bases_t += count_bases_dev(ids)
    ev1.record()
    barrier()
    t_region1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    if trace:
        prev = ev0
        for i, ev in enumerate(trace):
            sys.stderr.write("step %d: %.2f ms\n" % (i, prev.elapsed_time(ev)))
            prev = ev
    launches = eng.launch_count
    clocks = sampler.stop((t_region0, t_region1)) if rank == 0 else None
    # ---------------- per-kernel pass for the roofline entry: same steps with CUDA-event brackets around
    # every launch of the dominant kernel (the engine then runs the decode loop on one stream, eagerly)
    eng.profile_enable([ROOF_CAT])
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(max(1, min(args.steps, 3))):
        step_device()
    ev3.record()
    torch.cuda.synchronize()
    prof_ms = ev2.elapsed_time(ev3)
    prof = eng.profile_read()
    eng.profile_enable([])

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(bases_t.item()), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # slowest rank defines the step time
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)      # per-rank record counts gathered over NCCL
    ms_max = float(t.item())
    bases_total = float(tot[0].item())
    value = bases_total / (ms_max / 1e3)

    # ---------------- timed region 2: end to end through the public API with HOST buffers
    opt = default_translate_opt(beam_size=BEAM, fast=BEAM > 1, batch_size=B, max_length=L, min_length=min_length(),
                                src_seq_length=T, gpu=local, gemm_mode=args.gemm_mode)
    tr = Translator(eng, {"tgt": _Field(Vocab(cfg.vocab))}, opt, cfg)
    h_chunks, h_len = chunks.pin_memory(), lengths
    tr.translate(src=(h_chunks, h_len), batch_size=B)               # warm
    barrier()
    t0 = time.perf_counter()
    e2e_bases = 0
    for _ in range(args.steps):
        _, preds = tr.translate(src=(h_chunks, h_len), batch_size=B)
        e2e_bases += sum((p[0].count(" ") + 1) if p[0] else 0 for p in preds)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    be = torch.tensor([float(e2e_bases)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(be, op=dist.ReduceOp.SUM)
    e2e_value = float(be.item()) / float(te.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        ca_ms, ca_n = prof.get(ROOF_CAT, (0.0, 0))
        d = cfg.d_model
        # algorithmic bytes per launch: K and V (or uh and H, or keys and values of the conv attention) of every
        # chunk once (2*T*d*4) + query in + context out per row; beams of a chunk share the K/V pass
        bytes_per_launch = B * (2 * T * d * 4 + 2 * BEAM * d * 4)
        achieved = bytes_per_launch / (ca_ms / max(ca_n, 1) * 1e-3) / 1e9 if ca_n else None
        traffic = None
        try:
            if WORKLOAD == "l2t_greedy_b1024":
                traffic = json.load(open(os.path.join(ROOT, "profiles", "cross_attn_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        ca_name = ("cross_attn_kernel<%d,1>" % (d // 32) if BEAM == 1 else
                   ("cross_attn_ring_kernel<%d,%d>" % ((4 if BEAM <= 4 else 5), 2) if d == 256 and BEAM <= 5 else
                    "cross_attn_ring_kernel<8,1>" if d == 256 else "cross_attn_kernel<%d,8>" % (d // 32)))
        kname = {"cross_attn": ca_name + " (decode cross-attention, fp32 K/V)",
                 "mlp_attn": "mlp_attn_kernel<%d,1> (decode global / conv attention, fp32)" % (d // 32)}[ROOF_CAT]
        roofline = {"kernel": kname, "bound": "hbm",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "launches_timed": ca_n, "avg_launch_us": 1e3 * ca_ms / max(ca_n, 1),
                    "share_of_step": ca_ms / prof_ms if prof_ms else None,
                    "timing": "CUDA events around each launch on its stream, profiled pass of %d steps "
                              "(1 decode stream, no graph)" % max(1, min(args.steps, 3))}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tcgen05 3xTF32 GEMMs, fp32 attention)"
                if args.gemm_mode == "3xtf32" else "f32/" + args.gemm_mode, "data": "synthetic",
                "config": base_config(world), "clocks": clocks,
                "chunks_per_s": B * world * args.steps / (ms_max / 1e3),
                "token_steps_per_s": B * world * L * args.steps / (ms_max / 1e3),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * T * 4 + B * 8),
                        "d2h_bytes_per_step": int(B * L * 8 + B * 4),
                        "api": "Translator.translate(src=(pinned chunks, lengths), batch_size=1024) -> base strings"},
                "n_model_params": int(sum(v.numel() for v in sd.values() if torch.is_floating_point(v))),
                "gpu_launches": int(tot[1].item()), "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_sample(50, 2, 1, target_s=15.0).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
