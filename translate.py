#!/usr/bin/env python
"""Basecall a directory of reads — same command line as the reference's translate.py
(translate.py:59-187): -model M -src_dir D -save_data S [--fast] [-beam_size K] ...

Differences in mechanism only: reads are pooled, normalised and chunked on the GPU
(no multiprocessing.Pool / text round trip), decoded by the CUDA engine, and the per-read
result/<read>.fasta, segment/<read>.txt and speed.txt files are written in the same formats.
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nanodecoder_b200 import opts                                        # noqa: E402


def write_read_files(opt, file_src, all_predictions):
    """result/<read>.fasta and segment/<read>.txt of one read (translate.py:81-93, same formats) -> assembled length."""
    from nanodecoder_b200.utils.labelop import index2base, simple_assembly
    name = file_src.split(".txt")[0]
    if opt.src_seq_stride < opt.src_seq_length:
        c_bpread = index2base(np.argmax(simple_assembly(all_predictions), axis=0))
    else:
        c_bpread = simple_assembly(all_predictions, flag_intersection=False)
    with open(os.path.join(opt.save_data, "result", name + ".fasta"), "w") as f:
        f.writelines(">%s\n%s" % (name, c_bpread))
    with open(os.path.join(opt.save_data, "segment", file_src), "w+") as f:
        for n_best_preds in all_predictions:
            f.write("\n".join(n_best_preds) + "\n")
    return len(c_bpread)


def speed_line(file_src, time_translate, n_bases):
    """one line of speed.txt (translate.py:94-96)."""
    return "%s\t%0.2f\t%d\t%0.2f\n" % (file_src.split(".txt")[0], float(time_translate), n_bases,
                                        n_bases / float(max(time_translate, 1e-9)))


def write_output(opt, file_src, all_predictions, time_translate):
    """translate.py:81-98 (same file formats)."""
    n_bases = write_read_files(opt, file_src, all_predictions)
    with open(os.path.join(opt.save_data, "speed.txt"), "a+") as f:
        f.writelines(speed_line(file_src, time_translate, n_bases))


def run_reads(opt, todo_mine, read_signal, frontend, translator):
    """The read loop of one rank.  todo_mine: [(global read index, (file name, suffix, out name))]; read_signal(path,
    suffix) -> int16 samples; frontend(list of reads) -> (chunks, lengths, chunk_read); translator.translate as in the
    reference.  While the GPU works on a group of reads, `opt.thread` host threads parse the files of the NEXT group and
    assemble / write the per-read files of the PREVIOUS ones (file I/O and the libnanodec calls release the GIL).
    -> [(global read index, speed.txt line)] of the reads written."""
    from concurrent.futures import ThreadPoolExecutor
    from nanodecoder_b200.inputters.nano_dataset import reference_pad_lengths
    # reads pooled per GPU front-end launch; -attn_debug writes attention/<read>.txt per read (translate.py:110-111
    # of the reference opens the file and calls setAttnFile before every translate call), so reads go one by one
    pool_reads = 1 if opt.attn_debug else max(1, opt.thread * 8)
    groups = [todo_mine[g0: g0 + pool_reads] for g0 in range(0, len(todo_mine), pool_reads)]

    def load_one(path, suffix, name):
        # a read that cannot be loaded (corrupt file, unsupported content) is reported and skipped like the
        # reference does per read (its pool.apply_async swallows the worker's exception, translate.py:154-161;
        # the writer prints '!!!error!!!', :97-98) -- it must not take the run (or, under torchrun, the other
        # ranks waiting in the final gather) down
        try:
            return read_signal(path, suffix)
        except Exception as e:                                   # noqa: BLE001
            print("!!!error!!!data src: %s (%s: %s)" % (name.split(".txt")[0], type(e).__name__, e))
            return np.zeros((0,), dtype=np.int16)

    def write_one(idx, out_name, preds, seconds):
        try:
            return idx, speed_line(out_name, seconds, write_read_files(opt, out_name, preds))
        except Exception:                                        # translate.py:97-98
            print("!!!error!!!data src: " + out_name.split(".txt")[0])
            return None

    writes = []
    with ThreadPoolExecutor(max_workers=max(1, opt.thread)) as ex:
        def load(group):
            return [ex.submit(load_one, os.path.join(opt.src_dir, fn), suffix, out) for _, (fn, suffix, out) in group]

        nxt = load(groups[0]) if groups else None
        for gi, group in enumerate(groups):
            start = time.time()
            futs, nxt = nxt, (load(groups[gi + 1]) if gi + 1 < len(groups) else None)
            reads = [f.result() for f in futs]
            keep = [i for i, r in enumerate(reads) if r.size > 0]
            if not keep:
                continue
            chunks, lengths, chunk_read = frontend([reads[i] for i in keep])
            # chunks of many reads share GPU batches; each keeps the padding width of its read-by-read reference batch
            h_len = lengths.cpu().numpy()
            pad_to = np.empty_like(h_len)
            sels = [np.nonzero(chunk_read == j)[0] for j in range(len(keep))]
            for sel in sels:
                pad_to[sel] = reference_pad_lengths(h_len[sel], opt.batch_size)
            attn_file = None
            if opt.attn_debug:
                attn_file = open(os.path.join(opt.save_data, "attention", group[keep[0]][1][2]), "w")
                translator.setAttnFile(attn_file)
            try:
                _, preds = translator.translate(src=(chunks, lengths, pad_to), tgt=None, src_dir=opt.save_data,
                                                batch_size=opt.batch_size, attn_debug=opt.attn_debug)
            finally:
                if attn_file is not None:
                    translator.setAttnFile(None)
                    attn_file.close()
            elapsed = time.time() - start
            total = max(1, len(chunk_read))
            for j, i in enumerate(keep):
                sel = sels[j]
                writes.append(ex.submit(write_one, group[i][0], group[i][1][2], [preds[k] for k in sel],
                                        elapsed * len(sel) / total))
        lines = [w.result() for w in writes]
    return sorted((l for l in lines if l is not None), key=lambda l: l[0])


def finish_lines(opt, lines):
    """speed.txt: one process appends its lines in global read order; under torchrun the lines of all ranks travel to
    rank 0 (the only collective of a run), which appends them in global read order.  The per-read files were written
    by the rank that decoded the read (independent files; assembling everything on rank 0 would make one host core
    the bottleneck of 8 GPUs)."""
    from nanodecoder_b200 import shard
    merged = shard.gather_records(lines, dst=0)
    if merged is not None:
        with open(os.path.join(opt.save_data, "speed.txt"), "a+") as f:
            f.writelines(line for _, line in merged)


def finish_records(opt, records):
    """records of THIS rank: (global read index, out name, predictions, seconds) -> files (synchronous form of what
    run_reads does with its thread pool)."""
    lines = []
    for idx, out_name, preds, seconds in sorted(records, key=lambda r: r[0]):
        try:
            lines.append((idx, speed_line(out_name, seconds, write_read_files(opt, out_name, preds))))
        except Exception:                                        # translate.py:97-98
            print("!!!error!!!data src: " + out_name.split(".txt")[0])
    finish_lines(opt, lines)


def list_reads(opt):
    """-> (todo [(file name, suffix, out name)], sizes, number of reads that already have a result).
    One entry per file like the reference (translate.py:136-152), except that a MULTI-read .fast5 file (which the
    reference cannot read) contributes one entry per read: suffix "fast5:<member name>", outputs named by the read id."""
    from nanodecoder_b200.utils.labelop import list_fast5_reads
    todo, sizes, done = [], [], 0

    def add(fn, suffix, out, size):
        nonlocal done
        if os.path.exists(os.path.join(opt.save_data, "result", out.split(".txt")[0] + ".fasta")):
            done += 1                                            # translate.py:152 (resume semantics)
        else:
            todo.append((fn, suffix, out))
            sizes.append(size)

    for fn in sorted(os.listdir(opt.src_dir)):
        for suffix in ("fast5", "signal"):
            if fn.endswith(suffix):
                path = os.path.join(opt.src_dir, fn)
                size = os.path.getsize(path)
                members = None
                if suffix == "fast5":
                    try:
                        layout, names = list_fast5_reads(path)
                        members = names if layout == 2 else None
                    except Exception:                            # noqa: BLE001 -- reported per read when it is loaded
                        members = None
                if members:
                    for name in members:
                        rid = name[5:] if name.startswith("read_") else name
                        add(fn, "fast5:" + name, rid + ".txt", size // len(members))
                else:
                    add(fn, suffix, fn.split("." + suffix)[0] + ".txt", size)     # translate.py:138-142 of the reference
    return todo, sizes, done


def main(opt, logger):
    from nanodecoder_b200.inputters.nano_dataset import SignalFrontend
    from nanodecoder_b200.translate.translator import build_translator
    from nanodecoder_b200.utils.labelop import read_raw_signal
    from nanodecoder_b200 import shard
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        opt.gpu = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(opt.gpu)
        dist.init_process_group("nccl", device_id=torch.device("cuda", opt.gpu))
    for sub in ("", "result", "segment") + (("attention",) if opt.attn_debug else ()):
        os.makedirs(os.path.join(opt.save_data, sub), exist_ok=True)
    opt.tgt = None
    opt.data_type = "nano"
    translator = build_translator(opt, report_score=False, logger=logger)
    frontend = SignalFrontend(translator.model, opt.normalization_raw, opt.src_seq_length, opt.src_seq_stride)

    rank, ws = shard.world()
    todo, sizes, done = list_reads(opt) if rank == 0 else ([], [], 0)
    # ONE listing for the whole job: rank 0 lists the directory (and applies the "already has result/<read>.fasta"
    # resume rule) and broadcasts the table.  Ranks listing on their own race with faster ranks that already write
    # .fasta files: the tables differ, the index-based partitions diverge and reads are skipped or decoded twice.
    todo, sizes, done = shard.broadcast_object((todo, sizes, done), src=0)
    logger.info("%d reads have already translated, remains %d read\n" % (done, len(todo)))
    # reads shard across the ranks of a torchrun launch (one process per GPU, no collective inside the step)
    mine = shard.partition_reads(sizes, ws)[rank]
    lines = []
    try:
        lines = run_reads(opt, [(i, todo[i]) for i in mine], read_raw_signal, frontend, translator)
    finally:
        finish_lines(opt, lines)                                 # every rank always enters the final gather
        if ws > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="translate.py", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    opts.config_opts(parser)                                     # -config FILE.yml / -save_config FILE.yml
    opts.translate_opts(parser)
    opt = opts.parse_args(parser)
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s %(levelname)s] %(message)s")
    log = logging.getLogger("translate")
    if opt.log_file:
        log.addHandler(logging.FileHandler(opt.log_file))
    main(opt, log)
