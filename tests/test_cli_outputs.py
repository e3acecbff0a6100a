"""CPU: the output side of the translate.py CLI (reference translate.py:81-98): per-read result/<read>.fasta,
segment/<read>.txt and speed.txt, written by one process and by two ranks (gloo) of a read-sharded run."""
import importlib.util
import os
import types

import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli():
    spec = importlib.util.spec_from_file_location("nd_translate_cli", os.path.join(ROOT, "translate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _opt(save, stride=512, length=512):
    for sub in ("", "result", "segment"):
        os.makedirs(os.path.join(save, sub), exist_ok=True)
    return types.SimpleNamespace(save_data=save, src_seq_stride=stride, src_seq_length=length)


def _records():
    # (global read index, out name, predictions per chunk (n_best lists), seconds)
    return [(0, "readA.txt", [["A C G T A C"], ["T A C G G"]], 0.5),
            (1, "readB.txt", [["G G G"], [""], ["T T"]], 0.25),
            (2, "readC.txt", [["A C G T T G C A"], ["T T G C A A A"]], 1.0),
            (3, "readD.txt", [["C"]], 0.125)]


def _read(path):
    with open(path) as f:
        return f.read()


def test_single_process_files_have_the_reference_formats(tmp_path):
    cli = _cli()
    opt = _opt(str(tmp_path))
    cli.finish_records(opt, list(reversed(_records())))               # written in global read order anyway
    assert _read(tmp_path / "result" / "readA.fasta") == ">readA\nACGTACTACGG"      # stride == length: concatenation
    assert _read(tmp_path / "result" / "readB.fasta") == ">readB\nGGGTT"
    assert _read(tmp_path / "segment" / "readB.txt") == "G G G\n\nT T\n"
    speed = _read(tmp_path / "speed.txt").splitlines()
    assert speed == ["readA\t0.50\t11\t22.00", "readB\t0.25\t5\t20.00", "readC\t1.00\t15\t15.00", "readD\t0.12\t1\t8.00"]


def test_overlapping_windows_go_through_the_assembly(tmp_path):
    cli = _cli()
    opt = _opt(str(tmp_path), stride=60, length=300)
    cli.finish_records(opt, _records()[2:3])
    # chunks ACGTTGCA and TTGCAAA overlap on TTGCA: consensus ACGTTGCAAA
    assert _read(tmp_path / "result" / "readC.fasta") == ">readC\nACGTTGCAAA"
    assert _read(tmp_path / "speed.txt") == "readC\t1.00\t10\t10.00\n"


def _worker(rank, ws, init_file, save):
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(ws)
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=ws)
    try:
        cli = _cli()
        recs = [r for r in _records() if r[0] % ws == (1 - rank)]     # rank 0 holds reads 1, 3; rank 1 holds 0, 2
        cli.finish_records(_opt(save), recs)
    finally:
        dist.destroy_process_group()


def test_two_ranks_write_their_own_reads_and_rank0_the_speed_file(tmp_path):
    ws = 2
    mp.spawn(_worker, args=(ws, str(tmp_path / "rdv"), str(tmp_path / "out")), nprocs=ws, join=True)
    ref = tmp_path / "ref"
    _cli().finish_records(_opt(str(ref)), _records())
    for sub, names in (("result", ["readA.fasta", "readB.fasta", "readC.fasta", "readD.fasta"]),
                       ("segment", ["readA.txt", "readB.txt", "readC.txt", "readD.txt"])):
        for n in names:
            assert _read(tmp_path / "out" / sub / n) == _read(ref / sub / n)
    assert _read(tmp_path / "out" / "speed.txt") == _read(ref / "speed.txt")      # global read order, each read once


import pytest


@pytest.mark.parametrize("threads", [1, 3])
def test_read_loop_with_prefetch_and_write_behind_threads(tmp_path, threads):
    """run_reads (the CLI's per-rank loop) end to end on CPU: real `.signal` parsing, the reference's chunk table, the
    real Translator mirror over a recording stub engine; host threads prefetch the next group's files and write the
    previous reads.  Every non-empty read gets its files, chunks keep their order and their read's padding width."""
    import numpy as np
    import torch
    from nanodecoder_b200.inputters.nano_dataset import chunk_table, reference_pad_lengths
    from nanodecoder_b200.utils.labelop import read_raw_signal
    from test_translator_host import EOS, VOCAB, _tokens, _translator

    cli = _cli()
    T, L, B = 64, 12, 8
    src = tmp_path / "reads"
    src.mkdir()
    rng = np.random.default_rng(11)
    sizes = [200, 0, 64, 333, 65, 700, 1, 129, 500, 90, 260]            # one empty read, several ragged tails
    raws = []
    import h5_writer
    fast5 = {3: "old", 5: "new", 8: "old"}                               # three reads arrive as .fast5 (HDF5) files
    for i, n in enumerate(sizes):
        x = rng.integers(300, 900, size=n)
        if i in fast5:
            (src / ("r%02d.fast5" % i)).write_bytes(h5_writer.make_fast5(
                x.astype(np.int16), read_name="Read_%d" % (100 + i), flavour=fast5[i], chunk=128 if fast5[i] == "old" else None))
        else:
            (src / ("r%02d.signal" % i)).write_text(" ".join(map(str, x)))
        raws.append(x.astype(np.int16))
    (src / "r11.signal").write_text("512 garbage 7")                     # a corrupt read: reported and skipped, per read
    (src / "r12.fast5").write_bytes(b"\x89HDF\r\n\x1a\n" + bytes(200))     # a truncated HDF5 file: same
    sizes += [0, 0]
    suffix = {i: ("fast5" if i in fast5 or i == 12 else "signal") for i in range(len(sizes))}
    raws += [np.zeros(0, np.int16)] * 2                                  # the two corrupt reads
    multi = {}                                                           # two more reads inside ONE multi-read file
    for i, n in ((13, 150), (14, 410)):
        x = rng.integers(300, 900, size=n).astype(np.int16)
        multi["r%02d" % i] = x
        raws.append(x)
        sizes.append(n)
        suffix[i] = "fast5:read_r%02d" % i
    (src / "zz_multi.fast5").write_bytes(h5_writer.make_multi_fast5(multi, chunk=100))
    opt = _opt(str(tmp_path / "out"), stride=T, length=T)
    opt.src_dir, opt.thread, opt.batch_size, opt.attn_debug = str(src), threads, B, False   # groups of 8 * threads reads

    def frontend(reads):                                                 # chunking only: samples / 1024 as "signal"
        cr, cs = chunk_table([r.size for r in reads], T, T)
        chunks = np.zeros((len(cr), T), np.float32)
        lens = np.zeros(len(cr), np.int64)
        for k, (r, s) in enumerate(zip(cr, cs)):
            seg = reads[r][s: s + T].astype(np.float32) / 1024.0
            chunks[k, : seg.size] = seg
            lens[k] = seg.size
        return torch.from_numpy(chunks), torch.from_numpy(lens), cr

    tr, eng = _translator(L, batch_size=B)
    todo = [(i, ("zz_multi.fast5" if ":" in suffix[i] else "r%02d.%s" % (i, suffix[i]), suffix[i], "r%02d.txt" % i))
            for i in range(len(sizes))]
    assert [(t[0], t[1]) for t in cli.list_reads(opt)[0]] == [t[1][:2] for t in todo]
    lines = cli.run_reads(opt, todo, read_raw_signal, frontend, tr)
    cli.finish_lines(opt, lines)

    live = [i for i, n in enumerate(sizes) if n > 0]
    assert [l[0] for l in lines] == live
    speed = _read(tmp_path / "out" / "speed.txt").splitlines()
    assert [s.split("\t")[0] for s in speed] == ["r%02d" % i for i in live]
    assert not (tmp_path / "out" / "result" / "r01.fasta").exists()
    for i in live:
        chunks, lens, _ = frontend([raws[i]])
        widths = reference_pad_lengths(lens.numpy(), B)
        want = []
        for k in range(len(lens)):
            ids = _tokens(chunks[k].numpy(), int(lens[k]), int(widths[k]), L)
            want.append(" ".join(VOCAB[t] for t in ids[: int((ids != EOS).sum())]))
        assert _read(tmp_path / "out" / "segment" / ("r%02d.txt" % i)) == "".join(w + "\n" for w in want)
        fasta = _read(tmp_path / "out" / "result" / ("r%02d.fasta" % i))
        assert fasta == ">r%02d\n%s" % (i, "".join(w.replace(" ", "") for w in want))
        assert int(speed[live.index(i)].split("\t")[2]) == sum(len(w.split()) for w in want)


def test_read_table_expands_multi_read_fast5(tmp_path):
    """list_reads: one entry per file as in the reference (translate.py:136-152); a multi-read .fast5 file (not readable by
    the reference) becomes one entry per read, named by the read id, with the same resume rule per read"""
    import numpy as np
    import h5_writer
    from nanodecoder_b200.utils.labelop import read_raw_signal
    cli = _cli()
    src = tmp_path / "reads"
    src.mkdir()
    rng = np.random.default_rng(5)
    sig = {k: rng.integers(300, 900, size=n).astype(np.int16) for k, n in (("a", 900), ("b", 1500), ("m1", 700), ("m2", 64),
                                                                          ("m3", 2000))}
    (src / "a.signal").write_text(" ".join(map(str, sig["a"])))
    (src / "b.fast5").write_bytes(h5_writer.make_fast5(sig["b"], read_name="Read_4", chunk=512))
    (src / "c.fast5").write_bytes(h5_writer.make_multi_fast5({"0003-m3": sig["m3"], "0001-m1": sig["m1"], "0002-m2": sig["m2"]},
                                                             chunk=256))
    (src / "d.fast5").write_bytes(b"\x89HDF\r\n\x1a\n" + bytes(100))          # corrupt: stays one entry, reported at load time
    (src / "notes.txt").write_text("ignored")
    opt = _opt(str(tmp_path / "out"))
    opt.src_dir = str(src)
    (tmp_path / "out" / "result" / "0002-m2.fasta").write_text(">0002-m2\nACGT")       # already translated: skipped
    todo, sizes, done = cli.list_reads(opt)
    assert done == 1
    assert todo == [("a.signal", "signal", "a.txt"), ("b.fast5", "fast5", "b.txt"),
                    ("c.fast5", "fast5:read_0001-m1", "0001-m1.txt"), ("c.fast5", "fast5:read_0003-m3", "0003-m3.txt"),
                    ("d.fast5", "fast5", "d.txt")]
    csize = os.path.getsize(src / "c.fast5")
    assert sizes == [os.path.getsize(src / "a.signal"), os.path.getsize(src / "b.fast5"), csize // 3, csize // 3,
                     os.path.getsize(src / "d.fast5")]
    got = [read_raw_signal(str(src / fn), suffix) for fn, suffix, _ in todo[:4]]
    for g, k in zip(got, ("a", "b", "m1", "m3")):
        assert g.dtype == np.int16 and np.array_equal(g, sig[k])
    with pytest.raises((IOError, RuntimeError)):
        read_raw_signal(str(src / "d.fast5"), "fast5")
