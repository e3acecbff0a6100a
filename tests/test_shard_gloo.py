"""Read-sharded multi-rank path on CPU: world_size 2, gloo backend.  The GPU step itself has
no collective; what is tested here is everything the N>1 path adds: the partition every rank derives on its
own, the gather of per-read records on rank 0 and the throughput reduction."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanodecoder_b200 import shard


def test_partition_is_deterministic_balanced_and_complete():
    sizes = [2000 + (i * 7919) % 190000 for i in range(101)]
    for ws in (1, 2, 4, 8):
        parts = shard.partition_reads(sizes, ws)
        assert parts == shard.partition_reads(list(sizes), ws)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)          # LPT bound
    assert shard.partition_reads([], 2) == [[], []]
    assert shard.partition_reads([5], 2) == [[0], []]
    with pytest.raises(ValueError):
        shard.partition_reads([1], 0)


def test_chunk_range_covers_everything():
    for n in (0, 1, 7, 1000003):
        for ws in (1, 2, 8):
            spans = [shard.chunk_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))


def _worker(rank, ws, init_file, out_dir):
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=ws)
    try:
        sizes = [100, 900, 400, 50, 650, 10, 300]
        mine = shard.partition_reads(sizes, ws)[rank]
        # fake "basecalls": one record per read, ragged payloads, rank 1 also sends an empty string
        recs = [(i, "read%d" % i, "ACGT"[i % 4] * sizes[i], float(sizes[i]) / 1000.0, sizes[i]) for i in mine]
        merged = shard.gather_records(recs, dst=0)
        raw = shard.gather_bytes(b"" if rank == 1 else b"\x00\x01rank0", dst=0)
        units, secs = shard.reduce_throughput(sum(sizes[i] for i in mine), 1.0 + rank)
        # the job's read table exists once (rank 0's listing): ranks never partition from their own directory listing
        table = shard.broadcast_object(([("a.signal", "signal", "a.txt")], [7], 3) if rank == 0 else None, src=0)
        assert table == ([("a.signal", "signal", "a.txt")], [7], 3)
        if rank == 0:
            assert [r[0] for r in merged] == list(range(len(sizes)))
            assert all(r[2] == "ACGT"[r[0] % 4] * sizes[r[0]] for r in merged)
            assert raw == [b"\x00\x01rank0", b""]
        else:
            assert merged is None and raw is None
        assert units == float(sum(sizes)) and secs == float(ws)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_records_and_reduce_world_size_2(tmp_path):
    ws = 2
    init_file = str(tmp_path / "rendezvous")
    mp.spawn(_worker, args=(ws, init_file, str(tmp_path)), nprocs=ws, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(ws))


def test_single_process_paths_need_no_process_group():
    assert shard.world() == (0, 1)
    assert shard.gather_bytes(b"abc") == [b"abc"]
    assert shard.gather_records([(1, "b"), (0, "a")]) == [(0, "a"), (1, "b")]
    assert shard.reduce_throughput(10, 2.0) == (10.0, 2.0)
    assert shard.broadcast_object({"x": 1}) == {"x": 1}
