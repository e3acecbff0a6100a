"""Read-wise sharding of the translate path over the GPUs of one box.

Reads (and the chunks inside a read) are independent: normalisation is per read
(utils/labelop.py:219-223), the model keeps no cross-chunk state and assembly is per read
(translate.py:84-87).  So each rank runs its own engine on its own reads, with NO collective inside the
step; torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is used only
  * to gather the per-read FASTA records and (read, seconds, bases) timing rows on rank 0, and
  * to reduce the per-rank counters / the slowest rank's time for the throughput report.
The reference gets its read-level parallelism from a multiprocessing.Pool feeding ONE translator
(translate.py:136-161); this module is its multi-GPU counterpart.
"""
from __future__ import annotations

import pickle
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_reads(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Deterministic longest-processing-time assignment of reads to ranks: reads sorted by size
    (descending, index as tie break), each given to the currently lightest rank (lowest rank on ties).
    Every rank computes the same table from the same sizes, so no communication is needed.
    -> per-rank list of read indices, each in ascending index order."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(x) for x in out]


def chunk_range(n_chunks: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous chunk range [lo, hi) of `rank` for workloads that are one long list of chunks
    (the synthetic 1M-chunk configuration)."""
    lo = n_chunks * rank // world_size
    hi = n_chunks * (rank + 1) // world_size
    return lo, hi


def _comm_device(group=None) -> torch.device:
    backend = dist.get_backend(group)
    if backend == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_bytes(payload: bytes, dst: int = 0, group=None) -> Optional[List[bytes]]:
    """Gather one variable-length byte string per rank on `dst` (others get None).
    Two collectives: all_gather of the sizes, then all_gather of the payloads padded to the largest
    (a few bytes per base: negligible next to the decode step)."""
    rank, ws = world()
    if ws == 1:
        return [payload]
    dev = _comm_device(group)
    n = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros((cap,), dtype=torch.uint8, device=dev)
    if payload:
        buf[: len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    bufs = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(bufs, buf, group=group)
    if rank != dst:
        return None
    return [bytes(bufs[r][: sizes[r]].cpu().numpy().tobytes()) for r in range(ws)]


def broadcast_object(obj, src: int = 0, group=None):
    """The same picklable object on every rank (the job's read table, built once on `src`)."""
    rank, ws = world()
    if ws == 1:
        return obj
    box = [obj if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group, device=_comm_device(group))
    return box[0]


def gather_records(records: Sequence[tuple], dst: int = 0, group=None) -> Optional[List[tuple]]:
    """Per-read result records, e.g. (read_index, name, fasta_text, segment_lines, seconds, n_bases), from
    every rank -> one list on `dst`, sorted by the first field (the global read index)."""
    parts = gather_bytes(pickle.dumps(list(records), protocol=4), dst=dst, group=group)
    if parts is None:
        return None
    merged: List[tuple] = []
    for p in parts:
        merged.extend(pickle.loads(p))
    merged.sort(key=lambda r: r[0])
    return merged


def reduce_throughput(units: float, seconds: float, group=None) -> Tuple[float, float]:
    """(total units over all ranks, slowest rank's seconds): whole-job throughput = units / seconds."""
    rank, ws = world()
    if ws == 1:
        return float(units), float(seconds)
    dev = _comm_device(group)
    u = torch.tensor([float(units)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(u.item()), float(t.item())
