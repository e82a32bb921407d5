"""Sentence sharding across the GPUs of one box: rows are independent, so every rank encodes a
contiguous range of rows of (nearly) equal BYTE size with a full replica of the tables, and nothing is exchanged on the
data path -- the per-rank ragged results are concatenated on the host with their row_splits rebased."""
import numpy as np


def shard_rows(row_offsets, world_size):
    """-> list of (row_lo, row_hi) per rank: contiguous row ranges balanced by bytes (not by row count)"""
    off = np.asarray(row_offsets, dtype=np.int64)
    n = off.size - 1
    total = int(off[-1] - off[0])
    cuts = [0]
    for r in range(1, world_size):
        target = off[0] + (total * r) // world_size
        i = int(np.searchsorted(off, target, side='left'))
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def chunk_rows(row_offsets, chunk_bytes):
    """-> list of (row_lo, row_hi): contiguous row ranges of about chunk_bytes each for a copy / compute / copy pipeline.
    (Shorter chunks at both ends -- to start computing earlier and drain less -- were measured and do not pay: the
    pipelined paths sit at the rate of the host link with both directions busy, 23.7 ms per GiB of BPE input.)"""
    off = np.asarray(row_offsets, dtype=np.int64)
    total = int(off[-1] - off[0])
    k = max(1, (total + chunk_bytes - 1) // chunk_bytes)
    return [r for r in shard_rows(off, k) if r[1] > r[0]] or [(0, 0)]


def take_shard(data, row_offsets, lo, hi):
    """rows [lo, hi) as their own (bytes, offsets-from-zero) pair"""
    off = np.asarray(row_offsets, dtype=np.int64)
    return data[off[lo]:off[hi]], off[lo:hi + 1] - off[lo]


def concat_ragged(parts):
    """[(values, splits)] per rank, in rank order -> (values, splits) of the whole batch"""
    vals, splits, base = [], [np.zeros(1, dtype=np.int64)], 0
    for v, s in parts:
        v = np.asarray(v)
        s = np.asarray(s, dtype=np.int64)
        vals.append(v)
        splits.append(s[1:] + base)
        base += int(s[-1])
    return (np.concatenate(vals) if vals else np.zeros(0, dtype=np.int32)), np.concatenate(splits)


def gather_ragged(values, splits, group=None):
    """host-side gather of every rank's ragged result on all ranks (torch.distributed, any backend that moves
    Python objects: gloo or nccl); returns the concatenated (values, splits)"""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(values), np.asarray(splits)), group=group)
    return concat_ragged(parts)


def encode_sharded(encode_fn, data, row_offsets, group=None, gather=True):
    """The N > 1 entry point: every rank of the process group (one process per GPU, torchrun) calls this with the SAME
    batch (`data` uint8, `row_offsets` int64, host arrays); rank r encodes the rows shard_rows(...)[r] with
    `encode_fn(shard_data, shard_offsets) -> (values, splits)` -- e.g. `aksharTokenizer.encode_batch_host` on this rank's
    GPU -- and, with gather=True, every rank gets the ragged result of the whole batch (host-side gather, nothing is
    exchanged on the data path).  gather=False returns (row_lo, row_hi, values, splits) of this rank's shard."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_rows(row_offsets, world)[rank]
    d, o = take_shard(np.asarray(data), row_offsets, lo, hi)
    values, splits = encode_fn(d, o)
    if not gather:
        return lo, hi, values, splits
    if world == 1:
        return np.asarray(values), np.asarray(splits, dtype=np.int64)
    return gather_ragged(values, splits, group)
