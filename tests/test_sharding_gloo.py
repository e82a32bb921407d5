"""The N > 1 path on the CPU: two gloo ranks shard a batch by sentence, each produces its ragged result (the oracle
stands in for the GPU here -- the test is about the sharding / gather logic), and the host-side gather reproduces the
single-process result."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import akshar_oracle as O
import synth_corpus as sc
from akshar_b200 import shard


def test_shard_rows_balances_bytes():
    off = np.array([0, 10, 10, 500, 520, 1000, 1001], dtype=np.int64)
    parts = shard.shard_rows(off, 2)
    assert parts[0][0] == 0 and parts[-1][1] == 6 and parts[0][1] == parts[1][0]
    sizes = [off[hi] - off[lo] for lo, hi in parts]
    assert abs(sizes[0] - sizes[1]) <= 500
    for w in (1, 3, 8):
        p = shard.shard_rows(off, w)
        assert p[0][0] == 0 and p[-1][1] == 6 and all(p[i][1] == p[i + 1][0] for i in range(w - 1))
    # empty batch
    assert shard.shard_rows(np.zeros(1, dtype=np.int64), 4) == [(0, 0)] * 4


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(here, '..', 'oracle'), os.path.join(here, '..', 'tools')):
        sys.path.insert(0, p)
    lines = sc.Corpus('hinglish', 4).lines(30000) + ['', 'x', '']
    data, off = sc.pack(lines)
    lo, hi = shard.shard_rows(off, world)[rank]
    d, o = shard.take_shard(data, off, lo, hi)
    b = d.tobytes()
    mine = [b[o[i]:o[i + 1]].decode('utf-8') for i in range(len(o) - 1)]
    assert mine == lines[lo:hi]
    m = O.BpeModel(os.path.join(here, 'golden', 'models', 'bpe_corpus.json'))
    ids = [O.bpe_encode(m, O.normalize_text(s)) for s in mine]
    splits = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in ids], out=splits[1:])
    vals = np.array([t for x in ids for t in x], dtype=np.int32)
    gv, gs = shard.gather_ragged(vals, splits)
    # the product-level entry point over the same batch: every rank passes the whole batch, encodes its share, all get all

    def enc(dd, oo):
        bb = dd.tobytes()
        rows = [O.bpe_encode(m, O.normalize_text(bb[oo[i]:oo[i + 1]].decode('utf-8'))) for i in range(len(oo) - 1)]
        sp = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in rows], out=sp[1:])
        return np.array([t for x in rows for t in x], dtype=np.int32), sp
    ev, es = shard.encode_sharded(enc, data, off)
    assert ev.tolist() == gv.tolist() and es.tolist() == gs.tolist()
    lo2, hi2, sv, ss = shard.encode_sharded(enc, data, off, gather=False)
    assert (lo2, hi2) == (lo, hi) and sv.tolist() == vals.tolist()
    if rank == 0:
        q.put((gv.tolist(), gs.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(models_dir):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gv, gs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lines = sc.Corpus('hinglish', 4).lines(30000) + ['', 'x', '']
    m = O.BpeModel(os.path.join(models_dir, 'bpe_corpus.json'))
    exp = [O.bpe_encode(m, O.normalize_text(s)) for s in lines]
    assert gs[-1] == len(gv) and len(gs) == len(lines) + 1
    assert [gv[gs[i]:gs[i + 1]] for i in range(len(lines))] == exp
