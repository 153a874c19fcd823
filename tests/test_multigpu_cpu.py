"""Host-side logic of the multi-GPU path on CPU: range planning and the variable-size record exchange over a
world_size-2 gloo group.  The per-rank sort/count is emulated with numpy on selftest-built records, and the
union of the ranks' solid edges must equal the oracle's - i.e. partitioning by top-byte ranges loses nothing."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from megahit_b200 import lib
from megahit_b200.multigpu import plan_ranges, split_counts


def test_plan_ranges_balances_and_covers():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for _ in range(20):
            hist = rng.integers(0, 1000, 256)
            hist[: rng.integers(0, 200)] = 0  # skew / empty prefixes
            b = plan_ranges(hist, world)
            assert b[0] == 0 and b[-1] == 256 and (np.diff(b) > 0).all() and len(b) == world + 1
            c = split_counts(hist, b)
            assert c.sum() == hist.sum()
    # skewed like canonical k-mers: half of the mass in the first quarter
    hist = np.concatenate([np.full(64, 300), np.full(192, 100)])
    c = split_counts(hist, plan_ranges(hist, 4))
    assert c.max() <= 1.1 * c.sum() / 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, m, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from megahit_b200 import lib as L2
    from megahit_b200.multigpu import exchange_records, plan_ranges as pr, split_counts as sc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    case = os.path.join(GOLDEN, "syn150_k27")
    bin_words = np.fromfile(os.path.join(case, "reads.lib.bin"), np.uint32).reshape(3000, -1)
    mine = bin_words[rank::world]  # any split of the reads works
    wr = L2.count_record_words(k)
    recs = []
    for row in mine:
        Lr = int(row[0])
        for pos in range(Lr - k):
            recs.append(L2.selftest_count_record(row[1:], Lr, k, pos)[0])
    recs = np.array(recs, np.uint32).reshape(-1, wr)
    top = recs[:, 0] >> 24
    hist = np.bincount(top, minlength=256).astype(np.int64)
    g = torch.from_numpy(hist.copy())
    dist.all_reduce(g)
    bounds = pr(g.numpy(), world)
    send = sc(hist, bounds)
    grouped = recs[np.argsort(top, kind="stable")]
    t = torch.from_numpy(grouped.view(np.int32).reshape(-1).copy())
    out, n_recv = exchange_records(t, wr, send)
    own = out[: n_recv * wr].numpy().view(np.uint32).reshape(-1, wr)
    # every received record belongs to this rank's range
    assert ((own[:, 0] >> 24) >= bounds[rank]).all() and ((own[:, 0] >> 24) < bounds[rank + 1]).all()
    # local "sort + count": numpy
    key = own.copy()
    key[:, wr - 1] &= np.uint32(0xFFFFFFC0)
    order = np.lexsort(tuple(key[:, j] for j in reversed(range(wr))))
    srt = key[order]
    head = np.ones(len(srt), bool)
    head[1:] = (srt[1:] != srt[:-1]).any(axis=1)
    starts = np.nonzero(head)[0]
    counts = np.diff(np.append(starts, len(srt)))
    solid = srt[starts[counts >= m]]
    q.put((rank, solid, counts[counts >= m]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_reproduces_oracle_edges():
    import torch.multiprocessing as mp
    from oracle_pipeline import load_reads, oracle_count
    k, m, world = 27, 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oc = oracle_count(load_reads(os.path.join(GOLDEN, "syn150_k27")), k, m)
    keys = np.concatenate([g[1] for g in got])  # rank order == ascending bucket ranges
    mult = np.concatenate([g[2] for g in got])
    assert len(keys) == oc["n_solid"]
    assert (keys[:, :2] == (oc["edges"][:, :2] & np.array([0xFFFFFFFF, 0xFFFFFF00], np.uint32))).all()
    assert (np.minimum(mult, 65535) == (oc["edges"][:, 2] & 0xFFFF)).all()


def test_reverse_rows_reproduces_the_reference_cand_file():
    """multigpu._reverse_rows (the `.cand` writer of the partitioned build): candidate reads in the REVERSED orientation
    KmerCounter holds them in (kmer_counter.cpp:387-401) - byte-identical to the `.cand` the reference wrote"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_pipeline as OP
    from megahit_b200.multigpu import _reverse_rows
    case = os.path.join(GOLDEN, "syn150_k27")
    rows = np.fromfile(os.path.join(case, "reads.lib.bin"), np.uint32).reshape(3000, -1)
    oc = OP.oracle_count(OP.load_reads(case), 27, 2)
    got = _reverse_rows(rows[oc["cand_ids"]], 150).tobytes()
    assert got == open(os.path.join(case, "cand.bin"), "rb").read() and len(got) > 0
    assert _reverse_rows(rows[:0], 150).shape == (0, rows.shape[1])
