"""CPU test of the N>1 host logic (world_size 2, gloo): split sizes, count exchange and the
all-to-all-v of hash-partitioned partial-aggregate rows in polars_b200/dist.py.  The device kernels are
stood in for by the oracle (partial aggregation + the reference's partition function), so what is
tested is the exchange plan itself: after it, rank p must own exactly the groups with
hash_to_partition(dirty_hash(key), 2) == p and their merged aggregates must equal a single-process run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROW_WORDS = 6   # key, len|first, sum_i64, sum_f64, nullcnt, meta  (groupby.cu partial-row format for sum/mean/len)


def _data(rank, n=20_000, k=700):
    rng = np.random.default_rng(100 + rank)
    return rng.integers(-k // 2, k // 2, n).astype(np.int64), rng.integers(-1000, 1000, n).astype(np.int64), rng.uniform(0, 100, n).round(6)


def _partial_rows(oracle, key, vi, vf, world):
    ek, _, outs, g = oracle.group_by_agg(key, None, [("sum", vi, None), ("sum", vf, None), ("len", None, None)], 1, True)
    rows = np.zeros((ek.size, ROW_WORDS), np.int64)
    rows[:, 0] = ek
    rows[:, 1] = (g.first.astype(np.int64) << 32) | outs[2][0].astype(np.int64)
    rows[:, 2] = outs[0][0]
    rows[:, 3] = outs[1][0].view(np.int64)
    part = oracle.hash_to_partition(oracle.dirty_hash(oracle.key_bits(ek)), world).astype(np.int64)
    order = np.argsort(part, kind="stable")
    return rows[order], np.bincount(part, minlength=world)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from polars_b200.dist import all_to_all_rows
    key, vi, vf = _data(rank)
    rows, counts = _partial_rows(oracle, key, vi, vf, world)
    recv, rc = all_to_all_rows(torch.from_numpy(rows.reshape(-1).copy()), counts, ROW_WORDS)
    got = recv.numpy().reshape(-1, ROW_WORDS)
    assert got.shape[0] == rc.sum()
    # merge (what k_gb_merge does): add len, sums; min first
    ks, inv = np.unique(got[:, 0], return_inverse=True)
    ln = np.zeros(ks.size, np.int64)
    si = np.zeros(ks.size, np.int64)
    sf = np.zeros(ks.size, np.float64)
    np.add.at(ln, inv, got[:, 1] & 0xFFFFFFFF)
    np.add.at(si, inv, got[:, 2])
    np.add.at(sf, inv, got[:, 3].copy().view(np.float64))
    q.put((rank, ks, ln, si, sf))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_partitioned_exchange_world2():
    import oracle
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=240)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    key = np.concatenate([_data(r)[0] for r in range(world)])
    vi = np.concatenate([_data(r)[1] for r in range(world)])
    vf = np.concatenate([_data(r)[2] for r in range(world)])
    ek, _, outs, _ = oracle.group_by_agg(key, None, [("sum", vi, None), ("sum", vf, None), ("len", None, None)], 4, True)
    part = oracle.hash_to_partition(oracle.dirty_hash(oracle.key_bits(ek)), world)
    for p in range(world):
        ks, ln, si, sf = res[p]
        sel = part == p
        order = np.argsort(ek[sel])
        assert np.array_equal(ks, ek[sel][order]), f"rank {p} owns the wrong groups"
        assert np.array_equal(ln, outs[2][0][sel][order].astype(np.int64)) and np.array_equal(si, outs[0][0][sel][order])
        assert np.allclose(sf, outs[1][0][sel][order], rtol=1e-9)


def _worker_rows(rank, world, port, q):
    """High-cardinality plan (ii): raw rows partitioned on the key, exchanged column by column, aggregated on the owner."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from polars_b200.dist import exchange_columns
    key, vi, vf = _data(rank, n=30_000, k=40_000)           # groups ~ rows: pre-aggregation would not shrink anything
    part = oracle.hash_to_partition(oracle.dirty_hash(oracle.key_bits(key)), world).astype(np.int64)
    order = np.argsort(part, kind="stable")                  # what K6 does: stable scatter by partition
    counts = np.bincount(part, minlength=world)
    (rk, ri, rf, r32), rc = exchange_columns([torch.from_numpy(key[order]), torch.from_numpy(vi[order]), torch.from_numpy(vf[order]),
                                              torch.from_numpy(vi[order].astype(np.int32))], counts)
    assert rk.numel() == rc.sum() and np.array_equal(r32.numpy(), ri.numpy().astype(np.int32))
    owner = oracle.hash_to_partition(oracle.dirty_hash(oracle.key_bits(rk.numpy())), world)
    assert (owner == rank).all()
    ek, _, outs, _ = oracle.group_by_agg(rk.numpy(), None, [("sum", ri.numpy(), None), ("mean", rf.numpy(), None), ("len", None, None)], 1, True)
    q.put((rank, ek, outs[0][0], outs[1][0], outs[2][0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_raw_row_exchange_world2():
    import oracle
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_rows, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=240)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    key = np.concatenate([_data(r, n=30_000, k=40_000)[0] for r in range(world)])
    vi = np.concatenate([_data(r, n=30_000, k=40_000)[1] for r in range(world)])
    vf = np.concatenate([_data(r, n=30_000, k=40_000)[2] for r in range(world)])
    ek, _, outs, _ = oracle.group_by_agg(key, None, [("sum", vi, None), ("mean", vf, None), ("len", None, None)], 4, True)
    got_k = np.concatenate([res[r][0] for r in range(world)])
    o, eo = np.argsort(got_k), np.argsort(ek)
    assert np.array_equal(got_k[o], ek[eo])                   # every group on exactly one rank
    assert np.array_equal(np.concatenate([res[r][1] for r in range(world)])[o], outs[0][0][eo])
    assert np.allclose(np.concatenate([res[r][2] for r in range(world)])[o], outs[1][0][eo], rtol=1e-9)
    assert np.array_equal(np.concatenate([res[r][3] for r in range(world)])[o], outs[2][0][eo])


def test_join_plan_rule():
    """Exchange-volume rule (dist.choose_join_plan): C3's small build side is broadcast, relations of similar size are partitioned."""
    from polars_b200 import dist as pdist
    assert pdist.choose_join_plan(100_000_000, 10_000_000, 2) == "broadcast"
    assert pdist.choose_join_plan(100_000_000, 10_000_000, 8) == "broadcast"
    assert pdist.choose_join_plan(100_000_000, 100_000_000, 8) == "partitioned"
    assert pdist.choose_join_plan(100_000_000, 70_000_000, 4) == "partitioned"
