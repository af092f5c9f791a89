"""torchrun --nproc-per-node N tools/mgpu_check.py — correctness of the partitioned group_by on N GPUs
(both exchanges) against a numpy reduction of the concatenated data."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_b200 as plb  # noqa: E402
from polars_b200 import dist as pdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
plb.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, K = 2_000_000, 50_000


def data(r):
    rng = np.random.default_rng(7 + r)
    return rng.integers(-K // 2, K // 2, n).astype(np.int64), rng.integers(-1000, 1000, n).astype(np.int64), rng.uniform(0, 100, n).round(6)


key, vi, vf = data(rank)
dk, dvi, dvf = plb.to_device(key), plb.to_device(vi), plb.to_device(vf)
spec = [("sum", np.int64), ("mean", np.float64), ("len", None)]
allk = np.concatenate([data(r)[0] for r in range(world)])
alli = np.concatenate([data(r)[1] for r in range(world)])
allf = np.concatenate([data(r)[2] for r in range(world)])
uk, inv = np.unique(allk, return_inverse=True)
esum = np.zeros(uk.size, np.int64); np.add.at(esum, inv, alli)
efs = np.zeros(uk.size); np.add.at(efs, inv, allf)
ecnt = np.bincount(inv, minlength=uk.size)
h = (uk.view(np.uint64) * np.uint64(0x55fbfd6bfc5458e9))
part = np.array([(int(x) * world) >> 64 for x in h])
ex = pdist.PeerExchange(plb, rows_per_src=K + 1024, row_words=6)  # sum, mean(+null counter), len -> 3 words + key, len|first, meta
for mode in ("nccl", "p2p", "p2p"):
    if mode == "nccl":
        ok, outs = pdist.partitioned_group_by(plb, dk.view(), [dvi.view(), dvf.view(), None], spec)
    else:
        ok, outs = pdist.partitioned_group_by_p2p(plb, ex, dk.view(), [dvi.view(), dvf.view(), None], spec)
    k, _ = ok.to_numpy()
    s, _ = outs[0].to_numpy(); m, _ = outs[1].to_numpy(); c, _ = outs[2].to_numpy()
    o = np.argsort(k)
    sel = part == rank
    assert np.array_equal(k[o], uk[sel]), f"{mode}: rank {rank} owns the wrong groups ({k.size} vs {sel.sum()})"
    assert np.array_equal(s[o], esum[sel]) and np.array_equal(c[o], ecnt[sel]), f"{mode}: sums/counts differ"
    assert np.allclose(m[o], efs[sel] / ecnt[sel], rtol=1e-9), f"{mode}: means differ"
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check {mode}: OK world={world} groups_here={k.size}", flush=True)
ex.close()

# ---- partitioned join: global probe rows x globally-unique build keys
npr, nbr = 1_000_000, 100_000
perm = np.random.default_rng(99).permutation(nbr * world * 2)[: nbr * world].astype(np.int64)     # same on every rank
bkeys = perm[rank * nbr:(rank + 1) * nbr].copy()
pkeys_all = [np.random.default_rng(500 + r).integers(0, nbr * world * 2, npr).astype(np.int64) for r in range(world)]
dpk, dbk = plb.to_device(pkeys_all[rank]), plb.to_device(bkeys)
gl, gr = pdist.partitioned_hash_join(plb, dpk.view(), dbk.view(), rank * npr, rank * nbr)
l, _ = gl.to_numpy(); r_, _ = gr.to_numpy()
allp = np.concatenate(pkeys_all)
assert np.array_equal(allp[l], perm[r_]), "join: keys of the emitted pairs differ"
kh = (allp[l].view(np.uint64) * np.uint64(0x55fbfd6bfc5458e9))
assert all(((int(x) * world) >> 64) == rank for x in kh[:2000]), "join: pair on the wrong rank"
tot = torch.tensor([l.size], dtype=torch.int64, device="cuda"); dist.all_reduce(tot)
expect = int(np.isin(allp, perm).sum())
assert int(tot.item()) == expect, f"join: {int(tot.item())} pairs, expected {expect}"
assert np.unique(l).size == l.size
if rank == 0:
    print(f"mgpu_check join: OK world={world} pairs={expect}", flush=True)

# ---- broadcast-build join: same relation, no probe-side exchange; rank r keeps the matches of its own probe rows
gl2, gr2 = pdist.broadcast_hash_join(plb, dpk.view(), dbk.view(), rank * npr)
l2, _ = gl2.to_numpy(); r2, rv2 = gr2.to_numpy()
assert np.array_equal(allp[l2], perm[r2]), "broadcast join: keys of the emitted pairs differ"
assert l2.size == int(np.isin(pkeys_all[rank], perm).sum()) and (l2.size == 0 or (l2.min() >= rank * npr and l2.max() < (rank + 1) * npr))
assert np.all(np.diff(l2.astype(np.int64)) > 0), "broadcast join: probe order lost"
if rank == 0:
    print(f"mgpu_check broadcast join: OK world={world}", flush=True)

# ---- high-cardinality plan (ii): raw rows partitioned + exchanged, one aggregation on the owning rank
hk = [np.random.default_rng(900 + r).integers(0, 5_000_000, 400_000).astype(np.int64) for r in range(world)]
hv = [np.random.default_rng(950 + r).integers(-1000, 1000, 400_000).astype(np.int64) for r in range(world)]
dk, dv = plb.to_device(hk[rank]), plb.to_device(hv[rank])
okey, oaggs = pdist.partitioned_group_by_rows(plb, dk.view(), [dv.view()], [("sum", 0), ("len", None)])
k_, _ = okey.to_numpy(); s_, _ = oaggs[0].to_numpy(); l_, _ = oaggs[1].to_numpy()
allk, allv = np.concatenate(hk), np.concatenate(hv)
uk, inv = np.unique(allk, return_inverse=True)
es, el = np.zeros(uk.size, np.int64), np.bincount(inv)
np.add.at(es, inv, allv)
pos = np.searchsorted(uk, k_)
assert np.array_equal(uk[pos], k_) and np.array_equal(es[pos], s_) and np.array_equal(el[pos], l_), "raw-row plan: aggregates differ"
tot = torch.tensor([k_.size], dtype=torch.int64, device="cuda"); dist.all_reduce(tot)
assert int(tot.item()) == uk.size, "raw-row plan: a group is missing or owned twice"
if rank == 0:
    print(f"mgpu_check raw-row group_by: OK world={world} groups={uk.size}", flush=True)
dist.destroy_process_group()
