"""One launch of (almost) every product kernel at benchmark size, for `ncu --set full` (SURVEY.md §8 / north star: "every
kernel has a committed ncu capture").  Run as
    ncu --set full --clock-control none --import-source on -o gpurun_out/r02_all python tools/ncu_all.py
and summarise with tools/ncu_summary.py.  Each section runs its operator once (no warm-up: the capture replays it)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_b200 as plb  # noqa: E402
import bench  # noqa: E402

N = int(os.environ.get("NCU_ROWS", 100_000_000))      # headline kernels at the benchmarked size
D = plb.DEVICE
plb.init(0)
rng = np.random.default_rng(0)

# ---- C2 group_by (L2 plan), Zipf keys (heavy-hitter plan), low cardinality (CTA-private tables), K5r (beyond L2)
key, vi, vf = bench.gen_groupby(N, 1_000_000, 1)
dkey, dvi, dvf = plb.to_device(key), plb.to_device(vi), plb.to_device(vf)
aggs = lambda: [("sum", dvi.view()), ("mean", dvf.view()), ("len", None)]      # noqa: E731
plb.group_by_agg(dkey.view(), aggs(), False, location=D)
zkey = plb.to_device(bench.gen_groupby(N, 1_000_000, 1, "zipf")[0])
plb.group_by_agg(zkey.view(), aggs(), False, location=D)
del zkey
lkey = plb.to_device((key % 100).astype(np.int64))
plb.group_by_agg(lkey.view(), aggs(), False, location=D)
del lkey
hkey = plb.to_device(rng.integers(0, 4_000_000, N, dtype=np.int64))
plb.group_by_agg(hkey.view(), aggs(), False, location=D)      # K5r: histogram, scatter (TMA stores), aggregate (TMA loads)
os.environ["BL_K5_RADIX"] = "2"
plb.group_by_agg(dkey.view(), aggs(), False, location=D)      # K5r on C2 itself (bulk-store path, 256..1024 buckets)
os.environ.pop("BL_K5_RADIX")
del hkey

# ---- string keys: device dictionary encoding (hash, first-row ids, verification) + gather of the distinct values
ids = rng.integers(0, 200_000, 4_000_000)
lens = 1 + (ids % 23).astype(np.int64)
offs = np.zeros(ids.size + 1, np.int64); np.cumsum(lens, out=offs[1:])
pos = np.arange(int(offs[-1]), dtype=np.int64) - np.repeat(offs[:-1], lens)
sdata = ((np.repeat(ids, lens) * 31 + pos * 7) % 251).astype(np.uint8)
scol = plb.StringColumn(offsets=offs, data=sdata)
codes, _nd = plb.string_encode(scol, location=D)
plb.string_gather(scol, np.unique(codes.to_numpy()[0])[:100_000].astype(np.uint32))
del codes, scol, sdata, pos, offs, lens, ids

# ---- multi-GPU export / merge kernels in one process: a window to ourselves (world size 1)
g = plb.GroupBy(np.int64, [("sum", np.int64), ("mean", np.float64), ("len", None)], nullable=[False, False, False])
g.consume(dkey.view(), [dvi.view(), dvf.view(), None])
rows_per_src, row_words = 1_000_000 + 1024, 5
win = plb.Window(1024 + rows_per_src * row_words * 8)
g.export_partials_p2p_async([win.ptr], 0, rows_per_src, 1)
f = plb.GroupBy(np.int64, [("sum", np.int64), ("mean", np.float64), ("len", None)], expected_groups=1_300_000, nullable=[False, False, False])
f.merge_window_async(win.ptr, 1, rows_per_src, 1)
f.finish(False, location=D)
del g, f

# ---- group tuples + deterministic folds (radix sort kernels)
small = plb.to_device(key[: N // 10].copy())
plb.set_deterministic(True)
plb.group_by_agg(small.view(), [("sum", plb.to_device(vf[: N // 10].copy()).view())], False, location=D)
plb.set_deterministic(False)
del small, dkey, dvi, dvf, key, vi, vf

# ---- C3 joins: dense, hashed (sparse keys), duplicates (two-pass probe + emit), semi, full
probe, build = bench.gen_join(N, N // 10, 2)
dp, db = plb.to_device(probe), plb.to_device(build)
plb.hash_join(dp.view(), db.view(), "inner", False, "none", location=D)
sp, sb = plb.to_device(bench.sparsify(probe)), plb.to_device(bench.sparsify(build))
plb.hash_join(sp.view(), sb.view(), "inner", False, "none", location=D)
plb.hash_join(sp.view(), sb.view(), "semi", False, "none", location=D)
plb.hash_join(sp.view(), sb.view(), "full", False, "none", location=D)
p4, b4 = bench.gen_join(N // 2, N // 10, 2, dup=4, sparse=True)
d4p, d4b = plb.to_device(p4), plb.to_device(b4)
plb.hash_join(d4p.view(), d4b.view(), "inner", False, "none", location=D)
del d4p, d4b, sp, sb, p4, b4

# ---- K6 partition, K4 gather, K3 filter, K2 compare, K1 arithmetic
a = plb.to_device(rng.integers(-10**6, 10**6, N).astype(np.int64))
fcol = plb.to_device(rng.uniform(0, 100, N))
plb.hash_partition(a.view(), [fcol.view()], 8, location=D)
idx = plb.to_device(rng.integers(0, N, N).astype(np.uint32))
plb.gather([fcol.view()], idx.view(), check_bounds=False, location=D)
plb.filter_cmp([a.view(), fcol.view()], 0, "gt", 0, location=D)
plb.compare("gt", a.view(), np.array([0], np.int64), location=D)
plb.elementwise("add", a.view(), a.view(), location=D)
plb.sync()
print("ncu_all done, launches", plb.launch_count())
