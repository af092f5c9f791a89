"""Per-kernel bandwidth sweep (K1 arithmetic, K2 compare, K3 filter, K4 gather, K6 partition) on device-resident
columns: algorithmic bytes / CUDA-event kernel time, against MEASURED_PEAKS.json.
One JSON line per kernel.  `--once` runs every kernel a single time (for `ncu --set full`)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_b200 as plb  # noqa: E402

once = "--once" in sys.argv
N = 100_000_000
REPS = 1 if once else 5
peak = 6567.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
plb.init()
rng = np.random.default_rng(0)
a = plb.to_device(rng.integers(-10**6, 10**6, N).astype(np.int64))
b = plb.to_device(rng.integers(1, 10**6, N).astype(np.int64))
f = plb.to_device(rng.uniform(0, 100, N))
g = plb.to_device(rng.uniform(1, 100, N))
idx = plb.to_device(rng.integers(0, 10_000_000, N).astype(np.uint32))
src = plb.to_device(rng.normal(size=10_000_000))
idx_big = plb.to_device(rng.integers(0, N, N).astype(np.uint32))
scalar0 = np.array([0], np.int64)


def run(name, kernel, fn, alg_bytes, note=""):
    fn()
    plb.profile_reset(); plb.profile_enable(True)
    for _ in range(REPS):
        out = fn()
    plb.sync()
    prof = plb.profile(); plb.profile_enable(False)
    k = prof.get(kernel, {"ms": 0, "launches": 1})
    ms = k["ms"] / max(k["launches"], 1)
    gbs = alg_bytes / 1e9 / (ms / 1e3) if ms else 0
    print(json.dumps({"op": name, "kernel": kernel, "rows": N, "kernel_ms": round(ms, 4), "algorithmic_GB": round(alg_bytes / 1e9, 3), "achieved_GBps": round(gbs, 1),
                      "frac_of_measured_peak": round(gbs / peak, 3), "note": note}), flush=True)
    del out


D = plb.DEVICE
run("add i64 (array+array)", "k1_arith", lambda: plb.elementwise("add", a.view(), b.view(), location=D), 24 * N)
run("mul f64 (array*array)", "k1_arith", lambda: plb.elementwise("mul", f.view(), g.view(), location=D), 24 * N)
run("mul f64 (array*scalar)", "k1_arith", lambda: plb.elementwise("mul", f.view(), np.array([1.5]), location=D), 16 * N)
run("floordiv i64 (array//array)", "k1_arith", lambda: plb.elementwise("floordiv", a.view(), b.view(), location=D), 24 * N, "64-bit integer division is ALU-bound")
run("gt i64 (array>scalar)", "k2_compare", lambda: plb.compare("gt", a.view(), scalar0, location=D), (8 + 1 / 8) * N)
run("lt f64 (array<array)", "k2_compare", lambda: plb.compare("lt", f.view(), g.view(), location=D), (16 + 1 / 8) * N)
for sel, thr in ((0.5, 0), (0.1, 800_000), (0.9, -800_000)):
    thr_arr = np.array([thr], np.int64)
    run(f"filter x>c, 2 columns, selectivity {sel}", "k3_compact", lambda: plb.filter_cmp([a.view(), f.view()], 0, "gt", thr_arr[0], location=D), 8 * 2 * N * (1 + sel))
run("gather f64, 10M-row source (L2-resident)", "k4_gather", lambda: plb.gather([src.view()], idx.view(), check_bounds=False, location=D), 20 * N)
run("gather f64, 1e8-row source (HBM)", "k4_gather", lambda: plb.gather([f.view()], idx_big.view(), check_bounds=False, location=D), 20 * N, "random 8-byte reads: 32-byte sectors fetched")
run("hash partition key+1 payload, P=8", "k6_part_scatter", lambda: plb.hash_partition(a.view(), [f.view()], 8, location=D), 2 * 16 * N)
