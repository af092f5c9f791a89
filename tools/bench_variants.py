"""C2 / C3 workload variants of SURVEY.md 8(d) in one process (device-resident inputs, CUDA-event-free wall
timing around synchronised steps): Zipf(1.1) keys, 5 % nulls, 50 % hit rate, 4 duplicates per build key,
with the heavy-hitter path on and off, sorted keys, frequent null keys.  Prints one JSON object per line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_b200 as plb  # noqa: E402
import bench  # noqa: E402

ROWS = int(os.environ.get("VAR_ROWS", 100_000_000))
KEYS = int(os.environ.get("VAR_KEYS", 1_000_000))
STEPS = 5


def timed(fn, label, rows, extra=None):
    for _ in range(3):
        n_out = fn()
    plb.sync()
    plb.profile_reset(); plb.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(STEPS):
        n_out = fn()
    plb.sync()
    ms = (time.perf_counter() - t0) / STEPS * 1e3
    prof = {k: round(v["ms"] / STEPS, 3) for k, v in plb.profile().items() if v["ms"] / STEPS > 0.05}
    plb.profile_enable(False)
    print(json.dumps({"variant": label, "ms_per_step": round(ms, 3), "rows_per_s": rows / ms * 1e3, "out_rows": int(n_out), "kernels_ms": prof, **(extra or {})}), flush=True)


def main():
    plb.init(0)
    key, vi, vf = bench.gen_groupby(ROWS, KEYS, 1)
    dkey, dvi, dvf = plb.to_device(key), plb.to_device(vi), plb.to_device(vf)

    def gb(k, a, b):
        return lambda: plb.group_by_agg(k.view(), [("sum", a.view()), ("mean", b.view()), ("len", None)], False, location=plb.DEVICE)[0].length

    timed(gb(dkey, dvi, dvf), "groupby uniform (C2)", ROWS)
    rng = np.random.default_rng(100)
    val_i, val_f = (plb.pack_bits(rng.random(ROWS) >= 0.05) for _ in range(2))
    nvi, nvf = plb.to_device(vi, val_i), plb.to_device(vf, val_f)
    timed(gb(dkey, nvi, nvf), "groupby uniform, 5% nulls in both value columns", ROWS)
    del nvi, nvf
    zkey = plb.to_device(bench.gen_groupby(ROWS, KEYS, 1, "zipf")[0])
    timed(gb(zkey, dvi, dvf), "groupby Zipf(1.1) keys", ROWS)
    os.environ["BL_K5_HOTKEYS"] = "0"
    timed(gb(zkey, dvi, dvf), "groupby Zipf(1.1) keys, heavy-hitter rows off (BL_K5_HOTKEYS=0)", ROWS, {"BL_K5_HOTKEYS": 0})
    os.environ.pop("BL_K5_HOTKEYS")
    z3 = plb.to_device((np.random.default_rng(5).zipf(1.1, ROWS) % 1000).astype(np.int64))
    timed(gb(z3, dvi, dvf), "groupby Zipf(1.1) keys folded into 1000 groups", ROWS)
    del z3
    nk = plb.to_device(key, np.random.default_rng(6).random(ROWS) >= 0.1)
    timed(gb(nk, dvi, dvf), "groupby uniform keys, 10% null keys (one hot null group)", ROWS)
    del nk
    skey = plb.to_device(np.sort(key))
    timed(gb(skey, dvi, dvf), "groupby sorted keys (runs of ~100 equal keys)", ROWS)
    del dkey, dvi, dvf, zkey, skey, key, vi, vf

    build_rows = ROWS // 10
    for hit, dup in ((1.0, 1), (0.5, 1), (1.0, 4)):
        probe, build = bench.gen_join(ROWS, build_rows, 2, hit, dup)
        dp, db = plb.to_device(probe), plb.to_device(build)
        timed(lambda: plb.hash_join(dp.view(), db.view(), "inner", False, "none", location=plb.DEVICE)[0].length,
              f"join {ROWS} x {build_rows}, hit {hit:.0%}, {dup} copies/build key", ROWS)
        del dp, db


if __name__ == "__main__":
    main()
