"""C2 group_by under the table-update knobs (device-resident inputs): BL_K5_BULK (TMA bulk reduce of {len, sum} cells),
BL_K5_BULK_LANES (lanes of a warp that use it), BL_K5_BPS (CTAs per SM).  The first run of every configuration is checked
against numpy (bench.verify_groupby); kernel times are the library's CUDA-event profile.  One JSON object per line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_b200 as plb  # noqa: E402
import bench  # noqa: E402

ROWS = int(os.environ.get("VAR_ROWS", 100_000_000))
KEYS = int(os.environ.get("VAR_KEYS", 1_000_000))
STEPS = 5


def run(label, env, key, vi, vf, dkey, dvi, dvf, nulls=None):
    for k in ("BL_K5_BULK", "BL_K5_BULK_LANES", "BL_K5_BPS", "BL_K5_HINT", "BL_K5_LEAN"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})

    def step():
        return plb.group_by_agg(dkey.view(), [("sum", dvi.view()), ("mean", dvf.view()), ("len", None)], False, location=plb.DEVICE)

    k, outs = step()
    verified = "skipped"
    if nulls is None:
        verified = bench.verify_groupby(key, vi, vf, None, None, KEYS, k.to_numpy()[0], [o.to_numpy()[0] for o in outs])[:14]
    else:   # conservation: total len, and the valid-row integer sum
        ln = outs[2].to_numpy()[0].astype(np.int64); si = outs[0].to_numpy()[0]
        vmask = plb.unpack_bits(nulls, ROWS)
        assert ln.sum() == ROWS and si.sum() == vi[vmask].sum(), "null variant: totals differ"
        verified = "totals"
    del k, outs
    for _ in range(2):
        step()
    plb.sync()
    plb.profile_reset(); plb.profile_enable(True)
    for _ in range(STEPS):
        step()
    plb.sync()
    prof = {k: round(v["ms"] / STEPS, 4) for k, v in plb.profile().items()}
    plb.profile_enable(False)
    main = {k: v for k, v in prof.items() if v > 0.05}
    print(json.dumps({"config": label, "env": env, "verified": verified, "sum_kernels_ms": round(sum(prof.values()), 3), "kernels_ms": main}), flush=True)


def main():
    plb.init(0)
    key, vi, vf = bench.gen_groupby(ROWS, KEYS, 1)
    dkey, dvi, dvf = plb.to_device(key), plb.to_device(vi), plb.to_device(vf)
    a = (key, vi, vf, dkey, dvi, dvf)
    run("3 RED (word-major planes)", {"BL_K5_BULK": 0}, *a)
    run("bulk reduce, lean kernel (default grid: 48 CTAs/SM)", {"BL_K5_BULK": 1}, *a)
    for bps in (64, 96):
        run(f"bulk reduce, lean kernel, {bps} CTAs/SM", {"BL_K5_BULK": 1, "BL_K5_BPS": bps}, *a)
    for bps in (32, 48):
        run(f"3 RED, {bps} CTAs/SM", {"BL_K5_BULK": 0, "BL_K5_BPS": bps}, *a)
    # nullable value columns: null counters are extra REDs
    rng = np.random.default_rng(100)
    val_i, val_f = (plb.pack_bits(rng.random(ROWS) >= 0.05) for _ in range(2))
    nvi, nvf = plb.to_device(vi, val_i), plb.to_device(vf, val_f)
    an = (key, vi, vf, dkey, nvi, nvf)
    run("5% nulls, 3 RED", {"BL_K5_BULK": 0}, *an, nulls=val_i)
    run("5% nulls, lean bulk kernel", {"BL_K5_BULK": 1}, *an, nulls=val_i)


if __name__ == "__main__":
    main()
