#!/usr/bin/env python
"""bench.py — the hot path on synthetic data, one JSON line (contract: task brief).

Default workload = BASELINE.json configs[1] (C2): hash group_by over 1e8 rows, 1e6 uniform Int64 keys,
aggregations sum(v_i64), mean(v_f64), len, on ONE B200; inputs (2.4 GB) are larger than L2 (126 MB), so
no explicit L2 flush is needed between timed iterations.

  value     rows/s, whole job, inputs resident in HBM when the timed region starts, through the C ABI
            with BL_DEVICE columns (estimate + table init + fused build/aggregate + extraction).
  e2e       same call with BL_HOST columns in pinned memory: H2D of the three columns and D2H of the
            result inside the timed region.
  roofline  dominant kernel (k5_groupby_agg): algorithmic bytes (24 B/row) / its CUDA-event duration
            on the library stream, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (restatement of the reference's Rayon algorithm, "port") on the host
            cores over a bounded sample.

--workload join: configs[2] (C3) inner hash join 1e8 x 1e7 on Int64 (single GPU here).
--gpus N (torchrun): weak scaling, every rank owns --rows rows; local pre-aggregation, hash partition of
the partial aggregates, ONE all-to-all (NCCL), final merge.
--impl reference: the oracle on the host cores (the reference itself cannot be installed: Rust, no
toolchain/wheel — see DESIGN.md), same metric/config, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="groupby", choices=["groupby", "join", "q1"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--keys", type=int, default=1_000_000)
    ap.add_argument("--build-rows", type=int, default=10_000_000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=20_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skew", default="uniform", choices=["uniform", "zipf"], help="group_by key distribution (SURVEY.md 8(d) C2 variants)")
    ap.add_argument("--null-frac", type=float, default=0.0, help="group_by: fraction of null rows in each value column")
    ap.add_argument("--hit-frac", type=float, default=1.0, help="join: fraction of probe rows with a match (C3 variant: 0.5)")
    ap.add_argument("--dup", type=int, default=1, help="join: copies of every build key (C3 variant: 4)")
    ap.add_argument("--acero-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="multi-GPU exchange of the partial aggregates")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def gen_groupby(rows: int, keys: int, seed: int, skew: str = "uniform"):
    """uniform keys (or the Zipf(1.1) variant folded into [0, keys)), v_i64 in
    [-1000,1000), v_f64 = U(0,100).round(6) (h2oai v3)."""
    rng = np.random.default_rng(seed)
    if skew == "zipf":
        key = (rng.zipf(1.1, rows) % keys).astype(np.int64)
    else:
        key = rng.integers(0, keys, rows, dtype=np.int64)
    vi = rng.integers(-1000, 1000, rows, dtype=np.int64)
    vf = rng.uniform(0, 100, rows).round(6)
    return key, vi, vf


def gen_join(rows: int, build_rows: int, seed: int, hit_frac: float = 1.0, dup: int = 1):
    """C3: build = permutation (unique keys, or `dup` copies of each), probe keys uniform over
    [0, distinct build keys / hit_frac): 100 % hit by default; variants 50 % hit and 4 duplicates per key."""
    rng = np.random.default_rng(seed)
    distinct = max(1, build_rows // dup)
    build = rng.permutation(distinct * dup).astype(np.int64) % distinct if dup > 1 else rng.permutation(build_rows).astype(np.int64)
    probe = rng.integers(0, max(1, int(distinct / hit_frac)), rows, dtype=np.int64)
    return probe, build


def gen_lineitem(rows: int, seed: int):
    """BASELINE.json configs[3] (C4): PDS-H / TPC-H lineitem columns used by Q1, synthetic with dbgen-like
    marginals (no dbgen binary here): quantity 1..50, extendedprice = quantity * U(900, 2100), discount
    0..0.10, tax 0..0.08 (2 decimals), returnflag in {A,N,R} / linestatus in {O,F} as dictionary codes,
    shipdate uniform over 1992-01-02 .. 1998-12-01 in days since epoch."""
    rng = np.random.default_rng(seed)
    qty = rng.integers(1, 51, rows).astype(np.float64)
    price = (qty * rng.uniform(900.0, 2100.0, rows)).round(2)
    disc = (rng.integers(0, 11, rows) / 100.0)
    tax = (rng.integers(0, 9, rows) / 100.0)
    ship = rng.integers(8036, 10561, rows).astype(np.int64)            # days: 1992-01-02 .. 1998-12-01
    rf = np.where(ship > 9298, 1, rng.integers(0, 2, rows) * 2).astype(np.int64)   # N after 1995-06-17, else A(0)/R(2)
    ls = (ship > 9298).astype(np.int64)                                  # O(1) / F(0)
    return {"qty": qty, "price": price, "disc": disc, "tax": tax, "ship": ship, "rf": rf, "ls": ls}


Q1_CUTOFF = 10471   # 1998-09-02


def q1_device(plb, d):
    """PDS-H Q1 through the C ABI on device columns: filter (K2+K3) -> expressions (K1) -> group_by/agg (K5)."""
    cols = [d["ship"].view(), d["qty"].view(), d["price"].view(), d["disc"].view(), d["tax"].view(), d["rf"].view(), d["ls"].view()]
    f = plb.filter_cmp(cols, 0, "le", Q1_CUTOFF, location=plb.DEVICE)
    _, qty, price, disc, tax, rf, ls = f
    one_minus = plb.elementwise("sub", np.array([1.0]), disc.view(), location=plb.DEVICE)
    disc_price = plb.elementwise("mul", price.view(), one_minus.view(), location=plb.DEVICE)
    one_plus = plb.elementwise("add", tax.view(), np.array([1.0]), location=plb.DEVICE)
    charge = plb.elementwise("mul", disc_price.view(), one_plus.view(), location=plb.DEVICE)
    key = plb.elementwise("add", plb.elementwise("mul", rf.view(), np.array([256], np.int64), location=plb.DEVICE).view(), ls.view(), location=plb.DEVICE)
    q, p_, dp, ch, di = qty.view(), price.view(), disc_price.view(), charge.view(), disc.view()
    return plb.group_by_agg(key.view(), [("sum", q), ("sum", p_), ("sum", dp), ("sum", ch), ("mean", q), ("mean", p_), ("mean", di), ("len", None)], False, location=plb.DEVICE)


def q1_numpy(h):
    m = h["ship"] <= Q1_CUTOFF
    key = h["rf"][m] * 256 + h["ls"][m]
    dp = h["price"][m] * (1.0 - h["disc"][m])
    ch = dp * (h["tax"][m] + 1.0)
    uk, inv = np.unique(key, return_inverse=True)
    out = {"key": uk, "len": np.bincount(inv)}
    for name, v in (("qty", h["qty"][m]), ("price", h["price"][m]), ("dp", dp), ("ch", ch), ("disc", h["disc"][m])):
        out[name] = np.bincount(inv, weights=v)
    return out


# ------------------------------------------------------------------------------------- reference arm
def run_reference(a):
    import oracle
    oracle.build()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = oracle.max_threads()
    sample = min(a.rows, a.cpu_sample)
    if a.workload == "groupby":
        key, vi, vf = gen_groupby(sample, a.keys, 1, a.skew)
        aggs = [("sum", vi, None), ("mean", vf, None), ("len", None, None)]
        fn = lambda: oracle.group_by_agg(key, None, aggs, cores, False)   # noqa: E731
        unit_rows = sample
        metric, wl = "group_by_agg_rows_per_sec", f"C2 hash group_by {a.rows} rows, {a.keys} Int64 keys, sum(i64)/mean(f64)/len"
    else:
        build_rows = max(1, int(a.build_rows * sample / a.rows))
        probe, build = gen_join(sample, build_rows, 2, a.hit_frac, a.dup)
        fn = lambda: oracle.hash_join(probe, build, None, None, "inner", False, "none", cores)   # noqa: E731
        unit_rows = sample
        metric, wl = "hash_join_probe_rows_per_sec", f"C3 inner hash join {a.rows} x {a.build_rows} Int64"
    for _ in range(min(a.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = (time.perf_counter() - t0) / a.steps
    v = unit_rows / dt
    line = {"impl": "reference", "metric": metric, "value": v, "unit": "rows/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/float64", "data": "synthetic",
            "config": {"workload": wl, "sample_rows_per_step": sample},
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} rows/step of the same workload; oracle = C restatement of the reference's partitioned Rayon algorithm (not Polars itself: no Rust toolchain / wheel)"},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- B200 arm
def main():
    a = parse()
    # NCCL prints its version banner to STDOUT for any NCCL_DEBUG level: keep the JSON line alone on stdout
    os.environ.pop("NCCL_DEBUG", None)
    if os.environ.get("BENCH_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
    if a.acero_child:
        acero_child(a)
        return
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    import polars_b200 as plb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    plb.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ext = torch.cuda.ExternalStream(plb.stream(), device=torch.device("cuda", local))
    peak_gbs, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        plb.sync()

    if a.workload == "groupby":
        key, vi, vf = gen_groupby(a.rows, a.keys, 1 + rank, a.skew)
        hkey, hvi, hvf = (plb.to_pinned(key), plb.to_pinned(vi), plb.to_pinned(vf)) if a.e2e_steps > 0 else (None, None, None)
        val_i = val_f = None
        if a.null_frac > 0:      # the "+5 % nulls" variant: independent validity bitmaps on both value columns
            nrng = np.random.default_rng(100 + rank)
            val_i, val_f = (plb.pack_bits(nrng.random(a.rows) >= a.null_frac) for _ in range(2))      # Arrow LSB bitmaps
        dkey, dvi, dvf = plb.to_device(key), plb.to_device(vi, val_i), plb.to_device(vf, val_f)
        del key, vi, vf
        spec = [("sum", np.int64), ("mean", np.float64), ("len", None)]
        out_bytes = 0
        peer_ex = None
        if world > 1:
            from polars_b200 import dist as pdist
            if a.exchange == "p2p":      # window region per source rank: every group of a rank could go to one peer
                try:
                    peer_ex = pdist.PeerExchange(plb, rows_per_src=min(a.keys, a.rows) + 1024, row_words=2 + 3)
                    ok_all = torch.tensor([1], device="cuda")
                except Exception as e:      # CUDA IPC unavailable (container policy): use the NCCL all-to-all instead
                    print(f"[bench] peer windows unavailable ({e}); falling back to --exchange nccl", file=sys.stderr)
                    ok_all = torch.tensor([0], device="cuda")
                dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
                if int(ok_all.item()) == 0:
                    a.exchange, peer_ex = "nccl", None

        def step_device():
            nonlocal out_bytes
            if world == 1:
                ok, outs = plb.group_by_agg(dkey.view(), [("sum", dvi.view()), ("mean", dvf.view()), ("len", None)], False, location=plb.DEVICE)
                out_bytes = ok.length * (8 + 8 + 8 + 4)
                return ok.length
            # partitioned plan (polars_b200/dist.py): local pre-aggregation -> hash partition of the partial
            # aggregates -> exchange (fused P2P stores over NVLink, or one NCCL all-to-all) -> merge
            vals = [dvi.view(), dvf.view(), None]
            if a.exchange == "p2p":
                ok, outs = pdist.partitioned_group_by_p2p(plb, peer_ex, dkey.view(), vals, spec, nullable=[False, False, False])
            else:
                ok, outs = pdist.partitioned_group_by(plb, dkey.view(), vals, spec, nullable=[False, False, False])
            out_bytes = ok.length * 28
            return ok.length

        def step_e2e():
            (k, _), outs = plb.group_by_agg(plb.Column(hkey), [("sum", plb.Column(hvi, val_i)), ("mean", plb.Column(hvf, val_f)), ("len", None)], False, location=plb.HOST)
            return k.size, k.nbytes + sum(o[0].nbytes for o in outs)

        unit_rows = a.rows
        alg_bytes_per_row = 24.0
        dominant = "k5_groupby_agg"
        metric = "group_by_agg_rows_per_sec"
        wl = (f"C2 hash group_by {a.rows} rows/GPU, {a.keys} {'Zipf(1.1)-skewed' if a.skew == 'zipf' else 'uniform'} Int64 keys, sum(v_i64)/mean(v_f64)/len"
              + (f", {a.null_frac:.0%} nulls per value column" if a.null_frac > 0 else "") + "; inputs 2.4 GB > L2 (no flush needed)")
        h2d = a.rows * 24
    elif a.workload == "q1":
        rows = a.rows if a.rows != 100_000_000 else 60_000_000      # SF10 lineitem ~ 6e7 rows
        a.rows = rows
        h = gen_lineitem(rows, 4 + rank)
        d = {k: plb.to_device(v) for k, v in h.items()}
        hp = {k: plb.to_pinned(v) for k, v in h.items()}
        exp = q1_numpy(h)
        ok, outs = q1_device(plb, d)
        k, _ = ok.to_numpy(); o = np.argsort(k)
        assert np.array_equal(k[o], exp["key"]) and np.array_equal(outs[7].to_numpy()[0][o], exp["len"]), "Q1 groups differ"
        for i, nm in ((0, "qty"), (1, "price"), (2, "dp"), (3, "ch")):
            assert np.allclose(outs[i].to_numpy()[0][o], exp[nm], rtol=1e-6), "Q1 sums differ: " + nm
        del h

        def step_device():
            ok, outs = q1_device(plb, d)
            return ok.length

        def step_e2e():
            dd = {k: plb.to_device(v) for k, v in hp.items()}
            ok, outs = q1_device(plb, dd)
            res = [ok.to_numpy()[0]] + [o.to_numpy()[0] for o in outs]
            return res[0].size, sum(r.nbytes for r in res)

        unit_rows = rows
        alg_bytes_per_row = 8.0 * 6        # K5 reads key + 5 distinct value columns
        dominant = "k5_groupby_agg_smem"
        metric = "pdsh_q1_rows_per_sec"
        wl = f"C4 PDS-H Q1 shape on {rows} synthetic lineitem rows (SF10-sized): filter + 4 expressions + group_by(returnflag,linestatus) with 8 aggregates"
        h2d = rows * 8 * 7
    else:
        probe, build = gen_join(a.rows, a.build_rows, 2 + rank, a.hit_frac, a.dup)
        hp, hb = plb.to_pinned(probe), plb.to_pinned(build)
        dp, db = plb.to_device(probe), plb.to_device(build)
        del probe, build

        def step_device():
            li, ri = plb.hash_join(dp.view(), db.view(), "inner", False, "none", location=plb.DEVICE)
            return li.length

        def step_e2e():
            (li, _), (ri, _) = plb.hash_join(plb.Column(hp), plb.Column(hb), "inner", False, "none", location=plb.HOST)
            return li.size, li.nbytes + ri.nbytes

        unit_rows = a.rows
        alg_bytes_per_row = 8.0 + 8.0 * a.hit_frac * a.dup      # probe key + (left_idx, right_idx) u32 per match
        dominant = "k8_join_probe"
        metric = "hash_join_probe_rows_per_sec"
        wl = (f"C3 inner hash join: probe {a.rows} x build {a.build_rows} Int64 keys ({'unique' if a.dup == 1 else str(a.dup) + ' copies of each'}), "
              f"{a.hit_frac:.0%} hit; outputs (left_idx,right_idx) u32")
        h2d = (a.rows + a.build_rows) * 8

    # ---- warm-up, then the timed region (device-resident inputs)
    for _ in range(max(a.warmup, 3)):
        n_out = step_device()
    barrier()
    plb.profile_reset()
    plb.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(ext):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            n_out = step_device()
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = plb.profile()
    launches = plb.launch_count()
    plb.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / a.steps
    value = unit_rows * world / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with pinned host buffers (H2D + compute + D2H per step)
    e2e_vals, d2h = [], 0
    for i in range(a.e2e_steps + 1 if a.e2e_steps > 0 else 0):
        barrier()
        t0 = time.perf_counter()
        _, d2h = step_e2e()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_vals.append(dt)
    te = torch.tensor([float(np.mean(e2e_vals)) if e2e_vals else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    clocks = sampler.stop() if rank == 0 else None      # sampled over the timed region and the e2e steps that follow it

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if a.workload == "join":      # hashed or dense probe, whichever ran
        dominant = max((k for k in prof if k.startswith("k8_") and "probe" in k), key=lambda k: prof[k]["ms"], default=dominant)
    traffic = None      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this workload
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{a.workload}:{a.rows}:{a.keys if a.workload == 'groupby' else a.build_rows}", {})
            traffic = t.get(dominant)
    except Exception:
        pass
    dom = prof.get(dominant, {"launches": 0, "ms": 0.0})
    dom_ms = dom["ms"] / max(dom["launches"], 1)
    achieved = (alg_bytes_per_row * unit_rows / 1e9) / (dom_ms / 1e3) if dom_ms > 0 else 0.0
    total_kernel_ms = sum(v["ms"] for v in prof.values()) / a.steps
    line = {
        "metric": metric, "value": value, "unit": "rows/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/float64", "data": "synthetic",
        "config": {"workload": wl, "rows_per_gpu": a.rows, "groups_out": int(n_out), "l2_policy": "inputs larger than L2",
                   "parallelism": "single GPU" if world == 1 else (f"hash-partitioned x{world}: local pre-agg + fused partition/P2P-store exchange over NVLink + merge" if a.exchange == "p2p" else f"hash-partitioned x{world}: local pre-agg + one NCCL all-to-all of partial aggregates + merge")},
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs if peak_gbs else None,
                     "traffic": traffic, "peak_source": peak_src, "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": alg_bytes_per_row * unit_rows,
                     "kernel_share_of_step": (dom_ms / total_kernel_ms) if total_kernel_ms else None},
        "kernels_ms_per_step": {k: v["ms"] / a.steps for k, v in prof.items()},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "knobs": {k: v for k, v in os.environ.items() if k.startswith("BL_")},
        "e2e": {"value": unit_rows * world / e2e_s if e2e_s else None, "unit": "rows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                "path": "bl_groupby_agg / bl_hash_join with BL_HOST columns in pinned memory -> BL_HOST outputs"},
    }
    if world == 1 and not a.no_cpu_baseline and a.workload in ("groupby", "join"):
        line["cpu_baseline"] = cpu_baseline(a)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _CudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def acero_child(a):
    """Child process of `acero_baseline`: times the query on pyarrow's Acero engine and prints one JSON object."""
    import pyarrow as pa
    if a.workload == "groupby":
        key, vi, vf = gen_groupby(a.rows, a.keys, 1, a.skew)
        t = pa.table({"key": key, "vi": vi, "vf": vf})
        t0 = time.perf_counter()
        t.group_by("key", use_threads=True).aggregate([("vi", "sum"), ("vf", "mean"), ([], "count_all")])
        dt = time.perf_counter() - t0
    else:
        probe, build = gen_join(a.rows, a.build_rows, 2, a.hit_frac, a.dup)
        lt, rt = pa.table({"key": probe}), pa.table({"key": build, "r": np.arange(build.size, dtype=np.int64)})
        t0 = time.perf_counter()
        lt.join(rt, keys="key", join_type="inner", use_threads=True)
        dt = time.perf_counter() - t0
    print(json.dumps({"engine": f"pyarrow-acero {pa.__version__}", "value": a.rows / dt, "unit": "rows/s", "threads": pa.cpu_count()}), flush=True)
    os._exit(0)      # Acero's worker threads occasionally abort the interpreter during static destruction


def acero_baseline(a, sample: int, build_rows: int):
    """Second, independent CPU reference: the same query on pyarrow's Acero engine with
    its default thread pool, in a child process.  Reported beside the oracle port; neither is the target."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--acero-child", "--workload", a.workload, "--rows", str(sample), "--keys", str(a.keys),
           "--build-rows", str(build_rows), "--skew", a.skew, "--hit-frac", str(a.hit_frac), "--dup", str(a.dup)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:      # optional evidence, never a reason to lose the bench line
        return {"engine": "pyarrow-acero", "unavailable": f"{type(e).__name__}: {e}"[:160]}


def cpu_baseline(a):
    import oracle
    oracle.build()
    cores = oracle.max_threads()
    sample = min(a.rows, a.cpu_sample)
    if a.workload == "groupby":
        key, vi, vf = gen_groupby(sample, a.keys, 1, a.skew)
        aggs = [("sum", vi, None), ("mean", vf, None), ("len", None, None)]
        t0 = time.perf_counter()
        oracle.group_by_agg(key, None, aggs, cores, False)
        dt = time.perf_counter() - t0
        n1 = min(sample, 4_000_000)      # single-thread point of the same port (all cores and 1 core)
        t1 = time.perf_counter()
        oracle.group_by_agg(key[:n1], None, [(k, None if v is None else v[:n1], m) for k, v, m in aggs], 1, False)
        one = n1 / (time.perf_counter() - t1)
        second = acero_baseline(a, sample, a.build_rows)
    else:
        build_rows = max(1, int(a.build_rows * sample / a.rows))
        probe, build = gen_join(sample, build_rows, 2, a.hit_frac, a.dup)
        t0 = time.perf_counter()
        oracle.hash_join(probe, build, None, None, "inner", False, "none", cores)
        dt = time.perf_counter() - t0
        n1 = min(sample, 4_000_000)
        t1 = time.perf_counter()
        oracle.hash_join(probe[:n1], build, None, None, "inner", False, "none", 1)
        one = n1 / (time.perf_counter() - t1)
        second = acero_baseline(a, sample, build_rows)
    return {"value": sample / dt, "unit": "rows/s", "cores": cores, "kind": "port",
            "sample": f"{sample} rows of the same workload, one pass; oracle = C/OpenMP restatement of the reference's partitioned algorithm (the Rust reference cannot be built here)",
            "single_thread": {"value": one, "unit": "rows/s", "cores": 1, "sample": f"{n1} rows"},
            "second_reference": second}


if __name__ == "__main__":
    main()
