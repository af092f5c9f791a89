"""Shared parity harness: the same golden / property checks run against the CPU oracle
(-m "not gpu") and against the CUDA path through the C ABI (-m gpu).

An "impl" is any object with
    group_by_agg(key, key_valid, aggs=[(kind, values, valid)], maintain_order) -> (keys, key_valid, [(vals, valid)])
    hash_join(lk, rk, lvalid, rvalid, how, nulls_equal, maintain_order) -> (left_idx, right_idx)
"""
from __future__ import annotations

import numpy as np

IDX_NULL = 0xFFFFFFFF
REL_TOL = 1e-6   # north_star: float aggregates within 1e-6 relative


def col(values, dtype):
    """list with None -> (numpy values, valid bool array | None)."""
    dt = np.dtype(dtype)
    valid = np.array([v is not None for v in values], dtype=np.bool_)
    arr = np.array([0 if v is None else v for v in values], dtype=dt)
    return arr, (None if valid.all() else valid)


def assert_close(got, exp, got_valid=None, exp_valid=None, what=""):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, f"{what}: shape {got.shape} != {exp.shape}"
    gv = np.ones(got.shape, bool) if got_valid is None else np.asarray(got_valid, bool)
    ev = np.ones(exp.shape, bool) if exp_valid is None else np.asarray(exp_valid, bool)
    assert np.array_equal(gv, ev), f"{what}: validity differs\n got {gv}\n exp {ev}"
    g, e = got[gv], exp[ev]
    if exp.dtype.kind == "f":
        nan_g, nan_e = np.isnan(g), np.isnan(e)
        assert np.array_equal(nan_g, nan_e), f"{what}: NaN pattern differs"
        g, e = g[~nan_g], e[~nan_e]
        inf = np.isinf(e)
        assert np.array_equal(g[inf], e[inf]), f"{what}: inf differs"
        g, e = g[~inf].astype(np.float64), e[~inf].astype(np.float64)
        tol = REL_TOL * np.maximum(np.abs(g), np.abs(e))
        bad = np.abs(g - e) > tol
        assert not bad.any(), f"{what}: {bad.sum()} values beyond rel {REL_TOL}: got {g[bad][:5]} exp {e[bad][:5]}"
    else:
        assert np.array_equal(g, e), f"{what}: integer mismatch\n got {g[:10]}\n exp {e[:10]}"


class OracleImpl:
    """Adapter: the CPU oracle behind the harness interface."""

    def __init__(self, n_threads=None):
        import oracle
        self.o = oracle
        self.n_threads = n_threads

    def group_by_agg(self, key, key_valid, aggs, maintain_order):
        k, kv, outs, _ = self.o.group_by_agg(key, key_valid, aggs, self.n_threads, maintain_order)
        return k, kv, outs

    def hash_join(self, lk, rk, lvalid=None, rvalid=None, how="inner", nulls_equal=False, maintain_order="none"):
        return self.o.hash_join(lk, rk, lvalid, rvalid, how, nulls_equal, maintain_order, self.n_threads)


def sort_groups(keys, key_valid, outs):
    """Order-insensitive comparison helper: sort groups by key (nulls last)."""
    kv = np.ones(keys.shape, bool) if key_valid is None else np.asarray(key_valid, bool)
    order = np.lexsort((keys, ~kv))
    outs2 = [(v[order], None if m is None else np.asarray(m)[order]) for v, m in outs]
    return keys[order], (None if key_valid is None else kv[order]), outs2


def run_group_by_kat(impl, case):
    # a KAT the reference runs "for dt in [Int8, Int16, Int32, Int64]" lists those widths in key_dtypes
    for kdt in case.get("key_dtypes", [case["key_dtype"]]):
        _run_group_by_kat_one(impl, dict(case, key_dtype=kdt))


def _run_group_by_kat_one(impl, case):
    if case.get("generated") == "overflow_mean":
        key = np.array([1, 2] * 50_000, dtype=case["key_dtype"])
        kvalid = None
        cols = [(np.full(100_000, 10_000_000, dtype=a["dtype"]), None) for a in case["aggs"]]
    else:
        key, kvalid = col(case["key"], case["key_dtype"])
        cols = [col(a["col"], a["dtype"]) if "col" in a else (None, None) for a in case["aggs"]]
    aggs = [(a["kind"], c[0], c[1]) for a, c in zip(case["aggs"], cols)]
    keys, kv, outs = impl.group_by_agg(key, kvalid, aggs, case["maintain_order"])
    if case.get("sort_by_key"):
        keys, kv, outs = sort_groups(keys, kv, outs)
    ek, ekv = col(case["expect_key"], case["key_dtype"])
    assert_close(keys, ek, kv, ekv, what=f"{case['cite']} keys")
    for a, (v, m) in zip(case["aggs"], outs):
        exp_dt = {"mean": "float64", "len": "uint32", "count": "uint32", "var": "float64", "std": "float64"}.get(a["kind"], a.get("dtype"))
        if a["kind"] == "mean" and a["dtype"] == "float32":
            exp_dt = "float32"
        e, em = col(a["expect"], exp_dt)
        assert v.dtype == np.dtype(exp_dt), f"{case['cite']} {a['kind']}: dtype {v.dtype} != {exp_dt}"
        assert_close(v, e, m, em, what=f"{case['cite']} {a['kind']}")


def run_join_kat(impl, case, set_threads=None):
    for kdt in case.get("key_dtypes", [case["key_dtype"]]):
        _run_join_kat_one(impl, dict(case, key_dtype=kdt), set_threads)


def _run_join_kat_one(impl, case, set_threads=None):
    lk, lv = col(case["left_key"], case["key_dtype"])
    rk, rv = col(case["right_key"], case["key_dtype"])
    for t in case.get("threads", [None]):
        if set_threads is not None:
            set_threads(t)
        li, ri = impl.hash_join(lk, rk, lv, rv, how=case["how"], nulls_equal=False, maintain_order=case["maintain_order"])
        assert li.dtype == np.uint32 and ri.dtype == np.uint32
        if case["exact_order"]:
            el = np.array(case["expect_left_idx"], np.uint32)
            er = np.array([IDX_NULL if x is None else x for x in case["expect_right_idx"]], np.uint32)
            assert np.array_equal(li, el), f"{case['cite']} threads={t}: left idx {li} != {el}"
            assert np.array_equal(ri, er), f"{case['cite']} threads={t}: right idx {ri} != {er}"
        elif "expect_height" in case:
            assert li.size == case["expect_height"] and ri.size == li.size, f"{case['cite']}: height {li.size}"
            assert int((ri == IDX_NULL).sum()) == case["expect_right_nulls"], case["cite"]
            hit = ri != IDX_NULL
            assert np.array_equal(lk[li[hit]], rk[ri[hit]]), f"{case['cite']}: joined keys differ"
        else:
            got = sorted(zip(li.tolist(), ri.tolist()))
            assert got == [tuple(p) for p in case["expect_pairs_sorted"]], f"{case['cite']}: {got}"
        if "expect_payload" in case:
            for name, exp in case["expect_payload"].items():
                if name.endswith("_right"):
                    src = np.array(case["payload_right"][name[:-6]])
                    got = src[ri]
                else:
                    src = np.array(case["payload_left"][name])
                    got = src[li]
                assert np.array_equal(got, np.array(exp)), f"{case['cite']} payload {name}"


def pairs_sorted(li, ri):
    p = (li.astype(np.uint64) << np.uint64(32)) | ri.astype(np.uint64)
    p.sort()
    return p


def run_group_by_multi_kat(group_by_agg_multi, case):
    """group_by_agg_multi(keys_list, valids_list, aggs, maintain_order) -> ([(key, valid)...], [(vals, valid)...])"""
    kv = [col(k, case["key_dtype"]) for k in case["keys"]]
    keys, valids = [k for k, _ in kv], [v for _, v in kv]
    if case["kind"] == "len":
        aggs = [("len", None, None)]
    else:
        c, cvalid = col(case["col"], case["dtype"])
        aggs = [(case["kind"], c, cvalid)]
    kouts, outs = group_by_agg_multi(keys, valids, aggs, True)
    vals = outs[0][0]
    if "expect_sorted" in case:
        assert sorted(vals.tolist()) == case["expect_sorted"], case["cite"]
    else:
        assert vals.tolist() == case["expect"], case["cite"]
    for (kvals, kvalid), exp in zip(kouts, case.get("expect_keys", [])):
        e, em = col(exp, case["key_dtype"])
        assert_close(kvals, e, kvalid, em, what=case["cite"] + " keys")
