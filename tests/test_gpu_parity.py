"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs, the reference's golden vectors, and size-independent properties.
Integer / index results must be bit-exact; float aggregates within 1e-6 relative."""
import os

import numpy as np
import pytest

import oracle
from helpers import IDX_NULL, assert_close, col, pairs_sorted, run_group_by_kat, run_join_kat, sort_groups

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plb():
    import polars_b200 as m
    m.init()
    return m


class GpuImpl:
    def __init__(self, plb):
        self.plb = plb

    def group_by_agg(self, key, key_valid, aggs, maintain_order):
        cols = {}
        specs = []
        for kind, vals, valid in aggs:
            if kind == "len":
                specs.append(("len", None))
                continue
            k = (id(vals), id(valid))
            if k not in cols:
                cols[k] = self.plb.Column(vals, valid)
            specs.append((kind, cols[k]))
        (k, kv), outs = self.plb.group_by_agg(self.plb.Column(key, key_valid), specs, maintain_order)
        return k, kv, outs

    def hash_join(self, lk, rk, lvalid=None, rvalid=None, how="inner", nulls_equal=False, maintain_order="none"):
        (li, _), (ri, _) = self.plb.hash_join(self.plb.Column(lk, lvalid), self.plb.Column(rk, rvalid), how, nulls_equal, maintain_order)
        return li, ri


# ------------------------------------------------------------------ golden vectors
def test_group_by_ordered_agg_kats(plb, kats):
    for case in kats["group_by_ordered"]:
        run_group_by_kat(GpuImpl(plb), case)


def test_group_by_kats(plb, kats):
    for case in kats["group_by"]:
        run_group_by_kat(GpuImpl(plb), case)


def test_join_kats(plb, kats):
    for case in kats["join"]:
        c = dict(case)
        c["threads"] = [None]
        run_join_kat(GpuImpl(plb), c)


def test_hash_partition_kat(plb, kats):
    vecs = kats["hash"][0]["vectors"]
    keys = np.array([int(v["key_u64"]) for v in vecs], dtype=np.uint64)
    payload = np.arange(keys.size, dtype=np.int64)
    for P in (1, 2, 3, 7, 8, 16):
        (k, _), [(pl, _)], offs = plb.hash_partition(keys, [payload], P)
        exp_part = np.array([v["part"][str(P)] for v in vecs])
        assert offs[-1] == keys.size
        for p in range(P):
            got = sorted(pl[offs[p]:offs[p + 1]].tolist())
            assert got == sorted(np.nonzero(exp_part == p)[0].tolist()), (P, p)
            assert np.array_equal(keys[pl[offs[p]:offs[p + 1]]], k[offs[p]:offs[p + 1]])


# ------------------------------------------------------------------ elementwise / compare
@pytest.mark.parametrize("dtype", ["int64", "int32", "uint64", "uint32", "float64", "float32"])
@pytest.mark.parametrize("n", [0, 1, 7, 255, 1000, 100_003])
def test_elementwise_vs_oracle(plb, dtype, n):
    rng = np.random.default_rng(n + 17)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        a = rng.normal(0, 100, n).astype(dt)
        b = rng.normal(0, 100, n).astype(dt)
        if n > 40:
            b[::17] = 0
            a[::31] = np.nan
            a[5], b[5] = np.inf, -np.inf
    else:
        lo = 0 if dt.kind == "u" else -1000
        a = rng.integers(lo, 1000, n).astype(dt)
        b = rng.integers(lo, 1000, n).astype(dt)
        if n > 40:
            b[::17] = 0
            if dt.kind == "i":
                a[0], b[0] = np.iinfo(dt).min, -1
    av = (rng.random(n) > 0.1) if n else None
    bv = (rng.random(n) > 0.2) if n else None
    for op in ("add", "sub", "mul", "floordiv", "mod", "truediv"):
        for (l, lv, r, rv) in [(a, av, b, bv), (a, None, b, None)]:
            exp, expv = oracle.arith(op, l, r, lv, rv)
            got, gotv = plb.elementwise(op, (l, lv), (r, rv))
            assert got.dtype == exp.dtype, (op, got.dtype, exp.dtype)
            if dt.kind == "f" or op == "truediv":
                assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)) or np.array_equal(got, exp, equal_nan=True), op
            else:
                assert np.array_equal(got, exp), op
            ev = np.ones(n, bool) if expv is None else expv
            gv = np.ones(n, bool) if gotv is None else gotv
            assert np.array_equal(gv, ev), op
        if n <= 1:      # (1, 1) is the array/array kernel in the reference too (ops/arity.rs:910-911)
            continue
        for s in ([3, 0, -1] if dt.kind == "i" else [3, 0]):
            sc = dt.type(s)
            exp, expv = oracle.arith(op, a, sc, av, None)
            got, gotv = plb.elementwise(op, (a, av), np.array([sc], dt))
            assert np.array_equal(got, exp, equal_nan=True), (op, s)
            assert np.array_equal(np.ones(n, bool) if gotv is None else gotv, np.ones(n, bool) if expv is None else expv), (op, s)
            exp, expv = oracle.arith(op, sc, b, None, bv)
            got, gotv = plb.elementwise(op, np.array([sc], dt), (b, bv))
            assert np.array_equal(got, exp, equal_nan=True), (op, s, "lhs")
            assert np.array_equal(np.ones(n, bool) if gotv is None else gotv, np.ones(n, bool) if expv is None else expv), (op, s, "lhs")


@pytest.mark.parametrize("dtype", ["int64", "int32", "uint64", "uint32", "float64", "float32"])
@pytest.mark.parametrize("n", [0, 1, 31, 64, 129, 4097, 100_003])
def test_compare_vs_oracle(plb, dtype, n):
    rng = np.random.default_rng(n + 3)
    dt = np.dtype(dtype)
    a = rng.integers(0, 50, n).astype(dt)
    b = rng.integers(0, 50, n).astype(dt)
    if dt.kind == "f" and n > 10:
        a[::7] = np.nan
        b[::11] = np.nan
        a[3], b[3] = -0.0, 0.0
    av = (rng.random(n) > 0.1) if n else None
    bv = (rng.random(n) > 0.2) if n else None
    for op in ("eq", "ne", "lt", "le", "gt", "ge"):
        exp, expv = oracle.compare(op, a, b, av, bv)
        got, gotv = plb.compare(op, (a, av), (b, bv))
        assert np.array_equal(got, exp), op
        assert np.array_equal(np.ones(n, bool) if gotv is None else gotv, np.ones(n, bool) if expv is None else expv), op
        if n:
            exp, expv = oracle.compare(op, a, dt.type(25), av, None)
            got, gotv = plb.compare(op, (a, av), np.array([25], dt))
            assert np.array_equal(got, exp), (op, "scalar")
    for op in ("eq", "ne"):
        exp, _ = oracle.compare(op, a, b, av, bv, missing=True)
        got, gotv = plb.compare(op, (a, av), (b, bv), missing=True)
        assert gotv is None and np.array_equal(got, exp), (op, "missing")


def test_sliced_inputs_bit_offsets(plb):
    # Arrow slices: element offset applies to values AND to the validity bitmap (arbitrary bit offset;
    # regression in the reference: crates/polars/tests/it/core/joins.rs:651-684 test_4_threads_bit_offset)
    rng = np.random.default_rng(0)
    n = 5000
    a = rng.integers(-100, 100, n).astype(np.int64)
    av = rng.random(n) > 0.3
    for off in (1, 3, 8, 13, 64, 77):
        ln = n - off - 5
        ca = plb.Column(a, av, offset=off, length=ln)
        got, gotv = plb.elementwise("add", ca, np.array([1], np.int64))
        assert np.array_equal(got, a[off:off + ln] + 1) and np.array_equal(gotv, av[off:off + ln])
        (k, kv), outs = plb.group_by_agg(ca, [("len", None), ("sum", ca)], True)
        ek, ekv, eo, _ = oracle.group_by_agg(a[off:off + ln], av[off:off + ln], [("len", None, None), ("sum", a[off:off + ln], av[off:off + ln])], 1, True)
        assert_close(k, ek, kv, ekv, "sliced keys")
        assert_close(outs[0][0], eo[0][0], what="len")
        assert_close(outs[1][0], eo[1][0], what="sum")


def test_chunked_inputs(plb):
    # py-polars/tests/unit/operations/test_group_by.py:1248-1257 (chunked == rechunked)
    rng = np.random.default_rng(1)
    n = 30_000
    key = rng.integers(0, 100, n).astype(np.int64)
    v = rng.normal(size=n)
    cuts = [0, 7, 5000, 5001, 20_000, n]
    kch = [plb.Column(key[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    vch = [plb.Column(v[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    (k1, _), o1 = plb.group_by_agg(kch, [("sum", vch), ("len", None)], True)
    (k2, _), o2 = plb.group_by_agg(key, [("sum", v), ("len", None)], True)
    assert np.array_equal(k1, k2) and np.array_equal(o1[1][0], o2[1][0])
    assert_close(o1[0][0], o2[0][0], what="chunked sum")


# ------------------------------------------------------------------ filter / gather
def test_filter_kat_generator(plb):
    # reference KAT generator py-polars/tests/unit/operations/test_filter.py:271-286
    for size in list(range(0, 64)) + [100, 1000, 10_000, 100_000]:
        for sel in (0.0, 0.01, 0.1, 0.5, 0.9, 0.99, 1.0):
            rng = np.random.Generator(np.random.PCG64(size * 100 + int(sel * 100)))
            mask = rng.random(size) < sel
            v64 = rng.integers(-2**62, 2**62, size).astype(np.int64)
            v32 = rng.normal(size=size).astype(np.float32)
            valid = rng.random(size) > 0.2
            mvalid = rng.random(size) > 0.1
            outs = plb.filter([(v64, valid), (v32, None)], (mask, mvalid))
            keep = mask & mvalid
            assert np.array_equal(outs[0][0], v64[keep]), (size, sel)
            gv = np.ones(keep.sum(), bool) if outs[0][1] is None else outs[0][1]
            assert np.array_equal(gv, valid[keep]), (size, sel)
            assert np.array_equal(outs[1][0], v32[keep]) and outs[1][1] is None


def test_filter_cmp_fused(plb):
    rng = np.random.default_rng(9)
    n = 1_000_003
    x = rng.integers(-10**6, 10**6, n).astype(np.int64)
    key = rng.integers(0, 1000, n).astype(np.int64)
    xv = rng.random(n) > 0.05
    outs = plb.filter_cmp([(x, xv), (key, None)], 0, "gt", 0)
    keep = (x > 0) & xv
    assert np.array_equal(outs[0][0], x[keep]) and np.array_equal(outs[1][0], key[keep])
    assert outs[0][1] is None    # kept rows of the predicate column are all valid


def test_gather(plb):
    rng = np.random.default_rng(4)
    for n, m in [(10, 0), (10, 5), (1000, 4099), (50_000, 200_001)]:
        v64 = rng.normal(size=n)
        v32 = rng.integers(-5, 5, n).astype(np.int32)
        valid = rng.random(n) > 0.3
        idx = rng.integers(0, n, m).astype(np.uint32)
        ivalid = rng.random(m) > 0.2
        outs = plb.gather([(v64, valid), (v32, None)], (idx, ivalid))
        e0, e0v = oracle.gather(v64, valid, idx, ivalid)
        e1, e1v = oracle.gather(v32, None, idx, ivalid)
        assert_close(outs[0][0], e0, outs[0][1], e0v, "gather f64")
        assert_close(outs[1][0], e1, outs[1][1], e1v, "gather i32")
        outs = plb.gather([(v64, None)], idx)
        assert np.array_equal(outs[0][0], v64[idx]) and outs[0][1] is None
    with pytest.raises(plb.OutOfBoundsError):
        plb.gather([np.arange(4.0)], np.array([1, 4], np.uint32))


# ------------------------------------------------------------------ group_by
def _gb_case(rng, n, k, nulls, key_dtype="int64"):
    key = rng.integers(-k // 2, k // 2 + 1, n).astype(key_dtype)
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    kvalid = (rng.random(n) > 0.05) if nulls else None
    ivalid = (rng.random(n) > 0.05) if nulls else None
    fvalid = (rng.random(n) > 0.3) if nulls else None
    return key, kvalid, vi, ivalid, vf, fvalid


@pytest.mark.parametrize("n,k,nulls", [(0, 5, False), (1, 1, False), (2, 1, True), (1001, 7, True), (50_001, 1000, True), (300_000, 100_000, False), (1_000_000, 3, True)])
def test_group_by_vs_oracle(plb, n, k, nulls):
    rng = np.random.default_rng(n + k)
    key, kvalid, vi, ivalid, vf, fvalid = _gb_case(rng, n, k, nulls)
    kinds = [("sum", vi, ivalid), ("mean", vf, fvalid), ("len", None, None), ("min", vi, ivalid), ("max", vf, fvalid),
             ("count", vf, fvalid), ("sum", vf, fvalid), ("mean", vi, ivalid), ("max", vi, ivalid), ("min", vf, fvalid)]
    for order in (True, False):
        for sub in (kinds[:3], kinds[3:7], kinds[7:]):
            keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, sub, order)
            ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, sub, 4, order)
            if not order:
                keys, kv, outs = sort_groups(keys, kv, outs)
                ek, ekv, eouts = sort_groups(ek, ekv, eouts)
            assert_close(keys, ek, kv, ekv, "keys")
            for (kind, _, _), (v, m), (ev, em) in zip(sub, outs, eouts):
                assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
                assert_close(v, ev, m, em, kind)


@pytest.mark.parametrize("key_dtype,val_dtype", [("int32", "int32"), ("uint32", "float32"), ("uint64", "uint64"), ("float64", "int64"), ("float32", "float64")])
def test_group_by_dtypes(plb, key_dtype, val_dtype):
    rng = np.random.default_rng(5)
    n = 20_000
    key = rng.integers(0, 50, n).astype(key_dtype)
    if np.dtype(key_dtype).kind == "f":
        key[::9] = np.nan
        key[1::9] = -0.0
        key[2::9] = 0.0
    val = (rng.uniform(0, 100, n) if np.dtype(val_dtype).kind == "f" else rng.integers(0, 1000, n)).astype(val_dtype)
    vvalid = rng.random(n) > 0.1
    aggs = [("sum", val, vvalid), ("mean", val, vvalid), ("min", val, vvalid), ("max", val, vvalid), ("count", val, vvalid)]
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, None, aggs, True)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, None, aggs, 1, True)
    assert keys.dtype == ek.dtype
    assert np.array_equal(keys.view(np.uint8), ek.view(np.uint8)), "key bits (first occurrence) differ"
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
        assert_close(v, ev, m, em, kind)


@pytest.mark.parametrize("key_dtype,val_dtype", [("int8", "int8"), ("int16", "uint16"), ("uint8", "int16"), ("uint16", "uint8"), ("int64", "int8")])
def test_group_by_small_ints(plb, key_dtype, val_dtype):
    # Int8/16, UInt8/16 value columns aggregate as Int64 (series/implementations/mod.rs:145-154): sum -> Int64,
    # min/max keep the dtype, mean -> Float64; small integer keys group on their bit pattern (into_groups.rs:178-185)
    rng = np.random.default_rng(9)
    n = 30_000
    ki, vi = np.iinfo(key_dtype), np.iinfo(val_dtype)
    key = rng.integers(max(ki.min, -60), min(ki.max, 60) + 1, n).astype(key_dtype)
    val = rng.integers(vi.min, int(vi.max) + 1, n).astype(val_dtype)
    vvalid = rng.random(n) > 0.1
    kvalid = rng.random(n) > 0.02
    aggs = [("sum", val, vvalid), ("mean", val, vvalid), ("min", val, vvalid), ("max", val, vvalid), ("count", val, vvalid), ("len", None, None)]
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, True)
    wide = val.astype(np.int64)
    eaggs = [(k, None if v is None else wide, m) for (k, v, m) in aggs]
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, eaggs, 1, True)
    assert keys.dtype == np.dtype(key_dtype)
    assert_close(keys, ek, kv, ekv, "keys")
    exp_dtype = {"sum": np.int64, "mean": np.float64, "min": val_dtype, "max": val_dtype, "count": np.uint32, "len": np.uint32}
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert v.dtype == np.dtype(exp_dtype[kind]), (kind, v.dtype)
        assert_close(v, ev.astype(v.dtype), m, em, kind)


@pytest.mark.parametrize("how,order", [("inner", "none"), ("inner", "left_right"), ("left", "none"), ("left", "right_left")])
def test_join_materialised_payloads(plb, how, order):
    # _finish_join (join/general.rs:17-49): both sides gathered at the join tuples; left-join misses are null rows
    rng = np.random.default_rng(44)
    nl, nr = 40_000, 9_000
    lk, rk = rng.integers(0, 12_000, nl).astype(np.int64), rng.integers(0, 12_000, nr).astype(np.int64)
    lvalid, rvalid = rng.random(nl) > 0.03, rng.random(nr) > 0.03
    lp, rp = rng.normal(size=nl), rng.integers(-9, 9, nr).astype(np.int32)
    rp_valid = rng.random(nr) > 0.2
    (lo_k, lo_p), (ro_p,) = plb.join(plb.Column(lk, lvalid), plb.Column(rk, rvalid), [plb.Column(lk, lvalid), plb.Column(lp)], [plb.Column(rp, rp_valid)],
                                     how=how, maintain_order=order)
    li, ri = oracle.hash_join(lk, rk, lvalid, rvalid, how, False, order, 4)
    assert np.array_equal(lo_k[0], lk[li]) and np.array_equal(lo_k[1] if lo_k[1] is not None else np.ones(li.size, bool), lvalid[li])
    assert np.array_equal(lo_p[0].view(np.uint64), lp[li].view(np.uint64))
    hit = ri != plb.IDX_NULL
    exp_valid = np.zeros(ri.size, bool); exp_valid[hit] = rp_valid[ri[hit]]
    got_valid = ro_p[1] if ro_p[1] is not None else np.ones(ri.size, bool)
    assert np.array_equal(got_valid, exp_valid)
    assert np.array_equal(ro_p[0][exp_valid], rp[ri[exp_valid]])


@pytest.mark.parametrize("dtype", [np.int64, np.float64, np.uint32])
@pytest.mark.parametrize("nulls_equal", [False, True])
def test_semi_anti_join_vs_oracle(plb, dtype, nulls_equal):
    # single_keys_semi_anti.rs:41-140: left rows in order with / without a match; duplicates on either side, nulls
    rng = np.random.default_rng(3)
    for nl, nr, kr in ((0, 5, 3), (7, 0, 3), (5000, 700, 1500), (200_000, 30_000, 50_000)):
        lk, rk = rng.integers(0, kr, nl).astype(dtype), rng.integers(0, kr, nr).astype(dtype)
        if np.dtype(dtype).kind == "f" and nl > 10:
            lk[::7] = np.nan; rk[::5] = np.nan; lk[1::9] = -0.0; rk[1::9] = 0.0
        lv, rv = rng.random(nl) > 0.1, rng.random(nr) > 0.1
        for how in ("semi", "anti"):
            li, ri = GpuImpl(plb).hash_join(lk, rk, lv, rv, how=how, nulls_equal=nulls_equal)
            eli, _ = oracle.hash_join(lk, rk, lv, rv, how, nulls_equal, "none", 4)
            assert ri.size == 0 and li.dtype == np.uint32
            assert np.array_equal(li, eli), (how, nl, nr)


def test_join_small_int_keys(plb):
    rng = np.random.default_rng(10)
    lk = rng.integers(-128, 128, 5000).astype(np.int8)
    rk = rng.integers(-128, 128, 300).astype(np.int8)
    lv, rv = rng.random(5000) > 0.05, rng.random(300) > 0.05
    for how in ("inner", "left"):
        li, ri = GpuImpl(plb).hash_join(lk, rk, lv, rv, how=how)
        eli, eri = oracle.hash_join(lk, rk, lv, rv, how, False, "none", 4)
        assert np.array_equal(li, eli) and np.array_equal(ri, eri)


def test_group_by_edge_semantics(plb):
    key = np.array([0, 0, 1, 1, 2, 2, 3, -2**63, -2**63], np.int64)
    vi = np.array([2**62, 2**62, 1, 2, 5, 6, 7, 1, 1], np.int64)
    valid = np.array([1, 1, 0, 0, 1, 0, 1, 1, 1], bool)
    aggs = [("sum", vi, valid), ("mean", vi, valid), ("min", vi, valid), ("count", vi, valid), ("len", None, None)]
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, None, aggs, True)
    assert keys.tolist() == [0, 1, 2, 3, -2**63]
    assert outs[0][0].tolist() == [-2**63, 0, 5, 7, 2] and outs[0][1] is None          # wrapping; all-null sum = 0
    assert outs[1][1].tolist() == [True, False, True, True, True]
    assert outs[2][0][2] == 5 and outs[3][0].tolist() == [2, 0, 1, 1, 2] and outs[4][0].tolist() == [2, 2, 2, 1, 2]
    vf = np.array([np.nan, 1.0, np.nan, np.nan, -0.0, 3.0, np.inf, 1.0, 2.0])
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, None, [("min", vf, None), ("max", vf, None), ("sum", vf, None)], True)
    assert outs[0][0][0] == 1.0 and np.isnan(outs[0][0][1]) and outs[1][0][2] == 3.0 and outs[0][0][3] == np.inf
    assert np.isnan(outs[2][0][0]) and np.isnan(outs[2][0][1])


def test_group_by_high_cardinality_restart(plb):
    # every key distinct: the sampled estimate must size (or regrow) the table correctly
    n = 400_000
    rng = np.random.default_rng(8)
    key = rng.permutation(n).astype(np.int64) * 7919
    v = np.ones(n, np.int64)
    (k, _), [(s, _), (c, _)] = plb.group_by_agg(key, [("sum", v), ("len", None)], False)
    assert k.size == n and np.array_equal(np.sort(k), np.sort(key)) and s.sum() == n and (c == 1).all()
    # heavy skew: a 64K sample sees few keys, the tail forces the regrow path
    key2 = np.where(rng.random(n) < 0.9, 5, key)
    (k, _), [(s, _), (c, _)] = plb.group_by_agg(key2, [("sum", v), ("len", None)], False)
    assert c.sum() == n and k.size == np.unique(key2).size


def test_group_by_smem_plan_overflow_falls_through(plb):
    # the sampled estimate (~900 groups) selects the shared-memory plan, but the tail holds ~4000
    # distinct keys: rows that do not fit the CTA-private table must take the global path, exactly
    n = 400_000
    rng = np.random.default_rng(21)
    key = np.where(rng.random(n) < 0.99, 5, rng.integers(10, 10**9, n)).astype(np.int64)
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    valid = rng.random(n) > 0.1
    aggs = [("sum", vi, valid), ("mean", vf, None), ("len", None, None), ("min", vf, None), ("max", vi, valid)]
    for order in (True, False):
        keys, kv, outs = GpuImpl(plb).group_by_agg(key, None, aggs, order)
        ek, ekv, eouts, _ = oracle.group_by_agg(key, None, aggs, 4, order)
        if not order:
            keys, kv, outs = sort_groups(keys, kv, outs)
            ek, ekv, eouts = sort_groups(ek, ekv, eouts)
        assert_close(keys, ek, kv, ekv, "keys")
        for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
            assert_close(v, ev, m, em, kind)


@pytest.mark.parametrize("hot", [0, 64, 1024])
def test_group_by_skewed_keys(plb, monkeypatch, hot):
    # Zipf(1.1) keys (SURVEY 8(d) C2 skew variant): a hot head plus a long tail of singletons.  The sampled
    # estimate (uniform inversion vs Chao's tail correction) only steers table sizing; the result must be
    # exact either way, also with the hot-table detour (BL_K5_HOT) and for sorted keys (runs of equal keys)
    if hot:
        monkeypatch.setenv("BL_K5_HOT", str(hot))
    n = 600_000
    rng = np.random.default_rng(33)
    key = (rng.zipf(1.1, n) % 200_000).astype(np.int64)
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    valid = rng.random(n) > 0.05
    aggs = [("sum", vi, valid), ("mean", vf, None), ("len", None, None), ("min", vf, None), ("max", vi, valid)]
    for k_in in (key, np.sort(key)):
        keys, kv, outs = GpuImpl(plb).group_by_agg(k_in, None, aggs, True)
        ek, ekv, eouts, _ = oracle.group_by_agg(k_in, None, aggs, 4, True)
        assert_close(keys, ek, kv, ekv, "keys")
        for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
            assert_close(v, ev, m, em, kind)


@pytest.mark.parametrize("n,k,nulls,dtype", [(0, 1, False, np.int64), (1, 1, False, np.int64), (33, 4, True, np.int64), (5000, 37, True, np.float64),
                                             (200_000, 5000, True, np.int64), (300_001, 250_000, False, np.int32), (1_000_000, 3, False, np.uint64)])
def test_group_tuples_vs_oracle(plb, n, k, nulls, dtype):
    # GroupsIdx{first, all} with sorted = true (hashing.rs:41-63,116-167): bit-exact first / offsets / index lists
    rng = np.random.default_rng(n + k)
    key = rng.integers(0, k, n).astype(dtype)
    if n > 10 and np.dtype(dtype).kind == "f":
        key[::7] = np.nan; key[1::11] = -0.0; key[2::11] = 0.0
    kvalid = (rng.random(n) > 0.1) if nulls else None
    g = oracle.group_by(key, kvalid, 4, True)
    first, offsets, all_ = plb.group_tuples(plb.Column(key, kvalid))
    assert first.dtype == np.uint32 and offsets.dtype == np.uint32 and all_.dtype == np.uint32
    assert np.array_equal(first, g.first)
    assert np.array_equal(offsets.astype(np.uint64), g.offsets)
    assert np.array_equal(all_, g.idx)


def _hot_case(rng, n, kind):
    """Keys with heavy hitters (warp-private accumulator rows in k_gb_consume_hot), incl. hot null / i64::MIN groups."""
    if kind == "zipf_i64":
        key = (rng.zipf(1.1, n) % 50_000).astype(np.int64)
        key[rng.random(n) < 0.15] = np.iinfo(np.int64).min          # the GB_EMPTY bit pattern as a frequent key
        kvalid = rng.random(n) > 0.2                                  # frequent null keys
    elif kind == "f64":
        key = (rng.zipf(1.3, n) % 1000).astype(np.float64)
        key[rng.random(n) < 0.1] = np.nan
        key[rng.random(n) < 0.1] = -0.0
        key[rng.random(n) < 0.1] = 0.0
        kvalid = None
    else:
        key = (rng.zipf(1.2, n) % 3000).astype(np.int32) - 7
        kvalid = rng.random(n) > 0.01
    return key, kvalid


@pytest.mark.parametrize("kind", ["zipf_i64", "f64", "i32"])
@pytest.mark.parametrize("hot_rows", [None, "0"])
def test_group_by_heavy_hitters(plb, monkeypatch, kind, hot_rows):
    # hot_rows "0": every key sampled >= 12 times is treated as a heavy hitter (exercises up to 62 rows per warp)
    if hot_rows is not None:
        monkeypatch.setenv("BL_K5_HOT_ROWS", hot_rows)
    rng = np.random.default_rng(77)
    n = 700_001
    key, kvalid = _hot_case(rng, n, kind)
    vi = rng.integers(-1000, 1000, n).astype(np.int32 if kind == "i32" else np.int64)
    vf = rng.uniform(-50, 100, n).round(6)
    vf[::97] = np.nan
    ivalid, fvalid = rng.random(n) > 0.1, rng.random(n) > 0.3
    aggs = [("sum", vi, ivalid), ("mean", vf, fvalid), ("len", None, None), ("min", vf, fvalid), ("max", vi, ivalid), ("count", vi, ivalid), ("max", vf, None), ("min", vi, None)]
    for order in (True, False):
        keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, order)
        ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 4, order)
        if not order:
            keys, kv, outs = sort_groups(keys, kv, outs)
            ek, ekv, eouts = sort_groups(ek, ekv, eouts)
        assert_close(keys, ek, kv, ekv, "keys")
        for (k_, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
            assert_close(v, ev, m, em, k_)
    # the same keys through the group-tuple path (hot rows carry len/first only) and a streamed two-batch consume
    g = oracle.group_by(key, kvalid, 4, True)
    first, offsets, all_ = plb.group_tuples(plb.Column(key, kvalid))
    assert np.array_equal(first, g.first) and np.array_equal(offsets.astype(np.uint64), g.offsets) and np.array_equal(all_, g.idx)
    st = plb.GroupBy(key.dtype, [("sum", vi.dtype), ("len", None)], track_first=True)
    half = n // 2 + 1
    for a, b in ((0, half), (half, n)):
        st.consume(plb.Column(key[a:b], None if kvalid is None else kvalid[a:b]), [plb.Column(vi[a:b], ivalid[a:b]), None], row_base=a)
    (k2, kv2), outs2 = st.finish(maintain_order=True)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, [("sum", vi, ivalid), ("len", None, None)], 4, True)
    assert_close(k2, ek, kv2, ekv, "stream keys")
    assert_close(outs2[0][0], eouts[0][0], outs2[0][1], eouts[0][1], "stream sum")
    assert_close(outs2[1][0], eouts[1][0], outs2[1][1], eouts[1][1], "stream len")


def _gpu_group_by_multi(plb):
    def run(keys, valids, aggs, order):
        kouts, outs = plb.group_by_agg_keys([plb.Column(k, v) for k, v in zip(keys, valids)], [(kind, None if vals is None else plb.Column(vals, valid)) for kind, vals, valid in aggs], order)
        return kouts, outs
    return run


def test_group_by_multi_kats(plb, kats):
    from helpers import run_group_by_multi_kat
    for case in kats["group_by_multi"]:
        run_group_by_multi_kat(_gpu_group_by_multi(plb), case)


@pytest.mark.parametrize("shape", ["two_i32", "u8_u8", "i64_f64_i16", "f32_i64_i64_u16", "nullable_wide"])
def test_group_by_multi_vs_oracle(plb, shape):
    # several key columns (row-encoding equality, group_by/mod.rs:88-94): packed into one 64-bit key when the widths
    # fit, otherwise through 32-bit group ids of the key so far; outputs = every key column at the group's first row
    rng = np.random.default_rng(len(shape))
    n = 150_001
    spec = {"two_i32": [(np.int32, 40, False), (np.int32, 30, True)],
            "u8_u8": [(np.uint8, 3, False), (np.uint8, 2, False)],
            "i64_f64_i16": [(np.int64, 50, False), (np.float64, 6, True), (np.int16, 4, False)],
            "f32_i64_i64_u16": [(np.float32, 5, False), (np.int64, 20, True), (np.int64, 7, False), (np.uint16, 3, True)],
            "nullable_wide": [(np.uint64, 9, True), (np.float64, 9, True), (np.int64, 9, True)]}[shape]
    keys, valids = [], []
    for dt, card, nullable in spec:
        k = rng.integers(-card // 2, card - card // 2, n) if np.dtype(dt).kind in "if" else rng.integers(0, card, n)
        k = k.astype(dt)
        if np.dtype(dt).kind == "f":
            k[rng.random(n) < 0.05] = np.nan; k[rng.random(n) < 0.05] = -0.0
        keys.append(k); valids.append((rng.random(n) > 0.1) if nullable else None)
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    vf = rng.uniform(0, 10, n)
    ivalid = rng.random(n) > 0.2
    aggs = [("sum", vi, ivalid), ("mean", vf, None), ("len", None, None), ("min", vi, ivalid), ("max", vf, None), ("count", vi, ivalid)]
    ekouts, eouts, _ = oracle.group_by_agg_multi(keys, valids, aggs, True)
    kouts, outs = _gpu_group_by_multi(plb)(keys, valids, aggs, True)
    for (k, kv), (ek, ekv) in zip(kouts, ekouts):
        assert k.dtype == ek.dtype
        assert_close(k, ek, kv, ekv, "keys")
        if k.dtype.kind == "f":      # first-occurrence bit patterns (-0.0 vs 0.0, NaN payload) survive
            ok = np.ones(k.size, bool) if ekv is None else ekv
            assert np.array_equal(k[ok].view(np.uint8), ek[ok].view(np.uint8))
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert_close(v, ev, m, em, kind)
    # unordered variant: same groups as a set
    kouts2, outs2 = _gpu_group_by_multi(plb)(keys, valids, [("len", None, None)], False)
    assert outs2[0][0].size == eouts[2][0].size and int(outs2[0][0].sum()) == n


def test_group_by_streaming_and_partials(plb):
    # streaming consume == one shot; export -> merge of partial aggregates == single table
    rng = np.random.default_rng(12)
    n = 200_000
    key, kvalid, vi, ivalid, vf, fvalid = _gb_case(rng, n, 5000, True)
    spec = [("sum", np.int64), ("mean", np.float64), ("len", None), ("min", np.float64), ("max", np.int64), ("count", np.int64)]
    aggs = [("sum", vi, ivalid), ("mean", vf, fvalid), ("len", None, None), ("min", vf, fvalid), ("max", vi, ivalid), ("count", vi, ivalid)]
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 4, True)

    def batch_cols(a, b):
        return [plb.Column(vi[a:b], ivalid[a:b]), plb.Column(vf[a:b], fvalid[a:b]), None, plb.Column(vf[a:b], fvalid[a:b]), plb.Column(vi[a:b], ivalid[a:b]), plb.Column(vi[a:b], ivalid[a:b])]

    g = plb.GroupBy(np.int64, spec, expected_groups=6000, track_first=True)
    cuts = [0, 50_001, 120_000, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        g.consume(plb.Column(key[a:b], kvalid[a:b]), batch_cols(a, b), row_base=a)
    (k, kv), outs = g.finish(maintain_order=True)
    assert_close(k, ek, kv, ekv, "stream keys")
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert_close(v, ev, m, em, "stream " + kind)
    # two "ranks": each pre-aggregates half, exports P=2 partitions, rank p merges partition p of both
    halves = [(0, n // 2), (n // 2, n)]
    states = []
    for a, b in halves:
        s = plb.GroupBy(np.int64, spec, expected_groups=6000, track_first=True)
        s.consume(plb.Column(key[a:b], kvalid[a:b]), batch_cols(a, b), row_base=a)
        states.append(s)
    exports = [s.export_partials(2) for s in states]
    # partition ids must equal the oracle's hash_to_partition on the key bits (null -> 0)
    finals = []
    for p in range(2):
        f = plb.GroupBy(np.int64, spec, expected_groups=6000, track_first=True)
        for ptr, rw, offs in exports:
            f.merge_partials(ptr + int(offs[p]) * rw * 8, int(offs[p + 1] - offs[p]))
        finals.append(f.finish(maintain_order=True))
    for ptr, _, _ in exports:
        plb.dev_free(ptr)
    part = oracle.hash_to_partition(oracle.dirty_hash(oracle.key_bits(ek)), 2)
    if ekv is not None:
        part = np.where(ekv, part, 0)
    for p in range(2):
        (k, kv), outs = finals[p]
        sel = part == p
        assert_close(k, ek[sel], kv, None if ekv is None or ekv[sel].all() else ekv[sel], f"partition {p} keys")
        for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
            assert_close(v, ev[sel], m, None if em is None or em[sel].all() else em[sel], f"partition {p} {kind}")


# ------------------------------------------------------------------ join
@pytest.mark.parametrize("nl,nr,krange,dups", [(0, 0, 10, 1), (5, 0, 10, 1), (0, 5, 10, 1), (300, 100, 150, 1), (3000, 2000, 500, 3), (40_000, 100_000, 20_000, 2), (300_000, 50_000, 50_000, 1)])
@pytest.mark.parametrize("key_dtype", ["int64", "int32", "float64"])
def test_join_vs_oracle(plb, nl, nr, krange, dups, key_dtype):
    rng = np.random.default_rng(nl * 7 + nr)
    lk = rng.integers(0, krange, nl).astype(key_dtype)
    rk = np.repeat(rng.permutation(max(krange, nr))[: max(nr // dups, 0)], dups)[:nr]
    rk = np.concatenate([rk, rng.integers(0, krange, nr - rk.size)]).astype(key_dtype)
    rng.shuffle(rk)
    if np.dtype(key_dtype).kind == "f" and nl > 10 and nr > 10:
        step_l, step_r = (13, 17) if nl <= 3000 else (nl // 40, nr // 30)     # NaN joins NaN: keep the cross product small
        lk[::step_l] = np.nan
        rk[::step_r] = np.nan
        lk[1::step_l] = -0.0
        rk[1::step_r] = 0.0
    lv = rng.random(nl) > 0.1
    rv = rng.random(nr) > 0.1
    impl = GpuImpl(plb)
    # nulls_equal joins every null key with every null key: quadratic output, small inputs only
    for nulls_equal in ((False, True) if nl <= 3000 else (False,)):
        for how in ("inner", "left"):
            orders = ["none", "left", "right", "left_right"] if how == "inner" else ["none", "left", "right", "right_left"]
            for order in orders:
                li, ri = impl.plb.hash_join(plb.Column(lk, lv), plb.Column(rk, rv), how, nulls_equal, order)
                li, ri = li[0], ri[0]
                eli, eri = oracle.hash_join(lk, rk, lv, rv, how, nulls_equal, order, 4)
                # exact sequence: the oracle's emission order is the reference's (hash_join/mod.rs:41-50)
                assert np.array_equal(li, eli), (how, nulls_equal, order, "left idx")
                assert np.array_equal(ri, eri), (how, nulls_equal, order, "right idx")


def test_join_dtype_mismatch_is_compute_error(plb):
    with pytest.raises(plb.ComputeError):
        plb.hash_join(np.arange(4, dtype=np.int64), np.arange(4, dtype=np.int32))


def test_join_then_gather_materialises(plb, kats):
    case = next(c for c in kats["join"] if "payload_left" in c)
    lk = np.array(case["left_key"], np.int32)
    rk = np.array(case["right_key"], np.int32)
    (li, _), (ri, _) = plb.hash_join(lk, rk)
    temp = np.array(case["payload_left"]["temp"])
    rain = np.array(case["payload_right"]["rain"])
    [(t, _)] = plb.gather([temp], li)
    [(r, _)] = plb.gather([rain], ri)
    assert t.tolist() == case["expect_payload"]["temp"] and r.tolist() == case["expect_payload"]["rain_right"]


# ------------------------------------------------------------------ device-resident path + properties at size
def test_device_resident_roundtrip(plb):
    rng = np.random.default_rng(2)
    n = 2_000_000
    key = rng.integers(0, 10_000, n).astype(np.int64)
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    dk, dv = plb.to_device(key), plb.to_device(v)
    ok, [osum, olen] = plb.group_by_agg(dk.view(), [("sum", dv.view()), ("len", None)], False, location=plb.DEVICE)
    k, _ = ok.to_numpy()
    s, _ = osum.to_numpy()
    c, _ = olen.to_numpy()
    order = np.argsort(k)
    exp = np.bincount(key, weights=None, minlength=10_000)
    assert np.array_equal(k[order], np.arange(10_000)) and np.array_equal(c[order], exp)
    esum = np.zeros(10_000, np.int64)
    np.add.at(esum, key, v)
    assert np.array_equal(s[order], esum)


def test_config_c1_filter_groupby_sum(plb):
    # BASELINE.json configs[0]: filter(x > 0).group_by(key).agg(x.sum()) on 10M Int64 rows, seed 0
    rng = np.random.default_rng(0)
    n = 10_000_000
    key = rng.integers(0, 10**3, n).astype(np.int64)
    x = rng.integers(-10**6, 10**6, n).astype(np.int64)
    fx, fk = plb.filter_cmp([x, key], 0, "gt", 0, location=plb.DEVICE)
    ok, [osum] = plb.group_by_agg(fk.view(), [("sum", fx.view())], False, location=plb.DEVICE)
    k, _ = ok.to_numpy()
    s, _ = osum.to_numpy()
    keep = x > 0
    esum = np.zeros(1000, np.int64)
    np.add.at(esum, key[keep], x[keep])
    order = np.argsort(k)
    assert np.array_equal(k[order], np.arange(1000)) and np.array_equal(s[order], esum)
    # size-independent checksum property: sum of group sums == sum of the filtered column
    assert int(s.sum()) == int(x[keep].sum()) and fx.length == int(keep.sum())


def test_large_properties_groupby_join(plb):
    # 2e7-row property checks (full-size configs run in bench.py): conservation of counts and sums,
    # and join round trip: gather(build_key, right_idx) == gather(probe_key, left_idx)
    rng = np.random.default_rng(3)
    n, K = 20_000_000, 1_000_000
    key = rng.integers(0, K, n).astype(np.int64)
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    dk, dv = plb.to_device(key), plb.to_device(v)
    ok, [osum, olen] = plb.group_by_agg(dk.view(), [("sum", dv.view()), ("len", None)], False, location=plb.DEVICE)
    k, _ = ok.to_numpy()
    s, _ = osum.to_numpy()
    c, _ = olen.to_numpy()
    assert int(c.astype(np.int64).sum()) == n and int(s.sum()) == int(v.sum()) and np.unique(k).size == k.size
    assert np.array_equal(np.sort(c), np.sort(np.bincount(key, minlength=K)[np.unique(key)]))
    nb = 2_000_000
    bkey = rng.permutation(nb).astype(np.int64)
    db = plb.to_device(bkey)
    pkey = rng.integers(0, 2 * nb, n).astype(np.int64)     # ~50 % hit
    dp = plb.to_device(pkey)
    li, ri = plb.hash_join(dp.view(), db.view(), "inner", False, "none", location=plb.DEVICE)
    [gl] = plb.gather([dp.view()], li.view(), location=plb.DEVICE)
    [gr] = plb.gather([db.view()], ri.view(), location=plb.DEVICE)
    a, _ = gl.to_numpy()
    b, _ = gr.to_numpy()
    l_idx, _ = li.to_numpy()
    assert np.array_equal(a, b) and a.size == int((pkey < nb).sum())
    assert np.all(np.diff(l_idx.astype(np.int64)) > 0)       # probe order, unique build keys


# ------------------------------------------------------------------ boundary behaviour
def test_join_chunked_sliced_nullable_keys(plb):
    # chunked ChunkedArray inputs with arbitrary validity bit offsets on both sides
    # (reference regression: crates/polars/tests/it/core/joins.rs:651-684 test_4_threads_bit_offset)
    rng = np.random.default_rng(31)
    nl, nr = 7001, 2503
    lk = rng.integers(0, 900, nl + 13).astype(np.int64)
    rk = rng.integers(0, 900, nr + 5).astype(np.int64)
    lv = rng.random(nl + 13) > 0.2
    rv = rng.random(nr + 5) > 0.2
    cuts_l = [0, 1, 64, 1999, nl]
    lch = [plb.Column(lk, lv, offset=13 + a, length=b - a) for a, b in zip(cuts_l[:-1], cuts_l[1:])]
    rch = [plb.Column(rk, rv, offset=5, length=1000), plb.Column(rk, rv, offset=1005, length=nr - 1000)]
    for how in ("inner", "left"):
        (li, _), (ri, _) = plb.hash_join(lch, rch, how)
        eli, eri = oracle.hash_join(lk[13:13 + nl], rk[5:5 + nr], lv[13:13 + nl], rv[5:5 + nr], how, False, "none", 3)
        assert np.array_equal(li, eli) and np.array_equal(ri, eri), how


def test_error_paths(plb):
    a = np.arange(10, dtype=np.int64)
    with pytest.raises(plb.B200Error) as e:
        plb.elementwise("add", a, np.arange(7, dtype=np.int64))           # lengths do not broadcast
    assert e.value.status == 1
    with pytest.raises(plb.ComputeError):
        plb.elementwise("add", a, a.astype(np.float64))                   # dtype mismatch
    with pytest.raises(plb.B200Error) as e:
        plb.group_by_agg(a > 4, [("len", None)])                          # Boolean keys are outside the hot path
    assert e.value.status == 4
    with pytest.raises(plb.B200Error):
        plb.filter([a], np.ones(9, bool))                                 # mask length mismatch
    # the library stays usable after errors
    out, _ = plb.elementwise("mul", a, np.array([3], np.int64))
    assert np.array_equal(out, a * 3)


def test_null_scalar_and_empty_inputs(plb):
    a = np.arange(5, dtype=np.float64)
    out, v = plb.elementwise("add", a, plb.Column(np.array([1.0]), np.array([False])))    # null scalar -> all null (arity.rs:916-922)
    assert v is not None and not v.any() and out.size == 5
    (k, _), outs = plb.group_by_agg(np.zeros(0, np.int64), [("sum", np.zeros(0, np.int64)), ("mean", np.zeros(0)), ("len", None)], True)
    assert k.size == 0 and all(o[0].size == 0 for o in outs)
    (li, _), (ri, _) = plb.hash_join(np.zeros(0, np.int64), np.arange(3, dtype=np.int64), "left")
    assert li.size == 0 and ri.size == 0


def test_group_by_multipass_beyond_l2(plb):
    # ~1.4e6 groups in 2e6 rows: the table (134 MB) cannot stay L2-resident, K5 runs two passes over slot sub-ranges;
    # null keys and the sentinel key (i64::MIN) must be handled by exactly one pass
    rng = np.random.default_rng(77)
    n = 2_000_001
    key = (rng.integers(0, 3_000_000, n) * 7919 - 10**9).astype(np.int64)
    key[::1000] = -2**63
    kvalid = rng.random(n) > 0.001
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    (k, kv), outs = plb.group_by_agg(plb.Column(key, kvalid), [("sum", plb.Column(vi)), ("mean", plb.Column(vf)), ("len", None)], False)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, [("sum", vi, None), ("mean", vf, None), ("len", None, None)], 8, False)
    k, kv, outs = sort_groups(k, kv, outs)
    ek, ekv, eouts = sort_groups(ek, ekv, eouts)
    assert_close(k, ek, kv, ekv, "keys")
    for name, (v, m), (ev, em) in zip(("sum", "mean", "len"), outs, eouts):
        assert_close(v, ev, m, em, name)


@pytest.mark.parametrize("shape", ["i64_c2", "i32_keys_u32_vals", "one_col_minmax", "high_card", "many_buckets"])
def test_group_by_radix_plan(plb, monkeypatch, shape):
    """K5r (groupby_radix.cu): histogram -> TMA tile scatter -> TMA-staged shared-memory aggregation.  Forced with
    BL_K5_RADIX=2 at sizes the oracle finishes quickly; the profile proves the partitioned kernels ran.  The sentinel key
    (i64::MIN = the pad marker of the record streams) and negative keys must survive."""
    monkeypatch.setenv("BL_K5_RADIX", "2")
    rng = np.random.default_rng(len(shape))
    n = 1_300_001
    if shape == "i64_c2":
        key = (rng.integers(0, 200_000, n) * 104729 - 10**10).astype(np.int64); key[::5000] = -2**63
        vi = rng.integers(-1000, 1000, n).astype(np.int64); vf = rng.uniform(0, 100, n).round(6)
        aggs = [("sum", vi, None), ("mean", vf, None), ("len", None, None)]
    elif shape == "i32_keys_u32_vals":
        key = rng.integers(-50_000, 50_000, n).astype(np.int32)
        vu = rng.integers(0, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32); vf = rng.normal(size=n).astype(np.float32)
        aggs = [("sum", vu, None), ("max", vu, None), ("mean", vf, None), ("min", vf, None)]
    elif shape == "one_col_minmax":
        key = rng.integers(0, 30_000, n).astype(np.uint64)
        vi = rng.integers(-2**62, 2**62, n).astype(np.int64)
        aggs = [("min", vi, None), ("max", vi, None), ("sum", vi, None), ("count", vi, None)]
    elif shape == "high_card":
        n = 2_500_000
        key = rng.integers(0, 2_000_000, n).astype(np.int64)
        vi = rng.integers(-1000, 1000, n).astype(np.int64); vf = rng.uniform(0, 100, n).round(6)
        aggs = [("sum", vi, None), ("mean", vf, None), ("len", None, None)]
    else:      # > 512 buckets: the unpadded store path of pass 1
        n = 3_000_000
        key = rng.integers(0, 1_500_000, n).astype(np.int64) * 3 + 1
        vi = rng.integers(-1000, 1000, n).astype(np.int64)
        aggs = [("sum", vi, None), ("len", None, None)]
    dk = plb.to_device(key)
    dv = {id(v): plb.to_device(v) for _, v, _ in aggs if v is not None}
    plb.profile_reset(); plb.profile_enable(True)
    ok, outs = plb.group_by_agg(dk.view(), [(kind, None if v is None else dv[id(v)].view()) for kind, v, _ in aggs], False, location=plb.DEVICE)
    prof = plb.profile(); plb.profile_enable(False)
    assert "k5r_scatter" in prof and "k5r_aggregate" in prof, sorted(prof)
    k, kv = ok.to_numpy()
    res = [o.to_numpy() for o in outs]
    ek, ekv, eouts, _ = oracle.group_by_agg(key, None, aggs, 8, False)
    k, kv, res = sort_groups(k, kv, res)
    ek, ekv, eouts = sort_groups(ek, ekv, eouts)
    assert_close(k, ek, kv, ekv, "keys")
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, res, eouts):
        assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
        assert_close(v, ev, m, em, kind)


@pytest.mark.parametrize("nl,nr,krange,dups", [(0, 0, 10, 1), (5, 0, 10, 1), (0, 5, 10, 1), (300, 100, 150, 1), (3000, 2000, 500, 3), (40_000, 100_000, 20_000, 2), (300_000, 50_000, 90_000, 1)])
@pytest.mark.parametrize("key_dtype", ["int64", "float64"])
def test_full_join_vs_oracle(plb, nl, nr, krange, dups, key_dtype):
    """BL_JOIN_FULL vs the oracle's hash_join_tuples_outer restatement: the probe-phase tuples are compared as an exact
    sequence, the drained build rows (order unpinned in the reference, ascending here and in the oracle) too."""
    rng = np.random.default_rng(nl * 11 + nr)
    lk = rng.integers(0, krange, nl).astype(key_dtype)
    rk = np.repeat(rng.permutation(max(krange, nr))[: max(nr // dups, 0)], dups)[:nr].astype(key_dtype)
    rng.shuffle(rk)
    nr = rk.size
    lv = rng.random(nl) > 0.1
    rv = rng.random(nr) > 0.1
    for nulls_equal in ((False, True) if nl <= 3000 else (False,)):
        (li, lvv), (ri, rvv) = plb.hash_join(plb.Column(lk, lv), plb.Column(rk, rv), "full", nulls_equal, "none")
        eli, eri = oracle.hash_join(lk, rk, lv, rv, "full", nulls_equal, "none", 4)
        assert np.array_equal(li, eli) and np.array_equal(ri, eri), (nulls_equal, li[:10], eli[:10], ri[:10], eri[:10])
        if li.size:
            assert np.array_equal(np.ones(li.size, bool) if lvv is None else lvv, eli != 0xFFFFFFFF)
            assert np.array_equal(np.ones(ri.size, bool) if rvv is None else rvv, eri != 0xFFFFFFFF)
    # materialised: unmatched sides come back as nulls
    if nl and nr:
        lp, rp = rng.normal(size=nl), rng.integers(0, 100, nr).astype(np.int64)
        louts, routs = plb.join(plb.Column(lk, lv), plb.Column(rk, rv), [lp], [rp], "full")
        eli, eri = oracle.hash_join(lk, rk, lv, rv, "full", False, "none", 4)
        lh, rh = eli != 0xFFFFFFFF, eri != 0xFFFFFFFF
        assert np.array_equal(louts[0][0][lh], lp[eli[lh]]) and np.array_equal(routs[0][0][rh], rp[eri[rh]])
        assert (louts[0][1] is None and lh.all()) or np.array_equal(louts[0][1], lh)
        assert (routs[0][1] is None and rh.all()) or np.array_equal(routs[0][1], rh)


def test_join_multi_kats_gpu(plb, kats):
    """Multi-column join keys (joins.rs:602-684 known answers) through bl_hash_join_keys."""
    from helpers import col
    for case in kats["join_multi"]:
        lk = [plb.Column(*col(k, case["key_dtype"])) for k in case["left_keys"]]
        rk = [plb.Column(*col(k, case["key_dtype"])) for k in case["right_keys"]]
        (li, _), (ri, _) = plb.hash_join_keys(lk, rk, case["how"], case["nulls_equal"], "none")
        assert li.tolist() == case["expect_left_idx"], case["cite"]
        assert ri.tolist() == [0xFFFFFFFF if x is None else x for x in case["expect_right_idx"]], case["cite"]


@pytest.mark.parametrize("shape", ["two_i32", "i64_f64", "u8_i16_i64", "three_i64_wide"])
@pytest.mark.parametrize("how", ["inner", "left", "semi", "anti", "full"])
def test_join_multi_vs_oracle(plb, shape, how):
    """prepare_keys_multiple semantics (join/mod.rs:658-678): nulls_equal = false -> a null in any key column nulls the row
    key; true -> nulls are part of the key.  Exact tuple sequence vs the oracle (hash_join_multi)."""
    rng = np.random.default_rng(len(shape) * 7 + len(how))
    nl, nr = 5000, 1200
    dts = {"two_i32": ["int32", "int32"], "i64_f64": ["int64", "float64"], "u8_i16_i64": ["uint8", "int16", "int64"], "three_i64_wide": ["int64", "int64", "int64"]}[shape]
    def gen(n, dt, wide):
        if np.dtype(dt).kind == "f":
            v = rng.integers(0, 6, n).astype(dt); v[rng.random(n) < 0.05] = np.nan; v[rng.random(n) < 0.05] = -0.0
            return v
        span = 6 if not wide else 3
        base = rng.integers(0, span, n)
        return (base * (2**40 + 12345) - 7 if wide and np.dtype(dt).itemsize == 8 else base).astype(dt)
    wide = shape == "three_i64_wide"
    L = [gen(nl, dt, wide) for dt in dts]; R = [gen(nr, dt, wide) for dt in dts]
    LV = [rng.random(nl) > 0.08 if i != 1 else None for i in range(len(dts))]
    RV = [rng.random(nr) > 0.08 if i != 0 else None for i in range(len(dts))]
    for nulls_equal in (False, True):
        (li, _), (ri, _) = plb.hash_join_keys([plb.Column(k, v) for k, v in zip(L, LV)], [plb.Column(k, v) for k, v in zip(R, RV)], how, nulls_equal, "none")
        eli, eri = oracle.hash_join_multi(L, R, LV, RV, how, nulls_equal, "none", 4)
        assert np.array_equal(li, eli), (shape, how, nulls_equal, li[:8], eli[:8])
        assert np.array_equal(ri, eri), (shape, how, nulls_equal, ri[:8], eri[:8])


def test_group_by_deterministic_mode_is_bit_exact(plb):
    """bl_set_deterministic(1): GroupsIdx + one sequential fold per group in row order with the reference's reducers
    (sequential Kahan, kahan_sum.rs:36-47) -> float sums and means are BIT-identical to the oracle (no tolerance), and
    identical from run to run.  Values span 12 orders of magnitude so that the addition order matters."""
    rng = np.random.default_rng(99)
    n = 400_000
    key = rng.integers(-500, 500, n).astype(np.int64)
    kvalid = rng.random(n) > 0.02
    vf = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 7, n)); fvalid = rng.random(n) > 0.1
    vs = (rng.normal(size=n) * 10.0 ** rng.integers(-3, 4, n)).astype(np.float32)
    vi = rng.integers(-2**62, 2**62, n).astype(np.int64)
    aggs = [("sum", vf, fvalid), ("mean", vf, fvalid), ("sum", vs, None), ("mean", vs, None), ("min", vf, fvalid), ("max", vs, None), ("sum", vi, None), ("mean", vi, None),
            ("count", vf, fvalid), ("len", None, None)]
    plb.set_deterministic(True)
    try:
        runs = []
        for _ in range(2):
            keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, True)
            runs.append((keys, kv, outs))
        ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 4, True)
        keys, kv, outs = runs[0]
        assert np.array_equal(keys[kv] if kv is not None else keys, ek[ekv] if ekv is not None else ek)
        for (kind, vals, _), (v, m), (ev, em), (v2, m2) in zip(aggs, outs, eouts, runs[1][2]):
            assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
            gv = np.ones(v.shape, bool) if m is None else m
            xv = np.ones(ev.shape, bool) if em is None else em
            assert np.array_equal(gv, xv), kind
            assert np.array_equal(v[gv].view(np.uint8), ev[xv].view(np.uint8)), (kind, "not bit-identical to the oracle")
            assert np.array_equal(v[gv].view(np.uint8), v2[gv].view(np.uint8)), (kind, "differs between runs")
        # several key columns take the same path
        (kouts, outs2) = plb.group_by_agg_keys([plb.Column(key, kvalid), plb.Column((key % 7).astype(np.int32))], [("sum", plb.Column(vf, fvalid)), ("len", None)], True)
        ekk, eo, _ = oracle.group_by_agg_multi([key, (key % 7).astype(np.int32)], [kvalid, None], [("sum", vf, fvalid), ("len", None, None)], True)
        assert np.array_equal(outs2[0][0].view(np.uint8), eo[0][0].view(np.uint8)) and np.array_equal(outs2[1][0], eo[1][0])
    finally:
        plb.set_deterministic(False)


@pytest.mark.parametrize("knob,value", [("BL_K5_SOA", "0"), ("BL_K5_PAIRS", "2"), ("BL_K5_HINT", "1"), ("BL_K5_HINT", "3"), ("BL_K5_MULTIPASS", "0"), ("BL_K5_SMEM", "0"),
                                        ("BL_K5_RADIX", "0"), ("BL_K5_LF", "30"), ("BL_K5_BULK", "0"), ("BL_K5_BULK", "2"), ("BL_K5_LEAN", "0")])
def test_group_by_knob_variants(plb, monkeypatch, knob, value):
    """Every documented BL_K5_* fallback (entry-major table, 4 rows per thread, L2 policy hints, single pass beyond L2, no
    CTA-private tables, ...) is read per call and has to give the reference's answer."""
    monkeypatch.setenv(knob, value)
    rng = np.random.default_rng(5)
    for n, k in ((200_001, 50_000), (2_000_001, 1_400_000)) if knob in ("BL_K5_MULTIPASS", "BL_K5_RADIX") else ((200_001, 50_000), (60_000, 300)):
        key = (rng.integers(0, k, n) * 7919 - 10**9).astype(np.int64); key[::997] = -2**63
        kvalid = rng.random(n) > 0.01
        vi = rng.integers(-1000, 1000, n).astype(np.int64); ivalid = rng.random(n) > 0.05
        vf = rng.uniform(0, 100, n).round(6)
        aggs = [("sum", vi, ivalid), ("mean", vf, None), ("len", None, None), ("max", vi, ivalid)]
        keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, False)
        ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 8, False)
        keys, kv, outs = sort_groups(keys, kv, outs)
        ek, ekv, eouts = sort_groups(ek, ekv, eouts)
        assert_close(keys, ek, kv, ekv, "keys")
        for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
            assert_close(v, ev, m, em, f"{knob}={value} {kind}")


@pytest.mark.parametrize("val_dtype,order", [("int64", False), ("int64", True), ("int32", False), ("uint64", True)])
def test_group_by_bulk_reduce_pair_layout(plb, monkeypatch, val_dtype, order):
    """General bulk kernel (BL_K5_BULK=2; by default only the lean kernel's shape takes the pair layout).
    Pair layout + TMA bulk reduce (len and the first integer sum share a 16-byte table cell): the integer sum is NOT the first
    accumulator here (the displaced word moves to the sum's plane), values wrap, columns carry nulls, maintain_order tracks
    `first` in the high half of the cell's first word, and an odd row count exercises the scalar tail."""
    monkeypatch.setenv("BL_K5_BULK", "2")
    monkeypatch.setenv("BL_K5_BULK_LANES", "32" if order else "20")
    rng = np.random.default_rng(11)
    n, k = 400_001, 120_000
    key = (rng.integers(0, k, n) * 104729 - 5 * 10**8).astype(np.int64); key[::1009] = -2**63
    kvalid = rng.random(n) > 0.02
    info = np.iinfo(val_dtype)
    vi = rng.integers(info.min // 2, info.max // 2, n, dtype=val_dtype); ivalid = rng.random(n) > 0.1
    vf = rng.uniform(-50, 50, n).round(6); fvalid = rng.random(n) > 0.2
    aggs = [("mean", vf, fvalid), ("max", vf, fvalid), ("sum", vi, ivalid), ("len", None, None), ("count", vi, ivalid), ("min", vi, None)]
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, order)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 8, order)
    if not order:
        keys, kv, outs = sort_groups(keys, kv, outs)
        ek, ekv, eouts = sort_groups(ek, ekv, eouts)
    assert_close(keys, ek, kv, ekv, "keys")
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
        assert_close(v, ev, m, em, f"pair layout {kind}")


@pytest.mark.parametrize("shape", ["c2", "three_cols", "sum_last", "uint_key", "c2_nulls", "null_keys"])
def test_group_by_lean_bulk_kernel(plb, shape):
    """k_gb_consume_lean (default for 8-byte keys / value columns): len + integer sum as one 16-byte bulk reduce per row, the other
    accumulators as REDs; i64::MIN keys (the table's EMPTY marker) take the special slot, wrapping sums, odd row count; with
    validity bitmaps on the values (null -> 0 in the paired sum, null counters) and on the key (null group)."""
    rng = np.random.default_rng(21)
    n, k = 500_001, 150_000
    key = (rng.integers(0, k, n) * 104729 - 5 * 10**8).astype(np.int64); key[::1013] = -2**63
    if shape == "uint_key":
        key = key.view(np.uint64)
    vi = rng.integers(-2**62, 2**62, n).astype(np.int64)
    vf = rng.uniform(-50, 50, n).round(6)
    vu = rng.integers(0, 2**63, n).astype(np.uint64)
    nulls = shape in ("c2_nulls", "null_keys")
    iv = (rng.random(n) > 0.05) if nulls else None
    fv = (rng.random(n) > 0.3) if nulls else None
    kvalid = (rng.random(n) > 0.001) if shape == "null_keys" else None
    aggs = {"c2": [("sum", vi, None), ("mean", vf, None), ("len", None, None)],
            "three_cols": [("sum", vi, None), ("max", vi, None), ("mean", vf, None), ("min", vf, None), ("sum", vu, None), ("len", None, None)],
            "sum_last": [("max", vf, None), ("mean", vf, None), ("count", vi, None), ("sum", vi, None)],
            "uint_key": [("sum", vu, None), ("min", vu, None), ("mean", vi, None), ("len", None, None)],
            "c2_nulls": [("sum", vi, iv), ("mean", vf, fv), ("len", None, None)],
            "null_keys": [("sum", vi, iv), ("count", vi, iv), ("mean", vf, fv), ("len", None, None)]}[shape]
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, False)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 8, False)
    keys, kv, outs = sort_groups(keys, kv, outs)
    ek, ekv, eouts = sort_groups(ek, ekv, eouts)
    assert_close(keys, ek, kv, ekv, "keys")
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
        assert_close(v, ev, m, em, f"lean {shape} {kind}")


def test_group_by_pair_layout_readers(plb):
    """Every reader of the pair table layout on one GPU: streaming consume with rehash (grow), hash-partitioned export + merge, and
    the fused P2P export into a peer window (here: our own) + device-side merge — the N-GPU plan's kernels with world size 1."""
    rng = np.random.default_rng(31)
    n, k = 600_000, 90_000
    key = (rng.permutation(n) % k).astype(np.int64) * 31 - 7
    key[: n // 3] = np.sort(key[: n // 3])           # the first batch sees few distinct keys: later batches force a rehash
    vi = rng.integers(-2**40, 2**40, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    spec = [("sum", np.int64), ("mean", np.float64), ("len", None)]
    nn = [False, False, False]
    uk, inv = np.unique(key, return_inverse=True)
    e_len = np.bincount(inv); e_sum = np.zeros(uk.size, np.int64); np.add.at(e_sum, inv, vi); e_mean = np.bincount(inv, weights=vf) / e_len

    def check(res, sel=None, tag=""):
        (kk, kv), outs = res
        o = np.argsort(kk, kind="stable")
        want = np.ones(uk.size, bool) if sel is None else sel
        assert kv is None and np.array_equal(kk[o], uk[want]), tag + " keys"
        assert np.array_equal(outs[0][0][o], e_sum[want]), tag + " sum"
        assert np.allclose(outs[1][0][o], e_mean[want], rtol=1e-9), tag + " mean"
        assert np.array_equal(outs[2][0][o].astype(np.int64), e_len[want]), tag + " len"

    # streaming with growth
    g = plb.GroupBy(np.int64, spec, nullable=nn)
    for a, b in ((0, 50_000), (50_000, 200_000), (200_000, n)):
        g.consume(key[a:b], [vi[a:b], vf[a:b], None], row_base=a)
    check(g.finish(), tag="stream")
    # export (2 partitions) + merge
    s = plb.GroupBy(np.int64, spec, nullable=nn)
    s.consume(key, [vi, vf, None])
    ptr, rw, offs = s.export_partials(2)
    part = oracle.hash_to_partition(oracle.dirty_hash(oracle.key_bits(uk)), 2)
    for p_ in range(2):
        f = plb.GroupBy(np.int64, spec, expected_groups=k + 1000, nullable=nn)
        f.merge_partials(ptr + int(offs[p_]) * rw * 8, int(offs[p_ + 1] - offs[p_]))
        check(f.finish(), part == p_, f"export partition {p_}")
    plb.dev_free(ptr)
    # fused P2P export into our own window + merge on the device
    rows_per_src = k + 1024
    win = plb.Window(1024 + rows_per_src * rw * 8)
    s.export_partials_p2p_async([win.ptr], 0, rows_per_src, 1)
    f = plb.GroupBy(np.int64, spec, expected_groups=k + 1000, nullable=nn)
    f.merge_window_async(win.ptr, 1, rows_per_src, 1)
    check(f.finish(), tag="p2p window")
    del s, f, g
    win.destroy()


@pytest.mark.parametrize("knob,value", [("BL_JOIN_FUSED", "0"), ("BL_JOIN_TABLE", "compact"), ("BL_JOIN_DENSE", "0")])
def test_join_knob_variants(plb, monkeypatch, knob, value):
    monkeypatch.setenv(knob, value)
    rng = np.random.default_rng(6)
    for nl, nr, dups in ((50_000, 20_000, 1), (30_000, 9_000, 3)):
        lk = rng.integers(0, 25_000, nl).astype(np.int64)
        rk = np.repeat(rng.permutation(25_000)[: nr // dups], dups).astype(np.int64); rng.shuffle(rk)
        lv = rng.random(nl) > 0.05; rv = rng.random(rk.size) > 0.05
        for how in ("inner", "left", "semi", "anti", "full"):
            (li, _), (ri, _) = plb.hash_join(plb.Column(lk, lv), plb.Column(rk, rv), how, False, "none")
            eli, eri = oracle.hash_join(lk, rk, lv, rv, how, False, "none", 4)
            assert np.array_equal(li, eli) and np.array_equal(ri, eri), (knob, value, how, dups)


@pytest.mark.parametrize("val_dtype", ["float64", "float32", "int64", "int32", "int16"])
def test_group_by_first_last_var_std(plb, val_dtype):
    """Aggregations outside the fused set fold per group over GroupsIdx in row order (groupby_exact.cu): first / last incl.
    nulls, Welford var / std with ddof (take_agg/var.rs:11-41) — bit-identical to the oracle's restatement."""
    rng = np.random.default_rng(17)
    n = 120_000
    key = rng.integers(-300, 300, n).astype(np.int64)
    kvalid = rng.random(n) > 0.02
    if np.dtype(val_dtype).kind == "f":
        x = (rng.normal(1e6, 3.0, n)).astype(val_dtype)       # large mean, small spread: the naive sum-of-squares formula would lose everything
    else:
        x = rng.integers(-30000, 30000, n).astype(val_dtype)
    xvalid = rng.random(n) > 0.15
    key[-5:] = 9999; xvalid[-5:] = False                      # an all-null group
    aggs = [("first", x, xvalid), ("last", x, xvalid), ("var", x, xvalid), ("std:0", x, xvalid), ("var:2", x, None), ("mean", x, xvalid), ("len", None, None)]
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, aggs, True)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, aggs, 4, True)
    assert_close(keys, ek, kv, ekv, "keys")
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert v.dtype == ev.dtype, (kind, v.dtype, ev.dtype)
        gv = np.ones(v.shape, bool) if m is None else m
        xv = np.ones(ev.shape, bool) if em is None else em
        assert np.array_equal(gv, xv), kind
        assert np.array_equal(v[gv].view(np.uint8), ev[xv].view(np.uint8)), (kind, "not bit-identical")


def test_pdsh_q1_from_the_reference_schema(plb):
    """PDS-H Q1 on the reference's own lineitem sample (examples/datasets/pds_heads/lineitem.feather -> tests/golden/
    pdsh_lineitem_head.json): string flags enter as dictionary codes (the physical representation Polars groups
    Categorical / Enum keys on), the timestamp[us] ship date as Int64, quantity as Int64 — filter (K2+K3), expressions (K1),
    two-column group_by (K5) — against a direct numpy evaluation of the query."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "pdsh_lineitem_head.json")) as f:
        li = json.load(f)["columns"]
    qty = np.array(li["l_quantity"], np.int64); price = np.array(li["l_extendedprice"]); disc = np.array(li["l_discount"]); tax = np.array(li["l_tax"])
    ship = np.array(li["l_shipdate"], np.int64)
    rf_dict, rf = np.unique(np.array(li["l_returnflag"]), return_inverse=True)      # dictionary encoding on the host
    ls_dict, ls = np.unique(np.array(li["l_linestatus"]), return_inverse=True)
    rf, ls = rf.astype(np.uint32), ls.astype(np.uint32)
    cutoff = np.int64((np.datetime64("1998-09-02") - np.datetime64("1970-01-01")) // np.timedelta64(1, "us"))
    D = plb.DEVICE
    f_ = plb.filter_cmp([ship, qty, price, disc, tax, rf, ls], 0, "le", cutoff, location=D)
    _, fq, fp, fd, ft, frf, fls = f_
    one_minus = plb.elementwise("sub", np.array([1.0]), fd.view(), location=D)
    disc_price = plb.elementwise("mul", fp.view(), one_minus.view(), location=D)
    charge = plb.elementwise("mul", disc_price.view(), plb.elementwise("add", ft.view(), np.array([1.0]), location=D).view(), location=D)
    kouts, outs = plb.group_by_agg_keys([frf.view(), fls.view()], [("sum", fq.view()), ("sum", fp.view()), ("sum", disc_price.view()), ("sum", charge.view()),
                                                                     ("mean", fq.view()), ("mean", fp.view()), ("mean", fd.view()), ("len", None)], True)
    m = ship <= cutoff
    seen, exp = [], {}
    for i in np.nonzero(m)[0]:
        k = (int(rf[i]), int(ls[i]))
        if k not in exp:
            seen.append(k); exp[k] = []
        exp[k].append(i)
    assert [(int(a), int(b)) for a, b in zip(kouts[0][0], kouts[1][0])] == seen
    dp = price * (1.0 - disc); ch = dp * (tax + 1.0)
    for g, k in enumerate(seen):
        r = np.array(exp[k])
        assert outs[0][0][g] == qty[r].sum() and outs[7][0][g] == r.size
        for o, ref in ((1, price[r].sum()), (2, dp[r].sum()), (3, ch[r].sum()), (4, qty[r].mean()), (5, price[r].mean()), (6, disc[r].mean())):
            assert abs(outs[o][0][g] - ref) <= 1e-9 * abs(ref), (o, k)
    assert [str(rf_dict[a]) + str(ls_dict[b]) for a, b in seen] == ["NO", "RF", "AF"]      # the groups of the sample, first-occurrence order


# ------------------------------------------------------------------ string keys
def test_string_keys_kats(plb, kats):
    """The reference's string-key group_by tests end to end on the device: bl_string_encode -> bl_groupby_agg on the codes ->
    bl_string_gather of the group keys."""
    for case in kats["group_by_strings"]:
        col_s = plb.StringColumn(case["key"])
        (codes, cvalid), nd = plb.string_encode(col_s)
        assert codes.dtype == np.uint32 and codes.tolist() == case["expect_codes"] and cvalid is None, case["cite"]
        assert nd == len(case["expect_key"])
        cols = [col(a["col"], a["dtype"]) for a in case["aggs"]]
        aggs = [(a["kind"], c[0], c[1]) for a, c in zip(case["aggs"], cols)]
        keys, kv, outs = GpuImpl(plb).group_by_agg(codes, None, aggs, True)
        assert [b.decode() for b in plb.string_gather(col_s, keys)] == case["expect_key"], case["cite"]
        for a, (v, m) in zip(case["aggs"], outs):
            e, em = col(a["expect"], {"mean": "float64", "len": "uint32", "count": "uint32"}.get(a["kind"], a["dtype"]))
            assert_close(v, e, m, em, what=f"{case['cite']} {a['kind']}")


def _random_strings(rng, n, distinct, nulls):
    pool = []
    for i in range(distinct):
        ln = int(rng.integers(0, 40)) if i % 50 else int(rng.integers(200, 3000))      # a few long values
        pool.append(bytes(rng.integers(0, 256, ln, dtype=np.uint8)) if i % 3 else ("key-%d-" % i * (ln // 8 + 1))[:ln].encode())
    pool[0] = b""                                                                       # the empty string is a value, not a null
    vals = [pool[int(j)] for j in rng.integers(0, distinct, n)]
    if nulls:
        vals = [None if rng.random() < 0.07 else v for v in vals]
    return vals


@pytest.mark.parametrize("n,distinct,nulls", [(0, 1, False), (1, 1, True), (33, 5, True), (5_000, 4_000, True), (200_000, 30_000, False), (200_001, 7, True)])
def test_string_encode_vs_oracle(plb, n, distinct, nulls):
    rng = np.random.default_rng(n * 7 + distinct)
    vals = _random_strings(rng, n, distinct, nulls)
    ecodes, evalid, end = oracle.string_codes(vals)
    (codes, cvalid), nd = plb.string_encode(plb.StringColumn(vals))
    assert nd == end
    assert_close(codes, ecodes, cvalid, evalid, "string codes")
    if n >= 33:
        # chunked + sliced input (Arrow offsets): chunk 0 = rows [0, a) of a column with 3 leading rows, chunk 1 = the rest
        a = n // 3
        lead = [b"zz", None, b"lead"]
        c0 = plb.StringColumn(lead + vals[:a], offset=3, length=a)
        c1 = plb.StringColumn(vals[a:] + [b"tail"], offset=0, length=n - a)
        (codes2, cvalid2), nd2 = plb.string_encode([c0, c1])
        assert nd2 == end
        assert_close(codes2, ecodes, cvalid2, evalid, "string codes (chunked, sliced)")


def test_string_group_by_and_gather(plb):
    """group_by on a string key with nulls == the oracle's group_by on the oracle's codes; gather materialises the keys
    (null group -> null key; BL_IDX_NULL / null index -> null; out-of-range index -> OutOfBoundsError)."""
    rng = np.random.default_rng(77)
    n = 120_001
    vals = _random_strings(rng, n, 9_000, True)
    vi = rng.integers(-10**6, 10**6, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    aggs = [("sum", vi, None), ("mean", vf, None), ("len", None, None)]
    scol = plb.StringColumn(vals)
    (codes, cvalid), _ = plb.string_encode(scol)
    keys, kv, outs = GpuImpl(plb).group_by_agg(codes, cvalid, aggs, True)
    ecodes, evalid, _ = oracle.string_codes(vals)
    ek, ekv, eouts, _ = oracle.group_by_agg(ecodes, evalid, aggs, 4, True)
    assert_close(keys, ek, kv, ekv, "string group keys (codes)")
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert_close(v, ev, m, em, "string group_by " + kind)
    got = plb.string_gather(scol, (keys, kv))
    exp = [None if (kv is not None and not kv[i]) else vals[int(keys[i])] for i in range(keys.size)]
    assert got == exp
    idx = np.array([0, plb.IDX_NULL, n - 1, 5, 5], np.uint32)
    got = plb.string_gather(scol, (idx, np.array([True, True, True, False, True])))
    assert got == [vals[0], None, vals[n - 1], None, vals[5]]
    with pytest.raises(plb.OutOfBoundsError):
        plb.string_gather(scol, np.array([n], np.uint32))


def test_string_group_by_one_call_and_join(plb):
    """bl_groupby_agg_strings == group_by on the oracle's codes with the keys gathered back; bl_hash_join_strings == the oracle's
    join on codes computed over both sides together (exact tuple order), for every join kind, with null keys."""
    rng = np.random.default_rng(78)
    n = 60_000
    vals = _random_strings(rng, n, 3_000, True)
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    vf = rng.uniform(0, 100, n).round(6)
    aggs = [("sum", vi, None), ("mean", vf, None), ("len", None, None)]
    gkeys, outs = plb.group_by_agg_strings(plb.StringColumn(vals), [("sum", plb.Column(vi)), ("mean", plb.Column(vf)), ("len", None)], True)
    ecodes, evalid, _ = oracle.string_codes(vals)
    ek, ekv, eouts, _ = oracle.group_by_agg(ecodes, evalid, aggs, 4, True)
    assert gkeys == [None if (ekv is not None and not ekv[i]) else vals[int(ek[i])] for i in range(ek.size)]
    for (kind, _, _), (v, m), (ev, em) in zip(aggs, outs, eouts):
        assert_close(v, ev, m, em, "one-call string group_by " + kind)
    # join
    nl, nr = 40_000, 9_000
    pool = _random_strings(rng, 6_000, 6_000, False)
    left = [None if rng.random() < 0.03 else pool[int(j)] for j in rng.integers(0, 6_000, nl)]
    right = [None if rng.random() < 0.03 else pool[int(j)] for j in rng.integers(0, 4_000, nr)]
    codes, cvalid, _ = oracle.string_codes(left + right)
    lk, rk = codes[:nl], codes[nl:]
    lv = None if cvalid is None else cvalid[:nl]
    rv = None if cvalid is None else cvalid[nl:]
    for how in ("inner", "left", "semi", "anti", "full"):
        for ne in (False, True):
            (li, _), (ri, _) = plb.hash_join_strings(plb.StringColumn(left), plb.StringColumn(right), how, ne, "none")
            eli, eri = oracle.hash_join(lk, rk, lv, rv, how, ne, "none", 4)
            assert np.array_equal(li, eli) and np.array_equal(ri, eri), (how, ne)


@pytest.mark.parametrize("val_dtype,nulls", [("int64", False), ("int64", True), ("float64", True), ("int32", True), ("uint64", False)])
def test_group_by_n_unique(plb, kats, val_dtype, nulls):
    """BL_AGG_N_UNIQUE (agg_n_unique: a null counts as one value, NaN == NaN, -0.0 == 0.0) next to fused aggregations, against
    the reference's KAT and the oracle's restatement; null keys form their own group."""
    for case in kats["group_by_n_unique"]:
        key, kvalid = col(case["key"], case["key_dtype"])
        for c in case["cols"]:
            v, valid = col(c["col"], c["dtype"])
            _, _, outs = GpuImpl(plb).group_by_agg(key, kvalid, [("n_unique", v, valid)], True)
            assert outs[0][0].dtype == np.uint32 and outs[0][0].tolist() == c["expect"] and outs[0][1] is None, case["cite"]
    rng = np.random.default_rng(41)
    n = 150_001
    key = rng.integers(-500, 500, n).astype(np.int64)
    kvalid = (rng.random(n) > 0.02) if nulls else None
    if val_dtype == "float64":
        v = rng.integers(-20, 20, n).astype(np.float64); v[rng.random(n) < 0.05] = np.nan; v[rng.random(n) < 0.05] = -0.0
    else:
        info = np.iinfo(val_dtype)
        v = rng.integers(max(info.min, -30), 30, n).astype(val_dtype)
    valid = (rng.random(n) > 0.1) if nulls else None
    vi = rng.integers(-1000, 1000, n).astype(np.int64)
    keys, kv, outs = GpuImpl(plb).group_by_agg(key, kvalid, [("sum", vi, None), ("n_unique", v, valid), ("len", None, None)], True)
    firsts, counts = oracle.group_n_unique(key, kvalid, v, valid)
    ek, ekv, eouts, _ = oracle.group_by_agg(key, kvalid, [("sum", vi, None), ("len", None, None)], 4, True)
    assert_close(keys, ek, kv, ekv, "n_unique keys")
    assert_close(outs[0][0], eouts[0][0], outs[0][1], eouts[0][1], "sum beside n_unique")
    assert outs[1][0].dtype == np.uint32 and outs[1][1] is None and np.array_equal(outs[1][0], counts), "n_unique"
    assert_close(outs[2][0], eouts[1][0], outs[2][1], eouts[1][1], "len beside n_unique")
