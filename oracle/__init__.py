"""CPU oracle — numpy front-end over oracle/oracle.c (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package, and only as the checker / CPU baseline.  polars_b200/ never imports it.

Each wrapper cites the reference file:line in oracle.c.  Conventions:
  * validity arrays are numpy bool (one byte per row) or None (= no nulls);
  * IdxSize = uint32, null index = 0xFFFFFFFF;
  * group_by keys are passed through key_bits() (unsigned bit repr / canonical float bits,
    polars-core/src/frame/group_by/into_groups.rs:142-191, polars-utils/src/total_ord.rs:37-47).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
IDX_NULL = np.uint32(0xFFFFFFFF)

OPS = {"add": 0, "sub": 1, "mul": 2, "floordiv": 3, "mod": 4, "truediv": 5}
CMPS = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}
_SUFFIX = {np.dtype("int64"): "i64", np.dtype("int32"): "i32", np.dtype("uint64"): "u64",
           np.dtype("uint32"): "u32", np.dtype("float64"): "f64", np.dtype("float32"): "f32"}


def build(force: bool = False) -> str:
    """Compile oracle.c with the committed Makefile (gcc; seconds)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.or_group_by.restype = C.c_int64
        _lib.or_filter.restype = C.c_int64
        _lib.or_hash_join.restype = C.c_int64
        _lib.or_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().or_max_threads())


def hw_threads() -> int:
    """Logical CPUs OpenMP can see (independent of OMP_NUM_THREADS)."""
    lib().or_hw_threads.restype = C.c_int
    return int(lib().or_hw_threads())


def set_threads(n: int) -> None:
    """Pool size of the aggregation / probe loops (the reference runs them on its Rayon pool)."""
    lib().or_set_threads(C.c_int(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _valid(v, n):
    if v is None:
        return None
    v = np.ascontiguousarray(v, dtype=np.bool_)
    assert v.shape == (n,)
    return v


# ------------------------------------------------------------------ hashing
def dirty_hash(keys_u64: np.ndarray) -> np.ndarray:
    k = np.ascontiguousarray(keys_u64, dtype=np.uint64)
    out = np.empty_like(k)
    lib().or_dirty_hash_u64(_p(k), C.c_int64(k.size), _p(out))
    return out


def hash_to_partition(h: np.ndarray, n_partitions: int) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.uint64)
    out = np.empty_like(h)
    lib().or_hash_to_partition(_p(h), C.c_int64(h.size), C.c_uint64(n_partitions), _p(out))
    return out


def key_bits(keys: np.ndarray) -> np.ndarray:
    """Unsigned bit representation used for hashing/equality of group/join keys."""
    keys = np.ascontiguousarray(keys)
    if keys.dtype == np.float64:
        out = np.empty(keys.size, np.uint64)
        lib().or_canonical_f64_bits(_p(keys), C.c_int64(keys.size), _p(out))
        return out
    if keys.dtype == np.float32:
        out = np.empty(keys.size, np.uint64)
        lib().or_canonical_f32_bits(_p(keys), C.c_int64(keys.size), _p(out))
        return out
    if keys.dtype.kind == "i":
        return keys.view(np.dtype(f"u{keys.dtype.itemsize}")).astype(np.uint64)
    if keys.dtype.kind in "ub":
        return keys.astype(np.uint64)
    raise TypeError(keys.dtype)


# ------------------------------------------------------------------ elementwise
def arith(op: str, lhs, rhs, lhs_valid=None, rhs_valid=None):
    """Returns (values, valid|None).  One side may be a Python/numpy scalar."""
    l_scalar, r_scalar = np.isscalar(lhs), np.isscalar(rhs)
    arr = rhs if l_scalar else lhs
    arr = np.ascontiguousarray(arr)
    dt, n = arr.dtype, arr.size
    sfx = _SUFFIX[dt]
    mode = 2 if l_scalar else (1 if r_scalar else 0)
    L = np.array([lhs], dtype=dt) if l_scalar else arr
    R = np.array([rhs], dtype=dt) if r_scalar else np.ascontiguousarray(rhs, dtype=dt)
    lv, rv = _valid(lhs_valid, n), _valid(rhs_valid, n)
    ov = np.ones(n, np.bool_)
    out = np.empty(n, dt)
    fn = getattr(lib(), f"or_arith_{sfx}")
    if dt.kind in "iu":
        out_f = np.empty(n, np.float64)
        fn(C.c_int(OPS[op]), C.c_int(mode), _p(L), _p(lv), _p(R), _p(rv), C.c_int64(n), _p(out), _p(out_f), _p(ov))
        if op == "truediv":
            out = out_f
    else:
        fn(C.c_int(OPS[op]), C.c_int(mode), _p(L), _p(lv), _p(R), _p(rv), C.c_int64(n), _p(out), _p(ov))
    return out, (None if ov.all() else ov)


def compare(op: str, lhs, rhs, lhs_valid=None, rhs_valid=None, missing: bool = False):
    """Returns (bool values, valid|None).  rhs may be a scalar."""
    lhs = np.ascontiguousarray(lhs)
    dt, n = lhs.dtype, lhs.size
    r_scalar = np.isscalar(rhs)
    R = np.array([rhs], dtype=dt) if r_scalar else np.ascontiguousarray(rhs, dtype=dt)
    lv, rv = _valid(lhs_valid, n), _valid(rhs_valid, n)
    out = np.empty(n, np.bool_)
    ov = np.ones(n, np.bool_)
    getattr(lib(), f"or_cmp_{_SUFFIX[dt]}")(C.c_int(CMPS[op]), C.c_int(1 if r_scalar else 0), C.c_int(int(missing)),
                                           _p(lhs), _p(lv), _p(R), _p(rv), C.c_int64(n), _p(out), _p(ov))
    return out, (None if ov.all() else ov)


def filter(values, valid, mask, mask_valid=None):
    values = np.ascontiguousarray(values)
    n = values.size
    mask = np.ascontiguousarray(mask, dtype=np.bool_)
    out = np.empty(n, values.dtype)
    ov = np.ones(n, np.bool_)
    k = lib().or_filter(_p(values), _p(_valid(valid, n)), C.c_int64(n), C.c_int(values.dtype.itemsize),
                        _p(mask), _p(_valid(mask_valid, n)), _p(out), _p(ov))
    return out[:k].copy(), (None if valid is None else ov[:k].copy())


def gather(values, valid, idx, idx_valid=None):
    values = np.ascontiguousarray(values)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    m = idx.size
    out = np.empty(m, values.dtype)
    ov = np.ones(m, np.bool_)
    lib().or_gather(_p(values), _p(_valid(valid, values.size)), C.c_int(values.dtype.itemsize), _p(idx),
                    _p(_valid(idx_valid, m)), C.c_int64(m), _p(out), _p(ov))
    return out, (None if (valid is None and idx_valid is None) else ov)


# ------------------------------------------------------------------ group_by
class Groups:
    """GroupsIdx {first, all} in CSR form (polars-core/src/frame/group_by/position.rs:16-22)."""

    def __init__(self, first, offsets, idx):
        self.first, self.offsets, self.idx = first, offsets, idx

    def __len__(self):
        return self.first.size


def group_by(keys, key_valid=None, n_partitions: int | None = None, maintain_order: bool = True, workspace: dict | None = None) -> Groups:
    """workspace (bench only): a dict that keeps the n-sized output buffers between calls, so that repeated timed calls
    do not pay 1.6 GB of fresh page faults each — the returned Groups then alias the workspace until the next call."""
    bits = key_bits(keys)
    n = bits.size
    P = n_partitions if n_partitions is not None else max_threads()
    if workspace is not None and workspace.get("n") == n:
        first, offsets, idx = workspace["first"], workspace["offsets"], workspace["idx"]
    else:
        first = np.empty(max(n, 1), np.uint32)
        offsets = np.empty(n + 1, np.uint64)
        idx = np.empty(max(n, 1), np.uint32)
        if workspace is not None:
            workspace.update(n=n, first=first, offsets=offsets, idx=idx)
    G = lib().or_group_by(_p(bits), _p(_valid(key_valid, n)), C.c_int64(n), C.c_int(P), C.c_int(int(maintain_order)),
                          _p(first), _p(offsets), _p(idx))
    if workspace is not None:
        return Groups(first[:G], offsets[:G + 1], idx[:n])
    return Groups(first[:G].copy(), offsets[:G + 1].copy(), idx[:n])


def group_by_multi(keys_list, valids_list=None, maintain_order: bool = True) -> Groups:
    """Several key columns.  The reference row-encodes the columns (polars-core/src/frame/group_by/mod.rs:88-94,
    polars-row/src/fixed/numeric.rs:100-145: canonical float bits, a validity sentinel per column) and groups on
    the encoded rows; two rows fall in one group iff every column agrees under that encoding.  Restated with
    numpy: per-column (validity, canonical key bits) pairs -> np.unique over the stacked rows; groups ordered by
    first occurrence (hashing.rs:41-63), row lists ascending."""
    n = np.ascontiguousarray(keys_list[0]).size
    valids_list = valids_list or [None] * len(keys_list)
    cols = []
    for k, v in zip(keys_list, valids_list):
        bits = key_bits(k)
        vv = (np.ones(n, np.uint64) if v is None else _valid(v, n).astype(np.uint64))
        cols += [vv, np.where(vv != 0, bits, np.uint64(0))]
    if n == 0:
        return Groups(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32))
    rows = np.stack(cols, axis=1)
    _, first, inverse, counts = np.unique(rows, axis=0, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")                   # groups by first occurrence
    rank = np.empty_like(order); rank[order] = np.arange(order.size)
    gid = rank[np.asarray(inverse).reshape(-1)]
    idx = np.argsort(gid, kind="stable").astype(np.uint32)    # rows grouped, ascending inside a group
    offsets = np.concatenate([[0], np.cumsum(counts[order])]).astype(np.uint64)
    return Groups(first[order].astype(np.uint32), offsets, idx)


def group_by_agg_multi(keys_list, valids_list, aggs, maintain_order=True):
    """-> ([(key values, key valid|None) per key column], [(vals, valid)...], Groups); key output = take(first)."""
    g = group_by_multi(keys_list, valids_list, maintain_order)
    valids_list = valids_list or [None] * len(keys_list)
    kouts = []
    for k, v in zip(keys_list, valids_list):
        k = np.ascontiguousarray(k)
        kouts.append((k[g.first], None if v is None else np.asarray(v, np.bool_)[g.first]))
    outs = [agg(kind, vals, valid, g) for (kind, vals, valid) in aggs]
    return kouts, outs, g


def agg(kind: str, values, valid, groups: Groups):
    """kind in sum/mean/min/max/count/len.  Returns (values, valid|None) per group."""
    G = len(groups)
    off, idx = groups.offsets, groups.idx
    L = lib()
    if kind == "len":
        out = np.empty(G, np.uint32)
        L.or_agg_len(_p(off), C.c_int64(G), _p(out))
        return out, None
    values = np.ascontiguousarray(values)
    v = _valid(valid, values.size)
    if kind == "count":
        out = np.empty(G, np.uint32)
        L.or_agg_count(_p(v), _p(off), _p(idx), C.c_int64(G), _p(out))
        return out, None
    if values.dtype in (np.dtype("int8"), np.dtype("int16"), np.dtype("uint8"), np.dtype("uint16")):
        # 8/16-bit integers aggregate after a cast to Int64 (polars-core/src/series/implementations/mod.rs:145-154);
        # sum stays Int64, mean is Float64 (aggregations/mod.rs:1227-1296), min/max return the input dtype
        out, ov = agg(kind, values.astype(np.int64), valid, groups)
        return (out.astype(values.dtype) if kind in ("min", "max", "first", "last") else out), ov
    sfx = _SUFFIX[values.dtype]
    if kind == "sum":
        out = np.empty(G, values.dtype)
        getattr(L, f"or_agg_sum_{sfx}")(_p(values), _p(v), _p(off), _p(idx), C.c_int64(G), _p(out))
        return out, None
    ov = np.empty(G, np.bool_)
    if kind == "mean":
        out = np.empty(G, np.float64)
        getattr(L, f"or_agg_mean_{sfx}")(_p(values), _p(v), _p(off), _p(idx), C.c_int64(G), _p(out), _p(ov))
        if values.dtype == np.float32:
            out = out.astype(np.float32)
        return out, (None if ov.all() else ov)
    if kind in ("first", "last"):
        # agg_first / agg_last (aggregations/agg_list.rs / dispatch.rs:57-120): the value at the group's first / last row, nulls included
        rows = groups.first if kind == "first" else idx[(off[1:] - np.uint64(1)).astype(np.int64)] if G else np.zeros(0, np.uint32)
        out = values[rows]
        ov = None if v is None else v[rows]
        return out, (None if ov is None or ov.all() else ov)
    if kind in ("var", "std") or kind.startswith(("var:", "std:")):
        name, _, dd = kind.partition(":")
        ddof = int(dd) if dd else 1
        out = np.empty(G, np.float64)
        getattr(L, f"or_agg_var_{sfx}")(_p(values), _p(v), _p(off), _p(idx), C.c_int64(G), C.c_int(ddof), C.c_int(int(name == "std")), _p(out), _p(ov))
        if values.dtype == np.float32:
            out = out.astype(np.float32)
        return out, (None if ov.all() else ov)
    if kind in ("min", "max"):
        out = np.empty(G, values.dtype)
        getattr(L, f"or_agg_{kind}_{sfx}")(_p(values), _p(v), _p(off), _p(idx), C.c_int64(G), _p(out), _p(ov))
        return out, (None if ov.all() else ov)
    raise ValueError(kind)


def group_by_agg(keys, key_valid, aggs, n_partitions=None, maintain_order=True, workspace=None):
    """aggs: list of (kind, values, valid).  Returns (key values, key valid, [(vals, valid)...], Groups).
    Key output = take(first) (polars-core/src/frame/group_by/mod.rs:258-266)."""
    g = group_by(keys, key_valid, n_partitions, maintain_order, workspace)
    keys = np.ascontiguousarray(keys)
    kout = keys[g.first] if len(g) else keys[:0]
    kv = None if key_valid is None else np.asarray(key_valid, np.bool_)[g.first]
    outs = [agg(kind, vals, valid, g) for (kind, vals, valid) in aggs]
    return kout, kv, outs, g


# ------------------------------------------------------------------ join
def _adopt_u32(ptr: C.c_void_p, m: int) -> np.ndarray:
    """numpy view of a malloc'd uint32 buffer returned by the C side; freed (or_free) when the array is collected."""
    import weakref
    if m == 0:
        lib().or_free(ptr)
        return np.zeros(0, np.uint32)
    buf = (C.c_uint32 * m).from_address(ptr.value)
    weakref.finalize(buf, lib().or_free, C.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=np.uint32, count=m)


def hash_join(left_keys, right_keys, left_valid=None, right_valid=None, how: str = "inner", nulls_equal: bool = False,
              maintain_order: str = "none", n_threads: int | None = None):
    """Returns (left_idx u32, right_idx u32); unmatched right idx (left join) = IDX_NULL.
    how = "semi" / "anti": (left_idx, empty)."""
    lk, rk = key_bits(left_keys), key_bits(right_keys)
    if how in ("semi", "anti"):
        # polars-ops/src/frame/join/hash_join/single_keys_semi_anti.rs:8-140: a hash SET of the right keys (null keys
        # only when nulls_equal, :27-29); every left row, in row order, is kept when the set holds (semi) / does not
        # hold (anti) its key.  Restated as set membership on the canonical key bits.
        lv = np.ones(lk.size, np.bool_) if left_valid is None else np.asarray(left_valid, np.bool_)
        rv = np.ones(rk.size, np.bool_) if right_valid is None else np.asarray(right_valid, np.bool_)
        match = lv & np.isin(lk, rk[rv])
        if nulls_equal and (~rv).any():
            match |= ~lv
        idx = np.nonzero(match if how == "semi" else ~match)[0].astype(np.uint32)
        return idx, np.zeros(0, np.uint32)
    T = n_threads if n_threads is not None else max_threads()
    if how == "full":
        # polars-ops/src/frame/join/hash_join/single_keys_outer.rs:100-260 + single_keys_dispatch.rs:653-694: the longer
        # relation probes (det_hash_prone_order!, tie -> right probes); the probe phase emits exactly the left-join tuples
        # of the probe side (a null probe key is unmatched unless nulls_equal, :144-150), then the build rows no probe key
        # reached are drained from the hash tables as (None, idx_b) — in hashbrown iteration order (UNPINNED); this
        # restatement drains them in ascending build-row order.
        swapped = not (lk.size > rk.size)
        pk, pv, bk, bv = (rk, right_valid, lk, left_valid) if swapped else (lk, left_valid, rk, right_valid)
        pi, bi = hash_join(pk, bk, pv, bv, "left", nulls_equal, "none", T)
        matched = np.zeros(bk.size, np.bool_)
        matched[bi[bi != IDX_NULL]] = True
        drained = np.nonzero(~matched)[0].astype(np.uint32)
        pi = np.concatenate([pi, np.full(drained.size, IDX_NULL, np.uint32)])
        bi = np.concatenate([bi, drained])
        return (bi, pi) if swapped else (pi, bi)
    pl, pr = C.c_void_p(), C.c_void_p()
    m = lib().or_hash_join(_p(lk), _p(_valid(left_valid, lk.size)), C.c_int64(lk.size), _p(rk),
                           _p(_valid(right_valid, rk.size)), C.c_int64(rk.size), C.c_int({"inner": 0, "left": 1}[how]),
                           C.c_int(int(nulls_equal)), C.c_int(T), C.byref(pl), C.byref(pr))
    li, ri = _adopt_u32(pl, m), _adopt_u32(pr, m)
    if how == "inner" and maintain_order != "none":
        # polars-ops/src/frame/join/mod.rs:577-642: left probes (sorted) iff len(left) > len(right)
        left_sorted = lk.size > rk.size
        if maintain_order in ("left", "left_right"):
            if not left_sorted:
                lib().or_stable_sort_pairs(_p(li), _p(ri), C.c_int64(m), C.c_int(0))
        elif maintain_order in ("right", "right_left"):
            lib().or_stable_sort_pairs(_p(li), _p(ri), C.c_int64(m), C.c_int(1))
        else:
            raise ValueError(maintain_order)
    if how == "left" and maintain_order in ("right", "right_left"):
        # polars-ops/src/frame/join/dispatch_left_right.rs:142-170: stable sort on the right idx (null = u32::MAX last)
        lib().or_stable_sort_pairs(_p(li), _p(ri), C.c_int64(m), C.c_int(1))
    return li, ri


def hash_join_multi(left_keys, right_keys, left_valids=None, right_valids=None, how: str = "inner", nulls_equal: bool = False,
                    maintain_order: str = "none", n_threads: int | None = None):
    """Several key columns per side.  The reference canonicalises float keys, row-encodes the columns of each side
    into one binary key (polars-ops/src/frame/join/mod.rs:658-678 `prepare_keys_multiple`) and runs the SAME
    single-key machinery on it; with nulls_equal = false a null in ANY key column makes the whole encoded key null
    (`encode_rows_vertical_par_unordered_broadcast_nulls`, :676), with nulls_equal = true nulls are part of the key.
    Restated: one dense id per distinct (validity, canonical bits) row over BOTH relations, then `hash_join`."""
    nl, nr = np.ascontiguousarray(left_keys[0]).size, np.ascontiguousarray(right_keys[0]).size
    left_valids = left_valids or [None] * len(left_keys)
    right_valids = right_valids or [None] * len(right_keys)
    cols, all_valid = [], np.ones(nl + nr, np.bool_)
    for lk, rk, lv, rv in zip(left_keys, right_keys, left_valids, right_valids):
        if np.asarray(lk).dtype != np.asarray(rk).dtype:
            raise TypeError("join key dtypes differ")                      # join/mod.rs:231-241
        bits = np.concatenate([key_bits(lk), key_bits(rk)])
        vv = np.concatenate([np.ones(nl, np.bool_) if lv is None else _valid(lv, nl), np.ones(nr, np.bool_) if rv is None else _valid(rv, nr)])
        all_valid &= vv
        cols += [vv.astype(np.uint64), np.where(vv, bits, np.uint64(0))]
    if nl + nr == 0:
        ids = np.zeros(0, np.uint64)
    else:
        _, inverse = np.unique(np.stack(cols, axis=1), axis=0, return_inverse=True)
        ids = np.asarray(inverse).reshape(-1).astype(np.uint64)
    row_valid = None if nulls_equal else all_valid
    return hash_join(ids[:nl], ids[nl:], None if row_valid is None else row_valid[:nl], None if row_valid is None else row_valid[nl:],
                     how, nulls_equal, maintain_order, n_threads)


def string_codes(values):
    """BinaryChunked::group_tuples (polars-core/src/frame/group_by/into_groups.rs:215-251): rows are grouped by their BYTES
    (to_bytes_hashes + equality; a null is its own group) and a group's `first` is the row of its first occurrence
    (hashing.rs:26-63).  Returns (codes, valid, n_distinct): codes[i] = first row with the same bytes (IdxSize u32), valid[i] =
    row i is non-null (None when there are no nulls), n_distinct = distinct non-null values.  Plain Python dict: test
    infrastructure for small inputs, never the product path."""
    first = {}
    codes = np.zeros(len(values), np.uint32)
    valid = np.ones(len(values), bool)
    for i, v in enumerate(values):
        if v is None:
            valid[i] = False
            continue
        b = v.encode() if isinstance(v, str) else bytes(v)
        codes[i] = first.setdefault(b, i)
    return codes, (None if valid.all() else valid), len(first)


def group_n_unique(key, key_valid, values, valid, maintain_order=True):
    """agg_n_unique (polars-core/src/frame/group_by/aggregations/dispatch.rs:285-345): per group, the number of distinct values
    among the group's rows; a null is one value; floats compare by total equality (NaN == NaN, -0.0 == 0.0,
    polars-utils/src/total_ord.rs:37-47).  Groups in first-occurrence order (null key = own group), as group_by_agg with
    maintain_order.  Returns (first_row_of_group, counts u32).  Plain Python: test infrastructure for small inputs."""
    key = np.asarray(key); values = np.asarray(values)
    groups, firsts = {}, []
    for i in range(key.size):
        k = None if (key_valid is not None and not key_valid[i]) else (key[i].item() if key.dtype.kind != "f" else _canon_float(key[i]))
        if valid is not None and not valid[i]:
            v = None
        else:
            v = values[i].item() if values.dtype.kind != "f" else _canon_float(values[i])
        if k not in groups:
            groups[k] = set(); firsts.append(i)
        groups[k].add(v)
    return np.array(firsts, np.uint32), np.array([len(s) for s in groups.values()], np.uint32)


def _canon_float(x):
    x = float(x)
    if x != x:
        return "nan"
    return 0.0 if x == 0.0 else x
