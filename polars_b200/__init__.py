"""polars_b200 — host-side Python binding of libpolars_b200.so (the C ABI in include/polars_b200.h).

This module is plumbing: ctypes structs, numpy <-> bl_column marshalling and error translation.
All compute happens in the CUDA library; if the library (or a GPU) is missing every call raises —
there is no CPU fallback and this package never imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_lib", "libpolars_b200.so")

HOST, DEVICE = 0, 1
IDX_NULL = 0xFFFFFFFF

DTYPES = {np.dtype("int8"): 0, np.dtype("int16"): 1, np.dtype("int32"): 2, np.dtype("int64"): 3,
          np.dtype("uint8"): 4, np.dtype("uint16"): 5, np.dtype("uint32"): 6, np.dtype("uint64"): 7,
          np.dtype("float32"): 8, np.dtype("float64"): 9, np.dtype("bool"): 10}
NP_OF = {v: k for k, v in DTYPES.items()}
BOOL = 10
OPS = {"add": 0, "sub": 1, "mul": 2, "floordiv": 3, "mod": 4, "truediv": 5}
CMPS = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}
AGGS = {"sum": 0, "mean": 1, "min": 2, "max": 3, "count": 4, "len": 5, "first": 6, "last": 7, "var": 8 | (1 << 16), "std": 9 | (1 << 16), "n_unique": 10}


def _agg_kind(kind: str) -> int:
    """"var" / "std" default to ddof = 1 (Polars); "var:0", "std:2" ... carry an explicit ddof (BL_AGG_WITH_DDOF)."""
    name, _, dd = kind.partition(":")
    return (AGGS[name] & 0xFFFF) | (int(dd) << 16) if dd else AGGS[name]
JOINS = {"inner": 0, "left": 1, "semi": 2, "anti": 3, "full": 4}
ORDERS = {"none": 0, "left": 1, "left_right": 2, "right": 3, "right_left": 4}
STATUS = {1: "INVALID", 2: "CUDA", 3: "OOM", 4: "UNSUPPORTED", 5: "DTYPE", 6: "BOUNDS"}


class BlColumn(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("location", C.c_int32), ("length", C.c_int64), ("offset", C.c_int64),
                ("null_count", C.c_int64), ("values", C.c_void_p), ("validity", C.c_void_p), ("owner", C.c_void_p)]


class BlStringColumn(C.Structure):
    _fields_ = [("location", C.c_int32), ("reserved", C.c_int32), ("length", C.c_int64), ("offset", C.c_int64), ("null_count", C.c_int64),
                ("offsets", C.c_void_p), ("data", C.c_void_p), ("validity", C.c_void_p), ("owner", C.c_void_p)]


class BlAgg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_chunks", C.c_int32), ("values", C.POINTER(BlColumn))]


class B200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[{STATUS.get(status, status)}] {msg}")
        self.status = status


class ComputeError(B200Error):
    """dtype mismatches etc. — the reference raises ComputeError (join/mod.rs:231-241)."""


class OutOfBoundsError(B200Error):
    pass


_lib = None


def lib() -> C.CDLL:
    """Loads libpolars_b200.so.  Fails loudly when it has not been built (python -m polars_b200.build)."""
    global _lib, _SO
    if _lib is None:
        _SO = os.environ.get("POLARS_B200_LIB", _SO)      # a freshly built copy (see __graft_entry__.smoke)
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: build it with `python -m polars_b200.build` (nvcc, sm_100a). "
                              "polars_b200 has no CPU fallback.")
        L = C.CDLL(_SO)
        L.bl_last_error.restype = C.c_char_p
        L.bl_profile_json.restype = C.c_int64
        L.bl_launch_count.restype = C.c_int64
        L.bl_stream.restype = C.c_void_p
        for name in ("bl_column_free", "bl_free_pinned", "bl_dev_free", "bl_profile_enable", "bl_profile_reset", "bl_shutdown",
                     "bl_groupby_reset", "bl_groupby_destroy"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _check(st: int):
    if st != 0:
        msg = lib().bl_last_error().decode("utf-8", "replace")
        if st == 5:
            raise ComputeError(st, msg)
        if st == 6:
            raise OutOfBoundsError(st, msg)
        raise B200Error(st, msg)


def init(device: int = -1):
    _check(lib().bl_init(C.c_int32(device)))


def device_info() -> dict:
    sm, l2, tot, free = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib().bl_device_info(C.byref(sm), C.byref(l2), C.byref(tot), C.byref(free)))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "hbm_total": tot.value, "hbm_free": free.value}


def set_deterministic(on: bool = True):
    """Bit-stable group_by aggregation in the reference's own order (bl_set_deterministic)."""
    lib().bl_set_deterministic.restype = None
    lib().bl_set_deterministic(C.c_int32(int(on)))


def use_library(path: str) -> None:
    """Bind this module to another build of libpolars_b200.so (a fresh build made by __graft_entry__.smoke()).  A copy that
    is already loaded keeps running for whoever still holds its objects; everything created afterwards uses `path`."""
    global _lib, _SO
    _SO, _lib = path, None
    os.environ["POLARS_B200_LIB"] = path
    lib()


def loaded_library() -> str:
    """Path of the shared library this process has loaded (after the first call into it)."""
    lib()
    return _SO


def sync():
    _check(lib().bl_sync())


def stream() -> int:
    return int(lib().bl_stream() or 0)


def profile_enable(on: bool):
    lib().bl_profile_enable(C.c_int32(int(on)))


def profile_reset():
    lib().bl_profile_reset()


def profile() -> dict:
    n = lib().bl_profile_json(None, C.c_int64(0))
    buf = C.create_string_buffer(int(n) + 16)
    lib().bl_profile_json(buf, C.c_int64(len(buf)))
    return json.loads(buf.value.decode() or "{}")


def launch_count() -> int:
    return int(lib().bl_launch_count())


# ---------------------------------------------------------------------------------- pinned memory
def pinned_empty(n: int, dtype) -> np.ndarray:
    """numpy array backed by library pinned host memory (DMA-able without staging).
    The buffer is returned to the library's pinned cache when the array is garbage collected."""
    dt = np.dtype(dtype)
    nbytes = max(int(n) * dt.itemsize, 1)
    p = C.c_void_p()
    _check(lib().bl_alloc_pinned(C.c_size_t(nbytes), C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    import weakref
    weakref.finalize(buf, lib().bl_free_pinned, C.c_void_p(p.value))
    return arr


def to_pinned(a: np.ndarray) -> np.ndarray:
    out = pinned_empty(a.size, a.dtype)
    out[:] = a
    return out


# ---------------------------------------------------------------------------------- columns
def pack_bits(valid: np.ndarray) -> np.ndarray:
    return np.packbits(np.asarray(valid, dtype=np.bool_), bitorder="little")


def unpack_bits(buf: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(buf, count=n, bitorder="little").astype(np.bool_)


class Column:
    """A caller-owned host (numpy) or device column view passed INTO the library.

    values: numpy array (host) or an int device pointer; valid: bool array / packed bitmap (host) or
    device pointer to a bitmap.  `offset` exercises Arrow slicing semantics."""

    def __init__(self, values, valid=None, *, dtype=None, length=None, offset: int = 0, location: int = HOST, null_count: int = -1):
        self.location = location
        self.offset = int(offset)
        self._keep = []
        if location == HOST:
            values = np.asarray(values)
            if values.dtype == np.bool_:
                self.length = int(values.size - offset) if length is None else int(length)
                bits = pack_bits(values)
                self._keep.append(bits)
                self.dtype = BOOL
                self._vptr = bits.ctypes.data
            else:
                values = np.ascontiguousarray(values)
                self._keep.append(values)
                self.dtype = DTYPES[values.dtype]
                self.length = int(values.size - offset) if length is None else int(length)
                self._vptr = values.ctypes.data
            if valid is None:
                self._mptr, self.null_count = None, 0
            else:
                valid = np.asarray(valid)
                bits = pack_bits(valid) if valid.dtype == np.bool_ else np.ascontiguousarray(valid, dtype=np.uint8)
                self._keep.append(bits)
                self._mptr, self.null_count = bits.ctypes.data, null_count
        else:
            self.dtype = DTYPES[np.dtype(dtype)] if not isinstance(dtype, int) else dtype
            self.length = int(length)
            self._vptr = int(values)
            self._mptr = None if valid is None else int(valid)
            self.null_count = 0 if valid is None else null_count

    def struct(self) -> BlColumn:
        return BlColumn(self.dtype, self.location, self.length, self.offset, self.null_count, self._vptr, self._mptr, None)


class OutColumn:
    """A library-owned output column.  `.to_numpy()` copies host outputs out; device outputs expose
    raw pointers (`.values_ptr`, `.validity_ptr`).  Freed on `.free()` / garbage collection."""

    def __init__(self, st: BlColumn):
        self.st = st

    @property
    def length(self):
        return int(self.st.length)

    @property
    def location(self):
        return int(self.st.location)

    @property
    def values_ptr(self):
        return int(self.st.values or 0)

    @property
    def validity_ptr(self):
        return int(self.st.validity or 0)

    def view(self) -> Column:
        """Re-use a device output as an input column (no copy)."""
        assert self.location == DEVICE
        return Column(self.values_ptr, self.validity_ptr or None, dtype=int(self.st.dtype), length=self.length, location=DEVICE, null_count=int(self.st.null_count))

    def to_numpy(self):
        """-> (values, valid|None).  BOOL columns come back as numpy bool arrays."""
        st, n = self.st, int(self.st.length)
        if st.location == DEVICE:
            host = BlColumn()
            _check(lib().bl_column_to(C.byref(st), C.c_int32(1), C.c_int32(HOST), C.byref(host)))
            return OutColumn(host).to_numpy()      # the returned view keeps the host copy alive
        if st.dtype == BOOL:
            raw = np.ctypeslib.as_array(C.cast(st.values, C.POINTER(C.c_uint8)), shape=((n + 7) // 8 or 1,)) if n else np.zeros(0, np.uint8)
            vals = unpack_bits(raw, n)
        else:
            dt = NP_OF[int(st.dtype)]
            if n:
                # zero-copy: the array views the library-owned pinned buffer and keeps this OutColumn alive
                buf = (C.c_char * (n * dt.itemsize)).from_address(st.values)
                buf._owner = self
                vals = np.frombuffer(buf, dtype=dt, count=n)
                self._exported = True
            else:
                vals = np.zeros(0, dt)
        valid = None
        if st.validity:
            raw = np.ctypeslib.as_array(C.cast(st.validity, C.POINTER(C.c_uint8)), shape=((n + 7) // 8 or 1,)) if n else np.zeros(0, np.uint8)
            valid = unpack_bits(raw, n)
            if valid.all():
                valid = None
        return vals, valid

    def free(self):
        if self.st.owner:
            lib().bl_column_free(C.byref(self.st))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _as_col(x) -> Column:
    if isinstance(x, Column):
        return x
    if isinstance(x, OutColumn):
        return x.view()
    if isinstance(x, tuple):
        return Column(x[0], x[1])
    if np.isscalar(x):
        raise TypeError("scalars must be passed as length-1 arrays with the column's dtype")
    return Column(x)


def _scalar_col(x, like: Column) -> Column:
    if isinstance(x, (Column, OutColumn, tuple, np.ndarray)):
        return _as_col(x)
    return Column(np.array([x], dtype=NP_OF[like.dtype]))


def _finish(outs, location):
    res = [OutColumn(o) for o in outs]
    if location == HOST:
        np_res = [r.to_numpy() for r in res]
        for r in res:
            if not getattr(r, "_exported", False):
                r.free()       # value arrays that view the buffer keep their OutColumn alive instead
        return np_res
    return res


# ---------------------------------------------------------------------------------- string keys
class StringColumn:
    """A caller-owned host string / binary column in Arrow LargeUtf8 layout, built from a sequence of str / bytes / None
    (or from ready-made `offsets` (int64, n + 1), `data` (uint8) and an optional bool `valid`).  `offset` / `length`
    exercise Arrow slicing."""

    def __init__(self, values=None, *, offsets=None, data=None, valid=None, offset: int = 0, length: int | None = None):
        if values is not None:
            enc = [None if v is None else (v.encode() if isinstance(v, str) else bytes(v)) for v in values]
            lens = np.fromiter((0 if b is None else len(b) for b in enc), dtype=np.int64, count=len(enc))
            offsets = np.zeros(len(enc) + 1, np.int64)
            np.cumsum(lens, out=offsets[1:])
            data = np.frombuffer(b"".join(b for b in enc if b is not None), dtype=np.uint8)
            valid = None if all(b is not None for b in enc) else np.array([b is not None for b in enc], bool)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.data = np.ascontiguousarray(data, dtype=np.uint8) if data is not None and len(data) else np.zeros(1, np.uint8)
        self.bits = None if valid is None else pack_bits(np.asarray(valid, bool))
        self.offset = int(offset)
        self.length = int(self.offsets.size - 1 - offset) if length is None else int(length)

    def struct(self) -> BlStringColumn:
        return BlStringColumn(HOST, 0, self.length, self.offset, 0 if self.bits is None else -1, self.offsets.ctypes.data, self.data.ctypes.data,
                              None if self.bits is None else self.bits.ctypes.data, None)


def _str_array(cols: Sequence[StringColumn]):
    return (BlStringColumn * len(cols))(*[c.struct() for c in cols])


def string_encode(col, location: int = HOST):
    """bl_string_encode: -> ((codes, valid) | OutColumn, n_distinct).  codes[i] = first row holding the same bytes as row i."""
    chunks = col if isinstance(col, list) else [col]
    arr = _str_array(chunks)
    out, nd = BlColumn(), C.c_int64()
    _check(lib().bl_string_encode(arr, C.c_int32(len(chunks)), C.c_int32(location), C.byref(out), C.byref(nd)))
    return _finish([out], location)[0], int(nd.value)


def _string_out_to_list(out: BlStringColumn) -> list:
    """host BlStringColumn (library-owned) -> list of bytes / None; frees it."""
    try:
        n = int(out.length)
        offs = np.ctypeslib.as_array(C.cast(out.offsets, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        total = int(offs[-1])
        data = bytes(np.ctypeslib.as_array(C.cast(out.data, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total])
        valid = None
        if out.validity:
            valid = unpack_bits(np.ctypeslib.as_array(C.cast(out.validity, C.POINTER(C.c_uint8)), shape=((n + 7) // 8 or 1,)), n)
        return [None if (valid is not None and not valid[i]) else data[offs[i]:offs[i + 1]] for i in range(n)]
    finally:
        lib().bl_string_column_free.restype = None
        lib().bl_string_column_free(C.byref(out))


def string_gather(col, idx) -> list:
    """bl_string_gather with host output: -> list of bytes / None."""
    chunks = col if isinstance(col, list) else [col]
    arr = _str_array(chunks)
    ic = _as_col(idx)
    ist = ic.struct()
    out = BlStringColumn()
    _check(lib().bl_string_gather(arr, C.c_int32(len(chunks)), C.byref(ist), C.c_int32(HOST), C.byref(out)))
    return _string_out_to_list(out)


def group_by_agg_strings(key, aggs: Sequence, maintain_order: bool = False):
    """bl_groupby_agg_strings (host outputs): key = StringColumn or list of chunks; aggs as in group_by_agg.
    -> (group keys as a list of bytes / None, [agg outputs])."""
    chunks = key if isinstance(key, list) else [key]
    karr = _str_array(chunks)
    keep, agg_structs, cache = [], [], {}
    for kind, vals in aggs:
        if kind == "len" or vals is None:
            agg_structs.append(BlAgg(_agg_kind(kind), 0, None))
            continue
        ident = id(vals)
        if ident not in cache:
            cs = [_as_col(c) for c in (vals if isinstance(vals, list) else [vals])]
            cache[ident] = (cs, _col_array(cs))
        cs, arr = cache[ident]
        keep.append((cs, arr))
        agg_structs.append(BlAgg(_agg_kind(kind), len(cs), C.cast(arr, C.POINTER(BlColumn))))
    aarr = (BlAgg * max(len(agg_structs), 1))(*agg_structs)
    out_key, out_aggs = BlStringColumn(), (BlColumn * max(len(agg_structs), 1))()
    _check(lib().bl_groupby_agg_strings(karr, C.c_int32(len(chunks)), aarr, C.c_int32(len(agg_structs)), C.c_int32(int(maintain_order)), C.c_int32(HOST),
                                        C.byref(out_key), out_aggs))
    return _string_out_to_list(out_key), _finish(list(out_aggs)[: len(agg_structs)], HOST)


def hash_join_strings(left, right, how: str = "inner", nulls_equal: bool = False, maintain_order: str = "none", location: int = HOST):
    """bl_hash_join_strings: row-index tuples of a join on a string key (both sides encoded together on the device)."""
    lc = left if isinstance(left, list) else [left]
    rc = right if isinstance(right, list) else [right]
    la, ra = _str_array(lc), _str_array(rc)
    ol, orr = BlColumn(), BlColumn()
    _check(lib().bl_hash_join_strings(la, C.c_int32(len(lc)), ra, C.c_int32(len(rc)), C.c_int32(JOINS[how]), C.c_int32(int(nulls_equal)), C.c_int32(ORDERS[maintain_order]),
                                      C.c_int32(location), C.byref(ol), C.byref(orr)))
    return tuple(_finish([ol, orr], location))


# ---------------------------------------------------------------------------------- operators
def elementwise(op: str, lhs, rhs, location: int = HOST):
    l = _as_col(lhs) if not np.isscalar(lhs) else None
    r = _as_col(rhs) if not np.isscalar(rhs) else None
    if l is None:
        l = _scalar_col(lhs, r)
    if r is None:
        r = _scalar_col(rhs, l)
    out = BlColumn()
    ls, rs = l.struct(), r.struct()
    _check(lib().bl_elementwise(C.c_int32(OPS[op]), C.byref(ls), C.byref(rs), C.c_int32(location), C.byref(out)))
    return _finish([out], location)[0]


def compare(op: str, lhs, rhs, missing: bool = False, location: int = HOST):
    l = _as_col(lhs)
    r = _scalar_col(rhs, l)
    out = BlColumn()
    ls, rs = l.struct(), r.struct()
    _check(lib().bl_compare(C.c_int32(CMPS[op]), C.byref(ls), C.byref(rs), C.c_int32(int(missing)), C.c_int32(location), C.byref(out)))
    return _finish([out], location)[0]


def _col_array(cols: Sequence[Column]):
    arr = (BlColumn * len(cols))(*[c.struct() for c in cols])
    return arr


def filter(cols: Sequence, mask, location: int = HOST):
    cs = [_as_col(c) for c in cols]
    m = _as_col(mask)
    arr, outs = _col_array(cs), (BlColumn * len(cs))()
    ms = m.struct()
    _check(lib().bl_filter(arr, C.c_int32(len(cs)), C.byref(ms), C.c_int32(location), outs))
    return _finish(list(outs), location)


def filter_cmp(cols: Sequence, pred_col: int, op: str, scalar, location: int = HOST):
    cs = [_as_col(c) for c in cols]
    s = _scalar_col(scalar, cs[pred_col])
    arr, outs = _col_array(cs), (BlColumn * len(cs))()
    ss = s.struct()
    _check(lib().bl_filter_cmp(arr, C.c_int32(len(cs)), C.c_int32(pred_col), C.c_int32(CMPS[op]), C.byref(ss), C.c_int32(location), outs))
    return _finish(list(outs), location)


def gather(cols: Sequence, idx, check_bounds: bool = True, location: int = HOST):
    cs = [_as_col(c) for c in cols]
    ix = _as_col(idx)
    arr, outs = _col_array(cs), (BlColumn * len(cs))()
    ixs = ix.struct()
    _check(lib().bl_gather(arr, C.c_int32(len(cs)), C.byref(ixs), C.c_int32(int(check_bounds)), C.c_int32(location), outs))
    return _finish(list(outs), location)


def group_by_agg(key, aggs: Sequence, maintain_order: bool = False, location: int = HOST):
    """key: column or list of chunks; aggs: [(kind, column | [chunks] | None)].
    Returns (key_out, [agg_outs])."""
    kchunks = [_as_col(c) for c in (key if isinstance(key, list) else [key])]
    karr = _col_array(kchunks)
    keep, agg_structs = [], []
    cache = {}
    for kind, vals in aggs:
        if kind == "len" or vals is None:
            agg_structs.append(BlAgg(_agg_kind(kind), 0, None))
            continue
        ident = id(vals)
        if ident not in cache:
            chunks = [_as_col(c) for c in (vals if isinstance(vals, list) else [vals])]
            cache[ident] = (chunks, _col_array(chunks))
        chunks, arr = cache[ident]
        keep.append((chunks, arr))
        agg_structs.append(BlAgg(_agg_kind(kind), len(chunks), C.cast(arr, C.POINTER(BlColumn))))
    aarr = (BlAgg * max(len(agg_structs), 1))(*agg_structs)
    out_key, out_aggs = BlColumn(), (BlColumn * max(len(agg_structs), 1))()
    _check(lib().bl_groupby_agg(karr, C.c_int32(len(kchunks)), aarr, C.c_int32(len(agg_structs)), C.c_int32(int(maintain_order)), C.c_int32(location),
                                C.byref(out_key), out_aggs))
    res = _finish([out_key] + list(out_aggs)[: len(agg_structs)], location)
    return res[0], res[1:]


def group_by_agg_partitioned(key, aggs: Sequence, my_rank: int, peer_halves: Sequence[int], own_half: int, rows_per_src: int, epoch: int,
                             expected_groups: int = 0, location: int = HOST):
    """One rank's step of the multi-GPU group_by in one C call (bl_groupby_agg_partitioned): local pre-aggregation, fused
    partition + P2P exchange, merge of the rows this rank owns, finish.  aggs as in group_by_agg (one chunk per column)."""
    k = _as_col(key)
    ks = k.struct()
    keep, agg_structs, cache = [], [], {}
    for kind, vals in aggs:
        if kind == "len" or vals is None:
            agg_structs.append(BlAgg(_agg_kind(kind), 0, None))
            continue
        ident = id(vals)
        if ident not in cache:
            chunks = [_as_col(vals)]
            cache[ident] = (chunks, _col_array(chunks))
        chunks, arr = cache[ident]
        keep.append((chunks, arr))
        agg_structs.append(BlAgg(_agg_kind(kind), 1, C.cast(arr, C.POINTER(BlColumn))))
    aarr = (BlAgg * max(len(agg_structs), 1))(*agg_structs)
    n = len(peer_halves)
    parr = (C.c_void_p * n)(*[C.c_void_p(w) for w in peer_halves])
    out_key, out_aggs = BlColumn(), (BlColumn * max(len(agg_structs), 1))()
    _check(lib().bl_groupby_agg_partitioned(C.byref(ks), aarr, C.c_int32(len(agg_structs)), C.c_int32(n), C.c_int32(my_rank), parr, C.c_void_p(own_half),
                                            C.c_int64(rows_per_src), C.c_uint64(epoch), C.c_int64(expected_groups), C.c_int32(location), C.byref(out_key), out_aggs))
    res = _finish([out_key] + list(out_aggs)[: len(agg_structs)], location)
    return res[0], res[1:]


def group_by_agg_keys(keys: Sequence, aggs: Sequence, maintain_order: bool = False, location: int = HOST):
    """Several key columns (one chunk each); aggs as in group_by_agg.  Returns ([key_outs], [agg_outs])."""
    kcols = [_as_col(c) for c in keys]
    karr = _col_array(kcols)
    keep, agg_structs = [], []
    cache = {}
    for kind, vals in aggs:
        if kind == "len" or vals is None:
            agg_structs.append(BlAgg(_agg_kind(kind), 0, None))
            continue
        ident = id(vals)
        if ident not in cache:
            chunks = [_as_col(c) for c in (vals if isinstance(vals, list) else [vals])]
            cache[ident] = (chunks, _col_array(chunks))
        chunks, arr = cache[ident]
        keep.append((chunks, arr))
        agg_structs.append(BlAgg(_agg_kind(kind), len(chunks), C.cast(arr, C.POINTER(BlColumn))))
    aarr = (BlAgg * max(len(agg_structs), 1))(*agg_structs)
    out_keys, out_aggs = (BlColumn * len(kcols))(), (BlColumn * max(len(agg_structs), 1))()
    _check(lib().bl_groupby_agg_keys(karr, C.c_int32(len(kcols)), aarr, C.c_int32(len(agg_structs)), C.c_int32(int(maintain_order)), C.c_int32(location),
                                     out_keys, out_aggs))
    res = _finish(list(out_keys) + list(out_aggs)[: len(agg_structs)], location)
    return res[: len(kcols)], res[len(kcols):]


def group_tuples(key, location: int = HOST):
    """GroupsIdx of the reference (first, offsets, all): groups in first-occurrence order, rows ascending."""
    kch = [_as_col(c) for c in (key if isinstance(key, list) else [key])]
    ka = _col_array(kch)
    of, oo, oa = BlColumn(), BlColumn(), BlColumn()
    _check(lib().bl_group_tuples(ka, C.c_int32(len(kch)), C.c_int32(location), C.byref(of), C.byref(oo), C.byref(oa)))
    res = _finish([of, oo, oa], location)
    if location == HOST:
        return res[0][0], res[1][0], res[2][0]
    return res[0], res[1], res[2]


def hash_join(left_key, right_key, how: str = "inner", nulls_equal: bool = False, maintain_order: str = "none", location: int = HOST):
    lch = [_as_col(c) for c in (left_key if isinstance(left_key, list) else [left_key])]
    rch = [_as_col(c) for c in (right_key if isinstance(right_key, list) else [right_key])]
    la, ra = _col_array(lch), _col_array(rch)
    ol, orr = BlColumn(), BlColumn()
    _check(lib().bl_hash_join(la, C.c_int32(len(lch)), ra, C.c_int32(len(rch)), C.c_int32(JOINS[how]), C.c_int32(int(nulls_equal)),
                              C.c_int32(ORDERS[maintain_order]), C.c_int32(location), C.byref(ol), C.byref(orr)))
    res = _finish([ol, orr], location)
    return res[0], res[1]


def hash_join_keys(left_keys: Sequence, right_keys: Sequence, how: str = "inner", nulls_equal: bool = False, maintain_order: str = "none", location: int = HOST):
    """Join on several key columns per side (bl_hash_join_keys)."""
    lc, rc = [_as_col(c) for c in left_keys], [_as_col(c) for c in right_keys]
    assert len(lc) == len(rc) and lc
    la, ra = _col_array(lc), _col_array(rc)
    ol, orr = BlColumn(), BlColumn()
    _check(lib().bl_hash_join_keys(la, ra, C.c_int32(len(lc)), C.c_int32(JOINS[how]), C.c_int32(int(nulls_equal)), C.c_int32(ORDERS[maintain_order]), C.c_int32(location),
                                   C.byref(ol), C.byref(orr)))
    res = _finish([ol, orr], location)
    return res[0], res[1]


def join(left_key, right_key, left_cols: Sequence, right_cols: Sequence, how: str = "inner", nulls_equal: bool = False,
         maintain_order: str = "none", location: int = HOST):
    """hash_join + gather of both sides on the device (the reference's _finish_join).  Returns
    (left outputs, right outputs), each a list like `gather` returns."""
    lk, rk = _as_col(left_key), _as_col(right_key)
    lc, rc = [_as_col(c) for c in left_cols], [_as_col(c) for c in right_cols]
    lks, rks = lk.struct(), rk.struct()
    la, ra = (_col_array(lc) if lc else None), (_col_array(rc) if rc else None)
    lo, ro = (BlColumn * max(len(lc), 1))(), (BlColumn * max(len(rc), 1))()
    _check(lib().bl_join(C.byref(lks), C.byref(rks), la, C.c_int32(len(lc)), ra, C.c_int32(len(rc)), C.c_int32(JOINS[how]), C.c_int32(int(nulls_equal)),
                         C.c_int32(ORDERS[maintain_order]), C.c_int32(location), lo, ro))
    return _finish(list(lo)[:len(lc)], location), _finish(list(ro)[:len(rc)], location)


def hash_partition(key, payload: Sequence, n_partitions: int, location: int = HOST):
    k = _as_col(key)
    ps = [_as_col(p) for p in payload]
    parr = _col_array(ps) if ps else None
    ok, op = BlColumn(), (BlColumn * max(len(ps), 1))()
    offs = (C.c_int64 * (n_partitions + 1))()
    ks = k.struct()
    _check(lib().bl_hash_partition(C.byref(ks), parr, C.c_int32(len(ps)), C.c_int32(n_partitions), C.c_int32(location), C.byref(ok), op, offs))
    res = _finish([ok] + list(op)[: len(ps)], location)
    return res[0], res[1:], np.array(list(offs), dtype=np.int64)


class GroupBy:
    """Streaming group_by state (bl_groupby_*): consume batches, exchange partial aggregates, finish."""

    def __init__(self, key_dtype, aggs: Sequence, expected_groups: int = 0, track_first: bool = False, nullable: Sequence | None = None):
        """aggs: [(kind, value_dtype | None)]; nullable[i] = False promises a null-free column (smaller entries)"""
        self.kinds = [AGGS[k] for k, _ in aggs]
        dts = [DTYPES[np.dtype(d)] if d is not None else 3 for _, d in aggs]
        n = len(aggs)
        self.n = n
        self.h = C.c_void_p()
        nl = None if nullable is None else (C.c_int32 * max(n, 1))(*[int(bool(x)) for x in nullable])
        _check(lib().bl_groupby_create(C.c_int32(DTYPES[np.dtype(key_dtype)]), (C.c_int32 * max(n, 1))(*self.kinds), (C.c_int32 * max(n, 1))(*dts), nl,
                                       C.c_int32(n), C.c_int64(expected_groups), C.c_int32(int(track_first)), C.byref(self.h)))

    def consume(self, key, values: Sequence, row_base: int = 0):
        k = _as_col(key)
        cols = [(_as_col(v) if v is not None else Column(np.zeros(0, np.int64))) for v in values]
        arr = _col_array(cols) if cols else None
        ks = k.struct()
        _check(lib().bl_groupby_consume(self.h, C.byref(ks), arr, C.c_int64(row_base)))

    def export_partials(self, n_partitions: int):
        """-> (device pointer, row_words, offsets[n_partitions+1]); free the pointer with dev_free()."""
        p, rw = C.c_void_p(), C.c_int32()
        offs = (C.c_int64 * (n_partitions + 1))()
        _check(lib().bl_groupby_export_partials(self.h, C.c_int32(n_partitions), C.byref(p), C.byref(rw), offs))
        return int(p.value or 0), int(rw.value), np.array(list(offs), dtype=np.int64)

    def export_partials_p2p(self, windows: Sequence[int], my_rank: int, rows_per_src: int):
        """Fused partition + exchange: stores this rank's partial rows into the peers' windows.
        -> (row_words, sent_rows[n_ranks])"""
        n = len(windows)
        arr = (C.c_void_p * n)(*[C.c_void_p(w) for w in windows])
        rw = C.c_int32()
        sent = (C.c_int64 * n)()
        _check(lib().bl_groupby_export_partials_p2p(self.h, C.c_int32(n), C.c_int32(my_rank), arr, C.c_int64(rows_per_src), C.byref(rw), sent))
        return int(rw.value), np.array(list(sent), dtype=np.int64)

    def export_partials_p2p_async(self, window_halves: Sequence[int], my_rank: int, rows_per_src: int, epoch: int) -> int:
        """Fused partition + exchange + count publication (no host round trip).  -> row_words"""
        n = len(window_halves)
        arr = (C.c_void_p * n)(*[C.c_void_p(w) for w in window_halves])
        rw = C.c_int32()
        _check(lib().bl_groupby_export_partials_p2p_async(self.h, C.c_int32(n), C.c_int32(my_rank), arr, C.c_int64(rows_per_src), C.c_uint64(epoch), C.byref(rw)))
        return int(rw.value)

    def merge_window_async(self, own_half: int, n_ranks: int, rows_per_src: int, epoch: int):
        _check(lib().bl_groupby_merge_window_async(self.h, C.c_void_p(own_half), C.c_int32(n_ranks), C.c_int64(rows_per_src), C.c_uint64(epoch)))

    def defer_status(self, on: bool = True):
        lib().bl_groupby_defer_status.restype = None
        lib().bl_groupby_defer_status(self.h, C.c_int32(int(on)))

    def status(self) -> int:
        st = C.c_int32()
        _check(lib().bl_groupby_status(self.h, C.byref(st)))
        return int(st.value)

    def estimated_groups(self) -> int:
        lib().bl_groupby_estimated_groups.restype = C.c_int64
        return int(lib().bl_groupby_estimated_groups(self.h))

    def merge_partials(self, rows_dev_ptr: int, n_rows: int):
        _check(lib().bl_groupby_merge_partials(self.h, C.c_void_p(rows_dev_ptr), C.c_int64(n_rows)))

    def merge_partial_regions(self, ptrs: Sequence[int], counts: Sequence[int]):
        n = len(ptrs)
        pa = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in ptrs])
        ca = (C.c_int64 * n)(*[int(c) for c in counts])
        _check(lib().bl_groupby_merge_partial_regions(self.h, pa, ca, C.c_int32(n)))

    def finish(self, maintain_order: bool = False, location: int = HOST):
        ok, oa = BlColumn(), (BlColumn * max(self.n, 1))()
        _check(lib().bl_groupby_finish(self.h, C.c_int32(int(maintain_order)), C.c_int32(location), C.byref(ok), oa))
        res = _finish([ok] + list(oa)[: self.n], location)
        return res[0], res[1:]

    def reset(self):
        lib().bl_groupby_reset(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().bl_groupby_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Window:
    """Device memory exported over CUDA IPC so peer ranks can store into it (bl_window_*)."""

    def __init__(self, nbytes: int):
        self.h = C.c_void_p()
        buf = C.create_string_buffer(64)
        _check(lib().bl_window_create(C.c_size_t(nbytes), C.byref(self.h), buf))
        self.ipc_handle = bytes(buf.raw)
        lib().bl_window_ptr.restype = C.c_void_p
        self.ptr = int(lib().bl_window_ptr(self.h))
        self.nbytes = nbytes

    @staticmethod
    def open(ipc_handle: bytes) -> int:
        p = C.c_void_p()
        _check(lib().bl_window_open(C.create_string_buffer(ipc_handle, 64), C.byref(p)))
        return int(p.value)

    @staticmethod
    def close(ptr: int):
        lib().bl_window_close.restype = None
        lib().bl_window_close(C.c_void_p(ptr))

    def destroy(self):
        if self.h:
            lib().bl_window_destroy.restype = None
            lib().bl_window_destroy(self.h)
            self.h = None


def dev_free(ptr: int):
    lib().bl_dev_free(C.c_void_p(ptr))


def dev_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(lib().bl_dev_alloc(C.c_size_t(nbytes), C.byref(p)))
    return int(p.value)


def to_device(values: np.ndarray, valid=None) -> OutColumn:
    """Uploads a numpy column; returns a library-owned device column."""
    c = Column(values, valid)
    out = BlColumn()
    cs = c.struct()
    _check(lib().bl_column_to(C.byref(cs), C.c_int32(1), C.c_int32(DEVICE), C.byref(out)))
    return OutColumn(out)
