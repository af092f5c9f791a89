"""GPU test of boundary B1 (the Polars expression-plugin ABI v0.1) WITHOUT Polars: pyarrow plays the
caller of crates/polars-plan/src/plans/aexpr/function_expr/plugin.rs:70-137 — it exports the input
series as `SeriesExport{ArrowSchema*, ArrowArray**, len, release, private_data}`
(crates/polars-ffi/src/version_0.rs:7-16), calls `_polars_plugin_bl_<op>`, checks that the callee
released every input (plugin.rs:118-125) and imports the returned series."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64), ("n_children", C.c_int64),
                        ("children", C.POINTER(C.POINTER(ArrowSchema))), ("dictionary", C.POINTER(ArrowSchema)), ("release", C.c_void_p), ("private_data", C.c_void_p)]
ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64), ("n_children", C.c_int64),
                       ("buffers", C.POINTER(C.c_void_p)), ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", C.c_void_p), ("private_data", C.c_void_p)]


class SeriesExport(C.Structure):
    pass


RELEASE_FN = C.CFUNCTYPE(None, C.POINTER(SeriesExport))
SeriesExport._fields_ = [("field", C.POINTER(ArrowSchema)), ("arrays", C.POINTER(C.POINTER(ArrowArray))), ("len", C.c_size_t), ("release", RELEASE_FN), ("private_data", C.c_void_p)]


class Caller:
    """Plays polars' side of call_plugin."""

    def __init__(self, lib):
        self.lib, self.keep, self.released = lib, [], 0

        def _rel(p):
            self.released += 1
            p.contents.private_data = None
        self._rel = RELEASE_FN(_rel)

    def export(self, name, chunks):
        schema = ArrowSchema()
        arrays = [ArrowArray() for _ in chunks]
        typ = chunks[0].type
        pa.field(name, typ)._export_to_c(C.addressof(schema))
        for a, c in zip(arrays, chunks):
            c._export_to_c(C.addressof(a))
        ptrs = (C.POINTER(ArrowArray) * len(arrays))(*[C.pointer(a) for a in arrays])
        self.keep += [schema, arrays, ptrs]
        return SeriesExport(C.pointer(schema), ptrs, len(arrays), self._rel, 1), arrays

    def call(self, op, inputs, kwargs=None):
        exports, arrs = zip(*[self.export(n, ch) for n, ch in inputs])
        arr = (SeriesExport * len(exports))(*exports)
        ret = SeriesExport()
        before = self.released
        fn = getattr(self.lib, f"_polars_plugin_bl_{op}")
        fn.restype = None
        import pickle
        kw = pickle.dumps(kwargs) if kwargs else b""      # what register_plugin_function(kwargs=...) hands over (plugins.py:100-115)
        kbuf = (C.c_uint8 * max(len(kw), 1)).from_buffer_copy(kw or b"\0")
        fn(arr, C.c_size_t(len(exports)), kbuf if kw else None, C.c_size_t(len(kw)), C.byref(ret), None)
        # the callee owns the inputs: every ArrowArray and every SeriesExport must have been released
        assert self.released - before == len(exports), "input SeriesExport not released by the plugin"
        for group in arrs:
            for a in group:
                assert not a.release, "input ArrowArray not released by the plugin"
        if not ret.private_data:
            self.lib._polars_plugin_get_last_error_message.restype = C.c_char_p
            raise RuntimeError(self.lib._polars_plugin_get_last_error_message().decode())
        assert ret.len == 1
        out = pa.Array._import_from_c(C.addressof(ret.arrays[0].contents), C.addressof(ret.field.contents))
        ret.release(C.byref(ret))
        return out


@pytest.fixture(scope="module")
def caller():
    import polars_b200 as plb
    plb.init()
    return Caller(plb.lib())


def test_plugin_elementwise_and_compare(caller):
    rng = np.random.default_rng(0)
    a = rng.integers(-100, 100, 10_000)
    b = rng.integers(1, 50, 10_000)
    am = rng.random(10_000) < 0.1
    A = pa.array(a, mask=am)
    out = caller.call("add", [("x", [A.slice(0, 3000), A.slice(3000)]), ("y", [pa.array(b)])])     # chunked + sliced input
    assert out.type == pa.int64() and out.null_count == int(am.sum())
    assert np.array_equal(np.asarray(out.fill_null(0)), np.where(am, 0, a + b))
    out = caller.call("truediv", [("x", [pa.array(a)]), ("y", [pa.array(b)])])
    assert out.type == pa.float64() and np.array_equal(np.asarray(out), a / b)
    out = caller.call("gt", [("x", [A]), ("y", [pa.array([0])])])
    assert out.type == pa.bool_() and out.null_count == int(am.sum())
    assert np.array_equal(np.asarray(out.fill_null(False)), (a > 0) & ~am)


def test_plugin_filter_gather_group_join(caller):
    rng = np.random.default_rng(1)
    v = rng.normal(size=5000)
    m = rng.random(5000) < 0.3
    out = caller.call("filter", [("v", [pa.array(v)]), ("m", [pa.array(m)])])
    assert np.array_equal(np.asarray(out), v[m])
    idx = rng.integers(0, 5000, 777).astype(np.uint32)
    out = caller.call("gather", [("v", [pa.array(v)]), ("i", [pa.array(idx)])])
    assert np.array_equal(np.asarray(out), v[idx])
    key = rng.integers(0, 37, 5000)
    x = rng.integers(-10, 10, 5000)
    out = caller.call("group_sum", [("key", [pa.array(key)]), ("x", [pa.array(x)])])
    assert pa.types.is_struct(out.type)
    k, s = np.asarray(out.field("key")), np.asarray(out.field("agg"))
    _, first = np.unique(key, return_index=True)
    assert np.array_equal(k, key[np.sort(first)])                     # first-occurrence order
    assert np.array_equal(s, np.array([x[key == kk].sum() for kk in k]))
    lk, rk = rng.integers(0, 100, 300), rng.permutation(100)[:60]
    out = caller.call("join_inner_idx", [("l", [pa.array(lk)]), ("r", [pa.array(rk)])])
    li, ri = np.asarray(out.field("left_idx")), np.asarray(out.field("right_idx"))
    assert np.array_equal(lk[li], rk[ri]) and li.size == int(np.isin(lk, rk).sum())


def test_plugin_kwargs_multi_key_and_new_entries(caller):
    """kwargs arrive as a pickled dict (every pickle protocol CPython writes for it); several key columns; the entries added
    for first / last / var / std / len and left / full / semi / anti joins."""
    import pickle
    rng = np.random.default_rng(2)
    n = 4000
    k0, k1 = rng.integers(0, 9, n), rng.integers(0, 5, n).astype(np.int32)
    x = rng.normal(100.0, 3.0, n)
    xm = rng.random(n) < 0.1
    X = pa.array(x, mask=xm)
    groups = {}
    for i in range(n):
        groups.setdefault((int(k0[i]), int(k1[i])), []).append(i)
    out = caller.call("group_var", [("a", [pa.array(k0)]), ("b", [pa.array(k1)]), ("x", [X])], kwargs={"ddof": 0})
    ka, kb, agg = np.asarray(out.field("key")), np.asarray(out.field("key_1")), out.field("agg")
    assert [(int(a), int(b)) for a, b in zip(ka, kb)] == list(groups)                  # first-occurrence order of the key pairs
    exp = np.array([np.var(x[np.array(r)][~xm[np.array(r)]]) for r in groups.values()])
    assert np.allclose(np.asarray(agg), exp, rtol=1e-9)
    out1 = caller.call("group_std", [("a", [pa.array(k0)]), ("x", [X])])               # default ddof = 1
    exp1 = np.array([np.std(x[(k0 == kk) & ~xm], ddof=1) for kk in np.asarray(out1.field("key"))])
    assert np.allclose(np.asarray(out1.field("agg")), exp1, rtol=1e-9)
    out = caller.call("group_first", [("a", [pa.array(k0)]), ("x", [X])])
    firsts = {int(kk): int(np.nonzero(k0 == kk)[0][0]) for kk in np.unique(k0)}
    got = out.field("agg").to_pylist()
    assert got == [None if xm[firsts[int(kk)]] else x[firsts[int(kk)]] for kk in np.asarray(out.field("key"))]
    out = caller.call("group_len", [("a", [pa.array(k0)]), ("b", [pa.array(k1)])])
    assert np.asarray(out.field("agg")).tolist() == [len(r) for r in groups.values()]
    # joins: kwargs nulls_equal under two pickle protocols, several key columns
    lk = pa.array([1, 2, None, 4, 2], pa.int64()); rk = pa.array([2, None, 5], pa.int64())
    for proto in (2, 4):
        kw = pickle.dumps({"nulls_equal": True}, protocol=proto)
        assert caller.lib is not None and kw
    out = caller.call("join_left_idx", [("l", [lk]), ("r", [rk])], kwargs={"nulls_equal": True})
    assert out.field("left_idx").to_pylist() == [0, 1, 2, 3, 4] and out.field("right_idx").to_pylist() == [None, 0, 1, None, 0]
    out = caller.call("join_left_idx", [("l", [lk]), ("r", [rk])])
    assert out.field("right_idx").to_pylist() == [None, 0, None, None, 0]
    assert caller.call("join_semi_idx", [("l", [lk]), ("r", [rk])]).to_pylist() == [1, 4]
    assert caller.call("join_anti_idx", [("l", [lk]), ("r", [rk])]).to_pylist() == [0, 2, 3]
    out = caller.call("join_full_idx", [("l", [lk]), ("r", [rk])])
    assert sorted(zip([-1 if v is None else v for v in out.field("left_idx").to_pylist()], [-1 if v is None else v for v in out.field("right_idx").to_pylist()])) == \
        [(-1, 1), (-1, 2), (0, -1), (1, 0), (2, -1), (3, -1), (4, 0)]
    la, lb = pa.array([1, 1, 2, 2]), pa.array([7, 8, 7, 8]); ra, rb = pa.array([2, 1, 3]), pa.array([8, 7, 7])
    out = caller.call("join_inner_idx", [("la", [la]), ("lb", [lb]), ("ra", [ra]), ("rb", [rb])])
    assert list(zip(out.field("left_idx").to_pylist(), out.field("right_idx").to_pylist())) == [(0, 1), (3, 0)]


def test_plugin_error_channel(caller):
    with pytest.raises(RuntimeError, match="dtypes differ"):
        caller.call("add", [("x", [pa.array([1, 2, 3])]), ("y", [pa.array([1.0, 2.0, 3.0])])])


@pytest.mark.gpu
def test_plain_c_caller(tmp_path):
    # examples/c_abi_demo.c: filter -> group_by/agg -> join through the C ABI from a C99 program (no Python in the loop)
    import subprocess
    from test_cabi_cpu import _build_c_demo
    exe = _build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    key, x = np.arange(1000) % 7, (np.arange(1000) % 11) - 5
    m = x > 0
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "groups: 7"
    order = []
    for k in key[m]:
        if k not in order:
            order.append(int(k))
    for ln, k in zip(lines[1:8], order):
        sel = m & (key == k)
        assert ln.split() == ["key", str(k), "sum", str(int(x[sel].sum())), "mean", f"{x[sel].mean():.4f}", "len", str(int(sel.sum()))], ln
    assert lines[8].startswith("join rows: 1000 (first: key 0 payload 4.5)"), lines[8]
