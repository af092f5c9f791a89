// plugin.cu — boundary B1: the Polars expression-plugin ABI v0.1, served by the CUDA operators.
//
// A stock Polars build dlopen()s this library through `register_plugin_function`
// (py-polars/src/polars/plugins.py:24-37) and calls
//   _polars_plugin_get_version()                       crates/polars-plan/src/plans/aexpr/function_expr/plugin.rs:36-52
//   _polars_plugin_<name>(inputs, n, kwargs, kwargs_len, ret, ctx)          plugin.rs:70-137
//   _polars_plugin_field_<name>(fields, n, out, kwargs, kwargs_len)         plugin.rs:178-217
//   _polars_plugin_get_last_error_message()                                 plugin.rs:139-158
// with SeriesExport / CallerContext as in crates/polars-ffi/src/version_0.rs:7-16,134-140 and
// Arrow C Data Interface structs (crates/polars-arrow/src/ffi/generated.rs:6-34).
// Ownership (plugin.rs:118-125): the caller forgets the inputs, so this side releases every input
// array and every input SeriesExport; the return value carries its own release callbacks.
// Errors: `ret` is left untouched (private_data == NULL) and the message is kept thread-local.
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"
#include "groupby.h"

using namespace plb;

extern "C" {
struct ArrowSchema {
    const char* format; const char* name; const char* metadata; int64_t flags; int64_t n_children;
    struct ArrowSchema** children; struct ArrowSchema* dictionary; void (*release)(struct ArrowSchema*); void* private_data;
};
struct ArrowArray {
    int64_t length; int64_t null_count; int64_t offset; int64_t n_buffers; int64_t n_children;
    const void** buffers; struct ArrowArray** children; struct ArrowArray* dictionary; void (*release)(struct ArrowArray*); void* private_data;
};
struct SeriesExport {
    ArrowSchema* field; ArrowArray** arrays; size_t len; void (*release)(SeriesExport*); void* private_data;
};
struct CallerContext { uint64_t bitflags; };
}

static thread_local std::string t_plugin_error;

static int dtype_from_format(const char* f) {
    if (!f || !f[0] || f[1]) return -1;
    switch (f[0]) {
        case 'c': return BL_INT8; case 's': return BL_INT16; case 'i': return BL_INT32; case 'l': return BL_INT64;
        case 'C': return BL_UINT8; case 'S': return BL_UINT16; case 'I': return BL_UINT32; case 'L': return BL_UINT64;
        case 'f': return BL_FLOAT32; case 'g': return BL_FLOAT64; case 'b': return BL_BOOL;
        default: return -1;
    }
}
static const char* format_of(int dt) {
    static const char* f[] = {"c", "s", "i", "l", "C", "S", "I", "L", "f", "g", "b"};
    return f[dt];
}

// ---- schema / array construction with release callbacks -----------------------------------------
struct SchemaPriv { std::string name, format; std::vector<ArrowSchema*> kids; };
static void release_schema(ArrowSchema* s) {
    if (!s || !s->release) return;
    auto* p = reinterpret_cast<SchemaPriv*>(s->private_data);
    for (auto* k : p->kids) { if (k->release) k->release(k); free(k); }
    delete p;
    s->release = nullptr;
}
static void fill_schema(ArrowSchema* out, const std::string& name, const std::string& format, const std::vector<std::pair<std::string, int>>& children = {}) {
    auto* p = new SchemaPriv{name, format, {}};
    for (auto& c : children) {
        auto* k = reinterpret_cast<ArrowSchema*>(calloc(1, sizeof(ArrowSchema)));
        fill_schema(k, c.first, format_of(c.second));
        p->kids.push_back(k);
    }
    out->format = p->format.c_str(); out->name = p->name.c_str(); out->metadata = nullptr;
    out->flags = 2 /* ARROW_FLAG_NULLABLE */; out->n_children = (int64_t)p->kids.size();
    out->children = p->kids.empty() ? nullptr : p->kids.data(); out->dictionary = nullptr;
    out->release = release_schema; out->private_data = p;
}

struct ArrayPriv { bl_column col; const void* bufs[2]; std::vector<ArrowArray*> kids; };
static void release_array(ArrowArray* a) {
    if (!a || !a->release) return;
    auto* p = reinterpret_cast<ArrayPriv*>(a->private_data);
    for (auto* k : p->kids) { if (k->release) k->release(k); free(k); }
    bl_column_free(&p->col);
    delete p;
    a->release = nullptr;
}
static void fill_array(ArrowArray* out, const bl_column& col) {   // takes ownership of a host bl_column
    auto* p = new ArrayPriv();
    p->col = col; p->bufs[0] = col.validity; p->bufs[1] = col.values;
    out->length = col.length; out->null_count = col.validity ? -1 : 0; out->offset = 0; out->n_buffers = 2; out->n_children = 0;
    out->buffers = p->bufs; out->children = nullptr; out->dictionary = nullptr; out->release = release_array; out->private_data = p;
}
static void fill_struct_array(ArrowArray* out, const std::vector<bl_column>& kids) {
    auto* p = new ArrayPriv();
    memset(&p->col, 0, sizeof p->col);
    p->bufs[0] = nullptr;
    for (auto& k : kids) { auto* a = reinterpret_cast<ArrowArray*>(calloc(1, sizeof(ArrowArray))); fill_array(a, k); p->kids.push_back(a); }
    out->length = kids.empty() ? 0 : kids[0].length; out->null_count = 0; out->offset = 0; out->n_buffers = 1; out->n_children = (int64_t)p->kids.size();
    out->buffers = p->bufs; out->children = p->kids.data(); out->dictionary = nullptr; out->release = release_array; out->private_data = p;
}

struct SeriesPriv { ArrowSchema* schema; ArrowArray** arrays; size_t n; };
static void release_series(SeriesExport* e) {
    if (!e || !e->release) return;
    auto* p = reinterpret_cast<SeriesPriv*>(e->private_data);
    // the importer moved the ArrowArray structs out (ptr::read, version_0.rs:92-104): free the boxes only
    for (size_t i = 0; i < p->n; i++) free(p->arrays[i]);
    free(p->arrays);
    if (p->schema->release) p->schema->release(p->schema);
    free(p->schema);
    delete p;
    e->release = nullptr;
}
static void make_series(SeriesExport* ret, ArrowSchema* schema, ArrowArray* array) {
    auto* p = new SeriesPriv{schema, reinterpret_cast<ArrowArray**>(calloc(1, sizeof(ArrowArray*))), 1};
    p->arrays[0] = array;
    ret->field = schema; ret->arrays = p->arrays; ret->len = 1; ret->release = release_series; ret->private_data = p;
}

// ---- inputs --------------------------------------------------------------------------------------
static std::vector<bl_column> input_chunks(const SeriesExport& s, int* dtype_out) {
    int dt = dtype_from_format(s.field ? s.field->format : nullptr);
    PLB_REQUIRE(dt >= 0, BL_ERR_UNSUPPORTED, std::string("plugin: unsupported input dtype '") + (s.field && s.field->format ? s.field->format : "?") + "'");
    std::vector<bl_column> ch;
    for (size_t i = 0; i < s.len; i++) {
        const ArrowArray* a = s.arrays[i];
        bl_column c; memset(&c, 0, sizeof c);
        c.dtype = dt; c.location = BL_HOST; c.length = a->length; c.offset = a->offset; c.null_count = a->null_count;
        c.validity = a->n_buffers > 0 ? reinterpret_cast<const uint8_t*>(a->buffers[0]) : nullptr;
        c.values = a->n_buffers > 1 ? a->buffers[1] : nullptr;
        ch.push_back(c);
    }
    if (ch.empty()) { bl_column c; memset(&c, 0, sizeof c); c.dtype = dt; c.null_count = 0; static const uint64_t z = 0; c.values = &z; ch.push_back(c); }
    *dtype_out = dt;
    return ch;
}
static void release_inputs(SeriesExport* inputs, size_t n) {
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < inputs[i].len; j++) { ArrowArray* a = inputs[i].arrays[j]; if (a && a->release) a->release(a); }
        if (inputs[i].release) inputs[i].release(&inputs[i]);
    }
}

enum PluginOp { P_ARITH, P_CMP, P_FILTER, P_GATHER, P_GROUP, P_JOIN };

static void run_plugin(PluginOp kind, int op, SeriesExport* inputs, size_t n, SeriesExport* ret) {
    std::lock_guard<std::recursive_mutex> lk(ctx().mu);
    PLB_REQUIRE(n == 2, BL_ERR_INVALID, "plugin: expected exactly 2 input series");
    int dt0, dt1;
    std::vector<bl_column> c0 = input_chunks(inputs[0], &dt0), c1 = input_chunks(inputs[1], &dt1);
    const std::string name = inputs[0].field && inputs[0].field->name ? inputs[0].field->name : "";
    DevCol a = import_column(c0.data(), (int)c0.size()), b = import_column(c1.data(), (int)c1.size());
    auto* schema = reinterpret_cast<ArrowSchema*>(calloc(1, sizeof(ArrowSchema)));
    auto* array = reinterpret_cast<ArrowArray*>(calloc(1, sizeof(ArrowArray)));
    try {
        if (kind == P_ARITH || kind == P_CMP || kind == P_FILTER || kind == P_GATHER) {
            DevCol o;
            if (kind == P_ARITH) o = op_elementwise(op, a, b);
            else if (kind == P_CMP) o = op_compare(op, a, b, false);
            else if (kind == P_FILTER) { std::vector<DevCol> outs; op_filter({a}, b, outs); o = outs[0]; }
            else { std::vector<DevCol> outs; op_gather({a}, b, true, outs); o = outs[0]; }
            bl_column h; export_column(o, BL_HOST, &h);
            fill_array(array, h);
            fill_schema(schema, name, format_of(o.dtype));
        } else if (kind == P_GROUP) {
            // inputs: key, value  ->  struct {key, agg} in first-occurrence order
            GroupByState st(a.dtype, {op}, {b.dtype}, {b.validity != nullptr ? 1 : 0}, 0, true);
            st.consume_all(a, {&b});
            DevCol ok; std::vector<DevCol> oa;
            st.finish(true, &a, ok, oa);
            bl_column hk, hv; export_column(ok, BL_HOST, &hk);
            try { export_column(oa[0], BL_HOST, &hv); } catch (...) { bl_column_free(&hk); throw; }
            fill_struct_array(array, {hk, hv});
            fill_schema(schema, name, "+s", {{"key", ok.dtype}, {"agg", oa[0].dtype}});
        } else {
            JoinResult jr = op_hash_join(a, b, BL_JOIN_INNER, false, BL_ORDER_NONE);
            bl_column hl, hr; export_column(jr.left, BL_HOST, &hl);
            try { export_column(jr.right, BL_HOST, &hr); } catch (...) { bl_column_free(&hl); throw; }
            fill_struct_array(array, {hl, hr});
            fill_schema(schema, name, "+s", {{"left_idx", BL_UINT32}, {"right_idx", BL_UINT32}});
        }
    } catch (...) { free(schema); free(array); throw; }
    make_series(ret, schema, array);
}

static void plugin_entry(PluginOp kind, int op, SeriesExport* inputs, size_t n, SeriesExport* ret) {
    try { run_plugin(kind, op, inputs, n, ret); }
    catch (const std::exception& e) { t_plugin_error = e.what(); cudaGetLastError(); }
    catch (...) { t_plugin_error = "PANIC"; }      // special-cased by the caller (plugin.rs:219-221)
    release_inputs(inputs, n);
}

static void field_entry(PluginOp kind, int op, const ArrowSchema* fields, size_t n, ArrowSchema* out) {
    const std::string name = n > 0 && fields[0].name ? fields[0].name : "";
    const int dt = n > 0 ? dtype_from_format(fields[0].format) : -1;
    const int vdt = n > 1 ? dtype_from_format(fields[1].format) : -1;
    switch (kind) {
        case P_ARITH: fill_schema(out, name, format_of((op == BL_OP_TRUE_DIV && dt >= 0 && dt <= BL_UINT64) ? BL_FLOAT64 : (dt < 0 ? BL_INT64 : dt))); break;
        case P_CMP: fill_schema(out, name, "b"); break;
        case P_FILTER: case P_GATHER: fill_schema(out, name, format_of(dt < 0 ? BL_INT64 : dt)); break;
        case P_GROUP: {
            int adt = vdt < 0 ? BL_INT64 : vdt;
            if (op == BL_AGG_MEAN) adt = vdt == BL_FLOAT32 ? BL_FLOAT32 : BL_FLOAT64;
            if (op == BL_AGG_COUNT || op == BL_AGG_LEN) adt = BL_UINT32;
            fill_schema(out, name, "+s", {{"key", dt < 0 ? BL_INT64 : dt}, {"agg", adt}});
            break;
        }
        default: fill_schema(out, name, "+s", {{"left_idx", BL_UINT32}, {"right_idx", BL_UINT32}}); break;
    }
}

#define PLUGIN(NAME, KIND, OP)                                                                                            \
    void _polars_plugin_bl_##NAME(SeriesExport* inputs, size_t n, const uint8_t*, size_t, SeriesExport* ret, CallerContext*) { \
        plugin_entry(KIND, OP, inputs, n, ret);                                                                           \
    }                                                                                                                     \
    void _polars_plugin_field_bl_##NAME(const ArrowSchema* fields, size_t n, ArrowSchema* out, const uint8_t*, size_t) {  \
        try { field_entry(KIND, OP, fields, n, out); } catch (...) { t_plugin_error = "PANIC"; }                          \
    }

extern "C" {
uint32_t _polars_plugin_get_version(void) { return (0u << 16) + 1u; }     // (major 0, minor 1): polars-ffi/src/lib.rs:12-17
const char* _polars_plugin_get_last_error_message(void) { return t_plugin_error.c_str(); }

PLUGIN(add, P_ARITH, BL_OP_ADD)
PLUGIN(sub, P_ARITH, BL_OP_SUB)
PLUGIN(mul, P_ARITH, BL_OP_MUL)
PLUGIN(floordiv, P_ARITH, BL_OP_FLOOR_DIV)
PLUGIN(mod, P_ARITH, BL_OP_MOD)
PLUGIN(truediv, P_ARITH, BL_OP_TRUE_DIV)
PLUGIN(eq, P_CMP, BL_CMP_EQ)
PLUGIN(ne, P_CMP, BL_CMP_NE)
PLUGIN(lt, P_CMP, BL_CMP_LT)
PLUGIN(le, P_CMP, BL_CMP_LE)
PLUGIN(gt, P_CMP, BL_CMP_GT)
PLUGIN(ge, P_CMP, BL_CMP_GE)
PLUGIN(filter, P_FILTER, 0)
PLUGIN(gather, P_GATHER, 0)
PLUGIN(group_sum, P_GROUP, BL_AGG_SUM)
PLUGIN(group_mean, P_GROUP, BL_AGG_MEAN)
PLUGIN(group_min, P_GROUP, BL_AGG_MIN)
PLUGIN(group_max, P_GROUP, BL_AGG_MAX)
PLUGIN(group_count, P_GROUP, BL_AGG_COUNT)
PLUGIN(join_inner_idx, P_JOIN, 0)
}
