/*
 * oracle.c — CPU restatement of the reference (pola-rs/polars @ 4db92c1) semantics for the
 * hot path of SURVEY.md §8(a).  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library, and only as the checker / the CPU baseline.  The product path
 * (polars_b200/) never links, imports or calls it.
 *
 * Parity pinning: the reference itself cannot be built or imported in the authoring container
 * (no Rust toolchain, no wheel), so this restatement is pinned against the golden vectors lifted
 * from the reference's own tests (tests/golden/ JSON files, each citing file:line) and cross-checked
 * against pyarrow/pandas/numpy on random inputs (tests/test_oracle_*.py).
 * Third-party pieces absent from /root/reference: hashbrown 0.17.1 + foldhash 0.2.0 (hash table
 * and hasher; Cargo.lock:1934,1658).  They only decide the iteration order of UNORDERED group_by
 * output, which the reference's tests never pin; this file uses its own open-addressing table and
 * that order is "parity unpinned" (compare as a set).  Everything else is first-party and cited.
 *
 * Every function cites the reference file:line (relative to /root/reference/crates) it follows.
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#include <malloc.h>

/* Keep freed blocks in the heap instead of returning them to the kernel: the per-thread vectors of the group_by / join
 * restatements are re-allocated on every call, and with 64+ threads the mmap / page-fault traffic of glibc's default
 * policy (every block > 128 KB is its own mapping) serialises on the process' address-space lock — the Rayon original
 * keeps its thread-local buffers in jemalloc arenas and does not pay this. */
__attribute__((constructor)) static void or_malloc_policy(void) {
    mallopt(M_MMAP_THRESHOLD, 32 * 1024 * 1024);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 16 * 1024 * 1024);
}
#endif

typedef uint32_t idx_t;                 /* IdxSize = u32: polars-utils/src/index.rs:9 */
#define IDX_NULL 0xFFFFFFFFu            /* NullableIdxSize::null(): polars-utils/src/index.rs:50 */

/* ------------------------------------------------------------------------------------------
 * Bitmaps: LSB-first bit-packed, arbitrary bit offset.
 * polars-arrow/src/bitmap/utils/mod.rs:42-46, bitmap/immutable.rs:56-68,187-195
 * ------------------------------------------------------------------------------------------ */
static inline int get_bit(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
static inline void set_bit(uint8_t* b, int64_t i, int v) {
    if (v) b[i >> 3] |= (uint8_t)(1u << (i & 7)); else b[i >> 3] &= (uint8_t)~(1u << (i & 7));
}
/* valid(i) for an optional validity bitmap with bit offset */
static inline int is_valid(const uint8_t* validity, int64_t off, int64_t i) {
    return validity == NULL || get_bit(validity, off + i);
}

/* ------------------------------------------------------------------------------------------
 * Hashing / partitioning.  polars-utils/src/hashing.rs:62-69 (hash_to_partition),
 * :123-147 (DirtyHash, RANDOM_ODD), :183-187 (None -> 0); float canonicalisation
 * polars-utils/src/total_ord.rs:37-47,225-232.
 * ------------------------------------------------------------------------------------------ */
#define RANDOM_ODD 0x55fbfd6bfc5458e9ULL
static inline uint64_t dirty_hash_u64(uint64_t k) { return k * RANDOM_ODD; }
static inline uint64_t hash_to_partition(uint64_t h, uint64_t n) {
    return (uint64_t)(((unsigned __int128)h * (unsigned __int128)n) >> 64);
}
static inline double canonical_f64(double x) {
    double z = x + 0.0;                           /* -0.0 + 0.0 == +0.0 */
    if (z != z) { uint64_t b = 0x7ff8000000000000ULL; memcpy(&z, &b, 8); }
    return z;
}
static inline float canonical_f32(float x) {
    float z = x + 0.0f;
    if (z != z) { uint32_t b = 0x7fc00000u; memcpy(&z, &b, 4); }
    return z;
}

void or_dirty_hash_u64(const uint64_t* k, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = dirty_hash_u64(k[i]);
}
void or_hash_to_partition(const uint64_t* h, int64_t n, uint64_t n_partitions, uint64_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = hash_to_partition(h[i], n_partitions);
}
/* f64 keys -> canonical bit pattern (the TotalOrdItem that is hashed/compared) */
void or_canonical_f64_bits(const double* x, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; i++) { double c = canonical_f64(x[i]); memcpy(&out[i], &c, 8); }
}
void or_canonical_f32_bits(const float* x, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; i++) { float c = canonical_f32(x[i]); uint32_t b; memcpy(&b, &c, 4); out[i] = b; }
}

/* ------------------------------------------------------------------------------------------
 * Elementwise arithmetic.  polars-compute/src/arithmetic/signed.rs:23-232 (ints),
 * float.rs:21-122 (floats), polars-utils/src/floor_divmod.rs:38-66 (floor div/mod),
 * validity = AND of inputs (arity.rs:90), int //,% by zero -> null (signed.rs:35-70).
 * op codes shared with include/polars_b200.h.
 * mode: 0 = array (op) array, 1 = array (op) scalar, 2 = scalar (op) array.
 * For ints TRUE_DIV writes f64 output (binary.rs:74-94, signed.rs:13,216-228).
 * out_valid (may be NULL) is written as one byte per row (1 = valid) — test convenience.
 * ------------------------------------------------------------------------------------------ */
/* validity arguments of the or_* entry points are ONE BYTE PER ROW (numpy bool), NULL = all valid */
#define BV(v, i) ((v) == NULL || (v)[i])
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_FLOOR_DIV = 3, OP_MOD = 4, OP_TRUE_DIV = 5 };

#define DEF_INT_FLOORDIVMOD(NAME, T, UT)                                                    \
    static inline void NAME(T a, T b, T* d, T* m) {                                         \
        if (b == 0) { *d = 0; *m = 0; return; }                                             \
        T q, r;                                                                             \
        if (b == (T)-1) { q = (T)((UT)0 - (UT)a); r = 0; } /* wrapping_div / wrapping_rem */ \
        else { q = a / b; r = a % b; }                                                      \
        if (r != 0 && ((a < 0) != (b < 0))) { q -= 1; r += b; }                             \
        *d = q; *m = r;                                                                     \
    }
DEF_INT_FLOORDIVMOD(fdm_i64, int64_t, uint64_t)
DEF_INT_FLOORDIVMOD(fdm_i32, int32_t, uint32_t)
#define DEF_UINT_FLOORDIVMOD(NAME, T)                                                       \
    static inline void NAME(T a, T b, T* d, T* m) {                                         \
        if (b == 0) { *d = 0; *m = 0; return; }                                             \
        *d = a / b; *m = a % b;                                                             \
    }
DEF_UINT_FLOORDIVMOD(fdm_u64, uint64_t)
DEF_UINT_FLOORDIVMOD(fdm_u32, uint32_t)

#define DEF_INT_ARITH(NAME, T, UT, FDM)                                                     \
    void NAME(int op, int mode, const T* lhs, const uint8_t* lv, const T* rhs,              \
              const uint8_t* rv, int64_t n, T* out, double* out_f64, uint8_t* out_valid) {  \
        for (int64_t i = 0; i < n; i++) {                                                   \
            T a = (mode == 2) ? lhs[0] : lhs[i];                                            \
            T b = (mode == 1) ? rhs[0] : rhs[i];                                            \
            int va = (mode == 2) ? 1 : BV(lv, i);                                  \
            int vb = (mode == 1) ? 1 : BV(rv, i);                                  \
            int valid = va && vb;                                                           \
            T d, m;                                                                         \
            switch (op) {                                                                   \
                case OP_ADD: out[i] = (T)((UT)a + (UT)b); break;                            \
                case OP_SUB: out[i] = (T)((UT)a - (UT)b); break;                            \
                case OP_MUL: out[i] = (T)((UT)a * (UT)b); break;                            \
                case OP_FLOOR_DIV: FDM(a, b, &d, &m); out[i] = d; valid = valid && (b != 0); break; \
                case OP_MOD: FDM(a, b, &d, &m); out[i] = m; valid = valid && (b != 0); break; \
                case OP_TRUE_DIV:                                                           \
                    /* array/array: a as f64 / b as f64; scalar rhs: x * (1.0 / rhs) */     \
                    if (mode == 1) out_f64[i] = (double)a * (1.0 / (double)b);              \
                    else out_f64[i] = (double)a / (double)b;                                \
                    break;                                                                  \
            }                                                                               \
            if (out_valid) out_valid[i] = (uint8_t)valid;                                   \
        }                                                                                   \
    }
DEF_INT_ARITH(or_arith_i64, int64_t, uint64_t, fdm_i64)
DEF_INT_ARITH(or_arith_i32, int32_t, uint32_t, fdm_i32)
DEF_INT_ARITH(or_arith_u64, uint64_t, uint64_t, fdm_u64)
DEF_INT_ARITH(or_arith_u32, uint32_t, uint32_t, fdm_u32)

/* floats: float.rs:21-115; scalar-rhs floor_div/mod/true_div use the reciprocal (:80-83,
 * :98-101, :113-115). */
#define DEF_FLOAT_ARITH(NAME, T, FLOOR)                                                     \
    void NAME(int op, int mode, const T* lhs, const uint8_t* lv, const T* rhs,              \
              const uint8_t* rv, int64_t n, T* out, uint8_t* out_valid) {                   \
        for (int64_t i = 0; i < n; i++) {                                                   \
            T a = (mode == 2) ? lhs[0] : lhs[i];                                            \
            T b = (mode == 1) ? rhs[0] : rhs[i];                                            \
            int va = (mode == 2) ? 1 : BV(lv, i);                                  \
            int vb = (mode == 1) ? 1 : BV(rv, i);                                  \
            switch (op) {                                                                   \
                case OP_ADD: out[i] = a + b; break;                                         \
                case OP_SUB: out[i] = (mode == 1) ? a + (-b) : a - b; break;                \
                case OP_MUL: out[i] = a * b; break;                                         \
                case OP_FLOOR_DIV:                                                          \
                    if (mode == 1) { T inv = (T)1 / b; out[i] = FLOOR(a * inv); }           \
                    else out[i] = FLOOR(a / b);                                             \
                    break;                                                                  \
                case OP_MOD:                                                                \
                    if (mode == 1) { T inv = (T)1 / b; out[i] = a - b * FLOOR(a * inv); }   \
                    else out[i] = a - b * FLOOR(a / b);                                     \
                    break;                                                                  \
                case OP_TRUE_DIV:                                                           \
                    if (mode == 1) { T inv = (T)1 / b; out[i] = a * inv; }                  \
                    else out[i] = a / b;                                                    \
                    break;                                                                  \
            }                                                                               \
            if (out_valid) out_valid[i] = (uint8_t)(va && vb);                              \
        }                                                                                   \
    }
DEF_FLOAT_ARITH(or_arith_f64, double, floor)
DEF_FLOAT_ARITH(or_arith_f32, float, floorf)

/* ------------------------------------------------------------------------------------------
 * Comparisons -> boolean (one byte per row) + validity.
 * polars-compute/src/comparisons/mod.rs:4-76; total order for floats
 * polars-utils/src/total_ord.rs:317-364 (NaN == NaN, NaN is the greatest value);
 * nulls propagate (validity carried separately, polars-core/.../comparison/mod.rs:129-192).
 * missing=1: eq_missing / ne_missing (comparisons/mod.rs:14-52): null == null is true, result
 * never null.  mode as for arithmetic (0 array/array, 1 array/scalar).
 * ------------------------------------------------------------------------------------------ */
enum { CMP_EQ = 0, CMP_NE = 1, CMP_LT = 2, CMP_LE = 3, CMP_GT = 4, CMP_GE = 5 };

#define TOT_GE_INT(a, b) ((a) >= (b))
#define TOT_EQ_INT(a, b) ((a) == (b))
#define TOT_GE_FLT(a, b) (((a) != (a)) | ((a) >= (b)))             /* total_ord.rs:355-362 */
#define TOT_EQ_FLT(a, b) (((a) != (a)) ? ((b) != (b)) : ((a) == (b))) /* total_ord.rs:319-326 */

#define DEF_CMP(NAME, T, GE, EQ)                                                            \
    void NAME(int op, int mode, int missing, const T* lhs, const uint8_t* lv, const T* rhs, \
              const uint8_t* rv, int64_t n, uint8_t* out, uint8_t* out_valid) {             \
        for (int64_t i = 0; i < n; i++) {                                                   \
            T a = lhs[i];                                                                   \
            T b = (mode == 1) ? rhs[0] : rhs[i];                                            \
            int va = BV(lv, i);                                                    \
            int vb = (mode == 1) ? 1 : BV(rv, i);                                  \
            int r;                                                                          \
            switch (op) {                                                                   \
                case CMP_EQ: r = EQ(a, b); break;                                           \
                case CMP_NE: r = !EQ(a, b); break;                                          \
                case CMP_LT: r = !GE(a, b); break;              /* tot_lt = !tot_ge */      \
                case CMP_LE: r = GE(b, a); break;               /* tot_le = other.tot_ge */ \
                case CMP_GT: r = !GE(b, a); break;              /* tot_gt = other.tot_lt */ \
                default: r = GE(a, b); break;                                               \
            }                                                                               \
            if (missing) {                                                                  \
                if (op == CMP_EQ) r = (va && vb) ? r : (va == vb);                          \
                else if (op == CMP_NE) r = (va && vb) ? r : (va != vb);                     \
                if (out_valid) out_valid[i] = 1;                                            \
            } else if (out_valid) out_valid[i] = (uint8_t)(va && vb);                       \
            out[i] = (uint8_t)r;                                                            \
        }                                                                                   \
    }
DEF_CMP(or_cmp_i64, int64_t, TOT_GE_INT, TOT_EQ_INT)
DEF_CMP(or_cmp_i32, int32_t, TOT_GE_INT, TOT_EQ_INT)
DEF_CMP(or_cmp_u64, uint64_t, TOT_GE_INT, TOT_EQ_INT)
DEF_CMP(or_cmp_u32, uint32_t, TOT_GE_INT, TOT_EQ_INT)
DEF_CMP(or_cmp_f64, double, TOT_GE_FLT, TOT_EQ_FLT)
DEF_CMP(or_cmp_f32, float, TOT_GE_FLT, TOT_EQ_FLT)

/* ------------------------------------------------------------------------------------------
 * Filter.  polars-compute/src/filter/mod.rs:18-60: a null mask slot counts as false (:21-27);
 * values and validity of selected rows are compacted in row order.
 * mask / mask_valid are one byte per row; validity in/out one byte per row (NULL = all valid).
 * Returns the number of selected rows.  elem_size in bytes (4 or 8).
 * ------------------------------------------------------------------------------------------ */
int64_t or_filter(const void* values, const uint8_t* valid, int64_t n, int elem_size,
                  const uint8_t* mask, const uint8_t* mask_valid, void* out, uint8_t* out_valid) {
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        if (mask[i] && (mask_valid == NULL || mask_valid[i])) {
            memcpy((char*)out + k * elem_size, (const char*)values + i * elem_size, elem_size);
            if (out_valid) out_valid[k] = valid ? valid[i] : 1;
            k++;
        }
    }
    return k;
}

/* ------------------------------------------------------------------------------------------
 * Gather.  polars-compute/src/gather/primitive.rs:9-66: out[i] = values[idx[i]]; a null index
 * yields T::default() (0) and a null slot; validity of the source is gathered too.
 * ------------------------------------------------------------------------------------------ */
void or_gather(const void* values, const uint8_t* valid, int elem_size, const idx_t* idx,
               const uint8_t* idx_valid, int64_t m, void* out, uint8_t* out_valid) {
    for (int64_t i = 0; i < m; i++) {
        int iv = idx_valid == NULL || idx_valid[i];
        if (iv) memcpy((char*)out + i * elem_size, (const char*)values + (int64_t)idx[i] * elem_size, elem_size);
        else memset((char*)out + i * elem_size, 0, elem_size);
        if (out_valid) out_valid[i] = (uint8_t)(iv && (valid == NULL || valid[idx[i]]));
    }
}

/* ------------------------------------------------------------------------------------------
 * A small open-addressing map u64 key (+null flag) -> u32 payload.  Stand-in for
 * PlHashMap = hashbrown + foldhash (polars-utils/src/aliases.rs:5-13); only value semantics
 * matter.  Iteration order = slot order (unpinned, see header).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t* keys; uint32_t* vals; uint8_t* used; uint64_t cap, len;
    int has_null; uint32_t null_val;
} map_t;
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static void map_init(map_t* m, uint64_t expect) {
    uint64_t cap = 16; while (cap < expect * 2) cap <<= 1;
    m->cap = cap; m->len = 0; m->has_null = 0; m->null_val = 0;
    m->keys = (uint64_t*)malloc(cap * 8); m->vals = (uint32_t*)malloc(cap * 4);
    m->used = (uint8_t*)calloc(cap, 1);
}
static void map_free(map_t* m) { free(m->keys); free(m->vals); free(m->used); }
static void map_grow(map_t* m);
/* returns pointer to the payload; *inserted = 1 when the key was new (payload uninitialised) */
static inline uint32_t* map_entry(map_t* m, uint64_t key, int* inserted) {
    if ((m->len + 1) * 2 > m->cap) map_grow(m);
    uint64_t mask = m->cap - 1, s = mix64(key) & mask;
    while (m->used[s]) { if (m->keys[s] == key) { *inserted = 0; return &m->vals[s]; } s = (s + 1) & mask; }
    m->used[s] = 1; m->keys[s] = key; m->len++; *inserted = 1; return &m->vals[s];
}
static inline const uint32_t* map_get(const map_t* m, uint64_t key) {
    uint64_t mask = m->cap - 1, s = mix64(key) & mask;
    while (m->used[s]) { if (m->keys[s] == key) return &m->vals[s]; s = (s + 1) & mask; }
    return NULL;
}
static void map_grow(map_t* m) {
    map_t n; uint64_t oc = m->cap; map_init(&n, m->cap);   /* doubles */
    for (uint64_t s = 0; s < oc; s++) if (m->used[s]) { int ins; *map_entry(&n, m->keys[s], &ins) = m->vals[s]; }
    n.has_null = m->has_null; n.null_val = m->null_val;
    map_free(m); *m = n;
}

/* ------------------------------------------------------------------------------------------
 * Group-by build.  polars-core/src/frame/group_by/hashing.rs:116-167 (group_by_threaded_slice /
 * _iter: P tables, every "thread" scans ALL keys and keeps those with
 * hash_to_partition(dirty_hash(k), P) == thread_no), :76-113 (single-threaded group_by),
 * :26-73 (finish_group_order: sorted=true => groups ordered by first row index),
 * into_groups.rs:25-58 (multithreaded only if len > 1000; null key is its own group; keys are
 * the unsigned bit representation :179-186 / canonical float bits).
 * Null keys: Option<T>::dirty_hash() == 0 (hashing.rs:183-187) -> partition 0.
 *
 * keys: u64 bit representation; key_valid: one byte per row or NULL.
 * Outputs (caller-allocated, capacity n / n+1 / n): first[g], offsets[g..g+1], idx[].
 * Within a group indices ascend (row-order scan).  Group order: sorted => ascending first;
 * otherwise partition-major, table slot order (UNPINNED).
 * Returns number of groups.
 * ------------------------------------------------------------------------------------------ */
typedef struct { idx_t first; idx_t gid; } firstgid_t;
static int cmp_firstgid(const void* a, const void* b) {
    idx_t x = ((const firstgid_t*)a)->first, y = ((const firstgid_t*)b)->first;
    return (x > y) - (x < y);
}

int64_t or_group_by(const uint64_t* keys, const uint8_t* key_valid, int64_t n, int n_partitions,
                    int sorted, idx_t* first, uint64_t* offsets, idx_t* idx) {
    if (n == 0) { offsets[0] = 0; return 0; }
    int P = n_partitions < 1 ? 1 : n_partitions;
    if (n <= 1000) P = 1;                                /* into_groups.rs:25-28 */
    /* per-partition: local group count, first idx / count per local group, and the partition's rows in scan
     * order with their local group id — the restatement of the per-group IdxVec pushes of hashing.rs:142-153.
     * Everything a thread writes during the scan is private to it (the reference's threads own their tables
     * and vectors the same way), so the port scales with threads like the Rayon original. */
    int64_t* part_ngroups = (int64_t*)calloc(P, sizeof(int64_t));
    int64_t* part_nrows = (int64_t*)calloc(P, sizeof(int64_t));
    idx_t** part_first = (idx_t**)calloc(P, sizeof(idx_t*));
    idx_t** part_count = (idx_t**)calloc(P, sizeof(idx_t*));
    idx_t** part_rows = (idx_t**)calloc(P, sizeof(idx_t*));
    idx_t** part_gids = (idx_t**)calloc(P, sizeof(idx_t*));

#pragma omp parallel for schedule(static, 1) num_threads(P > 1 ? P : 1) if (P > 1)
    for (int t = 0; t < P; t++) {
        map_t m; map_init(&m, 512);                      /* _HASHMAP_INIT_SIZE */
        int64_t cap = 1024, ng = 0;
        int64_t rcap = n / P + n / (8 * P) + 1024, nr = 0;
        idx_t* pf = (idx_t*)malloc(cap * sizeof(idx_t));
        idx_t* pc = (idx_t*)malloc(cap * sizeof(idx_t));
        idx_t* rows = (idx_t*)malloc((size_t)rcap * sizeof(idx_t));
        idx_t* gids = (idx_t*)malloc((size_t)rcap * sizeof(idx_t));
        for (int64_t i = 0; i < n; i++) {
            int valid = key_valid == NULL || key_valid[i];
            uint64_t k = keys[i];
            uint64_t h = valid ? dirty_hash_u64(k) : 0;
            if (P > 1 && (int)hash_to_partition(h, (uint64_t)P) != t) continue;
            idx_t g;
            if (!valid) {
                if (!m.has_null) { m.has_null = 1; m.null_val = (idx_t)ng; g = (idx_t)ng; goto newgroup; }
                g = m.null_val;
            } else {
                int ins; uint32_t* slot = map_entry(&m, k, &ins);
                if (ins) { *slot = (uint32_t)ng; g = (idx_t)ng; goto newgroup; }
                g = *slot;
            }
            pc[g]++; goto push;
        newgroup:
            if (ng == cap) { cap *= 2; pf = (idx_t*)realloc(pf, cap * sizeof(idx_t)); pc = (idx_t*)realloc(pc, cap * sizeof(idx_t)); }
            pf[ng] = (idx_t)i; pc[ng] = 1; ng++;
        push:
            if (nr == rcap) { rcap += rcap / 2 + 1024; rows = (idx_t*)realloc(rows, (size_t)rcap * sizeof(idx_t)); gids = (idx_t*)realloc(gids, (size_t)rcap * sizeof(idx_t)); }
            rows[nr] = (idx_t)i; gids[nr] = g; nr++;
        }
        /* Local group ids are in first-occurrence order; the reference iterates the hashbrown
         * table instead (hashing.rs:157-160) — order unpinned when !sorted. */
        map_free(&m);
        part_ngroups[t] = ng; part_nrows[t] = nr; part_first[t] = pf; part_count[t] = pc; part_rows[t] = rows; part_gids[t] = gids;
    }
    /* global group numbering: partition-major (flatten, hashing.rs:66-71), then optional sort */
    int64_t* part_base = (int64_t*)malloc((P + 1) * sizeof(int64_t));
    part_base[0] = 0;
    for (int t = 0; t < P; t++) part_base[t + 1] = part_base[t] + part_ngroups[t];
    int64_t G = part_base[P];
    firstgid_t* fg = (firstgid_t*)malloc((size_t)G * sizeof(firstgid_t));
    idx_t* counts = (idx_t*)malloc((size_t)G * sizeof(idx_t));
#pragma omp parallel for schedule(static, 1) num_threads(P > 1 ? P : 1) if (P > 1)
    for (int t = 0; t < P; t++)
        for (int64_t j = 0; j < part_ngroups[t]; j++) {
            fg[part_base[t] + j].first = part_first[t][j]; fg[part_base[t] + j].gid = (idx_t)(part_base[t] + j);
            counts[part_base[t] + j] = part_count[t][j];
        }
    if (sorted) qsort(fg, (size_t)G, sizeof(firstgid_t), cmp_firstgid);   /* hashing.rs:61 */
    idx_t* newpos = (idx_t*)malloc((size_t)G * sizeof(idx_t));             /* old gid -> output position */
    offsets[0] = 0;
    for (int64_t p = 0; p < G; p++) {
        newpos[fg[p].gid] = (idx_t)p; first[p] = fg[p].first; offsets[p + 1] = offsets[p] + counts[fg[p].gid];
    }
    /* fill idx lists: every partition replays ITS rows in scan order => ascending inside each group */
    uint64_t* cursor = (uint64_t*)malloc((size_t)G * sizeof(uint64_t));
    memcpy(cursor, offsets, (size_t)G * sizeof(uint64_t));
#pragma omp parallel for schedule(static, 1) num_threads(P > 1 ? P : 1) if (P > 1)
    for (int t = 0; t < P; t++) {
        const idx_t* rows = part_rows[t]; const idx_t* gids = part_gids[t];
        const int64_t base = part_base[t];
        for (int64_t j = 0; j < part_nrows[t]; j++) {
            idx_t p = newpos[base + gids[j]];
            idx[cursor[p]++] = rows[j];
        }
    }
    for (int t = 0; t < P; t++) { free(part_first[t]); free(part_count[t]); free(part_rows[t]); free(part_gids[t]); }
    free(part_first); free(part_count); free(part_rows); free(part_gids); free(part_ngroups); free(part_nrows); free(part_base);
    free(fg); free(counts); free(newpos); free(cursor);
    return G;
}

/* ------------------------------------------------------------------------------------------
 * Per-group aggregations over GroupsIdx {first, all}.
 * sum:  polars-core/src/frame/group_by/aggregations/mod.rs:854-879 — ints fold a+b (wrapping in
 *       release builds, polars-compute/src/sum.rs:13-49), floats sequential KahanSum in row order
 *       (polars-utils/src/kahan_sum.rs:36-47); nulls skipped; empty / all-null -> 0, never null.
 * mean: :939-977 (floats) / :1227-1267 (ints): Kahan f64 sum / (len - null_count); single-row
 *       group returns the value itself (or null); all-null -> null (take_agg/mod.rs:80-84).
 * min/max: :486-518 / :669-703: reduce with min_ignore_nan / max_ignore_nan
 *       (polars-utils/src/min_max.rs:41-48 ints, :96-108 floats = f64::min/max, NaN ignored
 *       unless all NaN); nulls skipped; all-null -> null.
 * len:  position.rs:555-569 — group size incl. nulls (IdxSize).
 * count: aggregations/dispatch.rs:25-55 — non-null rows per group.
 * valid: one byte per row or NULL.  out_valid one byte per group (NULL allowed for sum/len).
 * ------------------------------------------------------------------------------------------ */
typedef struct { double sum, err; } kahan_t;
static inline void kahan_add(kahan_t* k, double rhs) {      /* kahan_sum.rs:36-47 */
    double y = rhs - k->err; double new_sum = k->sum + y; double new_err = (new_sum - k->sum) - y;
    k->sum = new_sum; if (isfinite(new_err)) k->err = new_err;
}
typedef struct { float sum, err; } kahanf_t;
static inline void kahanf_add(kahanf_t* k, float rhs) {
    float y = rhs - k->err; float new_sum = k->sum + y; float new_err = (new_sum - k->sum) - y;
    k->sum = new_sum; if (isfinite(new_err)) k->err = new_err;
}

void or_agg_len(const uint64_t* offsets, int64_t G, idx_t* out) {
    for (int64_t g = 0; g < G; g++) out[g] = (idx_t)(offsets[g + 1] - offsets[g]);
}
void or_agg_count(const uint8_t* valid, const uint64_t* offsets, const idx_t* idx, int64_t G, idx_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; g++) {
        idx_t c = 0;
        for (uint64_t j = offsets[g]; j < offsets[g + 1]; j++) c += (valid == NULL || valid[idx[j]]);
        out[g] = c;
    }
}
void or_agg_sum_i64(const int64_t* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,
                    int64_t G, int64_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; g++) {
        uint64_t s = 0;
        for (uint64_t j = offsets[g]; j < offsets[g + 1]; j++)
            if (valid == NULL || valid[idx[j]]) s += (uint64_t)v[idx[j]];
        out[g] = (int64_t)s;
    }
}
void or_agg_sum_i32(const int32_t* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,
                    int64_t G, int32_t* out) {   /* Int32 sums stay Int32 (series/implementations/mod.rs:145-154) */
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; g++) {
        uint32_t s = 0;
        for (uint64_t j = offsets[g]; j < offsets[g + 1]; j++)
            if (valid == NULL || valid[idx[j]]) s += (uint32_t)v[idx[j]];
        out[g] = (int32_t)s;
    }
}
void or_agg_sum_u64(const uint64_t* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,
                    int64_t G, uint64_t* out) { or_agg_sum_i64((const int64_t*)v, valid, offsets, idx, G, (int64_t*)out); }
void or_agg_sum_u32(const uint32_t* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,
                    int64_t G, uint32_t* out) { or_agg_sum_i32((const int32_t*)v, valid, offsets, idx, G, (int32_t*)out); }
void or_agg_sum_f64(const double* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,
                    int64_t G, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; g++) {
        uint64_t a = offsets[g], b = offsets[g + 1];
        if (b == a) { out[g] = 0.0; continue; }
        if (b - a == 1) { out[g] = (valid == NULL || valid[idx[a]]) ? v[idx[a]] : 0.0; continue; }
        kahan_t k = {0.0, 0.0};
        for (uint64_t j = a; j < b; j++) if (valid == NULL || valid[idx[j]]) kahan_add(&k, v[idx[j]]);
        out[g] = k.sum;
    }
}
void or_agg_sum_f32(const float* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,
                    int64_t G, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; g++) {
        uint64_t a = offsets[g], b = offsets[g + 1];
        if (b == a) { out[g] = 0.0f; continue; }
        if (b - a == 1) { out[g] = (valid == NULL || valid[idx[a]]) ? v[idx[a]] : 0.0f; continue; }
        kahanf_t k = {0.0f, 0.0f};
        for (uint64_t j = a; j < b; j++) if (valid == NULL || valid[idx[j]]) kahanf_add(&k, v[idx[j]]);
        out[g] = k.sum;
    }
}
/* mean over values viewed as f64 (is_int: values are int64; else double) */
#define DEF_AGG_MEAN(NAME, T)                                                               \
    void NAME(const T* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,  \
              int64_t G, double* out, uint8_t* out_valid) {                                 \
        _Pragma("omp parallel for schedule(static)")                                        \
        for (int64_t g = 0; g < G; g++) {                                                   \
            uint64_t a = offsets[g], b = offsets[g + 1];                                    \
            if (b == a) { out[g] = 0; out_valid[g] = 0; continue; }                         \
            if (b - a == 1) {                                                               \
                int ok = valid == NULL || valid[idx[a]];                                    \
                out[g] = ok ? (double)v[idx[a]] : 0; out_valid[g] = (uint8_t)ok; continue;  \
            }                                                                               \
            kahan_t k = {0.0, 0.0}; uint64_t nulls = 0;                                     \
            for (uint64_t j = a; j < b; j++) {                                              \
                if (valid == NULL || valid[idx[j]]) kahan_add(&k, (double)v[idx[j]]);       \
                else nulls++;                                                               \
            }                                                                               \
            if (nulls == b - a) { out[g] = 0; out_valid[g] = 0; }                           \
            else { out[g] = k.sum / ((double)(b - a) - (double)nulls); out_valid[g] = 1; }  \
        }                                                                                   \
    }
DEF_AGG_MEAN(or_agg_mean_i64, int64_t)
DEF_AGG_MEAN(or_agg_mean_i32, int32_t)
DEF_AGG_MEAN(or_agg_mean_u64, uint64_t)
DEF_AGG_MEAN(or_agg_mean_u32, uint32_t)
DEF_AGG_MEAN(or_agg_mean_f64, double)
DEF_AGG_MEAN(or_agg_mean_f32, float)   /* caller casts the f64 result back to f32 (:976) */

/* var / std: polars-core/src/frame/group_by/aggregations/mod.rs:1020-1178 -> take_var_*_primitive_iter_unchecked
 * (polars-arrow/src/legacy/kernels/take_agg/var.rs:11-41): Welford's online update in row order over the group's
 * non-null values converted to f64; None when count <= ddof; std = sqrt(var).  Output f64 (Float32 input: caller casts). */
#define DEF_AGG_VAR(NAME, T)                                                                \
    void NAME(const T* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,  \
              int64_t G, int ddof, int is_std, double* out, uint8_t* out_valid) {           \
        _Pragma("omp parallel for schedule(static)")                                        \
        for (int64_t g = 0; g < G; g++) {                                                   \
            double m2 = 0.0, mean = 0.0; uint64_t count = 0;                                \
            for (uint64_t j = offsets[g]; j < offsets[g + 1]; j++) {                        \
                if (valid != NULL && !valid[idx[j]]) continue;                              \
                double value = (double)v[idx[j]];                                           \
                uint64_t new_count = count + 1;                                             \
                double delta_1 = value - mean;                                              \
                double new_mean = delta_1 / (double)new_count + mean;                       \
                double delta_2 = value - new_mean;                                          \
                m2 = m2 + delta_1 * delta_2; count = new_count; mean = new_mean;            \
            }                                                                               \
            if (count <= (uint64_t)ddof) { out[g] = 0; out_valid[g] = 0; }                  \
            else { double r = m2 / ((double)count - (double)ddof); out[g] = is_std ? sqrt(r) : r; out_valid[g] = 1; } \
        }                                                                                   \
    }
DEF_AGG_VAR(or_agg_var_i64, int64_t)
DEF_AGG_VAR(or_agg_var_i32, int32_t)
DEF_AGG_VAR(or_agg_var_u64, uint64_t)
DEF_AGG_VAR(or_agg_var_u32, uint32_t)
DEF_AGG_VAR(or_agg_var_f64, double)
DEF_AGG_VAR(or_agg_var_f32, float)

#define MIN_INT(a, b) ((a) < (b) ? (a) : (b))
#define MAX_INT(a, b) ((a) < (b) ? (b) : (a))
#define DEF_AGG_MINMAX(NAME, T, RED)                                                        \
    void NAME(const T* v, const uint8_t* valid, const uint64_t* offsets, const idx_t* idx,  \
              int64_t G, T* out, uint8_t* out_valid) {                                      \
        _Pragma("omp parallel for schedule(static)")                                        \
        for (int64_t g = 0; g < G; g++) {                                                   \
            int have = 0; T acc = 0;                                                        \
            for (uint64_t j = offsets[g]; j < offsets[g + 1]; j++) {                        \
                if (valid != NULL && !valid[idx[j]]) continue;                              \
                T x = v[idx[j]];                                                            \
                if (!have) { acc = x; have = 1; } else acc = RED(acc, x);                   \
            }                                                                               \
            out[g] = have ? acc : 0; out_valid[g] = (uint8_t)have;                          \
        }                                                                                   \
    }
DEF_AGG_MINMAX(or_agg_min_i64, int64_t, MIN_INT)
DEF_AGG_MINMAX(or_agg_max_i64, int64_t, MAX_INT)
DEF_AGG_MINMAX(or_agg_min_i32, int32_t, MIN_INT)
DEF_AGG_MINMAX(or_agg_max_i32, int32_t, MAX_INT)
DEF_AGG_MINMAX(or_agg_min_u64, uint64_t, MIN_INT)
DEF_AGG_MINMAX(or_agg_max_u64, uint64_t, MAX_INT)
DEF_AGG_MINMAX(or_agg_min_u32, uint32_t, MIN_INT)
DEF_AGG_MINMAX(or_agg_max_u32, uint32_t, MAX_INT)
DEF_AGG_MINMAX(or_agg_min_f64, double, fmin)     /* f64::min == IEEE minNum == C fmin */
DEF_AGG_MINMAX(or_agg_max_f64, double, fmax)
DEF_AGG_MINMAX(or_agg_min_f32, float, fminf)
DEF_AGG_MINMAX(or_agg_max_f32, float, fmaxf)

/* ------------------------------------------------------------------------------------------
 * Hash join, single u64 key.
 * build_tables: polars-ops/src/frame/join/hash_join/single_keys.rs:16-167 — radix partition of
 *   the build keys into P = n_threads partitions (count :52-66, cumulative offsets :69-93,
 *   scatter keys + row idx :96-121), one table per partition mapping key -> ascending row idx
 *   list (:123-166); null keys skipped unless nulls_equal (:41,:148); < 2*128 keys => one table
 *   (:35-48, MIN_ELEMS_PER_THREAD = 128 in release).
 * probe: single_keys_inner.rs:11-38,78-148 — each probe slice in row order; for a hit emit
 *   (probe idx, build idx) for every build idx in ascending order; swap_fn restores (left,right).
 * which side builds: hash_join/mod.rs:41-50 — left probes iff left.len() > right.len(); else
 *   right probes and swapped = true (tie => swapped).
 * left join: single_keys_left.rs:106-195 — left always probes; miss => (idx, NULL).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int P;
    map_t* maps;            /* key -> list id (per partition) */
    uint64_t** list_off;    /* per partition: offsets into list_idx */
    idx_t** list_idx;
} jtables_t;

static void build_tables(const uint64_t* keys, const uint8_t* valid, int64_t n, int n_threads,
                         int nulls_equal, jtables_t* jt) {
    int P = n_threads < 1 ? 1 : n_threads;
    if (n < 2 * 128) P = 1;
    jt->P = P;
    jt->maps = (map_t*)calloc(P, sizeof(map_t));
    jt->list_off = (uint64_t**)calloc(P, sizeof(uint64_t*));
    jt->list_idx = (idx_t**)calloc(P, sizeof(idx_t*));
    /* count -> cumsum -> scatter (thread t owns the t-th contiguous slice of the build keys) */
    int64_t* slice_lo = (int64_t*)malloc((P + 1) * sizeof(int64_t));
    for (int t = 0; t <= P; t++) slice_lo[t] = (int64_t)((__int128)n * t / P);
    int64_t* sizes = (int64_t*)calloc((size_t)P * P, sizeof(int64_t));   /* [t][p] */
#pragma omp parallel for schedule(static, 1) num_threads(P) if (P > 1)
    for (int t = 0; t < P; t++)
        for (int64_t i = slice_lo[t]; i < slice_lo[t + 1]; i++) {
            int v = valid == NULL || valid[i];
            uint64_t h = v ? dirty_hash_u64(keys[i]) : 0;
            sizes[(size_t)t * P + hash_to_partition(h, (uint64_t)P)]++;
        }
    int64_t* off = (int64_t*)malloc((size_t)P * P * sizeof(int64_t));
    int64_t* part_off = (int64_t*)malloc((P + 1) * sizeof(int64_t));
    int64_t cum = 0;
    for (int p = 0; p < P; p++) {
        part_off[p] = cum;
        for (int t = 0; t < P; t++) { off[(size_t)t * P + p] = cum; cum += sizes[(size_t)t * P + p]; }
    }
    part_off[P] = cum;
    uint64_t* sk = (uint64_t*)malloc((size_t)(n ? n : 1) * 8);
    idx_t* si = (idx_t*)malloc((size_t)(n ? n : 1) * sizeof(idx_t));
    uint8_t* sv = (uint8_t*)malloc((size_t)(n ? n : 1));
#pragma omp parallel for schedule(static, 1) num_threads(P) if (P > 1)
    for (int t = 0; t < P; t++) {
        int64_t* o = &off[(size_t)t * P];
        for (int64_t i = slice_lo[t]; i < slice_lo[t + 1]; i++) {
            int v = valid == NULL || valid[i];
            uint64_t h = v ? dirty_hash_u64(keys[i]) : 0;
            int64_t d = o[hash_to_partition(h, (uint64_t)P)]++;
            sk[d] = keys[i]; si[d] = (idx_t)i; sv[d] = (uint8_t)v;
        }
    }
    /* per-partition tables; scatter order keeps row idx ascending inside a partition */
#pragma omp parallel for schedule(static, 1) num_threads(P) if (P > 1)
    for (int p = 0; p < P; p++) {
        int64_t lo = part_off[p], hi = part_off[p + 1];
        map_t* m = &jt->maps[p]; map_init(m, 512);
        /* pass 1: list ids + counts */
        int64_t cap = 1024, nl = 0; uint64_t* cnt = (uint64_t*)calloc(cap, 8);
        idx_t* lid = (idx_t*)malloc((size_t)(hi - lo + 1) * sizeof(idx_t));
        for (int64_t i = lo; i < hi; i++) {
            if (!sv[i] && !nulls_equal) { lid[i - lo] = IDX_NULL; continue; }
            idx_t l;
            if (!sv[i]) {
                if (!m->has_null) { m->has_null = 1; m->null_val = (uint32_t)nl; l = (idx_t)nl++; }
                else l = m->null_val;
            } else {
                int ins; uint32_t* s = map_entry(m, sk[i], &ins);
                if (ins) { *s = (uint32_t)nl; l = (idx_t)nl++; } else l = *s;
            }
            if (nl > cap) { cnt = (uint64_t*)realloc(cnt, cap * 2 * 8); memset(cnt + cap, 0, cap * 8); cap *= 2; }
            cnt[l]++; lid[i - lo] = l;
        }
        uint64_t* lo_ = (uint64_t*)malloc((size_t)(nl + 1) * 8); lo_[0] = 0;
        for (int64_t l = 0; l < nl; l++) lo_[l + 1] = lo_[l] + cnt[l];
        idx_t* li = (idx_t*)malloc((size_t)(lo_[nl] + 1) * sizeof(idx_t));
        memset(cnt, 0, (size_t)nl * 8);
        for (int64_t i = lo; i < hi; i++) { idx_t l = lid[i - lo]; if (l != IDX_NULL) li[lo_[l] + cnt[l]++] = si[i]; }
        free(cnt); free(lid);
        jt->list_off[p] = lo_; jt->list_idx[p] = li;
    }
    free(slice_lo); free(sizes); free(off); free(part_off); free(sk); free(si); free(sv);
}
static void free_tables(jtables_t* jt) {
    for (int p = 0; p < jt->P; p++) { map_free(&jt->maps[p]); free(jt->list_off[p]); free(jt->list_idx[p]); }
    free(jt->maps); free(jt->list_off); free(jt->list_idx);
}
/* look a probe key up: returns list [*b, *e) of ascending build idxs, or 0 */
static inline int probe_key(const jtables_t* jt, uint64_t k, int valid, int nulls_equal,
                            const idx_t** b, const idx_t** e) {
    uint64_t h = valid ? dirty_hash_u64(k) : 0;
    int p = (int)hash_to_partition(h, (uint64_t)jt->P);
    const map_t* m = &jt->maps[p]; uint32_t l;
    if (!valid) { if (!nulls_equal || !m->has_null) return 0; l = m->null_val; }
    else { const uint32_t* s = map_get(m, k); if (!s) return 0; l = *s; }
    *b = jt->list_idx[p] + jt->list_off[p][l]; *e = jt->list_idx[p] + jt->list_off[p][l + 1];
    return 1;
}

typedef struct { idx_t* a; idx_t* b; int64_t len, cap; } pairs_t;
static inline void pairs_push(pairs_t* p, idx_t a, idx_t b) {
    if (p->len == p->cap) { p->cap = p->cap ? p->cap * 2 : 1024; p->a = (idx_t*)realloc(p->a, p->cap * sizeof(idx_t)); p->b = (idx_t*)realloc(p->b, p->cap * sizeof(idx_t)); }
    p->a[p->len] = a; p->b[p->len] = b; p->len++;
}

/* how: 0 inner, 1 left.  Outputs are malloc'd (free with or_free); returns number of tuples.
 * Right idx of an unmatched left row (left join) = IDX_NULL. */
int64_t or_hash_join(const uint64_t* lk, const uint8_t* lvalid, int64_t nl, const uint64_t* rk,
                     const uint8_t* rvalid, int64_t nr, int how, int nulls_equal, int n_threads,
                     idx_t** out_left, idx_t** out_right) {
    int swapped = 0;
    const uint64_t *pk = lk, *bk = rk; const uint8_t *pv = lvalid, *bv = rvalid; int64_t np = nl, nb = nr;
    if (how == 0 && !(nl > nr)) { swapped = 1; pk = rk; pv = rvalid; np = nr; bk = lk; bv = lvalid; nb = nl; }
    jtables_t jt; build_tables(bk, bv, nb, n_threads, nulls_equal, &jt);
    int T = n_threads < 1 ? 1 : n_threads;
    pairs_t* res = (pairs_t*)calloc(T, sizeof(pairs_t));
#pragma omp parallel for schedule(static, 1) num_threads(T) if (T > 1)
    for (int t = 0; t < T; t++) {
        int64_t lo = (int64_t)((__int128)np * t / T), hi = (int64_t)((__int128)np * (t + 1) / T);
        pairs_t* r = &res[t];
        for (int64_t i = lo; i < hi; i++) {
            int v = pv == NULL || pv[i];
            const idx_t *b, *e;
            if (probe_key(&jt, pk[i], v, nulls_equal, &b, &e)) {
                for (const idx_t* q = b; q < e; q++) { if (swapped) pairs_push(r, *q, (idx_t)i); else pairs_push(r, (idx_t)i, *q); }
            } else if (how == 1) pairs_push(r, (idx_t)i, IDX_NULL);
        }
    }
    int64_t total = 0; for (int t = 0; t < T; t++) total += res[t].len;
    idx_t* L = (idx_t*)malloc((size_t)(total ? total : 1) * sizeof(idx_t));
    idx_t* R = (idx_t*)malloc((size_t)(total ? total : 1) * sizeof(idx_t));
    int64_t* starts = (int64_t*)malloc((size_t)(T + 1) * sizeof(int64_t));
    starts[0] = 0;
    for (int t = 0; t < T; t++) starts[t + 1] = starts[t] + res[t].len;
    /* flatten (single_keys_inner.rs:118-148 flattens the per-thread vectors in parallel too) */
#pragma omp parallel for schedule(static, 1) num_threads(T) if (T > 1)
    for (int t = 0; t < T; t++) {
        if (res[t].len) { memcpy(L + starts[t], res[t].a, (size_t)res[t].len * sizeof(idx_t)); memcpy(R + starts[t], res[t].b, (size_t)res[t].len * sizeof(idx_t)); }
        free(res[t].a); free(res[t].b);
    }
    free(starts);
    free(res); free_tables(&jt);
    *out_left = L; *out_right = R;
    return total;
}
void or_free(void* p) { free(p); }

/* maintain_order: stable sort of the tuples on the requested side
 * (polars-ops/src/frame/join/mod.rs:583-642).  by_right = 0 sorts on the left idx. */
typedef struct { idx_t k, o; int64_t pos; } sortrec_t;
static int cmp_sortrec(const void* a, const void* b) {
    const sortrec_t *x = (const sortrec_t*)a, *y = (const sortrec_t*)b;
    if (x->k != y->k) return (x->k > y->k) - (x->k < y->k);
    return (x->pos > y->pos) - (x->pos < y->pos);
}
void or_stable_sort_pairs(idx_t* left, idx_t* right, int64_t n, int by_right) {
    sortrec_t* r = (sortrec_t*)malloc((size_t)(n ? n : 1) * sizeof(sortrec_t));
    for (int64_t i = 0; i < n; i++) { r[i].k = by_right ? right[i] : left[i]; r[i].o = by_right ? left[i] : right[i]; r[i].pos = i; }
    qsort(r, (size_t)n, sizeof(sortrec_t), cmp_sortrec);
    for (int64_t i = 0; i < n; i++) { if (by_right) { right[i] = r[i].k; left[i] = r[i].o; } else { left[i] = r[i].k; right[i] = r[i].o; } }
    free(r);
}

int or_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* Thread count of the aggregation / elementwise loops (the Rayon pool size of the reference: POOL.install). */
void or_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int or_hw_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
