/*
 * polars_b200.h — C ABI of the B200-native hot-path library (libpolars_b200.so).
 *
 * This is the drop-in boundary B3 of SURVEY.md §8(b): the operator-level entry points a thin
 * Rust `extern "C"` crate inside polars-mem-engine binds in place of the Rayon dispatch of
 *   FilterExec::execute        crates/polars-mem-engine/src/executors/filter.rs:60-144
 *   group_by_helper            crates/polars-mem-engine/src/executors/group_by.rs:60-98
 *   DataFrameJoinOps::_join_impl  crates/polars-ops/src/frame/join/mod.rs:125-459
 *   DataFrame::take_unchecked  crates/polars-core/src/frame/mod.rs:1238-1294
 *   apply_operator             crates/polars-expr/src/expressions/binary.rs:61-131
 * (binding stubs: INTEGRATION.md).  Plain C: pointers, sizes, PODs.  No torch / C++ types.
 *
 * Memory model.  A `bl_column` describes one contiguous Arrow-layout primitive array
 * (crates/polars-arrow/src/array/primitive/mod.rs:56-60): a values buffer, an optional
 * LSB-first bit-packed validity bitmap (crates/polars-arrow/src/bitmap/immutable.rs:56-68) and
 * a logical element `offset` that applies to both (so arbitrary bitmap bit offsets are legal).
 * `location` says where the buffers live: BL_HOST (pageable or pinned host memory; pinned
 * buffers — bl_alloc_pinned or cudaHostRegister'ed — are DMA'd directly, pageable ones are staged
 * through an internal pinned ring) or BL_DEVICE (device pointers on the library's device).
 * A ChunkedArray (crates/polars-core/src/chunked_array/mod.rs:139-148) is an array of
 * bl_column chunks; the library concatenates chunks while uploading (the "rechunk" the
 * reference does first: executors/group_by.rs:70, join/mod.rs:194-218).
 *
 * Ownership.  Inputs are caller-owned and only read.  Outputs are written into caller-provided
 * `bl_column` structs; their buffers are library-owned (`owner != NULL`) in the requested
 * `out_location` and released with bl_column_free().  Nothing is retained across calls.
 *
 * Errors.  Every entry point returns a bl_status; BL_OK == 0.  On failure outputs are untouched
 * and bl_last_error() returns a thread-local NUL-terminated message (the channel the plugin ABI
 * forwards through _polars_plugin_get_last_error_message).  No exceptions cross the boundary,
 * the process is never aborted.  If the CUDA runtime or a device is missing every call fails
 * with BL_ERR_CUDA — there is no CPU fallback.
 *
 * Threading.  Entry points are thread-safe (calls are serialised per library context).
 */
#ifndef POLARS_B200_H
#define POLARS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BL_ABI_VERSION 1

typedef int32_t bl_status;
enum {
    BL_OK = 0,
    BL_ERR_INVALID = 1,     /* bad argument (null pointer, length mismatch, ...) */
    BL_ERR_CUDA = 2,        /* CUDA runtime / driver error, or no device */
    BL_ERR_OOM = 3,         /* device or pinned-host allocation failed */
    BL_ERR_UNSUPPORTED = 4, /* dtype / mode outside the hot path (caller falls back to its CPU path) */
    BL_ERR_DTYPE = 5,       /* key dtypes differ etc. (reference: ComputeError, join/mod.rs:231-241) */
    BL_ERR_BOUNDS = 6       /* gather index out of bounds (reference: check_bounds_ca, gather.rs:14-39) */
};

/* physical dtypes (Arrow primitive types) */
enum {
    BL_INT8 = 0, BL_INT16 = 1, BL_INT32 = 2, BL_INT64 = 3,
    BL_UINT8 = 4, BL_UINT16 = 5, BL_UINT32 = 6, BL_UINT64 = 7,
    BL_FLOAT32 = 8, BL_FLOAT64 = 9,
    BL_BOOL = 10            /* bit-packed values buffer (BooleanArray) */
};

enum { BL_HOST = 0, BL_DEVICE = 1 };

typedef struct bl_column {
    int32_t dtype;            /* BL_INT64 ... */
    int32_t location;         /* BL_HOST | BL_DEVICE */
    int64_t length;           /* logical number of rows */
    int64_t offset;           /* logical element offset into values and validity */
    int64_t null_count;       /* -1 = unknown */
    const void* values;       /* length+offset elements (BL_BOOL: bits) */
    const uint8_t* validity;  /* NULL = no nulls */
    void* owner;              /* NULL = caller-owned; else release with bl_column_free */
} bl_column;

/* IdxSize = u32 (crates/polars-utils/src/index.rs:9); null index (left join, gather) */
#define BL_IDX_NULL 0xFFFFFFFFu

/* ---- lifecycle ------------------------------------------------------------------------- */
int32_t bl_abi_version(void);
/* Binds the library to a CUDA device (-1 = current device / LOCAL_RANK env).  Idempotent. */
bl_status bl_init(int32_t device);
void bl_shutdown(void);
const char* bl_last_error(void);
/* device facts for the host side: writes sm count, L2 bytes, total/free HBM bytes */
bl_status bl_device_info(int32_t* sm_count, int64_t* l2_bytes, int64_t* hbm_total, int64_t* hbm_free);

/* Deterministic aggregation: on != 0 makes bl_groupby_agg / bl_groupby_agg_keys build the reference's
 * GroupsIdx and fold every group sequentially in row order with the reference's reducers (sequential Kahan float sums,
 * aggregations/mod.rs:854-977): results are bit-identical from run to run and to the reference's in-memory engine, floats
 * included, at a fraction of the fused path's speed.  Also enabled by the environment variable BL_DETERMINISTIC=1. */
void bl_set_deterministic(int32_t on);

/* ---- memory ---------------------------------------------------------------------------- */
/* Pinned host buffers (the SharedStorage::ForeignOwner seam, crates/polars-buffer/src/storage.rs:37-53). */
bl_status bl_alloc_pinned(size_t bytes, void** out);
void bl_free_pinned(void* p);
bl_status bl_dev_alloc(size_t bytes, void** out);
void bl_dev_free(void* p);
bl_status bl_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
bl_status bl_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
/* Copies a (possibly chunked) column to the other location; result is library-owned. */
bl_status bl_column_to(const bl_column* chunks, int32_t n_chunks, int32_t location, bl_column* out);
void bl_column_free(bl_column* col);
/* Blocks until all work queued by this thread's calls has finished (outputs are already
 * complete when an entry point returns; this is for bl_dev_* / profiling users). */
bl_status bl_sync(void);
/* The CUDA stream (cudaStream_t) every kernel of this library is launched on. */
void* bl_stream(void);

/* ---- K1: elementwise arithmetic  (ArithmeticKernel, polars-compute/src/arithmetic/mod.rs:8-76) */
enum { BL_OP_ADD = 0, BL_OP_SUB = 1, BL_OP_MUL = 2, BL_OP_FLOOR_DIV = 3, BL_OP_MOD = 4, BL_OP_TRUE_DIV = 5 };
/* lhs (op) rhs.  A length-1 side broadcasts as a scalar (apply_operator, binary.rs:61-131) and
 * takes the reference's scalar code path (e.g. float x / c == x * (1/c), float.rs:113-115).
 * Integer FLOOR_DIV / MOD by zero yield null (signed.rs:35-70); integer TRUE_DIV yields FLOAT64.
 * Output validity = AND of input validities (arity.rs:90). */
bl_status bl_elementwise(int32_t op, const bl_column* lhs, const bl_column* rhs, int32_t out_location, bl_column* out);

/* ---- K2: comparisons -> BooleanArray  (TotalOrdKernel/TotalEqKernel, comparisons/mod.rs:4-76) */
enum { BL_CMP_EQ = 0, BL_CMP_NE = 1, BL_CMP_LT = 2, BL_CMP_LE = 3, BL_CMP_GT = 4, BL_CMP_GE = 5 };
/* Total order on floats: NaN == NaN, NaN greatest (polars-utils/src/total_ord.rs:317-364).
 * missing != 0 selects eq_missing / ne_missing (null == null, never-null result). */
bl_status bl_compare(int32_t op, const bl_column* lhs, const bl_column* rhs, int32_t missing, int32_t out_location, bl_column* out);

/* ---- K3: filter  (polars-compute/src/filter/mod.rs:18-110; DataFrame::filter frame/mod.rs:1148-1180) */
/* mask: BL_BOOL column; a null mask slot counts as false.  All n_cols columns are compacted with
 * one mask pass.  outs[i] keeps cols[i].dtype. */
bl_status bl_filter(const bl_column* cols, int32_t n_cols, const bl_column* mask, int32_t out_location, bl_column* outs);
/* Fused predicate + compaction: keep rows where cols[pred_col] (cmp_op) scalar
 * (FilterExec over BinaryExpr(col, op, lit), executors/filter.rs:117-144).  scalar: length-1 column. */
bl_status bl_filter_cmp(const bl_column* cols, int32_t n_cols, int32_t pred_col, int32_t cmp_op, const bl_column* scalar,
                        int32_t out_location, bl_column* outs);

/* ---- K4: gather  (take_primitive_unchecked, polars-compute/src/gather/primitive.rs:9-78) */
/* idx: BL_UINT32 column (IdxSize).  A null idx slot or idx == BL_IDX_NULL gives a null row with
 * value 0.  check_bounds != 0 validates idx < col.length first (BL_ERR_BOUNDS). */
bl_status bl_gather(const bl_column* cols, int32_t n_cols, const bl_column* idx, int32_t check_bounds, int32_t out_location, bl_column* outs);

/* ---- K5: hash group_by + aggregation --------------------------------------------------- */
/* (group_by_threaded_slice hashing.rs:116-167 + agg_sum/mean/min/max aggregations/mod.rs:486-1018,
 *  fused: index lists are never materialised) */
enum { BL_AGG_SUM = 0, BL_AGG_MEAN = 1, BL_AGG_MIN = 2, BL_AGG_MAX = 3, BL_AGG_COUNT = 4, BL_AGG_LEN = 5,
       /* evaluated per group over the reference's GroupsIdx (row lists in row order), not by the fused atomics path: */
       BL_AGG_FIRST = 6, BL_AGG_LAST = 7,   /* value at the group's first / last row, nulls included (dispatch.rs:57-120) */
       BL_AGG_VAR = 8, BL_AGG_STD = 9,      /* Welford in row order, f64; null when count <= ddof (aggregations/mod.rs:1020-1178, take_agg/var.rs:11-41) */
       BL_AGG_N_UNIQUE = 10 };              /* distinct values per group, a null counts as one (aggregations/dispatch.rs:285-345); UInt32; bl_groupby_agg only */
/* delta degrees of freedom of VAR / STD travel in bits 16..23 of `kind` (Polars' default is 1) */
#define BL_AGG_WITH_DDOF(kind, ddof) ((kind) | ((ddof) << 16))
typedef struct bl_agg {
    int32_t kind;           /* BL_AGG_* */
    int32_t n_chunks;       /* chunks of the value column (ignored for BL_AGG_LEN) */
    const bl_column* values;
} bl_agg;
/* Single numeric key (ints are grouped on their bit pattern, floats canonicalised: -0 == +0,
 * all NaNs equal; a null key is its own group — into_groups.rs:25-58,142-191).
 * maintain_order != 0: groups ordered by first occurrence (hashing.rs:41-63); else unspecified.
 * out_key = key taken at each group's first row (group_by/mod.rs:258-266).
 * Output dtypes: SUM keeps the dtype (Int8/16,UInt8/16 -> Int64), ints wrap; MEAN -> FLOAT64
 * (FLOAT32 stays); MIN/MAX keep dtype; COUNT/LEN -> UINT32.  All-null group: SUM 0, MEAN/MIN/MAX null.
 * FIRST/LAST keep dtype; VAR/STD -> FLOAT64 (FLOAT32 stays).  A call that asks for any of FIRST/LAST/VAR/STD is evaluated
 * whole over GroupsIdx (groups then come in first-occurrence order), like bl_set_deterministic. */
bl_status bl_groupby_agg(const bl_column* key_chunks, int32_t n_key_chunks, const bl_agg* aggs, int32_t n_aggs,
                         int32_t maintain_order, int32_t out_location, bl_column* out_key, bl_column* out_aggs);

/* Several key columns (DataFrame::group_by_with_series routes them through row encoding,
 * polars-core/src/frame/group_by/mod.rs:88-94; polars-row/src/fixed/numeric.rs:100-145): numeric columns of the
 * dtypes above plus Int8/16, UInt8/16, one chunk each.  Equality per column as for a single key (null == null,
 * -0 == +0, NaN == NaN).  out_keys[i] = keys[i] taken at each group's first row; the rest as bl_groupby_agg. */
bl_status bl_groupby_agg_keys(const bl_column* keys, int32_t n_keys, const bl_agg* aggs, int32_t n_aggs,
                              int32_t maintain_order, int32_t out_location, bl_column* out_keys, bl_column* out_aggs);

/* Group tuples: the reference's GroupsIdx{first, all} (polars-core/src/frame/group_by/position.rs:16-22)
 * as built by group_by_threaded_slice with sorted = true (hashing.rs:116-167, finish_group_order :41-63),
 * for aggregations outside the fused set above.  Groups come in first-occurrence order; group g owns
 * out_all[out_offsets[g] .. out_offsets[g+1]) (row indices ascending), out_first[g] = its first row.
 * All three outputs are BL_UINT32 (IdxSize); out_offsets has n_groups + 1 entries.  Null key = own group. */
bl_status bl_group_tuples(const bl_column* key_chunks, int32_t n_key_chunks, int32_t out_location,
                          bl_column* out_first, bl_column* out_offsets, bl_column* out_all);

/* ---- string / binary keys: device-side dictionary encoding ------------------------ */
/* Arrow LargeUtf8 / LargeBinary layout (polars-arrow/src/array/binary/mod.rs): value i = data[offsets[offset+i] ..
 * offsets[offset+i+1]); validity bit (offset + i).  Polars hands strings out as view arrays (plugin.rs:165-166); the glue
 * casts them with polars_compute::cast::utf8view_to_utf8::<i64> (crates/polars-compute/src/cast/binview_to.rs:56) first. */
typedef struct bl_string_column {
    int32_t location;         /* BL_HOST | BL_DEVICE */
    int32_t reserved;
    int64_t length;           /* logical number of rows */
    int64_t offset;           /* logical element offset into offsets and validity */
    int64_t null_count;       /* -1 = unknown */
    const int64_t* offsets;   /* offset + length + 1 entries */
    const uint8_t* data;
    const uint8_t* validity;  /* NULL = no nulls */
    void* owner;              /* NULL = caller-owned; else release with bl_string_column_free */
} bl_string_column;

/* BinaryChunked::group_tuples (polars-core/src/frame/group_by/into_groups.rs:215-251) groups rows by their BYTES (hash of
 * the bytes + equality, nulls = own group).  bl_string_encode materialises that relation as a UInt32 code column:
 * out_codes[i] = index of the first row holding the same bytes as row i (null rows: null code).  The codes are an ordinary
 * key for bl_groupby_agg / bl_hash_join / bl_group_tuples (for a join, encode the concatenation of both sides' chunks and
 * split the codes), and the key values a group_by returns are the gather indices of the group keys (bl_string_gather).
 * *n_distinct (optional) = number of distinct non-null values.  Exact: equal codes <=> equal bytes (verified on the device). */
bl_status bl_string_encode(const bl_string_column* chunks, int32_t n_chunks, int32_t out_location, bl_column* out_codes, int64_t* n_distinct);
/* out[i] = the string at row idx[i] of the (concatenated) chunks; a null index, BL_IDX_NULL or a null source row gives a
 * null.  idx: BL_UINT32.  Out-of-range indices: BL_ERR_BOUNDS. */
bl_status bl_string_gather(const bl_string_column* chunks, int32_t n_chunks, const bl_column* idx, int32_t out_location, bl_string_column* out);
/* group_by / join on a string key in ONE call: encode -> bl_groupby_agg on the codes -> gather of the group keys
 * (out_key: LargeUtf8, one value per group, null = the null group), resp. encode both sides together -> bl_hash_join on the
 * codes (row-index tuples as bl_hash_join; with nulls_equal a null matches a null, single_keys_dispatch.rs:20-60). */
bl_status bl_groupby_agg_strings(const bl_string_column* key_chunks, int32_t n_key_chunks, const bl_agg* aggs, int32_t n_aggs, int32_t maintain_order,
                                 int32_t out_location, bl_string_column* out_key, bl_column* out_aggs);
bl_status bl_hash_join_strings(const bl_string_column* left_chunks, int32_t n_left_chunks, const bl_string_column* right_chunks, int32_t n_right_chunks, int32_t how,
                               int32_t nulls_equal, int32_t maintain_order, int32_t out_location, bl_column* out_left_idx, bl_column* out_right_idx);
/* copy / move a string column between host and device (concatenates chunks) */
bl_status bl_string_column_to(const bl_string_column* chunks, int32_t n_chunks, int32_t location, bl_string_column* out);
void bl_string_column_free(bl_string_column* col);

/* ---- K7/K8: hash join on one numeric key ----------------------------------------------- */
/* (build_tables single_keys.rs:16-167, probe_inner single_keys_inner.rs:11-149,
 *  hash_join_tuples_left single_keys_left.rs:106-195) */
enum { BL_JOIN_INNER = 0, BL_JOIN_LEFT = 1, BL_JOIN_SEMI = 2, BL_JOIN_ANTI = 3, BL_JOIN_FULL = 4 };
enum { BL_ORDER_NONE = 0, BL_ORDER_LEFT = 1, BL_ORDER_LEFT_RIGHT = 2, BL_ORDER_RIGHT = 3, BL_ORDER_RIGHT_LEFT = 4 };
/* Returns the join tuples as two UINT32 columns.  BL_ORDER_NONE reproduces the in-memory engine's
 * order (probe = longer relation, tie -> right probes; probe-row order; matches ascending build
 * idx — hash_join/mod.rs:41-50).  Null keys match only if nulls_equal.  Left join: unmatched
 * right idx = BL_IDX_NULL (and a null slot).  BL_JOIN_SEMI / BL_JOIN_ANTI (single_keys_semi_anti.rs:41-140):
 * out_left_idx = the left rows, in row order, with / without a match; out_right_idx is an empty column.
 * BL_JOIN_FULL (hash_join_tuples_outer, single_keys_outer.rs:100-260): the longer relation probes; first its left-join
 * tuples in probe order, then the build rows no probe key matched with a null on the probe side (ascending build row;
 * the reference drains its hash tables there, order unspecified).  Both outputs are nullable; BL_ORDER_NONE only. */
bl_status bl_hash_join(const bl_column* left_key, int32_t n_left_chunks, const bl_column* right_key, int32_t n_right_chunks,
                       int32_t how, int32_t nulls_equal, int32_t maintain_order, int32_t out_location,
                       bl_column* out_left_idx, bl_column* out_right_idx);

/* Several key columns per side (prepare_keys_multiple, polars-ops/src/frame/join/mod.rs:658-678: the reference row-encodes
 * the key columns of each side and runs the single-key machinery on the encoded rows).  left_keys[i] pairs with
 * right_keys[i] (same dtype; any numeric dtype incl. 8/16-bit; one chunk each).  nulls_equal == 0: a null in ANY key
 * column makes the row's key null (it matches nothing); != 0: nulls are part of the key and match each other per column.
 * Everything else as bl_hash_join. */
bl_status bl_hash_join_keys(const bl_column* left_keys, const bl_column* right_keys, int32_t n_keys, int32_t how, int32_t nulls_equal,
                            int32_t maintain_order, int32_t out_location, bl_column* out_left_idx, bl_column* out_right_idx);

/* Join + materialisation (_finish_join, polars-ops/src/frame/join/general.rs:17-49; JoinExec,
 * polars-mem-engine/src/executors/join.rs:39-120): bl_hash_join on the key columns followed by one K4
 * gather per side, the tuples never leaving the device.  out_left_cols[i] = left_cols[i] taken at the left
 * idx, out_right_cols[j] = right_cols[j] taken at the right idx (left join: unmatched rows are null).
 * Column naming (the `_right` suffix, dropping the right key) is the caller's metadata.  One chunk per column. */
bl_status bl_join(const bl_column* left_key, const bl_column* right_key,
                  const bl_column* left_cols, int32_t n_left_cols, const bl_column* right_cols, int32_t n_right_cols,
                  int32_t how, int32_t nulls_equal, int32_t maintain_order, int32_t out_location,
                  bl_column* out_left_cols, bl_column* out_right_cols);

/* ---- K6: radix hash partition (multi-GPU exchange step) --------------------------------- */
/* partition id = hash_to_partition(dirty_hash(key), n_partitions)
 *              = ((key * 0x55fbfd6bfc5458e9 mod 2^64) * n_partitions) >> 64   (hashing.rs:62-69,132-142),
 * null keys -> partition 0 (hashing.rs:113-115,183-187).  Rows are scattered so that partition p
 * occupies [offsets[p], offsets[p+1]) of every output column (the same permutation for every column; the row order
 * inside a partition is unspecified).
 * offsets: caller array of n_partitions+1 int64 (host). */
bl_status bl_hash_partition(const bl_column* key, const bl_column* payload, int32_t n_payload, int32_t n_partitions,
                            int32_t out_location, bl_column* out_key, bl_column* out_payload, int64_t* offsets);

/* ---- streaming group_by state (device-resident; used for chunked H2D overlap and multi-GPU) */
typedef struct bl_groupby bl_groupby;
/* key_dtype / value dtypes fix the plan; expected_groups <= 0 lets the library estimate.
 * track_first != 0 records each group's first row index (one extra 32-bit atomic per row); it is
 * required for bl_groupby_finish(maintain_order != 0). */
/* value_nullable[i] == 0 promises that aggregation i's column never carries nulls (saves its null
 * counter: smaller table entries); NULL = every column may be nullable. */
bl_status bl_groupby_create(int32_t key_dtype, const int32_t* agg_kinds, const int32_t* value_dtypes, const int32_t* value_nullable,
                            int32_t n_aggs, int64_t expected_groups, int32_t track_first, bl_groupby** out);
/* Accumulate one batch: key + one value column per agg (values[i] ignored for LEN).  Columns may
 * be BL_HOST or BL_DEVICE.  row_base = global index of the batch's first row. */
bl_status bl_groupby_consume(bl_groupby* g, const bl_column* key, const bl_column* values, int64_t row_base);
/* Partial-aggregate exchange for the partitioned multi-GPU plan:
 * export the table as dense rows of `*row_words` 64-bit words, scattered by key partition
 * (device memory, library-owned: free with bl_dev_free); offsets: n_partitions+1 int64 (host). */
bl_status bl_groupby_export_partials(bl_groupby* g, int32_t n_partitions, void** out_rows_dev, int32_t* row_words, int64_t* offsets);
/* Merge partial rows (from any rank) into this state. */
bl_status bl_groupby_merge_partials(bl_groupby* g, const void* rows_dev, int64_t n_rows);
/* Same for several regions (one per source rank) in one launch: rows_dev[i] holds n_rows[i] rows. */
bl_status bl_groupby_merge_partial_regions(bl_groupby* g, const void* const* rows_dev, const int64_t* n_rows, int32_t n_regions);
bl_status bl_groupby_finish(bl_groupby* g, int32_t maintain_order, int32_t out_location, bl_column* out_key, bl_column* out_aggs);
void bl_groupby_reset(bl_groupby* g);
void bl_groupby_destroy(bl_groupby* g);

/* ---- peer windows: fused partition + exchange over NVLink (one process per GPU) ----------- */
/* A window is plain device memory (cudaMalloc) exported with CUDA IPC so that the partition
 * kernels of the OTHER ranks can store into it directly (P2P stores over NVLink / NVSwitch):
 * partition and transfer are one kernel, no staging buffer, no NCCL on the data path.
 * Layout of a window used by bl_groupby_export_partials_p2p: n_ranks regions of
 * `rows_per_src` rows x row_words 64-bit words; region s is written only by rank s. */
typedef struct bl_window bl_window;
#define BL_IPC_HANDLE_BYTES 64
bl_status bl_window_create(size_t bytes, bl_window** out, void* ipc_handle_out /* BL_IPC_HANDLE_BYTES */);
void* bl_window_ptr(bl_window* w);
void bl_window_destroy(bl_window* w);
/* Maps a peer rank's window into this process (cudaIpcOpenMemHandle). */
bl_status bl_window_open(const void* ipc_handle, void** peer_ptr_out);
void bl_window_close(void* peer_ptr);
/* Fused K6 + exchange for the partitioned group_by: scatters this rank's partial-aggregate rows by
 * key partition straight into region `my_rank` of the destination rank's window.
 * windows[p] = device pointer of rank p's window as mapped in THIS process (own window for
 * p == my_rank).  sent_rows[p] (host) = rows written to rank p; the caller exchanges these counts and
 * merges region s of its own window with bl_groupby_merge_partials.  Returns after the kernel and
 * its peer stores have completed. */
bl_status bl_groupby_export_partials_p2p(bl_groupby* g, int32_t n_ranks, int32_t my_rank, void* const* windows, int64_t rows_per_src,
                                         int32_t* row_words, int64_t* sent_rows);

/* Exchange without host round trips.  window_halves[p] = base of the half of rank p's window used by this step (as
 * mapped in THIS process); a half is BL_WINDOW_HEADER_BYTES of header followed by n_ranks regions of rows_per_src
 * rows.  The export kernel stores the rows AND, from its last thread block, the per-destination row counts plus an
 * epoch flag into the destination headers (release, system scope); nothing is read back and the call returns as
 * soon as the kernel is queued.  `epoch` must grow from step to step (flags are never reset); windows must be
 * zero-initialised (bl_window_create does). */
#define BL_WINDOW_HEADER_BYTES 1024
bl_status bl_groupby_export_partials_p2p_async(bl_groupby* g, int32_t n_ranks, int32_t my_rank, void* const* window_halves, int64_t rows_per_src,
                                               uint64_t epoch, int32_t* row_words);
/* The owner's side: queues a kernel that waits (acquire, system scope, bounded) until every source rank has published
 * `epoch` in own_half's header, then merges the n_ranks regions into this state using the published counts.  Errors
 * (peer region overflow, peer timeout, table overflow) surface at bl_groupby_finish / bl_groupby_status. */
bl_status bl_groupby_merge_window_async(bl_groupby* g, const void* own_half, int32_t n_ranks, int64_t rows_per_src, uint64_t epoch);
/* on != 0: bl_groupby_consume no longer synchronises to check for table overflow; the check happens at
 * bl_groupby_finish (same state) or bl_groupby_status. */
void bl_groupby_defer_status(bl_groupby* g, int32_t on);
/* Synchronises and reports the state's device status word: 0 ok, 1 table overflow, 2 peer region overflow, 3 peer timeout. */
bl_status bl_groupby_status(bl_groupby* g, int32_t* status_out);
/* Sampled cardinality estimate of the consumed batches (0 before the first consume; no synchronisation). */
int64_t bl_groupby_estimated_groups(bl_groupby* g);

/* One rank's whole step of the partitioned group_by in ONE call (what polars_b200/dist.py composes from the pieces above):
 * local pre-aggregation of (key, aggs) -> bl_groupby_export_partials_p2p_async into the peers' window halves -> merge of this
 * rank's own half -> finish.  The outputs are the groups this rank owns (hash_to_partition(dirty_hash(key), n_ranks) ==
 * my_rank).  peer_halves[p] / own_half / rows_per_src / epoch as for the two calls it fuses; value columns one chunk each. */
bl_status bl_groupby_agg_partitioned(const bl_column* key, const bl_agg* aggs, int32_t n_aggs, int32_t n_ranks, int32_t my_rank,
                                     void* const* peer_halves, const void* own_half, int64_t rows_per_src, uint64_t epoch,
                                     int64_t expected_groups, int32_t out_location, bl_column* out_key, bl_column* out_aggs);

/* ---- profiling (CUDA events on the library stream) -------------------------------------- */
/* enable != 0: every kernel launch is bracketed by events; totals accumulate per kernel name. */
void bl_profile_enable(int32_t enable);
void bl_profile_reset(void);
/* Writes a JSON object {"kernel": {"launches": n, "ms": t}, ...} into buf; returns bytes needed. */
int64_t bl_profile_json(char* buf, int64_t cap);
/* Kernel launches issued by this library since bl_profile_reset (counted even when timing is off). */
int64_t bl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* POLARS_B200_H */
