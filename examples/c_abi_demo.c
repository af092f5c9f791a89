/*
 * c_abi_demo.c — the boundary used from plain C (no Python, no C++, no torch):
 *   df.filter(x > 0).group_by(key).agg(sum(x), mean(x), len())      BASELINE.json configs[0] (C1)
 *   left.join(right, on = k, how = "inner") with both sides materialised
 * A Rust `extern "C"` binding makes exactly these calls with pointers taken from
 * PrimitiveArray::values() / validity().
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lpolars_b200/_lib -lpolars_b200 -Wl,-rpath,$PWD/polars_b200/_lib -o c_abi_demo
 *
 * tests/test_cabi_cpu.py compiles and links this file on CPU (header = valid C99, every symbol resolves);
 * running it needs a GPU: without one the first call reports BL_ERR_CUDA and the program exits 2.
 */
#include <stdio.h>
#include <stdlib.h>

#include "polars_b200.h"

#define CHECK(call)                                                                       \
    do { bl_status st_ = (call);                                                          \
         if (st_ != BL_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, bl_last_error()); return st_ == BL_ERR_CUDA ? 2 : 1; } } while (0)

static bl_column host_col(int32_t dtype, const void* values, int64_t n) {
    bl_column c;
    c.dtype = dtype; c.location = BL_HOST; c.length = n; c.offset = 0; c.null_count = 0;
    c.values = values; c.validity = NULL; c.owner = NULL;
    return c;
}

int main(void) {
    enum { N = 1000, NB = 10 };
    static int64_t key[N], x[N], rkey[NB];
    static double payload[NB];
    int64_t i;
    for (i = 0; i < N; i++) { key[i] = i % 7; x[i] = (i % 11) - 5; }
    for (i = 0; i < NB; i++) { rkey[i] = NB - 1 - i; payload[i] = 0.5 * (double)i; }

    CHECK(bl_init(-1));

    /* filter(x > 0): all columns compacted by one fused compare + compaction (K2 + K3), kept on the device */
    bl_column cols[2], zero, filtered[2];
    int64_t zero_v = 0;
    cols[0] = host_col(BL_INT64, key, N);
    cols[1] = host_col(BL_INT64, x, N);
    zero = host_col(BL_INT64, &zero_v, 1);
    CHECK(bl_filter_cmp(cols, 2, /*pred_col=*/1, BL_CMP_GT, &zero, BL_DEVICE, filtered));

    /* group_by(key).agg(sum(x), mean(x), len()), groups in first-occurrence order (K5) */
    bl_agg aggs[3];
    bl_column out_key, out_aggs[3];
    aggs[0].kind = BL_AGG_SUM;  aggs[0].n_chunks = 1; aggs[0].values = &filtered[1];
    aggs[1].kind = BL_AGG_MEAN; aggs[1].n_chunks = 1; aggs[1].values = &filtered[1];
    aggs[2].kind = BL_AGG_LEN;  aggs[2].n_chunks = 0; aggs[2].values = NULL;
    CHECK(bl_groupby_agg(&filtered[0], 1, aggs, 3, /*maintain_order=*/1, BL_HOST, &out_key, out_aggs));
    printf("groups: %lld\n", (long long)out_key.length);
    for (i = 0; i < out_key.length; i++)
        printf("  key %lld  sum %lld  mean %.4f  len %u\n", (long long)((const int64_t*)out_key.values)[i], (long long)((const int64_t*)out_aggs[0].values)[i],
               ((const double*)out_aggs[1].values)[i], (unsigned)((const uint32_t*)out_aggs[2].values)[i]);

    /* inner join on key with the right payload materialised (K7 + K8 + K4); tuples never leave the device */
    bl_column lk = host_col(BL_INT64, key, N), rk = host_col(BL_INT64, rkey, NB), rp = host_col(BL_FLOAT64, payload, NB);
    bl_column out_left[1], out_right[1];
    CHECK(bl_join(&lk, &rk, &lk, 1, &rp, 1, BL_JOIN_INNER, /*nulls_equal=*/0, BL_ORDER_NONE, BL_HOST, out_left, out_right));
    printf("join rows: %lld (first: key %lld payload %.1f)\n", (long long)out_left[0].length,
           out_left[0].length ? (long long)((const int64_t*)out_left[0].values)[0] : -1LL, out_left[0].length ? ((const double*)out_right[0].values)[0] : 0.0);

    bl_column_free(&filtered[0]); bl_column_free(&filtered[1]);
    bl_column_free(&out_key);
    for (i = 0; i < 3; i++) bl_column_free(&out_aggs[i]);
    bl_column_free(&out_left[0]); bl_column_free(&out_right[0]);
    bl_shutdown();
    return 0;
}
