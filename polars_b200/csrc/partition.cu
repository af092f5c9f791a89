// partition.cu — K6: radix hash partition of (key, payload...) rows into P contiguous regions —
// the device half of the multi-GPU exchange (the other half is one all-to-all over NVLink).
//
// Reference precedent: the in-memory radix exchange of build_tables — count -> cumulative offsets
// -> scatter (crates/polars-ops/src/frame/join/hash_join/single_keys.rs:52-121) with
// partition = hash_to_partition(dirty_hash(key), P) (crates/polars-utils/src/hashing.rs:62-69,132-142;
// null -> partition 0, :113-115).  The partition function is restated bit-exactly so per-partition
// counts can be checked as integers against the oracle.
//
// B200 design: two streaming passes.  Pass 1: per-CTA shared-memory histogram (32-bit smem atomics),
// one global add per (CTA, partition).  Pass 2: each CTA re-reads its tile, reserves one contiguous
// range per partition with a single global atomicAdd, and its threads take slots inside the range
// from shared-memory cursors — so global atomics are O(CTAs * P), not O(rows) — and stages every column through
// shared memory in partition order so that the stores are coalesced runs.  Row order inside a partition is
// unspecified.  Algorithmic bytes: 2 * row_bytes per row (+8 for the key re-read).
#include "common.cuh"
#include "dev_utils.cuh"

namespace plb {

constexpr int P_MAX = 64;
constexpr int P_TILE = 2048;
constexpr int P_MAX_COLS = 9;   // key + 8 payload columns
struct PartCol { const void* in; void* out; int elem; int pad; };
struct PartArgs { PartCol c[P_MAX_COLS]; int ncols; };

__device__ __forceinline__ int part_of(const void* keys, const uint32_t* valid, int dtype, int64_t row, int P) {
    if (valid != nullptr && !bit_get(valid, row)) return 0;
    uint64_t k;
    switch (dtype) {
        case BL_INT64: case BL_UINT64: k = reinterpret_cast<const uint64_t*>(keys)[row]; break;
        case BL_FLOAT64: k = canonical_f64_bits(reinterpret_cast<const double*>(keys)[row]); break;
        case BL_FLOAT32: k = canonical_f32_bits(reinterpret_cast<const float*>(keys)[row]); break;
        default: k = (uint64_t)reinterpret_cast<const uint32_t*>(keys)[row]; break;
    }
    return (int)hash_to_partition(dirty_hash(k), (uint32_t)P);
}

__global__ void __launch_bounds__(256) k_part_count(const void* __restrict__ keys, const uint32_t* __restrict__ valid, int dtype, int64_t n, int P, unsigned long long* __restrict__ counts) {
    __shared__ unsigned hist[P_MAX];
    if (threadIdx.x < P_MAX) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[part_of(keys, valid, dtype, r, P)], 1u);
    __syncthreads();
    if (threadIdx.x < P && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

// Tile sort in shared memory, then coalesced run stores: the 2048 rows of a tile are ranked per partition (shared-memory
// atomics), every column is staged into shared memory in partition order and written out so that consecutive threads
// store consecutive elements of a run (round 1 stored every row straight to its slot: 8-byte scattered stores, 0.40 of
// the copy peak).  One global atomic per (tile, partition).
__global__ void __launch_bounds__(256) k_part_scatter(PartArgs a, const uint32_t* __restrict__ valid, uint32_t* __restrict__ out_valid, int dtype, int64_t n, int P,
                                                      const unsigned long long* __restrict__ part_off, unsigned long long* __restrict__ cursor) {
    __shared__ unsigned hist[P_MAX], start[P_MAX];
    __shared__ unsigned long long base[P_MAX];
    __shared__ uint64_t stage[P_TILE];
    __shared__ uint8_t sp[P_TILE];                 // partition of every sorted slot
    __shared__ uint8_t sv[P_TILE];                 // key validity of every sorted slot (nullable keys)
    constexpr int RPT = P_TILE / 256;
    const int64_t ntiles = (n + P_TILE - 1) / P_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (threadIdx.x < P_MAX) hist[threadIdx.x] = 0;
        __syncthreads();
        int p[RPT]; unsigned local[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int64_t r = t * P_TILE + k * 256 + threadIdx.x;
            p[k] = -1; local[k] = 0;
            if (r < n) { p[k] = part_of(a.c[0].in, valid, dtype, r, P); local[k] = atomicAdd(&hist[p[k]], 1u); }
        }
        __syncthreads();
        if (threadIdx.x == 0) { unsigned run = 0; for (int q = 0; q < P; q++) { start[q] = run; run += hist[q]; } }
        if (threadIdx.x < P && hist[threadIdx.x]) base[threadIdx.x] = part_off[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
        __syncthreads();
        const int rows = (int)min((int64_t)P_TILE, n - t * P_TILE);
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            if (p[k] < 0) continue;
            const unsigned pos = start[p[k]] + local[k];
            sp[pos] = (uint8_t)p[k];
            if (out_valid != nullptr) sv[pos] = bit_get(valid, t * P_TILE + k * 256 + threadIdx.x) ? 1 : 0;
        }
        for (int c = 0; c < a.ncols; c++) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                if (p[k] < 0) continue;
                const int64_t r = t * P_TILE + k * 256 + threadIdx.x;
                stage[start[p[k]] + local[k]] = a.c[c].elem == 8 ? __ldcs(reinterpret_cast<const unsigned long long*>(a.c[c].in) + r) : (uint64_t)__ldcs(reinterpret_cast<const unsigned int*>(a.c[c].in) + r);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < rows; i += 256) {
                const unsigned q = sp[i];
                const uint64_t dst = base[q] + (i - start[q]);
                if (a.c[c].elem == 8) reinterpret_cast<uint64_t*>(a.c[c].out)[dst] = stage[i];
                else reinterpret_cast<uint32_t*>(a.c[c].out)[dst] = (uint32_t)stage[i];
                if (c == 0 && out_valid != nullptr && sv[i]) atomicOr(&out_valid[dst >> 5], 1u << (dst & 31));
            }
            __syncthreads();
        }
    }
}

void op_hash_partition(const DevCol& key, const std::vector<DevCol>& payload, int P, DevCol& out_key, std::vector<DevCol>& out_payload, int64_t* offsets_host) {
    PLB_REQUIRE(P >= 1 && P <= P_MAX, BL_ERR_INVALID, "hash_partition: 1..64 partitions");
    PLB_REQUIRE(payload.size() + 1 <= P_MAX_COLS, BL_ERR_UNSUPPORTED, "hash_partition: at most 8 payload columns");
    const int dt = key.dtype;
    PLB_REQUIRE(dt == BL_INT64 || dt == BL_UINT64 || dt == BL_INT32 || dt == BL_UINT32 || dt == BL_FLOAT64 || dt == BL_FLOAT32, BL_ERR_UNSUPPORTED, "hash_partition: key dtype outside the hot path");
    const int64_t n = key.len;
    for (auto& c : payload) {
        PLB_REQUIRE(c.len == n, BL_ERR_INVALID, "hash_partition: payload length differs from key length");
        PLB_REQUIRE(c.validity == nullptr, BL_ERR_UNSUPPORTED, "hash_partition: nullable payload columns are outside the hot path");
        PLB_REQUIRE(dtype_size(c.dtype) == 4 || dtype_size(c.dtype) == 8, BL_ERR_UNSUPPORTED, "hash_partition: payload dtype outside the hot path");
    }
    Context& cx = ctx();
    DevPtr counts = dev_alloc(8 * P_MAX), cursor = dev_alloc(8 * P_MAX), off = dev_alloc(8 * P_MAX);
    dev_memset(counts->p, 0, 8 * P_MAX); dev_memset(cursor->p, 0, 8 * P_MAX);
    if (n > 0) PLB_LAUNCH("k6_part_count", k_part_count, grid_for(n, 256), 256, 0, key.v(), key.vm(), dt, n, P, as<unsigned long long>(counts));
    unsigned long long h[P_MAX], ho[P_MAX + 1];
    PLB_CUDA(cudaMemcpyAsync(h, counts->p, 8 * P, cudaMemcpyDeviceToHost, cx.stream));
    PLB_CUDA(cudaStreamSynchronize(cx.stream));
    ho[0] = 0;
    for (int p = 0; p < P; p++) ho[p + 1] = ho[p] + h[p];
    for (int p = 0; p <= P; p++) offsets_host[p] = (int64_t)ho[p];
    PLB_CUDA(cudaMemcpyAsync(off->p, ho, 8 * P, cudaMemcpyHostToDevice, cx.stream));
    out_key = make_col(dt, n, key.validity != nullptr);
    if (out_key.validity) dev_memset(out_key.validity->p, 0, out_key.validity->bytes);
    out_payload.clear();
    for (auto& c : payload) out_payload.push_back(make_col(c.dtype, n, false));
    if (n > 0) {
        PartArgs a; memset(&a, 0, sizeof a);
        a.ncols = 1 + (int)payload.size();
        a.c[0].in = key.v(); a.c[0].out = out_key.values->p; a.c[0].elem = dtype_size(dt);
        for (size_t i = 0; i < payload.size(); i++) { a.c[i + 1].in = payload[i].v(); a.c[i + 1].out = out_payload[i].values->p; a.c[i + 1].elem = dtype_size(payload[i].dtype); }
        const int64_t ntiles = (n + P_TILE - 1) / P_TILE;
        PLB_LAUNCH("k6_part_scatter", k_part_scatter, (int)std::min<int64_t>(ntiles, (int64_t)cx.sm_count * 8), 256, 0, a, key.vm(), as<uint32_t>(out_key.validity), dt, n, P,
                   as<unsigned long long>(off), as<unsigned long long>(cursor));
    }
    PLB_CUDA(cudaStreamSynchronize(cx.stream));
}

}  // namespace plb
