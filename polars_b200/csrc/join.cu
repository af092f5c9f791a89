// join.cu — K7 (hash-join build) and K8 (probe + deterministic tuple emission).
//
// Reference path being replaced (paths relative to /root/reference/crates):
//   build_tables            polars-ops/src/frame/join/hash_join/single_keys.rs:16-167
//   probe_inner / hash_join_tuples_inner   .../single_keys_inner.rs:11-149
//   hash_join_tuples_left   .../single_keys_left.rs:106-195
//   which side builds       .../hash_join/mod.rs:41-50 (probe = longer relation, tie -> right probes)
//   maintain_order sort     polars-ops/src/frame/join/mod.rs:577-642
// The reference radix-partitions the build side over threads and keeps a hashbrown map
// key -> ascending row-index vector per partition; probing walks the probe side in row order and
// emits (probe idx, build idx) for every build idx in ascending order.
//
// B200 design.  One open-addressing table in HBM of 16-byte entries {key, val, cnt} (one 128-bit
// load per probe step; slot = mulhi(key * RANDOM_ODD, cap) — the reference's own
// hash_to_partition — so the capacity is exactly 2x the build rows, no power-of-two padding).
//   build:  claim the key's entry (CAS), cnt += 1, val = min(val, row).  Unique build keys (the
//           primary-key case) need nothing else: val is the build row.  With duplicates the rows
//           are stably sorted by entry and an exclusive scan of cnt turns val into a CSR offset,
//           so every entry owns an ascending row list — the reference's IdxVec.
//   probe:  pass 1 looks every probe row up once and stores its match handle (4 B/row) plus
//           per-tile match counts; an exclusive scan of the tile counts gives every tile its
//           output offset; pass 2 expands the handles into (probe idx, build idx) tuples.  The
//           output is therefore in exact reference order with no atomics or spin-waits.
// Algorithmic bytes: build 8 B read + 16 B table write per build row; probe
// 8 B key read + 8 B tuple write per match.  Bound: random 32-byte sector reads of the table
// (HBM when the table exceeds L2, L2 otherwise).
#include <cstdlib>

#include "common.cuh"
#include "dev_utils.cuh"

namespace plb {

constexpr uint64_t J_EMPTY = 0x8000000000000000ULL;
constexpr uint32_t J_NONE = 0xFFFFFFFFu;
constexpr int J_TILE = 2048;

struct JoinTableDev { uint4* entries; uint64_t cap; };   // entries[cap] = null-key entry, [cap+1] = J_EMPTY-key entry

__device__ __forceinline__ uint64_t j_load_key(const void* keys, int dtype, int64_t row) {
    switch (dtype) {
        case BL_INT64: case BL_UINT64: return reinterpret_cast<const uint64_t*>(keys)[row];
        case BL_FLOAT64: return canonical_f64_bits(reinterpret_cast<const double*>(keys)[row]);   // NaN joins NaN (single_keys_dispatch.rs:316-322)
        case BL_FLOAT32: return canonical_f32_bits(reinterpret_cast<const float*>(keys)[row]);
        default: return (uint64_t)reinterpret_cast<const uint32_t*>(keys)[row];
    }
}

__global__ void k_join_init(uint4* entries, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        entries[i] = make_uint4((uint32_t)J_EMPTY, (uint32_t)(J_EMPTY >> 32), J_NONE, 0u);
}

// ---------------------------------------------------------------------------- K7 build
__global__ void __launch_bounds__(256) k_join_build(JoinTableDev T, const void* __restrict__ keys, const uint32_t* __restrict__ valid, int key_dtype, int64_t n,
                                                    int nulls_equal, uint32_t* __restrict__ slot_of_row, int* __restrict__ has_dups) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const bool v = valid == nullptr || bit_get(valid, r);
        uint64_t slot;
        if (!v) {
            if (!nulls_equal) { slot_of_row[r] = J_NONE; continue; }     // null keys are not inserted (single_keys.rs:41,148)
            slot = T.cap;
            atomicCAS(reinterpret_cast<unsigned long long*>(&T.entries[slot]), (unsigned long long)J_EMPTY, 0ull);
        } else {
            const uint64_t key = j_load_key(keys, key_dtype, r);
            if (key == J_EMPTY) { slot = T.cap + 1; atomicCAS(reinterpret_cast<unsigned long long*>(&T.entries[slot]), (unsigned long long)J_EMPTY, 1ull); }
            else {
                slot = __umul64hi(dirty_hash(key), T.cap);
                while (true) {
                    unsigned long long* kp = reinterpret_cast<unsigned long long*>(&T.entries[slot]);
                    unsigned long long k = __ldcg(kp);
                    if (k == key) break;
                    if (k == J_EMPTY) { unsigned long long old = atomicCAS(kp, (unsigned long long)J_EMPTY, (unsigned long long)key); if (old == J_EMPTY || old == key) break; }
                    if (++slot == T.cap) slot = 0;
                }
            }
        }
        uint32_t* w = reinterpret_cast<uint32_t*>(&T.entries[slot]);
        const uint32_t old = atomicAdd(w + 3, 1u);
        atomicMin(w + 2, (uint32_t)r);
        if (old != 0) *has_dups = 1;
        slot_of_row[r] = (uint32_t)slot;
    }
}

// duplicates: val <- CSR offset (exclusive scan of cnt over the entries, in entry order)
__global__ void __launch_bounds__(256) k_join_tile_sums(const uint4* __restrict__ entries, int64_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t ws[8];
    const int64_t ntiles = (n + J_TILE - 1) / J_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint32_t c = 0;
        for (int k = 0; k < J_TILE / 256; k++) { int64_t i = t * J_TILE + k * 256 + threadIdx.x; if (i < n) c += entries[i].w; }
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane_id() == 0) ws[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t s = 0; for (int w = 0; w < 8; w++) s += ws[w]; sums[t] = s; }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_join_csr_offsets(uint4* __restrict__ entries, int64_t n, const uint64_t* __restrict__ tile_off) {
    __shared__ uint32_t ws[8];
    __shared__ uint32_t carry;
    const int64_t ntiles = (n + J_TILE - 1) / J_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (threadIdx.x == 0) carry = (uint32_t)tile_off[t];
        __syncthreads();
        for (int k = 0; k < J_TILE / 256; k++) {
            const int64_t i = t * J_TILE + k * 256 + threadIdx.x;
            const uint32_t c = i < n ? entries[i].w : 0;
            uint32_t x = c;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane_id() >= (unsigned)o) x += y; }
            if (lane_id() == 31) ws[threadIdx.x >> 5] = x;
            __syncthreads();
            uint32_t wbase = 0;
            for (unsigned w = 0; w < (threadIdx.x >> 5); w++) wbase += ws[w];
            if (i < n) entries[i].z = carry + wbase + x - c;
            __syncthreads();
            if (threadIdx.x == 255) carry += wbase + x;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------- K8 probe, pass 1
// handle[i] = unique mode: build row (J_NONE on miss); CSR mode: entry index (J_NONE on miss).
template <int KEY_ELEM, int KEY_CANON, bool KEY_NULLS>
__global__ void __launch_bounds__(256) k_join_probe(JoinTableDev T, const void* __restrict__ keys, const uint32_t* __restrict__ valid, int64_t n, int nulls_equal, int csr_mode,
                                                    int left_join, uint32_t* __restrict__ handle, unsigned long long* __restrict__ tile_counts) {
    // 2 row pairs (4 rows) per thread and iteration; the first table probe of all 4 rows is issued
    // before any is resolved: the kernel is bound by random-sector latency, so MLP is what counts.
    constexpr int PAIRS = 2, R = 2 * PAIRS;
    const int64_t npairs = (n + 1) >> 1;
    const int64_t rounded = (npairs + 31) / 32 * 32;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p0 < rounded; p0 += gstride * PAIRS) {
        uint64_t key[R]; uint64_t slot[R]; uint4 e[R]; int kind[R];   // kind: 0 hashed, 1 special slot, -1 no lookup (miss / no row)
#pragma unroll
        for (int u = 0; u < PAIRS; u++) {
            const int64_t r0 = 2 * (p0 + u * gstride);
            uint64_t kraw[2] = {0, 0};
            if (r0 + 1 < n) {
                if (KEY_ELEM == 8) { ulonglong2 t = ld_stream_u64x2(reinterpret_cast<const uint64_t*>(keys) + r0); kraw[0] = t.x; kraw[1] = t.y; }
                else { uint2 t = ld_stream_u32x2(reinterpret_cast<const uint32_t*>(keys) + r0); kraw[0] = t.x; kraw[1] = t.y; }
            } else if (r0 < n) kraw[0] = KEY_ELEM == 8 ? reinterpret_cast<const uint64_t*>(keys)[r0] : (uint64_t)reinterpret_cast<const uint32_t*>(keys)[r0];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int r = 2 * u + j;
                const int64_t row = r0 + j;
                kind[r] = -1; slot[r] = 0; key[r] = 0;
                if (row >= n) continue;
                bool v = true;
                if (KEY_NULLS) v = bit_get(valid, row);
                uint64_t k = kraw[j];
                if (KEY_CANON == 1) k = canonical_f64_bits(__longlong_as_double((long long)k));
                if (KEY_CANON == 2) k = canonical_f32_bits(__uint_as_float((uint32_t)k));
                key[r] = k;
                if (!v) { if (nulls_equal) { kind[r] = 1; slot[r] = T.cap; } }
                else if (k == J_EMPTY) { kind[r] = 1; slot[r] = T.cap + 1; }
                else { kind[r] = 0; slot[r] = __umul64hi(dirty_hash(k), T.cap); }
                if (kind[r] >= 0) e[r] = __ldg(&T.entries[slot[r]]);
            }
        }
#pragma unroll
        for (int u = 0; u < PAIRS; u++) {
            const int64_t r0 = 2 * (p0 + u * gstride);
            if (r0 >= 2 * rounded) continue;                       // warp-uniform
            uint32_t h[2] = {J_NONE, J_NONE};
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int r = 2 * u + j;
                if (r0 + j >= n) continue;
                bool hit = false;
                if (kind[r] == 1) hit = e[r].w != 0;
                else if (kind[r] == 0) {
                    while (true) {
                        const uint64_t k = ((uint64_t)e[r].y << 32) | e[r].x;
                        if (k == key[r]) { hit = true; break; }
                        if (k == J_EMPTY) break;
                        if (++slot[r] == T.cap) slot[r] = 0;
                        e[r] = __ldg(&T.entries[slot[r]]);
                    }
                }
                if (hit) { h[j] = csr_mode ? (uint32_t)slot[r] : e[r].z; cnt += csr_mode ? e[r].w : 1u; }
                else if (left_join) cnt += 1u;
            }
            if (r0 + 1 < n) *reinterpret_cast<uint2*>(handle + r0) = make_uint2(h[0], h[1]);
            else if (r0 < n) handle[r0] = h[0];
            // 32 lanes x 2 rows = 64 consecutive rows: always inside one J_TILE
            unsigned long long c = cnt;
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane_id() == 0 && c) atomicAdd(&tile_counts[r0 / J_TILE], c);   // one address per 2048-row tile; the grand total comes from the scan
        }
    }
}

// ---------------------------------------------------------------------------- K8 probe, pass 2 (emit)
// Rows are mapped lane-strided (row = tile + j*256 + tid) so that loads of the handles and — when
// every probe row has <= 1 match, the primary-key case — the tuple stores are fully coalesced.
// Per (iteration j, warp) segment: shuffle scan of the per-row match counts; the 64 segment totals
// of the tile are scanned by one warp; the tile base comes from the tile-count scan.
__global__ void __launch_bounds__(256) k_join_emit(JoinTableDev T, const uint32_t* __restrict__ handle, int64_t n, int csr_mode, int left_join,
                                                   const uint32_t* __restrict__ sorted_rows, const uint64_t* __restrict__ tile_off,
                                                   uint32_t* __restrict__ out_probe, uint32_t* __restrict__ out_build) {
    constexpr int ITERS = J_TILE / 256;           // 8
    __shared__ uint32_t seg[ITERS * 8];           // segment = j * 8 + warp, in row order
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t ntiles = (n + J_TILE - 1) / J_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint32_t h[ITERS], c[ITERS], off[ITERS], lane_excl[ITERS];
#pragma unroll
        for (int j = 0; j < ITERS; j++) {
            const int64_t row = t * J_TILE + j * 256 + threadIdx.x;
            h[j] = row < n ? handle[row] : J_NONE;
            uint32_t ck = 0; off[j] = 0;
            if (row < n) {
                if (h[j] != J_NONE) { if (csr_mode) { uint4 e = __ldg(&T.entries[h[j]]); ck = e.w; off[j] = e.z; } else ck = 1; }
                else if (left_join) ck = 1;
            }
            c[j] = ck;
            uint32_t x = ck;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
            lane_excl[j] = x - ck;
            if (lane == 31) seg[j * 8 + warp] = x;
        }
        __syncthreads();
        if (warp == 0) {      // exclusive scan of the 64 segment totals (2 per lane)
            uint32_t a = seg[2 * lane], b = seg[2 * lane + 1], s2 = a + b, x = s2;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
            seg[2 * lane] = x - s2; seg[2 * lane + 1] = x - s2 + a;
        }
        __syncthreads();
        const uint64_t base = tile_off[t];
#pragma unroll
        for (int j = 0; j < ITERS; j++) {
            if (c[j] == 0) continue;
            uint64_t pos = base + seg[j * 8 + warp] + lane_excl[j];
            const uint32_t pi = (uint32_t)(t * J_TILE + j * 256 + threadIdx.x);
            if (h[j] == J_NONE) { out_probe[pos] = pi; out_build[pos] = J_NONE; }
            else if (!csr_mode) { out_probe[pos] = pi; out_build[pos] = h[j]; }
            else for (uint32_t q = 0; q < c[j]; q++) { out_probe[pos + q] = pi; out_build[pos + q] = sorted_rows[off[j] + q]; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------- dense (direct-address) mode
// When the build keys are integers whose value range is at most a few times the build rows (dense
// surrogate / primary keys) the hash table degenerates to a perfect hash: slot = key - min.  The
// table is then 4 bytes per key value (10^7 keys -> 40 MB, L2-resident on B200) and a probe is ONE
// L2 hit instead of a random HBM sector.  Same outputs as the hashed path; taken only for unique
// build keys (duplicates fall back to the hashed table + CSR lists).
__device__ __forceinline__ uint64_t j_ordered(uint64_t raw, int sign_bits) {
    // order-preserving map to u64: signed types flip the sign bit (32-bit patterns are sign-extended first)
    if (sign_bits == 64) return raw ^ 0x8000000000000000ULL;
    if (sign_bits == 32) return (uint64_t)(int64_t)(int32_t)(uint32_t)raw ^ 0x8000000000000000ULL;
    return raw;
}
__global__ void k_join_minmax(const void* __restrict__ keys, const uint32_t* __restrict__ valid, int elem, int sign_bits, int64_t n, unsigned long long* mm) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (valid != nullptr && !bit_get(valid, r)) continue;
        uint64_t raw = elem == 8 ? reinterpret_cast<const uint64_t*>(keys)[r] : (uint64_t)reinterpret_cast<const uint32_t*>(keys)[r];
        unsigned long long k = j_ordered(raw, sign_bits);
        lo = k < lo ? k : lo; hi = k > hi ? k : hi;
    }
    for (int o = 16; o; o >>= 1) {
        unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo; hi = b > hi ? b : hi;
    }
    if (lane_id() == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}
__global__ void k_fill_u32j(uint32_t* p, uint32_t v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(256) k_join_dense_build(uint32_t* __restrict__ table, const void* __restrict__ keys, const uint32_t* __restrict__ valid, int elem, int sign_bits,
                                                          int64_t n, uint64_t kmin, int* __restrict__ has_dups) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (valid != nullptr && !bit_get(valid, r)) continue;
        uint64_t raw = elem == 8 ? reinterpret_cast<const uint64_t*>(keys)[r] : (uint64_t)reinterpret_cast<const uint32_t*>(keys)[r];
        const uint32_t old = atomicMin(&table[j_ordered(raw, sign_bits) - kmin], (uint32_t)r);
        if (old != J_NONE) *has_dups = 1;
    }
}
template <int KEY_ELEM, bool KEY_NULLS>
__global__ void __launch_bounds__(256) k_join_dense_probe(const uint32_t* __restrict__ table, uint64_t kmin, uint64_t range, int sign_bits, const void* __restrict__ keys,
                                                          const uint32_t* __restrict__ valid, int64_t n, int left_join, uint32_t* __restrict__ handle,
                                                          unsigned long long* __restrict__ tile_counts) {
    constexpr int PAIRS = 2;
    const int64_t npairs = (n + 1) >> 1;
    const int64_t rounded = (npairs + 31) / 32 * 32;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p0 < rounded; p0 += gstride * PAIRS) {
        uint64_t kraw[PAIRS][2];
#pragma unroll
        for (int u = 0; u < PAIRS; u++) {
            const int64_t r0 = 2 * (p0 + u * gstride);
            kraw[u][0] = kraw[u][1] = 0;
            if (r0 + 1 < n) {
                if (KEY_ELEM == 8) { ulonglong2 t = ld_stream_u64x2(reinterpret_cast<const uint64_t*>(keys) + r0); kraw[u][0] = t.x; kraw[u][1] = t.y; }
                else { uint2 t = ld_stream_u32x2(reinterpret_cast<const uint32_t*>(keys) + r0); kraw[u][0] = t.x; kraw[u][1] = t.y; }
            } else if (r0 < n) kraw[u][0] = KEY_ELEM == 8 ? reinterpret_cast<const uint64_t*>(keys)[r0] : (uint64_t)reinterpret_cast<const uint32_t*>(keys)[r0];
        }
        uint32_t h[PAIRS][2];
#pragma unroll
        for (int u = 0; u < PAIRS; u++) {
            const int64_t r0 = 2 * (p0 + u * gstride);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                h[u][j] = J_NONE;
                const int64_t row = r0 + j;
                if (row >= n) continue;
                bool v = true;
                if (KEY_NULLS) v = bit_get(valid, row);
                const uint64_t d = j_ordered(kraw[u][j], sign_bits) - kmin;
                if (v && d < range) h[u][j] = __ldg(&table[d]);
            }
        }
#pragma unroll
        for (int u = 0; u < PAIRS; u++) {
            const int64_t r0 = 2 * (p0 + u * gstride);
            if (r0 >= 2 * rounded) continue;                       // warp-uniform
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < 2; j++) if (r0 + j < n) cnt += (h[u][j] != J_NONE || left_join) ? 1u : 0u;
            if (r0 + 1 < n) *reinterpret_cast<uint2*>(handle + r0) = make_uint2(h[u][0], h[u][1]);
            else if (r0 < n) handle[r0] = h[u][0];
            unsigned long long c = cnt;
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane_id() == 0 && c) atomicAdd(&tile_counts[r0 / J_TILE], c);
        }
    }
}

// ---------------------------------------------------------------------------- K8 fused probe + emit
// Unique build keys (dense table or hashed table without duplicates): every probe row yields at most
// one tuple, so probe and emission fuse into ONE pass with a decoupled look-back scan over 2048-row
// tiles (tiles are handed out by an atomic counter, so every predecessor of a tile is already
// running: the look-back cannot wait on an unscheduled CTA).  HBM traffic: 8 B key in + 8 B tuple out
// per match — exactly the algorithmic 16 B/row, instead of 24 B/row for the two-pass form.
// status[t]: bits 63..62 = 0 empty / 1 tile aggregate / 2 inclusive prefix, bits 61..0 = value.
constexpr unsigned long long LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_VAL = (1ull << 62) - 1;
template <int KEY_ELEM, int KEY_CANON, bool KEY_NULLS, bool DENSE>
__global__ void __launch_bounds__(256) k_join_probe_emit(JoinTableDev T, const uint32_t* __restrict__ dense_table, uint64_t kmin, uint64_t range, int sign_bits,
                                                         const void* __restrict__ keys, const uint32_t* __restrict__ valid, int64_t n, int nulls_equal, int left_join,
                                                         unsigned long long* __restrict__ status, unsigned int* __restrict__ tile_counter, int* __restrict__ error,
                                                         uint32_t* __restrict__ out_probe, uint32_t* __restrict__ out_build) {
    constexpr int ITERS = J_TILE / 256;
    __shared__ uint32_t seg[ITERS * 8];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_prefix;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t ntiles = (n + J_TILE - 1) / J_TILE;
    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(tile_counter, 1u);
        __syncthreads();
        const int64_t t = s_tile;
        if (t >= ntiles) break;
        uint32_t h[ITERS], lane_excl[ITERS]; bool emit[ITERS];
        uint64_t kraw[ITERS];
#pragma unroll
        for (int j = 0; j < ITERS; j++) {        // all key loads of the tile first (coalesced, 8 in flight per thread)
            const int64_t row = t * J_TILE + j * 256 + threadIdx.x;
            kraw[j] = 0;
            if (row < n) kraw[j] = KEY_ELEM == 8 ? __ldcs(reinterpret_cast<const unsigned long long*>(keys) + row) : (uint64_t)__ldcs(reinterpret_cast<const unsigned int*>(keys) + row);
        }
#pragma unroll
        for (int j = 0; j < ITERS; j++) {
            const int64_t row = t * J_TILE + j * 256 + threadIdx.x;
            h[j] = J_NONE;
            if (row < n) {
                bool v = true;
                if (KEY_NULLS) v = bit_get(valid, row);
                if (DENSE) {
                    const uint64_t d = j_ordered(kraw[j], sign_bits) - kmin;
                    if (v && d < range) h[j] = __ldg(&dense_table[d]);
                } else {
                    uint64_t key = kraw[j];
                    if (KEY_CANON == 1) key = canonical_f64_bits(__longlong_as_double((long long)key));
                    if (KEY_CANON == 2) key = canonical_f32_bits(__uint_as_float((uint32_t)key));
                    uint4 e;
                    if (!v) { if (nulls_equal) { e = __ldg(&T.entries[T.cap]); if (e.w) h[j] = e.z; } }
                    else if (key == J_EMPTY) { e = __ldg(&T.entries[T.cap + 1]); if (e.w) h[j] = e.z; }
                    else {
                        uint64_t slot = __umul64hi(dirty_hash(key), T.cap);
                        while (true) {
                            e = __ldg(&T.entries[slot]);
                            const uint64_t k = ((uint64_t)e.y << 32) | e.x;
                            if (k == key) { h[j] = e.z; break; }
                            if (k == J_EMPTY) break;
                            if (++slot == T.cap) slot = 0;
                        }
                    }
                }
            }
            emit[j] = row < n && (h[j] != J_NONE || left_join);
            const uint32_t b = __ballot_sync(0xffffffffu, emit[j]);
            lane_excl[j] = __popc(b & lanemask_lt());
            if (lane == 0) seg[j * 8 + warp] = __popc(b);
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive scan of the 64 segment counts; tile total
            uint32_t a = seg[2 * lane], b2 = seg[2 * lane + 1], s2 = a + b2, x = s2;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
            seg[2 * lane] = x - s2; seg[2 * lane + 1] = x - s2 + a;
            const unsigned long long tile_total = __shfl_sync(0xffffffffu, x, 31);
            // decoupled look-back
            if (lane == 0) atomicExch(&status[t], (t == 0 ? LB_INC : LB_AGG) | tile_total);
            unsigned long long exclusive = 0;
            if (t > 0) {
                int64_t look = t - 1;
                while (true) {
                    const int64_t idx = look - lane;
                    unsigned long long st = LB_INC;          // virtual predecessor before tile 0: inclusive 0
                    int spins = 0;
                    if (idx >= 0) st = *reinterpret_cast<volatile unsigned long long*>(&status[idx]);
                    while (__any_sync(0xffffffffu, (st >> 62) == 0)) {
                        if (idx >= 0 && (st >> 62) == 0) st = *reinterpret_cast<volatile unsigned long long*>(&status[idx]);
                        if (++spins > (1 << 22) || ((spins & 1023) == 0 && *reinterpret_cast<volatile int*>(error))) { *error = 1; break; }   // never hang the device
                    }
                    const unsigned inc = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                    const unsigned upto = inc ? (unsigned)(__ffs(inc) - 1) : 31u;     // nearest inclusive predecessor
                    unsigned long long v = lane <= upto ? (st & LB_VAL) : 0ull;
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    exclusive += v;
                    if (inc || *reinterpret_cast<volatile int*>(error)) break;
                    look -= 32;
                }
                if (lane == 0) atomicExch(&status[t], LB_INC | ((exclusive + tile_total) & LB_VAL));
            }
            if (lane == 0) s_prefix = exclusive;
        }
        __syncthreads();
        const uint64_t base = s_prefix;
#pragma unroll
        for (int j = 0; j < ITERS; j++) {
            if (!emit[j]) continue;
            const uint64_t pos = base + seg[j * 8 + warp] + lane_excl[j];
            out_probe[pos] = (uint32_t)(t * J_TILE + j * 256 + threadIdx.x);
            out_build[pos] = h[j];
        }
        __syncthreads();
    }
}

template <int KEY_ELEM, int KEY_CANON>
static void launch_probe(bool kn, int grid, JoinTableDev T, const DevCol& probe, int nulls_equal, int csr, int left, uint32_t* handle, unsigned long long* tc) {
    if (kn) PLB_LAUNCH("k8_join_probe", (k_join_probe<KEY_ELEM, KEY_CANON, true>), grid, 256, 0, T, probe.v(), probe.vm(), probe.len, nulls_equal, csr, left, handle, tc);
    else PLB_LAUNCH("k8_join_probe", (k_join_probe<KEY_ELEM, KEY_CANON, false>), grid, 256, 0, T, probe.v(), probe.vm(), probe.len, nulls_equal, csr, left, handle, tc);
}

static DevCol idx_col(DevPtr p, int64_t n, int64_t null_count) { DevCol c; c.dtype = BL_UINT32; c.len = n; c.values = p; c.null_count = null_count; return c; }

// ---------------------------------------------------------------------------- semi / anti
// hash_join_tuples_left_semi / _anti (polars-ops/src/frame/join/hash_join/single_keys_semi_anti.rs:41-140):
// the left rows, in row order, that have (semi) / do not have (anti) a key match on the right; a null left key
// never matches unless nulls_equal.  Derived from the left-join tuples, which already come in left-row order with
// BL_IDX_NULL for misses: anti keeps the misses, semi the first tuple of every matched left row.
__global__ void __launch_bounds__(256) k_semi_anti_mask(const uint32_t* __restrict__ li, const uint32_t* __restrict__ ri, int64_t M, int64_t m_round, int anti, uint32_t* __restrict__ mask_words) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m_round; i += (int64_t)gridDim.x * blockDim.x) {
        bool keep = false;
        if (i < M) {
            const bool miss = ri[i] == J_NONE;
            keep = anti ? miss : (!miss && (i == 0 || li[i] != li[i - 1]));
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if ((threadIdx.x & 31) == 0) mask_words[i >> 5] = b;
    }
}

static JoinResult hash_join_inner_left(const DevCol& left, const DevCol& right, int how, bool nulls_equal, int maintain_order);

JoinResult op_hash_join(const DevCol& left, const DevCol& right, int how, bool nulls_equal, int maintain_order) {
    if (how != BL_JOIN_SEMI && how != BL_JOIN_ANTI) return hash_join_inner_left(left, right, how, nulls_equal, maintain_order);
    JoinResult lj = hash_join_inner_left(left, right, BL_JOIN_LEFT, nulls_equal, BL_ORDER_NONE);
    const int64_t M = lj.left.len;
    JoinResult r;
    r.right = idx_col(dev_alloc(16), 0, 0);
    if (M == 0) { r.left = idx_col(dev_alloc(16), 0, 0); return r; }
    DevCol mask = make_col(BL_BOOL, M, false);
    const int64_t m_round = (M + 31) / 32 * 32;
    PLB_LAUNCH("k8_semi_anti_mask", k_semi_anti_mask, grid_for(m_round, 256, 16), 256, 0, as<uint32_t>(lj.left.values), as<uint32_t>(lj.right.values), M, m_round, how == BL_JOIN_ANTI ? 1 : 0,
               as<uint32_t>(mask.values));
    DevCol li = idx_col(lj.left.values, M, 0);
    std::vector<DevCol> in{li}, out;
    op_filter(in, mask, out);
    r.left = out[0];
    return r;
}

static JoinResult hash_join_inner_left(const DevCol& left, const DevCol& right, int how, bool nulls_equal, int maintain_order) {
    PLB_REQUIRE(how == BL_JOIN_INNER || how == BL_JOIN_LEFT, BL_ERR_UNSUPPORTED, "join: only inner, left, semi and anti joins are on the hot path");
    PLB_REQUIRE(left.dtype == right.dtype, BL_ERR_DTYPE, std::string("join: key dtypes differ (") + dtype_name(left.dtype) + " vs " + dtype_name(right.dtype) + ")");   // join/mod.rs:231-241
    const int dt = left.dtype;
    PLB_REQUIRE(dt == BL_INT64 || dt == BL_UINT64 || dt == BL_INT32 || dt == BL_UINT32 || dt == BL_FLOAT64 || dt == BL_FLOAT32, BL_ERR_UNSUPPORTED,
                std::string("join: key dtype ") + dtype_name(dt) + " is outside the hot path");
    PLB_REQUIRE(left.len < 0xFFFFFFFFll && right.len < 0xFFFFFFFFll, BL_ERR_UNSUPPORTED, "join: more than 2^32-2 rows (IdxSize = u32)");
    // hash_join/mod.rs:41-50: probe the longer relation; on a tie the right side probes (swapped)
    const bool swapped = how == BL_JOIN_INNER && !(left.len > right.len);
    const DevCol& probe = swapped ? right : left;
    const DevCol& build = swapped ? left : right;
    const int64_t nb = build.len, np = probe.len;
    Context& c = ctx();

    trace_point("join:start");
    // ---- build: dense direct-address table when the build keys are dense unique integers
    const bool is_int = dtype_is_int(dt);
    const int elem = dtype_size(dt);
    const int sign_bits = dtype_is_signed(dt) ? elem * 8 : 0;
    const char* env_dense = getenv("BL_JOIN_DENSE");
    bool dense = false; uint64_t kmin = 0, range = 0; DevPtr dense_table;
    if (is_int && !nulls_equal && nb >= 1024 && !(env_dense && env_dense[0] == '0')) {
        DevPtr mm = dev_alloc(16);
        unsigned long long init_mm[2] = {~0ull, 0ull};
        PLB_CUDA(cudaMemcpyAsync(mm->p, init_mm, 16, cudaMemcpyHostToDevice, c.stream));
        PLB_LAUNCH("k7_join_minmax", k_join_minmax, grid_for(nb, 256), 256, 0, build.v(), build.vm(), elem, sign_bits, nb, as<unsigned long long>(mm));
        unsigned long long hmm[2];
        PLB_CUDA(cudaMemcpyAsync(hmm, mm->p, 16, cudaMemcpyDeviceToHost, c.stream));
        PLB_CUDA(cudaStreamSynchronize(c.stream));
        if (hmm[0] <= hmm[1] && hmm[1] - hmm[0] < (unsigned long long)8 * (unsigned long long)nb) {
            kmin = hmm[0]; range = hmm[1] - hmm[0] + 1;
            dense_table = dev_alloc((size_t)range * 4 + 16);
            DevPtr dups = dev_alloc(4); dev_memset(dups->p, 0, 4);
            PLB_LAUNCH("k7_dense_init", k_fill_u32j, grid_for((int64_t)range, 256), 256, 0, as<uint32_t>(dense_table), J_NONE, (int64_t)range);
            PLB_LAUNCH("k7_dense_build", k_join_dense_build, grid_for(nb, 256), 256, 0, as<uint32_t>(dense_table), build.v(), build.vm(), elem, sign_bits, nb, (uint64_t)kmin, as<int>(dups));
            dense = read_scalar(as<int>(dups)) == 0;       // duplicates -> hashed table + CSR lists
        }
    }
    JoinTableDev T; T.cap = (uint64_t)std::max<int64_t>(2 * nb, 16); T.entries = nullptr;
    DevPtr entries, slot_of_row, sorted_rows;
    bool csr = false;
    if (!dense) {
        entries = dev_alloc((size_t)(T.cap + 2) * 16);
        T.entries = as<uint4>(entries);
        PLB_LAUNCH("k7_join_init", k_join_init, grid_for((int64_t)T.cap + 2, 256), 256, 0, T.entries, (int64_t)T.cap + 2);
        slot_of_row = dev_alloc((size_t)std::max<int64_t>(nb, 1) * 4);
        DevPtr has_dups = dev_alloc(4);
        dev_memset(has_dups->p, 0, 4);
        if (nb > 0)
            PLB_LAUNCH("k7_join_build", k_join_build, grid_for(nb, 256), 256, 0, T, build.v(), build.vm(), dt, nb, nulls_equal ? 1 : 0, as<uint32_t>(slot_of_row), as<int>(has_dups));
        csr = nb > 0 && read_scalar(as<int>(has_dups)) != 0;
        if (csr) {
            const int64_t ne = (int64_t)T.cap + 2, ntiles_e = (ne + J_TILE - 1) / J_TILE;
            DevPtr sums = dev_alloc((size_t)ntiles_e * 4), offs = dev_alloc((size_t)ntiles_e * 8);
            PLB_LAUNCH("k7_tile_sums", k_join_tile_sums, grid_for(ntiles_e * 256, 256), 256, 0, T.entries, ne, as<uint32_t>(sums));
            exclusive_scan_u32_to_u64(as<uint32_t>(sums), as<uint64_t>(offs), ntiles_e, nullptr);
            PLB_LAUNCH("k7_csr_offsets", k_join_csr_offsets, grid_for(ntiles_e * 256, 256), 256, 0, T.entries, ne, as<uint64_t>(offs));
            // ascending row lists: stable sort of the build rows by entry index (skipped null rows sort last)
            sorted_rows = dev_alloc((size_t)nb * 4);
            iota_u32(as<uint32_t>(sorted_rows), nb, 0);
            sort_pairs_u32(as<uint32_t>(slot_of_row), as<uint32_t>(sorted_rows), nb);
        }
    }

    trace_point("join:build");
    // ---- unique build keys: fused single-pass probe + emit
    static const int fused_on = [] { const char* e = getenv("BL_JOIN_FUSED"); return e ? atoi(e) : 1; }();
    const int64_t ntiles = (np + J_TILE - 1) / J_TILE;
    uint64_t M = 0;
    DevPtr out_probe, out_build;
    bool done = false;
    if (fused_on && !csr && np > 0) {
        out_probe = dev_alloc((size_t)np * 4 + 16); out_build = dev_alloc((size_t)np * 4 + 16);
        DevPtr st = dev_alloc((size_t)ntiles * 8 + 16), ctl = dev_alloc(8);
        dev_memset(st->p, 0, (size_t)ntiles * 8 + 16); dev_memset(ctl->p, 0, 8);
        const bool kn = probe.validity != nullptr;
        const int left_join = how == BL_JOIN_LEFT ? 1 : 0;
        const int grid = (int)std::min<int64_t>(ntiles, (int64_t)c.sm_count * 6);
        unsigned long long* stp = as<unsigned long long>(st); unsigned* cnt = as<unsigned>(ctl); int* err = as<int>(ctl) + 1;
        const uint32_t* tb = as<uint32_t>(dense_table);
#define PE_LAUNCH(E, C, D)                                                                                                                          \
        do { if (kn) PLB_LAUNCH("k8_join_probe_emit", (k_join_probe_emit<E, C, true, D>), grid, 256, 0, T, tb, kmin, range, sign_bits, probe.v(), probe.vm(), np, nulls_equal ? 1 : 0, left_join, stp, cnt, err, as<uint32_t>(out_probe), as<uint32_t>(out_build)); \
             else PLB_LAUNCH("k8_join_probe_emit", (k_join_probe_emit<E, C, false, D>), grid, 256, 0, T, tb, kmin, range, sign_bits, probe.v(), probe.vm(), np, nulls_equal ? 1 : 0, left_join, stp, cnt, err, as<uint32_t>(out_probe), as<uint32_t>(out_build)); } while (0)
        if (dense) { if (elem == 8) PE_LAUNCH(8, 0, true); else PE_LAUNCH(4, 0, true); }
        else if (dt == BL_FLOAT64) PE_LAUNCH(8, 1, false);
        else if (dt == BL_FLOAT32) PE_LAUNCH(4, 2, false);
        else if (elem == 8) PE_LAUNCH(8, 0, false);
        else PE_LAUNCH(4, 0, false);
#undef PE_LAUNCH
        unsigned long long last = 0; int herr[2] = {0, 0};
        PLB_CUDA(cudaMemcpyAsync(&last, stp + (ntiles - 1), 8, cudaMemcpyDeviceToHost, c.stream));
        PLB_CUDA(cudaMemcpyAsync(herr, ctl->p, 8, cudaMemcpyDeviceToHost, c.stream));
        PLB_CUDA(cudaStreamSynchronize(c.stream));
        if (herr[1] == 0) { M = last & LB_VAL; done = true; }     // else: look-back gave up -> two-pass path below
    }
    trace_point("join:probe");
    if (!done) {
    // ---- probe pass 1
    DevPtr handle = dev_alloc((size_t)std::max<int64_t>(np, 1) * 4 + 16), tc = dev_alloc((size_t)std::max<int64_t>(ntiles, 1) * 8), toff = dev_alloc((size_t)std::max<int64_t>(ntiles, 1) * 8), total = dev_alloc(8);
    dev_memset(tc->p, 0, (size_t)std::max<int64_t>(ntiles, 1) * 8); dev_memset(total->p, 0, 8);
    if (np > 0) {
        const int grid = grid_for((np + 1) / 2, 256);
        const bool kn = probe.validity != nullptr;
        const int left_join = how == BL_JOIN_LEFT ? 1 : 0;
        uint32_t* hp = as<uint32_t>(handle); unsigned long long* tcp = as<unsigned long long>(tc);
        if (dense) {
            const uint32_t* tb = as<uint32_t>(dense_table);
            if (elem == 8) { if (kn) PLB_LAUNCH("k8_dense_probe", (k_join_dense_probe<8, true>), grid, 256, 0, tb, kmin, range, sign_bits, probe.v(), probe.vm(), np, left_join, hp, tcp);
                             else PLB_LAUNCH("k8_dense_probe", (k_join_dense_probe<8, false>), grid, 256, 0, tb, kmin, range, sign_bits, probe.v(), probe.vm(), np, left_join, hp, tcp); }
            else { if (kn) PLB_LAUNCH("k8_dense_probe", (k_join_dense_probe<4, true>), grid, 256, 0, tb, kmin, range, sign_bits, probe.v(), probe.vm(), np, left_join, hp, tcp);
                   else PLB_LAUNCH("k8_dense_probe", (k_join_dense_probe<4, false>), grid, 256, 0, tb, kmin, range, sign_bits, probe.v(), probe.vm(), np, left_join, hp, tcp); }
        }
        else if (dt == BL_FLOAT64) launch_probe<8, 1>(kn, grid, T, probe, nulls_equal, csr, left_join, hp, tcp);
        else if (dt == BL_FLOAT32) launch_probe<4, 2>(kn, grid, T, probe, nulls_equal, csr, left_join, hp, tcp);
        else if (dtype_size(dt) == 8) launch_probe<8, 0>(kn, grid, T, probe, nulls_equal, csr, left_join, hp, tcp);
        else launch_probe<4, 0>(kn, grid, T, probe, nulls_equal, csr, left_join, hp, tcp);
        exclusive_scan_u64(as<uint64_t>(tc), as<uint64_t>(toff), ntiles, as<uint64_t>(total));
        M = read_scalar(as<unsigned long long>(total));
    }
    PLB_REQUIRE(M < 0xFFFFFFFFull, BL_ERR_UNSUPPORTED, "join: result has more than 2^32-2 rows (IdxSize = u32)");
    // ---- probe pass 2
    out_probe = dev_alloc((size_t)std::max<uint64_t>(M, 1) * 4 + 16); out_build = dev_alloc((size_t)std::max<uint64_t>(M, 1) * 4 + 16);
    if (M > 0)
        PLB_LAUNCH("k8_join_emit", k_join_emit, (int)std::min<int64_t>(ntiles, (int64_t)c.sm_count * 8), 256, 0, T, as<uint32_t>(handle), np, csr ? 1 : 0, how == BL_JOIN_LEFT ? 1 : 0,
                   as<uint32_t>(sorted_rows), as<uint64_t>(toff), as<uint32_t>(out_probe), as<uint32_t>(out_build));
    }
    if (!out_probe) { out_probe = dev_alloc(16); out_build = dev_alloc(16); }
    trace_point("join:emit");
    JoinResult r;
    r.left = idx_col(swapped ? out_build : out_probe, (int64_t)M, 0);
    r.right = idx_col(swapped ? out_probe : out_build, (int64_t)M, how == BL_JOIN_LEFT ? -1 : 0);
    // maintain_order (join/mod.rs:577-642): stable sort on the requested side unless already in that order
    if (how == BL_JOIN_INNER && maintain_order != BL_ORDER_NONE && M > 1) {
        const bool by_left = maintain_order == BL_ORDER_LEFT || maintain_order == BL_ORDER_LEFT_RIGHT;
        const bool left_sorted = !swapped;
        if (by_left && !left_sorted) sort_pairs_u32(as<uint32_t>(r.left.values), as<uint32_t>(r.right.values), (int64_t)M);
        else if (!by_left && !swapped) sort_pairs_u32(as<uint32_t>(r.right.values), as<uint32_t>(r.left.values), (int64_t)M);
    }
    // left join: only Right / RightLeft reorder (stable sort on the right idx, unmatched rows = u32::MAX last;
    // dispatch_left_right.rs:142-170); the probe order already is the left order
    if (how == BL_JOIN_LEFT && (maintain_order == BL_ORDER_RIGHT || maintain_order == BL_ORDER_RIGHT_LEFT) && M > 1)
        sort_pairs_u32(as<uint32_t>(r.right.values), as<uint32_t>(r.left.values), (int64_t)M);
    if (how == BL_JOIN_LEFT && M > 0) {
        // Arrow-proper validity for the nullable right index: bit = (idx != BL_IDX_NULL)
        DevCol none = make_col(BL_UINT32, 1, false);
        const uint32_t nv = J_NONE;
        PLB_CUDA(cudaMemcpyAsync(none.values->p, &nv, 4, cudaMemcpyHostToDevice, c.stream));
        PLB_CUDA(cudaStreamSynchronize(c.stream));
        DevCol m = op_compare(BL_CMP_NE, r.right, none, false);
        r.right.validity = m.values;
        r.right.null_count = -1;
    }
    trace_point("join:done");
    return r;
}

}  // namespace plb
