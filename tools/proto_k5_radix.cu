// tools/proto_k5_radix.cu — design experiment for the next round (NOT part of the product, not built by build()):
// can a two-pass radix-partitioned group_by beat the single-pass L2-atomic K5 on the C2 shape
// (1e8 rows, 1e6 Int64 keys, sum(i64) / sum(f64) / len)?
//
// K5 today issues 1 key load + 3 REDs per row against an L2-resident table and sits at the L2's random-op
// ceiling (1.95 ms, profiles/README.md).  The alternative moves 3x the bytes but replaces every L2 atomic by a
// shared-memory atomic:
//   pass 1  k_scatter   rows -> P partitions by the top hash bits.  Each CTA sorts a tile of T rows by partition
//                        in shared memory (counting sort), reserves space per partition with ONE global atomic
//                        per (tile, partition) and writes each partition's run contiguously.
//   pass 2  k_agg       one CTA per partition: the partition's ~G/P groups fit a shared-memory table; rows are
//                        aggregated with shared-memory atomics and the groups written out compactly.
// Break-even: (2.4 GB read + 2.4 GB scattered write + 2.4 GB read) must finish in < 1.95 ms, i.e. the scatter
// pass has to sustain ~4.5 TB/s combined.  This program measures exactly that, next to the L2-atomic baseline,
// and checks the results against it.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/proto_k5_radix tools/proto_k5_radix.cu
// Run:   tools/proto_k5_radix [rows=100000000] [keys=1000000] [log2_partitions=10]     (log2_partitions in 8..11)
// Expected reading: radix_total_ms < baseline_l2_red_ms (and < K5's 1.95 ms) makes the integration worth it; the
// scatter pass is the uncertain part (32..128-byte runs per partition and tile).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

#ifndef TILE
#define TILE 4096      // rows per CTA tile in pass 1 (-DTILE=8192: 216 KB of shared memory, runs twice as long)
#endif
static constexpr uint64_t RANDOM_ODD = 0x55fbfd6bfc5458e9ULL;      // the reference's DirtyHash multiplier
static constexpr uint64_t EMPTY = 0x8000000000000000ULL;
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// ---------------------------------------------------------------- synthetic C2 columns (device-side generator)
__global__ void k_gen(uint64_t* key, int64_t* vi, double* vf, int64_t n, uint64_t keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix((uint64_t)i + 0x9e3779b97f4a7c15ULL);
        key[i] = h % keys;
        vi[i] = (int64_t)(mix(h) % 2000) - 1000;
        vf[i] = (double)(mix(h + 1) % 100000000ULL) * 1e-6;
    }
}

// ---------------------------------------------------------------- baseline: dense arrays indexed by key, L2 REDs
// (keys are < `keys` in this harness, so the table lookup of the real K5 degenerates to the key itself: this is
//  the optimistic version of the single-pass design — no key-plane load at all)
__global__ void k_base(const uint64_t* __restrict__ key, const int64_t* __restrict__ vi, const double* __restrict__ vf, int64_t n,
                       unsigned long long* si, double* sf, unsigned* len) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = __ldcs(key + i);
        atomicAdd(si + k, (unsigned long long)__ldcs(vi + i));
        atomicAdd(sf + k, __ldcs(vf + i));
        atomicAdd(len + k, 1u);
    }
}

// ---------------------------------------------------------------- pass 1: tile-sorted scatter
// Partition buffers are SoA: column c of partition p lives at buf_c[p * cap ... p * cap + cursor[p]).
template <int LOGP, int T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_scatter(const uint64_t* __restrict__ key, const uint64_t* __restrict__ v1, const uint64_t* __restrict__ v2, int64_t n,
                                                     uint64_t* __restrict__ pk, uint64_t* __restrict__ p1, uint64_t* __restrict__ p2, int64_t cap,
                                                     unsigned* __restrict__ cursor, int* __restrict__ overflow) {
    constexpr int P = 1 << LOGP, R = T / THREADS;
    extern __shared__ uint64_t smem[];
    uint64_t* sk = smem;                 // T keys, sorted by partition
    uint64_t* s1 = sk + T;
    uint64_t* s2 = s1 + T;
    unsigned* hist = reinterpret_cast<unsigned*>(s2 + T);      // P counts -> run starts inside the tile
    unsigned* fill = hist + P;                                  // P fill cursors
    unsigned* gbase = fill + P;                                 // P global run starts
    uint16_t* sp = reinterpret_cast<uint16_t*>(gbase + P);     // T partition ids of the sorted slots
    __shared__ unsigned warp_tot[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ntiles = (n + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * T;
        const int rows = (int)min((int64_t)T, n - base);
        for (int p = tid; p < P; p += THREADS) { hist[p] = 0; fill[p] = 0; }
        __syncthreads();
        uint64_t k[R]; unsigned part[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int r = j * THREADS + tid;
            if (r < rows) { k[j] = __ldcs(key + base + r); part[j] = (unsigned)((k[j] * RANDOM_ODD) >> (64 - LOGP)); atomicAdd(&hist[part[j]], 1u); }
        }
        __syncthreads();
        // exclusive scan of hist (P <= 2 * THREADS handled generally: each thread owns P / THREADS consecutive bins)
        constexpr int BINS = (P + THREADS - 1) / THREADS;
        unsigned cnt[BINS], mine = 0;
#pragma unroll
        for (int b = 0; b < BINS; b++) { const int p = tid * BINS + b; cnt[b] = p < P ? hist[p] : 0; mine += cnt[b]; }
        unsigned x = mine;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned w = lane < THREADS / 32 ? warp_tot[lane] : 0, s = w;
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            if (lane < THREADS / 32) warp_tot[lane] = s - w;
        }
        __syncthreads();
        unsigned run = warp_tot[warp] + x - mine;
#pragma unroll
        for (int b = 0; b < BINS; b++) {
            const int p = tid * BINS + b;
            if (p < P) {
                hist[p] = run;                                                     // run start inside the tile
                gbase[p] = cnt[b] ? atomicAdd(&cursor[p], cnt[b]) : 0u;           // ONE global atomic per (tile, partition)
                run += cnt[b];
            }
        }
        __syncthreads();
        // place the rows (order inside a run is arbitrary: sums and counts do not care)
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int r = j * THREADS + tid;
            if (r < rows) {
                const unsigned pos = hist[part[j]] + atomicAdd(&fill[part[j]], 1u);
                sk[pos] = k[j]; s1[pos] = __ldcs(v1 + base + r); s2[pos] = __ldcs(v2 + base + r); sp[pos] = (uint16_t)part[j];
            }
        }
        __syncthreads();
        // write the runs: consecutive slots of one partition go to consecutive addresses
        for (int i = tid; i < rows; i += THREADS) {
            const unsigned p = sp[i];
            const int64_t g = (int64_t)gbase[p] + (i - (int)hist[p]);
            if (g < cap) { const int64_t at = (int64_t)p * cap + g; pk[at] = sk[i]; p1[at] = s1[i]; p2[at] = s2[i]; }
            else *overflow = 1;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- pass 2: one CTA per partition, shared-memory table
__device__ __forceinline__ void s_add_u64(uint64_t* a, uint64_t v) {       // exact 64-bit add from two 32-bit shared-memory atomics
    unsigned* w = reinterpret_cast<unsigned*>(a);
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(w, lo);
    const unsigned up = hi + (((unsigned)(old + lo) < old) ? 1u : 0u);
    if (up) atomicAdd(w + 1, up);
}
template <int LOGP, int SLOTS, int THREADS>
__global__ void __launch_bounds__(THREADS) k_agg(const uint64_t* __restrict__ pk, const uint64_t* __restrict__ p1, const uint64_t* __restrict__ p2, int64_t cap,
                                                 const unsigned* __restrict__ cursor, uint64_t* __restrict__ out_key, uint64_t* __restrict__ out_si, double* __restrict__ out_sf,
                                                 unsigned* __restrict__ out_len, unsigned long long* __restrict__ out_cursor, int* __restrict__ overflow) {
    extern __shared__ uint64_t smem[];
    uint64_t* tkey = smem;
    uint64_t* tsi = tkey + SLOTS;
    double* tsf = reinterpret_cast<double*>(tsi + SLOTS);
    unsigned* tlen = reinterpret_cast<unsigned*>(tsf + SLOTS);
    __shared__ unsigned s_used, s_base;
    const int tid = threadIdx.x;
    const int p = blockIdx.x;
    for (int i = tid; i < SLOTS; i += THREADS) { tkey[i] = EMPTY; tsi[i] = 0; tsf[i] = 0.0; tlen[i] = 0; }
    if (tid == 0) s_used = 0;
    __syncthreads();
    const int64_t rows = min((int64_t)cursor[p], cap);
    const uint64_t* k = pk + (int64_t)p * cap; const uint64_t* a = p1 + (int64_t)p * cap; const uint64_t* b = p2 + (int64_t)p * cap;
    for (int64_t i = tid; i < rows; i += THREADS) {
        const uint64_t key = __ldcs(k + i);
        unsigned slot = (unsigned)(((key * RANDOM_ODD) << LOGP) >> (64 - 11)) & (SLOTS - 1);      // the hash bits below the partition bits
        int probes = 0;
        for (; probes < SLOTS; probes++) {
            const uint64_t cur = *reinterpret_cast<volatile uint64_t*>(tkey + slot);
            if (cur == key) break;
            if (cur == EMPTY) {
                const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(tkey + slot), (unsigned long long)EMPTY, (unsigned long long)key);
                if (old == EMPTY) { atomicAdd(&s_used, 1u); break; }
                if (old == key) break;
            }
            slot = (slot + 1) & (SLOTS - 1);
        }
        if (probes == SLOTS) { *overflow = 2; continue; }
        s_add_u64(tsi + slot, __ldcs(a + i));
        atomicAdd(tsf + slot, __longlong_as_double((long long)__ldcs(b + i)));
        atomicAdd(tlen + slot, 1u);
    }
    __syncthreads();
    if (tid == 0) s_base = (unsigned)atomicAdd(out_cursor, (unsigned long long)s_used);
    __syncthreads();
    // compact the used slots (order inside the partition is irrelevant: unordered group_by output)
    for (int i = tid; i < SLOTS; i += THREADS) {
        if (tkey[i] == EMPTY) continue;
        const unsigned at = s_base + atomicSub(&s_used, 1u) - 1u;
        out_key[at] = tkey[i]; out_si[at] = tsi[i]; out_sf[at] = tsf[i]; out_len[at] = tlen[i];
    }
}

// ---------------------------------------------------------------- check: scatter the compact result back by key
__global__ void k_check(const uint64_t* out_key, const uint64_t* out_si, const double* out_sf, const unsigned* out_len, int64_t G,
                        const unsigned long long* si, const double* sf, const unsigned* len, unsigned long long* bad) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = out_key[g];
        const bool ok = out_si[g] == si[k] && out_len[g] == len[k] && fabs(out_sf[g] - sf[k]) <= 1e-9 * fabs(sf[k]) + 1e-9;
        if (!ok) atomicAdd(bad, 1ull);
    }
}

template <int LOGP>
static void run(int64_t n, uint64_t keys) {
    constexpr int P = 1 << LOGP, T = TILE, THREADS = 512, SLOTS = 2048;
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    uint64_t *key, *pk, *p1, *p2, *out_key, *out_si; int64_t* vi; double *vf, *sf, *out_sf; unsigned long long *si, *out_cursor, *bad; unsigned *len, *cursor, *out_len; int* overflow;
    const int64_t cap = (int64_t)((double)n / P * 1.15) + 4096;
    CK(cudaMalloc(&key, n * 8)); CK(cudaMalloc(&vi, n * 8)); CK(cudaMalloc(&vf, n * 8));
    CK(cudaMalloc(&si, keys * 8)); CK(cudaMalloc(&sf, keys * 8)); CK(cudaMalloc(&len, keys * 4));
    CK(cudaMalloc(&pk, (size_t)P * cap * 8)); CK(cudaMalloc(&p1, (size_t)P * cap * 8)); CK(cudaMalloc(&p2, (size_t)P * cap * 8));
    CK(cudaMalloc(&cursor, P * 4)); CK(cudaMalloc(&overflow, 4)); CK(cudaMalloc(&out_cursor, 8)); CK(cudaMalloc(&bad, 8));
    CK(cudaMalloc(&out_key, keys * 8)); CK(cudaMalloc(&out_si, keys * 8)); CK(cudaMalloc(&out_sf, keys * 8)); CK(cudaMalloc(&out_len, keys * 4));
    k_gen<<<sms * 8, 256>>>(key, vi, vf, n, keys); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const size_t smem1 = (size_t)3 * T * 8 + (size_t)3 * P * 4 + (size_t)T * 2;
    const size_t smem2 = (size_t)SLOTS * (8 + 8 + 8 + 4);
    auto* ks = k_scatter<LOGP, T, THREADS>; auto* ka = k_agg<LOGP, SLOTS, THREADS>;
    CK(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    CK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    float best_base = 1e9f, best_s = 1e9f, best_a = 1e9f;
    for (int it = 0; it < 5; it++) {
        CK(cudaMemset(si, 0, keys * 8)); CK(cudaMemset(sf, 0, keys * 8)); CK(cudaMemset(len, 0, keys * 4));
        CK(cudaEventRecord(e0));
        k_base<<<sms * 8, 256>>>(key, vi, vf, n, si, sf, len);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best_base = fminf(best_base, ms);
        CK(cudaMemset(cursor, 0, P * 4)); CK(cudaMemset(overflow, 0, 4)); CK(cudaMemset(out_cursor, 0, 8));
        CK(cudaEventRecord(e0));
        ks<<<sms * 1, THREADS, smem1>>>(key, (const uint64_t*)vi, (const uint64_t*)vf, n, pk, p1, p2, cap, cursor, overflow);
        CK(cudaEventRecord(e1));
        ka<<<P, THREADS, smem2>>>(pk, p1, p2, cap, cursor, out_key, out_si, out_sf, out_len, out_cursor, overflow);
        CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2));
        float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
        best_s = fminf(best_s, a); best_a = fminf(best_a, b);
    }
    int h_over = 0; unsigned long long G = 0, h_bad = 0;
    CK(cudaMemcpy(&h_over, overflow, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&G, out_cursor, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(bad, 0, 8));
    k_check<<<sms * 4, 256>>>(out_key, out_si, out_sf, out_len, (int64_t)G, si, sf, len, bad);
    CK(cudaMemcpy(&h_bad, bad, 8, cudaMemcpyDeviceToHost));
    const double gb = (double)n * 24 / 1e9;
    printf("{\"rows\": %lld, \"keys\": %llu, \"partitions\": %d, \"tile_rows\": %d, \"baseline_l2_red_ms\": %.3f, \"scatter_ms\": %.3f, \"scatter_GBps_rw\": %.0f, \"agg_ms\": %.3f, "
           "\"radix_total_ms\": %.3f, \"groups\": %llu, \"overflow\": %d, \"mismatching_groups\": %llu}\n",
           (long long)n, (unsigned long long)keys, P, T, best_base, best_s, 2 * gb / (best_s / 1e3), best_a, best_s + best_a, G, h_over, h_bad);
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 100000000LL;
    const uint64_t keys = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1000000ULL;
    const int logp = argc > 3 ? atoi(argv[3]) : 10;
    switch (logp) {
        case 8: run<8>(n, keys); break;
        case 9: run<9>(n, keys); break;
        case 11: run<11>(n, keys); break;
        default: run<10>(n, keys); break;
    }
    return 0;
}
