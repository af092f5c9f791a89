// tools/proto_radix2.cu — round-2 design experiment (NOT part of the product; not built by build()):
// the two-pass radix plan for the C2 group_by with TMA on both sides of shared memory.
//
//   pass 1  k_rp_scatter   rows -> B buckets by the top hash bits of key * RANDOM_ODD.  Every CTA sorts a tile of T
//                          rows by bucket in shared memory as ROW-MAJOR records [key, v1, v2] (24 B), reserves one
//                          run per (tile, bucket) with a single global atomic and writes the run out either
//                            BULK=1: with ONE cp.async.bulk (TMA) shared->global copy per run (runs padded to an
//                                    even row count with a GB_EMPTY-key pad row so both ends stay 16-byte aligned), or
//                            BULK=0: with coalesced 8-byte stores (thread i owns word i of the sorted tile).
//   pass 2  k_rp_agg       one CTA per bucket: the bucket's record stream is staged into shared memory by
//                          cp.async.bulk (TMA) global->shared copies on an mbarrier ring; rows are aggregated into a
//                          shared-memory open-addressing table with shared-memory atomics; groups leave compacted.
// Compared against the single-pass L2-atomic baseline (3 REDs per row) on the same data, results checked against it.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/proto_radix2 tools/proto_radix2.cu
// Run:   tools/proto_radix2 [rows=100000000] [keys=1000000]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static constexpr uint64_t RANDOM_ODD = 0x55fbfd6bfc5458e9ULL;
static constexpr uint64_t EMPTY = 0x8000000000000000ULL;
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

__global__ void k_gen(uint64_t* key, int64_t* vi, double* vf, int64_t n, uint64_t keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix((uint64_t)i + 0x9e3779b97f4a7c15ULL);
        key[i] = mix(h % keys + 77);        // sparse 64-bit keys (no dense-range shortcut possible)
        vi[i] = (int64_t)(mix(h) % 2000) - 1000;
        vf[i] = (double)(mix(h + 1) % 100000000ULL) * 1e-6;
    }
}

// ---------------------------------------------------------------- PTX wrappers (TMA bulk copies, mbarrier)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ ulonglong2 ld_stream2(const uint64_t* p) {
    ulonglong2 v; asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p)); return v;
}

// ---------------------------------------------------------------- baseline: single pass, L2 atomics (dense arrays by key id)
__global__ void k_base(const uint64_t* __restrict__ kid, const int64_t* __restrict__ vi, const double* __restrict__ vf, int64_t n,
                       unsigned long long* si, double* sf, unsigned* len) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = __ldcs(kid + i);
        atomicAdd(si + k, (unsigned long long)__ldcs(vi + i));
        atomicAdd(sf + k, __ldcs(vf + i));
        atomicAdd(len + k, 1u);
    }
}
__global__ void k_keyid(const uint64_t* key, uint64_t* kid, int64_t n, uint64_t keys) {      // for the baseline / checker only
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix((uint64_t)i + 0x9e3779b97f4a7c15ULL);
        kid[i] = h % keys;
    }
}

// ---------------------------------------------------------------- pass 1
// Record stream of bucket p: out[p * cap_rows * ROWW ...], cursor[p] rows used (including pad rows).
template <int ROWW, int RPT, int THREADS, bool BULK>
__global__ void __launch_bounds__(THREADS) k_rp_scatter(const uint64_t* __restrict__ key, const uint64_t* __restrict__ v1, const uint64_t* __restrict__ v2, int64_t n, int logB,
                                                        uint64_t* __restrict__ out, int64_t cap_rows, unsigned* __restrict__ cursor, int* __restrict__ overflow) {
    constexpr int T = THREADS * RPT;
    const int B = 1 << logB;
    extern __shared__ __align__(16) uint64_t smem[];
    const int stage_rows = T + (BULK ? B : 0);
    uint64_t* stage = smem;
    unsigned* hist = reinterpret_cast<unsigned*>(stage + (size_t)stage_rows * ROWW);
    unsigned* start = hist + B;
    unsigned* gpos = start + B;
    uint16_t* sp = reinterpret_cast<uint16_t*>(gpos + B);      // BULK=0 only: bucket of each sorted slot
    __shared__ unsigned warp_tot[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ntiles = (n + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * T;
        for (int p = tid; p < B; p += THREADS) hist[p] = 0;
        __syncthreads();
        uint64_t k[RPT]; unsigned pk[RPT];
#pragma unroll
        for (int j = 0; j < RPT / 2; j++) {
            const int64_t r0 = base + 2 * (int64_t)(j * THREADS + tid);
            k[2 * j] = k[2 * j + 1] = 0;
            if (r0 + 1 < n) { const ulonglong2 t = ld_stream2(key + r0); k[2 * j] = t.x; k[2 * j + 1] = t.y; }
            else if (r0 < n) k[2 * j] = key[r0];
        }
#pragma unroll
        for (int j = 0; j < RPT; j++) {
            const int64_t r = base + 2 * (int64_t)((j >> 1) * THREADS + tid) + (j & 1);
            pk[j] = 0xFFFFFFFFu;
            if (r < n) { const unsigned b = (unsigned)((k[j] * RANDOM_ODD) >> (64 - logB)); pk[j] = (b << 16) | atomicAdd(&hist[b], 1u); }
        }
        __syncthreads();
        // exclusive scan of the (padded) counts; one global reservation per non-empty bucket
        const int bins = (B + THREADS - 1) / THREADS;
        unsigned mine = 0;
        for (int q = 0; q < bins; q++) { const int p = tid * bins + q; if (p < B) { unsigned c = hist[p]; if (BULK) c = (c + 1u) & ~1u; mine += c; } }
        unsigned x = mine;
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned w = lane < THREADS / 32 ? warp_tot[lane] : 0, s = w;
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            if (lane < THREADS / 32) warp_tot[lane] = s - w;
        }
        __syncthreads();
        unsigned run = warp_tot[warp] + x - mine;
        for (int q = 0; q < bins; q++) {
            const int p = tid * bins + q;
            if (p < B) {
                const unsigned c = hist[p], cp = BULK ? ((c + 1u) & ~1u) : c;
                start[p] = run;
                unsigned g = 0;
                if (cp) { g = atomicAdd(&cursor[p], cp); if ((int64_t)g + cp > cap_rows) { *overflow = 1; g = 0xFFFFFFFFu; } }
                gpos[p] = g;
                run += cp;
            }
        }
        if (BULK) bulk_wait_read0();        // the previous tile's copies have finished reading the staging buffer
        __syncthreads();
        if (BULK) for (int p = tid; p < B; p += THREADS) { const unsigned c = hist[p]; if (c & 1u) { uint64_t* pad = stage + (size_t)(start[p] + c) * ROWW; pad[0] = EMPTY; pad[1] = 0; pad[2] = 0; } }
        // place the records (order inside a run is arbitrary)
#pragma unroll
        for (int j = 0; j < RPT / 2; j++) {
            const int64_t r0 = base + 2 * (int64_t)(j * THREADS + tid);
            uint64_t a[2] = {0, 0}, b[2] = {0, 0};
            if (r0 + 1 < n) { const ulonglong2 t = ld_stream2(v1 + r0), u = ld_stream2(v2 + r0); a[0] = t.x; a[1] = t.y; b[0] = u.x; b[1] = u.y; }
            else if (r0 < n) { a[0] = v1[r0]; b[0] = v2[r0]; }
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const unsigned q = pk[2 * j + e];
                if (q == 0xFFFFFFFFu) continue;
                const unsigned bkt = q >> 16, pos = start[bkt] + (q & 0xFFFFu);
                uint64_t* rec = stage + (size_t)pos * ROWW;
                rec[0] = k[2 * j + e]; rec[1] = a[e]; rec[2] = b[e];
                if (!BULK) sp[pos] = (uint16_t)bkt;
            }
        }
        if (BULK) fence_async_smem();
        __syncthreads();
        if (BULK) {
            for (int p = tid; p < B; p += THREADS) {
                const unsigned c = hist[p], cp = (c + 1u) & ~1u, g = gpos[p];
                if (cp && g != 0xFFFFFFFFu) bulk_s2g(out + ((size_t)p * cap_rows + g) * ROWW, stage + (size_t)start[p] * ROWW, cp * ROWW * 8);
            }
            bulk_commit();
        } else {
            const int rows = (int)min((int64_t)T, n - base);
            for (int w = tid; w < rows * ROWW; w += THREADS) {
                const int row = w / ROWW, c = w - row * ROWW;
                const unsigned p = sp[row], g = gpos[p];
                if (g != 0xFFFFFFFFu) out[((size_t)p * cap_rows + g + (row - start[p])) * ROWW + c] = stage[w];
            }
            __syncthreads();
        }
    }
    if (BULK) bulk_wait0();
}

// ---------------------------------------------------------------- pass 2
__device__ __forceinline__ void s_add_u64(uint64_t* a, uint64_t v) {
    unsigned* w = reinterpret_cast<unsigned*>(a);
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(w, lo);
    const unsigned up = hi + (((unsigned)(old + lo) < old) ? 1u : 0u);
    if (up) atomicAdd(w + 1, up);
}
template <int ROWW, int SLOTS, int CR, int NST, int THREADS>
__global__ void __launch_bounds__(THREADS) k_rp_agg(const uint64_t* __restrict__ recs, int64_t cap_rows, const unsigned* __restrict__ cursor, int logB,
                                                    uint64_t* __restrict__ out_key, uint64_t* __restrict__ out_si, double* __restrict__ out_sf, unsigned* __restrict__ out_len,
                                                    unsigned long long* __restrict__ out_cursor, int* __restrict__ overflow) {
    extern __shared__ __align__(128) uint64_t smem[];
    uint64_t* ring = smem;                                   // NST x CR x ROWW
    uint64_t* tkey = ring + (size_t)NST * CR * ROWW;
    uint64_t* tsi = tkey + SLOTS;
    double* tsf = reinterpret_cast<double*>(tsi + SLOTS);
    unsigned* tlen = reinterpret_cast<unsigned*>(tsf + SLOTS);
    __shared__ uint64_t full[NST];
    __shared__ unsigned s_used, s_base;
    const int tid = threadIdx.x;
    const int B = 1 << logB;
    int lgS = 0; while ((1 << lgS) < SLOTS) lgS++;
    if (tid == 0) { for (int s = 0; s < NST; s++) mbar_init(&full[s], 1); fence_mbar_init(); }
    __syncthreads();
    unsigned q_issue = 0, q_wait = 0;      // chunk sequence numbers of this CTA (stage = q % NST, parity = (q / NST) & 1)
    for (int p = blockIdx.x; p < B; p += gridDim.x) {
        for (int i = tid; i < SLOTS; i += THREADS) { tkey[i] = EMPTY; tsi[i] = 0; tsf[i] = 0.0; tlen[i] = 0; }
        if (tid == 0) s_used = 0;
        const int64_t rows = min((int64_t)cursor[p], cap_rows);
        const uint64_t* src = recs + (size_t)p * cap_rows * ROWW;
        const int nch = (int)((rows + CR - 1) / CR);
        __syncthreads();
        // prologue: NST - 1 chunks in flight
        if (tid == 0) {
            for (int c = 0; c < NST - 1 && c < nch; c++) {
                const unsigned bytes = (unsigned)(min((int64_t)CR, rows - (int64_t)c * CR) * ROWW * 8);
                const int st = q_issue % NST;
                mbar_expect_tx(&full[st], bytes);
                bulk_g2s(ring + (size_t)st * CR * ROWW, src + (size_t)c * CR * ROWW, bytes, &full[st]);
                q_issue++;
            }
        }
        for (int c = 0; c < nch; c++) {
            if (tid == 0 && c + NST - 1 < nch) {          // keep the ring full: the stage being refilled was released by the barrier that ended chunk c - 1
                const int cc = c + NST - 1;
                const unsigned bytes = (unsigned)(min((int64_t)CR, rows - (int64_t)cc * CR) * ROWW * 8);
                const int st = q_issue % NST;
                mbar_expect_tx(&full[st], bytes);
                bulk_g2s(ring + (size_t)st * CR * ROWW, src + (size_t)cc * CR * ROWW, bytes, &full[st]);
                q_issue++;
            }
            const int st = q_wait % NST; const unsigned par = (q_wait / NST) & 1u;
            while (!mbar_try_wait(&full[st], par)) {}
            q_wait++;
            const uint64_t* buf = ring + (size_t)st * CR * ROWW;
            const int crow = (int)min((int64_t)CR, rows - (int64_t)c * CR);
            for (int r = tid; r < crow; r += THREADS) {
                const uint64_t key = buf[r * ROWW];
                if (key == EMPTY) continue;                 // pad row
                const uint64_t a = buf[r * ROWW + 1], b = buf[r * ROWW + 2];
                unsigned slot = (unsigned)(((key * RANDOM_ODD) << logB) >> (64 - lgS));
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
                s_add_u64(tsi + slot, a);
                atomicAdd(tsf + slot, __longlong_as_double((long long)b));
                atomicAdd(tlen + slot, 1u);
            }
            __syncthreads();        // everyone is done with this stage (and, for the last chunk, with the table)
        }
        if (tid == 0) s_base = (unsigned)atomicAdd(out_cursor, (unsigned long long)s_used);
        __syncthreads();
        for (int i = tid; i < SLOTS; i += THREADS) {
            if (tkey[i] == EMPTY) continue;
            const unsigned at = s_base + atomicSub(&s_used, 1u) - 1u;
            out_key[at] = tkey[i]; out_si[at] = tsi[i]; out_sf[at] = tsf[i]; out_len[at] = tlen[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- check against the baseline (key -> id through the generator's inverse table)
__global__ void k_mark(const uint64_t* key, const uint64_t* kid, int64_t n, uint64_t* key_of_id) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) key_of_id[kid[i]] = key[i];
}

// ---------------------------------------------------------------- pass 2, variant V1: plain coalesced loads, no staging
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct AggOut { uint64_t* key; uint64_t* si; double* sf; unsigned* len; unsigned long long* cursor; int* overflow; };

__device__ __forceinline__ void agg_row(uint64_t* tkey, uint64_t* tsi, double* tsf, unsigned* tlen, unsigned S, int logB, unsigned* s_used, int* overflow, uint64_t key, uint64_t a, uint64_t b) {
    if (key == EMPTY) return;                 // pad row
    const uint64_t h = key * RANDOM_ODD;
    unsigned slot = __umulhi((unsigned)((h << logB) >> 32), S);
    unsigned probes = 0;
    for (; probes < S; probes++) {
        const uint64_t cur = *reinterpret_cast<volatile uint64_t*>(tkey + slot);
        if (cur == key) break;
        if (cur == EMPTY) {
            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(tkey + slot), (unsigned long long)EMPTY, (unsigned long long)key);
            if (old == EMPTY) { atomicAdd(s_used, 1u); break; }
            if (old == key) break;
        }
        if (++slot == S) slot = 0;
    }
    if (probes == S) { *overflow = 2; return; }
    s_add_u64(tsi + slot, a);
    atomicAdd(tsf + slot, __longlong_as_double((long long)b));
    atomicAdd(tlen + slot, 1u);
}
__device__ __forceinline__ void agg_flush(const uint64_t* tkey, const uint64_t* tsi, const double* tsf, const unsigned* tlen, unsigned S, unsigned* s_used, unsigned base, const AggOut& o, int tid, int nthr) {
    for (unsigned i = tid; i < S; i += nthr) {
        if (tkey[i] == EMPTY) continue;
        const unsigned at = base + atomicSub(s_used, 1u) - 1u;
        o.key[at] = tkey[i]; o.si[at] = tsi[i]; o.sf[at] = tsf[i]; o.len[at] = tlen[i];
    }
}

template <int ROWW, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_rp_agg_direct(const uint64_t* __restrict__ recs, int64_t cap_rows, const unsigned* __restrict__ cursor, int logB, unsigned S, AggOut o) {
    extern __shared__ __align__(16) uint64_t smem[];
    uint64_t* tkey = smem; uint64_t* tsi = tkey + S; double* tsf = reinterpret_cast<double*>(tsi + S); unsigned* tlen = reinterpret_cast<unsigned*>(tsf + S);
    __shared__ unsigned s_used, s_base;
    const int tid = threadIdx.x, B = 1 << logB;
    for (int p = blockIdx.x; p < B; p += gridDim.x) {
        for (unsigned i = tid; i < S; i += THREADS) { tkey[i] = EMPTY; tsi[i] = 0; tsf[i] = 0.0; tlen[i] = 0; }
        if (tid == 0) s_used = 0;
        __syncthreads();
        const int64_t rows = min((int64_t)cursor[p], cap_rows);
        const uint64_t* src = recs + (size_t)p * cap_rows * ROWW;
        for (int64_t r0 = tid; r0 < rows; r0 += (int64_t)THREADS * UNROLL) {
            uint64_t k[UNROLL], a[UNROLL], b[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int64_t r = r0 + (int64_t)u * THREADS;
                k[u] = EMPTY; a[u] = 0; b[u] = 0;
                if (r < rows) { k[u] = __ldg(src + r * ROWW); a[u] = __ldg(src + r * ROWW + 1); b[u] = __ldg(src + r * ROWW + 2); }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) agg_row(tkey, tsi, tsf, tlen, S, logB, &s_used, o.overflow, k[u], a[u], b[u]);
        }
        __syncthreads();
        if (tid == 0) s_base = (unsigned)atomicAdd(o.cursor, (unsigned long long)s_used);
        __syncthreads();
        agg_flush(tkey, tsi, tsf, tlen, S, &s_used, s_base, o, tid, THREADS);
        __syncthreads();
    }
}

// ---------------------------------------------------------------- pass 2, variant V2: TMA ring, producer warp, warp-granular consumption
// 1024 threads: warps 0..30 consume (each owns K x 32 rows of every chunk), warp 31 issues the bulk copies.
// full[s]: TMA complete_tx; empty[s]: one arrival per consumer warp once its rows of the stage sit in registers.
template <int ROWW, int K, int NST>
__global__ void __launch_bounds__(1024) k_rp_agg_tma(const uint64_t* __restrict__ recs, int64_t cap_rows, const unsigned* __restrict__ cursor, int logB, unsigned S, AggOut o) {
    constexpr int NCW = 31, CR = NCW * 32 * K, NCT = NCW * 32;
    extern __shared__ __align__(128) uint64_t smem[];
    uint64_t* ring = smem;
    uint64_t* tkey = ring + (size_t)NST * CR * ROWW; uint64_t* tsi = tkey + S; double* tsf = reinterpret_cast<double*>(tsi + S); unsigned* tlen = reinterpret_cast<unsigned*>(tsf + S);
    __shared__ uint64_t full[NST], empty[NST];
    __shared__ unsigned s_used, s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, B = 1 << logB;
    if (tid == 0) { for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); } fence_mbar_init(); }
    __syncthreads();
    if (warp == NCW) {                      // producer
        if (lane == 0) {
            unsigned q = 0;
            for (int p = blockIdx.x; p < B; p += gridDim.x) {
                const int64_t rows = min((int64_t)cursor[p], cap_rows);
                const uint64_t* src = recs + (size_t)p * cap_rows * ROWW;
                const int nch = (int)((rows + CR - 1) / CR);
                for (int c = 0; c < nch; c++, q++) {
                    const int st = q % NST; const unsigned use = q / NST;
                    if (use > 0) while (!mbar_try_wait(&empty[st], (use - 1) & 1u)) {}
                    const unsigned bytes = (unsigned)(min((int64_t)CR, rows - (int64_t)c * CR) * ROWW * 8);
                    mbar_expect_tx(&full[st], bytes);
                    bulk_g2s(ring + (size_t)st * CR * ROWW, src + (size_t)c * CR * ROWW, bytes, &full[st]);
                }
            }
        }
        return;
    }
    unsigned q = 0;
    for (int p = blockIdx.x; p < B; p += gridDim.x) {
        for (unsigned i = tid; i < S; i += NCT) { tkey[i] = EMPTY; tsi[i] = 0; tsf[i] = 0.0; tlen[i] = 0; }
        if (tid == 0) s_used = 0;
        named_bar_sync(1, NCT);
        const int64_t rows = min((int64_t)cursor[p], cap_rows);
        const int nch = (int)((rows + CR - 1) / CR);
        for (int c = 0; c < nch; c++, q++) {
            const int st = q % NST; const unsigned par = (q / NST) & 1u;
            while (!mbar_try_wait(&full[st], par)) {}
            const uint64_t* buf = ring + (size_t)st * CR * ROWW;
            const int crow = (int)min((int64_t)CR, rows - (int64_t)c * CR);
            uint64_t k[K], a[K], b[K];
#pragma unroll
            for (int u = 0; u < K; u++) {
                const int r = (warp * K + u) * 32 + lane;
                k[u] = EMPTY; a[u] = 0; b[u] = 0;
                if (r < crow) { k[u] = buf[r * ROWW]; a[u] = buf[r * ROWW + 1]; b[u] = buf[r * ROWW + 2]; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
#pragma unroll
            for (int u = 0; u < K; u++) agg_row(tkey, tsi, tsf, tlen, S, logB, &s_used, o.overflow, k[u], a[u], b[u]);
        }
        named_bar_sync(1, NCT);
        if (tid == 0) s_base = (unsigned)atomicAdd(o.cursor, (unsigned long long)s_used);
        named_bar_sync(1, NCT);
        agg_flush(tkey, tsi, tsf, tlen, S, &s_used, s_base, o, tid, NCT);
        named_bar_sync(1, NCT);
    }
}

template <typename KFN>
static float time_agg(KFN kfn, int threads, size_t smem, int B, int sms, int* occ_out, unsigned long long* out_cursor, const uint64_t* recs, int64_t cap_rows, const unsigned* cursor, int logB, unsigned S, AggOut o) {
    if (smem > 227 * 1024) { *occ_out = 0; return -1.f; }
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, threads, smem));
    *occ_out = occ;
    if (occ < 1) return -1.f;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(out_cursor, 0, 8));
        CK(cudaEventRecord(e0));
        kfn<<<std::min(B, sms * occ), threads, smem>>>(recs, cap_rows, cursor, logB, S, o);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

template <int ROWW, int RPT, int THREADS, bool BULK>
static float run_scatter(const uint64_t* key, const uint64_t* v1, const uint64_t* v2, int64_t n, int logB, uint64_t* out, int64_t cap_rows, unsigned* cursor, int* overflow, int sms, int* ctas_per_sm) {
    const int B = 1 << logB, T = RPT * THREADS;
    const size_t smem = (size_t)(T + (BULK ? B : 0)) * ROWW * 8 + (size_t)3 * B * 4 + (BULK ? 0 : (size_t)T * 2);
    auto* kfn = k_rp_scatter<ROWW, RPT, THREADS, BULK>;
    if (smem > 227 * 1024) { *ctas_per_sm = 0; return -1.f; }
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, THREADS, smem));
    if (occ < 1) { printf("scatter config does not fit (smem %zu)\n", smem); return -1.f; }
    *ctas_per_sm = occ;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(cursor, 0, B * 4)); CK(cudaMemset(overflow, 0, 4));
        CK(cudaEventRecord(e0));
        kfn<<<sms * occ, THREADS, smem>>>(key, v1, v2, n, logB, out, cap_rows, cursor, overflow);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}
template <int ROWW, int SLOTS, int CR, int NST, int THREADS>
static float run_agg(const uint64_t* recs, int64_t cap_rows, const unsigned* cursor, int logB, uint64_t* out_key, uint64_t* out_si, double* out_sf, unsigned* out_len,
                     unsigned long long* out_cursor, int* overflow, int sms, int* ctas_per_sm) {
    const size_t smem = (size_t)NST * CR * ROWW * 8 + (size_t)SLOTS * 28;
    auto* kfn = k_rp_agg<ROWW, SLOTS, CR, NST, THREADS>;
    if (smem > 227 * 1024) { *ctas_per_sm = 0; return -1.f; }
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, THREADS, smem));
    if (occ < 1) { printf("agg config does not fit (smem %zu)\n", smem); return -1.f; }
    *ctas_per_sm = occ;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    const int B = 1 << logB;
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(out_cursor, 0, 8));
        CK(cudaEventRecord(e0));
        kfn<<<min(B, sms * occ), THREADS, smem>>>(recs, cap_rows, cursor, logB, out_key, out_si, out_sf, out_len, out_cursor, overflow);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 100000000LL;
    const uint64_t keys = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1000000ULL;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint64_t *key, *kid, *recs, *out_key, *out_si, *key_of_id; int64_t* vi; double *vf, *sf, *out_sf; unsigned long long *si, *out_cursor; unsigned *len, *cursor, *out_len; int* overflow;
    CK(cudaMalloc(&key, n * 8)); CK(cudaMalloc(&kid, n * 8)); CK(cudaMalloc(&vi, n * 8)); CK(cudaMalloc(&vf, n * 8));
    CK(cudaMalloc(&si, keys * 8)); CK(cudaMalloc(&sf, keys * 8)); CK(cudaMalloc(&len, keys * 4)); CK(cudaMalloc(&key_of_id, keys * 8));
    const int64_t rec_rows = (int64_t)((double)n * 1.30) + (1 << 20);
    CK(cudaMalloc(&recs, (size_t)rec_rows * 24));
    CK(cudaMalloc(&cursor, 4096 * 4)); CK(cudaMalloc(&overflow, 4)); CK(cudaMalloc(&out_cursor, 8));
    CK(cudaMalloc(&out_key, keys * 8)); CK(cudaMalloc(&out_si, keys * 8)); CK(cudaMalloc(&out_sf, keys * 8)); CK(cudaMalloc(&out_len, keys * 4));
    k_gen<<<sms * 8, 256>>>(key, vi, vf, n, keys); k_keyid<<<sms * 8, 256>>>(key, kid, n, keys); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_base = 1e9f;
    for (int it = 0; it < 3; it++) {
        CK(cudaMemset(si, 0, keys * 8)); CK(cudaMemset(sf, 0, keys * 8)); CK(cudaMemset(len, 0, keys * 4));
        CK(cudaEventRecord(e0));
        k_base<<<sms * 8, 256>>>(kid, vi, vf, n, si, sf, len);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best_base = fminf(best_base, ms);
    }
    k_mark<<<sms * 8, 256>>>(key, kid, n, key_of_id); CK(cudaDeviceSynchronize());
    printf("{\"rows\": %lld, \"keys\": %llu, \"baseline_l2_red_ms\": %.3f}\n", (long long)n, (unsigned long long)keys, best_base);
    // host copies for the check
    std::vector<uint64_t> h_koi(keys); std::vector<unsigned long long> h_si(keys); std::vector<double> h_sf(keys); std::vector<unsigned> h_len(keys);
    CK(cudaMemcpy(h_koi.data(), key_of_id, keys * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h_si.data(), si, keys * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_sf.data(), sf, keys * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h_len.data(), len, keys * 4, cudaMemcpyDeviceToHost));
    std::vector<std::pair<uint64_t, uint64_t>> sorted(keys);
    for (uint64_t i = 0; i < keys; i++) sorted[i] = {h_koi[i], i};
    std::sort(sorted.begin(), sorted.end());

    auto check = [&](const char* tag) {
        unsigned long long G = 0; int h_over = 0;
        CK(cudaMemcpy(&G, out_cursor, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&h_over, overflow, 4, cudaMemcpyDeviceToHost));
        std::vector<uint64_t> ok(G), osi(G); std::vector<double> osf(G); std::vector<unsigned> ol(G);
        CK(cudaMemcpy(ok.data(), out_key, G * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(osi.data(), out_si, G * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(osf.data(), out_sf, G * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ol.data(), out_len, G * 4, cudaMemcpyDeviceToHost));
        unsigned long long bad = 0;
        for (unsigned long long g = 0; g < G; g++) {
            auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(ok[g], (uint64_t)0));
            if (it == sorted.end() || it->first != ok[g]) { bad++; continue; }
            const uint64_t id = it->second;
            if (osi[g] != h_si[id] || ol[g] != h_len[id] || fabs(osf[g] - h_sf[id]) > 1e-9 * fabs(h_sf[id]) + 1e-9) bad++;
        }
        unsigned long long live = 0; for (uint64_t i = 0; i < keys; i++) live += h_len[i] != 0;
        printf("  check[%s]: groups %llu (expected %llu), mismatching %llu, overflow %d\n", tag, G, live, bad, h_over);
    };

    const uint64_t* k64 = key; const uint64_t* a64 = (const uint64_t*)vi; const uint64_t* b64 = (const uint64_t*)vf;
    AggOut o{out_key, out_si, out_sf, out_len, out_cursor, overflow};
    for (int logB = 8; logB <= 10; logB++) {
        const int B = 1 << logB;
        int64_t cap_rows = (int64_t)((double)n / B * 1.25) + 64; cap_rows &= ~(int64_t)1;
        if (cap_rows * B > rec_rows) { printf("record buffer too small\n"); return 1; }
        int occ = 0;
        float t;
        printf("{\"buckets\": %d", B);
        const double gb = (double)n * 24 / 1e9;
        t = run_scatter<3, 8, 480, true>(k64, a64, b64, n, logB, recs, cap_rows, cursor, overflow, sms, &occ); printf(", \"scatter_bulk_T3840\": [%.3f, %d, %.0f]", t, occ, 2 * gb / (t / 1e3));
        t = run_scatter<3, 6, 512, true>(k64, a64, b64, n, logB, recs, cap_rows, cursor, overflow, sms, &occ); printf(", \"scatter_bulk_T3072\": [%.3f, %d, %.0f]", t, occ, 2 * gb / (t / 1e3));
        t = run_scatter<3, 4, 384, true>(k64, a64, b64, n, logB, recs, cap_rows, cursor, overflow, sms, &occ); printf(", \"scatter_bulk_T1536\": [%.3f, %d, %.0f]", t, occ, 2 * gb / (t / 1e3));
        t = run_scatter<3, 4, 512, true>(k64, a64, b64, n, logB, recs, cap_rows, cursor, overflow, sms, &occ); printf(", \"scatter_bulk_T2048\": [%.3f, %d, %.0f]", t, occ, 2 * gb / (t / 1e3));
        printf("}\n");
        // the record stream left in `recs` is the last variant's (bulk_T2048, padded runs): aggregate it
        const unsigned S = logB == 8 ? 6144u : (logB == 9 ? 3072u : 1536u);
        const size_t tab = (size_t)S * 28;
        float ta;
        ta = time_agg(k_rp_agg_direct<3, 1024, 2>, 1024, tab, B, sms, &occ, out_cursor, recs, cap_rows, cursor, logB, S, o);
        printf("{\"buckets\": %d, \"slots\": %u, \"agg\": \"direct_1024t_u2\", \"agg_ms\": %.3f, \"ctas_per_sm\": %d, \"GBps\": %.0f}\n", B, S, ta, occ, gb / (ta / 1e3)); if (ta > 0) check("direct_1024t_u2");
        ta = time_agg(k_rp_agg_direct<3, 512, 4>, 512, tab, B, sms, &occ, out_cursor, recs, cap_rows, cursor, logB, S, o);
        printf("{\"buckets\": %d, \"slots\": %u, \"agg\": \"direct_512t_u4\", \"agg_ms\": %.3f, \"ctas_per_sm\": %d, \"GBps\": %.0f}\n", B, S, ta, occ, gb / (ta / 1e3)); if (ta > 0) check("direct_512t_u4");
        ta = time_agg(k_rp_agg_tma<3, 1, 2>, 1024, tab + (size_t)2 * 992 * 24, B, sms, &occ, out_cursor, recs, cap_rows, cursor, logB, S, o);
        printf("{\"buckets\": %d, \"slots\": %u, \"agg\": \"tma_K1_x2\", \"agg_ms\": %.3f, \"ctas_per_sm\": %d, \"GBps\": %.0f}\n", B, S, ta, occ, gb / (ta / 1e3)); if (ta > 0) check("tma_K1_x2");
        if (logB >= 9) {
            ta = time_agg(k_rp_agg_tma<3, 2, 2>, 1024, tab + (size_t)2 * 1984 * 24, B, sms, &occ, out_cursor, recs, cap_rows, cursor, logB, S, o);
            printf("{\"buckets\": %d, \"slots\": %u, \"agg\": \"tma_K2_x2\", \"agg_ms\": %.3f, \"ctas_per_sm\": %d, \"GBps\": %.0f}\n", B, S, ta, occ, gb / (ta / 1e3)); if (ta > 0) check("tma_K2_x2");
            ta = time_agg(k_rp_agg_tma<3, 1, 4>, 1024, tab + (size_t)4 * 992 * 24, B, sms, &occ, out_cursor, recs, cap_rows, cursor, logB, S, o);
            printf("{\"buckets\": %d, \"slots\": %u, \"agg\": \"tma_K1_x4\", \"agg_ms\": %.3f, \"ctas_per_sm\": %d, \"GBps\": %.0f}\n", B, S, ta, occ, gb / (ta / 1e3)); if (ta > 0) check("tma_K1_x4");
        }
    }
    return 0;
}
