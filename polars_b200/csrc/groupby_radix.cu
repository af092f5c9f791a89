// groupby_radix.cu — K5r: the partitioned group_by plan for tables that cannot stay resident in L2.
//
// Reference path being replaced: the same as groupby.cu (group_by_threaded_slice, polars-core/src/frame/group_by/
// hashing.rs:116-167, which ALSO partitions the keys by hash_to_partition before it builds one table per partition, and
// agg_sum/mean/min/max, aggregations/mod.rs:486-1018).  groupby.cu aggregates with L2 atomics into one open-addressing
// table; past ~2.5e6 groups that table leaves L2 and every RED becomes a DRAM read-modify-write (1e7 groups: 19 ms per
// 1e8 rows, profiles/README.md).  This plan moves the rows instead of the atomics:
//   pass 0  k_gbr_hist      rows per bucket (bucket = top bits of table_hash(key): the slot function of the L2 plan)
//   pass 1  k_gbr_scatter   every CTA sorts a 2048-row tile by bucket in shared memory as ROW-MAJOR records
//                           [key, v0, v1, ...] and writes each (tile, bucket) run with ONE cp.async.bulk (TMA)
//                           shared->global copy (runs padded to an even record count with a GB_EMPTY-key record so both
//                           ends stay 16-byte aligned); 74 % of the measured copy peak at 256 buckets
//                           (profiles/r02_proto_radix.md).  Past 512 buckets the runs shrink to single records and the
//                           tile is written with coalesced 8-byte stores instead (no padding).
//   pass 2  k_gbr_agg       one CTA per bucket: the bucket's record stream is staged into shared memory by
//                           cp.async.bulk (TMA) global->shared copies on an mbarrier ring (a producer warp issues,
//                           31 consumer warps each release their slice of a stage on an `empty` mbarrier as soon as it
//                           sits in registers); rows aggregate into a shared-memory open-addressing table; a bucket's
//                           groups are final, so they leave compacted straight into the dense output arrays — the
//                           global table, its initialisation and its extraction disappear.
// Measured on C2 (1e6 groups) the plan loses to the L2 plan (2.3 vs 1.97 ms: 64-bit shared-memory atomics on random
// slots cost 3.9 SM-cycles per row), so it is only taken when the L2 plan's table would exceed the L2 budget.
// Restrictions (anything else stays on the L2 plan): no validity bitmaps, no first-row tracking, <= 4 value columns,
// no heavy hitters in the sample.  Exact for any input: bucket capacities come from the exact histogram; a bucket with
// more groups than its shared-memory table holds raises the status flag and the caller redoes the batch on the L2 plan.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "dev_utils.cuh"
#include "groupby.h"
#include "groupby_dev.cuh"

namespace plb {

constexpr int GBR_THREADS = 512;      // pass 1: 2048-row tiles (bulk stores) / 4096-row tiles (many buckets: longer runs per bucket)
constexpr int GBR_MAX_LOGB = 13;
constexpr int GBR_NCW = 31, GBR_NCT = GBR_NCW * 32;                                    // pass 2: 31 consumer warps + 1 producer warp

struct GbRadixDev {
    uint64_t* recs;                 // record streams, bucket b at recs + off[b] * roww
    const unsigned long long* off;  // first record of bucket b (even)
    unsigned* cursor;               // records written to bucket b so far (pads included)
    unsigned* counts;               // pass 0: rows of bucket b
    int logB, roww;
    uint64_t* special;              // accumulator row of the GB_EMPTY-key group: [len, words...] (RED target; rare rows only)
    int* status;
};

template <int KEY_ELEM> __device__ __forceinline__ void gbr_load_pair(const void* col, int64_t r0, int64_t n, uint64_t& a, uint64_t& b) {
    a = 0; b = 0;
    if (KEY_ELEM == 8) {
        if (r0 + 1 < n) { const ulonglong2 t = ld_stream_u64x2(reinterpret_cast<const uint64_t*>(col) + r0); a = t.x; b = t.y; }
        else if (r0 < n) a = reinterpret_cast<const uint64_t*>(col)[r0];
    } else {
        if (r0 + 1 < n) { const uint2 t = ld_stream_u32x2(reinterpret_cast<const uint32_t*>(col) + r0); a = t.x; b = t.y; }
        else if (r0 < n) a = reinterpret_cast<const uint32_t*>(col)[r0];
    }
}
__device__ __forceinline__ void gbr_load_pair_rt(const void* col, int elem, int64_t r0, int64_t n, uint64_t& a, uint64_t& b) {
    if (elem == 8) gbr_load_pair<8>(col, r0, n, a, b); else gbr_load_pair<4>(col, r0, n, a, b);
}

// ---------------------------------------------------------------------------- pass 0: exact bucket sizes
template <int KEY_ELEM, int KEY_CANON>
__global__ void __launch_bounds__(512) k_gbr_hist(const void* __restrict__ keys, int64_t n, int logB, unsigned* __restrict__ counts) {
    extern __shared__ unsigned h_s[];
    const int B = 1 << logB;
    for (int i = threadIdx.x; i < B; i += blockDim.x) h_s[i] = 0;
    __syncthreads();
    const int64_t npairs = (n + 1) >> 1;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
        uint64_t k0, k1;
        gbr_load_pair<KEY_ELEM>(keys, 2 * p, n, k0, k1);
        k0 = canon_key<KEY_CANON>(k0); k1 = canon_key<KEY_CANON>(k1);
        if (k0 != GB_EMPTY) atomicAdd(&h_s[(unsigned)(table_hash(k0) >> (64 - logB))], 1u);
        if (2 * p + 1 < n && k1 != GB_EMPTY) atomicAdd(&h_s[(unsigned)(table_hash(k1) >> (64 - logB))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x) if (h_s[i]) atomicAdd(&counts[i], h_s[i]);
}
// bucket capacities (worst-case padding: one pad record per tile that touches the bucket) -> exclusive offsets, one CTA
__global__ void __launch_bounds__(1024) k_gbr_offsets(const unsigned* __restrict__ counts, int B, unsigned long long ntiles, int pad, unsigned long long* __restrict__ off, unsigned* __restrict__ cursor) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    for (int base = 0; base < B; base += 1024) {
        const int i = base + threadIdx.x;
        unsigned long long c = 0;
        if (i < B) { c = counts[i]; if (pad) c += c < ntiles ? c : ntiles; c = (c + 1ull) & ~1ull; cursor[i] = 0; }      // even: every stream starts 16-byte aligned
        unsigned long long x = c;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long s = wsum[lane], t = s;
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= (unsigned)o) t += y; }
            wsum[lane] = t - s;
        }
        __syncthreads();
        const unsigned long long incl = carry_s + wsum[warp] + x;
        if (i < B) off[i] = incl - c;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[B] = carry_s;
}

// ---------------------------------------------------------------------------- pass 1: tile sort + TMA bulk stores
__device__ __forceinline__ void gbr_apply_special(const GbLayout& L, const GbBatch& Bt, uint64_t* sp, const uint64_t* raw) {
    atomicAdd(reinterpret_cast<unsigned long long*>(sp), 1ull);
    for (int c = 0; c < L.n_cols; c++)
        for (int k = L.col_kbegin[c]; k < L.col_kbegin[c + 1]; k++) {
            uint64_t* a = sp + 1 + L.wslot[k];
            const int dt = Bt.cols[c].dtype;
            switch (L.wop[k]) {
                case W_ADD_INT: atomicAdd(reinterpret_cast<unsigned long long*>(a), (unsigned long long)raw_to_int(dt, raw[c])); break;
                case W_ADD_F64: atomicAdd(reinterpret_cast<double*>(a), raw_to_f64(dt, raw[c])); break;
                case W_MIN_S64: atomicMin(reinterpret_cast<long long*>(a), (long long)raw_to_int(dt, raw[c])); break;
                case W_MAX_S64: atomicMax(reinterpret_cast<long long*>(a), (long long)raw_to_int(dt, raw[c])); break;
                case W_MIN_U64: atomicMin(reinterpret_cast<unsigned long long*>(a), (unsigned long long)raw[c]); break;
                case W_MAX_U64: atomicMax(reinterpret_cast<unsigned long long*>(a), (unsigned long long)raw[c]); break;
                case W_MIN_F64: { const double f = raw_to_f64(dt, raw[c]); if (f == f) atomicMin(reinterpret_cast<unsigned long long*>(a), (unsigned long long)f64_to_ordered(f)); break; }
                case W_MAX_F64: { const double f = raw_to_f64(dt, raw[c]); if (f == f) atomicMax(reinterpret_cast<unsigned long long*>(a), (unsigned long long)f64_to_ordered(f)); break; }
                default: break;
            }
        }
}

template <int ROWW, int KEY_ELEM, int KEY_CANON, bool BULK, int RPT>
__global__ void __launch_bounds__(GBR_THREADS) k_gbr_scatter(const __grid_constant__ GbLayout L, const __grid_constant__ GbBatch Bt, const __grid_constant__ GbRadixDev R) {
    constexpr int T = GBR_THREADS * RPT, THREADS = GBR_THREADS, NC = ROWW - 1;
    const int logB = R.logB, B = 1 << logB;
    extern __shared__ __align__(16) uint64_t gbr_smem[];
    uint64_t* stage = gbr_smem;                                        // (T + (BULK ? B : 0)) records
    unsigned* hist = reinterpret_cast<unsigned*>(stage + (size_t)(T + (BULK ? B : 0)) * ROWW);
    unsigned* start = hist + B;
    unsigned* gpos = start + B;
    uint16_t* sp = reinterpret_cast<uint16_t*>(gpos + B);             // !BULK: bucket of every sorted slot
    __shared__ unsigned warp_tot[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = Bt.n;
    const int64_t ntiles = (n + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * T;
        for (int p = tid; p < B; p += THREADS) hist[p] = 0;
        __syncthreads();
        uint64_t k[RPT]; unsigned pk[RPT];      // pk: bucket << 16 | rank in the tile's run;  ~0 = no row;  ~0 - 1 = GB_EMPTY-key row
#pragma unroll
        for (int j = 0; j < RPT / 2; j++) gbr_load_pair<KEY_ELEM>(Bt.keys, base + 2 * (int64_t)(j * THREADS + tid), n, k[2 * j], k[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < RPT; j++) {
            const int64_t r = base + 2 * (int64_t)((j >> 1) * THREADS + tid) + (j & 1);
            pk[j] = 0xFFFFFFFFu;
            if (r < n) {
                k[j] = canon_key<KEY_CANON>(k[j]);
                if (k[j] == GB_EMPTY) pk[j] = 0xFFFFFFFEu;
                else { const unsigned b = (unsigned)(table_hash(k[j]) >> (64 - logB)); pk[j] = (b << 16) | atomicAdd(&hist[b], 1u); }
            }
        }
        __syncthreads();
        // exclusive scan of the (padded) run lengths; one global reservation per non-empty (tile, bucket)
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
                gpos[p] = cp ? atomicAdd(&R.cursor[p], cp) : 0u;
                run += cp;
            }
        }
        if (BULK) bulk_wait_read0();        // the previous tile's copies have finished reading the staging buffer
        __syncthreads();
        if (BULK) for (int p = tid; p < B; p += THREADS) { const unsigned c = hist[p]; if (c & 1u) { uint64_t* pad = stage + (size_t)(start[p] + c) * ROWW; pad[0] = GB_EMPTY;
#pragma unroll
                                                                                                      for (int w = 1; w < ROWW; w++) pad[w] = 0; } }
        // place the records (order inside a run is arbitrary)
#pragma unroll
        for (int j = 0; j < RPT / 2; j++) {
            const int64_t r0 = base + 2 * (int64_t)(j * THREADS + tid);
            uint64_t v[NC > 0 ? NC : 1][2];
#pragma unroll
            for (int c = 0; c < NC; c++) gbr_load_pair_rt(Bt.cols[c].values, Bt.cols[c].elem, r0, n, v[c][0], v[c][1]);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const unsigned q = pk[2 * j + e];
                if (q == 0xFFFFFFFFu) continue;
                if (q == 0xFFFFFFFEu) {      // the GB_EMPTY key is the pad marker of the record streams: its (rare) rows aggregate right here
                    uint64_t raw[NC > 0 ? NC : 1];
#pragma unroll
                    for (int c = 0; c < NC; c++) raw[c] = v[c][e];
                    gbr_apply_special(L, Bt, R.special, raw);
                    continue;
                }
                const unsigned bkt = q >> 16, pos = start[bkt] + (q & 0xFFFFu);
                uint64_t* rec = stage + (size_t)pos * ROWW;
                rec[0] = k[2 * j + e];
#pragma unroll
                for (int c = 0; c < NC; c++) rec[1 + c] = v[c][e];
                if (!BULK) sp[pos] = (uint16_t)bkt;
            }
        }
        if (BULK) fence_async_smem();
        __syncthreads();
        if (BULK) {
            for (int p = tid; p < B; p += THREADS) {
                const unsigned c = hist[p], cp = (c + 1u) & ~1u;
                if (cp) bulk_s2g(R.recs + (R.off[p] + gpos[p]) * ROWW, stage + (size_t)start[p] * ROWW, cp * ROWW * 8);
            }
            bulk_commit();
        } else {
            const unsigned total = start[B - 1] + hist[B - 1];
            for (unsigned w = tid; w < total * ROWW; w += THREADS) {
                const unsigned row = w / ROWW, c = w - row * ROWW;
                const unsigned p = sp[row];
                R.recs[(R.off[p] + gpos[p] + (row - start[p])) * ROWW + c] = stage[w];
            }
            __syncthreads();
        }
    }
    if (BULK) bulk_wait0();
}

// ---------------------------------------------------------------------------- pass 2: TMA ring -> shared-memory table -> dense output
struct GbDenseDev { uint64_t* keys; uint32_t* first; uint32_t* len; uint64_t* words; int64_t Gb; unsigned long long* cursor; };

template <int ROWW, int K, int NST>
__global__ void __launch_bounds__(1024) k_gbr_agg(const __grid_constant__ GbLayout L, const __grid_constant__ GbBatch Bt, const __grid_constant__ GbRadixDev R, const __grid_constant__ GbDenseDev D, unsigned S) {
    constexpr int CR = GBR_NCW * 32 * K, NC = ROWW - 1;
    extern __shared__ __align__(128) uint64_t gbr_smem2[];
    uint64_t* ring = gbr_smem2;                                       // NST x CR records
    uint64_t* tkey = ring + (size_t)NST * CR * ROWW;                  // S keys
    uint64_t* tacc = tkey + S;                                        // n_words planes of S
    unsigned* tlen = reinterpret_cast<unsigned*>(tacc + (size_t)L.n_words * S);
    __shared__ uint64_t full[NST], empty[NST];
    __shared__ unsigned s_used, s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, B = 1 << R.logB;
    if (tid == 0) { for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], GBR_NCW); } mbar_fence_init(); }
    __syncthreads();
    if (warp == GBR_NCW) {                      // producer warp: keeps the ring full across bucket boundaries
        if (lane == 0) {
            unsigned q = 0;
            for (int p = blockIdx.x; p < B; p += gridDim.x) {
                const int64_t rows = (int64_t)R.cursor[p];
                const uint64_t* src = R.recs + R.off[p] * ROWW;
                const int nch = (int)((rows + CR - 1) / CR);
                for (int c = 0; c < nch; c++, q++) {
                    const int st = q % NST; const unsigned use = q / NST;
                    if (use > 0) while (!mbar_try_wait(&empty[st], (use - 1) & 1u)) {}
                    const int64_t crow = min((int64_t)CR, rows - (int64_t)c * CR);
                    const unsigned bytes = (unsigned)(((crow + 1) & ~(int64_t)1) * ROWW * 8);      // whole 16-byte units (the buffer is padded)
                    mbar_expect_tx(&full[st], bytes);
                    bulk_g2s(ring + (size_t)st * CR * ROWW, src + (size_t)c * CR * ROWW, bytes, &full[st]);
                }
            }
        }
        return;
    }
    const unsigned max_used = S - (S >> 2);
    unsigned q = 0;
    for (int p = blockIdx.x; p < B; p += gridDim.x) {
        for (unsigned i = tid; i < S; i += GBR_NCT) { tkey[i] = GB_EMPTY; tlen[i] = 0; for (int w = 0; w < L.n_words; w++) tacc[(size_t)w * S + i] = L.init[w]; }
        if (tid == 0) s_used = 0;
        named_bar_sync(1, GBR_NCT);
        const int64_t rows = (int64_t)R.cursor[p];
        const int nch = (int)((rows + CR - 1) / CR);
        for (int c = 0; c < nch; c++, q++) {
            const int st = q % NST; const unsigned par = (q / NST) & 1u;
            while (!mbar_try_wait(&full[st], par)) {}
            const uint64_t* buf = ring + (size_t)st * CR * ROWW;
            const int crow = (int)min((int64_t)CR, rows - (int64_t)c * CR);
            uint64_t rec[K][ROWW];
#pragma unroll
            for (int u = 0; u < K; u++) {
                const int r = (warp * K + u) * 32 + lane;
                rec[u][0] = GB_EMPTY;
                if (r < crow) {
#pragma unroll
                    for (int w = 0; w < ROWW; w++) rec[u][w] = buf[r * ROWW + w];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
#pragma unroll
            for (int u = 0; u < K; u++) {
                const uint64_t key = rec[u][0];
                if (key == GB_EMPTY) continue;              // pad record
                unsigned slot = __umulhi((unsigned)((table_hash(key) << R.logB) >> 32), S);
                bool found = false;
                for (unsigned probes = 0; probes < 128u; probes++) {
                    const uint64_t cur = *reinterpret_cast<volatile uint64_t*>(tkey + slot);
                    if (cur == key) { found = true; break; }
                    if (cur == GB_EMPTY) {
                        if (*reinterpret_cast<volatile unsigned*>(&s_used) >= max_used) break;
                        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(tkey + slot), (unsigned long long)GB_EMPTY, (unsigned long long)key);
                        if (old == GB_EMPTY) { atomicAdd(&s_used, 1u); found = true; break; }
                        if (old == key) { found = true; break; }
                    }
                    if (++slot == S) slot = 0;
                }
                if (!found) { *R.status = 1; continue; }    // more groups in this bucket than its table holds: the caller falls back
                if (L.need_len) atomicAdd(tlen + slot, 1u);
#pragma unroll
                for (int cix = 0; cix < NC; cix++) {
                    const int dt = Bt.cols[cix].dtype;
                    for (int kk = L.col_kbegin[cix]; kk < L.col_kbegin[cix + 1]; kk++) gb_apply_smem<false>(L.wop[kk], tacc + (size_t)L.wslot[kk] * S + slot, dt, rec[u][1 + cix], true);
                }
            }
        }
        named_bar_sync(1, GBR_NCT);
        if (tid == 0) s_base = (unsigned)atomicAdd(D.cursor, (unsigned long long)s_used);
        named_bar_sync(1, GBR_NCT);
        // the bucket's groups are final: compact them straight into the dense output (order inside a bucket = slot order)
        const unsigned long long base = s_base;
        for (unsigned i0 = 0; i0 < S; i0 += GBR_NCT) {
            const unsigned i = i0 + tid;
            const bool used = i < S && tkey[i] != GB_EMPTY;
            unsigned at = 0;
            if (used) at = atomicSub(&s_used, 1u) - 1u;
            if (used) {
                const unsigned long long pos = base + at;
                if ((int64_t)pos >= D.Gb) *R.status = 1;
                else {
                    D.keys[pos] = tkey[i]; D.len[pos] = tlen[i]; D.first[pos] = 0xFFFFFFFFu;
                    for (int w = 0; w < L.n_words; w++) D.words[(int64_t)w * D.Gb + pos] = tacc[(size_t)w * S + i];
                }
            }
        }
        named_bar_sync(1, GBR_NCT);
    }
}

// the GB_EMPTY-key group (if any row carried that key) joins the dense output
__global__ void k_gbr_append_special(const uint64_t* __restrict__ special, int n_words, GbDenseDev D, int* status) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || special[0] == 0) return;
    const unsigned long long pos = atomicAdd(D.cursor, 1ull);
    if ((int64_t)pos >= D.Gb) { *status = 1; return; }
    D.keys[pos] = GB_EMPTY; D.len[pos] = (uint32_t)special[0]; D.first[pos] = 0xFFFFFFFFu;
    for (int w = 0; w < n_words; w++) D.words[(int64_t)w * D.Gb + pos] = special[1 + w];
}

// =============================================================================================
// Host side
// =============================================================================================
constexpr int GBR_RPT_BULK = 4, GBR_RPT_PLAIN = 4;      // 8 rows per thread (4096-row tiles, 1 CTA / SM) measured slower: 2.93 vs 2.71 ms at 4096 buckets
static int64_t gbr_tile_rows(bool bulk) { return (int64_t)GBR_THREADS * (bulk ? GBR_RPT_BULK : GBR_RPT_PLAIN); }
template <int ROWW, int KEY_ELEM, int KEY_CANON>
static void launch_scatter(const GbLayout& L, const GbBatch& Bt, const GbRadixDev& R, bool bulk) {
    const int B = 1 << R.logB;
    const size_t tile = (size_t)gbr_tile_rows(bulk);
    const size_t smem = (tile + (bulk ? B : 0)) * ROWW * 8 + (size_t)3 * B * 4 + (bulk ? 0 : tile * 2);
    int occ = 0;
    if (bulk) {
        auto kfn = k_gbr_scatter<ROWW, KEY_ELEM, KEY_CANON, true, GBR_RPT_BULK>;
        PLB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, GBR_THREADS, smem));
        PLB_LAUNCH("k5r_scatter", kfn, ctx().sm_count * std::max(occ, 1), GBR_THREADS, smem, L, Bt, R);
    } else {
        auto kfn = k_gbr_scatter<ROWW, KEY_ELEM, KEY_CANON, false, GBR_RPT_PLAIN>;
        PLB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, GBR_THREADS, smem));
        PLB_LAUNCH("k5r_scatter", kfn, ctx().sm_count * std::max(occ, 1), GBR_THREADS, smem, L, Bt, R);
    }
}
template <int ROWW>
static void launch_agg(const GbLayout& L, const GbBatch& Bt, const GbRadixDev& R, const GbDenseDev& D, unsigned S, int stages) {
    const size_t table = (size_t)S * (8 + 4 + 8 * L.n_words);
    const int B = 1 << R.logB;
    auto go = [&](auto kfn, int k, int nst) {
        const size_t smem = (size_t)nst * GBR_NCW * 32 * k * ROWW * 8 + table + 16;
        PLB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        PLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, 1024, smem));
        PLB_LAUNCH("k5r_aggregate", kfn, std::min(B, ctx().sm_count * std::max(occ, 1)), 1024, smem, L, Bt, R, D, S);
    };
    if (stages >= 4) go(k_gbr_agg<ROWW, 1, 4>, 1, 4); else go(k_gbr_agg<ROWW, 1, 2>, 1, 2);
}

// Returns false when the plan does not apply (or gave up): nothing is left behind and the caller runs the L2 plan.
bool GroupByState::consume_radix(const DevCol& key, const std::vector<const DevCol*>& values, uint64_t planned_cap) {
    const int mode = [] { const char* e = getenv("BL_K5_RADIX"); return e ? atoi(e) : 1; }();      // 0 never, 1 when the table would leave L2, 2 whenever eligible (read per call)
    if (mode == 0 || L.need_first || key.validity != nullptr || hot.rows > 0 || key.len < (1 << 20)) return false;
    // column binding (same rule as launch_batch: aggregations over one buffer share a record word)
    GbBatch Bt; memset(&Bt, 0, sizeof Bt);
    Bt.keys = key.v(); Bt.n = key.len; Bt.key_dtype = key.dtype;
    std::vector<const void*> col_ptr; std::vector<int> col_of_agg(plans.size(), -1);
    for (size_t i = 0; i < plans.size(); i++) {
        if (plans[i].kind == BL_AGG_LEN) continue;
        const DevCol* v = values[i];
        if (v == nullptr || v->len != key.len || v->dtype != plans[i].in_dtype || v->validity != nullptr) return false;
        if (dtype_size(v->dtype) != 4 && dtype_size(v->dtype) != 8) return false;
        int c = -1;
        for (size_t j = 0; j < col_ptr.size(); j++) if (col_ptr[j] == v->v()) c = (int)j;
        if (c < 0) {
            if (col_ptr.size() >= 4) return false;
            c = (int)col_ptr.size(); col_ptr.push_back(v->v());
            Bt.cols[c].values = v->v(); Bt.cols[c].validity = nullptr; Bt.cols[c].dtype = v->dtype; Bt.cols[c].elem = dtype_size(v->dtype);
        }
        col_of_agg[i] = c;
    }
    GbLayout Lb = L;
    Lb.n_cols = (int)col_ptr.size();
    int kk = 0;
    for (int c = 0; c < Lb.n_cols; c++) {
        Lb.col_kbegin[c] = kk;
        for (size_t i = 0; i < plans.size(); i++) {
            if (col_of_agg[i] != c) continue;
            if (plans[i].main >= 0) { Lb.wslot[kk] = plans[i].main; Lb.wop[kk] = L.slot_op[plans[i].main]; kk++; }
        }
    }
    for (int c = Lb.n_cols; c <= GB_MAX_COLS; c++) Lb.col_kbegin[c] = kk;
    const int roww = 1 + Lb.n_cols;
    const int64_t n = key.len;
    // shared-memory table per bucket: what is left of ~110 KB (2 CTAs / SM: measured best, profiles/r02_proto_radix.md) after a
    // 2-stage ring; when even 8192 buckets of that size cannot take the estimated groups, one CTA / SM with a ~200 KB table
    const size_t entry = 8 + 4 + 8 * (size_t)L.n_words;
    const size_t ring2 = (size_t)2 * GBR_NCW * 32 * roww * 8;
    unsigned S = 0; int logB = 6;
    for (const size_t total_kb : {(size_t)110, (size_t)222}) {
        if (total_kb * 1024 < ring2 + 1024 + 512 * entry) continue;
        S = (unsigned)((total_kb * 1024 - ring2 - 1024) / entry) & ~31u;
        const double per_bucket = 0.55 * (double)S;                   // groups per bucket the table takes comfortably
        logB = 6;
        while (logB < GBR_MAX_LOGB && (double)est_groups / (double)(1 << logB) > per_bucket) logB++;
        if ((double)est_groups / (double)(1 << logB) <= 0.7 * (double)S) break;
        S = 0;
    }
    if (S == 0) return false;                                         // too many groups even for 8192 buckets of the large table
    if (mode == 1) {
        // the L2 plan fills a table of up to 2x its L2 budget in two slot-range passes at full speed (launch_batch: pass_bits);
        // beyond that it needs four passes or misses L2.  Measured on C2-shaped rows (profiles/r02_k5r_cardinality_sweep_v3.txt):
        // 3e6 keys 3.6 (L2 plan, 134 MB table) vs 4.3 ms; 4e6 keys 6.6 (268 MB) vs 4.4 ms; 1e7 keys 15.7 vs 6.1 ms
        const double l2_budget = 0.55 * (double)ctx().l2_bytes;
        if ((double)planned_cap * L.stride * 8 <= 2.0 * l2_budget) return false;
    }
    const int B = 1 << logB;
    const bool bulk = logB <= 9;
    const int64_t ntiles = (n + gbr_tile_rows(bulk) - 1) / gbr_tile_rows(bulk);
    const int64_t rec_rows = n + (bulk ? std::min<int64_t>(n, (int64_t)B * ntiles) : 0) + 2 * B + 16;
    const int elem = dtype_size(key.dtype);
    const int canon = key.dtype == BL_FLOAT64 ? 1 : (key.dtype == BL_FLOAT32 ? 2 : 0);
    DevPtr recs, ctl;
    try {
        recs = dev_alloc((size_t)rec_rows * roww * 8);
        ctl = dev_alloc((size_t)B * 4 * 2 + (size_t)(B + 1) * 8 + (size_t)(1 + GB_MAX_WORDS) * 8 + 64);
    } catch (const Error&) { cudaGetLastError(); return false; }       // not enough free HBM for the record streams
    GbRadixDev R; memset(&R, 0, sizeof R);
    R.recs = as<uint64_t>(recs); R.logB = logB; R.roww = roww; R.status = as<int>(status);
    char* cp = reinterpret_cast<char*>(ctl->p);
    R.off = reinterpret_cast<unsigned long long*>(cp); cp += (size_t)(B + 1) * 8;
    R.special = reinterpret_cast<uint64_t*>(cp); cp += (size_t)(1 + GB_MAX_WORDS) * 8;
    R.counts = reinterpret_cast<unsigned*>(cp); cp += (size_t)B * 4;
    R.cursor = reinterpret_cast<unsigned*>(cp);
    dev_memset(ctl->p, 0, ctl->bytes);
    {   // identities of the special accumulator row
        uint64_t h[1 + GB_MAX_WORDS]; h[0] = 0; for (int w = 0; w < L.n_words; w++) h[1 + w] = L.init[w];
        PLB_CUDA(cudaMemcpyAsync(R.special, h, (size_t)(1 + L.n_words) * 8, cudaMemcpyHostToDevice, ctx().stream));
        PLB_CUDA(cudaStreamSynchronize(ctx().stream));      // h lives on this frame
    }
    dev_memset(status->p, 0, 4);
    const int hgrid = ctx().sm_count * 4;
#define GBR_KEYS(CALL) do { if (elem == 8) { if (canon == 1) { CALL(8, 1); } else { CALL(8, 0); } } else { if (canon == 2) { CALL(4, 2); } else { CALL(4, 0); } } } while (0)
#define HIST(E, CN) PLB_LAUNCH("k5r_histogram", (k_gbr_hist<E, CN>), hgrid, 512, (size_t)B * 4, key.v(), n, logB, R.counts)
    GBR_KEYS(HIST);
#undef HIST
    PLB_LAUNCH("k5r_offsets", k_gbr_offsets, 1, 1024, 0, R.counts, B, (unsigned long long)ntiles, bulk ? 1 : 0, const_cast<unsigned long long*>(R.off), R.cursor);
#define SCAT(E, CN) do { switch (roww) { case 1: launch_scatter<1, E, CN>(Lb, Bt, R, bulk); break; case 2: launch_scatter<2, E, CN>(Lb, Bt, R, bulk); break; case 3: launch_scatter<3, E, CN>(Lb, Bt, R, bulk); break; \
                                          case 4: launch_scatter<4, E, CN>(Lb, Bt, R, bulk); break; default: launch_scatter<5, E, CN>(Lb, Bt, R, bulk); break; } } while (0)
    GBR_KEYS(SCAT);
#undef SCAT
#undef GBR_KEYS
    // dense output, sized by a generous bound on the group count (overflow -> status -> fall back)
    const int64_t Gb = std::max<int64_t>(1024, std::min<int64_t>(n + 1, 3 * est_groups + (1 << 16)));
    dense.keys = dev_alloc((size_t)Gb * 8); dense.first = dev_alloc((size_t)Gb * 4); dense.len = dev_alloc((size_t)Gb * 4);
    dense.words = dev_alloc((size_t)Gb * 8 * std::max(L.n_words, 1)); dense.ctl = dev_alloc(16); dense.Gb = Gb;
    const long long ctl_init[2] = {0, -1};
    PLB_CUDA(cudaMemcpyAsync(dense.ctl->p, ctl_init, 16, cudaMemcpyHostToDevice, ctx().stream));
    GbDenseDev D{as<uint64_t>(dense.keys), as<uint32_t>(dense.first), as<uint32_t>(dense.len), as<uint64_t>(dense.words), Gb, as<unsigned long long>(dense.ctl)};
    switch (roww) {
        case 1: launch_agg<1>(Lb, Bt, R, D, S, 2); break; case 2: launch_agg<2>(Lb, Bt, R, D, S, 2); break; case 3: launch_agg<3>(Lb, Bt, R, D, S, 2); break;
        case 4: launch_agg<4>(Lb, Bt, R, D, S, 2); break; default: launch_agg<5>(Lb, Bt, R, D, S, 2); break;
    }
    PLB_LAUNCH("k5r_special", k_gbr_append_special, 1, 32, 0, R.special, L.n_words, D, as<int>(status));
    const int st = read_scalar(as<int>(status));      // also orders ctl_init / recs lifetimes
    if (getenv("BL_K5_DEBUG")) fprintf(stderr, "[k5r] rows=%lld est_groups=%lld buckets=%d slots=%u bulk=%d status=%d\n", (long long)n, (long long)est_groups, B, S, (int)bulk, st);
    if (st != 0) { dense = GbDense{}; dev_memset(status->p, 0, 4); return false; }
    dense.ready = true;
    rows_seen = n;
    return true;
}

}  // namespace plb
