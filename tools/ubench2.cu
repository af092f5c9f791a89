// tools/ubench2.cu — round-2 micro-benchmarks for the group_by accumulate step (not part of the product):
// how many table updates per second does the part retire when the three accumulator updates of a row are issued as
//   A  key load + 3 RED (u32 len, u64 sum, f64 sum; word-major planes)            — the shipped K5
//   B  key load + 2 RED (u64 sum with len packed in the high bits, f64 sum)        — "packed" plan
//   C  key load + 1 bulk reduce of 16 B (cp.reduce.async.bulk .add.u64 on {sum,len}) + 1 RED f64
//   D  key load + 1 bulk reduce of 32 B (.add.u64 on {sum, len, fixed-point hi, fixed-point lo})
//   E  bulk reduce only, 16 / 32 B                                                 — TMA small-op issue rate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ void red_u64(uint64_t* p, uint64_t v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_u32(unsigned* p, unsigned v) { asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_f64(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }

__global__ void k_A(const uint64_t* keys, unsigned* len, uint64_t* si, double* sf, uint64_t mask, int64_t nops) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        if (__ldcg(keys + slot) == 0xdeadbeefULL) continue;
        red_u32(len + slot, 1u); red_u64(si + slot, (uint64_t)i); red_f64(sf + slot, 1.5);
    }
}
__global__ void k_B(const uint64_t* keys, uint64_t* si, double* sf, uint64_t mask, int64_t nops) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        if (__ldcg(keys + slot) == 0xdeadbeefULL) continue;
        red_u64(si + slot, ((uint64_t)i & 0xffff) + (1ull << 37)); red_f64(sf + slot, 1.5);
    }
}
// one RED only + key load (lower bound of any L2-atomic plan that keeps the probe)
__global__ void k_B1(const uint64_t* keys, uint64_t* si, uint64_t mask, int64_t nops) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        if (__ldcg(keys + slot) == 0xdeadbeefULL) continue;
        red_u64(si + slot, (uint64_t)i);
    }
}

// Bulk reduce: every thread owns RING staging slots of BYTES in shared memory; one cp.reduce.async.bulk per row.
// MODE 0: bulk only; 1: key load + bulk; 2: key load + bulk + RED f64
template <int BYTES, int MODE, int RING>
__global__ void k_bulk(const uint64_t* keys, uint64_t* entries, double* sf, uint64_t mask, int64_t nops) {
    extern __shared__ __align__(128) uint64_t sm[];
    constexpr int W = BYTES / 8;
    uint64_t* mine = sm + (size_t)threadIdx.x * W * RING;
    int r = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        if (MODE >= 1) { if (__ldcg(keys + slot) == 0xdeadbeefULL) continue; }
        // wait until the slot we are about to overwrite has been read by the TMA unit
        asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(RING - 1) : "memory");
        uint64_t* s = mine + r * W;
        s[0] = (uint64_t)i; s[1] = 1ull;
        if (W == 4) { s[2] = (uint64_t)i >> 7; s[3] = (uint64_t)i & 127; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" :: "l"(entries + slot * W), "r"(sa), "n"(BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (MODE == 2) red_f64(sf + slot, 1.5);
        r = (r + 1 == RING) ? 0 : r + 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// warp-batched: lanes stage their records, then lane 0..31 each still issue their own op but the commit/wait is per 4 ops
template <int BYTES>
__global__ void k_bulk_b4(uint64_t* entries, uint64_t mask, int64_t nops) {
    extern __shared__ __align__(128) uint64_t sm[];
    constexpr int W = BYTES / 8;
    uint64_t* mine = sm + (size_t)threadIdx.x * W * 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += 4 * stride) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 4; u++) { uint64_t* s = mine + u * W; s[0] = (uint64_t)i + u; s[1] = 1ull; if (W == 4) { s[2] = 3; s[3] = 4; } }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (i + u * stride >= nops) break;
            uint64_t slot = mix((uint64_t)(i + u * stride)) & mask;
            uint32_t sa = (uint32_t)__cvta_generic_to_shared(mine + u * W);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" :: "l"(entries + slot * W), "r"(sa), "n"(BYTES) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// closer to K5: 2 rows per thread per iteration from streamed 128-bit loads of three 8-byte columns, single staging ring
// (wait_group.read 0), grid-stride with `g` CTAs per SM.  MODE 0: 3 REDs; 1: bulk {len,sum} + RED f64; 2: bulk only; 3: bulk, RED issued before the fence
template <int MODE>
__global__ void __launch_bounds__(256) k_real(const ulonglong2* __restrict__ kcol, const ulonglong2* __restrict__ icol, const ulonglong2* __restrict__ fcol, const uint64_t* keys, uint64_t* pair, unsigned* len, uint64_t* si, double* sf, uint64_t mask, int64_t npairs) {
    extern __shared__ __align__(128) uint64_t sm[];
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
        const ulonglong2 k2 = __ldcs(kcol + p), i2 = __ldcs(icol + p), f2 = __ldcs(fcol + p);
        const uint64_t kk[2] = {k2.x, k2.y}, iv[2] = {i2.x, i2.y}; const double fv[2] = {__longlong_as_double((long long)f2.x), __longlong_as_double((long long)f2.y)};
        uint64_t slot[2], k0[2];
#pragma unroll
        for (int r = 0; r < 2; r++) { slot[r] = mix(kk[r]) & mask; k0[r] = __ldcg(keys + slot[r]); }
        if (MODE >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (MODE == 3) {
#pragma unroll
            for (int r = 0; r < 2; r++) if (k0[r] != 0xdeadbeefULL) red_f64(sf + slot[r], fv[r]);
        }
        if (MODE >= 1) {
#pragma unroll
            for (int r = 0; r < 2; r++) { uint64_t* c = sm + 2 * (r * 256 + threadIdx.x); c[0] = 1; c[1] = iv[r]; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (k0[r] == 0xdeadbeefULL) continue;
                uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm + 2 * (r * 256 + threadIdx.x));
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], 16;" :: "l"(pair + 2 * slot[r]), "r"(sa) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            if (k0[r] == 0xdeadbeefULL) continue;
            if (MODE == 0) { red_u32(len + slot[r], 1u); red_u64(si + slot[r], iv[r]); }
            if (MODE <= 1) red_f64(sf + slot[r], fv[r]);
        }
    }
    if (MODE >= 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> float timeit(F f, int reps = 4) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
    const int64_t nops = argc > 1 ? atoll(argv[1]) : 100000000;
    const int lg = 21; const uint64_t n = 1ull << lg, mask = n - 1;
    uint64_t *keys, *si, *ent; unsigned* len; double* sf;
    CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&si, n * 8)); CK(cudaMalloc(&sf, n * 8)); CK(cudaMalloc(&len, n * 4)); CK(cudaMalloc(&ent, n * 32));
    CK(cudaMemset(keys, 0, n * 8)); CK(cudaMemset(si, 0, n * 8)); CK(cudaMemset(sf, 0, n * 8)); CK(cudaMemset(len, 0, n * 4)); CK(cudaMemset(ent, 0, n * 32));
    const int sms = p.multiProcessorCount;
    for (int g : {8, 16}) {
        const int grid = sms * g, block = 256;
        float a = timeit([&] { k_A<<<grid, block>>>(keys, len, si, sf, mask, nops); });
        float b = timeit([&] { k_B<<<grid, block>>>(keys, si, sf, mask, nops); });
        float b1 = timeit([&] { k_B1<<<grid, block>>>(keys, si, mask, nops); });
        printf("{\"test\": \"red_plans\", \"ctas_per_sm\": %d, \"A_key_3red_Grows\": %.2f, \"B_key_2red_Grows\": %.2f, \"B1_key_1red_Grows\": %.2f}\n", g, nops / a / 1e6, nops / b / 1e6, nops / b1 / 1e6);
    }
    // correctness probe of the bulk reduce: sum of word 1 over all entries must equal nops
    auto check = [&](int W, const char* name) {
        uint64_t* h = (uint64_t*)malloc(n * W * 8); CK(cudaMemcpy(h, ent, n * W * 8, cudaMemcpyDeviceToHost));
        uint64_t tot = 0; for (uint64_t i = 0; i < n; i++) tot += h[i * W + 1]; free(h);
        printf("{\"test\": \"bulk_check\", \"kernel\": \"%s\", \"len_total\": %llu, \"expected\": %lld}\n", name, (unsigned long long)tot, (long long)nops);
    };
#define RUN_BULK(BYTES, MODE, RING, BLOCK, G) do { \
        auto kern = k_bulk<BYTES, MODE, RING>; const int smem = BLOCK * BYTES * RING; \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        CK(cudaMemset(ent, 0, n * 32)); kern<<<sms * G, BLOCK, smem>>>(keys, ent, sf, mask, nops); CK(cudaDeviceSynchronize()); check(BYTES / 8, "k_bulk<" #BYTES "," #MODE "," #RING ">"); \
        float ms = timeit([&] { kern<<<sms * G, BLOCK, smem>>>(keys, ent, sf, mask, nops); }); \
        printf("{\"test\": \"bulk_reduce\", \"bytes\": %d, \"mode\": %d, \"ring\": %d, \"block\": %d, \"ctas_per_sm\": %d, \"Grows\": %.2f, \"ms\": %.3f}\n", BYTES, MODE, RING, BLOCK, G, nops / ms / 1e6, ms); \
    } while (0)
    RUN_BULK(16, 0, 2, 256, 4);
    RUN_BULK(16, 0, 4, 256, 4);
    RUN_BULK(16, 0, 4, 256, 8);
    RUN_BULK(16, 0, 8, 128, 8);
    RUN_BULK(32, 0, 2, 256, 4);
    RUN_BULK(32, 0, 4, 256, 4);
    RUN_BULK(32, 0, 4, 256, 6);
    RUN_BULK(32, 0, 8, 128, 6);
    RUN_BULK(16, 1, 4, 256, 8);
    RUN_BULK(32, 1, 4, 256, 6);
    RUN_BULK(16, 2, 4, 256, 8);
    {
        auto kern = k_bulk_b4<32>; const int smem = 256 * 32 * 4;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        float ms = timeit([&] { kern<<<sms * 6, 256, smem>>>(ent, mask, nops); });
        printf("{\"test\": \"bulk_reduce_batched4\", \"bytes\": 32, \"Grows\": %.2f, \"ms\": %.3f}\n", nops / ms / 1e6, ms);
        auto kern16 = k_bulk_b4<16>; const int smem16 = 256 * 16 * 4;
        ms = timeit([&] { kern16<<<sms * 8, 256, smem16>>>(ent, mask, nops); });
        printf("{\"test\": \"bulk_reduce_batched4\", \"bytes\": 16, \"Grows\": %.2f, \"ms\": %.3f}\n", nops / ms / 1e6, ms);
    }
    {
        const int64_t npairs = nops / 2;
        ulonglong2 *kc, *ic, *fc; CK(cudaMalloc(&kc, npairs * 16)); CK(cudaMalloc(&ic, npairs * 16)); CK(cudaMalloc(&fc, npairs * 16));
        CK(cudaMemset(ic, 1, npairs * 16)); CK(cudaMemset(fc, 0x3f, npairs * 16));
        {   // keys = row index (mix() makes the slots random)
            uint64_t* h = (uint64_t*)malloc(nops * 8); for (int64_t i = 0; i < nops; i++) h[i] = (uint64_t)i * 0x9E3779B97F4A7C15ull; CK(cudaMemcpy(kc, h, nops * 8, cudaMemcpyHostToDevice)); free(h);
        }
        for (int g : {4, 8}) {
            float m0 = timeit([&] { k_real<0><<<sms * g, 256, 8192>>>(kc, ic, fc, keys, ent, len, si, sf, mask, npairs); });
            float m1 = timeit([&] { k_real<1><<<sms * g, 256, 8192>>>(kc, ic, fc, keys, ent, len, si, sf, mask, npairs); });
            float m2 = timeit([&] { k_real<2><<<sms * g, 256, 8192>>>(kc, ic, fc, keys, ent, len, si, sf, mask, npairs); });
            float m3 = timeit([&] { k_real<3><<<sms * g, 256, 8192>>>(kc, ic, fc, keys, ent, len, si, sf, mask, npairs); });
            printf("{\"test\": \"k5_like\", \"ctas_per_sm\": %d, \"three_red_ms\": %.3f, \"bulk_plus_red_ms\": %.3f, \"bulk_only_ms\": %.3f, \"red_before_fence_ms\": %.3f}\n", g, m0, m1, m2, m3);
        }
    }
    return 0;
}
