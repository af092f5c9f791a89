// tools/ubench.cu — B200 micro-benchmarks that size the hot-path design (not part of the product):
// streaming bandwidth, L2 atomic (RED) throughput on random entries, shared-memory atomics, random
// sector reads.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

__global__ void k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) __stcs(out + i, __ldcs(in + i));
}
__global__ void k_read(const uint4* __restrict__ in, int64_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * s < n; i += 4 * s) { uint4 a = __ldcs(in + i), b = __ldcs(in + i + s), c = __ldcs(in + i + 2 * s), d = __ldcs(in + i + 3 * s); acc += a.x + b.y + c.z + d.w; }
    for (; i < n; i += s) acc += __ldcs(in + i).x;
    if (acc == 0x1234567) *out = acc;
}
// mode 0: one RED.ADD.U64 per op on a dense u64 array; 1: entry stride 32 B, one RED; 2: key load + 3 REDs (u32 add, u64 add, f64 add) in one 32-B entry
template <int MODE>
__global__ void k_red(uint64_t* table, uint64_t mask, int64_t nops, int stride_words) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        uint64_t* e = table + slot * stride_words;
        if (MODE == 0 || MODE == 1) atomicAdd((unsigned long long*)e, 1ull);
        else {
            unsigned long long k = __ldcg((const unsigned long long*)e);
            if (k == 0xdeadbeefULL) continue;
            atomicAdd((unsigned*)(e + 1), 1u);
            atomicAdd((unsigned long long*)(e + 2), (unsigned long long)i);
            atomicAdd((double*)(e + 3), 1.5);
        }
    }
}
// SoA variant of mode 2: key load from keys[slot], REDs into three separate dense arrays
__global__ void k_red_soa(const uint64_t* keys, unsigned* len, uint64_t* si, double* sf, uint64_t mask, int64_t nops) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        unsigned long long k = __ldcg((const unsigned long long*)(keys + slot));
        if (k == 0xdeadbeefULL) continue;
        atomicAdd(len + slot, 1u);
        atomicAdd((unsigned long long*)(si + slot), (unsigned long long)i);
        atomicAdd(sf + slot, 1.5);
    }
}
__global__ void k_smem_atom(int64_t nops, int use64, unsigned long long* out) {
    __shared__ unsigned long long t[4096];   // 32 KB
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) t[i] = 0;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t h = mix((uint64_t)i);
        if (use64 == 1) atomicAdd(&t[h & 4095], 1ull);
        else if (use64 == 2) atomicAdd((double*)&t[h & 4095], 1.0);
        else atomicAdd(((unsigned*)t) + (h & 8191), 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0 && t[0] == 0x123456789ULL) *out = 1;
}
template <typename V>
__global__ void k_rand_load(const V* table, uint64_t mask, int64_t nops, unsigned long long* out) {
    unsigned long long acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nops; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t slot = mix((uint64_t)i) & mask;
        V v = __ldg(table + slot);
        acc += *reinterpret_cast<unsigned*>(&v);
    }
    if (acc == 0x1234567) *out = acc;
}

template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"l2_mb\": %.1f}\n", p.name, p.multiProcessorCount, p.l2CacheSize / 1048576.0);
    const int grid = p.multiProcessorCount * 8, block = 256;
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    {   // streaming
        const int64_t n16 = (int64_t)1 << 26;   // 1 GiB
        uint4 *a, *b; CK(cudaMalloc(&a, n16 * 16)); CK(cudaMalloc(&b, n16 * 16)); CK(cudaMemset(a, 1, n16 * 16));
        float ms = timeit([&] { k_copy<<<grid, block>>>(a, b, n16); });
        printf("{\"test\": \"copy_1GiB\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, 2.0 * n16 * 16 / ms / 1e6);
        ms = timeit([&] { k_read<<<grid, block>>>(a, n16, sink); });
        printf("{\"test\": \"read_1GiB\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, 1.0 * n16 * 16 / ms / 1e6);
        for (int g : {2, 4, 16}) { ms = timeit([&] { k_read<<<p.multiProcessorCount * g, block>>>(a, n16, sink); }); printf("{\"test\": \"read_1GiB_grid%dxSM\", \"ms\": %.4f, \"GBps\": %.1f}\n", g, ms, 1.0 * n16 * 16 / ms / 1e6); }
        cudaFree(a); cudaFree(b);
    }
    const int64_t nops = 100000000;
    for (int lg : {10, 16, 20, 21, 22, 24, 26}) {   // entries
        uint64_t n = 1ull << lg; uint64_t* t; CK(cudaMalloc(&t, n * 32)); CK(cudaMemset(t, 0, n * 32));
        float ms0 = timeit([&] { k_red<0><<<grid, block>>>(t, n - 1, nops, 1); });
        float ms1 = timeit([&] { k_red<1><<<grid, block>>>(t, n - 1, nops, 4); });
        float ms2 = timeit([&] { k_red<2><<<grid, block>>>(t, n - 1, nops, 4); });
        printf("{\"test\": \"red_random\", \"log2_entries\": %d, \"dense_u64_Gops\": %.2f, \"stride32_u64_Gops\": %.2f, \"keyload_plus_3red_Grows\": %.2f, \"table_MB_stride32\": %.1f}\n",
               lg, nops / ms0 / 1e6, nops / ms1 / 1e6, nops / ms2 / 1e6, n * 32 / 1048576.0);
        cudaFree(t);
    }
    for (int lg : {20, 21, 22}) {
        uint64_t n = 1ull << lg; uint64_t *keys, *si; unsigned* len; double* sf;
        CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&si, n * 8)); CK(cudaMalloc(&sf, n * 8)); CK(cudaMalloc(&len, n * 4));
        CK(cudaMemset(keys, 0, n * 8)); CK(cudaMemset(si, 0, n * 8)); CK(cudaMemset(sf, 0, n * 8)); CK(cudaMemset(len, 0, n * 4));
        float ms = timeit([&] { k_red_soa<<<grid, block>>>(keys, len, si, sf, n - 1, nops); });
        printf("{\"test\": \"keyload_plus_3red_SoA\", \"log2_entries\": %d, \"Grows\": %.2f, \"total_MB\": %.1f}\n", lg, nops / ms / 1e6, n * 28 / 1048576.0);
        cudaFree(keys); cudaFree(si); cudaFree(sf); cudaFree(len);
    }
    {
        float ms = timeit([&] { k_smem_atom<<<grid, block>>>(nops, 0, sink); });
        float ms64 = timeit([&] { k_smem_atom<<<grid, block>>>(nops, 1, sink); });
        float msf = timeit([&] { k_smem_atom<<<grid, block>>>(nops, 2, sink); });
        printf("{\"test\": \"smem_atomics_random_32KB\", \"u32_Gops\": %.2f, \"u64_cas_Gops\": %.2f, \"f64_cas_Gops\": %.2f}\n", nops / ms / 1e6, nops / ms64 / 1e6, nops / msf / 1e6);
    }
    for (int lg : {21, 23, 24, 25, 26}) {   // 16-byte entries: 32 MB .. 1 GB
        uint64_t n = 1ull << lg; uint4* t; CK(cudaMalloc(&t, n * 16)); CK(cudaMemset(t, 1, n * 16));
        float ms = timeit([&] { k_rand_load<uint4><<<grid, block>>>(t, n - 1, nops, sink); });
        float ms8 = timeit([&] { k_rand_load<uint2><<<grid, block>>>((const uint2*)t, 2 * n - 1, nops, sink); });
        printf("{\"test\": \"random_load\", \"table_MB\": %.0f, \"load16B_Gops\": %.2f, \"load8B_Gops\": %.2f}\n", n * 16 / 1048576.0, nops / ms / 1e6, nops / ms8 / 1e6);
        cudaFree(t);
    }
    return 0;
}
