// groupby_dev.cuh — device helpers shared by the group_by kernels (groupby.cu: L2-resident table plans; groupby_radix.cu:
// partitioned plan).  Semantics cited in groupby.cu's header.
#pragma once
#include "common.cuh"
#include "dev_utils.cuh"
#include "groupby.h"

namespace plb {

__host__ __device__ inline uint64_t word_identity(int op) {
    switch (op) {
        case W_MIN_S64: return 0x7FFFFFFFFFFFFFFFULL;
        case W_MAX_S64: return 0x8000000000000000ULL;
        case W_MIN_U64: case W_MIN_F64: return 0xFFFFFFFFFFFFFFFFULL;
        default: return 0;   // adds, MAX_U64, MAX_F64
    }
}


template <int KEY_CANON> __device__ __forceinline__ uint64_t canon_key(uint64_t raw) {
    if (KEY_CANON == 1) return canonical_f64_bits(__longlong_as_double((long long)raw));
    if (KEY_CANON == 2) return canonical_f32_bits(__uint_as_float((uint32_t)raw));
    return raw;
}
__device__ __forceinline__ uint64_t load_key_rt(const void* keys, int dtype, int64_t row) {
    switch (dtype) {
        case BL_INT64: case BL_UINT64: return reinterpret_cast<const uint64_t*>(keys)[row];
        case BL_FLOAT64: return canonical_f64_bits(reinterpret_cast<const double*>(keys)[row]);
        case BL_FLOAT32: return canonical_f32_bits(reinterpret_cast<const float*>(keys)[row]);
        default: return (uint64_t)reinterpret_cast<const uint32_t*>(keys)[row];   // i32/u32 bit pattern, zero-extended
    }
}


__device__ __forceinline__ double raw_to_f64(int dtype, uint64_t raw) {
    switch (dtype) {
        case BL_INT64: return (double)(long long)raw;
        case BL_UINT64: return (double)(unsigned long long)raw;
        case BL_INT32: return (double)(int)(uint32_t)raw;
        case BL_UINT32: return (double)(uint32_t)raw;
        case BL_FLOAT64: return __longlong_as_double((long long)raw);
        default: return (double)__uint_as_float((uint32_t)raw);
    }
}
__device__ __forceinline__ uint64_t raw_to_int(int dtype, uint64_t raw) {
    return dtype == BL_INT32 ? (uint64_t)(long long)(int)(uint32_t)raw : raw;   // sign-extend i32; u32 already zero-extended
}


// ---- shared-memory accumulators: 32-bit native ATOMS; 64-bit integer adds as two 32-bit adds with carry (exact,
//      order-free); f64 add and 64-bit min/max are CAS loops (ATOMS.CAST.SPIN.64).
__device__ __forceinline__ void s_add_u64(uint64_t* a, uint64_t v) {
    uint32_t* w = reinterpret_cast<uint32_t*>(a);
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    const uint32_t old = atomicAdd(w, lo);
    const uint32_t up = hi + (((uint32_t)(old + lo) < old) ? 1u : 0u);
    if (up) atomicAdd(w + 1, up);
}
// FAST: every value column is 8 bytes wide and has no validity bitmap (the common analytic case): the
// dtype dispatch collapses to one select and the null checks disappear (this kernel is issue-bound).
template <bool FAST>
__device__ __forceinline__ void gb_apply_smem(int op, uint64_t* addr, int dtype, uint64_t raw, bool valid) {
    switch (op) {
        case W_ADD_INT: { uint64_t v = FAST ? raw : raw_to_int(dtype, raw); if (valid && v) s_add_u64(addr, v); break; }
        case W_ADD_F64: { double f = FAST ? (dtype == BL_FLOAT64 ? __longlong_as_double((long long)raw) : (dtype == BL_INT64 ? (double)(long long)raw : (double)(unsigned long long)raw)) : raw_to_f64(dtype, raw);
                          if (valid && f != 0.0) atomicAdd(reinterpret_cast<double*>(addr), f); break; }
        case W_MIN_S64: if (valid) atomicMin(reinterpret_cast<long long*>(addr), (long long)raw_to_int(dtype, raw)); break;
        case W_MAX_S64: if (valid) atomicMax(reinterpret_cast<long long*>(addr), (long long)raw_to_int(dtype, raw)); break;
        case W_MIN_U64: if (valid) atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)raw); break;
        case W_MAX_U64: if (valid) atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)raw); break;
        case W_MIN_F64: { double f = raw_to_f64(dtype, raw); if (valid && f == f) atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)f64_to_ordered(f)); break; }
        case W_MAX_F64: { double f = raw_to_f64(dtype, raw); if (valid && f == f) atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)f64_to_ordered(f)); break; }
        default: if (!valid) atomicAdd(reinterpret_cast<unsigned*>(addr), 1u); break;   // W_NULLCNT (< 2^32 per CTA)
    }
}


}  // namespace plb
