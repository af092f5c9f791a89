// elementwise.cu — K1 (arithmetic) and K2 (comparison -> bitmask) plus bitmap utilities.
//
// Reference semantics restated (paths relative to /root/reference/crates):
//   ArithmeticKernel / prim_binary_values   polars-compute/src/arithmetic/mod.rs:8-76, arity.rs:47-127
//   integer rules                            polars-compute/src/arithmetic/signed.rs:23-232
//   float rules (scalar div = mul by 1/c)    polars-compute/src/arithmetic/float.rs:21-122
//   TotalOrd/TotalEq comparison kernels      polars-compute/src/comparisons/mod.rs:4-76,
//                                            polars-utils/src/total_ord.rs:317-364
// Both kernels are pure streaming: 128-bit loads/stores, 4 independent vectors in flight per
// thread, grid = SMs x 8 resident CTAs of 256 threads.  Bound: HBM (8*n_in + 8 B/row for K1,
// 8 + 1/8 B/row for K2).
#include "common.cuh"
#include "dev_utils.cuh"

namespace plb {

enum { MODE_AA = 0, MODE_AS = 1, MODE_SA = 2 };

// ---------------------------------------------------------------------------- K1
template <typename T, int OP, int MODE, bool IS_FLOAT> struct ArithFn {
    T s_inv;   // 1/scalar for float scalar-rhs division forms
    __device__ __forceinline__ T operator()(T a, T b) const {
        if constexpr (IS_FLOAT) {
            if constexpr (OP == BL_OP_ADD) return a + b;
            else if constexpr (OP == BL_OP_SUB) return MODE == MODE_AS ? a + (-b) : a - b;
            else if constexpr (OP == BL_OP_MUL) return a * b;
            else if constexpr (OP == BL_OP_FLOOR_DIV) return MODE == MODE_AS ? floor(a * s_inv) : floor(a / b);
            else if constexpr (OP == BL_OP_MOD) return MODE == MODE_AS ? a - b * floor(a * s_inv) : a - b * floor(a / b);
            else return MODE == MODE_AS ? a * s_inv : a / b;
        } else {
            using U = typename make_unsigned_t<T>::type;
            if constexpr (OP == BL_OP_ADD) return (T)((U)a + (U)b);
            else if constexpr (OP == BL_OP_SUB) return (T)((U)a - (U)b);
            else if constexpr (OP == BL_OP_MUL) return (T)((U)a * (U)b);
            else if constexpr (OP == BL_OP_FLOOR_DIV) { T d, m; floor_divmod<T>(a, b, d, m); return d; }
            else { T d, m; floor_divmod<T>(a, b, d, m); return m; }
        }
    }
};

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); T v[N]; };

template <typename T, int OP, int MODE, bool IS_FLOAT>
__global__ void __launch_bounds__(256) k_arith(const T* __restrict__ lhs, const T* __restrict__ rhs, T* __restrict__ out, int64_t n, T scalar) {
    constexpr int VN = 16 / sizeof(T);
    constexpr int UNROLL = 4;
    ArithFn<T, OP, MODE, IS_FLOAT> f;
    f.s_inv = IS_FLOAT ? (T)1 / scalar : (T)0;
    const int64_t nvec = n / VN;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    using V = Vec16<T>;
    for (; i + (UNROLL - 1) * stride < nvec; i += UNROLL * stride) {
        uint4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (MODE != MODE_SA) a[u] = ld_stream_u32x4(reinterpret_cast<const uint4*>(lhs) + i + u * stride);
            if (MODE != MODE_AS) b[u] = ld_stream_u32x4(reinterpret_cast<const uint4*>(rhs) + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            V va, vb, vo;
            if (MODE != MODE_SA) va = *reinterpret_cast<V*>(&a[u]);
            if (MODE != MODE_AS) vb = *reinterpret_cast<V*>(&b[u]);
#pragma unroll
            for (int k = 0; k < VN; k++) vo.v[k] = f(MODE == MODE_SA ? scalar : va.v[k], MODE == MODE_AS ? scalar : vb.v[k]);
            st_stream_u32x4(reinterpret_cast<uint4*>(out) + i + u * stride, *reinterpret_cast<uint4*>(&vo));
        }
    }
    for (; i < nvec; i += stride) {
        V va, vb, vo;
        if (MODE != MODE_SA) { uint4 t = ld_stream_u32x4(reinterpret_cast<const uint4*>(lhs) + i); va = *reinterpret_cast<V*>(&t); }
        if (MODE != MODE_AS) { uint4 t = ld_stream_u32x4(reinterpret_cast<const uint4*>(rhs) + i); vb = *reinterpret_cast<V*>(&t); }
#pragma unroll
        for (int k = 0; k < VN; k++) vo.v[k] = f(MODE == MODE_SA ? scalar : va.v[k], MODE == MODE_AS ? scalar : vb.v[k]);
        st_stream_u32x4(reinterpret_cast<uint4*>(out) + i, *reinterpret_cast<uint4*>(&vo));
    }
    // scalar tail (< VN elements)
    int64_t t = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = f(MODE == MODE_SA ? scalar : lhs[t], MODE == MODE_AS ? scalar : rhs[t]);
}

// integer true division -> f64 (signed.rs:216-228: a as f64 / b as f64; scalar rhs: x * (1.0 / rhs))
template <typename T, int MODE>
__global__ void __launch_bounds__(256) k_int_true_div(const T* __restrict__ lhs, const T* __restrict__ rhs, double* __restrict__ out, int64_t n, T scalar) {
    double inv = 1.0 / (double)scalar;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (MODE == MODE_AS) out[i] = (double)lhs[i] * inv;
        else if (MODE == MODE_SA) out[i] = (double)scalar / (double)rhs[i];
        else out[i] = (double)lhs[i] / (double)rhs[i];
    }
}

template <typename T, int OP, bool IS_FLOAT>
static void launch_arith_mode(int mode, const T* l, const T* r, T* o, int64_t n, T scalar) {
    int grid = grid_for(n / (16 / sizeof(T)) / 4 + 1, 256);
    if (mode == MODE_AA) PLB_LAUNCH("k1_arith", (k_arith<T, OP, MODE_AA, IS_FLOAT>), grid, 256, 0, l, r, o, n, scalar);
    else if (mode == MODE_AS) PLB_LAUNCH("k1_arith", (k_arith<T, OP, MODE_AS, IS_FLOAT>), grid, 256, 0, l, r, o, n, scalar);
    else PLB_LAUNCH("k1_arith", (k_arith<T, OP, MODE_SA, IS_FLOAT>), grid, 256, 0, l, r, o, n, scalar);
}
template <typename T, bool IS_FLOAT>
static void launch_arith(int op, int mode, const T* l, const T* r, T* o, int64_t n, T scalar) {
    switch (op) {
        case BL_OP_ADD: launch_arith_mode<T, BL_OP_ADD, IS_FLOAT>(mode, l, r, o, n, scalar); break;
        case BL_OP_SUB: launch_arith_mode<T, BL_OP_SUB, IS_FLOAT>(mode, l, r, o, n, scalar); break;
        case BL_OP_MUL: launch_arith_mode<T, BL_OP_MUL, IS_FLOAT>(mode, l, r, o, n, scalar); break;
        case BL_OP_FLOOR_DIV: launch_arith_mode<T, BL_OP_FLOOR_DIV, IS_FLOAT>(mode, l, r, o, n, scalar); break;
        case BL_OP_MOD: launch_arith_mode<T, BL_OP_MOD, IS_FLOAT>(mode, l, r, o, n, scalar); break;
        default:
            if constexpr (IS_FLOAT) launch_arith_mode<T, BL_OP_TRUE_DIV, true>(mode, l, r, o, n, scalar);
            break;
    }
}

// ---------------------------------------------------------------------------- bitmaps
__global__ void k_bitmap_and3(const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out, int64_t nwords) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w = 0xFFFFFFFFu;
        if (a) w &= a[i];
        if (b) w &= b[i];
        if (c) w &= c[i];
        out[i] = w;
    }
}
DevPtr bitmap_and(const uint32_t* a, const uint32_t* b, const uint32_t* c, int64_t bits) {
    if (!a && !b && !c) return nullptr;
    DevPtr out = dev_alloc(bitmap_bytes(bits) + 16);
    int64_t nw = (bits + 31) / 32;
    if (nw) PLB_LAUNCH("bitmap_and", k_bitmap_and3, grid_for(nw, 256), 256, 0, a, b, c, as<uint32_t>(out), nw);
    return out;
}
__global__ void k_bitmap_popcount(const uint32_t* bm, int64_t bits, unsigned long long* out) {
    int64_t nw = (bits + 31) / 32;
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w = bm[i];
        if (i == nw - 1 && (bits & 31)) w &= (1u << (bits & 31)) - 1u;
        c += __popc(w);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane_id() == 0 && c) atomicAdd(out, c);
}
int64_t bitmap_popcount(const uint32_t* bm, int64_t bits) {
    if (bits == 0) return 0;
    DevPtr d = dev_alloc(8); dev_memset(d->p, 0, 8);
    PLB_LAUNCH("bitmap_popcount", k_bitmap_popcount, grid_for((bits + 31) / 32, 256), 256, 0, bm, bits, as<unsigned long long>(d));
    return (int64_t)read_scalar(as<unsigned long long>(d));
}
// fill a bitmap with `bits` ones / zeros
__global__ void k_fill_u32(uint32_t* p, uint32_t v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_iota_u32(uint32_t* p, int64_t n, uint32_t base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = base + (uint32_t)i;
}
void iota_u32(uint32_t* p, int64_t n, uint32_t base) { if (n) PLB_LAUNCH("iota", k_iota_u32, grid_for(n, 256), 256, 0, p, n, base); }

// ---------------------------------------------------------------------------- K2
// total-order predicates (total_ord.rs:317-364): NaN == NaN, NaN is the greatest value
template <typename T> __device__ __forceinline__ bool tot_ge(T a, T b) { return a >= b; }
template <> __device__ __forceinline__ bool tot_ge<double>(double a, double b) { return (a != a) | (a >= b); }
template <> __device__ __forceinline__ bool tot_ge<float>(float a, float b) { return (a != a) | (a >= b); }
template <typename T> __device__ __forceinline__ bool tot_eq(T a, T b) { return a == b; }
template <> __device__ __forceinline__ bool tot_eq<double>(double a, double b) { return (a != a) ? (b != b) : (a == b); }
template <> __device__ __forceinline__ bool tot_eq<float>(float a, float b) { return (a != a) ? (b != b) : (a == b); }
template <typename T, int OP> __device__ __forceinline__ bool cmp_op(T a, T b) {
    if constexpr (OP == BL_CMP_EQ) return tot_eq<T>(a, b);
    else if constexpr (OP == BL_CMP_NE) return !tot_eq<T>(a, b);
    else if constexpr (OP == BL_CMP_LT) return !tot_ge<T>(a, b);
    else if constexpr (OP == BL_CMP_LE) return tot_ge<T>(b, a);
    else if constexpr (OP == BL_CMP_GT) return !tot_ge<T>(b, a);
    else return tot_ge<T>(a, b);
}

// One warp-step covers 32 vectors of 16 bytes = 32*VN rows; lane l owns rows [VN*l, VN*l+VN).
// VN ballots give VN interleaved 32-bit masks; spread_bits() re-interleaves them into row order.
// 64-bit types: VN = 2 -> one 64-bit mask word per warp-step.  32-bit types: VN = 4 -> 128 bits.
template <typename T, int OP, bool SCALAR>
__global__ void __launch_bounds__(256) k_compare(const T* __restrict__ lhs, const T* __restrict__ rhs, uint32_t* __restrict__ out, int64_t n, T scalar) {
    constexpr int VN = 16 / sizeof(T);
    constexpr int ROWS_PER_STEP = 32 * VN;
    using V = Vec16<T>;
    const int64_t nsteps = n / ROWS_PER_STEP;                       // full warp-steps
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lane = lane_id();
    constexpr int UNROLL = 4;      // independent 128-bit loads in flight per thread
    auto emit = [&](int64_t s, const V& va, const V& vb) {
        uint32_t m[VN];
#pragma unroll
        for (int k = 0; k < VN; k++) m[k] = __ballot_sync(0xffffffffu, cmp_op<T, OP>(va.v[k], SCALAR ? scalar : vb.v[k]));
        if constexpr (VN == 2) {
            uint64_t w = spread_bits(m[0]) | (spread_bits(m[1]) << 1);
            if (lane == 0) *reinterpret_cast<uint2*>(out + s * 2) = make_uint2((uint32_t)w, (uint32_t)(w >> 32));
        } else {
            // rows 4l+k: bit position 4l+k of a 128-bit word.  First interleave (0,2) and (1,3) pairs to
            // 2-spaced 64-bit words, then interleave those.
            uint64_t e = spread_bits(m[0]) | (spread_bits(m[2]) << 1);   // bit 2l+j  <- m[2j] bit l
            uint64_t o = spread_bits(m[1]) | (spread_bits(m[3]) << 1);   // bit 2l+j  <- m[2j+1] bit l
            uint64_t e_lo = spread_bits((uint32_t)e), e_hi = spread_bits((uint32_t)(e >> 32));
            uint64_t o_lo = spread_bits((uint32_t)o), o_hi = spread_bits((uint32_t)(o >> 32));
            uint64_t lo = e_lo | (o_lo << 1), hi = e_hi | (o_hi << 1);
            if (lane == 0) *reinterpret_cast<uint4*>(out + s * 4) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
        }
    };
    int64_t s = warp;
    for (; s + (UNROLL - 1) * nwarps < nsteps; s += UNROLL * nwarps) {
        uint4 ta[UNROLL], tb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            ta[u] = ld_stream_u32x4(reinterpret_cast<const uint4*>(lhs) + (s + u * nwarps) * 32 + lane);
            if (!SCALAR) tb[u] = ld_stream_u32x4(reinterpret_cast<const uint4*>(rhs) + (s + u * nwarps) * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            V va = *reinterpret_cast<V*>(&ta[u]), vb;
            if (!SCALAR) vb = *reinterpret_cast<V*>(&tb[u]);
            emit(s + u * nwarps, va, vb);
        }
    }
    for (; s < nsteps; s += nwarps) {
        uint4 ta = ld_stream_u32x4(reinterpret_cast<const uint4*>(lhs) + s * 32 + lane);
        V va = *reinterpret_cast<V*>(&ta), vb;
        if (!SCALAR) { uint4 tb = ld_stream_u32x4(reinterpret_cast<const uint4*>(rhs) + s * 32 + lane); vb = *reinterpret_cast<V*>(&tb); }
        emit(s, va, vb);
    }
    // tail rows: one warp, 32 rows per ballot
    if (warp == 0) {
        for (int64_t base = nsteps * ROWS_PER_STEP; base < n; base += 32) {
            int64_t i = base + lane;
            bool p = false;
            if (i < n) p = cmp_op<T, OP>(lhs[i], SCALAR ? scalar : rhs[i]);
            uint32_t w = __ballot_sync(0xffffffffu, p);
            if (lane == 0) out[base >> 5] = w;
        }
    }
}

// 8-byte types, second form: lane l of a warp-step owns rows l and l + 32 of a 64-row group (two coalesced 256-byte
// loads), so the two ballots ARE the two mask words in row order — no bit interleave at all.  The 128-bit form above spends
// ~40 integer instructions per 16 bytes on ballot + spread_bits, which is what capped the scalar compare (8 B/row in,
// 1 bit/row out) at 0.65 of the copy peak; array-vs-array (16 B/row) was already memory-bound.
template <typename T, int OP, bool SCALAR>
__global__ void __launch_bounds__(256) k_compare64(const T* __restrict__ lhs, const T* __restrict__ rhs, uint32_t* __restrict__ out, int64_t n, T scalar) {
    static_assert(sizeof(T) == 8, "8-byte elements");
    const int64_t nsteps = n / 64;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lane = lane_id();
    constexpr int UNROLL = 4;
    int64_t s = warp;
    for (; s + (UNROLL - 1) * nwarps < nsteps; s += UNROLL * nwarps) {
        T a0[UNROLL], a1[UNROLL], b0[UNROLL], b1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t r = (s + u * nwarps) * 64 + lane;
            a0[u] = __ldcs(lhs + r); a1[u] = __ldcs(lhs + r + 32);
            if (!SCALAR) { b0[u] = __ldcs(rhs + r); b1[u] = __ldcs(rhs + r + 32); }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t m0 = __ballot_sync(0xffffffffu, cmp_op<T, OP>(a0[u], SCALAR ? scalar : b0[u]));
            const uint32_t m1 = __ballot_sync(0xffffffffu, cmp_op<T, OP>(a1[u], SCALAR ? scalar : b1[u]));
            if (lane == 0) *reinterpret_cast<uint2*>(out + (s + u * nwarps) * 2) = make_uint2(m0, m1);
        }
    }
    for (; s < nsteps; s += nwarps) {
        const int64_t r = s * 64 + lane;
        const uint32_t m0 = __ballot_sync(0xffffffffu, cmp_op<T, OP>(lhs[r], SCALAR ? scalar : rhs[r]));
        const uint32_t m1 = __ballot_sync(0xffffffffu, cmp_op<T, OP>(lhs[r + 32], SCALAR ? scalar : rhs[r + 32]));
        if (lane == 0) *reinterpret_cast<uint2*>(out + s * 2) = make_uint2(m0, m1);
    }
    if (warp == 0) {
        for (int64_t base = nsteps * 64; base < n; base += 32) {
            const int64_t i = base + lane;
            bool p = false;
            if (i < n) p = cmp_op<T, OP>(lhs[i], SCALAR ? scalar : rhs[i]);
            const uint32_t w = __ballot_sync(0xffffffffu, p);
            if (lane == 0) out[base >> 5] = w;
        }
    }
}

template <typename T, int OP>
static void launch_cmp_s(bool scalar_rhs, const T* l, const T* r, uint32_t* o, int64_t n, T scalar) {
    int grid = grid_for(n / (16 / sizeof(T)) + 1, 256);
    if constexpr (sizeof(T) == 8) {
        if (scalar_rhs) { PLB_LAUNCH("k2_compare", (k_compare64<T, OP, true>), grid, 256, 0, l, r, o, n, scalar); return; }      // instruction-bound in the 128-bit form
    }
    if (scalar_rhs) PLB_LAUNCH("k2_compare", (k_compare<T, OP, true>), grid, 256, 0, l, r, o, n, scalar);
    else PLB_LAUNCH("k2_compare", (k_compare<T, OP, false>), grid, 256, 0, l, r, o, n, scalar);
}
template <typename T>
static void launch_cmp(int op, bool scalar_rhs, const T* l, const T* r, uint32_t* o, int64_t n, T scalar) {
    switch (op) {
        case BL_CMP_EQ: launch_cmp_s<T, BL_CMP_EQ>(scalar_rhs, l, r, o, n, scalar); break;
        case BL_CMP_NE: launch_cmp_s<T, BL_CMP_NE>(scalar_rhs, l, r, o, n, scalar); break;
        case BL_CMP_LT: launch_cmp_s<T, BL_CMP_LT>(scalar_rhs, l, r, o, n, scalar); break;
        case BL_CMP_LE: launch_cmp_s<T, BL_CMP_LE>(scalar_rhs, l, r, o, n, scalar); break;
        case BL_CMP_GT: launch_cmp_s<T, BL_CMP_GT>(scalar_rhs, l, r, o, n, scalar); break;
        default: launch_cmp_s<T, BL_CMP_GE>(scalar_rhs, l, r, o, n, scalar); break;
    }
}

// missing-aware eq/ne fix-up (comparisons/mod.rs:14-52): out = both valid ? out : (va == vb) [eq] / (va != vb) [ne]
__global__ void k_cmp_missing_fix(uint32_t* out, const uint32_t* va, const uint32_t* vb, int is_ne, int64_t nwords) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t a = va ? va[i] : 0xFFFFFFFFu, b = vb ? vb[i] : 0xFFFFFFFFu;
        uint32_t both = a & b;
        uint32_t alt = is_ne ? (a ^ b) : ~(a ^ b);
        out[i] = (out[i] & both) | (alt & ~both);
    }
}

template <typename T> static T read_host_scalar(const DevCol& c) {
    T h; PLB_CUDA(cudaMemcpyAsync(&h, c.v(), sizeof(T), cudaMemcpyDeviceToHost, ctx().stream));
    PLB_CUDA(cudaStreamSynchronize(ctx().stream)); return h;
}
static bool scalar_is_null(const DevCol& c) {
    if (!c.validity) return false;
    uint32_t w = read_host_scalar<uint32_t>(DevCol{BL_UINT32, 1, c.validity, nullptr, 0});
    return (w & 1u) == 0;
}
static DevCol full_null(int dtype, int64_t n) {
    DevCol o = make_col(dtype, n, true);
    dev_memset(o.values->p, 0, o.values->bytes); dev_memset(o.validity->p, 0, o.validity->bytes);
    o.null_count = n; return o;
}

// ---------------------------------------------------------------------------- operators
template <typename T, bool IS_FLOAT>
static DevCol elementwise_typed(int op, const DevCol& lhs, const DevCol& rhs, int mode, int64_t n) {
    const DevCol& arr = mode == MODE_SA ? rhs : lhs;
    T scalar = 0;
    if (mode == MODE_AS) scalar = read_host_scalar<T>(rhs);
    if (mode == MODE_SA) scalar = read_host_scalar<T>(lhs);
    int out_dt = (!IS_FLOAT && op == BL_OP_TRUE_DIV) ? BL_FLOAT64 : arr.dtype;
    // reference scalar shortcuts that change results: int // 0 and % 0 -> all null (signed.rs:103-105,173-175)
    if (!IS_FLOAT && mode == MODE_AS && (op == BL_OP_FLOOR_DIV || op == BL_OP_MOD) && scalar == 0) return full_null(out_dt, n);
    DevCol out = make_col(out_dt, n, false);
    const T* l = (const T*)lhs.v(); const T* r = (const T*)rhs.v();
    if (n > 0) {
        if (!IS_FLOAT && op == BL_OP_TRUE_DIV) {
            int grid = grid_for(n, 256);
            if (mode == MODE_AA) PLB_LAUNCH("k1_int_true_div", (k_int_true_div<T, MODE_AA>), grid, 256, 0, l, r, as<double>(out.values), n, scalar);
            else if (mode == MODE_AS) PLB_LAUNCH("k1_int_true_div", (k_int_true_div<T, MODE_AS>), grid, 256, 0, l, r, as<double>(out.values), n, scalar);
            else PLB_LAUNCH("k1_int_true_div", (k_int_true_div<T, MODE_SA>), grid, 256, 0, l, r, as<double>(out.values), n, scalar);
        } else launch_arith<T, IS_FLOAT>(op, mode, l, r, as<T>(out.values), n, scalar);
    }
    // validity = AND of inputs; integer //,% additionally AND (rhs != 0)
    const uint32_t* lv = mode == MODE_SA ? nullptr : lhs.vm();
    const uint32_t* rv = mode == MODE_AS ? nullptr : rhs.vm();
    DevPtr nz;
    if (!IS_FLOAT && (op == BL_OP_FLOOR_DIV || op == BL_OP_MOD) && mode != MODE_AS && n > 0) {
        nz = dev_alloc(bitmap_bytes(n) + 16);
        launch_cmp<T>(BL_CMP_NE, true, r, nullptr, as<uint32_t>(nz), n, (T)0);
    }
    out.validity = bitmap_and(lv, rv, nz ? as<uint32_t>(nz) : nullptr, n);
    out.null_count = out.validity ? -1 : 0;
    return out;
}

DevCol op_elementwise(int op, const DevCol& lhs, const DevCol& rhs) {
    PLB_REQUIRE(op >= BL_OP_ADD && op <= BL_OP_TRUE_DIV, BL_ERR_INVALID, "elementwise: unknown op");
    PLB_REQUIRE(lhs.dtype == rhs.dtype, BL_ERR_DTYPE, std::string("elementwise: dtypes differ (") + dtype_name(lhs.dtype) + " vs " + dtype_name(rhs.dtype) + ")");
    int mode; int64_t n;
    if (lhs.len == rhs.len) { mode = MODE_AA; n = lhs.len; }
    else if (rhs.len == 1) { mode = MODE_AS; n = lhs.len; }
    else if (lhs.len == 1) { mode = MODE_SA; n = rhs.len; }
    else fail(BL_ERR_INVALID, "elementwise: lengths " + std::to_string(lhs.len) + " and " + std::to_string(rhs.len) + " do not broadcast");
    // a null scalar makes everything null
    if ((mode == MODE_AS && scalar_is_null(rhs)) || (mode == MODE_SA && scalar_is_null(lhs))) {
        int out_dt = (dtype_is_int(lhs.dtype) && op == BL_OP_TRUE_DIV) ? BL_FLOAT64 : lhs.dtype;
        return full_null(out_dt, n);
    }
    switch (lhs.dtype) {
        case BL_INT64: return elementwise_typed<int64_t, false>(op, lhs, rhs, mode, n);
        case BL_INT32: return elementwise_typed<int32_t, false>(op, lhs, rhs, mode, n);
        case BL_UINT64: return elementwise_typed<uint64_t, false>(op, lhs, rhs, mode, n);
        case BL_UINT32: return elementwise_typed<uint32_t, false>(op, lhs, rhs, mode, n);
        case BL_FLOAT64: return elementwise_typed<double, true>(op, lhs, rhs, mode, n);
        case BL_FLOAT32: return elementwise_typed<float, true>(op, lhs, rhs, mode, n);
        default: fail(BL_ERR_UNSUPPORTED, std::string("elementwise: dtype ") + dtype_name(lhs.dtype) + " is outside the hot path");
    }
}

template <typename T>
static DevCol compare_typed(int op, const DevCol& lhs, const DevCol& rhs, bool scalar_rhs, bool missing) {
    int64_t n = lhs.len;
    DevCol out = make_col(BL_BOOL, n, false);
    T scalar = scalar_rhs ? read_host_scalar<T>(rhs) : (T)0;
    if (n > 0) launch_cmp<T>(op, scalar_rhs, (const T*)lhs.v(), (const T*)rhs.v(), as<uint32_t>(out.values), n, scalar);
    const uint32_t* lv = lhs.vm();
    const uint32_t* rv = scalar_rhs ? nullptr : rhs.vm();
    if (missing && (op == BL_CMP_EQ || op == BL_CMP_NE)) {
        if ((lv || rv) && n > 0)
            PLB_LAUNCH("k2_missing_fix", k_cmp_missing_fix, grid_for((n + 31) / 32, 256), 256, 0, as<uint32_t>(out.values), lv, rv, op == BL_CMP_NE ? 1 : 0, (n + 31) / 32);
        out.null_count = 0;
    } else {
        out.validity = bitmap_and(lv, rv, nullptr, n);
        out.null_count = out.validity ? -1 : 0;
    }
    return out;
}

DevCol op_compare(int op, const DevCol& lhs, const DevCol& rhs, bool missing) {
    PLB_REQUIRE(op >= BL_CMP_EQ && op <= BL_CMP_GE, BL_ERR_INVALID, "compare: unknown op");
    PLB_REQUIRE(lhs.dtype == rhs.dtype, BL_ERR_DTYPE, "compare: dtypes differ");
    bool scalar_rhs = rhs.len == 1 && lhs.len != 1;
    PLB_REQUIRE(scalar_rhs || lhs.len == rhs.len, BL_ERR_INVALID, "compare: lengths do not broadcast (only a length-1 rhs is a scalar)");
    if (scalar_rhs && scalar_is_null(rhs)) {
        if (!missing) return full_null(BL_BOOL, lhs.len);
        fail(BL_ERR_UNSUPPORTED, "compare: eq_missing against a null scalar is outside the hot path");
    }
    switch (lhs.dtype) {
        case BL_INT64: return compare_typed<int64_t>(op, lhs, rhs, scalar_rhs, missing);
        case BL_INT32: return compare_typed<int32_t>(op, lhs, rhs, scalar_rhs, missing);
        case BL_UINT64: return compare_typed<uint64_t>(op, lhs, rhs, scalar_rhs, missing);
        case BL_UINT32: return compare_typed<uint32_t>(op, lhs, rhs, scalar_rhs, missing);
        case BL_FLOAT64: return compare_typed<double>(op, lhs, rhs, scalar_rhs, missing);
        case BL_FLOAT32: return compare_typed<float>(op, lhs, rhs, scalar_rhs, missing);
        default: fail(BL_ERR_UNSUPPORTED, std::string("compare: dtype ") + dtype_name(lhs.dtype) + " is outside the hot path");
    }
}

// predicate mask for the fused filter: rows where col (op) scalar, nulls -> false
DevCol op_cmp_scalar_mask(const DevCol& col, int cmp_op, const DevCol& scalar) {
    DevCol m = op_compare(cmp_op, col, scalar, false);
    if (m.validity) {   // null -> false (filter/mod.rs:21-27)
        DevPtr v = bitmap_and(as<uint32_t>(m.values), m.vm(), nullptr, m.len);
        m.values = v; m.validity = nullptr; m.null_count = 0;
    }
    return m;
}

// ---------------------------------------------------------------------------- small-integer casts
// The reference aggregates Int8/16 and UInt8/16 columns after a cast to Int64
// (polars-core/src/series/implementations/mod.rs:145-154) and groups / joins integer keys on their
// zero-extended bit representation (into_groups.rs:178-185 `to_bit_repr`).  The hot kernels only
// take 4- and 8-byte elements, so the operator entry points widen such columns first and narrow
// key / min / max outputs back (truncation restores the original bit pattern exactly).
template <typename S, typename D>
__global__ void __launch_bounds__(256) k_cast_int(const S* __restrict__ in, D* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (D)in[i];
}
template <typename S, typename D> static void launch_cast(const DevCol& in, DevCol& out) {
    if (in.len > 0) PLB_LAUNCH("k1_cast_int", (k_cast_int<S, D>), grid_for(in.len, 256, 16), 256, 0, reinterpret_cast<const S*>(in.v()), reinterpret_cast<D*>(out.values->p), in.len);
}
bool dtype_is_small_int(int dt) { return dt == BL_INT8 || dt == BL_INT16 || dt == BL_UINT8 || dt == BL_UINT16; }

// small int -> BL_UINT32 (`bits`: zero-extended bit pattern, for keys) or BL_INT64 (numeric value, for
// aggregated columns); BL_UINT32 / BL_INT64 -> small int (truncation).  The validity bitmap is shared.
DevCol op_cast_small_int(const DevCol& in, int to_dtype, bool bits) {
    DevCol out = make_col(to_dtype, in.len, false);
    out.validity = in.validity; out.null_count = in.null_count;
    const int from = in.dtype;
    if (to_dtype == BL_UINT32 && bits) {
        if (dtype_size(from) == 1) launch_cast<uint8_t, uint32_t>(in, out); else launch_cast<uint16_t, uint32_t>(in, out);
    } else if (to_dtype == BL_INT64 && dtype_is_small_int(from)) {
        switch (from) {
            case BL_INT8: launch_cast<int8_t, int64_t>(in, out); break;
            case BL_INT16: launch_cast<int16_t, int64_t>(in, out); break;
            case BL_UINT8: launch_cast<uint8_t, int64_t>(in, out); break;
            default: launch_cast<uint16_t, int64_t>(in, out); break;
        }
    } else if (from == BL_UINT32 && dtype_is_small_int(to_dtype)) {
        if (dtype_size(to_dtype) == 1) launch_cast<uint32_t, uint8_t>(in, out); else launch_cast<uint32_t, uint16_t>(in, out);
    } else if (from == BL_INT64 && dtype_is_small_int(to_dtype)) {
        if (dtype_size(to_dtype) == 1) launch_cast<int64_t, uint8_t>(in, out); else launch_cast<int64_t, uint16_t>(in, out);
    } else fail(BL_ERR_UNSUPPORTED, std::string("cast ") + dtype_name(from) + " -> " + dtype_name(to_dtype) + " is outside the hot path");
    return out;
}

}  // namespace plb
