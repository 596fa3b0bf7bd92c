// ni_kernels.cu -- hand-written sm_100a kernels + the C ABI of libni_b200.so (include/ni_b200.h).
//
// The hot path of blairstar/NaturalDiffusion's Natural Inference sampler is a <=K-term
// weighted sum over stored x0 / noise tensors plus an affine model-output conversion: a pure
// HBM-streaming problem (<= ~2 flop/byte).  No tensor cores.  What matters here is
//   * one pass: model outputs, current input, history, noise are each read once and
//     x0 / x_next / (optional) kept noise are each written once per step;
//   * 128-bit coalesced accesses, every load of a thread issued before the first use so
//     >= 6-10 x 16 B are in flight per thread (B200 needs ~35 KB in flight per SM);
//   * Philox4x32-10 + Box-Muller evaluated while those loads are in flight;
//   * tables (pointers, coefficients) in kernel parameters: no allocation, no sync,
//     CUDA-graph capturable.
// Reference semantics: see include/ni_b200.h (each entry point cites file:line).
#include "ni_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>

#ifndef NI_LOAD_POLICY
#define NI_LOAD_POLICY 1 /* 0 plain ld.global, 1 ld.global.L1::no_allocate, 2 ld.global.cs (evict-first) */
#endif
#ifndef NI_STORE_POLICY
#define NI_STORE_POLICY 0 /* 0 plain st.global, 1 st.global.cs */
#endif
#ifndef NI_BLOCK
#define NI_BLOCK 256
#endif

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ------------------------------------------------------------------------------------------
// raw vector loads / stores with cache policy
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld128(const void *p)
{
    uint4 r;
#if NI_LOAD_POLICY == 1
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#elif NI_LOAD_POLICY == 2
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#else
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ uint2 ld64(const void *p)
{
    uint2 r;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st128(void *p, uint4 v)
{
#if NI_STORE_POLICY == 1
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}
__device__ __forceinline__ void st64(void *p, uint2 v)
{
    asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float f);
template <> __device__ __forceinline__ float from_f<float>(float f) { return f; }
template <> __device__ __forceinline__ __half from_f<__half>(float f) { return __float2half_rn(f); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }

// VEC elements of T <-> registers.  VEC*sizeof(T) is 2, 4 (scalar path), 8, 16 or 32 bytes.
template <typename T, int VEC> struct Raw {
    static constexpr int BYTES = VEC * (int)sizeof(T);
    static constexpr int WORDS = BYTES >= 4 ? BYTES / 4 : 1;
    uint32_t w[WORDS];
};

template <typename T, int VEC> __device__ __forceinline__ Raw<T, VEC> load_raw(const T *p)
{
    Raw<T, VEC> r;
    constexpr int BYTES = Raw<T, VEC>::BYTES;
    if constexpr (BYTES == 32) {
        uint4 a = ld128(p), b = ld128(reinterpret_cast<const char *>(p) + 16);
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w; r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
    } else if constexpr (BYTES == 16) {
        uint4 a = ld128(p);
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
    } else if constexpr (BYTES == 8) {
        uint2 a = ld64(p);
        r.w[0] = a.x; r.w[1] = a.y;
    } else if constexpr (BYTES == 4) {
        r.w[0] = *reinterpret_cast<const uint32_t *>(p);
    } else {
        r.w[0] = *reinterpret_cast<const uint16_t *>(p);
    }
    return r;
}

template <typename T, int VEC> __device__ __forceinline__ void unpack(const Raw<T, VEC> &r, float (&f)[VEC])
{
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) f[i] = __uint_as_float(r.w[i]);
    } else if constexpr (VEC == 1) {
        unsigned short s = (unsigned short)r.w[0];
        T t;
        memcpy(&t, &s, 2);
        f[0] = to_f<T>(t);
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
                __half2 h;
                memcpy(&h, &r.w[i], 4);
                float2 v = __half22float2(h);
                f[2 * i] = v.x; f[2 * i + 1] = v.y;
            } else {
                // bf16 -> f32 is a 16-bit shift
                f[2 * i] = __uint_as_float(r.w[i] << 16);
                f[2 * i + 1] = __uint_as_float(r.w[i] & 0xffff0000u);
            }
        }
    }
}

template <typename T, int VEC> __device__ __forceinline__ Raw<T, VEC> pack(const float (&f)[VEC])
{
    Raw<T, VEC> r;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) r.w[i] = __float_as_uint(f[i]);
    } else if constexpr (VEC == 1) {
        T t = from_f<T>(f[0]);
        unsigned short s;
        memcpy(&s, &t, 2);
        r.w[0] = s;
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            if constexpr (std::is_same<T, __half>::value) {
                __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                memcpy(&r.w[i], &h, 4);
            } else {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                memcpy(&r.w[i], &h, 4);
            }
        }
    }
    return r;
}

template <typename T, int VEC> __device__ __forceinline__ void store_raw(T *p, const Raw<T, VEC> &r)
{
    constexpr int BYTES = Raw<T, VEC>::BYTES;
    if constexpr (BYTES == 32) {
        st128(p, make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]));
        st128(reinterpret_cast<char *>(p) + 16, make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]));
    } else if constexpr (BYTES == 16) {
        st128(p, make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]));
    } else if constexpr (BYTES == 8) {
        st64(p, make_uint2(r.w[0], r.w[1]));
    } else if constexpr (BYTES == 4) {
        *reinterpret_cast<uint32_t *>(p) = r.w[0];
    } else {
        *reinterpret_cast<uint16_t *>(p) = (uint16_t)r.w[0];
    }
}

// round-trip through the storage type (what a later step will read back)
template <typename T> __device__ __forceinline__ float round_to(float f)
{
    if constexpr (sizeof(T) == 4) return f;
    else return to_f<T>(from_f<T>(f));
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (noise contract in include/ni_b200.h; CPU twin: oracle/philox_oracle.c)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float &za, float &zb)
{
    const float u = fmaf(__uint2float_rn(ra), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float v = fmaf(__uint2float_rn(rb), 4.6566128730773926e-10f, 2.3283064365386963e-10f);
#ifdef NI_FAST_NORMAL
    const float rad = sqrtf(-1.3862943611198906f * __log2f(u));
    float s, c;
    __sincosf(3.14159265358979f * v, &s, &c);
#else
    const float rad = sqrtf(-2.0f * logf(u));
    float s, c;
    sincospif(v, &s, &c);
#endif
    za = rad * c;
    zb = rad * s;
}

__device__ __forceinline__ void normal4(uint64_t group, uint64_t tensor_id, uint32_t k0, uint32_t k1, float (&z)[4])
{
    const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)tensor_id, (uint32_t)(tensor_id >> 32)), k0, k1);
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
}

// VEC normals for global elements [e, e+VEC)
template <int VEC> __device__ __forceinline__ void normal_vec(uint64_t e, uint64_t tensor_id, uint32_t k0, uint32_t k1, float (&z)[VEC])
{
    if constexpr (VEC == 1) {
        float q[4];
        normal4(e >> 2, tensor_id, k0, k1, q);
        const int lane = (int)(e & 3);
        z[0] = lane == 0 ? q[0] : lane == 1 ? q[1] : lane == 2 ? q[2] : q[3];
    } else {
#pragma unroll
        for (int j = 0; j < VEC / 4; ++j) {
            float q[4];
            normal4((e >> 2) + j, tensor_id, k0, k1, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) z[4 * j + i] = q[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// the fused step
// ------------------------------------------------------------------------------------------
template <int CAP> struct TermTable {
    const void *ptr[CAP];
    float c[CAP];
};

struct StepArgs {
    int64_t nvec; // vectors of VEC elements
    int64_t per_sample, out_sample_stride;
    const void *x_in, *out0, *out1;
    void *x0_dst, *x_next, *x_next_lp;
    float *sumsq;
    void *gen_dst[NI_MAX_GEN];
    uint64_t gen_tid[NI_MAX_GEN];
    uint64_t elem_offset;
    float gen_c[NI_MAX_GEN];
    float a, b0, b1, c_x0;
    uint32_t k0, k1;
    int n_terms, n_gen;
    int has_x0, out_strided, accumulate, lp_dtype;
};

template <typename T, typename TO, int VEC, int CAP>
__global__ void __launch_bounds__(NI_BLOCK) ni_step_kernel(const __grid_constant__ StepArgs s, const __grid_constant__ TermTable<CAP> tab)
{
    const int64_t v = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (v >= s.nvec) return;
    const int64_t e = v * VEC; // first element of this thread

    // 1. issue the x0-stage loads (consumed after the term loop)
    Raw<T, VEC> rx;
    Raw<TO, VEC> ro0, ro1;
    const bool has_x0 = s.has_x0 != 0;
    const bool has_x = has_x0 && s.x_in != nullptr;
    const bool has_o1 = has_x0 && s.out1 != nullptr;
    int64_t sample = 0;
    if (s.out_strided || s.sumsq != nullptr) sample = e / s.per_sample;
    if (has_x0) {
        int64_t eo = e;
        if (s.out_strided) eo = sample * s.out_sample_stride + (e - sample * s.per_sample);
        ro0 = load_raw<TO, VEC>(static_cast<const TO *>(s.out0) + eo);
        if (has_o1) ro1 = load_raw<TO, VEC>(static_cast<const TO *>(s.out1) + eo);
        if (has_x) rx = load_raw<T, VEC>(static_cast<const T *>(s.x_in) + e);
    }

    float acc[VEC];
    if (s.accumulate) {
        unpack<T, VEC>(load_raw<T, VEC>(static_cast<const T *>(s.x_next) + e), acc);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    }

    // 2. stored terms: batches of independent 128-bit loads, then the FMAs
    const int n = s.n_terms;
    int t = 0;
#define NI_TERM_BATCH(NB)                                                                      \
    for (; t + NB <= n; t += NB) {                                                             \
        Raw<T, VEC> rr[NB];                                                                    \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) rr[j] = load_raw<T, VEC>(static_cast<const T *>(tab.ptr[t + j]) + e); \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) {                                       \
            float f[VEC];                                                                      \
            unpack<T, VEC>(rr[j], f);                                                          \
            const float c = tab.c[t + j];                                                      \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c, f[i], acc[i]);    \
        }                                                                                      \
    }
    NI_TERM_BATCH(8)
    NI_TERM_BATCH(4)
    NI_TERM_BATCH(2)
    NI_TERM_BATCH(1)
#undef NI_TERM_BATCH

    // 3. generated noise (pure ALU; overlaps the loads still in flight)
    for (int g = 0; g < s.n_gen; ++g) {
        float z[VEC];
        normal_vec<VEC>(s.elem_offset + (uint64_t)e, s.gen_tid[g], s.k0, s.k1, z);
        if (s.gen_dst[g] != nullptr) {
            store_raw<T, VEC>(static_cast<T *>(s.gen_dst[g]) + e, pack<T, VEC>(z));
#pragma unroll
            for (int i = 0; i < VEC; ++i) z[i] = round_to<T>(z[i]);
        }
        const float c = s.gen_c[g];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c, z[i], acc[i]);
    }

    // 4. x0 = a*x + b0*out0 + b1*out1, kept in the ring, enters the sum with A[k,k]
    if (has_x0) {
        float x0[VEC], f[VEC];
        unpack<TO, VEC>(ro0, f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) x0[i] = s.b0 * f[i];
        if (has_o1) {
            unpack<TO, VEC>(ro1, f);
#pragma unroll
            for (int i = 0; i < VEC; ++i) x0[i] = fmaf(s.b1, f[i], x0[i]);
        }
        if (has_x) {
            unpack<T, VEC>(rx, f);
#pragma unroll
            for (int i = 0; i < VEC; ++i) x0[i] = fmaf(s.a, f[i], x0[i]);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) x0[i] = round_to<T>(x0[i]);
        if (s.x0_dst != nullptr) store_raw<T, VEC>(static_cast<T *>(s.x0_dst) + e, pack<T, VEC>(x0));
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(s.c_x0, x0[i], acc[i]);
    }

    // 5. x_{k+1}
    store_raw<T, VEC>(static_cast<T *>(s.x_next) + e, pack<T, VEC>(acc));
    if (s.x_next_lp != nullptr) {
        if (s.lp_dtype == NI_BF16) store_raw<__nv_bfloat16, VEC>(static_cast<__nv_bfloat16 *>(s.x_next_lp) + e, pack<__nv_bfloat16, VEC>(acc));
        else store_raw<__half, VEC>(static_cast<__half *>(s.x_next_lp) + e, pack<__half, VEC>(acc));
    }

    // 6. per-sample sum of squares: warp shuffle, one atomic per warp (per lane only where a warp straddles samples)
    if (s.sumsq != nullptr) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float r = round_to<T>(acc[i]);
            ss = fmaf(r, r, ss);
        }
        // full, converged warp inside one sample: butterfly + one atomic; otherwise (tail warp,
        // warp straddling a sample boundary) one atomic per lane.
        const unsigned mask = __activemask();
        bool fast = mask == 0xffffffffu;
        if (fast) {
            const long long s0 = __shfl_sync(0xffffffffu, (long long)sample, 0);
            fast = __all_sync(0xffffffffu, (long long)sample == s0);
        }
        if (fast) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(s.sumsq + sample, ss);
        } else {
            atomicAdd(s.sumsq + sample, ss);
        }
    }
}

// ------------------------------------------------------------------------------------------
// stand-alone weighted sum (drop-in for the reference's weighted_sum functions)
// ------------------------------------------------------------------------------------------
template <int CAP> struct WsumTable {
    const void *ptr[CAP];
    double c[CAP];
};

template <typename TS> struct SrcIO {
    template <int VEC> static __device__ __forceinline__ void load(const void *p, int64_t e, float (&f)[VEC])
    {
        unpack<TS, VEC>(load_raw<TS, VEC>(static_cast<const TS *>(p) + e), f);
    }
};

template <typename TS, typename TD, int VEC, int CAP>
__global__ void __launch_bounds__(NI_BLOCK) ni_wsum_kernel(const __grid_constant__ WsumTable<CAP> tab, int n, void *dst, int64_t nvec, double scale)
{
    const int64_t v = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (v >= nvec) return;
    const int64_t e = v * VEC;
    if constexpr (std::is_same<TS, double>::value) {
        // fp64 history (CIFAR loop): fp64 accumulate, like the reference
        double acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.0;
        int t = 0;
        for (; t + 4 <= n; t += 4) {
            Raw<double, VEC> rr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) rr[j] = load_raw<double, VEC>(static_cast<const double *>(tab.ptr[t + j]) + e);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fma(tab.c[t + j], __hiloint2double(rr[j].w[2 * i + 1], rr[j].w[2 * i]), acc[i]);
        }
        for (; t < n; ++t) {
            Raw<double, VEC> r1 = load_raw<double, VEC>(static_cast<const double *>(tab.ptr[t]) + e);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = fma(tab.c[t], __hiloint2double(r1.w[2 * i + 1], r1.w[2 * i]), acc[i]);
        }
        if constexpr (std::is_same<TD, double>::value) {
            Raw<double, VEC> o;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const double r = acc[i] * scale;
                o.w[2 * i] = __double2loint(r);
                o.w[2 * i + 1] = __double2hiint(r);
            }
            store_raw<double, VEC>(static_cast<double *>(dst) + e, o);
        } else {
            float f[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) f[i] = (float)(acc[i] * scale);
            store_raw<TD, VEC>(static_cast<TD *>(dst) + e, pack<TD, VEC>(f));
        }
    } else {
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        int t = 0;
        for (; t + 8 <= n; t += 8) {
            Raw<TS, VEC> rr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) rr[j] = load_raw<TS, VEC>(static_cast<const TS *>(tab.ptr[t + j]) + e);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float f[VEC];
                unpack<TS, VEC>(rr[j], f);
                const float c = (float)tab.c[t + j];
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c, f[i], acc[i]);
            }
        }
        for (; t < n; ++t) {
            float f[VEC];
            unpack<TS, VEC>(load_raw<TS, VEC>(static_cast<const TS *>(tab.ptr[t]) + e), f);
            const float c = (float)tab.c[t];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c, f[i], acc[i]);
        }
        const float sc = (float)scale;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] *= sc;
        store_raw<TD, VEC>(static_cast<TD *>(dst) + e, pack<TD, VEC>(acc));
    }
}

// ------------------------------------------------------------------------------------------
// noise only, and the pixel output stage
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(NI_BLOCK) ni_normal_kernel(T *dst, int64_t nvec, uint32_t k0, uint32_t k1, uint64_t tensor_id, uint64_t elem_offset)
{
    const int64_t v = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (v >= nvec) return;
    float z[VEC];
    normal_vec<VEC>(elem_offset + (uint64_t)(v * VEC), tensor_id, k0, k1, z);
    store_raw<T, VEC>(dst + v * VEC, pack<T, VEC>(z));
}

// NCHW -> NHWC uint8.  One thread per (n, h, w) pixel reads C planes (coalesced along w) and
// writes C consecutive bytes; for C == 3 a warp writes 96 contiguous bytes.
template <typename T>
__global__ void __launch_bounds__(NI_BLOCK) ni_pixel_kernel(const T *__restrict__ x, uint8_t *__restrict__ dst, int64_t npix_total, int C, int64_t HW, float scale, float shift)
{
    const int64_t p = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (p >= npix_total) return;
    const int64_t n = p / HW, hw = p - n * HW;
    const T *src = x + n * C * HW + hw;
    uint8_t *out = dst + p * C;
    for (int c = 0; c < C; ++c) {
        const float val = to_f<T>(src[(int64_t)c * HW]);
        // reference order: y = (x+1)/2 then clip(y*255, 0, 255) then truncate (numpy astype)
        const float y = (val * scale + shift) * 255.0f;
        out[c] = (uint8_t)(int)fminf(fmaxf(y, 0.f), 255.f);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int dtype_size(int d) { return d == NI_F32 ? 4 : d == NI_F64 ? 8 : (d == NI_F16 || d == NI_BF16) ? 2 : 0; }

int check_launch(const char *what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(NI_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NI_OK;
}

template <typename T, typename TO, int VEC>
int launch_step_cap(const StepArgs &a, const NiStepDesc *d, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((a.nvec + NI_BLOCK - 1) / NI_BLOCK);
    if (d->n_terms <= 32) {
        TermTable<32> tab;
        memset(&tab, 0, sizeof(tab));
        for (int i = 0; i < d->n_terms; ++i) { tab.ptr[i] = d->term_ptrs_host[i]; tab.c[i] = d->term_coeffs_host[i]; }
        ni_step_kernel<T, TO, VEC, 32><<<blocks, NI_BLOCK, 0, st>>>(a, tab);
    } else {
        static thread_local TermTable<NI_MAX_TERMS> tab;
        for (int i = 0; i < d->n_terms; ++i) { tab.ptr[i] = d->term_ptrs_host[i]; tab.c[i] = d->term_coeffs_host[i]; }
        ni_step_kernel<T, TO, VEC, NI_MAX_TERMS><<<blocks, NI_BLOCK, 0, st>>>(a, tab);
    }
    return check_launch("ni_step launch");
}

template <typename T, typename TO> int launch_step(StepArgs &a, const NiStepDesc *d, bool vec_ok, cudaStream_t st)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    if (vec_ok) {
        a.nvec = d->numel / VEC;
        return launch_step_cap<T, TO, VEC>(a, d, st);
    }
    a.nvec = d->numel;
    return launch_step_cap<T, TO, 1>(a, d, st);
}

} // namespace

extern "C" {

int ni_version(void) { return NI_ABI_VERSION; }
const char *ni_last_error(void) { return g_err; }
int64_t ni_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ni_step(const NiStepDesc *d, void *stream)
{
    if (d == nullptr) return fail(NI_ERR_INVALID, "ni_step: NULL descriptor");
    if (d->numel < 0 || d->per_sample <= 0) return fail(NI_ERR_INVALID, "ni_step: bad sizes numel=%lld per_sample=%lld", (long long)d->numel, (long long)d->per_sample);
    if (d->numel % d->per_sample != 0) return fail(NI_ERR_INVALID, "ni_step: numel %lld is not a multiple of per_sample %lld", (long long)d->numel, (long long)d->per_sample);
    if (d->n_terms < 0 || d->n_terms > NI_MAX_TERMS) return fail(NI_ERR_TOO_MANY, "ni_step: n_terms=%d exceeds NI_MAX_TERMS=%d (chain launches with accumulate=1)", d->n_terms, NI_MAX_TERMS);
    if (d->n_gen < 0 || d->n_gen > NI_MAX_GEN) return fail(NI_ERR_TOO_MANY, "ni_step: n_gen=%d exceeds NI_MAX_GEN=%d", d->n_gen, NI_MAX_GEN);
    if (d->x_next == nullptr) return fail(NI_ERR_INVALID, "ni_step: x_next is NULL");
    if (d->n_terms > 0 && (d->term_ptrs_host == nullptr || d->term_coeffs_host == nullptr)) return fail(NI_ERR_INVALID, "ni_step: term tables are NULL");
    if (d->has_x0) {
        if (d->out0 == nullptr) return fail(NI_ERR_INVALID, "ni_step: has_x0 but out0 is NULL");
        if (d->x_in == nullptr && d->a != 0.f) return fail(NI_ERR_INVALID, "ni_step: a != 0 but x_in is NULL");
        if (d->out_sample_stride < d->per_sample) return fail(NI_ERR_INVALID, "ni_step: out_sample_stride < per_sample");
        if (d->x_next == d->x_in || d->x_next == d->out0 || d->x_next == d->out1 || (d->x0_dst != nullptr && d->x0_dst == d->x_next))
            return fail(NI_ERR_INVALID, "ni_step: x_next aliases an input or x0_dst");
    }
    for (int i = 0; i < d->n_terms; ++i) {
        if (d->term_ptrs_host[i] == nullptr) return fail(NI_ERR_INVALID, "ni_step: term %d pointer is NULL", i);
        if (d->term_ptrs_host[i] == d->x_next) return fail(NI_ERR_INVALID, "ni_step: x_next aliases term %d", i);
    }
    if (d->numel == 0) return NI_OK;

    const int ds = dtype_size(d->dtype);
    if (!(d->dtype == NI_F32 || d->dtype == NI_F16 || d->dtype == NI_BF16)) return fail(NI_ERR_DTYPE, "ni_step: dtype %d not supported", d->dtype);
    const int od = d->has_x0 ? d->out_dtype : d->dtype;
    const bool combo = od == d->dtype || (d->dtype == NI_F32 && (od == NI_F16 || od == NI_BF16));
    if (!combo) return fail(NI_ERR_DTYPE, "ni_step: (dtype=%d, out_dtype=%d) not built", d->dtype, od);
    if (d->x_next_lp != nullptr && !(d->lp_dtype == NI_F16 || d->lp_dtype == NI_BF16)) return fail(NI_ERR_DTYPE, "ni_step: lp_dtype %d not supported", d->lp_dtype);

    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.per_sample = d->per_sample;
    a.out_sample_stride = d->has_x0 ? d->out_sample_stride : d->per_sample;
    a.out_strided = d->has_x0 && d->out_sample_stride != d->per_sample;
    a.x_in = (d->has_x0 && d->a != 0.f) ? d->x_in : nullptr;
    a.out0 = d->out0; a.out1 = d->out1;
    a.x0_dst = d->x0_dst; a.x_next = d->x_next; a.x_next_lp = d->x_next_lp; a.sumsq = d->sumsq;
    a.a = d->a; a.b0 = d->b0; a.b1 = d->b1; a.c_x0 = d->c_x0;
    a.k0 = (uint32_t)d->philox_seed; a.k1 = (uint32_t)(d->philox_seed >> 32);
    a.elem_offset = d->elem_offset;
    a.n_terms = d->n_terms; a.n_gen = d->n_gen;
    a.has_x0 = d->has_x0; a.accumulate = d->accumulate; a.lp_dtype = d->lp_dtype;
    for (int g = 0; g < d->n_gen; ++g) { a.gen_tid[g] = d->gen_tensor_ids[g]; a.gen_c[g] = d->gen_coeffs[g]; a.gen_dst[g] = d->gen_dst[g]; }

    // 128-bit path needs every pointer 16 B aligned (8 B for half outputs next to fp32 state) and vector-sized shapes
    const int VEC = 16 / ds;
    bool vec_ok = d->numel % VEC == 0 && d->per_sample % VEC == 0 && aligned16(d->x_next) && (d->n_gen == 0 || d->elem_offset % 4 == 0);
    if (d->has_x0) {
        const uintptr_t omask = (uintptr_t)(VEC * dtype_size(od) - 1);
        vec_ok = vec_ok && d->out_sample_stride % VEC == 0 && (reinterpret_cast<uintptr_t>(d->out0) & omask) == 0 &&
                 (d->out1 == nullptr || (reinterpret_cast<uintptr_t>(d->out1) & omask) == 0) && (a.x_in == nullptr || aligned16(a.x_in)) &&
                 (d->x0_dst == nullptr || aligned16(d->x0_dst));
    }
    for (int i = 0; i < d->n_terms && vec_ok; ++i) vec_ok = aligned16(d->term_ptrs_host[i]);
    for (int g = 0; g < d->n_gen && vec_ok; ++g) vec_ok = d->gen_dst[g] == nullptr || aligned16(d->gen_dst[g]);
    if (d->x_next_lp != nullptr) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(d->x_next_lp) & (uintptr_t)(VEC * 2 - 1)) == 0;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == NI_F32 && od == NI_F32) return launch_step<float, float>(a, d, vec_ok, st);
    if (d->dtype == NI_F32 && od == NI_F16) return launch_step<float, __half>(a, d, vec_ok, st);
    if (d->dtype == NI_F32 && od == NI_BF16) return launch_step<float, __nv_bfloat16>(a, d, vec_ok, st);
    if (d->dtype == NI_F16) return launch_step<__half, __half>(a, d, vec_ok, st);
    return launch_step<__nv_bfloat16, __nv_bfloat16>(a, d, vec_ok, st);
}

} // extern "C"

namespace {
template <typename TS, typename TD>
int launch_wsum(const void *const *src, const double *coeffs, int n, void *dst, int64_t numel, double scale, bool vec_ok, cudaStream_t st)
{
    static thread_local WsumTable<NI_MAX_TERMS> big;
    constexpr int VEC = 16 / (int)sizeof(TS);
    if (n <= 32) {
        WsumTable<32> tab;
        memset(&tab, 0, sizeof(tab));
        for (int i = 0; i < n; ++i) { tab.ptr[i] = src[i]; tab.c[i] = coeffs[i]; }
        if (vec_ok) {
            const int64_t nvec = numel / VEC;
            ni_wsum_kernel<TS, TD, VEC, 32><<<(unsigned)((nvec + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(tab, n, dst, nvec, scale);
        } else {
            ni_wsum_kernel<TS, TD, 1, 32><<<(unsigned)((numel + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(tab, n, dst, numel, scale);
        }
    } else {
        for (int i = 0; i < n; ++i) { big.ptr[i] = src[i]; big.c[i] = coeffs[i]; }
        if (vec_ok) {
            const int64_t nvec = numel / VEC;
            ni_wsum_kernel<TS, TD, VEC, NI_MAX_TERMS><<<(unsigned)((nvec + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(big, n, dst, nvec, scale);
        } else {
            ni_wsum_kernel<TS, TD, 1, NI_MAX_TERMS><<<(unsigned)((numel + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(big, n, dst, numel, scale);
        }
    }
    return check_launch("ni_weighted_sum launch");
}
} // namespace

extern "C" {

int ni_weighted_sum(const void *const *src, const double *coeffs, int n_terms, void *dst, int64_t numel, int src_dtype, int dst_dtype, double scale, void *stream)
{
    if (n_terms < 0 || n_terms > NI_MAX_TERMS) return fail(NI_ERR_TOO_MANY, "ni_weighted_sum: n_terms=%d exceeds NI_MAX_TERMS=%d", n_terms, NI_MAX_TERMS);
    if (numel == 0) return NI_OK;
    if (numel < 0 || dst == nullptr || (n_terms > 0 && (src == nullptr || coeffs == nullptr))) return fail(NI_ERR_INVALID, "ni_weighted_sum: bad arguments");
    for (int i = 0; i < n_terms; ++i) {
        if (src[i] == nullptr) return fail(NI_ERR_INVALID, "ni_weighted_sum: src[%d] is NULL", i);
        if (src[i] == dst) return fail(NI_ERR_INVALID, "ni_weighted_sum: dst aliases src[%d]", i);
    }
    const int ss = dtype_size(src_dtype), dsz = dtype_size(dst_dtype);
    if (ss == 0 || dsz == 0) return fail(NI_ERR_DTYPE, "ni_weighted_sum: unknown dtype");
    if (numel == 0) return NI_OK;
    const int VEC = 16 / ss;
    bool vec_ok = numel % VEC == 0 && (reinterpret_cast<uintptr_t>(dst) & (uintptr_t)(VEC * dsz >= 16 ? 15 : VEC * dsz - 1)) == 0;
    for (int i = 0; i < n_terms && vec_ok; ++i) vec_ok = aligned16(src[i]);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define NI_WS(S, D, TS, TD) if (src_dtype == S && dst_dtype == D) return launch_wsum<TS, TD>(src, coeffs, n_terms, dst, numel, scale, vec_ok, st);
    NI_WS(NI_F32, NI_F32, float, float)
    NI_WS(NI_F16, NI_F16, __half, __half)
    NI_WS(NI_BF16, NI_BF16, __nv_bfloat16, __nv_bfloat16)
    NI_WS(NI_F16, NI_F32, __half, float)
    NI_WS(NI_BF16, NI_F32, __nv_bfloat16, float)
    NI_WS(NI_F64, NI_F32, double, float)
    NI_WS(NI_F64, NI_F64, double, double)
#undef NI_WS
    return fail(NI_ERR_DTYPE, "ni_weighted_sum: (src=%d, dst=%d) not built", src_dtype, dst_dtype);
}

int ni_philox_normal(void *dst, int64_t numel, int dst_dtype, uint64_t seed, uint64_t tensor_id, uint64_t elem_offset, void *stream)
{
    if (numel == 0) return NI_OK;
    if (dst == nullptr || numel < 0) return fail(NI_ERR_INVALID, "ni_philox_normal: bad arguments");
    const int ds = dtype_size(dst_dtype);
    if (!(dst_dtype == NI_F32 || dst_dtype == NI_F16 || dst_dtype == NI_BF16)) return fail(NI_ERR_DTYPE, "ni_philox_normal: dtype %d not supported", dst_dtype);
    if (numel == 0) return NI_OK;
    const int VEC = 16 / ds;
    const bool vec_ok = numel % VEC == 0 && aligned16(dst) && elem_offset % 4 == 0;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nvec = vec_ok ? numel / VEC : numel;
    const unsigned blocks = (unsigned)((nvec + NI_BLOCK - 1) / NI_BLOCK);
    if (dst_dtype == NI_F32) {
        if (vec_ok) ni_normal_kernel<float, 4><<<blocks, NI_BLOCK, 0, st>>>((float *)dst, nvec, k0, k1, tensor_id, elem_offset);
        else ni_normal_kernel<float, 1><<<blocks, NI_BLOCK, 0, st>>>((float *)dst, nvec, k0, k1, tensor_id, elem_offset);
    } else if (dst_dtype == NI_F16) {
        if (vec_ok) ni_normal_kernel<__half, 8><<<blocks, NI_BLOCK, 0, st>>>((__half *)dst, nvec, k0, k1, tensor_id, elem_offset);
        else ni_normal_kernel<__half, 1><<<blocks, NI_BLOCK, 0, st>>>((__half *)dst, nvec, k0, k1, tensor_id, elem_offset);
    } else {
        if (vec_ok) ni_normal_kernel<__nv_bfloat16, 8><<<blocks, NI_BLOCK, 0, st>>>((__nv_bfloat16 *)dst, nvec, k0, k1, tensor_id, elem_offset);
        else ni_normal_kernel<__nv_bfloat16, 1><<<blocks, NI_BLOCK, 0, st>>>((__nv_bfloat16 *)dst, nvec, k0, k1, tensor_id, elem_offset);
    }
    return check_launch("ni_philox_normal launch");
}

int ni_to_pixel_u8(const void *x, int src_dtype, uint8_t *dst, int64_t batch, int channels, int height, int width, float scale, float shift, void *stream)
{
    if (x == nullptr || dst == nullptr || batch < 0 || channels <= 0 || height <= 0 || width <= 0) return fail(NI_ERR_INVALID, "ni_to_pixel_u8: bad arguments");
    if (batch == 0) return NI_OK;
    const int64_t HW = (int64_t)height * width, npix = batch * HW;
    const unsigned blocks = (unsigned)((npix + NI_BLOCK - 1) / NI_BLOCK);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (src_dtype == NI_F32) ni_pixel_kernel<float><<<blocks, NI_BLOCK, 0, st>>>((const float *)x, dst, npix, channels, HW, scale, shift);
    else if (src_dtype == NI_F16) ni_pixel_kernel<__half><<<blocks, NI_BLOCK, 0, st>>>((const __half *)x, dst, npix, channels, HW, scale, shift);
    else if (src_dtype == NI_BF16) ni_pixel_kernel<__nv_bfloat16><<<blocks, NI_BLOCK, 0, st>>>((const __nv_bfloat16 *)x, dst, npix, channels, HW, scale, shift);
    else return fail(NI_ERR_DTYPE, "ni_to_pixel_u8: dtype %d not supported", src_dtype);
    return check_launch("ni_to_pixel_u8 launch");
}

} // extern "C"
