// ni_common.cuh -- device helpers shared by every kernel of libni_b200.so: 128-bit loads/stores with cache
// policy, storage-type packing, the Philox4x32-10 + Box-Muller noise contract (include/ni_b200.h), and the
// host-side bookkeeping (error string, launch counter, options) that all translation units use.
#pragma once
#include "ni_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <type_traits>

// 128-bit load flavours (SASS): 0 plain ld.global (LDG.E.128), 1 ld.global.L1::no_allocate (LDG.E.NA.128), 2 ld.global.cs
// (evict-first), 4 plain + L2::256B prefetch.  Measured on B200 (profiles/r01_policy_sweep.txt): when a launch streams far
// more than the 126 MB L2 can hold (C2, C3) plain loads are 1.2% / 3.8% faster than NA loads -- NA-loaded lines are the
// first to leave L2, so the dirty lines of the stores pile up and drain in bursts -- while on launches whose tensors fit
// in L2 (SD3 first-order path, 33 MB tensors) NA loads are 11% faster because the x_{k+1} just written survives until the
// next step reads it.  The step kernels are therefore built in both flavours and the host picks per launch.
#ifndef NI_LOAD_POLICY
#define NI_LOAD_POLICY 1
#endif
#ifndef NI_STREAM_LOAD_POLICY
#define NI_STREAM_LOAD_POLICY 0 /* flavour of the step kernel's loads when the launch footprint is >> L2 */
#endif
#ifndef NI_STORE_POLICY
#define NI_STORE_POLICY 0 /* 0 plain st.global, 1 st.global.cs */
#endif
// Geometry of the direct-load kernels (profiles/r01_sweep.txt): 128-thread CTAs, registers capped at 48 (10 CTAs = 40
// warps per SM) and up to 8 independent 128-bit loads per thread.
#ifndef NI_BLOCK
#define NI_BLOCK 128
#endif
#ifndef NI_TERM_BATCH_MAX
#define NI_TERM_BATCH_MAX 8
#endif
#ifndef NI_MIN_BLOCKS
#define NI_MIN_BLOCKS 10
#endif

namespace ni {

// ---- host-side shared state (defined in ni_kernels.cu) ----------------------------------------------------
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);
int opt_pdl();
int opt_wide();
void count_lean_launch();
struct DevInfo { int sms; int64_t l2_bytes; };
const DevInfo &dev_info();
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int dtype_size(int d) { return d == NI_F32 ? 4 : d == NI_F64 ? 8 : (d == NI_F16 || d == NI_BF16) ? 2 : 0; }

// ---- raw vector loads / stores with cache policy ----------------------------------------------------------
template <int POL> __device__ __forceinline__ uint4 ld128_pol(const void *p)
{
    uint4 r;
    if constexpr (POL == 1) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if constexpr (POL == 2) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if constexpr (POL == 4) asm volatile("ld.global.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 256-bit global access: Blackwell only (PTX ld/st.global.v8.b32, SASS LDG.E.ENL2.256 / STG.E.ENL2.256); 32-byte aligned
struct Vec256 { uint32_t w[8]; };
template <int POL> __device__ __forceinline__ Vec256 ld256_pol(const void *p)
{
    Vec256 r;
    if constexpr (POL == 1)
        asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    else
        asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(void *p, const uint32_t (&w)[8])
{
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
template <int POL> __device__ __forceinline__ uint2 ld64_pol(const void *p)
{
    uint2 r;
    if constexpr (POL == 1) asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    else asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
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

template <typename T, int VEC, int POL = NI_LOAD_POLICY> __device__ __forceinline__ Raw<T, VEC> load_raw(const T *p)
{
    Raw<T, VEC> r;
    constexpr int BYTES = Raw<T, VEC>::BYTES;
    if constexpr (BYTES == 32) {
        const Vec256 a = ld256_pol<POL>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.w[i] = a.w[i];
    } else if constexpr (BYTES == 16) {
        uint4 a = ld128_pol<POL>(p);
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
    } else if constexpr (BYTES == 8) {
        uint2 a = ld64_pol<POL>(p);
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
        st256(p, r.w);
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

// acc += c * (VEC elements of T)
template <typename T, int VEC> __device__ __forceinline__ void fma_term(float (&acc)[VEC], const Raw<T, VEC> &r, float c)
{
    float f[VEC];
    unpack<T, VEC>(r, f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c, f[i], acc[i]);
}

// ---- Philox4x32-10 + Box-Muller (noise contract in include/ni_b200.h; CPU twin: oracle/philox_oracle.c) -----
// The ten round keys are the same for every thread of a launch: the host expands the seed once and they reach the
// kernel as constant-bank operands of the XORs (no per-thread key schedule).
struct PhiloxKeys {
    uint32_t k[20];
};
inline PhiloxKeys philox_keys(uint64_t seed)
{
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k[2 * r] = k0;
        K.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return K;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one IMAD.WIDE.U32 per product gives both halves
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ K.k[2 * r], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ K.k[2 * r + 1], (uint32_t)p0);
    }
    return c;
}

// -2 ln(u) for u in (0, 1].  MUFU.LG2 (lg2.approx) is relative-2^-22 accurate below 0.5 but only ABSOLUTE-2^-22
// accurate on (0.5, 2), useless where ln u -> 0.  There w = 1 - u is exact (Sterbenz) and
//   -2 ln(1 - w) = w (2 + w (1 + w (2/3 + w/2))) (1 + O(w^4 / 5)),
// used for w < 1/32 (relative error < 2e-7); above, the lg2 path is within 5.2e-6 relative, i.e. the radius
// sqrt(-2 ln u) >= 0.25 is within 6.5e-7 absolute.  9 instructions instead of ~28 for the precise logf; steps that draw
// fresh noise for every element (DDPM ancestral on the first-order path) are issue-bound by exactly this code.
// NI_PRECISE_NORMAL restores logf / sqrtf / sincospif.
__device__ __forceinline__ float neg2_log(float u)
{
    const float w = 1.0f - u;
    float p = fmaf(w, 0.5f, 0.66666668653488159f);
    p = fmaf(p, w, 1.0f);
    p = fmaf(p, w, 2.0f);
    const float t_small = p * w;
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    const float t_big = l2 * -1.3862943611198906f; // -2 ln 2
    return w < 0.03125f ? t_small : t_big;
}

// Box-Muller on two Philox words.  Radius: neg2_log + sqrt.approx (rel. 2^-23); angle: sin/cos.approx on an argument
// reduced to (-pi, pi] (abs. 2^-20.9) -- worst case |z - exact| < 5e-6 at the 6.7-sigma tail, ~3e-7 typical.
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float &za, float &zb)
{
    const float u = fmaf(__uint2float_rn(ra), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float v = fmaf(__uint2float_rn(rb), 4.6566128730773926e-10f, 2.3283064365386963e-10f);
#ifdef NI_PRECISE_NORMAL
    const float rad = sqrtf(-2.0f * logf(u));
    float s, c;
    sincospif(v, &s, &c);
    za = rad * c;
    zb = rad * s;
#else
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(neg2_log(u)));
    // cospi(v) = -cos(pi (v - 1)), sinpi(v) = -sin(pi (v - 1)); v in (0, 2] -> argument in (-pi, pi]
    const float ang = 3.14159265358979323846f * (v - 1.0f);
    za = -rad * __cosf(ang);
    zb = -rad * __sinf(ang);
#endif
}

__device__ __forceinline__ void normal4(uint64_t group, uint64_t tensor_id, const PhiloxKeys &K, float (&z)[4])
{
    const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)tensor_id, (uint32_t)(tensor_id >> 32)), K);
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
}

// VEC normals for global elements [e, e+VEC).  e need not be a multiple of 4 (an offset read from device memory cannot
// be checked on the host): the misaligned case draws one more Philox group and shifts -- a grid-uniform branch.
template <int VEC> __device__ __forceinline__ void normal_vec(uint64_t e, uint64_t tensor_id, const PhiloxKeys &K, float (&z)[VEC])
{
    if constexpr (VEC == 1) {
        float q[4];
        normal4(e >> 2, tensor_id, K, q);
        const int lane = (int)(e & 3);
        z[0] = lane == 0 ? q[0] : lane == 1 ? q[1] : lane == 2 ? q[2] : q[3];
    } else {
        const int sh = (int)(e & 3);
        if (sh == 0) {
#pragma unroll
            for (int j = 0; j < VEC / 4; ++j) {
                float q[4];
                normal4((e >> 2) + j, tensor_id, K, q);
#pragma unroll
                for (int i = 0; i < 4; ++i) z[4 * j + i] = q[i];
            }
        } else {
            float q[VEC + 4];
#pragma unroll
            for (int j = 0; j <= VEC / 4; ++j) {
                float t[4];
                normal4((e >> 2) + j, tensor_id, K, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) q[4 * j + i] = t[i];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) z[i] = sh == 1 ? q[i + 1] : sh == 2 ? q[i + 2] : q[i + 3];
        }
    }
}

// the element offset of a launch: host value plus (optionally) a counter in device memory, so a captured CUDA graph
// draws new noise on every replay (ni_counter_add advances the counter inside the graph)
__device__ __forceinline__ uint64_t effective_offset(uint64_t host_off, const uint64_t *dev_off)
{
    return dev_off == nullptr ? host_off : host_off + __ldg(dev_off);
}

// ---- tables ------------------------------------------------------------------------------------------------
template <int CAP> struct TermTable {
    const void *ptr[CAP];
    float c[CAP];
};

// programmatic dependent launch (PDL): let the next grid start launching; do not touch global memory before the
// previous grid is complete
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename Kern, typename... Args>
inline void launch_pdl(Kern kern, unsigned blocks, unsigned threads, size_t smem, cudaStream_t st, bool pdl, const Args &...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
}

// ni_step_lean.cuh: the specialised step kernels.  *used = false when the launch is not eligible (generic kernel takes it).
template <typename T, typename TO>
int launch_step_lean(const NiStepDesc *d, const void *x_in_eff, bool stream, cudaStream_t st, bool *used);

} // namespace ni
