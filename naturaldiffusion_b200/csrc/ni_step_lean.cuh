// ni_step_lean.cuh -- the production instantiations of the fused Natural Inference step (ni_step, include/ni_b200.h).
//
// Same arithmetic, same accumulation order and therefore the same bits as the generic kernel in ni_kernels.cu
// (tests/test_gpu_lean.py::test_lean_kernel_is_bit_identical_to_generic), with the per-thread overhead removed.  The
// generic kernel executes ~210 SASS instructions per thread on a 7-tensor step (64-bit index arithmetic, a 64-bit
// division per thread for the sample index, runtime loops over constant-bank tables); at 30 instructions per 16 B moved
// the SMs burn enough power for the 1 kW cap to pull the clock down in sustained runs.  Here:
//   * the row shape is a template parameter: NT stored terms (0..8, or a runtime loop), NG generated noise terms (0/1,
//     or a runtime loop), M model outputs (1/2) -- every pointer and coefficient is a constant-bank operand;
//   * 32-bit vector indices (one IMAD.WIDE per address);
//   * the sample index (strided model outputs, per-sample norms) comes from blockIdx through a multiply-shift when every
//     CTA lies inside one sample (it does for every named shape), else one 32-bit division;
//   * Philox round keys expanded on the host; -2 ln u through MUFU.LG2 with a series near u = 1 (ni_common.cuh);
//   * the fused uint8 output stage (last step) has its own instantiation in which a thread owns the C = 3 channel
//     planes of 4 consecutive pixels, the warp stages its 384 bytes in shared memory and stores them as 24 x 16 B.
//   * 128-bit or, on Blackwell, 256-bit global accesses per thread per tensor (NI_LEAN_WIDE below).
// Still an HBM-streaming kernel: no tensor cores, one vector load per tensor per thread, all issued before the first FMA.
#pragma once
#include "ni_common.cuh"

namespace ni {
namespace {

#ifndef NI_LEAN_MIN_BLOCKS
#define NI_LEAN_MIN_BLOCKS 10
#endif
// Bytes of the storage type each thread moves per tensor: 16 (LDG.E.128) or 32 (Blackwell's 256-bit global access,
// PTX ld/st.global.v8.b32 -> SASS LDG.E.ENL2.256 / STG.E.ENL2.256).  Measured on B200 (profiles/r02_vec256.jsonl): with
// fp32 state the 32-byte kernels win everywhere -- half the threads, half the per-thread overhead per byte: C2 +0.7 %,
// C3 +0.4 %, C4 dense +0.6 %, and +34 % on the DDPM-250 first-order step whose cost is the in-kernel noise generator --
// with fp16 state (16 elements per thread) they gain 1.3-2.3 % on the launches that stream far more than the L2 holds (SD3 dense
// rows, B 256) and lose 4-12 % on the L2-resident ones (SD3 first-order / sharp at B 64: 33 MB tensors).  So: fp32 state takes
// the 256-bit instantiation whenever everything is 32-byte aligned, 16-bit state only for launches that get the streaming load
// flavour (launch_streams()).  NI_LEAN_WIDE=0 builds without the 256-bit kernels.
#ifndef NI_LEAN_WIDE
#define NI_LEAN_WIDE 1
#endif

// everything a launch needs besides the term table; 32-bit addressing
struct LeanArgs {
    uint32_t nvec;            // vectors of VEC elements (< 2^31)
    uint32_t vec_per_sample;  // per_sample / VEC
    uint32_t out_extra_vec;   // (out_sample_stride - per_sample) / VEC: out vector index = v + sample * out_extra_vec
    uint32_t tile_in_sample;  // 1: every CTA lies inside one sample -> sample = (blockIdx.x * tps_mul) >> (32 + tps_shr)
    uint32_t tps_mul, tps_shr;
    uint32_t hw_vec;          // PIX: vectors per channel plane
    uint32_t n_pix_threads;   // PIX: batch * hw_vec
    const void *x_in, *out0, *out1;
    void *x0_dst, *x_next, *x_next_lp;
    float *sumsq;
    uint8_t *pixels;
    void *gen_dst[NI_MAX_GEN];
    uint64_t gen_tid[NI_MAX_GEN];
    float gen_c[NI_MAX_GEN];
    uint64_t elem_offset;
    const uint64_t *elem_offset_dev;
    PhiloxKeys keys;
    float a, b0, b1, c_x0, c_xin, bias, px_scale, px_shift;
    int n_terms, n_gen, lp_dtype, accumulate;
};

// One vector (VEC elements) of one step: every load issued before the first FMA, the sum in table order, generated noise,
// the x0 stage, the stores of x0_k / x_{k+1} / the low-precision copy.  Leaves x_{k+1} (fp32, before storage rounding) in acc.
// NT >= 0: exactly NT stored terms, NG/M exact.  NT < 0: runtime n_terms / n_gen, M still exact.
template <typename T, typename TO, int NT, int NG, int M, int POL, int CAP, int VB>
__device__ __forceinline__ void lean_vector(const LeanArgs &s, const TermTable<CAP> &tab, uint32_t v, uint32_t vo, float (&acc)[VB / (int)sizeof(T)])
{
    constexpr int VEC = VB / (int)sizeof(T);
    const bool has_x = s.x_in != nullptr;
    Raw<TO, VEC> ro0, ro1;
    Raw<T, VEC> rx;
    ro0 = load_raw<TO, VEC, POL>(static_cast<const TO *>(s.out0) + (size_t)vo * VEC);
    if constexpr (M == 2) ro1 = load_raw<TO, VEC, POL>(static_cast<const TO *>(s.out1) + (size_t)vo * VEC);
    if (has_x) rx = load_raw<T, VEC, POL>(static_cast<const T *>(s.x_in) + (size_t)v * VEC);
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;

    // stored terms, table order
    if constexpr (NT >= 0) {
        Raw<T, VEC> rr[NT > 0 ? NT : 1];
#pragma unroll
        for (int t = 0; t < NT; ++t) rr[t] = load_raw<T, VEC, POL>(static_cast<const T *>(tab.ptr[t]) + (size_t)v * VEC);
#pragma unroll
        for (int t = 0; t < NT; ++t) fma_term<T, VEC>(acc, rr[t], tab.c[t]);
    } else {
        const int n = s.n_terms;
        if (s.accumulate) unpack<T, VEC>(load_raw<T, VEC, POL>(static_cast<const T *>(s.x_next) + (size_t)v * VEC), acc);
        int t = 0;
#define NI_TERM_BATCH(NB)                                                                                                          \
    for (; t + NB <= n; t += NB) {                                                                                                 \
        Raw<T, VEC> rr[NB];                                                                                                        \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) rr[j] = load_raw<T, VEC, POL>(static_cast<const T *>(tab.ptr[t + j]) + (size_t)v * VEC); \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) fma_term<T, VEC>(acc, rr[j], tab.c[t + j]);                                 \
    }
        NI_TERM_BATCH(8)
        NI_TERM_BATCH(4)
        NI_TERM_BATCH(2)
        NI_TERM_BATCH(1)
#undef NI_TERM_BATCH
    }

    // generated noise (pure ALU; overlaps loads still in flight)
    const int n_gen = NT >= 0 ? NG : s.n_gen;
    if (n_gen > 0) {
        const uint64_t eoff = effective_offset(s.elem_offset, s.elem_offset_dev);
#pragma unroll
        for (int g = 0; g < (NT >= 0 ? NG : NI_MAX_GEN); ++g) {
            if (g < n_gen) {
                float z[VEC];
                normal_vec<VEC>(eoff + (uint64_t)v * VEC, s.gen_tid[g], s.keys, z);
                if (s.gen_dst[g] != nullptr) {
                    store_raw<T, VEC>(static_cast<T *>(s.gen_dst[g]) + (size_t)v * VEC, pack<T, VEC>(z));
#pragma unroll
                    for (int j = 0; j < VEC; ++j) z[j] = round_to<T>(z[j]);
                }
                const float c = s.gen_c[g];
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] = fmaf(c, z[j], acc[j]);
            }
        }
    }

    // x0 = a*x + b0*out0 + b1*out1, kept in the ring, enters the sum with A[k,k]; first-order rows add c_xin * x_k
    float x0[VEC], f[VEC];
    unpack<TO, VEC>(ro0, f);
#pragma unroll
    for (int j = 0; j < VEC; ++j) x0[j] = s.b0 * f[j];
    if constexpr (M == 2) {
        unpack<TO, VEC>(ro1, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) x0[j] = fmaf(s.b1, f[j], x0[j]);
    }
    if (has_x) {
        unpack<T, VEC>(rx, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) x0[j] = fmaf(s.a, f[j], x0[j]);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) x0[j] = round_to<T>(x0[j]);
    if (s.x0_dst != nullptr) store_raw<T, VEC>(static_cast<T *>(s.x0_dst) + (size_t)v * VEC, pack<T, VEC>(x0));
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = fmaf(s.c_x0, x0[j], acc[j]);
    if (has_x && s.c_xin != 0.f) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = fmaf(s.c_xin, f[j], acc[j]);
    }
    if (s.bias != 0.f) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += s.bias;
    }
    if (s.x_next != nullptr) store_raw<T, VEC>(static_cast<T *>(s.x_next) + (size_t)v * VEC, pack<T, VEC>(acc));
    if (s.x_next_lp != nullptr) {
        if (s.lp_dtype == NI_BF16) store_raw<__nv_bfloat16, VEC>(static_cast<__nv_bfloat16 *>(s.x_next_lp) + (size_t)v * VEC, pack<__nv_bfloat16, VEC>(acc));
        else store_raw<__half, VEC>(static_cast<__half *>(s.x_next_lp) + (size_t)v * VEC, pack<__half, VEC>(acc));
    }
}

template <typename T, typename TO, int NT, int NG, int M, int POL, bool PIX, int CAP, int VB>
__global__ void __launch_bounds__(NI_BLOCK, (VB == 32 ? 6 : PIX ? 8 : (NG >= 1 && NT >= 6) ? 8 : NI_LEAN_MIN_BLOCKS)) ni_step_lean_kernel(const __grid_constant__ LeanArgs s, const __grid_constant__ TermTable<CAP> tab)
{
    static_assert(!(PIX && VB != 16), "the output-stage instantiation works on 4-pixel groups");
    constexpr int VEC = VB / (int)sizeof(T);
    pdl_launch_dependents();

    if constexpr (PIX) {
        // fused output stage of the LAST step: NCHW float -> NHWC uint8 with the reference's truncating cast.  Thread g owns
        // pixels [4q, 4q+4) of sample n: it walks the three channel planes one after the other (each pass is an ordinary
        // vector of the step), keeps 4 bytes per channel, and the warp's 384 output bytes go out as 24 x 16 B.
        __shared__ __align__(16) uint32_t stage[NI_BLOCK * 3];
        const uint32_t g = blockIdx.x * NI_BLOCK + threadIdx.x;
        if (g >= s.n_pix_threads) return; // host guarantees whole warps (hw_vec % 32 == 0)
        const uint32_t sample = g / s.hw_vec;
        const uint32_t q = g - sample * s.hw_vec;
        pdl_wait();
        uint8_t b[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t v = sample * s.vec_per_sample + c * s.hw_vec + q;
            float acc[VEC];
            lean_vector<T, TO, NT, NG, M, POL, CAP, VB>(s, tab, v, v + sample * s.out_extra_vec, acc);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float y = (round_to<T>(acc[p]) * s.px_scale + s.px_shift) * 255.0f;
                b[p * 3 + c] = (uint8_t)(int)fminf(fmaxf(y, 0.f), 255.f);
            }
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t *ws = stage + warp * 96;
#pragma unroll
        for (int w = 0; w < 3; ++w) ws[lane * 3 + w] = (uint32_t)b[4 * w] | ((uint32_t)b[4 * w + 1] << 8) | ((uint32_t)b[4 * w + 2] << 16) | ((uint32_t)b[4 * w + 3] << 24);
        __syncwarp();
        if (lane < 24) {
            const uint32_t g0 = blockIdx.x * NI_BLOCK + warp * 32; // first thread of this warp
            const uint4 o = *reinterpret_cast<const uint4 *>(ws + lane * 4);
            st128(s.pixels + (size_t)g0 * 12 + lane * 16, o);
        }
        return;
    } else {
        const uint32_t v = blockIdx.x * NI_BLOCK + threadIdx.x;
        if (v >= s.nvec) return;
        uint32_t sample = 0;
        if (s.out_extra_vec != 0 || s.sumsq != nullptr) {
            if (s.tile_in_sample) sample = s.tps_mul == 0 ? blockIdx.x : (__umulhi(blockIdx.x, s.tps_mul) >> s.tps_shr);
            else sample = v / s.vec_per_sample;
        }
        pdl_wait();
        float acc[VEC];
        lean_vector<T, TO, NT, NG, M, POL, CAP, VB>(s, tab, v, v + sample * s.out_extra_vec, acc);

        // per-sample sum of squares: warp shuffle, one atomic per warp (per lane only where a warp straddles samples)
        if (s.sumsq != nullptr) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float r = round_to<T>(acc[j]);
                ss = fmaf(r, r, ss);
            }
            const unsigned mask = __activemask();
            bool fast = mask == 0xffffffffu;
            if (fast) {
                const uint32_t s0 = __shfl_sync(0xffffffffu, sample, 0);
                fast = __all_sync(0xffffffffu, sample == s0);
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
}

// ---- host dispatch ---------------------------------------------------------------------------------------------

// q = n / d for 0 <= n < 2^31 as (umulhi(n, mul) >> shr); mul == 0 means d == 1
void fast_divisor(uint32_t d, uint32_t *mul, uint32_t *shr)
{
    if (d <= 1) { *mul = 0; *shr = 0; return; }
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg; // ceil(log2 d)
    const int p = 31 + (int)lg;
    *mul = (uint32_t)(((1ull << p) + d - 1) / d);
    *shr = (uint32_t)(p - 32);
}

template <typename T, typename TO, int NT, int NG, int M, int POL, bool PIX, int CAP, int VB = 16>
int launch_one(const LeanArgs &a, const NiStepDesc *d, cudaStream_t st)
{
    TermTable<CAP> tab;
    memset(&tab, 0, sizeof(tab));
    for (int i = 0; i < d->n_terms; ++i) { tab.ptr[i] = d->term_ptrs_host[i]; tab.c[i] = d->term_coeffs_host[i]; }
    const uint32_t work = PIX ? a.n_pix_threads : a.nvec;
    launch_pdl(ni_step_lean_kernel<T, TO, NT, NG, M, POL, PIX, CAP, VB>, (work + NI_BLOCK - 1) / NI_BLOCK, NI_BLOCK, 0, st, opt_pdl() != 0, a, tab);
    count_lean_launch();
    return check_launch("ni_step (lean) launch");
}

template <typename T, typename TO, int NT, int NG, int M, int CAP, int VB>
int launch_pol(const LeanArgs &a, const NiStepDesc *d, bool stream, cudaStream_t st)
{
    if (stream) return launch_one<T, TO, NT, NG, M, NI_STREAM_LOAD_POLICY, false, CAP, VB>(a, d, st);
    return launch_one<T, TO, NT, NG, M, NI_LOAD_POLICY, false, CAP, VB>(a, d, st);
}

template <typename T, typename TO, int NT, int M, int VB>
int launch_ng(const LeanArgs &a, const NiStepDesc *d, bool stream, cudaStream_t st)
{
    if (d->n_gen == 0) return launch_pol<T, TO, NT, 0, M, (NT > 0 ? NT : 1), VB>(a, d, stream, st);
    return launch_pol<T, TO, NT, 1, M, (NT > 0 ? NT : 1), VB>(a, d, stream, st);
}

template <typename T, typename TO, int M, int VB>
int launch_nt(const LeanArgs &a, const NiStepDesc *d, bool stream, cudaStream_t st)
{
    const bool exact = d->n_terms <= 8 && d->n_gen <= 1 && !d->accumulate;
    if constexpr (std::is_same<T, TO>::value) { // mixed storage/output dtypes: runtime-loop instantiation only
        if (exact) {
            switch (d->n_terms) {
            case 0: return launch_ng<T, TO, 0, M, VB>(a, d, stream, st);
            case 1: return launch_ng<T, TO, 1, M, VB>(a, d, stream, st);
            case 2: return launch_ng<T, TO, 2, M, VB>(a, d, stream, st);
            case 3: return launch_ng<T, TO, 3, M, VB>(a, d, stream, st);
            case 4: return launch_ng<T, TO, 4, M, VB>(a, d, stream, st);
            case 5: return launch_ng<T, TO, 5, M, VB>(a, d, stream, st);
            case 6: return launch_ng<T, TO, 6, M, VB>(a, d, stream, st);
            case 7: return launch_ng<T, TO, 7, M, VB>(a, d, stream, st);
            default: return launch_ng<T, TO, 8, M, VB>(a, d, stream, st);
            }
        }
    }
    if (d->n_terms <= 32) return launch_pol<T, TO, -1, 0, M, 32, VB>(a, d, stream, st);
    return launch_pol<T, TO, -1, 0, M, NI_MAX_TERMS, VB>(a, d, stream, st);
}

} // namespace

// Called by ni_step after validation when the 128-bit path applies.  *used = false when this launch is not eligible
// (the generic kernel takes it): no x0 stage, > 2^31 vectors, or an output stage the lean kernels do not implement.
template <typename T, typename TO>
int launch_step_lean(const NiStepDesc *d, const void *x_in_eff, bool stream, cudaStream_t st, bool *used)
{
    const bool pix = d->pixels_u8 != nullptr;
    *used = false;
    if (!d->has_x0 || d->out0 == nullptr) return NI_OK;
    // 256-bit accesses: 32-byte alignment and whole 32-byte vectors everywhere (ni_step checked the 16-byte conditions)
    bool wide = false;
    if constexpr (NI_LEAN_WIDE && std::is_same<T, TO>::value) {
        auto al = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
        constexpr int V32 = 32 / (int)sizeof(T);
        wide = !pix && opt_wide() != 0 && (sizeof(T) == 4 || stream) && d->numel % V32 == 0 && d->per_sample % V32 == 0 && d->out_sample_stride % V32 == 0 &&
               al(x_in_eff) && al(d->out0) && al(d->out1) && al(d->x0_dst) && al(d->x_next) && al(d->x_next_lp);
        for (int i = 0; i < d->n_terms && wide; ++i) wide = al(d->term_ptrs_host[i]);
        for (int g = 0; g < d->n_gen && wide; ++g) wide = al(d->gen_dst[g]);
    }
    const int VEC = (wide ? 32 : 16) / (int)sizeof(T);
    const int64_t nvec = d->numel / VEC;
    const int64_t batch = d->numel / d->per_sample;
    const int64_t out_vec_total = batch * (d->out_sample_stride / VEC);
    if (nvec >= (1ll << 31) || out_vec_total >= (1ll << 32)) return NI_OK;

    LeanArgs a;
    memset(&a, 0, sizeof(a));
    a.nvec = (uint32_t)nvec;
    a.vec_per_sample = (uint32_t)(d->per_sample / VEC);
    a.out_extra_vec = (uint32_t)((d->out_sample_stride - d->per_sample) / VEC);
    a.tile_in_sample = a.vec_per_sample % NI_BLOCK == 0 ? 1u : 0u;
    if (a.tile_in_sample) fast_divisor(a.vec_per_sample / NI_BLOCK, &a.tps_mul, &a.tps_shr);
    a.x_in = x_in_eff; a.out0 = d->out0; a.out1 = d->out1;
    a.x0_dst = d->x0_dst; a.x_next = d->x_next; a.x_next_lp = d->x_next_lp; a.sumsq = d->sumsq; a.pixels = d->pixels_u8;
    for (int g = 0; g < d->n_gen; ++g) { a.gen_dst[g] = d->gen_dst[g]; a.gen_tid[g] = d->gen_tensor_ids[g]; a.gen_c[g] = d->gen_coeffs[g]; }
    a.elem_offset = d->elem_offset; a.elem_offset_dev = d->elem_offset_dev;
    a.keys = philox_keys(d->philox_seed);
    a.a = d->a; a.b0 = d->b0; a.b1 = d->b1; a.c_x0 = d->c_x0; a.c_xin = d->c_xin; a.bias = d->bias; a.px_scale = d->px_scale; a.px_shift = d->px_shift;
    a.n_terms = d->n_terms; a.n_gen = d->n_gen; a.lp_dtype = d->lp_dtype; a.accumulate = d->accumulate;

    if (d->pixels_u8 != nullptr) {
        // output-stage instantiation: fp32 state, 3 channels, whole warps per channel plane, <= 8 stored terms, no extras
        if constexpr (std::is_same<T, float>::value && std::is_same<TO, float>::value) {
            const int64_t hw_vec = d->per_sample / 3 / VEC;
            const bool ok = d->px_channels == 3 && d->per_sample % (3 * VEC) == 0 && hw_vec % 32 == 0 && d->n_terms <= 8 && d->n_gen == 0 && !d->accumulate &&
                            d->sumsq == nullptr && d->x_next_lp == nullptr && aligned16(d->pixels_u8) && batch * hw_vec < (1ll << 31);
            if (!ok) return NI_OK;
            a.hw_vec = (uint32_t)hw_vec;
            a.n_pix_threads = (uint32_t)(batch * hw_vec);
            *used = true;
#define NI_PIX(NTV)                                                                                                                        \
    case NTV:                                                                                                                              \
        if (d->out1 != nullptr)                                                                                                            \
            return stream ? launch_one<float, float, NTV, 0, 2, NI_STREAM_LOAD_POLICY, true, (NTV > 0 ? NTV : 1)>(a, d, st)                \
                          : launch_one<float, float, NTV, 0, 2, NI_LOAD_POLICY, true, (NTV > 0 ? NTV : 1)>(a, d, st);                      \
        return stream ? launch_one<float, float, NTV, 0, 1, NI_STREAM_LOAD_POLICY, true, (NTV > 0 ? NTV : 1)>(a, d, st)                    \
                      : launch_one<float, float, NTV, 0, 1, NI_LOAD_POLICY, true, (NTV > 0 ? NTV : 1)>(a, d, st);
            switch (d->n_terms) {
                NI_PIX(0) NI_PIX(1) NI_PIX(2) NI_PIX(3) NI_PIX(4) NI_PIX(5) NI_PIX(6) NI_PIX(7)
            default:
                NI_PIX(8)
            }
#undef NI_PIX
        }
        return NI_OK;
    }
    *used = true;
    if constexpr (NI_LEAN_WIDE && std::is_same<T, TO>::value) {
        if (wide) return d->out1 != nullptr ? launch_nt<T, TO, 2, 32>(a, d, stream, st) : launch_nt<T, TO, 1, 32>(a, d, stream, st);
    }
    if (d->out1 != nullptr) return launch_nt<T, TO, 2, 16>(a, d, stream, st);
    return launch_nt<T, TO, 1, 16>(a, d, stream, st);
}

} // namespace ni
