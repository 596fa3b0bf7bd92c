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
#include "ni_common.cuh"

#include <cstdarg>
#include <cstdio>

// launch_streams(): L2-friendly loads iff written <= NI_L2_KEEP_NUM/NI_L2_KEEP_DEN of the L2 and written >= traffic / NI_L2_KEEP_SHARE
#ifndef NI_L2_KEEP_NUM
#define NI_L2_KEEP_NUM 3
#endif
#ifndef NI_L2_KEEP_DEN
#define NI_L2_KEEP_DEN 5
#endif
#ifndef NI_L2_KEEP_SHARE
#define NI_L2_KEEP_SHARE 16
#endif

namespace ni {

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_lean_launches{0};
// tuning knobs (ni_set_option): which step kernel, and the TMA kernel's geometry
std::atomic<int> g_variant{0};       // 0 auto (lean kernels, generic fallback), 1 always the generic direct-load kernel, 2 the TMA kernel whenever eligible
std::atomic<int> g_tma_max_stages{32};
std::atomic<int> g_tma_warps{8};
std::atomic<int> g_tma_smem_kb{200};
std::atomic<int> g_tma_ctas_per_sm{2};
std::atomic<int> g_tma_tile_kb{2};   // bytes per source per stage: 2 or 4 KB
std::atomic<int> g_tma_l2_hint{0};   // cp.async.bulk L2 cache hint: 0 none, 1 evict_first, 2 evict_last, 3 evict_normal
std::atomic<int> g_tma_dynamic{0};   // 1: tiles claimed from a global counter instead of blockIdx.x + q * gridDim.x
std::atomic<int> g_wide{1};          // fp32 state: 256-bit (LDG.E.ENL2.256) instantiations of the specialised step kernels when eligible
std::atomic<int> g_pdl{1};           // programmatic dependent launch for the direct-load step kernels
std::atomic<int> g_load_policy{0};   // step-kernel load flavour: 0 by launch footprint vs L2, 1 always L2-friendly (NA), 2 always streaming
} // namespace

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(NI_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NI_OK;
}

int opt_pdl() { return g_pdl.load(std::memory_order_relaxed); }
int opt_wide() { return g_wide.load(std::memory_order_relaxed); }
void count_lean_launch() { g_lean_launches.fetch_add(1, std::memory_order_relaxed); }

const DevInfo &dev_info()
{
    static thread_local int cached_dev = -1;
    static thread_local DevInfo info = {148, 126 << 20};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev) {
        int sms = 0, l2 = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        info.sms = sms > 0 ? sms : 148;
        info.l2_bytes = l2 > 0 ? l2 : (126 << 20);
        cached_dev = dev;
    }
    return info;
}

namespace {

// ------------------------------------------------------------------------------------------
// the fused step
// ------------------------------------------------------------------------------------------
struct StepArgs {
    int64_t nvec; // vectors of VEC elements
    int64_t per_sample, out_sample_stride;
    const void *x_in, *out0, *out1;
    void *x0_dst, *x_next, *x_next_lp;
    float *sumsq;
    void *gen_dst[NI_MAX_GEN];
    uint64_t gen_tid[NI_MAX_GEN];
    uint64_t elem_offset;
    const uint64_t *elem_offset_dev;
    float gen_c[NI_MAX_GEN];
    float a, b0, b1, c_x0, c_xin, bias, px_scale, px_shift;
    uint8_t *pixels;
    int px_channels;
    PhiloxKeys keys;
    int n_terms, n_gen;
    int has_x0, out_strided, accumulate, lp_dtype;
};

// Everything after the stored terms are summed (shared by the LDG and the TMA kernels so both produce the same bits):
// generated noise, x0 conversion + ring store, x_{k+1} store(s), per-sample sum of squares.
template <typename T, typename TO, int VEC>
__device__ __forceinline__ void step_epilogue(const StepArgs &s, int64_t e, int64_t sample, float (&acc)[VEC], const Raw<TO, VEC> &ro0,
                                              const Raw<TO, VEC> &ro1, const Raw<T, VEC> &rx, bool has_x0, bool has_x, bool has_o1)
{
    // generated noise (pure ALU; overlaps loads still in flight)
    const uint64_t eoff = s.n_gen > 0 ? effective_offset(s.elem_offset, s.elem_offset_dev) : 0;
    for (int g = 0; g < s.n_gen; ++g) {
        float z[VEC];
        normal_vec<VEC>(eoff + (uint64_t)e, s.gen_tid[g], s.keys, z);
        if (s.gen_dst[g] != nullptr) {
            store_raw<T, VEC>(static_cast<T *>(s.gen_dst[g]) + e, pack<T, VEC>(z));
#pragma unroll
            for (int i = 0; i < VEC; ++i) z[i] = round_to<T>(z[i]);
        }
        const float c = s.gen_c[g];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c, z[i], acc[i]);
    }

    // x0 = a*x + b0*out0 + b1*out1, kept in the ring, enters the sum with A[k,k]
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
        // first-order (Markov) rows: the whole history collapses into c_xin * x_k, which is already in registers
        if (has_x && s.c_xin != 0.f) {
            unpack<T, VEC>(rx, f);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = fmaf(s.c_xin, f[i], acc[i]);
        }
    }

    // x_{k+1} (+ constant: the latent shift of the output stage, 0 otherwise)
    if (s.bias != 0.f) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += s.bias;
    }
    if (s.x_next != nullptr) store_raw<T, VEC>(static_cast<T *>(s.x_next) + e, pack<T, VEC>(acc));
    // fused output stage of the LAST step: NCHW float -> NHWC uint8 with the reference's truncating cast.  The VEC
    // elements of a thread are consecutive w of one channel, so they land C bytes apart.
    if (s.pixels != nullptr) {
        const int64_t n = e / s.per_sample, r = e - n * s.per_sample;
        const int64_t hw_total = s.per_sample / s.px_channels;
        const int64_t c = r / hw_total, hw = r - c * hw_total;
        uint8_t *dst = s.pixels + (n * hw_total + hw) * s.px_channels + c;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float y = (round_to<T>(acc[i]) * s.px_scale + s.px_shift) * 255.0f;
            dst[(int64_t)i * s.px_channels] = (uint8_t)(int)fminf(fmaxf(y, 0.f), 255.f);
        }
    }
    if (s.x_next_lp != nullptr) {
        if (s.lp_dtype == NI_BF16) store_raw<__nv_bfloat16, VEC>(static_cast<__nv_bfloat16 *>(s.x_next_lp) + e, pack<__nv_bfloat16, VEC>(acc));
        else store_raw<__half, VEC>(static_cast<__half *>(s.x_next_lp) + e, pack<__half, VEC>(acc));
    }

    // per-sample sum of squares: warp shuffle, one atomic per warp (per lane only where a warp straddles samples)
    if (s.sumsq != nullptr) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float r = round_to<T>(acc[i]);
            ss = fmaf(r, r, ss);
        }
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

// ---- variant 1: direct 128-bit global loads, one vector per thread ---------------------------------
// STREAM selects the load flavour (cache hints only, same arithmetic): true when the launch footprint is >> L2
template <typename T, typename TO, int VEC, int CAP, bool STREAM>
__global__ void __launch_bounds__(NI_BLOCK, NI_MIN_BLOCKS) ni_step_kernel(const __grid_constant__ StepArgs s, const __grid_constant__ TermTable<CAP> tab)
{
    constexpr int POL = STREAM ? NI_STREAM_LOAD_POLICY : NI_LOAD_POLICY;
    pdl_launch_dependents();
    const int64_t v = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (v >= s.nvec) return;
    const int64_t e = v * VEC; // first element of this thread
    pdl_wait();

    // issue the x0-stage loads (consumed after the term loop)
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
        ro0 = load_raw<TO, VEC, POL>(static_cast<const TO *>(s.out0) + eo);
        if (has_o1) ro1 = load_raw<TO, VEC, POL>(static_cast<const TO *>(s.out1) + eo);
        if (has_x) rx = load_raw<T, VEC, POL>(static_cast<const T *>(s.x_in) + e);
    }

    float acc[VEC];
    if (s.accumulate) {
        unpack<T, VEC>(load_raw<T, VEC, POL>(static_cast<const T *>(s.x_next) + e), acc);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    }

    // stored terms: batches of independent 128-bit loads, then the FMAs
    const int n = s.n_terms;
    int t = 0;
#define NI_TERM_BATCH(NB)                                                                                                     \
    for (; t + NB <= n; t += NB) {                                                                                            \
        Raw<T, VEC> rr[NB];                                                                                                   \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) rr[j] = load_raw<T, VEC, POL>(static_cast<const T *>(tab.ptr[t + j]) + e); \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) fma_term<T, VEC>(acc, rr[j], tab.c[t + j]);                            \
    }
#if NI_TERM_BATCH_MAX >= 8
    NI_TERM_BATCH(8)
#endif
#if NI_TERM_BATCH_MAX >= 4
    NI_TERM_BATCH(4)
#endif
#if NI_TERM_BATCH_MAX >= 2
    NI_TERM_BATCH(2)
#endif
    NI_TERM_BATCH(1)
#undef NI_TERM_BATCH

    step_epilogue<T, TO, VEC>(s, e, sample, acc, ro0, ro1, rx, has_x0, has_x, has_o1);
}

// ---- variant 2: TMA bulk copies into a shared-memory ring, warp-specialised -------------------------
// Persistent CTAs: a producer warp streams [source][tile] slabs global->shared with cp.async.bulk (one lane per
// source tensor, completion counted on an mbarrier, optional L2 cache-hint operand), NW consumer warps each own whole
// tiles (TILE_BYTES per source = 4 or 8 vectors per lane), so consumer warps run decoupled from each other with 4 / 8
// independent accumulation chains per thread.  A consumer reads the stored terms from shared memory, runs the same
// epilogue as the direct-load kernels (reading out0/out1/x_k from shared memory when it needs them), stores straight to
// global and hands the stage back.  stages x n_src x TILE_BYTES (<= ~200 KB) of loads are in flight per SM regardless of
// register pressure.  Source order in a stage: out0, [out1], [x_k], term 0..n-1.
// Tiles are handed out statically (tile = blockIdx.x + q * gridDim.x) or, with "tma_dynamic", claimed from a global
// counter by the producer (atomicInc that wraps after ntiles + gridDim claims: the counter is back at 0 when the grid
// ends, no per-launch reset), the claimed index travelling to the consumers in shared memory next to the stage.
constexpr int TMA_MAX_WARPS = 16;
constexpr int TMA_MAX_STAGES = 32;
constexpr int TMA_MAX_SRC = 32;
constexpr int TMA_BAR_BYTES = 2 * TMA_MAX_STAGES * 8 + TMA_MAX_STAGES * 4; // full + empty barriers, claimed tile per stage
constexpr int TMA_COUNTERS = 64;
__device__ unsigned int g_tma_tile_counter[TMA_COUNTERS]; // zero-initialised; each launch uses one slot (rotating)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nNI_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra NI_DONE;\nbra NI_WAIT;\nNI_DONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// L2 eviction-priority policies for the bulk copies (the encodings CUTLASS ships as TMA::CacheHintSm90)
constexpr uint64_t TMA_HINT[4] = {0ull, 0x12F0000000000000ull /* evict_first */, 0x14F0000000000000ull /* evict_last */, 0x1000000000000000ull /* evict_normal */};
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t hint)
{
    if (hint == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(hint) : "memory");
}

template <typename T, int TILE_BYTES>
__global__ void __launch_bounds__((TMA_MAX_WARPS + 1) * 32, 1) ni_step_tma_kernel(const __grid_constant__ StepArgs s, const __grid_constant__ TermTable<32> tab, int n_src, int stages, int64_t ntiles,
                                                                                      uint64_t l2_hint, int counter_slot /* < 0: static tiles */)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int TILE_ELEMS = TILE_BYTES / (int)sizeof(T);
    constexpr int VPL = TILE_BYTES / 16 / 32; // vectors per consumer lane per tile
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + TMA_MAX_STAGES;
    int *stage_tile = reinterpret_cast<int *>(empty + TMA_MAX_STAGES);
    unsigned char *tiles = smem + TMA_BAR_BYTES;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nw = (int)(blockDim.x >> 5) - 1; // consumer warps; the last warp is the producer
    const int64_t numel = s.nvec * VEC;
    const bool has_x0 = s.has_x0 != 0;
    const bool has_x = has_x0 && s.x_in != nullptr;
    const bool has_o1 = has_x0 && s.out1 != nullptr;
    const int n_pre = has_x0 ? 1 + (has_o1 ? 1 : 0) + (has_x ? 1 : 0) : 0; // sources ahead of the terms
    const bool dynamic = counter_slot >= 0;

    if (tid == 0) {
        for (int st = 0; st < stages; ++st) {
            mbar_init(&full[st], 1);
            mbar_init(&empty[st], 32); // every lane of the consuming warp arrives: each lane's reads are ordered by its own release
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == nw) {
        // ---------------- producer warp: lane i owns source i
        const unsigned char *gbase = nullptr;
        bool is_out = false;
        if (lane < n_src) {
            if (lane < n_pre) {
                if (lane == 0) { gbase = static_cast<const unsigned char *>(s.out0); is_out = true; }
                else if (has_o1 && lane == 1) { gbase = static_cast<const unsigned char *>(s.out1); is_out = true; }
                else gbase = static_cast<const unsigned char *>(s.x_in);
            } else {
                gbase = static_cast<const unsigned char *>(tab.ptr[lane - n_pre]);
            }
        }
        int st = 0;
        uint32_t ph = 0;
        int sentinels = 0;
        for (int64_t q = 0;; ++q) {
            int64_t tile;
            if (dynamic) {
                if (sentinels > 0) tile = ntiles;
                else {
                    unsigned int c = 0;
                    if (lane == 0) c = atomicInc(&g_tma_tile_counter[counter_slot], (unsigned int)(ntiles + gridDim.x - 1));
                    tile = (int64_t)__shfl_sync(0xffffffffu, c, 0);
                }
            } else {
                tile = blockIdx.x + q * (int64_t)gridDim.x;
            }
            const bool done = tile >= ntiles;
            if (done && !dynamic) break;
            mbar_wait(&empty[st], ph ^ 1u); // every lane acquires the stage itself before it issues a bulk copy into it
            __syncwarp();
            if (done) {
                // tell one consumer warp that the tiles are gone; every consumer warp needs its own notice
                // The notice travels exactly like a tile -- claimed index (-1) in stage_tile, arrive.expect_tx, a 16-byte bulk copy whose
                // completion flips the phase -- so there is ONE hand-off protocol.  (compute-sanitizer racecheck: 0 hazards,
                // profiles/r02_sanitizer_racecheck.txt; what it did find in the first version was the consumer side, see below.)
                if (lane == 0) {
                    stage_tile[st] = -1;
                    mbar_expect_tx(&full[st], 16);
                    bulk_g2s(tiles + (size_t)st * n_src * TILE_BYTES, gbase, 16, &full[st], 0);
                }
                if (++sentinels == nw) break;
            } else {
                const int64_t e0 = tile * TILE_ELEMS;
                const int64_t rem = numel - e0;
                const uint32_t bytes = (uint32_t)((rem < TILE_ELEMS ? rem : TILE_ELEMS) * (int64_t)sizeof(T));
                if (lane == 0) {
                    stage_tile[st] = (int)tile;
                    mbar_expect_tx(&full[st], bytes * (uint32_t)n_src);
                }
                __syncwarp();
                if (lane < n_src) {
                    int64_t off = e0;
                    if (is_out && s.out_strided) { // tiles never straddle samples (host guarantees per_sample % TILE_ELEMS == 0)
                        const int64_t smp = e0 / s.per_sample;
                        off = smp * s.out_sample_stride + (e0 - smp * s.per_sample);
                    }
                    bulk_g2s(tiles + ((size_t)st * n_src + lane) * TILE_BYTES, gbase + off * (int64_t)sizeof(T), bytes, &full[st], l2_hint);
                }
            }
            if (++st == stages) { st = 0; ph ^= 1u; }
        }
    } else {
        // ---------------- consumer warp `warp` owns stage-sequence numbers it = warp, warp + nw, ... of this CTA
        const int n = s.n_terms;
        for (int64_t it = warp;; it += nw) {
            if (!dynamic && blockIdx.x + it * (int64_t)gridDim.x >= ntiles) break;
            const int st = (int)(it % stages);
            const uint32_t ph = (uint32_t)((it / stages) & 1);
            mbar_wait(&full[st], ph);
            // Lane 0 alone reads the claimed index and broadcasts it: its own arrive on empty[st] below then orders that read before
            // the producer's next write to the slot by program order.  (With all 32 lanes reading it, racecheck reported a WAR hazard
            // between lane 31's read and the producer's next write -- ordered only through __syncwarp + lane 0's arrive.)
            int tile32 = 0;
            if (lane == 0) tile32 = stage_tile[st];
            const int64_t tile = __shfl_sync(0xffffffffu, tile32, 0);
            if (tile < 0) break;
            const int64_t e_tile = tile * TILE_ELEMS;
            const unsigned char *sp = tiles + (size_t)st * n_src * TILE_BYTES + lane * 16;
            auto lds = [&](int src, int v) {
                const uint4 q = *reinterpret_cast<const uint4 *>(sp + (size_t)src * TILE_BYTES + v * 512);
                Raw<T, VEC> r;
                r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
                return r;
            };
            float acc[VPL][VEC];
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[v][i] = 0.f;
            for (int t = 0; t < n; ++t) {
                const float c = tab.c[t];
                Raw<T, VEC> rr[VPL];
#pragma unroll
                for (int v = 0; v < VPL; ++v) rr[v] = lds(n_pre + t, v);
#pragma unroll
                for (int v = 0; v < VPL; ++v) fma_term<T, VEC>(acc[v], rr[v], c);
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int64_t e = e_tile + (int64_t)(v * 32 + lane) * VEC;
                if (e < numel) {
                    Raw<T, VEC> rx, ro0, ro1;
                    if (has_x0) {
                        ro0 = lds(0, v);
                        if (has_o1) ro1 = lds(1, v);
                        if (has_x) rx = lds(n_pre - 1, v);
                    }
                    int64_t sample = 0;
                    if (s.sumsq != nullptr) sample = e / s.per_sample;
                    step_epilogue<T, T, VEC>(s, e, sample, acc[v], ro0, ro1, rx, has_x0, has_x, has_o1);
                }
            }
            // hand the stage back: all 32 lanes arrive (count 32), so every lane's shared-memory reads of this stage are ordered
            // before the producer's next bulk copy into it by that lane's own release -- with __syncwarp + one elected arrive
            // (the usual pipeline idiom) racecheck reports the other 31 lanes' reads as WAR hazards against the async-proxy write
            mbar_arrive(&empty[st]);
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

// POL: load flavour, picked per launch like the step kernels' (plain ld.global when the call streams far more than the L2
// holds, L1::no_allocate when what it writes can stay resident for the next reader)
template <typename TS, typename TD, int VEC, int CAP, int POL>
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
#define NI_WSUM64_BATCH(NB)                                                                                                              \
    for (; t + NB <= n; t += NB) {                                                                                                       \
        Raw<double, VEC> rr[NB];                                                                                                         \
        _Pragma("unroll") for (int j = 0; j < NB; ++j) rr[j] = load_raw<double, VEC, POL>(static_cast<const double *>(tab.ptr[t + j]) + e); \
        _Pragma("unroll") for (int j = 0; j < NB; ++j)                                                                                   \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) acc[i] = fma(tab.c[t + j], __hiloint2double(rr[j].w[2 * i + 1], rr[j].w[2 * i]), acc[i]); \
    }
        NI_WSUM64_BATCH(8)
        NI_WSUM64_BATCH(4)
#undef NI_WSUM64_BATCH
        for (; t < n; ++t) {
            Raw<double, VEC> r1 = load_raw<double, VEC, POL>(static_cast<const double *>(tab.ptr[t]) + e);
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
            for (int j = 0; j < 8; ++j) rr[j] = load_raw<TS, VEC, POL>(static_cast<const TS *>(tab.ptr[t + j]) + e);
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
            unpack<TS, VEC>(load_raw<TS, VEC, POL>(static_cast<const TS *>(tab.ptr[t]) + e), f);
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
__global__ void __launch_bounds__(NI_BLOCK) ni_normal_kernel(T *dst, int64_t nvec, const __grid_constant__ PhiloxKeys keys, uint64_t tensor_id, uint64_t elem_offset,
                                                             const uint64_t *elem_offset_dev)
{
    const int64_t v = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (v >= nvec) return;
    float z[VEC];
    normal_vec<VEC>(effective_offset(elem_offset, elem_offset_dev) + (uint64_t)(v * VEC), tensor_id, keys, z);
    store_raw<T, VEC>(dst + v * VEC, pack<T, VEC>(z));
}

// counter += delta (the Philox element offset of a captured graph; one thread)
__global__ void ni_counter_add_kernel(uint64_t *counter, uint64_t delta) { *counter += delta; }

// the Box-Muller transform of the noise contract on caller-chosen Philox words (edge-case tests)
__global__ void __launch_bounds__(NI_BLOCK) ni_box_muller_kernel(const uint32_t *ra, const uint32_t *rb, float *za, float *zb, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * NI_BLOCK + threadIdx.x;
    if (i >= n) return;
    float a, b;
    box_muller(ra[i], rb[i], a, b);
    za[i] = a;
    zb[i] = b;
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

// Launch with programmatic dependent launch (PDL): the kernel calls griddepcontrol.launch_dependents at its top and
// griddepcontrol.wait before its first global access, so the next ni_step's CTAs are already resident and parked
// when this grid drains -- back-to-back steps (CUDA-graph replay, small tensors) lose no launch bubble.  After a
// kernel that never triggers (a torch denoiser kernel) it degrades to ordinary stream order.
// (ni::launch_pdl in ni_common.cuh.)

// Which load flavour a launch gets.  What the L2 can usefully keep between steps is what this launch WRITES (x_{k+1},
// x0_k, kept noise: the next step and the denoiser read them first).  L2-friendly loads keep those lines resident (the
// streamed history leaves first), which pays when they fit -- measured: up to ~0.6 of the L2, 67 MB written per launch
// on the SD3 shapes (+2..11%) but not C2's 100 MB (-1.2%) -- and when they are a visible share of the traffic (DDPM-250
// dense rows write 3 of ~250 tensors: streaming +1.7%).  Everything else streams.  profiles/r01_policy_sweep.txt.
bool launch_streams(const NiStepDesc *d)
{
    const int pol = g_load_policy.load(std::memory_order_relaxed);
    if (pol != 0) return pol == 2;
    const int64_t esz = dtype_size(d->dtype), osz = dtype_size(d->has_x0 ? d->out_dtype : d->dtype);
    const bool reads_x = d->has_x0 && d->x_in != nullptr && (d->a != 0.f || d->c_xin != 0.f);
    int64_t n_written = (d->has_x0 && d->x0_dst != nullptr ? 1 : 0) + (d->x_next != nullptr ? 1 : 0);
    for (int g = 0; g < d->n_gen; ++g) n_written += d->gen_dst[g] != nullptr ? 1 : 0;
    int64_t written = n_written * d->numel * esz;
    if (d->x_next_lp != nullptr) written += d->numel * 2;
    if (d->pixels_u8 != nullptr) written += d->numel;
    int64_t read = (int64_t)(d->n_terms + (d->accumulate ? 1 : 0) + (reads_x ? 1 : 0)) * d->numel * esz;
    if (d->has_x0) read += (int64_t)(1 + (d->out1 != nullptr ? 1 : 0)) * d->numel * osz;
    const int64_t l2 = dev_info().l2_bytes;
    return NI_L2_KEEP_DEN * written > NI_L2_KEEP_NUM * l2 || NI_L2_KEEP_SHARE * written < read + written;
}

template <typename T, typename TO, int VEC, bool STREAM>
int launch_step_pol(const StepArgs &a, const NiStepDesc *d, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((a.nvec + NI_BLOCK - 1) / NI_BLOCK);
    if (d->n_terms <= 32) {
        TermTable<32> tab;
        memset(&tab, 0, sizeof(tab));
        for (int i = 0; i < d->n_terms; ++i) { tab.ptr[i] = d->term_ptrs_host[i]; tab.c[i] = d->term_coeffs_host[i]; }
        launch_pdl(ni_step_kernel<T, TO, VEC, 32, STREAM>, blocks, NI_BLOCK, 0, st, opt_pdl() != 0, a, tab);
    } else {
        static thread_local TermTable<NI_MAX_TERMS> tab;
        for (int i = 0; i < d->n_terms; ++i) { tab.ptr[i] = d->term_ptrs_host[i]; tab.c[i] = d->term_coeffs_host[i]; }
        launch_pdl(ni_step_kernel<T, TO, VEC, NI_MAX_TERMS, STREAM>, blocks, NI_BLOCK, 0, st, opt_pdl() != 0, a, tab);
    }
    return check_launch("ni_step launch");
}

template <typename T, typename TO, int VEC>
int launch_step_cap(const StepArgs &a, const NiStepDesc *d, cudaStream_t st)
{
    if constexpr (VEC > 1) { // the scalar (misaligned / ragged) path is not a bandwidth path: one flavour
        if (launch_streams(d)) return launch_step_pol<T, TO, VEC, true>(a, d, st);
    }
    return launch_step_pol<T, TO, VEC, false>(a, d, st);
}

int sm_count() { return dev_info().sms; }

// TMA variant: eligible when the row fits one shared-memory stage set and tiles map 1:1 onto sources
template <typename T, int TILE_BYTES> int launch_step_tma_tile(StepArgs &a, const NiStepDesc *d, cudaStream_t st, bool *used)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int TILE_ELEMS = TILE_BYTES / (int)sizeof(T);
    *used = false;
    const int n_pre = d->has_x0 ? 1 + (d->out1 != nullptr ? 1 : 0) + (a.x_in != nullptr ? 1 : 0) : 0;
    const int n_src = n_pre + d->n_terms;
    if (n_src < 1 || n_src > TMA_MAX_SRC || d->n_terms > 32 || d->accumulate) return NI_OK;
    if (a.out_strided && d->per_sample % TILE_ELEMS != 0) return NI_OK;
    const int64_t ntiles = (d->numel + TILE_ELEMS - 1) / TILE_ELEMS;
    if (ntiles >= (1ll << 31) - 4096) return NI_OK;
    const int ctas = g_tma_ctas_per_sm.load();
    int nw = g_tma_warps.load();
    const int budget = g_tma_smem_kb.load() * 1024 / ctas - TMA_BAR_BYTES;
    int stages = budget / (n_src * TILE_BYTES);
    if (stages > g_tma_max_stages.load()) stages = g_tma_max_stages.load();
    if (stages > TMA_MAX_STAGES) stages = TMA_MAX_STAGES;
    if (stages < 2) return NI_OK;
    if (nw > stages) nw = stages;      // a consumer warp without a stage of its own would only wait
    // Stage of sequence number `it` is it % stages, its consumer is warp it % nw.  stages must be a multiple of nw so that every
    // barrier has ONE waiter walking its phases in order: mbarrier parity waits only tell the current phase from the
    // previous one, and with an unaligned ring a warp running ahead would wait on a phase two steps in the future, get
    // the stale parity of the preceding phase, read an unfilled stage and release it early (observed: launch failure).
    stages -= stages % nw;
    const size_t smem = TMA_BAR_BYTES + (size_t)stages * n_src * TILE_BYTES;
    static thread_local int attr_dev = -1; // the opt-in is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != attr_dev) {
        cudaError_t err = cudaFuncSetAttribute(ni_step_tma_kernel<T, TILE_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err != cudaSuccess) return fail(NI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(err));
        attr_dev = dev;
    }
    a.nvec = d->numel / VEC;
    int64_t grid = (int64_t)sm_count() * ctas;
    if (grid > ntiles) grid = ntiles;
    TermTable<32> tab;
    memset(&tab, 0, sizeof(tab));
    for (int i = 0; i < d->n_terms; ++i) { tab.ptr[i] = d->term_ptrs_host[i]; tab.c[i] = d->term_coeffs_host[i]; }
    static std::atomic<unsigned> next_slot{0};
    const int slot = g_tma_dynamic.load() ? (int)(next_slot.fetch_add(1) % TMA_COUNTERS) : -1;
    const uint64_t hint = TMA_HINT[g_tma_l2_hint.load() & 3];
    launch_pdl(ni_step_tma_kernel<T, TILE_BYTES>, (unsigned)grid, (unsigned)((nw + 1) * 32), smem, st,
               false /* measured: PDL slows the persistent kernel down (profiles/r01_sweep.txt) */, a, tab, n_src, stages, ntiles, hint, slot);
    *used = true;
    return check_launch("ni_step (TMA) launch");
}

template <typename T> int launch_step_tma(StepArgs &a, const NiStepDesc *d, cudaStream_t st, bool *used)
{
    if (g_tma_tile_kb.load() == 4) {
        const int rc = launch_step_tma_tile<T, 4096>(a, d, st, used);
        if (rc != NI_OK || *used) return rc;
    }
    return launch_step_tma_tile<T, 2048>(a, d, st, used);
}

template <typename T, typename TO> int launch_step(StepArgs &a, const NiStepDesc *d, bool vec_ok, cudaStream_t st)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    if constexpr (std::is_same<T, TO>::value) {
        if (vec_ok && g_variant.load() == 2) {
            bool used = false;
            const int rc = launch_step_tma<T>(a, d, st, &used);
            if (rc != NI_OK || used) return rc;
        }
    }
    if (vec_ok) {
        if (g_variant.load() == 0) { // the specialised kernels (ni_step_lean.cuh) take every launch they are built for
            bool used = false;
            const int rc = launch_step_lean<T, TO>(d, a.x_in, launch_streams(d), st, &used);
            if (rc != NI_OK || used) return rc;
        }
        a.nvec = d->numel / VEC;
        return launch_step_cap<T, TO, VEC>(a, d, st);
    }
    a.nvec = d->numel;
    return launch_step_cap<T, TO, 1>(a, d, st);
}

} // namespace
} // namespace ni

using namespace ni;

extern "C" {

int ni_version(void) { return NI_ABI_VERSION; }
const char *ni_last_error(void) { return g_err; }
int64_t ni_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int64_t ni_lean_launch_count(void) { return g_lean_launches.load(std::memory_order_relaxed); }

int ni_set_option(const char *name, int value)
{
    if (name == nullptr) return fail(NI_ERR_INVALID, "ni_set_option: NULL name");
    if (!strcmp(name, "variant")) { if (value < 0 || value > 2) return fail(NI_ERR_INVALID, "variant must be 0..2"); g_variant = value; return NI_OK; }
    if (!strcmp(name, "tma_max_stages")) { if (value < 2 || value > TMA_MAX_STAGES) return fail(NI_ERR_INVALID, "tma_max_stages must be 2..%d", TMA_MAX_STAGES); g_tma_max_stages = value; return NI_OK; }
    if (!strcmp(name, "tma_warps")) { if (value < 1 || value > TMA_MAX_WARPS) return fail(NI_ERR_INVALID, "tma_warps must be 1..%d", TMA_MAX_WARPS); g_tma_warps = value; return NI_OK; }
    if (!strcmp(name, "tma_smem_kb")) { if (value < 16 || value > 226) return fail(NI_ERR_INVALID, "tma_smem_kb must be 16..226"); g_tma_smem_kb = value; return NI_OK; }
    if (!strcmp(name, "pdl")) { g_pdl = value ? 1 : 0; return NI_OK; }
    if (!strcmp(name, "wide")) { g_wide = value ? 1 : 0; return NI_OK; }
    if (!strcmp(name, "load_policy")) { if (value < 0 || value > 2) return fail(NI_ERR_INVALID, "load_policy must be 0..2"); g_load_policy = value; return NI_OK; }
    if (!strcmp(name, "tma_tile_kb")) { if (value != 2 && value != 4) return fail(NI_ERR_INVALID, "tma_tile_kb must be 2 or 4"); g_tma_tile_kb = value; return NI_OK; }
    if (!strcmp(name, "tma_l2_hint")) { if (value < 0 || value > 3) return fail(NI_ERR_INVALID, "tma_l2_hint must be 0..3"); g_tma_l2_hint = value; return NI_OK; }
    if (!strcmp(name, "tma_dynamic")) { g_tma_dynamic = value ? 1 : 0; return NI_OK; }
    if (!strcmp(name, "tma_ctas_per_sm")) { if (value < 1 || value > 4) return fail(NI_ERR_INVALID, "tma_ctas_per_sm must be 1..4"); g_tma_ctas_per_sm = value; return NI_OK; }
    return fail(NI_ERR_INVALID, "ni_set_option: unknown option '%s'", name);
}

int ni_step_flavour(const NiStepDesc *d)
{
    if (d == nullptr) return fail(NI_ERR_INVALID, "ni_step_flavour: NULL descriptor");
    if (d->numel < 0 || dtype_size(d->dtype) == 0 || (d->has_x0 && dtype_size(d->out_dtype) == 0)) return fail(NI_ERR_INVALID, "ni_step_flavour: bad sizes or dtypes");
    if (d->n_terms < 0 || d->n_gen < 0 || d->n_gen > NI_MAX_GEN) return fail(NI_ERR_TOO_MANY, "ni_step_flavour: bad term counts");
    return launch_streams(d) ? 1 : 0;
}

int ni_step(const NiStepDesc *d, void *stream)
{
    if (d == nullptr) return fail(NI_ERR_INVALID, "ni_step: NULL descriptor");
    if (d->numel < 0 || d->per_sample <= 0) return fail(NI_ERR_INVALID, "ni_step: bad sizes numel=%lld per_sample=%lld", (long long)d->numel, (long long)d->per_sample);
    if (d->numel % d->per_sample != 0) return fail(NI_ERR_INVALID, "ni_step: numel %lld is not a multiple of per_sample %lld", (long long)d->numel, (long long)d->per_sample);
    if (d->n_terms < 0 || d->n_terms > NI_MAX_TERMS) return fail(NI_ERR_TOO_MANY, "ni_step: n_terms=%d exceeds NI_MAX_TERMS=%d (chain launches with accumulate=1)", d->n_terms, NI_MAX_TERMS);
    if (d->n_gen < 0 || d->n_gen > NI_MAX_GEN) return fail(NI_ERR_TOO_MANY, "ni_step: n_gen=%d exceeds NI_MAX_GEN=%d", d->n_gen, NI_MAX_GEN);
    if ((reinterpret_cast<uintptr_t>(d->elem_offset_dev) & 7u) != 0) return fail(NI_ERR_INVALID, "ni_step: elem_offset_dev must be 8-byte aligned");
    if (d->numel == 0) return NI_OK; // an empty shard: nothing to do (its tensors have NULL data pointers)
    if (d->x_next == nullptr && d->pixels_u8 == nullptr) return fail(NI_ERR_INVALID, "ni_step: x_next is NULL (allowed only with pixels_u8)");
    if (d->pixels_u8 != nullptr && (d->px_channels <= 0 || d->per_sample % d->px_channels != 0)) return fail(NI_ERR_INVALID, "ni_step: pixels_u8 needs px_channels dividing per_sample");
    if (d->x_next == nullptr && (d->accumulate || d->x_next_lp != nullptr || d->sumsq != nullptr)) return fail(NI_ERR_INVALID, "ni_step: accumulate / x_next_lp / sumsq need x_next");
    if (d->n_terms > 0 && (d->term_ptrs_host == nullptr || d->term_coeffs_host == nullptr)) return fail(NI_ERR_INVALID, "ni_step: term tables are NULL");
    if (d->has_x0) {
        if (d->out0 == nullptr) return fail(NI_ERR_INVALID, "ni_step: has_x0 but out0 is NULL");
        if (d->x_in == nullptr && (d->a != 0.f || d->c_xin != 0.f)) return fail(NI_ERR_INVALID, "ni_step: a or c_xin != 0 but x_in is NULL");
        if (d->out_sample_stride < d->per_sample) return fail(NI_ERR_INVALID, "ni_step: out_sample_stride < per_sample");
        if (d->x_next != nullptr && (d->x_next == d->x_in || d->x_next == d->out0 || d->x_next == d->out1 || d->x0_dst == d->x_next))
            return fail(NI_ERR_INVALID, "ni_step: x_next aliases an input or x0_dst");
    }
    for (int i = 0; i < d->n_terms; ++i) {
        if (d->term_ptrs_host[i] == nullptr) return fail(NI_ERR_INVALID, "ni_step: term %d pointer is NULL", i);
        if (d->term_ptrs_host[i] == d->x_next || d->term_ptrs_host[i] == d->x0_dst) return fail(NI_ERR_INVALID, "ni_step: x_next / x0_dst aliases term %d", i);
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
    a.x_in = (d->has_x0 && (d->a != 0.f || d->c_xin != 0.f)) ? d->x_in : nullptr;
    a.out0 = d->out0; a.out1 = d->out1;
    a.x0_dst = d->x0_dst; a.x_next = d->x_next; a.x_next_lp = d->x_next_lp; a.sumsq = d->sumsq;
    a.a = d->a; a.b0 = d->b0; a.b1 = d->b1; a.c_x0 = d->c_x0; a.c_xin = d->has_x0 ? d->c_xin : 0.f;
    a.bias = d->bias; a.pixels = d->pixels_u8; a.px_scale = d->px_scale; a.px_shift = d->px_shift; a.px_channels = d->px_channels;
    a.keys = philox_keys(d->philox_seed);
    a.elem_offset = d->elem_offset; a.elem_offset_dev = d->elem_offset_dev;
    a.n_terms = d->n_terms; a.n_gen = d->n_gen;
    a.has_x0 = d->has_x0; a.accumulate = d->accumulate; a.lp_dtype = d->lp_dtype;
    for (int g = 0; g < d->n_gen; ++g) { a.gen_tid[g] = d->gen_tensor_ids[g]; a.gen_c[g] = d->gen_coeffs[g]; a.gen_dst[g] = d->gen_dst[g]; }

    // 128-bit path needs every pointer 16 B aligned (8 B for half outputs next to fp32 state) and vector-sized shapes
    const int VEC = 16 / ds;
    bool vec_ok = d->numel % VEC == 0 && d->per_sample % VEC == 0 && (d->x_next == nullptr || aligned16(d->x_next));
    if (d->pixels_u8 != nullptr) vec_ok = vec_ok && (d->per_sample / d->px_channels) % VEC == 0; // a vector must stay inside one channel plane
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
template <typename TS, typename TD, int POL>
int launch_wsum_pol(const void *const *src, const double *coeffs, int n, void *dst, int64_t numel, double scale, bool vec_ok, cudaStream_t st)
{
    static thread_local WsumTable<NI_MAX_TERMS> big;
    constexpr int VEC = 16 / (int)sizeof(TS);
    if (n <= 32) {
        WsumTable<32> tab;
        memset(&tab, 0, sizeof(tab));
        for (int i = 0; i < n; ++i) { tab.ptr[i] = src[i]; tab.c[i] = coeffs[i]; }
        if (vec_ok) {
            const int64_t nvec = numel / VEC;
            ni_wsum_kernel<TS, TD, VEC, 32, POL><<<(unsigned)((nvec + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(tab, n, dst, nvec, scale);
        } else {
            ni_wsum_kernel<TS, TD, 1, 32, NI_LOAD_POLICY><<<(unsigned)((numel + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(tab, n, dst, numel, scale);
        }
    } else {
        for (int i = 0; i < n; ++i) { big.ptr[i] = src[i]; big.c[i] = coeffs[i]; }
        if (vec_ok) {
            const int64_t nvec = numel / VEC;
            ni_wsum_kernel<TS, TD, VEC, NI_MAX_TERMS, POL><<<(unsigned)((nvec + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(big, n, dst, nvec, scale);
        } else {
            ni_wsum_kernel<TS, TD, 1, NI_MAX_TERMS, NI_LOAD_POLICY><<<(unsigned)((numel + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, st>>>(big, n, dst, numel, scale);
        }
    }
    return check_launch("ni_weighted_sum launch");
}

template <typename TS, typename TD>
int launch_wsum(const void *const *src, const double *coeffs, int n, void *dst, int64_t numel, double scale, bool vec_ok, cudaStream_t st)
{
    // same rule as launch_streams(): keep the L2-friendly loads only when the result fits in 0.6 of the L2 and is a visible share of the traffic
    const int pol = g_load_policy.load(std::memory_order_relaxed);
    const int64_t written = numel * (int64_t)sizeof(TD), read = (int64_t)n * numel * (int64_t)sizeof(TS);
    const bool streams = pol != 0 ? pol == 2 : (NI_L2_KEEP_DEN * written > NI_L2_KEEP_NUM * dev_info().l2_bytes || NI_L2_KEEP_SHARE * written < read + written);
    if (streams) return launch_wsum_pol<TS, TD, NI_STREAM_LOAD_POLICY>(src, coeffs, n, dst, numel, scale, vec_ok, st);
    return launch_wsum_pol<TS, TD, NI_LOAD_POLICY>(src, coeffs, n, dst, numel, scale, vec_ok, st);
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

static int philox_normal_impl(void *dst, int64_t numel, int dst_dtype, uint64_t seed, uint64_t tensor_id, uint64_t elem_offset, const uint64_t *elem_offset_dev, void *stream)
{
    if ((reinterpret_cast<uintptr_t>(elem_offset_dev) & 7u) != 0) return fail(NI_ERR_INVALID, "ni_philox_normal_at: elem_offset_dev must be 8-byte aligned");
    if (numel == 0) return NI_OK;
    if (dst == nullptr || numel < 0) return fail(NI_ERR_INVALID, "ni_philox_normal: bad arguments");
    const int ds = dtype_size(dst_dtype);
    if (!(dst_dtype == NI_F32 || dst_dtype == NI_F16 || dst_dtype == NI_BF16)) return fail(NI_ERR_DTYPE, "ni_philox_normal: dtype %d not supported", dst_dtype);
    const int VEC = 16 / ds;
    const bool vec_ok = numel % VEC == 0 && aligned16(dst);
    const PhiloxKeys keys = philox_keys(seed);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nvec = vec_ok ? numel / VEC : numel;
    const unsigned blocks = (unsigned)((nvec + NI_BLOCK - 1) / NI_BLOCK);
    if (dst_dtype == NI_F32) {
        if (vec_ok) ni_normal_kernel<float, 4><<<blocks, NI_BLOCK, 0, st>>>((float *)dst, nvec, keys, tensor_id, elem_offset, elem_offset_dev);
        else ni_normal_kernel<float, 1><<<blocks, NI_BLOCK, 0, st>>>((float *)dst, nvec, keys, tensor_id, elem_offset, elem_offset_dev);
    } else if (dst_dtype == NI_F16) {
        if (vec_ok) ni_normal_kernel<__half, 8><<<blocks, NI_BLOCK, 0, st>>>((__half *)dst, nvec, keys, tensor_id, elem_offset, elem_offset_dev);
        else ni_normal_kernel<__half, 1><<<blocks, NI_BLOCK, 0, st>>>((__half *)dst, nvec, keys, tensor_id, elem_offset, elem_offset_dev);
    } else {
        if (vec_ok) ni_normal_kernel<__nv_bfloat16, 8><<<blocks, NI_BLOCK, 0, st>>>((__nv_bfloat16 *)dst, nvec, keys, tensor_id, elem_offset, elem_offset_dev);
        else ni_normal_kernel<__nv_bfloat16, 1><<<blocks, NI_BLOCK, 0, st>>>((__nv_bfloat16 *)dst, nvec, keys, tensor_id, elem_offset, elem_offset_dev);
    }
    return check_launch("ni_philox_normal launch");
}

int ni_philox_normal(void *dst, int64_t numel, int dst_dtype, uint64_t seed, uint64_t tensor_id, uint64_t elem_offset, void *stream)
{
    return philox_normal_impl(dst, numel, dst_dtype, seed, tensor_id, elem_offset, nullptr, stream);
}

int ni_philox_normal_at(void *dst, int64_t numel, int dst_dtype, uint64_t seed, uint64_t tensor_id, uint64_t elem_offset, const uint64_t *elem_offset_dev, void *stream)
{
    return philox_normal_impl(dst, numel, dst_dtype, seed, tensor_id, elem_offset, elem_offset_dev, stream);
}

int ni_counter_add(uint64_t *counter_dev, uint64_t delta, void *stream)
{
    if (counter_dev == nullptr || (reinterpret_cast<uintptr_t>(counter_dev) & 7u) != 0) return fail(NI_ERR_INVALID, "ni_counter_add: counter must be a non-NULL, 8-byte aligned device pointer");
    ni_counter_add_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(counter_dev, delta);
    return check_launch("ni_counter_add launch");
}

int ni_debug_box_muller(const uint32_t *ra, const uint32_t *rb, float *za, float *zb, int64_t n, void *stream)
{
    if (n == 0) return NI_OK;
    if (ra == nullptr || rb == nullptr || za == nullptr || zb == nullptr || n < 0) return fail(NI_ERR_INVALID, "ni_debug_box_muller: bad arguments");
    ni_box_muller_kernel<<<(unsigned)((n + NI_BLOCK - 1) / NI_BLOCK), NI_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(ra, rb, za, zb, n);
    return check_launch("ni_debug_box_muller launch");
}

int ni_to_pixel_u8(const void *x, int src_dtype, uint8_t *dst, int64_t batch, int channels, int height, int width, float scale, float shift, void *stream)
{
    if (batch < 0 || channels <= 0 || height <= 0 || width <= 0) return fail(NI_ERR_INVALID, "ni_to_pixel_u8: bad sizes");
    if (batch == 0) return NI_OK; // an empty shard: its tensors have NULL data pointers
    if (x == nullptr || dst == nullptr) return fail(NI_ERR_INVALID, "ni_to_pixel_u8: NULL pointer");
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
