/* ni_b200.h -- C ABI of libni_b200.so: the B200 (sm_100a) Natural Inference sampling step.
 *
 * This is the drop-in boundary.  The reference (blairstar/NaturalDiffusion) has no FFI of
 * its own: its "sampler API" is a set of module-level Python functions that the three
 * sampling scripts resolve at call time.  Each entry point below names the reference
 * function / loop lines it replaces; the Python host layer (naturaldiffusion_b200/) binds
 * them with ctypes and re-exposes the reference's exact signatures.
 *
 * Conventions
 *   - plain C types only; every data pointer is a DEVICE pointer unless the name ends in
 *     `_host`; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all calls are asynchronous on `stream`, never allocate, never synchronise, never throw;
 *     they are CUDA-graph capturable (tables travel in kernel parameters).
 *   - return 0 on success, a negative NI_ERR_* otherwise; ni_last_error() gives a
 *     thread-local message.  There is NO CPU fallback: without a CUDA device every compute
 *     entry point fails with NI_ERR_CUDA.
 *   - 16-byte aligned pointers + numel/per_sample multiples of the vector width take the
 *     128-bit path; anything else silently takes the (slower, same results) scalar path.
 *
 * Noise contract (shared with oracle/philox_oracle.c, bit-exact in the integer part):
 *   element e = elem_offset + i of noise tensor `tensor_id` under `seed`:
 *     group g = e >> 2, lane = e & 3
 *     (r0,r1,r2,r3) = Philox4x32-10(counter = (g.lo, g.hi, tensor_id.lo, tensor_id.hi),
 *                                   key = (seed.lo, seed.hi))
 *     u(r) = fma((float)r, 2^-32, 2^-33);  v(r) = fma((float)r, 2^-31, 2^-32)
 *     rad = sqrt(-2 ln u(r0)); z0 = rad*cospi(v(r1)); z1 = rad*sinpi(v(r1)); (r2,r3) -> z2,z3
 *     (device: -2 ln u through MUFU.LG2 with a 4-term series for u > 31/32, SFU sqrt/sin/cos -- within 6e-6 of the
 *     fp64 evaluation of these formulas, 3e-7 typical; ni_debug_box_muller exposes the transform for edge-case tests)
 *   Keyed by the GLOBAL element index, so a batch sharded over G GPUs (each shard passing
 *   its own elem_offset) draws exactly the tensor a single GPU would.
 */
#ifndef NI_B200_H
#define NI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NI_ABI_VERSION 4
#define NI_MAX_TERMS 512 /* stored history/noise terms per launch; longer rows: call twice with accumulate */
#define NI_MAX_GEN 4     /* noise terms generated in-kernel per launch */

typedef enum { NI_F32 = 0, NI_F16 = 1, NI_BF16 = 2, NI_F64 = 3 } ni_dtype;

enum {
    NI_OK = 0,
    NI_ERR_INVALID = -1,   /* NULL pointer, negative size, aliasing x_next with an input ... */
    NI_ERR_DTYPE = -2,     /* dtype combination not built */
    NI_ERR_TOO_MANY = -3,  /* n_terms > NI_MAX_TERMS or n_gen > NI_MAX_GEN */
    NI_ERR_CUDA = -4       /* a CUDA runtime call failed (message has the cudaError string) */
};

/* One fused NI step:
 *     x0      = a * x_in + b0 * out0 + b1 * out1                 (model I/O scaling + CFG)
 *     x_next  = c_x0 * x0 + sum_t term_coeffs[t] * term[t]        (row k of A against the x0 ring,
 *             + sum_g gen_coeffs[g] * N(seed, gen_tensor_ids[g])   row k of B against stored / fresh noise)
 *             + c_xin * x_in                                       (first-order rows only, see c_xin)
 * in ONE pass over HBM.  Replaces, per step,
 *   src/CIFAR10NaturalInference.py:298-304 (data_fn :219-230 + weighted_sum :233-238 + noise mix),
 *   src/ValidateNaturalInference.py:352-366 (CFG fuse :193, x0 :355, randn_like :359, 2x weighted_sum :198-204),
 *   src/SD3NaturalInference.py:207-221 (weighted_sum :157-168 x2, input mix :209, x0 + CFG :215-217).
 * Accumulation is fp32 FMA in the order written above; x0 is rounded to `dtype` before it
 * enters the sum, so later steps that re-read it from the ring see the same value. */
typedef struct NiStepDesc {
    int64_t numel;             /* elements in this shard (batch_local * per_sample) */
    int64_t per_sample;        /* C*H*W */
    int32_t dtype;             /* ni_dtype of x_in, x0_dst, terms, gen_dst, x_next: F32 | F16 | BF16 */
    int32_t out_dtype;         /* ni_dtype of out0/out1: same as dtype, or F16/BF16 with dtype F32 */

    int32_t has_x0;            /* 0: skip the x0 stage (pure weighted sum + noise) */
    const void *x_in;          /* current model input x_k; may be NULL iff a == 0 */
    const void *out0;          /* model output (cond / text / score-net h) */
    const void *out1;          /* second model output for CFG, or NULL */
    int64_t out_sample_stride; /* elements between samples in out0/out1; per_sample if dense.
                                  (DiT returns [B,8,32,32] and the path reads channels [:4]) */
    float a, b0, b1;
    void *x0_dst;              /* ring slot receiving x0_k, or NULL if no later row reads it */
    float c_x0;                /* A[k,k] */
    float c_xin;               /* usually 0.  Rows of first-order samplers (DDPM, DDIM, Euler, flow Euler) satisfy
                                  row_k[:k] = c_k * row_{k-1}[:k], i.e. sum_{j<k} A[k,j] x0_j + sum_{j<=k} B[k,j] eps_j
                                  = c_k * x_k: pass c_k here and no history terms -- O(1) reads per step */

    int32_t n_terms;           /* stored terms: earlier x0 slots and stored noise tensors */
    const void *const *term_ptrs_host; /* HOST array[n_terms] of device pointers */
    const float *term_coeffs_host;     /* HOST array[n_terms] */

    int32_t n_gen;             /* noise terms generated in-kernel (fresh eps_{k+1}; optionally eps_0) */
    uint64_t gen_tensor_ids[NI_MAX_GEN];
    float gen_coeffs[NI_MAX_GEN];
    void *gen_dst[NI_MAX_GEN]; /* where to keep the generated tensor for later rows, or NULL */
    uint64_t philox_seed;
    uint64_t elem_offset;      /* global index of this shard's element 0 */
    const uint64_t *elem_offset_dev; /* optional DEVICE counter added to elem_offset when the kernel runs: a captured CUDA graph
                                  draws new noise on every replay (advance it with ni_counter_add inside the graph), which is how
                                  successive batches get fresh randn like src/ValidateNaturalInference.py:345,359 per call */

    int32_t accumulate;        /* 1: x_next += (this launch) -- used to chain rows longer than NI_MAX_TERMS */
    float bias;                /* constant added to x_next (output stage of latent models: x/scaling_factor + shift_factor,
                                  src/SD3NaturalInference.py:238; the 1/scaling_factor goes into the row coefficients) */
    void *x_next;              /* x_{k+1}; must not alias any input.  May be NULL when pixels_u8 is given */
    void *x_next_lp;           /* optional copy of x_next in lp_dtype for a reduced-precision denoiser */
    int32_t lp_dtype;
    float *sumsq;              /* optional [batch_local] fp32: += sum over the sample of x_next^2 (caller zeroes) */
    uint8_t *pixels_u8;        /* optional, last step: NHWC uint8 = trunc(clip((x_next*px_scale + px_shift)*255, 0, 255)), the
                                  output stage of src/CIFAR10NaturalInference.py:308-309,212-216 fused into the step */
    float px_scale, px_shift;
    int32_t px_channels;       /* C of the NCHW sample (per_sample = C*H*W) */
} NiStepDesc;

int ni_version(void);
const char *ni_last_error(void);
/* kernels launched by this library in this process so far (bench.py's `gpu_launches`) */
int64_t ni_launch_count(void);
/* how many of those were ni_step launches served by the row-shape-specialised kernels (ni_step_lean.cuh) */
int64_t ni_lean_launch_count(void);

/* Tuning knobs, process-wide: "variant" 0 auto (row-shape-specialised kernels, generic kernel for what they do not cover)
 * | 1 generic direct-load kernel | 2 TMA-staged kernel when eligible;
 * "tma_max_stages" 2..32; "tma_warps" 1..16; "tma_smem_kb" 16..226; "tma_ctas_per_sm" 1..4; "pdl" 0|1 (programmatic
 * dependent launch of the direct-load step kernel, default 1); "wide" 0|1 (fp32 state: the 256-bit LDG.E.ENL2.256 / STG.E.ENL2.256
 * instantiations of the specialised step kernels whenever every pointer is 32-byte aligned, default 1); "tma_tile_kb" 2|4,
 * "tma_l2_hint" 0..3, "tma_dynamic" 0|1 (TMA variant); "load_policy" 0 auto | 1 L2-friendly loads
 * (ld.global.L1::no_allocate) | 2 streaming loads (plain ld.global) -- auto picks L2-friendly loads when what the
 * launch writes fits in 0.6 of the L2 and is >= 1/16 of its traffic, streaming otherwise (measured:
 * profiles/r01_policy_sweep.txt).  Results do not depend on them. */
int ni_set_option(const char *name, int value);

int ni_step(const NiStepDesc *desc_host, void *stream);

/* Which load flavour ni_step would pick for this descriptor on the current device: 1 = streaming (plain ld.global),
 * 0 = L2-friendly (ld.global.L1::no_allocate), negative = NI_ERR_*.  Launches nothing; reporting / tests only. */
int ni_step_flavour(const NiStepDesc *desc_host);

/* dst = scale * sum_t coeffs[t] * src[t].  The three reference `weighted_sum`s and
 * `euler_weighted_sum` (src/CIFAR10NaturalInference.py:233-238, src/ValidateNaturalInference.py:198-204,
 * src/SD3NaturalInference.py:157-168, :61-69) as one launch.  src_dtype F64 (the CIFAR loop keeps an
 * fp64 history) accumulates in fp64; everything else in fp32.  coeffs are fp64 host values. */
int ni_weighted_sum(const void *const *src_ptrs_host, const double *coeffs_host, int n_terms,
                    void *dst, int64_t numel, int src_dtype, int dst_dtype, double scale, void *stream);

/* dst[i] = N(0,1) of the noise contract above; dst_dtype F32 | F16 | BF16.  Stands in for
 * torch.randn / randn_like at src/CIFAR10NaturalInference.py:290, src/ValidateNaturalInference.py:345,359,
 * src/SD3NaturalInference.py:182 and lets tests feed the reference the very tensors the fused step draws. */
int ni_philox_normal(void *dst, int64_t numel, int dst_dtype, uint64_t seed, uint64_t tensor_id,
                     uint64_t elem_offset, void *stream);

/* Same, with a DEVICE counter added to elem_offset at run time (see NiStepDesc.elem_offset_dev). */
int ni_philox_normal_at(void *dst, int64_t numel, int dst_dtype, uint64_t seed, uint64_t tensor_id,
                        uint64_t elem_offset, const uint64_t *elem_offset_dev, void *stream);

/* *counter_dev += delta on `stream` (one tiny launch; graph-capturable).  The loop-level sampler ends each captured
 * trajectory with it so that replay i draws the noise of samples [first + i*B, first + (i+1)*B). */
int ni_counter_add(uint64_t *counter_dev, uint64_t delta, void *stream);

/* Test hook: the Box-Muller transform of the noise contract on caller-chosen Philox words, (za, zb)[i] from
 * (ra, rb)[i] -- lets tests hit u -> 1, the series/LG2 switch-over and the 6.7-sigma tail directly. */
int ni_debug_box_muller(const uint32_t *ra, const uint32_t *rb, float *za, float *zb, int64_t n, void *stream);

/* Output stage (src/CIFAR10NaturalInference.py:308-309 with :212-216): NCHW float -> NHWC uint8,
 * u8 = trunc(clip((x*scale + shift)*255, 0, 255)); scale = shift = 0.5 is the inverse scaler of
 * centered data (deps/score_sde_pytorch/datasets.py:32-38). */
int ni_to_pixel_u8(const void *x, int src_dtype, uint8_t *dst_nhwc, int64_t batch, int channels,
                   int height, int width, float scale, float shift, void *stream);

/* FID sufficient statistics (the evaluation step that follows sampling).  The reference computes
 * mu = np.mean(act, 0), sigma = np.cov(act, rowvar=False) over all pool3 activations on the host
 * (src/CIFAR10NaturalInference.py:73-86).  Here each rank accumulates  stats = [n | sum x (d) | sum x x^T (d*d, row-major)]
 * in fp64 on its GPU -- feats is [m, d] fp32 with row pitch ld (elements), widened in registers; the rank-k update runs on the
 * fp64 tensor cores (DMMA), upper triangle computed once and mirrored, deterministic -- and one all-reduce of `stats` merges
 * the ranks.  stats += (this batch); the caller zeroes it once. */
int ni_fid_accumulate(const float *feats, int64_t m, int d, int64_t ld, double *stats, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NI_B200_H */
