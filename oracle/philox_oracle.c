/* CPU oracle: counter-based Gaussian noise of the NI step (Philox4x32-10 + Box-Muller)
 * and a plain-C fp64 restatement of the weighted sum.
 *
 * TEST INFRASTRUCTURE ONLY -- linked by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker; never by the product path (naturaldiffusion_b200/).
 *
 * Why it exists: the reference draws noise with torch.randn
 * (src/CIFAR10NaturalInference.py:290, src/ValidateNaturalInference.py:345,359,
 * src/SD3NaturalInference.py:182) whose stream is device dependent; the north-star asks
 * the fused kernel to generate noise in-kernel with a counter-based Philox.  This file
 * restates that generator on the CPU so tests can (i) check the kernel's bits against the
 * published Random123 known-answer vectors and (ii) regenerate the very tensors the kernel
 * draws and feed them to the reference's loops.
 *
 * Algorithm (D. E. Shaw Research Random123 v1.09, philox.h, philox4x32-10 -- third party,
 * not in /root/reference; published algorithm restated):
 *   round:  (c0,c1,c2,c3) <- (hi(M1*c2)^c1^k0, lo(M1*c2), hi(M0*c0)^c3^k1, lo(M0*c0))
 *   key schedule: k0 += 0x9E3779B9, k1 += 0xBB67AE85 between rounds; 10 rounds.
 * Noise layout (ours; include/ni_b200.h "noise contract"):
 *   element e (GLOBAL index = elem_offset + i) belongs to group g = e>>2, lane e&3
 *   counter = (g lo32, g hi32, tensor_id lo32, tensor_id hi32), key = (seed lo32, seed hi32)
 *   u(r) = (float)r * 2^-32 + 2^-33  (one fp32 fma);  v(r) = (float)r * 2^-31 + 2^-32
 *   rad = sqrt(-2 ln u(r0)),  z0 = rad*cos(pi*v(r1)), z1 = rad*sin(pi*v(r1)); same for (r2,r3)->z2,z3
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void ni_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void box_muller(uint32_t ra, uint32_t rb, float *za, float *zb)
{
    const double PI = 3.14159265358979323846;
    float u = fmaf((float)ra, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   /* 2^-32, 2^-33 */
    float v = fmaf((float)rb, 4.6566128730773926e-10f, 2.3283064365386963e-10f);   /* 2^-31, 2^-32 */
    double rad = sqrt(-2.0 * log((double)u));
    *za = (float)(rad * cos(PI * (double)v));
    *zb = (float)(rad * sin(PI * (double)v));
}

/* the transform alone, on caller-chosen Philox words (edge cases: u -> 1, u -> 2^-33, the tails) */
void ni_oracle_box_muller(const uint32_t *ra, const uint32_t *rb, float *za, float *zb, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) box_muller(ra[i], rb[i], &za[i], &zb[i]);
}

/* dst[i] = N(0,1) sample of global element elem_offset+i of noise tensor `tensor_id`. */
void ni_oracle_philox_normal_f32(float *dst, int64_t numel, uint64_t seed, uint64_t tensor_id, uint64_t elem_offset)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint64_t g_prev = ~(uint64_t)0;
    float z[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < numel; ++i) {
        uint64_t e = elem_offset + (uint64_t)i;
        uint64_t g = e >> 2;
        if (g != g_prev) {
            uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)tensor_id, (uint32_t)(tensor_id >> 32)};
            uint32_t r[4];
            ni_oracle_philox4x32_10(ctr, key, r);
            box_muller(r[0], r[1], &z[0], &z[1]);
            box_muller(r[2], r[3], &z[2], &z[3]);
            g_prev = g;
        }
        dst[i] = z[e & 3];
    }
}

/* Plain-C fp64 weighted sum: dst[i] = sum_t coeff[t]*src[t][i] (the common form of
 * src/CIFAR10NaturalInference.py:233-238 / src/ValidateNaturalInference.py:198-204). */
void ni_oracle_weighted_sum_f32(const float *const *src, const double *coeff, int n_terms, float *dst, int64_t numel)
{
    for (int64_t i = 0; i < numel; ++i) {
        double acc = 0.0;
        for (int t = 0; t < n_terms; ++t) acc += coeff[t] * (double)src[t][i];
        dst[i] = (float)acc;
    }
}
