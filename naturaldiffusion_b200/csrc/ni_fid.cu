// ni_fid.cu -- FID sufficient statistics on the GPU: n, sum x, sum x x^T in fp64.
//
// The reference gathers all 50 000 Inception pool3 activations on the host and calls np.mean / np.cov
// (src/CIFAR10NaturalInference.py:73-86).  With sampling sharded by batch each rank keeps (n, sum x, sum x x^T) of its own
// activations on its own GPU and ONE all-reduce merges them (naturaldiffusion_b200/fid.py).  The 2048-wide rank-k update
// S += X^T X is the only GEMM-shaped work on the whole path, and it must be fp64 (covariances of 50 000 samples lose the
// small eigenvalues in fp32) -- so it runs on the fp64 tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA; tcgen05 has no fp64
// kind).  Hand-written rather than a cuBLAS dsyrk so that
//   * X is read as the fp32 the feature network produced and widened to fp64 in registers: no fp64 copy of the
//     activations is ever materialised (the torch path wrote and re-read m x d x 8 bytes per update);
//   * only the upper-triangular 64x64 tiles are computed (528 of 1024 for d = 2048) and each CTA owns its tile and the
//     mirrored one, so the accumulation into the persistent statistics buffer is a plain deterministic += (no atomics);
//   * column sums and the sample count ride in the same call.
// CTA = 128 threads = 4 warps, each warp a 32x32 block of the 64x64 tile (4x4 DMMA tiles), samples streamed 32 at a time
// through a cp.async double buffer (zero-filled past the last sample / column).
#include "ni_common.cuh"

namespace ni {
namespace {

constexpr int FID_TILE = 64;   // S tile edge
constexpr int FID_KC = 32;     // samples per stage
constexpr int FID_LD = 72;     // smem row pitch in floats: 72 mod 32 = 8 -> the 4 x 8 fragment reads hit 32 distinct banks

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// one stage: rows [s0, s0+32) x columns [c0, c0+64) of X -> smem[32][72]
__device__ __forceinline__ void load_stage(float *dst, const float *x, int64_t ld, int64_t m, int d, int64_t s0, int c0)
{
    // 32 rows x 16 float4 = 512 chunks, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = threadIdx.x + i * 128;
        const int r = q >> 4, c4 = (q & 15) * 4;
        const int64_t row = s0 + r;
        const int col = c0 + c4;
        int bytes = 0;
        if (row < m && col < d) bytes = (d - col >= 4 ? 4 : d - col) * 4;
        const float *src = bytes ? x + row * ld + col : x; // keep the address valid when nothing is read
        cp_async16(dst + r * FID_LD + c4, src, bytes);
    }
}

// grid.x = upper-triangular tile index; stats = [n | sum x (d) | S (d*d, row-major)]
__global__ void __launch_bounds__(128) ni_fid_syrk_kernel(const float *__restrict__ x, int64_t ld, int64_t m, int d, double *__restrict__ S, int ntile)
{
    __shared__ __align__(16) float sa[2][FID_KC * FID_LD];
    __shared__ __align__(16) float sb[2][FID_KC * FID_LD];
    // tile (bi, bj), bi <= bj, from the linear index
    int t = blockIdx.x, bi = 0;
    while (t >= ntile - bi) { t -= ntile - bi; ++bi; }
    const int bj = bi + t;
    const int i0 = bi * FID_TILE, j0 = bj * FID_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32; // this warp's 32x32 block inside the tile
    const int fr = lane >> 2, fk = lane & 3;               // fragment row/col (0..7) and k (0..3)

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int64_t nstage = (m + FID_KC - 1) / FID_KC;
    load_stage(sa[0], x, ld, m, d, 0, i0);
    load_stage(sb[0], x, ld, m, d, 0, j0);
    cp_async_commit();
    for (int64_t st = 0; st < nstage; ++st) {
        const int cur = (int)(st & 1);
        if (st + 1 < nstage) {
            load_stage(sa[cur ^ 1], x, ld, m, d, (st + 1) * FID_KC, i0);
            load_stage(sb[cur ^ 1], x, ld, m, d, (st + 1) * FID_KC, j0);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float *pa = sa[cur] + wi + fr, *pb = sb[cur] + wj + fr;
#pragma unroll
        for (int k0 = 0; k0 < FID_KC; k0 += 4) {
            double fa[4], fb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) fa[a] = (double)pa[(k0 + fk) * FID_LD + a * 8]; // A[i][k] = X[s0+k][i0+i]
#pragma unroll
            for (int b = 0; b < 4; ++b) fb[b] = (double)pb[(k0 + fk) * FID_LD + b * 8]; // B[k][j] = X[s0+k][j0+j]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b], fa[a], fb[b]);
        }
        __syncthreads();
    }

    // S[tile] += acc, and the mirrored tile for off-diagonal blocks; C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = i0 + wi + a * 8 + fr, j = j0 + wj + b * 8 + fk * 2 + e;
                if (i < d && j < d) {
                    S[(int64_t)i * d + j] += acc[a][b][e];
                    if (bi != bj) S[(int64_t)j * d + i] += acc[a][b][e];
                }
            }
}

// stats[0] += m, stats[1 + c] += sum_s X[s][c]: one thread per column over a slice of the rows, fp64 atomics across
// the row slices (order-dependent only in the last fp64 bit; the syrk above is fully deterministic)
__global__ void __launch_bounds__(256) ni_fid_colsum_kernel(const float *__restrict__ x, int64_t ld, int64_t m, int d, double *__restrict__ stats, int rows_per_block)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < m ? r0 + rows_per_block : m;
    if (c < d) {
        double s = 0.0;
        for (int64_t r = r0; r < r1; ++r) s += (double)x[r * ld + c];
        atomicAdd(stats + 1 + c, s);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) stats[0] += (double)m;
}

} // namespace
} // namespace ni

using namespace ni;

extern "C" int ni_fid_accumulate(const float *feats, int64_t m, int d, int64_t ld, double *stats, void *stream)
{
    if (m == 0) return NI_OK;
    if (feats == nullptr || stats == nullptr || m < 0 || d <= 0 || ld < d) return fail(NI_ERR_INVALID, "ni_fid_accumulate: bad arguments");
    if (d % 4 != 0 || ld % 4 != 0 || !aligned16(feats)) return fail(NI_ERR_INVALID, "ni_fid_accumulate: d and ld must be multiples of 4 and feats 16-byte aligned");
    if ((reinterpret_cast<uintptr_t>(stats) & 7u) != 0) return fail(NI_ERR_INVALID, "ni_fid_accumulate: stats must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ntile = (d + FID_TILE - 1) / FID_TILE;
    const int rows_per_block = 2048;
    dim3 cg((unsigned)((d + 255) / 256), (unsigned)((m + rows_per_block - 1) / rows_per_block));
    ni_fid_colsum_kernel<<<cg, 256, 0, st>>>(feats, ld, m, d, stats, rows_per_block);
    int rc = check_launch("ni_fid_accumulate (column sums) launch");
    if (rc != NI_OK) return rc;
    ni_fid_syrk_kernel<<<(unsigned)(ntile * (ntile + 1) / 2), 128, 0, st>>>(feats, ld, m, d, stats + 1 + d, ntile);
    return check_launch("ni_fid_accumulate (syrk) launch");
}
