// Minimal producer/consumer hand-off through shared memory ordered ONLY by an mbarrier (arrive = release, try_wait = acquire):
// the pattern of the TMA-staged kernel's "tiles are gone" notice.  Correct under the PTX memory model; used to show what
// compute-sanitizer --tool racecheck reports for it (profiles/r02_sanitizer_racecheck.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/racecheck_repro scripts/racecheck_mbarrier_repro.cu
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void handoff(int *out, int rounds)
{
    __shared__ uint64_t full, empty;
    __shared__ int slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto wait = [](uint64_t *b, uint32_t par) {
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(par) : "memory");
    };
    for (int r = 0; r < rounds; ++r) {
        const uint32_t ph = r & 1;
        if (warp == 0) { // producer
            if (lane == 0) {
                wait(&empty, ph ^ 1u);
                slot = r * 7 + 1;
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&full)) : "memory");
            }
        } else { // consumer
            wait(&full, ph);
            const int v = slot;
            if (lane == 0) out[r] = v;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty)) : "memory");
        }
    }
}
int main()
{
    int *d, h[8];
    cudaMalloc(&d, sizeof(h));
    handoff<<<1, 64>>>(d, 8);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int r = 0; r < 8; ++r) ok &= h[r] == r * 7 + 1;
    printf("handoff %s\n", ok ? "correct" : "WRONG");
    return !ok;
}
