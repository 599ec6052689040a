// Minimal producer/consumer hand-off through an mbarrier (release arrive / acquire try_wait), correct under the PTX
// memory model.  Used to find out whether compute-sanitizer's racecheck models mbarrier synchronisation: if it
// reports hazards here, the hazards it reports on the E hand-off of logmel_ws_kernel are the same tool limitation.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(float* out, int rounds) {
    __shared__ __align__(16) float buf[2][64];
    __shared__ __align__(8) unsigned long long full[2], empty[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + b)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(empty + b)));
        }
    }
    __syncthreads();
    float acc = 0.f;
    for (int k = 0; k < rounds; ++k) {
        const int b = k & 1;
        if (warp == 0) {                                        // producer
            if (k >= 2) {
                unsigned done = 0;
                while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(empty + b)), "r"(((k - 2) >> 1) & 1) : "memory");
            }
            buf[b][lane] = (float)(k * 32 + lane);
            buf[b][lane + 32] = (float)k;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(full + b)) : "memory");
            __syncwarp();
        } else {                                                // consumer
            unsigned done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(full + b)), "r"((k >> 1) & 1) : "memory");
            acc += buf[b][31 - lane] + buf[b][lane + 32];
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + b)) : "memory");
            __syncwarp();
        }
    }
    if (warp == 1) out[lane] = acc;
}
int main() {
    float* d; cudaMalloc(&d, 32 * sizeof(float));
    probe<<<1, 64>>>(d, 100);
    float h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double want = 0; for (int k = 0; k < 100; ++k) want += k * 32 + 31 + k;          // lane 0 reads element 31
    printf("probe %s: lane0 %.0f expected %.0f (%s)\n", cudaGetErrorString(cudaGetLastError()), h[0], want, h[0] == want ? "ok" : "MISMATCH");
    return 0;
}
