// microbench.cu — pipe throughputs on the target GPU that decide the kernel design
// (FP32 scalar vs packed f32x2, shuffle, shared-memory widths, MUFU.LG2).  Build & run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o /tmp/mb && /tmp/mb
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
#define UNROLL 8

template <int MODE>
__global__ void __launch_bounds__(512) k_pipe(float* out, long long* cycles, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.000001f, c = 0.5f;
    unsigned long long p0, p1, p2, p3, pb, pc;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {          // scalar FFMA, 8 independent chains
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        } else if (MODE == 1) {   // packed fma.rn.f32x2, 4 independent chains (= 8 fp32 FMAs)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pb), "l"(pc));
        } else if (MODE == 2) {   // scalar FADD
            a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
        } else if (MODE == 3) {   // packed add.f32x2
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pc));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pc));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pc));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pc));
        } else if (MODE == 4) {   // shuffle
            a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
            a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 8);
            a4 = __shfl_xor_sync(0xffffffffu, a4, 16); a5 = __shfl_xor_sync(0xffffffffu, a5, 1);
            a6 = __shfl_xor_sync(0xffffffffu, a6, 2); a7 = __shfl_xor_sync(0xffffffffu, a7, 4);
        } else if (MODE == 5) {   // MUFU.LG2
            a0 = __log2f(a0); a1 = __log2f(a1); a2 = __log2f(a2); a3 = __log2f(a3);
            a4 = __log2f(a4); a5 = __log2f(a5); a6 = __log2f(a6); a7 = __log2f(a7);
        } else if (MODE == 6) {   // LDS.32 conflict-free
            int i = (threadIdx.x + it) & 1023;
            const float* s = reinterpret_cast<const float*>(sm);
            a0 += s[i]; a1 += s[i + 32]; a2 += s[i + 64]; a3 += s[i + 96];
            a4 += s[i + 128]; a5 += s[i + 160]; a6 += s[i + 192]; a7 += s[i + 224];
        } else if (MODE == 7) {   // LDS.128 conflict-free
            int i = (threadIdx.x + it) & 511;
            float4 v0 = sm[i], v1 = sm[i + 64], v2 = sm[i + 128], v3 = sm[i + 192];
            a0 += v0.x + v0.y; a1 += v0.z + v0.w; a2 += v1.x + v1.y; a3 += v1.z + v1.w;
            a4 += v2.x + v2.y; a5 += v2.z + v2.w; a6 += v3.x + v3.y; a7 += v3.z + v3.w;
        } else if (MODE == 8) {   // FFMA + LDS.128 mix: 16 FFMA per LDS.128
            int i = (threadIdx.x + it) & 1023;
            float4 v0 = sm[i];
            a0 = fmaf(a0, b, v0.x); a1 = fmaf(a1, b, v0.y); a2 = fmaf(a2, b, v0.z); a3 = fmaf(a3, b, v0.w);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        } else if (MODE == 9) {   // scalar FFMA with 3 distinct register sources per op (no immediates / reuse)
            a0 = fmaf(a0, a1, a2); a1 = fmaf(a1, a2, a3); a2 = fmaf(a2, a3, a4); a3 = fmaf(a3, a4, a5);
            a4 = fmaf(a4, a5, a6); a5 = fmaf(a5, a6, a7); a6 = fmaf(a6, a7, a0); a7 = fmaf(a7, a0, a1);
        } else if (MODE == 10) {  // FFMA2 + scalar FADD interleaved (do they dual-issue?)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pb), "l"(pc));
            a0 += c;
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pb), "l"(pc));
            a1 += c;
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pb), "l"(pc));
            a2 += c;
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pb), "l"(pc));
            a3 += c;
        }
    }
    long long t1 = clock64();
    float q0, q1;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(p0 ^ p1 ^ p2 ^ p3));
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + q0 + q1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter, int threads) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks_per_sm = 2048 / threads >= 4 ? 4 : 2048 / threads;
    const int blocks = sms * blocks_per_sm;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k_pipe<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_pipe<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double ops_block = (double)ITERS * lane_ops_per_iter * threads;
    printf("%-34s threads/blk %4d blk/SM %d : %8.2f lane-ops/clk/SM (in-kernel clk)  %8.2f Tops/s wall  (%.3f ms, %.0f cyc)\n", name, threads,
           blocks_per_sm, ops_block * blocks_per_sm / avg, ops_block * blocks / (ms * 1e-3) / 1e12, ms, avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs %d  clock %d kHz  smem/SM %zu  L2 %d\n", p.name, p.multiProcessorCount, p.clockRate, p.sharedMemPerMultiprocessor, p.l2CacheSize);
    run<0>("FFMA scalar (imm/const srcs)", 8, 512);
    run<9>("FFMA scalar (3 reg srcs)", 8, 512);
    run<1>("FFMA2 packed f32x2 (fp32 lanes)", 8, 512);
    run<2>("FADD scalar", 8, 512);
    run<3>("FADD2 packed (fp32 lanes)", 8, 512);
    run<10>("FFMA2 + FADD interleaved (fp32 lanes)", 12, 512);
    run<4>("SHFL.BFLY", 8, 512);
    run<5>("MUFU.LG2", 8, 512);
    run<6>("LDS.32 (+FADD)", 8, 512);
    run<7>("LDS.128 x4 (+8 FADD) [loads]", 4, 512);
    run<8>("16 FFMA + 1 LDS.128 [ffma]", 16, 512);
    run<0>("FFMA scalar, 256 thr", 8, 256);
    run<1>("FFMA2 packed, 256 thr", 8, 256);
    return 0;
}
