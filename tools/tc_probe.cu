// tc_probe.cu — the tensor-pipe experiment BASELINE.json's north star asks for ("whether the DFT or mel stages run on
// the FP32 pipe or as a TF32 dense contraction is decided by ncu evidence"), as a stand-alone measurement:
//
//   one 20-point complex DFT stage of the 400-point transform (what a stage-2 consumer thread computes per exchange
//   row: 20 complex in -> 20 complex out -> 20 power values), operands in REGISTERS in both variants (no shared-memory
//   traffic for the data; the contraction variant fetches its constant matrix fragments from shared memory), as
//     (A) the FP32-pipe prime-factor FFT-20 of csrc/talfe_core.cuh (112 packed instructions per row),
//     (B) a dense [rows x 40] . [40 x 40] real contraction on the tensor pipe: mma.sync m16n8k8 TF32, once with plain
//         TF32 operands (x1) and once with the 3-product hi/lo split that recovers fp32-level accuracy (x3),
//   and the mel projection [frames x 200] . [200 x 80] (the one true contraction of the path) as
//     (C) the sparse FP32 form (392 FMAs per frame, both frames of a pair per FFMA2) and
//     (D) dense TF32 x3 on the tensor pipe.
//   Reported: rows (frames) per second per GPU, tensor / FMA pipe instruction counts per row, max error against a
//   float64 evaluation.  Run under `ncu --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active...`
//   for the pipe-utilisation evidence (tools/gpu_tc_probe.sh).  Not part of the product library.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tal_asrd_b200/csrc tools/tc_probe.cu -o /tmp/tc_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "talfe_core.cuh"

using namespace talfe;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__host__ __device__ inline float synth_val(unsigned row, unsigned col) {        // deterministic operand, |x| <= 1
    unsigned h = row * 2654435761u + col * 40503u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return ((int)(h & 0xFFFF) - 32768) * (1.0f / 32768.0f);
}

// ------------------------------------------------------------------------------------------ (A) FP32 FFT-20 + power
__global__ void __launch_bounds__(256) fft_fp32_kernel(float* out, int iters, float* sample /* [32][20] powers of block 0 warp 0 */) {
    const unsigned row0 = (blockIdx.x * blockDim.x + threadIdx.x) * 7u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        cf v[20];
#pragma unroll
        for (int j = 0; j < 20; ++j) v[j] = make_float2(synth_val(row0 + it, 2 * j) + acc * 1e-30f, synth_val(row0 + it, 2 * j + 1));
        fft20<true>(v);
#pragma unroll
        for (int j = 0; j < 20; ++j) {
            const float p = fmaf(v[j].x, v[j].x, v[j].y * v[j].y);
            acc += p;
            if (it == 0 && blockIdx.x == 0 && threadIdx.x < 32) sample[threadIdx.x * 20 + j] = p;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// ------------------------------------------------------------------------------------------ (B) TF32 mma.sync contraction
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ unsigned tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ unsigned tf32_lo(float x) { return __float_as_uint(x - __uint_as_float(tf32_hi(x))) & 0xFFFFE000u; }

// [rows x K] . [K x N]: K = 8 KS, N = 8 NT.  B fragments (hi and lo) live in shared memory in fragment order:
// s_b[((ks * NT + nt) * 2 + part) * 64 + lane * 2 + {0,1}]  -> one conflict-free LDS.64 per fragment and lane.
template <int KS, int NT, bool kSplit>
__global__ void __launch_bounds__(256) mma_kernel(const float* __restrict__ bmat /* [8 KS][8 NT] row-major */, float* out, int iters,
                                                  float* sample /* [16][8 NT] of block 0 warp 0 */, float scale_power) {
    extern __shared__ float2 s_b[];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
        const int l = i & 31, f = i >> 5, nt = f % NT, ks = f / NT;
        const float b0 = bmat[(8 * ks + (l & 3)) * (8 * NT) + 8 * nt + (l >> 2)];
        const float b1 = bmat[(8 * ks + (l & 3) + 4) * (8 * NT) + 8 * nt + (l >> 2)];
        const unsigned h0 = __float_as_uint(b0) & 0xFFFFE000u, h1 = __float_as_uint(b1) & 0xFFFFE000u;
        s_b[(f * 2 + 0) * 32 + l] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
        s_b[(f * 2 + 1) * 32 + l] = make_float2(b0 - __uint_as_float(h0), b1 - __uint_as_float(h1));
    }
    __syncthreads();
    const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const unsigned row0 = (warp_global * 7u + it) * 16u;
        float d[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            float av[4] = {synth_val(row0 + g, 8 * ks + t) + acc * 1e-30f, synth_val(row0 + g + 8, 8 * ks + t),
                           synth_val(row0 + g, 8 * ks + t + 4), synth_val(row0 + g + 8, 8 * ks + t + 4)};
            unsigned ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { ah[q] = tf32_hi(av[q]); al[q] = kSplit ? tf32_lo(av[q]) : 0u; }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 bh = s_b[((ks * NT + nt) * 2 + 0) * 32 + lane];
                const unsigned b_hi[2] = {__float_as_uint(bh.x), __float_as_uint(bh.y)};
                mma_tf32(d[nt], ah, b_hi);
                if (kSplit) {
                    const float2 bl = s_b[((ks * NT + nt) * 2 + 1) * 32 + lane];
                    const unsigned b_lo[2] = {__float_as_uint(bl.x) & 0xFFFFE000u, __float_as_uint(bl.y) & 0xFFFFE000u};
                    mma_tf32(d[nt], al, b_hi);
                    mma_tf32(d[nt], ah, b_lo);
                }
            }
        }
        // epilogue of the DFT stage: power of (re, im) column pairs (c0, c1 are columns 2t, 2t+1 = re, im of one output)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if (scale_power != 0.f) {
                acc += fmaf(d[nt][0], d[nt][0], d[nt][1] * d[nt][1]) + fmaf(d[nt][2], d[nt][2], d[nt][3] * d[nt][3]);
            } else {
                acc += d[nt][0] + d[nt][1] + d[nt][2] + d[nt][3];
            }
            if (it == 0 && warp_global == 0) {
                sample[g * (8 * NT) + 8 * nt + 2 * t] = d[nt][0];
                sample[g * (8 * NT) + 8 * nt + 2 * t + 1] = d[nt][1];
                sample[(g + 8) * (8 * NT) + 8 * nt + 2 * t] = d[nt][2];
                sample[(g + 8) * (8 * NT) + 8 * nt + 2 * t + 1] = d[nt][3];
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// ------------------------------------------------------------------------------------------ (C) sparse FP32 mel projection
// thread = (pair, mel lane): 4 mels (widths 2 / 4 / 7 / 13), both frames of the pair per FFMA2, weights in registers,
// power values synthesized in registers (the product kernel reads them from shared memory: that traffic is not the subject here)
__global__ void __launch_bounds__(256) mel_fp32_kernel(float* out, int iters) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    float w[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) w[i] = 0.5f + 0.01f * ((id + i) & 15);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        cf a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        const float base = acc * 1e-30f + it;
#pragma unroll
        for (int r = 0; r < 26; ++r) {
            const cf p = make_float2(base + r, base - r);
            if (r < 2) a0 = cfma_s(w[r], p, a0);
            else if (r < 6) a1 = cfma_s(w[r], p, a1);
            else if (r < 13) a2 = cfma_s(w[r], p, a2);
            else a3 = cfma_s(w[r], p, a3);
        }
        acc += a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F> static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 200;
    printf("%s  SMs %d\n", prop.name, sms);
    float *out, *sample; CK(cudaMalloc(&out, 4096)); CK(cudaMalloc(&sample, 16 * 80 * 4 * 4));
    std::vector<float> hs(16 * 80 * 4);

    // the DFT-20 stage as a real 40 x 40 matrix: row 2j + {0,1} = re / im of input j, column 2k + {0,1} of output k
    std::vector<float> F(40 * 40);
    for (int j = 0; j < 20; ++j) for (int k = 0; k < 20; ++k) {
        const double th = 2.0 * M_PI * ((j * k) % 20) / 20.0;
        F[(2 * j) * 40 + 2 * k] = (float)cos(th);      F[(2 * j) * 40 + 2 * k + 1] = (float)-sin(th);
        F[(2 * j + 1) * 40 + 2 * k] = (float)sin(th);  F[(2 * j + 1) * 40 + 2 * k + 1] = (float)cos(th);
    }
    float* dF; CK(cudaMalloc(&dF, F.size() * 4)); CK(cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice));

    // (A)
    const double rows_a = (double)blocks * threads * iters;
    float ms = time_ms([&] { fft_fp32_kernel<<<blocks, threads>>>(out, iters, sample); });
    CK(cudaMemcpy(hs.data(), sample, 32 * 20 * 4, cudaMemcpyDeviceToHost));
    double err_a = 0.0;
    for (int r = 0; r < 32; ++r) for (int k = 0; k < 20; ++k) {
        double re = 0, im = 0;
        for (int j = 0; j < 20; ++j) {
            const double th = 2.0 * M_PI * ((j * k) % 20) / 20.0, x = synth_val(r * 7u, 2 * j), y = synth_val(r * 7u, 2 * j + 1);
            re += x * cos(th) + y * sin(th); im += y * cos(th) - x * sin(th);
        }
        const double p = re * re + im * im;
        err_a = fmax(err_a, fabs(hs[r * 20 + k] - p) / fmax(1.0, p));
    }
    printf("(A) FFT-20 + power, FP32 pipe (packed f32x2)     : %8.2f G rows/s   112 FFMA2/FADD2/FMUL2 + 40 scalar per row   max rel err %.2e\n",
           rows_a / (ms * 1e-3) / 1e9, err_a);

    // (B) 40 x 40 contraction: KS = 5, NT = 5
    auto run_b = [&](bool split) {
        const double rows_b = (double)blocks * (threads / 32) * 16.0 * iters;
        const size_t smem = 5 * 5 * 2 * 32 * sizeof(float2);
        float t = split ? time_ms([&] { mma_kernel<5, 5, true><<<blocks, threads, smem>>>(dF, out, iters, sample, 1.f); })
                        : time_ms([&] { mma_kernel<5, 5, false><<<blocks, threads, smem>>>(dF, out, iters, sample, 1.f); });
        CK(cudaMemcpy(hs.data(), sample, 16 * 40 * 4, cudaMemcpyDeviceToHost));
        double err = 0.0, scale = 0.0;
        for (int r = 0; r < 16; ++r) for (int c = 0; c < 40; ++c) {
            double v = 0;
            for (int k = 0; k < 40; ++k) v += (double)synth_val(r, k) * F[k * 40 + c];
            err = fmax(err, fabs(hs[r * 40 + c] - v)); scale = fmax(scale, fabs(v));
        }
        printf("(B) DFT-20 stage as [rows x 40].[40 x 40], TF32 x%d : %8.2f G rows/s   %d mma.sync.m16n8k8 per 16 rows          max abs err %.2e (values up to %.1f)\n",
               split ? 3 : 1, rows_b / (t * 1e-3) / 1e9, split ? 75 : 25, err, scale);
    };
    run_b(false);
    run_b(true);

    // (C) / (D) mel projection
    const double frames_c = (double)blocks * threads * iters * 2.0 / 20.0;       // 20 mel lanes per frame pair
    ms = time_ms([&] { mel_fp32_kernel<<<blocks, threads>>>(out, iters); });
    printf("(C) mel projection, sparse FP32 (392 nz, FFMA2)     : %8.2f G frames/s  26 FFMA2 per thread and frame pair (x 20 lanes)\n",
           frames_c / (ms * 1e-3) / 1e9);
    std::vector<float> W(200 * 80);
    for (int b = 0; b < 200; ++b) for (int m = 0; m < 80; ++m) W[b * 80 + m] = (abs(b - (m * 5 / 2 + 2)) < 4 + m / 8) ? 0.25f + 0.001f * ((b * 7 + m) & 63) : 0.f;
    float* dW; CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    {
        const double frames_d = (double)blocks * (threads / 32) * 16.0 * iters;
        const size_t smem = 25 * 10 * 2 * 32 * sizeof(float2);                  // 128 KB of weight fragments (hi + lo)
        CK(cudaFuncSetAttribute(mma_kernel<25, 10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int blocks_d = sms;                                               // one CTA per SM (shared memory)
        const double frames_dd = (double)blocks_d * (threads / 32) * 16.0 * 40;
        float t = time_ms([&] { mma_kernel<25, 10, true><<<blocks_d, threads, smem>>>(dW, out, 40, sample, 0.f); });
        CK(cudaMemcpy(hs.data(), sample, 16 * 80 * 4, cudaMemcpyDeviceToHost));
        double err = 0.0;
        for (int r = 0; r < 16; ++r) for (int c = 0; c < 80; ++c) {
            double v = 0;
            for (int k = 0; k < 200; ++k) v += (double)synth_val(r, k) * W[k * 80 + c];
            err = fmax(err, fabs(hs[r * 80 + c] - v) / fmax(1.0, fabs(v)));
        }
        (void)frames_d;
        printf("(D) mel projection, dense TF32 x3 (mma.sync)        : %8.2f G frames/s  750 mma.sync.m16n8k8 per 16 frames         max rel err %.2e\n",
               frames_dd / (t * 1e-3) / 1e9, err);
    }
    return 0;
}
