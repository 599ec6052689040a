// tc5_probe.cu — the tensor-pipe experiment of tools/tc_probe.cu repeated with Blackwell's own instruction: the 20-point
// complex DFT stage of the 400-point transform as a [128 rows x 40] . [40 x 32] real contraction issued as
// tcgen05.mma.cta_group::1.kind::tf32 (M 128, N 32, K 8 per instruction, accumulator in tensor memory), once with plain TF32
// operands and once with the 3-product hi / lo split that fp32-level accuracy needs, with the A operand
//   (SS) in shared memory (canonical K-major no-swizzle layout, written by the CTA's threads: the operand preparation —
//        split + 16-byte stores — is part of what is timed, because the product kernel would have to do it per sample), or
//   (TS) in TENSOR MEMORY (tcgen05.st by the row's own thread; what round 1's verdict asked to be measured: "stage-2 FFT-20 as
//        tcgen05.mma with the A operand in TMEM").
// Per tile of 128 rows: operand preparation by 128 threads -> 5 (x1) or 15 (x3) MMAs by one thread -> tcgen05.commit ->
// tcgen05.ld of the 32 accumulator columns by the row's thread -> power of the 16 (re, im) pairs.  Four CTAs per SM overlap
// their phases.  Reported: rows per second, max error of the accumulator against a float64 evaluation; run under ncu for
// sm__pipe_tensor / FMA pipe utilisation (tools/gpu_tc_probe.sh).  Every wait is bounded (a wrong descriptor must not hang
// the GPU): on a timeout the kernel sets a flag and leaves.  Not part of the product library.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tc5_probe.cu -o /tmp/tc5_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__host__ __device__ inline float synth_val(unsigned row, unsigned col) {        // deterministic operand, |x| <= 1
    unsigned h = row * 2654435761u + col * 40503u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return ((int)(h & 0xFFFF) - 32768) * (1.0f / 32768.0f);
}

constexpr int kM = 128, kK = 40, kN = 32, kChunks = kK / 4;              // K in 16-byte chunks of 4 tf32
constexpr int kABytes = kChunks * kM * 16;                                // 20 480: one A part (hi or lo)
constexpr int kBBytes = kChunks * kN * 16;                                // 5 120: one B part
constexpr int kTmemCols = 128;                                            // D 32 | A hi 40 | A lo 40

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// canonical K-major, no swizzle: 16-byte unit (row r, K chunk c) at c * rows * 16 + r * 16  (8 consecutive rows = one 128-byte core matrix)
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (unsigned long long)((saddr & 0x3FFFF) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
           ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);   // version 1 (sm_100), swizzle none, base offset 0
}
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kN >> 3) << 17) | ((unsigned)(kM >> 4) << 24);   // f32 += tf32 x tf32, K-major A and B

__device__ __forceinline__ void mma_ss(unsigned d_tmem, unsigned long long da, unsigned long long db, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long db, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <bool kSplit, bool kTmemA>
__global__ void __launch_bounds__(128) tc5_kernel(const float* __restrict__ bmat /* [40][40] row-major, first kN columns used */,
                                                  float* out, int iters, float* sample /* [128][32] of block 0, tile 0 */, int* fail) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_a = smem;                                          // [2 parts][kABytes]   (SS variants)
    unsigned char* s_b = smem + 2 * kABytes;                            // [2 parts][kBBytes]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_b + 2 * kBBytes);
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    // B (hi / lo) in the canonical layout: "row" = output column n, K = input index k
    for (int i = tid; i < kN * kChunks; i += 128) {
        const int n = i % kN, c = i / kN;
        float4 h, l;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = bmat[(4 * c + q) * 40 + n];
        h = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
        l = make_float4(tf32_hi(v[0] - h.x), tf32_hi(v[1] - h.y), tf32_hi(v[2] - h.z), tf32_hi(v[3] - h.w));
        *reinterpret_cast<float4*>(s_b + c * kN * 16 + n * 16) = h;
        *reinterpret_cast<float4*>(s_b + kBBytes + c * kN * 16 + n * 16) = l;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tm = *s_tmem;
    const unsigned tm_lane = tm + ((unsigned)(32 * warp) << 16);        // this warp's 32 lanes
    const unsigned d_col = 0, ahi_col = 32, alo_col = 72;
    const unsigned long long db_hi = make_desc(smem_u32(s_b), kN * 16, 128), db_lo = make_desc(smem_u32(s_b + kBBytes), kN * 16, 128);
    const unsigned long long da_hi = make_desc(smem_u32(s_a), kM * 16, 128), da_lo = make_desc(smem_u32(s_a + kABytes), kM * 16, 128);

    float acc = 0.f;
    unsigned phase = 0;
    bool dead = false;
    for (int it = 0; it < iters && !dead; ++it) {
        const unsigned row = ((unsigned)blockIdx.x * 7u + (unsigned)it) * 128u + (unsigned)tid;
        // ---- operand preparation: this thread's row, 40 values, hi (and lo) parts
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = synth_val(row, 4 * c + q) + acc * 1e-30f;
            const float4 h = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
            const float4 l = make_float4(v[0] - h.x, v[1] - h.y, v[2] - h.z, v[3] - h.w);       // (the MMA truncates to tf32 itself)
            if (kTmemA) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tm_lane + ahi_col + 4 * c), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w));
                if (kSplit)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tm_lane + alo_col + 4 * c), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w));
            } else {
                *reinterpret_cast<float4*>(s_a + c * kM * 16 + tid * 16) = h;
                if (kSplit) *reinterpret_cast<float4*>(s_a + kABytes + c * kM * 16 + tid * 16) = l;
            }
        }
        if (kTmemA) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- the contraction: one thread issues, the accumulator lands in tensor-memory columns [0, 32)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kK / 8; ++ks) {
                const unsigned long long bh = db_hi + (unsigned long long)((2 * ks * kN * 16) >> 4), bl = db_lo + (unsigned long long)((2 * ks * kN * 16) >> 4);
                if (kTmemA) {
                    mma_ts(tm + d_col, tm + ahi_col + 8 * ks, bh, ks > 0);
                    if (kSplit) { mma_ts(tm + d_col, tm + alo_col + 8 * ks, bh, 1); mma_ts(tm + d_col, tm + ahi_col + 8 * ks, bl, 1); }
                } else {
                    const unsigned long long ah = da_hi + (unsigned long long)((2 * ks * kM * 16) >> 4), al = da_lo + (unsigned long long)((2 * ks * kM * 16) >> 4);
                    mma_ss(tm + d_col, ah, bh, ks > 0);
                    if (kSplit) { mma_ss(tm + d_col, al, bh, 1); mma_ss(tm + d_col, ah, bl, 1); }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(s_bar)) : "memory");
        }
        // ---- wait for the MMAs (bounded), then the epilogue: accumulator row -> registers -> power of the (re, im) pairs
        unsigned done = 0;
        for (int spin = 0; spin < (1 << 20) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(s_bar)), "r"(phase) : "memory");
        if (!done) { if (tid == 0) atomicExch(fail, 1 + it); dead = true; }
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!dead) {
            float d[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]), "=f"(d[8]), "=f"(d[9]),
                  "=f"(d[10]), "=f"(d[11]), "=f"(d[12]), "=f"(d[13]), "=f"(d[14]), "=f"(d[15]), "=f"(d[16]), "=f"(d[17]), "=f"(d[18]),
                  "=f"(d[19]), "=f"(d[20]), "=f"(d[21]), "=f"(d[22]), "=f"(d[23]), "=f"(d[24]), "=f"(d[25]), "=f"(d[26]), "=f"(d[27]),
                  "=f"(d[28]), "=f"(d[29]), "=f"(d[30]), "=f"(d[31])
                : "r"(tm_lane + d_col));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("" : "+f"(d[i]));
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += fmaf(d[2 * i], d[2 * i], d[2 * i + 1] * d[2 * i + 1]);
            if (it == 0 && blockIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) sample[tid * 32 + i] = d[i];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                // the accumulator and the operands may be overwritten
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (acc == 123.456f) out[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(kTmemCols) : "memory");
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
    const int sms = prop.multiProcessorCount, blocks = sms * 4, iters = 200;
    printf("%s  SMs %d\n", prop.name, sms);
    float *out, *sample; int* fail;
    CK(cudaMalloc(&out, 4096)); CK(cudaMalloc(&sample, 128 * 32 * 4)); CK(cudaMalloc(&fail, 4));
    std::vector<float> F(40 * 40), hs(128 * 32);
    for (int j = 0; j < 20; ++j) for (int k = 0; k < 20; ++k) {
        const double th = 2.0 * M_PI * ((j * k) % 20) / 20.0;
        F[(2 * j) * 40 + 2 * k] = (float)cos(th);      F[(2 * j) * 40 + 2 * k + 1] = (float)-sin(th);
        F[(2 * j + 1) * 40 + 2 * k] = (float)sin(th);  F[(2 * j + 1) * 40 + 2 * k + 1] = (float)cos(th);
    }
    float* dF; CK(cudaMalloc(&dF, F.size() * 4)); CK(cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = 2 * kABytes + 2 * kBBytes + 64;
    const double rows = (double)blocks * iters * 128.0;

    auto run = [&](auto kernel, const char* name, int mmas) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(fail, 0, 4));
        float t = time_ms([&] { kernel<<<blocks, 128, smem>>>(dF, out, iters, sample, fail); });
        int hf = 0; CK(cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hs.data(), sample, 128 * 32 * 4, cudaMemcpyDeviceToHost));
        double err = 0.0, scale = 0.0;
        for (int r = 0; r < 128; ++r) for (int c = 0; c < 32; ++c) {
            double v = 0;
            for (int k = 0; k < 40; ++k) v += (double)synth_val(r, k) * F[k * 40 + c];
            err = fmax(err, fabs(hs[r * 32 + c] - v)); scale = fmax(scale, fabs(v));
        }
        printf("%-58s: %8.2f G rows/s   %2d tcgen05.mma M128 N32 K8 per 128 rows   max abs err %.2e (values up to %.1f)%s\n", name,
               rows / (t * 1e-3) / 1e9, mmas, err, scale, hf ? "   ** MMA WAIT TIMED OUT **" : "");
    };
    run(tc5_kernel<false, false>, "(E) DFT-20 stage, tcgen05 TF32 x1, A in shared memory", 5);
    run(tc5_kernel<true, false>, "(E) DFT-20 stage, tcgen05 TF32 x3, A in shared memory", 15);
    run(tc5_kernel<false, true>, "(F) DFT-20 stage, tcgen05 TF32 x1, A in tensor memory", 5);
    run(tc5_kernel<true, true>, "(F) DFT-20 stage, tcgen05 TF32 x3, A in tensor memory", 15);
    return 0;
}
