// talfe_generic.cuh — log-mel for frame geometries other than the reference's (n_fft 400, hop 160): LogMelSpec(sr != 16000)
// (included by talfe.cu).  The reference derives n_fft = win = int(0.025 sr) and hop = int(0.010 sr) from its `sr`
// argument (tal/asr/models.py:22-32); it only ever runs at 16 kHz, which the specialised kernels serve.  This kernel
// keeps the constructor's full signature alive ON THE GPU (there is no CPU path anywhere): any n_fft <= 1280, odd or
// even, any hop, any filterbank with n_mels <= 80.  It is a correct, plain formulation — a direct DFT from a twiddle table in
// shared memory — not a tuned one: 5–30 M frames/s depending on n_fft, against 2.5 G for the 16 kHz kernel.
//
// One CTA = one tile of 32 consecutive frames of one row (the tile grid, the statistics slots and every later kernel are
// shared with the specialised path):
//   stage 0  the tile's samples -> shared memory as fp32 (reflection at the row's ends, zero beyond the buffer), element-wise
//   stage 1  thread = bin k: X_f[k] = sum_n w[n] x[f hop + n] W_N^(n k) for all 32 frames at once (64 accumulators; the
//            twiddle index n k mod N is carried, the sample reads are broadcasts) -> power P[f][k] in shared memory
//   stage 2  thread = (frame, mel): sum over the mel's support of fb[k][m] P[f][k], log(. + eps), store, partial sums
#pragma once

namespace {

constexpr int kGenThreads = 256;
constexpr int kGenFrames = kFramesPerTile;                             // 32
constexpr int kGenMaxNfft = 1280;                                      // 48 kHz (n_fft 1200) and a little more: shared-memory bound

struct GenericArgs {
    int nfft, hop, half, bins;
    const float* win;              // [nfft]
    const float2* tw;              // [nfft]: (cos, -sin)(2 pi i / nfft)
    const float* fb;               // [bins][n_mels]
    const int* mel_lo;             // first / last bin with a non-zero weight (lo > hi: empty filter)
    const int* mel_hi;
    int n_mels;
};

__host__ __device__ inline size_t generic_smem_bytes(int nfft, int hop) {
    const int bins = nfft / 2 + 1;
    return (size_t)(hop * (kGenFrames - 1) + nfft) * 4 + (size_t)nfft * 4 + (size_t)nfft * 8 + (size_t)kGenFrames * (bins + 1) * 4 + 64;
}

template <typename XT>
__global__ void __launch_bounds__(kGenThreads) logmel_generic_kernel(const KernelArgs a, const GenericArgs g) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int span = g.hop * (kGenFrames - 1) + g.nfft;
    float* s_x = reinterpret_cast<float*>(smem);
    float* s_win = s_x + span;
    float2* s_tw = reinterpret_cast<float2*>(s_win + g.nfft + ((span + g.nfft) & 1));       // 8-byte aligned
    float* s_p = reinterpret_cast<float*>(s_tw + g.nfft);
    const int pstride = g.bins + 1;
    __shared__ double s_red[2][kGenThreads / 32];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int row = tile / a.tiles_per_row, tq = tile - row * a.tiles_per_row;
    const int t0 = a.frame0 + tq * kGenFrames;
    int L, t_end;
    if (a.lens) {
        L = (int)min(a.lens[row], (long long)kMaxSamples);
        t_end = (int)min((long long)a.frame_end, frames_of(L, g.hop, g.nfft));
    } else {
        L = a.total_len;
        t_end = a.t_end_const;
    }
    const bool active = t0 < t_end;
    for (int i = tid; i < g.nfft; i += kGenThreads) { s_win[i] = g.win[i]; s_tw[i] = g.tw[i]; }
    cudaGridDependencySynchronize();
    double sum = 0.0, sumsq = 0.0;
    float* out_row = a.out + (a.out_offsets ? a.out_offsets[row] * g.n_mels : (long long)row * a.out_row_stride);
    const int nfr = min(kGenFrames, a.frame_end - t0);
    if (active) {
        const XT* rowp = reinterpret_cast<const XT*>(a.wave) + (long long)row * a.row_stride;
        const long long s0 = (long long)g.hop * t0 - g.half;
        const float scale = XLayout<XT>::kScale;
        for (int i = tid; i < span; i += kGenThreads) {
            long long gi = s0 + i;
            if (gi < 0) gi = -gi;                                       // reflect, no edge repeat
            if (gi >= L) gi = 2ll * (L - 1) - gi;
            const long long bi = gi - a.origin;
            float v = 0.f;
            if (gi >= 0 && gi < L && bi >= 0 && bi < a.buf_len) v = x_to_float(rowp[bi]) * scale;
            s_x[i] = v;
        }
        __syncthreads();
        for (int k = tid; k < g.bins; k += kGenThreads) {
            float re[kGenFrames], im[kGenFrames];
#pragma unroll
            for (int f = 0; f < kGenFrames; ++f) { re[f] = 0.f; im[f] = 0.f; }
            int idx = 0;
            for (int n = 0; n < g.nfft; ++n) {
                const float2 w = s_tw[idx];
                const float wn = s_win[n];
                const float c = wn * w.x, s = wn * w.y;
                const float* xp = s_x + n;
#pragma unroll
                for (int f = 0; f < kGenFrames; ++f) {
                    const float x = xp[f * g.hop];
                    re[f] = fmaf(x, c, re[f]);
                    im[f] = fmaf(x, s, im[f]);
                }
                idx += k;
                if (idx >= g.nfft) idx -= g.nfft;
            }
#pragma unroll
            for (int f = 0; f < kGenFrames; ++f) s_p[f * pstride + k] = fmaf(re[f], re[f], im[f] * im[f]);
        }
        __syncthreads();
    }
    // mel projection + log + store: thread = (frame, mel); frames of the tile beyond the row's own length are written as 0
    const bool mt = a.out_layout == TALFE_LAYOUT_MT;
    for (int i = tid; i < nfr * g.n_mels; i += kGenThreads) {
        int f, m;
        if (mt) { m = i / nfr; f = i - m * nfr; } else { f = i / g.n_mels; m = i - f * g.n_mels; }
        const bool valid = active && t0 + f < t_end;
        float y = 0.f;
        if (valid) {
            float acc = 0.f;
            const int lo = g.mel_lo[m], hi = g.mel_hi[m];
            for (int k = lo; k <= hi; ++k) acc = fmaf(g.fb[k * g.n_mels + m], s_p[f * pstride + k], acc);
            y = fast_log(acc + a.eps);
            sum += (double)y;
            sumsq += (double)y * (double)y;
        } else if (a.out_offsets) {
            continue;                                                   // packed output has no padding frames
        }
        if (mt) out_row[(long long)m * a.n_frames + (t0 - a.frame0) + f] = y;
        else out_row[(long long)(t0 - a.frame0 + f) * g.n_mels + m] = y;
    }
    // per-CTA partial sums, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = sum; s_red[1][tid >> 5] = sumsq; }
    __syncthreads();
    if (tid == 0) {
        double ts = 0.0, tq2 = 0.0;
        for (int w = 0; w < kGenThreads / 32; ++w) { ts += s_red[0][w]; tq2 += s_red[1][w]; }
        if (a.partials_per_tile) {
            a.partials[(long long)tile * kWarps] = make_double2(ts, tq2);
            for (int w = 1; w < kWarps; ++w) a.partials[(long long)tile * kWarps + w] = make_double2(0.0, 0.0);
        } else {
            a.partials[tile] = make_double2(ts, tq2);
        }
    }
}

}  // namespace
