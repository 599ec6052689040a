// talfe_core.cuh — per-thread building blocks of the fused log-mel kernel.
//
// Everything here is __host__ __device__ and free of CUDA built-ins so that the exact code the
// kernel inlines can also be driven, "thread" by "thread", by the host emulator in
// csrc/host_emul.cpp (CPU unit test of index maps and arithmetic; it is NOT a product path).
//
// Path being implemented (reference: /root/reference/tal/asr/models.py:22-53, whose arithmetic
// is torchaudio MelSpectrogram(n_fft=400, win=400, hop=160, n_mels=80) -> log(.+eps)):
// the 400-point real DFT of two consecutive Hann-windowed frames (a, b), computed by a group of
// 20 threads as a real-structured 20 x 20 Cooley-Tukey split:
//
//   n = j + 20 m,  k = k1 + 20 k2                                   (j, m, k1, k2 in 0..19)
//   X[k1 + 20 k2] = sum_j W20^(j k2) * [ W400^(j k1) * sum_m x[j + 20 m] W20^(m k1) ]
//
//   stage 1 (thread j):  ONE complex FFT-20 of (xa + i xb)[j + 20 m] yields, by conjugate
//                        symmetry, the real-input FFT-20 of both frames: A_a[k1], A_b[k1],
//                        k1 = 0..10.  Twiddle by W400^(j k1).  Rows for stage 2:
//                          row 0      : A_a[0] + i A_b[0]            (both real  -> packed)
//                          row 1..9   : A_a[k1] W^(j k1)             (frame a)
//                          row 10     : (A_a[10] + i A_b[10]) W^(10 j) (both real -> packed)
//                          row 11..19 : A_b[k1-10] W^(j (k1-10))     (frame b)
//   stage 2 (thread c = row): complex FFT-20 over j.  Rows 1..9 / 11..19 give 20 spectrum bins
//                        of one frame each (k = k1 + 20 k2 for k2 < 10, and 400 - k by conjugate
//                        symmetry for k2 >= 10); rows 0 and 10 give bins 20 q and 10 + 20 q of
//                        BOTH frames after an in-register untangle.  Exactly 20 FFTs for 20
//                        threads, 199 power bins per frame, no second exchange.
//
// The FFT-20 itself is a Good-Thomas (prime factor) 4 x 5 split: no internal twiddles.
#pragma once

#include <cuda_runtime.h>

#ifndef TALFE_HD
#define TALFE_HD __host__ __device__ __forceinline__
#endif

namespace talfe {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kHalf = 200;            // centre padding = n_fft / 2
constexpr int kBins = 201;
constexpr int kGroup = 20;            // threads per frame pair
constexpr int kMaxMels = 80;
constexpr int kMelSlots = 4;          // mel m is owned by thread m % 20, slot m / 20
constexpr int kPStride = 212;         // power bins per pair in shared memory (float2 each), padded
constexpr int kERow = 22;             // exchange row stride in float2 (20 + 2 pad -> conflict-free LDS.128)
constexpr int kEGroup = 452;          // exchange group stride in float2 (= 904 words, 8 mod 32)
constexpr int kMaxWeightsPerThread = 32;

typedef float2 cf;

TALFE_HD cf cadd(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
TALFE_HD cf csub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
TALFE_HD cf cmul(cf a, cf b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 5-point DFT constants (forward transform, W5 = exp(-2 pi i / 5))
#define TALFE_C1 0.30901699437494742f    /* cos(2 pi / 5) */
#define TALFE_C2 (-0.80901699437494742f) /* cos(4 pi / 5) */
#define TALFE_S1 0.95105651629515357f    /* sin(2 pi / 5) */
#define TALFE_S2 0.58778525229247313f    /* sin(4 pi / 5) */

TALFE_HD void dft5(cf a0, cf a1, cf a2, cf a3, cf a4, cf& y0, cf& y1, cf& y2, cf& y3, cf& y4) {
    cf t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
    y0 = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
    cf m1 = make_float2(fmaf(TALFE_C2, t2.x, fmaf(TALFE_C1, t1.x, a0.x)), fmaf(TALFE_C2, t2.y, fmaf(TALFE_C1, t1.y, a0.y)));
    cf m2 = make_float2(fmaf(TALFE_C1, t2.x, fmaf(TALFE_C2, t1.x, a0.x)), fmaf(TALFE_C1, t2.y, fmaf(TALFE_C2, t1.y, a0.y)));
    cf s1 = make_float2(fmaf(TALFE_S2, t4.x, TALFE_S1 * t3.x), fmaf(TALFE_S2, t4.y, TALFE_S1 * t3.y));
    cf s2 = make_float2(fmaf(-TALFE_S1, t4.x, TALFE_S2 * t3.x), fmaf(-TALFE_S1, t4.y, TALFE_S2 * t3.y));
    // y1 = m1 - i s1, y4 = m1 + i s1, y2 = m2 - i s2, y3 = m2 + i s2   ( -i (x + i y) = y - i x )
    y1 = make_float2(m1.x + s1.y, m1.y - s1.x);
    y4 = make_float2(m1.x - s1.y, m1.y + s1.x);
    y2 = make_float2(m2.x + s2.y, m2.y - s2.x);
    y3 = make_float2(m2.x - s2.y, m2.y + s2.x);
}

// In-place 20-point complex DFT, natural order in and out (all indices compile-time).
// Good-Thomas: n = (5 n1 + 4 n2) mod 20, k = (5 k1 + 16 k2) mod 20, n1,k1 in 0..3, n2,k2 in 0..4.
TALFE_HD void fft20(cf (&v)[20]) {
    cf t[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        cf a0 = v[(4 * n2) % 20], a1 = v[(5 + 4 * n2) % 20], a2 = v[(10 + 4 * n2) % 20], a3 = v[(15 + 4 * n2) % 20];
        cf s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
        t[0][n2] = cadd(s02, s13);
        t[2][n2] = csub(s02, s13);
        t[1][n2] = make_float2(d02.x + d13.y, d02.y - d13.x);   // d02 - i d13
        t[3][n2] = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cf y0, y1, y2, y3, y4;
        dft5(t[k1][0], t[k1][1], t[k1][2], t[k1][3], t[k1][4], y0, y1, y2, y3, y4);
        v[(5 * k1) % 20] = y0;
        v[(5 * k1 + 16) % 20] = y1;
        v[(5 * k1 + 32) % 20] = y2;
        v[(5 * k1 + 48) % 20] = y3;
        v[(5 * k1 + 64) % 20] = y4;
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 1.  xs points at the first sample of frame a inside the staged tile (frame b starts kHop
// later).  win_t[j*20 + m] = 0.5 * hann[j + 20 m]  (the 0.5 makes A_a = C[k] + conj C[20-k] exact
// scale); tw_t[j*10 + (k1-1)] = W400^(j k1) for k1 = 1..9 and 2 * W400^(10 j) for k1 = 10.
// Writes this thread's column j of the 20 exchange rows.
TALFE_HD void stage1(int j, const float* __restrict__ xs, const float* __restrict__ win_t,
                     const cf* __restrict__ tw_t, cf* __restrict__ e_group) {
    cf z[20];
    const float4* w4 = reinterpret_cast<const float4*>(win_t + j * 20);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        float4 w = w4[q];
        const float* p = xs + j + 80 * q;
        z[4 * q + 0] = make_float2(w.x * p[0], w.x * p[kHop]);
        z[4 * q + 1] = make_float2(w.y * p[20], w.y * p[20 + kHop]);
        z[4 * q + 2] = make_float2(w.z * p[40], w.z * p[40 + kHop]);
        z[4 * q + 3] = make_float2(w.w * p[60], w.w * p[60 + kHop]);
    }
    fft20(z);
    const float4* t4 = reinterpret_cast<const float4*>(tw_t + j * 10);
    cf* col = e_group + j;
    col[0] = make_float2(2.0f * z[0].x, 2.0f * z[0].y);                 // row 0: A_a[0] + i A_b[0]
#pragma unroll
    for (int h = 0; h < 5; ++h) {
        float4 tt = t4[h];                                              // twiddles k1 = 2h+1, 2h+2
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k1 = 2 * h + 1 + u;
            cf w = u == 0 ? make_float2(tt.x, tt.y) : make_float2(tt.z, tt.w);
            if (k1 < 10) {
                cf p = z[k1], q = z[20 - k1];
                cf aa = make_float2(p.x + q.x, p.y - q.y);              // A_a[k1] = C[k1] + conj C[20-k1]
                cf ab = make_float2(p.y + q.y, q.x - p.x);              // A_b[k1] = (C[k1] - conj C[20-k1]) / i
                col[k1 * kERow] = cmul(aa, w);
                col[(k1 + 10) * kERow] = cmul(ab, w);
            } else {
                col[10 * kERow] = cmul(z[10], w);                       // row 10 (w carries the factor 2)
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 2.  Thread c transforms exchange row c and writes |X|^2 into the pair's power array
// p2[k] = (P_a[k], P_b[k]).  `special` (c == 0 or c == 10) rows hold both frames packed.
TALFE_HD void stage2_load(int c, const cf* __restrict__ e_group, cf (&v)[20]) {
    const float4* row = reinterpret_cast<const float4*>(e_group + c * kERow);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        float4 r = row[q];
        v[2 * q] = make_float2(r.x, r.y);
        v[2 * q + 1] = make_float2(r.z, r.w);
    }
}

TALFE_HD void stage2_normal(int c, cf (&v)[20], float* __restrict__ p2) {
    fft20(v);
    const int frame = c >= 10 ? 1 : 0;
    const int k1 = c - 10 * frame;                                      // 1..9
    float* lo = p2 + 2 * k1 + frame;                                    // bins k1 + 20 q
    float* hi = p2 + 2 * (20 - k1) + frame;                             // bins (20 - k1) + 20 q
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        lo[40 * q] = fmaf(v[q].x, v[q].x, v[q].y * v[q].y);
        hi[40 * q] = fmaf(v[19 - q].x, v[19 - q].x, v[19 - q].y * v[19 - q].y);
    }
}

TALFE_HD void stage2_special(int c, cf (&v)[20], float* __restrict__ p2) {
    fft20(v);
    // c == 10: V[q] pairs with V[19-q] -> bins 10 + 20 q.   c == 0: V[q+1] pairs with V[19-q] -> bins 20 (q+1).
    const bool zero = (c == 0);
    cf* out = reinterpret_cast<cf*>(p2) + (zero ? 20 : 10);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        cf p = zero ? v[q + 1] : v[q];
        cf r = v[19 - q];
        float ar = p.x + r.x, ai = p.y - r.y;                           // 2 X_a
        float br = p.x - r.x, bi = p.y + r.y;                           // 2 i X_b
        // q == 9 with c == 0 is bin 200 (never weighted); it lands in a padding slot.
        out[20 * q] = make_float2(0.25f * fmaf(ar, ar, ai * ai), 0.25f * fmaf(br, br, bi * bi));
    }
}

// ---------------------------------------------------------------------------------------------
// Mel projection + log for the mels owned by thread c (m = c + 20 i).  mel_lo[m] = first bin,
// w_t[c * wstride + off_i + r] = fb[lo[m] + r, m] (zero padded to the slot's common width R_i).
// y[2*i + f] = log(mel_f[m] + eps) for frame f of the pair.
struct MelLayout {
    int n_mels;
    int n_slots;
    int width[kMelSlots];     // R_i
    int offset[kMelSlots];    // prefix sums of R_i
    int wstride;              // weights per thread (multiple of 4)
};

TALFE_HD float fast_log(float x) {
#ifdef __CUDA_ARCH__
    return __logf(x);
#else
    return logf(x);
#endif
}

TALFE_HD void mel_log(int c, const MelLayout& ml, const cf* __restrict__ p2, const float* __restrict__ w_t,
                      const int* __restrict__ mel_lo, float eps, float (&y)[2 * kMelSlots]) {
    const float* w = w_t + c * ml.wstride;
#pragma unroll
    for (int i = 0; i < kMelSlots; ++i) {
        float acc_a = 0.f, acc_b = 0.f;
        if (i < ml.n_slots) {
            const int m = c + 20 * i;
            const cf* p = p2 + mel_lo[m < ml.n_mels ? m : 0];
            const float* wi = w + ml.offset[i];
            for (int r = 0; r < ml.width[i]; ++r) {
                cf pw = p[r];
                float wr = wi[r];
                acc_a = fmaf(wr, pw.x, acc_a);
                acc_b = fmaf(wr, pw.y, acc_b);
            }
        }
        y[2 * i] = fast_log(acc_a + eps);
        y[2 * i + 1] = fast_log(acc_b + eps);
    }
}

}  // namespace talfe
