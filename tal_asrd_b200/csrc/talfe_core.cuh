// talfe_core.cuh — per-thread building blocks of the fused log-mel kernel.
//
// Everything here is __host__ __device__ and free of CUDA built-ins so that the exact code the
// kernel inlines can also be driven, "thread" by "thread", by the host emulator in
// csrc/host_emul.cu (CPU unit test of index maps and arithmetic; it is NOT a product path).
//
// Path being implemented (reference: /root/reference/tal/asr/models.py:22-53, whose arithmetic
// is torchaudio MelSpectrogram(n_fft=400, win=400, hop=160, n_mels=80) -> log(.+eps)):
// the 400-point real DFT of two consecutive Hann-windowed frames (a, b), computed by a group of
// 20 threads as a real-structured 20 x 20 Cooley-Tukey split:
//
//   n = j + 20 m,  k = k1 + 20 k2                                   (j, m, k1, k2 in 0..19)
//   X[k1 + 20 k2] = sum_j W20^(j k2) * [ W400^(j k1) * sum_m x[j + 20 m] W20^(m k1) ]
//
//   stage 1 (thread j):  the real-input FFT-20 of both frames' column j (frame a and frame b ride in the two halves
//                        of every packed register: rfft20_pair_windowed): A_a[k1], A_b[k1],
//                        k1 = 0..10.  Twiddle by W400^(j k1).  Exchange rows for stage 2:
//                          row 2(k1-1)   : A_a[k1] W^(j k1)   k1 = 1..9       (frame a)
//                          row 2(k1-1)+1 : A_b[k1] W^(j k1)   k1 = 1..9       (frame b)
//                          row 18        : (A_a[0] + i A_b[0]) / 2            (both real -> packed)
//                          row 19        : (A_a[10] + i A_b[10]) W^(10 j) / 2 (both real -> packed)
//   stage 2 (thread = row): complex FFT-20 over j.  Rows 0..17 give 20 spectrum bins of one frame
//                        each (k = k1 + 20 q for output q < 10, and 400 - k by conjugate symmetry
//                        for q >= 10); rows 18 and 19 give bins 20 q and 10 + 20 q of BOTH frames
//                        after an in-register untangle.  Exactly 20 FFTs for 20 threads, 199 power
//                        bins per frame, no second exchange.
//
// The FFT-20 itself is a Good-Thomas (prime factor) 4 x 5 split: no internal twiddles.
//
// Shared-memory layouts (bank-conflict analysis in DESIGN.md §4):
//   waveform tile : sample i of the tile lives at i + 20 * (i / 320)  -> thread (g, j) reads word
//                   340 g + j + const, i.e. consecutive lanes hit consecutive banks (16-bit tiles: skew 24);
//   exchange      : row stride 22 complex, group stride 452 complex (904 words = 8 mod 32);
//   power         : float2 (P_a[k], P_b[k]) at index k, group stride `pstride` = 9 (mod 16) so that
//                   lane r of group g writes word 18 g + r + const.
#pragma once

#include <cuda_fp16.h>
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
constexpr int kERow = 22;             // exchange row stride in float2 (20 + 2 pad -> conflict-free LDS.128)
constexpr int kEGroup = 452;          // exchange group stride in float2 (= 904 words, 8 mod 32)
constexpr int kXBlock = 320;          // waveform tile: a skew after every kXBlock samples (one TMA piece each)
// Element type of the staged waveform tile: fp32, or the narrow on-disk / decoding formats kept narrow in
// shared memory and widened in stage 1 (int16 PCM: the 1/32768 of torchaudio.load is folded into the window).
// The skew keeps every TMA piece 16-byte aligned and spreads consecutive pairs over distinct banks.
template <typename XT> struct XLayout {
    static constexpr int kSkew = sizeof(XT) == 4 ? 20 : 24;
    static constexpr int kGroup = kXBlock + kSkew;           // distance between the first samples of consecutive pairs
    static constexpr float kScale = 1.0f;
};
template <> struct XLayout<short> {
    static constexpr int kSkew = 24;
    static constexpr int kGroup = kXBlock + kSkew;
    static constexpr float kScale = 1.0f / 32768.0f;
};
TALFE_HD float x_to_float(float v) { return v; }
TALFE_HD float x_to_float(short v) { return (float)v; }
TALFE_HD float x_to_float(__half v) { return __half2float(v); }

// the reference configuration (80 HTK mels): common widths per slot, compile-time unrolled
constexpr int kRefW0 = 2, kRefW1 = 4, kRefW2 = 7, kRefW3 = 13;
constexpr int kRefWStride = 28;

typedef float2 cf;

// Complex helpers.  On the device the two components of a complex number ride in one 64-bit register
// pair and use Blackwell's packed fp32x2 arithmetic (add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 /
// FFMA2): the same IEEE operations as the scalar forms (bit-identical results), half the issue slots.
// The host emulator compiles the scalar forms.
#ifdef __CUDA_ARCH__
TALFE_HD cf cadd(cf a, cf b) {
    cf r;
    asm("{ .reg .b64 ra, rb, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rr, ra, rb; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
TALFE_HD cf csub(cf a, cf b) {
    cf r;
    asm("{ .reg .b64 ra, rb, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rr, ra, rb; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
// s * t + a and s * t for a real scalar s
TALFE_HD cf cfma_s(float s, cf t, cf a) {
    cf r;
    asm("{ .reg .b64 rs, rt, ra, rr; mov.b64 rs, {%2,%2}; mov.b64 rt, {%3,%4}; mov.b64 ra, {%5,%6}; fma.rn.f32x2 rr, rs, rt, ra; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(t.x), "f"(t.y), "f"(a.x), "f"(a.y));
    return r;
}
TALFE_HD cf cmul_s(float s, cf t) {
    cf r;
    asm("{ .reg .b64 rs, rt, rr; mov.b64 rs, {%2,%2}; mov.b64 rt, {%3,%4}; mul.rn.f32x2 rr, rs, rt; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(t.x), "f"(t.y));
    return r;
}
// s * t - a
TALFE_HD cf cfms_s(float s, cf t, cf a) {
    cf r;
    asm("{ .reg .b64 rs, rt, ra, rr; .reg .f32 n0, n1; neg.f32 n0, %5; neg.f32 n1, %6; mov.b64 rs, {%2,%2}; mov.b64 rt, {%3,%4}; mov.b64 ra, {n0,n1}; fma.rn.f32x2 rr, rs, rt, ra; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(s), "f"(t.x), "f"(t.y), "f"(a.x), "f"(a.y));
    return r;
}
#else
TALFE_HD cf cadd(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
TALFE_HD cf csub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
TALFE_HD cf cfma_s(float s, cf t, cf a) { return make_float2(fmaf(s, t.x, a.x), fmaf(s, t.y, a.y)); }
TALFE_HD cf cmul_s(float s, cf t) { return make_float2(s * t.x, s * t.y); }
TALFE_HD cf cfms_s(float s, cf t, cf a) { return make_float2(fmaf(s, t.x, -a.x), fmaf(s, t.y, -a.y)); }
#endif
// the same three forms with a DIFFERENT multiplier in each half (frame-per-lane kernel: the halves are two adjacent
// columns of one frame, whose window taps differ)
#ifdef __CUDA_ARCH__
TALFE_HD cf cmul_s(cf s, cf t) {
    cf r;
    asm("{ .reg .b64 rs, rt, rr; mov.b64 rs, {%2,%3}; mov.b64 rt, {%4,%5}; mul.rn.f32x2 rr, rs, rt; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(s.x), "f"(s.y), "f"(t.x), "f"(t.y));
    return r;
}
TALFE_HD cf cfma_s(cf s, cf t, cf a) {
    cf r;
    asm("{ .reg .b64 rs, rt, ra, rr; mov.b64 rs, {%2,%3}; mov.b64 rt, {%4,%5}; mov.b64 ra, {%6,%7}; fma.rn.f32x2 rr, rs, rt, ra; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(s.x), "f"(s.y), "f"(t.x), "f"(t.y), "f"(a.x), "f"(a.y));
    return r;
}
TALFE_HD cf cfms_s(cf s, cf t, cf a) {
    cf r;
    asm("{ .reg .b64 rs, rt, ra, rr; .reg .f32 n0, n1; neg.f32 n0, %6; neg.f32 n1, %7; mov.b64 rs, {%2,%3}; mov.b64 rt, {%4,%5}; mov.b64 ra, {n0,n1}; fma.rn.f32x2 rr, rs, rt, ra; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(s.x), "f"(s.y), "f"(t.x), "f"(t.y), "f"(a.x), "f"(a.y));
    return r;
}
#else
TALFE_HD cf cmul_s(cf s, cf t) { return make_float2(s.x * t.x, s.y * t.y); }
TALFE_HD cf cfma_s(cf s, cf t, cf a) { return make_float2(fmaf(s.x, t.x, a.x), fmaf(s.y, t.y, a.y)); }
TALFE_HD cf cfms_s(cf s, cf t, cf a) { return make_float2(fmaf(s.x, t.x, -a.x), fmaf(s.y, t.y, -a.y)); }
#endif
// explicit fused forms: the contraction the compiler would pick for a.x*b.x - a.y*b.y may differ from one kernel
// to the next; fixing it keeps every variant of the kernel (and the host emulator) bit-identical
TALFE_HD cf cmul(cf a, cf b) { return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)); }

// a - i b and a + i b.  Scalar form: two FADDs on mixed halves.  Packed form (kPacked, device only): ONE FFMA2 whose
// first operand is b with its halves swapped and a per-lane sign (SASS: Rb.F32x2.LO_HI.NP times the constant pair
// (1, 1)); fma(y, +-1, x) rounds exactly like x +- y, so both forms are bit-identical (tests: ws vs legacy kernel).
template <bool kPacked> TALFE_HD cf csub_i(cf a, cf b) {
#ifdef __CUDA_ARCH__
    if (kPacked) {
        cf r;
        asm("{ .reg .b64 ra, rb, rc, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%5,%4}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rr, rb, rc, ra; mov.b64 {%0,%1}, rr; }"
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(1.0f), "f"(-1.0f));
        return r;
    }
#endif
    return make_float2(a.x + b.y, a.y - b.x);
}
template <bool kPacked> TALFE_HD cf cadd_i(cf a, cf b) {
#ifdef __CUDA_ARCH__
    if (kPacked) {
        cf r;
        asm("{ .reg .b64 ra, rb, rc, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%5,%4}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rr, rb, rc, ra; mov.b64 {%0,%1}, rr; }"
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(-1.0f), "f"(1.0f));
        return r;
    }
#endif
    return make_float2(a.x - b.y, a.y + b.x);
}
// i w = (-w.y, w.x): one packed multiply by the constant pair (1, 1) with swap + per-lane sign on the device
TALFE_HD cf times_i(cf w) {
#ifdef __CUDA_ARCH__
    cf r;
    asm("{ .reg .b64 rb, rc, rr; mov.b64 rb, {%3,%2}; mov.b64 rc, {%4,%5}; mul.rn.f32x2 rr, rb, rc; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(w.x), "f"(w.y), "f"(-1.0f), "f"(1.0f));
    return r;
#else
    return make_float2(-w.y, w.x);
#endif
}
// -i w = (w.y, -w.x)
TALFE_HD cf times_minus_i(cf w) {
#ifdef __CUDA_ARCH__
    cf r;
    asm("{ .reg .b64 rb, rc, rr; mov.b64 rb, {%3,%2}; mov.b64 rc, {%4,%5}; mul.rn.f32x2 rr, rb, rc; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(w.x), "f"(w.y), "f"(1.0f), "f"(-1.0f));
    return r;
#else
    return make_float2(w.y, -w.x);
#endif
}
// s1 * u + s2 * v for real scalars s1, s2 (v's product is rounded first): FMUL2 + FFMA2 with scalar-broadcast operands.
// With u = w, v = i w this is the complex product (s1 + i s2) w, bit-identical to cmul((s1, s2), w).
TALFE_HD cf cfma_ss(float s1, cf u, float s2, cf v) { return cfma_s(s1, u, cmul_s(s2, v)); }

// 5-point DFT constants (forward transform, W5 = exp(-2 pi i / 5))
#define TALFE_C1 0.30901699437494742f    /* cos(2 pi / 5) */
#define TALFE_C2 (-0.80901699437494742f) /* cos(4 pi / 5) */
#define TALFE_S1 0.95105651629515357f    /* sin(2 pi / 5) */
#define TALFE_S2 0.58778525229247313f    /* sin(4 pi / 5) */

template <bool kPacked = false>
TALFE_HD void dft5(cf a0, cf a1, cf a2, cf a3, cf a4, cf& y0, cf& y1, cf& y2, cf& y3, cf& y4) {
    const cf t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
    y0 = cadd(cadd(a0, t1), t2);
    const cf m1 = cfma_s(TALFE_C2, t2, cfma_s(TALFE_C1, t1, a0));
    const cf m2 = cfma_s(TALFE_C1, t2, cfma_s(TALFE_C2, t1, a0));
    const cf s1 = cfma_s(TALFE_S2, t4, cmul_s(TALFE_S1, t3));
    const cf s2 = cfma_s(-TALFE_S1, t4, cmul_s(TALFE_S2, t3));
    // y1 = m1 - i s1, y4 = m1 + i s1, y2 = m2 - i s2, y3 = m2 + i s2   ( -i (x + i y) = y - i x )
    y1 = csub_i<kPacked>(m1, s1);
    y4 = cadd_i<kPacked>(m1, s1);
    y2 = csub_i<kPacked>(m2, s2);
    y3 = cadd_i<kPacked>(m2, s2);
}

// In-place 20-point complex DFT, natural order in and out (all indices compile-time).
// Good-Thomas: n = (5 n1 + 4 n2) mod 20, k = (5 k1 + 16 k2) mod 20, n1,k1 in 0..3, n2,k2 in 0..4.
// kPacked selects the single-instruction forms of the multiplications by -+i (same bits, fewer issue slots).
// Second half (the four 5-point DFTs) of the 20-point transform; t[k1][n2] come from the 4-point stage.
template <bool kPacked>
TALFE_HD void fft20_dft5s(cf (&t)[4][5], cf (&v)[20]);

// First half: the five 4-point DFTs.
template <bool kPacked>
TALFE_HD void fft20_fft4s(const cf (&v)[20], cf (&t)[4][5]) {
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        cf a0 = v[(4 * n2) % 20], a1 = v[(5 + 4 * n2) % 20], a2 = v[(10 + 4 * n2) % 20], a3 = v[(15 + 4 * n2) % 20];
        cf s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
        t[0][n2] = cadd(s02, s13);
        t[2][n2] = csub(s02, s13);
        t[1][n2] = csub_i<kPacked>(d02, d13);                   // d02 - i d13
        t[3][n2] = cadd_i<kPacked>(d02, d13);                   // d02 + i d13
    }
}

template <bool kPacked = false>
TALFE_HD void fft20(cf (&v)[20]) {
    cf t[4][5];
    fft20_fft4s<kPacked>(v, t);
    fft20_dft5s<kPacked>(t, v);
}

// The same transform of the WINDOWED input v[m] = win[m] x[m], with the window folded into the first butterflies:
//   a2 = w2 x2,  s02 = fma(w0, x0, a2),  d02 = fma(w0, x0, -a2)   (likewise a3, s13, d13)
// i.e. 10 multiplies + 20 fused multiply-adds instead of 20 multiplies + 20 adds, and one rounding fewer on half of
// the terms.  Written out explicitly (ptxas would contract a packed multiply feeding a packed add on its own, but
// which pairs it picks is its business) so that every kernel and the host emulator round identically.
template <bool kPacked = false>
TALFE_HD void fft20_windowed(const cf (&x)[20], const float (&win)[20], cf (&v)[20]) {
    cf t[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        const int i0 = (4 * n2) % 20, i1 = (5 + 4 * n2) % 20, i2 = (10 + 4 * n2) % 20, i3 = (15 + 4 * n2) % 20;
        const cf a2 = cmul_s(win[i2], x[i2]), a3 = cmul_s(win[i3], x[i3]);
        const cf s02 = cfma_s(win[i0], x[i0], a2), d02 = cfms_s(win[i0], x[i0], a2);
        const cf s13 = cfma_s(win[i1], x[i1], a3), d13 = cfms_s(win[i1], x[i1], a3);
        t[0][n2] = cadd(s02, s13);
        t[2][n2] = csub(s02, s13);
        t[1][n2] = csub_i<kPacked>(d02, d13);
        t[3][n2] = cadd_i<kPacked>(d02, d13);
    }
    fft20_dft5s<kPacked>(t, v);
}

template <bool kPacked>
TALFE_HD void fft20_dft5s(cf (&t)[4][5], cf (&v)[20]) {
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cf y0, y1, y2, y3, y4;
        dft5<kPacked>(t[k1][0], t[k1][1], t[k1][2], t[k1][3], t[k1][4], y0, y1, y2, y3, y4);
        v[(5 * k1) % 20] = y0;
        v[(5 * k1 + 16) % 20] = y1;
        v[(5 * k1 + 32) % 20] = y2;
        v[(5 * k1 + 48) % 20] = y3;
        v[(5 * k1 + 64) % 20] = y4;
    }
}

// ---------------------------------------------------------------------------------------------
// Stage-1 transform: the REAL-input 20-point DFT of both frames of a pair at once.  Every cf in this function is a
// (frame a, frame b) pair of real numbers — not a complex number — so each packed instruction advances the two
// frames' independent real transforms in lock step.  Compared with one complex FFT-20 of xa + i xb followed by the
// conjugate-symmetry untangle (round 1), the real-input structure needs no untangle and drops the redundant half of
// the prime-factor butterflies: 104 packed instructions per pair column instead of 122 + 18.
//
// Same Good-Thomas 4 x 5 index maps as fft20 (n = 5 n1 + 4 n2, k = 5 k1 + 16 k2 mod 20) and the same window folding as
// fft20_windowed.  The 4-point stage of a real sequence yields T0, T2 real and T1 = d02 - i d13 = conj T3, so
//   k1 = 0, 2 : real 5-point DFTs (14 instructions each)      -> bins 0, 4, 8 (12, 16 mirror) / 10, 6, 2 (14, 18 mirror)
//   k1 = 1    : one complex 5-point DFT (36 instructions)     -> bins 5, 1, 9 and, conjugated, 17 -> 3, 13 -> 7
//   k1 = 3    : the mirror image of k1 = 1, never computed.
// Output: re[k], im[k] for k = 0..10 (im[0], im[10] are identically zero and not written).  To keep every output a
// single instruction some imaginary parts come out NEGATED: im[k] holds -Im V[k] where rfft20_im_negated(k); the
// twiddle step that follows absorbs the sign by multiplying with -i w instead of i w (no extra instruction).
TALFE_HD constexpr bool rfft20_im_negated(int k) { return k == 1 || k == 2 || k == 5 || k == 6; }

template <typename WT /* float: one tap for both halves; cf: one per half */>
TALFE_HD void rfft20_pair_windowed(const cf (&x)[20], const WT (&win)[20], cf (&re)[11], cf (&im)[11]) {
    cf t0[5], t2[5], p[5], e[5];                       // T0, T2 (real), T1 = (p, -e)
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        const int i0 = (4 * n2) % 20, i1 = (5 + 4 * n2) % 20, i2 = (10 + 4 * n2) % 20, i3 = (15 + 4 * n2) % 20;
        const cf a2 = cmul_s(win[i2], x[i2]), a3 = cmul_s(win[i3], x[i3]);
        const cf s02 = cfma_s(win[i0], x[i0], a2), s13 = cfma_s(win[i1], x[i1], a3);
        p[n2] = cfms_s(win[i0], x[i0], a2);            // d02
        e[n2] = cfms_s(win[i1], x[i1], a3);            // d13
        t0[n2] = cadd(s02, s13);
        t2[n2] = csub(s02, s13);
    }
    {   // k1 = 0: real DFT-5 of T0 -> V[0] = y0, V[4] = y4 = (m1, s1), V[8] = y3 = (m2, s2)
        const cf u1 = cadd(t0[1], t0[4]), u2 = cadd(t0[2], t0[3]), u3 = csub(t0[1], t0[4]), u4 = csub(t0[2], t0[3]);
        re[0] = cadd(cadd(t0[0], u1), u2);
        re[4] = cfma_s(TALFE_C2, u2, cfma_s(TALFE_C1, u1, t0[0]));
        re[8] = cfma_s(TALFE_C1, u2, cfma_s(TALFE_C2, u1, t0[0]));
        im[4] = cfma_s(TALFE_S2, u4, cmul_s(TALFE_S1, u3));
        im[8] = cfma_s(-TALFE_S1, u4, cmul_s(TALFE_S2, u3));
    }
    {   // k1 = 2: real DFT-5 of T2 -> V[10] = y0, V[6] = y1 = (m1, -s1), V[2] = y2 = (m2, -s2)   (im held negated)
        const cf u1 = cadd(t2[1], t2[4]), u2 = cadd(t2[2], t2[3]), u3 = csub(t2[1], t2[4]), u4 = csub(t2[2], t2[3]);
        re[10] = cadd(cadd(t2[0], u1), u2);
        re[6] = cfma_s(TALFE_C2, u2, cfma_s(TALFE_C1, u1, t2[0]));
        re[2] = cfma_s(TALFE_C1, u2, cfma_s(TALFE_C2, u1, t2[0]));
        im[6] = cfma_s(TALFE_S2, u4, cmul_s(TALFE_S1, u3));
        im[2] = cfma_s(-TALFE_S1, u4, cmul_s(TALFE_S2, u3));
    }
    {   // k1 = 1: complex DFT-5 of T1[n2] = p[n2] - i e[n2]; the "n" quantities below carry MINUS the imaginary part
        const cf u1r = cadd(p[1], p[4]), u2r = cadd(p[2], p[3]), u3r = csub(p[1], p[4]), u4r = csub(p[2], p[3]);
        const cf u1n = cadd(e[1], e[4]), u2n = cadd(e[2], e[3]), u3n = csub(e[1], e[4]), u4n = csub(e[2], e[3]);
        re[5] = cadd(cadd(p[0], u1r), u2r);            // V[5] = y0
        im[5] = cadd(cadd(e[0], u1n), u2n);            //        (negated)
        const cf m1r = cfma_s(TALFE_C2, u2r, cfma_s(TALFE_C1, u1r, p[0])), m1n = cfma_s(TALFE_C2, u2n, cfma_s(TALFE_C1, u1n, e[0]));
        const cf m2r = cfma_s(TALFE_C1, u2r, cfma_s(TALFE_C2, u1r, p[0])), m2n = cfma_s(TALFE_C1, u2n, cfma_s(TALFE_C2, u1n, e[0]));
        const cf s1r = cfma_s(TALFE_S2, u4r, cmul_s(TALFE_S1, u3r)), s1n = cfma_s(TALFE_S2, u4n, cmul_s(TALFE_S1, u3n));
        const cf s2r = cfma_s(-TALFE_S1, u4r, cmul_s(TALFE_S2, u3r)), s2n = cfma_s(-TALFE_S1, u4n, cmul_s(TALFE_S2, u3n));
        // y1 = m1 - i s1, y4 = m1 + i s1, y2 = m2 - i s2, y3 = m2 + i s2 with Im m = -m.n, Im s = -s.n
        re[1] = csub(m1r, s1n); im[1] = cadd(m1n, s1r);            // V[1] = y1          (negated: Im y1 = -m1n - s1r)
        re[9] = cadd(m1r, s1n); im[9] = csub(s1r, m1n);            // V[9] = y4
        re[3] = csub(m2r, s2n); im[3] = cadd(m2n, s2r);            // V[3] = conj y2     (Im y2 = -m2n - s2r)
        re[7] = cadd(m2r, s2n); im[7] = csub(m2n, s2r);            // V[7] = conj y3     (Im y3 = s2r - m2n)
    }
}

// Exchange rows are stored in permuted slots: stage-2 lane `row` (thread 18 g + row reads with LDS.128,
// 8 lanes per wavefront) fetches slot row_slot(row).  The permutation makes every 8-lane window of the
// thread order 18 g + row hit 8 distinct 16-byte bank groups ((11 slot + 226 g) mod 8); found by search,
// see DESIGN.md §4.  Slots 2 and 9 hold the packed rows 18 and 19.
TALFE_HD constexpr int row_slot(int row) {
    constexpr int kSlot[20] = {11, 16, 6, 17, 15, 4, 10, 13, 19, 8, 1, 14, 7, 12, 18, 5, 3, 0, 2, 9};
    return kSlot[row];
}

// position of tile sample i inside the skewed waveform buffer
template <typename XT> TALFE_HD int xskew(int i) { return i + XLayout<XT>::kSkew * (i / kXBlock); }

// ---------------------------------------------------------------------------------------------
// Stage 1.  xg points at this pair's first sample inside the skewed tile (s_x + XLayout<XT>::kGroup * g);
// frame a = samples 0..399 of the pair, frame b = samples 160..559.
// win_t[j*20 + m] = 0.5 * hann[j + 20 m]: the transform runs at half scale;
// tw_t[j*10 + (k1-1)] = 2 W400^(j k1) for k1 = 1..9 restores it for rows 0..17, tw_t[j*10 + 9] = W400^(10 j).  Rows 18 / 19
// (the packed k1 = 0 / 10 rows) stay at half scale, which the power computation of stage 2 absorbs ((X/2 + X/2)^2 = |X|^2).
// Writes this thread's column j of the 20 exchange rows.
TALFE_HD void load_window(int j, const float* __restrict__ win_t, float scale, float (&win)[20]) {
    const float4* w4 = reinterpret_cast<const float4*>(win_t + j * 20);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const float4 w = w4[q];
        win[4 * q] = w.x * scale; win[4 * q + 1] = w.y * scale; win[4 * q + 2] = w.z * scale; win[4 * q + 3] = w.w * scale;
    }
}

template <typename XT>
TALFE_HD void stage1(int j, const XT* __restrict__ xg, const float (&win)[20],
                     const cf* __restrict__ tw_t, cf* __restrict__ e_group) {
    cf xin[20], re[11], im[11];
    const XT* p = xg + j;
    constexpr int kSkew = XLayout<XT>::kSkew;
#pragma unroll
    for (int m = 0; m < 20; ++m) {
        // sample j + 20 m of frame a, j + 20 m + 160 of the pair for frame b, with the block skew
        const int ia = 20 * m + (20 * m >= kXBlock ? kSkew : 0);
        const int ib = 20 * m + kHop + (20 * m + kHop >= kXBlock ? kSkew : 0);
        xin[m] = make_float2(x_to_float(p[ia]), x_to_float(p[ib]));
    }
    rfft20_pair_windowed(xin, win, re, im);
    const float4* t4 = reinterpret_cast<const float4*>(tw_t + j * 10);
    cf* col = e_group + j;
    col[row_slot(18) * kERow] = re[0];                                            // row 18: (A_a[0] + i A_b[0]) / 2
#pragma unroll
    for (int h = 0; h < 5; ++h) {
        const float4 tt = t4[h];                                        // twiddles k1 = 2h+1, 2h+2
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k1 = 2 * h + 1 + u;
            const cf w = u == 0 ? make_float2(tt.x, tt.y) : make_float2(tt.z, tt.w);
            if (k1 < 10) {
                const float sg = rfft20_im_negated(k1) ? -1.f : 1.f;
                col[row_slot(2 * (k1 - 1)) * kERow] = cmul(make_float2(re[k1].x, sg * im[k1].x), w);        // A_a[k1] W
                col[row_slot(2 * (k1 - 1) + 1) * kERow] = cmul(make_float2(re[k1].y, sg * im[k1].y), w);    // A_b[k1] W
            } else {
                col[row_slot(19) * kERow] = cmul(re[10], w);                      // row 19: (A_a[10] + i A_b[10]) W^(10 j) / 2
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 2.  Thread `row` transforms exchange row `row` and writes |X|^2 into the pair's power array
// p2[k] = (P_a[k], P_b[k]).  Rows 18 / 19 hold both frames packed.
TALFE_HD void stage2_load(const cf* __restrict__ e_row /* e_group + row_slot(row) * kERow */, cf (&v)[20]) {
    const float4* src = reinterpret_cast<const float4*>(e_row);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        const float4 r = src[q];
        v[2 * q] = make_float2(r.x, r.y);
        v[2 * q + 1] = make_float2(r.z, r.w);
    }
}

TALFE_HD void stage2_normal(int row, cf (&v)[20], float* __restrict__ p2) {
    fft20(v);
    const int frame = row & 1;
    const int k1 = 1 + (row >> 1);                                      // 1..9
    float* lo = p2 + 2 * k1 + frame;                                    // bins k1 + 20 q          -> word row + 2 + 40 q
    float* hi = p2 + 2 * (20 - k1) + frame;                             // bins (20 - k1) + 20 q
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        lo[40 * q] = fmaf(v[q].x, v[q].x, v[q].y * v[q].y);
        hi[40 * q] = fmaf(v[19 - q].x, v[19 - q].x, v[19 - q].y * v[19 - q].y);
    }
}

TALFE_HD void stage2_special(int row, cf (&v)[20], float* __restrict__ p2) {
    fft20(v);
    // row 19: V[q] pairs with V[19-q] -> bins 10 + 20 q.   row 18: V[q+1] pairs with V[19-q] -> bins 20 (q+1).
    const bool zero = (row == 18);
    cf* out = reinterpret_cast<cf*>(p2) + (zero ? 20 : 10);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        const cf p = zero ? v[q + 1] : v[q];
        const cf r = v[19 - q];
        const cf sm = cadd(p, r), df = csub(p, r);
        const float ar = sm.x, ai = df.y;                               // X_a   (rows 18/19 carry half scale)
        const float br = df.x, bi = sm.y;                               // i X_b
        // q == 9 on row 18 would be bin 200, which carries no mel weight: not stored (padding stays 0)
        if (!(zero && q == 9))
            out[20 * q] = make_float2(fmaf(ar, ar, ai * ai), fmaf(br, br, bi * bi));
    }
}

// ---------------------------------------------------------------------------------------------
// Mel projection + log.  Lane c owns one mel per slot i (which one: mel_id[i*20 + c], chosen on the
// host to minimise bank conflicts); mel_lo[i*20 + c] = its first bin;
// w_t[c * wstride + off_i + r] = fb[lo + r, mel] (zero padded to the slot's common width R_i).
// y[2*i + f] = log(mel_f + eps) for frame f of the pair.
struct MelLayout {
    int n_mels;
    int n_slots;
    int width[kMelSlots];     // R_i
    int offset[kMelSlots];    // prefix sums of R_i
    int wstride;              // weights per thread (multiple of 4)
};

TALFE_HD bool is_reference_layout(const MelLayout& ml) {
    return ml.n_mels == 80 && ml.n_slots == 4 && ml.width[0] == kRefW0 && ml.width[1] == kRefW1 &&
           ml.width[2] == kRefW2 && ml.width[3] == kRefW3 && ml.wstride == kRefWStride;
}

// ln(x) for x >= eps > 0 (power + eps is never subnormal for any sensible eps): the bare MUFU.LG2 without the
// subnormal-range rescue that __logf wraps around it (3 extra instructions per value)
TALFE_HD float fast_log(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.693147180559945309f;
#else
    return logf(x);
#endif
}
// the same for the two frames of a pair: two MUFU.LG2 and ONE packed multiply (bit-identical to two fast_log calls)
TALFE_HD cf fast_log2x(cf x) {
#ifdef __CUDA_ARCH__
    float ra, rb;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(x.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(x.y));
    return cmul_s(0.693147180559945309f, make_float2(ra, rb));
#else
    return make_float2(logf(x.x), logf(x.y));
#endif
}

// generic widths (any n_mels <= 80 / any filterbank with bounded support)
TALFE_HD void mel_log_generic(int c, const MelLayout& ml, const cf* __restrict__ p2, const float* __restrict__ w_t,
                              const int* __restrict__ mel_lo, float eps, float (&y)[2 * kMelSlots]) {
    const float* w = w_t + c * ml.wstride;
#pragma unroll
    for (int i = 0; i < kMelSlots; ++i) {
        float acc_a = 0.f, acc_b = 0.f;
        if (i < ml.n_slots) {
            const cf* p = p2 + mel_lo[i * 20 + c];
            const float* wi = w + ml.offset[i];
            for (int r = 0; r < ml.width[i]; ++r) {
                const cf pw = p[r];
                const float wr = wi[r];
                acc_a = fmaf(wr, pw.x, acc_a);
                acc_b = fmaf(wr, pw.y, acc_b);
            }
        }
        y[2 * i] = fast_log(acc_a + eps);
        y[2 * i + 1] = fast_log(acc_b + eps);
    }
}

// reference layout: widths (2, 4, 7, 13), 28 weights per thread fetched as 7 x 128 bit, fully unrolled
template <int W, int OFF>
TALFE_HD void mel_slot_ref(const cf* __restrict__ p, const float (&w)[kRefWStride], float eps, float& ya, float& yb) {
    float acc_a = 0.f, acc_b = 0.f;
#pragma unroll
    for (int r = 0; r < W; ++r) {
        const cf pw = p[r];
        acc_a = fmaf(w[OFF + r], pw.x, acc_a);
        acc_b = fmaf(w[OFF + r], pw.y, acc_b);
    }
    ya = fast_log(acc_a + eps);
    yb = fast_log(acc_b + eps);
}

TALFE_HD void mel_log_ref(int c, const cf* __restrict__ p2, const float* __restrict__ w_t, const int (&lo)[kMelSlots],
                          float eps, float (&y)[2 * kMelSlots]) {
    float w[kRefWStride];
    const float4* w4 = reinterpret_cast<const float4*>(w_t + c * kRefWStride);
#pragma unroll
    for (int q = 0; q < kRefWStride / 4; ++q) {
        const float4 t = w4[q];
        w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
    }
    mel_slot_ref<kRefW0, 0>(p2 + lo[0], w, eps, y[0], y[1]);
    mel_slot_ref<kRefW1, kRefW0>(p2 + lo[1], w, eps, y[2], y[3]);
    mel_slot_ref<kRefW2, kRefW0 + kRefW1>(p2 + lo[2], w, eps, y[4], y[5]);
    mel_slot_ref<kRefW3, kRefW0 + kRefW1 + kRefW2>(p2 + lo[3], w, eps, y[6], y[7]);
}


// =============================================================================================
// Warp-specialised path (csrc/talfe_ws.cuh): producer warps run stage 1, consumer warps stage 2 +
// mel.  Same arithmetic as above; what changes is who holds which constants (per-role registers) and
// the shared-memory layouts, all chosen so that EVERY shared access of the steady state is
// conflict-free by construction (checked on the CPU by tests/test_ws_layout.py):
//
//   producers  thread (g1 = t / 20, j = t % 20)   — group-major, as in the legacy kernel;
//   consumers  thread (g  = t % 16, r = t / 16)   — pair-minor: the 16 lanes of a half-warp are the 16
//              frame pairs of the tile, r (exchange row in stage 2, mel lane in the mel stage) is uniform.
//
//   exchange   E[g][row][j] at complex index ws_e_base(g) + 20 row + j, ws_e_base(g) = 404 g + 2 ((g>>2)&1):
//              404 = 4 (mod 16) keeps the group-major STS.64 of stage 1 on 16 distinct 8-byte banks wherever a
//              half-warp straddles two groups (straddles only happen inside blocks of four groups), and the
//              extra 16 bytes of every other block make 8 consecutive pairs hit 8 distinct 16-byte bank
//              groups for the pair-minor LDS.128 of stage 2;
//   power      P[bin][g] (float2: frame a, frame b): stage 2 writes words 32 bin + 2 g + f (one wavefront per
//              warp store), the mel stage reads 16 consecutive float2 per half-warp (one wavefront);
//   features   Y[frame][80] staged per tile, row fr at float offset 80 fr + 4 (fr >> 2) (16-byte aligned rows
//              for the bulk store to global memory; the pad leaves the scalar stores 2-way conflicted); for
//              [.., 80, T] outputs the tile is staged transposed (ws_yt_off) and leaves by coalesced stores.
constexpr int kWsGroups = 16;
constexpr int kWsFrames = 2 * kWsGroups;
constexpr int kWsERow = 20;
constexpr int kWsEGroup = 404;
constexpr int kWsECf = kWsGroups * kWsEGroup + 4;                       // complex entries per exchange buffer
constexpr int kWsPBins = 216;                                           // 200 bins + read padding of the widest slot
constexpr int kWsPCf = kWsPBins * kWsGroups;
constexpr int kWsYtStride = kWsFrames + 1;                              // transposed staging (layout [.., 80, T]): mel row stride 33
constexpr int kWsYFloats = kMaxMels * kWsYtStride;                      // 2640 >= 32 * 80 + 4 * 16 (the [.., T, 80] staging)
TALFE_HD constexpr int ws_e_base(int g) { return g * kWsEGroup + 2 * ((g >> 2) & 1); }
// (16 bytes of padding after every kWsYChunk frames: a chunk leaves as ONE bulk copy.  Issuing a bulk copy costs the
// issuing warp ~270 cycles (profiles/r02_timeline.json), so fewer and larger chunks matter; 4-frame chunks keep the
// scalar staging stores at the 2-way conflicts that 2-frame chunks had, 8-frame chunks would make them 4-way.)
constexpr int kWsYChunk = 4;
TALFE_HD constexpr int ws_y_off(int fr) { return fr * kMaxMels + 4 * (fr / kWsYChunk); }
// [.., 80, T] outputs stage the tile transposed, Yt[mel][frame] at mel * 33 + frame: the mel stage's stores (lanes = 16
// pairs x 2 adjacent mels -> words 2 g + f + 33 r) and the store loop's reads (lane = frame) are both conflict-free
TALFE_HD constexpr int ws_yt_off(int mel, int fr) { return mel * kWsYtStride + fr; }

// Stage 1, first half: 28 waveform samples -> window -> real-input FFT-20 of frame a and frame b (one in each half of
// every packed register), in registers.
template <typename XT>
TALFE_HD void stage1_ws_fft(const XT* __restrict__ p /* xg + j */, const float (&win)[20], cf (&re)[11], cf (&im)[11]) {
    constexpr int kSkew = XLayout<XT>::kSkew;
    cf xin[20];
#pragma unroll
    for (int m = 0; m < 20; ++m) {
        const int ia = 20 * m + (20 * m >= kXBlock ? kSkew : 0);
        const int ib = 20 * m + kHop + (20 * m + kHop >= kXBlock ? kSkew : 0);
        xin[m] = make_float2(x_to_float(p[ia]), x_to_float(p[ib]));
    }
    rfft20_pair_windowed(xin, win, re, im);             // window taps as scalar-broadcast operands of FMUL2 / FFMA2
}

// Stage 1, second half: twiddle by tw[k1-1] = W400^(j k1) and write column j of the pair's 20 exchange rows (same row
// meaning as the legacy stage1; rows in natural order).  With A = re + i im (one frame's half of the packed registers),
// w = tw[k1-1] and iw = i w:   A w = re w + im iw   — an FMUL2 + FFMA2 with scalar-broadcast operands per frame, with the
// same fused / unfused roundings as cmul(): bit-identical to the legacy stage1().  Where the transform left -im, the
// product uses -i w instead (rfft20_im_negated).  5 instructions per k1 for both frames.
TALFE_HD void stage1_ws_store(const cf (&re)[11], const cf (&im)[11], const cf (&tw)[10], cf* __restrict__ col /* E + ws_e_base(g1) + j */) {
    col[18 * kWsERow] = re[0];
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const cf w = tw[k1 - 1], iw = rfft20_im_negated(k1) ? times_minus_i(w) : times_i(w);
        col[(2 * (k1 - 1)) * kWsERow] = cfma_ss(re[k1].x, w, im[k1].x, iw);
        col[(2 * (k1 - 1) + 1) * kWsERow] = cfma_ss(re[k1].y, w, im[k1].y, iw);
    }
    col[19 * kWsERow] = cfma_ss(re[10].x, tw[9], re[10].y, times_i(tw[9]));
}

// Stage 2 (consumer thread (g, r)): |FFT-20(row r)|^2 kept in registers until the power array is free.
// Normal rows r < 18: pw[q] = (bin k1 + 20 q, bin (20 - k1) + 20 q) of frame r & 1, k1 = 1 + r / 2.
// (Measured and rejected in round 2, profiles/r02_ab_packed_power.json: leaving the last butterflies' results paired
// across two output bins — (Re y1, Re y4), (Im y1, Im y4), one FFMA2 each with two scalar-broadcast operands — makes
// |y|^2 of two bins one FMUL2 + one FFMA2 instead of four scalar instructions, 16 FMA-pipe slots fewer per row; the
// extra register traffic pushed the kernel over its 96-register budget: 81.8 us against 77.6 us.)
TALFE_HD void stage2_ws_power_normal(cf (&v)[20], cf (&pw)[10]) {
#if !(defined(TALFE_ABLATE) && (TALFE_ABLATE & 4))
    fft20<true>(v);
#endif
#pragma unroll
    for (int q = 0; q < 10; ++q)
        pw[q] = make_float2(fmaf(v[q].x, v[q].x, v[q].y * v[q].y), fmaf(v[19 - q].x, v[19 - q].x, v[19 - q].y * v[19 - q].y));
}
TALFE_HD void stage2_ws_store_normal(int k1, const cf (&pw)[10], float* __restrict__ pgf /* (float*)P + 2 g + f */) {
    float* lo = pgf + 2 * kWsGroups * k1;
    float* hi = pgf + 2 * kWsGroups * (20 - k1);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        lo[2 * kWsGroups * 20 * q] = pw[q].x;
        hi[2 * kWsGroups * 20 * q] = pw[q].y;
    }
}
// Packed rows: r == 18 (zero = true) -> bins 20 (q + 1); r == 19 -> bins 10 + 20 q; both frames per entry.
TALFE_HD void stage2_ws_power_special(bool zero, cf (&v)[20], cf (&pw)[10]) {
    fft20<true>(v);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        const cf p = zero ? v[q + 1] : v[q];
        const cf r = v[19 - q];
        const cf sm = cadd(p, r), df = csub(p, r);
        pw[q] = make_float2(fmaf(sm.x, sm.x, df.y * df.y), fmaf(df.x, df.x, sm.y * sm.y));
    }
}
TALFE_HD void stage2_ws_store_special(bool zero, const cf (&pw)[10], cf* __restrict__ pg /* P + g */) {
    cf* out = pg + kWsGroups * (zero ? 20 : 10);
#pragma unroll
    for (int q = 0; q < 10; ++q)
        if (!(zero && q == 9)) out[kWsGroups * 20 * q] = pw[q];         // bin 200 carries no mel weight
}

// Mel stage (consumer thread (g, c)): mels c, 20 + c, 40 + c, 60 + c; weights in registers.
template <int W, int OFF>
TALFE_HD void mel_slot_ws(const cf* __restrict__ p /* P + g + 16 lo */, const float (&w)[kRefWStride], float eps, float& ya, float& yb) {
    // both frames of the pair ride in one packed accumulator: FFMA2 with the weight as scalar-broadcast operand
    // (the same two IEEE fmas as the scalar form, half the issue slots)
    cf acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < W; ++r) acc = cfma_s(w[OFF + r], p[kWsGroups * r], acc);
    const cf le = cadd(acc, make_float2(eps, eps));
    const cf y2 = fast_log2x(le);
    ya = y2.x;
    yb = y2.y;
}
TALFE_HD void mel_log_ws(const cf* __restrict__ pg /* P + g */, const float (&w)[kRefWStride], const int (&lo)[kMelSlots], float eps,
                         float (&y)[2 * kMelSlots]) {
    mel_slot_ws<kRefW0, 0>(pg + kWsGroups * lo[0], w, eps, y[0], y[1]);
    mel_slot_ws<kRefW1, kRefW0>(pg + kWsGroups * lo[1], w, eps, y[2], y[3]);
    mel_slot_ws<kRefW2, kRefW0 + kRefW1>(pg + kWsGroups * lo[2], w, eps, y[4], y[5]);
    mel_slot_ws<kRefW3, kRefW0 + kRefW1 + kRefW2>(pg + kWsGroups * lo[3], w, eps, y[6], y[7]);
}

}  // namespace talfe
