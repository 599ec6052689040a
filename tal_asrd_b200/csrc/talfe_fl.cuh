// talfe_fl.cuh — "frame per lane" formulation of K1 (included by talfe.cu after talfe_ws.cuh; fp32 waveforms, reference
// filterbank shape).  EXPERIMENT, opt-in with TALFE_KERNEL=fl: parity-green on the whole GPU suite, measured 100 us on
// 64 x 30 s against 76 us for the warp-specialised kernel (profiles/r02_ab_fl_v3.json, r02_ncu_fl_v3_kernel.txt).
//
// Why a third formulation.  The warp-specialised kernel (talfe_ws.cuh) moves every sample through shared memory four times
// (tile, exchange write + read, power write + 2.6 x read) and its skeleton — every LDS/STS and hand-off, no arithmetic —
// already takes 53 us on 64 x 30 s: the shared-memory pipe (one 128-byte wavefront per clock and SM), not the FP32 pipe,
// bounds it.  Here ONE THREAD OWNS ONE FRAME from waveform to log-mel, and the 20 x 20 transposition
// between the two FFT stages happens inside the thread, through its own lane of TENSOR MEMORY used as a 512-word scratch
// pad (tcgen05.st after stage 1, tcgen05.ld before stage 2; its own datapath, no bank conflicts, no mbarriers).
//   * shared memory carries the waveform tile (100 LDS.128 per frame: 12.5 wavefronts), broadcast reads of the uniform
//     tables and the staged feature rows (20 STS.128): 40 wavefronts per frame instead of 75, 215 warp-instructions per
//     frame instead of 273;
//   * window taps, twiddles and mel weights are the same for every lane at every instruction (uniform tables);
//   * the power spectrum lives in tensor memory too (P[k] at column k), so the mel stage is 80 contiguous tcgen05.ld of
//     2 / 4 / 8 / 16 columns with run-time column addresses instead of 520 shared-memory loads;
//   * waveform tiles arrive by one tensor copy, feature tiles leave by one tensor store.
// What it costs: tensor memory holds ONE frame per lane (420 + 216 of 512 columns, overlapped), i.e. 128 frames in flight
// per SM — four quadrants of 32 lanes.  v1 (one warp per quadrant, everything unrolled: 75 KB of straight-line code) spent
// 45 % of its time waiting for instructions (178 us); v2 (the same as a handful of small loops, tables from the constant
// bank with run-time indices) 140 us: one warp per scheduler cannot hide dependent-issue and constant-load latency; v3
// (this file: TWO warps per quadrant that split every stage of the same 32 frames and meet at four named barriers per
// tile; tables staged in shared memory) 100 us at 41 % issue-slot utilisation — 10 % of it the quantisation tail (6 016
// tiles over 592 quadrants), 5 % the pair barriers.  More warps per quadrant need <= 170 registers per thread, which the
// four-column stage-1 pass (80 loaded samples + a 20-point transform) does not fit.  So on B200 the occupancy that tensor
// memory allows (8 warps) loses to the 20 warps of the shared-memory formulation although it issues 21 % fewer
// instructions and a fifth of the shared-memory wavefronts.
//
// Arithmetic: the same real-structured 20 x 20 Cooley-Tukey split as the other kernels (talfe_core.cuh), with the packed
// f32x2 halves carrying two adjacent COLUMNS of one frame in stage 1 (instead of two frames of one column) and the two
// components of a complex number in stage 2.  Bins k = k1 + 20 q with k1 = 1..9 and the mel sums are bit-identical to the
// other kernels; the rows k1 = 0 and 10 are computed per frame instead of per frame pair (differences <= 6e-5 where a
// filter's energy sits almost entirely in such a bin).
#pragma once

namespace {

constexpr int kFlQuads = 4;                                           // tensor-memory lane quadrants: 4 x 32 frames in flight per SM
constexpr int kFlSplit = 2;                                           // warps that share one quadrant's frames (same lanes, half the work each)
constexpr int kFlWarps = kFlQuads * kFlSplit;
constexpr int kFlThreads = 32 * kFlWarps;
constexpr int kFlFrames = 32;                                         // one frame per lane
constexpr int kFlTileSamples = kHop * kFlFrames + (kNfft - kHop);     // 5360
constexpr int kFlRowPitch = kHop + 4;                                 // 164 floats per 160-sample row: lane stride 41 x 16 bytes
constexpr int kFlRows = (kFlTileSamples + kHop - 1) / kHop;           // 34 rows of 160 samples
constexpr int kFlSpan = (kFlRows - 1) * kHop + kFlRowPitch;           // 5444 samples touched in global memory
constexpr int kFlXTxBytes = kFlRows * kFlRowPitch * (int)sizeof(float);   // 22 304: what one tensor copy delivers
constexpr int kFlXBytes = (kFlXTxBytes + 127) & ~127;                 // 22 400
constexpr int kFlYPitch = kMaxMels + 4;                               // 84 floats per staged feature row
constexpr int kFlYBytes = kFlFrames * kFlYPitch * (int)sizeof(float); // 10 752
constexpr int kFlWarpBytes = 2 * kFlXBytes + kFlYBytes;               // 55 552
constexpr int kFlCtrlBytes = 128;                                     // 8 mbarriers, tensor-memory base, reduction scratch
constexpr int kFlTabBytes = 6144;                                     // FlTables staged in shared memory (static_assert below)
constexpr int kFlSmemBytes = kFlQuads * kFlWarpBytes + kFlCtrlBytes + kFlWarps * (int)sizeof(double2) + kFlTabBytes;
static_assert(kFlSmemBytes <= 232448, "frame-per-lane kernel: shared memory");
static_assert(kFlFrames == kWsFrames, "the tile grid (32 frames) is shared with the other kernels: workspace layout, statistics slots");

// Tensor-memory layout of one frame (32-bit columns of the thread's lane).  Stage 1 parks its rows at the top:
//   row k1 = 1..10 (20 complex values each) at column 512 - 40 k1, row 0 (20 reals) at column 92;
// stage 2 then builds the power spectrum P[k] at column k (k = 1..199; columns 200..215 are read padding of the widest
// mel class).  The two regions overlap (420 + 216 > 512): rows 0, 10, 9 and 8 sit inside P's columns, so stage 2 loads
// exactly those four rows into registers before it stores its first power value; rows 7..1 live above column 216.
constexpr int kFlTmemCols = 512;
constexpr int kFlColRow0 = 92;
__host__ __device__ constexpr int fl_row_col(int k1) { return kFlTmemCols - 40 * k1; }
static_assert(fl_row_col(10) == kFlColRow0 + 20 && fl_row_col(7) >= 216, "rows 7..1 must not overlap the power spectrum");

// Uniform tables: window taps, twiddles and mel weights are the same for every lane at every instruction.  They travel BY
// VALUE as a kernel parameter and are staged in shared memory once per CTA; every use is a broadcast LDS (one wavefront).
// (Reading them from the constant bank with a loop-dependent index — LDC R, c[0x0][R + imm] — measured 2 x slower: ptxas
// places each such load right in front of its consumer, so a single warp eats the full latency every time.)
// Mel filters are grouped as in the other kernels: class c = mels 20 c .. 20 c + 19 with common widths 2 / 4 / 7 / 13
// (is_reference_layout), weights zero-padded to the class's load width 2 / 4 / 8 / 16.
struct FlTables {
    float2 win2[10][20];          // [column pair p][m]: 0.5 hann[2p + 20 m], 0.5 hann[2p + 1 + 20 m]   (x input scale)
    float4 tw4[10][10];           // [column pair p][k1 - 1]: Re w_j, Re w_j+1, Im w_j, Im w_j+1 with w_j = 2 W400^(j k1), j = 2p
    int mel_lo[kMaxMels];         // first bin of mel m
    float w0[20][2];
    float w1[20][4];
    float w2[20][8];
    float w3[20][16];
};
static_assert(sizeof(FlTables) <= kFlTabBytes && sizeof(FlTables) % 16 == 0, "FlTables staging");

// -------------------------------------------------------------------------------------------- tensor memory helpers
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kFlTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kFlTmemCols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st2(unsigned addr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b));
}
__device__ __forceinline__ void tmem_st4(unsigned addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void tmem_ld8(unsigned addr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_ld4(unsigned addr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld16(unsigned addr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_ld32(unsigned addr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(addr));
}
// the values a tcgen05.ld delivers are defined only after tcgen05.wait::ld: tie every register to a point after the wait
template <int N> __device__ __forceinline__ void tmem_pin(float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+f"(v[i]));
}

__device__ __forceinline__ void tma_store_3d(const void* tmap, unsigned smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(smem_src),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned smem_dst, const void* tmap, int c0, int c1, int c2, unsigned long long* bar,
                                            unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
            smem_dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// packed helpers with a different constant in each half (column pairs)
__device__ __forceinline__ cf fl_mul2(cf a, cf b) {
    cf r;
    asm("{ .reg .b64 ra, rb, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rr, ra, rb; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ cf fl_fma2(cf a, cf b, cf c) {
    cf r;
    asm("{ .reg .b64 ra, rb, rc, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rr, ra, rb, rc; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ cf fl_neg(cf a) { return make_float2(-a.x, -a.y); }

// -------------------------------------------------------------------------------------------- tiles
struct FlTile {
    int row, t0, t_end, L;
    int flags;                     // kWsActive | kWsFull | kWsBulkX | kWsBulkY, same meaning as in the ws kernel
    int c1;                        // tensor-copy row coordinate: 32 * tile index inside the row
};

__device__ __forceinline__ FlTile fl_describe(const KernelArgs& a, int tile) {
    FlTile d;
    const int row = tile / a.tiles_per_row, tq = tile - row * a.tiles_per_row;
    d.row = row;
    d.t0 = a.frame0 + tq * kFlFrames;
    if (a.lens) {
        d.L = (int)min(a.lens[row], (long long)kMaxSamples);
        d.t_end = min(a.frame_end, d.L > kHalf ? 1 + d.L / kHop : 0);
    } else {
        d.L = a.total_len;
        d.t_end = a.t_end_const;
    }
    const bool active = d.t0 < d.t_end;
    const bool full = d.t0 + kFlFrames <= d.t_end;
    const int s0 = kHop * d.t0 - kHalf;
    const int b0 = s0 - a.origin;
    const bool interior = s0 >= 0 && s0 + kFlTileSamples <= d.L && b0 >= 0 && b0 + kFlSpan <= a.buf_len;
    const bool bulk_x = active && interior && a.use_tma;
    const bool bulk_y = active && full && a.out_layout == TALFE_LAYOUT_TM && a.use_tma_out;
    d.flags = (active ? kWsActive : 0) | (full ? kWsFull : 0) | (bulk_x ? kWsBulkX : 0) | (bulk_y ? kWsBulkY : 0);
    d.c1 = tq * kFlFrames;
    return d;
}

// -------------------------------------------------------------------------------------------- stage 1
// Twiddle column pair p's transform and park it in tensor memory.  re / im: halves = columns 2p, 2p + 1.
// Same roundings as stage1_ws_store (fma(re, w, im * (i w))).
__device__ __forceinline__ void fl_store_pair(const FlTables& T, const int p, const cf (&re)[11], const cf (&im)[11], const unsigned tm) {
    {
        const cf t0 = cadd(re[0], re[0]);                               // the transform runs at half scale (window x 0.5)
        tmem_st2(tm + kFlColRow0 + 2 * p, t0.x, t0.y);
    }
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const float4 tw = T.tw4[p][k1 - 1];
        const cf wr = make_float2(tw.x, tw.y), wi = make_float2(tw.z, tw.w);
        const cf t1 = fl_mul2(im[k1], wi), t2 = fl_mul2(im[k1], wr);
        cf o_re, o_im;
        if (rfft20_im_negated(k1)) {                                    // im holds -Im: A w = re w + im (-i w)
            o_re = fl_fma2(re[k1], wr, t1);
            o_im = fl_fma2(re[k1], wi, fl_neg(t2));
        } else {
            o_re = fl_fma2(re[k1], wr, fl_neg(t1));
            o_im = fl_fma2(re[k1], wi, t2);
        }
        tmem_st4(tm + fl_row_col(k1) + 4 * p, o_re.x, o_im.x, o_re.y, o_im.y);
    }
    {
        const float4 tw = T.tw4[p][9];
        const cf o_re = fl_mul2(re[10], make_float2(tw.x, tw.y)), o_im = fl_mul2(re[10], make_float2(tw.z, tw.w));
        tmem_st4(tm + fl_row_col(10) + 4 * p, o_re.x, o_im.x, o_re.y, o_im.y);
    }
}

// One frame: 400 samples -> window -> 20 real-input FFT-20 (two columns per packed register) -> twiddle -> tensor memory.
// xb: this lane's frame inside the skewed tile (16-byte aligned: lane pitch 164 floats).  A real loop (five passes of four
// columns): the whole kernel is a handful of small loops so that a single warp per scheduler runs out of the instruction
// cache — the fully unrolled first version (75 KB of straight-line code) spent 45 % of its time waiting for instructions.
__device__ __forceinline__ void fl_stage1(const FlTables& T, const float4* __restrict__ xb, const unsigned tm, const int p_lo, const int p_hi) {
    // column pairs p_lo .. p_hi - 1 of this frame (the two warps of a quadrant take five pairs each; the middle group of
    // four columns is loaded by both)
#pragma unroll 1
    for (int c4 = p_lo >> 1; 2 * c4 < p_hi; ++c4) {                     // columns 4 c4 .. 4 c4 + 3
        float4 q[20];
#pragma unroll
        for (int m = 0; m < 20; ++m) {
            // sample 20 m + 4 c4 of the frame; + 4 floats of padding per 160 samples (20 m + 4 c4 < 160 <=> m < 8: c4 < 5)
            q[m] = xb[5 * m + c4 + (20 * m) / kHop];
        }
        if (2 * c4 >= p_lo) {
            cf xin[20], re[11], im[11];
#pragma unroll
            for (int m = 0; m < 20; ++m) xin[m] = make_float2(q[m].x, q[m].y);
            rfft20_pair_windowed(xin, T.win2[2 * c4], re, im);
            fl_store_pair(T, 2 * c4, re, im, tm);
        }
        if (2 * c4 + 1 < p_hi) {
            cf xin[20], re[11], im[11];
#pragma unroll
            for (int m = 0; m < 20; ++m) xin[m] = make_float2(q[m].z, q[m].w);
            rfft20_pair_windowed(xin, T.win2[2 * c4 + 1], re, im);
            fl_store_pair(T, 2 * c4 + 1, re, im, tm);
        }
    }
}

// -------------------------------------------------------------------------------------------- stage 2
__device__ __forceinline__ void fl_load_row(const unsigned addr, float (&r)[40]) {
    tmem_ld32(addr, r);
    tmem_ld8(addr + 32, r + 32);
}
__device__ __forceinline__ void tmem_st1(unsigned addr, float a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "f"(a));
}

// Row k1 = 1..9 (in r: 20 complex values) -> FFT-20 -> |.|^2 -> P[k1 + 20 q] and P[(20 - k1) + 20 q], q = 0..9.
// v[q] = X[k1 + 20 q]; v[19 - q] = X[k1 + 20 (19 - q)] = conj X[(20 - k1) + 20 q]   (same pairing as stage2_ws_power_normal)
__device__ __forceinline__ void fl_row(const float (&r)[40], const unsigned tm, const int k1) {
    cf v[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) v[j] = make_float2(r[2 * j], r[2 * j + 1]);
    fft20<true>(v);
    const unsigned lo = tm + k1, hi = tm + 20 - k1;
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        tmem_st1(lo + 20 * q, fmaf(v[q].x, v[q].x, v[q].y * v[q].y));
        tmem_st1(hi + 20 * q, fmaf(v[19 - q].x, v[19 - q].x, v[19 - q].y * v[19 - q].y));
    }
}

// rows k_hi, k_hi - 1, ..., k_lo (all in 1..9); ra / rb hold rows k_hi and k_hi - 1 already.  Two register arrays in
// turn: the load of the row after next is in flight while the other array is transformed.
__device__ __forceinline__ void fl_rows(const unsigned tm, float (&ra)[40], float (&rb)[40], const int k_hi, const int k_lo) {
#pragma unroll 1
    for (int k1 = k_hi; k1 >= k_lo; k1 -= 2) {
        fl_row(ra, tm, k1);
        if (k1 - 2 >= k_lo) fl_load_row(tm + fl_row_col(k1 - 2), ra);
        if (k1 - 1 >= k_lo) {
            fl_row(rb, tm, k1 - 1);
            if (k1 - 3 >= k_lo) fl_load_row(tm + fl_row_col(k1 - 3), rb);
        }
        tmem_wait_ld();
        tmem_pin(ra); tmem_pin(rb);
    }
}

// The two warps of a quadrant share stage 2: half 0 takes rows 7..3, half 1 the four rows that live inside P's columns
// (0, 10, 9, 8) and rows 2, 1.  `pair_sync` separates half 1's loads of those four rows from everybody's first power store.
template <typename Sync>
__device__ __forceinline__ void fl_stage2(const unsigned tm, const int half, Sync pair_sync) {
    float ra[40], rb[40], z[20], r10[40];
    if (half == 0) {
        fl_load_row(tm + fl_row_col(7), ra);
        fl_load_row(tm + fl_row_col(6), rb);
    } else {
        tmem_ld16(tm + kFlColRow0, z);
        tmem_ld4(tm + kFlColRow0 + 16, z + 16);
        fl_load_row(tm + fl_row_col(10), r10);
        fl_load_row(tm + fl_row_col(9), ra);
        fl_load_row(tm + fl_row_col(8), rb);
    }
    tmem_wait_ld();
    tmem_pin(ra); tmem_pin(rb);
    pair_sync();                                                        // (one call site for both halves: B2)
    if (half == 0) {
        fl_rows(tm, ra, rb, 7, 3);
    } else {
        tmem_pin(z); tmem_pin(r10);
        {   // row 0 (real): bins 20 q, q = 1..9 (bins 0 and 200 carry no mel weight)
            cf v[20];
#pragma unroll
            for (int j = 0; j < 20; ++j) v[j] = make_float2(z[j], 0.f);
            fft20<true>(v);
#pragma unroll
            for (int q = 1; q < 10; ++q) tmem_st1(tm + 20 * q, fmaf(v[q].x, v[q].x, v[q].y * v[q].y));
        }
        {   // row 10: bins 10 + 20 q, q = 0..9 (q = 10..19 are their mirror images)
            cf v[20];
#pragma unroll
            for (int j = 0; j < 20; ++j) v[j] = make_float2(r10[2 * j], r10[2 * j + 1]);
            fft20<true>(v);
#pragma unroll
            for (int q = 0; q < 10; ++q) tmem_st1(tm + 10 + 20 * q, fmaf(v[q].x, v[q].x, v[q].y * v[q].y));
        }
        fl_rows(tm, ra, rb, 9, 8);
        fl_load_row(tm + fl_row_col(2), ra);
        fl_load_row(tm + fl_row_col(1), rb);
        tmem_wait_ld();
        tmem_pin(ra); tmem_pin(rb);
        fl_rows(tm, ra, rb, 2, 1);
    }
}

// -------------------------------------------------------------------------------------------- mel + log + store
template <int N> __device__ __forceinline__ void tmem_ldn(unsigned addr, float (&v)[N]) {
    if constexpr (N == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
    else if constexpr (N == 4) tmem_ld4(addr, v);
    else if constexpr (N == 8) tmem_ld8(addr, v);
    else tmem_ld16(addr, v);
}
template <int W, int N> __device__ __forceinline__ float fl_mel_dot(const float (&w)[N], const float (&p)[N]) {
    float acc = 0.f;                                                    // the accumulation order of mel_slot_ws
#pragma unroll
    for (int r = 0; r < W; ++r) acc = fmaf(w[r], p[r], acc);
    return acc;
}
// a mel's weights -> registers, as whole 8- / 16-byte loads issued back to back (their latency overlaps that of the tensor-
// memory loads of the same pass: a multiply-add that waits for its own weight load costs a single warp dearly)
template <int N> __device__ __forceinline__ void fl_load_w(const float (&src)[N], float (&w)[N]) {
    if constexpr (N == 2) {
        const float2 t = *reinterpret_cast<const float2*>(src);
        w[0] = t.x; w[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 t = reinterpret_cast<const float4*>(src)[i];
            w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
    }
}
// The 80 mels in five passes; the two warps of a quadrant split every pass by width class — half 0 the four widest
// filters (class 3: 4 x 13 products), half 1 the other twelve (classes 0..2: 4 x (2 + 4 + 7) products) — so both
// issue the same number of multiply-adds.  All loads of P[lo .. lo + N) of a pass are in flight together; weights are
// uniform operands.  emit(class, pass, y01, y23): four consecutive mels 20 class + 4 pass + (0..3).
template <typename Emit>
__device__ __forceinline__ void fl_mel_emit(const int c, const int g, const float (&acc)[4], const float eps, float& sum, float& sumsq,
                                            const bool want_sumsq, Emit& emit) {
    const cf y01 = fast_log2x(cadd(make_float2(acc[0], acc[1]), make_float2(eps, eps)));
    const cf y23 = fast_log2x(cadd(make_float2(acc[2], acc[3]), make_float2(eps, eps)));
    sum += (y01.x + y01.y) + (y23.x + y23.y);
    if (want_sumsq) sumsq = fmaf(y23.y, y23.y, fmaf(y23.x, y23.x, fmaf(y01.y, y01.y, fmaf(y01.x, y01.x, sumsq))));
    emit(c, g, y01, y23);
}
template <typename Emit>
__device__ __forceinline__ void fl_mel_log(const FlTables& T, const unsigned tm, const int half, const float eps, float& sum, float& sumsq,
                                           const bool want_sumsq, Emit emit) {
    if (half == 0) {
#pragma unroll 1
        for (int g = 0; g < 5; ++g) {
            float p3[4][16], w3[4][16];
#pragma unroll
            for (int u = 0; u < 4; ++u) tmem_ldn<16>(tm + (unsigned)T.mel_lo[60 + 4 * g + u], p3[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) fl_load_w(T.w3[4 * g + u], w3[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) tmem_pin(w3[u]);
            tmem_wait_ld();
            float acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                tmem_pin(p3[u]);
                acc[u] = fl_mel_dot<kRefW3>(w3[u], p3[u]);
            }
            fl_mel_emit(3, g, acc, eps, sum, sumsq, want_sumsq, emit);
        }
    } else {
#pragma unroll 1
        for (int g = 0; g < 5; ++g) {
            float p0[4][2], p1[4][4], p2[4][8], w0[4][2], w1[4][4], w2[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                tmem_ldn<2>(tm + (unsigned)T.mel_lo[4 * g + u], p0[u]);
                tmem_ldn<4>(tm + (unsigned)T.mel_lo[20 + 4 * g + u], p1[u]);
                tmem_ldn<8>(tm + (unsigned)T.mel_lo[40 + 4 * g + u], p2[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                fl_load_w(T.w0[4 * g + u], w0[u]);
                fl_load_w(T.w1[4 * g + u], w1[u]);
                fl_load_w(T.w2[4 * g + u], w2[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { tmem_pin(w0[u]); tmem_pin(w1[u]); tmem_pin(w2[u]); }
            tmem_wait_ld();
            float acc[3][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                tmem_pin(p0[u]); tmem_pin(p1[u]); tmem_pin(p2[u]);
                acc[0][u] = fl_mel_dot<kRefW0>(w0[u], p0[u]);
                acc[1][u] = fl_mel_dot<kRefW1>(w1[u], p1[u]);
                acc[2][u] = fl_mel_dot<kRefW2>(w2[u], p2[u]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) fl_mel_emit(c, g, acc[c], eps, sum, sumsq, want_sumsq, emit);
        }
    }
}

// -------------------------------------------------------------------------------------------- the kernel
// 8 warps: warp w works for quadrant w & 3 as its half w >> 2.  The two halves of a quadrant own the same 32 frames
// (thread = frame in both) and the same tensor-memory lanes; they split every stage of a tile and meet at four named
// barriers (64 threads) per tile.  Two warps per scheduler instead of one is what hides the dependent-issue latency of
// a stream that is ~60 % packed FP32 instructions.
__global__ void __launch_bounds__(kFlThreads, 1)
logmel_fl_kernel(const KernelArgs a, const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ FlTables T_param) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quad = warp & (kFlQuads - 1), half = warp / kFlQuads;
    unsigned char* s_quad = smem + quad * kFlWarpBytes;
    float* s_x0 = reinterpret_cast<float*>(s_quad);
    float* s_y = reinterpret_cast<float*>(s_quad + 2 * kFlXBytes);
    unsigned char* s_ctrl = smem + kFlQuads * kFlWarpBytes;
    unsigned long long* s_full = reinterpret_cast<unsigned long long*>(s_ctrl) + 2 * quad;   // [2] per quadrant
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_ctrl + 64);
    double2* s_red = reinterpret_cast<double2*>(s_ctrl + kFlCtrlBytes);
    const FlTables& T = *reinterpret_cast<const FlTables*>(s_ctrl + kFlCtrlBytes + kFlWarps * sizeof(double2));
    {
        const int4* src = reinterpret_cast<const int4*>(&T_param);
        int4* dst = reinterpret_cast<int4*>(s_ctrl + kFlCtrlBytes + kFlWarps * sizeof(double2));
        for (int i = threadIdx.x; i < (int)(sizeof(FlTables) / 16); i += kFlThreads) dst[i] = src[i];
    }

    if (half == 0 && lane == 0) { mbar_init(s_full, 1); mbar_init(s_full + 1, 1); }
    if (warp == 0) tmem_alloc(s_tmem);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tm = *s_tmem + ((unsigned)(32 * quad) << 16);        // this quadrant's 32 lanes
    cudaGridDependencySynchronize();

    // both halves of the quadrant: tensor-memory and shared-memory accesses before / after
    auto pair_sync = [&]() {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        named_bar_sync(1 + quad, 32 * kFlSplit);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    const int n_units = kFlQuads * (int)gridDim.x;
    const int wg = quad * (int)gridDim.x + (int)blockIdx.x;             // consecutive tiles go to different SMs
    const int n_my = wg < a.n_tiles ? (a.n_tiles - wg + n_units - 1) / n_units : 0;
    const unsigned long long policy = l2_evict_first_policy();
    const float4* xb0 = reinterpret_cast<const float4*>(s_x0) + lane * (kFlRowPitch / 4);
    const bool mt = a.out_layout == TALFE_LAYOUT_MT;

    auto fetch = [&](int k) {                                           // tile k of this quadrant -> x[k & 1] (one lane, one instruction)
        const FlTile d = fl_describe(a, wg + k * n_units);
        if ((d.flags & kWsBulkX) && half == 0 && lane == 0) {
            mbar_expect_tx(s_full + (k & 1), kFlXTxBytes);
            tma_load_3d(smem_u32(s_x0) + (k & 1) * kFlXBytes, &tmap_in, 0, d.c1, d.row, s_full + (k & 1), policy);
        }
    };
    if (n_my > 0) fetch(0);
    if (n_my > 1) fetch(1);

    double acc_s = 0.0, acc_q = 0.0;
    bool store_pending = false;
    unsigned phase = 0;                                                 // bit b: parity of the next completion of s_full[b]
#pragma unroll 1
    for (int k = 0; k < n_my; ++k) {
        const int buf = k & 1;
        const int tile = wg + k * n_units;
        const FlTile d = fl_describe(a, tile);
        const bool active = d.flags & kWsActive;
        float* s_x = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_x0) + buf * kFlXBytes);
        if (d.flags & kWsBulkX) {
            mbar_wait_sleep(s_full + buf, (phase >> buf) & 1);          // (only tiles that travel by tensor copy complete a phase)
            phase ^= 1u << buf;
        } else if (active) {
            // edge tile (reflection), chunk boundary or unaligned row: element-wise staging by the quadrant's 64 threads
            const int s0 = kHop * d.t0 - kHalf;
            const float* rowp = reinterpret_cast<const float*>(a.wave) + (long long)d.row * a.row_stride;
#pragma unroll 4
            for (int i = lane + 32 * half; i < kFlTileSamples; i += 32 * kFlSplit) {
                int g = s0 + i;
                if (g < 0) g = -g;                                      // reflect, no edge repeat
                if (g >= d.L) g = 2 * (d.L - 1) - g;
                const int bi = g - a.origin;
                float v = 0.f;
                if (g >= 0 && g < d.L && bi >= 0 && bi < a.buf_len) v = __ldg(rowp + bi);
                s_x[i + 4 * (i / kHop)] = v;
            }
            pair_sync();
        }
        if (active) {
            fl_stage1(T, reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(xb0) + buf * kFlXBytes), tm,
                      half == 0 ? 0 : 5, half == 0 ? 5 : 10);
            tmem_wait_st();
        }
        pair_sync();                                                    // B1: all 20 columns parked; every lane has read x[buf]
        if (k + 2 < n_my) {
            if (!(d.flags & kWsBulkX)) fence_proxy_async();            // generic writes of x[buf] before the copy engine's
            fetch(k + 2);
        }
        float sum = 0.f, sumsq = 0.f;
        if (active) {
            fl_stage2(tm, half, pair_sync);                             // (B2 inside)
            tmem_wait_st();
            if (store_pending) {                                        // the previous tile's store has finished reading Y
                if (lane == 0) bulk_wait_read<0>();
                store_pending = false;
            }
            pair_sync();                                                // B3: the power spectrum is complete, Y is free
            const bool mine = d.t0 + lane < d.t_end;
            float4* yrow = reinterpret_cast<float4*>(s_y) + lane * (kFlYPitch / 4);
            float* out_row = a.out + (a.out_offsets ? a.out_offsets[d.row] * kMaxMels : (long long)d.row * a.out_row_stride);
            float* dst_mt = out_row + (d.t0 - a.frame0) + lane;
            const bool mt_ok = d.t0 + lane < a.frame_end;
            const long long nfr_ll = a.n_frames;
            fl_mel_log(T, tm, half, a.eps, sum, sumsq, a.want_sumsq != 0, [=](int cls, int g, cf y01, cf y23) {
                const int m0 = 20 * cls + 4 * g;
                if (!mt) {
                    yrow[m0 / 4] = make_float4(y01.x, y01.y, y23.x, y23.y);
                } else if (mt_ok) {                                     // [.., 80, T]: lane = frame, coalesced straight from registers
                    dst_mt[(long long)m0 * nfr_ll] = mine ? y01.x : 0.f;
                    dst_mt[(long long)(m0 + 1) * nfr_ll] = mine ? y01.y : 0.f;
                    dst_mt[(long long)(m0 + 2) * nfr_ll] = mine ? y23.x : 0.f;
                    dst_mt[(long long)(m0 + 3) * nfr_ll] = mine ? y23.y : 0.f;
                }
            });
            if (!mine) { sum = 0.f; sumsq = 0.f; }
            if (!mt) fence_proxy_async();                               // this thread's Y writes -> visible to the copy engine
            pair_sync();                                                // B4: Y complete, P consumed
            if (!mt) {
                if (d.flags & kWsBulkY) {
                    if (half == 0 && lane == 0) {
                        tma_store_3d(&tmap_out, smem_u32(s_y), 0, d.t0 - a.frame0, d.row);
                        bulk_commit();
                    }
                    store_pending = half == 0;
                } else {
                    const int nfr = min(kFlFrames, a.frame_end - d.t0);
                    for (int i = lane + 32 * half; i < nfr * kMaxMels; i += 32 * kFlSplit) {
                        const int f = i / kMaxMels, m = i - f * kMaxMels;
                        const bool valid = d.t0 + f < d.t_end;
                        if (!valid && a.out_offsets) continue;
                        out_row[(long long)(d.t0 + f - a.frame0) * kMaxMels + m] = valid ? s_y[f * kFlYPitch + m] : 0.f;
                    }
                }
            }
        } else if (!a.out_offsets && half == 0) {
            // a tile beyond the row's own length (ragged batch, padded output): zero fill
            float* out_row = a.out + (long long)d.row * a.out_row_stride;
            const int nfr = min(kFlFrames, a.frame_end - d.t0);
            if (mt) {
                if (lane < nfr)
                    for (int m = 0; m < kMaxMels; ++m) out_row[(long long)m * a.n_frames + (d.t0 - a.frame0) + lane] = 0.f;
            } else {
                for (int i = lane; i < nfr * kMaxMels; i += 32) out_row[(long long)(d.t0 - a.frame0) * kMaxMels + i] = 0.f;
            }
        }
        if (a.partials_per_tile) {
            double ds = (double)sum, dq = (double)sumsq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ds += __shfl_xor_sync(0xffffffffu, ds, o);
                dq += __shfl_xor_sync(0xffffffffu, dq, o);
            }
            // 10 slots per tile: slots 0 / 1 carry the two halves' sums, half 0 zeroes the rest
            if (lane == 0) a.partials[(long long)tile * kWsRoleWarps + half] = make_double2(ds, dq);
            if (half == 0 && lane >= kFlSplit && lane < kWsRoleWarps) a.partials[(long long)tile * kWsRoleWarps + lane] = make_double2(0.0, 0.0);
        } else {
            acc_s += (double)sum;
            acc_q += (double)sumsq;
        }
    }
    if (store_pending && lane == 0) bulk_wait_all<0>();                 // shared memory must outlive the copies that read it
    if (!a.partials_per_tile) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
            acc_q += __shfl_xor_sync(0xffffffffu, acc_q, o);
        }
        if (lane == 0) s_red[warp] = make_double2(acc_s, acc_q);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (!a.partials_per_tile && threadIdx.x == 0) {
        double ts = 0.0, tq2 = 0.0;
        for (int w2 = 0; w2 < kFlWarps; ++w2) { ts += s_red[w2].x; tq2 += s_red[w2].y; }
        a.partials[blockIdx.x] = make_double2(ts, tq2);
    }
    if (warp == 0) tmem_dealloc(*s_tmem);
}

}  // namespace
