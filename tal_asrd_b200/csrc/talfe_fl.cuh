// talfe_fl.cuh — "frame per lane" version of K1 (included by talfe.cu after talfe_ws.cuh; fp32 waveforms, reference filterbank).
//
// Why a third formulation.  The warp-specialised kernel (talfe_ws.cuh) moves every sample through shared memory four times
// (tile, exchange write + read, power write + 2.6 x read) and its skeleton — every LDS/STS and hand-off, no arithmetic —
// already takes 53 us on 64 x 30 s: the shared-memory pipe (one 128-byte wavefront per clock and SM), not the FP32 pipe,
// bounds it.  Here ONE THREAD OWNS ONE FRAME from waveform to log-mel, and the only thing that still
// crosses threads' register files is nothing at all: the 20 x 20 transposition between the two FFT stages happens inside
// the thread, through its own lane of TENSOR MEMORY used as a 420-word scratch pad (tcgen05.st after stage 1, tcgen05.ld
// before stage 2; 512 columns x 4 bytes per lane, its own datapath, no bank conflicts, no barriers, no mbarriers).
//   * shared memory carries the waveform tile (100 LDS.128 per frame: 12.5 wavefronts) and the staged feature rows
//     (20 STS.128: 2.5 wavefronts) — 15 wavefronts per frame instead of 75;
//   * window taps, twiddles and mel weights are the same for every lane at every instruction, so they are uniform
//     operands read from the kernel's parameter space (constant bank -> uniform registers): no table lives in
//     shared memory or in vector registers;
//   * the power spectrum never leaves the registers: each bin is folded into its (at most three) mel accumulators as
//     soon as it exists, with compile-time indices (80 accumulators per thread);
//   * a warp is a complete, independent pipeline (its own tile queue, its own double-buffered waveform tile fetched by
//     one tensor copy, its own feature tile leaving by one tensor store): there is no inter-warp synchronisation at all.
// Tensor memory holds one frame per lane, so an SM runs 4 warps x 32 frames at a time (one warp per scheduler,
// up to 255 registers per thread); latency is hidden by instruction-level parallelism inside the thread (20 independent
// column transforms, 11 independent row transforms), not by switching warps.
//
// Arithmetic: the same real-structured 20 x 20 Cooley-Tukey split as the other kernels (talfe_core.cuh), with the packed
// f32x2 halves carrying two adjacent COLUMNS of one frame in stage 1 (instead of two frames of one column) and the two
// components of a complex number in stage 2.  Rows k1 = 1..9 are bit-identical to the other kernels up to the order in
// which a mel filter's terms are added.
#pragma once

namespace {

constexpr int kFlWarps = 4;
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
constexpr int kFlSmemBytes = kFlWarps * kFlWarpBytes + kFlCtrlBytes + 4 * (int)sizeof(double2);
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

// Uniform tables, passed BY VALUE as a kernel parameter (about 6 KB of the constant bank): window taps, twiddles and mel
// weights are the same for every lane at every instruction, so each use is a load into a uniform register (LDCU).
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
__device__ __forceinline__ void fl_stage1(const FlTables& T, const float4* __restrict__ xb, const unsigned tm) {
#pragma unroll 1
    for (int c4 = 0; c4 < 5; ++c4) {                                    // columns 4 c4 .. 4 c4 + 3
        float4 q[20];
#pragma unroll
        for (int m = 0; m < 20; ++m) {
            // sample 20 m + 4 c4 of the frame; + 4 floats of padding per 160 samples (20 m + 4 c4 < 160 <=> m < 8: c4 < 5)
            q[m] = xb[5 * m + c4 + (20 * m) / kHop];
        }
        {
            cf xin[20], re[11], im[11];
#pragma unroll
            for (int m = 0; m < 20; ++m) xin[m] = make_float2(q[m].x, q[m].y);
            rfft20_pair_windowed(xin, T.win2[2 * c4], re, im);
            fl_store_pair(T, 2 * c4, re, im, tm);
        }
        {
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

__device__ __forceinline__ void fl_stage2(const unsigned tm) {
    float ra[40], rb[40];
    {
        // the four rows that live inside P's columns: all in registers before the first power value is stored
        float z[20], r10[40];
        tmem_ld16(tm + kFlColRow0, z);
        tmem_ld4(tm + kFlColRow0 + 16, z + 16);
        fl_load_row(tm + fl_row_col(10), r10);
        fl_load_row(tm + fl_row_col(9), ra);
        fl_load_row(tm + fl_row_col(8), rb);
        tmem_wait_ld();
        tmem_pin(z); tmem_pin(r10); tmem_pin(ra); tmem_pin(rb);
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
    }
    // rows 9, 8 | 7, 6 | 5, 4 | 3, 2 | 1: two register arrays in turn; the load of the row after next is in flight
    // while the other array is transformed
#pragma unroll 1
    for (int k1 = 9; k1 >= 1; k1 -= 2) {
        fl_row(ra, tm, k1);
        if (k1 >= 3) fl_load_row(tm + fl_row_col(k1 - 2), ra);
        if (k1 >= 2) {
            fl_row(rb, tm, k1 - 1);
            if (k1 >= 4) fl_load_row(tm + fl_row_col(k1 - 3), rb);
        }
        tmem_wait_ld();
        tmem_pin(ra); tmem_pin(rb);
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
// All 80 mels in five passes of 16 (four of each width class): 16 loads of P[lo .. lo + N) from tensor memory in flight
// together, 16 independent accumulation chains, weights as uniform operands.  emit(class, pass, y01, y23): four
// consecutive mels 20 class + 4 pass + (0..3).
template <typename Emit>
__device__ __forceinline__ void fl_mel_log(const FlTables& T, const unsigned tm, const float eps, float& sum, float& sumsq,
                                           const bool want_sumsq, Emit emit) {
#pragma unroll 1
    for (int g = 0; g < 5; ++g) {
        float p0[4][2], p1[4][4], p2[4][8], p3[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tmem_ldn<2>(tm + (unsigned)T.mel_lo[4 * g + u], p0[u]);
            tmem_ldn<4>(tm + (unsigned)T.mel_lo[20 + 4 * g + u], p1[u]);
            tmem_ldn<8>(tm + (unsigned)T.mel_lo[40 + 4 * g + u], p2[u]);
            tmem_ldn<16>(tm + (unsigned)T.mel_lo[60 + 4 * g + u], p3[u]);
        }
        tmem_wait_ld();
        float acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tmem_pin(p0[u]); tmem_pin(p1[u]); tmem_pin(p2[u]); tmem_pin(p3[u]);
            acc[0][u] = fl_mel_dot<kRefW0>(T.w0[4 * g + u], p0[u]);
            acc[1][u] = fl_mel_dot<kRefW1>(T.w1[4 * g + u], p1[u]);
            acc[2][u] = fl_mel_dot<kRefW2>(T.w2[4 * g + u], p2[u]);
            acc[3][u] = fl_mel_dot<kRefW3>(T.w3[4 * g + u], p3[u]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const cf y01 = fast_log2x(cadd(make_float2(acc[c][0], acc[c][1]), make_float2(eps, eps)));
            const cf y23 = fast_log2x(cadd(make_float2(acc[c][2], acc[c][3]), make_float2(eps, eps)));
            sum += (y01.x + y01.y) + (y23.x + y23.y);
            if (want_sumsq) sumsq = fmaf(y23.y, y23.y, fmaf(y23.x, y23.x, fmaf(y01.y, y01.y, fmaf(y01.x, y01.x, sumsq))));
            emit(c, g, y01, y23);
        }
    }
}

// -------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kFlThreads, 1)
logmel_fl_kernel(const KernelArgs a, const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ FlTables T) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* s_warp = smem + warp * kFlWarpBytes;
    float* s_x0 = reinterpret_cast<float*>(s_warp);
    float* s_y = reinterpret_cast<float*>(s_warp + 2 * kFlXBytes);
    unsigned char* s_ctrl = smem + kFlWarps * kFlWarpBytes;
    unsigned long long* s_full = reinterpret_cast<unsigned long long*>(s_ctrl) + 2 * warp;   // [2] per warp
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_ctrl + 64);
    double2* s_red = reinterpret_cast<double2*>(s_ctrl + kFlCtrlBytes);

    if (lane == 0) { mbar_init(s_full, 1); mbar_init(s_full + 1, 1); }
    if (warp == 0) tmem_alloc(s_tmem);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this warp's 32 lanes; the shuffle tells the compiler that the value is warp-uniform, so that every tcgen05 address
    // is uniform-register + immediate instead of one R2UR per access
    const unsigned tm = __shfl_sync(0xffffffffu, *s_tmem + ((unsigned)(32 * warp) << 16), 0);
    cudaGridDependencySynchronize();

    const int n_warps = kFlWarps * (int)gridDim.x;
    const int wg = warp * (int)gridDim.x + (int)blockIdx.x;             // consecutive tiles go to different SMs
    const int n_my = wg < a.n_tiles ? (a.n_tiles - wg + n_warps - 1) / n_warps : 0;
    const unsigned long long policy = l2_evict_first_policy();
    const float4* xb0 = reinterpret_cast<const float4*>(s_x0) + lane * (kFlRowPitch / 4);
    const bool mt = a.out_layout == TALFE_LAYOUT_MT;

    auto fetch = [&](int k) {                                           // tile k of this warp -> x[k & 1] (one lane, one instruction)
        const FlTile d = fl_describe(a, wg + k * n_warps);
        if ((d.flags & kWsBulkX) && lane == 0) {
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
        const int tile = wg + k * n_warps;
        const FlTile d = fl_describe(a, tile);
        const bool active = d.flags & kWsActive;
        float* s_x = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_x0) + buf * kFlXBytes);
        if (d.flags & kWsBulkX) {
            mbar_wait_sleep(s_full + buf, (phase >> buf) & 1);          // (only tiles that travel by tensor copy complete a phase)
            phase ^= 1u << buf;
        } else if (active) {
            // edge tile (reflection), chunk boundary or unaligned row: element-wise staging by the warp itself
            const int s0 = kHop * d.t0 - kHalf;
            const float* rowp = reinterpret_cast<const float*>(a.wave) + (long long)d.row * a.row_stride;
            for (int i = lane; i < kFlTileSamples; i += 32) {
                int g = s0 + i;
                if (g < 0) g = -g;                                      // reflect, no edge repeat
                if (g >= d.L) g = 2 * (d.L - 1) - g;
                const int bi = g - a.origin;
                float v = 0.f;
                if (g >= 0 && g < d.L && bi >= 0 && bi < a.buf_len) v = __ldg(rowp + bi);
                s_x[i + 4 * (i / kHop)] = v;
            }
            __syncwarp();
        }
        if (active) {
            fl_stage1(T, reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(xb0) + buf * kFlXBytes), tm);
            tmem_wait_st();
        }
        __syncwarp();                                                   // every lane has read x[buf]
        if (k + 2 < n_my) {
            if (!(d.flags & kWsBulkX)) fence_proxy_async();            // generic writes of x[buf] before the copy engine's
            fetch(k + 2);
        }
        float sum = 0.f, sumsq = 0.f;
        if (active) {
            fl_stage2(tm);
            tmem_wait_st();
            const bool mine = d.t0 + lane < d.t_end;
            if (store_pending) {                                        // the previous tile's store has finished reading Y
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                store_pending = false;
            }
            float4* yrow = reinterpret_cast<float4*>(s_y) + lane * (kFlYPitch / 4);
            float* out_row = a.out + (a.out_offsets ? a.out_offsets[d.row] * kMaxMels : (long long)d.row * a.out_row_stride);
            float* dst_mt = out_row + (d.t0 - a.frame0) + lane;
            const bool mt_ok = d.t0 + lane < a.frame_end;
            const long long nfr_ll = a.n_frames;
            fl_mel_log(T, tm, a.eps, sum, sumsq, a.want_sumsq != 0, [=](int cls, int g, cf y01, cf y23) {
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
            if (!mt) {
                if (d.flags & kWsBulkY) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tmap_out, smem_u32(s_y), 0, d.t0 - a.frame0, d.row);
                        bulk_commit();
                    }
                    store_pending = true;
                } else {
                    __syncwarp();
                    const int nfr = min(kFlFrames, a.frame_end - d.t0);
                    for (int i = lane; i < nfr * kMaxMels; i += 32) {
                        const int f = i / kMaxMels, m = i - f * kMaxMels;
                        const bool valid = d.t0 + f < d.t_end;
                        if (!valid && a.out_offsets) continue;
                        out_row[(long long)(d.t0 + f - a.frame0) * kMaxMels + m] = valid ? s_y[f * kFlYPitch + m] : 0.f;
                    }
                    __syncwarp();
                }
            }
        } else if (!a.out_offsets) {
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
            if (lane < kWsRoleWarps) a.partials[(long long)tile * kWsRoleWarps + lane] = lane == 0 ? make_double2(ds, dq) : make_double2(0.0, 0.0);
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
