// talfe_ws.cuh — warp-specialised version of K1 (included by talfe.cu after its helpers).
//
// One persistent CTA per SM, 640 threads:
//   warps  0..9   PRODUCERS  waveform tile -> window -> FFT-20 -> untangle/twiddle -> exchange E; in turn (tile k by
//                            warp k % 10) also the loader: tile descriptor, TMA fetch of the waveform tile one tile
//                            ahead (double-buffered), L2 prefetch two tiles ahead
//   warps 10..19  CONSUMERS  exchange E -> FFT-20 -> power P -> mel projection -> log -> staged feature tile Y
//                            -> cp.async.bulk store to global memory (640 bytes per frame pair); TWO groups of 5 warps
//                            on alternate tiles (group q owns E[q], P[q], Y[q]), two exchange rows per thread
// (a 21st warp for the loader would cost 16 registers per thread: 6 warps on one scheduler's 16K-register file)
// Producers keep their 20 window taps in registers; twiddles (5 LDS.128 per tile) and mel weights (7 LDS.128 per
// tile and mel lane) are re-read from shared memory, which measured faster than holding them in registers (B200,
// profiles/r01_ab_v7_ws_vs_legacy.json, profiles/r01_ab_v9_variants.json: register pressure in the FFT costs more
// than the loads).  The roles meet through mbarriers (x_full / x_empty / e_full / e_empty, two buffers each); each
// consumer group synchronises within itself with two named barriers per tile (ids 1 / 3, 160 threads); there is no
// CTA-wide barrier inside the tile loop.  The FP-heavy and the shared-memory-heavy phases of the three concurrent
// streams (producers, consumer group 0, consumer group 1) drift apart, so the SM's schedulers overlap them warp by
// warp; overlapping them INSIDE a warp (one fused instruction stream) measured slower, DESIGN.md §4.
//
// Shared-memory layouts and their bank-conflict properties: talfe_core.cuh ("Warp-specialised path").
// Arithmetic is identical to the legacy kernel (same functions for the FFT, untangle, power, mel and log).
#pragma once

namespace {

constexpr int kWsRoleThreads = kWsGroups * kGroup;          // 320
constexpr int kWsRoleWarps = kWsRoleThreads / 32;           // 10
constexpr int kWsThreads = 2 * kWsRoleThreads;              // 640: 10 producer warps, 10 consumer warps (5 per scheduler: 96 registers each)
// Other role / register splits that were built and measured against this one (all bit-identical; code in the history at the
// commits named in DESIGN.md §4, results under profiles/):
//   * a helper warpgroup behind `setmaxnreg` (768 threads launched at 80 registers, compute warps raised to 88, helpers at
//     40; a loader warp and a storer warp): the loader warp saves 2.6 us, the storer 0.4 us, 88 registers instead of 96 cost
//     5.0 us -> 79.9 us (r02_ab_helper_warpgroup.json);
//   * five producer warps with two frame pairs per thread (15 warps x 128 registers): 94.5 us (r02_ab_p2_five_producer_warps.json);
//   * the loader duty on a consumer warp, after or before its wait for the exchange buffer: 78.0 / 78.4 us
//     (r02_ab_consumer_loads*.json);
//   * one mbarrier arrival per phase through a shared-memory counter: 80.7 us (r02_ab_single_arrive.json); only the loader
//     waiting for a free exchange buffer: 77.9 us (r02_ab_loader_waits_e.json); an L2 evict_last hint on the feature stores:
//     no gain for the sweep that follows (76.1 / 89.2 us against 75.7 / 89.1).
constexpr int kWsTileSamples = kHop * kWsFrames + (kNfft - kHop);   // 5360
// fp32 tiles travel as ONE tensor copy (cp.async.bulk.tensor, SASS UTMALDG): the waveform is described to the copy engine
// as rows of 340 samples that start every 320 samples (a 4-D tensor map [68][5][rows][batch] with strides 272 B, 1 280 B,
// row pitch: the rows overlap by 20 samples), so a box of 17 rows lands in shared memory with exactly the 20-float skew
// per 320 samples that the conflict-free stage-1 reads need — what used to take 17 cp.async.bulk pieces, each of which
// costs its issuing warp ~275 cycles (profiles/r02_timeline.json).
constexpr int kWsTmaInner = 68, kWsTmaMid = 5, kWsTmaRows = 17;
constexpr int kWsTmaRowFloats = kWsTmaInner * kWsTmaMid;             // 340 = kXBlock + fp32 skew
constexpr int kWsTmaSpan = kXBlock * (kWsTmaRows - 1) + kWsTmaRowFloats;   // 5 460 samples touched in global memory
constexpr int kWsTmaBytes = kWsTmaRows * kWsTmaRowFloats * (int)sizeof(float);   // 23 120
constexpr int kWsXBufBytes = (kWsTmaBytes + 127) & ~127;              // 23 168: every x buffer starts 128-byte aligned
static_assert(kWsTmaRowFloats == XLayout<float>::kGroup && kXBlock * (kWsTmaRows - 1) + 240 == kWsTileSamples, "tensor box = skewed tile");
static_assert(kWsXBufBytes >= kXFloats * (int)sizeof(float), "x buffer holds the skewed tile of every element type");
static_assert(kWsTileSamples == kTileSamples && kWsRoleWarps == kWarps, "tile geometry is shared with the legacy kernel");

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, unsigned smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(unsigned smem_dst, const void* tmap, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar, unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(
            smem_dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// -DTALFE_TIMELINE: development build that records, for CTA 0, the SM clock at every phase boundary of every warp and
// tile into KernelArgs::timeline ([warp 20][tile 64][8] uint32): where a latency-bound pipeline loses its time cannot
// be read off aggregate counters (tools/timeline.py).  Never defined in the product build.
#ifdef TALFE_TIMELINE
#define TL_MARK(warp_, k_, slot_) do { if (a.timeline && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (k_) < 64) \
    a.timeline[(((warp_) * 64) + (k_)) * 8 + (slot_)] = (unsigned)clock64(); } while (0)
#else
#define TL_MARK(warp_, k_, slot_) do { } while (0)
#endif
#ifndef TALFE_WS_WAIT_HINT_NS
#define TALFE_WS_WAIT_HINT_NS 1000000
#endif
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of spinning on issue slots
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)TALFE_WS_WAIT_HINT_NS)
            : "memory");
    }
}

// The producers' wait for "E[buf] free" polls with mbarrier.test_wait and a FIXED nanosleep between polls instead of
// try_wait's event-driven suspension.  The ncu source view of the try_wait form showed that loop running 8.6 iterations
// per frame — 28 per producer warp and tile, two SYNCS operations each through the shared-memory instruction queue: a
// suspended warp is woken by every mbarrier event of the CTA (the tile copy's transaction updates included), not only by
// the arrivals it waits for.  Measured (profiles/r02_ab_wait_backoff.json, bit-identical): 76.4 us with try_wait, 75.5-75.9
// with 20 / 50 / 100 ns (100 shipped), 76.0 with 200 ns; the same treatment of the producers' wait for the tile (TALFE_WS_XFULL_NS) and of the consumers' wait for "E full" (TALFE_WS_EFULL_NS, 6
// iterations per warp and tile) changes nothing.  0 = try_wait with the suspend-time hint.
#ifndef TALFE_WS_EEMPTY_NS
#define TALFE_WS_EEMPTY_NS 100
#endif
#ifndef TALFE_WS_EFULL_NS
#define TALFE_WS_EFULL_NS 0
#endif
#ifndef TALFE_WS_XFULL_NS
#define TALFE_WS_XFULL_NS 0
#endif
template <int kNs>
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
    if constexpr (kNs > 0) {
        for (;;) {
            unsigned done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
            if (done) break;
            asm volatile("nanosleep.u32 %0;" ::"n"(kNs));
        }
    } else {
        mbar_wait_sleep(bar, parity);
    }
}

// Tile descriptor: written once per tile by the loader warp, read by the two compute roles (one LDS.128 + one LDS.64)
// so that no compute warp spends instructions on tile bookkeeping.  Ring of 8: the loader is at most two tiles
// ahead of the producers (x double buffer), the producers at most three ahead of the consumers' mel stage.
struct __align__(16) WsDesc {
    int t0, t_end, L, row;
    int flags;                     // bit 0 active, bit 1 full, bit 2 x tile fetched by the copy engine, bit 3 bulk store allowed
    int pad;                       // tensor copy: first 320-sample row of the tile (16 * tile index in the row)
    float* out_tile;               // &out[row][t0 - frame0][0] for [.., T, 80] outputs
    const void* src;               // first sample of the tile in global memory (bulk tiles)
    long long pad2;
};
constexpr int kWsDescRing = 16;
enum { kWsActive = 1, kWsFull = 2, kWsBulkX = 4, kWsBulkY = 8 };

__device__ __forceinline__ WsDesc ws_describe(const KernelArgs& a, int row, int tq, long long& src_off) {
    WsDesc d;
    d.t0 = a.frame0 + tq * kWsFrames;
    if (a.lens) {
        d.L = (int)min(a.lens[row], (long long)kMaxSamples);
        d.t_end = min(a.frame_end, d.L > kHalf ? 1 + d.L / kHop : 0);
    } else {
        d.L = a.total_len;
        d.t_end = a.t_end_const;
    }
    d.row = row;
    const bool active = d.t0 < d.t_end;
    const bool full = d.t0 + kWsFrames <= d.t_end;
    const int s0 = kHop * d.t0 - kHalf;
    const int b0 = s0 - a.origin;
    // (the tensor copy reads whole 340-sample rows: up to 100 samples past the tile, which must still be inside the buffer)
    const bool interior = s0 >= 0 && s0 + kWsTileSamples <= d.L && b0 >= 0 && b0 + (a.use_tma ? kWsTmaSpan : kWsTileSamples) <= a.buf_len;
    const bool bulk_x = active && interior && a.align_ok;
    const bool bulk_y = active && full && a.out_layout == TALFE_LAYOUT_TM && a.out_align_ok;
    d.flags = (active ? kWsActive : 0) | (full ? kWsFull : 0) | (bulk_x ? kWsBulkX : 0) | (bulk_y ? kWsBulkY : 0);
    d.pad = tq * (kWsTmaRows - 1);
    d.out_tile = a.out + (a.out_offsets ? a.out_offsets[row] * kMaxMels : (long long)row * a.out_row_stride) +
                 (long long)(d.t0 - a.frame0) * kMaxMels;
    src_off = (long long)row * a.row_stride + b0;
    d.src = nullptr;
    d.pad2 = 0;
    return d;
}

__device__ __forceinline__ void ws_advance(const KernelArgs& a, int& row, int& tq, int step) {
    tq += step;
    while (tq >= a.tiles_per_row) { tq -= a.tiles_per_row; ++row; }
}

// Tile index of this CTA's sequence -> (row, tile inside the row).  Packed ragged output (talfe_job::out_offsets): the
// tile list is COMPACT — only the tiles that hold frames of their row, enumerated by tile_map_kernel — so that a batch of
// 1 s .. 10 min utterances costs its own frames and not 64 x the longest row's worth of empty hand-offs.
template <bool kCompact>
__device__ __forceinline__ void ws_tile_coords(const KernelArgs& a, const int tile, int& row, int& tq) {
    if (kCompact) {
        const int2 rq = __ldg(a.tile_map + tile);
        row = rq.x; tq = rq.y;
    } else {
        row = tile / a.tiles_per_row;
        tq = tile - row * a.tiles_per_row;
    }
}

// ------------------------------------------------------------------------------------------ producers
// (ONE group of 10 warps, one frame pair per thread.  Splitting the producers into two groups on alternate tiles like
// the consumers — two pairs per thread, x[q] / E[q] per group — measured 91.7 us against 80.2 us: each group then holds
// its exchange buffer for two pairs' worth of work and the consumers wait for it.)
template <typename XT, bool kCompact>
__device__ __forceinline__ void ws_producer(const KernelArgs& a, const void* tmap, unsigned char* smem, XT* s_x0, cf* s_e0, WsDesc* s_desc,
                                            unsigned long long* s_bar, const int tid, const int n_my) {
    unsigned long long* x_full = s_bar;            // [2]
    unsigned long long* x_empty = s_bar + 2;       // [2]
    unsigned long long* e_full = s_bar + 4;        // [2]
    unsigned long long* e_empty = s_bar + 6;       // [2]
    const int warp = tid >> 5, lane = tid & 31;
    const int g1 = tid / kGroup, j = tid - g1 * kGroup;
    constexpr int kXG = XLayout<XT>::kGroup;
    constexpr int kXBufBytes = kWsXBufBytes;                            // every element size uses the fp32-sized buffer

    float win[20];
    load_window(j, a.win_global, XLayout<XT>::kScale, win);            // once per CTA, straight from global memory
    const cf* s_tw = reinterpret_cast<const cf*>(smem + a.off_tw) + j * 10;
    cf tw[10];
    const XT* xg = s_x0 + kXG * g1 + j;
    cf* col0 = s_e0 + ws_e_base(g1) + j;
    // Loader duty, taken in turn by the producer warps (tile kk by warp kk % 10, one tile ahead of the FFTs): all 32
    // lanes run it (no divergent waiting): descriptor, wait for x[kk & 1] to be free, byte count, the 17 pieces of the
    // tile fetch (lane i = piece i, one per 320-sample skew block), then an L2 prefetch of the tile after next.
    // (Measured and rejected in round 2, profiles/r02_ab_variants.json: issuing the fetch of tile k+2 from whichever
    // producer warp is the last to have read tile k — almost two tile times of lead instead of one — needs the pair's 28
    // samples or its transform live across the issue and spills: 85 us / 120 us against 76 us.)
    const int step = (int)gridDim.x;
    auto describe = [&](int kk) {
        const int tile = (int)blockIdx.x + kk * step;
        int row, tq;
        ws_tile_coords<kCompact>(a, tile, row, tq);
        long long src_off;
        WsDesc d = ws_describe(a, row, tq, src_off);
        d.src = reinterpret_cast<const XT*>(a.wave) + src_off;
        if (lane == 0) s_desc[kk & (kWsDescRing - 1)] = d;
        __syncwarp();
    };
    auto issue = [&](int kk) {
        const int lbuf = kk & 1;
        const WsDesc* dp = s_desc + (kk & (kWsDescRing - 1));
        const int flags = dp->flags;
#if defined(TALFE_ABLATE) && (TALFE_ABLATE & 16)
        const int fetch = 0;                                             // timing experiment only: no waveform fetch
#else
        const int fetch = flags & kWsBulkX;
#endif
        if (fetch && sizeof(XT) == 4 && a.use_tma) {
            if (lane == 0) {
                mbar_expect_tx(x_full + lbuf, kWsTmaBytes);                                // release: publishes the descriptor too
                tma_load_4d(smem_u32(s_x0) + lbuf * kXBufBytes, tmap, 0, 0, dp->pad, dp->row, x_full + lbuf, l2_evict_first_policy());
            }
        } else if (fetch) {
            const XT* src = reinterpret_cast<const XT*>(dp->src);
            if (lane == 0) mbar_expect_tx(x_full + lbuf, kWsTileSamples * (int)sizeof(XT));   // release: publishes the descriptor too
            __syncwarp();
            if (lane * kXBlock < kWsTileSamples)
                bulk_g2s_u32(smem_u32(s_x0 + lane * kXG) + lbuf * kXBufBytes, src + lane * kXBlock,
                             (unsigned)(min(kXBlock, kWsTileSamples - lane * kXBlock) * (int)sizeof(XT)), x_full + lbuf,
                             l2_evict_first_policy());
        } else if (lane == 0) {
            mbar_arrive(x_full + lbuf);                                                  // nothing in flight: descriptor only
        }
        __syncwarp();
    };
    auto load_duty = [&](int kk) {
        const int tile = (int)blockIdx.x + kk * step;
        const int row = tile / a.tiles_per_row, tq = tile - row * a.tiles_per_row;
        describe(kk);
        if (kk >= 2) mbar_wait_sleep(x_empty + (kk & 1), ((kk - 2) >> 1) & 1);           // tile kk-2 has left x[kk & 1]
        issue(kk);
        if (a.l2_prefetch && lane == 0 && tile + 2 * step < a.n_tiles) {                 // tile kk+2: HBM -> L2
            int r2 = row, q2 = tq;
            ws_advance(a, r2, q2, 2 * step);
            long long off2;
            const WsDesc d2 = ws_describe(a, r2, q2, off2);
            if (d2.flags & kWsBulkX) bulk_prefetch_l2(reinterpret_cast<const XT*>(a.wave) + off2, kWsTileSamples * (int)sizeof(XT));
        }
        __syncwarp();
    };
    if (warp == 0 && (!kCompact || n_my > 0)) load_duty(0);
    for (int k = 0; k < n_my; ++k) {
        const int buf = k & 1;
        if (k + 1 < n_my && (k + 1) % kWsRoleWarps == warp) load_duty(k + 1);
        TL_MARK(warp, k, 0);
        mbar_wait_backoff<TALFE_WS_XFULL_NS>(x_full + buf, (k >> 1) & 1);   // descriptor published, bulk tile landed
        TL_MARK(warp, k, 1);
        const int flags = s_desc[k & (kWsDescRing - 1)].flags;
        const bool active = flags & kWsActive;
        cf re[11], im[11];
        if (active) {
            if (!(flags & kWsBulkX)) {
                // edge tile (reflection), unaligned row or chunk boundary: element-wise staging by all producers
                // (x[buf] is free: the loader waited for x_empty before it arrived on x_full)
                const WsDesc d = s_desc[k & (kWsDescRing - 1)];
                XT* s_x = reinterpret_cast<XT*>(reinterpret_cast<unsigned char*>(s_x0) + buf * kXBufBytes);
                const int s0 = kHop * d.t0 - kHalf;
                const XT* rowp = reinterpret_cast<const XT*>(a.wave) + (long long)d.row * a.row_stride;
                for (int i = tid; i < kWsTileSamples; i += kWsRoleThreads) {
                    int g = s0 + i;
                    if (g < 0) g = -g;                                  // reflect, no edge repeat
                    if (g >= d.L) g = 2 * (d.L - 1) - g;
                    const int bi = g - a.origin;
                    XT v = XT(0.f);
                    if (g >= 0 && g < d.L && bi >= 0 && bi < a.buf_len) v = __ldg(rowp + bi);
                    s_x[xskew<XT>(i)] = v;
                }
                fence_proxy_async();                                    // these generic writes before the next tensor copy into x[buf]
                named_bar_sync(2, kWsRoleThreads);
            }
            stage1_ws_fft<XT>(reinterpret_cast<const XT*>(reinterpret_cast<const unsigned char*>(xg) + buf * kXBufBytes), win, re, im);
        }
        __syncwarp();
        TL_MARK(warp, k, 2);
        if (lane == 0) mbar_arrive(x_empty + buf);   // this warp no longer reads x[buf]
        __syncwarp();
        // every tile, active or not: a producer never runs more than one phase ahead of the consumers
        if (k >= 2) mbar_wait_backoff<TALFE_WS_EEMPTY_NS>(e_empty + buf, ((k - 2) >> 1) & 1); // consumers have loaded E[buf] of tile k-2
        TL_MARK(warp, k, 3);
        if (active) {
#pragma unroll
            for (int h = 0; h < 5; ++h) {
                const float4 tt = reinterpret_cast<const float4*>(s_tw)[h];
                tw[2 * h] = make_float2(tt.x, tt.y);
                tw[2 * h + 1] = make_float2(tt.z, tt.w);
            }
            stage1_ws_store(re, im, tw, col0 + buf * kWsECf);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(e_full + buf);
        __syncwarp();
        TL_MARK(warp, k, 4);
    }
}

// ------------------------------------------------------------------------------------------ consumers
// Tile (k-1) leaves shared memory: full tiles of a [.., T, 80] output go out as 8 bulk copies of 1 280 bytes (four
// frames) each (chunks w and w + 5 by lane 0 of consumer warp w of the group: per-lane bulk copies are serialised by the
// hardware interface, so they are spread over the warps); everything else (partial tiles, zero fill of frames
// beyond a row's own length, [.., 80, T] layout, unaligned output) takes the cooperative element-wise path.
// Returns whether bulk copies were issued.
// kNT threads (kNT / 32 warps) cooperate: the whole consumer role, or one of its two groups.
template <int kNT = kWsRoleThreads>
__device__ __forceinline__ bool ws_store_tile(const KernelArgs& a, const WsDesc* dp, const float* s_y, int tid) {
    constexpr int kNW = kNT / 32;
#if defined(TALFE_ABLATE) && (TALFE_ABLATE & 8)
    return false;                                                        // timing experiment only: features never leave shared memory
#endif
    const int flags = dp->flags;
    if (flags & kWsBulkY) {
        if ((tid & 31) == 0) {
            float* dst = dp->out_tile;
#pragma unroll
            for (int w = tid >> 5; w < kWsFrames / kWsYChunk; w += kNW)
                bulk_s2g(dst + kWsYChunk * w * kMaxMels, smem_u32(s_y + ws_y_off(kWsYChunk * w)), kWsYChunk * kMaxMels * (unsigned)sizeof(float));
            bulk_commit();
        }
        __syncwarp();
        return true;
    }
    const WsDesc t = *dp;
    const bool active = flags & kWsActive;
    if (!active && a.out_offsets) return false;                         // packed output has no padding frames
    float* out_row = a.out + (a.out_offsets ? a.out_offsets[t.row] * kMaxMels : (long long)t.row * a.out_row_stride);
    const int nfr = min(kWsFrames, a.frame_end - t.t0);
    if (a.out_layout == TALFE_LAYOUT_MT) {
        // transposed staging Yt[mel][frame]: lane = frame, one mel row (<= 128 contiguous bytes of the output) per warp store
        float* dst = out_row + (t.t0 - a.frame0);
        if (nfr == kWsFrames) {
            const int f = tid & 31;
            const float keep = (active && t.t0 + f < t.t_end) ? 1.f : 0.f;
#pragma unroll
            for (int q = 0; q < kMaxMels * kWsFrames / kNT; ++q) {
                const int m = (tid >> 5) + q * kNW;
                dst[(long long)m * a.n_frames + f] = keep != 0.f ? s_y[ws_yt_off(m, f)] : 0.f;
            }
        } else {
            for (int i = tid; i < nfr * kMaxMels; i += kNT) {
                const int m = i / nfr, f = i - m * nfr;
                const bool valid = active && t.t0 + f < t.t_end;
                dst[(long long)m * a.n_frames + f] = valid ? s_y[ws_yt_off(m, f)] : 0.f;
            }
        }
        return false;
    }
    for (int i = tid; i < nfr * kMaxMels; i += kNT) {
        const int f = i / kMaxMels, m = i - f * kMaxMels;
        const int t_abs = t.t0 + f;
        const bool valid = active && t_abs < t.t_end;
        if (!valid && a.out_offsets) continue;
        out_row[(long long)(t_abs - a.frame0) * kMaxMels + m] = valid ? s_y[ws_y_off(f) + m] : 0.f;
    }
    return false;
}

// The consumer role runs as TWO groups of 5 warps on alternate tiles (group q takes
// tiles q, q + 2, ...; buffers E[q], P[q], Y[q] are its own), every thread running two exchange rows (r0, r0 + 10) and
// two mel lanes per tile.  The groups drift apart in phase, so that one group's shared-memory phases (exchange loads,
// mel stage) can overlap the other's FP phase (FFT-20) across warps, which the hardware scheduler does for free and a
// fused instruction stream inside one warp does not.  Two named barriers per tile and group (160 threads):
//   A1: mel(k-2) done by all -> store of tile k-2 issued, P[q] free | wait E[q](k) | row r0 -> registers, FFT, power
//   -> P[q] | row r0+10 -> registers, release E[q] | FFT, power -> P[q] | wait for the store's reads
//   | A2: P[q](k) complete, Y[q] free | mel(k): P[q] -> Y[q] | fence
// (A1 and the store issue BEFORE the wait for E measured 79.3 us against 80.4 us for the other order; computing the
// first FFT ahead of A1 with its 20 power values parked in registers, 89.4 us)
// (measured against ONE group of 10 warps with one row per thread and one barrier per tile: 80.2 against 82.6 us;
// the ablation switches 32 / 64 of that version went with it)
constexpr int kCs2Threads = kWsRoleThreads / 2;                         // 160
template <typename XT, bool kApply>
__device__ __forceinline__ void ws_consumer(const KernelArgs& a, unsigned char* smem, const cf* s_e0, cf* s_p0, float* s_y0,
                                                const WsDesc* s_desc, unsigned long long* s_bar, const float2* s_norm, const int tid, const int n_my) {
    const int grp = tid >= kCs2Threads ? 1 : 0;
    const int gtid = tid - grp * kCs2Threads;
    unsigned long long* e_full = s_bar + 4 + grp;
    unsigned long long* e_empty = s_bar + 6 + grp;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = gtid & (kWsGroups - 1), r0 = gtid >> 4, r1 = r0 + 10;  // rows / mel lanes r0 (0..9) and r1 (10..19)
    const bool special1 = r1 >= 18;                                     // the group's last warp: packed rows as its second item
    const int bar_id = 1 + 2 * grp;                                     // 1 / 3 (2: producers' edge tiles, 4: all consumers)
    const float4* s_w4a = reinterpret_cast<const float4*>(smem + a.off_w_ws) + r0 * (kRefWStride / 4);
    const float4* s_w4b = reinterpret_cast<const float4*>(smem + a.off_w_ws) + r1 * (kRefWStride / 4);
    int lo0[kMelSlots], lo1[kMelSlots];
#pragma unroll
    for (int i = 0; i < kMelSlots; ++i) {
        lo0[i] = reinterpret_cast<const int*>(smem + a.off_lo_ws)[i * 20 + r0];
        lo1[i] = reinterpret_cast<const int*>(smem + a.off_lo_ws)[i * 20 + r1];
    }
    const cf* e_row0 = s_e0 + grp * kWsECf + ws_e_base(g) + r0 * kWsERow;
    const cf* e_row1 = e_row0 + 10 * kWsERow;
    cf* s_p = s_p0 + grp * kWsPCf;
    float* s_y = s_y0 + grp * kWsYFloats;
    const bool mt = a.out_layout == TALFE_LAYOUT_MT;
    float* yb_a = s_y + (mt ? ws_yt_off(r0, 2 * g) : ws_y_off(2 * g) + r0);
    float* yb_b = s_y + (mt ? ws_yt_off(r1, 2 * g) : ws_y_off(2 * g) + r1);
    double acc_s = 0.0, acc_q = 0.0;

    auto stage2_row = [&](cf (&v)[20], int r, bool special) {
        cf pw[10];
        if (!special) {
            stage2_ws_power_normal(v, pw);
            stage2_ws_store_normal(1 + (r >> 1), pw, reinterpret_cast<float*>(s_p) + 2 * g + (r & 1));
        } else {
            stage2_ws_power_special(r == 18, v, pw);
            stage2_ws_store_special(r == 18, pw, s_p + g);
        }
    };
    auto mel_lane = [&](const float4* s_w4, const int (&lo)[kMelSlots], float* yb, const WsDesc* dp, int flags, float& sum, float& sumsq, const float2* nrm) {
        float w[kRefWStride];
#pragma unroll
        for (int q = 0; q < kRefWStride / 4; ++q) {
            const float4 t = s_w4[q];
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
        float y[2 * kMelSlots];
        mel_log_ws(s_p + g, w, lo, a.eps, y);
        if (kApply) {
            // given statistics (talfe_job::given_stats): (y - mean[mel]) * rstd[mel], the two roundings of apply_stats_kernel
#pragma unroll
            for (int i = 0; i < kMelSlots; ++i) {
                const float2 nm = nrm[20 * i];                          // (mean, 1 / std) of mel r + 20 i
                y[2 * i] = (y[2 * i] - nm.x) * nm.y;
                y[2 * i + 1] = (y[2 * i + 1] - nm.x) * nm.y;
            }
        }
        if (!mt) {
#pragma unroll
            for (int i = 0; i < kMelSlots; ++i) {
                yb[20 * i] = y[2 * i];
                yb[kMaxMels + 20 * i] = y[2 * i + 1];
            }
        } else {
#pragma unroll
            for (int i = 0; i < kMelSlots; ++i) {
                yb[20 * i * kWsYtStride] = y[2 * i];
                yb[20 * i * kWsYtStride + 1] = y[2 * i + 1];
            }
        }
        if (flags & kWsFull) {
#pragma unroll
            for (int i = 0; i < 2 * kMelSlots; ++i) {
                sum += y[i];
                if (a.want_sumsq) sumsq = fmaf(y[i], y[i], sumsq);
            }
        } else {
            const int ta = dp->t0 + 2 * g, t_end = dp->t_end;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                if (ta + f < t_end) {
#pragma unroll
                    for (int i = 0; i < kMelSlots; ++i) {
                        sum += y[2 * i + f];
                        if (a.want_sumsq) sumsq = fmaf(y[2 * i + f], y[2 * i + f], sumsq);
                    }
                }
            }
        }
    };

    int k_last = -1;
    for (int k = grp; k < n_my; k += 2) {
        const WsDesc* dp = s_desc + (k & (kWsDescRing - 1));
        cf v[20];
        TL_MARK(10 + warp, k, 0);
        named_bar_sync(bar_id, kCs2Threads);                            // A1: mel(k-2) finished everywhere
        TL_MARK(10 + warp, k, 1);
        if (k >= 2) ws_store_tile<kCs2Threads>(a, s_desc + ((k - 2) & (kWsDescRing - 1)), s_y, gtid);
        mbar_wait_backoff<TALFE_WS_EFULL_NS>(e_full, (k >> 1) & 1);
        TL_MARK(10 + warp, k, 2);
        const int flags = dp->flags;
        const bool active = flags & kWsActive;
        if (active) {
            stage2_load(e_row0, v);
            stage2_row(v, r0, false);
            stage2_load(e_row1, v);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(e_empty);    // both rows of E[grp] are in registers
        __syncwarp();
        TL_MARK(10 + warp, k, 3);
        if (active) stage2_row(v, r1, special1);
        TL_MARK(10 + warp, k, 4);
        if (lane == 0) bulk_wait_read<0>();                             // this lane's store of tile k-2 has finished reading Y[grp]
        named_bar_sync(bar_id, kCs2Threads);                            // A2: P[grp](k) complete, Y[grp] free
        TL_MARK(10 + warp, k, 5);
        float sum = 0.f, sumsq = 0.f;
        if (active) {
            mel_lane(s_w4a, lo0, yb_a, dp, flags, sum, sumsq, s_norm + r0);
            mel_lane(s_w4b, lo1, yb_b, dp, flags, sum, sumsq, s_norm + r1);
        }
        fence_proxy_async();                                            // Y[grp] writes -> visible to the bulk-copy engine
        TL_MARK(10 + warp, k, 6);
        if (a.partials_per_tile) {
            double ds = (double)sum, dq = (double)sumsq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ds += __shfl_xor_sync(0xffffffffu, ds, o);
                dq += __shfl_xor_sync(0xffffffffu, dq, o);
            }
            // the tile's statistics slot = its index in the FULL tile grid (the tile list may be compact: tile_map_kernel)
            const long long tile = (long long)dp->row * a.tiles_per_row + (dp->t0 - a.frame0) / kWsFrames;
            if (lane == 0) {                                            // 10 slots per tile: this group's 5 warps fill 5, zero the rest
                a.partials[tile * kWsRoleWarps + (gtid >> 5)] = make_double2(ds, dq);
                a.partials[tile * kWsRoleWarps + (gtid >> 5) + kWsRoleWarps / 2] = make_double2(0.0, 0.0);
            }
        } else {
            acc_s += (double)sum;
            acc_q += (double)sumsq;
        }
        k_last = k;
    }
    named_bar_sync(bar_id, kCs2Threads);
    if (k_last >= 0) ws_store_tile<kCs2Threads>(a, s_desc + (k_last & (kWsDescRing - 1)), s_y, gtid);
    if (!a.partials_per_tile) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
            acc_q += __shfl_xor_sync(0xffffffffu, acc_q, o);
        }
        named_bar_sync(4, kWsRoleThreads);                              // both groups are done with the power arrays
        double2* s_red = reinterpret_cast<double2*>(s_p0);
        if (lane == 0) s_red[warp] = make_double2(acc_s, acc_q);
        named_bar_sync(4, kWsRoleThreads);
        if (tid == 0) {
            double ts = 0.0, tq2 = 0.0;
            for (int w2 = 0; w2 < kWsRoleWarps; ++w2) { ts += s_red[w2].x; tq2 += s_red[w2].y; }
            a.partials[blockIdx.x] = make_double2(ts, tq2);
        }
    }
    if (lane == 0) bulk_wait_all<0>();                                  // shared memory must outlive the copies that read it
}

// ------------------------------------------------------------------------------------------ fused normalisation
// Grid-wide barrier in global memory (one thread per CTA calls it; the launch is cooperative, so every CTA is resident).
// bar[0] counts arrivals and is reset by the last arriver, bar[1] is the generation the others wait on: nothing needs
// zeroing between launches, and the grid size may change from one launch to the next.
__device__ __forceinline__ void ws_grid_barrier(unsigned* bar, unsigned n_ctas) {
    unsigned gen, old;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    if (old == n_ctas - 1) {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
    } else {
        unsigned now;
        do {
            __nanosleep(64);
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(bar + 1) : "memory");
        } while (now == gen);
    }
}

// mel -= mel.mean() (tal/asr/models.py:52) without a second launch.  Called by all 640 threads after both roles have
// left their tile loops.  The reduction order is the one sub_scalar_flat_kernel uses (256 strided sums, binary tree),
// so the fused and the two-kernel paths produce bit-identical features.
struct WsNormArgs {            // passed by value (registers): a reference to the kernel parameters would force a stack copy
    const double2* partials; unsigned* grid_bar; double norm_count; double* stats_out;
    float* out; long long out_row_stride; int n_rows;
};
__device__ __noinline__ void ws_fused_batch_mean(const WsNormArgs a, unsigned char* smem_scratch) {
    // nothing of this epilogue stays live in registers across the tile loops
    const int tid = (int)threadIdx.x;
    double* s_a = reinterpret_cast<double*>(smem_scratch);             // [256] sums, [256] sums of squares, then the mean
    double* s_b = s_a + 256;
    // every bulk store of this CTA has completed (its issuing lane waited for the whole group); make the copy engine's
    // writes and the generic-proxy reads below agree, then meet: the partial of this CTA was written by consumer thread 0
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        ws_grid_barrier(a.grid_bar, gridDim.x);
        __threadfence();
    }
    __syncthreads();
    if (tid < 256) {
        double ra = 0.0, rb = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += 256) {
            const double2 v = __ldcg(a.partials + i);
            ra += v.x; rb += v.y;
        }
        s_a[tid] = ra; s_b[tid] = rb;
    }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { s_a[tid] += s_a[tid + o]; s_b[tid] += s_b[tid + o]; }
        __syncthreads();
    }
    const float mean = a.norm_count > 0.0 ? (float)(s_a[0] / a.norm_count) : 0.f;
    if (a.stats_out && blockIdx.x == 0 && tid == 0) { a.stats_out[0] = a.norm_count; a.stats_out[1] = s_a[0]; a.stats_out[2] = s_b[0]; }
    // flat sweep of the whole (contiguous) tensor, the same slice of it for every CTA whatever tiles it computed: a call of
    // this size sits in L2 entirely, and an even split finishes sooner than "own tiles" when some CTAs had one tile more
    // than others (1 x 60 s: 188 tiles on 148 CTAs).  kNormUnroll independent 128-bit loads in flight per thread.
    constexpr int kNormUnroll = 8;
    const long long total = (long long)a.n_rows * a.out_row_stride;
    const long long n4 = total >> 2, step = (long long)gridDim.x * kWsThreads;
    float4* p4 = reinterpret_cast<float4*>(a.out);
    long long i = (long long)blockIdx.x * kWsThreads + tid;
    for (; i + (kNormUnroll - 1) * step < n4; i += kNormUnroll * step) {
        float4 v[kNormUnroll];
#pragma unroll
        for (int u = 0; u < kNormUnroll; ++u) v[u] = __ldcg(p4 + i + u * step);
#pragma unroll
        for (int u = 0; u < kNormUnroll; ++u) {
            v[u].x -= mean; v[u].y -= mean; v[u].z -= mean; v[u].w -= mean;
            __stcs(p4 + i + u * step, v[u]);
        }
    }
    for (; i < n4; i += step) {
        float4 v = __ldcg(p4 + i);
        v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
        __stcs(p4 + i, v);
    }
    if (blockIdx.x == 0 && tid < (int)(total & 3)) a.out[4 * n4 + tid] -= mean;
}

__host__ __device__ constexpr size_t ws_x_offset(size_t table_bytes) { return (table_bytes + 127) & ~(size_t)127; }

template <typename XT, bool kFuse, bool kApply = false, bool kCompact = false>
__global__ void __launch_bounds__(kWsThreads, 1) logmel_ws_kernel(const KernelArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(16) unsigned char smem[];              // (the dynamic window itself starts 1 KB aligned: no static shared memory)
    // carve-up: tables (twiddles | mel weights | first bins) | x[2] (128-byte aligned) | E[2] | P[2] | Y[2] | descriptor ring | 8 mbarriers | (mean, 1/std) table
    const unsigned x_off = (unsigned)ws_x_offset((size_t)a.blob_bytes);
    XT* s_x0 = reinterpret_cast<XT*>(smem + x_off);
    cf* s_e0 = reinterpret_cast<cf*>(smem + x_off + 2 * kWsXBufBytes);
    cf* s_p0 = s_e0 + 2 * kWsECf;
    float* s_y0 = reinterpret_cast<float*>(s_p0 + 2 * kWsPCf);
    WsDesc* s_desc = reinterpret_cast<WsDesc*>(s_y0 + 2 * kWsYFloats);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_desc + kWsDescRing);
    float2* s_norm = reinterpret_cast<float2*>(s_bar + 8);            // [80] (mean, 1 / std) per mel: kApply only

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(s_bar + 0, 1); mbar_init(s_bar + 1, 1);                               // x_full: the loader's arrival (+ bytes)
        // x_empty / e_full: one arrival per producer warp; e_empty: one per warp of the buffer's consumer group
        constexpr unsigned kP = kWsRoleWarps, kC = kWsRoleWarps / 2;
        mbar_init(s_bar + 2, kP); mbar_init(s_bar + 3, kP);
        mbar_init(s_bar + 4, kP); mbar_init(s_bar + 5, kP);
        mbar_init(s_bar + 6, kC); mbar_init(s_bar + 7, kC);
    }
    {
        const int4* src = reinterpret_cast<const int4*>(a.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = tid; i < a.blob_bytes / 16; i += kWsThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < 2 * kWsPCf; i += kWsThreads) s_p0[i] = make_float2(0.f, 0.f);   // incl. the never-written read padding
    }
    __syncthreads();
    cudaGridDependencySynchronize();
    if (kApply) {                                                      // the block may come from the preceding kernel (an all-reduce)
        if (tid < kMaxMels) {
            float mean, rstd;
            stats_to_norm(a.given_stats, a.given_norm, kMaxMels, tid, mean, rstd);
            s_norm[tid] = make_float2(mean, rstd);
        }
        __syncthreads();
    }
    const int n_tiles = kCompact ? __ldg(a.n_tiles_dev) : a.n_tiles;     // (compact list: counted on the device by tile_prefix_kernel)
    const int n_my = (!kCompact || (int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;   // tiles blockIdx.x, + gridDim.x, ...
    if (tid < kWsRoleThreads) ws_producer<XT, kCompact>(a, &tmap, smem, s_x0, s_e0, s_desc, s_bar, tid, n_my);
    else ws_consumer<XT, kApply>(a, smem, s_e0, s_p0, s_y0, s_desc, s_bar, s_norm, tid - kWsRoleThreads, n_my);
    if (kFuse) {                                                       // the exchange buffers are free now: scratch for the reduction
        const WsNormArgs na{a.partials, a.grid_bar, a.norm_count, a.stats_out, a.out, a.out_row_stride, a.batch};
        ws_fused_batch_mean(na, smem + x_off + 2 * kWsXBufBytes);
    }
}

constexpr size_t ws_smem_bytes(size_t table_bytes) {
    return ws_x_offset(table_bytes) + 2 * (size_t)kWsXBufBytes + 2 * (size_t)kWsECf * sizeof(cf) + 2 * (size_t)kWsPCf * sizeof(cf) +
           2 * (size_t)kWsYFloats * sizeof(float) + kWsDescRing * sizeof(WsDesc) + 8 * sizeof(unsigned long long) + kMaxMels * sizeof(float2);
}
static_assert(kMaxMels * kWsFrames % kWsRoleThreads == 0 && kWsYFloats >= kWsFrames * kMaxMels + 4 * kWsGroups, "Y staging");
static_assert(ws_smem_bytes(4160) <= 232448, "the ws kernel's shared memory must fit one SM (227 KB)");

}  // namespace
