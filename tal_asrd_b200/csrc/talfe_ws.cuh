// talfe_ws.cuh — warp-specialised version of K1 (included by talfe.cu after its helpers).
//
// One persistent CTA per SM, 640 threads:
//   warps  0..9   PRODUCERS  waveform tile (TMA, double-buffered) -> window -> FFT-20 -> untangle/twiddle -> exchange E
//   warps 10..19  CONSUMERS  exchange E -> FFT-20 -> power P -> mel projection -> log -> staged feature tile Y
//                            -> cp.async.bulk store to global memory (one 320-byte row per frame)
// Each role keeps only its own constants in registers (producers: 20 window taps + 10 twiddles; consumers: 28
// mel weights), so the steady state reads NO tables from shared memory.  The roles meet through mbarriers
// (x_full / x_empty / e_full / e_empty, two buffers each); the consumers synchronise among themselves with one
// named barrier (id 1); there is no CTA-wide barrier inside the tile loop.  Because the FP-heavy stage 1 and the
// shared-memory-heavy stage 2 / mel stage now run in different warps, the SM's schedulers overlap them
// instruction by instruction instead of phase by phase.
//
// Shared-memory layouts and their bank-conflict properties: talfe_core.cuh ("Warp-specialised path").
// Arithmetic is identical to the legacy kernel (same functions for the FFT, untangle, power, mel and log).
#pragma once

namespace {

constexpr int kWsRoleThreads = kWsGroups * kGroup;          // 320
constexpr int kWsRoleWarps = kWsRoleThreads / 32;           // 10
constexpr int kWsThreads = 2 * kWsRoleThreads;              // 640
constexpr int kWsTileSamples = kHop * kWsFrames + (kNfft - kHop);   // 5360
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
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of spinning on issue slots
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
    }
}

// Everything about tile (row, tq) that the roles need; recomputed where it is used (a handful of integer
// operations) instead of being carried across the FFTs in registers.
struct WsTile {
    int t0, L, t_end;
    bool active, full, bulk;
    long long src_off;             // element offset of tile sample 0 from a.wave (bulk tiles)
};

__device__ __forceinline__ WsTile ws_tile(const KernelArgs& a, int row, int tq) {
    WsTile ti;
    ti.t0 = a.frame0 + tq * kWsFrames;
    if (a.lens) {
        ti.L = (int)min(a.lens[row], (long long)kMaxSamples);
        ti.t_end = min(a.frame_end, ti.L > kHalf ? 1 + ti.L / kHop : 0);
    } else {
        ti.L = a.total_len;
        ti.t_end = a.t_end_const;
    }
    ti.active = ti.t0 < ti.t_end;
    ti.full = ti.t0 + kWsFrames <= ti.t_end;
    const int s0 = kHop * ti.t0 - kHalf;
    const int b0 = s0 - a.origin;
    const bool interior = s0 >= 0 && s0 + kWsTileSamples <= ti.L && b0 >= 0 && b0 + kWsTileSamples <= a.buf_len;
    ti.bulk = ti.active && interior && a.align_ok;
    ti.src_off = (long long)row * a.row_stride + b0;
    return ti;
}

__device__ __forceinline__ void ws_advance(const KernelArgs& a, int& row, int& tq, int step) {
    tq += step;
    while (tq >= a.tiles_per_row) { tq -= a.tiles_per_row; ++row; }
}

// ------------------------------------------------------------------------------------------ producers
template <typename XT, bool kTwReg, bool kWinReg>
__device__ __forceinline__ void ws_producer(const KernelArgs& a, unsigned char* smem, XT* s_x0, cf* s_e0, unsigned long long* s_bar,
                                            const int tid) {
    unsigned long long* x_full = s_bar;            // [2]
    unsigned long long* x_empty = s_bar + 2;       // [2]
    unsigned long long* e_full = s_bar + 4;        // [2]
    unsigned long long* e_empty = s_bar + 6;       // [2]
    const int warp = tid >> 5, lane = tid & 31;
    const int g1 = tid / kGroup, j = tid - g1 * kGroup;
    constexpr int kXG = XLayout<XT>::kGroup;
    constexpr int kXBufBytes = kXFloats * (int)sizeof(float);           // both element sizes use the fp32-sized buffer

    float win[20];
    if (kWinReg) load_window(j, reinterpret_cast<const float*>(smem), XLayout<XT>::kScale, win);
    const cf* s_tw = reinterpret_cast<const cf*>(smem + a.off_tw) + j * 10;
    cf tw[10];
    if (kTwReg) {
#pragma unroll
        for (int h = 0; h < 5; ++h) {
            const float4 tt = reinterpret_cast<const float4*>(s_tw)[h];
            tw[2 * h] = make_float2(tt.x, tt.y);
            tw[2 * h + 1] = make_float2(tt.z, tt.w);
        }
    }
    // a tile fetch = 17 pieces of <= 320 samples (one per skew block); warp w issues pieces w and w + 10
    auto issue = [&](const WsTile& t, int buf) {                        // lane 0 of every producer warp
        const XT* src = reinterpret_cast<const XT*>(a.wave) + t.src_off;
        const unsigned long long policy = l2_evict_first_policy();
        if (tid == 0) mbar_expect_tx(x_full + buf, kWsTileSamples * (int)sizeof(XT));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pb = warp + h * kWsRoleWarps;
            if (pb * kXBlock < kWsTileSamples)
                bulk_g2s_u32(smem_u32(s_x0 + pb * kXG) + buf * kXBufBytes, src + pb * kXBlock,
                             (unsigned)(min(kXBlock, kWsTileSamples - pb * kXBlock) * (int)sizeof(XT)), x_full + buf, policy);
        }
    };

    const int step = (int)gridDim.x;
    int tile = blockIdx.x;
    int row = tile / a.tiles_per_row, tq = tile - row * a.tiles_per_row;
    if (lane == 0) {
        const WsTile t0 = ws_tile(a, row, tq);
        if (t0.bulk) issue(t0, 0);
    }
    __syncwarp();
    unsigned xfull_par = 0;                                             // bit b: parity of the next x_full[b] phase to wait for
    const XT* xg = s_x0 + kXG * g1 + j;
    cf* col0 = s_e0 + ws_e_base(g1) + j;
    for (int k = 0; tile < a.n_tiles; tile += step, ++k) {
        const int buf = k & 1;
        if (lane == 0 && tile + step < a.n_tiles) {                     // tile k+1: HBM -> shared memory; tile k+2: HBM -> L2
            int rn = row, qn = tq;
            ws_advance(a, rn, qn, step);
            const WsTile tn = ws_tile(a, rn, qn);
            if (tn.bulk) {
                if (k >= 1) mbar_wait_sleep(x_empty + (buf ^ 1), ((k - 1) >> 1) & 1);   // tile k-1 has left that buffer
                issue(tn, buf ^ 1);
            }
            if (a.l2_prefetch && tid == 0 && tile + 2 * step < a.n_tiles) {
                ws_advance(a, rn, qn, step);
                const WsTile tp = ws_tile(a, rn, qn);
                if (tp.bulk) bulk_prefetch_l2(reinterpret_cast<const XT*>(a.wave) + tp.src_off, kWsTileSamples * (int)sizeof(XT));
            }
        }
        __syncwarp();                                                   // reconverge before the FFT (lane 0 took a detour)
        const WsTile ti = ws_tile(a, row, tq);
        const bool active = ti.active;
        cf z[20];
        if (active) {
            XT* s_x = reinterpret_cast<XT*>(reinterpret_cast<unsigned char*>(s_x0) + buf * kXBufBytes);
            if (ti.bulk) {
                mbar_wait_sleep(x_full + buf, (xfull_par >> buf) & 1);
                xfull_par ^= 1u << buf;
            } else {
                // edge tile (reflection), unaligned row or chunk boundary: element-wise staging by all producers
                if (k >= 2) mbar_wait_sleep(x_empty + buf, ((k - 2) >> 1) & 1);
                const int s0 = kHop * ti.t0 - kHalf;
                const XT* rowp = reinterpret_cast<const XT*>(a.wave) + (long long)row * a.row_stride;
                for (int i = tid; i < kWsTileSamples; i += kWsRoleThreads) {
                    int g = s0 + i;
                    if (g < 0) g = -g;                                  // reflect, no edge repeat
                    if (g >= ti.L) g = 2 * (ti.L - 1) - g;
                    const int bi = g - a.origin;
                    XT v = XT(0.f);
                    if (g >= 0 && g < ti.L && bi >= 0 && bi < a.buf_len) v = __ldg(rowp + bi);
                    s_x[xskew<XT>(i)] = v;
                }
                named_bar_sync(2, kWsRoleThreads);
            }
            if (!kWinReg) load_window(j, reinterpret_cast<const float*>(smem), XLayout<XT>::kScale, win);
            stage1_ws_fft<XT>(reinterpret_cast<const XT*>(reinterpret_cast<const unsigned char*>(xg) + buf * kXBufBytes), win, z);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(x_empty + buf);                      // this warp no longer reads x[buf]
        __syncwarp();
        // every tile, active or not: a producer never runs more than one phase ahead of the consumers
        if (k >= 2) mbar_wait_sleep(e_empty + buf, ((k - 2) >> 1) & 1); // consumers have loaded E[buf] of tile k-2
        if (active) {
            if (!kTwReg) {
#pragma unroll
                for (int h = 0; h < 5; ++h) {
                    const float4 tt = reinterpret_cast<const float4*>(s_tw)[h];
                    tw[2 * h] = make_float2(tt.x, tt.y);
                    tw[2 * h + 1] = make_float2(tt.z, tt.w);
                }
            }
            stage1_ws_store(z, tw, col0 + buf * kWsECf);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(e_full + buf);
        __syncwarp();
        ws_advance(a, row, tq, step);
    }
}

// ------------------------------------------------------------------------------------------ consumers
// Tile (k-1) leaves shared memory: full tiles of a [.., T, 80] output go out as 16 bulk copies of one 640-byte
// frame pair each; everything else (partial tiles, zero fill of frames beyond a row's own length, [.., 80, T]
// layout, unaligned output) takes the cooperative element-wise path.  Returns whether bulk copies were issued.
__device__ __forceinline__ bool ws_store_tile(const KernelArgs& a, int row, const WsTile& t, const float* s_y, int tid) {
    float* out_row = a.out + (a.out_offsets ? a.out_offsets[row] * kMaxMels : (long long)row * a.out_row_stride);
    if (t.active && t.full && a.out_layout == TALFE_LAYOUT_TM && a.out_align_ok) {
        // the two frames of a pair are contiguous in Y (ws_y_off): 16 copies of 640 bytes, pair w and w + 10 by lane 0
        // of consumer warp w (per-lane bulk copies are serialised by the hardware interface, so spread them over warps)
        if ((tid & 31) == 0) {
            const int w = tid >> 5;
            float* dst = out_row + (long long)(t.t0 - a.frame0) * kMaxMels;
            bulk_s2g(dst + 2 * w * kMaxMels, smem_u32(s_y + ws_y_off(2 * w)), 2 * kMaxMels * (unsigned)sizeof(float));
            if (w + kWsRoleWarps < kWsGroups)
                bulk_s2g(dst + 2 * (w + kWsRoleWarps) * kMaxMels, smem_u32(s_y + ws_y_off(2 * (w + kWsRoleWarps))),
                         2 * kMaxMels * (unsigned)sizeof(float));
            bulk_commit();
        }
        __syncwarp();
        return true;
    }
    if (!t.active && a.out_offsets) return false;                       // packed output has no padding frames
    const int nfr = min(kWsFrames, a.frame_end - t.t0);
    for (int i = tid; i < nfr * kMaxMels; i += kWsRoleThreads) {
        int f, m;
        if (a.out_layout == TALFE_LAYOUT_TM) { f = i / kMaxMels; m = i - f * kMaxMels; }
        else { m = i / nfr; f = i - m * nfr; }                          // frame fastest: contiguous in [.., 80, T]
        const int t_abs = t.t0 + f;
        const bool valid = t.active && t_abs < t.t_end;
        if (!valid && a.out_offsets) continue;
        const float v = valid ? s_y[ws_y_off(f) + m] : 0.f;
        if (a.out_layout == TALFE_LAYOUT_TM) out_row[(long long)(t_abs - a.frame0) * kMaxMels + m] = v;
        else out_row[(long long)m * a.n_frames + (t_abs - a.frame0)] = v;
    }
    return false;
}

template <typename XT, bool kMelReg>
__device__ __forceinline__ void ws_consumer(const KernelArgs& a, unsigned char* smem, const cf* s_e0, cf* s_p, float* s_y0,
                                            unsigned long long* s_bar, const int tid) {
    unsigned long long* e_full = s_bar + 4;
    unsigned long long* e_empty = s_bar + 6;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = tid & (kWsGroups - 1), r = tid >> 4;                  // r: exchange row in stage 2, mel lane in the mel stage
    const bool special = r >= 18;                                       // warp 9: the packed rows, both frames
    const float4* s_w4 = reinterpret_cast<const float4*>(smem + a.off_w_ws) + r * (kRefWStride / 4);
    float w[kRefWStride];
    if (kMelReg) {
#pragma unroll
        for (int q = 0; q < kRefWStride / 4; ++q) {
            const float4 t = s_w4[q];
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
    }
    int lo[kMelSlots];
#pragma unroll
    for (int i = 0; i < kMelSlots; ++i) lo[i] = reinterpret_cast<const int*>(smem + a.off_lo_ws)[i * 20 + r];
    const cf* e_row0 = s_e0 + ws_e_base(g) + r * kWsERow;
    float* pgf = reinterpret_cast<float*>(s_p) + 2 * g + (r & 1);
    cf* pg = s_p + g;
    float* yb0 = s_y0 + ws_y_off(2 * g) + r;
    const int k1 = 1 + (r >> 1);
    double acc_s = 0.0, acc_q = 0.0;

    const int step = (int)gridDim.x;
    int tile = blockIdx.x;
    int row = tile / a.tiles_per_row, tq = tile - row * a.tiles_per_row;
    int prow = row, ptq = tq;
    int k = 0;
    for (; tile < a.n_tiles; tile += step, ++k) {
        const int buf = k & 1;
        const WsTile ti = ws_tile(a, row, tq);
        cf v[20];
        mbar_wait_sleep(e_full + buf, (k >> 1) & 1);
        if (ti.active) stage2_load(e_row0 + buf * kWsECf, v);
        __syncwarp();
        if (lane == 0) mbar_arrive(e_empty + buf);
        __syncwarp();
        cf pw[10];
        if (ti.active) {
            if (!special) stage2_ws_power_normal(v, pw);
            else stage2_ws_power_special(r == 18, v, pw);
        }
        named_bar_sync(1, kWsRoleThreads);                              // A: mel(k-1) done everywhere: P is free, Y[(k-1)&1] is complete
        bool issued = false;
        if (k >= 1) issued = ws_store_tile(a, prow, ws_tile(a, prow, ptq), s_y0 + (buf ^ 1) * kWsYFloats, tid);
        if (ti.active) {
            if (!special) stage2_ws_store_normal(k1, pw, pgf);
            else stage2_ws_store_special(r == 18, pw, pg);
        }
        if (lane == 0) {                                                // this lane's store of tile k-2 must have finished reading Y[buf]
            if (issued) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        named_bar_sync(1, kWsRoleThreads);                              // B: P(k) complete
        float sum = 0.f, sumsq = 0.f;
        if (ti.active) {
            if (!kMelReg) {
#pragma unroll
                for (int q = 0; q < kRefWStride / 4; ++q) {
                    const float4 t = s_w4[q];
                    w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
                }
            }
            float y[2 * kMelSlots];
            mel_log_ws(pg, w, lo, a.eps, y);
            float* yb = yb0 + buf * kWsYFloats;
#pragma unroll
            for (int i = 0; i < kMelSlots; ++i) {
                yb[20 * i] = y[2 * i];
                yb[kMaxMels + 20 * i] = y[2 * i + 1];
            }
            const int ta = ti.t0 + 2 * g;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                if (ti.full || ta + f < ti.t_end) {
#pragma unroll
                    for (int i = 0; i < kMelSlots; ++i) {
                        sum += y[2 * i + f];
                        if (a.want_sumsq) sumsq = fmaf(y[2 * i + f], y[2 * i + f], sumsq);
                    }
                }
            }
        }
        fence_proxy_async();                                            // Y[buf] writes -> visible to the bulk-copy engine
        if (a.partials_per_tile) {
            double ds = (double)sum, dq = (double)sumsq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ds += __shfl_xor_sync(0xffffffffu, ds, o);
                dq += __shfl_xor_sync(0xffffffffu, dq, o);
            }
            if (lane == 0) a.partials[(long long)tile * kWsRoleWarps + warp] = make_double2(ds, dq);
        } else {
            acc_s += (double)sum;
            acc_q += (double)sumsq;
        }
        prow = row; ptq = tq;
        ws_advance(a, row, tq, step);
    }
    named_bar_sync(1, kWsRoleThreads);
    if (k >= 1) ws_store_tile(a, prow, ws_tile(a, prow, ptq), s_y0 + ((k - 1) & 1) * kWsYFloats, tid);
    if (!a.partials_per_tile) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
            acc_q += __shfl_xor_sync(0xffffffffu, acc_q, o);
        }
        double2* s_red = reinterpret_cast<double2*>(s_p);               // the power array is free after the last barrier
        if (lane == 0) s_red[warp] = make_double2(acc_s, acc_q);
        named_bar_sync(1, kWsRoleThreads);
        if (tid == 0) {
            double ts = 0.0, tq2 = 0.0;
            for (int w2 = 0; w2 < kWsRoleWarps; ++w2) { ts += s_red[w2].x; tq2 += s_red[w2].y; }
            a.partials[blockIdx.x] = make_double2(ts, tq2);
        }
    }
    if (lane == 0) bulk_wait_all<0>();                                  // shared memory must outlive the copies that read it
}

// kCfg bit 0: twiddles in producer registers, bit 1: window taps in producer registers, bit 2: mel weights in consumer
// registers (a cleared bit = re-read from the shared-memory tables once per tile)
template <typename XT, int kCfg>
__global__ void __launch_bounds__(kWsThreads, 1) logmel_ws_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    // carve-up: tables | x[2] | E[2] | P | Y[2] | 8 mbarriers
    XT* s_x0 = reinterpret_cast<XT*>(smem + a.blob_bytes);
    cf* s_e0 = reinterpret_cast<cf*>(smem + a.blob_bytes + 2 * kXFloats * sizeof(float));
    cf* s_p = s_e0 + 2 * kWsECf;
    float* s_y0 = reinterpret_cast<float*>(s_p + kWsPCf);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_y0 + 2 * kWsYFloats);

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(s_bar + 0, 1); mbar_init(s_bar + 1, 1);                               // x_full: the expect_tx arrival
        mbar_init(s_bar + 2, kWsRoleWarps); mbar_init(s_bar + 3, kWsRoleWarps);         // x_empty: one arrival per producer warp
        mbar_init(s_bar + 4, kWsRoleWarps); mbar_init(s_bar + 5, kWsRoleWarps);         // e_full
        mbar_init(s_bar + 6, kWsRoleWarps); mbar_init(s_bar + 7, kWsRoleWarps);         // e_empty: one per consumer warp
    }
    {
        const int4* src = reinterpret_cast<const int4*>(a.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = tid; i < a.blob_bytes / 16; i += kWsThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < kWsPCf; i += kWsThreads) s_p[i] = make_float2(0.f, 0.f);   // incl. the never-written read padding
    }
    __syncthreads();
    cudaGridDependencySynchronize();
    if (tid < kWsRoleThreads) ws_producer<XT, (kCfg & 1) != 0, (kCfg & 2) != 0>(a, smem, s_x0, s_e0, s_bar, tid);
    else ws_consumer<XT, (kCfg & 4) != 0>(a, smem, s_e0, s_p, s_y0, s_bar, tid - kWsRoleThreads);
}

constexpr size_t ws_smem_bytes(size_t blob_bytes) {
    return blob_bytes + 2 * (size_t)kXFloats * sizeof(float) + 2 * (size_t)kWsECf * sizeof(cf) + (size_t)kWsPCf * sizeof(cf) +
           2 * (size_t)kWsYFloats * sizeof(float) + 8 * sizeof(unsigned long long);
}

}  // namespace
