// talfe.cu — fused log-mel front end for NVIDIA B200 (sm_100a) and its C ABI (include/talfe.h).
//
// Replaces, for one path only, /root/reference/tal/asr/models.py:36-53 (LogMelSpec.forward), i.e.
// the ~10 library launches torchaudio/torch make for reflect-pad -> frame x Hann -> rFFT-400 ->
// |.|^2 -> [201x80] mel matmul -> log(.+eps) -> mean subtraction, with
//   K1  logmel_kernel        everything up to un-normalised log-mel, written to HBM once, plus
//                            per-warp partial sums of the output (deterministic, no atomics);
//   K2  reduce_partials      fixed-order reduction of those partials into the statistics block;
//   K3  apply_stats          in-place (x - mean) [* rstd] sweep (L2-resident for batch-sized outputs).
//
// K1 work decomposition: one CTA = 16 groups of 20 threads = 16 frame pairs = 32 consecutive frames
// of one row per tile; persistent CTAs stride over the tiles.  Math per group: talfe_core.cuh.
#include <cuda.h>                  // CUtensorMap + the cuTensorMapEncodeTiled prototype only: the entry point is looked up at run time
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/talfe.h"
#include "talfe_core.cuh"
#include "talfe_tables.h"

using namespace talfe;

namespace {

constexpr int kGroupsPerCta = 16;
constexpr int kThreads = kGroupsPerCta * kGroup;            // 320
constexpr int kWarps = kThreads / 32;                       // 10
constexpr int kFramesPerTile = 2 * kGroupsPerCta;           // 32
constexpr int kTileSamples = kHop * kFramesPerTile + (kNfft - kHop);   // 5360
constexpr int kXFloats = (kTileSamples + 24 * ((kTileSamples - 1) / kXBlock) + 3) & ~3;   // skewed tile in elements (fp32 sizing)
constexpr int kNormalThreads = 18 * kGroupsPerCta;          // 288 = 9 full warps; warp 9 = packed rows 0/10
constexpr int kMaxSamples = 0x7fff0000;                     // sample / frame indices are 32-bit on the device
constexpr int kMaxBands = 16;                               // SpecAugment bands per row and axis
constexpr int kColChunk = 256;                              // frames per block in the per-mel statistics pass
static_assert(kNormalThreads % 32 == 0, "special rows must fill whole warps");

// frames torch.stft(center=True, pad_mode='reflect') makes of L samples: 1 + (L + 2 (n_fft / 2) - n_fft) / hop, and none at
// all unless L > n_fft / 2 (the reflection needs it).  n_fft = 400, hop = 160: 1 + L / 160.
__host__ __device__ __forceinline__ long long frames_of(long long L, int hop, int nfft) {
    const int half = nfft / 2;
    return L > half ? 1 + (L + 2 * half - nfft) / hop : 0;
}

thread_local int g_last_cuda_error = 0;
#ifdef TALFE_TIMELINE
unsigned* g_timeline = nullptr;
#endif

// runtime copy of talfe::row_slot() (the constexpr table would otherwise be materialised on the stack)
__constant__ int c_row_slot[20] = {11, 16, 6, 17, 15, 4, 10, 13, 19, 8, 1, 14, 7, 12, 18, 5, 3, 0, 2, 9};

struct FlTables;                   // talfe_fl.cuh
struct talfe_plan_impl {
    int device;
    int n_mels;
    int sm_count;
    int ctas_per_sm;
    int ref_layout;                // 80-mel reference filterbank shape -> fully unrolled mel stage
    int variant;                   // 0 = legacy kernel (2 CTAs/SM, every thread runs every stage), 1 = warp-specialised
    int n_fft, hop;                // frame geometry: 400 / 160 for the specialised kernels, anything else -> generic kernel
    int generic;
    float* g_win; float2* g_tw; float* g_fb; int* g_mel_lo; int* g_mel_hi;     // generic kernel's tables (device)
    int fl;                        // frame-per-lane kernel for fp32 waveforms (TALFE_KERNEL=fl; needs the reference filterbank support)
    FlTables* fl_tables;           // its uniform tables (host copy: they travel as a kernel parameter)
    int l2_prefetch;
    int use_tma;                   // TALFE_TMA (default 1): fp32 tiles by one tensor copy instead of 17 bulk pieces
    size_t off_ws, ws_bytes, off_tw_ws, off_w_ws, off_lo_ws, ws_smem;   // table section staged by the ws kernel
    MelLayout layout;
    int pstride;
    size_t off_tw, off_w, off_lo, off_id, blob_bytes;
    unsigned char* blob_dev;
    size_t smem_bytes;
    // streaming (talfe_stream_episode): a side stream for host->device chunk copies and the events that
    // hand the two staging buffers back and forth; created with the plan, never on the hot path
    cudaStream_t copy_stream;
    cudaEvent_t ev_ready[2], ev_free[2];
    // fused normalisation: grid-barrier words, one pair per stream this plan has been used on (launches on one stream
    // are ordered; two streams must not share a barrier), and what the driver accepted at the first cooperative launch
    int fuse_norm;                 // TALFE_FUSED_NORM (default 1) and cooperative launch available
    int coop_with_pdl;             // 1: cooperative + programmatic serialisation together, 0: cooperative only, -1: untried
    unsigned* bar_dev;             // [kBarSlots][2], zeroed once
    void* bar_stream[16];
    int bar_used;
    std::mutex* bar_mutex;
};
constexpr int kBarSlots = 16;
constexpr int kFuseMaxTilesPerCta = 1;          // TALFE_FUSED_NORM=1: fuse calls of at most one tile per CTA

struct KernelArgs {
    const void* wave;
    int dtype;
    long long row_stride;          // elements between rows (pointer arithmetic only)
    int batch, buf_len, origin, total_len;     // sample indices fit 31 bits (checked on the host; lens[] is clamped)
    const long long* lens;
    int frame0, n_frames, frame_end;           // frame_end = frame0 + n_frames
    int t_end_const;                           // min(frame_end, 1 + total_len / 160) when lens == nullptr
    int align_ok;                              // every interior tile of every row starts 16-byte aligned
    float* out;
    long long out_row_stride;
    const long long* out_offsets;   // packed ragged output: row r starts at frame out_offsets[r] (nullptr = padded rows)
    int out_layout;
    float eps;
    int tiles_per_row, n_tiles;
    double2* partials;             // [n_tiles][kWarps] (per-row statistics) or [grid][kWarps]
    int partials_per_tile;
    int want_sumsq;
    const unsigned char* blob;
    int blob_bytes, off_tw, off_w, off_lo, off_id;
    MelLayout layout;
    int pstride;
    // warp-specialised kernel only
    int off_w_ws, off_lo_ws;       // its mel tables inside the section it stages (blob = twiddles | w_ws | lo_ws)
    const float* win_global;       // window taps [20][20] in global memory (read once into producer registers)
    int out_align_ok;              // every frame row of `out` starts 16-byte aligned (bulk stores allowed)
    int l2_prefetch;               // prefetch tile k+2 into L2 while tile k+1 travels to shared memory
    int use_tma;                   // fp32 interior tiles arrive as one tensor copy (the kernel's CUtensorMap parameter is valid)
    int use_tma_out;               // frame-per-lane kernel: full tiles of a [.., T, 80] output leave as one tensor store
    // reference normalisation fused into the kernel (one launch per LogMelSpec.forward): after its last tile every CTA
    // publishes its partial sum, the grid meets at a barrier in global memory (cooperative launch: all CTAs resident),
    // every CTA derives the same scalar mean from the partials in a fixed order and subtracts it from ITS OWN tiles,
    // which it wrote moments ago and which therefore sit in L2
    int fuse_norm;
    unsigned* grid_bar;            // [0] arrivals of the current launch, [1] generation; self-resetting
    double norm_count;             // B * T * M
    double* stats_out;             // count, sum, sum of squares (or nullptr)
    // dataset-level normalisation applied inside the ws kernel (talfe_job::given_stats): ONE statistics block for every row
    const double* given_stats;
    int given_norm;                // talfe_norm that selects which entries of given_stats are used
    // packed ragged output, ws kernel: compact tile list (only tiles that hold frames of their row), built on the device
    const int2* tile_map;          // [n_tiles_dev] (row, tile inside the row), or nullptr = the full batch x tiles_per_row grid
    const int* n_tiles_dev;
    unsigned* timeline;            // -DTALFE_TIMELINE development builds only (nullptr otherwise)
};

// ------------------------------------------------------------------------------------------ K1
// mbarrier + bulk-copy (TMA, 1-D) helpers: the waveform tile of the NEXT iteration is fetched by the
// copy engine straight into shared memory while the SMs work on the current one.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// waveform samples are used once: fetch them with an L2 evict-first policy so that the features K1 writes
// stay L2-resident for the normalisation sweep that follows
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_u32(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar,
                                             unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

struct TileInfo {
    int row;
    int tq;                        // tile index inside the row
    int t0, L, t_end;              // first frame, row length, end of the row's valid frames
    bool active;                   // the tile has at least one frame the row really owns
    bool full;                     // every frame of the tile is valid and inside [frame0, frame0 + n_frames)
    bool bulk;                     // staged by the copy engine (completion on the mbarrier)
};

// Per-tile bookkeeping is executed by every warp (uniform datapath); it is kept to a handful of 32-bit
// operations: everything that does not depend on the tile is folded into KernelArgs on the host.
__device__ __forceinline__ void tile_fill(const KernelArgs& a, TileInfo& ti) {
    ti.t0 = a.frame0 + ti.tq * kFramesPerTile;
    if (a.lens) {
        ti.L = (int)min(a.lens[ti.row], (long long)kMaxSamples);
        ti.t_end = min(a.frame_end, ti.L > kHalf ? 1 + ti.L / kHop : 0);
    } else {
        ti.L = a.total_len;
        ti.t_end = a.t_end_const;                                       // min(frame_end, 1 + total_len / 160)
    }
    ti.active = ti.t0 < ti.t_end;
    ti.full = ti.t0 + kFramesPerTile <= ti.t_end;
    ti.bulk = false;
}

// what one warp contributes to a tile fetch: pieces `warp` and `warp + 10` of the 17 (constant per kernel)
struct WarpCopy {
    unsigned dst0, dst1;           // shared-memory addresses of the two pieces
    int off0, off1;                // source offsets in elements
    unsigned bytes0, bytes1;       // bytes1 == 0: this warp has a single piece
    unsigned long long policy;     // L2 evict-first
};

// Stage 0: waveform tile -> shared memory, skewed layout (talfe_core.cuh).  Interior tiles of an
// aligned row are fetched by the copy engine (cp.async.bulk, one piece per 320-sample block so that the
// skew can be inserted; completion is signalled on `bar`); one elected lane per warp issues that warp's
// one or two pieces, thread 0 posts the byte count.  Edge tiles (reflection) and unaligned rows take the
// synchronous element-wise path; the barrier that follows in program order publishes them.
template <typename XT>
__device__ __forceinline__ void load_tile(const KernelArgs& a, TileInfo& ti, XT* s_x, unsigned long long* bar,
                                          const WarpCopy& wc, int tid) {
    if (!ti.active) return;
    const int s0 = kHop * ti.t0 - kHalf;                                // episode index of tile sample 0
    const int b0 = s0 - a.origin;                                       // buffer index of tile sample 0
    const XT* rowp = reinterpret_cast<const XT*>(a.wave) + (long long)ti.row * a.row_stride;
    const bool interior = s0 >= 0 && s0 + kTileSamples <= ti.L && b0 >= 0 && b0 + kTileSamples <= a.buf_len;
    if (interior && a.align_ok) {
        ti.bulk = true;
        if ((tid & 31) == 0) {
            const XT* src = rowp + b0;
            if (tid == 0) mbar_expect_tx(bar, kTileSamples * (int)sizeof(XT));
            bulk_g2s_u32(wc.dst0, src + wc.off0, wc.bytes0, bar, wc.policy);
            if (wc.bytes1) bulk_g2s_u32(wc.dst1, src + wc.off1, wc.bytes1, bar, wc.policy);
        }
    } else {
        for (int i = tid; i < kTileSamples; i += kThreads) {
            int g = s0 + i;
            if (g < 0) g = -g;                                          // reflect, no edge repeat
            if (g >= ti.L) g = 2 * (ti.L - 1) - g;
            const int bi = g - a.origin;
            XT v = XT(0.f);
            if (g >= 0 && g < ti.L && bi >= 0 && bi < a.buf_len) v = __ldg(rowp + bi);
            s_x[xskew<XT>(i)] = v;
        }
    }
}

}  // namespace
// (mean, 1 / std) of mel m from a statistics block, as the normalisation mode defines them: the ONE place this is
// computed, so that the sweep (apply_stats_kernel) and the in-kernel application (logmel_ws_kernel<.., kApply>) agree bitwise
static __device__ __forceinline__ void stats_to_norm(const double* __restrict__ s, int norm, int n_mels, int m, float& mean, float& rstd) {
    mean = 0.f; rstd = 1.f;
    if (norm == TALFE_NORM_BATCH_MEAN || norm == TALFE_NORM_ROW_MEAN) {
        mean = s[0] > 0.0 ? (float)(s[1] / s[0]) : 0.f;
    } else if (norm == TALFE_NORM_ROW_MEL_MEAN || norm == TALFE_NORM_ROW_MEL_MEANVAR) {
        const double n = s[0] / n_mels;
        const double mu = n > 0.0 ? s[3 + m] / n : 0.0;
        mean = (float)mu;
        if (norm == TALFE_NORM_ROW_MEL_MEANVAR && n > 0.0) {
            double var = s[3 + n_mels + m] / n - mu * mu;
            if (var < 1e-10) var = 1e-10;
            rstd = (float)rsqrt(var);
        }
    }
}

#include "talfe_ws.cuh"
#include "talfe_fl.cuh"
#include "talfe_generic.cuh"
namespace {

template <bool kRef, typename XT>
__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    // shared-memory carve-up (all offsets multiples of 16 bytes)
    const float* s_win = reinterpret_cast<const float*>(smem);
    const cf* s_tw = reinterpret_cast<const cf*>(smem + a.off_tw);
    const float* s_w = reinterpret_cast<const float*>(smem + a.off_w);
    const int* s_lo = reinterpret_cast<const int*>(smem + a.off_lo);
    const int* s_id = reinterpret_cast<const int*>(smem + a.off_id);
    XT* s_x = reinterpret_cast<XT*>(smem + a.blob_bytes);
    cf* s_e = reinterpret_cast<cf*>(smem + a.blob_bytes + kXFloats * sizeof(float));
    cf* s_p = s_e + kGroupsPerCta * kEGroup;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_p + kGroupsPerCta * a.pstride);

    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(s_bar, 1);
    {   // constant tables -> shared memory; power array (incl. its padding) zeroed once
        const int4* src = reinterpret_cast<const int4*>(a.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = tid; i < a.blob_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < kGroupsPerCta * a.pstride; i += kThreads) s_p[i] = make_float2(0.f, 0.f);
    }
    __syncthreads();                                                    // tables + mbarrier visible
    cudaGridDependencySynchronize();       // launched programmatically: everything above overlapped the previous kernel's tail
    int tile = blockIdx.x;
    WarpCopy wc;
    {
        constexpr int kXG = XLayout<XT>::kGroup;
        const int b0 = tid >> 5, b1 = b0 + kWarps;
        wc.dst0 = smem_u32(s_x + b0 * kXG); wc.off0 = b0 * kXBlock;
        wc.bytes0 = (unsigned)(min(kXBlock, kTileSamples - b0 * kXBlock) * (int)sizeof(XT));
        wc.dst1 = smem_u32(s_x + b1 * kXG); wc.off1 = b1 * kXBlock;
        wc.bytes1 = b1 * kXBlock < kTileSamples ? (unsigned)(min(kXBlock, kTileSamples - b1 * kXBlock) * (int)sizeof(XT)) : 0u;
        wc.policy = l2_evict_first_policy();
    }
    TileInfo ti;
    ti.row = tile / a.tiles_per_row;
    ti.tq = tile - ti.row * a.tiles_per_row;
    tile_fill(a, ti);
    if (tile < a.n_tiles) load_tile(a, ti, s_x, s_bar, wc, tid);
    unsigned parity = 0;
    // roles: stage 1 and the mel stage use (pair g1, lane j); stage 2 uses (pair g2, exchange row)
    const int g1 = tid / kGroup, j = tid - g1 * kGroup;
    int g2, row;
    if (tid < kNormalThreads) { g2 = tid / 18; row = tid - g2 * 18; }
    else { const int s = tid - kNormalThreads; g2 = s >> 1; row = 18 + (s & 1); }
    const int warp = tid >> 5, lane = tid & 31;
    const int M = a.layout.n_mels;
    const XT* xg = s_x + XLayout<XT>::kGroup * g1;
    cf* e1 = s_e + g1 * kEGroup;
    const cf* e2 = s_e + g2 * kEGroup + c_row_slot[row] * kERow;
    float* p2w = reinterpret_cast<float*>(s_p + g2 * a.pstride);
    const cf* p2r = s_p + g1 * a.pstride;
    int lo[kMelSlots], mid[kMelSlots];
#pragma unroll
    for (int i = 0; i < kMelSlots; ++i) { lo[i] = s_lo[i * 20 + j]; mid[i] = s_id[i * 20 + j]; }
    float win[20];                                                      // this lane's 20 window taps stay in registers
    load_window(j, s_win, XLayout<XT>::kScale, win);
    double acc_s = 0.0, acc_q = 0.0;                                    // per-thread sums when partials are per CTA
    __syncthreads();                                                    // publishes an element-wise first tile

    // Software pipeline over tiles n = 0, 1, ...:
    //     phase A(n): stage 2 of tile n              (exchange rows -> power spectra; TMA fetch of tile n+1 in flight)
    //     phase B(n): mel/log/store of tile n  +  stage 1 of tile n+1   (independent: P(n) vs x(n+1) -> E(n+1))
    // one barrier after each phase.  In phase B even warps run the shared-memory-heavy mel stage first and
    // odd warps the FP-heavy stage 1 first, so that the two kinds of work overlap on the SM's pipes.
    auto run_stage1 = [&](const TileInfo& t) {
        if (t.bulk) { mbar_wait(s_bar, parity); parity ^= 1; }
        if (t.active) stage1(j, xg, win, s_tw, e1);
    };
    auto run_mel = [&](const TileInfo& cur, int cur_tile) {
        float* out_row = a.out + (a.out_offsets ? a.out_offsets[cur.row] * M : (long long)cur.row * a.out_row_stride);
        float sum = 0.f, sumsq = 0.f;
        if (cur.active) {
            float y[2 * kMelSlots];
            if (kRef) mel_log_ref(j, p2r, s_w, lo, a.eps, y);
            else mel_log_generic(j, a.layout, p2r, s_w, s_lo, a.eps, y);
            const int fa = cur.t0 - a.frame0 + 2 * g1;                 // output frame index of frame a
            if (cur.full && a.out_layout == TALFE_LAYOUT_TM) {
                float* o = out_row + (long long)fa * M;
#pragma unroll
                for (int i = 0; i < kMelSlots; ++i) {
                    if (kRef || mid[i] >= 0) {
                        o[mid[i]] = y[2 * i];
                        o[M + mid[i]] = y[2 * i + 1];
                        sum += y[2 * i] + y[2 * i + 1];
                        if (a.want_sumsq) sumsq = fmaf(y[2 * i], y[2 * i], fmaf(y[2 * i + 1], y[2 * i + 1], sumsq));
                    }
                }
            } else {
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const int t = cur.t0 + 2 * g1 + f;
                    if (t < a.frame_end) {
                        const bool valid = t < cur.t_end;
#pragma unroll
                        for (int i = 0; i < kMelSlots; ++i) {
                            const int m = mid[i];
                            if (m >= 0 && (valid || !a.out_offsets)) {
                                const float v = valid ? y[2 * i + f] : 0.f;
                                sum += v;
                                if (a.want_sumsq) sumsq = fmaf(v, v, sumsq);
                                if (a.out_layout == TALFE_LAYOUT_TM) out_row[(long long)(fa + f) * M + m] = v;
                                else out_row[(long long)m * a.n_frames + (fa + f)] = v;
                            }
                        }
                    }
                }
            }
        } else if (!a.out_offsets) {
            // tile lies entirely beyond this row's own frames ("each row as if alone"): zero fill
            const int nfr = min(kFramesPerTile, a.frame_end - cur.t0);
            for (int i = tid; i < nfr * M; i += kThreads) {
                const int f = i / M, m = i - f * M;
                if (a.out_layout == TALFE_LAYOUT_TM) out_row[(long long)(cur.t0 - a.frame0 + f) * M + m] = 0.f;
                else out_row[(long long)m * a.n_frames + (cur.t0 - a.frame0 + f)] = 0.f;
            }
        }
        if (a.partials_per_tile) {
            // per-warp partial sums (fixed shuffle tree -> bit-reproducible), one slot per (tile, warp)
            double ds = (double)sum, dq = (double)sumsq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ds += __shfl_xor_sync(0xffffffffu, ds, o);
                dq += __shfl_xor_sync(0xffffffffu, dq, o);
            }
            if (lane == 0) a.partials[(long long)cur_tile * kWarps + warp] = make_double2(ds, dq);
        } else {
            acc_s += (double)sum;
            acc_q += (double)sumsq;
        }
    };

    if (tile < a.n_tiles) run_stage1(ti);
    __syncthreads();                                                    // E(0) complete, s_x free
    for (; tile < a.n_tiles; tile += gridDim.x) {
        const TileInfo cur = ti;
        const bool has_next = tile + (int)gridDim.x < a.n_tiles;
        if (has_next) {                                                 // fetch tile n+1 behind stage 2 of tile n
            ti.tq += gridDim.x;
            while (ti.tq >= a.tiles_per_row) { ti.tq -= a.tiles_per_row; ++ti.row; }
            tile_fill(a, ti);
            load_tile(a, ti, s_x, s_bar, wc, tid);
        }
        if (cur.active) {                                               // phase A: stage 2
            cf v[20];
            stage2_load(e2, v);
            if (tid < kNormalThreads) stage2_normal(row, v, p2w);
            else stage2_special(row, v, p2w);
        }
        __syncthreads();                                                // P(n) complete, E free (and an element-wise x(n+1) published)
        if (warp & 1) {                                                 // phase B
            if (has_next) run_stage1(ti);
            run_mel(cur, tile);
        } else {
            run_mel(cur, tile);
            if (has_next) run_stage1(ti);
        }
        __syncthreads();                                                // E(n+1) complete, s_x and P free
    }
    if (!a.partials_per_tile) {
        // one slot per CTA: warp shuffle tree, then a fixed-order sum over the 10 warps (bit-reproducible)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
            acc_q += __shfl_xor_sync(0xffffffffu, acc_q, o);
        }
        double2* s_red = reinterpret_cast<double2*>(s_e);              // exchange buffer is free after the last barrier
        if (lane == 0) s_red[warp] = make_double2(acc_s, acc_q);
        __syncthreads();
        if (tid == 0) {
            double ts = 0.0, tq2 = 0.0;
            for (int w2 = 0; w2 < kWarps; ++w2) { ts += s_red[w2].x; tq2 += s_red[w2].y; }
            a.partials[blockIdx.x] = make_double2(ts, tq2);
        }
    }
}

// ------------------------------------------------------------------------------------------ compact tile list
// Packed ragged output: tiles of row r that hold frames of the row = ceil(T_r / 32), T_r from the row's own length.
// tile_prefix_kernel (one block): exclusive prefix over the rows -> prefix[0 .. B], total -> prefix[B] and *n_tiles;
// tile_map_kernel: compact tile t -> (row, tile inside the row) by binary search in the prefix.
// Frames of [frame0, frame_end) that row r's tiles must COMPUTE.  Per-row semantics: the row's own frames.  Padding hint
// (talfe_job::lens_are_padding_hint: reference semantics, the caller guarantees zeros beyond lens[r]): the frames whose
// window can see a sample below lens[r] — frame t sees samples [hop t - half, hop t - half + n_fft), and the reflection
// at the end of the PADDED row brings samples of its last n_fft / 2 back in, so a row that ends within n_fft of the padded
// length counts as full.  Every other frame is the constant log(eps).
__device__ __forceinline__ long long frames_to_compute(long long len, int hint, long long total_len, int frame0, int frame_end, int hop, int nfft) {
    len = min(len, (long long)kMaxSamples);
    long long t_row;
    if (!hint) t_row = frames_of(len, hop, nfft);
    else if (len <= 0) t_row = 0;
    else if (len + nfft + 2 >= total_len) t_row = frame_end;
    else t_row = (len + nfft / 2 + hop - 1) / hop;
    const long long v = min((long long)frame_end, t_row) - frame0;
    return v > 0 ? v : 0;
}
__global__ void __launch_bounds__(1024) tile_prefix_kernel(const long long* __restrict__ lens, int batch, int frame0, int frame_end, int hop,
                                                           int nfft, int* __restrict__ prefix, int* __restrict__ n_tiles, int hint, long long total_len) {
    __shared__ int s_part[1024];
    const int per = (batch + 1023) / 1024, lo = min(batch, (int)threadIdx.x * per), hi = min(batch, lo + per);
    auto tiles_of = [&](int r) {
        const long long v = frames_to_compute(lens[r], hint, total_len, frame0, frame_end, hop, nfft);
        return (int)((v + kFramesPerTile - 1) / kFramesPerTile);
    };
    int sum = 0;
    for (int r = lo; r < hi; ++r) sum += tiles_of(r);
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {                                            // 1 024 partial sums: a serial scan is a few microseconds
        int run = 0;
        for (int i = 0; i < 1024; ++i) { const int v = s_part[i]; s_part[i] = run; run += v; }
        prefix[batch] = run;
        *n_tiles = run;
    }
    __syncthreads();
    int run = s_part[threadIdx.x];
    for (int r = lo; r < hi; ++r) { prefix[r] = run; run += tiles_of(r); }
}
__global__ void __launch_bounds__(256) tile_map_kernel(const int* __restrict__ prefix, int batch, int2* __restrict__ map) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= prefix[batch]) return;
    int lo = 0, hi = batch;                                            // largest row with prefix[row] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid;
    }
    map[t] = make_int2(lo, t - prefix[lo]);
}

// Per-row lengths with a PADDED output: the compact tile list never visits the tiles beyond a row's own frames, so their
// zeros are written here, at streaming-store speed (block (row, y); frames from the end of the row's last tile onwards).
// (padding hint: the frames no tile computes are the constant log(eps) — the kernel's own expression for an all-zero frame —
// and const_frames[row] tells the mean how many of them there are)
__global__ void __launch_bounds__(256) zero_padding_kernel(float* __restrict__ out, long long out_row_stride, int out_layout, int n_mels,
                                                           const long long* __restrict__ lens, int frame0, int n_frames, int hop, int nfft,
                                                           int hint, long long total_len, float eps, long long* __restrict__ const_frames) {
    const long long row = blockIdx.x;
    const long long v = frames_to_compute(lens[row], hint, total_len, frame0, frame0 + n_frames, hop, nfft);
    const long long f0 = min((long long)n_frames, (v + kFramesPerTile - 1) / kFramesPerTile * kFramesPerTile);
    const float fill = hint ? fast_log(0.f + eps) : 0.f;
    if (const_frames && blockIdx.y == 0 && threadIdx.x == 0) const_frames[row] = n_frames - f0;
    float* base = out + row * out_row_stride;
    const long long tid = (long long)blockIdx.y * blockDim.x + threadIdx.x, nthr = (long long)gridDim.y * blockDim.x;
    if (out_layout == TALFE_LAYOUT_TM) {
        const long long lo = f0 * n_mels, hi = (long long)n_frames * n_mels;
        float* p = base + lo;
        long long n = hi - lo;
        if (n <= 0) return;
        const long long head = min(n, (long long)((16 - (reinterpret_cast<unsigned long long>(p) & 15ull)) & 15ull) / 4);
        if (tid < head) p[tid] = fill;
        p += head; n -= head;
        float4* p4 = reinterpret_cast<float4*>(p);
        const long long n4 = n / 4;
        for (long long i = tid; i < n4; i += nthr) __stcs(p4 + i, make_float4(fill, fill, fill, fill));
        if (tid < n - 4 * n4) p[4 * n4 + tid] = fill;
    } else {
        const long long w = n_frames - f0;
        for (long long i = tid; i < w * n_mels; i += nthr) {
            const long long m = i / w, f = i - m * w;
            base[m * n_frames + f0 + f] = fill;
        }
    }
}
// Padding hint with the batch-mean normalisation: the padding frames are written ONCE, already normalised.  Block (row, y):
// the scalar mean from the per-CTA partials + the constant frames' slot (fixed order), then x - mean over the frames the
// tiles computed and log(eps) - mean over the rest of the row (the same two values "fill, then sweep" would leave).
__global__ void __launch_bounds__(256) hint_count_kernel(const long long* __restrict__ lens, int batch, int frame0, int n_frames, int hop, int nfft,
                                                         long long total_len, long long* __restrict__ const_frames) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= batch) return;
    const long long v = frames_to_compute(lens[r], 1, total_len, frame0, frame0 + n_frames, hop, nfft);
    const_frames[r] = n_frames - min((long long)n_frames, (v + kFramesPerTile - 1) / kFramesPerTile * kFramesPerTile);
}
__global__ void __launch_bounds__(256) hint_finish_kernel(float* __restrict__ out, long long out_row_stride, int out_layout, int n_mels,
                                                          const long long* __restrict__ const_frames, int n_frames, float eps,
                                                          const double2* __restrict__ partials, int n_partials, double count,
                                                          double* __restrict__ stats_out) {
    __shared__ double s_a[256], s_b[256];
    double ra = 0.0, rb = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += 256) { const double2 v = partials[i]; ra += v.x; rb += v.y; }
    s_a[threadIdx.x] = ra; s_b[threadIdx.x] = rb;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
        __syncthreads();
    }
    const float mean = count > 0.0 ? (float)(s_a[0] / count) : 0.f;
    if (stats_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { stats_out[0] = count; stats_out[1] = s_a[0]; stats_out[2] = s_b[0]; }
    const float cm = fast_log(0.f + eps) - mean;
    const long long row = blockIdx.x;
    const long long f0 = n_frames - const_frames[row];                  // frames [0, f0) were computed, [f0, n_frames) are constant
    float* base = out + row * out_row_stride;
    const long long tid = (long long)blockIdx.y * blockDim.x + threadIdx.x, nthr = (long long)gridDim.y * blockDim.x;
    if (out_layout == TALFE_LAYOUT_TM && (n_mels & 3) == 0 && (reinterpret_cast<unsigned long long>(base) & 15ull) == 0) {
        float4* b4 = reinterpret_cast<float4*>(base);
        const long long n4 = (long long)n_frames * n_mels / 4, c4 = f0 * n_mels / 4;   // f0 is a multiple of 32 frames (or n_frames)
        const float4 fill = make_float4(cm, cm, cm, cm);
        for (long long i = tid; i < n4; i += nthr) {
            if (i < c4) {
                float4 v = b4[i];
                v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
                b4[i] = v;
            } else {
                __stcs(b4 + i, fill);
            }
        }
    } else if (out_layout == TALFE_LAYOUT_TM) {
        const long long n = (long long)n_frames * n_mels, c = f0 * n_mels;
        for (long long i = tid; i < n; i += nthr) base[i] = i < c ? base[i] - mean : cm;
    } else {
        const long long n = (long long)n_frames * n_mels;
        for (long long i = tid; i < n; i += nthr) {
            const long long f = i % n_frames;
            base[i] = f < f0 ? base[i] - mean : cm;
        }
    }
}
// The constant frames' contribution to the batch sums, as one more partial slot: integer frame count (exact, any order) x
// n_mels x c in double.
__global__ void __launch_bounds__(256) const_partial_kernel(const long long* __restrict__ const_frames, int batch, int n_mels, float eps,
                                                            double2* __restrict__ slot) {
    __shared__ long long s_n[256];
    long long n = 0;
    for (int r = threadIdx.x; r < batch; r += 256) n += const_frames[r];
    s_n[threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_n[threadIdx.x] += s_n[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double c = (double)fast_log(0.f + eps), cnt = (double)s_n[0] * n_mels;
        *slot = make_double2(cnt * c, cnt * c * c);
    }
}

// Collater padding found on the device: lens[r] = 1 + index of row r's last non-zero sample (0 for an all-zero row; lens is
// zeroed first).  Block (row, segment) scans ITS segment of the row backwards, every thread its own stride of samples until it
// meets a non-zero one: a thread inside the audio stops after one load, the padding is read once at memory speed (instead
// of being transformed), and the row's result is the maximum over its segments (atomicMax: order-independent).
constexpr int kPadSegments = 32, kPadThreads = 256;
template <typename XT>
__global__ void __launch_bounds__(kPadThreads) detect_padding_kernel(const XT* __restrict__ wave, long long row_stride, long long n_samples,
                                                                     unsigned long long* __restrict__ lens) {
    __shared__ long long s_last[kPadThreads / 32];
    const XT* row = wave + (long long)blockIdx.x * row_stride;
    const long long seg = (n_samples + kPadSegments - 1) / kPadSegments;
    const long long lo = (long long)blockIdx.y * seg, hi = min(n_samples, lo + seg);
    long long last = -1;
    long long i = hi - 1 - threadIdx.x;                                 // this thread's samples: hi-1-t, hi-1-t-256, ...
    if (i >= lo && x_to_float(row[i]) != 0.f) last = i;
    i -= kPadThreads;
    while (last < 0 && i >= lo) {
        constexpr int kU = 8;
        XT v[kU];
#pragma unroll
        for (int q = 0; q < kU; ++q) { const long long j = i - (long long)q * kPadThreads; v[q] = j >= lo ? row[j] : XT(0.f); }
#pragma unroll
        for (int q = kU - 1; q >= 0; --q) if (x_to_float(v[q]) != 0.f) last = max(last, i - (long long)q * kPadThreads);
        i -= (long long)kU * kPadThreads;
    }
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if ((threadIdx.x & 31) == 0) s_last[threadIdx.x >> 5] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long best = s_last[0];
#pragma unroll
        for (int w2 = 1; w2 < kPadThreads / 32; ++w2) best = max(best, s_last[w2]);
        if (best >= 0) atomicMax(lens + blockIdx.x, (unsigned long long)(best + 1));
    }
}

// ------------------------------------------------------------------------------------------ K2
// One block per statistics block (1 for batch-wide, B for per-row).  Fixed-order tree in shared
// memory: the result does not depend on scheduling.
constexpr int kReduceThreads = 1024;
constexpr int kReduceSplit = 64;                 // blocks per row of the first level when a row has many slots
constexpr long long kReduceSplitMin = 16384;     // rows with fewer slots are reduced by one block, as before
// First level for long rows: block (row, g) sums slots [g * slice, (g + 1) * slice) of the row in a fixed order.
__global__ void __launch_bounds__(256) reduce_slices_kernel(const double2* __restrict__ partials, long long slots_per_row, long long slice,
                                                            double2* __restrict__ out) {
    __shared__ double s_a[256], s_b[256];
    const long long row = blockIdx.x, g = blockIdx.y;
    const long long lo = g * slice, hi = min(lo + slice, slots_per_row);
    const double2* p = partials + row * slots_per_row;
    double a = 0.0, b = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) { const double2 v = p[i]; a += v.x; b += v.y; }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[row * gridDim.y + g] = make_double2(s_a[0], s_b[0]);
}
__global__ void __launch_bounds__(kReduceThreads) reduce_partials_kernel(const double2* __restrict__ partials, long long slots_per_block,
                                                              const long long* __restrict__ lens, long long total_len,
                                                              long long frame0, long long n_frames, long long rows_per_block,
                                                              int n_mels, int accumulate, double* __restrict__ stats, int hop, int nfft) {
    __shared__ double s_a[kReduceThreads], s_b[kReduceThreads];
    const long long blk = blockIdx.x;
    const double2* p = partials + blk * slots_per_block;
    double a = 0.0, b = 0.0;
    for (long long i = threadIdx.x; i < slots_per_block; i += kReduceThreads) { double2 v = p[i]; a += v.x; b += v.y; }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    for (int o = kReduceThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double count = 0.0;
        for (long long r = blk * rows_per_block; r < (blk + 1) * rows_per_block; ++r) {
            const long long L = lens ? lens[r] : total_len;
            const long long T_row = frames_of(L, hop, nfft);
            long long v = min(frame0 + n_frames, T_row) - frame0;
            if (v < 0) v = 0;
            count += (double)v * n_mels;
        }
        double* s = stats + blk * TALFE_STATS_DOUBLES(n_mels);
        if (accumulate) { s[0] += count; s[1] += s_a[0]; s[2] += s_b[0]; }
        else { s[0] = count; s[1] = s_a[0]; s[2] = s_b[0]; }
    }
}

// Per-mel column sums of un-normalised features (extension modes 3/4).  grid (B, chunks): the batch rides on grid.x,
// whose limit is 2^31 - 1; a row of the maximum supported length has ~52 k chunks, inside grid.y's 65 535.
__global__ void __launch_bounds__(320) colstats_kernel(const float* __restrict__ feats, long long out_row_stride, int out_layout,
                                                       long long n_frames, int n_mels, const long long* __restrict__ lens,
                                                       long long total_len, long long frame0, double* __restrict__ colpart,
                                                       const long long* __restrict__ out_offsets, int hop, int nfft, int chunks_per_block) {
    // block (row, g) sums the frames of chunks [g * chunks_per_block, (g + 1) * chunks_per_block): one partial per block, so
    // that the finishing kernel (ONE block per row) reads a few hundred partials even for an hour-long row (it used to read
    // 1 407 of them through a single SM: 33 us, more than this pass itself)
    __shared__ double s_sum[16][kMaxMels], s_sq[16][kMaxMels];
    const long long row = blockIdx.x, chunk = blockIdx.y;
    const long long L = lens ? lens[row] : total_len;
    const long long T_row = frames_of(L, hop, nfft);
    long long valid = min(frame0 + n_frames, T_row) - frame0;
    if (valid < 0) valid = 0;
    const long long f_lo = chunk * chunks_per_block * kColChunk, f_hi = min(f_lo + (long long)chunks_per_block * kColChunk, valid);
    const float* base = feats + (out_offsets ? out_offsets[row] * n_mels : row * out_row_stride);
    double* o = colpart + (row * gridDim.y + chunk) * 2 * kMaxMels;
    if (out_layout == TALFE_LAYOUT_TM && n_mels == kMaxMels && (reinterpret_cast<unsigned long long>(base) & 15ull) == 0) {
        // 20 float4 lanes x 16 frame lanes: every frame row is read as 320 contiguous bytes; fixed summation order
        const int q = threadIdx.x % 20, fl = threadIdx.x / 20;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        const float4* b4 = reinterpret_cast<const float4*>(base) + q;
#pragma unroll 4
        for (long long f = f_lo + fl; f < f_hi; f += 16) {
            const float4 v = b4[f * 20];
            a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
            b0 += (double)v.x * v.x; b1 += (double)v.y * v.y; b2 += (double)v.z * v.z; b3 += (double)v.w * v.w;
        }
        s_sum[fl][4 * q] = a0; s_sum[fl][4 * q + 1] = a1; s_sum[fl][4 * q + 2] = a2; s_sum[fl][4 * q + 3] = a3;
        s_sq[fl][4 * q] = b0; s_sq[fl][4 * q + 1] = b1; s_sq[fl][4 * q + 2] = b2; s_sq[fl][4 * q + 3] = b3;
        __syncthreads();
        if (threadIdx.x < kMaxMels) {
            const int m = threadIdx.x;
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { a += s_sum[i][m]; b += s_sq[i][m]; }
            o[m] = a; o[kMaxMels + m] = b;
        }
        return;
    }
    const int m = threadIdx.x % kMaxMels, lane_f = threadIdx.x / kMaxMels;       // 80 mels x 4 frame lanes
    double a = 0.0, b = 0.0;
    if (m < n_mels)
        for (long long f = f_lo + lane_f; f < f_hi; f += 4) {
            const float v = out_layout == TALFE_LAYOUT_TM ? base[f * n_mels + m] : base[(long long)m * n_frames + f];
            a += v; b += (double)v * v;
        }
    s_sum[lane_f][m] = a; s_sq[lane_f][m] = b;
    __syncthreads();
    if (lane_f == 0 && m < n_mels) {
        o[m] = (s_sum[0][m] + s_sum[1][m]) + (s_sum[2][m] + s_sum[3][m]);
        o[kMaxMels + m] = (s_sq[0][m] + s_sq[1][m]) + (s_sq[2][m] + s_sq[3][m]);
    }
}

constexpr int kFinishLanes = 64, kFinishMels = 16;                   // block (row, y): mels [16 y, 16 y + 16) x 64 partial lanes
__global__ void __launch_bounds__(kFinishLanes * kFinishMels) colstats_finish_kernel(const double* __restrict__ colpart, int chunks, int n_mels,
                                                                                     int accumulate, double* __restrict__ stats) {
    // lane l sums partials l, l + 64, ... of its mel; then a fixed-order sum over the 64 lanes (bit-reproducible)
    __shared__ double s_a[kFinishLanes][kFinishMels], s_b[kFinishLanes][kFinishMels];
    const long long row = blockIdx.x;
    const int mi = threadIdx.x % kFinishMels, l = threadIdx.x / kFinishMels, m = blockIdx.y * kFinishMels + mi;
    double a = 0.0, b = 0.0;
    if (m < n_mels)
        for (int ch = l; ch < chunks; ch += kFinishLanes) {
            const double* o = colpart + (row * chunks + ch) * 2 * kMaxMels;
            a += o[m]; b += o[kMaxMels + m];
        }
    s_a[l][mi] = a; s_b[l][mi] = b;
    __syncthreads();
    if (l != 0 || m >= n_mels) return;
    a = 0.0; b = 0.0;
#pragma unroll 8
    for (int i = 0; i < kFinishLanes; ++i) { a += s_a[i][mi]; b += s_b[i][mi]; }
    double* s = stats + row * TALFE_STATS_DOUBLES(n_mels);
    if (accumulate) { s[3 + m] += a; s[3 + n_mels + m] += b; }
    else { s[3 + m] = a; s[3 + n_mels + m] = b; }
}

// ------------------------------------------------------------------------------------------ K3
__global__ void __launch_bounds__(256) apply_stats_kernel(float* __restrict__ feats, long long batch, long long n_frames,
                                                          long long out_row_stride, int out_layout, int n_mels, int norm,
                                                          const double* __restrict__ stats, const long long* __restrict__ valid_frames,
                                                          const long long* __restrict__ lens, long long frame0,
                                                          const long long* __restrict__ out_offsets,
                                                          const int* __restrict__ freq_bands, const int* __restrict__ time_bands,
                                                          int n_bands, int hop, int nfft, int shared_block = 0) {
    __shared__ float s_mean[kMaxMels], s_rstd[kMaxMels];
    __shared__ int s_tb[2 * kMaxBands];
    const long long row = blockIdx.x;                  // batch on grid.x (no 65 535 limit), sweep blocks on grid.y
    const double* s = stats + ((norm == TALFE_NORM_BATCH_MEAN || shared_block) ? 0 : row * TALFE_STATS_DOUBLES(n_mels));
    long long valid = n_frames;
    if (valid_frames) valid = min(valid_frames[row], n_frames);
    else if (lens) {
        const long long L = lens[row];
        valid = max(0ll, min(frame0 + n_frames, frames_of(L, hop, nfft)) - frame0);
    }
    if (threadIdx.x < n_mels) {
        float mean, rstd;
        stats_to_norm(s, norm, n_mels, (int)threadIdx.x, mean, rstd);
        // SpecAugment frequency bands (tal/asr/models.py:531-548 sets them to 0 AFTER the mean subtraction):
        // a masked mel keeps its mean but gets a zero scale, so (x - mean) * 0 = 0
        for (int k = 0; k < n_bands; ++k) {
            const int lo_m = freq_bands[(row * n_bands + k) * 2], hi_m = freq_bands[(row * n_bands + k) * 2 + 1];
            if ((int)threadIdx.x >= lo_m && (int)threadIdx.x < hi_m) rstd = 0.f;
        }
        s_mean[threadIdx.x] = mean; s_rstd[threadIdx.x] = rstd;
    }
    if ((int)threadIdx.x < 2 * n_bands) s_tb[threadIdx.x] = time_bands[row * n_bands * 2 + threadIdx.x];
    __syncthreads();
    float* base = feats + (out_offsets ? out_offsets[row] * n_mels : row * out_row_stride);
    const long long total = valid * n_mels;
    auto time_masked = [&](int f) {                   // frame f inside a SpecAugment time band (models.py:550-566)
        bool hit = false;
        for (int k = 0; k < n_bands; ++k) hit |= (f >= s_tb[2 * k] && f < s_tb[2 * k + 1]);
        return hit;
    };
    if (out_layout == TALFE_LAYOUT_TM && (n_mels & 3) == 0 && ((reinterpret_cast<unsigned long long>(base) & 15ull) == 0)) {
        float4* b4 = reinterpret_cast<float4*>(base);
        const int m4 = n_mels >> 2;
        // ragged rows: a row is swept by its SHARE of the grid's y extent (blocks beyond it leave), so that a 10-minute row next
        // to 1-second rows is not left to the few blocks an even split would give it
        const long long n4 = total / 4, n4_max = (long long)n_frames * m4;
        const long long gy = (lens || valid_frames) ? max(1ll, ((long long)gridDim.y * n4 + n4_max - 1) / max(1ll, n4_max)) : (long long)gridDim.y;
        if ((long long)blockIdx.y >= gy) return;
        const long long step = gy * blockDim.x;
        long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x;
        if (!n_bands) {
            // HBM-sized sweeps (hour-long episodes, corpus passes): four independent 128-bit loads in flight per thread, and
            // the mel index carried along instead of a 64-bit division per element
            const int dm = (int)(step % m4);
            int mq = (int)(i % m4);
            auto fix = [&](float4 v, int q) {
                const int m = 4 * q;
                v.x = (v.x - s_mean[m]) * s_rstd[m];
                v.y = (v.y - s_mean[m + 1]) * s_rstd[m + 1];
                v.z = (v.z - s_mean[m + 2]) * s_rstd[m + 2];
                v.w = (v.w - s_mean[m + 3]) * s_rstd[m + 3];
                return v;
            };
            auto next = [&](int q) { q += dm; return q >= m4 ? q - m4 : q; };
            for (; i + 3 * step < n4; i += 4 * step) {
                const int q0 = mq, q1 = next(q0), q2 = next(q1), q3 = next(q2);
                mq = next(q3);
                const float4 a = b4[i], b = b4[i + step], c = b4[i + 2 * step], d = b4[i + 3 * step];
                b4[i] = fix(a, q0); b4[i + step] = fix(b, q1); b4[i + 2 * step] = fix(c, q2); b4[i + 3 * step] = fix(d, q3);
            }
            for (; i < n4; i += step) {
                b4[i] = fix(b4[i], mq);
                mq = next(mq);
            }
        } else {
            for (; i < n4; i += step) {
                const int f = (int)(i / m4);
                const int m = (int)(i - (long long)f * m4) * 4;
                float4 v = b4[i];
                v.x = (v.x - s_mean[m]) * s_rstd[m];
                v.y = (v.y - s_mean[m + 1]) * s_rstd[m + 1];
                v.z = (v.z - s_mean[m + 2]) * s_rstd[m + 2];
                v.w = (v.w - s_mean[m + 3]) * s_rstd[m + 3];
                if (time_masked(f)) v = make_float4(0.f, 0.f, 0.f, 0.f);
                b4[i] = v;
            }
        }
    } else if (out_layout == TALFE_LAYOUT_TM) {
        for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.y * blockDim.x) {
            const int f = (int)(i / n_mels);
            const int m = (int)(i - (long long)f * n_mels);
            base[i] = (n_bands && time_masked(f)) ? 0.f : (base[i] - s_mean[m]) * s_rstd[m];
        }
    } else {
        for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < (long long)n_mels * n_frames;
             i += (long long)gridDim.y * blockDim.x) {
            const int m = (int)(i / n_frames);
            const long long f = i - (long long)m * n_frames;
            if (f < valid) base[i] = (n_bands && time_masked((int)f)) ? 0.f : (base[i] - s_mean[m]) * s_rstd[m];
        }
    }
}

// Reference normalisation (one scalar for the whole contiguous tensor): flat float4 sweep, four
// independent 128-bit loads in flight per thread, most recently written data first (the tail of the
// tensor is what K1 left in L2 last).
__global__ void __launch_bounds__(256) sub_scalar_flat_kernel(float4* __restrict__ p, long long n4, float* __restrict__ tail,
                                                              int n_tail, const double2* __restrict__ partials, int n_partials,
                                                              double count, double* __restrict__ stats_out) {
    // every block re-derives the scalar from the few-hundred per-CTA partials K1 left in L2 (fixed order:
    // identical in every block and every run), which removes a separate reduction launch from the path
    __shared__ double s_a[256], s_b[256];
    cudaGridDependencySynchronize();                  // K1 (the preceding kernel in the stream) has completed and flushed
    double ra = 0.0, rb = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += 256) { const double2 v = partials[i]; ra += v.x; rb += v.y; }
    s_a[threadIdx.x] = ra; s_b[threadIdx.x] = rb;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
        __syncthreads();
    }
    const float mean = count > 0.0 ? (float)(s_a[0] / count) : 0.f;
    if (stats_out && blockIdx.x == 0 && threadIdx.x == 0) { stats_out[0] = count; stats_out[1] = s_a[0]; stats_out[2] = s_b[0]; }
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)(gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x;   // reversed block order
    long long i = first;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
        a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
        b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
        c.x -= mean; c.y -= mean; c.z -= mean; c.w -= mean;
        d.x -= mean; d.y -= mean; d.z -= mean; d.w -= mean;
        p[i] = a; p[i + stride] = b; p[i + 2 * stride] = c; p[i + 3 * stride] = d;
    }
    for (; i < n4; i += stride) {
        float4 a = p[i];
        a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
        p[i] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x < n_tail) tail[threadIdx.x] -= mean;
}

// ------------------------------------------------------------------------------------------ synthetic audio
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_kernel(void* wave, int dtype, long long rows, long long n, long long row_stride,
                             unsigned long long seed, long long first_episode, long long start) {
    const long long total = rows * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / n, col = idx - r * n;
        const unsigned long long key = splitmix64(seed ^ ((unsigned long long)(first_episode + r) * 0xD1B54A32D192ED03ull));
        const unsigned long long i = (unsigned long long)(start + col);
        const unsigned long long h = splitmix64(key + i);
        const long long s = (long long)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070;
        const long long amp = (s * 2838) >> 15;
        const unsigned long long eoff = (key >> 40) % 53333ull;
        const long long ph = (long long)((i + eoff) % 53333ull);
        long long tri = 2 * ph - 53333; if (tri < 0) tri = -tri;
        const long long env = 3277 + (62259 * tri) / 53333;
        long long k = (amp * env) >> 16;
        const unsigned long long g = splitmix64(key + 0x5851F42D4C957F2Dull * (i / 4000ull + 1ull));
        if ((g >> 32) % 10ull == 0ull) k = 0;
        k = max(-32767ll, min(32767ll, k));
        const long long o = r * row_stride + col;
        if (dtype == TALFE_F32) reinterpret_cast<float*>(wave)[o] = (float)k * (1.0f / 32768.0f);
        else if (dtype == TALFE_F16) reinterpret_cast<__half*>(wave)[o] = __float2half((float)k * (1.0f / 32768.0f));
        else reinterpret_cast<short*>(wave)[o] = (short)k;
    }
}

// ------------------------------------------------------------------------------------------ resampling (loader side)
// load_audio_segment resamples a file whose rate is not 16 kHz with torchaudio.transforms.Resample (tal/asr/data/util.py:44-48):
// a polyphase windowed-sinc FIR, out[b new + j] = sum_k x[b orig - width + k] K[j][k] with zeros beyond the ends
// (torchaudio/functional/functional.py:_apply_sinc_resample_kernel: pad (width, width + orig), conv1d with stride orig).
// The filter table K[new][2 width + orig] is built on the host by the binding with torchaudio's own formula.
template <typename XT>
__global__ void __launch_bounds__(256) resample_kernel(const XT* __restrict__ x, long long row_stride, long long n, int orig, int new_,
                                                       int width, const float* __restrict__ K, float* __restrict__ out,
                                                       long long out_len, long long out_row_stride, float scale) {
    const XT* xr = x + (long long)blockIdx.y * row_stride;
    float* orow = out + (long long)blockIdx.y * out_row_stride;
    const int taps = 2 * width + orig;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < out_len; o += (long long)gridDim.x * blockDim.x) {
        const long long blk = o / new_;
        const int j = (int)(o - blk * new_);
        const long long base = blk * orig - width;
        const float* kj = K + (long long)j * taps;
        float acc = 0.f;
        for (int k = 0; k < taps; ++k) {
            const long long i = base + k;
            if (i >= 0 && i < n) acc = fmaf(x_to_float(xr[i]) * scale, kj[k], acc);
        }
        orow[o] = acc;
    }
}

typedef void (*logmel_kernel_t)(const KernelArgs);
logmel_kernel_t kernel_for(bool ref_layout, int dtype) {
    if (ref_layout) return dtype == TALFE_F32 ? logmel_kernel<true, float> : dtype == TALFE_F16 ? logmel_kernel<true, __half> : logmel_kernel<true, short>;
    return dtype == TALFE_F32 ? logmel_kernel<false, float> : dtype == TALFE_F16 ? logmel_kernel<false, __half> : logmel_kernel<false, short>;
}

typedef void (*logmel_ws_kernel_t)(const KernelArgs, const CUtensorMap);
logmel_ws_kernel_t ws_kernel_for(int dtype, bool fuse = false, bool apply = false, bool compact = false) {
    if (compact) return dtype == TALFE_F32 ? logmel_ws_kernel<float, false, false, true> : dtype == TALFE_F16 ? logmel_ws_kernel<__half, false, false, true> : logmel_ws_kernel<short, false, false, true>;
    if (apply) return dtype == TALFE_F32 ? logmel_ws_kernel<float, false, true> : dtype == TALFE_F16 ? logmel_ws_kernel<__half, false, true> : logmel_ws_kernel<short, false, true>;
    if (fuse) return dtype == TALFE_F32 ? logmel_ws_kernel<float, true> : dtype == TALFE_F16 ? logmel_ws_kernel<__half, true> : logmel_ws_kernel<short, true>;
    return dtype == TALFE_F32 ? logmel_ws_kernel<float, false> : dtype == TALFE_F16 ? logmel_ws_kernel<__half, false> : logmel_ws_kernel<short, false>;
}

// Development / profiling knobs (read once per plan): TALFE_KERNEL=legacy|ws, TALFE_L2_PREFETCH=0|1.
int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

int cuda_fail(cudaError_t e) { g_last_cuda_error = (int)e; return TALFE_ERR_CUDA; }
#define TALFE_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__); } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// blocks along x for the in-place sweep: enough to fill the GPU ~8 deep across all rows
unsigned sweep_blocks(int sm_count, long long batch, long long dense_per_row, bool ragged = false) {
    long long want = (dense_per_row / 4 + 255) / 256;
    long long cap = std::max<long long>(1, (long long)sm_count * 8 / std::max<long long>(1, batch));
    if (ragged) cap = std::max<long long>(cap, (long long)sm_count * 2);   // each row uses its share of y (apply_stats_kernel)
    return (unsigned)std::max<long long>(1, std::min<long long>(std::min(want, cap), 65535));   // rides on grid.y
}

struct WorkspaceLayout { size_t partials, colpart, scratch_stats, partials2, tile_prefix, tile_map, const_frames, total; long long tiles_per_row, n_tiles; int chunks; };

WorkspaceLayout workspace_layout(int n_mels, long long batch, long long n_frames) {
    WorkspaceLayout w{};
    w.tiles_per_row = (n_frames + kFramesPerTile - 1) / kFramesPerTile;
    w.n_tiles = w.tiles_per_row * batch;
    w.chunks = (int)((n_frames + kColChunk - 1) / kColChunk);
    if (w.chunks < 1) w.chunks = 1;
    w.partials = 0;
    size_t off = align_up((size_t)w.n_tiles * kWarps * sizeof(double2), 256);
    w.colpart = off;
    off += align_up((size_t)batch * w.chunks * 2 * kMaxMels * sizeof(double), 256);
    w.scratch_stats = off;
    off += align_up((size_t)batch * TALFE_STATS_DOUBLES(n_mels) * sizeof(double), 256);
    w.partials2 = off;                                                  // first-level sums of long rows (reduce_slices_kernel)
    off += align_up((size_t)batch * kReduceSplit * sizeof(double2), 256);
    w.tile_prefix = off;                                                // compact tile list of packed ragged calls: prefix[B + 1], count
    off += align_up((size_t)(batch + 2) * sizeof(int), 256);
    w.tile_map = off;
    off += align_up((size_t)w.n_tiles * sizeof(int2), 256);
    w.const_frames = off;                                               // padding hint: frames per row that are the constant log(eps)
    off += align_up((size_t)batch * sizeof(long long), 256);
    w.total = off;
    return w;
}

// The waveform as the copy engine sees it (fp32 only): element (c0, c1, c2, c3) = sample 68 c1 + c0 of the 340-sample row that
// starts 320 c2 samples after the first interior tile position of batch row c3.  Rows overlap by 20 samples on purpose: a box
// of [68][5][17][1] written densely into shared memory IS the skewed tile (talfe_ws.cuh).  `shift` is the buffer index of
// the sample the tile grid starts at (tile tq of a row begins at shift + 5120 tq); it is negative when the grid starts in
// the reflected region, which only the first (edge) tile touches — that tile never uses the copy engine.
typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_t tensor_map_encoder() {
    static encode_tiled_t fn = []() -> encode_tiled_t {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<encode_tiled_t>(p);
    }();
    return fn;
}

bool encode_wave_map(CUtensorMap* tm, const float* wave, long long shift, long long buf_len, long long row_stride, long long batch) {
    const encode_tiled_t enc = tensor_map_encoder();
    if (!enc) return false;
    const long long rows = (buf_len - shift - kWsTmaRowFloats) / kXBlock + 1;      // 340-sample rows that end inside the buffer
    if (buf_len - shift < kWsTmaRowFloats || rows < kWsTmaRows) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)kWsTmaInner, (cuuint64_t)kWsTmaMid, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {kWsTmaInner * sizeof(float), kXBlock * sizeof(float), (cuuint64_t)row_stride * sizeof(float)};
    const cuuint32_t box[4] = {kWsTmaInner, kWsTmaMid, kWsTmaRows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    static const int l2 = env_int("TALFE_TMA_L2", 2);                    // development knob: L2 promotion none / 64 / 128 / 256 bytes
    const CUtensorMapL2promotion promo = l2 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2 == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : l2 == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(wave + shift), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- frame-per-lane kernel: host side
void fl_build_tables(const float* window, const float* fb, FlTables* t) {
    for (int p = 0; p < 10; ++p)
        for (int m = 0; m < 20; ++m) t->win2[p][m] = make_float2(0.5f * window[2 * p + 20 * m], 0.5f * window[2 * p + 1 + 20 * m]);
    for (int p = 0; p < 10; ++p)
        for (int k1 = 1; k1 <= 10; ++k1) {
            float w[4];
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * p + h;
                const double ang = -2.0 * M_PI * (double)((j * k1) % kNfft) / kNfft;   // same expression as build_tables (bit-identical twiddles)
                w[h] = (float)(2.0 * std::cos(ang));
                w[2 + h] = (float)(2.0 * std::sin(ang));
            }
            t->tw4[p][k1 - 1] = make_float4(w[0], w[1], w[2], w[3]);
        }
    // mel classes: first bin and zero-padded weights (bins 0 and 200 carry no weight: build_tables rejects such filterbanks)
    std::memset(t->w0, 0, sizeof(t->w0)); std::memset(t->w1, 0, sizeof(t->w1));
    std::memset(t->w2, 0, sizeof(t->w2)); std::memset(t->w3, 0, sizeof(t->w3));
    for (int m = 0; m < kMaxMels; ++m) {
        int first = -1;
        for (int f = 1; f < 200 && first < 0; ++f) if (fb[f * kMaxMels + m] != 0.f) first = f;
        if (first < 0) first = 1;                                       // empty filter -> log(eps)
        t->mel_lo[m] = first;
        const int c = m / 20, i = m % 20;
        const int width = c == 0 ? kRefW0 : c == 1 ? kRefW1 : c == 2 ? kRefW2 : kRefW3;
        for (int r = 0; r < width; ++r) {
            const float wv = first + r < 200 ? fb[(first + r) * kMaxMels + m] : 0.f;
            if (c == 0) t->w0[i][r] = wv; else if (c == 1) t->w1[i][r] = wv; else if (c == 2) t->w2[i][r] = wv; else t->w3[i][r] = wv;
        }
    }
}

// waveform as rows of 164 samples every 160 samples (box: 34 rows); features as [B][T][80] (box: 32 frames x 84 columns, the 4 columns past
// the 80 mels are never written)
bool fl_encode_maps(CUtensorMap* tm_in, CUtensorMap* tm_out, bool* out_ok, const float* wave, long long shift, long long buf_len,
                    long long row_stride, long long batch, float* out, long long n_frames, long long out_row_stride, bool want_out) {
    const encode_tiled_t enc = tensor_map_encoder();
    if (!enc) return false;
    // rows of 164 samples that start every 160 samples (they overlap by 4: the tile's padding columns carry the next row's
    // first samples, which nothing reads).  A box wider than the tensor, with the padding out of bounds, faulted for boxes
    // that cross a multiple of 256 rows on B200 (memcheck: illegal address in UTMALDG) — the overlapping form does not.
    const long long rows = (buf_len - shift - kFlRowPitch) / kHop + 1;  // rows that end inside the buffer
    if (buf_len - shift < kFlRowPitch || rows < kFlRows) return false;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)kFlRowPitch, (cuuint64_t)rows, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {kHop * sizeof(float), (cuuint64_t)row_stride * sizeof(float)};
        const cuuint32_t box[3] = {kFlRowPitch, kFlRows, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wave + shift), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    *out_ok = false;
    if (want_out) {
        const cuuint64_t dims[3] = {(cuuint64_t)kMaxMels, (cuuint64_t)n_frames, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {kMaxMels * sizeof(float), (cuuint64_t)out_row_stride * sizeof(float)};
        const cuuint32_t box[3] = {kFlYPitch, kFlFrames, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        *out_ok = enc(tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    return true;
}

}  // namespace


struct talfe_plan : talfe_plan_impl {};

extern "C" {

int talfe_version(void) { return TALFE_VERSION; }

size_t talfe_job_size(void) { return sizeof(talfe_job); }

#ifdef TALFE_TIMELINE
int talfe_debug_timeline(unsigned* host_out /* [20][64][8] */) {     // development builds only; not declared in talfe.h
    if (!g_timeline) return TALFE_ERR_INVALID;
    TALFE_CUDA(cudaDeviceSynchronize());
    TALFE_CUDA(cudaMemcpy(host_out, g_timeline, 20 * 64 * 8 * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return TALFE_OK;
}
#endif

int talfe_launches_per_forward(const talfe_plan* plan, int64_t batch, int64_t n_samples) {
    if (!plan || batch < 1 || n_samples <= plan->n_fft / 2) return TALFE_ERR_INVALID;
    const WorkspaceLayout w = workspace_layout(plan->n_mels, batch, frames_of(n_samples, plan->hop, plan->n_fft));
    const bool fused = !plan->generic && plan->variant == 1 && plan->fuse_norm &&
                       (plan->fuse_norm >= 2 || w.n_tiles <= (long long)kFuseMaxTilesPerCta * plan->sm_count);
    return fused ? 1 : 2;
}

const char* talfe_strerror(int status) {
    switch (status) {
        case TALFE_OK: return "ok";
        case TALFE_ERR_INVALID: return "invalid argument";
        case TALFE_ERR_TOO_SHORT: return "waveform too short: reflect padding of 200 needs more than 200 samples";
        case TALFE_ERR_UNSUPPORTED: return "configuration not supported by the sm_100a kernels";
        case TALFE_ERR_WORKSPACE: return "workspace missing or too small";
        case TALFE_ERR_CUDA: return "CUDA runtime error";
        case TALFE_ERR_NCCL: return "NCCL unavailable or collective failed";
        default: return "unknown status";
    }
}

int talfe_last_cuda_error(void) { return g_last_cuda_error; }

int64_t talfe_num_frames(int64_t n_samples) {
    if (n_samples <= kHalf) return TALFE_ERR_TOO_SHORT;
    return 1 + n_samples / kHop;
}

int64_t talfe_plan_num_frames(const talfe_plan* plan, int64_t n_samples) {
    if (!plan) return TALFE_ERR_INVALID;
    if (n_samples <= plan->n_fft / 2) return TALFE_ERR_TOO_SHORT;
    return frames_of(n_samples, plan->hop, plan->n_fft);
}

int talfe_plan_geometry(const talfe_plan* plan, int* n_fft, int* hop) {
    if (!plan) return TALFE_ERR_INVALID;
    if (n_fft) *n_fft = plan->n_fft;
    if (hop) *hop = plan->hop;
    return TALFE_OK;
}

int talfe_plan_create(talfe_plan** plan_out, int device, int n_mels, const float* window_host, const float* fb_host) {
    if (!plan_out) return TALFE_ERR_INVALID;
    *plan_out = nullptr;
    if (n_mels < 1 || n_mels > kMaxMels) return TALFE_ERR_UNSUPPORTED;
    std::vector<float> win(kNfft), fb((size_t)kBins * n_mels);
    if (window_host) win.assign(window_host, window_host + kNfft); else default_window(win.data());
    if (fb_host) fb.assign(fb_host, fb_host + (size_t)kBins * n_mels); else default_filterbank(n_mels, fb.data());
    HostTables t;
    int rc = build_tables(n_mels, win.data(), fb.data(), t);
    if (rc) return rc;

    int prev = 0;
    TALFE_CUDA(cudaGetDevice(&prev));
    TALFE_CUDA(cudaSetDevice(device));
    talfe_plan* p = new (std::nothrow) talfe_plan();
    if (!p) return TALFE_ERR_INVALID;
    p->device = device;
    p->n_mels = n_mels;
    p->n_fft = kNfft; p->hop = kHop; p->generic = 0;
    p->g_win = nullptr; p->g_tw = nullptr; p->g_fb = nullptr; p->g_mel_lo = nullptr; p->g_mel_hi = nullptr;
    p->layout = t.layout;
    p->pstride = t.pstride;
    p->off_tw = t.off_tw; p->off_w = t.off_w; p->off_lo = t.off_lo; p->off_id = t.off_id; p->blob_bytes = t.blob_bytes;
    p->ref_layout = is_reference_layout(t.layout) ? 1 : 0;
    p->off_ws = t.off_ws; p->ws_bytes = t.ws_bytes; p->off_tw_ws = t.off_tw_ws; p->off_w_ws = t.off_w_ws; p->off_lo_ws = t.off_lo_ws;
    p->ws_smem = ws_smem_bytes(t.ws_bytes);
    {
        const char* kv = std::getenv("TALFE_KERNEL");
        p->variant = p->ref_layout ? 1 : 0;                               // the warp-specialised kernel is unrolled for the reference filterbank shape
        if (kv && std::strcmp(kv, "legacy") == 0) p->variant = 0;
        p->fl = 0;
        p->fl_tables = nullptr;
        if (p->variant == 1 && n_mels == kMaxMels && kv && std::strcmp(kv, "fl") == 0) {
            p->fl_tables = new (std::nothrow) FlTables();
            if (p->fl_tables) { fl_build_tables(win.data(), fb.data(), p->fl_tables); p->fl = 1; }
        }
        p->l2_prefetch = env_int("TALFE_L2_PREFETCH", 0);
        p->use_tma = env_int("TALFE_TMA", 1);
    }
    // the legacy kernel stages only its own tables (the blob's prefix up to the ws section): 2 CTAs per SM need <= 113.5 KB each
    p->smem_bytes = t.off_ws + (size_t)kXFloats * sizeof(float) +
                    (size_t)kGroupsPerCta * kEGroup * sizeof(cf) + (size_t)kGroupsPerCta * t.pstride * sizeof(cf) + 16;
    cudaError_t e = cudaMalloc(&p->blob_dev, t.blob_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(p->blob_dev, t.blob.data(), t.blob_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    for (int dt = TALFE_F32; dt <= TALFE_I16 && e == cudaSuccess; ++dt)
        e = cudaFuncSetAttribute(kernel_for(p->ref_layout != 0, dt), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 std::max<int>((int)p->smem_bytes, 113 * 1024));   // the attribute is per kernel, plans of several table sizes coexist
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->ctas_per_sm, kernel_for(p->ref_layout != 0, TALFE_F32), kThreads, p->smem_bytes);
    if (p->variant == 1) {
        for (int dt = TALFE_F32; dt <= TALFE_I16 && e == cudaSuccess; ++dt) {
            e = cudaFuncSetAttribute(ws_kernel_for(dt, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->ws_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_kernel_for(dt, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->ws_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_kernel_for(dt, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->ws_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_kernel_for(dt, false, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->ws_smem);
        }
    }
    if (p->fl && e == cudaSuccess) e = cudaFuncSetAttribute(logmel_fl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlSmemBytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
    {
        int coop = 0;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        p->fuse_norm = (coop && p->variant == 1) ? env_int("TALFE_FUSED_NORM", 0) : 0;
        p->coop_with_pdl = -1;
        p->bar_used = 0;
        p->bar_mutex = new (std::nothrow) std::mutex();
        if (!p->bar_mutex) p->fuse_norm = 0;
        if (e == cudaSuccess) e = cudaMalloc(&p->bar_dev, kBarSlots * 2 * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMemset(p->bar_dev, 0, kBarSlots * 2 * sizeof(unsigned));
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p->ev_ready[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_free[i], cudaEventDisableTiming);
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) { talfe_plan_destroy(p); return cuda_fail(e); }
    if (p->ctas_per_sm < 1) { talfe_plan_destroy(p); return TALFE_ERR_UNSUPPORTED; }
    *plan_out = p;
    return TALFE_OK;
}

// Any other frame geometry (LogMelSpec(sr != 16000): n_fft = win = int(0.025 sr), hop = int(0.010 sr), tal/asr/models.py:24-32)
// gets a plan for the generic kernel (csrc/talfe_generic.cuh); n_fft 400 / hop 160 delegates to talfe_plan_create.
int talfe_plan_create_ex(talfe_plan** plan_out, int device, int n_fft, int hop, int n_mels, const float* window_host,
                         const float* fb_host) {
    if (!plan_out) return TALFE_ERR_INVALID;
    *plan_out = nullptr;
    if (n_fft == kNfft && hop == kHop) return talfe_plan_create(plan_out, device, n_mels, window_host, fb_host);
    if (!window_host || !fb_host || n_fft < 2 || hop < 1) return TALFE_ERR_INVALID;
    if (n_mels < 1 || n_mels > kMaxMels || n_fft > kGenMaxNfft || generic_smem_bytes(n_fft, hop) > 232448 - 1024) return TALFE_ERR_UNSUPPORTED;
    const int bins = n_fft / 2 + 1;
    std::vector<float2> tw(n_fft);
    for (int i = 0; i < n_fft; ++i) {
        const double ang = -2.0 * M_PI * (double)i / (double)n_fft;
        tw[i] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    std::vector<int> lo(n_mels, 1), hi(n_mels, 0);
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int f = 0; f < bins; ++f)
            if (fb_host[(size_t)f * n_mels + m] != 0.f) { if (first < 0) first = f; last = f; }
        if (first >= 0) { lo[m] = first; hi[m] = last; }
    }
    int prev = 0;
    TALFE_CUDA(cudaGetDevice(&prev));
    TALFE_CUDA(cudaSetDevice(device));
    talfe_plan* p = new (std::nothrow) talfe_plan();
    if (!p) return TALFE_ERR_INVALID;
    p->device = device; p->n_mels = n_mels; p->n_fft = n_fft; p->hop = hop; p->generic = 1;
    p->variant = 0; p->fl = 0; p->fuse_norm = 0; p->use_tma = 0; p->ref_layout = 0; p->ctas_per_sm = 1;
    cudaError_t e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    auto upload = [&](auto** dst, const void* src, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(dst), bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    upload(&p->g_win, window_host, sizeof(float) * n_fft);
    upload(&p->g_tw, tw.data(), sizeof(float2) * n_fft);
    upload(&p->g_fb, fb_host, sizeof(float) * (size_t)bins * n_mels);
    upload(&p->g_mel_lo, lo.data(), sizeof(int) * n_mels);
    upload(&p->g_mel_hi, hi.data(), sizeof(int) * n_mels);
    // (the attribute belongs to the kernel, not to the plan: plans of different geometries coexist, so ask for the most any of them may need)
    const int smem = 232448 - 1024;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_generic_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_generic_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_generic_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { talfe_plan_destroy(p); return cuda_fail(e); }
    *plan_out = p;
    return TALFE_OK;
}

void talfe_plan_destroy(talfe_plan* plan) {
    if (!plan) return;
    cudaFree(plan->g_win); cudaFree(plan->g_tw); cudaFree(plan->g_fb); cudaFree(plan->g_mel_lo); cudaFree(plan->g_mel_hi);
    if (plan->blob_dev) cudaFree(plan->blob_dev);
    if (plan->bar_dev) cudaFree(plan->bar_dev);
    delete plan->fl_tables;
    delete plan->bar_mutex;
    if (plan->copy_stream) cudaStreamDestroy(plan->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (plan->ev_ready[i]) cudaEventDestroy(plan->ev_ready[i]);
        if (plan->ev_free[i]) cudaEventDestroy(plan->ev_free[i]);
    }
    delete plan;
}

int talfe_plan_n_mels(const talfe_plan* plan) { return plan ? plan->n_mels : TALFE_ERR_INVALID; }

size_t talfe_workspace_bytes(const talfe_plan* plan, int64_t batch, int64_t n_frames) {
    if (!plan || batch < 1 || n_frames < 1) return 0;
    return workspace_layout(plan->n_mels, batch, n_frames).total;
}

int talfe_run(const talfe_plan* plan, const talfe_job* job) {
    if (!plan || !job || !job->wave || !job->out) return TALFE_ERR_INVALID;
    if (job->batch < 1 || job->n_frames < 1 || job->frame0 < 0 || job->buf_len < 1 || job->origin < 0) return TALFE_ERR_INVALID;
    if (job->wave_dtype < TALFE_F32 || job->wave_dtype > TALFE_I16) return TALFE_ERR_INVALID;
    if (job->norm < TALFE_NORM_NONE || job->norm > TALFE_NORM_ROW_MEL_MEANVAR) return TALFE_ERR_INVALID;
    if (job->out_layout != TALFE_LAYOUT_TM && job->out_layout != TALFE_LAYOUT_MT) return TALFE_ERR_INVALID;
    if (job->row_stride < job->buf_len) return TALFE_ERR_INVALID;
    if (!(job->eps >= 1.1754944e-38f)) return TALFE_ERR_UNSUPPORTED;    // the log is the bare MUFU.LG2 (subnormals flush): eps keeps its argument normal
    const int hop = plan->hop, nfft = plan->n_fft, half = nfft / 2;    // 160 / 400 / 200 unless the plan is a generic one
    if (!job->lens || job->lens_are_padding_hint) {
        if (job->total_len <= half) return TALFE_ERR_TOO_SHORT;
        if (job->frame0 + job->n_frames > frames_of(job->total_len, hop, nfft)) return TALFE_ERR_INVALID;
    }
    const int M = plan->n_mels;
    const long long dense = job->n_frames * M;
    const long long ors = job->out_row_stride ? job->out_row_stride : dense;
    if (ors < dense) return TALFE_ERR_INVALID;
    if (job->out_offsets && (!job->lens || job->out_layout != TALFE_LAYOUT_TM)) return TALFE_ERR_INVALID;   // packed = ragged [sum T_i, M]
    if (job->n_bands < 0 || job->n_bands > kMaxBands) return TALFE_ERR_INVALID;
    if (job->n_bands > 0 && (!job->freq_bands || !job->time_bands || job->norm == TALFE_NORM_NONE || job->defer_normalise))
        return TALFE_ERR_INVALID;                          // the masks ride on the normalisation sweep
    // lens as a PADDING HINT: reference (padded) semantics, rows guaranteed zero beyond lens[r]; scalar normalisations only
    const bool hint = job->lens_are_padding_hint != 0 && job->lens != nullptr;
    if (hint && (job->out_offsets || job->norm > TALFE_NORM_BATCH_MEAN || job->n_bands > 0 || job->given_stats || job->defer_normalise ||
                 (job->accumulate_stats && job->stats)))
        return TALFE_ERR_INVALID;
    const bool given = job->given_stats != nullptr;        // normalise with the caller's statistics block (dataset-level CMVN)
    if (given && (job->norm == TALFE_NORM_NONE || job->n_bands > 0 || job->defer_normalise)) return TALFE_ERR_INVALID;
    const WorkspaceLayout w = workspace_layout(M, job->batch, job->n_frames);
    if (!job->workspace || job->workspace_bytes < w.total) return TALFE_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(job->workspace) & 15) != 0) return TALFE_ERR_WORKSPACE;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(job->stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>(job->workspace);

    KernelArgs a{};
    a.wave = job->wave; a.dtype = job->wave_dtype;
    if (job->buf_len > kMaxSamples || job->origin > kMaxSamples || job->total_len > kMaxSamples ||
        job->frame0 + job->n_frames > kMaxSamples / hop || job->batch > 0x7fffffffLL)
        return TALFE_ERR_UNSUPPORTED;                         // ~37 h of 16 kHz audio per row: stream it in chunks instead
    a.batch = (int)job->batch; a.row_stride = job->row_stride; a.buf_len = (int)job->buf_len; a.origin = (int)job->origin;
    a.total_len = (int)job->total_len; a.lens = job->lens_are_padding_hint ? nullptr : reinterpret_cast<const long long*>(job->lens);
    a.frame0 = (int)job->frame0; a.n_frames = (int)job->n_frames; a.frame_end = a.frame0 + a.n_frames;
    a.t_end_const = (int)std::min<long long>(a.frame_end, frames_of(job->total_len, hop, nfft));
    {
        // 160 samples (one hop) are a multiple of 16 bytes for every element type, so whether an interior tile of
        // any row starts 16-byte aligned depends only on the base pointer, the row pitch and the chunk origin
        const long long elt = job->wave_dtype == TALFE_F32 ? 4 : 2;
        a.align_ok = ((reinterpret_cast<uintptr_t>(job->wave) & 15) == 0 && ((job->row_stride * elt) & 15) == 0 &&
                      (((kHalf + job->origin) * elt) & 15) == 0) ? 1 : 0;
    }
    a.out = job->out; a.out_row_stride = ors; a.out_layout = job->out_layout; a.eps = job->eps;
    a.out_offsets = reinterpret_cast<const long long*>(job->out_offsets);
    if (w.n_tiles > 0x7fffffffLL) return TALFE_ERR_UNSUPPORTED;
    a.tiles_per_row = (int)w.tiles_per_row; a.n_tiles = (int)w.n_tiles;
    a.partials = reinterpret_cast<double2*>(ws + w.partials);
    const bool use_ws = plan->variant == 1 && !plan->generic;
    a.blob = plan->blob_dev + (use_ws ? plan->off_ws : 0); a.blob_bytes = (int)(use_ws ? plan->ws_bytes : plan->off_ws);
    a.win_global = reinterpret_cast<const float*>(plan->blob_dev);
    a.off_tw = (int)(use_ws ? plan->off_tw_ws : plan->off_tw); a.off_w = (int)plan->off_w; a.off_lo = (int)plan->off_lo; a.off_id = (int)plan->off_id;
    a.layout = plan->layout; a.pstride = plan->pstride;

    a.off_w_ws = (int)plan->off_w_ws; a.off_lo_ws = (int)plan->off_lo_ws;
    a.l2_prefetch = plan->l2_prefetch;
    a.out_align_ok = ((reinterpret_cast<uintptr_t>(job->out) & 15) == 0 && (ors & 3) == 0) ? 1 : 0;
    long long grid = use_ws ? (long long)plan->sm_count : (long long)plan->sm_count * plan->ctas_per_sm;
    if (grid > w.n_tiles || plan->generic) grid = w.n_tiles;           // generic kernel: one CTA per tile, one partial per CTA
    CUtensorMap tmap{}, tmap_out{};
    bool use_fl = false;
    if (use_ws && plan->fl && a.dtype == TALFE_F32 && a.align_ok && job->row_stride < (1ll << 36)) {
        bool out_ok = false;
        const bool want_out = job->out_layout == TALFE_LAYOUT_TM && a.out_align_ok && !a.out_offsets;
        use_fl = fl_encode_maps(&tmap, &tmap_out, &out_ok, reinterpret_cast<const float*>(job->wave),
                                (long long)kHop * job->frame0 - kHalf - job->origin, job->buf_len, job->row_stride, job->batch,
                                job->out, job->n_frames, ors, want_out);
        static const int no_tma = env_int("TALFE_FL_NOTMA", 0);          // development: 1 = no tensor loads, 2 = no tensor stores
        if (use_fl) { a.use_tma = (no_tma & 1) ? 0 : 1; a.use_tma_out = (out_ok && !(no_tma & 2)) ? 1 : 0; }
    }
    if (use_fl) {
        grid = std::min<long long>(plan->sm_count, (w.n_tiles + kFlQuads - 1) / kFlQuads);
        static const int grid_cap = env_int("TALFE_FL_GRID", 0);        // development: fewer CTAs -> more tiles per warp
        if (grid_cap > 0) grid = std::min<long long>(grid, grid_cap);
    } else if (use_ws && plan->use_tma && a.dtype == TALFE_F32 && a.align_ok && job->row_stride < (1ll << 36))
        a.use_tma = encode_wave_map(&tmap, reinterpret_cast<const float*>(job->wave), (long long)kHop * job->frame0 - kHalf - job->origin,
                                    job->buf_len, job->row_stride, job->batch) ? 1 : 0;
    // packed ragged output through the ws kernel: a compact tile list built on the device (rows' own tiles only)
    static const int compact_ok = env_int("TALFE_COMPACT_TILES", 1);       // development: 0 = visit the full tile grid
    const bool compact = compact_ok && use_ws && !use_fl && !plan->generic && (a.lens || hint) && !given;   // (its own kernel instantiation)
    const bool hint_sum = compact && hint && job->norm == TALFE_NORM_BATCH_MEAN;       // the constant frames' share of the batch mean
    if (compact) {
        const long long* lens_dev = reinterpret_cast<const long long*>(job->lens);
        if (!a.out_offsets) {                                           // padded output: the frames no tile computes (zeros / log(eps))
            const unsigned zb = (unsigned)std::max<long long>(1, std::min<long long>(65535, (long long)plan->sm_count * 8 / job->batch));
            long long* const_frames = reinterpret_cast<long long*>(ws + w.const_frames);
            if (hint_sum)                                               // the padding is written once, normalised, after the transform
                hint_count_kernel<<<(unsigned)((job->batch + 255) / 256), 256, 0, stream>>>(lens_dev, a.batch, a.frame0, a.n_frames, hop, nfft,
                                                                                            job->total_len, const_frames);
            else
                zero_padding_kernel<<<dim3((unsigned)job->batch, zb), 256, 0, stream>>>(job->out, ors, job->out_layout, M, lens_dev, a.frame0,
                                                                                        a.n_frames, hop, nfft, hint ? 1 : 0, job->total_len, job->eps,
                                                                                        nullptr);
            TALFE_CUDA(cudaGetLastError());
            if (hint_sum) {                                             // one more partial slot behind the per-CTA ones
                const_partial_kernel<<<1, 256, 0, stream>>>(const_frames, a.batch, M, job->eps,
                                                            reinterpret_cast<double2*>(ws + w.partials) + grid);
                TALFE_CUDA(cudaGetLastError());
            }
        }
        int* prefix = reinterpret_cast<int*>(ws + w.tile_prefix);
        int2* map = reinterpret_cast<int2*>(ws + w.tile_map);
        tile_prefix_kernel<<<1, 1024, 0, stream>>>(lens_dev, a.batch, a.frame0, a.frame_end, hop, nfft, prefix, prefix + a.batch + 1, hint ? 1 : 0,
                                                   job->total_len);
        TALFE_CUDA(cudaGetLastError());
        tile_map_kernel<<<(unsigned)((w.n_tiles + 255) / 256), 256, 0, stream>>>(prefix, a.batch, map);
        TALFE_CUDA(cudaGetLastError());
        a.tile_map = map;
        a.n_tiles_dev = prefix + a.batch + 1;
        a.l2_prefetch = 0;
    }
    const bool want_stats = !given && (job->stats != nullptr || job->norm != TALFE_NORM_NONE);
    const bool per_row = job->norm >= TALFE_NORM_ROW_MEAN;
    // one slot per CTA is enough for batch-wide sums — and for per-row sums of a ONE-row call (an episode of a corpus pass,
    // a streamed chunk): per-tile slots (a double shuffle tree per warp and tile in the consumers, 16 bytes x 10 per tile
    // to reduce afterwards) cost 10 us per hour-long episode
    const bool tile_slots = per_row && !given && job->batch > 1;
    a.partials_per_tile = tile_slots ? 1 : 0;
    if (compact && tile_slots)                                          // tiles outside the compact list are never visited: their slots read as 0
        TALFE_CUDA(cudaMemsetAsync(ws + w.partials, 0, (size_t)w.n_tiles * kWarps * sizeof(double2), stream));
    a.want_sumsq = (job->stats != nullptr && !given) ? 1 : 0;   // the sum of squares is only ever reported, never needed by K3
    // given statistics are applied by the ws kernel's mel stage itself; the other kernels are followed by the sweep
    const bool apply_in_kernel = given && use_ws && !use_fl && !plan->generic;
    a.given_stats = apply_in_kernel ? job->given_stats : nullptr;
    a.given_norm = job->norm;
    // The reference case — one scalar over a contiguous [B, T, M] tensor — CAN be normalised inside the ws kernel itself
    // (cooperative launch, grid barrier, flat sweep split evenly over the CTAs): one launch per LogMelSpec.forward.
    const bool ref_norm = !given && job->norm == TALFE_NORM_BATCH_MEAN && !a.lens && !(job->accumulate_stats && job->stats) &&
                          !job->defer_normalise && ors == dense && job->n_bands == 0 &&
                          (reinterpret_cast<uintptr_t>(job->out) & 15) == 0;
    unsigned* bar = nullptr;
    // Opt-in since round 2's measurement across call sizes (tools/latency_fused_ab.py, profiles/r02_latency_fused_ab.json):
    // on the device the cooperative launch + grid barrier cost ~2 us more than the second, programmatically launched
    // kernel at EVERY size (1 x 1 s: 12.5 vs 10.4 us, 1 x 60 s: 14.4 vs 12.5, 64 x 30 s: 89.0 vs 87.2), and back to back
    // from Python the two-launch path is ahead as well once the host is warm (13.4 vs 14.5 us per call).
    // TALFE_FUSED_NORM: 0 (default) never, 1 calls of at most kFuseMaxTilesPerCta tiles per CTA, 2 always.
    const bool fuse_pays = plan->fuse_norm >= 2 || w.n_tiles <= (long long)kFuseMaxTilesPerCta * plan->sm_count;
    if (ref_norm && use_ws && !use_fl && !compact && plan->fuse_norm && fuse_pays && job->out_layout == TALFE_LAYOUT_TM && a.out_align_ok) {
        talfe_plan* pl = const_cast<talfe_plan*>(plan);
        std::lock_guard<std::mutex> lock(*pl->bar_mutex);
        int slot = -1;
        for (int i = 0; i < pl->bar_used; ++i) if (pl->bar_stream[i] == job->stream) slot = i;
        if (slot < 0 && pl->bar_used < kBarSlots) { slot = pl->bar_used++; pl->bar_stream[slot] = job->stream; }
        if (slot >= 0) bar = pl->bar_dev + 2 * slot;                  // more than 16 streams on one plan: two-kernel path
    }
    a.fuse_norm = bar ? 1 : 0;
    a.grid_bar = bar;
    a.norm_count = (double)dense * (double)job->batch;
    a.stats_out = job->stats;
#ifdef TALFE_TIMELINE
    {
        static unsigned* tl = nullptr;                                  // development build: one global buffer, read back by
        if (!tl) { cudaMalloc(&tl, 20 * 64 * 8 * sizeof(unsigned)); cudaMemset(tl, 0, 20 * 64 * 8 * sizeof(unsigned)); }
        a.timeline = tl;                                                // talfe_debug_timeline()
        g_timeline = tl;
    }
#endif
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(use_ws ? kWsThreads : kThreads);
        cfg.dynamicSmemBytes = use_ws ? plan->ws_smem : plan->smem_bytes; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (plan->generic) {
            GenericArgs g{nfft, hop, half, nfft / 2 + 1, plan->g_win, plan->g_tw, plan->g_fb, plan->g_mel_lo, plan->g_mel_hi, M};
            cfg.blockDim = dim3(kGenThreads); cfg.dynamicSmemBytes = generic_smem_bytes(nfft, hop);
            if (a.dtype == TALFE_F32) TALFE_CUDA(cudaLaunchKernelEx(&cfg, logmel_generic_kernel<float>, (const KernelArgs)a, (const GenericArgs)g));
            else if (a.dtype == TALFE_F16) TALFE_CUDA(cudaLaunchKernelEx(&cfg, logmel_generic_kernel<__half>, (const KernelArgs)a, (const GenericArgs)g));
            else TALFE_CUDA(cudaLaunchKernelEx(&cfg, logmel_generic_kernel<short>, (const KernelArgs)a, (const GenericArgs)g));
        } else if (use_fl) {
            cfg.blockDim = dim3(kFlThreads); cfg.dynamicSmemBytes = kFlSmemBytes;
            TALFE_CUDA(cudaLaunchKernelEx(&cfg, logmel_fl_kernel, (const KernelArgs)a, (const CUtensorMap)tmap, (const CUtensorMap)tmap_out,
                                          (const FlTables)*plan->fl_tables));
        } else if (!use_ws) {
            TALFE_CUDA(cudaLaunchKernelEx(&cfg, kernel_for(plan->ref_layout != 0, a.dtype), (const KernelArgs)a));
        } else if (a.fuse_norm) {
            const auto fn = ws_kernel_for(a.dtype, true);
            talfe_plan* pl = const_cast<talfe_plan*>(plan);
            cudaError_t e = cudaErrorUnknown;
            if (pl->coop_with_pdl != 0) {                              // first choice: keep the prologue overlap as well
                cfg.attrs = attr; cfg.numAttrs = 2;
                e = cudaLaunchKernelEx(&cfg, fn, (const KernelArgs)a, (const CUtensorMap)tmap);
                if (e != cudaSuccess && pl->coop_with_pdl < 0) { cudaGetLastError(); pl->coop_with_pdl = 0; }
                else if (e == cudaSuccess) pl->coop_with_pdl = 1;
            }
            if (pl->coop_with_pdl == 0) {
                cfg.attrs = attr + 1; cfg.numAttrs = 1;
                e = cudaLaunchKernelEx(&cfg, fn, (const KernelArgs)a, (const CUtensorMap)tmap);
            }
            TALFE_CUDA(e);
            return TALFE_OK;
        } else {
            TALFE_CUDA(cudaLaunchKernelEx(&cfg, ws_kernel_for(a.dtype, false, apply_in_kernel, compact), (const KernelArgs)a, (const CUtensorMap)tmap));
        }
    }
    if (given) {
        if (!apply_in_kernel) {                                         // legacy / frame-per-lane / generic kernel: the sweep applies the block
            apply_stats_kernel<<<dim3((unsigned)job->batch, sweep_blocks(plan->sm_count, job->batch, dense)), 256, 0, stream>>>(
                job->out, job->batch, job->n_frames, ors, job->out_layout, M, job->norm, job->given_stats, nullptr, a.lens, a.frame0,
                a.out_offsets, nullptr, nullptr, 0, hop, nfft, 1);
            TALFE_CUDA(cudaGetLastError());
        }
        return TALFE_OK;
    }

    if (!want_stats) return TALFE_OK;
    double* stats = job->stats ? job->stats : reinterpret_cast<double*>(ws + w.scratch_stats);
    const int accumulate = (job->accumulate_stats && job->stats) ? 1 : 0;
    if (hint_sum) {
        const unsigned yb = (unsigned)std::max<long long>(1, std::min<long long>(65535, (long long)plan->sm_count * 8 / job->batch));
        hint_finish_kernel<<<dim3((unsigned)job->batch, yb), 256, 0, stream>>>(job->out, ors, job->out_layout, M,
                                                                              reinterpret_cast<const long long*>(ws + w.const_frames), a.n_frames, job->eps,
                                                                              (const double2*)a.partials, (int)grid + 1, (double)dense * (double)job->batch,
                                                                              job->stats);
        TALFE_CUDA(cudaGetLastError());
        return TALFE_OK;
    }
    if (ref_norm) {
        // the reference case: one scalar over a contiguous [B, T, M] (or [B, M, T]) tensor; the sweep
        // derives the mean from the per-CTA partials itself (no separate reduction launch)
        const long long total = dense * job->batch, n4 = total / 4;
        static const int sweep_per_sm = env_int("TALFE_SWEEP_BLOCKS", 8);     // development knob (A/B: profiles/r02_ab_sweep.json)
        const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((n4 + 1023) / 1024, (long long)plan->sm_count * sweep_per_sm));
        // programmatic dependent launch: the sweep's blocks are scheduled while K1 drains and wait at
        // cudaGridDependencySynchronize() for its memory to be visible, which hides the launch gap
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        TALFE_CUDA(cudaLaunchKernelEx(&cfg, sub_scalar_flat_kernel, reinterpret_cast<float4*>(job->out), n4, job->out + 4 * n4,
                                      (int)(total - 4 * n4), (const double2*)a.partials, (int)grid + (hint_sum ? 1 : 0), (double)total, job->stats));
        return TALFE_OK;
    }
    const long long blocks = per_row ? job->batch : 1;
    const long long rows_per_block = per_row ? 1 : job->batch;
    const long long slots_per_block = tile_slots ? w.tiles_per_row * kWarps : grid;
    if (tile_slots && slots_per_block > kReduceSplitMin) {
        // a long row (an hour-long episode has 112 510 slots = 1.8 MB): kReduceSplit blocks per row sum contiguous slices in
        // a fixed order, the final block sums their kReduceSplit results (one block reading it all took 25 us)
        double2* part2 = reinterpret_cast<double2*>(ws + w.partials2);
        const long long slice = (slots_per_block + kReduceSplit - 1) / kReduceSplit;
        reduce_slices_kernel<<<dim3((unsigned)job->batch, kReduceSplit), 256, 0, stream>>>(a.partials, slots_per_block, slice, part2);
        TALFE_CUDA(cudaGetLastError());
        reduce_partials_kernel<<<(unsigned)blocks, kReduceThreads, 0, stream>>>(part2, kReduceSplit, a.lens, a.total_len,
                                                                      a.frame0, a.n_frames, rows_per_block, M, accumulate, stats, hop, nfft);
    } else {
        reduce_partials_kernel<<<(unsigned)blocks, kReduceThreads, 0, stream>>>(a.partials, slots_per_block, a.lens, a.total_len,
                                                                      a.frame0, a.n_frames, rows_per_block, M, accumulate, stats, hop, nfft);
    }
    TALFE_CUDA(cudaGetLastError());
    if (job->norm == TALFE_NORM_ROW_MEL_MEAN || job->norm == TALFE_NORM_ROW_MEL_MEANVAR) {
        double* colpart = reinterpret_cast<double*>(ws + w.colpart);
        // about four blocks per SM over the whole batch, each covering a contiguous range of 256-frame chunks
        const long long want_blocks = std::max<long long>(1, (long long)plan->sm_count * 4 / job->batch);
        const int cpb = (int)std::max<long long>(1, (w.chunks + want_blocks - 1) / want_blocks);
        const int col_blocks = (w.chunks + cpb - 1) / cpb;
        colstats_kernel<<<dim3((unsigned)job->batch, (unsigned)col_blocks), 320, 0, stream>>>(job->out, ors, job->out_layout, job->n_frames, M,
                                                                                             a.lens, a.total_len, a.frame0, colpart, a.out_offsets, hop, nfft, cpb);
        TALFE_CUDA(cudaGetLastError());
        colstats_finish_kernel<<<dim3((unsigned)job->batch, kMaxMels / kFinishMels), kFinishLanes * kFinishMels, 0, stream>>>(colpart, col_blocks, M, accumulate, stats);
        TALFE_CUDA(cudaGetLastError());
    }
    if (job->norm == TALFE_NORM_NONE || job->defer_normalise) return TALFE_OK;
    // rows keep their zero fill beyond their own length: the sweep derives valid frames from lens
    apply_stats_kernel<<<dim3((unsigned)job->batch, sweep_blocks(plan->sm_count, job->batch, dense, a.lens != nullptr)), 256, 0, stream>>>(
        job->out, job->batch, job->n_frames, ors, job->out_layout, M, job->norm, stats, nullptr, a.lens, a.frame0, a.out_offsets,
        job->freq_bands, job->time_bands, job->n_bands, hop, nfft);
    TALFE_CUDA(cudaGetLastError());
    return TALFE_OK;
}

int talfe_logmel_forward(const talfe_plan* plan, const void* wave, int wave_dtype, int64_t batch, int64_t n_samples,
                         int64_t row_stride, float* out, float eps, void* workspace, size_t workspace_bytes, void* stream) {
    if (!plan) return TALFE_ERR_INVALID;
    if (n_samples <= plan->n_fft / 2) return TALFE_ERR_TOO_SHORT;
    talfe_job job{};
    job.wave = wave; job.wave_dtype = wave_dtype; job.norm = TALFE_NORM_BATCH_MEAN;
    job.batch = batch; job.row_stride = row_stride; job.buf_len = n_samples; job.origin = 0; job.total_len = n_samples;
    job.lens = nullptr; job.frame0 = 0; job.n_frames = frames_of(n_samples, plan->hop, plan->n_fft);
    job.out = out; job.out_row_stride = 0; job.out_layout = TALFE_LAYOUT_TM; job.accumulate_stats = 0;
    job.eps = eps; job.defer_normalise = 0; job.stats = nullptr;
    job.workspace = workspace; job.workspace_bytes = workspace_bytes; job.stream = stream;
    return talfe_run(plan, &job);
}

int talfe_resample(const void* wave, int wave_dtype, int64_t batch, int64_t n_samples, int64_t row_stride, int orig, int new_,
                   int width, const float* kernel_dev, float* out, int64_t out_len, int64_t out_row_stride, void* stream) {
    if (!wave || !kernel_dev || !out || batch < 1 || batch > 65535 || n_samples < 1 || orig < 1 || new_ < 1 || width < 1) return TALFE_ERR_INVALID;
    if (wave_dtype < TALFE_F32 || wave_dtype > TALFE_I16 || row_stride < n_samples || out_row_stride < out_len) return TALFE_ERR_INVALID;
    if (out_len != (new_ * n_samples + orig - 1) / orig) return TALFE_ERR_INVALID;          // ceil(new L / orig), as torchaudio truncates
    const dim3 grid((unsigned)std::min<long long>((out_len + 255) / 256, 148 * 16), (unsigned)batch);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (wave_dtype == TALFE_F32)
        resample_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(wave), row_stride, n_samples, orig, new_, width, kernel_dev, out, out_len, out_row_stride, 1.0f);
    else if (wave_dtype == TALFE_F16)
        resample_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(wave), row_stride, n_samples, orig, new_, width, kernel_dev, out, out_len, out_row_stride, 1.0f);
    else
        resample_kernel<short><<<grid, 256, 0, st>>>(reinterpret_cast<const short*>(wave), row_stride, n_samples, orig, new_, width, kernel_dev, out, out_len, out_row_stride, 1.0f / 32768.0f);
    TALFE_CUDA(cudaGetLastError());
    return TALFE_OK;
}

int talfe_apply_stats(const talfe_plan* plan, float* feats, int64_t batch, int64_t n_frames, int64_t out_row_stride,
                      int out_layout, int norm, const double* stats, const int64_t* valid_frames, void* stream) {
    if (!plan || !feats || !stats || batch < 1 || n_frames < 1) return TALFE_ERR_INVALID;
    if (norm <= TALFE_NORM_NONE || norm > TALFE_NORM_ROW_MEL_MEANVAR) return TALFE_ERR_INVALID;
    if (out_layout != TALFE_LAYOUT_TM && out_layout != TALFE_LAYOUT_MT) return TALFE_ERR_INVALID;
    if (batch > 0x7fffffffLL) return TALFE_ERR_UNSUPPORTED;
    const int M = plan->n_mels;
    const long long dense = n_frames * M;
    const long long ors = out_row_stride ? out_row_stride : dense;
    if (ors < dense) return TALFE_ERR_INVALID;
    apply_stats_kernel<<<dim3((unsigned)batch, sweep_blocks(plan->sm_count, batch, dense, valid_frames != nullptr)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        feats, batch, n_frames, ors, out_layout, M, norm, stats, reinterpret_cast<const long long*>(valid_frames), nullptr, 0,
        nullptr, nullptr, nullptr, 0, plan->hop, plan->n_fft);
    TALFE_CUDA(cudaGetLastError());
    return TALFE_OK;
}

// ---- streaming of one long episode from (pinned) host memory, native loop: no per-chunk interpreter cost
size_t talfe_stream_staging_bytes(int wave_dtype, int64_t chunk_frames) {
    if (chunk_frames < 1 || wave_dtype < TALFE_F32 || wave_dtype > TALFE_I16) return 0;
    const size_t elt = wave_dtype == TALFE_F32 ? 4 : 2;
    const size_t per = align_up((size_t)(kHop * chunk_frames + kNfft + kHalf) * elt, 256);
    return 2 * per;
}

int talfe_stream_episode(const talfe_plan* plan, const void* wave_host, int wave_dtype, int64_t total_len,
                         int64_t chunk_frames, float* out, int norm, int defer_normalise, double* stats, float eps,
                         void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes, void* stream_v) {
    if (!plan || !wave_host || !out || !stats || chunk_frames < 1) return TALFE_ERR_INVALID;
    if (plan->generic) return TALFE_ERR_UNSUPPORTED;                   // the chunk arithmetic below is the 16 kHz geometry's
    if (total_len <= kHalf) return TALFE_ERR_TOO_SHORT;
    if (wave_dtype < TALFE_F32 || wave_dtype > TALFE_I16) return TALFE_ERR_INVALID;
    if (norm < TALFE_NORM_NONE || norm > TALFE_NORM_ROW_MEL_MEANVAR) return TALFE_ERR_INVALID;
    // an episode that already lives in device memory is transformed in place, chunk by chunk, with no staging copies
    cudaPointerAttributes pa{};
    bool on_device = false;
    if (cudaPointerGetAttributes(&pa, wave_host) == cudaSuccess) on_device = pa.type == cudaMemoryTypeDevice;
    else cudaGetLastError();
    if (!on_device && (!staging || staging_bytes < talfe_stream_staging_bytes(wave_dtype, chunk_frames))) return TALFE_ERR_WORKSPACE;
    const size_t elt = wave_dtype == TALFE_F32 ? 4 : 2;
    const size_t per = staging_bytes / 2 / 256 * 256;
    const int64_t T = 1 + total_len / kHop;
    const int M = plan->n_mels;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    talfe_plan* pl = const_cast<talfe_plan*>(plan);
    int k = 0;
    for (int64_t f0 = 0; f0 < T; f0 += chunk_frames, ++k) {
        const int64_t f1 = std::min<int64_t>(T, f0 + chunk_frames);
        // samples the frames span, widened where an edge reflects back into the episode
        int64_t lo = kHop * f0 - kHalf, hi = kHop * (f1 - 1) + kHalf, need_lo = lo, need_hi = hi;
        if (hi > total_len) need_lo = std::min(need_lo, 2 * (total_len - 1) - (hi - 1));
        if (lo < 0) need_hi = std::max(need_hi, -lo + 1);
        lo = std::max<int64_t>(0, need_lo);
        hi = std::min<int64_t>(total_len, need_hi);
        lo -= lo % 8;                                          // keep the chunk origin 16-byte aligned -> TMA fast path
        const int b = k & 1;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(wave_host) + lo * elt;
        const void* chunk = src;
        if (!on_device) {
            unsigned char* buf = reinterpret_cast<unsigned char*>(staging) + b * per;
            if ((size_t)(hi - lo) * elt > per) return TALFE_ERR_WORKSPACE;
            // staging buffer b is free once the transform that last read it has run (also orders against a previous episode;
            // waiting on a never-recorded event is a no-op)
            TALFE_CUDA(cudaStreamWaitEvent(pl->copy_stream, pl->ev_free[b], 0));
            TALFE_CUDA(cudaMemcpyAsync(buf, src, (size_t)(hi - lo) * elt, cudaMemcpyHostToDevice, pl->copy_stream));
            TALFE_CUDA(cudaEventRecord(pl->ev_ready[b], pl->copy_stream));
            TALFE_CUDA(cudaStreamWaitEvent(stream, pl->ev_ready[b], 0));
            chunk = buf;
        }
        talfe_job job{};
        job.wave = chunk; job.wave_dtype = wave_dtype; job.norm = norm; job.batch = 1; job.row_stride = hi - lo;
        job.buf_len = hi - lo; job.origin = lo; job.total_len = total_len; job.lens = nullptr;
        job.frame0 = f0; job.n_frames = f1 - f0; job.out = out + f0 * M; job.out_row_stride = 0;
        job.out_layout = TALFE_LAYOUT_TM; job.accumulate_stats = k > 0; job.eps = eps; job.defer_normalise = 1;
        job.stats = stats; job.workspace = workspace; job.workspace_bytes = workspace_bytes; job.stream = stream_v;
        const int rc = talfe_run(plan, &job);
        if (rc) return rc;
        if (!on_device) TALFE_CUDA(cudaEventRecord(pl->ev_free[b], stream));
    }
    // every copy out of the caller's host buffer has completed when this call returns (the transforms may still be
    // running): the caller may reuse or free `wave_host` at once — a pinned-memory allocator that only tracks copies
    // made through its own API cannot know about ours
    if (!on_device) TALFE_CUDA(cudaStreamSynchronize(pl->copy_stream));
    if (norm == TALFE_NORM_NONE || defer_normalise) return TALFE_OK;
    return talfe_apply_stats(plan, out, 1, T, 0, TALFE_LAYOUT_TM, norm, stats, nullptr, stream_v);
}

// ---- diagnostics: the FP32 pipe's measured FMA rate, the compute-side denominator of the roofline (SURVEY.md §8d asks
// for a measured figure; MEASURED_PEAKS.json has none).  Dependent chains of packed FFMA2, 8 chains per thread.
__global__ void __launch_bounds__(512) fp32_fma_probe_kernel(float2* out, int iters) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(1.0f + threadIdx.x * 1e-6f + i, 0.5f + i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("{ .reg .b64 ra, rm, rc; mov.b64 ra, {%0,%1}; mov.b64 rm, {%2,%3}; mov.b64 rc, {%4,%5}; fma.rn.f32x2 ra, ra, rm, rc; mov.b64 {%0,%1}, ra; }"
                         : "+f"(acc[i].x), "+f"(acc[i].y) : "f"(m.x), "f"(m.y), "f"(c.x), "f"(c.y));
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { r.x += acc[i].x; r.y += acc[i].y; }
    if (r.x == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;   // keeps the chains alive, practically never stores
}

int talfe_probe_fp32_fma_rate(int device, double* fma_per_second) {
    if (!fma_per_second) return TALFE_ERR_INVALID;
    int prev = 0, sms = 0;
    TALFE_CUDA(cudaGetDevice(&prev));
    TALFE_CUDA(cudaSetDevice(device));
    TALFE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float2* scratch = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int blocks = sms * 4, threads = 512, iters = 4096;
    cudaError_t e = cudaMalloc(&scratch, (size_t)blocks * threads * sizeof(float2));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0, 0);
        fp32_fma_probe_kernel<<<blocks, threads>>>(scratch, iters);
        cudaEventRecord(e1, 0);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess && rep > 0 && ms > 0.f)     // first repetition = warm-up
            best = std::max(best, (double)blocks * threads * iters * 8.0 * 2.0 / (ms * 1e-3));
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (scratch) cudaFree(scratch);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e);
    *fma_per_second = best;
    return TALFE_OK;
}

// ---- NCCL, resolved at run time so that the library has no link-time dependency on it
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static nccl_allreduce_fn g_nccl_allreduce = nullptr;
static std::once_flag g_nccl_once;

int talfe_allreduce_stats(double* stats_dev, int64_t count, void* nccl_comm, void* stream) {
    if (!stats_dev || count < 1 || !nccl_comm) return TALFE_ERR_INVALID;
    std::call_once(g_nccl_once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) g_nccl_allreduce = reinterpret_cast<nccl_allreduce_fn>(dlsym(h, "ncclAllReduce"));
    });
    if (!g_nccl_allreduce) return TALFE_ERR_NCCL;
    // ncclDouble = 8 (ncclFloat64), ncclSum = 0 in every NCCL 2.x release
    const int rc = g_nccl_allreduce(stats_dev, stats_dev, (size_t)count, 8, 0, nccl_comm, reinterpret_cast<cudaStream_t>(stream));
    return rc == 0 ? TALFE_OK : TALFE_ERR_NCCL;
}

int talfe_detect_padding(const void* wave, int wave_dtype, int64_t batch, int64_t n_samples, int64_t row_stride, int64_t* lens, void* stream) {
    if (!wave || !lens || batch < 1 || batch > 0x7fffffffLL || n_samples < 1 || row_stride < n_samples) return TALFE_ERR_INVALID;
    if (wave_dtype < TALFE_F32 || wave_dtype > TALFE_I16) return TALFE_ERR_INVALID;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long* out = reinterpret_cast<unsigned long long*>(lens);
    TALFE_CUDA(cudaMemsetAsync(out, 0, (size_t)batch * sizeof(unsigned long long), st));
    const dim3 grid((unsigned)batch, kPadSegments);
    if (wave_dtype == TALFE_F32) detect_padding_kernel<float><<<grid, kPadThreads, 0, st>>>(reinterpret_cast<const float*>(wave), row_stride, n_samples, out);
    else if (wave_dtype == TALFE_F16) detect_padding_kernel<__half><<<grid, kPadThreads, 0, st>>>(reinterpret_cast<const __half*>(wave), row_stride, n_samples, out);
    else detect_padding_kernel<short><<<grid, kPadThreads, 0, st>>>(reinterpret_cast<const short*>(wave), row_stride, n_samples, out);
    TALFE_CUDA(cudaGetLastError());
    return TALFE_OK;
}

int talfe_synth_fill(void* wave, int wave_dtype, int64_t rows, int64_t n_samples, int64_t row_stride, uint64_t seed,
                     int64_t first_episode, int64_t start, void* stream) {
    if (!wave || rows < 1 || n_samples < 1 || row_stride < n_samples) return TALFE_ERR_INVALID;
    if (wave_dtype < TALFE_F32 || wave_dtype > TALFE_I16) return TALFE_ERR_INVALID;
    const long long total = rows * n_samples;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
    synth_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(wave, wave_dtype, rows, n_samples, row_stride, seed,
                                                                         first_episode, start);
    TALFE_CUDA(cudaGetLastError());
    return TALFE_OK;
}

}  // extern "C"
