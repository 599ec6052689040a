// talfe_tables.h — host-side construction of the constant tables the kernel keeps in shared memory.
// Used by talfe_plan_create (product) and by the CPU emulator test (csrc/host_emul.cu).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "talfe_core.cuh"

namespace talfe {

// torch.hann_window(400, periodic=True): 0.5 - 0.5 cos(2 pi n / 400)   (reference buffer
// mel_transform.spectrogram.window, /root/reference/tal/asr/models.py:24-32 via torchaudio)
inline void default_window(float* w) {
    for (int n = 0; n < kNfft; ++n) w[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / kNfft));
}

// torchaudio melscale_fbanks(201, 0, 8000, n_mels, 16000, norm=None, mel_scale="htk")
// (reference buffer mel_transform.mel_scale.fb).  Computed in double here; the Python host passes
// the fp32-evaluated table instead so that weights are bit-identical to the reference's.
inline void default_filterbank(int n_mels, float* fb /* [201][n_mels] */) {
    const double top = 2595.0 * std::log10(1.0 + 8000.0 / 700.0);
    std::vector<double> pts(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) pts[i] = 700.0 * (std::pow(10.0, (top * i / (n_mels + 1)) / 2595.0) - 1.0);
    for (int f = 0; f < kBins; ++f) {
        const double hz = 8000.0 * f / (kBins - 1);
        for (int m = 0; m < n_mels; ++m) {
            const double up = (hz - pts[m]) / (pts[m + 1] - pts[m]);
            const double down = (pts[m + 2] - hz) / (pts[m + 2] - pts[m + 1]);
            fb[f * n_mels + m] = (float)std::fmax(0.0, std::fmin(up, down));
        }
    }
    // the band edges 0 Hz and 8000 Hz carry exactly zero weight; keep rounding residue out of them
    for (int m = 0; m < n_mels; ++m) fb[0 * n_mels + m] = fb[(kBins - 1) * n_mels + m] = 0.f;
}

struct HostTables {
    MelLayout layout;
    int pstride;                    // float2 entries per pair in the power array
    std::vector<float> win_t;       // [20][20]  0.5 * window[j + 20 m]
    std::vector<float> tw_t;        // [20][10] complex: 2 W400^(j k1), k1 = 1..9, and W400^(10 j)
    std::vector<float> w_t;         // [20][wstride]
    std::vector<int> mel_lo;        // [4][20] slot-major: first bin of the mel owned by (slot i, lane c)
    std::vector<int> mel_id;        // [4][20] slot-major: which mel (slot i, lane c) owns, -1 = none
    // warp-specialised kernel (talfe_ws.cuh): lane c owns mels c, 20 + c, 40 + c, 60 + c (no permutation: its
    // power reads are conflict-free for any ownership), weights live in the consumer threads' registers
    std::vector<float> w_ws;        // [20][wstride]
    std::vector<int> lo_ws;         // [4][20] slot-major first bins
    // section staged by the warp-specialised kernel: [twiddles | w_ws | lo_ws] starting at off_ws (its window taps are
    // read once from global memory into registers); off_tw_ws / off_w_ws / off_lo_ws are relative to off_ws
    size_t off_ws, ws_bytes, off_tw_ws, off_w_ws, off_lo_ws;
    // byte offsets inside the blob that is copied to shared memory
    size_t off_win, off_tw, off_w, off_lo, off_id, blob_bytes;
    std::vector<unsigned char> blob;
};

// returns 0, or -3 (TALFE_ERR_UNSUPPORTED)
inline int build_tables(int n_mels, const float* window, const float* fb, HostTables& t) {
    if (n_mels < 1 || n_mels > kMaxMels) return -3;
    t.win_t.assign(400, 0.f);
    for (int j = 0; j < 20; ++j)
        for (int m = 0; m < 20; ++m) t.win_t[j * 20 + m] = 0.5f * window[j + 20 * m];
    t.tw_t.assign(400, 0.f);
    for (int j = 0; j < 20; ++j)
        for (int k1 = 1; k1 <= 10; ++k1) {
            const double ang = -2.0 * M_PI * (double)((j * k1) % kNfft) / kNfft;
            const double s = k1 < 10 ? 2.0 : 1.0;              // the transform runs on 0.5 * window: rows 0..17 get their factor 2 here
            t.tw_t[(j * 10 + k1 - 1) * 2 + 0] = (float)(s * std::cos(ang));
            t.tw_t[(j * 10 + k1 - 1) * 2 + 1] = (float)(s * std::sin(ang));
        }
    // mel supports
    for (int m = 0; m < n_mels; ++m)
        if (fb[0 * n_mels + m] != 0.f || fb[200 * n_mels + m] != 0.f) return -3;   // bins 0 / 200 are not computed
    MelLayout& L = t.layout;
    L.n_mels = n_mels;
    L.n_slots = (n_mels + 19) / 20;
    std::vector<int> lo(kMaxMels, 1), hi(kMaxMels, 0);
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int f = 1; f < 200; ++f)
            if (fb[f * n_mels + m] != 0.f) { if (first < 0) first = f; last = f; }
        if (first < 0) { first = 1; last = 0; }                                   // empty filter -> log(eps)
        lo[m] = first; hi[m] = last;
    }
    int total = 0, max_w = 1;
    for (int i = 0; i < kMelSlots; ++i) {
        int w = 0;
        for (int c = 0; c < 20; ++c) {
            const int m = c + 20 * i;
            if (m < n_mels) w = std::max(w, hi[m] - lo[m] + 1);
        }
        if (i < L.n_slots && w < 1) w = 1;
        L.width[i] = i < L.n_slots ? w : 0;
        L.offset[i] = total;
        total += L.width[i];
        max_w = std::max(max_w, L.width[i]);
    }
    L.wstride = (total + 3) & ~3;
    if (L.wstride > 256) return -3;
    t.pstride = 201 + max_w;
    while (t.pstride % 16 != 9) ++t.pstride;                                       // 2*pstride = 18 (mod 32) words
    // Which lane owns which mel inside a slot is free; choose it to minimise shared-memory bank
    // conflicts of the mel-stage reads (LDS.64 of the power pairs, 16 lanes per wavefront, 16 banks of
    // 8 bytes; lanes of a half-warp collide when their first bins are congruent mod 16).  Deterministic
    // pairwise-swap descent from the identity; model and numbers in DESIGN.md §4 / tools/bank_search.py.
    std::vector<int> owner(kMaxMels, -1);                      // owner[i*20 + c] = mel
    for (int m = 0; m < n_mels; ++m) owner[(m / 20) * 20 + m % 20] = m;
    const int gshift = t.pstride % 16;
    auto read_cost = [&]() {
        int total = 0;
        for (int i = 0; i < L.n_slots; ++i)
            for (int r = 0; r < L.width[i]; ++r)
                for (int hw = 0; hw < 5; ++hw) {               // the 5 ways 16 consecutive threads straddle 20-thread groups
                    int addr[16], nb = 0, worst = 1;
                    for (int tt = 16 * hw; tt < 16 * hw + 16; ++tt) {
                        const int g = tt / 20, c = tt % 20, m = owner[i * 20 + c];
                        if (m < 0) continue;
                        addr[nb++] = lo[m] + r + g * (16 * 1000 + gshift);   // distinct groups never share an address
                    }
                    for (int x = 0; x < nb; ++x) {
                        int mult = 1;
                        for (int y = 0; y < x; ++y)
                            if (addr[y] != addr[x] && (addr[y] - addr[x]) % 16 == 0) {
                                bool dup = false;                            // count distinct addresses only
                                for (int z = 0; z < y; ++z) if (addr[z] == addr[y]) dup = true;
                                if (!dup) ++mult;
                            }
                        worst = std::max(worst, mult);
                    }
                    total += worst;
                }
        return total;
    };
    int cost = read_cost();
    for (bool improved = true; improved;) {
        improved = false;
        for (int i = 0; i < L.n_slots; ++i)
            for (int a = 0; a < 20; ++a)
                for (int b = a + 1; b < 20; ++b) {
                    std::swap(owner[i * 20 + a], owner[i * 20 + b]);
                    const int c2 = read_cost();
                    if (c2 < cost) { cost = c2; improved = true; }
                    else std::swap(owner[i * 20 + a], owner[i * 20 + b]);
                }
    }
    t.w_t.assign(20 * L.wstride, 0.f);
    t.mel_lo.assign(kMaxMels, 1);
    t.mel_id.assign(kMaxMels, -1);
    for (int i = 0; i < L.n_slots; ++i)
        for (int c = 0; c < 20; ++c) {
            const int m = owner[i * 20 + c];
            t.mel_id[i * 20 + c] = m;
            if (m < 0) continue;
            t.mel_lo[i * 20 + c] = lo[m];
            for (int f = lo[m]; f <= hi[m]; ++f) t.w_t[c * L.wstride + L.offset[i] + (f - lo[m])] = fb[f * n_mels + m];
        }
    t.w_ws.assign(20 * L.wstride, 0.f);
    t.lo_ws.assign(kMaxMels, 1);
    for (int i = 0; i < L.n_slots; ++i)
        for (int c = 0; c < 20; ++c) {
            const int m = 20 * i + c;
            if (m >= n_mels) continue;
            t.lo_ws[i * 20 + c] = lo[m];
            for (int f = lo[m]; f <= hi[m]; ++f) t.w_ws[c * L.wstride + L.offset[i] + (f - lo[m])] = fb[f * n_mels + m];
        }
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    t.off_win = 0;
    t.off_tw = align16(t.off_win + 400 * sizeof(float));
    t.off_w = align16(t.off_tw + 400 * sizeof(float));
    t.off_lo = align16(t.off_w + t.w_t.size() * sizeof(float));
    t.off_id = align16(t.off_lo + kMaxMels * sizeof(int));
    t.off_ws = align16(t.off_id + kMaxMels * sizeof(int));
    t.off_tw_ws = 0;
    t.off_w_ws = align16(t.off_tw_ws + 400 * sizeof(float));
    t.off_lo_ws = align16(t.off_w_ws + t.w_ws.size() * sizeof(float));
    t.ws_bytes = align16(t.off_lo_ws + kMaxMels * sizeof(int));
    t.blob_bytes = t.off_ws + t.ws_bytes;
    t.blob.assign(t.blob_bytes, 0);
    std::memcpy(t.blob.data() + t.off_win, t.win_t.data(), 400 * sizeof(float));
    std::memcpy(t.blob.data() + t.off_tw, t.tw_t.data(), 400 * sizeof(float));
    std::memcpy(t.blob.data() + t.off_w, t.w_t.data(), t.w_t.size() * sizeof(float));
    std::memcpy(t.blob.data() + t.off_lo, t.mel_lo.data(), kMaxMels * sizeof(int));
    std::memcpy(t.blob.data() + t.off_id, t.mel_id.data(), kMaxMels * sizeof(int));
    std::memcpy(t.blob.data() + t.off_ws + t.off_tw_ws, t.tw_t.data(), 400 * sizeof(float));
    std::memcpy(t.blob.data() + t.off_ws + t.off_w_ws, t.w_ws.data(), t.w_ws.size() * sizeof(float));
    std::memcpy(t.blob.data() + t.off_ws + t.off_lo_ws, t.lo_ws.data(), kMaxMels * sizeof(int));
    return 0;
}

}  // namespace talfe
