// talfe_tables.h — host-side construction of the constant tables the kernel keeps in shared memory.
// Used by talfe_plan_create (product) and by the CPU emulator test (csrc/host_emul.cu).
#pragma once

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
    std::vector<float> tw_t;        // [20][10] complex: W400^(j k1) k1=1..9, 2 W400^(10 j)
    std::vector<float> w_t;         // [20][wstride]
    std::vector<int> mel_lo;        // [80]
    // byte offsets inside the blob that is copied to shared memory
    size_t off_win, off_tw, off_w, off_lo, blob_bytes;
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
            const double s = k1 == 10 ? 2.0 : 1.0;
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
    t.w_t.assign(20 * L.wstride, 0.f);
    t.mel_lo.assign(kMaxMels, 1);
    for (int m = 0; m < n_mels; ++m) {
        const int c = m % 20, i = m / 20;
        t.mel_lo[m] = lo[m];
        for (int f = lo[m]; f <= hi[m]; ++f) t.w_t[c * L.wstride + L.offset[i] + (f - lo[m])] = fb[f * n_mels + m];
    }
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    t.off_win = 0;
    t.off_tw = align16(t.off_win + 400 * sizeof(float));
    t.off_w = align16(t.off_tw + 400 * sizeof(float));
    t.off_lo = align16(t.off_w + t.w_t.size() * sizeof(float));
    t.blob_bytes = align16(t.off_lo + kMaxMels * sizeof(int));
    t.blob.assign(t.blob_bytes, 0);
    std::memcpy(t.blob.data() + t.off_win, t.win_t.data(), 400 * sizeof(float));
    std::memcpy(t.blob.data() + t.off_tw, t.tw_t.data(), 400 * sizeof(float));
    std::memcpy(t.blob.data() + t.off_w, t.w_t.data(), t.w_t.size() * sizeof(float));
    std::memcpy(t.blob.data() + t.off_lo, t.mel_lo.data(), kMaxMels * sizeof(int));
    return 0;
}

}  // namespace talfe
