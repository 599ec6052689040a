// host_emul.cu — CPU emulator of ONE 20-thread group of the fused kernel (TEST ONLY).
//
// Drives the very same __host__ __device__ stage functions the kernel inlines (talfe_core.cuh),
// one emulated "thread" at a time with plain arrays standing in for shared memory, so that index
// maps, table layouts and arithmetic can be checked against the oracle on a machine without a
// GPU.  Built into tests/_build/libtalfe_emul.so by tests/test_host_emul.py; it is not linked
// into, loaded by, or reachable from the product library.
#include <cstdint>
#include <vector>

#include "talfe_core.cuh"
#include "talfe_tables.h"

using namespace talfe;

extern "C" int talfe_emul_logmel(const float* x, int64_t n_samples, int n_mels, const float* window,
                                 const float* fb, float eps, int force_generic, float* out /* [T][n_mels] un-normalised */) {
    if (n_samples <= kHalf) return -2;
    std::vector<float> win(kNfft), fbv;
    if (window) win.assign(window, window + kNfft); else default_window(win.data());
    if (fb) fbv.assign(fb, fb + kBins * n_mels); else { fbv.resize(kBins * n_mels); default_filterbank(n_mels, fbv.data()); }
    HostTables t;
    int rc = build_tables(n_mels, win.data(), fbv.data(), t);
    if (rc) return rc;
    const int64_t T = 1 + n_samples / kHop;
    std::vector<float> xs(xskew<float>(kNfft + kHop - 1) + 1);
    std::vector<cf> e(kEGroup), p2(t.pstride);
    const cf* tw = reinterpret_cast<const cf*>(t.tw_t.data());
    for (int64_t t0 = 0; t0 < T; t0 += 2) {
        for (int i = 0; i < kNfft + kHop; ++i) {
            int64_t g = kHop * t0 - kHalf + i;
            if (g < 0) g = -g;
            if (g >= n_samples) g = 2 * (n_samples - 1) - g;
            xs[xskew<float>(i)] = (g >= 0 && g < n_samples) ? x[g] : 0.f;
        }
        for (auto& v : e) v = make_float2(0.f, 0.f);
        for (auto& v : p2) v = make_float2(0.f, 0.f);
        for (int j = 0; j < 20; ++j) {
            float win[20];
            load_window(j, t.win_t.data(), 1.0f, win);
            stage1(j, xs.data(), win, tw, e.data());
        }
        for (int row = 0; row < 20; ++row) {
            cf v[20];
            stage2_load(e.data() + row_slot(row) * kERow, v);
            if (row >= 18) stage2_special(row, v, reinterpret_cast<float*>(p2.data()));
            else stage2_normal(row, v, reinterpret_cast<float*>(p2.data()));
        }
        for (int c = 0; c < 20; ++c) {
            float y[2 * kMelSlots];
            if (is_reference_layout(t.layout) && !force_generic) {
                const int lo[kMelSlots] = {t.mel_lo[c], t.mel_lo[c + 20], t.mel_lo[c + 40], t.mel_lo[c + 60]};   // slot-major
                mel_log_ref(c, p2.data(), t.w_t.data(), lo, eps, y);
            } else {
                mel_log_generic(c, t.layout, p2.data(), t.w_t.data(), t.mel_lo.data(), eps, y);
            }
            for (int i = 0; i < t.layout.n_slots; ++i) {
                const int m = t.mel_id[i * 20 + c];
                if (m < 0) continue;
                out[t0 * n_mels + m] = y[2 * i];
                if (t0 + 1 < T) out[(t0 + 1) * n_mels + m] = y[2 * i + 1];
            }
        }
    }
    return 0;
}

// raw 20-point DFT for a direct unit test of the prime-factor index maps
extern "C" void talfe_emul_fft20(float* reim /* 40 floats, interleaved, in place */) {
    cf v[20];
    for (int i = 0; i < 20; ++i) v[i] = make_float2(reim[2 * i], reim[2 * i + 1]);
    fft20(v);
    for (int i = 0; i < 20; ++i) { reim[2 * i] = v[i].x; reim[2 * i + 1] = v[i].y; }
}
