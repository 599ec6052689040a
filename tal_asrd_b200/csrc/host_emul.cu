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

// Warp-specialised path (csrc/talfe_ws.cuh): one whole 32-frame tile at a time, every producer thread then every
// consumer thread, through the very same stage functions and the same E / P / Y layouts the kernel uses.
extern "C" int talfe_emul_logmel_ws(const float* x, int64_t n_samples, const float* window, const float* fb, float eps,
                                    float* out /* [T][80] un-normalised */) {
    const int n_mels = 80;
    if (n_samples <= kHalf) return -2;
    std::vector<float> win(kNfft), fbv;
    if (window) win.assign(window, window + kNfft); else default_window(win.data());
    if (fb) fbv.assign(fb, fb + kBins * n_mels); else { fbv.resize(kBins * n_mels); default_filterbank(n_mels, fbv.data()); }
    HostTables t;
    int rc = build_tables(n_mels, win.data(), fbv.data(), t);
    if (rc) return rc;
    if (!is_reference_layout(t.layout)) return -3;
    const int64_t T = 1 + n_samples / kHop;
    const int tile_samples = kHop * kWsFrames + (kNfft - kHop);
    std::vector<float> xs(xskew<float>(tile_samples - 1) + 1);
    std::vector<cf> e(kWsECf), p(kWsPCf, make_float2(0.f, 0.f));
    std::vector<float> y(kWsYFloats);
    const cf* tw_all = reinterpret_cast<const cf*>(t.tw_t.data());
    for (int64_t t0 = 0; t0 < T; t0 += kWsFrames) {
        for (int i = 0; i < tile_samples; ++i) {
            int64_t g = kHop * t0 - kHalf + i;
            if (g < 0) g = -g;
            if (g >= n_samples) g = 2 * (n_samples - 1) - g;
            xs[xskew<float>(i)] = (g >= 0 && g < n_samples) ? x[g] : 0.f;
        }
        for (auto& v : e) v = make_float2(NAN, NAN);                   // every entry that is read must have been written
        for (int tid = 0; tid < kWsGroups * kGroup; ++tid) {           // producers: thread (g1, j)
            const int g1 = tid / kGroup, j = tid % kGroup;
            float wj[20];
            load_window(j, t.win_t.data(), 1.0f, wj);
            cf tw[10];
            for (int q = 0; q < 10; ++q) tw[q] = tw_all[j * 10 + q];
            cf re[11], im[11];
            stage1_ws_fft<float>(xs.data() + XLayout<float>::kGroup * g1 + j, wj, re, im);
            stage1_ws_store(re, im, tw, e.data() + ws_e_base(g1) + j);
        }
        for (int tid = 0; tid < kWsGroups * kGroup; ++tid) {           // consumers, stage 2: thread (g, r)
            const int g = tid & (kWsGroups - 1), r = tid >> 4;
            cf v[20], pw[10];
            stage2_load(e.data() + ws_e_base(g) + r * kWsERow, v);
            if (r < 18) {
                stage2_ws_power_normal(v, pw);
                stage2_ws_store_normal(1 + (r >> 1), pw, reinterpret_cast<float*>(p.data()) + 2 * g + (r & 1));
            } else {
                stage2_ws_power_special(r == 18, v, pw);
                stage2_ws_store_special(r == 18, pw, p.data() + g);
            }
        }
        for (int tid = 0; tid < kWsGroups * kGroup; ++tid) {           // consumers, mel stage: thread (g, c)
            const int g = tid & (kWsGroups - 1), c = tid >> 4;
            float w[kRefWStride];
            for (int q = 0; q < kRefWStride; ++q) w[q] = t.w_ws[c * kRefWStride + q];
            const int lo[kMelSlots] = {t.lo_ws[c], t.lo_ws[20 + c], t.lo_ws[40 + c], t.lo_ws[60 + c]};
            float yy[2 * kMelSlots];
            mel_log_ws(p.data() + g, w, lo, eps, yy);
            float* yb = y.data() + ws_y_off(2 * g) + c;
            for (int i = 0; i < kMelSlots; ++i) { yb[20 * i] = yy[2 * i]; yb[kMaxMels + 20 * i] = yy[2 * i + 1]; }
        }
        for (int f = 0; f < kWsFrames && t0 + f < T; ++f)
            for (int m = 0; m < n_mels; ++m) out[(t0 + f) * n_mels + m] = y[ws_y_off(f) + m];
    }
    return 0;
}

// Shared-memory address (in units of the access width: 4-byte words for kinds 0, 3, 6; 8-byte for 1, 4, 5; 16-byte for 2)
// that role thread `tid` touches with its `idx`-th access of the given kind — the same expressions the stage
// functions above use, exported so that tests/test_ws_layout.py can count bank conflicts per warp instruction.
extern "C" int64_t talfe_emul_ws_addr(int kind, int tid, int idx, int idx2) {
    HostTables t;
    static std::vector<float> win(kNfft), fbv(kBins * 80);
    static bool init = false;
    static HostTables tt;
    if (!init) { default_window(win.data()); default_filterbank(80, fbv.data()); build_tables(80, win.data(), fbv.data(), tt); init = true; }
    const int g1 = tid / kGroup, j = tid % kGroup;                    // producer view
    const int g = tid & (kWsGroups - 1), r = tid >> 4;                 // consumer view
    switch (kind) {
        case 0: {   // x load: sample j + 20 idx of the pair (idx 0..27)
            const int i = 20 * idx;
            return XLayout<float>::kGroup * g1 + j + i + (i >= kXBlock ? XLayout<float>::kSkew : 0);
        }
        case 1: return ws_e_base(g1) + j + idx * kWsERow;                                        // E store, row idx
        case 2: return (ws_e_base(g) + r * kWsERow) / 2 + idx;                                   // E load, 16-byte chunk idx 0..9
        case 3: {   // P store, normal row: idx 0..9 -> lo bins, 10..19 -> hi bins
            const int k1 = 1 + (r >> 1), q = idx % 10;
            const int bin = (idx < 10 ? k1 : 20 - k1) + 20 * q;
            return 2 * (kWsGroups * bin + g) + (r & 1);
        }
        case 4: return kWsGroups * ((r == 18 ? 20 : 10) + 20 * idx) + g;                         // P store, packed rows
        case 5: return kWsGroups * (tt.lo_ws[idx * 20 + r] + idx2) + g;                          // mel load: slot idx, tap idx2
        case 6: return ws_y_off(2 * g) + r + 20 * idx + kMaxMels * idx2;                         // Y store: slot idx, frame idx2
        default: return -1;
    }
}
