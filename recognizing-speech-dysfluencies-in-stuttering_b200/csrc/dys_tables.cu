// Host-side construction (float64, then rounded exactly where the reference's libraries
// round) and per-device upload of the front-end's lookup tables.
//
// Restates, in C++, the table builders the reference calls implicitly on every clip:
//   librosa.filters.mel(sr=16000, n_fft=2048)               (pipeline1.py:216 via mfcc)
//   scipy.fftpack.dct(type=2, norm='ortho')[:20]            (pipeline1.py:216)
//   librosa.filters.chroma(sr, n_fft, tuning) x 100 tunings (pipeline1.py:227 via chroma_stft)
//   scipy.signal.get_window('hann', N, fftbins=True)        (librosa.stft / istft)
//   noisereduce _smoothing_filter / get_time_smoothed_representation (pipeline1.py:140)
#include "dys_tables.h"

#include <algorithm>
#include <cmath>
#include <mutex>

#include "dys_error.h"

namespace dys {

namespace {

constexpr double kPi = 3.14159265358979323846;

std::vector<double> linspace(double start, double stop, int num, bool endpoint = true) {
    // numpy.linspace: y = arange(num) * step + start, last element forced to stop
    std::vector<double> y(num);
    const int div = endpoint ? num - 1 : num;
    const double step = (stop - start) / div;
    for (int i = 0; i < num; ++i) y[i] = i * step + start;
    if (endpoint && num > 1) y[num - 1] = stop;
    return y;
}

double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3;
    const double min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3;
    const double min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

void build_mel(HostTables& t) {
    const double val = 1.0 / (kNfft * (1.0 / kSR));          // numpy.fft.rfftfreq
    std::vector<double> fftfreqs(kBins);
    for (int k = 0; k < kBins; ++k) fftfreqs[k] = k * val;
    std::vector<double> mel_pts = linspace(hz_to_mel(0.0), hz_to_mel(kSR / 2.0), kMels + 2);
    std::vector<double> mel_f(kMels + 2);
    for (int i = 0; i < kMels + 2; ++i) mel_f[i] = mel_to_hz(mel_pts[i]);
    t.mel_dense.assign(size_t(kMels) * kBins, 0.f);
    t.mel_start.assign(kMels, 0);
    t.mel_len.assign(kMels, 0);
    t.mel_ptr.assign(kMels, 0);
    t.mel_w.clear();
    for (int i = 0; i < kMels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k) {
            const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
            const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
            const float w32 = float(std::fmax(0.0, std::fmin(lower, upper)));    // weights[i] = ... (float32)
            const float w = float(double(w32) * enorm);                          // weights *= enorm
            t.mel_dense[size_t(i) * kBins + k] = w;
            if (w != 0.f) { if (first < 0) first = k; last = k; }
        }
        t.mel_ptr[i] = int(t.mel_w.size());
        if (first >= 0) {
            t.mel_start[i] = first;
            t.mel_len[i] = last - first + 1;
            for (int k = first; k <= last; ++k) t.mel_w.push_back(t.mel_dense[size_t(i) * kBins + k]);
        }
    }
}

// Lane-per-filter layout of the sparse mel weights.  Filter f = lane + 32 g reads P[start + j], j < len, all 32 lanes at
// the same j: the reads are free of bank conflicts iff the 32 starts are distinct mod 32.  A filter may start up to
// `shift` bins early (leading zero weights change nothing), so each group picks shifts that make the residues distinct
// while keeping the longest padded filter as short as possible (bottleneck assignment, augmenting paths): for the
// slaney bank this costs no extra step at all (6 + 10 + 22 + 48 steps, as without padding).
bool assign_residues(const std::vector<int>& st, const std::vector<int>& ln, int thr, int (&res)[32]) {
    int owner[32];
    std::fill(owner, owner + 32, -1);
    auto fits = [&](int u, int r) { const int d = ((st[u] - r) % 32 + 32) % 32; return d <= st[u] && ln[u] + d <= thr; };
    struct Aug {
        decltype(fits)& ok; int* owner; bool seen[32];
        bool run(int u) {
            for (int r = 0; r < 32; ++r)
                if (ok(u, r) && !seen[r]) {
                    seen[r] = true;
                    if (owner[r] < 0 || run(owner[r])) { owner[r] = u; return true; }
                }
            return false;
        }
    } aug{fits, owner, {}};
    for (int u = 0; u < 32; ++u) {
        std::fill(aug.seen, aug.seen + 32, false);
        if (!aug.run(u)) return false;
    }
    for (int r = 0; r < 32; ++r) res[owner[r]] = r;
    return true;
}

void build_mel_step_major(HostTables& t) {
    t.mel_rstart = t.mel_start;
    t.mel_rlen = t.mel_len;
    int off = 0;
    for (int g = 0; g < 4; ++g) {
        std::vector<int> st(t.mel_start.begin() + 32 * g, t.mel_start.begin() + 32 * g + 32);
        std::vector<int> ln(t.mel_len.begin() + 32 * g, t.mel_len.begin() + 32 * g + 32);
        int L = *std::max_element(ln.begin(), ln.end());
        int res[32];
        while (!assign_residues(st, ln, L, res)) ++L;          // terminates: L = max len + 31 always fits bins >= 31
        for (int l = 0; l < 32; ++l) {
            const int d = ((st[l] - res[l]) % 32 + 32) % 32;
            t.mel_rstart[32 * g + l] = st[l] - d;
            t.mel_rlen[32 * g + l] = ln[l] + d;
        }
        t.mel_goff[g] = off;
        off += 32 * L;
    }
    t.mel_wt.assign(size_t(off), 0.f);
    for (int f = 0; f < kMels; ++f) {
        const int d = t.mel_start[f] - t.mel_rstart[f];
        for (int j = 0; j < t.mel_len[f]; ++j) t.mel_wt[t.mel_goff[f / 32] + 32 * (j + d) + f % 32] = t.mel_w[t.mel_ptr[f] + j];
    }
}

void build_chroma(HostTables& t) {
    t.chroma.assign(size_t(kTunings) * kBins * kChroma, 0.f);
    t.tuning_edges = linspace(-0.5, 0.5, kTunings + 1);
    std::vector<double> freqs = linspace(0.0, double(kSR), kNfft, /*endpoint=*/false);
    std::vector<double> frq(kNfft), bw(kNfft), col(kChroma);
    for (int ti = 0; ti < kTunings; ++ti) {
        const double tuning = t.tuning_edges[ti];
        const double a440 = 440.0 * std::pow(2.0, tuning / kChroma);
        for (int k = 1; k < kNfft; ++k) frq[k] = kChroma * std::log2(freqs[k] / (a440 / 16.0));
        frq[0] = frq[1] - 1.5 * kChroma;
        for (int k = 0; k + 1 < kNfft; ++k) bw[k] = std::fmax(frq[k + 1] - frq[k], 1.0);
        bw[kNfft - 1] = 1.0;
        for (int k = 0; k < kBins; ++k) {
            double ss = 0.0;
            for (int c = 0; c < kChroma; ++c) {
                double D = frq[k] - double(c);
                D = std::fmod(D + 6.0 + 10.0 * kChroma, double(kChroma)) - 6.0;
                const double z = 2.0 * D / bw[k];
                col[c] = std::exp(-0.5 * (z * z));
                ss += col[c] * col[c];
            }
            const double len = std::sqrt(ss);
            const double oct = (frq[k] / kChroma - 5.0) / 2.0;
            const double scale = std::exp(-0.5 * (oct * oct));
            for (int c = 0; c < kChroma; ++c) {
                double w = col[c] / len;
                w *= scale;
                const int row = (c - 3 + kChroma) % kChroma;            // np.roll(wts, -3, axis=0)
                t.chroma[(size_t(ti) * kBins + k) * kChroma + row] = float(w);
            }
        }
    }
}

std::vector<double> hann_periodic(int n) {
    std::vector<double> w(n);
    for (int k = 0; k < n; ++k) w[k] = 0.5 - 0.5 * std::cos(2.0 * kPi * k / n);
    return w;
}

std::vector<double> tri(int n) {
    // concatenate(linspace(0,1,n+1,endpoint=False), linspace(1,0,n+2))[1:-1]
    std::vector<double> up = linspace(0.0, 1.0, n + 1, false), down = linspace(1.0, 0.0, n + 2);
    std::vector<double> all(up);
    all.insert(all.end(), down.begin(), down.end());
    return std::vector<double>(all.begin() + 1, all.end() - 1);
}

void build(HostTables& t) {
    // windows
    std::vector<double> h2048 = hann_periodic(kNfft);
    t.hann2048.resize(kNfft);
    for (int k = 0; k < kNfft; ++k) t.hann2048[k] = float(h2048[k]);
    t.hann1024 = hann_periodic(kNrFft);
    // twiddles: exact quadrant symmetry is not needed, only correct rounding of each entry
    t.tw1024.resize(32 * 32);
    for (int kA = 0; kA < 32; ++kA)
        for (int l = 0; l < 32; ++l) {
            const double a = 2.0 * kPi * ((l * kA) % 1024) / 1024.0;
            t.tw1024[kA * 32 + l] = make_float2(float(std::cos(a)), float(-std::sin(a)));
        }
    t.split2048.resize(1024);
    for (int k = 0; k < 1024; ++k) {
        const double a = 2.0 * kPi * k / 2048.0;
        t.split2048[k] = make_float2(float(std::cos(a)), float(std::sin(a)));
    }
    t.tw512.resize(16 * 32);
    for (int kA = 0; kA < 16; ++kA)
        for (int l = 0; l < 32; ++l) {
            const double a = 2.0 * kPi * ((l * kA) % 512) / 512.0;
            t.tw512[kA * 32 + l] = make_double2(std::cos(a), -std::sin(a));
        }
    t.split1024.resize(512);
    for (int k = 0; k < 512; ++k) {
        const double a = 2.0 * kPi * k / 1024.0;
        t.split1024[k] = make_double2(std::cos(a), std::sin(a));
    }
    build_mel(t);
    build_mel_step_major(t);
    // ortho DCT-II rows 0..19 over 128 mel bands
    t.dct.resize(size_t(kMfcc) * kMels);
    for (int k = 0; k < kMfcc; ++k)
        for (int m = 0; m < kMels; ++m) {
            double c = std::cos(kPi * k * (2.0 * m + 1.0) / (2.0 * kMels)) * std::sqrt(2.0 / kMels);
            if (k == 0) c *= std::sqrt(0.5);
            t.dct[size_t(k) * kMels + m] = float(c);
        }
    build_chroma(t);
    t.chroma_planes.assign(size_t(kTunings) * 3 * kChromaPitch * 4, 0.f);
    for (int ti = 0; ti < kTunings; ++ti)
        for (int k = 0; k < kBins; ++k)
            for (int c = 0; c < kChroma; ++c)
                t.chroma_planes[((size_t(ti) * 3 + c / 4) * kChromaPitch + k) * 4 + c % 4] = t.chroma[(size_t(ti) * kBins + k) * kChroma + c];
    // istft normaliser in the interior: four squared-window taps added in ascending frame order
    t.wss.resize(kNrHop);
    for (int p = 0; p < kNrHop; ++p) {
        double s = 0.0;
        for (int j = 3; j >= 0; --j) { const double w = t.hann1024[p + kNrHop * j]; s += w * w; }
        t.wss[p] = s;
    }
    // separable factors of noisereduce's 33 x 7 smoothing filter
    const int n_grad_freq = int(500.0 / (kSR / (kNrFft / 2.0)));
    const int n_grad_time = int(50.0 / ((double(kNrHop) / kSR) * 1000.0));
    t.smooth_f = tri(n_grad_freq);
    t.smooth_t = tri(n_grad_time);
    double sf = 0, st = 0;
    for (double v : t.smooth_f) sf += v;
    for (double v : t.smooth_t) st += v;
    for (double& v : t.smooth_f) v /= sf;
    for (double& v : t.smooth_t) v /= st;
    const double t_frames = 2.0 * kSR / double(kNrHop);
    t.iir_b = (std::sqrt(1 + 4 * t_frames * t_frames) - 1) / (2 * t_frames * t_frames);
}

template <typename T>
bool upload(const std::vector<T>& h, const T** d) {
    T* p = nullptr;
    if (cudaMalloc(&p, h.size() * sizeof(T)) != cudaSuccess) return false;
    if (cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    *d = p;
    return true;
}

// uploads a float vector that the device reads through a wider vector type
template <typename V>
bool upload_as(const std::vector<float>& h, const V** d) {
    const float* p = nullptr;
    if (!upload(h, &p)) return false;
    *d = reinterpret_cast<const V*>(p);                      // cudaMalloc returns 256-byte aligned memory
    return true;
}

constexpr int kMaxDevices = 64;
std::mutex g_mu;
DeviceTables g_dev[kMaxDevices];
bool g_dev_ready[kMaxDevices] = {};

}  // namespace

const HostTables& host_tables() {
    static HostTables* t = [] { auto* p = new HostTables(); build(*p); return p; }();
    return *t;
}

const DeviceTables* device_tables() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error("device_tables: no usable CUDA device");
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_dev_ready[dev]) return &g_dev[dev];
    const HostTables& h = host_tables();
    if (int(h.mel_w.size()) != kMelNnz) {
        set_error("mel filterbank has " + std::to_string(h.mel_w.size()) + " non-zeros, expected 2020");
        return nullptr;
    }
    if (int(h.mel_wt.size()) > kMelWtMax) {
        set_error("step-major mel table has " + std::to_string(h.mel_wt.size()) + " entries, more than kMelWtMax");
        return nullptr;
    }
    DeviceTables d{};
    bool ok = upload(h.hann2048, &d.hann2048) && upload(h.tw1024, &d.tw1024) && upload(h.split2048, &d.split2048) &&
              upload(h.mel_rstart, &d.mel_start) && upload(h.mel_rlen, &d.mel_len) && upload(h.mel_ptr, &d.mel_ptr) &&
              upload(h.mel_w, &d.mel_w) && upload(h.mel_wt, &d.mel_wt) && upload(h.dct, &d.dct) && upload_as(h.chroma_planes, &d.chroma) &&
              upload(h.tuning_edges, &d.tuning_edges) && upload(h.hann1024, &d.hann1024) &&
              upload(h.tw512, &d.tw512) && upload(h.split1024, &d.split1024) &&
              upload(h.wss, &d.wss) && upload(h.smooth_f, &d.smooth_f) && upload(h.smooth_t, &d.smooth_t);
    if (!ok) {
        set_error(std::string("table upload failed: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    d.iir_b = h.iir_b;
    for (int g = 0; g < 4; ++g) d.mel_goff[g] = h.mel_goff[g];
    d.mel_wt_len = int(h.mel_wt.size());
    g_dev[dev] = d;
    g_dev_ready[dev] = true;
    return &g_dev[dev];
}

}  // namespace dys
