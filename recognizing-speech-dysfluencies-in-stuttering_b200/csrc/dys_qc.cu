// Per-file quality-control scalars (reporting only; not on the feature hot path).
//
// Replaces snr_db / spectral_flatness_mean / high_freq_energy_ratio of the reference
// (/root/reference/pipeline1.py:151-186), the three numbers it logs per raw and per cleaned file into
// output_results/per_file_analysis.csv (pipeline1.py:379-381, 394-396, 402-422).
//
//   k_qc_snr       : one CTA per clip: energies of the 400-sample frames (hop 160, float32 sequential sums like
//                    np.sum(frames**2, axis=0)), 25th percentile (numpy 'linear': radix select of the two order
//                    statistics + float32 lerp), noise = frames below it, 10 log10(mean / (noise mean + 1e-10)).
//   k_qc_flatness  : one warp per STFT frame of the power spectrogram k_frame_spectra left in scratch:
//                    exp(mean log max(1e-10, S^2)) / mean max(1e-10, S^2)  (librosa.feature.spectral_flatness).
//   k_qc_hf_bins   : high_freq_energy_ratio needs the FULL-LENGTH DFT of an arbitrary-length clip.  Only the band
//                    at or below 4 kHz is evaluated -- one thread per bin, float64 phasor recurrence re-seeded
//                    exactly every 1024 samples -- and the rest follows from Parseval's identity for the one-sided
//                    spectrum.  O(n^2 / 4) per clip: fine for a per-corpus report (0.2 ms per 3-s clip on a B200),
//                    deliberately not part of the batched feature path.
//   k_qc_finish    : deterministic sums of the partials, Parseval total, means -> out[clip][0..2].
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {

constexpr int kSnrFrame = 400;          // int(0.025 * 16000)  pipeline1.py:153
constexpr int kSnrHop = 160;            // int(0.010 * 16000)  pipeline1.py:154
constexpr int kHfTile = 1024;
constexpr int kHfThreads = 256;

__host__ __device__ inline int snr_frames_of(int n) { return n < kSnrFrame ? 0 : 1 + (n - kSnrFrame) / kSnrHop; }

struct QcScratch {
    float* energy;        // [n_clips][f_max]
    double* hf_part;      // [n_clips][hf_blocks]
    float* flat;          // [n_sub][t_max]   per-frame flatness of the current sub-batch
    int32_t* status;      // [n_clips]        k_feat_init / k_frame_spectra flags (non-finite samples)
    int f_max, hf_blocks, t_max;
};

// rank-th smallest (0-based) of N non-negative floats: 4-pass radix select on the bit patterns
__device__ float block_select(const float* __restrict__ v, int N, int rank, unsigned* hist, unsigned* s_prefix, unsigned* s_k) {
    const int tid = threadIdx.x;
    if (tid == 0) { *s_prefix = 0; *s_k = unsigned(rank); }
    unsigned mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = *s_prefix;
        for (int i = tid; i < N; i += 256) {
            const unsigned key = __float_as_uint(v[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned k = *s_k, cum = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] > k) break;
                cum += hist[b];
            }
            *s_k = k - cum;
            *s_prefix = prefix | (unsigned(b) << (8 * pass));
        }
        mask |= 0xffu << (8 * pass);
        __syncthreads();
    }
    return __uint_as_float(*s_prefix);
}

__global__ void __launch_bounds__(256)
k_qc_snr(const ClipView cv, QcScratch sc, float* __restrict__ out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_k;
    __shared__ double red[2][8];
    __shared__ int cnt[8];
    const int c = blockIdx.x, tid = threadIdx.x;
    int n = cv.lengths[c];
    if (n < 0 || n > cv.max_len) n = 0;
    const int N = snr_frames_of(n);
    if (N == 0) {                                            // pipeline1.py:155-156
        if (tid == 0) out[c * 3 + 0] = 0.f;
        return;
    }
    const float* y = cv.audio + cv.starts[c];
    float* en = sc.energy + size_t(c) * sc.f_max;
    for (int f = tid; f < N; f += 256) {
        const float* p = y + size_t(f) * kSnrHop;
        float e = 0.f;
        for (int i = 0; i < kSnrFrame; ++i) e = __fadd_rn(e, __fmul_rn(p[i], p[i]));     // np.sum(frames**2, axis=0): rows in order
        en[f] = e;
    }
    __syncthreads();
    // np.percentile(energy, 25), method 'linear', in float32
    const double vi = 0.25 * double(N - 1);
    const int lo = int(floor(vi)), hi = min(lo + 1, N - 1);
    const float g = float(vi - double(lo));
    const float a = block_select(en, N, lo, hist, &s_prefix, &s_k);
    __syncthreads();
    const float b = block_select(en, N, hi, hist, &s_prefix, &s_k);
    const float d = __fsub_rn(b, a);
    const float p25 = (g >= 0.5f) ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g))) : __fadd_rn(a, __fmul_rn(d, g));
    double s_all = 0.0, s_noise = 0.0;
    int n_noise = 0;
    for (int f = tid; f < N; f += 256) {
        const float e = en[f];
        s_all += double(e);
        if (e < p25) { s_noise += double(e); ++n_noise; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_all += __shfl_xor_sync(0xffffffffu, s_all, o);
        s_noise += __shfl_xor_sync(0xffffffffu, s_noise, o);
        n_noise += __shfl_xor_sync(0xffffffffu, n_noise, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_all; red[1][tid >> 5] = s_noise; cnt[tid >> 5] = n_noise; }
    __syncthreads();
    if (tid == 0) {
        double A = 0.0, B = 0.0;
        int M = 0;
        for (int w = 0; w < 8; ++w) { A += red[0][w]; B += red[1][w]; M += cnt[w]; }
        float r = 0.f;                                       // pipeline1.py:161-162: empty noise mask -> 0.0
        if (M > 0) {
            const float signal_power = float(A / double(N)), noise_power = float(B / double(M));
            r = 10.0f * log10f(__fdiv_rn(signal_power, __fadd_rn(noise_power, 1e-10f)));
        }
        out[c * 3 + 0] = r;
    }
}

// One warp per frame over the power rows of the current sub-batch.
__global__ void __launch_bounds__(256)
k_qc_flatness(const ClipView cv, int inst0, FeatScratch fs, QcScratch sc) {
    const int li = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n = cv.lengths[inst0 + li];
    if (n <= 0 || n > cv.max_len) return;
    const int T = frames_of(n);
    const int t = blockIdx.y * 8 + warp;
    if (t >= T) return;
    const float* row = fs.power + (size_t(li) * fs.t_max + t) * kBinsPad;
    double lg = 0.0, am = 0.0;
    for (int k = lane; k < kBins; k += 32) {
        const float s = sqrtf(row[k]);                       // S = |stft| in float32, then S ** 2
        const float p = fmaxf(1e-10f, __fmul_rn(s, s));
        lg += double(logf(p));
        am += double(p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lg += __shfl_xor_sync(0xffffffffu, lg, o);
        am += __shfl_xor_sync(0xffffffffu, am, o);
    }
    if (lane == 0) {
        const float gmean = expf(float(lg / double(kBins)));
        const float amean = float(am / double(kBins));
        sc.flat[size_t(li) * sc.t_max + t] = __fdiv_rn(gmean, amean);
    }
}

__global__ void __launch_bounds__(256)
k_qc_flat_reduce(const ClipView cv, int inst0, QcScratch sc, const int32_t* __restrict__ status, float* __restrict__ out) {
    __shared__ double red[8];
    const int li = blockIdx.x, tid = threadIdx.x, c = inst0 + li;
    int n = cv.lengths[c];
    const bool bad = n <= 0 || n > cv.max_len || (status[c] & kStatusNonFinite);       // the reference's except branch -> 0.0
    const int T = bad ? 0 : frames_of(n);
    double s = 0.0;
    for (int t = tid; t < T; t += 256) s += double(sc.flat[size_t(li) * sc.t_max + t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += red[w];
        out[c * 3 + 1] = bad ? 0.f : float(a / double(T));
    }
}

// |X[k]|^2 of the full-length DFT for the bins the reference does NOT count as "high" (f_k <= 4000 Hz).
__global__ void __launch_bounds__(kHfThreads)
k_qc_hf_bins(const ClipView cv, QcScratch sc) {
    __shared__ double tile[kHfTile];
    __shared__ double red[kHfThreads / 32];
    const int c = blockIdx.x, tid = threadIdx.x;
    int n = cv.lengths[c];
    if (n < 0 || n > cv.max_len) n = 0;
    double* part = sc.hf_part + size_t(c) * sc.hf_blocks + blockIdx.y;
    const int k = blockIdx.y * kHfThreads + tid;
    // np.fft.rfftfreq(n, 1/sr)[k] = k * (1 / (n * (1/sr))), evaluated exactly like numpy does (float64)
    const double val = n > 0 ? 1.0 / (double(n) * (1.0 / double(kSR))) : 0.0;
    const bool mine = n > 0 && k <= n / 2 && !(double(k) * val > 4000.0);
    if (n == 0 || blockIdx.y * kHfThreads > n / 4 + 1) {     // whole block above the band (or empty clip)
        if (tid == 0) *part = 0.0;
        return;
    }
    const float* y = cv.audio + cv.starts[c];
    const double inv_n = 1.0 / double(n);
    double sr_ = 0.0, si_ = 0.0, w1r, w1i;
    sincospi(2.0 * double(k) * inv_n, &w1i, &w1r);           // e^{+i 2 pi k / n}; the sign does not matter for |X|^2
    for (int m0 = 0; m0 < n; m0 += kHfTile) {
        const int cnt = min(kHfTile, n - m0);
        __syncthreads();
        for (int i = tid; i < cnt; i += kHfThreads) tile[i] = double(y[m0 + i]);
        __syncthreads();
        if (mine) {
            // phasor at m0, exact argument reduction: (k * m0) mod n in 64-bit integers
            const long long r = (static_cast<long long>(k) * m0) % n;
            double wr, wi;
            sincospi(2.0 * double(r) * inv_n, &wi, &wr);
            for (int i = 0; i < cnt; ++i) {
                const double x = tile[i];
                sr_ = fma(x, wr, sr_);
                si_ = fma(x, wi, si_);
                const double t = wr * w1r - wi * w1i;
                wi = wr * w1i + wi * w1r;
                wr = t;
            }
        }
    }
    double pw = mine ? sr_ * sr_ + si_ * si_ : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
    if ((tid & 31) == 0) red[tid >> 5] = pw;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int w = 0; w < kHfThreads / 32; ++w) a += red[w];
        *part = a;
    }
}

// sum_{k=0}^{n/2} |X[k]|^2 = (n sum y^2 + X[0]^2 + (n even) X[n/2]^2) / 2  for real y  (Parseval, one-sided spectrum)
__global__ void __launch_bounds__(256)
k_qc_finish(const ClipView cv, QcScratch sc, float* __restrict__ out) {
    __shared__ double red[3][8];
    const int c = blockIdx.x, tid = threadIdx.x;
    int n = cv.lengths[c];
    if (n <= 0 || n > cv.max_len) {                          // np.fft.rfft of an empty array raises -> 0.0
        if (tid == 0) { out[c * 3 + 2] = 0.f; if (n <= 0 || n > cv.max_len) out[c * 3 + 1] = 0.f; }
        return;
    }
    const float* y = cv.audio + cv.starts[c];
    double s2 = 0.0, s1 = 0.0, sa = 0.0;
    for (int i = tid; i < n; i += 256) {
        const double x = double(y[i]);
        s2 += x * x;
        s1 += x;
        sa += (i & 1) ? -x : x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s2; red[1][tid >> 5] = s1; red[2][tid >> 5] = sa; }
    __syncthreads();
    if (tid == 0) {
        double S2 = 0.0, S1 = 0.0, SA = 0.0, low = 0.0;
        for (int w = 0; w < 8; ++w) { S2 += red[0][w]; S1 += red[1][w]; SA += red[2][w]; }
        for (int b = 0; b < sc.hf_blocks; ++b) low += sc.hf_part[size_t(c) * sc.hf_blocks + b];
        const double total = 0.5 * (double(n) * S2 + S1 * S1 + ((n & 1) ? 0.0 : SA * SA));
        out[c * 3 + 2] = float((total - low) / (total + 1e-10));
    }
}

size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

// Lays the QC scratch out from byte offset 0; returns the offset where the feature scratch (power rows) starts.
size_t qc_layout(int n_clips, int max_len, int n_sub, QcScratch* q, unsigned char* base) {
    q->f_max = std::max(1, snr_frames_of(std::max(max_len, 0)));
    q->hf_blocks = (std::max(max_len, 0) / 4 + 2 + kHfThreads - 1) / kHfThreads;
    q->t_max = frames_of(std::max(max_len, 0));
    size_t off = 0;
    q->energy = reinterpret_cast<float*>(base + off); off += al256(size_t(n_clips) * q->f_max * 4);
    q->hf_part = reinterpret_cast<double*>(base + off); off += al256(size_t(n_clips) * q->hf_blocks * 8);
    q->flat = reinterpret_cast<float*>(base + off); off += al256(size_t(n_sub) * q->t_max * 4);
    q->status = reinterpret_cast<int32_t*>(base + off); off += al256(size_t(n_clips) * 4);
    return off;
}

}  // namespace

size_t qc_scratch_bytes(int n_clips, int max_len, int n_sub) {
    QcScratch q;
    const size_t off = qc_layout(n_clips, max_len, n_sub, &q, nullptr);
    return off + feat_scratch_bytes(n_sub, q.t_max);
}

cudaError_t launch_qc(const DeviceTables& tb, const ClipView& cv, float* out, void* scratch, size_t scratch_bytes, int n_sub,
                      cudaStream_t stream) {
    const int n = cv.n_clips;
    if (n <= 0) return cudaSuccess;
    if (scratch_bytes < qc_scratch_bytes(n, cv.max_len, n_sub)) return cudaErrorInvalidValue;
    QcScratch q;
    unsigned char* sbase = static_cast<unsigned char*>(scratch);
    void* feat_base = sbase + qc_layout(n, cv.max_len, n_sub, &q, sbase);
    { LaunchScope ls(kK_qc_snr, stream);
      k_qc_snr<<<n, 256, 0, stream>>>(cv, q, out); }
    { LaunchScope ls(kK_qc_hf_bins, stream);
      k_qc_hf_bins<<<dim3(n, q.hf_blocks), kHfThreads, 0, stream>>>(cv, q); }
    { LaunchScope ls(kK_qc_finish, stream);
      k_qc_finish<<<n, 256, 0, stream>>>(cv, q, out); }
    FeatScratch fs;
    feat_scratch_carve(feat_base, n_sub, q.t_max, &fs);
    for (int i0 = 0; i0 < n; i0 += n_sub) {
        const int cnt = std::min(n_sub, n - i0);
        cudaError_t e = launch_power_only(tb, cv, i0, cnt, fs, q.status, stream);
        if (e != cudaSuccess) return e;
        { LaunchScope ls(kK_qc_flatness, stream);
          k_qc_flatness<<<dim3(cnt, (q.t_max + 7) / 8), 256, 0, stream>>>(cv, i0, fs, q); }
        { LaunchScope ls(kK_qc_flat_reduce, stream);
          k_qc_flat_reduce<<<cnt, 256, 0, stream>>>(cv, i0, q, q.status, out); }
    }
    return cudaGetLastError();
}

}  // namespace dys
