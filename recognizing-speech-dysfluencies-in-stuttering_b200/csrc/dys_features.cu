// Feature kernels: 16 kHz clip -> float32[149]  (MFCC/delta/delta2/chroma statistics).
//
// Replaces extract_audio_features / extract_features of the reference
// (/root/reference/pipeline1.py:206-265; identical copy main1.py:665-715), i.e. the librosa
// call chain  feature.mfcc -> melspectrogram -> stft -> power_to_db -> dct,  feature.delta x2,
// feature.chroma_stft -> estimate_tuning -> piptrack -> filters.chroma.
//
//   k_frame_spectra : one warp per STFT frame.  The 32 coalesced 8-byte sample loads of a frame
//                     (PCM-16 on the clean branch: 4-byte) are all in flight at once and land directly
//                     in the FFT's register layout (4x frame overlap is served by L1/L2); windowed,
//                     transformed by the warp-resident 1024-point complex FFT (dys_fft.cuh), split into
//                     the 1025-bin power spectrum, and -- while the frame is still in shared memory /
//                     registers -- reduced to 128 log-mel values and to the piptrack peak list.
//   k_tuning        : one CTA per clip: exact median (radix select) of the peak magnitudes,
//                     100-bin residual histogram, first arg-max  -> tuning index.
//   k_frame_cepstra : one warp per 6 frames (every chroma weight load feeds 6 x 12 FMAs): top-dB clamp +
//                     DCT-II (20x128) and the 12x1025 chroma projection with the tuning's filterbank + inf-norm.
//   k_clip_stats    : one CTA per clip: delta / delta-delta (Savitzky-Golay taps, replicated
//                     edges) and mean / population std of every row -> the 149-vector.
#include <cfloat>
#include <cmath>

#include "dys_fft.cuh"
#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kFramesPerCta = 96;            // 12 frames per warp; one CTA covers a 3-s clip (94 frames; two CTAs of 48: +5 %)

__device__ __forceinline__ int enc_f32(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float dec_f32(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct InstSrc {
    const float* f32;    // float32 samples of the clip on this branch (raw clip, or raw clip as clean fallback) ...
    const int16_t* q16;  // ... or PCM-16: the caller's 16-bit samples, or what the reference would have written to
                         //     clear_audio/<stem>.wav (clean branch)
    int n;               // samples
    bool vec_ok;         // the active source allows the vector loads (f32: 8-byte aligned, q16: 4-byte aligned)
};

__device__ __forceinline__ InstSrc inst_source(const ClipView& cv, int inst) {
    InstSrc s;
    const bool clean = inst >= cv.n_clips;
    const int c = clean ? inst - cv.n_clips : inst;
    int n = cv.lengths[c];
    if (n < 0 || n > cv.max_len) n = 0;
    s.n = n;
    s.f32 = cv.audio ? cv.audio + cv.starts[c] : nullptr;
    s.q16 = cv.audio_q ? cv.audio_q + cv.starts[c] : nullptr;
    if (clean && cv.clean_flag[c] == 0) s.q16 = cv.clean_q + int64_t(c) * cv.clean_pitch;
    s.vec_ok = s.q16 ? (reinterpret_cast<uintptr_t>(s.q16) & 3u) == 0 : (reinterpret_cast<uintptr_t>(s.f32) & 7u) == 0;
    return s;
}

// ------------------------------------------------------------------------------------------
__global__ void k_feat_init(int* peak_count, int* lmax_enc, int32_t* status, const ClipView cv, int inst0, int n_inst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst) return;
    peak_count[i] = 0;
    lmax_enc[i] = enc_f32(-INFINITY);
    const int inst = inst0 + i;
    const bool clean = inst >= cv.n_clips;
    const int c = clean ? inst - cv.n_clips : inst;
    const int n = cv.lengths[c];
    int st = 0;
    if (n < 0 || n > cv.max_len) st |= kStatusBadLength | kStatusShort;
    else if (frames_of(n) < 9) st |= kStatusShort;
    if (clean && cv.clean_flag[c] != 0) st |= kStatusCleanFallback;
    status[inst] = st;
}

// ------------------------------------------------------------------------------------------
// Per-warp scratch inside the 32x33 float2 exchange tile (8448 B), reused along the frame:
//   [0, 8192)      windowed frame as 1024 complex pairs (before the FFT)
//   [0, 8448)      transposition tile (during the FFT)
//   [0, 4100)      power spectrum P[0..1024]           (after the FFT)
//   [4224, 8328)   upper half of the complex spectrum: zb[j] = Z[512 + j], j = 1..511, zb[512] = Z[0]
constexpr int kZbOffset = 528;             // in float2 units

struct SpectraSmem {
    float hann[kNfft];
    float2 tw[32 * 32];
    float2 split[1024];
    float mel_wt[kMelWtMax];            // step-major weights (dys_tables.h): conflict-free across the 32 filters of a group
    int mel_start[kMels];
    int mel_len[kMels];
    float2 xbuf[kWarps][kXbuf1024];
};

__global__ void __launch_bounds__(kThreads, 2)
k_frame_spectra(const DeviceTables tb, const ClipView cv, int inst0, FeatScratch sc, int32_t* __restrict__ status, int min_frames) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SpectraSmem& sm = *reinterpret_cast<SpectraSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;                 // local instance
    const int inst = inst0 + li;
    const InstSrc src = inst_source(cv, inst);
    const int T = frames_of(src.n);
    const int t_begin = blockIdx.y * kFramesPerCta;
    if (T < min_frames || src.n <= 0 || t_begin >= T) return;   // features: short clips (T < 9) emit zeros in k_clip_stats
    const int t_end = min(T, t_begin + kFramesPerCta);

    for (int i = tid; i < kNfft; i += kThreads) sm.hann[i] = tb.hann2048[i];
    for (int i = tid; i < 1024; i += kThreads) { sm.tw[i] = tb.tw1024[i]; sm.split[i] = tb.split2048[i]; }
    for (int i = tid; i < tb.mel_wt_len; i += kThreads) sm.mel_wt[i] = tb.mel_wt[i];
    for (int i = tid; i < kMels; i += kThreads) { sm.mel_start[i] = tb.mel_start[i]; sm.mel_len[i] = tb.mel_len[i]; }
    __syncthreads();

    float2* xbuf = sm.xbuf[warp];
    float* pbuf = reinterpret_cast<float*>(xbuf);
    float2* zb = xbuf + kZbOffset;
    const float2* hann2 = reinterpret_cast<const float2*>(sm.hann);
    float* g_power = sc.power + size_t(li) * sc.t_max * kBinsPad;
    float* g_logmel = sc.logmel + size_t(li) * sc.t_max * kMels;
    float2* g_peaks = sc.peaks + size_t(li) * sc.t_max * kMaxPeaksPerFrame;
    float warp_lmax = -INFINITY;
    bool nonfinite = false;
    // The clean branch reads PCM-16 and librosa.load returns int16 / 32768: the FFT runs on the integers (scaling by
    // a power of two commutes exactly with every rounding on the way) and the 2^-30 lands in the power's 1/4 factor.
    const float pscale = src.q16 ? 0.25f * (1.0f / 1073741824.0f) : 0.25f;
    const int n = src.n;

    for (int t = t_begin + warp; t < t_end; t += kWarps) {
        // ---- windowed frame straight into the FFT's register layout: v[m] = (x[2i], x[2i+1]) * hann, i = lane + 32 m.
        // Interior frames issue all 32 coalesced 8-byte (PCM-16: 4-byte) loads before the first use; frames that
        // overlap the zero centre padding or the clip end take the bounds-checked loop through shared memory.
        const int f0 = t * kHop - kNfft / 2;               // clip-relative index of the frame's first sample
        const bool interior = f0 >= 0 && f0 + kNfft <= n;
        if (t + kWarps < t_end) {                           // this warp's next frame: pull its lines towards L2 now
            const int s_next = f0 + kWarps * kHop + 64 * lane;          // 64 samples: 256 B of float32, 128 B of PCM-16
            if (s_next >= 0 && s_next < n) {
                if (src.q16) asm volatile("prefetch.global.L2 [%0];" ::"l"(src.q16 + s_next));
                else {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src.f32 + s_next));
                    if (s_next + 32 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(src.f32 + s_next + 32));
                }
            }
        }
        float2 v[32];
        if (interior && src.q16 && src.vec_ok) {            // PCM-16: int16 / 32768 (exact), as librosa.load reads a 16-bit WAV
            const short2* q2 = reinterpret_cast<const short2*>(src.q16 + f0);
            short2 raw[32];
            static_for<32>([&](auto im) { constexpr int m = decltype(im)::value; raw[m] = __ldg(q2 + lane + 32 * m); });
            static_for<32>([&](auto im) {
                constexpr int m = decltype(im)::value;
                v[m] = vmul(make_float2(float(raw[m].x), float(raw[m].y)), hann2[lane + 32 * m]);   // x 2^15, undone at the power
            });
        } else if (interior && !src.q16 && src.vec_ok) {
            const float2* x2 = reinterpret_cast<const float2*>(src.f32 + f0);
            static_for<32>([&](auto im) { constexpr int m = decltype(im)::value; v[m] = __ldg(x2 + lane + 32 * m); });
            static_for<32>([&](auto im) {
                constexpr int m = decltype(im)::value;
                v[m] = vmul(v[m], hann2[lane + 32 * m]);
            });
        } else {
#pragma unroll 2
            for (int i = lane; i < kNfft / 2; i += 32) {
                const int s = f0 + 2 * i;
                float a = 0.f, b = 0.f;
                if (src.q16) {
                    if (s >= 0 && s < n) a = float(__ldg(src.q16 + s));
                    if (s + 1 >= 0 && s + 1 < n) b = float(__ldg(src.q16 + s + 1));
                } else {
                    if (s >= 0 && s < n) a = __ldg(src.f32 + s);
                    if (s + 1 >= 0 && s + 1 < n) b = __ldg(src.f32 + s + 1);
                }
                const float2 w = hann2[i];
                xbuf[i] = make_float2(a * w.x, b * w.y);
            }
            __syncwarp();
            static_for<32>([&](auto im) {
                constexpr int m = decltype(im)::value;
                v[m] = xbuf[lane + 32 * m];
            });
            __syncwarp();
        }

        warp_fft1024_rolled(v, xbuf, sm.tw, lane);          // Z[lane + 32 q] = v[bitrev(q)]
        // Z[0] = (sum of even samples, sum of odd samples) * window: non-finite iff some sample is
        // (librosa.util.valid_audio raises on those -> zeros, pipeline1.py:237-239)
        if (lane == 0) nonfinite |= !(isfinite(v[0].x) && isfinite(v[0].y));

        // ---- real split, two bins per pair: with E = (Z[k] + conj Z[1024-k]) / 2 and
        // T = e^{-2 pi i k / 2048} * (-i) (Z[k] - conj Z[1024-k]) / 2 :  X[k] = E + T,  X[1024-k] = conj(E - T).
        // k = 0 pairs with the Nyquist bin through zb[512] = Z[0]; k = 512 is its own partner.
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value + 16;
            zb[lane + 32 * (q - 16)] = v[bitrev(q, 5)];     // Z[512 + lane + 32 (q - 16)]
        });
        if (lane == 0) zb[512] = v[0];
        __syncwarp();
        float* gp = g_power + size_t(t) * kBinsPad;
        float fmax_ = 0.f;
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value;
            const int k = lane + 32 * q;                    // 0 .. 511
            const float2 z = v[bitrev(q, 5)];
            const float2 p = zb[512 - k];                   // Z[1024 - k]  (k = 0: Z[0])
            const float2 pc = make_float2(p.x, -p.y);
            const float2 e = cadd(z, pc), d = csub(z, pc);  // packed FADD2: (ex, ey) = z + conj p, (dx, dy) = z - conj p
            const float2 cs = sm.split[k];
            const float2 tt = make_float2(cs.x * d.y - cs.y * d.x, -(cs.x * d.x + cs.y * d.y));
            const float2 a = cadd(e, tt), b = csub(e, tt);
            const float plo = pscale * (a.x * a.x + a.y * a.y);
            const float phi = pscale * (b.x * b.x + b.y * b.y);
            fmax_ = fmaxf(fmax_, fmaxf(plo, phi));
            gp[k] = plo;
            gp[1024 - k] = phi;
            pbuf[k] = plo;                                  // the power row lives below the zb region of the tile
            pbuf[1024 - k] = phi;
        });
        if (lane == 0) {
            const float2 z = v[bitrev(16, 5)];
            const float p_mid = (4.0f * pscale) * (z.x * z.x + z.y * z.y);      // X[512] = conj Z[512]
            gp[512] = p_mid;
            pbuf[512] = p_mid;
            fmax_ = fmaxf(fmax_, p_mid);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fmax_ = fmaxf(fmax_, __shfl_xor_sync(0xffffffffu, fmax_, o));
        __syncwarp();

        // ---- sparse slaney mel (<= 2 filters per bin) + 10 log10 ---------------------------------
        float* gl = g_logmel + size_t(t) * kMels;
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            const int f = lane + 32 * g;
            const int st = sm.mel_start[f], ln = sm.mel_len[f];
            const float* wt = sm.mel_wt + tb.mel_goff[g] + lane;
            float acc = 0.f;
#pragma unroll 4
            for (int j = 0; j < ln; ++j) acc = fmaf(wt[32 * j], pbuf[st + j], acc);
            const float L = 10.0f * log10f(fmaxf(acc, 1e-10f));
            gl[f] = L;
            warp_lmax = fmaxf(warp_lmax, L);
        }

        // ---- piptrack: thresholded local maxima in [150, 4000) Hz, parabolic refinement ----------
        const float ref = 0.1f * fmax_;
        unsigned pkmask = 0;
#pragma unroll 4
        for (int q = 0; q < 16; ++q) {
            const int k = lane + 32 * q;
            if (k >= kPipLo) {
                const float c = pbuf[k], l = pbuf[k - 1], r = pbuf[k + 1];
                const float qc = c > ref ? c : 0.f, ql = l > ref ? l : 0.f, qr = r > ref ? r : 0.f;
                if (qc > ql && qc >= qr) pkmask |= 1u << q;
            }
        }
        int npk = __popc(pkmask);
        int incl = npk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total > 0) {
            int base = 0;
            if (lane == 31) base = atomicAdd(&sc.peak_count[li], total);
            base = __shfl_sync(0xffffffffu, base, 31);
            int slot = base + incl - npk;
            while (pkmask) {
                const int q = __ffs(pkmask) - 1;
                pkmask &= pkmask - 1;
                const int k = lane + 32 * q;
                const float c = pbuf[k], l = pbuf[k - 1], r = pbuf[k + 1];
                // librosa >= 0.10 _parabolic_interpolation evaluates in float64 and stores float32
                const double a = double(r) + double(l) - 2.0 * double(c);
                const double b = (double(r) - double(l)) / 2.0;
                const float shift = (fabs(b) >= fabs(a)) ? 0.f : float(-b / a);
                const float avg = (r - l) * 0.5f;                       // np.gradient, float32
                const float dskew = (0.5f * avg) * shift;
                const float pitch = float((double(k) + double(shift)) * 16000.0 / 2048.0);
                g_peaks[slot++] = make_float2(pitch, c + dskew);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_lmax = fmaxf(warp_lmax, __shfl_xor_sync(0xffffffffu, warp_lmax, o));
    if (lane == 0 && warp_lmax > -INFINITY) atomicMax(&sc.lmax_enc[li], enc_f32(warp_lmax));
    if (__any_sync(0xffffffffu, nonfinite) && lane == 0) atomicOr(&status[inst], kStatusNonFinite);
}

// ------------------------------------------------------------------------------------------
// librosa.estimate_tuning: threshold = median(mag), histogram of mod(12 log2(f / 27.5), 1)
// folded to [-0.5, 0.5) over 100 bins, first arg-max.
__global__ void __launch_bounds__(256)
k_tuning(const DeviceTables tb, FeatScratch sc) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_k;
    __shared__ unsigned s_cle, s_mingt;
    const int li = blockIdx.x, tid = threadIdx.x;
    const int N = min(sc.peak_count[li], sc.t_max * kMaxPeaksPerFrame);
    const float2* pk = sc.peaks + size_t(li) * sc.t_max * kMaxPeaksPerFrame;
    if (N == 0) {
        if (tid == 0) sc.tuning_idx[li] = kTunings / 2;       // no pitches -> tuning 0.0
        return;
    }
    // radix select of rank r0 = (N-1)/2 on the (positive) float bit patterns
    if (tid == 0) { s_prefix = 0; s_k = unsigned(N - 1) / 2; }
    unsigned mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int i = tid; i < N; i += 256) {
            const unsigned key = __float_as_uint(pk[i].y);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned k = s_k, cum = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] > k) break;
                cum += hist[b];
            }
            s_k = k - cum;
            s_prefix = prefix | (unsigned(b) << (8 * pass));
        }
        mask |= 0xffu << (8 * pass);
        __syncthreads();
    }
    const unsigned v0 = s_prefix;
    if (tid == 0) { s_cle = 0; s_mingt = 0xffffffffu; }
    for (int i = tid; i < 100; i += 256) hist[i] = 0;
    __syncthreads();
    unsigned cle = 0, mingt = 0xffffffffu;
    for (int i = tid; i < N; i += 256) {
        const unsigned key = __float_as_uint(pk[i].y);
        if (key <= v0) ++cle; else mingt = min(mingt, key);
    }
    atomicAdd(&s_cle, cle);
    atomicMin(&s_mingt, mingt);
    __syncthreads();
    const unsigned r1 = unsigned(N) / 2;
    const unsigned v1 = (s_cle >= r1 + 1) ? v0 : s_mingt;
    const float thr = (__uint_as_float(v0) + __uint_as_float(v1)) * 0.5f;      // np.median: mean of the two middles
    for (int i = tid; i < N; i += 256) {
        const float2 p = pk[i];
        if (p.y >= thr && p.x > 0.f) {
            const float octs = float(log2(double(p.x / 27.5f)));               // float32 log2, correctly rounded
            float r = fmodf(12.0f * octs, 1.0f);
            if (r >= 0.5f) r -= 1.0f;
            const double x = double(r);
            int b = int(floor((x + 0.5) * 100.0));
            b = max(0, min(99, b));
            while (b > 0 && x < tb.tuning_edges[b]) --b;
            while (b < 99 && x >= tb.tuning_edges[b + 1]) ++b;
            atomicAdd(&hist[b], 1u);
        }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned best = 0; int bi = 0;
        for (int b = 0; b < 100; ++b) if (hist[b] > best) { best = hist[b]; bi = b; }
        sc.tuning_idx[li] = bi;
    }
}

// ------------------------------------------------------------------------------------------
#ifndef DYS_CEP_UNROLL
#define DYS_CEP_UNROLL 2
#endif
#ifndef DYS_CEP_PREFETCH
#define DYS_CEP_PREFETCH 1
#endif
#ifndef DYS_CEP_FRAMES
#define DYS_CEP_FRAMES 6
#endif
// frames per warp iteration: every chroma weight load feeds kCepFrames x 12 FMAs (8 warps x 6 frames: a 94-frame clip
// takes two balanced iterations)
constexpr int kCepFrames = DYS_CEP_FRAMES;
constexpr int kCepUnroll = DYS_CEP_UNROLL;     // bins-per-lane steps unrolled in the chroma projection

struct CepstraSmem {
    float dctT[kMels * kMfcc];          // [m][k]
    float4 lrow[kWarps][kCepFrames][kMels / 4];
};

__device__ __forceinline__ void chroma_fma(float (&acc)[kChroma], float p, const float4& w0, const float4& w1, const float4& w2) {
    acc[0] = fmaf(w0.x, p, acc[0]); acc[1] = fmaf(w0.y, p, acc[1]); acc[2] = fmaf(w0.z, p, acc[2]);
    acc[3] = fmaf(w0.w, p, acc[3]); acc[4] = fmaf(w1.x, p, acc[4]); acc[5] = fmaf(w1.y, p, acc[5]);
    acc[6] = fmaf(w1.z, p, acc[6]); acc[7] = fmaf(w1.w, p, acc[7]); acc[8] = fmaf(w2.x, p, acc[8]);
    acc[9] = fmaf(w2.y, p, acc[9]); acc[10] = fmaf(w2.z, p, acc[10]); acc[11] = fmaf(w2.w, p, acc[11]);
}

__global__ void __launch_bounds__(kThreads, 2)
k_frame_cepstra(const DeviceTables tb, const ClipView cv, int inst0, FeatScratch sc) {
    __shared__ CepstraSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;
    const InstSrc src = inst_source(cv, inst0 + li);
    const int T = frames_of(src.n);
    const int t_begin = blockIdx.y * kFramesPerCta;
    if (T < 9 || t_begin >= T) return;
    const int t_end = min(T, t_begin + kFramesPerCta);
    for (int i = tid; i < kMels * kMfcc; i += kThreads) {
        const int m = i / kMfcc, k = i % kMfcc;
        sm.dctT[i] = tb.dct[k * kMels + m];
    }
    __syncthreads();
    const float thr = dec_f32(sc.lmax_enc[li]) - 80.0f;                 // power_to_db top_db over the WHOLE clip
    const float4* wtab = tb.chroma + size_t(sc.tuning_idx[li]) * 3 * kChromaPitch;        // planes of chroma rows 0-3, 4-7, 8-11
#define DYS_W3(k) __ldg(&wtab[k]), w1 = __ldg(&wtab[kChromaPitch + (k)]), w2 = __ldg(&wtab[2 * kChromaPitch + (k)])
    const float* g_power = sc.power + size_t(li) * sc.t_max * kBinsPad;
    const float* g_logmel = sc.logmel + size_t(li) * sc.t_max * kMels;
    float* g_mfcc = sc.mfcc + size_t(li) * sc.t_max * kMfcc;
    float* g_chroma = sc.chroma + size_t(li) * sc.t_max * kChroma;
    float4 (*lrow)[kMels / 4] = sm.lrow[warp];

    for (int t0 = t_begin + warp * kCepFrames; t0 < t_end; t0 += kWarps * kCepFrames) {
        const int nf = min(kCepFrames, t_end - t0);                      // warp-uniform
        // ---- clamp + DCT-II: lane k < 20 owns coefficient k of the warp's frames ----------------------
#pragma unroll
        for (int f = 0; f < kCepFrames; ++f) {
            const int tt = min(t0 + f, t_end - 1);
            const float4 L = reinterpret_cast<const float4*>(g_logmel + size_t(tt) * kMels)[lane];
            lrow[f][lane] = make_float4(fmaxf(L.x, thr), fmaxf(L.y, thr), fmaxf(L.z, thr), fmaxf(L.w, thr));
        }
        __syncwarp();
        if (lane < kMfcc) {
            float a[kCepFrames];
#pragma unroll
            for (int f = 0; f < kCepFrames; ++f) a[f] = 0.f;
#pragma unroll 2
            for (int m4 = 0; m4 < kMels / 4; ++m4) {
                const float d0 = sm.dctT[(4 * m4 + 0) * kMfcc + lane], d1 = sm.dctT[(4 * m4 + 1) * kMfcc + lane];
                const float d2 = sm.dctT[(4 * m4 + 2) * kMfcc + lane], d3 = sm.dctT[(4 * m4 + 3) * kMfcc + lane];
#pragma unroll
                for (int f = 0; f < kCepFrames; ++f) {
                    const float4 l = lrow[f][m4];
                    a[f] = fmaf(l.x, d0, a[f]); a[f] = fmaf(l.y, d1, a[f]); a[f] = fmaf(l.z, d2, a[f]); a[f] = fmaf(l.w, d3, a[f]);
                }
            }
#pragma unroll
            for (int f = 0; f < kCepFrames; ++f)
                if (f < nf) g_mfcc[size_t(t0 + f) * kMfcc + lane] = a[f];
        }
        // ---- chroma: 12 x 1025 projection of the warp's frames, lane owns bins lane + 32 j -----------------
        float acc[kCepFrames][kChroma];
#pragma unroll
        for (int f = 0; f < kCepFrames; ++f)
#pragma unroll
            for (int c = 0; c < kChroma; ++c) acc[f][c] = 0.f;
#if DYS_CEP_PREFETCH
        {   // the warp's power rows are contiguous (kCepFrames x 4128 B, straight from DRAM): ask for all of their lines
            // now, so that the projection loop below finds them in L2 (it keeps only a few loads per lane in flight)
            const char* blk = reinterpret_cast<const char*>(g_power + size_t(t0) * kBinsPad);
            const int bytes = nf * kBinsPad * 4;
            for (int o = 128 * lane; o < bytes; o += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + o));
        }
#endif
        int go[kCepFrames];                                              // row offsets (frames past the end repeat the last row)
#pragma unroll
        for (int f = 0; f < kCepFrames; ++f) go[f] = min(t0 + f, t_end - 1) * kBinsPad + lane;
#pragma unroll kCepUnroll
        for (int j = 0; j < 32; ++j) {
            const int k = lane + 32 * j;
            const float4 w0 = DYS_W3(k);
#pragma unroll
            for (int f = 0; f < kCepFrames; ++f) chroma_fma(acc[f], g_power[go[f] + 32 * j], w0, w1, w2);
        }
        if (lane == 0) {
            const int k = 1024;
            const float4 w0 = DYS_W3(k);
#pragma unroll
            for (int f = 0; f < kCepFrames; ++f) chroma_fma(acc[f], g_power[go[f] + k], w0, w1, w2);
        }
#pragma unroll
        for (int f = 0; f < kCepFrames; ++f) {
            float cmax = 0.f;
#pragma unroll
            for (int c = 0; c < kChroma; ++c) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[f][c] += __shfl_xor_sync(0xffffffffu, acc[f][c], o);
                cmax = fmaxf(cmax, fabsf(acc[f][c]));
            }
            if (cmax < FLT_MIN) cmax = 1.0f;                             // util.normalize: below tiny -> unscaled
            float mine = 0.f;
#pragma unroll
            for (int c = 0; c < kChroma; ++c) if (lane == c) mine = acc[f][c];
            if (lane < kChroma && f < nf) g_chroma[size_t(t0 + f) * kChroma + lane] = __fdiv_rn(mine, cmax);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_clip_stats(const ClipView cv, int inst0, FeatScratch sc, float* __restrict__ out_raw, float* __restrict__ out_clean,
             const int32_t* __restrict__ status) {
    constexpr int kParts = 12;
    __shared__ double red[kParts][kMfcc][6];
    __shared__ double cred[21][kChroma][2];
    const int tid = threadIdx.x, li = blockIdx.x, inst = inst0 + li;
    const bool clean = inst >= cv.n_clips;
    const int c = clean ? inst - cv.n_clips : inst;
    float* out = (clean ? out_clean : out_raw) + size_t(c) * kFeat;
    const int st = status[inst];
    if (st & (kStatusShort | kStatusNonFinite | kStatusBadLength)) {
        for (int i = tid; i < kFeat; i += 256) out[i] = 0.f;
        return;
    }
    int n = cv.lengths[c];
    const int T = frames_of(n);
    const float* M = sc.mfcc + size_t(li) * sc.t_max * kMfcc;
    const float* C = sc.chroma + size_t(li) * sc.t_max * kChroma;
    if (tid < kParts * kMfcc) {
        const int k = tid % kMfcc, part = tid / kMfcc;
        const double x0 = double(M[k]);               // shift by the first frame: exact zero std for constant rows
        double s[6] = {0, 0, 0, 0, 0, 0};
        for (int t = part; t < T; t += kParts) {
            const int tc = min(max(t, 4), T - 5);     // savgol mode='interp' with polyorder == deriv: edges replicate
            float w[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) w[j] = M[size_t(tc - 4 + j) * kMfcc + k];
            const double d1 = (4.0 * (double(w[8]) - double(w[0])) + 3.0 * (double(w[7]) - double(w[1])) +
                               2.0 * (double(w[6]) - double(w[2])) + (double(w[5]) - double(w[3]))) / 60.0;
            const double d2 = (28.0 * (double(w[0]) + double(w[8])) + 7.0 * (double(w[1]) + double(w[7])) -
                               8.0 * (double(w[2]) + double(w[6])) - 17.0 * (double(w[3]) + double(w[5])) -
                               20.0 * double(w[4])) / 462.0;
            const float d1f = float(d1), d2f = float(d2);      // librosa.feature.delta returns float32
            const double x = double(M[size_t(t) * kMfcc + k]) - x0;
            s[0] += x; s[1] += x * x;
            s[2] += double(d1f); s[3] += double(d1f) * double(d1f);
            s[4] += double(d2f); s[5] += double(d2f) * double(d2f);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) red[part][k][j] = s[j];
    }
    if (tid < 21 * kChroma) {
        const int ch = tid % kChroma, part = tid / kChroma;
        double a = 0, b = 0;
        for (int t = part; t < T; t += 21) {
            const double x = double(C[size_t(t) * kChroma + ch]);
            a += x; b += x * x;
        }
        cred[part][ch][0] = a; cred[part][ch][1] = b;
    }
    __syncthreads();
    const double invT = 1.0 / double(T);
    if (tid < kMfcc * 3) {
        const int k = tid % kMfcc, which = tid / kMfcc;          // 0 mfcc, 1 delta, 2 delta2
        double a = 0, b = 0;
        for (int p = 0; p < kParts; ++p) { a += red[p][k][2 * which]; b += red[p][k][2 * which + 1]; }
        const double mean = a * invT;
        const double var = fmax(b * invT - mean * mean, 0.0);
        const double shift = which == 0 ? double(M[k]) : 0.0;
        out[which * 40 + k] = float(mean + shift);
        out[which * 40 + 20 + k] = float(sqrt(var));
    } else if (tid >= 64 && tid < 64 + kChroma) {
        const int ch = tid - 64;
        double a = 0, b = 0;
        for (int p = 0; p < 21; ++p) { a += cred[p][ch][0]; b += cred[p][ch][1]; }
        const double mean = a * invT;
        const double var = fmax(b * invT - mean * mean, 0.0);
        out[120 + ch] = float(mean);
        out[132 + ch] = float(sqrt(var));
    } else if (tid >= 96 && tid < 96 + (kFeat - kAudioFeat)) {
        out[kAudioFeat + tid - 96] = 0.f;                        // extract_text_features("") -> zeros(5)
    }
}

}  // namespace

size_t feat_scratch_bytes(int n_inst, int t_max) {
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t f = size_t(n_inst) * t_max;
    return al(f * kBinsPad * 4) + al(f * kMels * 4) + al(f * kMfcc * 4) + al(f * kChroma * 4) +
           al(f * kMaxPeaksPerFrame * 8) + 3 * al(size_t(n_inst) * 4);
}

void feat_scratch_carve(void* base, int n_inst, int t_max, FeatScratch* out) {
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    unsigned char* p = static_cast<unsigned char*>(base);
    const size_t f = size_t(n_inst) * t_max;
    out->power = reinterpret_cast<float*>(p); p += al(f * kBinsPad * 4);
    out->logmel = reinterpret_cast<float*>(p); p += al(f * kMels * 4);
    out->mfcc = reinterpret_cast<float*>(p); p += al(f * kMfcc * 4);
    out->chroma = reinterpret_cast<float*>(p); p += al(f * kChroma * 4);
    out->peaks = reinterpret_cast<float2*>(p); p += al(f * kMaxPeaksPerFrame * 8);
    out->peak_count = reinterpret_cast<int*>(p); p += al(size_t(n_inst) * 4);
    out->lmax_enc = reinterpret_cast<int*>(p); p += al(size_t(n_inst) * 4);
    out->tuning_idx = reinterpret_cast<int*>(p);
    out->t_max = t_max;
}

static cudaError_t spectra_attr() { return ensure_dynamic_smem<kK_frame_spectra>(k_frame_spectra, int(sizeof(SpectraSmem))); }

// Power spectrograms only (QC metrics): every clip with at least one sample, however short.
cudaError_t launch_power_only(const DeviceTables& tb, const ClipView& cv, int inst0, int n_inst, const FeatScratch& sc,
                              int32_t* status, cudaStream_t stream) {
    if (n_inst <= 0) return cudaSuccess;
    if (cudaError_t e = spectra_attr()) return e;
    const int gx = (sc.t_max + kFramesPerCta - 1) / kFramesPerCta;
    { LaunchScope ls(kK_feat_init, stream);
      k_feat_init<<<(n_inst + 255) / 256, 256, 0, stream>>>(sc.peak_count, sc.lmax_enc, status, cv, inst0, n_inst); }
    { LaunchScope ls(kK_frame_spectra, stream);
      k_frame_spectra<<<dim3(n_inst, gx), kThreads, sizeof(SpectraSmem), stream>>>(tb, cv, inst0, sc, status, 1); }
    return cudaGetLastError();
}

cudaError_t launch_features(const DeviceTables& tb, const ClipView& cv, int inst0, int n_inst, const FeatScratch& sc,
                            float* out_raw, float* out_clean, int32_t* status, cudaStream_t stream) {
    if (n_inst <= 0) return cudaSuccess;
    if (cudaError_t e = spectra_attr()) return e;
    const int gx = (sc.t_max + kFramesPerCta - 1) / kFramesPerCta;
    { LaunchScope ls(kK_feat_init, stream);
      k_feat_init<<<(n_inst + 255) / 256, 256, 0, stream>>>(sc.peak_count, sc.lmax_enc, status, cv, inst0, n_inst); }
    { LaunchScope ls(kK_frame_spectra, stream);
      k_frame_spectra<<<dim3(n_inst, gx), kThreads, sizeof(SpectraSmem), stream>>>(tb, cv, inst0, sc, status, 9); }
    { LaunchScope ls(kK_tuning, stream);
      k_tuning<<<n_inst, 256, 0, stream>>>(tb, sc); }
    { LaunchScope ls(kK_frame_cepstra, stream);
      k_frame_cepstra<<<dim3(n_inst, gx), kThreads, 0, stream>>>(tb, cv, inst0, sc); }
    { LaunchScope ls(kK_clip_stats, stream);
      k_clip_stats<<<n_inst, 256, 0, stream>>>(cv, inst0, sc, out_raw, out_clean, status); }
    return cudaGetLastError();
}

}  // namespace dys
