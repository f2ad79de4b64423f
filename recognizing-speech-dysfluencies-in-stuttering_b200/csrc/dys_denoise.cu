// Spectral-gate denoise kernels (float64): clip -> denoised float32 clip + peak.
//
// Replaces  nr.reduce_noise(y=y, sr=sr)  of the reference's clean_audio_and_cache
// (/root/reference/pipeline1.py:140; main1.py:605 with prop_decrease=0.8), i.e. noisereduce's
// SpectralGateNonStationary with its defaults: zero-pad 30 000 samples, STFT 1024/256 (periodic
// Hann, centred, zero pad), |D| smoothed over time by filtfilt([b],[1,b-1]) (forward + backward
// one-pole IIR), sigmoid((|D|-S)/S - 2) * 10), 33x7 triangular smoothing of the mask, D * mask,
// ISTFT (Hann synthesis, overlap-add / window-sum-square), crop.  The reference runs this in
// float64; so do these kernels, because the PCM-16 quantiser that follows (pipeline1.py:142)
// turns float32-level errors into LSB flips.
//
//   k_nr_stft_mag    : one warp per frame, 512-point complex fp64 FFT  -> |D|           [frames x 513]
//   k_nr_iir_mask    : one thread per (chunk, bin): forward IIR, closed-form zero tail, backward IIR,
//                      sigmoid -> raw mask (in place); NaN (0/0) raises the clip's fallback flag
//   k_nr_smooth      : 7-tap time then 33-tap frequency triangular smoothing, prop_decrease blend
//   k_nr_apply_istft : one warp per frame: FFT again, * mask, inverse FFT, synthesis window
//   k_nr_overlap_add : one thread per output sample: 4-frame overlap-add (ascending frame order),
//                      / window-sum-square, cast to float32, clip peak via atomicMax
// Only frames that overlap the un-padded samples are touched (191 of 422 for a 3-s clip).
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "dys_fft.cuh"
#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kFramesPerCta = 64;

struct NrGeom {
    int clip, n, c0, out_len, L, Tn, t_first, t_last;
    bool valid;
};

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

__device__ __forceinline__ NrGeom nr_geom(const ClipView& cv, int item, int cpc) {
    NrGeom g;
    g.clip = item / cpc;
    const int ch = item - g.clip * cpc;
    int n = cv.lengths[g.clip];
    if (n < 0 || n > cv.max_len) n = 0;
    g.n = n;
    g.valid = n > 0;
    int chunk_len;
    if (n <= kNrChunk) {                 // SpectralGate.get_traces: a single padded chunk of the whole clip
        g.c0 = 0; chunk_len = n;
        if (ch > 0) g.valid = false;
    } else {                             // 600 000-sample chunks, each padded with 30 000 neighbouring samples
        g.c0 = ch * kNrChunk; chunk_len = kNrChunk;
        if (g.c0 >= n) g.valid = false;
    }
    g.L = chunk_len + 2 * kNrPad;
    g.Tn = 1 + g.L / kNrHop;
    g.out_len = min(n, g.c0 + chunk_len) - g.c0;
    const int lo = max(0, g.c0 - kNrPad), hi = min(n, g.c0 + chunk_len + kNrPad);
    const int pa = lo - g.c0 + kNrPad, pb = hi - g.c0 + kNrPad;        // data extent in padded-chunk coordinates
    g.t_first = max(0, floor_div(pa - 512, kNrHop) + 1);               // frame t covers [256 t - 512, 256 t + 512)
    g.t_last = min(g.Tn - 1, (pb + 512 + kNrHop - 1) / kNrHop - 1);
    if (!g.valid) { g.t_first = 0; g.t_last = -1; g.out_len = 0; }
    return g;
}

struct NrSmem {
    double hann[kNrFft];
    double2 tw512[16 * 32];
    double2 tw32h[32];
    double2 split[512];
    double2 xbuf[kWarps][kXbuf512];
};

__device__ __forceinline__ void nr_load_tables(NrSmem& sm, const DeviceTables& tb, int tid) {
    for (int i = tid; i < kNrFft; i += kThreads) sm.hann[i] = tb.hann1024[i];
    for (int i = tid; i < 512; i += kThreads) { sm.tw512[i] = tb.tw512[i]; sm.split[i] = tb.split1024[i]; }
    if (tid < 32) sm.tw32h[tid] = tb.tw32h[tid];
}

// Windowed frame t of the zero-padded chunk -> STFT bins.  On return x[q] = D[lane + 32 q] (q < 16)
// and *nyq = D[512] (real).
__device__ __forceinline__ void nr_frame_stft(const NrSmem& sm, double2* xbuf, const float* __restrict__ base, bool vec_ok,
                                              const NrGeom& g, int t, int lane, double2 (&x)[16], double* nyq) {
    double2 v[16];
    const int p0 = t * kNrHop - kNrFft / 2;                 // padded-chunk coordinate of the frame's first sample
    static_for<16>([&](auto im) {
        constexpr int m = decltype(im)::value;
        const int j2 = 2 * (lane + 32 * m);
        const int p = p0 + j2;
        const int s = p - kNrPad + g.c0;                    // clip-relative sample index
        float a = 0.f, b = 0.f;
        const bool in0 = p >= 0 && p < g.L && s >= 0 && s < g.n;
        const bool in1 = p + 1 >= 0 && p + 1 < g.L && s + 1 >= 0 && s + 1 < g.n;
        if (in0 && in1 && vec_ok) {
            const float2 pr = __ldg(reinterpret_cast<const float2*>(base + s));
            a = pr.x; b = pr.y;
        } else {
            if (in0) a = __ldg(base + s);
            if (in1) b = __ldg(base + s + 1);
        }
        v[m] = make_double2(double(a) * sm.hann[j2], double(b) * sm.hann[j2 + 1]);
    });
    warp_fft512(v, xbuf, sm.tw512, sm.tw32h, lane);
    const int src_lane = (32 - lane) & 31;
    static_for<16>([&](auto iq) {
        constexpr int q = decltype(iq)::value;
        const int k = lane + 32 * q;
        const double2 z = v[q];
        double2 p;
        p.x = __shfl_sync(0xffffffffu, v[15 - q].x, src_lane);
        p.y = __shfl_sync(0xffffffffu, v[15 - q].y, src_lane);
        if (lane == 0) p = v[(16 - q) & 15];
        const double ex = z.x + p.x, ey = z.y - p.y, dx = z.x - p.x, dy = z.y + p.y;
        const double2 cs = sm.split[k];
        x[q] = make_double2(0.5 * (ex + (cs.x * dy - cs.y * dx)), 0.5 * (ey - (cs.x * dx + cs.y * dy)));
    });
    const double z0x = __shfl_sync(0xffffffffu, v[0].x, 0), z0y = __shfl_sync(0xffffffffu, v[0].y, 0);
    *nyq = z0x - z0y;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
k_nr_stft_mag(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NrSmem& sm = *reinterpret_cast<NrSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    const int t_begin = g.t_first + blockIdx.y * kFramesPerCta;
    if (t_begin > g.t_last) return;
    const int t_end = min(g.t_last + 1, t_begin + kFramesPerCta);
    nr_load_tables(sm, tb, tid);
    __syncthreads();
    const float* base = cv.audio + cv.starts[g.clip];
    const bool vec_ok = (reinterpret_cast<uintptr_t>(base) & 7u) == 0 && (g.c0 & 1) == 0;
    double* mag = sc.mag + size_t(li) * sc.ta_max * kNrBinsPad;
    for (int t = t_begin + warp; t < t_end; t += kWarps) {
        double2 x[16];
        double nyq;
        nr_frame_stft(sm, sm.xbuf[warp], base, vec_ok, g, t, lane, x, &nyq);
        double* row = mag + size_t(t - g.t_first) * kNrBinsPad;
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value;
            row[lane + 32 * q] = sqrt(x[q].x * x[q].x + x[q].y * x[q].y);
        });
        if (lane == 0) row[512] = fabs(nyq);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_nr_iir_mask(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc, int32_t* __restrict__ clean_flag) {
    const int li = blockIdx.x;
    const int k = blockIdx.y * 128 + threadIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    if (k >= kNrBins || g.t_last < g.t_first) return;
    const int Ta = g.t_last - g.t_first + 1;
    double* mag = sc.mag + size_t(li) * sc.ta_max * kNrBinsPad + k;
    double* fwd = sc.fwd + size_t(li) * sc.ta_max * kNrBinsPad + k;
    const double b = tb.iir_b, r = 1.0 - b;
    // forward: f[t] = b A[t] + (1 - b) f[t-1],  f[-1] := A[0]  (lfilter_zi steady state; A[0] = 0 when padded)
    double prev = (g.t_first == 0) ? mag[0] : 0.0;
    for (int i = 0; i < Ta; ++i) {
        prev = b * mag[size_t(i) * kNrBinsPad] + r * prev;
        fwd[size_t(i) * kNrBinsPad] = prev;
    }
    // frames t_last+1 .. Tn-1 hold zeros: f decays geometrically and the backward recursion over them,
    // started from S[Tn] := f[Tn-1], collapses to  S[t_last+1] = r F u,  u <- b + r^2 u  (m-1 times from 1).
    const int m = g.Tn - 1 - g.t_last;
    double nxt = prev;
    if (m > 0) {
        double u = 1.0;
        for (int i = 1; i < m; ++i) u = b + r * r * u;
        nxt = r * prev * u;
    }
    bool bad = false;
    for (int i = Ta - 1; i >= 0; --i) {
        const double S = b * fwd[size_t(i) * kNrBinsPad] + r * nxt;
        nxt = S;
        const double A = mag[size_t(i) * kNrBinsPad];
        const double above = (A - S) / S;
        const double m0 = 1.0 / (1.0 + exp(-(above + -2.0) * 10.0));
        bad |= isnan(m0);
        mag[size_t(i) * kNrBinsPad] = m0;
    }
    if (bad) atomicOr(&clean_flag[g.clip], 1);
}

// ------------------------------------------------------------------------------------------
constexpr int kSmoothRows = 8;
constexpr int kSmoothPitch = kNrBins + 32;      // 16 zero bins either side ('same' convolution)

__global__ void __launch_bounds__(kThreads)
k_nr_smooth(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc, double prop) {
    __shared__ double rows[kSmoothRows][kSmoothPitch];
    __shared__ double ff[kNrFreqTaps], ft[kNrTimeTaps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    const int Ta = g.t_last - g.t_first + 1;
    const int r_begin = blockIdx.y * kSmoothRows;
    if (r_begin >= Ta) return;
    if (tid < kNrFreqTaps) ff[tid] = tb.smooth_f[tid];
    if (tid < kNrTimeTaps) ft[tid] = tb.smooth_t[tid];
    __syncthreads();
    const double* m0 = sc.mag + size_t(li) * sc.ta_max * kNrBinsPad;
    double* out = sc.fwd + size_t(li) * sc.ta_max * kNrBinsPad;
    // frames outside the active range but inside [0, Tn) hold |D| = 0 over a positive floor: the raw
    // mask there is exactly sigmoid(-30); outside [0, Tn) the 'same' convolution pads zeros.
    const double c_pad = 1.0 / (1.0 + exp(30.0));
    for (int i = tid; i < kSmoothRows * kSmoothPitch; i += kThreads) {
        const int rr = i / kSmoothPitch, col = i - rr * kSmoothPitch;
        const int k = col - 16;
        const int row = r_begin + rr;
        double acc = 0.0;
        if (k >= 0 && k < kNrBins && row < Ta) {
#pragma unroll
            for (int bb = 0; bb < kNrTimeTaps; ++bb) {
                const int rsrc = row + 3 - bb;
                const int t = g.t_first + rsrc;
                double v;
                if (rsrc >= 0 && rsrc < Ta) v = m0[size_t(rsrc) * kNrBinsPad + k];
                else v = (t >= 0 && t < g.Tn) ? c_pad : 0.0;
                acc += ft[bb] * v;
            }
        }
        rows[rr][col] = acc;
    }
    __syncthreads();
    const int row = r_begin + warp;
    if (row < Ta) {
        for (int k = lane; k < kNrBins; k += 32) {
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < kNrFreqTaps; ++a) acc += ff[a] * rows[warp][k + 32 - a];
            out[size_t(row) * kNrBinsPad + k] = acc * prop + (1.0 - prop);
        }
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
k_nr_apply_istft(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NrSmem& sm = *reinterpret_cast<NrSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    const int t_begin = g.t_first + blockIdx.y * kFramesPerCta;
    if (t_begin > g.t_last) return;
    const int t_end = min(g.t_last + 1, t_begin + kFramesPerCta);
    nr_load_tables(sm, tb, tid);
    __syncthreads();
    const float* base = cv.audio + cv.starts[g.clip];
    const bool vec_ok = (reinterpret_cast<uintptr_t>(base) & 7u) == 0 && (g.c0 & 1) == 0;
    const double* mask = sc.fwd + size_t(li) * sc.ta_max * kNrBinsPad;
    double* frames = sc.frames + size_t(li) * sc.ta_max * kNrFft;
    const int src_lane = (32 - lane) & 31;
    for (int t = t_begin + warp; t < t_end; t += kWarps) {
        double2 x[16];
        double nyq;
        nr_frame_stft(sm, sm.xbuf[warp], base, vec_ok, g, t, lane, x, &nyq);
        const double* mrow = mask + size_t(t - g.t_first) * kNrBinsPad;
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value;
            const double mk_ = mrow[lane + 32 * q];
            x[q].x *= mk_; x[q].y *= mk_;
        });
        nyq *= mrow[512];
        // inverse real split: Z'[k] = (X[k] + conj X[512-k]) + i e^{+2 pi i k/1024} (X[k] - conj X[512-k]);
        // the inverse FFT is taken as conj(FFT(conj Z')), overall scale 1/1024.
        double2 v[16];
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value;
            const int k = lane + 32 * q;
            const double2 a = x[q];
            double2 p;
            p.x = __shfl_sync(0xffffffffu, x[15 - q].x, src_lane);
            p.y = __shfl_sync(0xffffffffu, x[15 - q].y, src_lane);
            if (lane == 0) {
                if constexpr (q == 0) p = make_double2(nyq, 0.0);
                else p = x[16 - q];
            }
            const double ex = a.x + p.x, ey = a.y - p.y, dx = a.x - p.x, dy = a.y + p.y;
            const double2 cs = sm.split[k];
            const double zr = ex - (dx * cs.y + dy * cs.x);
            const double zi = ey + (dx * cs.x - dy * cs.y);
            v[q] = make_double2(zr, -zi);
        });
        __syncwarp();
        warp_fft512(v, sm.xbuf[warp], sm.tw512, sm.tw32h, lane);
        double2* frow = reinterpret_cast<double2*>(frames + size_t(t - g.t_first) * kNrFft);
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value;
            const int j = lane + 32 * q;
            const double s0 = v[q].x * (1.0 / 1024.0), s1 = -v[q].y * (1.0 / 1024.0);
            frow[j] = make_double2(s0 * sm.hann[2 * j], s1 * sm.hann[2 * j + 1]);
        });
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_nr_overlap_add(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc, float* __restrict__ clean,
                 float* __restrict__ clean_peak, int32_t* __restrict__ clean_flag) {
    __shared__ float s_max[8];
    __shared__ int s_bad;
    const int li = blockIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    const int s_local = blockIdx.y * 256 + threadIdx.x;
    if (blockIdx.y * 256 >= g.out_len) return;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    float mine = 0.f;
    if (s_local < g.out_len) {
        const int p = s_local + kNrPad;                       // padded-chunk coordinate
        const int t_hi = (p + 512) / kNrHop;
        const double* frames = sc.frames + size_t(li) * sc.ta_max * kNrFft;
        double acc = 0.0;
#pragma unroll
        for (int d = 3; d >= 0; --d) {                        // ascending frame index, like librosa's __overlap_add
            const int t = t_hi - d;
            if (t >= g.t_first && t <= g.t_last) acc += frames[size_t(t - g.t_first) * kNrFft + (p - t * kNrHop + 512)];
        }
        const float y = float(acc / tb.wss[p & (kNrHop - 1)]);
        clean[size_t(g.clip) * cv.clean_pitch + g.c0 + s_local] = y;
        mine = fabsf(y);
        if (!isfinite(y)) { s_bad = 1; mine = 0.f; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = fmaxf(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = s_max[0];
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_max[w]);
        atomicMax(reinterpret_cast<unsigned*>(&clean_peak[g.clip]), __float_as_uint(mx));
        if (s_bad) atomicOr(&clean_flag[g.clip], 1);
    }
}

__global__ void k_clean_init(float* clean_peak, int32_t* clean_flag, const ClipView cv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cv.n_clips) return;
    clean_peak[i] = 0.f;
    const int n = cv.lengths[i];
    clean_flag[i] = (n <= 0 || n > cv.max_len) ? 1 : 0;        // nothing to clean -> the reference's except branch
}

// clean float32 -> the int16 PCM the reference writes to clear_audio/<stem>.wav:
// librosa.util.normalize (y / max|y|, float32 division) then libsndfile's clip(lrintf(x * 32768))  (pipeline1.py:141-142).
// The clean feature branch reads clean_q back as int16 / 32768, exactly what librosa.load returns for that WAV.
__global__ void k_quantize_pcm(const ClipView cv, int16_t* __restrict__ clean_q, int16_t* __restrict__ pcm,
                               const int64_t* __restrict__ pcm_starts) {
    const int c = blockIdx.x;
    const int n = cv.lengths[c];
    if (n <= 0 || n > cv.max_len) return;
    int16_t* user = pcm ? pcm + pcm_starts[c] : nullptr;
    if (cv.clean_flag[c] != 0) {           // reference wrote no WAV for this clip; the caller's buffer gets the raw samples quantised
        if (!user) return;
        const float* src = cv.audio + cv.starts[c];
        for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
            float q = rintf(src[i] * 32768.0f);
            user[i] = int16_t(fminf(fmaxf(q, -32768.0f), 32767.0f));
        }
        return;
    }
    float pk = cv.clean_peak[c];
    if (pk < FLT_MIN) pk = 1.0f;           // librosa.util.normalize: below tiny -> left unscaled
    const float* src = cv.clean + size_t(c) * cv.clean_pitch;
    int16_t* dst = clean_q + size_t(c) * cv.clean_pitch;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        float q = rintf(__fdiv_rn(src[i], pk) * 32768.0f);
        const int16_t v = int16_t(fminf(fmaxf(q, -32768.0f), 32767.0f));
        dst[i] = v;
        if (user) user[i] = v;
    }
}

}  // namespace

int nr_chunks_of(int max_len) { return max_len <= kNrChunk ? 1 : (max_len - 1) / kNrChunk + 1; }
int nr_ta_max(int max_len) {
    // single padded chunk: only frames overlapping the clip are active  (n / 256 + 4.2 at most);
    // chunked clips (> 600 000 samples) can have every frame of a chunk active.
    if (max_len <= kNrChunk) return max_len / kNrHop + 6;
    return 1 + (kNrChunk + 2 * kNrPad) / kNrHop;
}

size_t nr_scratch_bytes(int n_items, int ta_max) {
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t f = size_t(n_items) * ta_max;
    return 2 * al(f * kNrBinsPad * 8) + al(f * kNrFft * 8);
}

void nr_scratch_carve(void* base, int n_items, int ta_max, NrScratch* out) {
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    unsigned char* p = static_cast<unsigned char*>(base);
    const size_t f = size_t(n_items) * ta_max;
    out->mag = reinterpret_cast<double*>(p); p += al(f * kNrBinsPad * 8);
    out->fwd = reinterpret_cast<double*>(p); p += al(f * kNrBinsPad * 8);
    out->frames = reinterpret_cast<double*>(p);
    out->ta_max = ta_max;
}

cudaError_t launch_clean_init(const ClipView& cv, float* clean_peak, int32_t* clean_flag, cudaStream_t stream) {
    if (cv.n_clips <= 0) return cudaSuccess;
    LaunchScope ls(kK_clean_init, stream);
    k_clean_init<<<(cv.n_clips + 255) / 256, 256, 0, stream>>>(clean_peak, clean_flag, cv);
    return cudaGetLastError();
}

cudaError_t launch_denoise(const DeviceTables& tb, const ClipView& cv, float* clean, float* clean_peak, int32_t* clean_flag,
                           int cpc, int item0, int n_items, const NrScratch& sc, float prop_decrease, cudaStream_t stream) {
    if (n_items <= 0) return cudaSuccess;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_nr_stft_mag, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(NrSmem)));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_nr_apply_istft, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(NrSmem)));
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int gy = (sc.ta_max + kFramesPerCta - 1) / kFramesPerCta;
    ClipView cvw = cv;
    cvw.clean = clean; cvw.clean_peak = clean_peak; cvw.clean_flag = clean_flag;
    { LaunchScope ls(kK_nr_stft_mag, stream);
      k_nr_stft_mag<<<dim3(n_items, gy), kThreads, sizeof(NrSmem), stream>>>(tb, cvw, cpc, item0, sc); }
    { LaunchScope ls(kK_nr_iir_mask, stream);
      k_nr_iir_mask<<<dim3(n_items, (kNrBins + 127) / 128), 128, 0, stream>>>(tb, cvw, cpc, item0, sc, clean_flag); }
    { LaunchScope ls(kK_nr_smooth, stream);
      k_nr_smooth<<<dim3(n_items, (sc.ta_max + kSmoothRows - 1) / kSmoothRows), kThreads, 0, stream>>>(tb, cvw, cpc, item0, sc,
                                                                                                      double(prop_decrease)); }
    { LaunchScope ls(kK_nr_apply_istft, stream);
      k_nr_apply_istft<<<dim3(n_items, gy), kThreads, sizeof(NrSmem), stream>>>(tb, cvw, cpc, item0, sc); }
    const int max_out = std::min(cv.max_len, kNrChunk);
    { LaunchScope ls(kK_nr_overlap_add, stream);
      k_nr_overlap_add<<<dim3(n_items, (max_out + 255) / 256), 256, 0, stream>>>(tb, cvw, cpc, item0, sc, clean, clean_peak,
                                                                               clean_flag); }
    return cudaGetLastError();
}

cudaError_t launch_quantize_pcm(const ClipView& cv, int16_t* clean_q, int16_t* pcm, const int64_t* pcm_starts,
                                cudaStream_t stream) {
    if (cv.n_clips <= 0) return cudaSuccess;
    const int gy = std::max(1, std::min(64, (cv.max_len + 255) / 256));
    LaunchScope ls(kK_quantize_pcm, stream);
    k_quantize_pcm<<<dim3(cv.n_clips, gy), 256, 0, stream>>>(cv, clean_q, pcm, pcm_starts);
    return cudaGetLastError();
}

}  // namespace dys
