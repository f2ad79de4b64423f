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
//   k_nr_stft_mag  : one warp per frame, 512-point complex fp64 FFT + real split -> complex spectrum D and |D|
//                    [frames x 513 each]; PCM-16 or float32 samples.  Also the FORWARD half of filtfilt: every warp
//                    keeps Horner sums of b |D| r^k over its frames, one 513-value row per CTA (64 or 256 frames)
//   k_nr_iir_mask  : one thread per (chunk, bin), sequential in time: chains the CTA rows into the forward
//                    state at the interval ends (no forward sweep over |D|), closed-form zero tail, backward
//                    IIR with the forward state re-derived in reverse between those check-points, sigmoid, and the
//                    7-tap time smoothing of the mask through a register window -> time-smoothed mask, in place of
//                    |D|.  NaN (0/0) raises the clip's fallback flag.  <true>: rows through a cp.async.bulk.tensor
//                    ring (opt-in, measured slower)
//   k_nr_apply_ola : one warp per frame: 33-tap frequency smoothing of the mask row (two cascaded 17-bin
//                    running sums in registers, 16 bins per lane, halos through the warp's shared-memory
//                    tile), stored spectrum * mask, inverse FFT, synthesis window; the CTA overlap-adds its
//                    frames in shared memory (ascending frame order, like librosa's __overlap_add),
//                    scales by 1 / window-sum-square, stores float32 and maxes the clip peak.  Inverse
//                    frames never touch HBM.
// Only frames that overlap the un-padded samples are touched (191 of 422 for a 3-s clip).
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cuda.h>                 // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link)

#include "dys_async.cuh"
#include "dys_fft.cuh"
#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {

// spectrum bins (x 32 lanes) whose loads are issued before the second running sum of the mask smoothing
#ifndef DYS_APPLY_XEARLY
#define DYS_APPLY_XEARLY 12
#endif

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
// Launch groups with at least this many chunks fill the machine twice over with ONE CTA per chunk: k_nr_stft_mag then
// gives a CTA 256 frames instead of 64 (its 24 KB of tables are loaded once per CTA: -2 %) and k_nr_apply_ola 24 rounds
// instead of 8 (no frames transformed twice at CTA seams: -1.6 %).  Smaller groups keep the short CTAs: a single clip's
// latency is set by how many SMs its frames spread over.
constexpr int kBigGroup = 592;
constexpr int kFramesPerCtaSmall = 64, kFramesPerCtaBig = 256;   // powers of two: the check-point tests of the sweep are masks
constexpr int kIirSegShift = 6;
constexpr int kIirSeg = 1 << kIirSegShift;           // frames per forward-IIR interval = check-point spacing of the backward sweep
static_assert(kFramesPerCtaSmall == 1 << 6 && kFramesPerCtaBig == 1 << 8, "launch_denoise passes the shifts 6 / 8 to k_nr_iir_mask");

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

struct NrTables {
    double2 tw512[16 * 32];
};
// The forward kernel (k_nr_stft_mag) is bound by the FP64 pipe and keeps its window / split factors in tables;
// the inverse kernel (k_nr_apply_ola) is bound by shared-memory traffic and derives them arithmetically (below).
struct NrFwdTables {
    double hann[kNrFft];
    double2 split[512];
};

__device__ __forceinline__ void nr_load_tables(NrTables& sm, const DeviceTables& tb, int tid, int nthreads) {
    for (int i = tid; i < 512; i += nthreads) sm.tw512[i] = tb.tw512[i];
}
__device__ __forceinline__ void nr_load_fwd_tables(NrFwdTables& sm, const DeviceTables& tb, int tid, int nthreads) {
    for (int i = tid; i < kNrFft; i += nthreads) sm.hann[i] = tb.hann1024[i];
    for (int i = tid; i < 512; i += nthreads) sm.split[i] = tb.split1024[i];
}

// Window and real-split factors without table look-ups (the FFT kernels are bound by shared-memory wavefronts):
// every lane keeps the sine/cosine of its own base angles and rotates them by compile-time constants,
//   hann[2 (lane + 32 q) + e] = 0.5 - 0.5 cos(theta_e + q pi/8),   theta_e = 2 pi (2 lane + e) / 1024,
//   split[lane + 32 q]        = exp(i (psi + q pi/16)),            psi     = 2 pi lane / 1024,
// one multiply + one fused multiply-add per component, within ~1.5 ulp of the correctly rounded table entry.
struct LaneTrig {
    double c0, s0, c1, s1;      // cos / sin of theta_0, theta_1
    double ck, sk;              // cos / sin of psi
};
// Called once per frame: hides the six values from loop-invariant code motion, which would otherwise hoist all
// 16 + 16 derived factors out of the frame loop and spill them to local memory.
__device__ __forceinline__ LaneTrig per_frame(LaneTrig t) {
    asm volatile("" : "+d"(t.c0), "+d"(t.s0), "+d"(t.c1), "+d"(t.s1), "+d"(t.ck), "+d"(t.sk));
    return t;
}
__device__ __forceinline__ LaneTrig lane_trig(const DeviceTables& tb, int lane) {
    const double2 a = tb.split1024[2 * lane], b = tb.split1024[2 * lane + 1], c = tb.split1024[lane];
    return LaneTrig{a.x, a.y, b.x, b.y, c.x, c.y};
}
// cos / sin of 2 pi i / 64 for 0 <= i < 64
__host__ __device__ constexpr double kCos64w(int i) { return i <= 32 ? kCos64(i) : kCos64(64 - i); }
__host__ __device__ constexpr double kSin64w(int i) { return i <= 32 ? kSin64(i) : -kSin64(64 - i); }

template <int Q>
__device__ __forceinline__ double2 hann_pair(const LaneTrig& t) {
    constexpr double cq = kCos64w(4 * Q), sq = kSin64w(4 * Q);
    const double ca = fma(t.c0, cq, -(t.s0 * sq)), cb = fma(t.c1, cq, -(t.s1 * sq));
    return make_double2(fma(-0.5, ca, 0.5), fma(-0.5, cb, 0.5));
}
template <int Q>
__device__ __forceinline__ double2 split_factor(const LaneTrig& t) {
    constexpr double cq = kCos64w(2 * Q), sq = kSin64w(2 * Q);
    return make_double2(fma(t.ck, cq, -(t.sk * sq)), fma(t.sk, cq, t.ck * sq));
}

// Windowed frame t of the zero-padded chunk -> STFT bins.  sink(q, D[lane + 32 q]) is called for q = 0..15 as each bin
// leaves the real-FFT split (no second register array: the accumulators of the caller stay in registers), then
// *nyq = D[512] (real).
template <bool kPcm, typename Sink>
__device__ __forceinline__ void nr_frame_stft(const NrTables& sm, const NrFwdTables& fw, double2* xbuf,
                                              const float* __restrict__ base, const int16_t* __restrict__ base_q, bool vec_ok,
                                              const NrGeom& g, int t, int lane, Sink&& sink, double* nyq) {
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
        if constexpr (kPcm) {                               // int16 / 32768: exact in float32, like librosa.load on a 16-bit WAV
            if (in0 && in1 && vec_ok) {
                const short2 pr = __ldg(reinterpret_cast<const short2*>(base_q + s));
                a = float(pr.x) * (1.0f / 32768.0f); b = float(pr.y) * (1.0f / 32768.0f);
            } else {
                if (in0) a = float(__ldg(base_q + s)) * (1.0f / 32768.0f);
                if (in1) b = float(__ldg(base_q + s + 1)) * (1.0f / 32768.0f);
            }
        } else if (in0 && in1 && vec_ok) {
            const float2 pr = __ldg(reinterpret_cast<const float2*>(base + s));
            a = pr.x; b = pr.y;
        } else {
            if (in0) a = __ldg(base + s);
            if (in1) b = __ldg(base + s + 1);
        }
        v[m] = make_double2(double(a) * fw.hann[j2], double(b) * fw.hann[j2 + 1]);
    });
    warp_fft512_rolled(v, xbuf, sm.tw512, lane);  // Z[lane + 32 q] = v[bitrev(q, 4)]
    // real split through the (now free) exchange tile: every lane publishes its Z and fetches the partner
    // Z[512 - k] of each of its bins, so X[k] replaces Z[k] in place (no second register array, no shuffles).
    static_for<16>([&](auto iq) {
        constexpr int q = decltype(iq)::value;
        xbuf[lane + 32 * q] = v[bitrev(q, 4)];
    });
    if (lane == 0) xbuf[512] = v[0];                        // Z[512] := Z[0]
    __syncwarp();
    static_for<16>([&](auto iq) {
        constexpr int q = decltype(iq)::value;
        const int k = lane + 32 * q;
        const double2 z = v[bitrev(q, 4)];
        const double2 p = lds_once(&xbuf[512 - k]);
        const double ex = z.x + p.x, ey = z.y - p.y, dx = z.x - p.x, dy = z.y + p.y;
        const double2 cs = lds_once(&fw.split[k]);
        sink(iq, make_double2(0.5 * (ex + (cs.x * dy - cs.y * dx)), 0.5 * (ey - (cs.x * dx + cs.y * dy))));
    });
    const double2 z0 = xbuf[512];
    *nyq = z0.x - z0.y;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
struct MagSmem {
    NrTables tab;
    NrFwdTables fwd;
    double2 xbuf[kWarps][kXbuf512];
};

template <bool kPcm>
__global__ void __launch_bounds__(kThreads, 2)
k_nr_stft_mag(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc, int frames_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MagSmem& sm = *reinterpret_cast<MagSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    const int t_begin = g.t_first + blockIdx.y * frames_per_cta;
    if (t_begin > g.t_last) return;
    const int t_end = min(g.t_last + 1, t_begin + frames_per_cta);
    nr_load_tables(sm.tab, tb, tid, kThreads);
    nr_load_fwd_tables(sm.fwd, tb, tid, kThreads);
    __syncthreads();
    const float* base = kPcm ? nullptr : cv.audio + cv.starts[g.clip];
    const int16_t* base_q = kPcm ? cv.audio_q + cv.starts[g.clip] : nullptr;
    const bool vec_ok = (kPcm ? (reinterpret_cast<uintptr_t>(base_q) & 3u) == 0 : (reinterpret_cast<uintptr_t>(base) & 7u) == 0) &&
                        (g.c0 & 1) == 0;
    double* mag = sc.mag + size_t(li) * sc.ta_max * kNrBinsPad;
    double2* spec = sc.spec + size_t(li) * sc.ta_max * kNrBinsPad;
    // Forward half of filtfilt([b], [1, b - 1]) -- f[t] = b A[t] + r f[t-1], r = 1 - b -- without a sweep of its own: the
    // recursion is linear, so over the CTA's interval of frames  f[end] = r^len f[start - 1] + sum_u b A[u] r^(len-1-u).
    // Each warp keeps the Horner sum over ITS frames (every 8th: acc <- r^8 acc + b |D|) while the magnitudes are still
    // in registers; when the CTA is through, the 8 sums are weighted (r^0..r^7), added in warp order through the FFT
    // tiles and stored: 4 KB per CTA instead of re-reading the |D| rows it wrote (0.27 - 0.8 MB).  One interval per
    // CTA: the only block-wide barrier sits where the warps finish anyway (a barrier every 64 frames cost 0.2 ms).
    // k_nr_iir_mask chains the intervals (they are the check-points of its backward sweep).
    const double iir_b = tb.iir_b, iir_r = 1.0 - iir_b;
    double r8 = iir_r * iir_r; r8 *= r8; r8 *= r8;
    double* part = sc.part + size_t(li) * sc.n_seg_max * kNrBinsPad;
    {
        const int t0 = t_begin, seg_len = t_end - t_begin;
        double acc[16], acc_nyq = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = 0.0;
        int last_u = -1;
        for (int u = warp; u < seg_len; u += kWarps) {
            const int t = t0 + u;
            last_u = u;
            {   // this warp's next frame starts 2048 samples further on: pull its 32 lines towards L2 while this one is transformed
                const long long s_next = (long long)(t + kWarps) * kNrHop - kNrFft / 2 - kNrPad + g.c0 + 32 * lane;
                if (t + kWarps < t_end && s_next >= 0 && s_next < g.n) {
                    if constexpr (kPcm) { if ((lane & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base_q + s_next)); }
                    else asm volatile("prefetch.global.L2 [%0];" ::"l"(base + s_next));
                }
            }
            double* row = mag + size_t(t - g.t_first) * kNrBinsPad;
            double2* srow = spec + size_t(t - g.t_first) * kNrBinsPad;
            double nyq;
            nr_frame_stft<kPcm>(sm.tab, sm.fwd, sm.xbuf[warp], base, base_q, vec_ok, g, t, lane,
                                [&](auto iq, const double2 xq) {
                                    constexpr int q = decltype(iq)::value;
                                    srow[lane + 32 * q] = xq;
                                    const double m = sqrt(xq.x * xq.x + xq.y * xq.y);
                                    row[lane + 32 * q] = m;
                                    acc[q] = fma(r8, acc[q], iir_b * m);
                                }, &nyq);
            acc_nyq = fma(r8, acc_nyq, iir_b * fabs(nyq));
            if (lane == 0) { srow[512] = make_double2(nyq, 0.0); row[512] = fabs(nyq); }
            __syncwarp();
        }
        // weight r^(seg_len - 1 - last_u) brings the warp's sum to the interval's last frame (0..7 steps; 0 frames: no share)
        double w = 0.0;
        if (last_u >= 0) { w = 1.0; for (int e = seg_len - 1 - last_u; e > 0; --e) w *= iir_r; }
        double* pub = reinterpret_cast<double*>(sm.xbuf[warp]);
        static_for<16>([&](auto iq) {
            constexpr int q = decltype(iq)::value;
            pub[lane + 32 * q] = acc[q] * w;
        });
        if (lane == 0) pub[512] = acc_nyq * w;
        __syncthreads();
        double* prow = part + size_t(blockIdx.y) * kNrBinsPad;
        for (int k = tid; k < kNrBins; k += kThreads) {
            double sum = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < kWarps; ++w2) sum += reinterpret_cast<const double*>(sm.xbuf[w2])[k];
            prow[k] = sum;
        }
    }
}

// ------------------------------------------------------------------------------------------
// filtfilt([b], [1, b - 1]) over time + sigmoid + 7-tap time smoothing, one thread per bin.
constexpr int kIirThreads = 32;
constexpr int kIirMaxSeg = 48;                       // covers ta_max <= 3072 (a full 660 000-sample chunk has 2579)

// ---- bulk-tensor (TMA) ring for the backward sweep ------------------------------------------------------------------
// The sweep walks |D| from the last frame to the first, 8 rows at a time.  With kTma the rows do not go through the
// load/store unit and registers: lane 0 asks the copy engine for [8 rows x 32 bins] boxes of the [items x frames][520]
// float64 matrix (cp.async.bulk.tensor.2d, one instruction per 2 KB tile), kIirStages tiles ahead, each landing in
// shared memory and flipping an mbarrier; the warp waits on the barrier, takes its 8 values from the tile and hands
// the stage back to the engine.  Bins past 519 (the last slab) are filled with zeros by the engine.
constexpr int kIirRows = 8;                          // rows per tile = rows per straight-line batch
#ifndef DYS_IIR_STAGES
#define DYS_IIR_STAGES 4
#endif
constexpr int kIirStages = DYS_IIR_STAGES;
constexpr int kIirTileBytes = kIirRows * kIirThreads * 8;

template <bool kTma>
__global__ void __launch_bounds__(kIirThreads)
k_nr_iir_mask(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc, int32_t* __restrict__ clean_flag,
              int seg_shift, const __grid_constant__ CUtensorMap mag_map) {
    extern __shared__ __align__(128) unsigned char iir_smem[];
    // kTma: [kIirStages tiles of 8 x 32 doubles][kIirStages mbarriers], then in both cases
    // ck[n_seg_max][kIirThreads]: f at the last frame of every interval
    double* tiles = reinterpret_cast<double*>(iir_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(iir_smem + (kTma ? kIirStages * kIirTileBytes : 0));
    double (*ck)[kIirThreads] = reinterpret_cast<double (*)[kIirThreads]>(iir_smem + (kTma ? kIirStages * kIirTileBytes + 64 : 0));
    const int li = blockIdx.x;
    const int k = blockIdx.y * kIirThreads + threadIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    if (k >= kNrBins || g.t_last < g.t_first) return;
    const int Ta = g.t_last - g.t_first + 1;
    constexpr size_t P = kNrBinsPad;
    double* col = sc.mag + size_t(li) * sc.ta_max * P + k;
    const double b = tb.iir_b, r = 1.0 - b, rinv = 1.0 / r;
    // forward: f[t] = b A[t] + (1 - b) f[t-1],  f[-1] := A[0]  (lfilter_zi steady state; A[0] = 0 when padded).
    // k_nr_stft_mag left  sum_u b A[u] r^(len-1-u)  per interval: chain them,  f[end] = r^len f[start-1] + sum.
    double prev = (g.t_first == 0) ? col[0] : 0.0;
    {
        const int seg = 1 << seg_shift;               // frames per CTA of k_nr_stft_mag: 64 or 256
        double rseg = r;                              // r^seg by squaring
        for (int e = 0; e < seg_shift; ++e) rseg *= rseg;
        const double* part = sc.part + size_t(li) * sc.n_seg_max * P + k;
        const int n_seg = (Ta + seg - 1) >> seg_shift;
        double pnext = part[0];
        for (int j = 0; j < n_seg; ++j) {
            const double pj = pnext;
            if (j + 1 < n_seg) pnext = part[size_t(j + 1) * P];
            const int len = min(seg, Ta - (j << seg_shift));
            double rl = rseg;
            if (len < seg) {                          // the last, partial interval: r^len by squaring
                rl = 1.0;
                double sq = r;
                for (int e = len; e > 0; e >>= 1, sq *= sq) if (e & 1) rl *= sq;
            }
            prev = fma(rl, prev, pj);
            ck[j][threadIdx.x] = prev;
        }
    }
    // frames t_last+1 .. Tn-1 hold zeros: f decays geometrically and the backward recursion over them,
    // started from S[Tn] := f[Tn-1], collapses to  S[t_last+1] = r F u,  u <- b + r^2 u  (m-1 times from 1).
    const int m = g.Tn - 1 - g.t_last;
    double nxt = prev;
    if (m > 0) {
        double u = 1.0;
        for (int j = 1; j < m; ++j) u = b + r * r * u;
        nxt = r * prev * u;
    }
    // frames outside the active range but inside [0, Tn) hold |D| = 0 over a positive floor: the raw mask
    // there is exactly sigmoid(-30); outside [0, Tn) the 'same' convolution pads zeros.
    const double c_pad = 1.0 / (1.0 + exp(30.0));
    auto virt = [&](int rsrc) { const int t = g.t_first + rsrc; return (t >= 0 && t < g.Tn) ? c_pad : 0.0; };
    double ft[kNrTimeTaps];
#pragma unroll
    for (int j = 0; j < kNrTimeTaps; ++j) ft[j] = tb.smooth_t[j];
    double w[kNrTimeTaps];                            // w[j] = raw mask of row i + j
#pragma unroll
    for (int j = 0; j < kNrTimeTaps - 1; ++j) w[j] = virt(Ta + j);
    w[kNrTimeTaps - 1] = 0.0;
    double fcur = prev;                               // f[Ta-1]
    bool bad = false;
    // backward: S[t] = b f[t] + (1 - b) S[t+1]; the forward state is re-derived as f[t-1] = (f[t] - b A[t]) / (1 - b)
    // (error growth (1/r)^256 = 7.7 between check-points); row i+3 of the time-smoothed mask is complete once
    // the raw mask of row i is known, and overwrites |D| in place (row i+3 was consumed three steps earlier).
    auto emit = [&](int row, double m0) {                // push the raw mask of row - 3, write the smoothed row
#pragma unroll
        for (int j = kNrTimeTaps - 1; j > 0; --j) w[j] = w[j - 1];
        w[0] = m0;
        double acc = 0.0;
#pragma unroll
        for (int bb = 0; bb < kNrTimeTaps; ++bb) acc += ft[bb] * w[kNrTimeTaps - 1 - bb];
        if (row < Ta) col[size_t(row) * P] = acc;
    };
    auto gate = [&](double A) {                           // one backward step at the current row: raw sigmoid mask
        const double S = b * fcur + r * nxt;
        nxt = S;
        const double above = (A - S) * __drcp_rn(S);     // (A - S) / S to within one ulp; 0/0 and x/0 behave alike
        const double m0 = __drcp_rn(1.0 + exp(-(above + -2.0) * 10.0));
        bad |= isnan(m0);
        return m0;
    };
    const int seg_mask = (1 << seg_shift) - 1;
    auto rewind = [&](int i_, double A) {                 // f[i_ - 1] from f[i_]
        if (i_ > 0) {
            if ((i_ & seg_mask) == 0) fcur = ck[(i_ >> seg_shift) - 1][threadIdx.x];
            else fcur = (fcur - b * A) * rinv;
        }
    };
    // prologue: the top rows one at a time until the remaining count is a multiple of 8 and no store can
    // fall outside the clip's rows; then straight-line batches of 8 (rows 8m+7 .. 8m): the next batch's loads
    // are in flight, the only branch is the check-point reload at the batch's last row, and a batch's stores
    // touch rows >= 8m + 3, all already in registers or consumed.
    int i_ = Ta - 1;
    {
        int pro = Ta & 7;
        if (pro < 3 && Ta >= pro + 8) pro += 8;
        if (Ta < 8) pro = Ta;
        for (int c = 0; c < pro; ++c, --i_) {
            const double A = col[size_t(i_) * P];
            const double m0 = gate(A);
            rewind(i_, A);
            emit(i_ + 3, m0);
        }
    }
    if constexpr (kTma) {
        if (i_ >= 7) {
            const int lane = threadIdx.x;
            const int n_batch = (i_ + 1) >> 3;                        // rows i_ .. 0 in batches of 8 (i_ + 1 is a multiple of 8)
            const int col0 = blockIdx.y * kIirThreads;
            const int row_base = li * sc.ta_max;                      // this chunk's first row in the [items x frames] matrix
            if (lane == 0) {
                for (int s_ = 0; s_ < kIirStages; ++s_) mbar_init(&bars[s_], 1);
                mbar_fence_init();
            }
            __syncwarp();
            if (lane == 0) {
                for (int j = 0; j < min(kIirStages, n_batch); ++j) {
                    mbar_expect_tx(&bars[j], kIirTileBytes);
                    tma_load_2d(tiles + j * (kIirTileBytes / 8), &mag_map, col0, row_base + i_ - 7 - 8 * j, &bars[j]);
                }
            }
            for (int j = 0; j < n_batch; ++j, i_ -= 8) {
                const int st_ = j % kIirStages;
                mbar_wait(&bars[st_], unsigned(j / kIirStages) & 1u);
                const double* tile = tiles + st_ * (kIirTileBytes / 8) + lane;
                double a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = tile[(7 - u) * kIirThreads];        // a[u] = row i_ - u
                __syncwarp();                                         // every lane holds its 8 values: the stage is free again
                if (lane == 0 && j + kIirStages < n_batch) {
                    fence_proxy_async();                              // our reads before the engine's writes
                    mbar_expect_tx(&bars[st_], kIirTileBytes);
                    tma_load_2d(tiles + st_ * (kIirTileBytes / 8), &mag_map, col0, row_base + i_ - 7 - 8 * kIirStages, &bars[st_]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double m0 = gate(a[u]);
                    if (u < 7) fcur = (fcur - b * a[u]) * rinv;
                    else rewind(i_ - 7, a[7]);
                    emit(i_ - u + 3, m0);
                }
            }
        }
    } else if (i_ >= 7) {
        double a[8], an[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = col[size_t(i_ - u) * P];
        for (; i_ >= 7; i_ -= 8) {
            const bool more = i_ >= 15;
            if (more) {
#pragma unroll
                for (int u = 0; u < 8; ++u) an[u] = col[size_t(i_ - 8 - u) * P];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double m0 = gate(a[u]);
                if (u < 7) fcur = (fcur - b * a[u]) * rinv;      // rows 8m+7 .. 8m+1 are never check-point rows
                else rewind(i_ - 7, a[7]);
                emit(i_ - u + 3, m0);
            }
            if (more) {
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = an[u];
            }
        }
    }
    for (int v_ = -1; v_ >= -3; --v_) emit(v_ + 3, virt(v_));
    if (bad) atomicOr(&clean_flag[g.clip], 1);
}

// ------------------------------------------------------------------------------------------
// Mask row layout in the warp's tile: bin k (-16 <= k <= 527) at double 18 (floor(k / 16) + 1) + k mod 16.
__device__ __forceinline__ int seg_idx(int k) { return ((k >> 4) + 1) * 18 + (k & 15); }

// Per-warp tile: mask row / first running sum / smoothed mask (34 segments of 18 doubles, 4896 B), then the FFT
// exchange (8448 B), then the windowed output frame (8192 B).
constexpr int kApplyTile = 576;                        // double2 units = 9216 B

template <int W>
struct ApplySmem {
    NrTables tab;
    double wss[kNrHop];                                // 1 / window-sum-square
    double carry[(W * 32) / kNrHop][3][kNrHop];        // double-buffered only when two thread groups share it
    float red[W];
    int bad;
    double2 xbuf[W][kApplyTile];
};

template <int W>
__global__ void __launch_bounds__(W * 32, W == 8 ? 2 : 1)
k_nr_apply_ola(const DeviceTables tb, const ClipView cv, int cpc, int item0, NrScratch sc, double prop, int blocks_per_cta,
               float* __restrict__ clean, float* __restrict__ clean_peak, int32_t* __restrict__ clean_flag) {
    constexpr int kT = W * 32;
    constexpr int kGroups = kT / kNrHop;               // thread groups of 256 for the overlap-add
    static_assert(kT % kNrHop == 0 && W % kGroups == 0, "overlap-add needs whole groups of 256 threads");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ApplySmem<W>& sm = *reinterpret_cast<ApplySmem<W>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x;
    const NrGeom g = nr_geom(cv, item0 + li, cpc);
    if (g.out_len <= 0) return;
    // output blocks of 256 samples in padded-chunk coordinates: block h = [256 h, 256 h + 256)
    const int h_first = kNrPad / kNrHop, h_last = (kNrPad + g.out_len - 1) / kNrHop;
    const int hb = h_first + blockIdx.y * blocks_per_cta;
    if (hb > h_last) return;
    const int he = min(h_last + 1, hb + blocks_per_cta);      // exclusive
    nr_load_tables(sm.tab, tb, tid, kT);
    for (int i = tid; i < kNrHop; i += kT) sm.wss[i] = 1.0 / tb.wss[i];     // one reciprocal per CTA and sample phase
    if (tid == 0) sm.bad = 0;
    __syncthreads();
    const double* tsm = sc.mag + size_t(li) * sc.ta_max * kNrBinsPad;
    const double2* spec = sc.spec + size_t(li) * sc.ta_max * kNrBinsPad;
    float* out = clean + size_t(g.clip) * cv.clean_pitch + g.c0;
    double2* xb = sm.xbuf[warp];
    double* mrow = reinterpret_cast<double*>(xb);
    const double one_minus_prop = 1.0 - prop;
    const LaneTrig trig = lane_trig(tb, lane);
    float peak = 0.f;
    bool bad = false;
    int cbuf = 0;

    auto frame_ok = [&](int t) { return t >= g.t_first && t <= g.t_last && t <= he + 1; };
    // block h is the sum of frames h-1 .. h+2: frames hb-1 .. he+1 are needed
    for (int T0 = hb - 1; T0 - 2 < he; T0 += W) {
        const int t = T0 + warp;
        if (frame_ok(t)) {
            const double* trow = tsm + size_t(t - g.t_first) * kNrBinsPad;
            double2 mr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) mr[j] = __ldg(reinterpret_cast<const double2*>(trow + 2 * lane + 64 * j));
            const double mr512 = lane == 0 ? __ldg(trow + 512) : 0.0;
            const double2* srow = spec + size_t(t - g.t_first) * kNrBinsPad;
            const LaneTrig tr = per_frame(trig);
            if (t + W <= g.t_last && t + W <= he + 1) {         // this warp's next frame: pull both rows towards L2
                const char* nm = reinterpret_cast<const char*>(trow + size_t(W) * kNrBinsPad);
                const char* ns = reinterpret_cast<const char*>(srow + size_t(W) * kNrBinsPad);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nm + 128 * lane));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ns + 128 * lane));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ns + 4096 + 128 * lane));
                if (lane == 0) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nm + 4096));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ns + 8192));
                }
            }
            // ---- 33-tap frequency smoothing of the time-smoothed mask row ----------------------------
            // noisereduce's normalised triangle tri(16) is box17 (*) box17 / 289.  Lane L owns bins 16 L .. 16 L + 15
            // and runs both 17-bin running sums in registers; a pass needs 8 bins of halo on either side, which
            // come through the warp's tile (segment s = bins 16 s .. 16 s + 15 at doubles 18 (s + 1) .. : the two
            // pad slots keep 16-byte accesses of neighbouring lanes in different bank groups).  Every shared-memory
            // access moves 16 bytes: 230 wavefronts per frame against 324 for the scalar sliding window it replaces.
            // (Mask values lie in [0, 1]; the running sums stay within ~1e-15 of the direct sums.)
            double* rb = mrow;
            const double gain = prop * (1.0 / 289.0);
            {
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<double2*>(rb + seg_idx(2 * lane + 64 * j)) = mr[j];
                if (lane < 12) {                                // bins 512 .. 527 (only 512 is data) and the zero halo -8 .. -1
                    const int k = lane < 8 ? 512 + 2 * lane : 2 * (lane - 8) - 8;
                    *reinterpret_cast<double2*>(rb + seg_idx(k)) = make_double2(lane == 0 ? mr512 : 0.0, 0.0);
                }
            }
            __syncwarp();
            double b1[16];                                      // first running sum: B1[16 L + i] = sum of bins -8 .. +8 around it
            {
                double w[32];                                   // w[i] = mask bin 16 L - 8 + i
                const double* wl = rb + 18 * lane;
                static_for<16>([&](auto ic) {
                    constexpr int c = decltype(ic)::value;
                    constexpr int off = c < 4 ? 8 + 2 * c : (c < 12 ? 18 + 2 * (c - 4) : 36 + 2 * (c - 12));
                    const double2 v = lds_once(reinterpret_cast<const double2*>(wl + off));
                    w[2 * c] = v.x; w[2 * c + 1] = v.y;
                });
                double s_ = w[0];
#pragma unroll
                for (int j = 1; j <= 16; ++j) s_ += w[j];
                b1[0] = s_;
#pragma unroll
                for (int i = 0; i < 15; ++i) { s_ += w[i + 17] - w[i]; b1[i + 1] = s_; }      // one addition on the serial chain
                __syncwarp();                                   // every lane holds its window: the tile is free for B1
                double* own = rb + 18 * (lane + 1);
#pragma unroll
                for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(own + 2 * c) = make_double2(b1[2 * c], b1[2 * c + 1]);
                if (lane == 0) {                                // B1[-1 - u] = B1[-u] - m[8 - u]  (bins below 0 are zero)
                    double e = b1[0], bl[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { e -= w[16 - u]; bl[u] = e; }
#pragma unroll
                    for (int c = 0; c < 4; ++c) *reinterpret_cast<double2*>(rb + 8 + 2 * c) = make_double2(bl[7 - 2 * c], bl[6 - 2 * c]);
                }
                if (lane == 31) {                               // B1[512 + u] = B1[511 + u] - m[503 + u]  (bins above 512 are zero)
                    double e = b1[15], br[10];
#pragma unroll
                    for (int u = 0; u < 9; ++u) { e -= w[15 + u]; br[u] = e; }
                    br[9] = 0.0;
#pragma unroll
                    for (int c = 0; c < 5; ++c) *reinterpret_cast<double2*>(rb + 18 * 33 + 2 * c) = make_double2(br[2 * c], br[2 * c + 1]);
                }
            }
            __syncwarp();
            double2 x[16];
#if DYS_APPLY_XEARLY
            static_for<DYS_APPLY_XEARLY>([&](auto iq) {          // part of the spectrum row in flight during the second running sum
                constexpr int q = decltype(iq)::value;
                x[q] = __ldg(srow + lane + 32 * q);
            });
#endif
            {
                double B[33];                                   // B[i] = B1[16 L - 8 + i]
                static_for<4>([&](auto ic) {
                    constexpr int c = decltype(ic)::value;
                    const double2 l = lds_once(reinterpret_cast<const double2*>(rb + 18 * lane + 8 + 2 * c));
                    const double2 r = lds_once(reinterpret_cast<const double2*>(rb + 18 * (lane + 2) + 2 * c));
                    B[2 * c] = l.x; B[2 * c + 1] = l.y;
                    B[24 + 2 * c] = r.x; B[25 + 2 * c] = r.y;
                });
                B[32] = rb[18 * (lane + 2) + 8];
#pragma unroll
                for (int i = 0; i < 16; ++i) B[8 + i] = b1[i];
                __syncwarp();                                   // halos are in registers: the tile is free for the smoothed mask
                double o = B[0];
#pragma unroll
                for (int j = 1; j <= 16; ++j) o += B[j];
                double* own = rb + 18 * (lane + 1);
                double prev_out = 0.0;
#pragma unroll
                for (int i = 0; i <= 16; ++i) {
                    const double val = o * gain + one_minus_prop;      // smoothed bin 16 L + i
                    if (i & 1) *reinterpret_cast<double2*>(own + i - 1) = make_double2(prev_out, val);
                    else if (i == 16) { if (lane == 31) own[16 + 2] = val; }   // bin 512 -> segment 32, slot 0
                    prev_out = val;
                    if (i < 16) o += B[i + 17] - B[i];
                }
            }
            __syncwarp();
            static_for<16 - DYS_APPLY_XEARLY>([&](auto iq) {
                constexpr int q = decltype(iq)::value + DYS_APPLY_XEARLY;
                x[q] = __ldg(srow + lane + 32 * q);
            });
            double nyq = __ldg(&srow[512].x);
            static_for<16>([&](auto iq) {
                constexpr int q = decltype(iq)::value;
                const double mk_ = rb[18 * ((lane >> 4) + 2 * q + 1) + (lane & 15)];
                x[q].x *= mk_; x[q].y *= mk_;
            });
            nyq *= rb[18 * 33];
            // inverse real split: Z'[k] = (X[k] + conj X[512-k]) + i e^{+2 pi i k/1024} (X[k] - conj X[512-k]);
            // the inverse FFT is taken as conj(FFT(conj Z')), overall scale 1/1024.
            __syncwarp();
            static_for<16>([&](auto iq) {
                constexpr int q = decltype(iq)::value;
                xb[lane + 32 * q] = x[q];
            });
            if (lane == 0) xb[512] = make_double2(nyq, 0.0);
            __syncwarp();
            double2 v[16];
            static_for<16>([&](auto iq) {
                constexpr int q = decltype(iq)::value;
                const int k = lane + 32 * q;
                const double2 a = x[q];
                const double2 p = lds_once(&xb[512 - k]);       // X[512 - k]  (k = 0: the Nyquist bin)
                const double ex = a.x + p.x, ey = a.y - p.y, dx = a.x - p.x, dy = a.y + p.y;
                const double2 cs = split_factor<q>(tr);
                const double zr = ex - (dx * cs.y + dy * cs.x);
                const double zi = ey + (dx * cs.x - dy * cs.y);
                v[q] = make_double2(zr, -zi);
            });
            __syncwarp();
            warp_fft512_rolled(v, xb, sm.tab.tw512, lane);
            // windowed frame -> this warp's tile (sample 2 j, 2 j + 1 at double2 index j)
            static_for<16>([&](auto iq) {
                constexpr int q = decltype(iq)::value;
                constexpr int rq = bitrev(q, 4);
                const int j = lane + 32 * q;
                const double s0 = v[rq].x * (1.0 / 1024.0), s1 = -v[rq].y * (1.0 / 1024.0);
                const double2 w = hann_pair<q>(tr);
                xb[j] = make_double2(s0 * w.x, s1 * w.y);
            });
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) xb[lane + 32 * q] = make_double2(0.0, 0.0);
        }
        __syncthreads();
        // ---- overlap-add: block T0 - 2 + m gets segments 3,2,1,0 of local frames m-3 .. m (ascending frame order) ----
        {
            const int j = tid & (kNrHop - 1);
            const double* carry_in = &sm.carry[cbuf][0][0];
            double* carry_out = &sm.carry[kGroups > 1 ? cbuf ^ 1 : 0][0][0];   // one group: thread j reads before it writes
            for (int m = tid / kNrHop; m < W + 3; m += kGroups) {
                double acc = (m < 3) ? carry_in[m * kNrHop + j] : 0.0;
#pragma unroll
                for (int d = 3; d >= 0; --d) {
                    const int f = m - d;
                    if (f >= 0 && f < W) acc += reinterpret_cast<const double*>(sm.xbuf[f])[d * kNrHop + j];
                }
                if (m >= W) { carry_out[(m - W) * kNrHop + j] = acc; continue; }
                const int h = T0 - 2 + m;
                if (h < hb || h >= he) continue;
                const int s_local = h * kNrHop + j - kNrPad;
                if (s_local >= 0 && s_local < g.out_len) {
                    // librosa's istft divides by the window-sum-square in float64; multiplying by its correctly rounded
                    // reciprocal differs by at most one float64 ulp before the float32 rounding
                    const float y = float(acc * sm.wss[j]);
                    out[s_local] = y;
                    if (isfinite(y)) peak = fmaxf(peak, fabsf(y)); else bad = true;
                }
            }
            if (kGroups > 1) cbuf ^= 1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if (lane == 0) sm.red[warp] = peak;
    if (bad) sm.bad = 1;
    __syncthreads();
    if (tid == 0) {
        float mx = sm.red[0];
        for (int w2 = 1; w2 < W; ++w2) mx = fmaxf(mx, sm.red[w2]);
        atomicMax(reinterpret_cast<unsigned*>(&clean_peak[g.clip]), __float_as_uint(mx));
        if (sm.bad) atomicOr(&clean_flag[g.clip], 1);
    }
}

constexpr int kApplyWarps = 8;

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
        for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
            if (cv.audio_q) { user[i] = cv.audio_q[cv.starts[c] + i]; continue; }
            float q = rintf(cv.audio[cv.starts[c] + i] * 32768.0f);
            user[i] = int16_t(fminf(fmaxf(q, -32768.0f), 32767.0f));
        }
        return;
    }
    float pk = cv.clean_peak[c];
    if (pk < FLT_MIN) pk = 1.0f;           // librosa.util.normalize: below tiny -> left unscaled
    const float* src = cv.clean + size_t(c) * cv.clean_pitch;
    int16_t* dst = clean_q + size_t(c) * cv.clean_pitch;
    auto quant = [pk](float x) {
        const float q = rintf(__fdiv_rn(x, pk) * 32768.0f);
        return int16_t(fminf(fmaxf(q, -32768.0f), 32767.0f));
    };
    // rows of the workspace are 256-byte aligned: 4 samples per thread (16-byte load, 8-byte store)
    const int n4 = n >> 2;
    const bool user_vec = user && (reinterpret_cast<uintptr_t>(user) & 7u) == 0;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += gridDim.y * blockDim.x) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src) + i);
        short4 q;
        q.x = quant(x.x); q.y = quant(x.y); q.z = quant(x.z); q.w = quant(x.w);
        reinterpret_cast<short4*>(dst)[i] = q;
        if (user_vec) reinterpret_cast<short4*>(user)[i] = q;
        else if (user) { user[4 * i] = q.x; user[4 * i + 1] = q.y; user[4 * i + 2] = q.z; user[4 * i + 3] = q.w; }
    }
    if (blockIdx.y == 0 && threadIdx.x < (n & 3)) {
        const int i = 4 * n4 + threadIdx.x;
        const int16_t v = quant(src[i]);
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

static int nr_seg_max(int ta_max) { return (ta_max + kIirSeg - 1) / kIirSeg; }

size_t nr_scratch_bytes(int n_items, int ta_max) {
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    return al(size_t(n_items) * ta_max * kNrBinsPad * 8) + al(size_t(n_items) * ta_max * kNrBinsPad * 16) +
           al(size_t(n_items) * nr_seg_max(ta_max) * kNrBinsPad * 8);
}

void nr_scratch_carve(void* base, int n_items, int ta_max, NrScratch* out) {
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    unsigned char* p = static_cast<unsigned char*>(base);
    out->mag = reinterpret_cast<double*>(p); p += al(size_t(n_items) * ta_max * kNrBinsPad * 8);
    out->spec = reinterpret_cast<double2*>(p); p += al(size_t(n_items) * ta_max * kNrBinsPad * 16);
    out->part = reinterpret_cast<double*>(p);
    out->ta_max = ta_max;
    out->n_seg_max = nr_seg_max(ta_max);
}

cudaError_t launch_clean_init(const ClipView& cv, float* clean_peak, int32_t* clean_flag, cudaStream_t stream) {
    if (cv.n_clips <= 0) return cudaSuccess;
    LaunchScope ls(kK_clean_init, stream);
    k_clean_init<<<(cv.n_clips + 255) / 256, 256, 0, stream>>>(clean_peak, clean_flag, cv);
    return cudaGetLastError();
}

// [n_items x ta_max rows][520 columns] float64 view of the |D| scratch for cp.async.bulk.tensor; box = 8 rows x 32 bins.
// The encoder lives in the driver: fetched once through the runtime, no link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
static bool iir_tma_enabled() {
    // MEASURED (B200, 10 000 3-s clips, round 2): 4.09 - 4.15 ms per step with 2 - 4 stages against 3.97 ms for the
    // register-pipelined loads (6 / 8 stages: 4.65 / 5.49 ms).  The sweep is bound by its dependent float64 chain (exp and
    // two reciprocals per element) and wants as many resident warps as possible; the ring's shared memory costs
    // residency (24 instead of 32 single-warp CTAs per SM) and the loads were already off the critical path.
    // So the copy-engine path is opt-in: DYS_IIR_TMA=1.
    static const bool on = [] { const char* v = std::getenv("DYS_IIR_TMA"); return v && v[0] == '1'; }();
    return on;
}
static bool make_mag_map(const NrScratch& sc, int n_items, CUtensorMap* map) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cuuint64_t(kNrBinsPad), cuuint64_t(n_items) * cuuint64_t(sc.ta_max)};
    const cuuint64_t strides[1] = {cuuint64_t(kNrBinsPad) * 8};
    const cuuint32_t box[2] = {cuuint32_t(kIirThreads), cuuint32_t(kIirRows)};
    const cuuint32_t estr[2] = {1, 1};
    if (dims[1] >= (cuuint64_t(1) << 31)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, sc.mag, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_denoise(const DeviceTables& tb, const ClipView& cv, float* clean, float* clean_peak, int32_t* clean_flag,
                           int cpc, int item0, int n_items, const NrScratch& sc, float prop_decrease, cudaStream_t stream) {
    if (n_items <= 0) return cudaSuccess;
    if (cudaError_t e = ensure_dynamic_smem<kK_nr_stft_mag>(k_nr_stft_mag<false>, int(sizeof(MagSmem)))) return e;
    if (cudaError_t e = ensure_dynamic_smem<kK_nr_stft_mag + 32>(k_nr_stft_mag<true>, int(sizeof(MagSmem)))) return e;
    if (cudaError_t e = ensure_dynamic_smem<kK_nr_apply_ola>(k_nr_apply_ola<kApplyWarps>, int(sizeof(ApplySmem<kApplyWarps>)))) return e;
    if (sc.ta_max > kIirMaxSeg << kIirSegShift) return cudaErrorInvalidValue;
    // One CTA per chunk ("big") saves table loads and seam frames.  Alone, a batch below three waves of resident CTAs runs
    // a little faster with the 64-frame CTAs (tools/group_geometry_probe.py, per clip: 600 chunks 3.44 vs 3.66 us, 800: 3.24
    // vs 3.28, 888: 3.16 vs 3.10, 2 500: 3.05 vs 2.99, 10 000: 3.04 vs 2.97) -- but inside the host streaming path, where
    // three streams fill each other's tails, 592-clip chunks take 31.7 ms per 10 000 clips with one CTA per chunk and 32.7 ms
    // with the short CTAs, so the threshold stays at two waves.  DYS_GATE_BIG=0/1 forces a geometry.
    bool big = n_items >= kBigGroup;
    static const int forced = [] { const char* v = std::getenv("DYS_GATE_BIG"); return v ? (v[0] == '1' ? 1 : (v[0] == '0' ? 0 : -1)) : -1; }();
    if (forced >= 0) big = forced == 1;
    const int fpc = big ? kFramesPerCtaBig : kFramesPerCtaSmall;
    const int gy = (sc.ta_max + fpc - 1) / fpc;
    ClipView cvw = cv;
    cvw.clean = clean; cvw.clean_peak = clean_peak; cvw.clean_flag = clean_flag;
    { LaunchScope ls(kK_nr_stft_mag, stream);
      if (cv.audio_q) k_nr_stft_mag<true><<<dim3(n_items, gy), kThreads, sizeof(MagSmem), stream>>>(tb, cvw, cpc, item0, sc, fpc);
      else k_nr_stft_mag<false><<<dim3(n_items, gy), kThreads, sizeof(MagSmem), stream>>>(tb, cvw, cpc, item0, sc, fpc); }
    {
        const dim3 grid(n_items, (kNrBins + kIirThreads - 1) / kIirThreads);
        const size_t ck_bytes = size_t(sc.n_seg_max) * kIirThreads * sizeof(double);
        CUtensorMap map;
        const bool tma = iir_tma_enabled() && make_mag_map(sc, n_items, &map);
        LaunchScope ls(kK_nr_iir_mask, stream);
        if (tma) k_nr_iir_mask<true><<<grid, kIirThreads, kIirStages * kIirTileBytes + 64 + ck_bytes, stream>>>(tb, cvw, cpc, item0, sc,
                                                                                                              clean_flag, big ? 8 : 6, map);
        else { std::memset(&map, 0, sizeof(map));
               k_nr_iir_mask<false><<<grid, kIirThreads, ck_bytes, stream>>>(tb, cvw, cpc, item0, sc, clean_flag, big ? 8 : 6, map); }
    }
    // output blocks of 256 samples per chunk, split evenly over CTAs of about 16 W frames
    const int max_out = std::min(cv.max_len, kNrChunk);
    const int n_blocks = (kNrPad + std::max(max_out, 1) - 1) / kNrHop - kNrPad / kNrHop + 1;
    // A CTA that owns b blocks transforms b + 3 frames in rounds of kApplyWarps: pick the rounds per CTA that leave the
    // fewest idle warp slots over the whole chunk (a 3-s clip: 188 blocks -> 61 + 61 + 61 + 5 = 25 rounds, or one CTA of
    // 24 rounds in a big launch group; the even split 4 x 47 ran 28).
    int blocks_per_cta = kApplyWarps * 8 - 3, best_rounds = INT_MAX;
    for (int r : {big ? 24 : 8, 8, 7, 9, 6, 10, 5, 11, 12, 4}) {
        const int bpc = kApplyWarps * r - 3;
        const int full = (n_blocks - 1) / bpc, last = n_blocks - full * bpc;
        const int rounds = full * r + (last + 3 + kApplyWarps - 1) / kApplyWarps;
        if (rounds < best_rounds) { best_rounds = rounds; blocks_per_cta = bpc; }
    }
    const int n_cta = (n_blocks + blocks_per_cta - 1) / blocks_per_cta;
    { LaunchScope ls(kK_nr_apply_ola, stream);
      k_nr_apply_ola<kApplyWarps><<<dim3(n_items, n_cta), kApplyWarps * 32, sizeof(ApplySmem<kApplyWarps>), stream>>>(
          tb, cvw, cpc, item0, sc, double(prop_decrease), blocks_per_cta, clean, clean_peak, clean_flag); }
    return cudaGetLastError();
}

cudaError_t launch_quantize_pcm(const ClipView& cv, int16_t* clean_q, int16_t* pcm, const int64_t* pcm_starts,
                                cudaStream_t stream) {
    if (cv.n_clips <= 0) return cudaSuccess;
    const int gy = std::max(1, std::min(16, (cv.max_len / 4 + 255) / 256));
    LaunchScope ls(kK_quantize_pcm, stream);
    k_quantize_pcm<<<dim3(cv.n_clips, gy), 256, 0, stream>>>(cv, clean_q, pcm, pcm_starts);
    return cudaGetLastError();
}

}  // namespace dys
