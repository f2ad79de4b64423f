// Rate conversion to 16 kHz: the step between the decoder and the hot path.
//
// Replaces the resampling half of  librosa.load(path, sr=16000, mono=True)  (/root/reference/pipeline1.py:102):
// librosa decodes at the file's own rate (the corpus: 22 050 Hz) and calls soxr's HQ recipe.  soxr is un-vendored; its
// published HQ specification is restated (same restatement as oracle/resample.py, parity = statistical against the
// reference's *_raw_feats.npy, bit-level against the oracle): linear phase, pass-band to 0.9136 of the lower Nyquist,
// stop-band from the Nyquist, 126.4 dB, one Kaiser-windowed sinc designed on the grid of twice the lower rate; output
// sample m sits at input time m * sr_in / 16000, the input is zero outside the clip, length ceil(n * 16000 / sr_in).
//
//   k_resample : polyphase FIR, y[m] = sum_j h[(m down) mod up][j] x[(m down) / up + j - half].
//     A CTA stages the input span of 64 consecutive periods (one period = up' outputs = down' inputs) in shared memory;
//     lane r of a warp owns periods r and r + 32, the warp owns one group of 16 consecutive phases: every input sample
//     a lane loads (conflict-free for odd down', e.g. 441) feeds 16 FMAs, and the 16 coefficients it needs are the same
//     for all 32 lanes -- four 16-byte broadcast loads from the group's expanded table c[i][16] feed 32 FMAs per lane.
//     6 shared-memory wavefronts per 1024 FMAs; accumulation in float32 in tap order.
#include <algorithm>
#include <cmath>
#include <map>
#include <type_traits>
#include <mutex>
#include <vector>

#include "dys_async.cuh"
#include "dys_error.h"
#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr int kGroup = 16;            // phases per warp item
constexpr int kPeriods = 64;          // periods per CTA tile (two per lane)
constexpr int kRsWarps = 8;

double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = 0.25 * x * x;
    for (int k = 1; k < 200; ++k) {
        term *= q / (double(k) * double(k));
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

struct RsDesign {
    int sr_in = 0, up = 0, down = 0, half = 0, ntaps = 0;
    int gmul = 1;                       // periods fused so that up' = up * gmul is a multiple of kGroup
    int up2 = 0, down2 = 0;             // up', down'
    int n_groups = 0, wl = 0;           // phase groups per period', window length of a group (input samples)
    int tile = 0;                       // input samples staged per CTA
    std::vector<double> h;              // [up][ntaps] float64
    std::vector<float> expanded;        // [n_groups][wl][kGroup]
    std::vector<int> wstart;            // [n_groups] first input sample of a group's window, relative to the period base
    std::vector<int> i0;                // [up2] floor(p' down / up)
};

int gcd_i(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// soxr_quality_spec(SOXR_HQ): 20 bit.  Same arithmetic as oracle/resample.py::lowpass_design.
bool design(int sr_in, RsDesign* d) {
    const int sr_out = kSR;
    if (sr_in <= 0 || sr_in == sr_out) return false;
    const int g = gcd_i(sr_in, sr_out);
    d->sr_in = sr_in; d->up = sr_out / g; d->down = sr_in / g;
    if (d->up > 4096 || d->down > 8192) return false;
    const double db_per_bit = 20.0 * std::log10(2.0);
    const double rej = 20.0 * db_per_bit, att = 21.0 * db_per_bit;
    const double to3db = (1.6e-6 * rej - 7.5e-4) * rej + 0.646;
    const double fp = 1.0 - 0.05 / to3db, fs = 1.0;
    const double low = std::min(sr_in, sr_out), grid = 2.0 * low;
    const double tr = 0.5 * (fs - fp) * (low / 2.0) / (grid / 2.0);
    const double fc_n = (fs * (low / 2.0)) / (grid / 2.0) - tr;
    const double beta = 0.1102 * (att - 8.7);
    int taps = int(std::ceil((att - 7.95) / (2.285 * 2.0 * kPi * tr) + 1));
    taps = (taps + 2) / 4 * 4 + 1;
    const double half_width_s = (0.5 * (taps - 1) + 0.5) / grid;
    const double fc_hz = fc_n * grid / 2.0;
    const double t_in = half_width_s * sr_in;
    d->half = int(std::ceil(t_in));
    d->ntaps = 2 * d->half + 1;
    if (d->ntaps > 4096) return false;
    const double f = fc_hz / (sr_in / 2.0), i0b = bessel_i0(beta);
    d->h.assign(size_t(d->up) * d->ntaps, 0.0);
    for (int p = 0; p < d->up; ++p) {
        const double frac = double(p) / d->up;
        for (int j = 0; j < d->ntaps; ++j) {
            const double dd = double(j - d->half) - frac, u = dd / t_in;
            double w = 0.0;
            if (std::fabs(u) < 1.0) w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - u * u))) / i0b;
            const double a = kPi * f * dd;
            const double sinc = std::fabs(a) < 1e-300 ? 1.0 : std::sin(a) / a;
            d->h[size_t(p) * d->ntaps + j] = f * sinc * w;
        }
    }
    d->gmul = kGroup / gcd_i(d->up, kGroup);
    d->up2 = d->up * d->gmul; d->down2 = d->down * d->gmul;
    d->n_groups = d->up2 / kGroup;
    d->i0.resize(d->up2);
    for (int p = 0; p < d->up2; ++p) d->i0[p] = int((int64_t(p) * d->down) / d->up);
    int span = 0;
    for (int gi = 0; gi < d->n_groups; ++gi) span = std::max(span, d->i0[gi * kGroup + kGroup - 1] - d->i0[gi * kGroup]);
    d->wl = d->ntaps + span;
    d->wstart.resize(d->n_groups);
    d->expanded.assign(size_t(d->n_groups) * d->wl * kGroup, 0.f);
    for (int gi = 0; gi < d->n_groups; ++gi) {
        const int base = d->i0[gi * kGroup];
        d->wstart[gi] = base - d->half;
        for (int i = 0; i < d->wl; ++i)
            for (int pp = 0; pp < kGroup; ++pp) {
                const int p2 = gi * kGroup + pp;
                const int j = i - (d->i0[p2] - base);
                if (j >= 0 && j < d->ntaps)
                    d->expanded[(size_t(gi) * d->wl + i) * kGroup + pp] =
                        float(d->h[size_t((int64_t(p2) * d->down) % d->up) * d->ntaps + j]);
            }
    }
    int base_max = 0;
    for (int gi = 0; gi < d->n_groups; ++gi) base_max = std::max(base_max, d->i0[gi * kGroup]);
    d->tile = (kPeriods - 1) * d->down2 + base_max + d->wl + 1;
    return true;
}

struct RsDevice {
    const float* expanded = nullptr;
    const int* wstart = nullptr;
};

std::mutex g_mu;
std::map<int, RsDesign> g_designs;                       // by sr_in (host, shared by all devices)
std::map<std::pair<int, int>, RsDevice> g_device;        // by (device, sr_in)

const RsDesign* get_design(int sr_in) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_designs.find(sr_in);
    if (it != g_designs.end()) return it->second.up ? &it->second : nullptr;
    RsDesign d;
    const bool ok = design(sr_in, &d);
    if (!ok) d = RsDesign{};
    auto& slot = g_designs[sr_in];
    slot = std::move(d);
    return ok ? &slot : nullptr;
}

cudaError_t get_device(const RsDesign& d, RsDevice* out) {
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    std::lock_guard<std::mutex> lock(g_mu);
    auto key = std::make_pair(dev, d.sr_in);
    auto it = g_device.find(key);
    if (it != g_device.end()) { *out = it->second; return cudaSuccess; }
    float* ex = nullptr; int* ws = nullptr;
    if (cudaError_t e = cudaMalloc(&ex, d.expanded.size() * 4)) return e;
    if (cudaError_t e = cudaMalloc(&ws, d.wstart.size() * 4)) return e;
    if (cudaError_t e = cudaMemcpy(ex, d.expanded.data(), d.expanded.size() * 4, cudaMemcpyHostToDevice)) return e;
    if (cudaError_t e = cudaMemcpy(ws, d.wstart.data(), d.wstart.size() * 4, cudaMemcpyHostToDevice)) return e;
    RsDevice r; r.expanded = ex; r.wstart = ws;
    g_device[key] = r;
    *out = r;
    return cudaSuccess;
}

struct RsParams {
    int up2, down2, n_groups, wl, half, tile;
    int up, down;
};

// One group of 16 output phases for kP periods of this lane (period stride xstride input samples / ostride output samples):
// software pipeline -- the loads of tap i + 1 are issued before the kP x 16 FMAs of tap i.
template <typename TIn, int kP>
__device__ __forceinline__ void rs_group(const TIn* __restrict__ xa, int xstride, const float4* __restrict__ c4, int wl,
                                         float* __restrict__ dst, long long m0, long long ostride, long long n_out) {
    constexpr float kScale = std::is_same<TIn, int16_t>::value ? (1.0f / 32768.0f) : 1.0f;   // PCM-16: value = q / 32768 (exact)
    float acc[kP][kGroup];
#pragma unroll
    for (int h = 0; h < kP; ++h)
#pragma unroll
        for (int k = 0; k < kGroup; ++k) acc[h][k] = 0.f;
    float4 q0 = c4[0], q1 = c4[1], q2 = c4[2], q3 = c4[3];
    TIn r[kP];
#pragma unroll
    for (int h = 0; h < kP; ++h) r[h] = xa[h * xstride];
    for (int i = 0; i < wl; ++i) {
        const int nx = min(i + 1, wl - 1);
        const float4 n0 = c4[4 * nx], n1 = c4[4 * nx + 1], n2 = c4[4 * nx + 2], n3 = c4[4 * nx + 3];
        TIn rn[kP];
#pragma unroll
        for (int h = 0; h < kP; ++h) rn[h] = xa[h * xstride + nx];
        const float w[kGroup] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
        for (int h = 0; h < kP; ++h) {
            const float v = float(r[h]) * kScale;
#pragma unroll
            for (int k = 0; k < kGroup; ++k) acc[h][k] = fmaf(w[k], v, acc[h][k]);
        }
        q0 = n0; q1 = n1; q2 = n2; q3 = n3;
#pragma unroll
        for (int h = 0; h < kP; ++h) r[h] = rn[h];
    }
#pragma unroll
    for (int h = 0; h < kP; ++h) {
        const long long m = m0 + h * ostride;
        if (m + kGroup <= n_out && ((reinterpret_cast<uintptr_t>(dst + m) & 15u) == 0)) {
#pragma unroll
            for (int k = 0; k < kGroup; k += 4)
                *reinterpret_cast<float4*>(dst + m + k) = make_float4(acc[h][k], acc[h][k + 1], acc[h][k + 2], acc[h][k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < kGroup; ++k)
                if (m + k < n_out) dst[m + k] = acc[h][k];
        }
    }
}

// grid (clips, tiles of 64 periods); up to 8 warps; dynamic smem: [mbarriers 128 B][input tile, kept in the input's own
// type: 16-bit PCM tiles are half the size, which doubles the resident warps][warps x (wl x 16) coefficients].
// Lane r of a warp owns periods r and r + 32 of the tile: 2 x 16 accumulators, so every 16-byte broadcast of
// coefficients feeds 8 FMAs per lane and the kernel is bound by the FMA pipe, not by shared memory.
// The tile (57 - 114 KB of contiguous input) and every group's coefficient table (18 KB) are moved by the copy engine
// (cp.async.bulk + mbarrier): one instruction each instead of ~100 dependent load/store rounds per thread.
template <typename TIn>
__global__ void __launch_bounds__(kRsWarps * 32)
k_resample(const TIn* __restrict__ in, const int64_t* __restrict__ in_starts, const int32_t* __restrict__ in_lengths,
           float* __restrict__ out, const int64_t* __restrict__ out_starts, const float* __restrict__ expanded,
           const int* __restrict__ wstart, RsParams P) {
    extern __shared__ __align__(128) unsigned char rs_smem[];
    constexpr int kPer16 = 16 / int(sizeof(TIn));                   // elements per 16 bytes
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(rs_smem);        // [0] tile, [1 + warp] coefficient tables
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, n_warps = blockDim.x >> 5;
    const int n_in = in_lengths[c];
    if (n_in <= 0) return;
    const long long n_out = ((long long)n_in * P.up + P.down - 1) / P.down;
    const long long per0 = (long long)blockIdx.y * kPeriods;       // first period of this tile
    if (per0 * P.up2 >= n_out) return;
    const long long in0 = per0 * P.down2 - P.half;                 // input index of the tile's first element
    const TIn* src = in + in_starts[c];
    // the tile sits in shared memory with the same 16-byte phase as its source, so that whole 16-byte units can be bulk-copied
    const int lead = int((reinterpret_cast<uintptr_t>(src + in0) & 15u) / sizeof(TIn));
    TIn* xt = reinterpret_cast<TIn*>(rs_smem + 128) + lead;          // xt[i] = input sample in0 + i
    float* ct = reinterpret_cast<float*>(rs_smem + 128 + ((size_t(P.tile + kPer16) * sizeof(TIn) + 15) & ~size_t(15))) +
                warp * (P.wl * kGroup);
    const long long lo = max(in0, 0LL), hi = min(in0 + P.tile, (long long)n_in);       // valid input range of the tile
    // 16-byte aligned interior [alo, ahi) of [lo, hi) goes through the copy engine, the rest (zero padding, ragged ends) by hand
    long long alo = lo + ((kPer16 - int((reinterpret_cast<uintptr_t>(src + lo) & 15u) / sizeof(TIn))) % kPer16);
    long long ahi = hi - int((reinterpret_cast<uintptr_t>(src + hi) & 15u) / sizeof(TIn));
    if (ahi <= alo) { alo = hi; ahi = hi; }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        for (int w = 0; w < n_warps; ++w) mbar_init(&bars[1 + w], 1);
        mbar_fence_init();
        if (ahi > alo) {
            // the engine takes at most 2^20 - 16 bytes per copy: tiles are far below that
            mbar_expect_tx(&bars[0], unsigned((ahi - alo) * sizeof(TIn)));
            bulk_copy_g2s(xt + (alo - in0), src + alo, unsigned((ahi - alo) * sizeof(TIn)), &bars[0]);
        }
    }
    for (long long s_ = in0 + tid; s_ < in0 + P.tile; s_ += blockDim.x) {
        if (s_ >= alo && s_ < ahi) { s_ += ((ahi - s_ - 1) / blockDim.x) * blockDim.x; continue; }   // skip the engine's part
        xt[s_ - in0] = (s_ >= lo && s_ < hi) ? __ldg(src + s_) : TIn(0);
    }
    __syncthreads();                                                 // barriers initialised, hand-written part in place
    if (ahi > alo) mbar_wait(&bars[0], 0);
    float* dst = out + out_starts[c];
    // the clip's last tile often holds fewer than 33 periods (a 3-s clip: 64 + 64 + 22): then only the lanes' first period
    // exists and the second accumulator set is skipped
    const bool two = (per0 + 32) * P.up2 < n_out;
    unsigned phase = 0;
    for (int g = warp; g < P.n_groups; g += n_warps, phase ^= 1u) {
        __syncwarp();                                                // every lane is done with the previous table
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bars[1 + warp], unsigned(P.wl * kGroup * 4));
            bulk_copy_g2s(ct, expanded + size_t(g) * P.wl * kGroup, unsigned(P.wl * kGroup * 4), &bars[1 + warp]);
        }
        const TIn* xa = xt + lane * P.down2 + (__ldg(wstart + g) + P.half);     // window of period `lane` for this group
        mbar_wait(&bars[1 + warp], phase);
        const long long m0 = (per0 + lane) * P.up2 + (long long)g * kGroup;
        if (two) rs_group<TIn, 2>(xa, 32 * P.down2, reinterpret_cast<const float4*>(ct), P.wl, dst, m0, 32LL * P.up2, n_out);
        else rs_group<TIn, 1>(xa, 32 * P.down2, reinterpret_cast<const float4*>(ct), P.wl, dst, m0, 32LL * P.up2, n_out);
    }
}

}  // namespace

int64_t resample_table_host(int sr_in, double* h_out, int64_t max_elems, int32_t* meta) {
    const RsDesign* d = get_design(sr_in);
    if (!d) return 0;
    const int64_t n = int64_t(d->up) * d->ntaps;
    if (meta) { meta[0] = d->up; meta[1] = d->down; meta[2] = d->half; meta[3] = d->ntaps; }
    if (h_out) {
        if (max_elems < n) return -1;
        std::copy(d->h.begin(), d->h.end(), h_out);
    }
    return n;
}

cudaError_t launch_resample(const float* in_f32, const int16_t* in_q16, int sr_in, const int64_t* in_starts,
                            const int32_t* in_lengths, int n_clips, int max_in_len, float* out, const int64_t* out_starts,
                            cudaStream_t stream) {
    if (n_clips <= 0 || max_in_len <= 0) return cudaSuccess;
    const RsDesign* d = get_design(sr_in);
    if (!d) return cudaErrorInvalidValue;
    RsDevice dv;
    if (cudaError_t e = get_device(*d, &dv)) return e;
    RsParams P;
    P.up2 = d->up2; P.down2 = d->down2; P.n_groups = d->n_groups; P.wl = d->wl; P.half = d->half; P.tile = d->tile;
    P.up = d->up; P.down = d->down;
    const size_t esz = in_q16 ? 2 : 4;
    const size_t tile_b = 128 + ((size_t(d->tile) * esz + 16 + 15) & ~size_t(15)), table_b = size_t(d->wl) * kGroup * 4;
    if (tile_b + table_b > 227 * 1024) return cudaErrorInvalidValue;
    const int n_warps = int(std::max<size_t>(1, std::min<size_t>(std::min(kRsWarps, d->n_groups), (227 * 1024 - tile_b) / table_b)));
    const size_t smem = tile_b + size_t(n_warps) * table_b;
    static std::mutex attr_mu;
    {
        std::lock_guard<std::mutex> lock(attr_mu);
        if (cudaError_t e = in_q16 ? cudaFuncSetAttribute(k_resample<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))
                                   : cudaFuncSetAttribute(k_resample<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) return e;
    }
    const long long max_out = ((long long)max_in_len * d->up + d->down - 1) / d->down;
    const int tiles = int((max_out + (long long)kPeriods * d->up2 - 1) / ((long long)kPeriods * d->up2));
    LaunchScope ls(kK_resample, stream);
    if (in_q16) k_resample<int16_t><<<dim3(n_clips, tiles), n_warps * 32, smem, stream>>>(in_q16, in_starts, in_lengths, out, out_starts,
                                                                                           dv.expanded, dv.wstart, P);
    else k_resample<float><<<dim3(n_clips, tiles), n_warps * 32, smem, stream>>>(in_f32, in_starts, in_lengths, out, out_starts,
                                                                                  dv.expanded, dv.wstart, P);
    return cudaGetLastError();
}

}  // namespace dys
