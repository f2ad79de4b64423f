// Warp-resident FFTs for the dysfluency front-end (sm_100a).
//
// One warp transforms one STFT frame.  A length-N complex FFT (N = 32*R, R = 32 for the
// 2048-sample feature frames in fp32, R = 16 for the 1024-sample spectral-gate frames in
// fp64) is split as  j = lane + 32*m,  k = kA + R*kB :
//     pass 1: R-point FFT over m in registers (each lane owns one residue class),
//     twiddle W_N^(lane*kA), one transposition through a padded shared-memory tile,
//     pass 2: 32-point FFT over the lanes' residue, again entirely in registers.
// The transform leaves  Z[lane + 32*q]  in register bitrev(q) -- the same "stride-32" ownership
// the input had (the bit reversal is a compile-time renaming) -- so real-FFT split, masking and
// the inverse transform chain without re-layout, and every shared-memory access in the
// exchange is conflict-free (row pad 33).  Complex additions of the fp32 transform are packed
// FADD2 (add.f32x2, Blackwell), which halves their issue slots.
//
// Replaces: numpy.fft.rfft as called by librosa.stft inside librosa.feature.mfcc /
// chroma_stft (reference pipeline1.py:216,227) and by noisereduce (pipeline1.py:140).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include <utility>

namespace dys {

template <typename T> struct cx_of;
template <> struct cx_of<float>  { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename T> using cx = typename cx_of<T>::type;

template <typename T> __device__ __forceinline__ cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
// Blackwell packed fp32: one FADD2 adds both halves of a complex number (sm_100a add.f32x2, IEEE round-to-nearest
// per lane like two FADDs).  The FMA pipe time is the same, the issue slots are halved -- and the fp32 FFT kernel is
// issue-bound.
#ifndef DYS_NO_F32X2
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
// element-wise product (window multiply): one FMUL2
__device__ __forceinline__ float2 vmul(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#else
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}

// Shared-memory loads the compiler may neither duplicate nor re-issue: under register pressure ptxas
// re-materialises plain LDS (two copies of every partner / table load in the real-FFT split), which
// costs wavefronts on the unit these kernels are bound by.
// Ordering: the asm is volatile but carries no "memory" clobber on purpose (a clobber would also pin every
// independent load and store around it).  Every use reads data that other lanes stored BEFORE a __syncwarp() and
// is followed by a __syncwarp() before the location is overwritten; __syncwarp() is the compiler-level and
// hardware-level fence that orders the plain stores against these loads.
__device__ __forceinline__ double2 lds_once(const double2* p) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(unsigned(__cvta_generic_to_shared(p))));
    return r;
}
__device__ __forceinline__ float2 lds_once(const float2* p) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(unsigned(__cvta_generic_to_shared(p))));
    return r;
}

// ---- compile-time loops -------------------------------------------------------------------
template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- constant twiddles: cos(2*pi*i/64), i = 0..16 -----------------------------------------
__host__ __device__ constexpr double kCosQ64(int i) {
    constexpr double t[17] = {
        1.0, 0.995184726672196886245, 0.980785280403230449126, 0.956940335732208864936,
        0.923879532511286756128, 0.881921264348355029713, 0.831469612302545237079,
        0.773010453362736960811, 0.707106781186547524401, 0.634393284163645498215,
        0.555570233019602224743, 0.471396736825997648556, 0.382683432365089771728,
        0.290284677254462367636, 0.195090322016128267848, 0.0980171403295606019942, 0.0};
    return t[i];
}
// cos / sin of 2*pi*i/64 for 0 <= i <= 32
__host__ __device__ constexpr double kCos64(int i) { return i <= 16 ? kCosQ64(i) : -kCosQ64(32 - i); }
__host__ __device__ constexpr double kSin64(int i) { return i <= 16 ? kCosQ64(16 - i) : kCosQ64(i - 16); }

__host__ __device__ constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((x >> b) & 1) << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2(int n) { int b = 0; while ((1 << b) < n) ++b; return b; }

// d * W_N^I  with  W_N = exp(-2*pi*i/N),  0 <= I < N/2,  N in {2,...,64}
template <int I, int N, typename T>
__device__ __forceinline__ cx<T> mul_w(cx<T> d) {
    if constexpr (I == 0) {
        return d;
    } else if constexpr (4 * I == N) {                 // -i
        return mk<T>(d.y, -d.x);
    } else if constexpr (8 * I == N) {                 // (1 - i)/sqrt2
        constexpr T c = T(0.707106781186547524401);
        return mk<T>((d.x + d.y) * c, (d.y - d.x) * c);
    } else if constexpr (8 * I == 3 * N) {             // (-1 - i)/sqrt2
        constexpr T c = T(0.707106781186547524401);
        return mk<T>((d.y - d.x) * c, -(d.x + d.y) * c);
    } else {
        constexpr T c = T(kCos64(I * (64 / N)));
        constexpr T s = T(kSin64(I * (64 / N)));
        return mk<T>(d.x * c + d.y * s, d.y * c - d.x * s);
    }
}

// In-register radix-2 decimation-in-frequency FFT.  Output is bit-reversed:
// X[k] ends in v[bitrev(k, log2 N)].
template <int N, int LEN, typename T>
__device__ __forceinline__ void dif_stage(cx<T> (&v)[N]) {
    static_for<N / LEN>([&](auto ib) {
        constexpr int base = decltype(ib)::value * LEN;
        static_for<LEN / 2>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const cx<T> a = v[base + i];
            const cx<T> b = v[base + i + LEN / 2];
            v[base + i] = cadd(a, b);
            v[base + i + LEN / 2] = mul_w<i, LEN, T>(csub(a, b));
        });
    });
    if constexpr (LEN > 2) dif_stage<N, LEN / 2, T>(v);
}
template <int N, typename T>
__device__ __forceinline__ void fft_reg(cx<T> (&v)[N]) { dif_stage<N, N, T>(v); }

// Shared-memory footprints (in complex elements) of the exchange tiles.
constexpr int kXbuf1024 = 32 * 33;   // float2  ->  8448 B
constexpr int kXbuf512 = 16 * 33;    // double2 ->  8448 B

// ---- 1024-point complex FFT, fp32, one warp ------------------------------------------------
// in : v[m] = z[lane + 32 m]          out: Z[lane + 32 q] = v[bitrev(q, 5)]  (static renaming is free)
// xbuf: 32x33 float2 private to the warp; tw[kA*32 + l] = W_1024^(l*kA)
// The two 32-point register passes share ONE copy of the butterfly code (a 2-trip rolled loop), which keeps the
// frame loop inside the instruction cache (the fully unrolled form was 116 KB of SASS, I-cache hit rate 54 %).
__device__ __forceinline__ void warp_fft1024_rolled(float2 (&v)[32], float2* xbuf, const float2* __restrict__ tw, int lane) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fft_reg<32, float>(v);
        if (pass == 0) {
            static_for<32>([&](auto ik) {
                constexpr int kA = decltype(ik)::value;
                constexpr int r = bitrev(kA, 5);
                float2 val = v[r];
                if constexpr (kA != 0) val = cmul(val, tw[kA * 32 + lane]);
                xbuf[lane * 33 + kA] = val;
            });
            __syncwarp();
            static_for<32>([&](auto il) {
                constexpr int l = decltype(il)::value;
                v[l] = xbuf[l * 33 + lane];
            });
            __syncwarp();
        }
    }
}

// ---- 512-point complex FFT, fp64, one warp -------------------------------------------------
// in : v[m] = z[lane + 32 m], m < 16   out: Z[lane + 32 q] = v[bitrev(q, 4)], q < 16
// xbuf: 16x33 double2 private to the warp; tw512[kA*32 + l] = W_512^(l*kA).
// 16-point pass, exchange + radix-2 combine across half-warps (upper half * W_32^l', a compile-time constant), 16-point
// pass (one rolled copy of the butterfly code).
__device__ __forceinline__ void warp_fft512_rolled(double2 (&v)[16], double2* xbuf, const double2* __restrict__ tw512, int lane) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fft_reg<16, double>(v);
        if (pass == 0) {
            static_for<16>([&](auto ik) {
                constexpr int kA = decltype(ik)::value;
                constexpr int r = bitrev(kA, 4);
                double2 val = v[r];
                if constexpr (kA != 0) val = cmul(val, tw512[kA * 32 + lane]);
                xbuf[kA * 33 + lane] = val;
            });
            __syncwarp();
            const int kA = lane & 15, h = lane >> 4;
            const double sgn = h ? -1.0 : 1.0;
            static_for<16>([&](auto il) {
                constexpr int l = decltype(il)::value;
                const double2 a = xbuf[kA * 33 + l];
                const double2 b = xbuf[kA * 33 + l + 16];
                double2 d = mk<double>(a.x + sgn * b.x, a.y + sgn * b.y);
                // upper half-warp: * W_32^l (a compile-time constant); lower: * 1.  Register selects instead of a
                // 16-byte shared-memory load whose 32 lanes fetch only two distinct values (4 wavefronts each).
                if constexpr (l != 0) {
                    const double c = h ? kCos64(2 * l) : 1.0, s = h ? -kSin64(2 * l) : 0.0;
                    d = mk<double>(d.x * c - d.y * s, d.x * s + d.y * c);
                }
                v[l] = d;
            });
            __syncwarp();
        }
    }
}

}  // namespace dys
