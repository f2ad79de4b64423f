// Pipe-rate microbenchmark for sm_100a: how many warp-instructions per clock per SM the B200
// sustains for the instruction classes the front-end kernels are made of (FFMA vs packed FFMA2,
// FADD vs FADD2, DFMA/DADD/DMUL, LDS.64/LDS.128, SHFL).  Informs the FFT kernel design only;
// not part of the product.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kIters = 4096;
constexpr int kChains = 8;

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    float2 f = make_float2(a, b);
    return *reinterpret_cast<unsigned long long*>(&f);
}

__global__ void k_ffma(float* out, float a, float b) {
    float v[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += v[i];
    if (s == 12345.f) out[0] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long v[kChains];
    const unsigned long long A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = pk(threadIdx.x + i, i);
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(A), "l"(B));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s ^= v[i];
    if (s == 12345ull) out[0] = 1.f;
}

__global__ void k_fadd(float* out, float a) {
    float v[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += v[i];
    if (s == 12345.f) out[0] = s;
}

__global__ void k_fadd2(float* out, float a) {
    unsigned long long v[kChains];
    const unsigned long long A = pk(a, a);
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = pk(threadIdx.x + i, i);
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(A));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s ^= v[i];
    if (s == 12345ull) out[0] = 1.f;
}

__global__ void k_dfma(float* out, double a, double b) {
    double v[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += v[i];
    if (s == 12345.0) out[0] = float(s);
}

__global__ void k_dadd(float* out, double a) {
    double v[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += v[i];
    if (s == 12345.0) out[0] = float(s);
}

template <int W>
__global__ void k_lds(float* out, int stride) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const unsigned base = unsigned(__cvta_generic_to_shared(buf)) + ((threadIdx.x * stride * W * 4) & 8191);
    float s = 0;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const unsigned a = base + ((u * 32 * W * 4) & 8191);
            float x, y, z, w;
            if constexpr (W == 1) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); s += x; }
            if constexpr (W == 2) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); s += x; }
            if constexpr (W == 4) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); s += x; }
        }
    }
    if (s == 12345.f) out[0] = s;
}

__global__ void k_shfl(float* out) {
    float v[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + u);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += v[i];
    if (s == 12345.f) out[0] = s;
}

template <typename F>
void run(const char* name, double warp_instr_per_thread_iter, double flop_per_instr_lane, F launch) {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int ctas = p.multiProcessorCount * 4, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(ctas, threads);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        launch(ctas, threads);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double warps = double(ctas) * threads / 32;
    const double winstr = warps * kIters * warp_instr_per_thread_iter;
    const double per_s = winstr / (best * 1e-3);
    const double per_clk_sm = per_s / (double(clk_khz) * 1e3) / p.multiProcessorCount;
    printf("%-10s %8.3f ms  %7.2f Gwarp-instr/s  %6.3f warp-instr/clk/SM (at %d MHz nominal)  %8.2f TFLOP/s\n", name, best,
           per_s / 1e9, per_clk_sm, clk_khz / 1000, per_s * 32 * flop_per_instr_lane / 1e12);
}

int main() {
    float* out;
    CK(cudaMalloc(&out, 64));
    run("FFMA", 4 * kChains, 2, [&](int g, int t) { k_ffma<<<g, t>>>(out, 1.0001f, 0.5f); });
    run("FFMA2", 4 * kChains, 4, [&](int g, int t) { k_ffma2<<<g, t>>>(out, 1.0001f, 0.5f); });
    run("FADD", 4 * kChains, 1, [&](int g, int t) { k_fadd<<<g, t>>>(out, 0.5f); });
    run("FADD2", 4 * kChains, 2, [&](int g, int t) { k_fadd2<<<g, t>>>(out, 0.5f); });
    run("DFMA", 4 * kChains, 2, [&](int g, int t) { k_dfma<<<g, t>>>(out, 1.0001, 0.5); });
    run("DADD", 4 * kChains, 1, [&](int g, int t) { k_dadd<<<g, t>>>(out, 0.5); });
    run("LDS.32", 32, 0, [&](int g, int t) { k_lds<1><<<g, t>>>(out, 1); });
    run("LDS.64", 32, 0, [&](int g, int t) { k_lds<2><<<g, t>>>(out, 1); });
    run("LDS.128", 32, 0, [&](int g, int t) { k_lds<4><<<g, t>>>(out, 1); });
    run("SHFL", 4 * kChains, 0, [&](int g, int t) { k_shfl<<<g, t>>>(out); });
    CK(cudaDeviceSynchronize());
    return 0;
}
