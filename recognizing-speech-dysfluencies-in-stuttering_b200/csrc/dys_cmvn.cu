// Global feature standardisation ("CMVN") kernels.
//
// Replaces StandardScaler().fit(X) / .transform(X) of the reference
// (/root/reference/pipeline1.py:470-473, main1.py:848-852): per-feature float64 mean and
// population variance over ALL clips, scale = sqrt(var) with constant features -> 1.0.
// The per-GPU moments [n, sum(x - s), sum (x - s)^2] are what one ncclAllReduce(sum, f64, 299)
// combines across ranks; with s = the global mean (second pass) this is sklearn's corrected
// two-pass variance.  Summation order is fixed (row blocks -> ordered merge), so results are
// bit-reproducible run to run.
#include <cmath>

#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {

__global__ void __launch_bounds__(160)
k_cmvn_partial(const float* __restrict__ feats, int64_t n_rows, const double* __restrict__ shift, double* __restrict__ partials) {
    const int f = threadIdx.x;
    if (f >= kFeat) return;
    const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = int64_t(blockIdx.x) * per, r1 = min(n_rows, r0 + per);
    const double s = shift ? shift[f] : 0.0;
    double a = 0.0, b = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const double x = double(feats[r * kFeat + f]) - s;
        a += x;
        b += x * x;
    }
    partials[size_t(blockIdx.x) * 2 * kFeat + f] = a;
    partials[size_t(blockIdx.x) * 2 * kFeat + kFeat + f] = b;
}

__global__ void __launch_bounds__(160)
k_cmvn_merge(const double* __restrict__ partials, int n_partials, int64_t n_rows, double* __restrict__ acc) {
    const int f = threadIdx.x;
    if (f == 0) acc[0] = double(n_rows);
    if (f >= kFeat) return;
    double a = 0.0, b = 0.0;
    for (int p = 0; p < n_partials; ++p) {
        a += partials[size_t(p) * 2 * kFeat + f];
        b += partials[size_t(p) * 2 * kFeat + kFeat + f];
    }
    acc[1 + f] = a;
    acc[1 + kFeat + f] = b;
}

// acc holds moments about `shift` (null = 0):  mean = shift + S1/n,  var = S2/n - (S1/n)^2
__global__ void __launch_bounds__(160)
k_cmvn_finalize(const double* __restrict__ acc, const double* __restrict__ shift, double* __restrict__ mean,
                double* __restrict__ scale) {
    const int f = threadIdx.x;
    if (f >= kFeat) return;
    const double n = acc[0];
    const double m1 = acc[1 + f] / n;
    const double mu = (shift ? shift[f] : 0.0) + m1;
    const double var = fmax(acc[1 + kFeat + f] / n - m1 * m1, 0.0);
    // sklearn _is_constant_feature + _handle_zeros_in_scale
    const double eps = 2.220446049250313e-16;
    const double bound = n * eps * var + (n * mu * eps) * (n * mu * eps);
    double sc = sqrt(var);
    if (var <= bound || sc < 10.0 * eps) sc = 1.0;
    mean[f] = mu;
    scale[f] = sc;
}

__global__ void k_cmvn_apply(const float* __restrict__ feats, int64_t n_rows, const double* __restrict__ mean,
                             const double* __restrict__ scale, float* __restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_rows * kFeat) return;
    const int f = int(i % kFeat);
    // StandardScaler.transform on a float32 matrix: X -= mean_.astype(float32); X /= scale_.astype(float32)
    out[i] = __fdiv_rn(__fsub_rn(feats[i], float(mean[f])), float(scale[f]));
}

}  // namespace

cudaError_t launch_cmvn_accumulate(const float* feats, int64_t n_rows, const double* shift, double* acc, double* partials,
                                   cudaStream_t stream) {
    { LaunchScope ls(kK_cmvn_partial, stream);
      k_cmvn_partial<<<kCmvnPartials, 160, 0, stream>>>(feats, n_rows, shift, partials); }
    { LaunchScope ls(kK_cmvn_merge, stream);
      k_cmvn_merge<<<1, 160, 0, stream>>>(partials, kCmvnPartials, n_rows, acc); }
    return cudaGetLastError();
}

cudaError_t launch_cmvn_finalize(const double* acc, const double* shift, double* mean, double* scale, cudaStream_t stream) {
    LaunchScope ls(kK_cmvn_finalize, stream);
    k_cmvn_finalize<<<1, 160, 0, stream>>>(acc, shift, mean, scale);
    return cudaGetLastError();
}

cudaError_t launch_cmvn_apply(const float* feats, int64_t n_rows, const double* mean, const double* scale, float* out,
                              cudaStream_t stream) {
    const int64_t total = n_rows * kFeat;
    if (total <= 0) return cudaSuccess;
    LaunchScope ls(kK_cmvn_apply, stream);
    k_cmvn_apply<<<unsigned((total + 255) / 256), 256, 0, stream>>>(feats, n_rows, mean, scale, out);
    return cudaGetLastError();
}

}  // namespace dys
