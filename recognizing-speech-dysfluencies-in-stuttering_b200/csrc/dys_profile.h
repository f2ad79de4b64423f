// Launch accounting + optional per-kernel CUDA-event timing (behind dys_profile_* in the C ABI).
//
// Every kernel launch of the library goes through a LaunchScope: it bumps a per-kernel launch
// counter (always) and, when profiling is enabled, brackets the launch with two cudaEvents
// recorded on the launching stream, so bench.py can report the dominant kernel's average
// duration "live" without a profiler attached.
#pragma once
#include <cuda_runtime.h>

namespace dys {

enum KernelId : int {
    kK_feat_init = 0,
    kK_frame_spectra,
    kK_tuning,
    kK_frame_cepstra,
    kK_clip_stats,
    kK_clean_init,
    kK_nr_stft_mag,
    kK_nr_iir_mask,
    kK_nr_apply_ola,
    kK_quantize_pcm,
    kK_cmvn_partial,
    kK_cmvn_merge,
    kK_cmvn_finalize,
    kK_cmvn_apply,
    kK_qc_snr,
    kK_qc_hf_bins,
    kK_qc_finish,
    kK_qc_flatness,
    kK_qc_flat_reduce,
    kK_resample,
    kKernelCount
};

const char* kernel_name(int id);

class LaunchScope {
public:
    LaunchScope(int id, cudaStream_t stream);
    ~LaunchScope();
    LaunchScope(const LaunchScope&) = delete;
    LaunchScope& operator=(const LaunchScope&) = delete;
private:
    cudaEvent_t stop_;
    cudaStream_t stream_;
};

void profile_enable(bool on);
// Synchronises the recorded events and adds them into ms[] / launches[] (both [kKernelCount]).
// reset != 0 clears the counters afterwards.  Returns cudaSuccess or the failing call's error.
cudaError_t profile_read(double* ms, long long* launches, int reset);

}  // namespace dys
