// Internal launch interface between the C-ABI (dys_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "dys_tables.h"

#include <atomic>

namespace dys {

// Opt a kernel into more than 48 KB of dynamic shared memory, once per device; safe to call from several host
// threads (the attribute call is idempotent, the flag only saves the driver round trip).  Tag = one id per kernel
// (the flags are per instantiation, and two kernels may share a signature).
template <int Tag, typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes) {
    static std::atomic<bool> done[64] = {};
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    std::atomic<bool>& flag = done[dev & 63];
    if (!flag.load(std::memory_order_acquire)) {
        if (cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) return e;
        flag.store(true, std::memory_order_release);
    }
    return cudaSuccess;
}

// status bits written per clip instance (mirrors the reference's "log + zeros / None" conventions)
constexpr int kStatusShort = 1;       // T < 9 frames: librosa.feature.delta raises -> zeros(144)   (pipeline1.py:237-239)
constexpr int kStatusNonFinite = 2;   // librosa.util.valid_audio raises -> zeros(144)
constexpr int kStatusCleanFallback = 4;  // denoise/normalise failed -> raw clip used as "clean"     (pipeline1.py:385-387)
constexpr int kStatusBadLength = 8;   // length < 0 or > max_len passed by the caller

__host__ __device__ inline int frames_of(int n) { return 1 + n / kHop; }

// One "instance" = one clip on one branch (raw or clean).  Instances [0, n_clips) are the raw
// branch, [n_clips, 2 n_clips) the clean branch of clip (i - n_clips).
struct ClipView {
    const float* audio;          // raw samples (packed buffer), float32 ...
    const int16_t* audio_q;      // ... or PCM-16 (value = q / 32768, what librosa.load returns for a 16-bit WAV); one of the two is null
    const int64_t* starts;       // [n_clips] first sample of clip c in `audio`
    const int32_t* lengths;      // [n_clips]
    int n_clips;
    int max_len;                 // caller-supplied upper bound on lengths
    // clean branch (null when raw only)
    const float* clean;          // [n_clips][clean_pitch] denoised float32, before normalise/quantise
    int64_t clean_pitch;
    const float* clean_peak;     // [n_clips] max |clean|
    const int16_t* clean_q;      // [n_clips][clean_pitch] PCM-16 of the normalised clean clip (what sf.write stores)
    const int32_t* clean_flag;   // [n_clips] non-zero -> fall back to the raw samples
};

// Scratch of one feature sub-batch (all device pointers, sized for n_inst x t_max frames).
struct FeatScratch {
    float* power;        // [n_inst][t_max][kBinsPad]
    float* logmel;       // [n_inst][t_max][kMels]
    float* mfcc;         // [n_inst][t_max][kMfcc]
    float* chroma;       // [n_inst][t_max][kChroma]
    float2* peaks;       // [n_inst][t_max * kMaxPeaksPerFrame]  (pitch, mag)
    int* peak_count;     // [n_inst]
    int* lmax_enc;       // [n_inst] order-preserving int encoding of max log-mel
    int* tuning_idx;     // [n_inst]
    int t_max;
};
size_t feat_scratch_bytes(int n_inst, int t_max);
void feat_scratch_carve(void* base, int n_inst, int t_max, FeatScratch* out);

// Runs the feature pipeline for instances [inst0, inst0 + n_inst) and writes out[(inst) * 149].
// out_raw / out_clean: [n_clips][149]; status: [2 * n_clips] (or [n_clips] when raw only).
cudaError_t launch_features(const DeviceTables& tb, const ClipView& cv, int inst0, int n_inst, const FeatScratch& sc,
                            float* out_raw, float* out_clean, int32_t* status, cudaStream_t stream);

// k_feat_init + k_frame_spectra only: fills sc.power for every clip with >= 1 sample (QC metrics).
cudaError_t launch_power_only(const DeviceTables& tb, const ClipView& cv, int inst0, int n_inst, const FeatScratch& sc,
                              int32_t* status, cudaStream_t stream);

// Per-file QC scalars of the reference (pipeline1.py:151-186): out[n_clips][3] = snr_db, spectral_flatness_mean,
// high_freq_energy_ratio.  scratch layout is private to dys_qc.cu (qc_scratch_bytes).
size_t qc_scratch_bytes(int n_clips, int max_len, int n_sub);
cudaError_t launch_qc(const DeviceTables& tb, const ClipView& cv, float* out, void* scratch, size_t scratch_bytes, int n_sub,
                      cudaStream_t stream);

// Spectral-gate scratch for a sub-batch of chunks.
struct NrScratch {
    double* mag;       // [n_items][ta_max][kNrBinsPad]   |STFT|, then (in place) the time-smoothed sigmoid mask
    double2* spec;     // [n_items][ta_max][kNrBinsPad]   complex STFT (read back when the mask is applied)
    double* part;      // [n_items][n_seg_max][kNrBinsPad] forward-IIR sums, one row per CTA of k_nr_stft_mag (64 or 256 frames) -> k_nr_iir_mask
    int ta_max;
    int n_seg_max;
};
size_t nr_scratch_bytes(int n_items, int ta_max);
void nr_scratch_carve(void* base, int n_items, int ta_max, NrScratch* out);
int nr_ta_max(int max_len);            // active-frame bound for one chunk
int nr_chunks_of(int max_len);         // chunks per clip bound (1 unless max_len > 600000)

// Denoise items [item0, item0 + n_items) (item = clip * chunks_per_clip + chunk): writes
// clean[clip][...] (float32, pre-normalise), atomically maxes clean_peak[clip] and ORs clean_flag[clip].
cudaError_t launch_denoise(const DeviceTables& tb, const ClipView& cv, float* clean, float* clean_peak, int32_t* clean_flag,
                           int chunks_per_clip, int item0, int n_items, const NrScratch& sc, float prop_decrease,
                           cudaStream_t stream);
cudaError_t launch_clean_init(const ClipView& cv, float* clean_peak, int32_t* clean_flag, cudaStream_t stream);
// clean float32 -> PCM-16: always into clean_q (workspace, read by the clean feature branch), and into the caller's
// packed buffer when pcm != nullptr.
cudaError_t launch_quantize_pcm(const ClipView& cv, int16_t* clean_q, int16_t* pcm, const int64_t* pcm_starts,
                                cudaStream_t stream);

// CMVN: acc[0] = n, acc[1..149] = sum (x - shift), acc[150..298] = sum (x - shift)^2
// (float64, fixed summation order -> bit-reproducible); shift may be null (= 0).
constexpr int kCmvnPartials = 128;
constexpr int kCmvnAcc = 1 + 2 * kFeat;
cudaError_t launch_cmvn_accumulate(const float* feats, int64_t n_rows, const double* shift, double* acc, double* partials,
                                   cudaStream_t stream);
cudaError_t launch_cmvn_finalize(const double* acc, const double* shift, double* mean, double* scale, cudaStream_t stream);
cudaError_t launch_cmvn_apply(const float* feats, int64_t n_rows, const double* mean, const double* scale, float* out,
                              cudaStream_t stream);

// Rate conversion sr_in -> 16 kHz (librosa.load's soxr_hq step, pipeline1.py:102).  in_f32 / in_q16: one of the two is null.
// Clip c: n_in = in_lengths[c] samples at in_starts[c] -> ceil(n_in * 16000 / sr_in) float32 samples at out_starts[c].
cudaError_t launch_resample(const float* in_f32, const int16_t* in_q16, int sr_in, const int64_t* in_starts,
                            const int32_t* in_lengths, int n_clips, int max_in_len, float* out, const int64_t* out_starts,
                            cudaStream_t stream);
// Host copy of the polyphase table for the parity tests: returns up * ntaps (0 when sr_in is unsupported); meta = {up, down, half, ntaps}.
int64_t resample_table_host(int sr_in, double* h_out, int64_t max_elems, int32_t* meta);

}  // namespace dys
