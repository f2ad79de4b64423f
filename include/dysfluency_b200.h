/* dysfluency_b200.h -- C ABI of the B200-native audio front-end (libdysb200.so).
 *
 * Drop-in boundary for ONE path of kishormb/Recognizing-Speech-Dysfluencies-in-Stuttering:
 * 16 kHz clip -> `<stem>_raw_feats.npy` / `<stem>_clean_feats.npy` (float32[149]).
 * The reference has no FFI (three flat Python scripts); each entry point below names the
 * Python function / call site of the reference whose arithmetic it replaces.  The Python
 * host layer (recognizing-speech-dysfluencies-in-stuttering_b200/frontend.py) binds these
 * with ctypes and re-exports the reference's own function names; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; every `d_*` pointer is DEVICE memory owned by the caller;
 *     `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is
 *     stream-ordered and asynchronous, nothing is copied to or from the host.
 *   - return value: DYS_OK or a DYS_ERR_* code; dys_last_error() gives the text (thread-local).
 *   - sample rate is frozen at 16 000 Hz like the reference's TARGET_SR (pipeline1.py:78).
 *   - the library never falls back to a CPU path: without a CUDA device every call fails.
 */
#ifndef DYSFLUENCY_B200_H
#define DYSFLUENCY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DYS_API __attribute__((visibility("default")))
#else
#define DYS_API
#endif

#define DYS_OK 0
#define DYS_ERR_INVALID 1   /* bad argument                               */
#define DYS_ERR_CUDA 2      /* CUDA runtime error (see dys_last_error)    */
#define DYS_ERR_WORKSPACE 3 /* workspace too small                        */

#define DYS_FEATURE_LEN 149       /* pipeline1.py:86  TOTAL_FEATURE_LEN  */
#define DYS_AUDIO_FEATURE_LEN 144 /* pipeline1.py:84  AUDIO_FEATURE_LEN  */
#define DYS_SAMPLE_RATE 16000     /* pipeline1.py:78  TARGET_SR          */
#define DYS_CMVN_ACC_LEN 299      /* 1 + 2 * 149                          */

/* per-clip status bits (the reference logs and returns zeros / None in these cases) */
#define DYS_STATUS_SHORT 1          /* < 9 STFT frames (n < 4096): librosa.feature.delta raises -> zeros(144)  pipeline1.py:237-239 */
#define DYS_STATUS_NONFINITE 2      /* NaN/Inf sample: librosa.util.valid_audio raises -> zeros(144)                              */
#define DYS_STATUS_CLEAN_FALLBACK 4 /* cleaning failed (e.g. all-zero clip -> NaN): raw clip used as clean  pipeline1.py:385-387 */
#define DYS_STATUS_BAD_LENGTH 8     /* length < 0 or > max_len                                                                     */

/* Library version (major * 100 + minor). */
DYS_API int dys_version(void);

/* Text of the last error raised on the calling thread ("" if none). */
DYS_API const char* dys_last_error(void);

/* Builds the lookup tables (Hann windows, twiddles, slaney mel filterbank, DCT-II, the 100
 * chroma filterbanks, smoothing taps) on the CURRENT CUDA device.  Idempotent; the other calls
 * invoke it implicitly.  Replaces the filterbank construction librosa repeats on every clip
 * (librosa.filters.mel / filters.chroma under pipeline1.py:216,227). */
DYS_API int dys_init(void);

/* Bytes of device scratch that let a batch of n_clips clips of at most max_len samples run with
 * the default sub-batching (with_clean != 0: denoise + both branches).  Smaller workspaces are
 * accepted down to dys_workspace_min_bytes(); the batch is then processed in more sub-batches. */
DYS_API int64_t dys_workspace_bytes(int32_t n_clips, int32_t max_len, int32_t with_clean);
DYS_API int64_t dys_workspace_min_bytes(int32_t n_clips, int32_t max_len, int32_t with_clean);

/* on != 0: dys_features_raw_clean* run the spectral gate and the clean branch on a library-owned high-priority side
 * stream (forked from and joined to the caller's stream with events) while the raw clip's feature pass stays on the
 * caller's stream; dys_workspace_bytes() then includes the raw branch's own scratch region.  Measured slower than the
 * single-stream order on B200, so the default is off. */
DYS_API int dys_set_overlap(int32_t on);

/* extract_features(y, sr, "") for a batch                    [pipeline1.py:257-265, 206-239]
 *   d_audio   float32 samples; clip c = d_audio[d_starts[c] .. d_starts[c] + d_lengths[c])
 *             (clips may overlap -> sliding windows cost no copy; even d_starts enable 8-byte loads)
 *   d_out     float32 [n_clips][149]  ([0:144] audio statistics, [144:149] zeros: transcript is "")
 *   d_status  int32   [n_clips]       DYS_STATUS_* bits; rows with SHORT/NONFINITE/BAD_LENGTH are zeros
 */
DYS_API int dys_features_raw(const float* d_audio, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                     int32_t max_len, float* d_out, int32_t* d_status, void* d_workspace, int64_t workspace_bytes,
                     void* stream);

/* Both cache entries of a clip in one call:
 *   raw   = extract_features(y)                                                [pipeline1.py:449]
 *   clean = extract_features(load(write_pcm16(normalize(nr.reduce_noise(y)))))  [pipeline1.py:140-142, 450]
 *   prop_decrease   1.0 = pipeline1.py:140 (the committed artefacts), 0.8 = main1.py:605
 *   d_out_raw / d_out_clean  float32 [n_clips][149]
 *   d_status        int32 [2 * n_clips]: [0,n) raw branch, [n,2n) clean branch
 *   d_clean_pcm     optional (NULL to skip) int16 PCM exactly as the reference writes it to
 *                   clear_audio/<stem>.wav; clip c at d_clean_pcm[d_pcm_starts[c] ..]
 */
DYS_API int dys_features_raw_clean(const float* d_audio, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                           int32_t max_len, float prop_decrease, float* d_out_raw, float* d_out_clean,
                           int32_t* d_status, int16_t* d_clean_pcm, const int64_t* d_pcm_starts, void* d_workspace,
                           int64_t workspace_bytes, void* stream);

/* The same two entry points for 16-bit PCM input: sample value = d_pcm[i] / 32768, exactly what
 * librosa.load returns for a PCM-16 WAV (the reference's clean branch and its clear_audio/ corpus,
 * pipeline1.py:389, 437).  Results are bit-identical to the float32 entry points fed int16 / 32768.0f;
 * half the bytes cross PCIe and HBM.  Even d_starts keep the 4-byte vector loads. */
DYS_API int dys_features_raw_pcm16(const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                                   int32_t max_len, float* d_out, int32_t* d_status, void* d_workspace, int64_t workspace_bytes,
                                   void* stream);
DYS_API int dys_features_raw_clean_pcm16(const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                                         int32_t max_len, float prop_decrease, float* d_out_raw, float* d_out_clean,
                                         int32_t* d_status, int16_t* d_clean_pcm, const int64_t* d_pcm_starts, void* d_workspace,
                                         int64_t workspace_bytes, void* stream);

/* Rate conversion to 16 kHz: the resampling half of librosa.load(path, sr=16000)            [pipeline1.py:102]
 * (soxr "HQ": linear phase, pass-band to 0.9136 x the lower Nyquist, stop-band from it, 126 dB; restated from soxr's
 * published recipe -- parity with the reference's *_raw_feats.npy is statistical).
 *   d_in        float32 samples, or int16 PCM when in_is_pcm16 != 0 (value = q / 32768)
 *   clip c      d_in[d_in_starts[c] .. + d_in_lengths[c])  ->  d_out[d_out_starts[c] .. + dys_resampled_length(len, sr_in))
 *   sr_in       any rate whose ratio to 16000 reduces to at most 4096 / 8192 (22050, 44100, 48000, 32000, 24000, 11025, 8000, ...)
 */
DYS_API int64_t dys_resampled_length(int64_t n_in, int32_t sr_in);
DYS_API int dys_resample_to_16k(const void* d_in, int32_t in_is_pcm16, int32_t sr_in, const int64_t* d_in_starts,
                                const int32_t* d_in_lengths, int32_t n_clips, int32_t max_in_len, float* d_out,
                                const int64_t* d_out_starts, void* stream);
/* Host copy of the polyphase table h[up][ntaps] (float64) for the parity tests; h_meta = {up, down, half, ntaps}.
 * Returns up * ntaps, 0 for an unsupported rate, -1 when max_elems is too small.  h_out may be NULL. */
DYS_API int64_t dys_resample_table(int32_t sr_in, double* h_out, int64_t max_elems, int32_t* h_meta);

/* StandardScaler().fit building block                            [pipeline1.py:470-471]
 *   d_acc (float64[299]) <- [n_rows, sum_f (x - shift), sum_f (x - shift)^2]; d_shift NULL = 0.
 *   Per-GPU accumulators add across ranks (one NCCL all-reduce of 299 doubles).
 *   d_partials: float64 scratch [128 * 298]. */
DYS_API int dys_cmvn_accumulate(const float* d_feats, int64_t n_rows, const double* d_shift, double* d_acc, double* d_partials,
                        void* stream);
/* mean_/scale_ from (all-reduced) moments; constant features get scale 1.0 like sklearn. */
DYS_API int dys_cmvn_finalize(const double* d_acc, const double* d_shift, double* d_mean, double* d_scale, void* stream);
/* StandardScaler.transform on the float32 feature matrix      [pipeline1.py:472-473, main1.py:987]
 * (scikit-learn's arithmetic: mean_ and scale_ are cast to float32, then x - mean, then / scale, each rounded to float32) */
DYS_API int dys_cmvn_apply(const float* d_feats, int64_t n_rows, const double* d_mean, const double* d_scale, float* d_out,
                   void* stream);

/* Per-file QC scalars the reference logs to output_results/per_file_analysis.csv       [pipeline1.py:151-186, 379-396]
 *   d_out float32 [n_clips][3] = { snr_db(y), spectral_flatness_mean(y), high_freq_energy_ratio(y, 16000) }
 *   (clips shorter than 400 samples: snr 0.0; empty / non-finite clips: flatness 0.0 -- the reference's early return
 *   and except branches).  Reporting only: the high-frequency ratio evaluates a full-length DFT band per clip. */
DYS_API int64_t dys_qc_workspace_bytes(int32_t n_clips, int32_t max_len);
DYS_API int dys_qc_metrics(const float* d_audio, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                           int32_t max_len, float* d_out, void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- introspection used by the parity tests (stage-wise comparison, SURVEY.md section 4 iii) ---- */
/* Copies a host-side table: which = 0 mel filterbank float32[128*1025], 1 DCT float32[20*128],
 * 2 chroma filterbank of tuning index `arg` float32[1025*12] ([bin][chroma]), 3 hann2048 float32[2048],
 * 4 tuning edges float64[101], 5 smoothing taps float64[33+7], 6 istft window-sum-square float64[256],
 * 7 iir b float64[1], 8 lane-per-filter mel runs int32[128 starts + 128 lengths + 4 group offsets] (padded in front so
 * that the 32 starts of a group differ mod 32), 9 their step-major weights float32[32*88] (filter f = lane + 32 g, step j
 * at offset[g] + 32 j + lane).  Returns the element count, or -1. */
DYS_API int64_t dys_get_table(int32_t which, int32_t arg, void* h_out, int64_t max_elems);

/* Runs the feature path for ONE clip already on the device and exposes the intermediates:
 * d_power float32 [T][1032], d_logmel float32 [T][128] (before the top-dB clamp), d_mfcc float32 [T][20],
 * d_chroma float32 [T][12], d_scalars int32[4] = {T, tuning index, peak count, status}, d_out float32[149].
 * Any output pointer may be NULL. */
DYS_API int dys_debug_feature_stages(const float* d_audio, int32_t n, float* d_power, float* d_logmel, float* d_mfcc,
                             float* d_chroma, int32_t* d_scalars, float* d_out, void* stream);
/* Denoise ONE clip: d_clean float32[n] (reduce_noise output, before normalisation), d_info float32[2] = {peak, flag}. */
DYS_API int dys_debug_denoise(const float* d_audio, int32_t n, float prop_decrease, float* d_clean, float* d_info, void* stream);

/* ---- launch accounting (bench.py: "gpu_launches" and the live roofline timing) ---- */
/* Number of distinct kernels in the library; dys_kernel_name(i) names kernel i ("" out of range). */
DYS_API int dys_kernel_count(void);
DYS_API const char* dys_kernel_name(int32_t index);
/* on != 0: bracket every kernel launch with two cudaEvents on its stream (adds no synchronisation). */
DYS_API int dys_profile_enable(int32_t on);
/* Waits for the recorded events; h_ms[i] = total milliseconds spent in kernel i since the last reset
 * (only while profiling was enabled), h_launches[i] = launches of kernel i since the last reset (always
 * counted).  Both are HOST arrays of dys_kernel_count() elements; either may be NULL. */
DYS_API int dys_profile_read(double* h_ms, int64_t* h_launches, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* DYSFLUENCY_B200_H */
