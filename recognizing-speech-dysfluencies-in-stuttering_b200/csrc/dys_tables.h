// Read-only lookup tables of the dysfluency front-end, built once per device in float64 on
// the host and uploaded (the reference rebuilds its filterbanks on every call:
// librosa.filters.mel / librosa.filters.chroma inside pipeline1.py:216,227).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace dys {

constexpr int kSR = 16000;
// feature STFT (librosa defaults inherited by pipeline1.py:216,227)
constexpr int kNfft = 2048;
constexpr int kHop = 512;
constexpr int kBins = 1025;
constexpr int kBinsPad = 1032;          // row pitch of the power-spectrogram scratch (floats)
constexpr int kMels = 128;
constexpr int kMfcc = 20;               // pipeline1.py:79 MFCC_N
constexpr int kChroma = 12;
constexpr int kFeat = 149;              // pipeline1.py:86 TOTAL_FEATURE_LEN
constexpr int kAudioFeat = 144;         // pipeline1.py:84 AUDIO_FEATURE_LEN
constexpr int kMelNnz = 2020;
constexpr int kMelWtMax = 32 * 88;       // step-major mel weights: 32 x (6 + 10 + 22 + 48) used
constexpr int kTunings = 100;
constexpr int kPipLo = 20;              // 150 Hz <= k * 7.8125 < 4000 Hz  ->  k in [20, 511]
constexpr int kPipHi = 511;
constexpr int kMaxPeaksPerFrame = 246;  // strict-left / weak-right local maxima cannot be adjacent
// spectral gate (noisereduce defaults inherited by pipeline1.py:140)
constexpr int kChromaPitch = 1032;      // bins per plane of the device chroma table (float4 units)
constexpr int kNrFft = 1024;
constexpr int kNrHop = 256;
constexpr int kNrBins = 513;
constexpr int kNrBinsPad = 520;         // row pitch of the fp64 spectral-gate scratch (doubles)
constexpr int kNrPad = 30000;
constexpr int kNrChunk = 600000;
constexpr int kNrFreqTaps = 33;
constexpr int kNrTimeTaps = 7;

struct HostTables {
    std::vector<float> hann2048;            // [2048]
    std::vector<float2> tw1024;             // [32*32]  W_1024^(l*kA) at kA*32+l  (x = cos, y = -sin)
    std::vector<float2> split2048;          // [1024]   (cos, sin)(2 pi k / 2048)
    std::vector<int> mel_start, mel_len, mel_ptr;   // [128]  support of every filter
    std::vector<int> mel_rstart, mel_rlen;          // [128]  the same, padded in front so that a group's 32 starts differ mod 32
    std::vector<float> mel_w;               // [2020]
    // the same weights step-major for the lane-per-filter loop: filter f = lane + 32 g, its j-th bin at
    // mel_wt[mel_goff[g] + 32 j + lane] -> the 32 simultaneous weight reads are conflict-free
    std::vector<float> mel_wt;              // [32 * sum_g (longest filter of group g)]
    int mel_goff[4] = {0, 0, 0, 0};
    std::vector<float> mel_dense;           // [128*1025] (debug / tests only, host side)
    std::vector<float> dct;                 // [20*128]  ortho DCT-II rows
    std::vector<float> chroma;              // [100][1025][12]
    // device layout of the same weights: three planes of 4 chroma rows each, [100][3][kChromaPitch] float4, so the 32
    // lanes of a warp (one bin each) read 512 contiguous bytes per plane (the bin-major rows are 48 B apart)
    std::vector<float> chroma_planes;
    std::vector<double> tuning_edges;       // [101]
    std::vector<double> hann1024;           // [1024]
    std::vector<double2> tw512;             // [16*32]
    std::vector<double2> split1024;         // [512]
    std::vector<double> wss;                // [256]  istft window-sum-square, interior
    std::vector<double> smooth_f, smooth_t; // [33], [7]
    double iir_b = 0.0;
};

struct DeviceTables {
    const float* hann2048;
    const float2* tw1024;
    const float2* split2048;
    const int* mel_start;                   // HostTables::mel_rstart / mel_rlen (padded runs)
    const int* mel_len;
    const int* mel_ptr;
    const float* mel_w;
    const float* mel_wt;
    int mel_goff[4];
    int mel_wt_len;
    const float* dct;
    const float4* chroma;                   // plane-major (HostTables::chroma_planes)
    const double* tuning_edges;
    const double* hann1024;
    const double2* tw512;
    const double2* split1024;
    const double* wss;
    const double* smooth_f;
    const double* smooth_t;
    double iir_b;
};

const HostTables& host_tables();                 // built lazily, thread-safe
// Tables resident on the current CUDA device (uploaded on first use). nullptr + error on failure.
const DeviceTables* device_tables();

}  // namespace dys
