#include "dys_profile.h"

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

namespace dys {

namespace {

const char* const kNames[kKernelCount] = {
    "k_feat_init", "k_frame_spectra", "k_tuning", "k_frame_cepstra", "k_clip_stats", "k_clean_init",
    "k_nr_stft_mag", "k_nr_iir_mask", "k_nr_apply_ola", "k_quantize_pcm",
    "k_cmvn_partial", "k_cmvn_merge", "k_cmvn_finalize", "k_cmvn_apply",
    "k_qc_snr", "k_qc_hf_bins", "k_qc_finish", "k_qc_flatness", "k_qc_flat_reduce", "k_resample"};

struct Record {
    int id, dev;
    cudaEvent_t start, stop;
};

std::atomic<long long> g_launches[kKernelCount];
std::atomic<bool> g_enabled{false};
std::mutex g_mu;
std::vector<Record> g_records;          // events in flight (not yet read)
std::map<int, std::vector<cudaEvent_t>> g_free;   // recycled events, per device (an event belongs to the device it was created on)
double g_ms[kKernelCount] = {};

cudaEvent_t take_event(int dev) {
    std::vector<cudaEvent_t>& pool = g_free[dev];
    if (!pool.empty()) {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

}  // namespace

const char* kernel_name(int id) { return (id >= 0 && id < kKernelCount) ? kNames[id] : ""; }

LaunchScope::LaunchScope(int id, cudaStream_t stream) : stop_(nullptr), stream_(stream) {
    g_launches[id].fetch_add(1, std::memory_order_relaxed);
    if (!g_enabled.load(std::memory_order_relaxed)) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(g_mu);
    Record r{id, dev, take_event(dev), take_event(dev)};
    if (!r.start || !r.stop) return;
    cudaEventRecord(r.start, stream);
    stop_ = r.stop;                      // the handle, not an index: profile_read may clear g_records from another thread
    g_records.push_back(r);
}

LaunchScope::~LaunchScope() {
    if (stop_) cudaEventRecord(stop_, stream_);
}

void profile_enable(bool on) { g_enabled.store(on); }

cudaError_t profile_read(double* ms, long long* launches, int reset) {
    std::lock_guard<std::mutex> lock(g_mu);
    cudaError_t err = cudaSuccess;
    for (const Record& r : g_records) {
        cudaError_t e = cudaEventSynchronize(r.stop);
        float t = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.start, r.stop);
        if (e == cudaSuccess) g_ms[r.id] += double(t);
        else err = e;
        g_free[r.dev].push_back(r.start);
        g_free[r.dev].push_back(r.stop);
    }
    g_records.clear();
    for (int i = 0; i < kKernelCount; ++i) {
        if (ms) ms[i] = g_ms[i];
        if (launches) launches[i] = g_launches[i].load();
        if (reset) { g_ms[i] = 0.0; g_launches[i].store(0); }
    }
    return err;
}

}  // namespace dys
