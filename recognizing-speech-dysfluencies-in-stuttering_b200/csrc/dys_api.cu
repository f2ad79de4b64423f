// C ABI of libdysb200.so (declared in include/dysfluency_b200.h): argument checks, workspace
// carving, sub-batch loops.  No host copies, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "../../include/dysfluency_b200.h"
#include "dys_error.h"
#include "dys_kernels.h"
#include "dys_profile.h"

namespace dys {

namespace {
thread_local std::string g_err;
}
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }

namespace {

size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    const long x = std::strtol(v, nullptr, 10);
    return x > 0 ? int(x) : dflt;
}
// Sub-batch caps: how many clip instances / denoise chunks share one scratch arena per launch group.
// Large groups keep the last wave of CTAs full (measured on B200, 10k 3-s clips: 45.9 ms per pass at
// 1024/1024, 42.6 ms at 4096/4096, 41.9 ms unbounded in round 1; round 2: 29.83 ms with 8 GiB of scratch = 3 launch
// groups per kernel, 29.59 ms with 16 GiB, 29.50 ms with 32 GiB = one group); the arena is additionally bounded in
// bytes.  The default budget is sized for the 180 GB of a B200: a 10 000-clip batch asks for 27 GB of workspace.
int feat_cap() { return env_int("DYS_FEAT_SUBBATCH", 32768); }
int nr_cap() { return env_int("DYS_NR_SUBBATCH", 16384); }
size_t scratch_budget() { return size_t(env_int("DYS_SCRATCH_MB", 32768)) << 20; }
int sub_count(size_t per_item, int cap, int64_t total) {
    const size_t by_bytes = std::max<size_t>(1, scratch_budget() / std::max<size_t>(per_item, 1));
    return int(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(cap, total), int64_t(by_bytes))));
}

struct Layout {
    int64_t clean_pitch = 0;
    size_t clean_off = 0, cleanq_off = 0, peak_off = 0, flag_off = 0, scratch_off = 0;
    int t_max = 0, ta_max = 0, cpc = 1;
};

Layout make_layout(int n_clips, int max_len, bool with_clean) {
    Layout L;
    L.t_max = frames_of(std::max(max_len, 0));
    L.ta_max = nr_ta_max(std::max(max_len, 1));
    L.cpc = nr_chunks_of(std::max(max_len, 1));
    size_t off = 0;
    if (with_clean) {
        L.clean_pitch = (int64_t(std::max(max_len, 1)) + 63) & ~int64_t(63);
        L.clean_off = off; off += al256(size_t(n_clips) * L.clean_pitch * 4);
        L.cleanq_off = off; off += al256(size_t(n_clips) * L.clean_pitch * 2);
        L.peak_off = off; off += al256(size_t(n_clips) * 4);
        L.flag_off = off; off += al256(size_t(n_clips) * 4);
    }
    L.scratch_off = off;
    return L;
}

int check_common(const void* d_audio, const void* d_starts, const void* d_lengths, int n_clips, int max_len);
int features_raw_impl(const float* d_audio, const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                      int32_t max_len, float* d_out, int32_t* d_status, void* d_workspace, int64_t workspace_bytes, void* stream);
int features_raw_clean_impl(const float* d_audio, const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths,
                            int32_t n_clips, int32_t max_len, float prop_decrease, float* d_out_raw, float* d_out_clean,
                            int32_t* d_status, int16_t* d_clean_pcm, const int64_t* d_pcm_starts, void* d_workspace,
                            int64_t workspace_bytes, void* stream);

int check_common(const void* d_audio, const void* d_starts, const void* d_lengths, int n_clips, int max_len) {
    if (n_clips < 0 || max_len < 0) { set_error("n_clips and max_len must be >= 0"); return DYS_ERR_INVALID; }
    if (n_clips > 0 && (!d_audio || !d_starts || !d_lengths)) { set_error("null device pointer"); return DYS_ERR_INVALID; }
    if (int64_t(n_clips) * 2 > 0x7fff0000LL) { set_error("too many clips for one call"); return DYS_ERR_INVALID; }
    return DYS_OK;
}

// Feature pipeline over instances [inst_begin, inst_end) with `avail` bytes of scratch at `scratch`.
int run_features(const DeviceTables& tb, const ClipView& cv, int inst_begin, int inst_end, int t_max, unsigned char* scratch,
                 size_t avail, int cap, float* out_raw, float* out_clean, int32_t* status, cudaStream_t stream) {
    const int n_inst_total = inst_end - inst_begin;
    if (n_inst_total <= 0) return DYS_OK;
    const size_t per = feat_scratch_bytes(1, t_max);
    // the wanted group first (it fits exactly when the caller allocated dys_workspace_bytes()), the per-item estimate only
    // for smaller workspaces: the estimate rounds every array up separately and would cut a 10 000-clip group to 9 998 + 2
    int n_sub = sub_count(per, cap, n_inst_total);
    if (feat_scratch_bytes(n_sub, t_max) > avail) n_sub = int(std::min<size_t>(avail / per, size_t(n_sub)));
    if (n_sub < 1) { set_error("workspace too small for one clip"); return DYS_ERR_WORKSPACE; }
    while (n_sub > 1 && feat_scratch_bytes(n_sub, t_max) > avail) --n_sub;
    FeatScratch sc;
    feat_scratch_carve(scratch, n_sub, t_max, &sc);
    for (int i0 = inst_begin; i0 < inst_end; i0 += n_sub) {
        const int cnt = std::min(n_sub, inst_end - i0);
        DYS_CUDA_OK(launch_features(tb, cv, i0, cnt, sc, out_raw, out_clean, status, stream));
    }
    return DYS_OK;
}

// ---- two branches on two streams ---------------------------------------------------------------------------------
// The raw clip's feature pass does not depend on the spectral gate.  The gate and the clean branch run on a
// library-owned HIGH-priority side stream (one per caller stream and device, forked and joined with events), the raw
// branch stays on the caller's stream: its CTAs are scheduled only where the gate leaves an SM idle -- the
// latency-bound time-smoothing sweep (32-thread CTAs, little shared memory) and the tail of every wave.  The raw branch
// needs its own scratch region, so this is used when the caller's workspace has dys_workspace_bytes(); smaller
// workspaces run everything on the caller's stream.
// MEASURED (B200, 10 000 3-s clips, round 2): 31.9 ms per step with the fork against 31.2 ms without (raw branch forked
// at entry on an equal-priority stream: 31.6 against 31.0); the host streaming path loses 1 - 3 ms.  The co-running
// kernels take bandwidth and shared memory from the gate's FFT kernels, which costs more than the idle time they fill.
// It is therefore OFF by default; dys_set_overlap(1) enables it.
std::atomic<int> g_overlap{0};
int side_cap() { return env_int("DYS_SIDE_SUBBATCH", 4096); }
size_t side_scratch_bytes(int n_clips, int t_max) {
    return feat_scratch_bytes(sub_count(feat_scratch_bytes(1, t_max), side_cap(), std::max(n_clips, 1)), t_max);
}
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
std::mutex g_side_mu;
std::map<std::pair<int, cudaStream_t>, SideStream> g_side;
bool side_stream_for(cudaStream_t caller, SideStream* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(g_side_mu);
    auto key = std::make_pair(dev, caller);
    auto it = g_side.find(key);
    if (it == g_side.end()) {
        SideStream s;
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, greatest) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return false;
        it = g_side.emplace(key, s).first;
    }
    *out = it->second;
    return true;
}

}  // namespace
}  // namespace dys

using namespace dys;

extern "C" {

DYS_API int dys_version(void) { return 200; }

DYS_API const char* dys_last_error(void) { return last_error_cstr(); }

DYS_API int dys_init(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        set_error("no CUDA device: this library has no CPU path");
        return DYS_ERR_CUDA;
    }
    return device_tables() ? DYS_OK : DYS_ERR_CUDA;
}

DYS_API int64_t dys_workspace_bytes(int32_t n_clips, int32_t max_len, int32_t with_clean) {
    if (n_clips < 0 || max_len < 0) return -1;
    const Layout L = make_layout(n_clips, max_len, with_clean != 0);
    const int n_inst = sub_count(feat_scratch_bytes(1, L.t_max), feat_cap(), int64_t(n_clips) * (with_clean ? 2 : 1));
    size_t scratch = feat_scratch_bytes(n_inst, L.t_max);
    if (with_clean) {
        const int n_items = sub_count(nr_scratch_bytes(1, L.ta_max), nr_cap(), int64_t(n_clips) * L.cpc);
        scratch = std::max(scratch, nr_scratch_bytes(n_items, L.ta_max));
        if (g_overlap.load()) scratch += side_scratch_bytes(n_clips, L.t_max);   // the raw branch's own region (two-stream mode)
    }
    return int64_t(L.scratch_off + scratch);
}

DYS_API int dys_set_overlap(int32_t on) {
    g_overlap.store(on != 0 ? 1 : 0);
    return DYS_OK;
}

DYS_API int64_t dys_workspace_min_bytes(int32_t n_clips, int32_t max_len, int32_t with_clean) {
    if (n_clips < 0 || max_len < 0) return -1;
    const Layout L = make_layout(n_clips, max_len, with_clean != 0);
    size_t scratch = feat_scratch_bytes(1, L.t_max);
    if (with_clean) scratch = std::max(scratch, nr_scratch_bytes(1, L.ta_max));
    return int64_t(L.scratch_off + scratch);
}

DYS_API int dys_features_raw(const float* d_audio, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                     int32_t max_len, float* d_out, int32_t* d_status, void* d_workspace, int64_t workspace_bytes,
                     void* stream) {
    return features_raw_impl(d_audio, nullptr, d_starts, d_lengths, n_clips, max_len, d_out, d_status, d_workspace,
                             workspace_bytes, stream);
}

DYS_API int dys_features_raw_pcm16(const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                                   int32_t max_len, float* d_out, int32_t* d_status, void* d_workspace, int64_t workspace_bytes,
                                   void* stream) {
    return features_raw_impl(nullptr, d_pcm, d_starts, d_lengths, n_clips, max_len, d_out, d_status, d_workspace,
                             workspace_bytes, stream);
}

DYS_API int dys_features_raw_clean_pcm16(const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                                         int32_t max_len, float prop_decrease, float* d_out_raw, float* d_out_clean,
                                         int32_t* d_status, int16_t* d_clean_pcm, const int64_t* d_pcm_starts, void* d_workspace,
                                         int64_t workspace_bytes, void* stream) {
    return features_raw_clean_impl(nullptr, d_pcm, d_starts, d_lengths, n_clips, max_len, prop_decrease, d_out_raw, d_out_clean,
                                   d_status, d_clean_pcm, d_pcm_starts, d_workspace, workspace_bytes, stream);
}

DYS_API int64_t dys_resampled_length(int64_t n_in, int32_t sr_in) {
    if (n_in < 0 || sr_in <= 0) return -1;
    if (sr_in == kSR) return n_in;
    return (n_in * kSR + sr_in - 1) / sr_in;                 // librosa.resample: ceil(n * target_sr / orig_sr)
}

DYS_API int dys_resample_to_16k(const void* d_in, int32_t in_is_pcm16, int32_t sr_in, const int64_t* d_in_starts,
                                const int32_t* d_in_lengths, int32_t n_clips, int32_t max_in_len, float* d_out,
                                const int64_t* d_out_starts, void* stream) {
    if (n_clips < 0 || max_in_len < 0) { set_error("n_clips and max_in_len must be >= 0"); return DYS_ERR_INVALID; }
    if (n_clips == 0) return DYS_OK;
    if (!d_in || !d_in_starts || !d_in_lengths || !d_out || !d_out_starts) { set_error("null device pointer"); return DYS_ERR_INVALID; }
    if (sr_in == kSR) { set_error("sr_in is already 16000: nothing to resample"); return DYS_ERR_INVALID; }
    if (resample_table_host(sr_in, nullptr, 0, nullptr) <= 0) { set_error("unsupported input sample rate"); return DYS_ERR_INVALID; }
    if (!device_tables()) return DYS_ERR_CUDA;
    DYS_CUDA_OK(launch_resample(in_is_pcm16 ? nullptr : static_cast<const float*>(d_in),
                                in_is_pcm16 ? static_cast<const int16_t*>(d_in) : nullptr, sr_in, d_in_starts, d_in_lengths,
                                n_clips, max_in_len, d_out, d_out_starts, static_cast<cudaStream_t>(stream)));
    return DYS_OK;
}

DYS_API int64_t dys_resample_table(int32_t sr_in, double* h_out, int64_t max_elems, int32_t* h_meta) {
    return resample_table_host(sr_in, h_out, max_elems, h_meta);
}

}  // extern "C"

namespace dys { namespace {

int features_raw_impl(const float* d_audio, const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                      int32_t max_len, float* d_out, int32_t* d_status, void* d_workspace, int64_t workspace_bytes, void* stream) {
    if (int rc = check_common(d_audio ? static_cast<const void*>(d_audio) : d_pcm, d_starts, d_lengths, n_clips, max_len)) return rc;
    if (n_clips == 0) return DYS_OK;
    if (!d_out || !d_status || !d_workspace) { set_error("null output / workspace pointer"); return DYS_ERR_INVALID; }
    const DeviceTables* tb = device_tables();
    if (!tb) return DYS_ERR_CUDA;
    const Layout L = make_layout(n_clips, max_len, false);
    if (workspace_bytes < dys_workspace_min_bytes(n_clips, max_len, 0)) {
        set_error("workspace smaller than dys_workspace_min_bytes()");
        return DYS_ERR_WORKSPACE;
    }
    ClipView cv{};
    cv.audio = d_audio; cv.audio_q = d_pcm; cv.starts = d_starts; cv.lengths = d_lengths; cv.n_clips = n_clips; cv.max_len = max_len;
    return run_features(*tb, cv, 0, n_clips, L.t_max, static_cast<unsigned char*>(d_workspace) + L.scratch_off,
                        size_t(workspace_bytes) - L.scratch_off, feat_cap(), d_out, nullptr, d_status,
                        static_cast<cudaStream_t>(stream));
}

int features_raw_clean_impl(const float* d_audio, const int16_t* d_pcm, const int64_t* d_starts, const int32_t* d_lengths,
                            int32_t n_clips, int32_t max_len, float prop_decrease, float* d_out_raw, float* d_out_clean,
                            int32_t* d_status, int16_t* d_clean_pcm, const int64_t* d_pcm_starts, void* d_workspace,
                            int64_t workspace_bytes, void* stream) {
    if (int rc = check_common(d_audio ? static_cast<const void*>(d_audio) : d_pcm, d_starts, d_lengths, n_clips, max_len)) return rc;
    if (n_clips == 0) return DYS_OK;
    if (!d_out_raw || !d_out_clean || !d_status || !d_workspace) {
        set_error("null output / workspace pointer");
        return DYS_ERR_INVALID;
    }
    if (d_clean_pcm && !d_pcm_starts) { set_error("d_clean_pcm given without d_pcm_starts"); return DYS_ERR_INVALID; }
    if (!(prop_decrease >= 0.f && prop_decrease <= 1.f)) { set_error("prop_decrease must be in [0, 1]"); return DYS_ERR_INVALID; }
    const DeviceTables* tb = device_tables();
    if (!tb) return DYS_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Layout L = make_layout(n_clips, max_len, true);
    if (workspace_bytes < dys_workspace_min_bytes(n_clips, max_len, 1)) {
        set_error("workspace smaller than dys_workspace_min_bytes()");
        return DYS_ERR_WORKSPACE;
    }
    unsigned char* ws = static_cast<unsigned char*>(d_workspace);
    float* clean = reinterpret_cast<float*>(ws + L.clean_off);
    int16_t* clean_q = reinterpret_cast<int16_t*>(ws + L.cleanq_off);
    float* peak = reinterpret_cast<float*>(ws + L.peak_off);
    int32_t* flag = reinterpret_cast<int32_t*>(ws + L.flag_off);
    ClipView cv{};
    cv.audio = d_audio; cv.audio_q = d_pcm; cv.starts = d_starts; cv.lengths = d_lengths; cv.n_clips = n_clips; cv.max_len = max_len;
    cv.clean = clean; cv.clean_pitch = L.clean_pitch; cv.clean_peak = peak; cv.clean_flag = flag; cv.clean_q = clean_q;

    // ---- fork: the gate and the clean branch go to the library's high-priority side stream, the raw branch stays on the
    // caller's stream (streams created by the caller have the lowest priority: its CTAs only fill what the gate leaves idle)
    const size_t side_bytes = side_scratch_bytes(n_clips, L.t_max);
    const bool overlap = g_overlap.load() != 0 && workspace_bytes >= dys_workspace_bytes(n_clips, max_len, 1);
    const size_t main_bytes = size_t(workspace_bytes) - (overlap ? side_bytes : 0);     // [0, main_bytes): everything but the raw branch's region
    SideStream side;
    struct Joiner {
        cudaStream_t stream = nullptr, from = nullptr;
        cudaEvent_t event = nullptr;
        ~Joiner() { if (event && cudaEventRecord(event, from) == cudaSuccess) cudaStreamWaitEvent(stream, event, 0); }
    } joiner;                                                              // the caller's stream waits for the side stream on every return path
    const bool forked = overlap && side_stream_for(st, &side);
    cudaStream_t gate_st = st;
    if (forked) {
        DYS_CUDA_OK(cudaEventRecord(side.fork, st));                       // inputs (and earlier users of the workspace) are ready here
        DYS_CUDA_OK(cudaStreamWaitEvent(side.stream, side.fork, 0));
        gate_st = side.stream;
        joiner.stream = st; joiner.from = side.stream; joiner.event = side.join;
    }
    // ---- spectral gate over sub-batches of chunks ------------------------------------------
    DYS_CUDA_OK(launch_clean_init(cv, peak, flag, gate_st));
    {
        const size_t avail = main_bytes - L.scratch_off;
        const size_t per = nr_scratch_bytes(1, L.ta_max);
        const int64_t n_items = int64_t(n_clips) * L.cpc;
        int n_sub = sub_count(per, nr_cap(), n_items);
        if (nr_scratch_bytes(n_sub, L.ta_max) > avail) n_sub = int(std::min<size_t>(avail / per, size_t(n_sub)));
        if (n_sub < 1) { set_error("workspace too small for one denoise chunk"); return DYS_ERR_WORKSPACE; }
        while (n_sub > 1 && nr_scratch_bytes(n_sub, L.ta_max) > avail) --n_sub;
        NrScratch sc;
        nr_scratch_carve(ws + L.scratch_off, n_sub, L.ta_max, &sc);
        for (int64_t i0 = 0; i0 < n_items; i0 += n_sub) {
            const int cnt = int(std::min<int64_t>(n_sub, n_items - i0));
            DYS_CUDA_OK(launch_denoise(*tb, cv, clean, peak, flag, L.cpc, int(i0), cnt, sc, prop_decrease, gate_st));
            if (forked && i0 == 0) {                                       // the raw branch is queued behind the first gate launches
                if (int rc = run_features(*tb, cv, 0, n_clips, L.t_max, ws + main_bytes, side_bytes, side_cap(), d_out_raw,
                                          d_out_clean, d_status, st)) return rc;
            }
        }
    }
    DYS_CUDA_OK(launch_quantize_pcm(cv, clean_q, d_clean_pcm, d_pcm_starts, gate_st));
    // ---- features: the clean branch (and the raw one when it was not forked) -----------------------------------
    return run_features(*tb, cv, forked ? n_clips : 0, 2 * n_clips, L.t_max, ws + L.scratch_off, main_bytes - L.scratch_off,
                        feat_cap(), d_out_raw, d_out_clean, d_status, gate_st);
}

}}  // namespace dys::(anonymous)

extern "C" {

DYS_API int dys_features_raw_clean(const float* d_audio, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                           int32_t max_len, float prop_decrease, float* d_out_raw, float* d_out_clean,
                           int32_t* d_status, int16_t* d_clean_pcm, const int64_t* d_pcm_starts, void* d_workspace,
                           int64_t workspace_bytes, void* stream) {
    return features_raw_clean_impl(d_audio, nullptr, d_starts, d_lengths, n_clips, max_len, prop_decrease, d_out_raw, d_out_clean,
                                   d_status, d_clean_pcm, d_pcm_starts, d_workspace, workspace_bytes, stream);
}

DYS_API int dys_cmvn_accumulate(const float* d_feats, int64_t n_rows, const double* d_shift, double* d_acc, double* d_partials,
                        void* stream) {
    if (n_rows < 0 || !d_acc || !d_partials || (n_rows > 0 && !d_feats)) { set_error("bad cmvn arguments"); return DYS_ERR_INVALID; }
    DYS_CUDA_OK(launch_cmvn_accumulate(d_feats, n_rows, d_shift, d_acc, d_partials, static_cast<cudaStream_t>(stream)));
    return DYS_OK;
}

DYS_API int dys_cmvn_finalize(const double* d_acc, const double* d_shift, double* d_mean, double* d_scale, void* stream) {
    if (!d_acc || !d_mean || !d_scale) { set_error("bad cmvn arguments"); return DYS_ERR_INVALID; }
    DYS_CUDA_OK(launch_cmvn_finalize(d_acc, d_shift, d_mean, d_scale, static_cast<cudaStream_t>(stream)));
    return DYS_OK;
}

DYS_API int dys_cmvn_apply(const float* d_feats, int64_t n_rows, const double* d_mean, const double* d_scale, float* d_out,
                   void* stream) {
    if (n_rows < 0 || !d_mean || !d_scale || (n_rows > 0 && (!d_feats || !d_out))) { set_error("bad cmvn arguments"); return DYS_ERR_INVALID; }
    DYS_CUDA_OK(launch_cmvn_apply(d_feats, n_rows, d_mean, d_scale, d_out, static_cast<cudaStream_t>(stream)));
    return DYS_OK;
}

DYS_API int64_t dys_qc_workspace_bytes(int32_t n_clips, int32_t max_len) {
    if (n_clips < 0 || max_len < 0) return -1;
    const int n_sub = sub_count(feat_scratch_bytes(1, frames_of(max_len)), feat_cap(), std::max(n_clips, 1));
    return int64_t(qc_scratch_bytes(n_clips, max_len, n_sub));
}

DYS_API int dys_qc_metrics(const float* d_audio, const int64_t* d_starts, const int32_t* d_lengths, int32_t n_clips,
                           int32_t max_len, float* d_out, void* d_workspace, int64_t workspace_bytes, void* stream) {
    if (int rc = check_common(d_audio, d_starts, d_lengths, n_clips, max_len)) return rc;
    if (n_clips == 0) return DYS_OK;
    if (!d_out || !d_workspace) { set_error("null output / workspace pointer"); return DYS_ERR_INVALID; }
    const DeviceTables* tb = device_tables();
    if (!tb) return DYS_ERR_CUDA;
    if (workspace_bytes < dys_qc_workspace_bytes(n_clips, max_len)) {
        set_error("workspace smaller than dys_qc_workspace_bytes()");
        return DYS_ERR_WORKSPACE;
    }
    const int n_sub = sub_count(feat_scratch_bytes(1, frames_of(max_len)), feat_cap(), n_clips);
    ClipView cv{};
    cv.audio = d_audio; cv.starts = d_starts; cv.lengths = d_lengths; cv.n_clips = n_clips; cv.max_len = max_len;
    DYS_CUDA_OK(launch_qc(*tb, cv, d_out, d_workspace, size_t(workspace_bytes), n_sub, static_cast<cudaStream_t>(stream)));
    return DYS_OK;
}

DYS_API int64_t dys_get_table(int32_t which, int32_t arg, void* h_out, int64_t max_elems) {
    const HostTables& h = host_tables();
    const void* src = nullptr;
    int64_t n = 0;
    size_t esz = 4;
    double taps[kNrFreqTaps + kNrTimeTaps];
    int32_t runs[2 * kMels + 4];
    float wt[kMelWtMax];
    switch (which) {
        case 0: src = h.mel_dense.data(); n = int64_t(h.mel_dense.size()); break;
        case 1: src = h.dct.data(); n = int64_t(h.dct.size()); break;
        case 2:
            if (arg < 0 || arg >= kTunings) return -1;
            src = h.chroma.data() + size_t(arg) * kBins * kChroma; n = kBins * kChroma; break;
        case 3: src = h.hann2048.data(); n = kNfft; break;
        case 4: src = h.tuning_edges.data(); n = kTunings + 1; esz = 8; break;
        case 5:
            std::copy(h.smooth_f.begin(), h.smooth_f.end(), taps);
            std::copy(h.smooth_t.begin(), h.smooth_t.end(), taps + kNrFreqTaps);
            src = taps; n = kNrFreqTaps + kNrTimeTaps; esz = 8; break;
        case 6: src = h.wss.data(); n = kNrHop; esz = 8; break;
        case 7: src = &h.iir_b; n = 1; esz = 8; break;
        case 8:
            std::copy(h.mel_rstart.begin(), h.mel_rstart.end(), runs);
            std::copy(h.mel_rlen.begin(), h.mel_rlen.end(), runs + kMels);
            std::copy(h.mel_goff, h.mel_goff + 4, runs + 2 * kMels);
            src = runs; n = 2 * kMels + 4; break;
        case 9:
            std::fill(wt, wt + kMelWtMax, 0.f);
            std::copy(h.mel_wt.begin(), h.mel_wt.begin() + std::min<size_t>(h.mel_wt.size(), kMelWtMax), wt);
            src = wt; n = kMelWtMax; break;
        default: return -1;
    }
    if (!h_out || max_elems < n) return -1;
    std::memcpy(h_out, src, size_t(n) * esz);
    return n;
}

DYS_API int dys_debug_feature_stages(const float* d_audio, int32_t n, float* d_power, float* d_logmel, float* d_mfcc,
                             float* d_chroma, int32_t* d_scalars, float* d_out, void* stream) {
    if (!d_audio || n < 0) { set_error("bad debug arguments"); return DYS_ERR_INVALID; }
    const DeviceTables* tb = device_tables();
    if (!tb) return DYS_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int T = frames_of(n);
    unsigned char* ws = nullptr;
    const size_t bytes = feat_scratch_bytes(1, T) + 4096;
    DYS_CUDA_OK(cudaMalloc(&ws, bytes));
    int64_t h_start = 0;
    int32_t h_len = n;
    int64_t* d_start = reinterpret_cast<int64_t*>(ws);
    int32_t* d_len = reinterpret_cast<int32_t*>(ws + 64);
    int32_t* d_status = reinterpret_cast<int32_t*>(ws + 128);
    float* d_feat = reinterpret_cast<float*>(ws + 1024);
    cudaMemcpyAsync(d_start, &h_start, 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_len, &h_len, 4, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(ws + 4096, 0, bytes - 4096, st);
    ClipView cv{};
    cv.audio = d_audio; cv.starts = d_start; cv.lengths = d_len; cv.n_clips = 1; cv.max_len = n;
    FeatScratch sc;
    feat_scratch_carve(ws + 4096, 1, T, &sc);
    cudaError_t e = launch_features(*tb, cv, 0, 1, sc, d_feat, nullptr, d_status, st);
    if (e == cudaSuccess) {
        if (d_power) cudaMemcpyAsync(d_power, sc.power, size_t(T) * kBinsPad * 4, cudaMemcpyDeviceToDevice, st);
        if (d_logmel) cudaMemcpyAsync(d_logmel, sc.logmel, size_t(T) * kMels * 4, cudaMemcpyDeviceToDevice, st);
        if (d_mfcc) cudaMemcpyAsync(d_mfcc, sc.mfcc, size_t(T) * kMfcc * 4, cudaMemcpyDeviceToDevice, st);
        if (d_chroma) cudaMemcpyAsync(d_chroma, sc.chroma, size_t(T) * kChroma * 4, cudaMemcpyDeviceToDevice, st);
        if (d_out) cudaMemcpyAsync(d_out, d_feat, kFeat * 4, cudaMemcpyDeviceToDevice, st);
        if (d_scalars) {
            cudaMemcpyAsync(d_scalars + 1, sc.tuning_idx, 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(d_scalars + 2, sc.peak_count, 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(d_scalars + 3, d_status, 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(d_scalars, &T, 4, cudaMemcpyHostToDevice, st);
        }
        e = cudaStreamSynchronize(st);
    }
    cudaFree(ws);
    if (e != cudaSuccess) { set_error(std::string("debug_feature_stages: ") + cudaGetErrorString(e)); return DYS_ERR_CUDA; }
    return DYS_OK;
}

DYS_API int dys_debug_denoise(const float* d_audio, int32_t n, float prop_decrease, float* d_clean, float* d_info, void* stream) {
    if (!d_audio || n <= 0 || !d_clean) { set_error("bad debug arguments"); return DYS_ERR_INVALID; }
    const DeviceTables* tb = device_tables();
    if (!tb) return DYS_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cpc = nr_chunks_of(n), ta = nr_ta_max(n);
    unsigned char* ws = nullptr;
    const size_t bytes = nr_scratch_bytes(1, ta) + 4096;
    DYS_CUDA_OK(cudaMalloc(&ws, bytes));
    int64_t h_start = 0;
    int32_t h_len = n;
    int64_t* d_start = reinterpret_cast<int64_t*>(ws);
    int32_t* d_len = reinterpret_cast<int32_t*>(ws + 64);
    float* d_peak = reinterpret_cast<float*>(ws + 128);
    int32_t* d_flag = reinterpret_cast<int32_t*>(ws + 192);
    cudaMemcpyAsync(d_start, &h_start, 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_len, &h_len, 4, cudaMemcpyHostToDevice, st);
    ClipView cv{};
    cv.audio = d_audio; cv.starts = d_start; cv.lengths = d_len; cv.n_clips = 1; cv.max_len = n;
    cv.clean = d_clean; cv.clean_pitch = n; cv.clean_peak = d_peak; cv.clean_flag = d_flag;
    NrScratch sc;
    nr_scratch_carve(ws + 4096, 1, ta, &sc);
    cudaError_t e = launch_clean_init(cv, d_peak, d_flag, st);
    for (int it = 0; it < cpc && e == cudaSuccess; ++it)
        e = launch_denoise(*tb, cv, d_clean, d_peak, d_flag, cpc, it, 1, sc, prop_decrease, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && d_info) {
        float info[2];
        int32_t flag = 0;
        cudaMemcpy(&info[0], d_peak, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
        info[1] = float(flag);
        e = cudaMemcpy(d_info, info, 8, cudaMemcpyHostToDevice);
    }
    cudaFree(ws);
    if (e != cudaSuccess) { set_error(std::string("debug_denoise: ") + cudaGetErrorString(e)); return DYS_ERR_CUDA; }
    return DYS_OK;
}

DYS_API int dys_kernel_count(void) { return kKernelCount; }

DYS_API const char* dys_kernel_name(int32_t index) { return kernel_name(index); }

DYS_API int dys_profile_enable(int32_t on) {
    profile_enable(on != 0);
    return DYS_OK;
}

DYS_API int dys_profile_read(double* h_ms, int64_t* h_launches, int32_t reset) {
    long long counts[kKernelCount];
    DYS_CUDA_OK(profile_read(h_ms, counts, reset));
    if (h_launches)
        for (int i = 0; i < kKernelCount; ++i) h_launches[i] = counts[i];
    return DYS_OK;
}

}  // extern "C"
