// engine.h - the engine object behind the C ABI (include/b2piv.h) and the launch / dispatch entry points that the
// translation units of libb2piv.so share.  One .cu per kernel family (k_*.cu) so that they compile in parallel and stay
// reviewable; abi_piv.cu holds the PIV entry points and the host pipeline, abi_aux.cu the pre-processing / projection /
// mask / packing / predictor entry points.
#pragma once
#include "../../include/b2piv.h"
#include "piv_core.cuh"

#include <cuda.h>   // CUtensorMap (types only; the encoder is resolved at run time, libcuda is not linked)

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges show up when a tool (nsys / ncu --nvtx) is attached, otherwise no-ops
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// NVTX range around an ABI call (the reference has no tracing at all; its tqdm bars are the closest thing)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#include "work_partition.h"   // pick_run_len, partition_units: the host arithmetic behind the work-unit lists (no CUDA)
#include "stager.h"   // copy threads + cache-resident page-locked ring for pageable host frames; stage_copy_nt

// Ensemble accumulate (pyorc/velocimetry/ffpiv.py:200-243 thresholds, :361-363 accumulation)
struct EnsParams {
    float corr_min, s2n_min;
    float* plane_sum;   // [n_windows][WY][WX] fftshifted coordinates
    float* count;       // [n_windows]
};

// Round 1's staging pool (option "stage_mode" = 0; the default is the Stager of stager.h): worker threads that copy ordinary
// (pageable) host frames into three large page-locked buffers, one condition-variable round trip per buffer.  pyorc hands
// `frame_chunk.values` - plain numpy memory - to the engine (pyorc/velocimetry/ffpiv.py:223,451); a cudaMemcpyAsync from
// pageable memory is staged by the driver on ONE thread (measured: 11 GB/s, 18.7 ms per 100-pair 1080p step against 4.1 ms
// from pinned memory), so the staging is done here, sliced over a few threads, one chunk ahead of the H2D copy.
class CopyPool {
public:
    explicit CopyPool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // rows x row_bytes from src (pitch spitch) to dst (pitch dpitch), split by rows over the workers; returns when done
    void copy2d(unsigned char* dst, size_t dpitch, const unsigned char* src, size_t spitch, size_t row_bytes, size_t rows) {
        std::unique_lock<std::mutex> lk(m_);
        dst_ = dst; src_ = src; dpitch_ = dpitch; spitch_ = spitch; row_bytes_ = row_bytes; rows_ = rows;
        pending_ = (int)workers_.size();
        ++gen_;
        cv_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    int size() const { return (int)workers_.size(); }

private:
    void run(int idx) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            const size_t n = workers_.size(), per = (rows_ + n - 1) / n;
            const size_t r0 = per * idx < rows_ ? per * idx : rows_, r1 = r0 + per < rows_ ? r0 + per : rows_;
            unsigned char* d = dst_; const unsigned char* sp = src_;
            const size_t dp = dpitch_, spp = spitch_, rb = row_bytes_;
            lk.unlock();
            if (dp == rb && spp == rb) {
                if (r1 > r0) stage_copy_nt(d + r0 * rb, sp + r0 * rb, (r1 - r0) * rb);
            } else {
                for (size_t r = r0; r < r1; ++r) memcpy(d + r * dp, sp + r * spp, rb);
            }
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_ = false;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    unsigned char* dst_ = nullptr; const unsigned char* src_ = nullptr;
    size_t dpitch_ = 0, spitch_ = 0, row_bytes_ = 0, rows_ = 0;
};

struct b2piv_engine {
    int device = 0;
    std::string err;
    // options
    int clip_norm = 0, border_nan = 1, copy_chunks = 0;   // copy_chunks = 0: auto (about 10 MB of frames per H2D chunk)
      // clip_norm = 0 is what ffpiv does (pinned, tests/test_golden.py)
    int variant = 0;    // 0: auto, 1: generic shared-memory FFT kernel, 2: row-per-thread TMA kernel (error if
                        // ineligible), 3: direct any-size kernel
    int run_len = 0;    // frame pairs per work unit of the rows kernel (0: auto)
    int tmem = 1;       // 64x64 rows kernel: parked spectra in Tensor Memory (1) or in shared memory (0)
    int force_parts = 0; // > 0: force the even work partition of the rows kernels into this many parts (tests)
    int last_variant = 0;
    float gauss_eps = 1e-7f;
    // plan
    bool planned = false;
    int H = 0, W = 0, wy = 0, wx = 0, oy = 0, ox = 0, dtype = 0, n_rows = 0, n_cols = 0;
    float2 *d_twx = nullptr, *d_twy = nullptr;   // point into tw_cache
    std::map<int, float2*> tw_cache;              // transform length -> exp(-2 pi i j / n) table on the device
    int sm_count = 0;
    // streams / events
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    // device workspace for *_host calls
    unsigned char* d_frames = nullptr; size_t cap_frames = 0;
    // page-locked staging ring + copy threads for pageable host frames (pipeline_host)
    unsigned char* h_stage[3] = {nullptr, nullptr, nullptr}; size_t cap_stage = 0;
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};
    CopyPool* pool = nullptr;
    int stage_threads = 0;   // 0: auto (min(8, hardware threads))
    // stage_mode 1 (default): Stager - slices of `stage_slice_kb` per worker, one H2D per group of `threads` slices, a ring of
    // `stage_groups` groups, non-temporal stores unless `stage_nt` = 0.  Defaults from tools/stage_sweep.py on B200 boxes
    // (profiles/r02/stage_sweep_*.log): 4 MB per H2D and six of them in flight are the most even over the boxes of the pool; a small
    // cache-resident ring with plain stores loses
    int stage_mode = 1, stage_slice_kb = 512, stage_groups = 6, stage_nt = 1;
    Stager* stager = nullptr;
    unsigned char* h_ring = nullptr; size_t cap_ring = 0;
    std::vector<cudaEvent_t> ev_ring;
    float* d_out = nullptr; size_t cap_out = 0;       // 4 result fields
    float* d_planes = nullptr; size_t cap_planes = 0;
    float* d_planes_nat = nullptr; size_t cap_planes_nat = 0;   // padded rows kernel: W x W planes in natural lag order
    unsigned char* d_keep = nullptr; size_t cap_keep = 0;
    // ensemble accumulators
    float* d_pre_mean = nullptr; size_t cap_pre_mean = 0;   // pre-processing workspaces
    unsigned* d_pre_mm = nullptr; size_t cap_pre_mm = 0;
    // orthoprojection plan (CSR gather lists, project.cuh)
    int* d_proj_off = nullptr; size_t cap_proj_off = 0;
    int* d_proj_src = nullptr; size_t cap_proj_src = 0;
    int proj_h = 0, proj_w = 0, proj_out_h = 0, proj_out_w = 0; long long proj_samples = 0;
    b2piv::PeerOut peer = {};                                       // fused gather over peer memory (b2piv_set_peer_outputs)
    double* d_mp_ws = nullptr; size_t cap_mp_ws = 0;        // two-pass scheme: validated pass-1 fields
    double* d_dt = nullptr; size_t cap_dt = 0;              // per-pair time steps of b2piv_pairs_host_units
    float* d_direct_ws = nullptr; size_t cap_direct_ws = 0; // large-window direct kernel: one correlation plane per CTA
    float* d_mask_ws = nullptr; size_t cap_mask_ws = 0;     // mask stack: time statistics / window_replace ping-pong
    float* d_ens_sum = nullptr; float* d_ens_cnt = nullptr; size_t cap_ens = 0, cap_ens_windows = 0; bool ens_open = false;
    // last work enqueued on the accumulators, whatever its stream: every later user waits for it first (ens_begin / ens_add_device
    // / ens_finish may run on different streams - the engine's own or the caller's)
    cudaEvent_t ev_ens = nullptr; bool ens_pending = false;
    // explicit work-unit lists of the row-per-thread kernels (build_unit_table), cached per problem shape
    struct UnitTable { long long n_wp = -1; int n_pairs = -1, n_parts = -1, n_units = 0; long long cost = 0; int* d = nullptr; };
    UnitTable unit_tables[4];
    int unit_table_next = 0;
    // stats
    float last_kernel_ms = 0.f;
    long long launches = 0;
};

extern std::string g_create_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
            return B2PIV_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

static inline int fail(b2piv_engine* e, int code, const std::string& msg) {
    e->err = msg;
    return code;
}

template <class T>
static int ensure(b2piv_engine* e, T** ptr, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *ptr) return B2PIV_OK;
    if (*ptr) CK(cudaFree(*ptr));
    *ptr = nullptr; *cap = 0;
    CK(cudaMalloc((void**)ptr, bytes));
    *cap = bytes;
    return B2PIV_OK;
}

// The partition as a device table, cached in the engine; nullptr when the partition is not worth it / not possible
// (fewer than four items per part unless `forced`); *n_units and *cost are set.
static inline const int* build_unit_table(b2piv_engine* e, long long n_wp, int n_pairs, int n_parts, cudaStream_t st, int* n_units, long long* cost,
                                          bool forced = false) {
    const long long total = n_wp * (long long)n_pairs;
    if (n_parts < 1 || total < (forced ? 1LL : 4LL) * n_parts) return nullptr;
    for (auto& t : e->unit_tables)
        if (t.d && t.n_wp == n_wp && t.n_pairs == n_pairs && t.n_parts == n_parts) { *n_units = t.n_units; *cost = t.cost; return t.d; }
    std::vector<int> tab;
    long long worst = 0;
    const int units = partition_units(n_wp, n_pairs, n_parts, tab, &worst);
    b2piv_engine::UnitTable& t = e->unit_tables[e->unit_table_next++ % 4];
    if (t.d) { cudaStreamSynchronize(st); cudaFree(t.d); t.d = nullptr; }      // an older table may still be in use by a launch in flight
    if (cudaMalloc((void**)&t.d, tab.size() * sizeof(int)) != cudaSuccess) { t.d = nullptr; t.n_wp = -1; cudaGetLastError(); return nullptr; }
    if (cudaMemcpy(t.d, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(t.d); t.d = nullptr; t.n_wp = -1; cudaGetLastError(); return nullptr; }
    t.n_wp = n_wp; t.n_pairs = n_pairs; t.n_parts = n_parts; t.n_units = units; t.cost = worst;
    *n_units = t.n_units; *cost = t.cost;
    return t.d;
}

// ---- launch entry points of the kernel families (k_*.cu) ------------------------------------------------------------
using b2piv::Params;
bool fft_config(int wy, int wx);
void plane_shape(const b2piv_engine* e, int* py, int* px);
int launch_generic(b2piv_engine* e, const Params& p, cudaStream_t st);                               // k_generic.cu
int launch_generic_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st);      // k_generic_ens.cu
int launch_direct(b2piv_engine* e, const Params& p, cudaStream_t st);                                // k_direct.cu
int launch_direct_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st);
int launch_direct_big(b2piv_engine* e, const Params& p, cudaStream_t st);                            // windows with a side of 65 .. 128 px
int launch_direct_big_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st);
bool tma_available();                                                                                // k_rows_u8.cu
int launch_rows_u8(b2piv_engine* e, const Params& p, cudaStream_t st);                               // 32x32 / 64x64 uint8
int launch_rows_f32(b2piv_engine* e, const Params& p, cudaStream_t st);                              // k_rows_f32.cu
int launch_rows_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st);         // k_rows_ens.cu (uint8 and float32)
int launch_rows_pad(b2piv_engine* e, const Params& p, cudaStream_t st, const EnsParams* ep);         // k_rows_pad.cu
int launch_rows_pad_f32(b2piv_engine* e, const Params& p, cudaStream_t st, const EnsParams* ep);     // k_rows_pad_f32.cu
int launch_rows_shift(b2piv_engine* e, const Params& p, cudaStream_t st);                            // k_rows_shift.cu
int launch_rows128(b2piv_engine* e, const Params& p, cudaStream_t st, const EnsParams* ep = nullptr, bool pad = false);                               // k_rows128.cu
// dispatch (abi_piv.cu)
int dispatch_pairs(b2piv_engine* e, const Params& p, cudaStream_t st);
int dispatch_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st);
Params base_params(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch, int n_pairs);
