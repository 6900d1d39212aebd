// stager.h - copy threads that feed a page-locked ring for the H2D copy of ordinary (pageable) host frames.
//
// pyorc hands `frame_chunk.values` - plain numpy memory - to the engine (pyorc/velocimetry/ffpiv.py:223,451).  A
// cudaMemcpyAsync from such memory is staged by the driver on one thread (11 GB/s); the engine stages it itself.  What the
// round-2 measurements showed: with the frames copied once into a large page-locked ring with non-temporal stores
// every byte crosses the host's memory bus three times (source read, ring write, DMA read), and that - about 140-160 GB/s in
// total on this pool's boxes - is what bounds the call as soon as two or more GPUs of one host are fed at the same time.
//
// This stager is built so that the ring can live in the caches: the (frames x rows) of a call are cut into GROUPS of `parts`
// slices; a slice is a few hundred KB copied by whichever worker is free, a group is one dense piece of the ring and ONE
// cudaMemcpyAsync, and only `ring_groups` groups are in flight.  With plain stores a ring of a few MB stays dirty in L2 / L3,
// the DMA read is served from there and the next round of stores hits the same lines: the source read is the only DRAM
// traffic.  Non-temporal stores (the earlier design) remain selectable.  Workers are woken ONCE per call and hand slices over
// through atomics (the earlier pool paid a condition-variable round trip per 6 MB).
//
// No CUDA in this header: the owner passes three callables - `issue(group, ring_ptr, first_row, n_rows)` enqueues the H2D of a
// staged group, `landed(group)` says whether that copy has completed (its ring slot may be overwritten), `issued(rows)` is told
// how many rows have been enqueued so far (the engine launches a chunk's kernel when its frames are under way).  tests/emul
// drives it with a thread that plays the DMA engine.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define B2PIV_CPU_RELAX() _mm_pause()
__attribute__((target("avx2"))) static inline void stream_copy_avx2(unsigned char* d, const unsigned char* s, size_t n) {
    size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head > n) head = n;
    if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i)), b = _mm256_loadu_si256((const __m256i*)(s + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + i + 96));
        _mm256_stream_si256((__m256i*)(d + i), a); _mm256_stream_si256((__m256i*)(d + i + 32), b);
        _mm256_stream_si256((__m256i*)(d + i + 64), c); _mm256_stream_si256((__m256i*)(d + i + 96), e);
    }
    if (i < n) memcpy(d + i, s + i, n - i);
    _mm_sfence();
}
// Copy into a page-locked staging buffer with NON-TEMPORAL stores: the destination is neither fetched into the caches first
// (read-for-ownership) nor does it evict the source.  glibc switches to such stores only for single copies of many MB.
static inline void stage_copy_nt(unsigned char* d, const unsigned char* s, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) stream_copy_avx2(d, s, n); else memcpy(d, s, n);
}
#else
#define B2PIV_CPU_RELAX() ((void)0)
static inline void stage_copy_nt(unsigned char* d, const unsigned char* s, size_t n) { memcpy(d, s, n); }
#endif

class Stager {
public:
    struct Job {
        const unsigned char* src = nullptr;   // dense rows of `row_bytes`
        unsigned char* ring = nullptr;        // ring_groups * parts * slice_rows * row_bytes bytes, page-locked
        size_t row_bytes = 0, rows = 0;       // rows = frames * H
        size_t slice_rows = 0;                // rows per slice (one worker, one copy)
        int parts = 1;                        // slices per group (a group is one H2D copy)
        int ring_groups = 3;                  // groups in flight
        bool nt = false;                      // non-temporal stores
    };

    explicit Stager(int n_threads) {
        for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~Stager() {
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }

    static size_t group_rows(const Job& j) { return j.slice_rows * (size_t)j.parts; }
    static size_t n_groups(const Job& j) { return (j.rows + group_rows(j) - 1) / group_rows(j); }
    static size_t ring_bytes(const Job& j) { return (size_t)j.ring_groups * group_rows(j) * j.row_bytes; }

    // Runs one call on the calling thread (which makes all the `issue / landed / issued` calls, i.e. all CUDA calls) and the
    // workers.  Returns 0, or the first non-zero value an `issue` / `issued` call returned (the rest of the job is dropped;
    // copies already enqueued are the caller's to wait for).
    template <class Issue, class Landed, class Issued>
    int run(const Job& job, Issue&& issue, Landed&& landed, Issued&& issued) {
        const size_t ng = n_groups(job), gr = group_rows(job);
        if (ng == 0) return 0;
        job_ = job;
        n_slices_ = (job.rows + job.slice_rows - 1) / job.slice_rows;
        staged_.reset(new std::atomic<int>[ng]);
        for (size_t g = 0; g < ng; ++g) staged_[g].store(0, std::memory_order_relaxed);
        next_.store(0, std::memory_order_relaxed);
        freed_.store(0, std::memory_order_relaxed);
        abort_.store(false, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(m_);
            active_ = (int)workers_.size();
            ++gen_;
        }
        cv_.notify_all();
        int rc = 0;
        size_t retired = 0;   // groups whose H2D has landed
        for (size_t g = 0; g < ng && !rc; ++g) {
            const size_t r0 = g * gr, r1 = r0 + gr < job.rows ? r0 + gr : job.rows;
            const int want = (int)((r1 - r0 + job.slice_rows - 1) / job.slice_rows);
            unsigned spins = 0;
            for (;;) {
                // slots whose copy has landed go back to the workers - also while waiting: they may be waiting for exactly that
                while (retired < g && landed(retired)) freed_.store(++retired, std::memory_order_release);
                if (staged_[g].load(std::memory_order_acquire) == want) break;
                relax(spins);
            }
            rc = issue(g, job.ring + (g % (size_t)job.ring_groups) * gr * job.row_bytes, r0, r1 - r0);
            if (!rc) rc = issued(r1);
        }
        if (rc) abort_.store(true, std::memory_order_release);
        // the workers leave the job before its description may change
        std::unique_lock<std::mutex> lk(m_);
        idle_.wait(lk, [this] { return active_ == 0; });
        return rc;
    }

private:
    // Waiting (for a slice to be staged, for an H2D to land): a short spin, then give the core away, then sleep in steps of
    // ~20 us + timer slack - a waiting thread must not keep a copying one off a core when the host is shared by several ranks
    static void relax(unsigned& spins) {
        ++spins;
        if (spins < 128) B2PIV_CPU_RELAX();
        else if (spins < 144) std::this_thread::yield();
        else std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_;
            }
            const Job& j = job_;
            const size_t parts = (size_t)j.parts, gr = group_rows(j);
            for (;;) {
                const size_t s = next_.fetch_add(1, std::memory_order_relaxed);
                if (s >= n_slices_) break;
                const size_t g = s / parts, part = s % parts;
                unsigned spins = 0;
                while (g >= freed_.load(std::memory_order_acquire) + (size_t)j.ring_groups) {   // slot still read by an H2D
                    if (abort_.load(std::memory_order_acquire)) break;
                    relax(spins);
                }
                if (abort_.load(std::memory_order_acquire)) break;
                const size_t r0 = g * gr + part * j.slice_rows;
                const size_t r1 = r0 + j.slice_rows < j.rows ? r0 + j.slice_rows : j.rows;
                unsigned char* d = j.ring + ((g % (size_t)j.ring_groups) * gr + part * j.slice_rows) * j.row_bytes;
                const unsigned char* sp = j.src + r0 * j.row_bytes;
                if (j.nt) stage_copy_nt(d, sp, (r1 - r0) * j.row_bytes); else memcpy(d, sp, (r1 - r0) * j.row_bytes);
                staged_[g].fetch_add(1, std::memory_order_release);
            }
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) idle_.notify_one();
        }
    }

    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, idle_;
    bool quit_ = false;
    unsigned long long gen_ = 0;
    int active_ = 0;
    Job job_;
    size_t n_slices_ = 0;
    std::unique_ptr<std::atomic<int>[]> staged_;
    std::atomic<size_t> next_{0}, freed_{0};
    std::atomic<bool> abort_{false};
};
