// TEST-ONLY: drives pyorc_b200/csrc/stager.h without CUDA.  A thread plays the copy engine: it takes the enqueued "H2D copies"
// in order, waits a little, copies the ring slot into a pitched destination (like cudaMemcpy2DAsync) and marks the group as
// landed.  tests/test_host_logic.py checks that every byte arrives once whatever the slice / group / ring / thread counts are,
// that the chunk hook sees monotone row counts, and that an error from `issue` ends the job without a hang.
#include "../../pyorc_b200/csrc/stager.h"

#include <chrono>
#include <deque>

namespace {
struct Copy { size_t g; const unsigned char* ring; size_t row0, rows; };
}

extern "C" long long stager_selftest(long long rows, long long row_bytes, long long dpitch, long long slice_rows, int parts, int ring_groups,
                                     int threads, int nt, long long fail_at_group, int dma_delay_us, int repeats) {
    std::vector<unsigned char> src((size_t)rows * row_bytes), dst((size_t)rows * dpitch, 0xEE);
    unsigned x = 12345u;
    for (auto& b : src) { x = x * 1664525u + 1013904223u; b = (unsigned char)(x >> 24); }
    Stager::Job job;
    job.src = src.data(); job.row_bytes = (size_t)row_bytes; job.rows = (size_t)rows; job.slice_rows = (size_t)slice_rows;
    job.parts = parts; job.ring_groups = ring_groups; job.nt = nt != 0;
    std::vector<unsigned char> ring(Stager::ring_bytes(job) + 64);
    job.ring = ring.data() + ((64 - ((uintptr_t)ring.data() & 63)) & 63);
    const size_t ng = Stager::n_groups(job);
    Stager st(threads);
    long long bad = 0;
    for (int rep = 0; rep < repeats; ++rep) {
        std::fill(dst.begin(), dst.end(), (unsigned char)0xEE);
        std::unique_ptr<std::atomic<int>[]> landed(new std::atomic<int>[ng]);
        for (size_t g = 0; g < ng; ++g) landed[g].store(0);
        std::mutex qm;
        std::deque<Copy> q;
        std::atomic<bool> stop{false};
        std::atomic<long long> in_flight_max{0}, in_flight{0};
        std::thread dma([&] {
            for (;;) {
                Copy c;
                {
                    std::unique_lock<std::mutex> lk(qm);
                    if (q.empty()) {
                        if (stop.load()) return;
                        lk.unlock();
                        std::this_thread::yield();
                        continue;
                    }
                    c = q.front();
                    q.pop_front();
                }
                if (dma_delay_us) std::this_thread::sleep_for(std::chrono::microseconds(dma_delay_us));
                for (size_t r = 0; r < c.rows; ++r)
                    memcpy(dst.data() + (c.row0 + r) * (size_t)dpitch, c.ring + r * (size_t)row_bytes, (size_t)row_bytes);
                in_flight.fetch_sub(1);
                landed[c.g].store(1, std::memory_order_release);
            }
        });
        size_t last_rows = 0, expect_g = 0;
        const int rc = st.run(
            job,
            [&](size_t g, unsigned char* p, size_t row0, size_t n) -> int {
                if (g != expect_g++) ++bad;                                      // groups are issued in order
                if ((long long)g == fail_at_group) return 7;
                const long long f = in_flight.fetch_add(1) + 1;
                if (f > in_flight_max.load()) in_flight_max.store(f);
                std::lock_guard<std::mutex> lk(qm);
                q.push_back(Copy{g, p, row0, n});
                return 0;
            },
            [&](size_t g) -> bool { return landed[g].load(std::memory_order_acquire) != 0; },
            [&](size_t r) -> int {
                if (r <= last_rows || r > (size_t)rows) ++bad;                 // monotone, never beyond the job
                last_rows = r;
                return 0;
            });
        stop.store(true);
        dma.join();
        if (fail_at_group >= 0 && fail_at_group < (long long)ng) {
            if (rc != 7) ++bad;
            continue;
        }
        if (rc != 0 || last_rows != (size_t)rows) ++bad;
        if (in_flight_max.load() > ring_groups) ++bad;                            // a slot is never rewritten under a copy
        for (long long r = 0; r < rows; ++r) {
            if (memcmp(dst.data() + r * dpitch, src.data() + r * row_bytes, (size_t)row_bytes) != 0) ++bad;
            for (long long k = row_bytes; k < dpitch; ++k)
                if (dst[(size_t)(r * dpitch + k)] != 0xEE) ++bad;                 // the pitch padding stays untouched
        }
    }
    return bad;
}

// ---- work partition of the row-per-thread kernels (pyorc_b200/csrc/work_partition.h) -------------------------------------
#include "../../pyorc_b200/csrc/work_partition.h"

extern "C" int emul_pick_run_len(int n_pairs, long long n_wp, long long resident) { return pick_run_len(n_pairs, n_wp, resident); }

// writes at most `cap` ints of the [round][part][3] table; returns the number of units (rounds * n_parts), *cost as in the engine
extern "C" int emul_partition_units(long long n_wp, int n_pairs, int n_parts, int* out, long long cap, long long* cost) {
    std::vector<int> tab;
    const int units = partition_units(n_wp, n_pairs, n_parts, tab, cost);
    for (size_t i = 0; i < tab.size() && (long long)i < cap; ++i) out[i] = tab[i];
    return units;
}
