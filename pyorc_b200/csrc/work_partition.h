// work_partition.h - host arithmetic that deals the (window pair, frame pair) space to the resident groups of the row-per-thread
// kernels.  No CUDA in here: engine.h turns the tables into device memory, tests/emul/stager_emul.cpp exports both functions to the
// CPU tests (tests/test_host_logic.py::test_unit_table_*, test_run_length_*).
#pragma once
#include <algorithm>
#include <vector>

// Frame pairs per work unit of the row-per-thread kernels.  A unit follows its window pair through `run` consecutive frame
// pairs and pays ONE extra forward transform at its start, and the units are dealt to `resident` persistent groups in waves:
// cost ~ ceil(units / resident) * (run + 1) frame times.  The number of time chunks that minimises it is searched (round 1
// aimed at >= 8 waves, which for 100 pairs of 1080p - 944 window pairs on 592 groups - gave 6 chunks: 9.6 waves, the last one
// 57 % full, 180 frame times; 5 chunks fill 7.97 waves: 168).
static inline int pick_run_len(int n_pairs, long long n_wp, long long resident) {
    long long best_cost = -1;
    int best_run = n_pairs;
    for (int c = 1; c <= n_pairs && c <= 64; ++c) {
        const int run = (n_pairs + c - 1) / c;
        const long long chunks = (n_pairs + run - 1) / run;
        const long long waves = (n_wp * chunks + resident - 1) / resident;
        const long long cost = waves * (run + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_run = run; }
    }
    return best_run < 1 ? 1 : best_run;
}

// Even 1-D partition of the (window pair, frame pair) space over `n_parts` independent groups: part i gets the items
// [i * total / n_parts, (i + 1) * total / n_parts) of the window-pair-major list, cut at window-pair boundaries into segments
// (a segment = one work unit: a window pair followed through consecutive frame pairs, one extra forward transform at its start).
// Layout [round][part][3] = (window pair, first frame pair, one past the last) so that the kernels' round-robin walk
// (unit = part + round * n_parts) gives part i its own segments; (-1, 0, -1) pads the parts with fewer segments.
// Pure host arithmetic (tests/emul exports it: tests/test_host_logic.py::test_unit_table_*).  Returns the number of units
// (rounds * n_parts); *cost = frame times of the longest part (frame pairs + one per segment).
static inline int partition_units(long long n_wp, int n_pairs, int n_parts, std::vector<int>& tab, long long* cost) {
    const long long total = n_wp * (long long)n_pairs;
    std::vector<std::vector<int>> segs((size_t)n_parts);
    size_t max_seg = 0;
    long long worst = 0;
    for (int i = 0; i < n_parts; ++i) {
        long long pos = total * i / n_parts;
        const long long end = total * (i + 1) / n_parts;
        long long c = 0;
        while (pos < end) {
            const long long wp = pos / n_pairs;
            const int f0 = (int)(pos % n_pairs);
            const long long len = std::min<long long>(n_pairs - f0, end - pos);
            segs[i].push_back((int)wp); segs[i].push_back(f0); segs[i].push_back(f0 + (int)len);
            c += len + 1;
            pos += len;
        }
        max_seg = std::max(max_seg, segs[i].size() / 3);
        worst = std::max(worst, c);
    }
    tab.assign(max_seg * (size_t)n_parts * 3, 0);
    for (size_t rd = 0; rd < max_seg; ++rd)
        for (int i = 0; i < n_parts; ++i) {
            int* dst = &tab[(rd * (size_t)n_parts + i) * 3];
            if (rd < segs[i].size() / 3) { dst[0] = segs[i][3 * rd]; dst[1] = segs[i][3 * rd + 1]; dst[2] = segs[i][3 * rd + 2]; }
            else { dst[0] = -1; dst[1] = 0; dst[2] = -1; }
        }
    *cost = worst;
    return (int)(max_seg * (size_t)n_parts);
}
