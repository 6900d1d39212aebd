// k_rows_f32.cu - row-per-thread kernel, float32 frames (128B-swizzled TMA boxes, two-pass moments), per-time-step mode.
#include "rows_kernel.cuh"

int launch_rows_f32(b2piv_engine* e, const Params& p, cudaStream_t st) {
    if (e->wy == 64) return launch_rows<RCfg<64>, 1, true, true, true>(e, p, st);
    return launch_rows<RCfg<32>, 4, false, true, true>(e, p, st);
}
