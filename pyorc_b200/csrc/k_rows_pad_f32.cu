// k_rows_pad_f32.cu - padded mode of the row-per-thread kernel for float32 frames: any window (square or not, any stride) whose
// larger side is at most 32 px - what pyorc's own recipe hands over (normalize -> edge_detect -> minmax -> get_piv(window_size=25):
// float32 frames, 26 x 26 windows; examples/ngwerere/ngwerere.yml).  Un-swizzled (W/2 + 4 floats) x ny TMA boxes, two-pass moments.
#include "rows_kernel.cuh"

int launch_rows_pad_f32(b2piv_engine* e, const Params& p, cudaStream_t st, const EnsParams* ep) {
    const int m = e->wy > e->wx ? e->wy : e->wx;
    if (ep) {
        if (2 * m <= 32) return launch_rows<RCfg<32>, 4, false, false, true, true, true>(e, p, st, ep);
        return launch_rows<RCfg<64>, 1, true, false, true, true, true>(e, p, st, ep);
    }
    if (2 * m <= 32) return launch_rows<RCfg<32>, 4, false, false, true, false, true>(e, p, st, nullptr);
    return launch_rows<RCfg<64>, 1, true, false, true, false, true>(e, p, st, nullptr);
}
