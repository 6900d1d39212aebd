// k_rows_u8.cu - row-per-thread kernel, native 32x32 / 64x64 uint8 windows, per-time-step mode (the headline kernel).
#include "rows_kernel.cuh"

#ifndef B2_TM_GROUPS
#define B2_TM_GROUPS 6
#endif

bool tma_available() { return get_encode_tiled() != nullptr; }

// Compiled variants (measured on B200, profiles/r01/quick_sweeps.log): 64x64 is fastest with one group per CTA (four
// 64-thread CTAs per SM) and one shared FFT body, 32x32 with four single-warp groups and two FFT copies.
int launch_rows_u8(b2piv_engine* e, const Params& p, cudaStream_t st) {
    const bool aligned = ((e->wx - e->ox) & 15) == 0;
    // 64 x 64: parked spectra in Tensor Memory, B2_TM_GROUPS groups per SM instead of four (option "tmem" = 0 switches back)
    if (e->wy == 64 && e->tmem) return aligned ? launch_rows_tm<RCfg<64>, B2_TM_GROUPS, true>(e, p, st) : launch_rows_tm<RCfg<64>, B2_TM_GROUPS, false>(e, p, st);
    if (e->wy == 64) return aligned ? launch_rows<RCfg<64>, 1, true, true, false>(e, p, st) : launch_rows<RCfg<64>, 1, true, false, false>(e, p, st);
    return aligned ? launch_rows<RCfg<32>, 4, false, true, false>(e, p, st) : launch_rows<RCfg<32>, 4, false, false, false>(e, p, st);
}
