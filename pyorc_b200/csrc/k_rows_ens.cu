// k_rows_ens.cu - row-per-thread kernel in ensemble mode (planes added to the HBM accumulators, no per-pair peak search).
#include "rows_kernel.cuh"

int launch_rows_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
    const bool aligned = ((e->wx - e->ox) & 15) == 0;
    if (e->dtype == B2PIV_F32) {
        if (e->wy == 64) return launch_rows<RCfg<64>, 1, true, true, true, true>(e, p, st, &ep);
        return launch_rows<RCfg<32>, 4, false, true, true, true>(e, p, st, &ep);
    }
    if (e->wy == 64) return aligned ? launch_rows<RCfg<64>, 1, true, true, false, true>(e, p, st, &ep) : launch_rows<RCfg<64>, 1, true, false, false, true>(e, p, st, &ep);
    return aligned ? launch_rows<RCfg<32>, 4, false, true, false, true>(e, p, st, &ep) : launch_rows<RCfg<32>, 4, false, false, false, true>(e, p, st, &ep);
}
