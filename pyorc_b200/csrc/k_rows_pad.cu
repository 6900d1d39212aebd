// k_rows_pad.cu - padded mode of the row-per-thread kernel: any uint8 window (square or not, any stride) whose larger side is
// at most 32 px, i.e. at most half of a 64 x 64 (or 32 x 32) plane.
#include "rows_kernel.cuh"

// triage path of the padded rows kernel: W x W planes in natural lag order -> the reference's fftshifted ny x nx planes
__global__ void planes_reorder_kernel(const float* __restrict__ nat, float* __restrict__ out, long long n_planes, int W, int ny, int nx) {
    const long long n = n_planes * ny * nx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % nx), iy = (int)((i / nx) % ny);
        const long long pl = i / ((long long)nx * ny);
        const int hy = ny / 2, hx = nx / 2;
        const int qy = iy < hy ? iy + ny - hy : iy - hy, qx = ix < hx ? ix + nx - hx : ix - hx;   // lag (j + n - n/2) % n
        out[i] = nat[(pl * W + qy) * W + qx];
    }
}

int launch_rows_pad(b2piv_engine* e, const Params& p, cudaStream_t st, const EnsParams* ep) {
    const int m = e->wy > e->wx ? e->wy : e->wx;
    if (ep) {
        if (2 * m <= 32) return launch_rows<RCfg<32>, 4, false, false, false, true, true>(e, p, st, ep);
        return launch_rows<RCfg<64>, 1, true, false, false, true, true>(e, p, st, ep);
    }
    if (2 * m <= 32) return launch_rows<RCfg<32>, 4, false, false, false, false, true>(e, p, st, nullptr);
    return launch_rows<RCfg<64>, 1, true, false, false, false, true>(e, p, st, nullptr);
}
