/* piv_pairs.c - the C ABI of libb2piv.so (include/b2piv.h) from plain C, no Python, no torch: what a non-Python host (or the
 * cgo / JNI / FFI stub of another language) does for pyorc's _get_uv_timestep (pyorc/velocimetry/ffpiv.py:446-474).
 *
 *   gcc -std=c99 -I include examples/piv_pairs.c -o piv_pairs -L pyorc_b200 -l:libb2piv.so -Wl,-rpath,$PWD/pyorc_b200
 *   ./piv_pairs frames.raw n H W wy wx oy ox out.raw      (frames.raw: n x H x W uint8, row-major; out.raw: u, v, corr, s2n as
 *                                                          four float32 blocks [n-1][rows][cols])
 *
 * Without a B200 the engine cannot be created and the program says so and exits with the ABI's status code: there is no CPU
 * fallback (tests/test_abi.py::test_c_example_builds_links_and_fails_loudly_without_a_gpu). */
#include "b2piv.h"

#include <stdio.h>
#include <stdlib.h>

static int die(const b2piv_engine* e, const char* what, int rc) {
    fprintf(stderr, "%s failed (status %d): %s\n", what, rc, b2piv_last_error(e));
    return rc ? rc : 1;
}

int main(int argc, char** argv) {
    if (argc != 10) {
        fprintf(stderr, "usage: %s frames.raw n H W wy wx oy ox out.raw   (libb2piv ABI version %d)\n", argv[0], b2piv_version());
        return 64;
    }
    const int n = atoi(argv[2]), H = atoi(argv[3]), W = atoi(argv[4]);
    const int wy = atoi(argv[5]), wx = atoi(argv[6]), oy = atoi(argv[7]), ox = atoi(argv[8]);
    if (n < 2 || H < 1 || W < 1) { fprintf(stderr, "need at least two frames\n"); return 64; }

    b2piv_engine* e = NULL;
    int rc = b2piv_create(&e, 0);
    if (rc != B2PIV_OK) return die(NULL, "b2piv_create", rc);      /* no B200 / no driver: loud failure, nothing is computed */

    int rows = 0, cols = 0;
    rc = b2piv_plan(e, H, W, wy, wx, oy, ox, B2PIV_U8, &rows, &cols);
    if (rc != B2PIV_OK) { rc = die(e, "b2piv_plan", rc); b2piv_destroy(e); return rc; }

    const size_t n_px = (size_t)n * H * W, n_res = (size_t)(n - 1) * rows * cols;
    unsigned char* frames = (unsigned char*)b2piv_host_alloc(n_px);              /* page-locked: H2D at the full PCIe rate */
    float* out = (float*)malloc(4 * n_res * sizeof(float));
    FILE* f = fopen(argv[1], "rb");
    if (!frames || !out || !f || fread(frames, 1, n_px, f) != n_px) {
        fprintf(stderr, "cannot read %zu bytes of frames from %s\n", n_px, argv[1]);
        if (f) fclose(f);
        b2piv_host_free(frames); free(out); b2piv_destroy(e);
        return 66;
    }
    fclose(f);

    /* u, v in pixels per frame, corr = max of the clipped correlation plane, s2n = max / mean; signal_threshold < 0: off */
    rc = b2piv_pairs_host(e, frames, n, -1.0f, out, out + n_res, out + 2 * n_res, out + 3 * n_res);
    if (rc != B2PIV_OK) { rc = die(e, "b2piv_pairs_host", rc); b2piv_host_free(frames); free(out); b2piv_destroy(e); return rc; }

    float ms = 0.f;
    b2piv_last_kernel_ms(e, &ms);
    printf("%d pairs of %dx%d, %dx%d windows on a %dx%d grid: kernel family %d, %lld launch(es), last kernel %.3f ms\n", n - 1, H, W, wy, wx,
           rows, cols, b2piv_last_variant(e), b2piv_launch_count(e), ms);
    f = fopen(argv[9], "wb");
    if (!f || fwrite(out, sizeof(float), 4 * n_res, f) != 4 * n_res) { fprintf(stderr, "cannot write %s\n", argv[9]); rc = 73; }
    if (f) fclose(f);
    b2piv_host_free(frames);
    free(out);
    b2piv_destroy(e);
    return rc;
}
