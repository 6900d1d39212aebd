"""CPU, this container only: the reference's own EXAMPLE RESULT against this engine's, end to end.

pyorc ships `examples/ngwerere/ngwerere_piv.nc` - the Dataset its notebook 02 writes: frames 0..125 of the Ngwerere clip,
`normalize()` -> `project(method="numpy")` with `examples/ngwerere/ngwerere.json` -> `get_piv()` (window_size 25 of the camera
configuration).  The file was written by an older pyorc (OpenPIV engine: 25 px windows, stride 13) on stabilised frames, so it pins
no VALUE of today's ffpiv path - but it is the only place where the reference shows what `v_x` AND `v_y` of this river look like, and
pyorc's own tests never check `v_y` ("treat the sign as to be confirmed").  Here the same recipe runs through the
reference's own decode / camera / projection code (tests/ref_harness.py), this repository's `get_piv` (window 25 -> 26, overlap 12,
px -> m/s, result axes) with the float64 oracle as the engine (no GPU in this container; the CUDA engine equals the oracle to 2e-3 px,
tests/test_gpu_parity.py), and the time-median fields are compared with the shipped ones at the nearest grid points:

  * `v_x`: same sign at > 95 % of the points, medians within 25 %, spatial correlation > 0.5
  * `v_y`: POSITIVE spatial correlation > 0.5 - the sign convention of `v_y` (no flip after ffpiv: +v_y = increasing row) and the
    units (m/s = px * resolution / dt) agree with what the reference publishes.

All 125 pairs (python tests/test_reference_example.py): v_x correlation 0.73, medians 0.236 vs 0.25 m/s, sign agreement 99.3 %;
v_y correlation +0.79 (a flipped convention would give -0.79).  The netCDF4 files are read with tests/golden/h5min.py.
"""
import importlib
import os
import sys
import warnings

import numpy as np
import pytest

sys.path[:0] = [p for p in (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))) if p not in sys.path]
import ref_harness  # noqa: E402
from oracle import ffpiv_oracle as O, mask_oracle as MO, preprocess_oracle as PO  # noqa: E402
from pyorc_b200 import _xr, frames as b2frames, velocimetry, window as b2window  # noqa: E402
from test_host_logic import FakeEngine  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference is not present (GPU box): the reference's Python and example files cannot travel")

EXAMPLE = os.path.join(ref_harness.REF, "examples", "ngwerere")


def example_pipeline(n_pairs):
    """Frames 0 .. n_pairs of the clip as notebook 02 prepares them (without its stabilisation): grayscale -> normalize ->
    project(method="numpy") with the example's camera configuration.  Decode, camera model, index maps and group means are the
    reference's code; `normalize` is its numpy expression (oracle/preprocess_oracle.py, pinned in tests/test_preprocess.py)."""
    import cv2

    ref_harness.install_stubs()
    try:
        cv = importlib.import_module("pyorc.cv")
        cameraconfig = importlib.import_module("pyorc.api.cameraconfig")
        project = importlib.import_module("pyorc.project")
        cc = cameraconfig.load_camera_config(os.path.join(EXAMPLE, "ngwerere.json"))
        shape, res, win = tuple(cc.shape), float(cc.resolution), cc.window_size
        fn = os.path.join(EXAMPLE, "ngwerere_20191103.mp4")
        cap = cv2.VideoCapture(fn)
        times, numbers, _ = cv.get_time_frames(cap, 0, n_pairs, lazy=True, progress=False, method="bgr")
        cap.release()
        cap = cv2.VideoCapture(fn)
        cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
        imgs = []
        for _ in numbers:
            ret, img = cv.get_frame(cap, rotation=None, ms=None, method="grayscale")
            assert ret
            imgs.append(img)
        cap.release()
        norm = PO.normalize(np.array(imgs), 15)
        y = np.flipud(np.linspace(res / 2, res * (shape[0] - 0.5), shape[0]))
        x = np.linspace(res / 2, res * (shape[1] - 0.5), shape[1])
        z = cc.get_z_a(0.0)
        idx_img, idx_ortho = cc.map_idx_img_ortho(x, y, z)
        src_idx, uidx, norm_idx = cc.map_mean_idx_img_ortho(x, y, z)
        proj = np.stack([np.nan_to_num(project.img_to_ortho(im, x, y, idx_img, idx_ortho, src_idx, uidx, norm_idx), nan=0.0).astype(np.uint8)
                         for im in norm])
    finally:
        ref_harness.remove_stubs()
    return _xr.DataArray(proj, ("time", "y", "x"), {"time": np.array(times) * 0.001, "y": y, "x": x}), res, win


def shipped_result():
    sys.path.insert(0, ref_harness.GOLDEN)
    import h5min

    h = h5min.H5(os.path.join(EXAMPLE, "ngwerere_piv.nc"))
    names = h5min.dataset_names(h)
    r = {k: h.read(names[k]) for k in ("v_x", "v_y", "corr", "x", "y")}
    return {"v_x": MO.decode_int16(r["v_x"]), "v_y": MO.decode_int16(r["v_y"]), "corr": MO.decode_int16(r["corr"]), "x": r["x"], "y": r["y"]}


def compare(n_pairs):
    frames, res, win = example_pipeline(n_pairs)
    assert win == 25 and frames.shape[1:] == (785, 875)
    fe = FakeEngine()
    saved = velocimetry.get_engine, b2window.available_memory
    velocimetry.get_engine = lambda device=0, slot=0: fe
    b2window.available_memory = lambda device=None, **kw: 64e9
    O.CLIP_NORMALIZED = False
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ds = b2frames.get_piv(frames, window_size=win, resolution=res, chunksize=9)     # window 25 -> 26, overlap 12 (frames.py:159-171)
    finally:
        velocimetry.get_engine, b2window.available_memory = saved
    ref = shipped_result()
    vx, vy, c = (np.asarray(ds[k].values) for k in ("v_x", "v_y", "corr"))
    assert vx.shape == (n_pairs, 55, 61)
    xo, yo = np.asarray(ds.coords["x"]), np.asarray(ds.coords["y"])

    def med(a, cm):   # time median over the pairs whose correlation is decent
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return np.nanmedian(np.where(cm > 0.3, a, np.nan), axis=0)

    ours_x, ours_y = med(vx, c), med(vy, c)
    ref_x, ref_y = med(ref["v_x"][:n_pairs], ref["corr"][:n_pairs]), med(ref["v_y"][:n_pairs], ref["corr"][:n_pairs])
    jj = np.abs(ref["x"][:, None] - xo[None]).argmin(1)          # nearest grid point (local coordinates, m)
    ii = np.abs(ref["y"][:, None] - yo[None]).argmin(1)
    ox, oy = ours_x[np.ix_(ii, jj)], ours_y[np.ix_(ii, jj)]
    ok = np.isfinite(ox) & np.isfinite(oy) & np.isfinite(ref_x) & np.isfinite(ref_y) & (np.hypot(ref_x, ref_y) > 0.05)
    return {"points": int(ok.sum()),
            "vx_corr": float(np.corrcoef(ox[ok], ref_x[ok])[0, 1]), "vy_corr": float(np.corrcoef(oy[ok], ref_y[ok])[0, 1]),
            "vx_median": (float(np.median(ox[ok])), float(np.median(ref_x[ok]))),
            "vx_sign_agreement": float(np.mean(np.sign(ox[ok]) == np.sign(ref_x[ok])))}


def test_sign_conventions_and_units_agree_with_the_reference_example():
    pytest.importorskip("cv2")
    r = compare(40)
    assert r["points"] > 800
    assert r["vx_sign_agreement"] > 0.95
    assert 0.8 < r["vx_median"][0] / r["vx_median"][1] < 1.25      # m / s: px * resolution / dt
    assert r["vx_corr"] > 0.5
    assert r["vy_corr"] > 0.5                                          # positive: v_y has the reference's sign convention


if __name__ == "__main__":
    print(compare(int(sys.argv[1]) if len(sys.argv) > 1 else 125))
