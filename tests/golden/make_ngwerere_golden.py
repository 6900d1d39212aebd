#!/usr/bin/env python
"""Generate the Ngwerere golden fixture by running the REFERENCE's own Python (imported from /root/reference) for
everything up to the ffpiv call: video decode -> grayscale -> CameraConfig (PnP, bbox from corners) ->
Frames.project(method="numpy") index maps -> img_to_ortho.  This is the input of pyorc's only numerical pin of the
hot path, tests/test_frames.py:139-153 (window_size=10, s2n_min=corr_min=count_min=0, v_x mean over time, last 4).

pyorc's heavy optional dependencies that are absent offline (xarray, dask, shapely, rasterio, pyproj, geopandas,
matplotlib) are replaced by the minimal stand-ins below; only geometry primitives that the projection path really
executes are implemented (rotate, bounds, LineString.length, Affine indexing, pixel-centre rasterize).  Everything
numerical (cv2.solvePnP, projectPoints, undistortPoints, the index maps, the float32 group means) is the reference's
code, unmodified.  Runs only where /root/reference exists (this container); the GPU box uses the committed .npz.

Usage: python tests/golden/make_ngwerere_golden.py  ->  tests/golden/ngwerere_proj.npz
"""
import importlib
import os
import sys
import types
from unittest import mock

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ngwerere_proj.npz")
OUT_MAPS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ngwerere_maps.npz")
OUT_TAIL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ngwerere_gray_tail.npz")


# ---------------------------------------------------------------------------------------------------------------
# minimal stand-ins
# ---------------------------------------------------------------------------------------------------------------
class _Ring:
    def __init__(self, coords):
        self.coords = [tuple(float(v) for v in c) for c in coords]


class Polygon:
    def __init__(self, coords):
        coords = [tuple(float(v) for v in c) for c in np.asarray(coords, dtype=np.float64)]
        if coords[0] != coords[-1]:
            coords = coords + [coords[0]]
        self.exterior = _Ring(coords)

    @property
    def bounds(self):
        a = np.array(self.exterior.coords)
        return (a[:, 0].min(), a[:, 1].min(), a[:, 0].max(), a[:, 1].max())

    @property
    def has_z(self):
        return len(self.exterior.coords[0]) == 3

    @property
    def is_empty(self):
        return False


class LineString:
    def __init__(self, coords):
        self.coords = np.asarray(coords, dtype=np.float64)

    @property
    def length(self):
        return float(np.sqrt((np.diff(self.coords, axis=0) ** 2).sum(axis=1)).sum())


class Point:
    def __init__(self, *a):
        self.xy = a


def rotate(geom, angle, origin, use_radians=False):
    if not use_radians:
        angle = angle * np.pi / 180.0
    c, s = np.cos(angle), np.sin(angle)
    x0, y0 = origin[0], origin[1]
    out = []
    for p in geom.exterior.coords:
        x, y = p[0], p[1]
        # shapely.affinity.rotate: affine matrix [cos, -sin, sin, cos, xoff, yoff]
        xoff = x0 - x0 * c + y0 * s
        yoff = y0 - x0 * s - y0 * c
        out.append((c * x - s * y + xoff, s * x + c * y + yoff))
    return Polygon(out)


class Affine(tuple):
    def __new__(cls, a, b, c, d, e, f):
        return tuple.__new__(cls, (a, b, c, d, e, f, 0.0, 0.0, 1.0))


def rasterize(shapes, out_shape, **kw):
    """GDAL-style polygon burn (all_touched=False): a pixel is set when its CENTRE lies inside the polygon."""
    poly = shapes[0]
    pts = np.array(poly.exterior.coords)[:, :2]
    h, w = out_shape
    yy, xx = np.mgrid[0:h, 0:w]
    px, py = xx + 0.5, yy + 0.5
    inside = np.zeros(out_shape, dtype=bool)
    x0, y0 = pts[:-1, 0], pts[:-1, 1]
    x1, y1 = pts[1:, 0], pts[1:, 1]
    for ax, ay, bx, by in zip(x0, y0, x1, y1):
        if ay == by:
            continue
        cond = ((ay <= py) & (py < by)) | ((by <= py) & (py < ay))
        xint = ax + (py - ay) * (bx - ax) / (by - ay)
        inside ^= cond & (px < xint)
    return inside.astype(np.uint8)


class _CRS:
    is_geographic = 0

    @classmethod
    def from_user_input(cls, x):
        return cls()

    def to_wkt(self):
        return "EPSG:32735"


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "matplotlib.colors", "matplotlib.collections",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "geopandas", "xarray", "dask", "dask.array",
                 "rasterio.fill", "rasterio.warp", "rasterio.crs", "flox", "pooch"]:
        sys.modules[name] = mock.MagicMock(name=name)
    geom = mod("shapely.geometry", Polygon=Polygon, LineString=LineString, Point=Point)
    aff = mod("shapely.affinity", rotate=rotate)
    mod("shapely.ops")
    mod("shapely.wkt", loads=lambda s: None)
    mod("shapely", geometry=geom, affinity=aff, ops=sys.modules["shapely.ops"], wkt=sys.modules["shapely.wkt"])
    tr = mod("rasterio.transform", Affine=Affine, xy=None)
    ft = mod("rasterio.features", rasterize=rasterize)
    mod("rasterio", transform=tr, features=ft, fill=sys.modules["rasterio.fill"], warp=sys.modules["rasterio.warp"], crs=sys.modules["rasterio.crs"])
    exc = mod("pyproj.exceptions", CRSError=Exception)
    mod("pyproj", CRS=_CRS, Transformer=mock.MagicMock(), exceptions=exc)
    # a bare `pyorc` package so sub-modules import without running pyorc/__init__.py (which pulls in the whole API)
    pkg = types.ModuleType("pyorc")
    pkg.__path__ = [os.path.join(REF, "pyorc")]
    sys.modules["pyorc"] = pkg
    api = types.ModuleType("pyorc.api")
    api.__path__ = [os.path.join(REF, "pyorc", "api")]
    sys.modules["pyorc.api"] = api


def main():
    install_stubs()
    cv = importlib.import_module("pyorc.cv")
    cameraconfig = importlib.import_module("pyorc.api.cameraconfig")
    project = importlib.import_module("pyorc.project")

    # tests/conftest.py fixtures: gcps :112-123, lens_position :126-128, corners :147-159, camera_matrix :176-178,
    # cam_config :186-198 (window_size 25, resolution 0.01, crs 32735), dist_coeffs
    gcps = dict(src=[[1421, 1001], [1251, 460], [421, 432], [470, 607]],
                dst=[[642735.8076, 8304292.1190], [642737.5823, 8304295.593], [642732.7864, 8304298.4250], [642732.6705, 8304296.8580]],
                z_0=1182.2, h_ref=0.0)
    conftest = open(os.path.join(REF, "tests", "conftest.py")).read()
    assert "def dist_coeffs" in conftest
    ns = {}
    import re

    m = re.search(r"def dist_coeffs\(\):\n\s+return (.*)\n", conftest)
    dist_coeffs = eval(m.group(1), {"np": np})
    cc = cameraconfig.CameraConfig(
        height=1080, width=1920, gcps=gcps, lens_position=[642732.6705, 8304289.010, 1188.5], dist_coeffs=dist_coeffs,
        camera_matrix=np.array([[1550.0, 0.0, 960.0], [0.0, 1550.0, 540.0], [0.0, 0.0, 1.0]]),
        corners=[[500, 800], [400, 600], [1200, 550], [1350, 650]], window_size=25, resolution=0.01, crs=32735,
    )
    shape = cc.shape
    print("ortho shape", shape, "(tests/test_frames.py:36 expects (475, 371) at 0.01 m)")
    print("bbox", cc.bbox.exterior.coords[:4])
    # ---- video: frames 0..2 grayscale exactly as Video.get_frames_chunk does (api/video.py:442-451) ----
    fn = os.path.join(REF, "examples", "ngwerere", "ngwerere_20191103.mp4")
    cap = cv2.VideoCapture(fn)
    times, numbers, _ = cv.get_time_frames(cap, 0, 2, lazy=True, progress=False, method="bgr")
    cap.release()
    cap = cv2.VideoCapture(fn)
    cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
    imgs = []
    for n in numbers:
        ret, img = cv.get_frame(cap, rotation=None, ms=None, method="grayscale")
        assert ret
        imgs.append(img)
    cap.release()
    imgs = np.array(imgs)
    print("frames", imgs.shape, imgs.dtype, "frame[1][0,:4] =", imgs[1][0, :4], "(tests/test_video.py:49 pins [85 71 65 80])", "time ms", times)
    # ---- Frames.project(method='numpy', reducer='mean') (api/frames.py:234-265, project.py:160-230) ----
    res = cc.resolution
    y = np.flipud(np.linspace(res / 2, res * (shape[0] - 0.5), shape[0]))
    x = np.linspace(res / 2, res * (shape[1] - 0.5), shape[1])
    z = cc.get_z_a(0.0)
    idx_img, idx_ortho = cc.map_idx_img_ortho(x, y, z)
    src_idx, uidx, norm_idx = cc.map_mean_idx_img_ortho(x, y, z)
    proj = np.stack([project.img_to_ortho(im, x, y, idx_img, idx_ortho, src_idx, uidx, norm_idx) for im in imgs])
    proj = np.nan_to_num(proj, nan=0.0)  # .fillna(0.0)
    # xr.apply_ufunc(..., output_dtypes=[da.dtype], dask="parallelized") (project.py:205-227) hands the float64 group
    # means back as the frames' own dtype, uint8: the fractional part is truncated.  (Only the truncated frames
    # reproduce the pinned vectors - to 2e-8; the float64 ones are off by 1e-4.)
    proj = proj.astype(np.uint8)
    print("projected", proj.shape, proj.dtype, "mean", proj.mean(), "nonzero frac", (proj != 0).mean())
    np.savez_compressed(
        OUT, frames=proj, time_s=np.array(times) * 0.001, resolution=res,
        pinned_vx_timestep=np.array([0.10837663, 0.11250661, 0.11100861, 0.1231317]),
        pinned_vx_ensemble=np.array([0.10917795, 0.10898168, 0.11020568, 0.12450387]),
        source="pyorc @ be7d7c8 tests/test_frames.py:139-153; generated by tests/golden/make_ngwerere_golden.py",
    )
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    # ---- second fixture: the index maps themselves + the raw frame 0, so that the orthoprojection can be
    # checked against the reference's own img_to_ortho output.  Maps are delta-coded int32 (np.cumsum restores them); the
    # raw frame is cropped to the bounding box of every referenced source pixel (the rest is never read).
    allsrc = np.concatenate([idx_img, src_idx])
    r, c = allsrc // cc.width, allsrc % cc.width
    r0, r1, c0, c1 = int(r.min()), int(r.max()) + 1, int(c.min()), int(c.max()) + 1

    def delta(a):
        return np.diff(np.asarray(a, dtype=np.int64), prepend=0).astype(np.int32)

    ref0 = project.img_to_ortho(imgs[0], x, y, idx_img, idx_ortho, src_idx, uidx, norm_idx)   # float64 holding float32 means
    assert np.array_equal(ref0.astype(np.float32).astype(np.float64), ref0)
    assert np.array_equal(np.nan_to_num(ref0, nan=0.0).astype(np.uint8), proj[0])
    np.savez_compressed(
        OUT_MAPS, img_shape=np.array([cc.height, cc.width]), ortho_shape=np.array(shape), crop=np.array([r0, r1, c0, c1]),
        raw0_crop=imgs[0][r0:r1, c0:c1], d_idx_img=delta(idx_img), idx_ortho_bits=np.packbits(idx_ortho), n_ortho=idx_ortho.size,
        d_src_idx=delta(src_idx), d_uidx=delta(uidx), d_norm_idx=delta(norm_idx), ref0_f32=ref0.astype(np.float32),
        source="pyorc @ be7d7c8 CameraConfig.map_idx_img_ortho / map_mean_idx_img_ortho / project.img_to_ortho on "
               "examples/ngwerere frame 0; generated by tests/golden/make_ngwerere_golden.py",
    )
    print("wrote", OUT_MAPS, os.path.getsize(OUT_MAPS), "bytes; nearest", idx_img.size, "samples", src_idx.size, "groups", uidx.size)
    # ---- third fixture: the bottom-right corner of the raw grayscale frames (32 rows x 64 columns of each of the three
    # frames) - enough context for the pins of pyorc's filter tests on `frames_grayscale` (tests/test_frames.py:54-93:
    # the last four values of smooth(); time_diff / range are pinned on the same corner)
    np.savez_compressed(
        OUT_TAIL, gray_tail=imgs[:, -32:, -64:], frame_shape=np.array(imgs.shape[1:]),
        pinned_smooth_last4=np.array([158.125, 153.5, 151.375, 151.0]),
        pinned_edge_detect_proj_last4=np.array([-5.6953125, 4.0703125, 8.0625, 4.3125]),
        source="pyorc @ be7d7c8 tests/test_frames.py:86-103 (test_smooth on frames_grayscale, test_edge_detect on frames_proj); "
               "generated by tests/golden/make_ngwerere_golden.py",
    )
    print("wrote", OUT_TAIL, os.path.getsize(OUT_TAIL), "bytes")


if __name__ == "__main__":
    main()
