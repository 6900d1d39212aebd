"""TEST-ONLY harness that imports the REFERENCE's own Python (``/root/reference/pyorc``; this container only - the GPU box
has no reference tree and the reference cannot be pip-installed offline: its build backend ``flit_core`` is absent) with
minimal stand-ins for its heavy dependencies that are not installed here (xarray, dask, shapely, rasterio, pyproj,
geopandas, matplotlib, ffpiv, affine).  Only what the executed paths really touch is implemented: geometry primitives of the
projection path (rotate, bounds, LineString.length, Affine indexing, pixel-centre rasterize, ``rasterio.transform.xy``), WKT
of a polygon for the camera-configuration JSON round trip, and the xarray surface of ``pyorc_b200._xr``.  Everything numerical
(cv2.solvePnP, projectPoints, the index maps, the float32 group means, ``Frames.get_piv`` / ``get_ffpiv`` themselves) is the
reference's code, unmodified.

Used by tests/golden/make_ngwerere_golden.py (fixture generation) and tests/test_reference_dropin.py (the reference's
``Frames.get_piv`` body around ``pyorc_b200.frames.install()``).
"""
import importlib
import os
import re
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "pyorc"))


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

    def __str__(self):   # WKT with repr() precision, so that the JSON round trip of a camera configuration is exact
        return "POLYGON ((" + ", ".join(" ".join(repr(v) for v in c) for c in self.exterior.coords) + "))"


def wkt_loads(s):
    m = re.match(r"\s*POLYGON\s*(?:Z\s*)?\(\((.*)\)\)\s*$", s)
    if not m:
        raise ValueError(f"stand-in WKT reader handles POLYGON only, got {s[:40]!r}")
    return Polygon([[float(v) for v in pt.split()] for pt in m.group(1).split(",")])


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


def transform_xy(transform, rows, cols, offset="center"):
    """rasterio.transform.xy: map coordinates of pixel centres."""
    a, b, c, d, e, f = [float(v) for v in list(transform)[:6]]
    r, cc = np.asarray(rows, dtype=np.float64) + 0.5, np.asarray(cols, dtype=np.float64) + 0.5
    return a * cc + b * r + c, d * cc + e * r + f


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

    def __init__(self, *a, **k):
        pass

    @classmethod
    def from_user_input(cls, x):
        return cls()

    @classmethod
    def from_wkt(cls, x):
        return cls()

    def to_wkt(self, *a, **k):
        return "EPSG:32735"


def _warp_transform(src_crs, dst_crs, xs, ys):
    """Stand-in for rasterio.warp.transform: NOT a reprojection (pyproj is absent) - lon / lat coordinates of the result are
    placeholders here and no test reads them."""
    return list(np.asarray(xs, dtype=np.float64) * 0.0), list(np.asarray(ys, dtype=np.float64) * 0.0)


class _UncachedAccessor:
    """xarray.core.utils.UncachedAccessor"""

    def __init__(self, accessor):
        self._accessor = accessor

    def __get__(self, obj, cls):
        return self._accessor if obj is None else self._accessor(obj)


_STUBBED = []


def install_stubs(xarray_module=None, ffpiv_module=None):
    """Put the stand-ins into ``sys.modules`` (remembered, so that :func:`remove_stubs` can take them out again).
    ``xarray_module``: the module to serve as ``xarray`` (default: a MagicMock - enough for the projection path);
    ``ffpiv_module``: served as ``ffpiv`` (``window``, ``cross_corr``, ``u_v_displacement``)."""

    def put(name, m):
        if name not in sys.modules:
            _STUBBED.append(name)
        sys.modules[name] = m
        return m

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return put(name, m)

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "matplotlib.colors", "matplotlib.collections",
                 "matplotlib.ticker", "matplotlib.patheffects", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "geopandas",
                 "dask", "dask.array", "rasterio.fill", "rasterio.crs", "flox", "pooch", "affine", "pyproj.enums", "netCDF4", "typeguard"]:
        put(name, mock.MagicMock(name=name))
    if xarray_module is None:
        put("xarray", mock.MagicMock(name="xarray"))
    else:
        put("xarray", xarray_module)
        core = mod("xarray.core", utils=mod("xarray.core.utils", UncachedAccessor=_UncachedAccessor))
        xarray_module.core = core
    if ffpiv_module is not None:
        put("ffpiv", ffpiv_module)
    geom = mod("shapely.geometry", Polygon=Polygon, LineString=LineString, Point=Point)
    aff = mod("shapely.affinity", rotate=rotate)
    mod("shapely.ops")
    mod("shapely.wkt", loads=wkt_loads)
    mod("shapely", geometry=geom, affinity=aff, ops=sys.modules["shapely.ops"], wkt=sys.modules["shapely.wkt"])
    tr = mod("rasterio.transform", Affine=Affine, xy=transform_xy)
    ft = mod("rasterio.features", rasterize=rasterize)
    wp = mod("rasterio.warp", transform=_warp_transform)
    mod("rasterio", transform=tr, features=ft, fill=sys.modules["rasterio.fill"], warp=wp, crs=sys.modules["rasterio.crs"])
    exc = mod("pyproj.exceptions", CRSError=Exception)
    mod("pyproj", CRS=_CRS, Transformer=mock.MagicMock(), exceptions=exc, enums=sys.modules["pyproj.enums"])
    # a bare `pyorc` package so sub-modules import without running pyorc/__init__.py (which pulls in the whole API)
    pkg = types.ModuleType("pyorc")
    pkg.__path__ = [os.path.join(REF, "pyorc")]
    pkg.__version__ = "0.9.9"
    put("pyorc", pkg)
    api = types.ModuleType("pyorc.api")
    api.__path__ = [os.path.join(REF, "pyorc", "api")]
    put("pyorc.api", api)
    pkg.api = api


def remove_stubs():
    """Take the stand-ins (and every reference module imported on top of them) out of ``sys.modules`` again."""
    for name in list(sys.modules):
        if name == "pyorc" or name.startswith("pyorc.") or name in _STUBBED or any(name.startswith(s + ".") for s in _STUBBED):
            del sys.modules[name]
    _STUBBED.clear()


def ngwerere_camera_config(cameraconfig):
    """The camera configuration of the reference's test suite: tests/conftest.py fixtures gcps :112-123, lens_position
    :126-128, corners :147-159, camera_matrix :176-178, cam_config :186-198 (window_size 25, resolution 0.01, crs 32735)."""
    gcps = dict(src=[[1421, 1001], [1251, 460], [421, 432], [470, 607]],
                dst=[[642735.8076, 8304292.1190], [642737.5823, 8304295.593], [642732.7864, 8304298.4250], [642732.6705, 8304296.8580]],
                z_0=1182.2, h_ref=0.0)
    conftest = open(os.path.join(REF, "tests", "conftest.py")).read()
    m = re.search(r"def dist_coeffs\(\):\n\s+return (.*)\n", conftest)
    dist_coeffs = eval(m.group(1), {"np": np})
    return cameraconfig.CameraConfig(
        height=1080, width=1920, gcps=gcps, lens_position=[642732.6705, 8304289.010, 1188.5], dist_coeffs=dist_coeffs,
        camera_matrix=np.array([[1550.0, 0.0, 960.0], [0.0, 1550.0, 540.0], [0.0, 0.0, 1.0]]),
        corners=[[500, 800], [400, 600], [1200, 550], [1350, 650]], window_size=25, resolution=0.01, crs=32735,
    )


def import_ref(name):
    return importlib.import_module(name)
