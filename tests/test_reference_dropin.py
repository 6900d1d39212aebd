"""CPU, this container only: the REFERENCE's own ``Frames.get_piv`` body (pyorc/api/frames.py:114-197, imported from
/root/reference through tests/ref_harness.py) run around ``pyorc_b200.frames.install()`` - pyorc's ``test_get_piv``
(tests/test_frames.py:139-153) with ``engine="b200"``, both modes, on the projected Ngwerere frames and the camera
configuration of its conftest, against the pinned ``v_x`` vectors.

What is real here: CameraConfig, ``Frames.get_piv`` / ``get_piv_coords`` / ``ORCBase.add_xy_coords`` /
``Velocimetry.set_encoding``, ``pyorc.velocimetry.ffpiv.get_ffpiv`` (the ``engine="numba"`` arm), and on our side
``install()`` + ``get_b2piv``.  What is stood in: xarray (``pyorc_b200._xr``), the ``ffpiv`` package (the float64 oracle's
``cross_corr`` / ``u_v_displacement`` / ``window``) and - there being no GPU in this container - the CUDA engine (the
oracle-backed ``FakeEngine`` of tests/test_host_logic.py).  The same pins are checked with the REAL engine through
``pyorc_b200.frames.get_piv`` in tests/test_golden.py (-m gpu) and through ``install()`` on a pyorc-shaped module in
tests/test_gpu_parity.py::test_install_dispatch_real_engine.  The reference cannot travel to the GPU box (not installable
offline: flit_core is missing), so this test skips there.
"""
import json
import os
import threading
import types

import numpy as np
import pytest

import ref_harness
from oracle import ffpiv_oracle as O
from pyorc_b200 import _xr, frames as b2frames, velocimetry, window as b2window  # noqa: F401  (imported BEFORE the stand-ins go in)
from test_host_logic import FakeEngine

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference is not present (GPU box): the reference's Python cannot travel")

PIN_TIMESTEP = [0.10837663, 0.11250661, 0.11100861, 0.1231317]     # tests/test_frames.py:143
PIN_ENSEMBLE = [0.10917795, 0.10898168, 0.11020568, 0.12450387]    # tests/test_frames.py:142


@pytest.fixture(scope="module")
def ref():
    O.CLIP_NORMALIZED = False
    win = types.SimpleNamespace(round_to_even=O.round_to_even, get_rect_coordinates=lambda dim_size, window_size, search_area_size, overlap:
                                O.get_rect_coordinates(dim_size, window_size, overlap, search_area_size),
                                required_memory=O.required_memory, available_memory=O.available_memory)
    ffpiv_mod = types.ModuleType("ffpiv")
    ffpiv_mod.window, ffpiv_mod.cross_corr, ffpiv_mod.u_v_displacement = win, O.cross_corr, O.u_v_displacement
    ref_harness.install_stubs(xarray_module=_xr, ffpiv_module=ffpiv_mod)
    import sys

    cameraconfig = ref_harness.import_ref("pyorc.api.cameraconfig")
    sys.modules["pyorc"].get_camera_config = cameraconfig.get_camera_config
    ns = types.SimpleNamespace(
        cameraconfig=cameraconfig,
        frames=ref_harness.import_ref("pyorc.api.frames"),
        velocimetry=ref_harness.import_ref("pyorc.api.velocimetry"),
        ffpiv=ref_harness.import_ref("pyorc.velocimetry.ffpiv"),
        const=ref_harness.import_ref("pyorc.const"),
    )
    yield ns
    b2frames.uninstall()
    ref_harness.remove_stubs()


@pytest.fixture(scope="module")
def frames_proj(ref):
    """The `frames_proj` fixture of pyorc's conftest (:379-380) from the committed golden file (made by the reference's own decode
    and projection code, tests/golden/make_ngwerere_golden.py)."""
    g = np.load(os.path.join(ref_harness.GOLDEN, "ngwerere_proj.npz"))
    cc = ref_harness.ngwerere_camera_config(ref.cameraconfig)
    fr, res = g["frames"], float(g["resolution"])
    assert fr.shape[1:] == tuple(cc.shape)
    y = np.flipud(np.linspace(res / 2, res * (fr.shape[1] - 0.5), fr.shape[1]))
    x = np.linspace(res / 2, res * (fr.shape[2] - 0.5), fr.shape[2])
    attrs = {"camera_config": cc.to_json(), "camera_shape": str([cc.height, cc.width]), "h_a": json.dumps(0.0)}
    return _xr.DataArray(fr, ("time", "y", "x"), {"time": g["time_s"], "y": y, "x": x}, attrs=attrs)


@pytest.fixture()
def fake(monkeypatch):
    fe = FakeEngine()
    monkeypatch.setattr(velocimetry, "get_engine", lambda device=0, slot=0: fe)
    monkeypatch.setattr(b2window, "available_memory", lambda device=None, **kw: 64e9)
    O.CLIP_NORMALIZED = False
    return fe


@pytest.mark.parametrize("ensemble_corr,result", [(True, PIN_ENSEMBLE), (False, PIN_TIMESTEP)])
def test_get_piv_through_the_reference_body(ref, frames_proj, fake, ensemble_corr, result):
    kw = dict(window_size=10, ensemble_corr=ensemble_corr, s2n_min=0, corr_min=0, count_min=0)
    # the reference, end to end (its own get_ffpiv on the stood-in ffpiv package = the oracle): pyorc's test as it stands
    piv_ref = frames_proj.frames.get_piv(engine="numba", **kw)
    assert np.allclose(piv_ref.mean(dim="time", keep_attrs=True)["v_x"].values.flatten()[-4:], result, equal_nan=True, atol=2e-8)
    b2frames.uninstall()
    with pytest.raises(ValueError, match="Selected PIV engine b200 does not exist."):
        frames_proj.frames.get_piv(engine="b200", **kw)                      # not registered (yet / any more)
    assert b2frames.install() and b2frames.install()                         # idempotent
    n_before = len(fake.calls)
    piv = frames_proj.frames.get_piv(engine="b200", **kw)                    # same body, engine call redirected
    assert len(fake.calls) > n_before
    piv_mean = piv.mean(dim="time", keep_attrs=True)
    assert np.allclose(piv_mean["v_x"].values.flatten()[-4:], result, equal_nan=True, atol=2e-6)
    # the whole Dataset, against the reference arm: variables, coordinates (1-D and the 2-D mesh the body adds), attrs, encoding
    for k in ("v_x", "v_y", "corr", "s2n"):
        a, b = piv_ref[k].values, piv[k].values
        assert a.shape == b.shape and a.dtype == b.dtype == np.float32
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.allclose(a, b, rtol=2e-6, atol=1e-7, equal_nan=True), k
        assert piv[k].encoding == ref.const.ENCODING_PARAMS
    for k in ("time", "y", "x", "xp", "yp", "xs", "ys", "lon", "lat"):
        assert np.array_equal(piv_ref[k].values, piv[k].values), k
        assert piv_ref[k].attrs == piv[k].attrs
    assert piv.attrs == piv_ref.attrs and json.loads(piv.attrs["camera_config"])["window_size"] == 10
    # the registered wrapper leaves other engines alone
    again = frames_proj.frames.get_piv(engine="numba", **kw)
    assert np.array_equal(again["v_x"].values, piv_ref["v_x"].values, equal_nan=True)
    with pytest.raises(ValueError, match="Selected PIV engine openpiv does not exist."):
        frames_proj.frames.get_piv(engine="openpiv", **kw)


def test_b200_kwargs_reach_the_binding(ref, frames_proj, fake, monkeypatch):
    b2frames.install()
    seen = {}
    real = velocimetry.get_b2piv

    def spy(frames, y, x, dt, *a, **kw):
        seen.update(kw)
        seen["dt_type"] = type(dt).__name__
        return real(frames, y, x, dt, *a, **kw)

    monkeypatch.setattr(b2frames, "get_b2piv", spy)
    frames_proj.frames.get_piv(window_size=10, engine="b200", devices=[0], chunksize=2, signal_threshold=0.5)
    assert seen["engine"] == "b200" and seen["devices"] == [0] and seen["chunksize"] == 2 and seen["signal_threshold"] == 0.5
    assert seen["window_size"] == (10, 10) and seen["overlap"] == (5, 5) and seen["res_x"] == 0.01
    assert seen["dt_type"] == "DataArray"          # the reference hands `time.diff` over as a DataArray (frames.py:157)
    # the reference's own arm with an explicit chunksize hits its NameError (ffpiv.py:140); ours does not
    with pytest.raises(NameError):
        frames_proj.frames.get_piv(window_size=10, engine="numba", chunksize=2)


def test_concurrent_calls_with_different_engines_do_not_cross(ref, frames_proj, monkeypatch):
    """pyorc under dask's threaded scheduler: a `numba` call in one thread while a `b200` call is in flight in another must
    reach the reference's get_ffpiv, and vice versa (context variable, no module global swapped per call)."""
    b2frames.install()
    inside, release = threading.Event(), threading.Event()

    class SlowFake(FakeEngine):
        def pairs(self, *a, **k):
            inside.set()
            assert release.wait(30)
            return super().pairs(*a, **k)

    fe = SlowFake()
    monkeypatch.setattr(velocimetry, "get_engine", lambda device=0, slot=0: fe)
    monkeypatch.setattr(b2window, "available_memory", lambda device=None, **kw: 64e9)
    out = {}
    t = threading.Thread(target=lambda: out.setdefault("b200", frames_proj.frames.get_piv(window_size=10, engine="b200")))
    t.start()
    assert inside.wait(30)                       # the b200 call is now inside the engine
    calls_before = len(fe.calls)
    out["numba"] = frames_proj.frames.get_piv(window_size=10, engine="numba")     # must NOT be routed to the b200 engine
    assert len(fe.calls) == calls_before
    release.set()
    t.join(60)
    assert not t.is_alive() and len(fe.calls) >= 1
    assert np.allclose(out["b200"]["v_x"].values, out["numba"]["v_x"].values, rtol=2e-6, atol=1e-7, equal_nan=True)
