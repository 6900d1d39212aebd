"""CPU: host-side logic of the binding (chunking with 1-frame halo, unit conversion, Dataset layout, errors) with the
engine replaced by an oracle-backed fake - the arithmetic itself is covered by the -m gpu parity tests."""
import warnings

import numpy as np
import pytest

from oracle import ffpiv_oracle as O
from pyorc_b200 import _xr, frames as b2frames, synth, velocimetry, window


class FakeEngine:
    """Test double with the Engine interface, computing with the oracle (never used by the product)."""

    def __init__(self, device=0, log=None):
        self.device = device
        self.calls = [] if log is None else log

    def pairs(self, imgs, ws, ov, signal_threshold=None, stream=None):
        self.calls.append(np.asarray(imgs).shape[0])
        nr, nc = O.get_array_shape(np.asarray(imgs).shape[-2:], ws, ov)
        u, v, c, s = O.uv_timestep(np.asarray(imgs), nc, nr, ws, ov, signal_threshold=signal_threshold)
        return u.astype(np.float32), v.astype(np.float32), c, s

    def pairs_two_pass(self, imgs, coarse, fine, mode="offset"):
        from oracle import multipass_oracle as MP

        self.calls.append(("two_pass", np.asarray(imgs).shape[0], coarse, fine) + ((mode,) if mode != "offset" else ()))
        fn = MP.two_pass if mode == "offset" else MP.two_pass_deform
        u, v, c, s, _, _ = fn(np.asarray(imgs), coarse, fine)
        return u.astype(np.float32), v.astype(np.float32), c.astype(np.float32), s.astype(np.float32)

    def ens_begin(self, dim_size, ws, ov, dtype):
        nr, nc = O.get_array_shape(dim_size, ws, ov)
        self._ens = None
        self._shape = (nr, nc)
        return nr, nc

    def ens_add(self, imgs, ws, ov, corr_min=0.2, s2n_min=3.0, signal_threshold=None, stream=None):
        self.calls.append(("ens", np.asarray(imgs).shape[0]))
        if self._ens is None:
            self._ens = O.Ensemble(*self._shape, ws, ov, corr_min, s2n_min, 0.0, signal_threshold)
        self._ens.add_chunk(np.asarray(imgs))
        return self._ens.corr_chunks[-1].copy(), self._ens.s2n_chunks[-1].copy()

    def ens_finish(self, min_count):
        e = self._ens
        cnt = np.asarray(e.corr_count).reshape(-1).astype(np.float32)
        corr_sum = np.array(e.corr_sum, dtype=np.float32, copy=True)
        with np.errstate(all="ignore"):
            corr_sum[:, cnt < min_count] = np.nan
            mean = corr_sum / cnt[None, :, None, None]
        u, v = O.u_v_displacement(mean, *self._shape)
        return u.reshape(-1).astype(np.float32), v.reshape(-1).astype(np.float32), cnt


@pytest.fixture()
def fake(monkeypatch):
    fe = FakeEngine()
    fakes = {0: fe}

    def get_engine(device=0, slot=0):   # one fake per (device, slot), sharing device 0's call log
        key = device if slot == 0 else (device, slot)
        if key not in fakes:
            fakes[key] = FakeEngine(device, fe.calls)
        return fakes[key]

    def merge(engines, min_count):   # numpy stand-in for engine.merge_ensembles (peer adds + peak fit on the first device)
        e0 = engines[0]
        for e in engines[1:]:
            e0._ens.corr_sum = e0._ens.corr_sum + e._ens.corr_sum
            e0._ens.corr_count = e0._ens.corr_count + e._ens.corr_count
        return e0.ens_finish(min_count)

    monkeypatch.setattr(velocimetry, "get_engine", get_engine)
    monkeypatch.setattr(velocimetry, "merge_ensembles", merge)
    monkeypatch.setattr(window, "available_memory", lambda device=None, **kw: 64e9)
    O.CLIP_NORMALIZED = True
    return fe


def make_frames(n=7, H=100, W=140, fps=30.0):
    imgs = synth.particle_frames(n, H, W, dtype=np.uint8)
    t = np.arange(n) / fps
    res = 0.01
    y = np.flipud(np.linspace(res / 2, res * (H - 0.5), H))
    x = np.linspace(res / 2, res * (W - 0.5), W)
    return _xr.DataArray(imgs, ("time", "y", "x"), {"time": t, "y": y, "x": x}, attrs={"camera_config": "{}"}), res


def test_get_piv_matches_reference_semantics(fake):
    da, res = make_frames()
    ds = b2frames.get_piv(da, window_size=32, engine="b200", resolution=res)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    assert ds["v_x"].values.shape == (6, nr, nc) and ds["v_x"].values.dtype == np.float32
    u, v, c, s = O.uv_timestep(da.values, nc, nr, (32, 32), (16, 16))
    dt = 1 / 30.0
    assert np.allclose(ds["v_x"].values, (u * res / dt).astype(np.float32), equal_nan=True, rtol=1e-6)
    assert np.allclose(ds["v_y"].values, (v * res / dt).astype(np.float32), equal_nan=True, rtol=1e-6)
    assert np.array_equal(ds["corr"].values, c) and np.array_equal(ds["s2n"].values, s, equal_nan=True)
    # coordinates: integer window centres index the frame axes (helpers.get_axes)
    cols, rows = window.get_rect_coordinates((100, 140), (32, 32), (16, 16))
    assert np.array_equal(ds.coords["x"], da.coords["x"][cols]) and np.array_equal(ds.coords["y"], da.coords["y"][rows])
    assert np.array_equal(ds.coords["time"], da.coords["time"][1:])
    assert ds.attrs == da.attrs


def test_default_overlap_and_rounding(fake):
    da, res = make_frames(n=3)
    ds = b2frames.get_piv(da, window_size=31, engine="b200", resolution=res)   # -> window 32, overlap int(round(31)/2) = 15
    nr, nc = O.get_array_shape((100, 140), (32, 32), (15, 15))
    assert ds["v_x"].values.shape == (2, nr, nc)


def test_unknown_engine_raises_like_reference(fake):
    da, res = make_frames(n=3)
    with pytest.raises(ValueError, match="Selected PIV engine numba does not exist."):
        b2frames.get_piv(da, window_size=32, engine="numba", resolution=res)


def test_chunks_share_one_frame_halo(fake):
    da, res = make_frames(n=11)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    y, x = np.arange(nr), np.arange(nc)
    whole = velocimetry.get_b2piv(da, y, x, np.full(10, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, chunksize=100)
    fake.calls.clear()
    parts = velocimetry.get_b2piv(da, y, x, np.full(10, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, chunksize=4)
    assert fake.calls == [4, 5, 4]            # frames [0:4], [3:8], [7:11]  (ffpiv.py:140)
    for k in ("v_x", "v_y", "corr", "s2n"):
        assert np.array_equal(whole[k].values, parts[k].values, equal_nan=True)
    assert np.array_equal(parts.coords["time"], da.coords["time"][1:])


def test_chunksize_errors_and_warning(fake, monkeypatch):
    da, res = make_frames(n=6)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    y, x = np.arange(nr), np.arange(nc)
    with pytest.raises(OverflowError):
        velocimetry.get_b2piv(da, y, x, np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, chunksize=1)
    monkeypatch.setattr(window, "available_memory", lambda device=None, **kw: 1e5)   # tiny "device" -> chunksize floor 5 + warning
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ds = velocimetry.get_b2piv(da, y, x, np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res)
    assert any("Memory availability is poor" in str(m.message) for m in w)
    assert ds["v_x"].values.shape[0] == 5
    with pytest.raises(ValueError):
        velocimetry.get_b2piv(da, y[:-1], x, np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res)
    with pytest.raises(NotImplementedError):
        velocimetry.get_b2piv(da, y, x, np.full(5, 1 / 30), (32, 32), (16, 16), (64, 64), res, res)


def test_ensemble_mode_plumbing(fake):
    da, res = make_frames(n=9)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    y, x = np.arange(nr), np.arange(nc)
    ds = velocimetry.get_b2piv(da, y, x, np.full(8, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, chunksize=5,
                               ensemble_corr=True, corr_min=0.0, s2n_min=0.0, count_min=0.0)
    ens = O.Ensemble(nr, nc, (32, 32), (16, 16), 0.0, 0.0, 0.0)
    ens.add_chunk(da.values[0:5])
    ens.add_chunk(da.values[4:9])
    u, v, cm, sn = ens.finalize()
    assert ds["v_x"].values.shape == (1, nr, nc)
    assert np.allclose(ds["v_x"].values, (u * res * 30).astype(np.float32), equal_nan=True, rtol=1e-5)
    assert np.allclose(ds["corr"].values, cm, equal_nan=True) and np.allclose(ds["s2n"].values, sn, equal_nan=True)
    assert len(ds.coords["time"]) == 1


def test_window_memory_model():
    assert window.required_memory(101, (1080, 1920), (64, 64), (32, 32)) > 101 * 1080 * 1920
    with pytest.raises(ValueError):
        window.get_axis_shape(100, 16, 16)
    with pytest.raises(NotImplementedError):
        window.get_rect_coordinates((100, 100), (32, 32), (16, 16), search_area_size=(64, 64))


def test_pinned_result_pool_recycles_blocks_only_when_every_view_is_gone():
    """Engine.pairs returns four views of one page-locked block; the block goes back to the pool when all are collected."""
    import ctypes
    import gc

    from pyorc_b200.engine import _PinnedPool

    class FakeLib:
        def __init__(self):
            self.live, self.allocs = {}, 0

        def b2piv_host_alloc(self, size):
            b = ctypes.create_string_buffer(size)
            self.live[ctypes.addressof(b)] = b
            self.allocs += 1
            return ctypes.addressof(b)

        def b2piv_host_free(self, p):
            del self.live[p]

    lib = FakeLib()
    pool = _PinnedPool(lib)
    a = pool.empty((4, 10, 3, 5))
    a[...] = 1.5
    view = a[2]
    del a
    gc.collect()
    assert lib.allocs == 1 and not pool._free          # a view is still alive
    b = pool.empty((4, 10, 3, 5))
    assert lib.allocs == 2                              # so a second call must not reuse the block
    assert float(view[0, 0, 0]) == 1.5
    del view
    gc.collect()
    c = pool.empty((4, 10, 3, 5))
    assert lib.allocs == 2                              # recycled
    del b, c
    gc.collect()
    pool.close()
    assert not lib.live
    late = _PinnedPool(lib)
    d = late.empty((8,))
    late.close()
    del d
    gc.collect()
    assert not lib.live                                 # released after close: freed, not pooled


def test_coarse_pass_routes_chunks_through_the_two_pass_engine(fake):
    """get_b2piv(coarse_pass=...) - the two-pass scheme of BASELINE configs[2] behind the reference-shaped binding: chunks with
    a 1-frame halo, unit conversion and Dataset layout as in the single pass; refused in ensemble mode."""
    O.CLIP_NORMALIZED = False
    da, res = make_frames(n=6, H=150, W=200)
    ws, ov, coarse = (32, 32), (24, 24), ((64, 64), (48, 48))
    nr, nc = O.get_array_shape((150, 200), ws, ov)
    y, x = np.arange(nr), np.arange(nc)
    ds = velocimetry.get_b2piv(da, y, x, np.full(5, 1 / 30), ws, ov, ws, res, res, chunksize=4, coarse_pass=coarse)
    assert [c[:2] for c in fake.calls] == [("two_pass", 4), ("two_pass", 3)] and fake.calls[0][2:] == (coarse, (ws, ov))
    assert ds["v_x"].values.shape == (5, nr, nc) and ds["v_x"].values.dtype == np.float32
    from oracle import multipass_oracle as MP

    u, v, c, s, _, _ = MP.two_pass(da.values, coarse, (ws, ov))
    assert np.allclose(ds["v_x"].values, (u * res * 30).astype(np.float32), equal_nan=True, rtol=1e-6)
    assert np.allclose(ds["corr"].values, c, equal_nan=True)
    with pytest.raises(NotImplementedError):
        velocimetry.get_b2piv(da, y, x, np.full(5, 1 / 30), ws, ov, ws, res, res, coarse_pass=coarse, ensemble_corr=True)
    with pytest.raises(ValueError):
        velocimetry.get_b2piv(da, y, x, np.full(5, 1 / 30), ws, ov, ws, res, res, coarse_pass=((16, 16), (8, 8)))
    O.CLIP_NORMALIZED = True


# ---- devices=[...]: one call, several GPUs ------------------------------------------------------------------
def test_work_items_cover_every_pair_once_with_halo():
    for bounds in ([(0, 11)], [(0, 4), (3, 8), (7, 11)], [(0, 2)], [(0, 3), (2, 4)]):
        for n_dev in (1, 2, 3, 8):
            items = velocimetry._work_items(bounds, n_dev)
            pairs = [p for _, a, b in items for p in range(a, b - 1)]
            assert pairs == [p for a, b in bounds for p in range(a, b - 1)]
            assert all(b - a >= 2 for _, a, b in items)
            if n_dev == 1:
                assert [(a, b) for _, a, b in items] == list(bounds)
            for c, (a, b) in enumerate(bounds):
                assert len([1 for cc, _, _ in items if cc == c]) == min(n_dev, b - a - 1)


@pytest.mark.parametrize("devices", [[0, 1], [0, 1, 2], list(range(8)), [0, 0, 1]])
def test_devices_sharding_equals_single_device(fake, devices):
    da, res = make_frames(n=11)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    y, x = np.arange(nr), np.arange(nc)
    args = (da, y, x, np.full(10, 1 / 30), (32, 32), (16, 16), (32, 32), res, res)
    one = velocimetry.get_b2piv(*args, chunksize=6)
    fake.calls.clear()
    many = velocimetry.get_b2piv(*args, chunksize=6, devices=devices)
    assert sum(n - 1 for n in fake.calls) == 10 and len(fake.calls) == min(len(devices), 5) * 2
    for k in ("v_x", "v_y", "corr", "s2n"):
        assert np.array_equal(one[k].values, many[k].values, equal_nan=True)
    assert np.array_equal(one.coords["time"], many.coords["time"])


def test_devices_sharding_ensemble_equals_single_device(fake):
    da, res = make_frames(n=9)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    y, x = np.arange(nr), np.arange(nc)
    args = (da, y, x, np.full(8, 1 / 30), (32, 32), (16, 16), (32, 32), res, res)
    kw = dict(ensemble_corr=True, corr_min=0.1, s2n_min=1.0, count_min=0.2, chunksize=5)
    one = velocimetry.get_b2piv(*args, **kw)
    many = velocimetry.get_b2piv(*args, **kw, devices=[0, 1, 2])
    assert np.isfinite(one["v_x"].values).sum() > 0
    for k in ("corr", "s2n"):
        assert np.array_equal(one[k].values, many[k].values, equal_nan=True)
    for k in ("v_x", "v_y"):   # plane sums are added in a different order: float32 rounding of the sums only
        assert np.array_equal(np.isnan(one[k].values), np.isnan(many[k].values))
        assert np.nanmax(np.abs(one[k].values - many[k].values)) < 1e-4
    assert np.array_equal(one.coords["time"], many.coords["time"])


def test_device_thread_errors_propagate(fake):
    da, res = make_frames(n=6)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    with pytest.raises(ValueError, match="do not match"):
        velocimetry.get_b2piv(da, np.arange(nr + 1), np.arange(nc), np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, devices=[0, 1])

    def boom(*a, **k):
        raise RuntimeError("device lost")

    velocimetry.get_engine(1).pairs = boom
    with pytest.raises(RuntimeError, match="device lost"):
        velocimetry.get_b2piv(da, np.arange(nr), np.arange(nc), np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, devices=[0, 1])


def test_get_piv_sets_encoding_and_keeps_attrs(fake):
    da, res = make_frames(n=3)
    ds = b2frames.get_piv(da, window_size=32, engine="b200", resolution=res)
    for k in ("v_x", "v_y", "corr", "s2n"):
        assert ds[k].encoding == {"zlib": True, "dtype": "int16", "scale_factor": 0.01, "_FillValue": -9999}   # pyorc/const.py:80-83
    assert ds.attrs == da.attrs

    class Cfg:   # duck-typed camera configuration: window size / resolution defaults, JSON like CameraConfig.to_json
        window_size, resolution = 25, res

        def to_json(self):
            return '{"window_size": %d}' % self.window_size

    ds = b2frames.get_piv(da, engine="b200", camera_config=Cfg())
    nr, nc = O.get_array_shape((100, 140), (26, 26), (12, 12))     # 25 -> 26 (round_to_even), overlap int(round(25) / 2)
    assert ds["v_x"].values.shape == (2, nr, nc) and ds.attrs["camera_config"] == '{"window_size": 25}'
    ds = b2frames.get_piv(da, window_size=32, engine="b200", camera_config=Cfg())
    assert ds.attrs["camera_config"] == '{"window_size": 32}'       # frames.py:194-195: the window size actually used


def test_metrics_sidecar_one_json_line_per_call(fake, tmp_path, monkeypatch):
    import json

    path = tmp_path / "b2piv_metrics.jsonl"
    monkeypatch.setenv("B2PIV_METRICS", str(path))
    da, res = make_frames(n=6)
    nr, nc = O.get_array_shape((100, 140), (32, 32), (16, 16))
    velocimetry.get_b2piv(da, np.arange(nr), np.arange(nc), np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, chunksize=4)
    velocimetry.get_b2piv(da, np.arange(nr), np.arange(nc), np.full(5, 1 / 30), (32, 32), (16, 16), (32, 32), res, res, ensemble_corr=True)
    recs = [json.loads(line) for line in open(path)]
    assert [r["mode"] for r in recs] == ["per-time-step", "ensemble"]
    assert recs[0]["windows"] == 5 * nr * nc and recs[0]["chunks"] == 2 and recs[0]["alg_bytes"] == 5 * nr * nc * (2 * 16 * 16 + 16)
    assert recs[0]["windows_per_s"] > 0 and recs[0]["devices"] == [0]


def test_result_block_of_engine_fields():
    """parallel._result_block: the four fields Engine.pairs returns are views of one [4, pairs, rows, cols] block - the push gather
    forwards that block without a copy; anything else is stacked."""
    import torch

    from pyorc_b200 import parallel

    block = torch.arange(4 * 3 * 5 * 6, dtype=torch.float32).reshape(4, 3, 5, 6)
    views = (block[0], block[1], block[2], block[3])
    assert parallel._result_block(views).data_ptr() == block.data_ptr()
    assert parallel._result_block(block) is block
    loose = tuple(v.clone() for v in views)
    got = parallel._result_block(loose)
    assert got.data_ptr() != block.data_ptr() and torch.equal(got, block)
    partial = (block[0, 1:], block[1, 1:], block[2, 1:], block[3, 1:])     # views, but not of the whole block: must be copied
    assert torch.equal(parallel._result_block(partial), block[:, 1:])


@pytest.mark.parametrize(
    "rows,row_bytes,dpitch,slice_rows,parts,ring,threads,nt,delay_us",
    [(4370, 203, 208, 1291, 8, 4, 8, 0, 20),      # the GPU test's 190 x 203 frames: one group, re-pitched rows
     (3000, 1920, 1920, 136, 8, 4, 8, 0, 5),      # 1080p rows, 256 KB slices, plain stores
     (3000, 1920, 1920, 136, 8, 2, 3, 1, 5),      # fewer workers than slices per group, non-temporal stores, the smallest ring
     (777, 333, 336, 7, 5, 3, 2, 0, 0),           # ragged: last group and last slice short
     (50, 100, 112, 100, 4, 2, 4, 1, 0),          # less than one slice
     (3, 10, 16, 1, 1, 2, 1, 0, 0)],
)
def test_stager_delivers_every_row_once(rows, row_bytes, dpitch, slice_rows, parts, ring, threads, nt, delay_us):
    """csrc/stager.h (the copy threads and ring behind `b2piv_pairs_host` on pageable numpy memory) with a thread playing the
    copy engine: every row arrives once at its pitched place, groups are issued in order, a ring slot is never rewritten under
    a copy in flight, the chunk hook sees growing row counts - and an error from the issue callback ends the call."""
    import ctypes

    import __graft_entry__ as g

    lib = ctypes.CDLL(g.build_stager_emulator())
    f = lib.stager_selftest
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_longlong] * 4 + [ctypes.c_int] * 4 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    assert f(rows, row_bytes, dpitch, slice_rows, parts, ring, threads, nt, -1, delay_us, 3) == 0
    n_groups = -(-rows // (slice_rows * parts))
    assert f(rows, row_bytes, dpitch, slice_rows, parts, ring, threads, nt, n_groups // 2, delay_us, 2) == 0


# ---- work partition of the row-per-thread kernels (csrc/work_partition.h, compiled for the host) ----------------------------
def _partition_lib():
    import ctypes

    import __graft_entry__ as g

    lib = ctypes.CDLL(g.build_stager_emulator())
    lib.emul_pick_run_len.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong]
    lib.emul_partition_units.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong,
                                         ctypes.POINTER(ctypes.c_longlong)]
    return lib


def _partition(lib, n_wp, n_pairs, n_parts):
    import ctypes

    cap = 3 * (n_wp + n_parts + 2) * max(n_parts, 1) * 2
    out = np.full(cap, -9, np.int32)
    cost = ctypes.c_longlong()
    units = lib.emul_partition_units(n_wp, n_pairs, n_parts, out.ctypes.data, cap, ctypes.byref(cost))
    assert units % n_parts == 0 and 3 * units <= cap
    return out[: 3 * units].reshape(units // n_parts, n_parts, 3), int(cost.value)


@pytest.mark.parametrize(
    "n_wp,n_pairs,n_parts",
    [(944, 100, 888),      # BASELINE configs[1]: 1888 windows = 944 window pairs, 100 frame pairs, six groups on each of 148 SMs
     (944, 100, 592),      # ... four groups per SM (shared-memory variant)
     (3927, 125, 888),     # configs[3]: a 125-pair chunk of 4K
     (7, 3, 5), (1, 40, 8), (13, 1, 4), (5, 2, 10), (3, 3, 1), (2, 7, 14)],
)
def test_unit_table_covers_every_item_once_and_evenly(n_wp, n_pairs, n_parts):
    """Every (window pair, frame pair) belongs to exactly one segment, the parts are contiguous in window-pair-major order and
    hold total / n_parts items up to rounding, a segment never crosses a window pair, padding entries are (-1, 0, -1), and the
    reported cost is the longest part's frame pairs plus one transform per segment (what the engine compares with the wave
    schedule of pick_run_len)."""
    lib = _partition_lib()
    tab, cost = _partition(lib, n_wp, n_pairs, n_parts)
    total = n_wp * n_pairs
    seen = np.zeros((n_wp, n_pairs), np.int32)
    worst = 0
    pos = 0
    for part in range(n_parts):
        items, c = 0, 0
        padded = False
        for rd in range(tab.shape[0]):
            wp, f0, f1 = (int(x) for x in tab[rd, part])
            if wp < 0:
                assert (wp, f0, f1) == (-1, 0, -1)
                padded = True
                continue
            assert not padded                                   # a part's segments come first, then only padding
            assert 0 <= wp < n_wp and 0 <= f0 < f1 <= n_pairs
            assert wp * n_pairs + f0 == pos                     # contiguous: the parts tile the window-pair-major list in order
            seen[wp, f0:f1] += 1
            pos += f1 - f0
            items += f1 - f0
            c += f1 - f0 + 1
        assert items in (total // n_parts, -(-total // n_parts))
        worst = max(worst, c)
    assert pos == total and (seen == 1).all()
    assert cost == worst


def test_run_length_minimises_the_wave_cost():
    """pick_run_len against a brute-force search of its own cost model; the 1080p example of DESIGN.md 4.1b (944 window pairs,
    100 frame pairs, 592 resident groups: 5 chunks of 20 pairs, 7.97 waves, 168 frame times)."""
    lib = _partition_lib()

    def cost(n_pairs, n_wp, resident, run):
        chunks = -(-n_pairs // run)
        return -(-(n_wp * chunks) // resident) * (run + 1)

    assert lib.emul_pick_run_len(100, 944, 592) == 20
    assert cost(100, 944, 592, 20) == 168
    rng = np.random.default_rng(3)
    for _ in range(200):
        n_pairs, n_wp, resident = int(rng.integers(1, 300)), int(rng.integers(1, 5000)), int(rng.integers(1, 1200))
        run = lib.emul_pick_run_len(n_pairs, n_wp, resident)
        assert 1 <= run <= n_pairs
        reachable = {-(-n_pairs // c) for c in range(1, min(n_pairs, 64) + 1)}      # the run lengths the search visits
        assert run in reachable
        assert cost(n_pairs, n_wp, resident, run) == min(cost(n_pairs, n_wp, resident, r) for r in reachable)


def test_load_frame_chunk_retries_with_one_frame_less():
    """pyorc/velocimetry/ffpiv.py:13-21: a chunk whose lazy load raises TypeError (pyorc: the last frame of a video that cannot be
    decoded) is retried without its last frame, recursively; objects without .load() pass through; other exceptions propagate."""

    class Lazy:
        def __init__(self, n, bad_from, exc=TypeError):
            self.n, self.bad_from, self.exc, self.loads = n, bad_from, exc, []

        def __len__(self):
            return self.n

        def __getitem__(self, sl):
            assert sl == slice(None, -1)
            child = Lazy(self.n - 1, self.bad_from, self.exc)
            child.loads = self.loads
            return child

        def load(self):
            self.loads.append(self.n)
            if self.n > self.bad_from:
                raise self.exc("cannot decode")
            return self

    lazy = Lazy(6, 4)
    got = velocimetry.load_frame_chunk(lazy)
    assert len(got) == 4 and lazy.loads == [6, 5, 4]
    arr = np.zeros((3, 4, 5), np.uint8)
    assert velocimetry.load_frame_chunk(arr) is arr
    with pytest.raises(ValueError):
        velocimetry.load_frame_chunk(Lazy(3, 1, ValueError))


def test_chunk_shortened_by_the_loader_keeps_time_and_dt_aligned(fake):
    """A chunk that comes back one frame short (load_frame_chunk's retry) yields one time step less, with the time stamps and
    dt of the pairs that were actually computed (ffpiv.py:402-404: `time = da.time[1:]`, `dt.sel(time=time)`)."""
    da, res = make_frames(n=7)

    frames = da
    orig = velocimetry.load_frame_chunk
    try:
        velocimetry.load_frame_chunk = lambda chunk: chunk[:-1] if len(chunk) == 7 else orig(chunk)
        ws, ov = (32, 32), (16, 16)
        nr, nc = O.get_array_shape(frames.shape[-2:], ws, ov)
        t = np.asarray(frames["time"].values)
        dt = np.diff(t) * np.array([1.0, 1.1, 1.2, 1.3, 1.4, 1.5])
        ds = velocimetry.get_b2piv(frames, np.arange(nr), np.arange(nc), dt, ws, ov, ws, res, res, chunksize=7)
    finally:
        velocimetry.load_frame_chunk = orig
    assert ds["v_x"].shape == (5, nr, nc)
    assert np.array_equal(np.asarray(ds.coords["time"]), t[1:6])
    u, v, c, s = O.uv_timestep(np.asarray(frames.values)[:6], nc, nr, ws, ov)
    want = (u * res / dt[:5, None, None]).astype(np.float32)
    assert np.allclose(ds["v_x"].values, want, rtol=0, atol=1e-6, equal_nan=True)


def test_staging_threads_are_shared_between_engines_and_given_back(monkeypatch):
    """Several engines in one get_b2piv call split the host's cores for their staging threads; the option is sent only when it
    changes (it tears the copy threads down) and a later single-engine call on the same engine restores the engine's default."""
    import os

    class Opt:
        def __init__(self):
            self.sent = []

        def set_option(self, name, value):
            self.sent.append((name, value))

    monkeypatch.setattr(os, "cpu_count", lambda: 32)
    a, b, c, d = Opt(), Opt(), Opt(), Opt()
    velocimetry._share_host_cores([a])
    assert a.sent == []                                          # the single-engine path never touches the option
    velocimetry._share_host_cores([a, b, c, d])
    assert a.sent == b.sent == [("stage_threads", 8)]
    velocimetry._share_host_cores([a, b, c, d])
    assert a.sent == [("stage_threads", 8)]                      # unchanged: not sent again
    velocimetry._share_host_cores([a, b, c, d, Opt(), Opt(), Opt(), Opt()])
    assert a.sent[-1] == ("stage_threads", 4)
    velocimetry._share_host_cores([a])
    assert a.sent[-1] == ("stage_threads", 0) and len(a.sent) == 3
    velocimetry._share_host_cores([FakeEngine()])                # engines without options (test doubles) are left alone
