"""PIV engine binding for pyorc - B200 counterpart of ``pyorc/velocimetry/ffpiv.py``.

``get_b2piv`` has the signature, chunking rules, unit conversion, Dataset layout, warnings and errors of the
reference's ``get_ffpiv`` (pyorc/velocimetry/ffpiv.py:24-179); the arithmetic (``ffpiv.cross_corr`` +
``np.nanmax``/``np.nanmean`` + ``ffpiv.u_v_displacement``) runs fused in the CUDA engine
(:mod:`pyorc_b200.engine`).  No CPU fallback.
"""

from __future__ import annotations

import warnings
from typing import Optional, Sequence, Tuple

import numpy as np

try:  # a pyorc installation always has xarray; this build image does not
    import xarray as xr
except Exception:  # pragma: no cover - exercised in this image
    from . import _xr as xr

from . import window
from .engine import get_engine, merge_ensembles

__all__ = ["get_b2piv", "load_frame_chunk"]


def _metrics_sidecar(record: dict):
    """One JSON line per ``get_b2piv`` call in the file ``$B2PIV_METRICS`` names (the reference logs nothing about
    its PIV step; here: windows, windows/s, algorithmic bytes and flops of SURVEY.md 8d, wall time, devices)."""
    import json
    import os

    path = os.environ.get("B2PIV_METRICS")
    if not path:
        return
    try:
        with open(path, "a") as f:
            f.write(json.dumps(record) + "\n")
    except OSError:
        pass


def load_frame_chunk(da):
    """Load a frame chunk into memory; on ``TypeError`` retry with one frame less (ffpiv.py:13-21)."""
    if not hasattr(da, "load"):
        return da
    try:
        da_loaded = da.load()
    except TypeError:
        da_loaded = load_frame_chunk(da[:-1])
    return da_loaded


def _values(da) -> np.ndarray:
    return da.values if hasattr(da, "values") else np.asarray(da)


def _time_of(da, n):
    if hasattr(da, "coords") and "time" in getattr(da, "coords", {}):
        return np.asarray(da["time"].values if hasattr(da["time"], "values") else da["time"])
    return np.arange(n)


def get_b2piv(
    frames,
    y: np.ndarray,
    x: np.ndarray,
    dt,
    window_size: Tuple[int, int],
    overlap: Tuple[int, int],
    search_area_size: Tuple[int, int],
    res_y: float,
    res_x: float,
    chunksize: Optional[int] = None,
    memory_factor: float = 4,
    engine: str = "b200",
    ensemble_corr: bool = False,
    corr_min: float = 0.2,
    s2n_min: float = 3,
    count_min: float = 0.2,
    signal_threshold: Optional[float] = None,
    device: int = 0,
    coarse_pass: Optional[Tuple[Tuple[int, int], Tuple[int, int]]] = None,
    devices: Optional[Sequence[int]] = None,
    multipass: str = "offset",
):
    """Time-resolved or ensemble PIV on the B200 engine; same contract as ``get_ffpiv`` (ffpiv.py:24-179).

    Returns a Dataset with ``s2n``, ``corr``, ``v_x``, ``v_y`` on ``(time, y, x)``; velocities in m/s
    (``u * res_x / dt``), float32.  ``engine`` must be ``"b200"``; ``device`` selects the GPU.

    ``devices=[0, 1, ...]`` shards the work over several GPUs of the box from this ONE call: the frame pairs of
    every chunk of the reference's chunk list (ffpiv.py:140-142) are cut into contiguous sub-ranges with the same 1-frame
    halo, one per device, each handled by that device's engine on its own host thread (the ABI's "one engine per thread
    and device"; a device named twice gets two engines); results are put back in time order.  Per-time-step results are bit-identical to ``devices=[0]`` (frame
    pairs are independent); in ensemble mode the devices' plane sums are added before the peak fit (same result up to the
    float32 summation order), and the count filter keeps using the reference's chunk count.

    ``coarse_pass=((wy, wx), (oy, ox))`` (no reference counterpart - ffpiv is single pass): two-pass PIV with a discrete
    window offset, BASELINE.json ``configs[2]``: a first pass on that coarse grid gives a validated, interpolated
    whole-pixel predictor, ``window_size`` / ``overlap`` are the grid of the second pass (``Engine.pairs_two_pass``);
    per-time-step mode only.  ``multipass="deform"``: the second pass correlates frame k with frame k+1 RESAMPLED by the
    per-pixel predictor (bilinear window deformation) instead of displacing whole windows by whole pixels.
    """
    CHUNK_SIZE_ERROR = (
        "Chunk size with selected nr of chunks ({chunks}) is 2 or less. If you manually "
        "selected `chunks={chunks}` then consider increasing chunk size to at least 2, and preferrably more. If memory "
        "is limited, consider closing memory intensive applications. If pyorc crashes, then this is due to "
        " insufficient memory."
    )
    CHUNK_SIZE_WARNING = (
        "Memory availability is poor ({avail_mem} GB). Chunk size is automatically set to {chunksize} to avoid "
        "memory issues. If pyorc crashes, then this is due to insufficient memory. Consider to manually set a lower "
        "chunk size e.g using `get_piv(engine={engine}, chunk=2)` or `get_piv(engine={engine}, chunk=3)` or close "
        "memory intensive applications."
    )
    if engine != "b200":
        raise ValueError(f"Selected PIV engine {engine} does not exist.")
    if tuple(search_area_size) != tuple(window_size):
        raise NotImplementedError("search_area_size must equal window_size (pyorc/api/frames.py:168)")
    n_total = len(frames)
    dim_size = frames[0].shape
    devices = [int(device)] if devices is None else [int(d) for d in devices]
    if not devices:
        raise ValueError("devices must name at least one GPU")
    req_mem = window.required_memory(
        n_frames=n_total, dim_size=dim_size, window_size=window_size, overlap=overlap, search_area_size=search_area_size
    )
    if chunksize is None:
        # the smallest free HBM of the devices used sizes the chunks (every device holds one sub-range of a chunk at a time)
        # (a recent answer is reused while it is at least twice what this call needs - window.available_memory)
        avail_mem = min(window.available_memory(d, need=req_mem * memory_factor) for d in devices) / memory_factor
        chunks = int((req_mem // avail_mem) + 1)
        chunksize = int(np.ceil(n_total / chunks))
        if chunksize <= 5:
            warnings.warn(
                CHUNK_SIZE_WARNING.format(avail_mem=avail_mem / 1e9, chunksize=chunksize, engine=engine), stacklevel=2
            )
            chunksize = 5
            chunks = int(np.ceil(n_total / chunksize))
    else:
        # the reference leaves `chunks` undefined here (NameError at ffpiv.py:140); a given chunksize just works
        chunksize = int(chunksize)
        chunks = int(np.ceil(n_total / max(chunksize, 1)))
    if chunksize < 2:
        raise OverflowError(CHUNK_SIZE_ERROR.format(chunks=chunks))
    # 1-frame halo between consecutive chunks (ffpiv.py:140), positions kept for dt / time bookkeeping
    bounds = [(max(c * chunksize - 1, 0), min((c + 1) * chunksize, n_total)) for c in range(chunks)]
    bounds = [(a, b) for a, b in bounds if b - a >= 2]
    n_rows, n_cols = len(y), len(x)
    exp_rows, exp_cols = window.get_array_shape(dim_size, window_size, overlap)
    if (n_rows, n_cols) != (exp_rows, exp_cols):
        raise ValueError(f"y/x lengths {(n_rows, n_cols)} do not match the PIV field shape {(exp_rows, exp_cols)}")
    times = _time_of(frames, n_total)
    dt_vals = np.asarray(_values(dt), dtype=np.float64).reshape(-1)
    if dt_vals.size != n_total - 1:
        raise ValueError("dt must hold one interval per frame pair")
    # a device named k times gets k engines (slots): each engine is driven by one host thread only
    engs = [get_engine(d, devices[:i].count(d)) if devices[:i].count(d) else get_engine(d) for i, d in enumerate(devices)]
    _share_host_cores(engs)
    common = (frames, bounds, times, dt_vals, y, x, res_y, res_x, n_rows, n_cols, window_size, overlap, engs)
    if coarse_pass is not None:
        if ensemble_corr:
            raise NotImplementedError("coarse_pass (two-pass PIV) is available in per-time-step mode only")
        if signal_threshold is not None:
            raise NotImplementedError("coarse_pass (two-pass PIV) does not take a signal_threshold")
        (cwy, cwx), (coy, cox) = coarse_pass
        if cwy < window_size[0] or cwx < window_size[1] or cwy > dim_size[0] or cwx > dim_size[1]:
            raise ValueError("the coarse window must be at least as large as window_size and fit the frame")
        if multipass not in ("offset", "deform"):
            raise ValueError("multipass must be 'offset' or 'deform'")
        coarse_pass = ((int(cwy), int(cwx)), (int(coy), int(cox)), multipass)
    import time as _time

    t0 = _time.perf_counter()
    if ensemble_corr:
        ds = _get_b2piv_mean(*common, corr_min, s2n_min, count_min, signal_threshold)
    else:
        ds = _get_b2piv_timestep(*common, signal_threshold, coarse_pass)
    wall = _time.perf_counter() - t0
    n_win = (n_total - 1) * n_rows * n_cols
    wy, wx = window_size
    s_in = 1 if getattr(frames, "dtype", None) == np.uint8 else 4
    _metrics_sidecar({
        "call": "get_b2piv", "mode": "ensemble" if ensemble_corr else ("two-pass " + coarse_pass[2] if coarse_pass else "per-time-step"),
        "frames": n_total, "frame_shape": [int(dim_size[0]), int(dim_size[1])], "window": [int(wy), int(wx)], "overlap": [int(overlap[0]), int(overlap[1])],
        "chunks": len(bounds), "devices": devices, "windows": int(n_win), "wall_s": wall, "windows_per_s": n_win / wall if wall > 0 else None,
        "alg_bytes": int(n_win * (2 * (wy - overlap[0]) * (wx - overlap[1]) * s_in + 16)),
        "alg_flops": float(n_win * (3 * 2.5 * wy * wx * np.log2(wy * wx) + 6 * wy * (wx // 2 + 1))),
        "kernel_launches": int(sum(getattr(e, "launch_count", 0) for e in engs)),
    })
    return ds


def _share_host_cores(engs):
    """Every engine stages pageable frames with its own copy threads: several engines in one call share the host's cores instead
    of oversubscribing them; a later single-engine call on the same (process-wide) engine gets the engine's own default back.
    The option is only sent when it changes - setting it tears down the engine's copy threads."""
    import os

    per = 0 if len(engs) == 1 else max(2, min(8, (os.cpu_count() or 8) // len(engs)))      # 0: the engine's default
    for e in engs:
        if hasattr(e, "set_option") and getattr(e, "_stage_threads_shared", 0) != per:
            e.set_option("stage_threads", per)
            e._stage_threads_shared = per


def _get_uv_timestep(da, n_cols, n_rows, window_size, overlap, search_area_size, engine, signal_threshold=None, coarse_pass=None, units=None):
    """``u, v`` [px/frame], ``corr_max``, ``s2n`` - the narrow waist (ffpiv.py:446-474), fused on the GPU.  ``units=(res_x, res_y,
    dt)``: the engine converts to m / s itself (single pass only)."""
    if coarse_pass is not None:
        mode = coarse_pass[2] if len(coarse_pass) > 2 else "offset"
        kw = {"mode": mode} if mode != "offset" else {}
        u, v, corr_max, s2n = engine.pairs_two_pass(_values(da), tuple(coarse_pass[:2]), (tuple(window_size), tuple(overlap)), **kw)
    elif units is not None:
        u, v, corr_max, s2n = engine.pairs(_values(da), window_size, overlap, signal_threshold=signal_threshold, units=units)
    else:
        u, v, corr_max, s2n = engine.pairs(_values(da), window_size, overlap, signal_threshold=signal_threshold)
    assert u.shape[1:] == (n_rows, n_cols)
    return u, v, corr_max, s2n


def _work_items(bounds, n_dev):
    """Cut every chunk ``(a, b)`` (frames a .. b-1, pairs a .. b-2) into at most ``n_dev`` contiguous sub-ranges of frame
    pairs, each with its 1-frame halo (the rule of ffpiv.py:140 applied once more): ``[(chunk index, a_k, b_k)]`` in time
    order.  One device -> the reference's chunk list unchanged."""
    items = []
    for c, (a, b) in enumerate(bounds):
        n_pairs = b - a - 1
        parts = max(1, min(n_dev, n_pairs))
        base, rem = divmod(n_pairs, parts)
        p0 = a
        for k in range(parts):
            p1 = p0 + base + (1 if k < rem else 0)
            items.append((c, p0, p1 + 1))
            p0 = p1
    return items


def _run_items(items, engs, job):
    """``job(item, engine)`` for every work item; item i runs on device i % D.  One host thread per device, each working
    through ITS items in order (an engine is used by one thread only); the first exception is re-raised.  Returns the results
    in item order."""
    n_dev = len(engs)
    results = [None] * len(items)
    if n_dev == 1:
        for i, it in enumerate(items):
            results[i] = job(it, engs[0])
        return results
    import threading

    errors = []

    def worker(k):
        try:
            for i in range(k, len(items), n_dev):
                if errors:
                    return
                results[i] = job(items[i], engs[k])
        except BaseException as exc:  # noqa: BLE001 - re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,), name=f"b2piv-dev{engs[k].device}") for k in range(min(n_dev, len(items)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def _get_b2piv_timestep(frames, bounds, times, dt_vals, y, x, res_y, res_x, n_rows, n_cols, window_size, overlap, engs,
                        signal_threshold, coarse_pass=None):
    """Per-time-step mode (ffpiv.py:379-443)."""

    def job(item, eng):
        _, a, b = item
        da = load_frame_chunk(frames[a:b])
        if len(da) < 2:
            return None
        b = a + len(da)
        time = times[a + 1 : b]
        dt_chunk = dt_vals[a : b - 1]
        if coarse_pass is None and getattr(eng, "fused_units", False):
            # px -> m/s inside the engine call (ffpiv.py:418-419 with numpy's own float32 / float64 arithmetic), no host passes
            u, v, corr_max, s2n = _get_uv_timestep(da, n_cols, n_rows, window_size, overlap, window_size, eng, signal_threshold, None,
                                                   units=(res_x, res_y, dt_chunk))
        else:
            u, v, corr_max, s2n = _get_uv_timestep(da, n_cols, n_rows, window_size, overlap, window_size, eng, signal_threshold, coarse_pass)
            u = (u * res_x / np.expand_dims(dt_chunk, (1, 2))).astype(np.float32)
            v = (v * res_y / np.expand_dims(dt_chunk, (1, 2))).astype(np.float32)
        ds = xr.Dataset(
            {
                "s2n": (["time", "y", "x"], s2n),
                "corr": (["time", "y", "x"], corr_max),
                "v_x": (["time", "y", "x"], u),
                "v_y": (["time", "y", "x"], v),
            },
            coords={"time": time, "y": y, "x": x},
        )
        del da
        return ds

    if signal_threshold is not None and len(engs) > 1:
        # the reference scores a window over ALL frames of a chunk (ffpiv.py:93-97): sub-ranges would change the score
        engs = engs[:1]
    ds_piv_chunks = [ds for ds in _run_items(_work_items(bounds, len(engs)), engs, job) if ds is not None]
    # (the reference runs gc.collect() after every chunk to get its window stacks and correlation planes - GBs - out of RAM,
    # ffpiv.py:436-438; nothing of that size exists here, and a full collection costs ~35 ms in a process that has torch loaded,
    # nine times the whole 100-pair 1080p call)
    if len(ds_piv_chunks) == 1:       # nothing to concatenate: no copy of the four fields
        return ds_piv_chunks[0]
    return xr.concat(ds_piv_chunks, dim="time")


def _get_b2piv_mean(frames, bounds, times, dt_vals, y, x, res_y, res_x, n_rows, n_cols, window_size, overlap, engs,
                    corr_min, s2n_min, count_min, signal_threshold):
    """Ensemble-correlation mode (ffpiv.py:182-376): thresholds and plane sums on the device(s), the tiny
    per-pair statistics aggregated on the host exactly like ``aggregate_results``."""
    from .parallel import aggregate_ensemble

    if signal_threshold is not None and len(engs) > 1:
        engs = engs[:1]   # see _get_b2piv_timestep
    items = _work_items(bounds, len(engs))
    opened = set()      # id() of the engines that hold accumulators of this call
    chunk_seen = set()

    def job(item, eng):
        c, a, b = item
        da = load_frame_chunk(frames[a:b])
        if len(da) < 2:
            return None
        vals = _values(da)
        if id(eng) not in opened:
            dtype = vals.dtype if vals.dtype in (np.uint8, np.float32) else np.float32
            eng.ens_begin(vals.shape[-2:], window_size, overlap, dtype)
            opened.add(id(eng))
        chunk_seen.add(c)
        corr_max, s2n = eng.ens_add(vals, window_size, overlap, corr_min=corr_min, s2n_min=s2n_min, signal_threshold=signal_threshold)
        del da
        return corr_max, s2n

    results = _run_items(items, engs, job)
    done = [(it, r) for it, r in zip(items, results) if r is not None]
    corr_chunks = [r[0] for _, r in done]
    s2n_chunks = [r[1] for _, r in done]
    # time coordinate: first stamp of the last chunk processed (ffpiv.py:336: time[0:1] of the last loop iteration)
    c_last = max(c for (c, _, _), _ in done)
    a_last = min(a for (c, a, _), _ in done if c == c_last)
    time = times[a_last + 1 : a_last + 2]
    dt_av = dt_vals.mean()
    n_frames = len(chunk_seen)  # number of CHUNKS, as in the reference (ffpiv.py:373) - not of device sub-ranges
    used = [e for e in engs if id(e) in opened]
    u, v, corr_count = merge_ensembles(used, count_min * n_frames)
    corr_max_mean, s2n_mean = aggregate_ensemble(np.concatenate(corr_chunks, axis=0), np.concatenate(s2n_chunks, axis=0), corr_count,
                                                 count_min * n_frames, n_rows, n_cols)
    u = (u.reshape(-1, n_rows, n_cols) * res_x / dt_av).astype(np.float32)
    v = (v.reshape(-1, n_rows, n_cols) * res_y / dt_av).astype(np.float32)
    return xr.Dataset(
        {
            "s2n": (["time", "y", "x"], s2n_mean),
            "corr": (["time", "y", "x"], corr_max_mean),
            "v_x": (["time", "y", "x"], u),
            "v_y": (["time", "y", "x"], v),
        },
        coords={"time": time, "y": y, "x": x},
    )
