"""CPU oracle: float64 numpy restatement of the ffpiv path that pyorc's ``get_piv`` runs.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under ``pyorc_b200/``
imports, calls or links anything under ``oracle/``; the product path fails loudly without its CUDA library.

What it restates
----------------
pyorc (localdevices/pyorc @ be7d7c8) delegates all PIV arithmetic to the un-vendored PyPI dependency
``ffpiv>=0.2.1`` (pyproject.toml:19) which in turn calls pocketfft through ``rocket-fft`` (pyproject.toml:38).
Neither package is present under /root/reference nor installable offline, so this file restates their published
algorithm (ffpiv's numba engine, which follows OpenPIV's ``fft_correlate_images`` / ``normalize_intensity`` /
3-point Gaussian ``find_subpixel_peak_position``), anchored on pyorc's own call sites:

* ``cross_corr``          <- pyorc/velocimetry/ffpiv.py:222-231, :450-459 (call signature, normalize=False)
* ``u_v_displacement``    <- pyorc/velocimetry/ffpiv.py:324, :471
* ``window.*``            <- pyorc/api/frames.py:85-90, :167 ; pyorc/velocimetry/ffpiv.py:120-129
* ``uv_timestep``         <- pyorc/velocimetry/ffpiv.py:446-474 (nanmax / nanmean / s2n, verbatim semantics)
* ``Ensemble``            <- pyorc/velocimetry/ffpiv.py:200-243, :245-288, :290-338, :345-376

PARITY PIN STATUS: PINNED against the reference's own golden vectors.  tests/golden/make_ngwerere_golden.py runs
the reference's Python (imported from /root/reference) up to the ffpiv call and this oracle then reproduces both
pinned ``v_x`` vectors of pyorc's tests/test_frames.py:139-153 (window 10, per-time-step and ensemble mode) to
< 2e-8 m/s (the pins are printed with 8 digits) - tests/test_golden.py.  ``v_y``, ``corr`` and ``s2n`` are not
value-checked by any reference test; the SIGN convention of ``v_y`` and the units are confirmed on the result file the reference
ships (examples/ngwerere/ngwerere_piv.nc, an older engine's output: spatial correlation +0.79 for v_y, 0.73 for v_x over 125 pairs -
tests/test_reference_example.py).  The switches below mark the details that pyorc's source does not determine
(they live in ffpiv); CLIP_NORMALIZED is decided by the pin, GAUSS_EPS and BORDER_RULE are not observable in it.

All FFTs are pocketfft (numpy.fft / scipy.fft) in float64 - the same algorithm family rocket-fft binds.
"""

from __future__ import annotations

import os
import warnings

import numpy as np

try:  # scipy's pocketfft has a `workers=` argument (threaded batch FFT); numpy.fft is the fallback.
    import scipy.fft as _fft

    _HAVE_SCIPY = True
except Exception:  # pragma: no cover
    import numpy.fft as _fft

    _HAVE_SCIPY = False

# --------------------------------------------------------------------------------------------------------------
# Details that live in ffpiv (not determinable from pyorc's source).  One documented default each.
# --------------------------------------------------------------------------------------------------------------
#: OpenPIV's ``normalize_intensity`` clips the zero-mean/unit-std window at 0 (``np.clip(w, 0, w.max())``); ffpiv
#: does NOT: only the un-clipped variant reproduces pyorc's golden vectors (tests/test_frames.py:142-143) - to
#: 2e-8 in both modes, see tests/test_golden.py - the clipped one is off by 25 %.  PINNED.
CLIP_NORMALIZED = False
#: epsilon added to the five correlation samples before the logs of the Gaussian fit.
GAUSS_EPS = 1e-7
#: what a correlation peak on the plane border yields: "nan" (no sub-pixel fit possible -> NaN displacement)
#: or "integer" (integer peak minus centre).
BORDER_RULE = "nan"


# --------------------------------------------------------------------------------------------------------------
# ffpiv.window  (geometry + memory model)           call sites: frames.py:85-90,167 ; ffpiv.py:120-129
# --------------------------------------------------------------------------------------------------------------
def round_to_even(input_tuple):
    """Round every entry to an even integer (odd -> next even).  pyorc/api/frames.py:167."""
    return tuple(int(round(x)) if int(round(x)) % 2 == 0 else int(round(x)) + 1 for x in input_tuple)


def get_axis_shape(dim_size: int, window_size: int, overlap: int) -> int:
    """Number of interrogation windows along one axis (OpenPIV ``get_field_shape`` rule)."""
    return (dim_size - window_size) // (window_size - overlap) + 1


def get_array_shape(dim_size, window_size, overlap):
    """(n_rows, n_cols) of the PIV field."""
    return (
        get_axis_shape(dim_size[0], window_size[0], overlap[0]),
        get_axis_shape(dim_size[1], window_size[1], overlap[1]),
    )


def get_axis_coords(dim_size: int, window_size: int, overlap: int) -> np.ndarray:
    """Integer window-centre coordinates along one axis: ``i*(w-o) + w//2``."""
    n = get_axis_shape(dim_size, window_size, overlap)
    return np.int64(np.arange(n) * (window_size - overlap) + window_size / 2.0)


def get_rect_coordinates(dim_size, window_size, overlap, search_area_size=None):
    """(cols_vector, rows_vector) - integer centres; pyorc indexes ``x[cols]``, ``y[rows]`` with them
    (pyorc/helpers.py:166-167)."""
    y = get_axis_coords(dim_size[0], window_size[0], overlap[0])
    x = get_axis_coords(dim_size[1], window_size[1], overlap[1])
    return x, y


def required_memory(n_frames, dim_size, window_size, overlap, search_area_size=None, safety=1.0):
    """Bytes the CPU path needs: window stack (f64 during FFT) + f32 correlation planes.  ffpiv.py:120-126."""
    n_rows, n_cols = get_array_shape(dim_size, window_size, overlap)
    per_plane = n_rows * n_cols * window_size[0] * window_size[1]
    return safety * (n_frames * per_plane * 8 + max(n_frames - 1, 0) * per_plane * 4)


def available_memory():
    """Available RAM in bytes.  ffpiv.py:129."""
    try:
        import psutil

        return float(psutil.virtual_memory().available)
    except Exception:  # pragma: no cover
        return float(os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_AVPHYS_PAGES"))


# --------------------------------------------------------------------------------------------------------------
# ffpiv.cross_corr
# --------------------------------------------------------------------------------------------------------------
def window_origins(dim_size, window_size, overlap):
    """Top-left corner (row, col) vectors of the windows: ``i*(w-o)``; windows are flattened row-major
    (index r*n_cols + c), see the reshape at pyorc/velocimetry/ffpiv.py:469-470."""
    n_rows, n_cols = get_array_shape(dim_size, window_size, overlap)
    y0 = np.arange(n_rows) * (window_size[0] - overlap[0])
    x0 = np.arange(n_cols) * (window_size[1] - overlap[1])
    return y0, x0


def subwindows(imgs: np.ndarray, window_size, overlap) -> np.ndarray:
    """Gather ``[n, H, W] -> [n, n_rows*n_cols, wy, wx]`` (the 4-16x expansion the GPU engine never builds)."""
    imgs = np.asarray(imgs)
    y0, x0 = window_origins(imgs.shape[-2:], window_size, overlap)
    wy, wx = window_size
    yy = (y0[:, None, None, None] + np.arange(wy)[None, None, :, None])  # [r,1,wy,1]
    xx = (x0[None, :, None, None] + np.arange(wx)[None, None, None, :])  # [1,c,1,wx]
    out = imgs[:, yy, xx]  # [n, r, c, wy, wx]
    return out.reshape(imgs.shape[0], len(y0) * len(x0), wy, wx)


def normalize_intensity(win: np.ndarray) -> np.ndarray:
    """Per-window ``(w-mean)/std`` (std==0 -> zeros), optionally clipped at 0 (see CLIP_NORMALIZED); float64."""
    win = win.astype(np.float64)
    win = win - win.mean(axis=(-2, -1), keepdims=True)
    std = win.std(axis=(-2, -1), keepdims=True)
    out = np.divide(win, std, out=np.zeros_like(win), where=(std != 0))
    if CLIP_NORMALIZED:
        out = np.clip(out, 0.0, None)
    return out


def ncc(win_a: np.ndarray, win_b: np.ndarray, workers: int = 1) -> np.ndarray:
    """Normalised circular cross-correlation planes, fftshifted, /N, clipped to [0, 1]; float64 in, float64 out."""
    n_px = win_a.shape[-2] * win_a.shape[-1]
    a = normalize_intensity(win_a)
    b = normalize_intensity(win_b)
    kw = {"workers": workers} if _HAVE_SCIPY else {}
    f2a = np.conj(_fft.rfft2(a, **kw))
    f2b = _fft.rfft2(b, **kw)
    c = _fft.irfft2(f2a * f2b, s=a.shape[-2:], **kw)
    c = np.fft.fftshift(c, axes=(-2, -1))
    return np.clip(c / n_px, 0.0, 1.0)


def signal_mask(win_stack: np.ndarray, signal_threshold) -> np.ndarray | None:
    """Windows whose fraction of non-zero pixels over the whole window stack (all frames of the call) is below
    ``signal_threshold`` get NaN planes (pyorc/velocimetry/ffpiv.py:93-97).  Returns bool [n_windows] keep-mask."""
    if signal_threshold is None:
        return None
    score = (win_stack != 0).mean(axis=(0, 2, 3))
    return score >= signal_threshold


def cross_corr(
    imgs,
    window_size=(64, 64),
    overlap=(32, 32),
    search_area_size=None,
    normalize=False,
    engine="numba",
    signal_threshold=None,
    verbose=False,
    workers: int = 1,
):
    """Restatement of ``ffpiv.cross_corr`` as pyorc calls it (ffpiv.py:222-231, :450-459).

    Returns ``x`` (cols), ``y`` (rows) centre vectors and ``corr`` float32 ``[n-1, n_windows, wy, wx]``.
    """
    imgs = np.asarray(imgs)
    search_area_size = window_size if search_area_size is None else search_area_size
    if tuple(search_area_size) != tuple(window_size):
        raise NotImplementedError("pyorc always passes search_area_size == window_size (api/frames.py:168)")
    if normalize:
        raise NotImplementedError("pyorc always passes normalize=False (ffpiv.py:227,455)")
    x, y = get_rect_coordinates(imgs.shape[-2:], window_size, overlap)
    stack = subwindows(imgs, window_size, overlap)
    keep = signal_mask(stack, signal_threshold)
    corr = np.empty((stack.shape[0] - 1,) + stack.shape[1:], dtype=np.float32)
    for n in range(stack.shape[0] - 1):
        corr[n] = ncc(stack[n], stack[n + 1], workers=workers)
    if keep is not None:
        corr[:, ~keep] = np.nan
    return x, y, corr


# --------------------------------------------------------------------------------------------------------------
# ffpiv.u_v_displacement
# --------------------------------------------------------------------------------------------------------------
def peak_position(corr: np.ndarray) -> np.ndarray:
    """Sub-pixel peak (row, col) of correlation planes ``[..., wy, wx]`` - first-occurrence argmax + 3-point
    Gaussian fit.  Vectorised over leading dims; float64 arithmetic on the float32 samples."""
    wy, wx = corr.shape[-2:]
    flat = corr.reshape(-1, wy * wx)
    nanplane = np.isnan(flat).all(axis=1)
    safe = np.where(np.isnan(flat), -np.inf, flat)
    idx = np.argmax(safe, axis=1)  # first occurrence, row-major
    pi, pj = idx // wx, idx % wx
    border = (pi == 0) | (pi == wy - 1) | (pj == 0) | (pj == wx - 1)
    ii = np.clip(pi, 1, wy - 2)
    jj = np.clip(pj, 1, wx - 2)
    planes = flat.reshape(-1, wy, wx).astype(np.float64)
    n = np.arange(planes.shape[0])
    c = planes[n, ii, jj] + GAUSS_EPS
    cl = planes[n, ii - 1, jj] + GAUSS_EPS
    cr = planes[n, ii + 1, jj] + GAUSS_EPS
    cd = planes[n, ii, jj - 1] + GAUSS_EPS
    cu = planes[n, ii, jj + 1] + GAUSS_EPS
    with np.errstate(all="ignore"):
        di = (np.log(cl) - np.log(cr)) / (2 * np.log(cl) - 4 * np.log(c) + 2 * np.log(cr))
        dj = (np.log(cd) - np.log(cu)) / (2 * np.log(cd) - 4 * np.log(c) + 2 * np.log(cu))
    sub_i = pi + di
    sub_j = pj + dj
    if BORDER_RULE == "nan":
        sub_i = np.where(border, np.nan, sub_i)
        sub_j = np.where(border, np.nan, sub_j)
    else:
        sub_i = np.where(border, pi.astype(np.float64), sub_i)
        sub_j = np.where(border, pj.astype(np.float64), sub_j)
    sub_i = np.where(nanplane, np.nan, sub_i)
    sub_j = np.where(nanplane, np.nan, sub_j)
    return np.stack([sub_i, sub_j], axis=-1).reshape(corr.shape[:-2] + (2,))


def u_v_displacement(corr: np.ndarray, n_rows: int, n_cols: int, engine="numba"):
    """``u`` = column (x) shift, ``v`` = row (y, image-down) shift in pixels/frame, ``[..., n_rows, n_cols]``.
    No sign flip is applied by pyorc afterwards (ffpiv.py:325-326, :418-419)."""
    wy, wx = corr.shape[-2:]
    pk = peak_position(corr)
    v = pk[..., 0] - wy // 2
    u = pk[..., 1] - wx // 2
    lead = corr.shape[:-3]
    return u.reshape(lead + (n_rows, n_cols)), v.reshape(lead + (n_rows, n_cols))


# --------------------------------------------------------------------------------------------------------------
# pyorc-side wrappers (semantics verbatim from pyorc/velocimetry/ffpiv.py)
# --------------------------------------------------------------------------------------------------------------
def uv_timestep(imgs, n_cols, n_rows, window_size, overlap, search_area_size=None, signal_threshold=None, workers=1):
    """``_get_uv_timestep`` (ffpiv.py:446-474): returns ``u, v`` [px/frame], ``corr_max``, ``s2n`` (float32),
    each ``[n-1, n_rows, n_cols]``."""
    _, _, corr = cross_corr(
        imgs,
        window_size=window_size,
        overlap=overlap,
        search_area_size=search_area_size,
        normalize=False,
        signal_threshold=signal_threshold,
        verbose=False,
        workers=workers,
    )
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        corr_max = np.nanmax(corr, axis=(-1, -2))
        with np.errstate(all="ignore"):
            s2n = corr_max / np.nanmean(corr, axis=(-1, -2))
    s2n = (s2n.reshape(-1, n_rows, n_cols)).astype(np.float32)
    corr_max = (corr_max.reshape(-1, n_rows, n_cols)).astype(np.float32)
    u, v = u_v_displacement(corr, n_rows, n_cols)
    return u, v, corr_max, s2n


class Ensemble:
    """Ensemble-correlation mode, ``_get_ffpiv_mean`` (ffpiv.py:182-376), chunk by chunk."""

    def __init__(self, n_rows, n_cols, window_size, overlap, corr_min=0.2, s2n_min=3.0, count_min=0.2, signal_threshold=None):
        self.n_rows, self.n_cols = n_rows, n_cols
        self.window_size, self.overlap = tuple(window_size), tuple(overlap)
        self.corr_min, self.s2n_min, self.count_min = corr_min, s2n_min, count_min
        self.signal_threshold = signal_threshold
        self.corr_sum, self.corr_count = 0.0, 0.0  # ffpiv.py:345
        self.corr_chunks, self.s2n_chunks = [], []

    def add_chunk(self, imgs, workers=1):
        """``process_frame_chunk`` + accumulation (ffpiv.py:200-243, :359-365)."""
        _, _, corr = cross_corr(
            imgs, window_size=self.window_size, overlap=self.overlap, signal_threshold=self.signal_threshold, workers=workers
        )
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore", category=RuntimeWarning)
            corr_max = np.max(corr, axis=(-1, -2))
            s2n = corr_max / np.mean(corr, axis=(-1, -2))
            masks = (corr_max >= self.corr_min) & (s2n >= self.s2n_min) & (np.isfinite(corr_max))
        corr[~masks] = 0.0
        corr_max[~masks] = 0.0
        s2n[~masks] = 0.0
        self.corr_sum = self.corr_sum + np.sum(corr, axis=0, keepdims=True)
        self.corr_count = self.corr_count + np.sum(corr_max > 1e-6, axis=0, keepdims=True)
        self.corr_chunks.append(corr_max)
        self.s2n_chunks.append(s2n)

    def finalize(self):
        """``aggregate_results`` + ``u_v_displacement`` (ffpiv.py:245-288, :324).  NB: ``n_frames`` is the number
        of *chunks* in the reference (ffpiv.py:373) - mirrored here."""
        n_frames = len(self.corr_chunks)
        s2n_concat = np.concatenate(self.s2n_chunks, axis=0)
        corr_max_concat = np.concatenate(self.corr_chunks, axis=0)
        corr_sum = np.array(self.corr_sum, dtype=np.float32, copy=True)
        corr_count = np.asarray(self.corr_count)
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore", category=RuntimeWarning)
            corr_sum[corr_count < self.count_min * n_frames] = np.nan
            corr_max_concat[:, corr_count.flatten() < self.count_min * n_frames] = np.nan
            corr_mean = np.divide(corr_sum, corr_count[..., None, None])
            corr_max_mean = np.nanmean(corr_max_concat, axis=0).reshape(-1, self.n_rows, self.n_cols)
            s2n_mean = np.nanmean(s2n_concat, axis=0).reshape(-1, self.n_rows, self.n_cols)
        u, v = u_v_displacement(corr_mean, self.n_rows, self.n_cols)
        return u, v, corr_max_mean, s2n_mean


# --------------------------------------------------------------------------------------------------------------
# Timed CPU baseline: same passes over memory as pyorc performs, pocketfft threaded over the window batch.
# --------------------------------------------------------------------------------------------------------------
def _pair_job(args):
    a, b, window_size, overlap, n_rows, n_cols = args
    wa = subwindows(a[None], window_size, overlap)[0]
    wb = subwindows(b[None], window_size, overlap)[0]
    corr = ncc(wa, wb, workers=1).astype(np.float32)[None]
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore", category=RuntimeWarning)
        corr_max = np.nanmax(corr, axis=(-1, -2))
        s2n = corr_max / np.nanmean(corr, axis=(-1, -2))
    u, v = u_v_displacement(corr, n_rows, n_cols)
    return u[0], v[0], corr_max.reshape(n_rows, n_cols).astype(np.float32), s2n.reshape(n_rows, n_cols).astype(np.float32)


def cpu_reference_pairs(imgs, window_size, overlap, workers=None, pool=None):
    """The reference's per-time-step CPU path on ``imgs`` using all host cores the way ffpiv's numba engine does
    (``prange`` over frame pairs): one worker per frame pair, each doing gather -> normalise -> rfft2.conj.irfft2 ->
    fftshift,/N,clip -> f32 corr -> nanmax, nanmean -> argmax + Gaussian, i.e. the same passes over memory pyorc performs
    (ffpiv.py:446-474).  ``pool``: a :func:`make_pool` PROCESS pool (every core busy: the window gather is numpy fancy indexing
    and holds the GIL, so threads top out at a fraction of the cores - round 1 measured 16 threads of 32 cores, and a 128-core box
    SLOWER than a 16-core one); without a pool: threads."""
    from concurrent.futures import ThreadPoolExecutor

    imgs = np.asarray(imgs)
    n_rows, n_cols = get_array_shape(imgs.shape[-2:], window_size, overlap)
    jobs = [(imgs[k], imgs[k + 1], tuple(window_size), tuple(overlap), n_rows, n_cols) for k in range(imgs.shape[0] - 1)]
    if pool is not None:
        res = list(pool.map(_pair_job, jobs))
    else:
        workers = workers or os.cpu_count() or 1
        with ThreadPoolExecutor(max_workers=max(1, min(workers, len(jobs)))) as ex:
            res = list(ex.map(_pair_job, jobs))
    return tuple(np.stack([r[i] for r in res]) for i in range(4))


def _pool_warm(_):
    import time

    np.fft.rfft2(np.zeros((8, 8)))
    time.sleep(0.05)   # keeps the worker busy long enough for every worker of the pool to take one warm-up job
    return os.getpid()


def make_pool(workers=None):
    """Process pool (fork) for :func:`cpu_reference_pairs`, every worker started and warmed.  Create it BEFORE the parent
    initialises CUDA (bench.py does): the children only ever run numpy."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    workers = workers or os.cpu_count() or 1
    pool = ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork"))
    pids = set(pool.map(_pool_warm, range(4 * workers)))
    pool.n_workers = workers
    pool.n_started = len(pids)
    return pool
