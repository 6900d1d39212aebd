"""ctypes binding of ``libb2piv.so`` (C ABI in ``include/b2piv.h``) - the B200 LSPIV engine.

This is the host side of the drop-in boundary for pyorc's ``_get_uv_timestep`` / ``_get_ffpiv_mean``
(pyorc/velocimetry/ffpiv.py:446-474, :182-376): numpy (host) or torch-CUDA (device) frame stacks in, the four
result fields out.  There is NO CPU fallback: without the CUDA library or without a B200 every call raises.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2PIV_LIB: development override (A/B builds of the kernels); the default is the in-tree library
_LIB_PATH = os.environ.get("B2PIV_LIB") or os.path.join(_HERE, "libb2piv.so")
_lib = None

B2PIV_U8, B2PIV_F32 = 0, 1
_ERR = {1: ValueError, 2: RuntimeError, 3: NotImplementedError, 4: RuntimeError}

# every symbol include/b2piv.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "b2piv_version",
    "b2piv_create",
    "b2piv_destroy",
    "b2piv_last_error",
    "b2piv_set_option",
    "b2piv_plan",
    "b2piv_pairs_host",
    "b2piv_pairs_host_units",
    "b2piv_pairs_device",
    "b2piv_corr_planes_host",
    "b2piv_ens_begin",
    "b2piv_ens_begin_device",
    "b2piv_ens_add_host",
    "b2piv_ens_add_device",
    "b2piv_ens_accum",
    "b2piv_ens_finish_host",
    "b2piv_ens_finish_device",
    "b2piv_peaks_host",
    "b2piv_pre_normalize_device",
    "b2piv_pre_time_diff_device",
    "b2piv_pre_minmax_device",
    "b2piv_pre_gauss_device",
    "b2piv_project_plan",
    "b2piv_project_device",
    "b2piv_mask_elementwise",
    "b2piv_time_stats",
    "b2piv_mask_count",
    "b2piv_mask_outliers",
    "b2piv_mask_variance",
    "b2piv_mask_rolling",
    "b2piv_mask_window_nan",
    "b2piv_mask_window_mean",
    "b2piv_window_replace",
    "b2piv_mask_apply",
    "b2piv_encode_int16",
    "b2piv_decode_int16",
    "b2piv_rotate_uv",
    "b2piv_predictor_device",
    "b2piv_pairs_shifted_device",
    "b2piv_deform_device",
    "b2piv_pairs_interleaved_device",
    "b2piv_set_peer_outputs",
    "b2piv_peer_push",
    "b2piv_host_alloc",
    "b2piv_host_free",
    "b2piv_last_kernel_ms",
    "b2piv_launch_count",
    "b2piv_last_variant",
    "b2piv_fp32_peak",
)


def load_library(path: Optional[str] = None):
    """Load ``libb2piv.so`` and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or _LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(pyorc_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(path)
    vp, ci, cf, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
    fp = ctypes.POINTER(ctypes.c_float)
    lib.b2piv_version.restype = ci
    lib.b2piv_create.argtypes = [ctypes.POINTER(vp), ci]
    lib.b2piv_destroy.argtypes = [vp]
    lib.b2piv_destroy.restype = None
    lib.b2piv_last_error.argtypes = [vp]
    lib.b2piv_last_error.restype = ctypes.c_char_p
    lib.b2piv_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_double]
    lib.b2piv_plan.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.b2piv_pairs_host.argtypes = [vp, vp, ci, cf, vp, vp, vp, vp]
    lib.b2piv_pairs_host_units.argtypes = [vp, vp, ci, cf, cf, cf, vp, vp, vp, vp, vp]
    lib.b2piv_pairs_device.argtypes = [vp, vp, cll, ci, ci, cf, vp, vp, vp, vp, vp]
    lib.b2piv_corr_planes_host.argtypes = [vp, vp, ci, cf, vp]
    lib.b2piv_ens_begin.argtypes = [vp]
    lib.b2piv_ens_begin_device.argtypes = [vp, vp]
    lib.b2piv_ens_add_host.argtypes = [vp, vp, ci, cf, cf, cf, vp, vp]
    lib.b2piv_ens_add_device.argtypes = [vp, vp, cll, ci, ci, cf, cf, cf, vp, vp, vp]
    lib.b2piv_ens_accum.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(cll), ctypes.POINTER(cll)]
    lib.b2piv_ens_finish_host.argtypes = [vp, cf, vp, vp, vp]
    lib.b2piv_ens_finish_device.argtypes = [vp, cf, cll, cll, vp, vp, vp]
    lib.b2piv_peaks_host.argtypes = [vp, vp, cll, ci, ci, vp, vp]
    lib.b2piv_pre_normalize_device.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp]
    lib.b2piv_pre_time_diff_device.argtypes = [vp, vp, ci, ci, ci, ci, cf, ci, vp, vp]
    lib.b2piv_pre_minmax_device.argtypes = [vp, vp, ci, cll, cf, cf, vp, vp]
    lib.b2piv_pre_gauss_device.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, vp]
    lib.b2piv_project_plan.argtypes = [vp, ci, ci, ci, ci, vp, vp, cll, vp, vp, cll, vp, cll]
    lib.b2piv_project_device.argtypes = [vp, vp, ci, ci, vp, ci, vp]
    cd, ip, vpp = ctypes.c_double, ctypes.POINTER(ci), ctypes.POINTER(vp)
    lib.b2piv_mask_elementwise.argtypes = [vp, ci, vp, vp, cll, cf, cf, vp, vp]
    lib.b2piv_time_stats.argtypes = [vp, vp, ci, cll, vp, vp, vp, vp]
    lib.b2piv_mask_count.argtypes = [vp, vp, ci, cll, cd, vp, vp]
    lib.b2piv_mask_outliers.argtypes = [vp, vp, vp, ci, cll, cf, ci, vp, vp]
    lib.b2piv_mask_variance.argtypes = [vp, vp, vp, ci, cll, cf, ci, vp, vp]
    lib.b2piv_mask_rolling.argtypes = [vp, vp, vp, ci, cll, ci, cf, vp, vp]
    lib.b2piv_mask_window_nan.argtypes = [vp, vp, ci, ci, ci, ip, cd, vp, vp]
    lib.b2piv_mask_window_mean.argtypes = [vp, vp, vp, ci, ci, ci, ip, cf, ci, vp, vp]
    lib.b2piv_window_replace.argtypes = [vp, vpp, ci, ci, ci, ci, ip, ci, vp]
    lib.b2piv_mask_apply.argtypes = [vp, vpp, ci, ci, cll, vp, ci, vp]
    lib.b2piv_encode_int16.argtypes = [vp, vp, cll, cf, ci, vp, vp]
    lib.b2piv_decode_int16.argtypes = [vp, vp, cll, cf, ci, vp, vp]
    lib.b2piv_rotate_uv.argtypes = [vp, vp, vp, cll, cd, vp, vp, vp]
    lib.b2piv_predictor_device.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp]
    lib.b2piv_pairs_shifted_device.argtypes = [vp, vp, cll, ci, ci, vp, vp, vp, vp, vp, vp]
    lib.b2piv_deform_device.argtypes = [vp, vp, cll, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    lib.b2piv_pairs_interleaved_device.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp]
    lib.b2piv_set_peer_outputs.argtypes = [vp, ci, vpp, cll, cll]
    lib.b2piv_peer_push.argtypes = [vp, vp, ci, ci, vpp, cll, cll, vp]
    lib.b2piv_host_alloc.argtypes = [ctypes.c_size_t]
    lib.b2piv_host_alloc.restype = vp
    lib.b2piv_host_free.argtypes = [vp]
    lib.b2piv_host_free.restype = None
    lib.b2piv_last_kernel_ms.argtypes = [vp, fp]
    lib.b2piv_launch_count.argtypes = [vp]
    lib.b2piv_launch_count.restype = cll
    lib.b2piv_last_variant.argtypes = [vp]
    lib.b2piv_fp32_peak.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double)]
    for name in ABI_SYMBOLS:
        getattr(lib, name)
    if path == _LIB_PATH:
        _lib = lib
    return lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _stream_ctx(torch, device, stream):
    """(raw stream handle, context manager that makes it torch's current stream) for an optional ``stream=`` argument (a raw
    ``cudaStream_t`` value or a ``torch.cuda.Stream``).  Tensors allocated inside the context belong to that stream for the
    caching allocator, so a kernel launched on it may use them without ``record_stream`` bookkeeping."""
    import contextlib

    if stream is None:
        return torch.cuda.current_stream(device).cuda_stream, contextlib.nullcontext()
    ts = stream if isinstance(stream, torch.cuda.Stream) else torch.cuda.ExternalStream(int(stream), device=device)
    return ts.cuda_stream, torch.cuda.stream(ts)


class _CudaView:
    """Expose a raw device pointer through ``__cuda_array_interface__`` (so torch can wrap it without a copy)."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class _PinnedPool:
    """Page-locked result buffers, recycled: a block goes back to the pool when the last numpy view of it is collected.

    Results copied into pageable ``np.empty`` arrays go through the driver's staging buffer (measured: ~0.3 ms for the 3 MB of
    a 100-pair 1080p call, 7 % of the whole end-to-end step); blocks are only ever allocated while the caller still holds
    every earlier result, so a steady loop allocates nothing."""

    def __init__(self, lib):
        import threading

        self._lib, self._free, self._lock, self._closed = lib, {}, threading.Lock(), False

    @staticmethod
    def _cls(nbytes: int) -> int:
        return max(1 << 16, 1 << (int(nbytes) - 1).bit_length())          # power-of-two size classes >= 64 KB

    def empty(self, shape, dtype=np.float32) -> np.ndarray:
        import weakref

        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        size = self._cls(max(count * dtype.itemsize, 1))
        with self._lock:
            lst = self._free.get(size)
            ptr = lst.pop() if lst else None
        if ptr is None:
            ptr = self._lib.b2piv_host_alloc(size)
            if not ptr:
                raise MemoryError("cudaHostAlloc failed")
        buf = (ctypes.c_ubyte * size).from_address(ptr)
        weakref.finalize(buf, self._give, ptr, size)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _give(self, ptr, size):
        with self._lock:
            if not self._closed:
                self._free.setdefault(size, []).append(ptr)
                return
        self._lib.b2piv_host_free(ptr)

    def close(self):
        with self._lock:
            self._closed = True
            blocks = [p for lst in self._free.values() for p in lst]
            self._free = {}
        for ptr in blocks:
            self._lib.b2piv_host_free(ptr)


def _serialised(fn):
    """Run an Engine method under the engine's lock (re-entrant: the composite calls nest)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self.lock:
            return fn(self, *args, **kwargs)

    return wrapper


class Engine:
    """One PIV engine bound to one CUDA device (mirrors the role of ffpiv's ``engine=`` back-ends)."""

    fused_units = True      # pairs(..., units=...) converts px / frame to m / s on the device

    def __init__(self, device: int = 0, clip_normalized: Optional[bool] = None, border_nan: Optional[bool] = None,
                 gauss_eps: Optional[float] = None):
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.b2piv_create(ctypes.byref(h), int(device))
        if rc != 0:
            msg = self._lib.b2piv_last_error(None).decode()
            raise _ERR.get(rc, RuntimeError)(f"b2piv_create failed: {msg}")
        self._h = h
        self.device = int(device)
        self._plan = None
        # An engine owns workspaces (device copy of the frames, result block, keep mask, ensemble accumulators ...) that its calls
        # share: one host thread at a time.  The lock makes a second thread WAIT instead of racing on them (pyorc under dask's
        # threaded scheduler); independent work belongs on a second engine (get_engine(device, slot)).
        import threading

        self.lock = threading.RLock()
        self._pinned = []
        self._results = _PinnedPool(self._lib)
        if clip_normalized is not None:
            self.set_option("clip_normalized", float(bool(clip_normalized)))
        if border_nan is not None:
            self.set_option("border_nan", float(bool(border_nan)))
        if gauss_eps is not None:
            self.set_option("gauss_eps", float(gauss_eps))

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.b2piv_last_error(self._h).decode()
            raise _ERR.get(rc, RuntimeError)(f"{what}: {msg}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            for p in self._pinned:
                self._lib.b2piv_host_free(p)
            self._pinned = []
            self._results.close()
            self._lib.b2piv_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, name: str, value: float):
        self._check(self._lib.b2piv_set_option(self._h, name.encode(), float(value)), "b2piv_set_option")

    @property
    def launch_count(self) -> int:
        return int(self._lib.b2piv_launch_count(self._h))

    @property
    def last_variant(self) -> int:
        """Kernel family of the last PIV call: 1 shared-memory FFT, 2 row-per-thread (native sizes), 3 direct, 4 row-per-thread padded."""
        return int(self._lib.b2piv_last_variant(self._h))

    @property
    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        self._check(self._lib.b2piv_last_kernel_ms(self._h, ctypes.byref(ms)), "b2piv_last_kernel_ms")
        return float(ms.value)

    def fp32_peak(self, iters: int = 20000) -> float:
        """Measured fp32 FMA throughput of this device in TFLOP/s (micro-benchmark of independent FFMA chains)."""
        t = ctypes.c_double()
        self._check(self._lib.b2piv_fp32_peak(self._h, int(iters), ctypes.byref(t)), "b2piv_fp32_peak")
        return float(t.value)

    def pinned_empty(self, shape, dtype=np.uint8) -> np.ndarray:
        """A page-locked numpy array (freed with the engine) - H2D at full PCIe rate."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = self._lib.b2piv_host_alloc(max(nbytes, 1))
        if not p:
            raise MemoryError("cudaHostAlloc failed")
        self._pinned.append(p)
        buf = (ctypes.c_ubyte * max(nbytes, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # ---- plan -----------------------------------------------------------------------------------------------
    @_serialised
    def plan(self, dim_size: Tuple[int, int], window_size: Tuple[int, int], overlap: Tuple[int, int], dtype) -> Tuple[int, int]:
        """Fix frame and window geometry; returns ``(n_rows, n_cols)``."""
        dt = np.dtype(dtype) if not isinstance(dtype, int) else None
        if dt is not None:
            if dt == np.uint8:
                code = B2PIV_U8
            elif dt == np.float32:
                code = B2PIV_F32
            else:
                raise TypeError(f"frames must be uint8 or float32, got {dt}")
        else:
            code = dtype
        key = (tuple(dim_size), tuple(window_size), tuple(overlap), code)
        if self._plan and self._plan[0] == key:
            return self._plan[1]
        nr, nc = ctypes.c_int(), ctypes.c_int()
        self._check(
            self._lib.b2piv_plan(self._h, int(dim_size[0]), int(dim_size[1]), int(window_size[0]), int(window_size[1]),
                                 int(overlap[0]), int(overlap[1]), code, ctypes.byref(nr), ctypes.byref(nc)),
            "b2piv_plan",
        )
        self._plan = (key, (nr.value, nc.value))
        return self._plan[1]

    def _prep(self, frames, window_size, overlap):
        """Validate a frame stack, (re)plan, return (frames, n, n_rows, n_cols, is_torch)."""
        if _is_torch(frames):
            import torch

            if not frames.is_cuda:
                raise TypeError("torch frames must live on the engine's CUDA device (pass numpy for host data)")
            if frames.dim() != 3:
                raise ValueError("frames must be [n, H, W]")
            if frames.dtype not in (torch.uint8, torch.float32):
                raise TypeError("frames must be uint8 or float32")
            if frames.stride(2) != 1:
                frames = frames.contiguous()
            dt = np.uint8 if frames.dtype == torch.uint8 else np.float32
            n, H, W = frames.shape
            nr, nc = self.plan((H, W), window_size, overlap, dt)
            return frames, n, nr, nc, True
        frames = np.asarray(frames)
        if frames.ndim != 3:
            raise ValueError("frames must be [n, H, W]")
        if frames.dtype not in (np.uint8, np.float32):
            # the reference accepts any real dtype and promotes to float64; the engine computes in float32
            frames = frames.astype(np.float32)
        frames = np.ascontiguousarray(frames)
        n, H, W = frames.shape
        nr, nc = self.plan((H, W), window_size, overlap, frames.dtype)
        return frames, n, nr, nc, False

    def _same_device(self, t):
        if t.device.index != self.device:
            raise ValueError(f"frames are on cuda:{t.device.index}, engine on cuda:{self.device}")

    # ---- per-time-step ----------------------------------------------------------------------------------------
    @_serialised
    def pairs(self, frames, window_size, overlap, signal_threshold: Optional[float] = None, stream=None, units=None):
        """``u, v, corr_max, s2n`` for every consecutive frame pair, each ``[n-1, n_rows, n_cols]`` float32.

        numpy in -> numpy out (H2D/D2H inside, synchronous); torch CUDA tensor in -> torch CUDA tensors out
        (stream-ordered on ``stream`` or torch's current stream, no synchronisation).
        ``units=(res_x, res_y, dt)`` (host frames): ``u, v`` come back in m / s, ``u * res_x / dt[k]`` with numpy's float32 /
        float64 arithmetic (ffpiv.py:418-419), converted on the device before the D2H copy."""
        frames, n, nr, nc, on_dev = self._prep(frames, window_size, overlap)
        if n < 2:
            raise ValueError("need at least 2 frames (one frame pair)")
        thr = -1.0 if signal_threshold is None else float(signal_threshold)
        if on_dev:
            import torch

            if units is not None:
                raise TypeError("units= applies to host (numpy) frames; device results stay in px / frame")
            self._same_device(frames)
            st, ctx = _stream_ctx(torch, frames.device, stream)
            with ctx:
                out = torch.empty((4, n - 1, nr, nc), dtype=torch.float32, device=frames.device)
                if stream is not None:
                    frames.record_stream(torch.cuda.current_stream(frames.device))   # also covers a .contiguous() temporary
            es = frames.element_size()
            self._check(
                self._lib.b2piv_pairs_device(self._h, frames.data_ptr(), frames.stride(0) * es, frames.stride(1) * es, n, thr,
                                             out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), st),
                "b2piv_pairs_device",
            )
            return out[0], out[1], out[2], out[3]
        # the four fields are views of ONE page-locked block that returns to the engine's pool when they are all released
        block = self._results.empty((4, n - 1, nr, nc), np.float32)
        outs = [block[k] for k in range(4)]
        if units is not None:
            res_x, res_y, dt = units
            dt = np.ascontiguousarray(dt, dtype=np.float64).reshape(-1)
            if dt.size != n - 1:
                raise ValueError("units: dt must hold one time step per frame pair")
            self._check(
                self._lib.b2piv_pairs_host_units(self._h, frames.ctypes.data, n, thr, float(np.float32(res_x)), float(np.float32(res_y)),
                                                 dt.ctypes.data, *[o.ctypes.data for o in outs]),
                "b2piv_pairs_host_units",
            )
            return tuple(outs)
        self._check(
            self._lib.b2piv_pairs_host(self._h, frames.ctypes.data, n, thr, *[o.ctypes.data for o in outs]),
            "b2piv_pairs_host",
        )
        return tuple(outs)

    # ---- fused multi-GPU gather ---------------------------------------------------------------------------------
    @_serialised
    def set_peer_outputs(self, peer_ptrs, pairs_total: int, pair_offset: int):
        """Make :meth:`pairs` (device tensors) also store its results straight into every peer's gather buffer
        ``[4, pairs_total, n_rows, n_cols]`` float32 (raw device pointers, e.g. ``symm_mem_handle.buffer_ptrs``) at this
        rank's ``pair_offset``; ``peer_ptrs=None`` or ``[]`` switches back.  Call after :meth:`plan`."""
        ptrs = list(peer_ptrs or [])
        arr = (ctypes.c_void_p * max(len(ptrs), 1))(*[int(p) for p in ptrs])
        self._check(self._lib.b2piv_set_peer_outputs(self._h, len(ptrs), arr, int(pairs_total), int(pair_offset)),
                    "b2piv_set_peer_outputs")

    @_serialised
    def peer_push(self, local, peer_ptrs, pairs_total: int, pair_offset: int, stream=None):
        """Copy a contiguous result block ``[4, n_pairs, n_rows, n_cols]`` (CUDA tensor, e.g. the base of what :meth:`pairs`
        returned) into every peer's gather buffer at ``pair_offset``, stream-ordered on ``stream`` / torch's current stream."""
        import torch

        if local.dim() != 4 or local.shape[0] != 4 or not local.is_contiguous() or local.dtype != torch.float32:
            raise ValueError("local must be a contiguous float32 [4, n_pairs, n_rows, n_cols] tensor")
        ptrs = [int(p) for p in peer_ptrs]
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        st, _ = _stream_ctx(torch, local.device, stream)
        self._check(self._lib.b2piv_peer_push(self._h, local.data_ptr(), int(local.shape[1]), len(ptrs), arr, int(pairs_total), int(pair_offset), st),
                    "b2piv_peer_push")

    # ---- two-pass (BASELINE configs[2]) -----------------------------------------------------------------------
    @_serialised
    def predictor(self, u1, v1, dim_size, coarse, fine, dtype=np.uint8):
        """Whole-pixel window shifts ``[n_pairs, n_rows, n_cols, 2]`` (dy, dx; int16, on the device) for a second pass on the
        ``fine = (window_size, overlap)`` grid from the pass-1 fields ``u1, v1`` of the ``coarse`` grid (CUDA tensors):
        universal outlier detection on 3x3 neighbourhoods, bilinear interpolation, rounding, clamping to the frame."""
        import torch

        (ws1, ov1), (ws2, ov2) = coarse, fine
        nr, nc = self.plan(dim_size, ws2, ov2, dtype)
        u1 = u1.contiguous().float()
        v1 = v1.contiguous().float()
        P, r1, c1 = u1.shape
        shift = torch.empty((P, nr, nc, 2), dtype=torch.int16, device=u1.device)
        st = torch.cuda.current_stream(u1.device).cuda_stream
        self._check(self._lib.b2piv_predictor_device(self._h, u1.data_ptr(), v1.data_ptr(), P, r1, c1, int(ws1[0]), int(ws1[1]),
                                                     int(ov1[0]), int(ov1[1]), shift.data_ptr(), st), "b2piv_predictor_device")
        return shift

    @_serialised
    def pairs_shifted(self, frames, window_size, overlap, shift):
        """Second pass: frame k+1's window of every (pair, window) displaced by ``shift[pair, row, col] = (dy, dx)``;
        returns ``u, v`` (= shift + residual), ``corr_max, s2n`` as CUDA tensors.  ``frames``: CUDA tensor [n, H, W]."""
        import torch

        frames, n, nr, nc, on_dev = self._prep(frames, window_size, overlap)
        if not on_dev:
            raise TypeError("pairs_shifted takes device-resident frames (use pairs_two_pass for host arrays)")
        if tuple(shift.shape) != (n - 1, nr, nc, 2) or shift.dtype != torch.int16 or not shift.is_cuda:
            raise ValueError(f"shift must be a CUDA int16 tensor of shape {(n - 1, nr, nc, 2)}")
        self._same_device(frames)
        shift = shift.contiguous()
        out = torch.empty((4, n - 1, nr, nc), dtype=torch.float32, device=frames.device)
        st = torch.cuda.current_stream(frames.device).cuda_stream
        es = frames.element_size()
        self._check(self._lib.b2piv_pairs_shifted_device(self._h, frames.data_ptr(), frames.stride(0) * es, frames.stride(1) * es, n,
                                                         shift.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                                         out[3].data_ptr(), st), "b2piv_pairs_shifted_device")
        return out[0], out[1], out[2], out[3]

    @_serialised
    def pairs_two_pass(self, frames, coarse=((64, 64), (32, 32)), fine=((32, 32), (24, 24)), mode: str = "offset", chunk_pairs: int = 64):
        """Two-pass PIV (BASELINE.json configs[2]; defined in DESIGN.md, no reference counterpart):
        pass 1 on the coarse grid -> validated predictor -> pass 2 on the fine grid.  ``mode="offset"``: frame k+1's windows are
        displaced by the predictor rounded to whole pixels; ``mode="deform"``: frame k+1 is resampled with the per-pixel
        predictor (bilinear window deformation), pass 2 sees the residual only.  numpy in -> numpy out; CUDA tensor in -> CUDA
        tensors out.  Returns ``u, v, corr_max, s2n`` on the fine grid (u, v = predictor + residual, px / frame)."""
        import torch

        if mode not in ("offset", "deform"):
            raise ValueError("mode must be 'offset' or 'deform'")
        was_np = not _is_torch(frames)
        if was_np:
            a = np.asarray(frames)
            if a.dtype not in (np.uint8, np.float32):
                a = a.astype(np.float32)
            frames = torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{self.device}")
        dt = np.uint8 if frames.dtype == torch.uint8 else np.float32
        u1, v1, _, _ = self.pairs(frames, coarse[0], coarse[1])
        if mode == "offset":
            shift = self.predictor(u1, v1, tuple(frames.shape[-2:]), coarse, fine, dt)
            res = self.pairs_shifted(frames, fine[0], fine[1], shift)
        else:
            res = self._pairs_deformed(frames, u1, v1, coarse, fine, chunk_pairs)
        return tuple(r.cpu().numpy() for r in res) if was_np else res

    @_serialised
    def deform(self, frames, u1, v1, coarse, fine):
        """The deformation step alone: interleaved float32 stack ``[2 (n - 1), H, W]`` = (frame k, warped frame k+1) and the
        un-rounded predictor ``[n - 1, n_rows, n_cols, 2]`` = (dv, du) at the fine window centres (CUDA tensors)."""
        import torch

        (ws1, ov1), (ws2, ov2) = coarse, fine
        if frames.stride(2) != 1:
            frames = frames.contiguous()
        self._same_device(frames)
        n, H, W = frames.shape
        nr, nc = self.plan((H, W), ws2, ov2, np.float32)
        u1 = u1.contiguous().float()
        v1 = v1.contiguous().float()
        P, r1, c1 = u1.shape
        if P != n - 1:
            raise ValueError("pass-1 fields must hold one field per frame pair")
        stack = torch.empty((2 * P, H, W), dtype=torch.float32, device=frames.device)
        pred = torch.empty((P, nr, nc, 2), dtype=torch.float32, device=frames.device)
        st = torch.cuda.current_stream(frames.device).cuda_stream
        es = frames.element_size()
        code = B2PIV_U8 if frames.dtype == torch.uint8 else B2PIV_F32
        self._check(self._lib.b2piv_deform_device(self._h, frames.data_ptr(), frames.stride(0) * es, frames.stride(1) * es, code, n,
                                                  u1.data_ptr(), v1.data_ptr(), r1, c1, int(ws1[0]), int(ws1[1]), int(ov1[0]), int(ov1[1]),
                                                  stack.data_ptr(), pred.data_ptr(), st), "b2piv_deform_device")
        return stack, pred

    def _pairs_deformed(self, frames, u1, v1, coarse, fine, chunk_pairs):
        """Pass 2 of the deformation scheme, ``chunk_pairs`` frame pairs at a time (the float32 stack of a chunk is
        2 * chunk_pairs frames: 1 GB for 64 pairs of 1080p)."""
        import torch

        n = frames.shape[0]
        nr, nc = self.plan(tuple(frames.shape[-2:]), fine[0], fine[1], np.float32)
        out = torch.empty((4, n - 1, nr, nc), dtype=torch.float32, device=frames.device)
        st = torch.cuda.current_stream(frames.device).cuda_stream
        for a in range(0, n - 1, max(int(chunk_pairs), 1)):
            b = min(a + max(int(chunk_pairs), 1), n - 1)
            stack, pred = self.deform(frames[a : b + 1], u1[a:b], v1[a:b], coarse, fine)
            self._check(self._lib.b2piv_pairs_interleaved_device(self._h, stack.data_ptr(), b - a, pred.data_ptr(), out[0, a:b].data_ptr(),
                                                                 out[1, a:b].data_ptr(), out[2, a:b].data_ptr(), out[3, a:b].data_ptr(), st),
                        "b2piv_pairs_interleaved_device")
            del stack, pred
        return out[0], out[1], out[2], out[3]

    @_serialised
    def corr_planes(self, frames, window_size, overlap, signal_threshold: Optional[float] = None) -> np.ndarray:
        """Full correlation planes ``[n-1, n_windows, wy, wx]`` float32 as ``ffpiv.cross_corr`` returns them
        (triage / parity only - the fast path never materialises them)."""
        frames, n, nr, nc, on_dev = self._prep(frames, window_size, overlap)
        if on_dev:
            frames = frames.cpu().numpy()
        if n < 2:
            raise ValueError("need at least 2 frames (one frame pair)")
        thr = -1.0 if signal_threshold is None else float(signal_threshold)
        corr = np.empty((n - 1, nr * nc, window_size[0], window_size[1]), dtype=np.float32)
        self._check(self._lib.b2piv_corr_planes_host(self._h, frames.ctypes.data, n, thr, corr.ctypes.data), "b2piv_corr_planes_host")
        return corr

    @_serialised
    def peaks(self, corr) -> Tuple[np.ndarray, np.ndarray]:
        """``u, v`` (pixels, shape ``corr.shape[:-2]``) of arbitrary correlation planes ``[..., wy, wx]``: first-occurrence
        argmax + 3-point Gaussian fit minus the centre - ``ffpiv.u_v_displacement`` without the reshape."""
        corr = np.ascontiguousarray(corr, dtype=np.float32)
        if corr.ndim < 2:
            raise ValueError("corr must be [..., wy, wx]")
        lead = corr.shape[:-2]
        n = int(np.prod(lead)) if lead else 1
        u = np.empty(n, dtype=np.float32)
        v = np.empty(n, dtype=np.float32)
        self._check(self._lib.b2piv_peaks_host(self._h, corr.ctypes.data, n, corr.shape[-2], corr.shape[-1], u.ctypes.data, v.ctypes.data),
                    "b2piv_peaks_host")
        return u.reshape(lead), v.reshape(lead)

    # ---- ensemble ---------------------------------------------------------------------------------------------
    @_serialised
    def ens_begin(self, dim_size, window_size, overlap, dtype, stream=None, device_ordered: bool = False):
        """Zero the accumulators.  ``device_ordered`` (or a ``stream``): stream-ordered on torch's current stream (or
        ``stream``) without synchronising; otherwise on the engine's own stream, synchronous."""
        nr, nc = self.plan(dim_size, window_size, overlap, dtype)
        if stream is not None or device_ordered:
            import torch

            st, _ = _stream_ctx(torch, torch.device("cuda", self.device), stream)
            self._check(self._lib.b2piv_ens_begin_device(self._h, st), "b2piv_ens_begin_device")
        else:
            self._check(self._lib.b2piv_ens_begin(self._h), "b2piv_ens_begin")
        return nr, nc

    @_serialised
    def ens_add(self, frames, window_size, overlap, corr_min=0.2, s2n_min=3.0, signal_threshold=None, stream=None):
        """Accumulate one chunk; returns masked per-pair ``corr_max, s2n`` ``[n-1, n_windows]``."""
        frames, n, nr, nc, on_dev = self._prep(frames, window_size, overlap)
        if n < 2:
            raise ValueError("need at least 2 frames (one frame pair)")
        thr = -1.0 if signal_threshold is None else float(signal_threshold)
        if on_dev:
            import torch

            self._same_device(frames)
            st, ctx = _stream_ctx(torch, frames.device, stream)
            with ctx:
                out = torch.empty((2, n - 1, nr * nc), dtype=torch.float32, device=frames.device)
                if stream is not None:
                    frames.record_stream(torch.cuda.current_stream(frames.device))
            es = frames.element_size()
            self._check(
                self._lib.b2piv_ens_add_device(self._h, frames.data_ptr(), frames.stride(0) * es, frames.stride(1) * es, n,
                                               float(corr_min), float(s2n_min), thr, out[0].data_ptr(), out[1].data_ptr(), st),
                "b2piv_ens_add_device",
            )
            return out[0], out[1]
        cm = np.empty((n - 1, nr * nc), dtype=np.float32)
        sn = np.empty((n - 1, nr * nc), dtype=np.float32)
        self._check(
            self._lib.b2piv_ens_add_host(self._h, frames.ctypes.data, n, float(corr_min), float(s2n_min), thr, cm.ctypes.data, sn.ctypes.data),
            "b2piv_ens_add_host",
        )
        return cm, sn

    @_serialised
    def ens_accumulators(self):
        """torch views of the device accumulators ``(plane_sum [n_windows, wy * wx], count [n_windows])`` so that ranks can
        reduce them before the peak fit.  The engine orders its own calls on the accumulators whatever their streams; work
        the caller enqueues on these views (a collective on torch's current stream) is ordered by finishing on the same
        stream: :meth:`ens_finish_device`."""
        import torch

        ps, pc = ctypes.c_void_p(), ctypes.c_void_p()
        nf, nw = ctypes.c_longlong(), ctypes.c_longlong()
        self._check(self._lib.b2piv_ens_accum(self._h, ctypes.byref(ps), ctypes.byref(pc), ctypes.byref(nf), ctypes.byref(nw)), "b2piv_ens_accum")
        dev = torch.device("cuda", self.device)
        plane = torch.as_tensor(_CudaView(ps.value, (nf.value,)), device=dev)
        count = torch.as_tensor(_CudaView(pc.value, (nw.value,)), device=dev)
        return plane.view(nw.value, -1), count

    @_serialised
    def ens_finish(self, min_count: float):
        """Count filter + mean plane + peak fit; returns ``u, v, count`` each ``[n_windows]`` (numpy; synchronous, ordered
        after every earlier ``ens_add`` whatever stream it ran on)."""
        nr, nc = self._plan[1]
        u = np.empty(nr * nc, dtype=np.float32)
        v = np.empty(nr * nc, dtype=np.float32)
        cnt = np.empty(nr * nc, dtype=np.float32)
        self._check(self._lib.b2piv_ens_finish_host(self._h, float(min_count), u.ctypes.data, v.ctypes.data, cnt.ctypes.data), "b2piv_ens_finish_host")
        return u, v, cnt

    @_serialised
    def ens_finish_device(self, min_count: float, first_window: int = 0, n_windows: Optional[int] = None, stream=None):
        """Peak fit of the windows ``[first_window, first_window + n_windows)`` on torch's current stream (or ``stream``):
        returns CUDA tensors ``u, v`` ``[n_windows]``, no synchronisation.  After a reduce-scatter of the plane sums over the
        window axis each rank finishes its own slice (:func:`pyorc_b200.parallel.ensemble_reduce_finish`)."""
        import torch

        nr, nc = self._plan[1]
        if n_windows is None:
            n_windows = nr * nc - first_window
        dev = torch.device("cuda", self.device)
        st, ctx = _stream_ctx(torch, dev, stream)
        with ctx:
            out = torch.empty((2, max(int(n_windows), 0)), dtype=torch.float32, device=dev)
        self._check(self._lib.b2piv_ens_finish_device(self._h, float(min_count), int(first_window), int(n_windows),
                                                      out[0].data_ptr(), out[1].data_ptr(), st), "b2piv_ens_finish_device")
        return out[0], out[1]


_default_engines = {}


_default_lock = __import__("threading").Lock()


def get_engine(device: int = 0, slot: int = 0) -> Engine:
    """Process-wide engine per ``(device, slot)`` (created on first use).  An engine is used by one host thread at a time
    (the ABI's "one engine per thread and device"); ``slot > 0`` gives further engines on the same GPU -
    ``get_b2piv(devices=[0, 0])`` runs two of them side by side, each with its own streams and staging buffers."""
    key = (int(device), int(slot))
    with _default_lock:
        if key not in _default_engines:
            _default_engines[key] = Engine(int(device))
        return _default_engines[key]


def merge_ensembles(engines, min_count: float):
    """Finish an ensemble whose frame pairs were accumulated on several devices of ONE process (``get_b2piv(devices=...)``):
    the other engines' plane sums and counts are added to the first engine's accumulators (peer copies over NVLink, on its
    current torch stream) and the peak fit runs there, ordered behind the additions.  Returns numpy ``u, v, count``
    ``[n_windows]``.  (One process per GPU: :func:`pyorc_b200.parallel.ensemble_sharded`.)"""
    if len(engines) == 1:
        return engines[0].ens_finish(min_count)
    import torch

    e0 = engines[0]
    plane0, count0 = e0.ens_accumulators()
    with torch.cuda.device(e0.device):
        for e in engines[1:]:
            pk, ck = e.ens_accumulators()
            torch.cuda.synchronize(e.device)   # its ens_add_host calls have returned, i.e. are complete; belt and braces
            plane0.add_(pk.to(plane0.device))
            count0.add_(ck.to(count0.device))
        u, v = e0.ens_finish_device(min_count)
        return u.cpu().numpy(), v.cpu().numpy(), count0.cpu().numpy()
