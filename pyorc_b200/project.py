"""Orthoprojection of camera frames onto the PIV grid on the GPU - the counterpart of ``pyorc.project.project_numpy``
(pyorc/project.py:160-230) and its per-frame worker ``img_to_ortho`` (project.py:123-157, numba ``_group_average``
:19-53), so that decoded frames can go decode -> project -> pre-process -> ``get_piv`` without leaving HBM.

The index maps are NOT computed here: they are what ``CameraConfig.map_idx_img_ortho`` and
``CameraConfig.map_mean_idx_img_ortho`` return (pyorc/api/cameraconfig.py:739-860; camera geometry, out of scope), once
per camera configuration and water level.  :class:`OrthoProjector` merges them into per-target-pixel gather lists on the
device (``b2piv_project_plan``); projecting a stack of frames is then one kernel launch.

* :func:`img_to_ortho`   same name, arguments and float result as the reference worker (one image)
* :func:`project_numpy`  the ``[time, y, x]`` stack version with the reference's result dtype (= input dtype; uint8
  truncates the float means exactly like ``apply_ufunc(..., output_dtypes=[da.dtype])``, project.py:205-227)

numpy in -> numpy out; torch CUDA tensor in -> torch CUDA tensor out (stream-ordered).  No CPU fallback.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from .engine import B2PIV_F32, B2PIV_U8, Engine, get_engine

__all__ = ["OrthoProjector", "img_to_ortho", "project_numpy"]


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.int64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p) if a.size else None


class OrthoProjector:
    """Index maps of one camera configuration, resident on the device.

    Parameters mirror ``img_to_ortho`` (project.py:123-145): ``idx_img`` / ``idx_ortho`` are the nearest-neighbour map
    (``idx_ortho`` either the reference's boolean mask over the flattened target grid or integer positions),
    ``src_idx`` / ``uidx`` / ``norm_idx`` the group-mean map (all ``None`` for ``reducer != "mean"``).
    """

    def __init__(self, img_shape, ortho_shape, idx_img, idx_ortho, src_idx=None, uidx=None, norm_idx=None, device: int = 0,
                 engine: Optional[Engine] = None):
        self.engine = engine if engine is not None else get_engine(device)
        self.img_shape = (int(img_shape[0]), int(img_shape[1]))
        self.ortho_shape = (int(ortho_shape[0]), int(ortho_shape[1]))
        idx_ortho = np.asarray(idx_ortho)
        if idx_ortho.dtype == np.bool_:
            if idx_ortho.size != self.ortho_shape[0] * self.ortho_shape[1]:
                raise ValueError("boolean idx_ortho must cover the flattened target grid")
            idx_ortho = np.flatnonzero(idx_ortho.ravel())
        nn_img, nn_ortho = _i64(idx_img), _i64(idx_ortho)
        if nn_img.size != nn_ortho.size:
            raise ValueError("idx_img and idx_ortho select a different number of pixels")
        if (src_idx is None) != (norm_idx is None) or (src_idx is None) != (uidx is None):
            raise ValueError("src_idx, uidx and norm_idx go together")
        s, g, u = (_i64(src_idx), _i64(norm_idx), _i64(uidx)) if src_idx is not None else (_i64([]),) * 3
        if s.size != g.size:
            raise ValueError("src_idx and norm_idx must have the same length")
        eng = self.engine
        eng._check(eng._lib.b2piv_project_plan(eng._h, self.img_shape[0], self.img_shape[1], self.ortho_shape[0], self.ortho_shape[1],
                                               _ptr(nn_img), _ptr(nn_ortho), nn_img.size, _ptr(s), _ptr(g), s.size, _ptr(u), u.size),
                   "b2piv_project_plan")
        self._token = object()
        eng._project_owner = self._token       # one plan per engine: re-plan if another projector took it over
        self._maps = (nn_img, nn_ortho, s, g, u)

    def _ensure_plan(self):
        eng = self.engine
        if getattr(eng, "_project_owner", None) is not self._token:
            nn_img, nn_ortho, s, g, u = self._maps
            eng._check(eng._lib.b2piv_project_plan(eng._h, self.img_shape[0], self.img_shape[1], self.ortho_shape[0], self.ortho_shape[1],
                                                   _ptr(nn_img), _ptr(nn_ortho), nn_img.size, _ptr(s), _ptr(g), s.size, _ptr(u), u.size),
                       "b2piv_project_plan")
            eng._project_owner = self._token

    def __call__(self, frames, out_float: bool = False):
        """Project ``[time, y, x]`` (or one ``[y, x]``) frames; result dtype = input dtype, or float32 if ``out_float``."""
        import torch

        eng = self.engine
        self._ensure_plan()
        if type(frames).__module__.startswith("torch"):
            if not frames.is_cuda:
                raise TypeError("torch frames must be CUDA tensors (pass numpy for host data)")
            t, was_np = frames, False
        else:
            a = np.asarray(frames)
            if a.dtype not in (np.uint8, np.float32):
                a = a.astype(np.float32)
            t, was_np = torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{eng.device}"), True
        single = t.dim() == 2
        if single:
            t = t[None]
        if t.dim() != 3 or tuple(t.shape[1:]) != self.img_shape:
            raise ValueError(f"frames must be [time, {self.img_shape[0]}, {self.img_shape[1]}], got {tuple(t.shape)}")
        if t.dtype not in (torch.uint8, torch.float32):
            t = t.float()
        t = t.contiguous()
        code = B2PIV_U8 if t.dtype == torch.uint8 else B2PIV_F32
        out_dtype, out_code = (torch.float32, B2PIV_F32) if out_float else (t.dtype, code)
        out = torch.empty((t.shape[0],) + self.ortho_shape, dtype=out_dtype, device=t.device)
        eng._check(eng._lib.b2piv_project_device(eng._h, t.data_ptr(), code, int(t.shape[0]), out.data_ptr(), out_code,
                                                 torch.cuda.current_stream(t.device).cuda_stream), "b2piv_project_device")
        if single:
            out = out[0]
        return out.cpu().numpy() if was_np else out


def img_to_ortho(img, x, y, idx_img, idx_ortho, src_idx=None, uidx=None, norm_idx=None, device: int = 0):
    """One image -> ortho grid ``[len(y), len(x)]`` of float means / nearest values, 0 where nothing maps
    (project.py:123-157).  The reference returns float64 holding float32 values; this returns the float32 values."""
    img = np.asarray(img)
    proj = OrthoProjector(img.shape, (len(y), len(x)), idx_img, idx_ortho, src_idx, uidx, norm_idx, device=device)
    return proj(img, out_float=True)


def project_numpy(frames, x, y, idx_img, idx_ortho, src_idx=None, uidx=None, norm_idx=None, device: int = 0):
    """``project_numpy`` on a ``[time, y, x]`` stack with the maps already evaluated (``cc.map_idx_img_ortho(x, y, z)``,
    ``cc.map_mean_idx_img_ortho(x, y, z)`` - project.py:196-201); result dtype = input dtype (project.py:222)."""
    shape = tuple(frames.shape[-2:])
    proj = OrthoProjector(shape, (len(y), len(x)), idx_img, idx_ortho, src_idx, uidx, norm_idx, device=device)
    return proj(frames)
