"""TEST INFRASTRUCTURE ONLY - CPU restatement of pyorc's index-map orthoprojection, the checker for
pyorc_b200/csrc/project.cuh.  Nothing in the product path may import this module.

Follows pyorc @ be7d7c8:
  img_to_ortho      pyorc/project.py:123-157   nearest scatter-assign, then group means written over it
  _group_average    pyorc/project.py:19-53     float32 sums accumulated in sample order, int64 counts, sum / count
  result dtype      pyorc/project.py:205-227   apply_ufunc(vectorize=True, output_dtypes=[da.dtype]) -> astype(dtype)

Pinned: tests/test_golden.py::test_project_oracle_reproduces_reference compares it with the output of the reference's
own ``img_to_ortho`` on the Ngwerere frame and camera configuration (tests/golden/make_ngwerere_golden.py).
"""
import numpy as np


def group_average(data: np.ndarray, idx: np.ndarray, num_groups: int) -> np.ndarray:
    """project.py:19-53.  np.add.at is unbuffered and applies the additions in index order, i.e. the reference's loop."""
    sums = np.zeros(num_groups, dtype=np.float32)
    counts = np.zeros(num_groups, dtype=np.int64)
    np.add.at(sums, idx, data.astype(np.float32))
    np.add.at(counts, idx, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (sums.astype(np.float64) / counts).astype(np.float32)   # numba: float32 / int64 -> float64, stored as float32


def img_to_ortho(img, x, y, idx_img, idx_ortho, src_idx=None, uidx=None, norm_idx=None) -> np.ndarray:
    """project.py:123-157; float64 array [len(y), len(x)] like the reference."""
    flat = np.float32(np.asarray(img).flatten())
    new_arr = np.zeros(len(y) * len(x))
    new_arr[idx_ortho] = flat[idx_img]
    if src_idx is not None:
        new_arr[uidx] = group_average(flat[src_idx], np.asarray(norm_idx), len(uidx))
    return new_arr.reshape(len(y), -1)


def project_stack(frames, x, y, idx_img, idx_ortho, src_idx=None, uidx=None, norm_idx=None) -> np.ndarray:
    """project_numpy's result for a [time, y, x] stack: per-frame img_to_ortho cast to the input dtype."""
    frames = np.asarray(frames)
    out = np.stack([img_to_ortho(f, x, y, idx_img, idx_ortho, src_idx, uidx, norm_idx) for f in frames])
    return np.nan_to_num(out, nan=0.0).astype(frames.dtype)     # Frames.project: .fillna(0.0), frames.py:263
