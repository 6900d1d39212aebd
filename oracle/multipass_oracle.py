"""TEST INFRASTRUCTURE ONLY - CPU statement of the two-pass PIV of BASELINE.json configs[2] ("2-pass deform").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

PARITY UNPINNED - there is NO reference implementation: pyorc's engine (ffpiv) is single pass (pyorc/velocimetry/ffpiv.py:
446-474 calls ``cross_corr`` once per chunk).  The algorithm is therefore DEFINED here from the classical
multi-pass scheme with a discrete window offset (Westerweel, Dabiri & Gharib 1997; Scarano & Riethmuller 1999), every
correlation being ffpiv's own (oracle/ffpiv_oracle.py):

  1. pass 1: ffpiv correlation at the coarse window / overlap -> (u1, v1) per coarse window, px / frame;
  2. validation: universal outlier detection on the 3 x 3 neighbourhood (Westerweel & Scarano 2005): with the median m of the
     valid neighbours (centre excluded) and r = median |neighbour - m|, a vector is replaced by m when it is NaN or
     |value - m| / (r + 0.1) > 2 for u or for v; a vector without valid neighbours is kept (NaN -> 0);
  3. predictor: bilinear interpolation of the validated field at the centres of the fine windows (coarse centres
     r * stride + size / 2; positions outside the coarse grid take the edge value), rounded half-to-even to whole pixels and
     clamped so that the shifted window stays inside the frame;
  4. pass 2: ffpiv correlation of the fine window of frame k at (y0, x0) with the window of frame k+1 at (y0 + dy, x0 + dx);
  5. result: u = dx + u2, v = dy + v2; corr / s2n are those of pass 2.

The accuracy claim is checked against the IMPOSED synthetic displacement field, not against another implementation.

WINDOW DEFORMATION (``two_pass_deform`` - the "deform" of configs[2]; Scarano 2002, one-sided image deformation):
  3'. the validated field is interpolated bilinearly to EVERY pixel centre (y + 0.5, x + 0.5) (edge values outside the coarse
      grid) and frame k+1 is resampled at (y + dv, x + du): bilinear between the four neighbours, sample coordinates clamped
      to the frame, float64 arithmetic, the result rounded once to float32 -> B'_k;
  4'. pass 2: ffpiv correlation of the fine windows of frame k (as float32) with the SAME windows of B'_k;
  5'. result: the un-rounded predictor at the fine window centre (float32) + the residual of pass 2.
"""
from __future__ import annotations

import numpy as np

from oracle import ffpiv_oracle as O

EPS_MEDIAN = 0.1
THRESHOLD = 2.0


def validate(u, v):
    """Step 2 on [n_pairs, rows, cols] float arrays (float64 arithmetic)."""
    u, v = np.asarray(u, np.float64), np.asarray(v, np.float64)
    P, R, C = u.shape
    ou, ov = u.copy(), v.copy()
    for k in range(P):
        for r in range(R):
            for c in range(C):
                nu, nv = [], []
                for dr in (-1, 0, 1):
                    for dc in (-1, 0, 1):
                        rr, cc = r + dr, c + dc
                        if (dr or dc) and 0 <= rr < R and 0 <= cc < C and np.isfinite(u[k, rr, cc]) and np.isfinite(v[k, rr, cc]):
                            nu.append(u[k, rr, cc])
                            nv.append(v[k, rr, cc])
                bad = not (np.isfinite(u[k, r, c]) and np.isfinite(v[k, r, c]))
                if not nu:
                    if bad:
                        ou[k, r, c], ov[k, r, c] = 0.0, 0.0
                    continue
                mu, mv = lower_median(nu), lower_median(nv)
                ru = lower_median([abs(x - mu) for x in nu]) + EPS_MEDIAN
                rv = lower_median([abs(x - mv) for x in nv]) + EPS_MEDIAN
                if bad or abs(u[k, r, c] - mu) / ru > THRESHOLD or abs(v[k, r, c] - mv) / rv > THRESHOLD:
                    ou[k, r, c], ov[k, r, c] = mu, mv
    return ou, ov


def lower_median(vals):
    """Median as the element of rank (n - 1) // 2 of the sorted values (no averaging: the result is one of the samples)."""
    s = sorted(vals)
    return s[(len(s) - 1) // 2]


def predictor(u, v, dim_size, coarse, fine):
    """Step 3.  `coarse` / `fine` = (window_size, overlap).  Returns integer shifts dy, dx [n_pairs, rows2, cols2]."""
    (ws1, ov1), (ws2, ov2) = coarse, fine
    H, W = dim_size
    y1, x1 = O.window_origins(dim_size, ws1, ov1)
    y2, x2 = O.window_origins(dim_size, ws2, ov2)
    sy1, sx1 = ws1[0] - ov1[0], ws1[1] - ov1[1]
    R, C = len(y1), len(x1)
    # fractional coarse index of every fine centre, clamped to the coarse grid
    fy = np.clip(((y2 + ws2[0] / 2.0) - ws1[0] / 2.0) / sy1, 0.0, R - 1.0)
    fx = np.clip(((x2 + ws2[1] / 2.0) - ws1[1] / 2.0) / sx1, 0.0, C - 1.0)
    iy = np.minimum(np.floor(fy).astype(int), max(R - 2, 0))
    ix = np.minimum(np.floor(fx).astype(int), max(C - 2, 0))
    ty, tx = fy - iy, fx - ix
    iy1, ix1 = np.minimum(iy + 1, R - 1), np.minimum(ix + 1, C - 1)

    def interp(f):
        f00 = f[:, iy][:, :, ix]
        f01 = f[:, iy][:, :, ix1]
        f10 = f[:, iy1][:, :, ix]
        f11 = f[:, iy1][:, :, ix1]
        top = f00 + (f01 - f00) * tx[None, None, :]
        bot = f10 + (f11 - f10) * tx[None, None, :]
        return top + (bot - top) * ty[None, :, None]

    du, dv = interp(u), interp(v)
    dx = np.rint(du).astype(np.int64)
    dy = np.rint(dv).astype(np.int64)
    dy = np.clip(dy, -y2[None, :, None], (H - ws2[0] - y2)[None, :, None])
    dx = np.clip(dx, -x2[None, None, :], (W - ws2[1] - x2)[None, None, :])
    return dy, dx


def shifted_pass(imgs, dy, dx, window_size, overlap):
    """Steps 4 and 5: ffpiv's correlation with frame k+1's window displaced by (dy, dx)."""
    imgs = np.asarray(imgs)
    n = imgs.shape[0]
    y0, x0 = O.window_origins(imgs.shape[-2:], window_size, overlap)
    wy, wx = window_size
    R, C = len(y0), len(x0)
    u = np.empty((n - 1, R, C), np.float64)
    v, cm, sn = np.empty_like(u), np.empty_like(u), np.empty_like(u)
    for k in range(n - 1):
        wa = np.stack([imgs[k, y:y + wy, x:x + wx] for y in y0 for x in x0])
        wb = np.stack([imgs[k + 1, y + dy[k, r, c]:y + dy[k, r, c] + wy, x + dx[k, r, c]:x + dx[k, r, c] + wx]
                       for r, y in enumerate(y0) for c, x in enumerate(x0)])
        corr = O.ncc(wa, wb).astype(np.float32)
        cmax = corr.max(axis=(-1, -2))
        with np.errstate(all="ignore"):
            s2n = cmax / corr.mean(axis=(-1, -2))
        pk = O.peak_position(corr)
        v[k] = (pk[:, 0] - wy // 2).reshape(R, C) + dy[k]
        u[k] = (pk[:, 1] - wx // 2).reshape(R, C) + dx[k]
        cm[k] = cmax.reshape(R, C)
        sn[k] = s2n.reshape(R, C)
    return u, v, cm, sn


def two_pass(imgs, coarse, fine):
    """The whole scheme: returns (u, v, corr_max, s2n) on the fine grid plus the integer predictor (dy, dx)."""
    imgs = np.asarray(imgs)
    (ws1, ov1), (ws2, ov2) = coarse, fine
    nr1, nc1 = O.get_array_shape(imgs.shape[-2:], ws1, ov1)
    u1, v1, _, _ = O.uv_timestep(imgs, nc1, nr1, ws1, ov1)
    u1, v1 = validate(u1, v1)
    dy, dx = predictor(u1, v1, imgs.shape[-2:], coarse, fine)
    u, v, cm, sn = shifted_pass(imgs, dy, dx, ws2, ov2)
    return u, v, cm, sn, dy, dx


# ---- window deformation ------------------------------------------------------------------------------------------------------
def _cell(pos, w1, s1, n1):
    f = np.clip((pos - w1 / 2.0) / s1, 0.0, n1 - 1.0)
    i = np.minimum(np.floor(f).astype(int), max(n1 - 2, 0))
    return i, np.minimum(i + 1, n1 - 1), f - i


def _bilin(f, iy, iy1, ix, ix1, ty, tx):
    """f [P, R, C] at cells (iy, ix) with weights (ty [ny], tx [nx]) -> [P, ny, nx]; operation order of the device kernel."""
    f00 = f[:, iy][:, :, ix]
    f01 = f[:, iy][:, :, ix1]
    f10 = f[:, iy1][:, :, ix]
    f11 = f[:, iy1][:, :, ix1]
    top = f00 + (f01 - f00) * tx[None, None, :]
    bot = f10 + (f11 - f10) * tx[None, None, :]
    return top + (bot - top) * ty[None, :, None]


def deform(imgs, u, v, coarse):
    """Step 3': interleaved float32 stack [2 (n - 1), H, W] = (frame k, frame k+1 resampled with the per-pixel predictor)."""
    imgs = np.asarray(imgs)
    n, H, W = imgs.shape
    (ws1, ov1) = coarse
    R, C = u.shape[1:]
    iy, iy1, ty = _cell(np.arange(H) + 0.5, ws1[0], ws1[0] - ov1[0], R)
    ix, ix1, tx = _cell(np.arange(W) + 0.5, ws1[1], ws1[1] - ov1[1], C)
    du = _bilin(np.asarray(u, np.float64), iy, iy1, ix, ix1, ty, tx)
    dv = _bilin(np.asarray(v, np.float64), iy, iy1, ix, ix1, ty, tx)
    yy = np.clip(np.arange(H, dtype=np.float64)[None, :, None] + dv, 0.0, H - 1.0)
    xx = np.clip(np.arange(W, dtype=np.float64)[None, None, :] + du, 0.0, W - 1.0)
    y0 = np.minimum(np.floor(yy).astype(int), max(H - 2, 0))
    x0 = np.minimum(np.floor(xx).astype(int), max(W - 2, 0))
    y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
    wy, wx = yy - y0, xx - x0
    stack = np.empty((2 * (n - 1), H, W), np.float32)
    for k in range(n - 1):
        fb = imgs[k + 1].astype(np.float64)
        b00, b01, b10, b11 = fb[y0[k], x0[k]], fb[y0[k], x1[k]], fb[y1[k], x0[k]], fb[y1[k], x1[k]]
        top = b00 + (b01 - b00) * wx[k]
        bot = b10 + (b11 - b10) * wx[k]
        stack[2 * k] = imgs[k].astype(np.float32)
        stack[2 * k + 1] = (top + (bot - top) * wy[k]).astype(np.float32)
    return stack


def predictor_float(u, v, dim_size, coarse, fine):
    """Un-rounded predictor (dv, du) float32 [P, rows2, cols2] at the fine window centres."""
    (ws1, ov1), (ws2, ov2) = coarse, fine
    y2, x2 = O.window_origins(dim_size, ws2, ov2)
    R, C = u.shape[1:]
    iy, iy1, ty = _cell(y2 + ws2[0] / 2.0, ws1[0], ws1[0] - ov1[0], R)
    ix, ix1, tx = _cell(x2 + ws2[1] / 2.0, ws1[1], ws1[1] - ov1[1], C)
    du = _bilin(np.asarray(u, np.float64), iy, iy1, ix, ix1, ty, tx).astype(np.float32)
    dv = _bilin(np.asarray(v, np.float64), iy, iy1, ix, ix1, ty, tx).astype(np.float32)
    return dv, du


def two_pass_deform(imgs, coarse, fine):
    """Two-pass PIV with window deformation: returns (u, v, corr_max, s2n) on the fine grid, the stack and the predictor."""
    imgs = np.asarray(imgs)
    (ws1, ov1), (ws2, ov2) = coarse, fine
    nr1, nc1 = O.get_array_shape(imgs.shape[-2:], ws1, ov1)
    nr2, nc2 = O.get_array_shape(imgs.shape[-2:], ws2, ov2)
    u1, v1, _, _ = O.uv_timestep(imgs, nc1, nr1, ws1, ov1)
    u1, v1 = validate(u1, v1)
    stack = deform(imgs, u1, v1, coarse)
    dv, du = predictor_float(u1, v1, imgs.shape[-2:], coarse, fine)
    P = imgs.shape[0] - 1
    u = np.empty((P, nr2, nc2), np.float64)
    v, cm, sn = np.empty_like(u), np.empty_like(u), np.empty_like(u)
    for k in range(P):
        uk, vk, ck, sk = O.uv_timestep(stack[2 * k : 2 * k + 2], nc2, nr2, ws2, ov2)
        u[k], v[k], cm[k], sn[k] = uk[0] + du[k], vk[0] + dv[k], ck[0], sk[0]
    return u, v, cm, sn, stack, (dv, du)
