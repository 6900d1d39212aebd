"""Seeded synthetic particle-image frame stacks for parity tests and bench.py.

``particle_frames`` renders on the host with numpy (small sizes, tests); ``particle_frames_torch`` renders on the
device from (seed, frame index) so 4K/8K shards never exist on the host."""

from __future__ import annotations

import numpy as np

SEED = 20260925


def displacement_field(H: int, W: int, y, x):
    """Imposed displacement in px/frame at positions (y, x): dx = 3.30 + 1.5 sin(2 pi y/H), dy = -1.70 + 0.8 cos(2 pi x/W)."""
    dx = 3.30 + 1.5 * np.sin(2 * np.pi * y / H)
    dy = -1.70 + 0.8 * np.cos(2 * np.pi * x / W)
    return dx, dy


def particle_frames(n_frames: int, H: int, W: int, dtype=np.uint8, seed: int = SEED, density: float = 0.02, sigma: float = 1.2) -> np.ndarray:
    """``[n_frames, H, W]`` Gaussian-blob particle images advected by :func:`displacement_field`."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    n = int(density * H * W)
    px = rng.uniform(0, W, n)
    py = rng.uniform(0, H, n)
    amp = rng.uniform(120, 255, n)
    out = np.empty((n_frames, H, W), dtype=np.float32)
    r = 4
    oy, ox = np.mgrid[-r:r + 1, -r:r + 1]
    for k in range(n_frames):
        img = np.full((H + 2 * r, W + 2 * r), 10.0, dtype=np.float64)
        iy = np.floor(py).astype(np.int64)
        ix = np.floor(px).astype(np.int64)
        fy = (py - iy)[:, None, None]
        fx = (px - ix)[:, None, None]
        blob = amp[:, None, None] * np.exp(-((oy[None] - fy) ** 2 + (ox[None] - fx) ** 2) / (2 * sigma**2))
        yy = (iy[:, None, None] + oy[None] + r).ravel()
        xx = (ix[:, None, None] + ox[None] + r).ravel()
        np.add.at(img, (yy, xx), blob.ravel())
        img = img[r:-r, r:-r] + rng.normal(0, 2.0, (H, W))
        out[k] = np.clip(img, 0, 255)
        dx, dy = displacement_field(H, W, py, px)
        px = (px + dx) % W
        py = (py + dy) % H
        # 2 % drop-in / drop-out
        swap = rng.random(n) < 0.02
        ns = int(swap.sum())
        px[swap] = rng.uniform(0, W, ns)
        py[swap] = rng.uniform(0, H, ns)
    if np.dtype(dtype) == np.uint8:
        return np.floor(out).astype(np.uint8)
    return out.astype(dtype)


def particle_frames_torch(n_frames: int, H: int, W: int, device, dtype="uint8", seed: int = SEED, first_frame: int = 0,
                          density: float = 0.02, sigma: float = 1.2):
    """Device-side generator: frame ``first_frame + k`` depends only on ``(seed, first_frame + k)`` so a rank can
    render its own shard (particles advance by a constant-in-time field; positions are closed-form in k for the
    uniform part and integrated stepwise for the sinusoidal part)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = int(density * H * W)
    px0 = torch.rand(n, generator=g, device=device) * W
    py0 = torch.rand(n, generator=g, device=device) * H
    amp = 120 + 135 * torch.rand(n, generator=g, device=device)
    frames = torch.empty((n_frames, H, W), dtype=torch.uint8 if dtype == "uint8" else torch.float32, device=device)
    r = 3
    offs = torch.arange(-r, r + 1, device=device)
    oy, ox = torch.meshgrid(offs, offs, indexing="ij")
    oy = oy.reshape(1, -1)
    ox = ox.reshape(1, -1)
    px, py = px0.clone(), py0.clone()
    two_pi = 2 * np.pi
    for k in range(first_frame + n_frames):
        if k >= first_frame:
            img = torch.full((H * W,), 10.0, device=device)
            iy = torch.floor(py).long()
            ix = torch.floor(px).long()
            fy = (py - iy).unsqueeze(1)
            fx = (px - ix).unsqueeze(1)
            blob = amp.unsqueeze(1) * torch.exp(-((oy - fy) ** 2 + (ox - fx) ** 2) / (2 * sigma**2))
            yy = (iy.unsqueeze(1) + oy) % H
            xx = (ix.unsqueeze(1) + ox) % W
            img.index_add_(0, (yy * W + xx).reshape(-1), blob.reshape(-1))
            gn = torch.Generator(device=device)
            gn.manual_seed(seed * 1000003 + k)
            img = img + 2.0 * torch.randn(H * W, generator=gn, device=device)
            img = img.clamp_(0, 255).reshape(H, W)
            frames[k - first_frame] = img.to(frames.dtype)
        dx = 3.30 + 1.5 * torch.sin(two_pi * py / H)
        dy = -1.70 + 0.8 * torch.cos(two_pi * px / W)
        px = (px + dx) % W
        py = (py + dy) % H
    return frames
