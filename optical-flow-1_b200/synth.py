"""Deterministic synthetic frame pairs.

I0 is a band-limited multi-octave sinusoid texture; I1 is the same analytic texture evaluated
at displaced coordinates: the background moves by (+2.3, +1.4) px, a central disc of radius
ny/4 moves by (-3.3, +2.6) px.  Motions are non-integer on purpose: integer motions park
samples exactly on the warp-validity edge (src/bicubic_interpolation.cpp:214-215), where even
the reference's float and double builds disagree.

Works on numpy arrays (tests, fixtures) and on torch tensors (bench: 256 x 1080p pairs are
generated directly in HBM).  Values are float32 in about [0, 255].
"""
import math

import numpy as np

N_TERMS = 24
BG_MOTION = (2.3, 1.4)
DISC_MOTION = (-3.3, 2.6)


def texture_params(seed):
    """24 sinusoids: 8 octaves (0.02 * 1.25^m cycles/px) x 3 random orientations/phases."""
    rs = np.random.RandomState(seed)  # MT19937
    theta = rs.uniform(0.0, math.pi, N_TERMS)
    phase = rs.uniform(0.0, 2.0 * math.pi, N_TERMS)
    m = np.arange(N_TERMS) // 3
    freq = 0.02 * 1.25 ** m
    amp = 30.0 / (1.0 + 0.3 * m)
    return freq * np.cos(theta), freq * np.sin(theta), phase, amp


def _texture(xp, X, Y, seed):
    fx, fy, ph, amp = texture_params(seed)
    out = None
    for k in range(N_TERMS):
        t = xp.sin((2.0 * math.pi * fx[k]) * X + (2.0 * math.pi * fy[k]) * Y + ph[k]) * amp[k]
        out = t if out is None else out + t
    return out + 128.0


def make_pair(nx, ny, seed=1234, scale=1.0):
    """numpy float32 (I0, I1), shape (ny, nx).  `scale` multiplies both motions (small test
    images use scale < 1 so that the motion stays inside the pyramid's capture range)."""
    Y, X = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64),
                       indexing="ij")
    I0 = _texture(np, X, Y, seed)
    disc = (X - 0.5 * nx) ** 2 + (Y - 0.5 * ny) ** 2 < (0.25 * ny) ** 2
    dx = np.where(disc, DISC_MOTION[0], BG_MOTION[0]) * scale
    dy = np.where(disc, DISC_MOTION[1], BG_MOTION[1]) * scale
    I1 = _texture(np, X - dx, Y - dy, seed)
    return I0.astype(np.float32), I1.astype(np.float32)


def make_triple(nx, ny, seed=1234, scale=1.0):
    """numpy float32 (I_1, I0, I1): the texture one step back, now and one step ahead along the same
    two motions -- the three frames src/tvl1occflow.cpp takes (previous, source, target)."""
    Y, X = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64),
                       indexing="ij")
    disc = (X - 0.5 * nx) ** 2 + (Y - 0.5 * ny) ** 2 < (0.25 * ny) ** 2
    dx = np.where(disc, DISC_MOTION[0], BG_MOTION[0]) * scale
    dy = np.where(disc, DISC_MOTION[1], BG_MOTION[1]) * scale
    I_1 = _texture(np, X + dx, Y + dy, seed)
    I0 = _texture(np, X, Y, seed)
    I1 = _texture(np, X - dx, Y - dy, seed)
    return I_1.astype(np.float32), I0.astype(np.float32), I1.astype(np.float32)


def make_batch_torch(npairs, nx, ny, seed=1234, device="cuda"):
    """torch float32 (I0[npairs, ny, nx], I1[...]) on `device`; pair b uses seed + b."""
    import torch
    ys = torch.arange(ny, dtype=torch.float32, device=device)
    xs = torch.arange(nx, dtype=torch.float32, device=device)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    disc = (X - 0.5 * nx) ** 2 + (Y - 0.5 * ny) ** 2 < (0.25 * ny) ** 2
    dx = torch.where(disc, torch.tensor(DISC_MOTION[0], device=device), torch.tensor(BG_MOTION[0], device=device))
    dy = torch.where(disc, torch.tensor(DISC_MOTION[1], device=device), torch.tensor(BG_MOTION[1], device=device))
    I0 = torch.empty((npairs, ny, nx), dtype=torch.float32, device=device)
    I1 = torch.empty_like(I0)
    for b in range(npairs):
        I0[b] = _texture(torch, X, Y, seed + b)
        I1[b] = _texture(torch, X - dx, Y - dy, seed + b)
    return I0, I1
