"""Synthetic inputs of SURVEY.md section 8d, shared by bench.py, the measurement scripts and the parity tests."""
import math

import numpy as np


def synthetic_field(num_theta: int, num_phi: int, seed: int = 1234, h0_nm: float = 500.0, noise_nm: float = 5.0,
                    dtype=np.float64, row0: int = 0, rows=None, polar_noise: bool = False) -> np.ndarray:
    """h0(th,ph) = h0 [1 + 0.10 cos th + 0.05 sin^2 th sin 3ph] + noise u(th,ph) sin^2 th, u ~ U(-1,1).

    The noise of global row i comes from its own PCG64 stream seeded with (seed, i), so any latitude band
    [row0, row0+rows) of the global grid can be generated on its own and bands agree with the whole.
    polar_noise=True drops the sin^2 factor of the noise: the rings next to the poles are then rough, which is where
    the explicit phi-diffusion makes increments that plain f32 storage resolves even at dt = 0.25 dt_max of a large
    grid (the increment-parity tests use it)."""
    if rows is None:
        rows = num_theta - row0
    gi = np.arange(row0, row0 + rows)
    th = (gi.astype(np.float64) * (math.pi / (num_theta - 1)))[:, None]
    ph = (np.arange(num_phi, dtype=np.float64) * (2.0 * math.pi / num_phi))[None, :]
    s2 = np.sin(th) ** 2
    out = np.empty((rows, num_phi), dtype=dtype)
    base_phi = 0.05 * np.sin(3.0 * ph)
    for k, i in enumerate(gi):
        u = np.random.Generator(np.random.PCG64([seed, int(i)])).uniform(-1.0, 1.0, size=num_phi)
        row = h0_nm * 1e-9 * (1.0 + 0.10 * math.cos(th[k, 0]) + s2[k, 0] * base_phi[0]) + noise_nm * 1e-9 * u * (1.0 if polar_noise else s2[k, 0])
        out[k] = row
    return out


def dt_max(num_theta: int, num_phi: int, diffusion: float = 1e-9, radius: float = 0.025) -> float:
    """Forward-Euler stability bound of the diffusion term in the ring next to the pole
    (SURVEY.md section 7 hard part (ii)): dt <= 1 / (2 (a_theta + a_phi))."""
    dth = math.pi / (num_theta - 1)
    dph = 2.0 * math.pi / num_phi
    s1 = math.sin(dth)
    a_t = diffusion / (radius * radius * dth * dth)
    a_p = diffusion / (radius * radius * dph * dph * s1 * s1)
    return 1.0 / (2.0 * (a_t + a_p))
