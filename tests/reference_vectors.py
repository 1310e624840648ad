"""Reader / writer of the reference-vector files scripts/make_reference_vectors.rs produces (format in its header).

The writer exists so that the format itself can be tested without a Rust toolchain (it is fed by the ORACLE and its
files are never placed under tests/golden/); files under tests/golden/drainage_*.bin / interference_*.bin are
outputs of the real reference and turn the drainage half of the parity statement from "unpinned" to pinned.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

DRAINAGE_MAGIC = b"SBXVEC01"
INTERFERENCE_MAGIC = b"SBXINT01"


@dataclass
class DrainageVector:
    resolution: int
    num_theta: int
    num_phi: int
    film_thickness_nm: float
    diameter: float
    critical_thickness_nm: float
    density: float
    viscosity: float
    surfactant_diffusion: float
    gravity: float
    dt: float
    tfe_enabled: bool
    # (steps_done, current_time, active_tfe_count, field[num_theta, num_phi])
    checkpoints: List[Tuple[int, float, int, np.ndarray]] = field(default_factory=list)

    @property
    def initial(self) -> np.ndarray:
        """ThicknessField::new (drainage.rs:121-122): uniform film_thickness_nm * 1e-9 (config.rs:369)."""
        return np.full((self.num_theta, self.num_phi), self.film_thickness_nm * 1e-9)

    def oracle_kwargs(self) -> dict:
        return dict(h0_m=self.film_thickness_nm * 1e-9, density=self.density, gravity=self.gravity,
                    viscosity=self.viscosity, diffusion=self.surfactant_diffusion,
                    critical_thickness_m=self.critical_thickness_nm * 1e-9, radius_m=self.diameter / 2.0)


def read_vector(path: str) -> DrainageVector:
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != DRAINAGE_MAGIC:
        raise ValueError(f"{path}: not a drainage vector file")
    res, nt, nphi = struct.unpack_from("<3I", raw, 8)
    vals = struct.unpack_from("<8d", raw, 20)
    tfe, ncp = struct.unpack_from("<2I", raw, 84)
    v = DrainageVector(res, nt, nphi, *vals, bool(tfe))
    off = 92
    cells = nt * nphi
    for _ in range(ncp):
        steps, t, cnt, _pad = struct.unpack_from("<QdII", raw, off)
        off += 24
        fld = np.frombuffer(raw, dtype="<f8", count=cells, offset=off).reshape(nt, nphi).copy()
        off += 8 * cells
        v.checkpoints.append((steps, t, cnt, fld))
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} trailing bytes")
    return v


def write_vector(path: str, v: DrainageVector) -> None:
    with open(path, "wb") as f:
        f.write(DRAINAGE_MAGIC)
        f.write(struct.pack("<3I", v.resolution, v.num_theta, v.num_phi))
        f.write(struct.pack("<8d", v.film_thickness_nm, v.diameter, v.critical_thickness_nm, v.density, v.viscosity,
                            v.surfactant_diffusion, v.gravity, v.dt))
        f.write(struct.pack("<2I", int(v.tfe_enabled), len(v.checkpoints)))
        for steps, t, cnt, fld in v.checkpoints:
            f.write(struct.pack("<QdII", steps, t, cnt, 0))
            f.write(np.ascontiguousarray(fld, dtype="<f8").tobytes())


def read_interference(path: str):
    """-> (n_film, points[n, 8]): d_nm, cos_i, gamma rgb, linear rgb."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != INTERFERENCE_MAGIC:
        raise ValueError(f"{path}: not an interference vector file")
    (n_film,) = struct.unpack_from("<d", raw, 8)
    n, _pad = struct.unpack_from("<2I", raw, 16)
    pts = np.frombuffer(raw, dtype="<f8", count=8 * n, offset=24).reshape(n, 8).copy()
    return n_film, pts


def write_interference(path: str, n_film: float, pts: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(INTERFERENCE_MAGIC)
        f.write(struct.pack("<d", n_film))
        f.write(struct.pack("<2I", len(pts), 0))
        f.write(np.ascontiguousarray(pts, dtype="<f8").tobytes())
