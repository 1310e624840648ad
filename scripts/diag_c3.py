#!/usr/bin/env python
"""Diagnostic: repeat the C3 fused step from fresh handles, compare the runs bit for bit and against the oracle,
and say where any difference sits (row, column, band, strip)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import soap_bubble_simulation_b200 as S  # noqa: E402
from soap_bubble_simulation_b200.synth import dt_max, synthetic_field  # noqa: E402

nt, nphi = 8192, 16384
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 6
f = synthetic_field(nt, nphi, dtype=np.float32)
dt = 0.25 * dt_max(nt, nphi)
view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
ora = O.OracleDrainage(nt, nphi, threads=max(1, len(os.sched_getaffinity(0))))
ora.data[:] = f
ora.step(dt, 1)


def where(mask, tag):
    r, c = np.nonzero(mask)
    if len(r) == 0:
        print(f"  {tag}: none")
        return
    print(f"  {tag}: {len(r)} cells, rows {r.min()}..{r.max()} ({len(np.unique(r))} distinct: {np.unique(r)[:12]}), "
          f"cols {c.min()}..{c.max()} ({len(np.unique(c))} distinct; first {np.unique(c)[:8]}), strips {np.unique(c // 1024)[:16]}")


first = None
for t in range(trials):
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=nt, num_phi=nphi, h_ref=0.0)
    sim.upload(f)
    sim.step_fused(dt, view, 1)
    got = sim.download(np.float32)[0]
    px = sim.download_rgb()[0]
    rel = np.abs(got.astype(np.float64) / ora.data - 1)
    print(f"trial {t}: max rel err vs oracle {rel.max():.3e}", flush=True)
    where(rel > 1e-5, "rel > 1e-5")
    if rel.max() > 1e-5:
        r, c = np.unravel_index(np.argmax(rel), rel.shape)
        print(f"  worst cell ({r},{c}): got {got[r, c]:.9e} oracle {ora.data[r, c]:.9e} input {f[r, c]:.9e}; "
              f"input neighbours up {f[r - 1, c]:.9e} dn {f[min(r + 1, nt - 1), c]:.9e}")
    if first is None:
        first, first_px = got.copy(), px.copy()
    else:
        where(got != first, "differs from trial 0 (h)")
        where((px != first_px).any(axis=-1), "differs from trial 0 (rgba)")
    # a second step on the same handle
    sim.step_fused(dt, view, 1)
    sim.close()
