"""One fused step configuration for ncu: python scripts/profile_mode.py <RGB_MODE> [PREC]  (8192x16384, 6 steps)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import soap_bubble_simulation_b200 as S  # noqa: E402
from soap_bubble_simulation_b200.synth import dt_max, synthetic_field  # noqa: E402

mode = getattr(S, sys.argv[1] if len(sys.argv) > 1 else "RGB_EXACT3")
prec = getattr(S, sys.argv[2] if len(sys.argv) > 2 else "PREC_F32")
nt, nphi = 8192, 16384
sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, num_theta=nt, num_phi=nphi, h_ref=0.0)
sim.upload(synthetic_field(nt, nphi, dtype=np.float32))
sim.step_fused(0.25 * dt_max(nt, nphi), S.View(mode=mode, format=S.OUT_RGBA8), 6)
sim.sync()
