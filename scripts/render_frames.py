#!/usr/bin/env python
"""Human-checkable artefacts of rows f3/f4: the reference's default headless scene rendered by k_bubble_frame, and the
same bubble coloured from a drainage simulator's thickness field, written with sbx_export_frame_png."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import soap_bubble_simulation_b200 as S  # noqa: E402
from soap_bubble_simulation_b200.synth import synthetic_field  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
p = S.HeadlessRenderPipeline.new(256, 256)
p.set_thickness(500.0)
p.set_time(0.0)
p.set_camera_distance(0.2)
S.export_frame(os.path.join(out, "frame_default_bubble.png"), 256, 256, p.render_to_buffer())

sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=512, num_phi=1024, h_ref=0.0)
sim.upload(synthetic_field(512, 1024, dtype=np.float32))
sim.step(1e-5, 50)
S.export_frame(os.path.join(out, "frame_drainage_field_512x1024.png"), 256, 256, S.render_drainage_frame(sim, 256, 256))
view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
sim.step_fused(1e-5, view, 1)
S.export_frame(os.path.join(out, "equirect_rgba8_512x1024.png"), 1024, 512, sim.download_rgb()[0])
print("wrote", sorted(f for f in os.listdir(out) if f.endswith(".png")))
