"""Small workload touching every kernel family once; run under compute-sanitizer (scripts/sanitize.sh)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import soap_bubble_simulation_b200 as S  # noqa: E402
from soap_bubble_simulation_b200.synth import dt_max, synthetic_field  # noqa: E402

view = S.View()
for nt, nphi in ((130, 1028), (64, 2048), (37, 50)):
    f = synthetic_field(nt, nphi, dtype=np.float32)
    dt = min(1e-3, 0.25 * dt_max(nt, nphi))
    for prec in (S.PREC_F32, S.PREC_F64, S.PREC_F64_STRICT):
        sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, num_theta=nt, num_phi=nphi, h_ref=0.0)
        sim.upload(f)
        sim.step(dt, 2)
        sim.step_fused(dt, view, 2)
        sim.step_fused(dt, S.View(mode=S.RGB_EXACT3, format=S.OUT_RGB_F32), 1)
        sim.shade(view)
        sim.minmax()
        sim.get_thickness(1.0, 2.0)
        sim.download()
        sim.download_rgb()
        sim.close()
# TFE fronts active (t > 0.5 s) on the TMA path, offset storage, batch
sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=64, num_phi=256, batch=3)
sim.step_fused(5e-3, view, 130)
assert sim.active_tfe_count() >= 1
sim.download()
sim.close()
# shared-memory-resident multi-step path: strict / fast, fronts active, batch (cluster per bubble), then consumers
for prec in (S.PREC_F32, S.PREC_F64_STRICT):
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, num_theta=64, num_phi=256, batch=3)
    sim.step(5e-3, 160)
    assert sim.active_tfe_count() >= 1
    sim.caustic_map()
    sim.sample_uv(np.random.default_rng(0).uniform(0, 1, (64, 2)))
    sim.sample_normal(np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0]]))
    sim.download()
    sim.close()
sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=37, num_phi=50)
sim.step(1e-3, 40)  # ragged: scalar st.async pushes
sim.close()
g = S.GPUDrainageSimulator(500e-9, 128, 64)
g.dispatch(50)      # WGSL resident
g.download_concentration()
g.close()
g = S.GPUDrainageSimulator(500e-9, 256, 64)
g.set_marangoni_enabled(True)
g.dispatch(3, view)
g.download_concentration()
g.close()
out = np.empty((1, 1300, 256), np.float32)
px = np.empty((1, 1300, 256, 4), np.uint8)
sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=1300, num_phi=256, h_ref=0.0)
sim.step_host(1e-6, 1, view, synthetic_field(1300, 256, dtype=np.float32)[None].copy(), out, px)
sim.close()
# min/max left behind by the step kernel (forced on), batch with per-bubble offsets, fronts alive on the streaming kernel
sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=48, num_phi=1024, batch=3,
                          batch_h0_m=np.array([300e-9, 500e-9, 700e-9]))
sim.track_minmax(1)
sim.step(5e-3, 120)
for _ in range(3):
    sim.step_fused(5e-3, view, 1)
    sim.minmax()
    sim.has_critical_region()
assert sim.active_tfe_count() >= 1
# rows f3/f4: the fragment-shader frame (procedural film and drainage-field film)
S.HeadlessRenderPipeline.new(96, 64).render_to_buffer()
S.render_drainage_frame(sim, 64, 64, bubble=2)
sim.close()
# round 2: resident colour epilogue (CPU + WGSL semantics, fronts alive), direct spectral pass, foam host-buffer step
# (offset storage: in-place conversions), checkpoint / resume
sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=64, num_phi=256)
sim.step_fused(5e-3, view, 120)
sim.step_fused(5e-3, S.View(mode=S.RGB_EXACT3), 10)
st = sim.get_state()
sim.set_state(*st)
sim.shade_spectral(view, 380.0 + 400.0 * np.arange(9) / 8.0)
sim.step_fused(5e-3, S.View(mode=S.RGB_SPECTRAL), 2)
sim.download_rgb()
sim.close()
g = S.GPUDrainageSimulator(500e-9, 128, 64)
g.dispatch(10, view)
g.download_rgb()
g.close()
os.environ["SBX_HOST_CHUNK_BUBBLES"] = "2"
B = 5
sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=48, num_phi=256, batch=B, h_ref=1.0,
                          batch_h0_m=np.linspace(300e-9, 700e-9, B))
hin = np.full((B, 48, 256), 4e-7, np.float32)
hout, hpx = np.empty_like(hin), np.empty((B, 48, 256, 4), np.uint8)
sim.step_host(1e-3, 1, view, hin, hout, hpx)
sim.close()
S.generate_interference_lut(1.33, 1.0)
S.InterferenceCalculator(1.33).batch(np.linspace(0, 2000, 100), np.linspace(0, 1, 100))
print("SANITIZE_TARGET_DONE")
