#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json config and every precision/colour mode (one GPU).

Prints one JSON object per line and a summary table; results are committed under profiles/.  Timing: CUDA
events on the launching stream, warm-up, inputs larger than L2 for every config except C1.
"""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import soap_bubble_simulation_b200 as S  # noqa: E402
from soap_bubble_simulation_b200.synth import dt_max, synthetic_field  # noqa: E402

PEAK = 6532.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rows = []


def timed(fn, steps, warm=5):
    fn(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, cells, bytes_per_cell, ms, note=""):
    gcell = cells / (ms * 1e-3) / 1e9
    gbs = gcell * bytes_per_cell
    r = {"config": name, "cells": cells, "ms_per_step": round(ms, 5), "gcell_per_s": round(gcell, 2),
         "bytes_per_cell": bytes_per_cell, "achieved_gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 4),
         "note": note}
    rows.append(r)
    print(json.dumps(r), flush=True)


def single(name, nt, nphi, prec, mode, steps, bpc, h_ref=0.0, note="", semantics="cpu", marangoni=False):
    cfg = S.SimulationConfig.default()
    dt = min(1e-3, 0.25 * dt_max(nt, nphi))
    view = S.View(mode=mode, format=S.OUT_RGBA8) if mode != S.RGB_NONE else None
    if semantics == "wgsl":
        sim = S.GPUDrainageSimulator(500e-9, nphi, nt)
        sim.set_marangoni_enabled(marangoni)
        sim.set_dt(dt)
        sim.upload(synthetic_field(nt, nphi, dtype=np.float32))
        sim.set_stream(stream.cuda_stream)
        fn = lambda n: sim.dispatch(n, view)  # noqa: E731
    else:
        sim = S.DrainageSimulator(cfg, precision=prec, num_theta=nt, num_phi=nphi, h_ref=h_ref)
        f = synthetic_field(nt, nphi, dtype=np.float32)
        sim.upload(f)
        del f
        sim.set_stream(stream.cuda_stream)
        fn = (lambda n: sim.step(dt, n)) if view is None else (lambda n: sim.step_fused(dt, view, n))
    ms = timed(fn, steps)
    report(name, (nt - 2) * nphi if semantics == "cpu" else (nt - 2) * nphi, bpc, ms, note)
    sim.close()


which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5", "modes"]

if "c1" in which:
    # config 1: config/default.json, 128x256, 1000 steps.  A multi-step call keeps the field resident in shared memory
    # (sbx_resident.cu); one step per call streams through HBM and is launch-bound.
    def loop(f):
        def run(n):
            for _ in range(n):
                f()
        return run

    sim = S.DrainageSimulator(S.SimulationConfig.default())
    sim.set_stream(stream.cuda_stream)
    ms = timed(lambda n: sim.step(1e-3, n), 1000, warm=50)
    report("C1 128x256 default.json drainage (f32 offset storage, TFE on), 1000 steps per call", 126 * 256, 8, ms,
           "shared-memory resident: no HBM traffic inside the call; exchange-latency-bound")
    ms = timed(loop(lambda: sim.step(1e-3, 1)), 1000, warm=50)
    report("C1 128x256 drainage, one step per call", 126 * 256, 8, ms, "launch-bound")
    view = S.View()
    ms = timed(loop(lambda: sim.step_fused(1e-3, view, 1)), 1000, warm=50)
    report("C1 128x256 fused drainage + LUT interference, one frame per step", 126 * 256, 12, ms, "launch-bound")
    ms = timed(loop(lambda: sim.step_fused(1e-3, view, 10)), 100, warm=10) / 10.0
    report("C1 128x256 fused, 10 steps per frame (9 resident + 1 fused), per step", 126 * 256, 12, ms, "")
    sim.close()
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F64_STRICT)
    sim.set_stream(stream.cuda_stream)
    ms = timed(lambda n: sim.step(1e-3, n), 1000, warm=50)
    report("C1 128x256 f64 STRICT (bit-identical to the reference), 1000 steps per call", 126 * 256, 16, ms, "resident")
    sim.close()
    g = S.GPUDrainageSimulator(500e-9, 128, 64)
    g.set_stream(stream.cuda_stream)
    ms = timed(lambda n: g.dispatch(n), 1000, warm=50)
    report("reference GPU grid 128x64 WGSL semantics, 1000 dispatches per call", 62 * 128, 16, ms, "resident")
    g.close()

if "c2" in which:
    single("C2 2048x4096 f32 fused LUT->RGBA8", 2048, 4096, S.PREC_F32, S.RGB_LUT, 400, 12, note="64 MiB working set: L2-resident")
    single("C2 2048x4096 f32 drainage only", 2048, 4096, S.PREC_F32, S.RGB_NONE, 400, 8, note="L2-resident")

if "c3" in which:
    single("C3 8192x16384 f32 fused LUT->RGBA8 (headline)", 8192, 16384, S.PREC_F32, S.RGB_LUT, 100, 12)
    single("C3 8192x16384 f32 drainage only", 8192, 16384, S.PREC_F32, S.RGB_NONE, 100, 8)

if "hl" in which:  # headline only (A/B runs)
    single("C3 8192x16384 f32 fused LUT->RGBA8 (headline)", 8192, 16384, S.PREC_F32, S.RGB_LUT, 200, 12)
    single("C2 2048x4096 f32 fused LUT->RGBA8", 2048, 4096, S.PREC_F32, S.RGB_LUT, 400, 12)

if "exact_hl" in which:  # the headline alone (A/B of run-time knobs)
    single("C3 8192x16384 f32 fused LUT->RGBA8 (headline)", 8192, 16384, S.PREC_F32, S.RGB_LUT, 200, 12)
    single("C3 8192x16384 f32 drainage only", 8192, 16384, S.PREC_F32, S.RGB_NONE, 100, 8)

for w in which:  # "grid:<rows>:<cols>": the fused f32 step on any grid (band-height / launch-shape studies)
    if w.startswith("grid:"):
        _, gnt, gnphi = w.split(":")
        single(f"grid {gnt}x{gnphi} f32 fused LUT->RGBA8", int(gnt), int(gnphi), S.PREC_F32, S.RGB_LUT, 300, 12)

if "tfe" in which:
    # The headline run starts at t = 0, where no marginal-regeneration front exists yet (first spawn at t = 0.5 s,
    # drainage.rs:397-403), so it runs the kernel's no-front instantiation.  This line times the TFE = true
    # instantiation with fronts alive: the simulator is advanced to t = 15 s with a dt far above the stability bound
    # of this grid (the field values are meaningless by then -- timing only, same memory traffic and instruction path).
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=8192, num_phi=16384, h_ref=0.0)
    sim.upload(synthetic_field(8192, 16384, dtype=np.float32))
    sim.set_stream(stream.cuda_stream)
    sim.step(2.5e-3, 6000)
    view = S.View()
    ms = timed(lambda n: sim.step_fused(2.5e-3, view, n), 100)
    report("C3 f32 fused LUT->RGBA8 with marginal-regeneration fronts alive (timing only)", 8190 * 16384, 12, ms,
           f"{sim.active_tfe_count()} fronts alive at t = {sim.current_time():.2f} s")
    sim.close()

if "frame" in which:
    # The reference's per-frame call sequence around the path (RenderPipeline::update, pipeline.rs:1289-1322):
    # step, thickness_field().min_thickness(), .max_thickness(), get_thickness(pi/2, 0), has_critical_region().
    import time
    for nt, nphi, B, tag in ((8192, 16384, 1, "C3 8192x16384"), (256, 512, 512, "C4 foam 512 x 256x512")):
        kw = dict(batch=B, batch_h0_m=np.linspace(300e-9, 700e-9, B), h_ref=1.0) if B > 1 else dict(h_ref=0.0)
        sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=nt, num_phi=nphi, **kw)
        if B == 1:
            sim.upload(synthetic_field(nt, nphi, dtype=np.float32))
        sim.set_stream(stream.cuda_stream)
        view = S.View()
        dt = min(1e-3, 0.25 * dt_max(nt, nphi))
        fld = sim.thickness_field()

        def frame():
            sim.step_fused(dt, view, 1)
            lo, hi = fld.min_thickness(), fld.max_thickness()
            mid = sim.get_thickness(math.pi / 2, 0.0)
            return lo, hi, mid, sim.has_critical_region()

        for _ in range(5):
            frame()
        t0 = time.perf_counter()
        for _ in range(50):
            frame()
        ms = (time.perf_counter() - t0) / 50 * 1e3
        ms_mm = timed(lambda n: [sim.step_fused(dt, view, 1) or sim.minmax() for _ in range(n)], 20) - timed(lambda n: sim.step_fused(dt, view, n), 20)
        report(f"{tag}: one frame of the reference's update loop (fused step + min + max + get_thickness + has_critical), host wall clock",
               B * (nt - 2) * nphi, 12, ms, f"min/max query after a step: {ms_mm:.4f} ms (a pass over the field would be {B * nt * nphi * 4 / 6.3e9:.4f} ms at HBM speed; "
               "by default the step kernel leaves min/max behind and the three reductions share them)")
        sim.close()

if "l2ab" in which:  # the other fields-larger-than-L2 cases of the march kernel (A/B of the L2 policy)
    single("C3 f64 fused LUT->RGBA8", 8192, 16384, S.PREC_F64, S.RGB_LUT, 50, 20)
    single("C3 f32 drainage only", 8192, 16384, S.PREC_F32, S.RGB_NONE, 100, 8)
    single("C3 f32 fused two-beam Schlick->RGBA8", 8192, 16384, S.PREC_F32, S.RGB_SCHLICK_TWO_BEAM, 50, 12)

if "modes" in which:
    single("C3 f32 fused EXACT3 (3-lambda exact Fresnel + gamma)->RGBA8", 8192, 16384, S.PREC_F32, S.RGB_EXACT3, 30, 12,
           note="SFU/issue-bound: 3 sinpi + 3 pow per cell, f64 phase")
    single("C3 f32 fused two-beam Schlick (the documented original model)->RGBA8", 8192, 16384, S.PREC_F32,
           S.RGB_SCHLICK_TWO_BEAM, 50, 12, note="analytic 3-lambda colour, no LUT")
    single("C3 f64 fused LUT->RGBA8", 8192, 16384, S.PREC_F64, S.RGB_LUT, 50, 20)
    single("C3 f64 drainage only", 8192, 16384, S.PREC_F64, S.RGB_NONE, 50, 16)
    single("C3 f64 STRICT drainage only (bit-identical mode)", 8192, 16384, S.PREC_F64_STRICT, S.RGB_NONE, 20, 16,
           note="5 IEEE f64 divisions per cell: FP64-bound")
    single("C3 WGSL semantics (h + concentration), no Marangoni", 8192, 16384, S.PREC_F32, S.RGB_NONE, 50, 16, semantics="wgsl")
    single("C3 WGSL semantics + Marangoni + LUT->RGBA8", 8192, 16384, S.PREC_F32, S.RGB_LUT, 50, 20, semantics="wgsl", marangoni=True)

if "exact3" in which:
    single("C3 f32 fused EXACT3 (3-lambda exact Fresnel + gamma)->RGBA8", 8192, 16384, S.PREC_F32, S.RGB_EXACT3, 30, 12)

if "wgsl" in which:
    single("C3 WGSL semantics (h + concentration), no Marangoni", 8192, 16384, S.PREC_F32, S.RGB_NONE, 50, 16, semantics="wgsl")
    single("C3 WGSL semantics + Marangoni + LUT->RGBA8", 8192, 16384, S.PREC_F32, S.RGB_LUT, 50, 20, semantics="wgsl", marangoni=True)

if "c4" in which:
    # config 4: foam of independent bubbles, 256x512 each; per-bubble h0 and radius, offset storage
    for B in (512, 4096):
        idx = np.arange(B) * (4095.0 / max(B - 1, 1))
        h0 = (300.0 + 400.0 * idx / 4095.0) * 1e-9
        rad = 0.0125 + 0.0875 * idx / 4095.0
        sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=256, num_phi=512,
                                  h_ref=1.0, batch=B, batch_h0_m=h0, batch_radius_m=rad)
        sim.set_stream(stream.cuda_stream)
        view = S.View()
        ms = timed(lambda n: sim.step_fused(1e-3, view, n), 50)
        report(f"C4 foam {B} bubbles x 256x512 f32 fused LUT->RGBA8 (one GPU)", B * 254 * 512, 12, ms,
               "independent bubbles, no exchange" + ("; 8-GPU share of the 4096-bubble cluster" if B == 512 else ""))
        sim.close()

if "c5" in which:
    # config 5: 64-wavelength LUT build + shading of an 8192x16384 field for 8 yaw angles
    wl = 380.0 + 400.0 * np.arange(64) / 63.0
    import time
    S.generate_spectral_lut(1.33, 1.0, wl, 256, 64, 2000.0)  # warm-up: module load, first-launch set-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        lut = S.generate_spectral_lut(1.33, 1.0, wl, 256, 64, 2000.0)
    t_lut = (time.perf_counter() - t0) / 20
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=8192, num_phi=16384, h_ref=0.0)
    sim.upload(synthetic_field(8192, 16384, dtype=np.float32))
    sim.set_lut(lut)
    sim.set_stream(stream.cuda_stream)
    views = [S.View(cam_pos=S.camera_position(0.2, 2 * math.pi * k / 8, 0.3)) for k in range(8)]

    def sweep(n):
        for i in range(n):
            sim.shade(views[i % 8])

    ms = timed(sweep, 80, warm=8)
    report("C5 spectral sweep: LUT shading of 8192x16384 per view (64-lambda LUT)", 8192 * 16384, 8, ms,
           f"LUT build 256x64x64 lambda incl. D2H: {t_lut * 1e3:.3f} ms per build (host wall clock)")
    sim.close()

if "f2" in which or "c3" in which:
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=8192, num_phi=16384, h_ref=0.0)
    sim.upload(synthetic_field(8192, 16384, dtype=np.float32))
    sim.set_stream(stream.cuda_stream)
    import ctypes as C
    from soap_bubble_simulation_b200 import _lib as L
    cp = L.CausticParamsC()
    L.check(L.lib().sbx_caustic_params_default(C.byref(cp)))

    def caus(n):
        for _ in range(n):
            L.check(L.lib().sbx_drainage_caustics(sim._h, C.byref(cp), None))

    ms = timed(caus, 50)
    report("f2 caustic map (9-point stencil consumer) of 8192x16384", 8192 * 16384, 8, ms,
           "direct loads" if os.environ.get("SBX_NO_TMA") == "1" else "TMA ring")
    sim.close()

print("\n| config | ms/step | Gcell/s | B/cell | GB/s | % of measured HBM peak | note |")
print("|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['config']} | {r['ms_per_step']} | {r['gcell_per_s']} | {r['bytes_per_cell']} | {r['achieved_gbs']} | "
          f"{100 * r['frac_of_measured_hbm']:.1f} | {r['note']} |")
