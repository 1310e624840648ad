#!/usr/bin/env python
"""Where a frame of the reference's steps_per_frame loop (10 steps + colour frame, one resident launch) spends its time:
host enqueue cost per call (no synchronisation inside the loop), device-bound time per call (synchronised at the end of a
long loop), for the CPU-semantics default grid and the reference's GPU grid."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soap_bubble_simulation_b200 as S  # noqa: E402

view = S.View()


def probe(name, call, sync, n=300):
    for _ in range(20):
        call()
    sync()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    t_enq = (time.perf_counter() - t0) / n
    sync()
    t_all = (time.perf_counter() - t0) / n
    print(f"{name:70s} host enqueue {t_enq * 1e6:7.2f} us/call   wall incl. device {t_all * 1e6:7.2f} us/call", flush=True)


sim = S.DrainageSimulator(S.SimulationConfig.default())
probe("C1 128x256 step_fused(dt, view, 10): 10 steps + frame, one launch", lambda: sim.step_fused(1e-3, view, 10), sim.sync)
probe("C1 128x256 step(dt, 10): 10 steps, no frame", lambda: sim.step(1e-3, 10), sim.sync)
probe("C1 128x256 step(dt, 100)", lambda: sim.step(1e-3, 100), sim.sync, n=100)
probe("C1 128x256 step_fused(dt, view, 1): streaming fused step", lambda: sim.step_fused(1e-3, view, 1), sim.sync)
sim.set_marginal_regeneration(False)
probe("C1 128x256 step_fused(dt, view, 10), marginal regeneration off", lambda: sim.step_fused(1e-3, view, 10), sim.sync)
g = S.GPUDrainageSimulator(500e-9, 128, 64)
probe("GPU grid 128x64 WGSL dispatch(10, view): steps_per_frame + frame", lambda: g.dispatch(10, view), g.sync)
probe("GPU grid 128x64 WGSL dispatch(10)", lambda: g.dispatch(10), g.sync)
torch.cuda.synchronize()
