import time, numpy as np, os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import soap_bubble_simulation_b200 as S
def mk(prec, nt=128, nphi=256, **kw):
    return S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, num_theta=nt, num_phi=nphi, **kw)
def t(sim, n, f):
    f(sim, 10); sim.sync()
    t0 = time.perf_counter(); f(sim, n); t1 = time.perf_counter(); sim.sync(); t2 = time.perf_counter()
    return (t2 - t0) / n * 1e6, (t1 - t0) / n * 1e6
step = lambda s, n: s.step(1e-3, n)
for tfe in (True, False):
    for nt, nphi in ((128, 256), (64, 128), (32, 64), (16, 32)):
        a = mk(S.PREC_F32, nt, nphi); a.set_marginal_regeneration(tfe)
        tot, host = t(a, 4000, step)
        print(f"f32 {nt}x{nphi} tfe={tfe}: total {tot:.3f} us/step, host-side call {host:.3f} us/step", flush=True)
