import time, numpy as np, os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import soap_bubble_simulation_b200 as S
def mk(prec, nt=128, nphi=256, **kw):
    return S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, num_theta=nt, num_phi=nphi, **kw)
def t(sim, n, f):
    f(sim, 10 if n > 10 else n); sim.sync()
    t0 = time.perf_counter(); f(sim, n); sim.sync(); return (time.perf_counter() - t0) / n * 1e6
for name, prec in (("f32", S.PREC_F32), ("f64", S.PREC_F64), ("strict", S.PREC_F64_STRICT)):
    for nt, nphi in ((128, 256), (100, 200), (64, 128)):
        a = mk(prec, nt, nphi)
        os.environ["SBX_NO_RESIDENT"] = "1"; b = mk(prec, nt, nphi); del os.environ["SBX_NO_RESIDENT"]
        step = lambda s, n: s.step(1e-3, n)
        print(f"{name} {nt}x{nphi}: resident {t(a, 2000, step):.3f} us/step, streaming {t(b, 2000, step):.3f} us/step", flush=True)
for w, h in ((128, 64), (256, 128)):
    a = S.GPUDrainageSimulator(500e-9, w, h)
    os.environ["SBX_NO_RESIDENT"] = "1"; b = S.GPUDrainageSimulator(500e-9, w, h); del os.environ["SBX_NO_RESIDENT"]
    d = lambda s, n: s.dispatch(n)
    print(f"wgsl {w}x{h}: resident {t(a, 2000, d):.3f} us/step, streaming {t(b, 2000, d):.3f} us/step", flush=True)
# foam batch 16 x 128x256
a = mk(S.PREC_F32, 128, 256, batch=16)
os.environ["SBX_NO_RESIDENT"] = "1"; b = mk(S.PREC_F32, 128, 256, batch=16); del os.environ["SBX_NO_RESIDENT"]
print(f"f32 16 bubbles 128x256: resident {t(a, 2000, step):.3f} us/step, streaming {t(b, 2000, step):.3f}", flush=True)
