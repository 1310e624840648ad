"""The CUDA path against known answers that come from no restatement (tests/closed_form.py): one step of each drainage
semantics on a field a + eps * theta * cos(m phi), whose finite differences have closed forms.  The oracle passes the same
checks on CPU (tests/test_oracle_reference_tests.py::test_kat_*_discrete_mode_closed_form); here the kernels behind the C
ABI do -- the strict f64 mode at the rounding of the stored field, the streaming f64 and f32 kernels (k_march_tma) and
the WGSL-semantics kernel (k_wgsl_tma) within a few ulps of their storage type."""
import numpy as np
import pytest

import oracle_lib as O
from closed_form import cpu_mode, wgsl_mode

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import soap_bubble_simulation_b200 as S

    return S


@pytest.mark.parametrize("prec_name,tol,tol_pole", [("PREC_F64_STRICT", 2e-22, 1e-21), ("PREC_F64", 1e-21, 2e-21),
                                                    ("PREC_F32", 1.2e-13, 1.2e-13)])
def test_cpu_semantics_step_matches_the_closed_form(S, prec_name, tol, tol_pole):
    nt, nphi, m, dt = 33, 128, 3, 0.05
    d = O.DEFAULTS  # config/default.json, what SimulationConfig.default() holds
    f32 = prec_name == "PREC_F32"
    h, inc, top, bot = cpu_mode(nt, nphi, m, 500e-9, 40e-9, dt, d["density"], d["gravity"], d["viscosity"], d["diffusion"],
                                d["radius_m"], storage=np.float32 if f32 else np.float64)
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=getattr(S, prec_name), num_theta=nt, num_phi=nphi,
                              h_ref=0.0 if f32 else None)
    sim.upload(h)
    sim.step(dt, 1)
    got = sim.download()[0].astype(np.float64)
    sim.close()
    assert np.abs(inc).max() > 5e-12  # 160 ulps of the f32 field, 8e10 of the f64 one
    assert np.abs(got[1:-1] - h[1:-1].astype(np.float64) - inc).max() <= tol
    assert abs(got[0, 0] - top) <= tol_pole and np.all(got[0] == got[0, 0])
    assert abs(got[-1, 7] - bot) <= tol_pole and np.all(got[-1] == got[-1, 0])


def test_wgsl_semantics_step_matches_the_closed_form(S):
    H, W = 33, 128
    sim = S.GPUDrainageSimulator(500e-9, W, H)
    q = sim.params()
    sim.set_dt(1.0)
    sim.set_physics(q.viscosity, q.density, q.gravity, 1e-7)
    sim.set_marangoni_params(q.gamma_air, q.gamma_reduction, 1e-7, 2e3)
    sim.set_marangoni_enabled(True)
    p = sim.params()
    assert p.grid_width == W and p.grid_height == H and p.marangoni_enabled == 1
    h, c, inc_h, inc_c, grav, mar = wgsl_mode(H, W, 3, 500e-9, 40e-9, 0.4, 0.05, p)
    sim.upload(h)
    sim.upload_concentration(c)
    sim.dispatch(1)
    gh, gc = sim.download(np.float32)[0].astype(np.float64), sim.download_concentration()[0].astype(np.float64)
    assert np.array_equal(gh[[0, -1]], h[[0, -1]]) and np.array_equal(gc[[0, -1]], c[[0, -1]])  # poles: copied
    # the f32 differences of neighbouring values carry ~ulp / dphi^2 of noise: the oracle meets the closed form to
    # 7e-13 m and 4.5e-7 at this grid; increments are up to 1.8e-8 m and 0.023
    tol_h, tol_c = 2.5e-12, 1.5e-6
    assert np.abs(mar).max() > 1000 * tol_h and np.abs(inc_h - mar).max() > 1000 * tol_h
    assert np.abs(gh[1:-1] - h[1:-1].astype(np.float64) - inc_h).max() <= tol_h
    assert np.abs(gc[1:-1] - c[1:-1].astype(np.float64) - inc_c).max() <= tol_c
