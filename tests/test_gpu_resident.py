"""GPU parity tests of the shared-memory-resident multi-step path (csrc/sbx_resident.cu).

`step(dt, n)` with n >= 2 on a small whole-sphere grid takes all n steps inside one cluster launch; the same calls
with SBX_NO_RESIDENT=1 (read when the handle is created) take the streaming kernels one launch per step.  Both are
checked against the CPU oracle, and against each other: strict f64 bit-identical, fast modes within the per-step
tolerance of BASELINE.json (1e-5 relative in fp32).
"""
import numpy as np
import pytest

import oracle_lib as O
from synth import dt_max, synthetic_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import soap_bubble_simulation_b200 as S

    return S


def make_sim(S, nt, np_, precision, h_ref=None, **kw):
    cfg = S.SimulationConfig.default()
    return S.DrainageSimulator(cfg, precision=precision, num_theta=nt, num_phi=np_, h_ref=h_ref, **kw)


def launches(sim):
    return sim.launch_count()


def resident_launches(nsteps):
    """A multi-step call is cut into launches of at most 1024 steps (the host computes the next chunk's front tables
    while the current one runs)."""
    return (nsteps + 1023) // 1024


@pytest.mark.parametrize("nt,nphi,steps", [(128, 256, 1000), (33, 50, 300), (16, 20, 2500), (4, 2, 50), (7, 5, 64)])
def test_strict_resident_is_bit_identical_and_single_launch(S, nt, nphi, steps):
    """C1 default config incl. TFE fronts: resident strict steps are bit-identical to the oracle, one launch per call."""
    sim = make_sim(S, nt, nphi, S.PREC_F64_STRICT)
    ora = O.OracleDrainage(nt, nphi)
    l0 = launches(sim)
    sim.step(1e-3, steps)
    assert launches(sim) - l0 == resident_launches(steps), "the multi-step call did not take the resident path"
    ora.step(1e-3, steps)
    assert np.array_equal(sim.download()[0], ora.data)
    assert sim.fronts() == ora.fronts()
    assert sim.current_time() == ora.current_time()
    # and mixing: single steps (streaming) after resident ones continue the same trajectory
    sim.step(1e-3, 1)
    sim.step(1e-3, 3)
    ora.step(1e-3, 4)
    assert np.array_equal(sim.download()[0], ora.data)


def test_strict_resident_equals_streaming(S, monkeypatch):
    nt, nphi = 64, 128
    f = synthetic_field(nt, nphi)
    dt = 0.25 * dt_max(nt, nphi)
    a = make_sim(S, nt, nphi, S.PREC_F64_STRICT)
    monkeypatch.setenv("SBX_NO_RESIDENT", "1")
    b = make_sim(S, nt, nphi, S.PREC_F64_STRICT)
    monkeypatch.delenv("SBX_NO_RESIDENT")
    for s in (a, b):
        s.upload(f)
    la, lb = launches(a), launches(b)
    a.step(dt, 150)
    b.step(dt, 150)
    assert launches(a) - la == 1 and launches(b) - lb >= 150
    assert np.array_equal(a.download()[0], b.download()[0])


@pytest.mark.parametrize("prec_name,tol", [("PREC_F32", 1e-5), ("PREC_F64", 1e-12)])
@pytest.mark.parametrize("nt,nphi", [(128, 256), (100, 200), (37, 50), (5, 7), (256, 512), (17, 1000)])
def test_fast_resident_matches_oracle(S, prec_name, tol, nt, nphi):
    """Non-uniform fields, ragged shapes, cluster widths 1..8: per-call error against the f64 oracle and the
    increment check (a frozen field cannot pass)."""
    prec = getattr(S, prec_name)
    f = synthetic_field(nt, nphi)
    if prec == S.PREC_F32:
        f = f.astype(np.float32).astype(np.float64)  # start from values the storage type holds exactly
    dt = 0.25 * dt_max(nt, nphi)
    sim = make_sim(S, nt, nphi, prec, h_ref=0.0 if prec == S.PREC_F32 else None)
    ora = O.OracleDrainage(nt, nphi)
    ora.set_marginal_regeneration(False)
    sim.set_marginal_regeneration(False)
    sim.upload(f)
    ora.data[:] = f
    n = 40
    sim.step(dt, n)
    ora.step(dt, n)
    got = sim.download()[0].astype(np.float64)
    assert np.abs(got / ora.data - 1).max() <= tol * (n if prec == S.PREC_F32 else 1)
    inc_ref, inc_got = ora.data - f, got - f
    assert np.abs(inc_ref).max() > 0
    assert np.abs(inc_got - inc_ref).max() <= (5e-2 if prec == S.PREC_F32 else 1e-8) * np.abs(inc_ref).max()
    assert np.all(got[0] == got[0, 0]) and np.all(got[-1] == got[-1, 0])


def test_f32_resident_default_config_with_tfe_long_run(S):
    """C1 in f32 offset storage with TFE fronts: 1000 steps in one launch, drift as in the streaming path."""
    nt, nphi = 128, 256
    sim = make_sim(S, nt, nphi, S.PREC_F32)
    ora = O.OracleDrainage(nt, nphi)
    l0 = launches(sim)
    sim.step(1e-3, 1000)
    assert launches(sim) - l0 == 1
    ora.step(1e-3, 1000)
    got = sim.download()[0]
    assert np.abs(got / ora.data - 1).max() <= 1e-6
    assert sim.fronts() == ora.fronts()
    assert sim.active_tfe_count() == ora.active_tfe_count()


def test_f32_resident_equals_streaming_within_tolerance(S, monkeypatch):
    nt, nphi = 128, 256
    f = synthetic_field(nt, nphi).astype(np.float32)
    dt = 0.25 * dt_max(nt, nphi)
    a = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    monkeypatch.setenv("SBX_NO_RESIDENT", "1")
    b = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    monkeypatch.delenv("SBX_NO_RESIDENT")
    for s in (a, b):
        s.upload(f)
        s.step(dt, 100)
    ga, gb = a.download()[0], b.download()[0]
    # same per-cell arithmetic; only the summation order of the pole means differs
    assert np.abs(ga / gb - 1).max() <= 5e-6
    assert np.array_equal(ga[2:-2], gb[2:-2]) or np.abs(ga[2:-2] / gb[2:-2] - 1).max() <= 5e-6


def test_resident_batch_of_independent_bubbles(S):
    """One cluster per bubble; per-bubble radius and thickness as in the foam path."""
    nt, nphi, nb = 32, 64, 12
    radii = np.array([0.02 + 0.002 * i for i in range(nb)])
    h0 = np.array([400e-9 + 20e-9 * i for i in range(nb)])
    sim = make_sim(S, nt, nphi, S.PREC_F64, batch=nb, batch_h0_m=h0, batch_radius_m=radii)
    fields = np.stack([synthetic_field(nt, nphi, seed=i, h0_nm=h0[i] * 1e9) for i in range(nb)])
    sim.upload(fields)
    dt = 0.1 * dt_max(nt, nphi)
    l0 = launches(sim)
    sim.step(dt, 30)
    assert launches(sim) - l0 == 1
    got = sim.download()
    for i in range(nb):
        ora = O.OracleDrainage(nt, nphi, h0_m=h0[i], radius_m=radii[i])
        ora.data[:] = fields[i]
        ora.step(dt, 30)
        assert np.abs(got[i] / ora.data - 1).max() < 1e-12, i


@pytest.mark.parametrize("prec_name", ["PREC_F32", "PREC_F64"])
def test_resident_result_does_not_depend_on_cluster_width(S, prec_name):
    """The batch size picks the cluster width (16, 8, 4, 2, 1 CTAs per bubble).  Cell arithmetic and the order of the
    pole sums do not depend on it, so every bubble of every batch equals the lone bubble bit for bit."""
    nt, nphi = 48, 64
    prec = getattr(S, prec_name)
    f = synthetic_field(nt, nphi).astype(np.float32).astype(np.float64)
    dt = 0.25 * dt_max(nt, nphi)
    ref = None
    for nb in (1, 12, 20, 40, 100):
        sim = make_sim(S, nt, nphi, prec, h_ref=0.0 if prec == S.PREC_F32 else None, batch=nb)
        sim.upload(np.broadcast_to(f, (nb, nt, nphi)).copy())
        l0 = launches(sim)
        sim.step(dt, 50)
        assert launches(sim) - l0 == 1
        got = sim.download()
        if ref is None:
            ref = got[0].copy()
        for b in range(nb):
            assert np.array_equal(got[b], ref), (nb, b)


@pytest.mark.parametrize("marangoni", [False, True])
@pytest.mark.parametrize("w,h", [(128, 64), (50, 37), (256, 128)])
def test_wgsl_resident_matches_f32_oracle(S, w, h, marangoni, monkeypatch):
    """The reference's GPU drainage (128 x 64, steps_per_frame dispatches per frame) in one launch."""
    sim = S.GPUDrainageSimulator(500e-9, w, h)
    monkeypatch.setenv("SBX_NO_RESIDENT", "1")
    ref = S.GPUDrainageSimulator(500e-9, w, h)
    monkeypatch.delenv("SBX_NO_RESIDENT")
    rng = np.random.default_rng(3)
    f = synthetic_field(h, w, h0_nm=5000.0, noise_nm=20.0).astype(np.float32)
    c = np.clip(0.5 + 0.2 * rng.standard_normal((h, w)), 0, 1).astype(np.float32)
    for s in (sim, ref):
        s.set_marangoni_enabled(marangoni)
        s.upload(f)
        s.upload_concentration(c)
    n = 24
    l0 = launches(sim)
    sim.dispatch(n)
    assert launches(sim) - l0 == 1
    ref.dispatch(n)
    p = O.wgsl_params(grid_width=w, grid_height=h, marangoni_enabled=int(marangoni), dt=0.001)
    oh, oc = f, c
    for _ in range(n):
        oh, oc = O.wgsl_step(p, oh, oc)
    gh, gc = sim.download(np.float32)[0], sim.download_concentration()[0]
    rh, rc = ref.download(np.float32)[0], ref.download_concentration()[0]
    assert np.abs(gh.astype(np.float64) / oh - 1).max() <= 1e-5 and np.abs(gc.astype(np.float64) - oc).max() <= 1e-5
    assert np.abs(gh.astype(np.float64) / rh - 1).max() <= 1e-5 and np.abs(gc - rc).max() <= 1e-5
    assert np.array_equal(gh[0], f[0]) and np.array_equal(gh[-1], f[-1])  # frozen pole rows


def test_fused_multi_step_call_uses_resident_steps_then_one_fused(S):
    nt, nphi = 128, 256
    f = synthetic_field(nt, nphi).astype(np.float32)
    dt = 0.25 * dt_max(nt, nphi)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    a = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    a.upload(f)
    l0 = launches(a)
    a.step_fused(dt, view, 25)
    assert launches(a) - l0 <= 4  # 1 resident + the fused step (+ LUT build)
    rgb_a = a.download_rgb()
    # the same final field shaded on its own gives the same frame
    b = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    b.upload(a.download()[0])
    b.shade(view)
    rgb_b = b.download_rgb()
    assert np.abs(rgb_a.astype(int) - rgb_b.astype(int)).max() <= 1


def test_strict_resident_equals_streaming_on_random_shapes(S, monkeypatch):
    """Twenty random (rows, columns, batch, steps, fronts on/off) combinations, odd and prime sizes included: the
    resident and the streaming path give the same bits in strict mode, and the same front state."""
    rng = np.random.default_rng(20261017)
    for case in range(20):
        nt = int(rng.integers(4, 140))
        nphi = int(rng.integers(2, 300))
        nb = int(rng.choice([1, 1, 2, 5, 17]))
        n = int(rng.integers(2, 60))
        tfe = bool(rng.integers(0, 2))
        dt = 0.2 * dt_max(nt, nphi) if not tfe else min(0.05, 0.2 * dt_max(nt, nphi))
        f = np.stack([synthetic_field(nt, nphi, seed=100 + case + b) for b in range(nb)])
        sims = []
        for no_res in ("0", "1"):
            monkeypatch.setenv("SBX_NO_RESIDENT", no_res)
            sim = make_sim(S, nt, nphi, S.PREC_F64_STRICT, batch=nb)
            sim.set_marginal_regeneration(tfe)
            sim.upload(f)
            sims.append(sim)
        monkeypatch.delenv("SBX_NO_RESIDENT")
        if tfe:  # bring fronts to life first (both handles through the same calls)
            for sim in sims:
                sim.step(0.3, 2)
        l0 = launches(sims[0])
        for sim in sims:
            sim.step(dt, n)
        assert launches(sims[0]) - l0 == 1, (case, nt, nphi, nb, n)
        a, b = sims[0].download(), sims[1].download()
        assert np.array_equal(a, b, equal_nan=True), (case, nt, nphi, nb, n, tfe)
        assert sims[0].fronts() == sims[1].fronts() and sims[0].current_time() == sims[1].current_time()
