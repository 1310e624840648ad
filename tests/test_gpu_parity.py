"""GPU parity tests: the CUDA path behind the C ABI against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): thickness max relative error <= 1e-5 per step in fp32 (plus the
increment check of SURVEY.md 8d so that a frozen field cannot pass); interference RGB <= 1e-4 absolute,
<= 1 LSB in RGBA8; integer/byte work (LUT bytes) within 1 LSB of a truncating generator; the strict f64
mode is bit-identical.
"""
import math

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


# ------------------------------------------------------------------ strict f64: bit-identical
@pytest.mark.parametrize("nt,nphi,steps", [(128, 256, 1000), (33, 50, 300), (16, 20, 2500)])
def test_strict_f64_bit_identical_default_config(S, nt, nphi, steps):
    """C1 (config/default.json, uniform 500 nm, dt 1e-3, TFE on): every bit of the field after every
    checkpoint, TFE fronts included; (16,20) runs long enough for several fronts and front removal."""
    sim = make_sim(S, nt, nphi, S.PREC_F64_STRICT)
    ora = O.OracleDrainage(nt, nphi)
    done = 0
    for chunk in (1, 9, 90, steps - 100):
        sim.step(1e-3, chunk)
        ora.step(1e-3, chunk)
        done += chunk
        got = sim.download()[0]
        assert np.array_equal(got, ora.data), f"mismatch after {done} steps: max |d| = {np.abs(got - ora.data).max()}"
        assert sim.fronts() == ora.fronts()
    assert sim.current_time() == ora.current_time()
    assert sim.active_tfe_count() == ora.active_tfe_count()


def test_strict_f64_bit_identical_nonuniform(S):
    nt, nphi = 64, 128
    f = synthetic_field(nt, nphi)
    dt = 0.25 * dt_max(nt, nphi)
    sim = make_sim(S, nt, nphi, S.PREC_F64_STRICT)
    ora = O.OracleDrainage(nt, nphi)
    sim.upload(f)
    ora.data[:] = f
    sim.step(dt, 200)
    ora.step(dt, 200)
    assert np.array_equal(sim.download()[0], ora.data)


# ------------------------------------------------------------------ fast f64
def test_fast_f64_matches_oracle(S):
    nt, nphi = 128, 256
    f = synthetic_field(nt, nphi)
    dt = 0.25 * dt_max(nt, nphi)
    sim = make_sim(S, nt, nphi, S.PREC_F64)
    ora = O.OracleDrainage(nt, nphi)
    sim.upload(f)
    ora.data[:] = f
    prev = f.copy()
    for _ in range(20):
        sim.step(dt, 1)
        ora.step(dt, 1)
        got = sim.download()[0]
        inc_ref = ora.data - prev
        inc_got = got - prev
        assert np.abs(got / ora.data - 1).max() < 1e-13
        assert np.abs(inc_got - inc_ref).max() <= 1e-9 * np.abs(inc_ref).max()
        prev = ora.data.copy()
        sim.upload(prev)  # same input for the next single-step comparison


# ------------------------------------------------------------------ f32
def test_f32_per_step_error_and_increments_resolvable_regime(S):
    """fp32 storage, plain (h_ref = 0): per-step max relative error <= 1e-5 vs the f64 oracle from the same
    f32-representable input, and the increment check in a regime where increments are above the fp32 ulp
    (thick film, large dt)."""
    nt, nphi = 128, 256
    cfg = S.SimulationConfig.default()
    cfg.bubble.film_thickness_nm = 20000.0  # 20 um: drainage increment 2.6e-8 relative per step
    sim = S.DrainageSimulator(cfg, precision=S.PREC_F32, num_theta=nt, num_phi=nphi, h_ref=0.0)
    ora = O.OracleDrainage(nt, nphi, h0_m=cfg.film_thickness_meters())
    sim.set_marginal_regeneration(False)
    ora.set_marginal_regeneration(False)
    f = (synthetic_field(nt, nphi, h0_nm=20000.0, noise_nm=50.0)).astype(np.float32)
    dt = 2.0  # well inside the diffusion bound (0.115 s would be for D-dominated; here we test drainage)
    dt = min(dt, 0.25 * dt_max(nt, nphi))
    sim.upload(f)
    ora.data[:] = f.astype(np.float64)
    sim.step(dt, 1)
    ora.step(dt, 1)
    got = sim.download(np.float32)[0].astype(np.float64)
    rel = np.abs(got / ora.data - 1).max()
    assert rel <= 1e-5, rel
    inc_ref = ora.data - f
    inc_got = got - f
    # fp32 storage rounds the *result* to 6e-8 relative: compare increments with that floor
    floor = 6e-8 * np.abs(ora.data).max()
    assert np.abs(inc_got - inc_ref).max() <= 1e-3 * np.abs(inc_ref).max() + floor


def test_f32_delta_storage_keeps_sub_ulp_increments_c1(S):
    """C1: uniform 500 nm, dt 1e-3.  The per-step change (8e-10 relative) is far below the fp32 ulp, so plain
    f32 freezes; with h_ref = h0 the device stores h - h_ref and the increments survive: per-step increment
    error <= 1e-3 of the oracle's, 1000-step drift reported and bounded."""
    nt, nphi = 128, 256
    sim = make_sim(S, nt, nphi, S.PREC_F32)  # h_ref defaults to the initial thickness
    ora = O.OracleDrainage(nt, nphi)
    h0 = ora.data.copy()
    sim.step(1e-3, 1)
    ora.step(1e-3, 1)
    got = sim.download()[0]
    inc_ref, inc_got = ora.data - h0, got - h0
    assert np.abs(inc_ref).max() > 0
    assert np.abs(inc_got - inc_ref).max() <= 1e-3 * np.abs(inc_ref).max()
    assert np.abs(got / ora.data - 1).max() <= 1e-5
    sim.step(1e-3, 999)
    ora.step(1e-3, 999)
    got = sim.download()[0]
    drift = np.abs(got / ora.data - 1).max()
    print(f"\n[drift] C1 f32(delta) vs f64 oracle after 1000 steps: max rel {drift:.3e}")
    assert drift <= 1e-5
    assert sim.fronts() == ora.fronts()
    # and the plain-f32 twin really is frozen outside the TFE fronts (documented trap (i))
    plain = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    plain.set_marginal_regeneration(False)
    p0 = plain.download(np.float32)[0].copy()
    plain.step(1e-3, 10)
    assert np.array_equal(plain.download(np.float32)[0][2:-2], p0[2:-2])


@pytest.mark.parametrize("nt,nphi", [(64, 128), (37, 50), (5, 7), (3, 4), (130, 1028)])
def test_f32_nonuniform_shapes_incl_ragged(S, nt, nphi):
    """vectorised (num_phi % 4 == 0) and scalar (ragged) paths, tiny grids, TFE on."""
    f = synthetic_field(nt, nphi).astype(np.float32)
    dt = min(1e-3, 0.25 * dt_max(nt, nphi))
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    ora = O.OracleDrainage(nt, nphi)
    sim.upload(f)
    ora.data[:] = f.astype(np.float64)
    for _ in range(3):
        sim.step(dt, 1)
        ora.step(dt, 1)
        got = sim.download()[0]
        assert np.abs(got / ora.data - 1).max() <= 1e-5
        sim.upload(ora.data)  # re-sync so each step is judged on its own
        ora.data[:] = sim.download()[0]


def test_f32_with_tfe_fronts_long_run(S):
    nt, nphi = 64, 128
    sim = make_sim(S, nt, nphi, S.PREC_F32)
    ora = O.OracleDrainage(nt, nphi)
    sim.step(1e-3, 3000)
    ora.step(1e-3, 3000)
    assert ora.active_tfe_count() >= 1
    assert sim.fronts() == ora.fronts()
    got = sim.download()[0]
    assert np.abs(got / ora.data - 1).max() <= 1e-5


@pytest.mark.parametrize("prec,tol", [("PREC_F32", 1e-5), ("PREC_F64", 1e-12)])
@pytest.mark.parametrize("nt,nphi", [(48, 2048), (48, 2050), (40, 100)])
def test_streaming_kernels_with_fronts_alive(S, prec, tol, nt, nphi):
    """One launch per step (n = 1 never takes the resident path) while several marginal-regeneration fronts are alive,
    wrapped ones included: exercises the TFE instantiation of k_march_tma (two strips of 1024 columns) and of the
    direct-load k_march (ragged width 2050: one column per thread; 100 columns: too narrow for the ring), whose front
    membership is a per-row mask ANDed with a per-thread column mask, against the oracle's f64 predicate."""
    sim = make_sim(S, nt, nphi, getattr(S, prec), h_ref=0.0)
    ora = O.OracleDrainage(nt, nphi)
    f = synthetic_field(nt, nphi)
    ff = f.astype(np.float32).astype(np.float64) if prec == "PREC_F32" else f
    sim.upload(ff)
    ora.data[:] = ff
    sim.step(1e-3, 6000)
    ora.step(1e-3, 6000)
    assert ora.active_tfe_count() >= 3 and sim.fronts() == ora.fronts()
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    touched = np.zeros((nt, nphi), bool)
    for i in range(40):
        sim.upload(ora.data)  # judge every step on its own
        before = ora.data.copy()
        if i % 2:
            sim.step(1e-3, 1)
        else:
            sim.step_fused(1e-3, view, 1)
        ora.step(1e-3, 1)
        got = sim.download()[0]
        assert np.abs(got / ora.data - 1).max() <= tol, i
        assert sim.fronts() == ora.fronts()
        # cells a front thinned this step: (1 - 0.5 (1 - ratio) dt) = 0.999925, far outside the drainage increment
        touched |= (ora.data / before) < 0.99995
    assert touched.any() and not touched.all()
    cols = np.nonzero(touched.any(axis=0))[0]
    assert len(cols) > nphi // 16  # fronts are 0.3-0.7 rad wide


def test_critical_cells_are_skipped_and_clamped(S):
    nt, nphi = 32, 64
    f = synthetic_field(nt, nphi)
    f[10:14, 5:9] = 5e-9   # below the 10 nm critical thickness: must be left untouched
    f[20, 30] = 0.0
    for prec in (S.PREC_F32, S.PREC_F64, S.PREC_F64_STRICT):
        sim = make_sim(S, nt, nphi, prec, h_ref=0.0)
        ora = O.OracleDrainage(nt, nphi)
        ff = f.astype(np.float32).astype(np.float64) if prec == S.PREC_F32 else f
        sim.upload(ff)
        ora.data[:] = ff
        sim.step(1e-3, 2)
        ora.step(1e-3, 2)
        got = sim.download()[0]
        assert np.array_equal(got[10:14, 5:9], ff[10:14, 5:9])
        assert got[20, 30] == 0.0
        assert (got >= 0).all()
        assert np.allclose(got, ora.data, rtol=1e-5 if prec == S.PREC_F32 else 1e-12, atol=0)
        assert sim.has_critical_region() == ora.has_critical_region() == True  # noqa: E712


# ------------------------------------------------------------------ API surface
def test_drainage_simulator_api_mirrors_reference_tests(S):
    """The reference's own drainage unit tests (drainage.rs:636-719) through the product path."""
    cfg = S.SimulationConfig.default()
    sim = S.DrainageSimulator.new(cfg)
    assert abs(sim.current_time()) < 1e-16 and not sim.has_critical_region()
    sim.step(0.001)
    assert abs(sim.current_time() - 0.001) < 2.3e-16
    sim2 = S.DrainageSimulator.new(cfg)
    assert abs(sim2.get_thickness(0.5, 1.0) - sim2.get_thickness(1.5, 3.0)) < 1e-15
    top0 = sim2.get_thickness(0.1, 0.0)
    sim2.step(cfg.dt, 100)
    assert sim2.get_thickness(0.1, 0.0) < top0
    assert abs(sim2.get_thickness(1.0, 0.0) - sim2.get_thickness(1.0, 2.0 * math.pi)) < 1e-12
    sim2.reset(600e-9)
    assert abs(sim2.current_time()) < 1e-16
    assert abs(sim2.get_thickness(0.5, 0.5) - 600e-9) < 1e-15
    fld = sim2.thickness_field()
    assert fld.num_theta() == 128 and fld.num_phi() == 256
    fld.set(10, 20, 400e-9)
    assert abs(fld.get(10, 20) - 400e-9) < 1e-15
    assert abs(fld.min_thickness() - 400e-9) < 1e-13
    sim2.set_tfe_thickness_ratio(0.99)
    assert sim2.tfe_thickness_ratio() == 0.95
    sim2.set_marginal_regeneration(False)
    assert not sim2.marginal_regeneration_enabled() and sim2.active_tfe_count() == 0
    assert sim2.critical_thickness() == 10.0 * 1e-9


def test_get_thickness_and_minmax_match_oracle(S):
    nt, nphi = 48, 96
    f = synthetic_field(nt, nphi)
    sim = make_sim(S, nt, nphi, S.PREC_F64)
    ora = O.OracleDrainage(nt, nphi)
    sim.upload(f)
    ora.data[:] = f
    rng = np.random.default_rng(7)
    for th, ph in list(zip(rng.uniform(-0.5, 3.6, 40), rng.uniform(-7, 14, 40))) + [(0, 0), (math.pi, 0), (math.pi / 2, 2 * math.pi)]:
        assert sim.get_thickness(th, ph) == ora.get_thickness(th, ph)
    assert sim.minmax() == (ora.min_thickness(), ora.max_thickness())


@pytest.mark.parametrize("prec", ["PREC_F32", "PREC_F64"])
@pytest.mark.parametrize("nt,nphi,B", [(32, 64, 5), (17, 50, 3), (256, 512, 64), (40, 4100, 1)])
def test_minmax_one_pass_over_the_batch_and_cache(S, prec, nt, nphi, B):
    """min/max over all bubbles in one launch (per-bubble storage offsets, ghost rows skipped, ragged widths on the
    scalar path), and the cached result is dropped by everything that changes the field: the reference's frame loop
    calls min_thickness, max_thickness and has_critical_region back to back (pipeline.rs:1298-1312)."""
    rng = np.random.default_rng(nt * 1000 + nphi + B)
    h0 = np.linspace(300e-9, 700e-9, B)
    f = h0[:, None, None] * (1.0 + 0.2 * rng.uniform(-1, 1, (B, nt, nphi)))
    if prec == "PREC_F32":
        f = f.astype(np.float32).astype(np.float64)
    kw = dict(batch=B, batch_h0_m=h0) if B > 1 else dict(h_ref=0.0)
    sim = make_sim(S, nt, nphi, getattr(S, prec), **kw)
    sim.upload(f)
    n0 = sim.launch_count()
    mn, mx = sim.minmax()
    got = sim.download()
    assert (mn, mx) == (float(got.min()), float(got.max()))
    assert abs(mn / f.min() - 1) < 1e-6 and abs(mx / f.max() - 1) < 1e-6
    launches = sim.launch_count() - n0
    assert sim.minmax() == (mn, mx) and sim.has_critical_region() is False
    assert sim.launch_count() - n0 == launches  # the second and third query reused the pass
    # a step, an upload and a reset each invalidate it
    sim.step(1e-4, 1)
    g = sim.download()
    assert sim.minmax() == (float(g.min()), float(g.max()))
    f2 = f.copy()
    f2[B - 1, nt // 2, nphi - 1] = 5e-9  # below the critical thickness, in the last column of the last bubble
    f2[0, 1, 0] = 2e-6
    sim.upload(f2)
    mn2, mx2 = sim.minmax()
    assert abs(mn2 - 5e-9) < 1e-12 and abs(mx2 - 2e-6) < 1e-12 and sim.has_critical_region() is True
    sim.reset(450e-9)
    mn3, mx3 = sim.minmax()
    assert mn3 == mx3 and abs(mn3 / 450e-9 - 1) < 1e-7
    sim.close()


@pytest.mark.parametrize("prec", ["PREC_F32", "PREC_F64"])
@pytest.mark.parametrize("nt,nphi,B,offset", [(64, 256, 1, False), (48, 2048, 1, False), (32, 128, 7, True), (256, 512, 16, True)])
def test_minmax_from_the_step_kernel_equals_the_pass(S, prec, nt, nphi, B, offset):
    """min/max as a by-product of k_march_tma (ordered-word atomics, pole rows and TFE-thinned cells included) must be
    exactly what a pass over the downloaded field gives -- forced on, in automatic mode (switches on after two
    step+query rounds), with fronts alive, and for one launch per step as well as the last step of a multi-step call."""
    rng = np.random.default_rng(nt + nphi + B)
    h0 = np.linspace(300e-9, 700e-9, B)
    f = h0[:, None, None] * (1.0 + 0.2 * rng.uniform(-1, 1, (B, nt, nphi)))
    if prec == "PREC_F32":
        f = f.astype(np.float32).astype(np.float64)
    kw = dict(batch=B, batch_h0_m=h0) if offset else dict(h_ref=0.0, batch=B)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)

    def check(sim, what):
        g = sim.download()
        assert sim.minmax() == (float(g.min()), float(g.max())), what

    for mode in (1, 2):
        sim = make_sim(S, nt, nphi, getattr(S, prec), **kw)
        sim.track_minmax(mode)
        sim.upload(f)
        for i in range(6):
            n0 = sim.launch_count()
            if i % 2:
                sim.step(1e-4, 1)
            else:
                sim.step_fused(1e-4, view, 1)
            n1 = sim.launch_count()
            check(sim, (mode, i))
            assert n1 - n0 >= 1
        sim.step(1e-4, 3)          # multi-step call: the last step's by-product (or a pass, on the resident path)
        check(sim, (mode, "multi"))
        sim.step(1e-3, 700)        # past t = 0.5 s: fronts alive
        sim.step_fused(1e-3, view, 1)
        assert sim.active_tfe_count() >= 1
        check(sim, (mode, "fronts"))
        f2 = f.copy()
        f2[B - 1, 0, :] = 900e-9   # a pole row above everything: the pole mean of the next step comes from ring 1 instead
        f2[0, nt // 2, 3] = 50e-9
        sim.upload(f2)
        check(sim, (mode, "upload"))
        sim.step_fused(1e-4, view, 1)
        check(sim, (mode, "after upload"))
        sim.close()


def test_minmax_tracking_switches_itself_on_and_off(S):
    """Automatic mode: the third step+query round no longer needs a pass (one launch per step and the one-thread
    get_thickness sample the mirror's frame fetch includes, no reduction pass); eight steps without a query switch it off
    again."""
    sim = make_sim(S, 64, 256, S.PREC_F32, h_ref=0.0)
    sim.upload(synthetic_field(64, 256))
    per_round = []
    for _ in range(5):
        n0 = sim.launch_count()
        sim.step(1e-4, 1)
        sim.minmax()
        sim.has_critical_region()
        per_round.append(sim.launch_count() - n0)
    assert per_round[0] >= 3 and per_round[-1] == 2, per_round  # step + sample (+ a pass in the first rounds)
    for _ in range(10):
        sim.step(1e-4, 1)
    n0 = sim.launch_count()
    sim.minmax()
    assert sim.launch_count() - n0 == 2  # a pass again (+ the sample)
    sim.close()


def test_batched_bubbles_match_independent_oracles(S):
    """foam batch (BASELINE config 4): independent bubbles with their own h0 and radius, no exchange."""
    nt, nphi, B = 32, 64, 5
    h0 = np.linspace(300e-9, 700e-9, B)
    rad = np.linspace(0.0125, 0.1, B)
    for prec, tol in ((S.PREC_F32, 1e-5), (S.PREC_F64, 1e-12)):
        sim = make_sim(S, nt, nphi, prec, batch=B, batch_h0_m=h0, batch_radius_m=rad)
        fields = np.stack([synthetic_field(nt, nphi, seed=b, h0_nm=h0[b] * 1e9) for b in range(B)])
        if prec == S.PREC_F32:
            fields = fields.astype(np.float32).astype(np.float64)
            sim2 = make_sim(S, nt, nphi, prec, h_ref=0.0, batch=B, batch_h0_m=h0, batch_radius_m=rad)
        else:
            sim2 = sim
        sim2.upload(fields)
        sim2.step(1e-3, 5)
        got = sim2.download()
        for b in range(B):
            ora = O.OracleDrainage(nt, nphi, h0_m=h0[b], radius_m=rad[b])
            ora.data[:] = fields[b]
            ora.step(1e-3, 5)
            assert np.abs(got[b] / ora.data - 1).max() <= 5 * tol, (prec, b)
        # uniform start with per-bubble reference (delta storage for f32)
        fresh = make_sim(S, nt, nphi, prec, batch=B, batch_h0_m=h0, batch_radius_m=rad)
        fresh.step(1e-3, 3)
        got = fresh.download()
        for b in range(B):
            ora = O.OracleDrainage(nt, nphi, h0_m=h0[b], radius_m=rad[b])
            ora.step(1e-3, 3)
            inc_ref = ora.data - h0[b]
            assert np.abs((got[b] - h0[b]) - inc_ref).max() <= 2e-3 * np.abs(inc_ref).max()


# ------------------------------------------------------------------ WGSL semantics
@pytest.mark.parametrize("marangoni", [False, True])
@pytest.mark.parametrize("w,h", [(128, 64), (50, 37)])
def test_wgsl_semantics_match_f32_oracle(S, w, h, marangoni):
    sim = S.GPUDrainageSimulator(500e-9, w, h)
    rng = np.random.default_rng(3)
    hf = synthetic_field(h, w, h0_nm=5000.0, noise_nm=20.0).astype(np.float32)
    cf = np.clip(0.5 + 0.2 * rng.standard_normal((h, w)), 0, 1).astype(np.float32)
    p = O.wgsl_params(grid_width=w, grid_height=h, marangoni_enabled=int(marangoni), dt=0.001)
    sim.set_marangoni_enabled(marangoni)
    sim.upload(hf)
    sim.upload_concentration(cf)
    ho, co = hf, cf
    for _ in range(4):
        ho, co = O.wgsl_step(p, ho, co)
    sim.dispatch(4)
    gh, gc = sim.download(np.float32)[0], sim.download_concentration()[0]
    assert np.array_equal(gh[0], hf[0]) and np.array_equal(gh[-1], hf[-1])  # poles frozen
    assert np.abs(gh.astype(np.float64) / ho - 1).max() <= 1e-5
    assert np.abs(gc.astype(np.float64) - co).max() <= 1e-5


def test_gpu_drainage_simulator_api(S):
    sim = S.GPUDrainageSimulator(500e-9, 128, 64)
    assert sim.grid_size() == (128, 64)
    assert sim.params().critical_thickness == pytest.approx(30e-9)
    sim.step(1.0 / 60.0)  # disabled: no-op (gpu_drainage.rs:339-341)
    assert sim.current_time() == 0.0
    sim.enabled = True
    sim.step(1.0 / 60.0)
    assert sim.current_time() == pytest.approx(100.0 / 60.0, rel=1e-6)
    sim.set_dt(0.002)
    sim.set_physics(0.002, 1100.0, 9.0, 2e-9)
    sim.set_bubble_radius(0.03)
    sim.set_marangoni_params(0.07, 0.04, 2e-9, 0.02)
    assert sim.params().dt == 0.002 and sim.params().bubble_radius == 0.03
    sim.reset(400e-9)
    assert sim.current_time() == 0.0
    assert np.all(sim.download(np.float32) == np.float32(400e-9))
    assert np.all(sim.download_concentration() == np.float32(0.5))
    assert sim.current_thickness_buffer() != 0 and sim.concentration_buffer() != 0


# ------------------------------------------------------------------ interference
def test_interference_calculator_matches_oracle(S):
    ic = S.InterferenceCalculator(1.33)
    oc = O.OracleInterference(1.33)
    rng = np.random.default_rng(11)
    d = np.concatenate([rng.uniform(0, 2000, 4000), [0.0, -5.0, 500.0, 300.0, 100.0, 1000.0]])
    c = np.concatenate([rng.uniform(0, 1, 4000), [1.0, 0.5, 1.0, 1.0, 1.0, 0.0]])
    for gamma in (True, False):
        got, ref = ic.batch(d, c, gamma), oc.batch(d, c, gamma)
        assert np.abs(got - ref).max() <= 1e-9
    fr = ic.calculate_fresnel_reflection(1.0)
    assert abs(fr.reflectance_average - ((1.0 - 1.33) / (1.0 + 1.33)) ** 2) < 1e-15  # interference.rs:672-679
    assert abs(ic.calculate_transmission_angle_cos(1.0) - 1.0) < 1e-15
    assert ic.calculate_reflected_intensity(0.0, 1.0, 550.0) == 0.0
    assert ic.calculate_interference_color(1000.0, 0.0).to_u8_tuple() == (255, 255, 255)
    assert S.RgbColor.new(1.0, 0.5, 0.0).to_u8_tuple() == (255, 128, 0)  # interference.rs:627-633
    t = ic.generate_color_lookup_table(0.0, 1000.0, 100, 1.0)
    assert len(t) == 100
    t2 = ic.generate_2d_color_lookup_table(0.0, 500.0, 10, 5)
    assert len(t2) == 10 and len(t2[0]) == 5
    ref2 = oc.lut_2d(0.0, 500.0, 10, 5)
    assert abs(t2[3][2].green - ref2[3, 2, 1]) < 1e-9


def test_lut_bytes_match_oracle_within_one_lsb(S):
    for n, inten in ((1.33, 1.0), (1.4, 2.5)):
        got = S.generate_interference_lut(n, inten).reshape(64, 256, 4)
        ref = O.generate_interference_lut(n, inten)
        diff = np.abs(got.astype(int) - ref.astype(int))
        print(f"\n[lut] n={n}: {np.count_nonzero(diff)} of {diff.size} bytes differ, max {diff.max()}")
        assert got.size == 65536 and diff.max() <= 1
        assert np.count_nonzero(diff) <= 0.002 * diff.size
    wl = 380.0 + 400.0 * np.arange(64) / 63.0
    got = S.generate_spectral_lut(1.33, 1.0, wl, 256, 64, 2000.0)
    ref = O.generate_lut_spectral(1.33, 1.0, wl, 256, 64, 2000.0)
    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1


def _oracle_view_cos(nt, nphi, radius, cam):
    th = np.arange(nt) * (math.pi / (nt - 1))
    ph = np.arange(nphi) * (2 * math.pi / nphi)
    out = np.empty((nt, nphi))
    for i in range(nt):
        for j in range(nphi):
            out[i, j] = O.view_cos(th[i], ph[j], radius, cam)
    return out


@pytest.mark.parametrize("mode_name", ["EXACT3", "EXACT3_LINEAR"])
def test_fused_exact3_rgb_within_1e4(S, mode_name):
    nt, nphi = 48, 96
    mode = getattr(S, "RGB_" + mode_name)
    f = synthetic_field(nt, nphi, h0_nm=600.0, noise_nm=40.0)
    dt = 0.25 * dt_max(nt, nphi)
    view = S.View(mode=mode, format=S.OUT_RGB_F32)
    cos = _oracle_view_cos(nt, nphi, 0.025, view.cam_pos)
    oc = O.OracleInterference(1.33)
    for prec in (S.PREC_F32, S.PREC_F64, S.PREC_F64_STRICT):
        sim = make_sim(S, nt, nphi, prec, h_ref=0.0)
        ora = O.OracleDrainage(nt, nphi)
        ff = f.astype(np.float32).astype(np.float64) if prec == S.PREC_F32 else f
        sim.upload(ff)
        ora.data[:] = ff
        sim.step_fused(dt, view, 2)
        ora.step(dt, 2)
        rgb = sim.download_rgb(S.OUT_RGB_F32)[0]
        ref = oc.batch(ora.data * 1e9, cos, gamma=(mode == S.RGB_EXACT3)).reshape(nt, nphi, 3)
        err = np.abs(rgb - ref).max()
        print(f"\n[rgb] {mode_name} prec={prec}: max abs err {err:.2e}")
        assert err <= 1e-4
    # RGBA8 output: <= 1 LSB of round(255 x)
    view8 = S.View(mode=mode, format=S.OUT_RGBA8)
    sim.shade(view8)
    px = sim.download_rgb(S.OUT_RGBA8)[0]
    ref8 = np.floor(ref * 255.0 + 0.5)
    assert np.abs(px[..., :3].astype(int) - ref8).max() <= 1
    assert (px[..., 3] == 255).all()


def test_fused_lut_rgba8_within_one_lsb(S):
    nt, nphi = 64, 128
    f = synthetic_field(nt, nphi, h0_nm=700.0, noise_nm=60.0).astype(np.float32)
    dt = 0.25 * dt_max(nt, nphi)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8, intensity=4.0)
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    ora = O.OracleDrainage(nt, nphi)
    sim.upload(f)
    ora.data[:] = f.astype(np.float64)
    sim.step_fused(dt, view, 1)
    ora.step(dt, 1)
    px = sim.download_rgb()[0]
    lut = O.generate_interference_lut(1.33, 1.0)
    cos = _oracle_view_cos(nt, nphi, 0.025, view.cam_pos)
    ref = np.empty((nt, nphi, 3))
    for i in range(nt):
        for j in range(nphi):
            ref[i, j] = O.lut_shade(lut, float(np.float32(ora.data[i, j] * 1e9)), float(np.float32(cos[i, j])), 4.0,
                                    weight_bits=8)  # texture-unit filtering, as the reference's textureSample
    ref8 = np.floor(ref * 255.0 + 0.5)
    diff = np.abs(px[..., :3].astype(int) - ref8)
    print(f"\n[lut-shade] max diff {diff.max()} LSB, {np.count_nonzero(diff)} of {diff.size} channels differ")
    assert diff.max() <= 1
    assert (px[..., 3] == 255).all()
    # float output of the same mode: texture filtering uses 8-bit weights -> 2e-3 absolute at intensity 4
    sim.shade(S.View(mode=S.RGB_LUT, format=S.OUT_RGB_F32, intensity=1.0))
    rgbf = sim.download_rgb(S.OUT_RGB_F32)[0]
    ref1 = np.empty((nt, nphi, 3))
    for i in range(nt):
        for j in range(nphi):
            ref1[i, j] = O.lut_shade(lut, float(np.float32(ora.data[i, j] * 1e9)), float(np.float32(cos[i, j])), 1.0)
    assert np.abs(rgbf - ref1).max() <= 2.5e-3


def test_schlick_wall_mode_matches_oracle(S):
    nt, nphi = 24, 48
    f = synthetic_field(nt, nphi, h0_nm=400.0, noise_nm=30.0).astype(np.float32)
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    sim.upload(f)
    view = S.View(mode=S.RGB_SCHLICK_WALL, format=S.OUT_RGB_F32)
    sim.shade(view)
    rgb = sim.download_rgb(S.OUT_RGB_F32)[0]
    cos = _oracle_view_cos(nt, nphi, 0.025, view.cam_pos)
    for i in range(0, nt, 3):
        for j in range(0, nphi, 5):
            ref = O.schlick_wall_rgb(float(f[i, j] * np.float32(1e9)), float(np.float32(cos[i, j])), 1.33)
            assert np.abs(rgb[i, j] - np.array(ref)).max() <= 2e-4


def test_errors_are_reported_not_fatal(S):
    cfg = S.SimulationConfig.default()
    with pytest.raises(S.SbxError):
        S.DrainageSimulator(cfg, num_theta=2, num_phi=8)
    with pytest.raises(S.SbxError):
        S.DrainageSimulator(cfg, num_theta=16, num_phi=32, row_begin=1, row_end=8)
    sim = S.DrainageSimulator(cfg, num_theta=16, num_phi=32)
    with pytest.raises(ValueError):
        sim.upload(np.zeros((3, 3)))
    with pytest.raises(S.SbxError):
        sim.download_rgb()


def test_step_host_pipelined_equals_separate_calls(S, monkeypatch):
    """sbx_drainage_step_host (chunk-pipelined H2D / kernels / D2H) against upload + step_fused + download."""
    nt, nphi = 1300, 256
    monkeypatch.setenv("SBX_HOST_CHUNK_ROWS", "504")  # three chunks (the default is ~32 MiB per chunk)
    f = synthetic_field(nt, nphi, dtype=np.float32)
    dt = 0.25 * dt_max(nt, nphi)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    ref = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    out_h = np.empty((1, nt, nphi), np.float32)
    out_px = np.empty((1, nt, nphi, 4), np.uint8)
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    cur = f[None].copy()
    for step in range(3):
        ref.upload(cur)
        ref.step_fused(dt, view, 1)
        want_h, want_px = ref.download(np.float32), ref.download_rgb()
        sim.step_host(dt, 1, view, cur, out_h, out_px)
        assert np.array_equal(out_h, want_h), step
        assert np.array_equal(out_px, want_px), step
        cur = want_h.copy()
    assert sim.current_time() == ref.current_time()
    # the non-pipelined fallback (f64 host buffers) gives the same field
    sim64 = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    o64 = np.empty((1, nt, nphi), np.float64)
    sim64.step_host(dt, 1, view, f[None].astype(np.float64), o64, out_px)
    sim.step_host(dt, 1, view, f[None].copy(), out_h, None)
    assert np.array_equal(o64.astype(np.float32), out_h)


def test_wgsl_sin_safe_threshold_on_weak_scaled_grid(S):
    """65,536 theta rows (the 8-GPU weak-scaled grid): sin(theta_1) = 4.8e-5 is below the shader's 1e-4 guard
    (drainage.wgsl:101), so the rings next to the poles run on the clamped value -- and must still match."""
    w, h = 16, 65536
    sim = S.GPUDrainageSimulator(500e-9, w, h)
    rng = np.random.default_rng(5)
    hf = (5e-6 * (1 + 0.05 * rng.standard_normal((h, w)))).astype(np.float32)
    cf = np.clip(0.5 + 0.1 * rng.standard_normal((h, w)), 0, 1).astype(np.float32)
    p = O.wgsl_params(grid_width=w, grid_height=h, marangoni_enabled=1, dt=1e-9)
    sim.set_marangoni_enabled(True)
    sim.set_dt(1e-9)
    sim.upload(hf)
    sim.upload_concentration(cf)
    ho, co = O.wgsl_step(p, hf, cf)
    ho, co = O.wgsl_step(p, ho, co)
    sim.dispatch(2)
    gh, gc = sim.download(np.float32)[0], sim.download_concentration()[0]
    assert np.abs(gh[:4].astype(np.float64) / ho[:4] - 1).max() <= 1e-5
    assert np.abs(gh.astype(np.float64) / ho - 1).max() <= 1e-5
    assert np.abs(gc.astype(np.float64) - co).max() <= 1e-5


def test_nan_inf_and_huge_cells_follow_the_reference(S):
    """Non-finite and extreme cells: f64::max(NaN, 0.0) is 0.0 in Rust (drainage.rs:378), comparisons with NaN are
    false (so a NaN cell is not 'below critical'), +inf stays +inf.  The strict mode must agree bit for bit, f32 in
    the same places."""
    nt, nphi = 24, 32
    f = synthetic_field(nt, nphi)
    f[5, 7] = np.nan
    f[9, 0] = np.inf
    f[12, 31] = 1e-3      # a millimetre-thick cell: huge h^3 term, clamps to 0
    f[15, 3] = 0.0
    f[1, 4] = np.nan      # in the ring that feeds the top pole mean
    ora = O.OracleDrainage(nt, nphi)
    ora.data[:] = f
    ora.step(1e-3, 3)
    strict = make_sim(S, nt, nphi, S.PREC_F64_STRICT)
    strict.upload(f)
    strict.step(1e-3, 3)
    got = strict.download()[0]
    assert np.array_equal(np.isnan(got), np.isnan(ora.data))
    assert np.array_equal(got[~np.isnan(got)], ora.data[~np.isnan(ora.data)])
    f32sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    f32sim.upload(f)
    f32sim.step(1e-3, 3)
    g32 = f32sim.download()[0]
    assert np.array_equal(np.isnan(g32), np.isnan(ora.data))
    fin = np.isfinite(ora.data) & (ora.data > 0)
    assert np.array_equal(np.isinf(g32), np.isinf(ora.data))
    assert np.array_equal(g32 == 0, ora.data == 0)
    assert np.abs(g32[fin] / ora.data[fin] - 1).max() <= 1e-5


def test_caustic_map_consumer_matches_oracle(S):
    """f2: the caustic compute pass that reads the thickness buffer right after the drainage step
    (pipeline.rs:1808-1810, caustics_compute.wgsl:29-126)."""
    nt, nphi = 64, 128  # the reference's fixed GPU grid (pipeline.rs:673-674)
    rng = np.random.default_rng(2)
    for scale, kw in ((1.0, {}), (1e6, dict(focal_length=0.7, branch_threshold=1e-4)), (3e5, dict(caustic_sharpness=2.5))):
        # thickness in metres gives a flat map (2.0); scaled-up fields exercise focusing, branches and the clamp
        f = (synthetic_field(nt, nphi, noise_nm=40.0) * scale).astype(np.float32)
        f += (scale * 2e-7 * rng.standard_normal((nt, nphi))).astype(np.float32)
        sim = S.GPUDrainageSimulator(500e-9, nphi, nt)
        sim.upload(f)
        got = sim.caustic_map(**kw)[0]
        ref = O.caustic_map(O.caustic_params(grid_width=nphi, grid_height=nt, **kw), f)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), scale
        assert 0.0 <= got.min() and got.max() <= 10.0
        if scale == 1.0:
            assert np.allclose(got, 2.0, atol=1e-4)
    # through a CPU-semantics handle with offset storage and f64 storage, after a fused step
    cpu = make_sim(S, nt, nphi, S.PREC_F32)
    cpu.step(1e-3, 3)
    m = cpu.caustic_map()[0]
    assert np.allclose(m, 2.0, atol=1e-4)
    d = make_sim(S, nt, nphi, S.PREC_F64)
    d.upload(synthetic_field(nt, nphi) * 1e6)
    ref = O.caustic_map(O.caustic_params(grid_width=nphi, grid_height=nt), (synthetic_field(nt, nphi) * 1e6).astype(np.float32))
    assert np.abs(d.caustic_map()[0] - ref).max() <= 1e-4 * np.abs(ref).max()


def test_two_beam_schlick_documented_model(S):
    """SBX_RGB_SCHLICK_TWO_BEAM: the model README/north_star describe (docs/key-insights.md:288-309); checked
    against the restated formula, fused and shade-only, float and RGBA8 output, f32 and f64 storage."""
    nt, nphi = 48, 128
    f = synthetic_field(nt, nphi, h0_nm=450.0, noise_nm=60.0).astype(np.float32)
    view = S.View(mode=S.RGB_SCHLICK_TWO_BEAM, format=S.OUT_RGB_F32)
    cos = _oracle_view_cos(nt, nphi, 0.025, view.cam_pos)
    ref = np.empty((nt, nphi, 3))
    for i in range(nt):
        for j in range(nphi):
            ref[i, j] = O.schlick_two_beam_rgb(float(f[i, j] * np.float32(1e9)), float(np.float32(cos[i, j])), 1.33)
    ref = np.clip(ref, 0.0, 1.0)
    for prec in (S.PREC_F32, S.PREC_F64):
        sim = make_sim(S, nt, nphi, prec, h_ref=0.0)
        sim.upload(f)
        sim.shade(view)
        rgb = sim.download_rgb(S.OUT_RGB_F32)[0]
        assert np.abs(rgb - ref).max() <= 1e-4
        sim.shade(S.View(mode=S.RGB_SCHLICK_TWO_BEAM, format=S.OUT_RGBA8))
        px = sim.download_rgb()[0]
        assert np.abs(px[..., :3].astype(int) - np.floor(ref * 255 + 0.5)).max() <= 1
    # fused through the TMA kernel: colour of the stepped field equals shade-only of the same field
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    sim.upload(f)
    v8 = S.View(mode=S.RGB_SCHLICK_TWO_BEAM, format=S.OUT_RGBA8)
    sim.step_fused(1e-4, v8, 2)
    fused = sim.download_rgb()[0]
    sim.shade(v8)
    assert np.abs(fused.astype(int) - sim.download_rgb()[0].astype(int)).max() <= 1


def test_branched_flow_sampling_side_matches_oracle(S):
    """f2: nearest-texel thickness and eps = 0.01 central-difference gradient at uv positions
    (branched_flow_compute.wgsl:187-210), on the reference's 128x64 grid and on a larger one."""
    rng = np.random.default_rng(9)
    for w, h in ((128, 64), (1024, 512)):
        f = (synthetic_field(h, w, noise_nm=30.0)).astype(np.float32)
        sim = S.GPUDrainageSimulator(500e-9, w, h)
        sim.upload(f)
        uv = np.concatenate([rng.uniform(-0.05, 1.05, (4000, 2)), [[0, 0], [1, 1], [0.5, 0.0], [0.999999, 0.5]]]).astype(np.float32)
        gh, gg = sim.sample_uv(uv)
        rh, rg = O.sample_thickness_uv(f, uv)
        assert np.array_equal(gh, rh)
        assert np.array_equal(gg, rg)
        # sample_thickness(normal): libm and CUDA atan2f/acosf differ by an ulp or two, which can move a sample sitting
        # on a texel boundary to the neighbouring texel; everywhere else the values are identical
        nrm = rng.normal(size=(4000, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        nrm = np.concatenate([nrm, [[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 1e-9]]]).astype(np.float32)
        gn = sim.sample_normal(nrm, scale=2.5)
        rn, ruv = O.sample_thickness_normal(f, nrm, 2.5)
        fx, fy = ruv[:, 0] * np.float32(w - 1), ruv[:, 1] * np.float32(h - 1)
        tie = (np.abs(fx - np.round(fx)) < 1e-3) | (np.abs(fy - np.round(fy)) < 1e-3)
        assert np.array_equal(gn[~tie], rn[~tie])
        assert tie.mean() < 0.02


def test_many_strips_fall_back_to_the_pole_kernel(S):
    """More than 128 strips per ring: the pole rows are not folded into the TMA kernel (the strips could not all be
    co-resident) and k_poles finishes them; results are the same."""
    nt, nphi = 6, 129 * 1024
    f = synthetic_field(nt, nphi).astype(np.float32)
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    ora = O.OracleDrainage(nt, nphi)
    sim.upload(f)
    ora.data[:] = f.astype(np.float64)
    dt = 0.25 * dt_max(nt, nphi)
    sim.step(dt, 2)
    ora.step(dt, 2)
    got = sim.download()[0]
    assert np.abs(got / ora.data - 1).max() <= 1e-5
    assert np.all(got[0] == got[0, 0]) and np.all(got[-1] == got[-1, 0])
