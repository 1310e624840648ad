"""GPU tests added in round 2: checkpoint / resume of the TFE state machine, batched ThicknessField views, the
device-side wait status, full-size increment parity (a frozen field must fail), reference-vector consumption."""
import glob
import json
import math
import os

import numpy as np
import pytest

import oracle_lib as O
from synth import dt_max, synthetic_field

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def S():
    import soap_bubble_simulation_b200 as S

    return S


def make_sim(S, nt, np_, precision, h_ref=None, **kw):
    return S.DrainageSimulator(S.SimulationConfig.default(), precision=precision, num_theta=nt, num_phi=np_, h_ref=h_ref, **kw)


# ------------------------------------------------------------------ checkpoint / resume (sbx_drainage_get_state / set_state)
@pytest.mark.parametrize("prec", [2, 0])
def test_checkpoint_resume_is_exact(S, prec):
    """Field (download / upload) + (current_time, next_tfe_spawn_time, tfe_fronts) restore a simulator exactly:
    a resumed run and an uninterrupted run agree bit for bit, fronts included (drainage.rs:237-255)."""
    nt, nphi = 48, 96
    a = make_sim(S, nt, nphi, prec)
    a.step(5e-3, 700)  # t = 3.5 s: fronts alive
    assert a.active_tfe_count() >= 1
    field = a.download(np.float64)
    t, nxt, fronts = a.get_state()
    assert t == a.current_time() and fronts == a.fronts()
    b = make_sim(S, nt, nphi, prec)
    b.upload(field)
    b.set_state(t, nxt, fronts)
    a.step(5e-3, 900)  # crosses spawns and a retirement
    b.step(5e-3, 900)
    assert a.fronts() == b.fronts() and a.current_time() == b.current_time()
    assert np.array_equal(a.download(np.float64), b.download(np.float64))
    if prec == 2:  # and both equal the oracle run from t = 0
        ora = O.OracleDrainage(nt, nphi)
        ora.step(5e-3, 1600)
        assert np.array_equal(b.download()[0], ora.data) and b.fronts() == ora.fronts()
    with pytest.raises(ValueError):
        b.set_state(0.0, 0.5, [(0.0, 0.0, 0.3, 0.05)] * 9)
    with pytest.raises(S.SbxError):
        b.set_state(float("nan"), 0.5, [])


def test_restored_fronts_drive_the_tma_kernel_like_the_oracle(S):
    """bench.py times the fronts-alive instantiation by restoring the TFE state of t = 15 s on the big grid; the
    same move at a size the oracle finishes: restored state + k_march_tma<TFE=true> vs the oracle given the same state."""
    nt, nphi = 96, 1024
    small = make_sim(S, 32, 64, 0)
    small.step(5e-3, 3000)
    t, nxt, fronts = small.get_state()
    assert len(fronts) >= 4
    f = synthetic_field(nt, nphi, dtype=np.float32)
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    sim.upload(f)
    sim.set_state(t, nxt, fronts)
    ora = O.OracleDrainage(nt, nphi)
    ora.data[:] = f.astype(np.float64)
    ora.set_state(t, nxt, fronts)
    dt = 0.25 * dt_max(nt, nphi)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    sim.step_fused(dt, view, 5)
    ora.step(dt, 5)
    assert sim.fronts() == ora.fronts()
    assert float(np.abs(sim.download()[0] / ora.data - 1).max()) <= 1e-5


# ------------------------------------------------------------------ host mirror fixes
def test_thickness_field_view_on_a_batched_handle(S):
    """thickness_field() on a foam handle mirrors bubble 0; edits are written back into the whole batch and are
    seen by step_host / shade (they used to raise or run on the stale device field)."""
    nt, nphi, B = 16, 32, 3
    h0 = np.array([300e-9, 500e-9, 700e-9])
    sim = make_sim(S, nt, nphi, S.PREC_F64, batch=B, batch_h0_m=h0)
    fld = sim.thickness_field()
    assert fld.get(3, 4) == h0[0]
    fld.set(3, 4, 350e-9)
    assert sim.get_thickness(3 * math.pi / (nt - 1), 4 * 2 * math.pi / nphi) == pytest.approx(350e-9, rel=1e-12)
    got = sim.download()
    assert got[0, 3, 4] == 350e-9 and np.all(got[1] == h0[1]) and np.all(got[2] == h0[2])
    fld.data_mut()[5, :] = 123e-9
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    sim.shade(view)  # flushes the pending edit first
    assert np.all(sim.download()[0, 5] == 123e-9)
    fld.set(6, 0, 99e-9)
    out = np.empty((B, nt, nphi), np.float64)
    sim.step_host(0.0, 0, None, None, out, None)  # no input buffer: the edited field is the input
    assert out[0, 6, 0] == 99e-9


def test_wgsl_radius_change_reaches_the_pole_rows_colour(S):
    """update_params with a new bubble_radius: pole rows (shaded from BubbleParam.radius) and interior rows (shaded
    from params.bubble_radius) must use the same radius afterwards."""
    w, h = 64, 32
    sim = S.GPUDrainageSimulator(500e-9, w, h)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    sim.dispatch(1, view)
    sim.set_bubble_radius(0.08)
    sim.dispatch(1, view)
    px = sim.download_rgb()[0]
    ref = S.GPUDrainageSimulator(500e-9, w, h)
    ref.set_bubble_radius(0.08)
    ref.dispatch(2, view)
    assert np.array_equal(px, ref.download_rgb()[0])


def test_device_wait_status_is_clean_after_folded_pole_steps(S):
    """The folded pole rows wait for their sibling strips with a bounded spin; a healthy run reports 0."""
    from soap_bubble_simulation_b200.sharding import Comm

    sim = make_sim(S, 64, 4096, S.PREC_F32, h_ref=0.0)
    sim.upload(synthetic_field(64, 4096, dtype=np.float32))
    sim.step_fused(1e-7, S.View(), 20)
    assert Comm.peer_status(sim) == 0


# ------------------------------------------------------------------ full-size parity in a regime where f32 increments are resolved
def _increment_case(S, nt, nphi, steps, prec, h_nm=30000.0):
    """A 30 um film with 1 % noise that reaches the poles.  At dt = 0.25 dt_max of a large grid almost every increment
    of the default 500 nm field is below the f32 ulp (SURVEY 7 trap (i)); what plain f32 storage does resolve is the
    phi-diffusion of rough rings next to the poles (coefficient x dt = 0.125 / k^2 in ring k), so that is where the
    sup norm of the increment lives.  Rows are sampled over the whole sphere (oracle: window mode around each band);
    the gate is the global one of SURVEY 8d: ||dh_gpu - dh_ref||inf <= 1e-3 ||dh_ref||inf."""
    f32 = synthetic_field(nt, nphi, dtype=np.float32, h0_nm=h_nm, noise_nm=h_nm * 0.01, polar_noise=True)
    dt = 0.25 * dt_max(nt, nphi)
    sim = make_sim(S, nt, nphi, prec, h_ref=0.0)
    sim.upload(f32)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    sim.step_fused(dt, view, steps)
    got = sim.download(np.float64)[0]
    del sim
    worst_rel, err_max, inc_max, moved = 0.0, 0.0, 0.0, 0
    bands = [0, 40, nt // 7, nt // 2 - 20, nt - nt // 5, nt - 80, nt - 40]  # both poles, the equator, strips in between
    for g0 in bands:
        w0, w1 = max(0, g0 - steps - 1), min(nt, g0 + 40 + steps + 1)
        ora = O.OracleDrainage(nt, nphi, window=(w0, w1 - w0), threads=os.cpu_count() or 1)
        start = f32[w0:w1].astype(np.float64)
        ora.data[:] = start
        ora.step(dt, steps)
        lo, hi = g0 - w0, g0 - w0 + 40
        ref, ini, g = ora.data[lo:hi], start[lo:hi], got[g0:g0 + 40]
        worst_rel = max(worst_rel, float(np.abs(g / ref - 1).max()))
        err_max = max(err_max, float(np.abs((g - ini) - (ref - ini)).max()))
        inc_max = max(inc_max, float(np.abs(ref - ini).max()))
        moved += int(np.count_nonzero(g != ini))
    assert moved > 1000, "the field did not move"
    return worst_rel, err_max / inc_max


@pytest.mark.parametrize("nt,nphi,steps", [(2048, 4096, 3), (8192, 16384, 1)])
def test_full_size_f32_increments_match_the_oracle(S, nt, nphi, steps):
    """SURVEY 8d's increment gate at the sizes of configs 2 and 3 through k_march_tma (multi-strip, multi-band, PDL,
    L2 hints): ||dh_gpu - dh_ref||inf <= 1e-3 ||dh_ref||inf.  An identity kernel fails this by construction
    (its increment error is 100 %)."""
    rel, inc = _increment_case(S, nt, nphi, steps, S.PREC_F32)
    assert rel <= 1e-5, rel
    assert inc <= 1e-3, inc


def test_full_size_f64_increments_match_the_oracle(S):
    rel, inc = _increment_case(S, 2048, 4096, 3, S.PREC_F64)
    assert rel <= 1e-12 and inc <= 1e-9, (rel, inc)


def test_c2_thousand_step_drift(S):
    """Stated 1000-step drift at config 2's size (2048 x 4096, 30 um film, dt = 0.25 dt_max): f32 and f64 storage vs
    the f64 oracle on latitude slabs.  The numbers are recorded in BASELINE.md; the gate is north_star's 1e-5 per
    step accumulated linearly at worst (1e-2), in practice orders of magnitude below."""
    nt, nphi, steps = 2048, 4096, 1000
    h_nm = 30000.0
    f32 = synthetic_field(nt, nphi, dtype=np.float32, h0_nm=h_nm, noise_nm=h_nm * 0.01, polar_noise=True)
    dt = 0.25 * dt_max(nt, nphi)
    res = {}
    # the window oracle needs `steps` rows of margin on each cut side: two slabs, one holding each pole
    slabs = [(0, 24), (nt - 24, nt)]
    refs = []
    for a, b in slabs:
        w0, w1 = max(0, a - steps - 1), min(nt, b + steps + 1)
        ora = O.OracleDrainage(nt, nphi, window=(w0, w1 - w0), threads=os.cpu_count() or 1)
        ora.data[:] = f32[w0:w1].astype(np.float64)
        ora.set_marginal_regeneration(False)
        ora.step(dt, steps)
        refs.append(np.array(ora.data[a - w0:b - w0]))
    for name, prec in (("f32", S.PREC_F32), ("f64", S.PREC_F64)):
        sim = make_sim(S, nt, nphi, prec, h_ref=0.0)
        sim.set_marginal_regeneration(False)
        sim.upload(f32)
        sim.step(dt, steps)
        got = sim.download(np.float64)[0]
        res[name] = max(float(np.abs(got[a:b] / r - 1).max()) for (a, b), r in zip(slabs, refs))
        del sim
    print("C2 1000-step drift (max rel, polar slabs):", json.dumps(res))
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "c2_drift.json"), "w") as fh:
        json.dump(res, fh)
    assert res["f32"] <= 1e-3 and res["f64"] <= 1e-10, res


# ------------------------------------------------------------------ reference vectors (written by scripts/make_reference_vectors.rs)
def _vector_files():
    return sorted(glob.glob(os.path.join(HERE, "golden", "drainage_*.bin")))


@pytest.mark.skipif(not _vector_files(), reason="no tests/golden/drainage_*.bin (run scripts/make_reference_vectors.rs "
                                                "with cargo on any machine to pin the drainage half against the real reference)")
@pytest.mark.parametrize("path", _vector_files() or ["-"])
def test_strict_mode_reproduces_reference_vectors(S, path):
    """Vectors produced by the reference's own DrainageSimulator::step: the strict f64 mode must be bit-identical."""
    from reference_vectors import read_vector

    v = read_vector(path)
    sim = make_sim(S, v.num_theta, v.num_phi, S.PREC_F64_STRICT)
    sim.upload(v.initial)
    done = 0
    for steps, field in v.checkpoints:
        sim.step(v.dt, steps - done)
        done = steps
        assert np.array_equal(sim.download()[0], field), (path, steps)


# ------------------------------------------------------------------ direct per-cell spectral evaluation (config 5's SFU-bound half)
@pytest.mark.parametrize("nl", [7, 64])
def test_spectral_direct_evaluation_matches_the_oracle(S, nl):
    """SBX_RGB_SPECTRAL: thin_film_interference of the fragment shaders (bubble_instanced.wgsl:190-218) per cell over
    nl wavelengths, against the oracle's f32 restatement of the same function.  Away from grazing incidence
    (cos theta_i > 0.15; at the limb the Airy fringes are needle-sharp and the shader's own f32 sine is ill-conditioned)
    RGBA8 within 1 LSB and float colour within 1e-4."""
    nt, nphi = 96, 512
    f = synthetic_field(nt, nphi, h0_nm=900.0, noise_nm=40.0)
    f *= np.linspace(0.05, 2.0, nt)[:, None]  # 45 nm ... 1.8 um over the sphere
    wl = (380.0 + 400.0 * np.arange(64) / 63.0) if nl == 64 else np.arange(400.0, 701.0, 50.0)
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    sim.upload(f)
    view = S.View(mode=S.RGB_SPECTRAL, format=S.OUT_RGBA8)
    if nl == 7:
        sim.shade(view)  # the handle's default spectrum is the reference's seven wavelengths
    else:
        sim.shade_spectral(view, wl)
    px = sim.download_rgb()[0]
    f32 = f.astype(np.float32).astype(np.float64)
    ref = O.shade_field_spectral(f32, nt, 0, 0.025, view.cam_pos, wl)
    th = (np.arange(nt) * math.pi / (nt - 1))[:, None]
    ph = (np.arange(nphi) * 2 * math.pi / nphi)[None, :]
    ci = np.vectorize(lambda t, p: O.view_cos(t, p, 0.025, view.cam_pos))(th, ph)
    ok = ci > 0.15
    d = np.abs(px.astype(int) - ref.astype(int))[..., :3].max(axis=-1)
    assert d[ok].max() <= 1, int(d[ok].max())
    assert (px[..., 3] == 255).all()
    assert np.count_nonzero(d[ok]) <= 0.02 * ok.sum()
    viewf = S.View(mode=S.RGB_SPECTRAL, format=S.OUT_RGB_F32)
    sim.shade(viewf)
    rgbf = sim.download_rgb(S.OUT_RGB_F32)[0]
    reff = O.shade_field_spectral(f32, nt, 0, 0.025, view.cam_pos, wl, as_float=True)
    assert float(np.abs(rgbf - reff)[ok].max()) <= 1e-4
    # the fused entry point with this mode = drainage steps, then the colour pass
    sim.step_fused(1e-6, view, 2)
    assert sim.download_rgb().shape == (1, nt, nphi, 4)
    with pytest.raises(S.SbxError):
        sim.set_spectrum(np.zeros(3))


# ------------------------------------------------------------------ one launch per frame on small grids (resident colour epilogue)
def test_steps_per_frame_loop_is_one_launch_with_the_frame(S):
    """step_fused(dt, view, n >= 2) on a grid that fits shared memory: n steps AND the colour frame of the final field in
    ONE launch (the reference's steps_per_frame loop + draw, gpu_drainage.rs:338-370); the frame is the one a separate
    colour pass of the same field produces, bit for bit, pole rows included."""
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    sim = make_sim(S, 128, 256, S.PREC_F32)  # config/default.json, offset storage
    sim.step_fused(1e-3, view, 10)            # builds the LUT, column table, ...
    l0 = sim.launch_count()
    sim.step_fused(1e-3, view, 10)
    assert sim.launch_count() - l0 == 1
    frame = sim.download_rgb()
    sim.shade(view)
    assert np.array_equal(frame, sim.download_rgb())
    ora = O.OracleDrainage(128, 256)
    ora.step(1e-3, 20)
    ref = O.shade_field(ora.data, 128, 0, 0.025, view.cam_pos, intensity=4.0, mode=1)
    assert int(np.abs(frame[0].astype(int) - ref.astype(int)).max()) <= 1
    # EXACT3 epilogue, f64 storage
    v3 = S.View(mode=S.RGB_EXACT3, format=S.OUT_RGBA8)
    s64 = make_sim(S, 64, 128, S.PREC_F64)
    s64.step_fused(1e-3, v3, 5)
    l0 = s64.launch_count()
    s64.step_fused(1e-3, v3, 5)
    assert s64.launch_count() - l0 == 1
    fr = s64.download_rgb()
    s64.shade(v3)
    assert np.array_equal(fr, s64.download_rgb())
    # the reference's GPU grid, WGSL semantics: steps_per_frame dispatches + frame
    g = S.GPUDrainageSimulator(500e-9, 128, 64)
    g.dispatch(10, view)
    l0 = g.launch_count()
    g.dispatch(10, view)
    assert g.launch_count() - l0 == 1
    fr = g.download_rgb()
    g.shade(view)
    assert np.array_equal(fr, g.download_rgb())


# ------------------------------------------------------------------ host-buffer step of a foam batch, pipelined over groups of bubbles
@pytest.mark.parametrize("offset_storage", [False, True])
def test_foam_step_host_pipelined_equals_separate_calls(S, monkeypatch, offset_storage):
    """sbx_drainage_step_host on a batch of independent bubbles (H2D / [convert] / step / [convert] / D2H pipelined over
    groups of bubbles) against upload + step_fused + download, plain f32 and offset storage, pole rows folded per bubble."""
    monkeypatch.setenv("SBX_HOST_CHUNK_BUBBLES", "2")  # 5 bubbles: groups of 2, 2, 1
    nt, nphi, B = 48, 256, 5
    h0 = np.linspace(300e-9, 700e-9, B)
    rad = np.linspace(0.0125, 0.1, B)
    kw = dict(batch=B, batch_h0_m=h0, batch_radius_m=rad, h_ref=1.0 if offset_storage else 0.0)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    rng = np.random.default_rng(7)
    cur = (h0[:, None, None] * (1.0 + 0.01 * rng.standard_normal((B, nt, nphi)))).astype(np.float32)
    sim, ref = make_sim(S, nt, nphi, S.PREC_F32, **kw), make_sim(S, nt, nphi, S.PREC_F32, **kw)
    out_h = np.empty((B, nt, nphi), np.float32)
    out_px = np.empty((B, nt, nphi, 4), np.uint8)
    for step in range(3):
        ref.upload(cur)
        ref.step_fused(1e-3, view, 1)
        want_h, want_px = ref.download(np.float32), ref.download_rgb()
        sim.step_host(1e-3, 1, view, cur, out_h, out_px)
        assert np.array_equal(out_h, want_h), step
        assert np.array_equal(out_px, want_px), step
        cur = want_h.copy()
    # and against the oracle, bubble by bubble
    for b in (0, B - 1):
        ora = O.OracleDrainage(nt, nphi, h0_m=float(h0[b]), radius_m=float(rad[b]))
        ora.data[:] = cur[b].astype(np.float64)
        ora.step(1e-3, 1)
        sim.step_host(1e-3, 1, view, cur, out_h, out_px)
        assert float(np.abs(out_h[b] / ora.data - 1).max()) <= 1e-5
        break


# ------------------------------------------------------------------ the reference's frame loop behind one synchronisation
def test_frame_stats_equal_the_separate_queries(S):
    """sbx_drainage_frame_stats = min_thickness, max_thickness, get_thickness, has_critical_region, current_time of
    RenderPipeline::update (pipeline.rs:1298-1312) in one call; the Python mirror serves its separate methods from it."""
    from soap_bubble_simulation_b200 import _lib as L
    import ctypes as C

    nt, nphi = 96, 1024
    sim = make_sim(S, nt, nphi, S.PREC_F32, h_ref=0.0)
    f = synthetic_field(nt, nphi, dtype=np.float32)
    f[40, 7] = 4e-9  # one sub-critical cell
    sim.upload(f)
    ora = O.OracleDrainage(nt, nphi)
    ora.data[:] = f.astype(np.float64)
    dt = 0.25 * dt_max(nt, nphi)
    view = S.View()
    for frame in range(4):  # the tracked min/max of the step kernel switches itself on after two frames
        sim.step_fused(dt, view, 1)
        ora.step(dt, 1)
        st = sim.frame_stats(math.pi / 2, 0.3)
        # the separate C entry points give the same numbers
        mn, mx, th, crit = C.c_double(), C.c_double(), C.c_double(), C.c_int32()
        L.check(L.lib().sbx_drainage_minmax(sim._h, C.byref(mn), C.byref(mx)))
        L.check(L.lib().sbx_drainage_get_thickness(sim._h, 0, math.pi / 2, 0.3, C.byref(th)))
        L.check(L.lib().sbx_drainage_has_critical(sim._h, C.byref(crit)))
        assert (st["min_thickness"], st["max_thickness"], st["thickness"], st["has_critical"]) == (mn.value, mx.value, th.value, bool(crit.value))
        assert st["time"] == sim.current_time() and st["has_critical"]
        assert abs(st["thickness"] / ora.get_thickness(math.pi / 2, 0.3) - 1) <= 1e-5
        assert abs(st["max_thickness"] / ora.max_thickness() - 1) <= 1e-5
        # the mirror's reference-shaped methods: served from one fetch per frame
        fld = sim.thickness_field()
        assert (fld.min_thickness(), fld.max_thickness()) == (mn.value, mx.value)
        assert sim.get_thickness(math.pi / 2, 0.3) == th.value and sim.has_critical_region()
    assert sim.get_thickness(1.0, 2.0) == pytest.approx(ora.get_thickness(1.0, 2.0), rel=1e-5)  # other arguments: refetched
