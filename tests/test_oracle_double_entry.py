"""Double-entry check of the drainage oracle: oracle/sbx_oracle.c and tests/numpy_restatement.py were transcribed
separately from src/physics/drainage.rs and must agree BIT FOR BIT.  The reference holds no golden vectors for drainage
and cannot be built here, so this is the strongest pin available for that half: a transcription
slip would have to be made identically twice."""
import math

import numpy as np
import pytest

import oracle_lib as O
from numpy_restatement import NumpyDrainage
from synth import synthetic_field


def _pair(nt, nphi, field=None):
    a, b = O.OracleDrainage(nt, nphi), NumpyDrainage(nt, nphi)
    if field is not None:
        a.data[:] = field
        b.h[:] = field
    return a, b


@pytest.mark.parametrize("nt,nphi,steps,dt", [(128, 256, 1000, 1e-3), (33, 50, 1500, 1e-3), (16, 20, 6000, 2e-3)])
def test_default_config_runs_with_fronts(nt, nphi, steps, dt):
    """uniform 500 nm start (config/default.json): fronts spawn from t = 0.5 s, several are retired in the long run"""
    a, b = _pair(nt, nphi)
    for s in range(steps):
        a.step(dt, 1)
        b.step(dt)
        if s % 250 == 249 or s == steps - 1:
            assert np.array_equal(a.data, b.h), s
            assert a.fronts() == [(f.phi_center, f.extent, f.width, f.velocity) for f in b.fronts], s
    assert a.current_time() == b.time
    if steps * dt > 0.6:
        assert len(b.fronts) >= 1


def test_noisy_field_critical_cells_and_queries():
    nt, nphi = 48, 96
    f = synthetic_field(nt, nphi)
    f[10:13, 5:9] = 5e-9   # below the critical thickness: skipped
    f[20, 30] = 0.0
    f[30, 95] = 3e-6       # a spike at the periodic seam
    a, b = _pair(nt, nphi, f)
    for s in range(400):
        a.step(1e-3, 1)
        b.step(1e-3)
    assert np.array_equal(a.data, b.h)
    rng = np.random.default_rng(3)
    for th, ph in list(zip(rng.uniform(-0.5, 3.6, 50), rng.uniform(-7, 14, 50))) + [(0, 0), (math.pi, 0), (math.pi / 2, 2 * math.pi)]:
        assert a.get_thickness(th, ph) == b.get_thickness(th, ph)


def test_unstable_run_stays_identical_even_through_overflow():
    """dt far above the stability bound: values blow up to inf / NaN and the max(.., 0) clamp; both must follow the
    same IEEE path (NaN handling of f64::max included)."""
    nt, nphi = 24, 64
    a, b = _pair(nt, nphi, synthetic_field(nt, nphi))
    with np.errstate(all="ignore"):
        for s in range(60):
            a.step(0.5, 1)
            b.step(0.5)
    assert np.array_equal(a.data, b.h, equal_nan=True)


def test_interference_calculator_double_entry():
    """The f64 calculator: oracle (C) vs the separate Python transcription, bit for bit, over thickness x angle incl.
    grazing incidence, zero / negative thickness and out-of-range cosines."""
    from numpy_restatement import PyInterference

    for n_film, wl in ((1.33, None), (1.45, (700.0, 546.1, 435.8))):
        ora = O.OracleInterference(n_film, wl)
        py = PyInterference(n_film, wl or (650.0, 532.0, 450.0))
        rng = np.random.default_rng(11)
        ds = list(rng.uniform(0.0, 2000.0, 300)) + [0.0, -5.0, 1e-3, 500.0, 1000.0, 1e5]
        cs = list(rng.uniform(0.0, 1.0, 300)) + [0.0, 1.0, 1e-9, -0.3, 1.7, 0.5]
        for d, c in zip(ds, cs):
            assert ora.transmission_angle_cos(c) == py.transmission_angle_cos(c)
            assert tuple(ora.fresnel(c)) == py.fresnel(c)
            for w in py.wl:
                assert ora.reflected_intensity(d, c, w) == py.reflected_intensity(d, c, w), (d, c, w)
            assert tuple(ora.color_linear(d, c)) == py.color_linear(d, c)
            assert tuple(ora.color(d, c)) == py.color(d, c)


@pytest.mark.parametrize("marangoni", [0, 1])
@pytest.mark.parametrize("H,W", [(64, 128), (33, 50)])
def test_wgsl_step_double_entry(H, W, marangoni):
    """drainage.wgsl::drainage_step: the oracle's f32 restatement vs the separate numpy float32 transcription over 20
    dispatches of a noisy thickness + concentration pair (sub-critical cells included)."""
    from numpy_restatement import wgsl_drainage_step

    rng = np.random.default_rng(H * 100 + W + marangoni)
    h = (synthetic_field(H, W) * (1.0 + 0.05 * rng.uniform(-1, 1, (H, W)))).astype(np.float32)
    c = (0.5 + 0.3 * rng.uniform(-1, 1, (H, W))).astype(np.float32)
    h[5:8, 3:6] = 20e-9  # below the GPU path's 30 nm critical thickness
    p = O.wgsl_params(grid_width=W, grid_height=H, marangoni_enabled=marangoni, dt=1e-3)
    ha, ca, hb, cb = h, c, h, c
    for _ in range(20):
        ha, ca = O.wgsl_step(p, ha, ca)
        hb, cb = wgsl_drainage_step(p, hb, cb)
    assert np.array_equal(ha.reshape(H, W), hb) and np.array_equal(ca.reshape(H, W), cb)


@pytest.mark.parametrize("n,intensity", [(1.33, 1.0), (1.40, 4.0)])
def test_interference_lut_double_entry(n, intensity):
    """generate_interference_lut: every one of the 65,536 bytes, oracle vs the separate numpy float32 transcription."""
    from numpy_restatement import interference_lut

    a = O.generate_interference_lut(n, intensity).reshape(64, 256, 4)
    b = interference_lut(n, intensity)
    assert np.array_equal(a, b), int(np.count_nonzero(a != b))


@pytest.mark.parametrize("H,W", [(64, 128), (37, 50)])
def test_caustic_map_double_entry(H, W):
    """caustics_compute.wgsl (row f2): oracle vs the separate numpy float32 transcription.  The shader's `length()` and
    `pow()` have no exact host definition, so this one is compared to 2 ulp instead of bit for bit."""
    from numpy_restatement import caustic_map

    rng = np.random.default_rng(H + W)
    # thickness in the unit the consumer sees (the reference feeds metres; variations of 1e-8 would vanish in f32
    # against focal_length, so scale the noise up to exercise every branch)
    t = (1.0 + 0.3 * rng.uniform(-1, 1, (H, W))).astype(np.float32)
    p = O.caustic_params(grid_width=W, grid_height=H)
    a = O.caustic_map(p, t).reshape(H, W)
    b = caustic_map(p, t)
    assert (a > 0).all() and len(np.unique(a)) > 100
    assert np.abs(a - b).max() <= 2 * np.spacing(np.abs(a).max())
    assert np.count_nonzero(a != b) <= 0.01 * a.size


def test_branched_flow_sampling_double_entry():
    """sample_thickness_at_uv + thickness_gradient_uv (row f2): oracle vs the separate numpy transcription, exact."""
    from numpy_restatement import sample_thickness_at_uv

    rng = np.random.default_rng(5)
    for H, W in ((64, 128), (100, 37)):
        t = synthetic_field(H, W).astype(np.float32)
        uv = np.concatenate([rng.uniform(-0.2, 1.2, (4000, 2)), [[0, 0], [1, 1], [0.5, 0.5], [0.995, 0.005]]]).astype(np.float32)
        ha, ga = O.sample_thickness_uv(t, uv)
        hb, gb = sample_thickness_at_uv(t, uv)
        assert np.array_equal(ha, hb) and np.array_equal(ga, gb)
