"""The reference's own unit tests for the hot path, restated against the CPU oracle.

Each test cites the reference test it restates (file:line under /root/reference).  These are the
only pins the reference offers for this path: property tests plus the
closed-form Fresnel value.  Closed-form known answers derived from the cited formulas are added
at the end and frozen in tests/golden/kat_interference.json.
"""
import json
import math
import os

import numpy as np
import pytest

import oracle_lib as O

EPS = 1e-6
F64_EPS = 2.220446049250313e-16
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- drainage.rs:616-720
def test_thickness_field_creation():  # drainage.rs:621-626
    s = O.OracleDrainage(64, 128, 500e-9)
    assert s.data.shape == (64, 128)
    assert abs(s.data[0, 0] - 500e-9) < F64_EPS


def test_thickness_field_get_set():  # drainage.rs:629-633
    s = O.OracleDrainage(64, 128, 500e-9)
    s.data[10, 20] = 400e-9
    assert abs(s.data[10, 20] - 400e-9) < F64_EPS


def test_drainage_simulator_creation():  # drainage.rs:636-642
    s = O.OracleDrainage()
    assert s.num_theta == 128 and s.num_phi == 256
    assert abs(s.current_time()) < F64_EPS
    assert not s.has_critical_region()


def test_drainage_step_advances_time():  # drainage.rs:645-653
    s = O.OracleDrainage()
    t0 = s.current_time()
    s.step(0.001)
    assert abs(s.current_time() - t0 - 0.001) < F64_EPS


def test_get_thickness_interpolation_uniform():  # drainage.rs:656-665
    s = O.OracleDrainage()
    assert abs(s.get_thickness(0.5, 1.0) - s.get_thickness(1.5, 3.0)) < 1e-15


def test_drainage_causes_thinning_at_top():  # drainage.rs:668-688
    s = O.OracleDrainage()
    before = s.get_thickness(0.1, 0.0)
    s.step(O.DEFAULTS["dt"], 100)
    assert s.get_thickness(0.1, 0.0) < before


def test_phi_periodicity():  # drainage.rs:691-702
    s = O.OracleDrainage()
    s.step(1e-3, 5)
    assert abs(s.get_thickness(1.0, 0.0) - s.get_thickness(1.0, 2.0 * math.pi)) < 1e-12


def test_reset():  # drainage.rs:705-719
    s = O.OracleDrainage()
    s.step(O.DEFAULTS["dt"], 10)
    s.reset(600e-9)
    assert abs(s.current_time()) < F64_EPS
    assert abs(s.get_thickness(0.5, 0.5) - 600e-9) < 1e-15
    assert s.active_tfe_count() == 0 and s.next_spawn() == 0.5


# ---------------------------------------------------------------- interference.rs:612-792
def test_rgb_color_to_u8():  # interference.rs:627-633
    assert O.rgb_to_u8((1.0, 0.5, 0.0)) == (255, 128, 0)


def test_calculator_creation():  # interference.rs:636-647
    ic = O.OracleInterference(1.33)
    assert abs(ic.c.n_film - 1.33) < EPS and abs(ic.c.n_medium - 1.0) < EPS
    assert tuple(ic.c.wavelengths_nm) == (650.0, 532.0, 450.0)


def test_transmission_angle_normal_incidence():  # interference.rs:649-654
    assert abs(O.OracleInterference(1.33).transmission_angle_cos(1.0) - 1.0) < EPS


def test_transmission_angle_snells_law():  # interference.rs:657-669
    ic = O.OracleInterference(1.5)
    ci = math.sqrt(2.0) / 2.0
    ct = ic.transmission_angle_cos(ci)
    ratio = math.sqrt(1 - ci * ci) / math.sqrt(1 - ct * ct)
    assert abs(ratio - 1.5) < 1e-4


def test_fresnel_normal_incidence():  # interference.rs:672-679 -- the one closed-form KAT
    rs, rp, avg = O.OracleInterference(1.33).fresnel(1.0)
    expected = ((1.0 - 1.33) / (1.0 + 1.33)) ** 2
    assert abs(avg - expected) < 1e-4
    assert abs(avg - expected) < 1e-15  # and in fact to rounding


def test_zero_thickness_gives_zero_intensity():  # interference.rs:682-686
    assert abs(O.OracleInterference().reflected_intensity(0.0, 1.0, 550.0)) < EPS


def test_interference_produces_colors():  # interference.rs:689-699
    ic = O.OracleInterference()
    for d in (100.0, 200.0, 300.0, 400.0, 500.0):
        assert max(ic.color(d, 1.0)) > 0.01


def test_color_varies_with_thickness():  # interference.rs:702-716
    ic = O.OracleInterference()
    a, b = ic.color(100.0, 1.0), ic.color(300.0, 1.0)
    assert sum(abs(x - y) for x, y in zip(a, b)) > 0.01


def test_color_varies_with_angle():  # interference.rs:719-731
    ic = O.OracleInterference()
    a, b = ic.color(300.0, 1.0), ic.color(300.0, 0.5)
    assert sum(abs(x - y) for x, y in zip(a, b)) > 0.001


def test_lookup_table_generation():  # interference.rs:734-747
    t = O.OracleInterference().lut_1d(0.0, 1000.0, 100, 1.0)
    assert t.shape == (100, 3) and (t >= 0).all() and (t <= 1).all()


def test_2d_lookup_table_generation():  # interference.rs:750-755
    t = O.OracleInterference().lut_2d(0.0, 500.0, 10, 5)
    assert t.shape[:2] == (10, 5)


def test_linear_vs_gamma_color():  # interference.rs:758-769
    ic = O.OracleInterference()
    lin, gam = ic.color_linear(300.0, 1.0), ic.color(300.0, 1.0)
    if 0.01 < lin[0] < 0.99:
        assert gam[0] >= lin[0]


def test_custom_wavelengths():  # interference.rs:772-780
    c = O.OracleInterference(1.33, (700.0, 550.0, 400.0)).color(300.0, 1.0)
    assert all(0.0 <= x <= 1.0 for x in c)


# ---------------------------------------------------------------- interference_lut.rs:165-205
def test_lut_generation_produces_valid_data():  # interference_lut.rs:170-174
    lut = O.generate_interference_lut(1.33, 1.0)
    assert lut.size == 256 * 64 * 4
    assert (lut[..., 3] == 255).all()


def test_thin_film_produces_colors():  # interference_lut.rs:177-185
    rgb = O.thin_film_interference(500.0, 0.9, 1.33, 1.0)
    assert all(0.0 <= x <= 1.0 for x in rgb) and sum(rgb) > 0.01


def test_lut_snells_law():  # interference_lut.rs:188-196
    assert abs(O.lib().orc_snells_law(1.0, 1.33) - 1.0) < 0.01
    assert 0.0 < O.lib().orc_snells_law(0.1, 1.33) < 1.0


def test_lut_fresnel_at_normal_incidence():  # interference_lut.rs:199-204
    assert 0.01 < O.lib().orc_fresnel_unpolarized(1.0, 1.0, 1.0, 1.33) < 0.05


# ---------------------------------------------------------------- closed-form known answers
def _airy(R, cycles):
    c = math.cos(2.0 * math.pi * cycles)
    return 2 * R * (1 - c) / (1 + R * R - 2 * R * c)


def test_kat_linear_rgb_closed_form():
    """Independent re-derivation (pure Python, straight from the formulas at
    interference.rs:197-211,226-254,266-298,337-352) of the linear RGB for a few points."""
    ic = O.OracleInterference(1.33)
    n = 1.33
    for d, ci in [(500.0, 1.0), (300.0, 0.5), (250.0, math.sqrt(0.5)), (400.0, 1.0), (1234.5, 0.3)]:
        st = math.sqrt(1 - ci * ci) / n
        ct = math.sqrt(1 - st * st)
        rs = (ci - n * ct) / (ci + n * ct)
        rp = (n * ci - ct) / (n * ci + ct)
        R = (rs * rs + rp * rp) / 2
        exp = [min(max(_airy(R, (2 * n * d * ct + wl / 2) / wl), 0.0), 1.0) for wl in (650.0, 532.0, 450.0)]
        got = ic.color_linear(d, ci)
        assert np.allclose(got, exp, rtol=0, atol=1e-12)


def test_kat_frozen_vectors():
    """Frozen known answers -- regenerate with tests/golden/make_kat.py."""
    with open(os.path.join(GOLDEN, "kat_interference.json")) as f:
        kat = json.load(f)
    ic = O.OracleInterference(kat["n_film"])
    for row in kat["linear"]:
        assert np.allclose(ic.color_linear(row["d_nm"], row["cos"]), row["rgb"], rtol=0, atol=1e-9)
    for row in kat["gamma"]:
        assert np.allclose(ic.color(row["d_nm"], row["cos"]), row["rgb"], rtol=0, atol=1e-9)
    assert abs(ic.fresnel(1.0)[2] - kat["fresnel_normal"]) < 1e-15


def test_kat_drainage_single_step_uniform():
    """Uniform 500 nm, default config, one step: the discrete Laplacian of a uniform field is exactly 0,
    so interior dh = -dt*rho*g/(3*eta)*h^3*sin(theta_i); poles = mean of ring 1 (drainage.rs:347-348,480-503)."""
    s = O.OracleDrainage()
    h0 = 500.0 * 1e-9  # film_thickness_meters(), config.rs:369-371 (== 5.000000000000001e-07, not 500e-9)
    s.step(1e-3)
    c = 1000.0 * 9.81 / (3.0 * 0.001)
    i = np.arange(1, 127)
    theta = i * (math.pi / 127.0)
    # math.sin is glibc's sin (what the oracle and Rust call); numpy's SIMD sin may differ by an ulp
    expect = np.array([h0 + 1e-3 * (-c * (h0 * h0 * h0) * math.sin(t)) for t in theta])
    assert np.array_equal(s.data[1:-1, 0], expect)
    assert np.array_equal(s.data[1:-1, 17], expect)
    assert s.data[0, 0] == s.data[0, 255]
    acc = 0.0
    for v in s.data[1]:  # left-to-right f64 accumulation, drainage.rs:484-488
        acc += float(v)
    assert s.data[0, 0] == acc / 256.0
    assert abs((s.data[64, 0] - h0) + 4.0875e-16) < 1e-19


def test_kat_drainage_discrete_mode_closed_form():
    """A known answer that does not come from running any restatement (tests/closed_form.py): on h = a + eps * theta *
    cos(m phi) the reference's finite differences (drainage.rs:352-367) have closed forms, so one Euler step
    (drainage.rs:346-378) is

        h' = h + dt * ( -rho g / (3 eta) * h^3 sin(theta)
                        + D / R^2 * eps cos(m phi) * (cot(theta) + theta * lam_m / sin^2(theta)) ),   lam_m = (2 cos(m dphi) - 2) / dphi^2

    and each pole is the mean of the updated ring next to it (drainage.rs:480-503).  Checked on the increments."""
    from closed_form import cpu_mode

    nt, nphi, m, dt = 33, 48, 3, 0.05
    p = O.DEFAULTS
    h, inc, top, bot = cpu_mode(nt, nphi, m, 500e-9, 40e-9, dt, p["density"], p["gravity"], p["viscosity"], p["diffusion"],
                                p["radius_m"])
    s = O.OracleDrainage(nt, nphi)
    s.data[:] = h
    s.step(dt)
    got = s.data[1:-1] - h[1:-1]
    # increments: diffusion up to 9e-12 m, gravity up to 3e-14 m, on a 5e-7 m field whose ulp is 1e-22 m -- the
    # agreement (5e-23 m measured) is at the rounding of the stored field, 2e-9 of the smaller of the two terms
    assert np.abs(inc).max() > 5e-12
    assert np.abs(got - inc).max() <= 2e-22
    assert abs(s.data[0, 0] - top) <= 1e-21 and np.all(s.data[0] == s.data[0, 0])
    assert abs(s.data[-1, 5] - bot) <= 1e-21 and np.all(s.data[-1] == s.data[-1, 0])


def test_kat_wgsl_discrete_mode_closed_form():
    """The same kind of known answer for drainage.wgsl::drainage_step (drainage.wgsl:72-200), thickness and surfactant
    concentration both of the form a + eps * theta * cos(m phi): Laplacians as above, the Marangoni term from the exact
    central differences d/dtheta = eps cos(m phi), d/dphi = -eps theta sin(m phi) sin(m dphi) / dphi (:157-174), the
    concentration update with its thickness coupling (:180-197), poles copied (:84-88).  The shader is f32, so the
    closed form (f64) is met to a few ulps of the stored f32 values: 1e-13 m on increments up to 1.8e-8 m, 1e-7 on
    concentration increments up to 0.023 (the 0.001 coupling term, 6e-8 at most here, is below that and not checked)."""
    from closed_form import wgsl_mode

    H, W = 33, 48
    p = O.wgsl_params(grid_width=W, grid_height=H, marangoni_enabled=1, dt=1.0, diffusion_coeff=1e-7,
                      surfactant_diffusion=1e-7, marangoni_coeff=2e3)
    h, c, inc_h, inc_c, grav, mar = wgsl_mode(H, W, 3, 500e-9, 40e-9, 0.4, 0.05, p)
    ho, co = O.wgsl_step(p, h, c)
    ho, co = ho.reshape(H, W).astype(np.float64), co.reshape(H, W).astype(np.float64)
    assert np.array_equal(ho[[0, -1]], h[[0, -1]]) and np.array_equal(co[[0, -1]], c[[0, -1]])  # poles: copied
    got_h, got_c = ho[1:-1] - h[1:-1].astype(np.float64), co[1:-1] - c[1:-1].astype(np.float64)
    # tolerances: ~4 ulp of h (5.7e-14 m) and of c (6e-8).  Leaving out the Marangoni term (4.5e-9 m), the diffusion term
    # (1.8e-8 m) or the gravity term (6e-13 m) would each miss by more than that.
    tol_h, tol_c = 2.5e-13, 2.5e-7
    assert np.abs(mar).max() > 1000 * tol_h and np.abs(inc_h - mar).max() > 1000 * tol_h and np.abs(grav).max() > 2 * tol_h
    assert np.abs(inc_c).max() > 1000 * tol_c
    assert np.abs(got_h - inc_h).max() <= tol_h
    assert np.abs(got_c - inc_c).max() <= tol_c


def test_tfe_spawn_sequence():
    """First TFE front appears once 0.5 s have elapsed (drainage.rs:285,397-403) with the
    time-hashed parameters of drainage.rs:466-468 evaluated at current_time before the increment."""
    s = O.OracleDrainage(32, 64)
    n = 0
    while s.active_tfe_count() == 0:
        s.step(1e-3)
        n += 1
        assert n < 600
    t = (n - 1) * 1e-3  # accumulated the same way the simulator does
    tt = 0.0
    for _ in range(n - 1):
        tt += 1e-3
    (phi_c, extent, width, vel), = s.fronts()
    assert phi_c == abs(math.sin(tt * 7.89 + 1.23)) * 2.0 * math.pi
    assert width == 0.3 + abs(math.sin(tt * 3.21)) * 0.4
    assert vel == 0.05 + abs(math.cos(tt * 5.67)) * 0.05
    assert extent == min(0.0 + vel * 1e-3, math.pi / 2)
    assert abs(t - tt) < 1e-12
    assert s.next_spawn() == 2.0 * (0.5 + abs(math.sin(tt * 12.345)))


def test_openmp_rows_bit_identical():
    a, b = O.OracleDrainage(64, 128), O.OracleDrainage(64, 128, threads=4)
    rng = np.random.default_rng(1)
    f = 500e-9 * (1 + 0.1 * rng.standard_normal((64, 128)))
    a.data[:] = f
    b.data[:] = f
    a.step(1e-3, 20)
    b.step(1e-3, 20)
    assert np.array_equal(a.data, b.data)
