"""The product's per-cell drainage arithmetic, compiled for the HOST from the very text the kernels include
(cell_update / cell_update_strict in csrc/sbx_device.cuh), against the oracle -- a CPU-side guard for the device math:
the strict form must reproduce the oracle's interior update bit for bit, the fast form to rounding.  Only the
`__device__` qualifiers and the IEEE intrinsics (__dmul_rn ... = plain IEEE ops under -ffp-contract=off) are shimmed.
The GPU tests remain the parity tests proper; this one needs no GPU."""
import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from synth import synthetic_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cmath>
#include <cstddef>
#define __device__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
using std::fma;
namespace sbx {
%s
}
// one interior pass over a [nt][nphi] field with the coefficient tables the library builds on the host
// (sbx_api.cu: coefX = {sin, cos/sin_safe, sin_safe^2}; coefS = {G, B, E}; A = D/(R^2 dth^2))
extern "C" void pass_strict(const double *in, double *out, int nt, int nphi, double rho, double g, double eta,
                            double D, double R, double crit, double dt)
{
    const double PI = 3.14159265358979323846, dth = PI / (nt - 1), dph = 2.0 * PI / nphi;
    const double coef = rho * g / (3.0 * eta);
    for (int i = 1; i < nt - 1; ++i) {
        const double th = i * dth, s = std::sin(th), c = std::cos(th);
        const double ss = std::fabs(s) < 1e-10 ? std::copysign(1e-10, s) : s;
        const sbx::StrictRow r{s, c / ss, ss * ss};
        for (int j = 0; j < nphi; ++j) {
            const size_t k = (size_t)i * nphi + j;
            out[k] = sbx::cell_update_strict(in[k], in[k - nphi], in[k + nphi], in[(size_t)i * nphi + (j + nphi - 1) %% nphi],
                                             in[(size_t)i * nphi + (j + 1) %% nphi], r, coef, dth * dth, 2.0 * dth, dph * dph,
                                             R * R, D, dt, crit);
        }
    }
}
template <typename T>
static void pass_fast_t(const T *in, T *out, int nt, int nphi, double rho, double g, double eta, double D, double R,
                        double crit, double dt, double href)
{
    const double PI = 3.14159265358979323846, dth = PI / (nt - 1), dph = 2.0 * PI / nphi;
    const double A = D / (R * R * dth * dth);
    for (int i = 1; i < nt - 1; ++i) {
        const double th = i * dth, s = std::sin(th), c = std::cos(th);
        const double ss = std::fabs(s) < 1e-10 ? std::copysign(1e-10, s) : s;
        const double G = rho * g / (3.0 * eta) * s, B = D * (c / ss) / (R * R * 2.0 * dth), E = D / (R * R * dph * dph * ss * ss);
        for (int j = 0; j < nphi; ++j) {
            const size_t k = (size_t)i * nphi + j;
            out[k] = sbx::cell_update<T>(in[k], in[k - nphi], in[k + nphi], in[(size_t)i * nphi + (j + nphi - 1) %% nphi],
                                         in[(size_t)i * nphi + (j + 1) %% nphi], (T)G, (T)A, (T)B, (T)E, (T)1, (T)href, (T)dt, (T)crit);
        }
    }
}
extern "C" void pass_fast_f64(const double *in, double *out, int nt, int nphi, double rho, double g, double eta, double D,
                              double R, double crit, double dt, double href)
{ pass_fast_t<double>(in, out, nt, nphi, rho, g, eta, D, R, crit, dt, href); }
extern "C" void pass_fast_f32(const float *in, float *out, int nt, int nphi, double rho, double g, double eta, double D,
                              double R, double crit, double dt, double href)
{ pass_fast_t<float>(in, out, nt, nphi, rho, g, eta, D, R, crit, dt, href); }
"""


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    with open(os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc", "sbx_device.cuh")) as f:
        src = f.read()
    m = re.search(r"template <typename T>\n__device__ __forceinline__ T cell_update\(.*?\n}\n.*?"
                  r"__device__ __forceinline__ double cell_update_strict\(.*?\n}\n", src, flags=re.S)
    assert m, "cell_update / cell_update_strict not found in sbx_device.cuh"
    d = tmp_path_factory.mktemp("hostmath")
    cpp = d / "hostmath.cpp"
    cpp.write_text(HARNESS % m.group(0))
    so = d / "libhostmath.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           str(cpp), "-o", str(so)])
    return ctypes.CDLL(str(so))


PHYS = (1000.0, 9.81, 0.001, 1e-9, 0.05 / 2.0, 10.0 * 1e-9)  # rho, g, eta, D, R, crit (config/default.json)


def _oracle_step(f, dt):
    nt, nphi = f.shape
    ora = O.OracleDrainage(nt, nphi)
    ora.set_marginal_regeneration(False)
    ora.data[:] = f
    ora.step(dt, 1)
    return ora.data.copy()


@pytest.mark.parametrize("nt,nphi", [(64, 128), (33, 50)])
def test_strict_form_is_bit_identical_to_the_oracle(hostlib, nt, nphi):
    f = synthetic_field(nt, nphi)
    f[10:12, 5:9] = 5e-9  # sub-critical cells
    out = f.copy()
    c = ctypes.c_double
    hostlib.pass_strict(f.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), nt, nphi,
                        *(c(v) for v in PHYS), c(1e-3))
    ref = _oracle_step(f, 1e-3)
    assert np.array_equal(out[1:-1], ref[1:-1])


@pytest.mark.parametrize("nt,nphi", [(64, 128), (33, 50)])
def test_fast_form_matches_the_oracle_to_rounding(hostlib, nt, nphi):
    f = synthetic_field(nt, nphi)
    f[10:12, 5:9] = 5e-9
    ref = _oracle_step(f, 1e-3)
    c = ctypes.c_double
    out = f.copy()
    hostlib.pass_fast_f64(f.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), nt, nphi,
                          *(c(v) for v in PHYS), c(1e-3), c(0.0))
    assert np.abs(out[1:-1] / ref[1:-1] - 1).max() <= 1e-13
    # increments, not just values: the fast form must resolve the per-step change itself
    inc, inc_ref = out[1:-1] - f[1:-1], ref[1:-1] - f[1:-1]
    assert np.abs(inc - inc_ref).max() <= 1e-9 * np.abs(inc_ref).max()
    # f32, plain storage (h_ref = 0) and offset storage (h_ref = 500 nm keeps the sub-ulp increments)
    f32 = f.astype(np.float32)
    o32 = f32.copy()
    hostlib.pass_fast_f32(f32.ctypes.data_as(ctypes.c_void_p), o32.ctypes.data_as(ctypes.c_void_p), nt, nphi,
                          *(c(v) for v in PHYS), c(1e-3), c(0.0))
    ref32 = _oracle_step(f32.astype(np.float64), 1e-3)
    assert np.abs(o32[1:-1].astype(np.float64) / ref32[1:-1] - 1).max() <= 1.2e-7
    href = 500.0 * 1e-9
    u = np.full((nt, nphi), href)  # the default configuration: uniform film, increments 8e-10 relative per step
    d32 = (u - href).astype(np.float32)
    od = d32.copy()
    hostlib.pass_fast_f32(d32.ctypes.data_as(ctypes.c_void_p), od.ctypes.data_as(ctypes.c_void_p), nt, nphi,
                          *(c(v) for v in PHYS), c(1e-3), c(href))
    refu = _oracle_step(u, 1e-3)
    inc, inc_ref = od[1:-1].astype(np.float64), refu[1:-1] - href
    assert np.abs(inc_ref).max() > 0 and np.abs(inc - inc_ref).max() <= 1e-3 * np.abs(inc_ref).max()


def test_frame_kernel_noise_matches_the_shader_restatement(tmp_path):
    """k_bubble_frame's simplex / FBM film thickness (csrc/sbx_frame.cu, from bubble.wgsl:104-229), host-compiled from
    the kernel's own text, against the numpy restatement the golden-screenshot tests use (tests/golden_render.py)."""
    import golden_render as G

    with open(os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc", "sbx_frame.cu")) as f:
        src = f.read()
    body = src[src.index("struct F3"):src.index("// Thickness of the drainage field")]
    cpp = tmp_path / "noise.cpp"
    cpp.write_text("#include <cmath>\n#define __device__\n#define __forceinline__ inline\n"
                   "struct FrameArgs { float base_nm, film_time, swirl, drainage_speed, pattern_scale; };\n" + body +
                   'extern "C" void ft(const float *n, int cnt, float base, float t, float *out) {\n'
                   "  FrameArgs a{base, t, 1.0f, 0.5f, 1.0f};\n"
                   "  for (int k = 0; k < cnt; ++k) out[k] = film_thickness_procedural(a, F3{n[3 * k], n[3 * k + 1], n[3 * k + 2]});\n}\n")
    so = tmp_path / "libnoise.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", str(cpp), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(1)
    n = rng.normal(size=(100000, 3))
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    for t in (0.0, 3.7):
        out = np.empty(len(n), np.float32)
        lib.ft(n.ctypes.data_as(ctypes.c_void_p), len(n), ctypes.c_float(500.0), ctypes.c_float(t), out.ctypes.data_as(ctypes.c_void_p))
        ref = G.film_thickness(n, 500.0, film_time=t)
        assert np.abs(out - ref).max() <= 1e-3  # nm; f32 rounding of a ~500 nm value is 3e-5


COLOUR_SHIMS = r"""
#include <cmath>
#define __device__
#define __forceinline__ inline
#include <cstring>
#include <cstdint>
struct RgbArgs { float n_film, inv_n2, two_n, r0, omr0; double inv_n2_d; double two_n_over_wl[3]; float two_pi_over_wl_f[3];
                 float inv_n2_hi, inv_n2_lo, w_hi[3], w_lo[3]; };
static inline float __fmul_rn(float a, float b) { return a * b; }  // plain IEEE ops under -ffp-contract=off
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float rsqrt_approx(float x) { return (float)(1.0 / std::sqrt((double)x)) * (1.0f + 1.2e-7f); }  // ~1 ulp off, like the SFU
static inline float clamp01(float x) { return std::fmin(std::fmax(x, 0.f), 1.f); }
static inline float rcp_approx(float x) { return 1.0f / x; }       // the SFU approximations, exact here: this test
static inline float sqrt_approx(float x) { return std::sqrt(x); }  // checks the algebra, the GPU tests the hardware
static inline float ex2_approx(float x) { return std::exp2(x); }
static inline float lg2_approx(float x) { return std::log2(x); }
static inline float __saturatef(float x) { return clamp01(x); }
#define __cosf(x) std::cos((float)(x))  /* glibc declares a __cosf of its own */
using std::fma; using std::fmaf; using std::fabs; using std::rint; using std::fmaxf; using std::fminf;
"""


def test_colour_stage_algebra_matches_the_oracle(tmp_path):
    """rgb_exact3 / rgb_two_beam / rgb_schlick_wall (csrc/sbx_device.cuh) host-compiled from the kernel text with exact
    stand-ins for the SFU approximations: the algebra (f64 fringe phase folded to the nearest zero, 1 - R from the
    transmittances, Airy, gamma) must agree with InterferenceCalculator (interference.rs:380-426) within the
    north-star tolerance of 1e-4 -- measured here it is ~1e-6 -- over thickness x angle incl. grazing incidence."""
    with open(os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc", "sbx_device.cuh")) as f:
        src = f.read()
    body = src[src.index("// sin(pi x) for |x| <= 0.5"):src.index("// bubble.wgsl:425-439 through the texture unit")]
    cpp = tmp_path / "colour.cpp"
    cpp.write_text(COLOUR_SHIMS + body + r"""
extern "C" void eval(int mode, double n, const double *d, const float *ci, int cnt, float *out) {
  RgbArgs a{};
  a.n_film = (float)n; a.inv_n2 = (float)(1.0 / (n * n)); a.inv_n2_d = 1.0 / (n * n); a.two_n = (float)(2.0 * n);
  a.r0 = (float)(((n - 1.0) / (n + 1.0)) * ((n - 1.0) / (n + 1.0))); a.omr0 = 1.0f - a.r0;
  const double wl[3] = {650.0, 532.0, 450.0};
  for (int k = 0; k < 3; ++k) { a.two_n_over_wl[k] = 2.0 * n / wl[k]; a.two_pi_over_wl_f[k] = (float)(2.0 * 3.14159265358979323846 / wl[k]); }
  a.inv_n2_hi = (float)a.inv_n2_d; a.inv_n2_lo = (float)(a.inv_n2_d - (double)a.inv_n2_hi);
  for (int k = 0; k < 3; ++k) { a.w_hi[k] = (float)a.two_n_over_wl[k]; a.w_lo[k] = (float)(a.two_n_over_wl[k] - (double)a.w_hi[k]); }
  for (int i = 0; i < cnt; ++i) {
    if (mode == 0) rgb_exact3(d[i], ci[i], a, true, out + 3 * i);
    else if (mode == 1) rgb_exact3(d[i], ci[i], a, false, out + 3 * i);
    else if (mode == 2) rgb_schlick_wall(d[i], ci[i], a, out + 3 * i);
    else rgb_two_beam((float)d[i], ci[i], a, out + 3 * i);
  }
}
// the float-float phase of the f32 fast paths: stored thickness h and offset h_ref in metres
extern "C" void eval_ff(int gamma, double n, const float *h, const float *href, const float *ci, int cnt, float *out) {
  RgbArgs a{};
  a.n_film = (float)n; a.inv_n2_d = 1.0 / (n * n);
  const double wl[3] = {650.0, 532.0, 450.0};
  for (int k = 0; k < 3; ++k) a.two_n_over_wl[k] = 2.0 * n / wl[k];
  a.inv_n2_hi = (float)a.inv_n2_d; a.inv_n2_lo = (float)(a.inv_n2_d - (double)a.inv_n2_hi);
  for (int k = 0; k < 3; ++k) { a.w_hi[k] = (float)a.two_n_over_wl[k]; a.w_lo[k] = (float)(a.two_n_over_wl[k] - (double)a.w_hi[k]); }
  for (int i = 0; i < cnt; ++i) rgb_exact3_ff(h[i], href[i], ci[i], a, gamma != 0, out + 3 * i);
}
""")
    so = tmp_path / "libcolour.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", str(cpp), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(2)
    d = np.concatenate([rng.uniform(0.0, 2000.0, 20000), [0.0, 500.0, 1000.0, 1e-3]])
    ci = np.concatenate([rng.uniform(0.0, 1.0, 20000), [1.0, 1.0, 0.0, 0.02]]).astype(np.float32)
    out = np.empty((len(d), 3), np.float32)

    def run(mode):
        lib.eval(mode, ctypes.c_double(1.33), d.ctypes.data_as(ctypes.c_void_p), ci.ctypes.data_as(ctypes.c_void_p), len(d),
                 out.ctypes.data_as(ctypes.c_void_p))
        return out.astype(np.float64)

    ora = O.OracleInterference(1.33)
    assert np.abs(run(0) - ora.batch(d, ci.astype(np.float64), gamma=True)).max() <= 1e-4
    assert np.abs(run(1) - ora.batch(d, ci.astype(np.float64), gamma=False)).max() <= 1e-4
    # float-float phase (the f32 fast paths): thickness given as stored value + offset, plain (h_ref = 0), offset
    # storage (h_ref = 500 nm) and a thick film (30 um, ~180 fringe orders); the oracle sees the thickness the pair holds
    for href_m, scale in ((0.0, 1.0), (500e-9, 1.0), (0.0, 15.0)):
        href = np.full(len(d), href_m, np.float32)
        hs = (d * scale * 1e-9 - href.astype(np.float64)).astype(np.float32)
        d_pair = (hs.astype(np.float64) + href.astype(np.float64)) * 1e9
        for gamma in (1, 0):
            lib.eval_ff(gamma, ctypes.c_double(1.33), hs.ctypes.data_as(ctypes.c_void_p), href.ctypes.data_as(ctypes.c_void_p),
                        ci.ctypes.data_as(ctypes.c_void_p), len(d), out.ctypes.data_as(ctypes.c_void_p))
            ref = ora.batch(d_pair, ci.astype(np.float64), gamma=bool(gamma))
            err = np.abs(out.astype(np.float64) - ref).max()
            assert err <= 1e-4, (href_m, scale, gamma, err)
    sub = np.r_[0:2000, len(d) - 4:len(d)]  # the scalar oracle entry points: a subset
    wall, two = run(2).copy(), run(3).copy()
    for i in sub:
        assert np.abs(wall[i] - np.array(O.schlick_wall_rgb(float(d[i]), float(ci[i])))).max() <= 1e-4, i
        assert np.abs(two[i] - np.clip(O.schlick_two_beam_rgb(float(d[i]), float(ci[i])), 0.0, 1.0)).max() <= 1e-4, i  # the kernel saturates


def test_band_height_model_picks_whole_waves(tmp_path):
    """auto_band_rows (csrc/sbx_launch.h, host code): the band height that fills whole waves of CTAs on launches of a few
    waves, the default (24 rows) from 8 waves up.  The documented choices -- 24 rows at 8192 x 16384 (16 strips, 592 CTA
    slots: 9.2 waves; round 1 used the whole-wave model's 56 rows there), 14 rows at 2048 x 4096 (4 strips: one wave of
    588 CTAs) -- are what DESIGN.md 4.1 / BASELINE.md quote."""
    cpp = tmp_path / "band.cpp"
    cpp.write_text('#include "sbx_launch.h"\nextern "C" int band(int rows, long long per_band, int slots, int dflt)\n'
                   "{ return sbx::auto_band_rows(rows, per_band, slots, dflt); }\n")
    so = tmp_path / "libband.so"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                           "-I" + os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc"), str(cpp), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    lib.band.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    slots = 148 * 4
    b3 = lib.band(8190, 16, slots, 24)
    assert b3 == 24  # 9.2 waves: the default height (measured best, profiles/r02_ab_band_rows.txt)
    assert lib.band(4094, 16, slots, 24) == 24  # 4.6 waves at 24 rows: the default as well
    assert lib.band(2046, 16, slots, 24) == 28  # 2.3 waves at 24 rows -> 28 rows = exactly two waves (1184 CTAs)
    assert lib.band(1022, 16, slots, 24) == 28  # one wave: the whole-wave model (592 CTAs)
    b2 = lib.band(2046, 4, slots, 24)
    assert b2 == 14 and math.ceil(2046 / b2) * 4 <= slots
    assert lib.band(254, 4096, slots, 24) == 24   # foam of 4096 bubbles: many waves, default height
    assert lib.band(10, 1, slots, 24) == 6    # a tiny grid is split into two bands that run side by side
    assert lib.band(0, 1, slots, 24) == 1
    for rows, per_band in ((8190, 16), (2046, 4), (1022, 2), (254, 512), (62, 1)):
        b = lib.band(rows, per_band, slots, 24)
        assert 1 <= b <= max(rows, 1) and b <= 64
