"""shader_code mirror (src/physics/interference.rs:525-610): the reference only checks that the four function
names appear (interference.rs:783-790); here the GLSL text is additionally *executed* -- compiled as C++ behind a
small GLSL prelude -- and compared with the oracle's f64 calculator."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from soap_bubble_simulation_b200 import shader_code

PRELUDE = r"""
#include <cmath>
#include <algorithm>
struct vec3 { float x, y, z; vec3(float a) : x(a), y(a), z(a) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} };
static inline float clamp(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
static inline float max(float a, float b) { return std::max(a, b); }
static inline float abs(float a) { return std::fabs(a); }
static inline float sqrt(float a) { return std::sqrt(a); }
static inline float cos(float a) { return std::cos(a); }
static inline vec3 pow(vec3 a, vec3 b) { return vec3(std::pow(a.x, b.x), std::pow(a.y, b.y), std::pow(a.z, b.z)); }
"""

WRAPPERS = r"""
extern "C" void glsl_color(float d, float ci, float n, int gamma, float *out) {
    vec3 c = gamma ? calculateInterferenceColorGamma(d, ci, n) : calculateInterferenceColor(d, ci, n);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
extern "C" float glsl_cos_t(float ci, float n) { return calculateTransmissionAngleCos(ci, 1.0f, n); }
extern "C" float glsl_fresnel(float ci, float n) { return calculateFresnelReflectance(ci, 1.0f, n); }
"""


def test_shader_code_generation():  # interference.rs:783-790
    code = shader_code.complete_shader_code()
    for name in ("calculateTransmissionAngleCos", "calculateFresnelReflectance", "calculateInterferenceIntensity",
                 "calculateInterferenceColor"):
        assert name in code
    assert code.index("calculateTransmissionAngleCos") < code.index("float calculateFresnelReflectance")
    for part in (shader_code.SNELLS_LAW_GLSL, shader_code.FRESNEL_GLSL, shader_code.INTERFERENCE_INTENSITY_GLSL,
                 shader_code.INTERFERENCE_COLOR_GLSL):
        assert part in code


@pytest.fixture(scope="module")
def glsl(tmp_path_factory):
    d = tmp_path_factory.mktemp("glsl")
    src = d / "glsl.cpp"
    src.write_text(PRELUDE + shader_code.complete_shader_code() + WRAPPERS)
    so = d / "glsl.so"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(os.fspath(so))
    lib.glsl_color.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.glsl_cos_t.restype = ctypes.c_float
    lib.glsl_cos_t.argtypes = [ctypes.c_float, ctypes.c_float]
    lib.glsl_fresnel.restype = ctypes.c_float
    lib.glsl_fresnel.argtypes = [ctypes.c_float, ctypes.c_float]
    return lib


def test_glsl_text_matches_the_calculator(glsl):
    calc = O.OracleInterference(1.33)
    assert abs(glsl.glsl_fresnel(1.0, 1.33) - 0.020059312199524774) < 1e-7  # SURVEY 8c KAT
    assert glsl.glsl_cos_t(0.0, 1.33) == pytest.approx(calc.transmission_angle_cos(0.0), abs=1e-6)
    out = (ctypes.c_float * 3)()
    rng = np.random.default_rng(7)
    worst = 0.0
    for d, ci in zip(rng.uniform(0.0, 1500.0, 400), rng.uniform(0.05, 1.0, 400)):
        glsl.glsl_color(d, ci, 1.33, 0, out)
        ref = calc.color_linear(float(d), float(ci))
        worst = max(worst, max(abs(out[k] - ref[k]) for k in range(3)))
    # f32 phase (relative 1e-7 on up to ~60 rad) and the 1e-4 denominator guard: well inside 1e-3 of the f64 result
    assert worst < 1e-3, worst
    glsl.glsl_color(500.0, 1.0, 1.33, 1, out)
    ref = calc.color(500.0, 1.0)
    assert abs(out[0] - ref[0]) < 2e-3 and abs(out[2] - ref[2]) < 2e-3
    glsl.glsl_color(0.0, 1.0, 1.33, 0, out)
    assert tuple(out) == (0.0, 0.0, 0.0)  # interference.rs:682-686
