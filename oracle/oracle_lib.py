"""ctypes binding of oracle/libsbx_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference hot path (oracle/sbx_oracle.h).  Only tests/ (through the alias tests/oracle_lib.py),
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = _HERE
_SO = os.path.join(ORACLE_DIR, "libsbx_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("sbx_oracle.c", "sbx_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return _SO


class TfeFront(C.Structure):
    _fields_ = [("phi_center", C.c_double), ("extent", C.c_double), ("width", C.c_double),
                ("velocity", C.c_double)]


class WgslParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("gravity", C.c_float), ("viscosity", C.c_float),
                ("density", C.c_float), ("diffusion_coeff", C.c_float), ("bubble_radius", C.c_float),
                ("critical_thickness", C.c_float), ("grid_width", C.c_uint32),
                ("grid_height", C.c_uint32), ("marangoni_enabled", C.c_uint32),
                ("gamma_air", C.c_float), ("gamma_reduction", C.c_float),
                ("surfactant_diffusion", C.c_float), ("marangoni_coeff", C.c_float),
                ("_padding1", C.c_uint32), ("_padding2", C.c_uint32)]


class CausticParams(C.Structure):
    _fields_ = [("grid_width", C.c_uint32), ("grid_height", C.c_uint32), ("refractive_index", C.c_float),
                ("film_thickness_scale", C.c_float), ("light_dir_x", C.c_float), ("light_dir_y", C.c_float),
                ("light_dir_z", C.c_float), ("light_intensity", C.c_float), ("focal_length", C.c_float),
                ("caustic_intensity", C.c_float), ("caustic_sharpness", C.c_float), ("branch_threshold", C.c_float),
                ("ground_y", C.c_float), ("bubble_radius", C.c_float), ("time", C.c_float), ("_padding", C.c_uint32)]


class Interference(C.Structure):
    _fields_ = [("n_film", C.c_double), ("n_medium", C.c_double), ("wavelengths_nm", C.c_double * 3)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    d, i, p, f = C.c_double, C.c_int, C.c_void_p, C.c_float
    dp = C.POINTER(C.c_double)
    sig = {
        "orc_drainage_create": (p, [i, i, d, d, d, d, d, d, d]),
        "orc_drainage_create_window": (p, [i, i, i, i, d, d, d, d, d, d, d]),
        "orc_drainage_set_state": (i, [p, d, d, C.POINTER(TfeFront), i]),
        "orc_drainage_destroy": (None, [p]),
        "orc_drainage_step": (None, [p, d]),
        "orc_drainage_set_threads": (None, [p, i]),
        "orc_drainage_data": (dp, [p]),
        "orc_drainage_num_theta": (i, [p]),
        "orc_drainage_num_phi": (i, [p]),
        "orc_drainage_delta_theta": (d, [p]),
        "orc_drainage_delta_phi": (d, [p]),
        "orc_drainage_get_thickness": (d, [p, d, d]),
        "orc_drainage_min": (d, [p]),
        "orc_drainage_max": (d, [p]),
        "orc_drainage_has_critical": (i, [p]),
        "orc_drainage_critical": (d, [p]),
        "orc_drainage_time": (d, [p]),
        "orc_drainage_reset": (None, [p, d]),
        "orc_drainage_set_marginal_regeneration": (None, [p, i]),
        "orc_drainage_marginal_regeneration": (i, [p]),
        "orc_drainage_set_tfe_ratio": (None, [p, d]),
        "orc_drainage_tfe_ratio": (d, [p]),
        "orc_drainage_active_tfe_count": (i, [p]),
        "orc_drainage_fronts": (i, [p, C.POINTER(TfeFront)]),
        "orc_drainage_next_spawn": (d, [p]),
        "orc_wgsl_params_default": (None, [C.POINTER(WgslParams)]),
        "orc_wgsl_step": (None, [C.POINTER(WgslParams), p, p, p, p]),
        "orc_interference_new": (None, [C.POINTER(Interference), d]),
        "orc_transmission_angle_cos": (d, [C.POINTER(Interference), d]),
        "orc_optical_path_difference": (d, [C.POINTER(Interference), d, d, d]),
        "orc_phase_difference": (d, [d, d]),
        "orc_fresnel_reflection": (None, [C.POINTER(Interference), d, dp]),
        "orc_reflected_intensity": (d, [C.POINTER(Interference), d, d, d]),
        "orc_interference_color": (None, [C.POINTER(Interference), d, d, dp]),
        "orc_interference_color_linear": (None, [C.POINTER(Interference), d, d, dp]),
        "orc_rgb_to_u8": (None, [dp, C.POINTER(C.c_uint8)]),
        "orc_color_lut_1d": (None, [C.POINTER(Interference), d, d, i, d, dp]),
        "orc_color_lut_2d": (None, [C.POINTER(Interference), d, d, i, i, dp]),
        "orc_interference_batch": (None, [C.POINTER(Interference), p, p, C.c_size_t, i, p]),
        "orc_schlick_wall_rgb": (None, [f, f, f, C.POINTER(C.c_float)]),
        "orc_schlick_two_beam_rgb": (None, [f, f, f, C.POINTER(C.c_float)]),
        "orc_thin_film_interference": (None, [f, f, f, f, C.POINTER(C.c_float)]),
        "orc_generate_interference_lut": (None, [f, f, p]),
        "orc_generate_lut_spectral": (None, [f, f, p, i, i, i, f, p]),
        "orc_snells_law": (f, [f, f]),
        "orc_fresnel_unpolarized": (f, [f, f, f, f]),
        "orc_lut_shade": (None, [p, i, i, f, f, f, f, C.POINTER(C.c_float)]),
        "orc_lut_shade_w": (None, [p, i, i, f, f, f, f, i, C.POINTER(C.c_float)]),
        "orc_view_cos": (d, [d, d, d, dp]),
        "orc_caustic_params_default": (None, [C.POINTER(CausticParams)]),
        "orc_caustic_map": (None, [C.POINTER(CausticParams), p, p]),
        "orc_sample_thickness_uv": (None, [p, i, i, p, C.c_size_t, p, p]),
        "orc_sample_thickness_normal": (None, [p, i, i, p, C.c_size_t, C.c_float, p, p]),
        "orc_shade_field": (None, [p, i, i, i, i, d, dp, d, p, i, i, f, f, i, p, i]),
        "orc_shade_field_spectral": (None, [p, i, i, i, i, d, dp, f, f, p, i, p, p, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# default physical parameters: config/default.json, src/config.rs:26-35,56-66,93-103,305-316
DEFAULTS = dict(density=1000.0, gravity=9.81, viscosity=0.001, diffusion=1e-9,
                critical_thickness_m=10.0 * 1e-9, radius_m=0.05 / 2.0, h0_m=500.0 * 1e-9,
                dt=0.001, resolution=128, refractive_index=1.33)


class OracleDrainage:
    """DrainageSimulator (src/physics/drainage.rs:212-614) over the C oracle."""

    def __init__(self, num_theta=None, num_phi=None, h0_m=None, threads=1, window=None, **kw):
        """window=(row0, rows): hold only that latitude slab of the num_theta-row sphere (orc_drainage_create_window);
        ``data`` then has shape (rows, num_phi) and after k steps the k rows next to a cut edge are stale."""
        p = dict(DEFAULTS)
        p.update(kw)
        if num_theta is None:
            num_theta = p["resolution"]
        if num_phi is None:
            num_phi = 2 * num_theta
        if h0_m is None:
            h0_m = p["h0_m"]
        self._L = lib()
        if window is None:
            self._s = self._L.orc_drainage_create(num_theta, num_phi, h0_m, p["density"], p["gravity"],
                                                  p["viscosity"], p["diffusion"],
                                                  p["critical_thickness_m"], p["radius_m"])
            self.num_theta, self.num_phi = num_theta, num_phi
        else:
            row0, rows = int(window[0]), int(window[1])
            self._s = self._L.orc_drainage_create_window(num_theta, num_phi, row0, rows, h0_m, p["density"],
                                                         p["gravity"], p["viscosity"], p["diffusion"],
                                                         p["critical_thickness_m"], p["radius_m"])
            if not self._s:
                raise ValueError(f"bad window {window} for {num_theta} rows")
            self.num_theta, self.num_phi = rows, num_phi
            self.nt_global, self.row0 = num_theta, row0
        if threads > 1:
            self._L.orc_drainage_set_threads(self._s, threads)

    def __del__(self):
        if getattr(self, "_s", None):
            self._L.orc_drainage_destroy(self._s)
            self._s = None

    @property
    def data(self) -> np.ndarray:
        """Live view of the f64 field, shape (num_theta, num_phi)."""
        ptr = self._L.orc_drainage_data(self._s)
        return np.ctypeslib.as_array(ptr, shape=(self.num_theta, self.num_phi))

    def step(self, dt, n=1):
        for _ in range(n):
            self._L.orc_drainage_step(self._s, dt)

    def get_thickness(self, theta, phi):
        return self._L.orc_drainage_get_thickness(self._s, theta, phi)

    def min_thickness(self):
        return self._L.orc_drainage_min(self._s)

    def max_thickness(self):
        return self._L.orc_drainage_max(self._s)

    def has_critical_region(self):
        return bool(self._L.orc_drainage_has_critical(self._s))

    def critical_thickness(self):
        return self._L.orc_drainage_critical(self._s)

    def current_time(self):
        return self._L.orc_drainage_time(self._s)

    def reset(self, h0):
        self._L.orc_drainage_reset(self._s, h0)

    def set_marginal_regeneration(self, on):
        self._L.orc_drainage_set_marginal_regeneration(self._s, int(bool(on)))

    def marginal_regeneration_enabled(self):
        return bool(self._L.orc_drainage_marginal_regeneration(self._s))

    def set_tfe_thickness_ratio(self, r):
        self._L.orc_drainage_set_tfe_ratio(self._s, r)

    def tfe_thickness_ratio(self):
        return self._L.orc_drainage_tfe_ratio(self._s)

    def active_tfe_count(self):
        return self._L.orc_drainage_active_tfe_count(self._s)

    def set_state(self, time, next_spawn, fronts):
        buf = (TfeFront * 9)()
        for k, f in enumerate(fronts):
            buf[k].phi_center, buf[k].extent, buf[k].width, buf[k].velocity = f
        if self._L.orc_drainage_set_state(self._s, time, next_spawn, buf, len(fronts)) != 0:
            raise ValueError("bad state")

    def fronts(self):
        buf = (TfeFront * 9)()
        n = self._L.orc_drainage_fronts(self._s, buf)
        return [(buf[k].phi_center, buf[k].extent, buf[k].width, buf[k].velocity) for k in range(n)]

    def next_spawn(self):
        return self._L.orc_drainage_next_spawn(self._s)


class OracleInterference:
    """InterferenceCalculator (src/physics/interference.rs:127-513) over the C oracle."""

    def __init__(self, n_film=1.33, wavelengths=None):
        self._L = lib()
        self.c = Interference()
        self._L.orc_interference_new(C.byref(self.c), n_film)
        if wavelengths is not None:
            for k in range(3):
                self.c.wavelengths_nm[k] = wavelengths[k]

    def transmission_angle_cos(self, ci):
        return self._L.orc_transmission_angle_cos(C.byref(self.c), ci)

    def optical_path_difference(self, d, ct, wl):
        return self._L.orc_optical_path_difference(C.byref(self.c), d, ct, wl)

    def phase_difference(self, opd, wl):
        return self._L.orc_phase_difference(opd, wl)

    def fresnel(self, ci):
        out = (C.c_double * 3)()
        self._L.orc_fresnel_reflection(C.byref(self.c), ci, out)
        return tuple(out)

    def reflected_intensity(self, d, ci, wl):
        return self._L.orc_reflected_intensity(C.byref(self.c), d, ci, wl)

    def color(self, d, ci):
        out = (C.c_double * 3)()
        self._L.orc_interference_color(C.byref(self.c), d, ci, out)
        return tuple(out)

    def color_linear(self, d, ci):
        out = (C.c_double * 3)()
        self._L.orc_interference_color_linear(C.byref(self.c), d, ci, out)
        return tuple(out)

    def batch(self, d_nm, cos_i, gamma=True) -> np.ndarray:
        d = np.ascontiguousarray(d_nm, dtype=np.float64).ravel()
        c = np.ascontiguousarray(cos_i, dtype=np.float64).ravel()
        assert d.size == c.size
        out = np.empty((d.size, 3), dtype=np.float64)
        self._L.orc_interference_batch(C.byref(self.c), d.ctypes.data, c.ctypes.data, d.size,
                                       int(gamma), out.ctypes.data)
        return out

    def lut_1d(self, lo, hi, n, cos_theta):
        out = np.empty((n, 3), dtype=np.float64)
        self._L.orc_color_lut_1d(C.byref(self.c), lo, hi, n, cos_theta,
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def lut_2d(self, lo, hi, T, A):
        out = np.empty((T, A, 3), dtype=np.float64)
        self._L.orc_color_lut_2d(C.byref(self.c), lo, hi, T, A,
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
        return out


def rgb_to_u8(rgb):
    a = (C.c_double * 3)(*rgb)
    o = (C.c_uint8 * 3)()
    lib().orc_rgb_to_u8(a, o)
    return tuple(o)


def generate_interference_lut(n=1.33, intensity=1.0) -> np.ndarray:
    out = np.empty((64, 256, 4), dtype=np.uint8)
    lib().orc_generate_interference_lut(n, intensity, out.ctypes.data)
    return out


def generate_lut_spectral(n, intensity, wavelengths, T, A, max_nm) -> np.ndarray:
    wl = np.ascontiguousarray(wavelengths, dtype=np.float32)
    out = np.empty((A, T, 4), dtype=np.uint8)
    lib().orc_generate_lut_spectral(n, intensity, wl.ctypes.data, wl.size, T, A, max_nm, out.ctypes.data)
    return out


def thin_film_interference(d_nm, cos_theta, n=1.33, intensity=1.0):
    o = (C.c_float * 3)()
    lib().orc_thin_film_interference(d_nm, cos_theta, n, intensity, o)
    return tuple(o)


def schlick_wall_rgb(d_nm, cos_i, n=1.33):
    o = (C.c_float * 3)()
    lib().orc_schlick_wall_rgb(d_nm, cos_i, n, o)
    return tuple(o)


def schlick_two_beam_rgb(d_nm, cos_i, n=1.33):
    o = (C.c_float * 3)()
    lib().orc_schlick_two_beam_rgb(d_nm, cos_i, n, o)
    return tuple(o)


def lut_shade(lut: np.ndarray, d_nm, cos_theta, intensity=4.0, max_nm=2000.0, weight_bits=0):
    """weight_bits=0: ideal f32 bilinear; 8: weights rounded to 8 fractional bits like a GPU texture unit."""
    A, T = lut.shape[0], lut.shape[1]
    o = (C.c_float * 3)()
    lib().orc_lut_shade_w(lut.ctypes.data, T, A, max_nm, d_nm, cos_theta, intensity, weight_bits, o)
    return tuple(o)


def view_cos(theta, phi, radius, cam):
    a = (C.c_double * 3)(*cam)
    return lib().orc_view_cos(theta, phi, radius, a)


def wgsl_params(**kw) -> WgslParams:
    p = WgslParams()
    lib().orc_wgsl_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def wgsl_step(p: WgslParams, h: np.ndarray, c: np.ndarray):
    h = np.ascontiguousarray(h, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    ho, co = np.empty_like(h), np.empty_like(c)
    lib().orc_wgsl_step(C.byref(p), h.ctypes.data, c.ctypes.data, ho.ctypes.data, co.ctypes.data)
    return ho, co


def shade_field(h: np.ndarray, num_theta: int, row0: int, radius, cam, n_film=1.33, lut=None, intensity=4.0,
                mode=1, threads=1, max_nm=2000.0) -> np.ndarray:
    """Colour frame (rows, num_phi, 4) uint8 of the band h (rows, num_phi) starting at global row row0."""
    h = np.ascontiguousarray(h, dtype=np.float64)
    rows, nphi = h.shape
    out = np.empty((rows, nphi, 4), dtype=np.uint8)
    a = (C.c_double * 3)(*cam)
    if lut is None:
        lut = generate_interference_lut(n_film, 1.0)
    A, T = lut.shape[0], lut.shape[1]
    lib().orc_shade_field(h.ctypes.data, num_theta, nphi, row0, rows, radius, a, n_film, lut.ctypes.data, T, A,
                          max_nm, intensity, mode, out.ctypes.data, threads)
    return out


def shade_field_spectral(h: np.ndarray, num_theta: int, row0: int, radius, cam, wavelengths, n_film=1.33,
                         intensity=4.0, threads=1, as_float=False) -> np.ndarray:
    """Per-cell spectral colour (bubble_instanced.wgsl:190-218 with the wavelength list as a parameter) of the band h:
    RGBA8 (rows, num_phi, 4) or, as_float, (rows, num_phi, 3) float32."""
    h = np.ascontiguousarray(h, dtype=np.float64)
    rows, nphi = h.shape
    wl = np.ascontiguousarray(wavelengths, dtype=np.float32)
    a = (C.c_double * 3)(*cam)
    if as_float:
        out = np.empty((rows, nphi, 3), dtype=np.float32)
        lib().orc_shade_field_spectral(h.ctypes.data, num_theta, nphi, row0, rows, radius, a, n_film, intensity,
                                       wl.ctypes.data, len(wl), None, out.ctypes.data, threads)
    else:
        out = np.empty((rows, nphi, 4), dtype=np.uint8)
        lib().orc_shade_field_spectral(h.ctypes.data, num_theta, nphi, row0, rows, radius, a, n_film, intensity,
                                       wl.ctypes.data, len(wl), out.ctypes.data, None, threads)
    return out


def caustic_params(**kw) -> CausticParams:
    p = CausticParams()
    lib().orc_caustic_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def caustic_map(p: CausticParams, thickness: np.ndarray) -> np.ndarray:
    t = np.ascontiguousarray(thickness, dtype=np.float32)
    out = np.empty_like(t)
    lib().orc_caustic_map(C.byref(p), t.ctypes.data, out.ctypes.data)
    return out


def sample_thickness_uv(thickness: np.ndarray, uv: np.ndarray):
    t = np.ascontiguousarray(thickness, dtype=np.float32)
    q = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
    h = np.empty(len(q), np.float32)
    g = np.empty((len(q), 2), np.float32)
    lib().orc_sample_thickness_uv(t.ctypes.data, t.shape[1], t.shape[0], q.ctypes.data, len(q), h.ctypes.data, g.ctypes.data)
    return h, g


def sample_thickness_normal(thickness: np.ndarray, normals: np.ndarray, scale: float = 1.0):
    t = np.ascontiguousarray(thickness, dtype=np.float32)
    q = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    h = np.empty(len(q), np.float32)
    uv = np.empty((len(q), 2), np.float32)
    lib().orc_sample_thickness_normal(t.ctypes.data, t.shape[1], t.shape[0], q.ctypes.data, len(q), scale, h.ctypes.data, uv.ctypes.data)
    return h, uv
