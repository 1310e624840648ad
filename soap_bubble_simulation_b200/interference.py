"""InterferenceCalculator, RgbColor and the interference LUT -- mirror of src/physics/interference.rs
and src/render/interference_lut.rs over libsbx.

Every evaluation runs on the GPU (f64 kernel k_interference_eval for the calculator, f32 kernel
k_lut_build for the LUT); the scalar methods of the Rust API are one-element batches of the same
kernels.  Use ``batch`` for arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L

# wavelengths (interference.rs:24-31) and constants (:34-39)
RED_NM, GREEN_NM, BLUE_NM = 650.0, 532.0, 450.0
REFRACTIVE_INDEX_AIR = 1.0
REFRACTIVE_INDEX_SOAP_FILM = 1.33

# interference_lut.rs:9-13
LUT_THICKNESS_SAMPLES = 256
LUT_ANGLE_SAMPLES = 64
LUT_MAX_THICKNESS_NM = 2000.0


def _clamp01(x: float) -> float:
    return min(max(float(x), 0.0), 1.0)


@dataclass(frozen=True)
class RgbColor:
    """RgbColor (interference.rs:43-95)."""
    red: float = 0.0
    green: float = 0.0
    blue: float = 0.0

    @classmethod
    def new(cls, red, green, blue) -> "RgbColor":  # :54-60 (clamping constructor)
        return cls(_clamp01(red), _clamp01(green), _clamp01(blue))

    @classmethod
    def from_f32(cls, red, green, blue) -> "RgbColor":  # :63-65
        return cls.new(float(np.float32(red)), float(np.float32(green)), float(np.float32(blue)))

    def to_f32_tuple(self):  # :68-70
        return (float(np.float32(self.red)), float(np.float32(self.green)), float(np.float32(self.blue)))

    def to_u8_tuple(self):  # :73-79 -- round half away from zero
        return tuple(int(np.floor(c * 255.0 + 0.5)) for c in (self.red, self.green, self.blue))

    def to_f32_array_with_alpha(self, alpha: float):  # :82-84
        return [*self.to_f32_tuple(), float(np.float32(alpha))]


@dataclass(frozen=True)
class FresnelCoefficients:  # interference.rs:99-106
    reflectance_s: float
    reflectance_p: float
    reflectance_average: float


class InterferenceCalculator:
    """InterferenceCalculator (interference.rs:127-513)."""

    def __init__(self, refractive_index_film: float = REFRACTIVE_INDEX_SOAP_FILM, *, device: int = 0,
                 wavelengths_nm=(RED_NM, GREEN_NM, BLUE_NM)):
        self._n = float(refractive_index_film)
        self._n_medium = REFRACTIVE_INDEX_AIR
        self._wl = tuple(float(w) for w in wavelengths_nm)
        self._device = device

    @classmethod
    def new(cls, refractive_index_film: float, **kw):  # :141-151
        return cls(refractive_index_film, **kw)

    @classmethod
    def default_soap_film(cls, **kw):  # :154-156
        return cls(REFRACTIVE_INDEX_SOAP_FILM, **kw)

    @classmethod
    def with_custom_wavelengths(cls, refractive_index_film, red_nm, green_nm, blue_nm, **kw):  # :165-176
        return cls(refractive_index_film, wavelengths_nm=(red_nm, green_nm, blue_nm), **kw)

    def refractive_index_film(self) -> float:  # :179-181
        return self._n

    def refractive_index_medium(self) -> float:  # :184-186
        return self._n_medium

    # -- batched device evaluation
    def batch(self, thickness_nm, cos_theta_incident, gamma: bool = True) -> np.ndarray:
        """(n,3) float64 colours for n (thickness, cos) pairs: calculate_interference_color (gamma) or
        _linear."""
        d = np.ascontiguousarray(np.broadcast_arrays(thickness_nm, cos_theta_incident)[0], dtype=np.float64).ravel()
        c = np.ascontiguousarray(np.broadcast_arrays(thickness_nm, cos_theta_incident)[1], dtype=np.float64).ravel()
        out = np.empty((d.size, 3), dtype=np.float64)
        wl = (C.c_double * 3)(*self._wl)
        L.check(L.lib().sbx_interference_eval(self._n, wl, d.ctypes.data, c.ctypes.data, d.size, int(gamma),
                                              out.ctypes.data, self._device))
        return out

    def _terms(self, d_nm, cos_i, wavelength_nm, want):
        d = np.ascontiguousarray(np.atleast_1d(d_nm), dtype=np.float64)
        c = np.ascontiguousarray(np.atleast_1d(cos_i), dtype=np.float64)
        n = d.size
        ct = np.empty(n) if want == "ct" else None
        fr = np.empty((n, 3)) if want == "fr" else None
        it = np.empty(n) if want == "it" else None
        L.check(L.lib().sbx_interference_terms(
            self._n, d.ctypes.data, c.ctypes.data, n, float(wavelength_nm),
            ct.ctypes.data if ct is not None else None, fr.ctypes.data if fr is not None else None,
            it.ctypes.data if it is not None else None, self._device))
        return ct if ct is not None else (fr if fr is not None else it)

    # -- the Rust API, scalar
    def calculate_transmission_angle_cos(self, cos_theta_incident: float) -> float:  # :197-211
        return float(self._terms(0.0, cos_theta_incident, 550.0, "ct")[0])

    def calculate_optical_path_difference(self, film_thickness_nm, cos_theta_transmitted, wavelength_nm) -> float:
        # :226-238 -- two multiplications and an addition on scalars already in host registers
        geometric_path = 2.0 * self._n * film_thickness_nm * cos_theta_transmitted
        return geometric_path + wavelength_nm / 2.0

    def calculate_phase_difference(self, optical_path_difference_nm, wavelength_nm) -> float:  # :248-254
        return 2.0 * np.pi * optical_path_difference_nm / wavelength_nm

    def calculate_fresnel_reflection(self, cos_theta_incident: float) -> FresnelCoefficients:  # :266-298
        fr = self._terms(0.0, cos_theta_incident, 550.0, "fr")[0]
        return FresnelCoefficients(float(fr[0]), float(fr[1]), float(fr[2]))

    def calculate_reflected_intensity(self, film_thickness_nm, cos_theta_incident, wavelength_nm) -> float:  # :312-353
        return float(self._terms(film_thickness_nm, cos_theta_incident, wavelength_nm, "it")[0])

    def calculate_interference_color(self, film_thickness_nm, cos_theta_incident) -> RgbColor:  # :380-400
        return RgbColor.new(*self.batch([film_thickness_nm], [cos_theta_incident], True)[0])

    def calculate_interference_color_linear(self, film_thickness_nm, cos_theta_incident) -> RgbColor:  # :413-426
        return RgbColor.new(*self.batch([film_thickness_nm], [cos_theta_incident], False)[0])

    def calculate_interference_color_with_index(self, film_thickness_nm, cos_theta_incident, refractive_index):  # :439-447
        return InterferenceCalculator(refractive_index, device=self._device).calculate_interference_color(
            film_thickness_nm, cos_theta_incident)

    def generate_color_lookup_table(self, min_thickness_nm, max_thickness_nm, num_samples, cos_theta):  # :461-477
        step = (max_thickness_nm - min_thickness_nm) / (num_samples - 1)
        d = [min_thickness_nm + i * step for i in range(num_samples)]
        return [RgbColor.new(*row) for row in self.batch(d, [cos_theta] * num_samples, True)]

    def generate_2d_color_lookup_table(self, min_thickness_nm, max_thickness_nm, thickness_samples, angle_samples):  # :491-512
        tstep = (max_thickness_nm - min_thickness_nm) / max(thickness_samples - 1, 1)
        astep = 1.0 / max(angle_samples - 1, 1)
        d = np.repeat([min_thickness_nm + i * tstep for i in range(thickness_samples)], angle_samples)
        c = np.tile([j * astep for j in range(angle_samples)], thickness_samples)
        rgb = self.batch(d, c, True).reshape(thickness_samples, angle_samples, 3)
        return [[RgbColor.new(*rgb[i, j]) for j in range(angle_samples)] for i in range(thickness_samples)]


def generate_interference_lut(refractive_index: float, intensity: float, *, device: int = 0) -> np.ndarray:
    """generate_interference_lut (interference_lut.rs:18-41): 65,536 bytes, row-major [angle][thickness] RGBA8."""
    out = np.empty(LUT_THICKNESS_SAMPLES * LUT_ANGLE_SAMPLES * 4, dtype=np.uint8)
    L.check(L.lib().sbx_lut_build(refractive_index, intensity, out.ctypes.data, device))
    return out


def generate_spectral_lut(refractive_index: float, intensity: float, wavelengths_nm, thickness_samples: int,
                          angle_samples: int, max_thickness_nm: float, *, device: int = 0) -> np.ndarray:
    """The same algorithm over an arbitrary wavelength list (BASELINE config 5); (A, T, 4) uint8."""
    wl = np.ascontiguousarray(wavelengths_nm, dtype=np.float32)
    out = np.empty((angle_samples, thickness_samples, 4), dtype=np.uint8)
    L.check(L.lib().sbx_lut_build_spectral(refractive_index, intensity, wl.ctypes.data, wl.size,
                                           thickness_samples, angle_samples, max_thickness_nm,
                                           out.ctypes.data, device))
    return out


def thin_film_interference(lut_rgba: np.ndarray, thickness_nm, cos_theta, interference_intensity: float = 4.0,
                           max_thickness_nm: float = LUT_MAX_THICKNESS_NM, *, device: int = 0) -> np.ndarray:
    """thin_film_interference() of the bubble fragment shader (src/render/shaders/bubble.wgsl:425-439) for arrays of
    (thickness, cos theta): bilinear LUT fetch through the GPU's texture unit, times intensity, clamped.
    lut_rgba: (A, T, 4) uint8 as returned by generate_interference_lut(...).reshape(64, 256, 4); returns (n, 3) f32."""
    lut = np.ascontiguousarray(lut_rgba, dtype=np.uint8)
    A, T = lut.shape[0], lut.shape[1]
    d = np.ascontiguousarray(thickness_nm, dtype=np.float32).ravel()
    c = np.ascontiguousarray(cos_theta, dtype=np.float32).ravel()
    out = np.empty((d.size, 3), dtype=np.float32)
    L.check(L.lib().sbx_lut_shade_points(lut.ctypes.data, T, A, max_thickness_nm, d.ctypes.data, c.ctypes.data,
                                         d.size, interference_intensity, out.ctypes.data, device))
    return out
