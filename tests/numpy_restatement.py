"""A SECOND, independent restatement of the reference's CPU drainage path, written in numpy straight from the Rust
source (src/physics/drainage.rs) without looking at oracle/sbx_oracle.c.  It exists only to cross-check the C oracle:
two transcriptions made separately must agree bit for bit (both are plain IEEE f64, no FMA contraction, glibc libm for
the per-row sin/cos and the TFE pseudo-random numbers).  TEST INFRASTRUCTURE ONLY.

Vectorised over phi; the operation order inside a cell follows the Rust expressions token for token."""
import math

import numpy as np

PI = math.pi
FRAC_PI_2 = math.pi / 2.0


class TfeFront:  # drainage.rs:23-88
    def __init__(self, phi_center, width, velocity):
        self.phi_center, self.extent, self.width, self.velocity = phi_center, 0.0, width, velocity

    def advance(self, dt):  # :47-49
        self.extent = min(self.extent + self.velocity * dt, FRAC_PI_2)

    def contains_row(self, theta, phi):  # :55-71, for one theta and an array of phi
        if abs(theta - FRAC_PI_2) > self.extent:
            return np.zeros(phi.shape, bool)
        phi_diff = np.abs(phi - self.phi_center)
        phi_diff = np.where(phi_diff > PI, 2.0 * PI - phi_diff, phi_diff)
        return phi_diff < self.width / 2.0

    def theta_range(self):  # :76-79
        return FRAC_PI_2 - self.extent, FRAC_PI_2 + self.extent


class NumpyDrainage:
    """DrainageSimulator (drainage.rs:212-614) with config/default.json's physical parameters."""

    def __init__(self, num_theta, num_phi, h0=500.0 * 1e-9, density=1000.0, gravity=9.81, viscosity=0.001,
                 diffusion=1e-9, critical=10.0 * 1e-9, radius=0.05 / 2.0):
        # defaults as DrainageSimulator::new derives them (drainage.rs:266-279): film_thickness_nm * 1e-9,
        # critical_thickness_nm * 1e-9, diameter / 2 -- 500.0 * 1e-9 is one ulp away from the literal 500e-9
        self.nt, self.nphi = num_theta, num_phi
        self.delta_theta = PI / (num_theta - 1)        # :123
        self.delta_phi = 2.0 * PI / num_phi            # :124
        self.h = np.full((num_theta, num_phi), h0, np.float64)
        self.density, self.gravity, self.viscosity = density, gravity, viscosity
        self.diffusion, self.critical, self.radius = diffusion, critical, radius
        self.time = 0.0
        self.fronts = []
        self.ratio, self.next_spawn, self.spawn_interval, self.tfe = 0.85, 0.5, 2.0, True  # :280-286

    def step(self, dt):  # :298-392
        nt, nphi, dth, dph = self.nt, self.nphi, self.delta_theta, self.delta_phi
        drainage_coefficient = self.density * self.gravity / (3.0 * self.viscosity)
        scratch = self.h.copy()
        for i in range(1, nt - 1):
            theta = i * dth                                # theta_at, :143-145
            sin_theta, cos_theta = math.sin(theta), math.cos(theta)
            sin_safe = math.copysign(1e-10, sin_theta) if abs(sin_theta) < 1e-10 else sin_theta
            cur = scratch[i]
            tm, tp = scratch[i - 1], scratch[i + 1]
            pm, pp = np.roll(cur, 1), np.roll(cur, -1)     # (phi + n - 1) % n, (phi + 1) % n
            cubed = cur * cur * cur
            drainage_term = -drainage_coefficient * cubed * sin_theta
            radius_squared = self.radius * self.radius
            d2h_dtheta2 = (tp - 2.0 * cur + tm) / (dth * dth)
            dh_dtheta = (tp - tm) / (2.0 * dth)
            d2h_dphi2 = (pp - 2.0 * cur + pm) / (dph * dph)
            laplacian = (d2h_dtheta2 + cos_theta / sin_safe * dh_dtheta + d2h_dphi2 / (sin_safe * sin_safe)) / radius_squared
            dh_dt = drainage_term + self.diffusion * laplacian
            new = np.fmax(cur + dt * dh_dt, 0.0)           # f64::max: a NaN operand yields the other one
            self.h[i] = np.where(cur < self.critical, cur, new)   # `continue` leaves the cell as it was
        self._update_poles()
        if self.tfe:
            self._marginal_regeneration(dt)
        self.time += dt

    def _update_poles(self):  # :480-503, sequential left-to-right sums
        n = self.nphi
        self.h[0, :] = np.cumsum(self.h[1])[-1] / float(n)
        self.h[-1, :] = np.cumsum(self.h[-2])[-1] / float(n)

    def _spawn(self):  # :464-477
        t = self.time
        phi = abs(math.sin(t * 7.89 + 1.23)) * 2.0 * PI
        width = 0.3 + abs(math.sin(t * 3.21)) * 0.4
        velocity = 0.05 + abs(math.cos(t * 5.67)) * 0.05
        self.fronts.append(TfeFront(phi, width, velocity))
        if len(self.fronts) > 8:
            self.fronts.pop(0)

    def _marginal_regeneration(self, dt):  # :395-461
        self.next_spawn -= dt
        if self.next_spawn <= 0.0:
            self._spawn()
            self.next_spawn = self.spawn_interval * (0.5 + abs(math.sin(self.time * 12.345)))
        for f in self.fronts:
            f.advance(dt)
        self.fronts = [f for f in self.fronts if f.extent < FRAC_PI_2 * 0.95]
        if not self.fronts:
            return
        nt, dth = self.nt, self.delta_theta
        lo, hi = nt, 0
        for f in self.fronts:
            mn, mx = f.theta_range()
            lo = min(lo, max(0, math.floor(mn / dth)))                 # `as usize` saturates negatives to 0
            hi = max(hi, min(max(0, math.ceil(mx / dth)), nt - 1))
        lo = min(lo, nt - 1)
        phi = np.arange(self.nphi, dtype=np.float64) * self.delta_phi
        reduction_rate = (1.0 - self.ratio) * 0.5
        for i in range(lo, hi + 1):
            theta = i * dth
            hit = np.zeros(self.nphi, bool)
            for f in self.fronts:                                       # one front per point: any() is enough
                hit |= f.contains_row(theta, phi)
            self.h[i] = np.where(hit, self.h[i] * (1.0 - reduction_rate * dt), self.h[i])

    def get_thickness(self, theta, phi):  # :516-552
        tc = min(max(theta, 0.0), PI) / self.delta_theta
        pc = (phi % (2.0 * PI)) / self.delta_phi      # rem_euclid for a positive modulus
        ti = min(int(math.floor(tc)), self.nt - 2)
        pj = int(math.floor(pc)) % self.nphi
        pj_hi = (pj + 1) % self.nphi
        tf, pf = tc - ti, pc - pj
        v0 = self.h[ti, pj] * (1.0 - pf) + self.h[ti, pj_hi] * pf
        v1 = self.h[ti + 1, pj] * (1.0 - pf) + self.h[ti + 1, pj_hi] * pf
        return v0 * (1.0 - tf) + v1 * tf


class PyInterference:
    """InterferenceCalculator (src/physics/interference.rs:127-447) in scalar Python floats (IEEE f64, glibc libm via
    the math module), transcribed from the Rust expressions; second entry for the oracle's calculator."""

    def __init__(self, n_film=1.33, wavelengths=(650.0, 532.0, 450.0), n_medium=1.0):  # :26-30, :141-151
        self.n_film, self.n_medium, self.wl = n_film, n_medium, tuple(wavelengths)

    @staticmethod
    def _clamp(x, lo, hi):
        return lo if x < lo else hi if x > hi else x

    def transmission_angle_cos(self, ci):  # :197-211
        c = self._clamp(ci, -1.0, 1.0)
        sin_i = math.sqrt(1.0 - c * c)
        sin_t = (self.n_medium / self.n_film) * sin_i
        if abs(sin_t) > 1.0:
            return 0.0
        return math.sqrt(1.0 - sin_t * sin_t)

    def optical_path_difference(self, d, ct, wl):  # :226-238
        return 2.0 * self.n_film * d * ct + wl / 2.0

    @staticmethod
    def phase_difference(opd, wl):  # :248-254
        return 2.0 * PI * opd / wl

    def fresnel(self, ci):  # :266-298
        c = self._clamp(ci, -1.0, 1.0)
        ct = self.transmission_angle_cos(c)
        n1, n2 = self.n_medium, self.n_film
        n1ci, n2ct, n2ci, n1ct = n1 * c, n2 * ct, n2 * c, n1 * ct
        rs = (n1ci - n2ct) / (n1ci + n2ct)
        rp = (n2ci - n1ct) / (n2ci + n1ct)
        Rs, Rp = rs * rs, rp * rp
        return Rs, Rp, (Rs + Rp) / 2.0

    def reflected_intensity(self, d, ci, wl):  # :312-353
        if d <= 0.0:
            return 0.0
        ct = self.transmission_angle_cos(ci)
        phase = self.phase_difference(self.optical_path_difference(d, ct, wl), wl)
        R = self.fresnel(ci)[2]
        cos_phase = math.cos(phase)
        num = 2.0 * R * (1.0 - cos_phase)
        den = 1.0 + R * R - 2.0 * R * cos_phase
        if abs(den) < 2.220446049250313e-16:
            return R
        return self._clamp(num / den, 0.0, 1.0)

    def color_linear(self, d, ci):  # :413-426 (RgbColor::new clamps, :52-61)
        return tuple(self._clamp(self.reflected_intensity(d, ci, w), 0.0, 1.0) for w in self.wl)

    def color(self, d, ci):  # :380-400
        return tuple(self._clamp(math.pow(self.reflected_intensity(d, ci, w), 1.0 / 2.2), 0.0, 1.0) for w in self.wl)


def wgsl_drainage_step(p, h, c):
    """drainage.wgsl::drainage_step (src/render/shaders/drainage.wgsl:72-202) in numpy float32, transcribed from the
    shader text: every intermediate is an f32 array or scalar, so each +, -, *, / rounds once to f32 like the shader's;
    sin/cos of the row angle come from glibc's sinf/cosf.  `p` is any object with the DrainageParams fields
    (gpu_drainage.rs:10-44); h, c: (grid_height, grid_width) float32.  Returns (h_out, c_out)."""
    import ctypes

    f32 = np.float32
    libm = ctypes.CDLL("libm.so.6")
    for fn in (libm.sinf, libm.cosf):
        fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
    H, W = int(p.grid_height), int(p.grid_width)
    h = np.ascontiguousarray(h, f32).reshape(H, W)
    c = np.ascontiguousarray(c, f32).reshape(H, W)
    ho, co = h.copy(), c.copy()                      # pole rows and sub-critical cells copy through (:85-89, :108-112)
    PI32 = f32(3.14159265358979323846)
    delta_theta = PI32 / f32(H - 1)
    delta_phi = f32(2.0) * PI32 / f32(W)
    dt, R = f32(p.dt), f32(p.bubble_radius)
    drainage_coeff = f32(p.density) * f32(p.gravity) / (f32(3.0) * f32(p.viscosity))
    radius_squared = R * R
    mar = int(p.marangoni_enabled) != 0
    for i in range(1, H - 1):
        theta = f32(i) * delta_theta
        sin_theta, cos_theta = f32(libm.sinf(float(theta))), f32(libm.cosf(float(theta)))
        sin_safe = f32(np.sign(sin_theta)) * f32(0.0001) if abs(sin_theta) < f32(0.0001) else sin_theta
        hh, cc = h[i], c[i]
        h_tm, h_tp, h_pm, h_pp = h[i - 1], h[i + 1], np.roll(hh, 1), np.roll(hh, -1)
        c_tm, c_tp, c_pm, c_pp = c[i - 1], c[i + 1], np.roll(cc, 1), np.roll(cc, -1)
        h_cubed = hh * hh * hh
        drainage_term = -drainage_coeff * h_cubed * sin_theta
        d2h_dtheta2 = (h_tp - f32(2.0) * hh + h_tm) / (delta_theta * delta_theta)
        dh_dtheta = (h_tp - h_tm) / (f32(2.0) * delta_theta)
        d2h_dphi2 = (h_pp - f32(2.0) * hh + h_pm) / (delta_phi * delta_phi)
        h_lap = (d2h_dtheta2 + cos_theta / sin_safe * dh_dtheta + d2h_dphi2 / (sin_safe * sin_safe)) / radius_squared
        diffusion_term = f32(p.diffusion_coeff) * h_lap
        marangoni_term = np.zeros(W, f32)
        dc_dtheta = (c_tp - c_tm) / (f32(2.0) * delta_theta)
        if mar:
            dc_dphi = (c_pp - c_pm) / (f32(2.0) * delta_phi)
            g_t = dc_dtheta / R
            g_p = dc_dphi / (R * sin_safe)
            grad_c_mag = np.sqrt(g_t * g_t + g_p * g_p)
            marangoni_term = f32(p.marangoni_coeff) * (f32(p.gamma_reduction) * grad_c_mag) * hh * hh
        dh_dt = drainage_term + diffusion_term + marangoni_term
        new_h = np.fmax(hh + dt * dh_dt, f32(0.0))
        d2c_dtheta2 = (c_tp - f32(2.0) * cc + c_tm) / (delta_theta * delta_theta)
        d2c_dphi2 = (c_pp - f32(2.0) * cc + c_pm) / (delta_phi * delta_phi)
        c_lap = (d2c_dtheta2 + cos_theta / sin_safe * dc_dtheta + d2c_dphi2 / (sin_safe * sin_safe)) / radius_squared
        dc_dt = f32(p.surfactant_diffusion) * c_lap
        if mar:
            dc_dt = dc_dt + f32(0.001) * (h_lap * R * R) * cc
        new_c = np.clip(cc + dt * dc_dt, f32(0.0), f32(1.0))
        skip = hh < f32(p.critical_thickness)
        ho[i] = np.where(skip, hh, new_h)
        co[i] = np.where(skip, cc, new_c)
    return ho, co


def interference_lut(refractive_index=1.33, intensity=1.0, T=256, A=64, max_nm=2000.0):
    """generate_interference_lut (src/render/interference_lut.rs:18-163) in numpy float32, vectorised over the thickness
    axis; transcribed from the Rust text (f32 operations in its order, glibc sinf/expf, `as u8` truncation)."""
    import ctypes

    f32 = np.float32
    libm = ctypes.CDLL("libm.so.6")
    for fn in (libm.sinf, libm.expf):
        fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
    sinf = np.vectorize(lambda x: libm.sinf(float(x)), otypes=[np.float32])
    PI32 = f32(math.pi)
    n = f32(refractive_index)

    def gaussian(x, mean, sigma):  # :151-155
        t = (f32(x) - f32(mean)) / f32(sigma)
        return f32(libm.expf(float(f32(-0.5) * t * t)))

    def cie(w):  # :136-148
        x = f32(1.056) * gaussian(w, 599.8, 37.9) + f32(0.362) * gaussian(w, 442.0, 16.0) - f32(0.065) * gaussian(w, 501.1, 20.4)
        y = f32(0.821) * gaussian(w, 568.8, 46.9) + f32(0.286) * gaussian(w, 530.9, 31.1)
        z = f32(1.217) * gaussian(w, 437.0, 11.8) + f32(0.681) * gaussian(w, 459.0, 26.0)
        return max(x, f32(0.0)), max(y, f32(0.0)), max(z, f32(0.0))

    out = np.empty((A, T, 4), np.uint8)
    thickness = (np.arange(T, dtype=f32) / f32(T - 1)) * f32(max_nm)                    # :25-26
    for j in range(A):
        ci = f32(j) / f32(A - 1)                                                          # :27
        sin_i = np.sqrt(max(f32(1.0) - ci * ci, f32(0.0)))                                # :96-100
        sin_t = sin_i / n
        ct = np.sqrt(max(f32(1.0) - sin_t * sin_t, f32(0.0)))
        n1ci, n2ct, n2ci, n1ct = f32(1.0) * ci, n * ct, n * ci, f32(1.0) * ct             # :103-120
        r_s = (n1ci - n2ct) / max(n1ci + n2ct, f32(0.0001))
        r_p = (n2ci - n1ct) / max(n2ci + n1ct, f32(0.0001))
        R = (r_s * r_s + r_p * r_p) * f32(0.5)
        optical_path = f32(2.0) * n * thickness * ct                                      # :61
        one_minus_r = max(f32(1.0) - R, f32(0.001))                                       # :123-133
        F = f32(4.0) * R / (one_minus_r * one_minus_r)
        xyz = [np.zeros(T, f32), np.zeros(T, f32), np.zeros(T, f32)]
        for w in (400.0, 450.0, 500.0, 550.0, 600.0, 650.0, 700.0):                       # :52, :66-78
            phase = f32(2.0) * PI32 * optical_path / f32(w) + PI32
            s = sinf(phase * f32(0.5))
            sin2 = s * s
            airy = (F * sin2) / np.maximum(f32(1.0) + F * sin2, f32(0.001))
            cx, cy, cz = cie(w)
            xyz[0] = xyz[0] + cx * airy
            xyz[1] = xyz[1] + cy * airy
            xyz[2] = xyz[2] + cz * airy
        X, Y, Z = (v / f32(7.0) for v in xyz)                                             # :81-83
        r = f32(3.2404542) * X - f32(1.5371385) * Y - f32(0.4985314) * Z                  # :158-163
        g = f32(-0.969_266) * X + f32(1.8760108) * Y + f32(0.0415560) * Z
        b = f32(0.0556434) * X - f32(0.2040259) * Y + f32(1.0572252) * Z
        for k, ch in enumerate((r, g, b)):
            out[j, :, k] = (np.clip(ch * f32(intensity), f32(0.0), f32(1.0)) * f32(255.0)).astype(np.uint8)  # :33-35
        out[j, :, 3] = 255
    return out


def caustic_map(p, thickness):
    """caustics_compute.wgsl::main (src/render/shaders/caustics_compute.wgsl:29-126) in numpy float32 over the whole
    grid: x = phi wraps, y = theta clamps at the poles; powf from glibc.  `p`: CausticParams fields; returns (H, W) f32."""
    import ctypes

    f32 = np.float32
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype, libm.powf.argtypes = ctypes.c_float, [ctypes.c_float, ctypes.c_float]
    powf = np.vectorize(lambda x, y: libm.powf(float(x), float(y)), otypes=[np.float32])
    t = np.ascontiguousarray(thickness, f32)
    H, W = t.shape

    def s(dx, dy):  # sample_thickness(x + dx, y + dy), :29-43
        rows = np.clip(np.arange(H) + dy, 0, H - 1)
        return np.roll(t[rows], -dx, axis=1)

    center, left, right, up, down = s(0, 0), s(-1, 0), s(1, 0), s(0, -1), s(0, 1)
    h = f32(1.0)
    h2 = h * h
    gx = (right - left) / (f32(2.0) * h)                                   # :46-51
    gy = (down - up) / (f32(2.0) * h)
    laplacian = (left + right + up + down - f32(4.0) * center) / h2        # :54-65
    d2x = (left + right - f32(2.0) * center) / h2                          # :68-88
    d2y = (up + down - f32(2.0) * center) / h2
    ul, ur, dl, dr = s(-1, -1), s(1, -1), s(-1, 1), s(1, 1)
    dxy = (dr - dl - ur + ul) / (f32(4.0) * h2)
    hessian_det = d2x * d2y - dxy * dxy
    curvature_factor = -laplacian * f32(p.focal_length)                    # :91-112
    focusing = f32(1.0) / np.maximum(np.abs(f32(1.0) - curvature_factor), f32(0.01))
    branch_factor = np.where(hessian_det < -f32(p.branch_threshold), f32(1.5), f32(1.0))
    grad_mag = np.sqrt(gx * gx + gy * gy)
    refraction_factor = f32(1.0) + grad_mag * f32(0.5)
    intensity = focusing * branch_factor * refraction_factor
    intensity = powf(intensity, f32(p.caustic_sharpness))
    intensity = intensity * f32(p.caustic_intensity)
    return np.clip(intensity, f32(0.0), f32(10.0)).astype(f32)


def sample_thickness_at_uv(t, uv):
    """sample_thickness_at_uv / thickness_gradient_uv (branched_flow_compute.wgsl:187-210), generalised from 128 x 64
    to t.shape: nearest texel by truncation, eps = 0.01 central differences.  Returns (h, grad)."""
    f32 = np.float32
    t = np.ascontiguousarray(t, f32)
    H, W = t.shape
    uv = np.ascontiguousarray(uv, f32).reshape(-1, 2)

    def at(u, v):
        x = (np.clip(u, f32(0.0), f32(1.0)) * f32(W - 1)).astype(np.uint32)   # u32() truncates
        y = (np.clip(v, f32(0.0), f32(1.0)) * f32(H - 1)).astype(np.uint32)
        return t[y, x]

    u, v = uv[:, 0], uv[:, 1]
    eps = f32(0.01)
    gx = (at(u + eps, v) - at(u - eps, v)) / (f32(2.0) * eps)
    gy = (at(u, v + eps) - at(u, v - eps)) / (f32(2.0) * eps)
    return at(u, v), np.stack([gx, gy], axis=1)
