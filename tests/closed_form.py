"""Known answers written down from the reference's formulas alone (no restatement is run to produce them).

On a field of the form  a + eps * theta_i * cos(m phi_j)  the reference's finite differences have closed forms:

  second theta difference of a linear function            = 0
  central first theta difference                          = eps cos(m phi_j)
  periodic second phi difference of cos(m phi_j)          = cos(m phi_j) (2 cos(m dphi) - 2) / dphi^2
  central first phi difference of cos(m phi_j)            = -sin(m phi_j) sin(m dphi) / dphi

so one Euler step of DrainageSimulator::step (src/physics/drainage.rs:346-378, poles :480-503) and of
drainage.wgsl::drainage_step (src/render/shaders/drainage.wgsl:72-200) can be evaluated directly.  Used by
tests/test_oracle_reference_tests.py (the oracle) and tests/test_gpu_closed_form.py (the CUDA path).
"""
import math

import numpy as np


def _shape(theta_in, phi, m, dph):
    """R^2 x surface Laplacian of theta cos(m phi) on the interior rows (drainage.rs:352-367)."""
    lam = (2.0 * math.cos(m * dph) - 2.0) / (dph * dph)
    cm = np.cos(m * phi)[None, :]
    return cm * (np.cos(theta_in) / np.sin(theta_in) + theta_in * lam / np.sin(theta_in) ** 2)


def cpu_mode(nt, nphi, m, a, eps, dt, density, gravity, viscosity, diffusion, radius_m, storage=np.float64):
    """-> (field as stored, increments of the interior rows, new top pole value, new bottom pole value), all f64
    except the field, which is cast to `storage` first (the gravity term uses the stored values)."""
    dth, dph = math.pi / (nt - 1), 2.0 * math.pi / nphi
    theta, phi = np.arange(nt) * dth, np.arange(nphi) * dph
    h = (a + eps * theta[:, None] * np.cos(m * phi)[None, :]).astype(storage)
    th = theta[1:-1, None]
    hh = h[1:-1].astype(np.float64)
    c = density * gravity / (3.0 * viscosity)
    inc = dt * (-c * hh ** 3 * np.sin(th) + diffusion * eps * _shape(th, phi, m, dph) / (radius_m * radius_m))
    new = hh + inc
    return h, inc, float(new[0].mean()), float(new[-1].mean())


def wgsl_mode(H, W, m, a, eps, c0, delta, p):
    """p: the shader's DrainageParams (attributes dt, gravity, ...).  -> (h f32, c f32, thickness increments,
    concentration increments) with the increments in f64 on the interior rows; pole rows are copied by the shader."""
    dth, dph = math.pi / (H - 1), 2.0 * math.pi / W
    theta, phi = np.arange(H) * dth, np.arange(W) * dph
    mode = theta[:, None] * np.cos(m * phi)[None, :]
    h, c = (a + eps * mode).astype(np.float32), (c0 + delta * mode).astype(np.float32)
    R, dt = float(p.bubble_radius), float(p.dt)
    th, cm, sm = theta[1:-1, None], np.cos(m * phi)[None, :], np.sin(m * phi)[None, :]
    shape = _shape(th, phi, m, dph)
    hh, cc = h[1:-1].astype(np.float64), c[1:-1].astype(np.float64)
    grav = -(float(p.density) * float(p.gravity) / (3.0 * float(p.viscosity))) * hh ** 3 * np.sin(th)
    mar = 0.0
    if p.marangoni_enabled:
        grad = np.sqrt((delta * cm / R) ** 2 + (-delta * th * sm * math.sin(m * dph) / dph / (R * np.sin(th))) ** 2)
        mar = float(p.marangoni_coeff) * float(p.gamma_reduction) * grad * hh * hh
    inc_h = dt * (grav + float(p.diffusion_coeff) * eps * shape / (R * R) + mar)
    inc_c = dt * float(p.surfactant_diffusion) * delta * shape / (R * R)
    if p.marangoni_enabled:
        inc_c = inc_c + dt * 0.001 * (eps * shape) * cc
    return h, c, inc_h, inc_c, dt * grav, dt * mar
