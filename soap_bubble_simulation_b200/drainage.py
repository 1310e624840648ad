"""DrainageSimulator / ThicknessField -- host-side mirror of src/physics/drainage.rs over libsbx.

Same names, argument meaning and error behaviour as the Rust API (drainage.rs:96-204, 212-614); the
thickness field lives in HBM and every step runs the hand-written sm_100a kernels behind the C ABI
(include/sbx.h).  ``thickness_field()`` returns a host view that is refreshed from the device on
demand, which is what a Rust shim returning ``&ThicknessField`` has to do as well.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field as _field
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .config import SimulationConfig


@dataclass
class View:
    """Camera + optics of the colour stage (sbx_view).  Defaults: Camera::new (src/render/camera.rs:49-68:
    distance 0.2, yaw 0, pitch 0.3), refractive index 1.33, interference_intensity 4.0 (pipeline.rs:75)."""
    cam_pos: Sequence[float] = _field(default_factory=lambda: camera_position(0.2, 0.0, 0.3))
    n_film: float = 1.33
    intensity: float = 4.0
    mode: int = L.RGB_LUT
    format: int = L.OUT_RGBA8

    def to_c(self) -> L.ViewC:
        v = L.ViewC()
        for k in range(3):
            v.cam_pos[k] = float(self.cam_pos[k])
        v.n_film, v.intensity, v.mode, v.format = self.n_film, self.intensity, self.mode, self.format
        return v


def camera_position(distance: float, yaw: float, pitch: float):
    """Camera::position (src/render/camera.rs:63-68) for target = origin; f32 like the reference."""
    d, y, p = np.float32(distance), np.float32(yaw), np.float32(pitch)
    return (float(d * np.cos(p) * np.sin(y)), float(d * np.sin(p)), float(d * np.cos(p) * np.cos(y)))


class ThicknessField:
    """ThicknessField (drainage.rs:96-204).  Standalone it is a plain host container; obtained from
    ``DrainageSimulator.thickness_field()`` it is a view of the device field (reductions run on the GPU,
    ``data()`` downloads, ``set``/``data_mut`` edits are uploaded before the next step).  The reference has one
    bubble per simulator; on a batched (foam) handle the view shows bubble 0 and edits to it are written back
    into the full batch."""

    def __init__(self, num_theta: int, num_phi: int, initial_thickness: float, _sim=None):
        self._num_theta, self._num_phi = int(num_theta), int(num_phi)
        self._delta_theta = math.pi / (num_theta - 1)   # drainage.rs:123
        self._delta_phi = 2.0 * math.pi / num_phi       # drainage.rs:124
        self._sim = _sim
        self._host: Optional[np.ndarray] = None
        self._dirty = False
        if _sim is None:
            self._host = np.full((num_theta, num_phi), initial_thickness, dtype=np.float64)

    # -- synchronisation with the device (attached fields only)
    def _refresh(self) -> np.ndarray:
        if self._sim is not None and (self._host is None or self._sim._host_stale) and not self._dirty:
            self._full = self._sim.download(np.float64)  # (batch, rows, phi)
            self._host = self._full[0]                   # a view: edits land in the full array
            self._sim._host_stale = False
        return self._host

    def _flush(self) -> None:
        if self._sim is not None and self._dirty:
            self._dirty = False
            self._sim.upload(self._full)

    def get(self, theta_index: int, phi_index: int) -> float:  # drainage.rs:137-139
        return float(self._refresh()[theta_index, phi_index])

    def set(self, theta_index: int, phi_index: int, value: float) -> None:  # drainage.rs:143-145
        self._refresh()[theta_index, phi_index] = value
        self._dirty = self._sim is not None

    def theta_at(self, theta_index: int) -> float:  # drainage.rs:149-151
        return theta_index * self._delta_theta

    def phi_at(self, phi_index: int) -> float:  # drainage.rs:155-157
        return phi_index * self._delta_phi

    def num_theta(self) -> int:
        return self._num_theta

    def num_phi(self) -> int:
        return self._num_phi

    def delta_theta(self) -> float:
        return self._delta_theta

    def delta_phi(self) -> float:
        return self._delta_phi

    def min_thickness(self) -> float:  # drainage.rs:180-185
        if self._sim is not None:
            self._flush()
            return self._sim.minmax()[0]
        return float(self._host.min())

    def max_thickness(self) -> float:  # drainage.rs:188-193
        if self._sim is not None:
            self._flush()
            return self._sim.minmax()[1]
        return float(self._host.max())

    def data(self) -> np.ndarray:  # drainage.rs:196-198 (row-major [theta][phi], metres)
        return self._refresh()

    def data_mut(self) -> np.ndarray:  # drainage.rs:201-203
        a = self._refresh()
        self._dirty = self._sim is not None
        return a


class _DeviceField:
    """Shared handle plumbing for the two simulator mirrors."""

    # every path that changes the device field sets _host_stale = True; the per-frame statistics cache follows it
    @property
    def _host_stale(self):
        return self.__dict__.get("_host_stale_flag", True)

    @_host_stale.setter
    def _host_stale(self, v):
        self.__dict__["_host_stale_flag"] = v
        if v:
            self.__dict__["_host_stale_stats"] = True

    def __init__(self, cfg: L.DrainageCfgC, device: int, keep=()):
        self._keep = keep  # keeps numpy arrays behind cfg pointers alive during create
        self._h = C.c_void_p()
        L.check(L.lib().sbx_drainage_create(C.byref(cfg), device, C.byref(self._h)))
        self.num_theta, self.num_phi, self.batch = cfg.num_theta, cfg.num_phi, cfg.batch
        rb, re = cfg.row_begin, cfg.row_end
        if rb == 0 and re == 0:
            re = cfg.num_theta
        self.row_begin, self.row_end = rb, re
        self.rows_local = re - rb
        self.device = device
        self._host_stale = True
        self._view_c = None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            L.lib().sbx_drainage_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- field I/O (shape (batch, rows_local, num_phi))
    def _shape(self):
        return (self.batch, self.rows_local, self.num_phi)

    def upload(self, field: np.ndarray) -> None:
        a = np.ascontiguousarray(field)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        if a.size != self.batch * self.rows_local * self.num_phi:
            raise ValueError(f"field has {a.size} cells, expected {self._shape()}")
        L.check(L.lib().sbx_drainage_upload(self._h, a.ctypes.data, L.F64 if a.dtype == np.float64 else L.F32))
        self._host_stale = True
        if getattr(self, "_field", None) is not None:
            self._field._dirty = False  # an explicit upload replaces pending ThicknessField edits

    def download(self, dtype=np.float64) -> np.ndarray:
        out = np.empty(self._shape(), dtype=dtype)
        L.check(L.lib().sbx_drainage_download(self._h, out.ctypes.data, L.F64 if out.dtype == np.float64 else L.F32))
        return out

    def download_rgb(self, fmt: int = L.OUT_RGBA8) -> np.ndarray:
        if fmt == L.OUT_RGBA8:
            out = np.empty(self._shape() + (4,), dtype=np.uint8)
        else:
            out = np.empty(self._shape() + (3,), dtype=np.float32)
        L.check(L.lib().sbx_drainage_download_rgb(self._h, out.ctypes.data))
        return out

    def minmax(self):
        mn, mx = C.c_double(), C.c_double()
        L.check(L.lib().sbx_drainage_minmax(self._h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def frame_stats(self, theta: float, phi: float, bubble: int = 0) -> dict:
        """min_thickness, max_thickness, get_thickness(theta, phi), has_critical_region and current_time -- what
        RenderPipeline::update reads after every step (pipeline.rs:1298-1312) -- behind one synchronisation."""
        st = L.FrameStatsC()
        L.check(L.lib().sbx_drainage_frame_stats(self._h, bubble, theta, phi, C.byref(st)))
        return {"min_thickness": st.min_thickness, "max_thickness": st.max_thickness, "thickness": st.thickness,
                "time": st.time, "has_critical": bool(st.has_critical)}

    def track_minmax(self, mode: int) -> None:
        """0 off, 1 on, 2 auto (default): min/max as a by-product of the step kernel instead of a pass over the field."""
        L.check(L.lib().sbx_drainage_track_minmax(self._h, int(mode)))

    def sync(self) -> None:
        L.check(L.lib().sbx_drainage_sync(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        L.check(L.lib().sbx_drainage_set_stream(self._h, C.c_void_p(cuda_stream)))

    def device_ptr(self, kind: int) -> int:
        p = C.c_void_p()
        L.check(L.lib().sbx_drainage_device_ptr(self._h, kind, C.byref(p)))
        if kind == L.PTR_THICKNESS:
            self._host_stale = True  # the caller may write through the pointer
        return p.value

    def launch_count(self) -> int:
        n = C.c_int64()
        L.check(L.lib().sbx_drainage_launch_count(self._h, C.byref(n)))
        return n.value

    def set_lut(self, lut_rgba: np.ndarray, max_nm: float = 2000.0) -> None:
        a = np.ascontiguousarray(lut_rgba, dtype=np.uint8)
        A, T = a.shape[0], a.shape[1]
        L.check(L.lib().sbx_drainage_set_lut(self._h, a.ctypes.data, T, A, max_nm))

    def _flush_field(self) -> None:
        f = getattr(self, "_field", None)
        if f is not None:
            f._flush()

    def shade(self, view: View) -> None:
        self._flush_field()
        vc = view.to_c()
        L.check(L.lib().sbx_drainage_shade(self._h, C.byref(vc)))

    def set_spectrum(self, wavelengths_nm) -> None:
        """Wavelength list of RGB_SPECTRAL (the reference's shaders use 400, 450, ..., 700 nm)."""
        wl = np.ascontiguousarray(wavelengths_nm, dtype=np.float32)
        L.check(L.lib().sbx_drainage_set_spectrum(self._h, wl.ctypes.data_as(C.POINTER(C.c_float)), len(wl)))

    def shade_spectral(self, view: View, wavelengths_nm=None) -> None:
        """Colour frame of the current field by direct per-cell spectral evaluation (no LUT):
        thin_film_interference of bubble_instanced.wgsl:190-218 over the given wavelengths."""
        if wavelengths_nm is not None:
            self.set_spectrum(wavelengths_nm)
        v = View(cam_pos=view.cam_pos, n_film=view.n_film, intensity=view.intensity, mode=L.RGB_SPECTRAL, format=view.format)
        self.shade(v)

    def step_fused(self, dt: float, view: View, nsteps: int = 1) -> None:
        vc = view.to_c()
        L.check(L.lib().sbx_drainage_step_fused(self._h, dt, nsteps, C.byref(vc)))
        self._host_stale = True

    def step_host(self, dt: float, nsteps: int, view: Optional[View], host_in: Optional[np.ndarray],
                  host_out_h: Optional[np.ndarray], host_out_rgb: Optional[np.ndarray]) -> None:
        """Host-buffer round trip (sbx_drainage_step_host): upload, nsteps fused steps, download."""
        if host_in is None:
            self._flush_field()  # pending ThicknessField edits are the input
        elif getattr(self, "_field", None) is not None:
            self._field._dirty = False  # the uploaded field replaces them
        vc = view.to_c() if view is not None else None

        def dt_of(a):
            return L.F64 if a is not None and a.dtype == np.float64 else L.F32

        L.check(L.lib().sbx_drainage_step_host(
            self._h, dt, nsteps, C.byref(vc) if vc is not None else None,
            host_in.ctypes.data if host_in is not None else None, dt_of(host_in),
            host_out_h.ctypes.data if host_out_h is not None else None, dt_of(host_out_h),
            host_out_rgb.ctypes.data if host_out_rgb is not None else None))
        self._host_stale = True

    def caustic_map(self, **params) -> np.ndarray:
        """CausticRenderer::compute (src/render/caustics.rs:441-457, caustics_compute.wgsl:114-126) on the latest
        thickness buffer; keyword arguments override CausticParams::default (caustics.rs:58-79).  Returns the
        (batch, rows, num_phi) float32 caustic intensity map."""
        p = L.CausticParamsC()
        L.check(L.lib().sbx_caustic_params_default(C.byref(p)))
        for k, v in params.items():
            setattr(p, k, v)
        out = np.empty(self._shape(), dtype=np.float32)
        L.check(L.lib().sbx_drainage_caustics(self._h, C.byref(p), out.ctypes.data))
        return out

    def sample_uv(self, uv, bubble: int = 0):
        """sample_thickness_at_uv / thickness_gradient_uv (branched_flow_compute.wgsl:187-210) at (n, 2) uv positions:
        returns (thickness (n,) f32 in metres, gradient (n, 2) f32 per unit uv)."""
        q = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        hh = np.empty(len(q), np.float32)
        gg = np.empty((len(q), 2), np.float32)
        L.check(L.lib().sbx_drainage_sample_uv(self._h, bubble, q.ctypes.data, len(q), hh.ctypes.data, gg.ctypes.data))
        return hh, gg

    def sample_normal(self, normals, scale: float = 1.0, bubble: int = 0):
        """sample_thickness(normal) (branched_flow_compute.wgsl:110-116, 294-300): thickness * scale at the texel each
        of the (n, 3) unit normals points at."""
        q = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        hh = np.empty(len(q), np.float32)
        L.check(L.lib().sbx_drainage_sample_normal(self._h, bubble, q.ctypes.data, len(q), float(scale), hh.ctypes.data))
        return hh

    def attach_comm(self, comm) -> None:
        L.check(L.lib().sbx_drainage_attach_comm(self._h, comm._c))
        self._comm = comm

    def minmax_global(self):
        mn, mx = C.c_double(), C.c_double()
        L.check(L.lib().sbx_drainage_minmax_global(self._h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value


class DrainageSimulator(_DeviceField):
    """DrainageSimulator (drainage.rs:212-614).

    ``precision``: ``PREC_F32`` (4 B/cell; the field is stored relative to ``h_ref`` -- by default the
    initial uniform thickness -- so that the sub-ulp increments of the default configuration are kept),
    ``PREC_F64`` or ``PREC_F64_STRICT`` (bit-identical to the f64 CPU path).
    """

    def __init__(self, config: Optional[SimulationConfig] = None, *, precision: int = L.PREC_F32,
                 device: int = 0, h_ref: Optional[float] = None, num_theta: Optional[int] = None,
                 num_phi: Optional[int] = None, batch: int = 1, batch_h0_m=None, batch_radius_m=None,
                 row_begin: int = 0, row_end: int = 0):
        config = config or SimulationConfig.default()
        self.config = config
        cfg = L.DrainageCfgC()
        L.check(L.lib().sbx_drainage_cfg_default(C.byref(cfg)))
        cfg.num_theta = int(num_theta if num_theta is not None else config.resolution)      # drainage.rs:264
        cfg.num_phi = int(num_phi if num_phi is not None else config.resolution * 2)        # drainage.rs:265
        cfg.batch = batch
        cfg.h0_m = config.film_thickness_meters()                                            # drainage.rs:266
        cfg.density = config.fluid.density
        cfg.gravity = config.environment.gravity
        cfg.viscosity = config.fluid.viscosity
        cfg.diffusion = config.fluid.surfactant_diffusion
        cfg.critical_thickness_m = config.bubble.critical_thickness_nm * 1e-9               # drainage.rs:278
        cfg.radius_m = config.bubble_radius()
        cfg.semantics = L.SEM_CPU
        cfg.precision = precision
        cfg.h_ref_m = cfg.h0_m if h_ref is None else h_ref
        cfg.row_begin, cfg.row_end = row_begin, row_end
        keep = []
        if batch_h0_m is not None:
            a = np.ascontiguousarray(batch_h0_m, dtype=np.float64)
            cfg.batch_h0_m = a.ctypes.data_as(C.POINTER(C.c_double))
            keep.append(a)
        if batch_radius_m is not None:
            a = np.ascontiguousarray(batch_radius_m, dtype=np.float64)
            cfg.batch_radius_m = a.ctypes.data_as(C.POINTER(C.c_double))
            keep.append(a)
        super().__init__(cfg, device, keep)
        self.precision = precision
        self._field = ThicknessField(cfg.num_theta, cfg.num_phi, cfg.h0_m, _sim=self)

    @classmethod
    def new(cls, config: SimulationConfig, **kw) -> "DrainageSimulator":  # drainage.rs:263
        return cls(config, **kw)

    # The reference's frame loop calls step, thickness_field().min_thickness(), .max_thickness(), get_thickness(pi/2, 0)
    # and has_critical_region() one after the other (pipeline.rs:1289-1322).  The mirror keeps those separate methods
    # but serves them from ONE sbx_drainage_frame_stats call: the first query after the field changed fetches all of
    # them, with the (theta, phi) the caller asked for last frame.
    def _frame(self, theta=None, phi=None, bubble=0):
        self._field._flush()
        c = getattr(self, "_frame_cache", None)
        if self._host_stale_stats or c is None or (theta is not None and c["args"] != (theta, phi, bubble)):
            if theta is None:
                theta, phi, bubble = getattr(self, "_last_query", (math.pi / 2.0, 0.0, 0))
                if not (self.row_begin * math.pi / (self.num_theta - 1) <= theta <= (self.row_end - 2) * math.pi / (self.num_theta - 1)):
                    theta = (self.row_begin + 0.5 * (self.rows_local - 1)) * math.pi / (self.num_theta - 1)  # a row this band owns
            c = self.frame_stats(theta, phi, bubble)
            c["args"] = (theta, phi, bubble)
            self._frame_cache = c
            self._host_stale_stats = False
        return c

    def minmax(self):
        c = self._frame()
        return c["min_thickness"], c["max_thickness"]

    def step(self, dt: float, nsteps: int = 1) -> None:  # drainage.rs:298-392
        self._field._flush()
        L.check(L.lib().sbx_drainage_step(self._h, dt, nsteps))
        self._host_stale = True

    def step_fused(self, dt: float, view: View, nsteps: int = 1) -> None:
        self._field._flush()
        super().step_fused(dt, view, nsteps)

    def get_thickness(self, theta: float, phi: float, bubble: int = 0) -> float:  # drainage.rs:516-552
        self._last_query = (theta, phi, bubble)
        return self._frame(theta, phi, bubble)["thickness"]

    def thickness_field(self) -> ThicknessField:  # drainage.rs:555-557
        return self._field

    def current_time(self) -> float:  # drainage.rs:560-562
        out = C.c_double()
        L.check(L.lib().sbx_drainage_current_time(self._h, C.byref(out)))
        return out.value

    def has_critical_region(self) -> bool:  # drainage.rs:565-567
        return self._frame()["has_critical"]

    def critical_thickness(self) -> float:  # drainage.rs:570-572
        out = C.c_double()
        L.check(L.lib().sbx_drainage_critical_thickness(self._h, C.byref(out)))
        return out.value

    def reset(self, initial_thickness: float) -> None:  # drainage.rs:575-583
        L.check(L.lib().sbx_drainage_reset(self._h, initial_thickness))
        self._field._dirty = False
        self._host_stale = True

    def set_marginal_regeneration(self, enabled: bool) -> None:  # drainage.rs:588-593
        L.check(L.lib().sbx_drainage_set_marginal_regeneration(self._h, int(bool(enabled))))

    def marginal_regeneration_enabled(self) -> bool:  # drainage.rs:596-598
        out = C.c_int32()
        L.check(L.lib().sbx_drainage_marginal_regeneration(self._h, C.byref(out)))
        return bool(out.value)

    def set_tfe_thickness_ratio(self, ratio: float) -> None:  # drainage.rs:601-603
        L.check(L.lib().sbx_drainage_set_tfe_ratio(self._h, ratio))

    def tfe_thickness_ratio(self) -> float:  # drainage.rs:606-608
        out = C.c_double()
        L.check(L.lib().sbx_drainage_tfe_ratio(self._h, C.byref(out)))
        return out.value

    def active_tfe_count(self) -> int:  # drainage.rs:611-613
        out = C.c_int32()
        L.check(L.lib().sbx_drainage_active_tfe_count(self._h, C.byref(out)))
        return out.value

    def fronts(self):
        buf = (L.TfeFrontC * 8)()
        n = C.c_int32()
        L.check(L.lib().sbx_drainage_fronts(self._h, buf, C.byref(n)))
        return [(buf[k].phi_center, buf[k].extent, buf[k].width, buf[k].velocity) for k in range(n.value)]

    def get_state(self):
        """(current_time, next_tfe_spawn_time, fronts) -- the host state machine beside the field (drainage.rs:237-255)."""
        st = L.TfeStateC()
        L.check(L.lib().sbx_drainage_get_state(self._h, C.byref(st)))
        fr = [(st.fronts[k].phi_center, st.fronts[k].extent, st.fronts[k].width, st.fronts[k].velocity)
              for k in range(st.count)]
        return st.time, st.next_spawn, fr

    def set_state(self, time: float, next_spawn: float, fronts) -> None:
        """Restore a checkpoint taken with get_state (field: download / upload)."""
        st = L.TfeStateC()
        st.time, st.next_spawn, st.count = float(time), float(next_spawn), len(fronts)
        if len(fronts) > 8:
            raise ValueError("at most 8 fronts (drainage.rs:474)")
        for k, f in enumerate(fronts):
            st.fronts[k].phi_center, st.fronts[k].extent, st.fronts[k].width, st.fronts[k].velocity = map(float, f)
        L.check(L.lib().sbx_drainage_set_state(self._h, C.byref(st)))
        self._host_stale = True

    def set_physics(self, density, gravity, viscosity, diffusion, critical_thickness_m, radius_m) -> None:
        L.check(L.lib().sbx_drainage_set_physics(self._h, density, gravity, viscosity, diffusion,
                                                 critical_thickness_m, radius_m))
        self._host_stale_stats = True  # has_critical_region depends on the critical thickness
