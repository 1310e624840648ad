"""GPUDrainageSimulator / DrainageParams -- mirror of src/render/gpu_drainage.rs over libsbx.

Same public fields and methods as the Rust struct (gpu_drainage.rs:81-491) minus the wgpu
device/queue/encoder arguments: the compute shader drainage.wgsl::drainage_step is replaced by the
sm_100a kernel k_wgsl (SBX_SEM_WGSL semantics: f32, thickness + concentration, Marangoni coupling,
poles frozen).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .drainage import _DeviceField


@dataclass
class DrainageParams:
    """DrainageParams (gpu_drainage.rs:10-68), defaults from ``impl Default``."""
    dt: float = 0.001
    gravity: float = 9.81
    viscosity: float = 0.001
    density: float = 1000.0
    diffusion_coeff: float = 1e-9
    bubble_radius: float = 0.025
    critical_thickness: float = 30e-9
    grid_width: int = 128
    grid_height: int = 64
    marangoni_enabled: int = 0
    gamma_air: float = 0.072
    gamma_reduction: float = 0.045
    surfactant_diffusion: float = 1e-9
    marangoni_coeff: float = 0.01

    def to_c(self) -> L.DrainageParamsC:
        p = L.DrainageParamsC()
        for k, v in self.__dict__.items():
            setattr(p, k, v)
        return p

    @classmethod
    def from_c(cls, p: L.DrainageParamsC) -> "DrainageParams":
        return cls(**{k: getattr(p, k) for k in cls.__dataclass_fields__})


class GPUDrainageSimulator(_DeviceField):
    """GPUDrainageSimulator (gpu_drainage.rs:81-491)."""

    def __init__(self, initial_thickness: float, grid_width: int, grid_height: int, *, device: int = 0,
                 batch: int = 1, row_begin: int = 0, row_end: int = 0):
        cfg = L.DrainageCfgC()
        L.check(L.lib().sbx_drainage_cfg_default(C.byref(cfg)))
        cfg.num_phi, cfg.num_theta, cfg.batch = int(grid_width), int(grid_height), batch
        cfg.h0_m = float(np.float32(initial_thickness))
        cfg.semantics, cfg.precision, cfg.h_ref_m = L.SEM_WGSL, L.PREC_F32, 0.0
        cfg.row_begin, cfg.row_end = row_begin, row_end
        self._params = DrainageParams(grid_width=int(grid_width), grid_height=int(grid_height))
        cfg.wgsl = self._params.to_c()
        super().__init__(cfg, device)
        # public fields, gpu_drainage.rs:100-107, 326-329
        self.enabled = False
        self.time_scale = 100.0
        self.steps_per_frame = 10
        self.marangoni_enabled = False

    def step(self, dt: float) -> None:
        """gpu_drainage.rs:338-370: if enabled, steps_per_frame dispatches with params.dt; the frame dt only
        advances current_time (by dt * time_scale in total)."""
        if not self.enabled:
            return
        scaled_dt = np.float32(dt) * np.float32(self.time_scale)
        step_dt = float(scaled_dt / np.float32(self.steps_per_frame))
        L.check(L.lib().sbx_drainage_step(self._h, step_dt, int(self.steps_per_frame)))
        self._host_stale = True

    def dispatch(self, nsteps: int = 1, view=None) -> None:
        """nsteps raw dispatches of drainage_step (no enabled/time_scale logic), optionally fused with colour."""
        if view is None:
            L.check(L.lib().sbx_drainage_step(self._h, float(self._params.dt), nsteps))
        else:
            vc = view.to_c()
            L.check(L.lib().sbx_drainage_step_fused(self._h, float(self._params.dt), nsteps, C.byref(vc)))
        self._host_stale = True

    def _push(self) -> None:
        pc = self._params.to_c()
        L.check(L.lib().sbx_drainage_update_params(self._h, C.byref(pc)))

    def update_params(self, params: DrainageParams) -> None:  # gpu_drainage.rs:377-380
        self._params = params
        self._push()

    def set_dt(self, dt: float) -> None:  # gpu_drainage.rs:383-386
        self._params.dt = dt
        self._push()

    def set_physics(self, viscosity: float, density: float, gravity: float, diffusion: float) -> None:  # :389-402
        self._params.viscosity, self._params.density = viscosity, density
        self._params.gravity, self._params.diffusion_coeff = gravity, diffusion
        self._push()

    def set_bubble_radius(self, radius: float) -> None:  # gpu_drainage.rs:405-408
        self._params.bubble_radius = radius
        self._push()

    def reset(self, initial_thickness: float) -> None:  # gpu_drainage.rs:415-432
        L.check(L.lib().sbx_drainage_reset(self._h, float(np.float32(initial_thickness))))
        self._host_stale = True

    def set_marangoni_enabled(self, enabled: bool) -> None:  # gpu_drainage.rs:435-440
        self.marangoni_enabled = bool(enabled)
        self._params.marangoni_enabled = 1 if enabled else 0
        self._push()

    def set_marangoni_params(self, gamma_air, gamma_reduction, surfactant_diffusion, marangoni_coeff) -> None:  # :443-456
        self._params.gamma_air, self._params.gamma_reduction = gamma_air, gamma_reduction
        self._params.surfactant_diffusion, self._params.marangoni_coeff = surfactant_diffusion, marangoni_coeff
        self._push()

    def concentration_buffer(self) -> int:  # gpu_drainage.rs:459-461 (device pointer; latest, see sbx.h)
        return self.device_ptr(L.PTR_CONCENTRATION)

    def current_time(self) -> float:  # gpu_drainage.rs:464-466
        out = C.c_double()
        L.check(L.lib().sbx_drainage_current_time(self._h, C.byref(out)))
        return out.value

    def current_buffer(self) -> int:  # gpu_drainage.rs:471-475 (device pointer; latest, see sbx.h)
        return self.device_ptr(L.PTR_THICKNESS)

    def current_thickness_buffer(self) -> int:  # gpu_drainage.rs:478-480
        return self.current_buffer()

    def grid_size(self):  # gpu_drainage.rs:483-485
        return (self._params.grid_width, self._params.grid_height)

    def params(self) -> DrainageParams:  # gpu_drainage.rs:488-490
        return self._params

    # field I/O helpers (the reference reads the wgpu buffers back through map_async)
    def upload_concentration(self, conc: np.ndarray) -> None:
        a = np.ascontiguousarray(conc, dtype=np.float32)
        L.check(L.lib().sbx_drainage_upload_concentration(self._h, a.ctypes.data))

    def download_concentration(self) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float32)
        L.check(L.lib().sbx_drainage_download_concentration(self._h, out.ctypes.data))
        return out
