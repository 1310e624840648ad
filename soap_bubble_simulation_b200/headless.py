"""Host-side mirror of the reference's headless bubble renderer and PNG export over the C ABI.

``HeadlessRenderPipeline`` keeps the public surface of ``src/render/headless.rs:27-552`` (``new(width, height)``,
``render_to_buffer``, ``size``, ``set_thickness``, ``set_refractive_index``, ``orbit_camera``, ``zoom_camera``,
``set_camera_distance``, ``set_time``, public ``camera`` / ``bubble_uniform``); the frame is produced by
``sbx_render_bubble_frame`` (the fragment shader's shading path as a CUDA kernel, ``csrc/sbx_frame.cu``) instead of
wgpu.  ``export_frame`` is ``src/export/image_export.rs:42-76``.  No CPU fallback: without a CUDA device
``render_to_buffer`` raises ``SbxError``.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib as L
from .interference import generate_interference_lut


class ExportError(Exception):
    """ExportError (image_export.rs:6-28): ``kind`` is 'InvalidDimensions', 'BufferCreation' or 'SaveError'."""

    def __init__(self, kind: str, msg: str):
        super().__init__(msg)
        self.kind = kind


class Camera:
    """Orbit camera, src/render/camera.rs:28-105 (f32 state like the reference)."""

    def __init__(self, aspect: float):
        self.distance = 0.2
        self.yaw = 0.0
        self.pitch = 0.3
        self.fov = math.radians(45.0)
        self.aspect = float(aspect)
        self.near, self.far = 0.001, 100.0

    def position(self):
        d, y, p = np.float32(self.distance), np.float32(self.yaw), np.float32(self.pitch)
        return (float(d * np.cos(p) * np.sin(y)), float(d * np.sin(p)), float(d * np.cos(p) * np.cos(y)))

    def orbit(self, delta_x: float, delta_y: float) -> None:  # camera.rs:71-78
        self.yaw = float(np.float32(self.yaw) + np.float32(delta_x) * np.float32(0.01))
        self.pitch = float(np.float32(self.pitch) + np.float32(delta_y) * np.float32(0.01))
        max_pitch = float(np.float32(math.pi / 2.0) - np.float32(0.01))
        self.pitch = min(max(self.pitch, -max_pitch), max_pitch)

    def zoom(self, delta: float) -> None:  # camera.rs:81-84
        self.distance = float(np.float32(self.distance) * (np.float32(1.0) - np.float32(delta) * np.float32(0.1)))
        self.distance = min(max(self.distance, 0.05), 10.0)

    def set_aspect(self, aspect: float) -> None:
        self.aspect = float(aspect)


class BubbleUniform:
    """The BubbleUniform fields the shading path reads, with BubbleUniform::default (pipeline.rs:69-112)."""

    def __init__(self):
        self.refractive_index = 1.33
        self.base_thickness_nm = 500.0
        self.time = 0.0
        self.interference_intensity = 4.0
        self.base_alpha = 0.3
        self.edge_alpha = 0.6
        self.background_r, self.background_g, self.background_b = 0.1, 0.1, 0.15
        self.film_time = 0.0
        self.swirl_intensity = 1.0
        self.drainage_speed = 0.5
        self.pattern_scale = 1.0
        self.edge_smoothing_mode = 1
        self.branched_flow_enabled = 0


def frame_params(width: int, height: int, camera: Camera, bu: BubbleUniform, radius: float = 0.025) -> L.FrameParamsC:
    p = L.FrameParamsC()
    L.check(L.lib().sbx_frame_params_default(C.byref(p)))
    p.width, p.height = int(width), int(height)
    p.cam_distance, p.cam_yaw, p.cam_pitch, p.cam_fov_rad = camera.distance, camera.yaw, camera.pitch, camera.fov
    p.radius = radius
    p.base_thickness_nm = bu.base_thickness_nm
    p.film_time, p.swirl_intensity = bu.film_time, bu.swirl_intensity
    p.drainage_speed, p.pattern_scale = bu.drainage_speed, bu.pattern_scale
    p.interference_intensity = bu.interference_intensity
    p.base_alpha, p.edge_alpha = bu.base_alpha, bu.edge_alpha
    p.edge_smoothing_mode = int(bu.edge_smoothing_mode)
    p.background[0], p.background[1], p.background[2] = bu.background_r, bu.background_g, bu.background_b
    return p


class HeadlessRenderPipeline:
    """HeadlessRenderPipeline (headless.rs:27-552) for the bubble.  As in the reference, the interference LUT is
    built once at construction for the default refractive index (headless.rs:143-146); ``set_refractive_index``
    only changes the uniform."""

    def __init__(self, width: int, height: int, device: int = 0):
        if width <= 0 or height <= 0:
            raise ValueError(f"Invalid dimensions: {width}x{height}")
        self.width, self.height, self.device = int(width), int(height), int(device)
        self.camera = Camera(width / height)
        self.bubble_uniform = BubbleUniform()
        self._lut = np.ascontiguousarray(generate_interference_lut(self.bubble_uniform.refractive_index, 1.0, device=device)
                                         .reshape(64, 256, 4))

    @classmethod
    def new(cls, width: int, height: int, device: int = 0) -> "HeadlessRenderPipeline":
        return cls(width, height, device)

    def size(self):
        return (self.width, self.height)

    def set_thickness(self, thickness_nm: float) -> None:
        self.bubble_uniform.base_thickness_nm = float(thickness_nm)

    def set_refractive_index(self, n: float) -> None:
        self.bubble_uniform.refractive_index = float(n)

    def orbit_camera(self, delta_x: float, delta_y: float) -> None:
        self.camera.orbit(delta_x, delta_y)

    def zoom_camera(self, delta: float) -> None:
        self.camera.zoom(delta)

    def set_camera_distance(self, distance: float) -> None:
        self.camera.distance = min(max(float(distance), 0.05), 10.0)

    def set_time(self, time: float) -> None:
        self.bubble_uniform.time = float(time)
        self.bubble_uniform.film_time = float(time)

    def render_to_buffer(self) -> np.ndarray:
        """(height, width, 4) uint8, row 0 at the top (the reference returns the same bytes as a flat Vec<u8>)."""
        p = frame_params(self.width, self.height, self.camera, self.bubble_uniform)
        out = np.empty((self.height, self.width, 4), np.uint8)
        L.check(L.lib().sbx_render_bubble_frame(C.byref(p), self._lut.ctypes.data, 256, 64, out.ctypes.data, self.device))
        return out


def render_drainage_frame(sim, width: int = 256, height: int = 256, camera: Camera | None = None,
                          bubble_uniform: BubbleUniform | None = None, bubble: int = 0, radius: float = 0.025) -> np.ndarray:
    """The same frame with the film thickness read from a DrainageSimulator's device field
    (``sbx_drainage_render_frame``) instead of the shader's procedural noise."""
    cam = camera or Camera(width / height)
    bu = bubble_uniform or BubbleUniform()
    p = frame_params(width, height, cam, bu, radius=radius)
    out = np.empty((height, width, 4), np.uint8)
    L.check(L.lib().sbx_drainage_render_frame(sim._h, int(bubble), C.byref(p), float(bu.refractive_index), out.ctypes.data))
    return out


def export_frame(path, width: int, height: int, data) -> None:
    """export_frame (image_export.rs:42-76): RGBA8 bytes -> PNG file; raises ExportError like the reference returns it."""
    if width == 0 or height == 0:
        raise ExportError("InvalidDimensions", f"Invalid dimensions: {width}x{height}")
    buf = np.ascontiguousarray(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray, memoryview))
                               else np.asarray(data, np.uint8)).reshape(-1)
    expected = int(width) * int(height) * 4
    if buf.size != expected:
        raise ExportError("BufferCreation", f"Failed to create image buffer: Data length {buf.size} doesn't match "
                                            f"expected {expected} ({width}x{height}x4)")
    rc = L.lib().sbx_export_frame_png(os.fsencode(os.fspath(path)), int(width), int(height), buf.ctypes.data, buf.size)
    if rc == L.ERR_IO:
        raise ExportError("SaveError", L.lib().sbx_last_error().decode("utf-8", "replace"))
    L.check(rc)
