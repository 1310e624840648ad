"""ctypes binding of libsbx.so (include/sbx.h).  No torch types cross this boundary.

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
soap_bubble_simulation_b200/csrc``).  There is no CPU fallback: if the shared object is missing,
importing this module raises, and every compute call fails loudly when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsbx.so")

# enums (include/sbx.h)
SBX_OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_COMM, ERR_IO = -1, -2, -3, -4, -5, -6
SEM_CPU, SEM_WGSL = 0, 1
PREC_F32, PREC_F64, PREC_F64_STRICT = 0, 1, 2
F32, F64 = 0, 1
RGB_NONE, RGB_LUT, RGB_EXACT3, RGB_EXACT3_LINEAR, RGB_SCHLICK_WALL, RGB_SCHLICK_TWO_BEAM, RGB_SPECTRAL = 0, 1, 2, 3, 4, 5, 6
OUT_RGBA8, OUT_RGB_F32 = 0, 1
PTR_THICKNESS, PTR_CONCENTRATION, PTR_RGB, PTR_CAUSTICS = 0, 1, 2, 3
IPC_BYTES = 512


class SbxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsbx error {code}: {msg}")
        self.code = code


class DrainageParamsC(C.Structure):
    """sbx_drainage_params == DrainageParams (src/render/gpu_drainage.rs:10-44), 64 bytes."""
    _fields_ = [("dt", C.c_float), ("gravity", C.c_float), ("viscosity", C.c_float),
                ("density", C.c_float), ("diffusion_coeff", C.c_float), ("bubble_radius", C.c_float),
                ("critical_thickness", C.c_float), ("grid_width", C.c_uint32),
                ("grid_height", C.c_uint32), ("marangoni_enabled", C.c_uint32),
                ("gamma_air", C.c_float), ("gamma_reduction", C.c_float),
                ("surfactant_diffusion", C.c_float), ("marangoni_coeff", C.c_float),
                ("_padding1", C.c_uint32), ("_padding2", C.c_uint32)]


class CausticParamsC(C.Structure):
    """sbx_caustic_params == CausticParams (src/render/caustics.rs:19-56), 64 bytes."""
    _fields_ = [("grid_width", C.c_uint32), ("grid_height", C.c_uint32), ("refractive_index", C.c_float),
                ("film_thickness_scale", C.c_float), ("light_dir_x", C.c_float), ("light_dir_y", C.c_float),
                ("light_dir_z", C.c_float), ("light_intensity", C.c_float), ("focal_length", C.c_float),
                ("caustic_intensity", C.c_float), ("caustic_sharpness", C.c_float), ("branch_threshold", C.c_float),
                ("ground_y", C.c_float), ("bubble_radius", C.c_float), ("time", C.c_float), ("_padding", C.c_uint32)]


class FrameParamsC(C.Structure):
    """sbx_frame_params: Camera (camera.rs:49-60) + the BubbleUniform fields bubble.wgsl::fs_main's shading reads."""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("cam_distance", C.c_float), ("cam_yaw", C.c_float),
                ("cam_pitch", C.c_float), ("cam_fov_rad", C.c_float), ("radius", C.c_float),
                ("base_thickness_nm", C.c_float), ("film_time", C.c_float), ("swirl_intensity", C.c_float),
                ("drainage_speed", C.c_float), ("pattern_scale", C.c_float), ("interference_intensity", C.c_float),
                ("base_alpha", C.c_float), ("edge_alpha", C.c_float), ("edge_smoothing_mode", C.c_uint32),
                ("background", C.c_float * 3), ("lut_max_nm", C.c_float)]


class DrainageCfgC(C.Structure):
    _fields_ = [("num_theta", C.c_int32), ("num_phi", C.c_int32), ("batch", C.c_int32),
                ("h0_m", C.c_double), ("density", C.c_double), ("gravity", C.c_double),
                ("viscosity", C.c_double), ("diffusion", C.c_double),
                ("critical_thickness_m", C.c_double), ("radius_m", C.c_double),
                ("semantics", C.c_int32), ("precision", C.c_int32), ("h_ref_m", C.c_double),
                ("tfe_enabled", C.c_int32), ("tfe_ratio", C.c_double),
                ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("batch_h0_m", C.POINTER(C.c_double)), ("batch_radius_m", C.POINTER(C.c_double)),
                ("wgsl", DrainageParamsC)]


class ViewC(C.Structure):
    _fields_ = [("cam_pos", C.c_double * 3), ("n_film", C.c_double), ("intensity", C.c_float),
                ("mode", C.c_int32), ("format", C.c_int32)]


class TfeFrontC(C.Structure):
    _fields_ = [("phi_center", C.c_double), ("extent", C.c_double), ("width", C.c_double),
                ("velocity", C.c_double)]


class FrameStatsC(C.Structure):
    _fields_ = [("min_thickness", C.c_double), ("max_thickness", C.c_double), ("thickness", C.c_double),
                ("time", C.c_double), ("has_critical", C.c_int32), ("_pad", C.c_int32)]


class TfeStateC(C.Structure):
    _fields_ = [("time", C.c_double), ("next_spawn", C.c_double), ("count", C.c_int32), ("_pad", C.c_int32),
                ("fronts", TfeFrontC * 8)]


_SIGS = {
    # name: (argtypes)   -- every function returns int32 except sbx_last_error
    "sbx_version": [C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "sbx_device_count": [C.POINTER(C.c_int32)],
    "sbx_host_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "sbx_host_free": [C.c_void_p],
    "sbx_drainage_cfg_default": [C.POINTER(DrainageCfgC)],
    "sbx_drainage_create": [C.POINTER(DrainageCfgC), C.c_int32, C.POINTER(C.c_void_p)],
    "sbx_drainage_destroy": [C.c_void_p],
    "sbx_drainage_set_stream": [C.c_void_p, C.c_void_p],
    "sbx_drainage_sync": [C.c_void_p],
    "sbx_drainage_upload": [C.c_void_p, C.c_void_p, C.c_int32],
    "sbx_drainage_download": [C.c_void_p, C.c_void_p, C.c_int32],
    "sbx_drainage_upload_concentration": [C.c_void_p, C.c_void_p],
    "sbx_drainage_download_concentration": [C.c_void_p, C.c_void_p],
    "sbx_drainage_step": [C.c_void_p, C.c_double, C.c_int32],
    "sbx_drainage_step_fused": [C.c_void_p, C.c_double, C.c_int32, C.POINTER(ViewC)],
    "sbx_drainage_shade": [C.c_void_p, C.POINTER(ViewC)],
    "sbx_drainage_step_host": [C.c_void_p, C.c_double, C.c_int32, C.POINTER(ViewC), C.c_void_p,
                               C.c_int32, C.c_void_p, C.c_int32, C.c_void_p],
    "sbx_drainage_download_rgb": [C.c_void_p, C.c_void_p],
    "sbx_drainage_get_thickness": [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_double)],
    "sbx_drainage_minmax": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "sbx_drainage_track_minmax": [C.c_void_p, C.c_int32],
    "sbx_drainage_has_critical": [C.c_void_p, C.POINTER(C.c_int32)],
    "sbx_drainage_critical_thickness": [C.c_void_p, C.POINTER(C.c_double)],
    "sbx_drainage_current_time": [C.c_void_p, C.POINTER(C.c_double)],
    "sbx_drainage_reset": [C.c_void_p, C.c_double],
    "sbx_drainage_set_marginal_regeneration": [C.c_void_p, C.c_int32],
    "sbx_drainage_marginal_regeneration": [C.c_void_p, C.POINTER(C.c_int32)],
    "sbx_drainage_set_tfe_ratio": [C.c_void_p, C.c_double],
    "sbx_drainage_tfe_ratio": [C.c_void_p, C.POINTER(C.c_double)],
    "sbx_drainage_active_tfe_count": [C.c_void_p, C.POINTER(C.c_int32)],
    "sbx_drainage_fronts": [C.c_void_p, C.POINTER(TfeFrontC), C.POINTER(C.c_int32)],
    "sbx_drainage_frame_stats": [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.POINTER(FrameStatsC)],
    "sbx_drainage_get_state": [C.c_void_p, C.POINTER(TfeStateC)],
    "sbx_drainage_set_state": [C.c_void_p, C.POINTER(TfeStateC)],
    "sbx_drainage_update_params": [C.c_void_p, C.POINTER(DrainageParamsC)],
    "sbx_drainage_params_get": [C.c_void_p, C.POINTER(DrainageParamsC)],
    "sbx_drainage_set_physics": [C.c_void_p] + [C.c_double] * 6,
    "sbx_drainage_device_ptr": [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)],
    "sbx_drainage_set_lut": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float],
    "sbx_drainage_set_spectrum": [C.c_void_p, C.POINTER(C.c_float), C.c_int32],
    "sbx_drainage_launch_count": [C.c_void_p, C.POINTER(C.c_int64)],
    "sbx_caustic_params_default": [C.POINTER(CausticParamsC)],
    "sbx_drainage_caustics": [C.c_void_p, C.POINTER(CausticParamsC), C.c_void_p],
    "sbx_drainage_sample_uv": [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p],
    "sbx_drainage_sample_normal": [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p],
    "sbx_frame_params_default": [C.POINTER(FrameParamsC)],
    "sbx_render_bubble_frame": [C.POINTER(FrameParamsC), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32],
    "sbx_drainage_render_frame": [C.c_void_p, C.c_int32, C.POINTER(FrameParamsC), C.c_double, C.c_void_p],
    "sbx_export_frame_png": [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t],
    "sbx_lut_build": [C.c_float, C.c_float, C.c_void_p, C.c_int32],
    "sbx_lut_build_spectral": [C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                               C.c_float, C.c_void_p, C.c_int32],
    "sbx_lut_shade_points": [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t,
                             C.c_float, C.c_void_p, C.c_int32],
    "sbx_interference_eval": [C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_size_t,
                              C.c_int32, C.c_void_p, C.c_int32],
    "sbx_interference_terms": [C.c_double, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_int32],
    "sbx_comm_unique_id": [C.c_void_p],
    "sbx_comm_init": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)],
    "sbx_comm_destroy": [C.c_void_p],
    "sbx_drainage_attach_comm": [C.c_void_p, C.c_void_p],
    "sbx_drainage_minmax_global": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "sbx_drainage_ipc_export": [C.c_void_p, C.c_void_p],
    "sbx_drainage_ipc_attach": [C.c_void_p, C.c_void_p, C.c_void_p],
    "sbx_drainage_peer_status": [C.c_void_p, C.POINTER(C.c_int32)],
}

EXPORTED_SYMBOLS = ["sbx_last_error"] + list(_SIGS)

_lib = None


def lib() -> C.CDLL:
    """Load libsbx.so once.  Raises if the extension has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C soap_bubble_simulation_b200/csrc`; this package has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.sbx_last_error.restype = C.c_char_p
        L.sbx_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = C.c_int32
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != SBX_OK:
        raise SbxError(rc, lib().sbx_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int32(0)
    check(lib().sbx_device_count(C.byref(n)))
    return n.value
