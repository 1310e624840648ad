"""The host-side mirrors keep the reference's public names for the path (SURVEY.md section 8b "Rust signatures to
preserve"), so that a caller of the reference finds every method under the same name.  Name lists are transcribed from
the reference sources cited beside them; no GPU and no reference checkout needed."""
import inspect
import os
import re

import soap_bubble_simulation_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SURFACE = {
    # src/physics/drainage.rs:263-613
    S.DrainageSimulator: ["new", "step", "get_thickness", "thickness_field", "current_time", "has_critical_region",
                          "critical_thickness", "reset", "set_marginal_regeneration", "marginal_regeneration_enabled",
                          "set_tfe_thickness_ratio", "tfe_thickness_ratio", "active_tfe_count"],
    # src/physics/drainage.rs:121-203
    S.ThicknessField: ["get", "set", "theta_at", "phi_at", "num_theta", "num_phi", "delta_theta", "delta_phi",
                       "min_thickness", "max_thickness", "data", "data_mut"],
    # src/render/gpu_drainage.rs:118-491 (minus the wgpu device/queue/encoder arguments)
    S.GPUDrainageSimulator: ["step", "update_params", "set_dt", "set_physics", "set_bubble_radius", "reset",
                             "set_marangoni_enabled", "set_marangoni_params", "concentration_buffer", "current_buffer",
                             "current_thickness_buffer", "grid_size", "params", "current_time"],
    # src/physics/interference.rs:127-512
    S.InterferenceCalculator: ["new", "default_soap_film", "with_custom_wavelengths", "refractive_index_film",
                               "refractive_index_medium", "calculate_transmission_angle_cos",
                               "calculate_optical_path_difference", "calculate_phase_difference",
                               "calculate_fresnel_reflection", "calculate_reflected_intensity",
                               "calculate_interference_color", "calculate_interference_color_linear",
                               "calculate_interference_color_with_index", "generate_color_lookup_table",
                               "generate_2d_color_lookup_table"],
    # src/physics/interference.rs:43-95
    S.RgbColor: ["new", "to_f32_tuple", "to_u8_tuple", "to_f32_array_with_alpha"],
    # src/render/headless.rs:50-548 (bubble only)
    S.HeadlessRenderPipeline: ["new", "render_to_buffer", "size", "set_thickness", "set_refractive_index", "orbit_camera",
                               "zoom_camera", "set_camera_distance", "set_time"],
    # src/render/camera.rs:49-105
    S.Camera: ["position", "orbit", "zoom", "set_aspect"],
}


def test_python_mirrors_keep_the_reference_names():
    for cls, names in SURFACE.items():
        missing = [n for n in names if not callable(getattr(cls, n, None))]
        assert not missing, (cls.__name__, missing)
    # free functions and constants: interference_lut.rs:9-41, image_export.rs:42, gpu_drainage.rs:46-68
    assert callable(S.generate_interference_lut) and callable(S.export_frame)
    assert (S.LUT_THICKNESS_SAMPLES, S.LUT_ANGLE_SAMPLES, S.LUT_MAX_THICKNESS_NM) == (256, 64, 2000.0)
    p = S.DrainageParams()
    assert (p.dt, p.gravity, p.viscosity, p.density, p.grid_width, p.grid_height) == (0.001, 9.81, 0.001, 1000.0, 128, 64)
    assert abs(p.critical_thickness - 30e-9) < 1e-15 and (p.gamma_air, p.gamma_reduction) == (0.072, 0.045)
    # GPUDrainageSimulator's public fields (gpu_drainage.rs:95-104, defaults :325-330)
    src = inspect.getsource(S.GPUDrainageSimulator.__init__)
    for field in ("enabled", "time_scale", "steps_per_frame", "marangoni_enabled"):
        assert f"self.{field}" in src, field


def test_cxx_mirror_keeps_the_reference_names():
    with open(os.path.join(ROOT, "include", "sbx.hpp")) as f:
        hpp = f.read()
    for cls in ("DrainageSimulator", "ThicknessField", "GPUDrainageSimulator", "InterferenceCalculator", "RgbColor",
                "HeadlessRenderPipeline", "Camera"):
        assert re.search(rf"\b(class|struct) {cls}\b", hpp), cls
    wanted = SURFACE[S.DrainageSimulator][1:] + SURFACE[S.ThicknessField] + \
        ["set_dt", "set_physics", "set_bubble_radius", "set_marangoni_enabled", "set_marangoni_params", "grid_size",
         "calculate_transmission_angle_cos", "calculate_reflected_intensity", "calculate_interference_color",
         "generate_interference_lut", "render_to_buffer", "set_camera_distance", "orbit_camera", "export_frame"]
    missing = [n for n in wanted if not re.search(rf"\b{n}\s*\(", hpp)]
    assert not missing, missing
