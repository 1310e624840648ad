//! Raw FFI of libsbx (include/sbx.h).  Source only: this image has no Rust toolchain, so this file is not
//! compiled here; it is what `bindgen include/sbx.h` produces (every entry point of the header; tests/test_abi_exports.py checks that
//! the two lists agree).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct sbx_drainage_params {
    pub dt: f32, pub gravity: f32, pub viscosity: f32, pub density: f32, pub diffusion_coeff: f32,
    pub bubble_radius: f32, pub critical_thickness: f32, pub grid_width: u32, pub grid_height: u32,
    pub marangoni_enabled: u32, pub gamma_air: f32, pub gamma_reduction: f32,
    pub surfactant_diffusion: f32, pub marangoni_coeff: f32, pub _padding1: u32, pub _padding2: u32,
}

#[repr(C)]
pub struct sbx_drainage_cfg {
    pub num_theta: i32, pub num_phi: i32, pub batch: i32,
    pub h0_m: f64, pub density: f64, pub gravity: f64, pub viscosity: f64, pub diffusion: f64,
    pub critical_thickness_m: f64, pub radius_m: f64,
    pub semantics: i32, pub precision: i32, pub h_ref_m: f64,
    pub tfe_enabled: i32, pub tfe_ratio: f64,
    pub row_begin: i32, pub row_end: i32,
    pub batch_h0_m: *const f64, pub batch_radius_m: *const f64,
    pub wgsl: sbx_drainage_params,
}

#[repr(C)]
pub struct sbx_view { pub cam_pos: [f64; 3], pub n_film: f64, pub intensity: f32, pub mode: i32, pub format: i32 }

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct sbx_tfe_front { pub phi_center: f64, pub extent: f64, pub width: f64, pub velocity: f64 }

/// what RenderPipeline::update reads after a step (pipeline.rs:1298-1312), fetched behind one synchronisation
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct sbx_frame_stats { pub min_thickness: f64, pub max_thickness: f64, pub thickness: f64, pub time: f64, pub has_critical: i32, pub _pad: i32 }

/// current_time, next_tfe_spawn_time and tfe_fronts of DrainageSimulator (drainage.rs:237-255): checkpoint / resume
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct sbx_tfe_state { pub time: f64, pub next_spawn: f64, pub count: i32, pub _pad: i32, pub fronts: [sbx_tfe_front; 8] }

/// One headless frame: Camera (camera.rs:49-60) + the BubbleUniform fields the fragment shader's shading path reads
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct sbx_frame_params {
    pub width: u32, pub height: u32,
    pub cam_distance: f32, pub cam_yaw: f32, pub cam_pitch: f32, pub cam_fov_rad: f32,
    pub radius: f32, pub base_thickness_nm: f32,
    pub film_time: f32, pub swirl_intensity: f32, pub drainage_speed: f32, pub pattern_scale: f32,
    pub interference_intensity: f32, pub base_alpha: f32, pub edge_alpha: f32,
    pub edge_smoothing_mode: u32,
    pub background: [f32; 3],
    pub lut_max_nm: f32,
}

/// CausticParams, src/render/caustics.rs:19-79 (64-byte POD)
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct sbx_caustic_params {
    pub grid_width: u32, pub grid_height: u32,
    pub refractive_index: f32, pub film_thickness_scale: f32,
    pub light_dir_x: f32, pub light_dir_y: f32, pub light_dir_z: f32, pub light_intensity: f32,
    pub focal_length: f32, pub caustic_intensity: f32, pub caustic_sharpness: f32, pub branch_threshold: f32,
    pub ground_y: f32, pub bubble_radius: f32, pub time: f32, pub _padding: u32,
}

pub enum sbx_drainage {}
pub enum sbx_comm {}
pub const SBX_IPC_BYTES: usize = 512;

pub const SBX_SEM_CPU: i32 = 0;
pub const SBX_SEM_WGSL: i32 = 1;
pub const SBX_PREC_F32: i32 = 0;
pub const SBX_PREC_F64: i32 = 1;
pub const SBX_PREC_F64_STRICT: i32 = 2;
pub const SBX_F32: i32 = 0;
pub const SBX_F64: i32 = 1;
pub const SBX_PTR_THICKNESS: i32 = 0;
pub const SBX_PTR_CONCENTRATION: i32 = 1;
pub const SBX_PTR_RGB: i32 = 2;
pub const SBX_PTR_CAUSTICS: i32 = 3;
pub const SBX_RGB_NONE: i32 = 0;
pub const SBX_RGB_LUT: i32 = 1;
pub const SBX_RGB_EXACT3: i32 = 2;
pub const SBX_RGB_EXACT3_LINEAR: i32 = 3;
pub const SBX_RGB_SCHLICK_WALL: i32 = 4;
pub const SBX_RGB_SCHLICK_TWO_BEAM: i32 = 5;
pub const SBX_RGB_SPECTRAL: i32 = 6;
pub const SBX_OUT_RGBA8: i32 = 0;
pub const SBX_OUT_RGB_F32: i32 = 1;

extern "C" {
    pub fn sbx_last_error() -> *const c_char;
    pub fn sbx_drainage_cfg_default(cfg: *mut sbx_drainage_cfg) -> i32;
    pub fn sbx_drainage_create(cfg: *const sbx_drainage_cfg, device: i32, out: *mut *mut sbx_drainage) -> i32;
    pub fn sbx_drainage_destroy(h: *mut sbx_drainage) -> i32;
    pub fn sbx_drainage_upload(h: *mut sbx_drainage, host: *const c_void, dtype: i32) -> i32;
    pub fn sbx_drainage_download(h: *mut sbx_drainage, host: *mut c_void, dtype: i32) -> i32;
    pub fn sbx_drainage_step(h: *mut sbx_drainage, dt: f64, nsteps: i32) -> i32;
    pub fn sbx_drainage_step_fused(h: *mut sbx_drainage, dt: f64, nsteps: i32, view: *const sbx_view) -> i32;
    pub fn sbx_drainage_get_thickness(h: *mut sbx_drainage, bubble: i32, theta: f64, phi: f64, out: *mut f64) -> i32;
    pub fn sbx_drainage_minmax(h: *mut sbx_drainage, min_out: *mut f64, max_out: *mut f64) -> i32;
    pub fn sbx_drainage_track_minmax(h: *mut sbx_drainage, mode: i32) -> i32;
    pub fn sbx_drainage_has_critical(h: *mut sbx_drainage, out: *mut i32) -> i32;
    pub fn sbx_drainage_critical_thickness(h: *mut sbx_drainage, out: *mut f64) -> i32;
    pub fn sbx_drainage_current_time(h: *mut sbx_drainage, out: *mut f64) -> i32;
    pub fn sbx_drainage_reset(h: *mut sbx_drainage, h0_m: f64) -> i32;
    pub fn sbx_drainage_set_marginal_regeneration(h: *mut sbx_drainage, enabled: i32) -> i32;
    pub fn sbx_drainage_marginal_regeneration(h: *mut sbx_drainage, out: *mut i32) -> i32;
    pub fn sbx_drainage_set_tfe_ratio(h: *mut sbx_drainage, ratio: f64) -> i32;
    pub fn sbx_drainage_tfe_ratio(h: *mut sbx_drainage, out: *mut f64) -> i32;
    pub fn sbx_drainage_active_tfe_count(h: *mut sbx_drainage, out: *mut i32) -> i32;
    pub fn sbx_drainage_update_params(h: *mut sbx_drainage, p: *const sbx_drainage_params) -> i32;
    pub fn sbx_drainage_device_ptr(h: *mut sbx_drainage, kind: i32, out: *mut *mut c_void) -> i32;
    pub fn sbx_lut_build(refractive_index: f32, intensity: f32, out65536: *mut u8, device: i32) -> i32;
    pub fn sbx_version(major: *mut i32, minor: *mut i32) -> i32;
    pub fn sbx_device_count(count: *mut i32) -> i32;
    pub fn sbx_host_alloc(bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn sbx_host_free(p: *mut c_void) -> i32;
    pub fn sbx_drainage_set_stream(h: *mut sbx_drainage, cuda_stream: *mut c_void) -> i32;
    pub fn sbx_drainage_sync(h: *mut sbx_drainage) -> i32;
    pub fn sbx_drainage_upload_concentration(h: *mut sbx_drainage, host: *const f32) -> i32;
    pub fn sbx_drainage_download_concentration(h: *mut sbx_drainage, host: *mut f32) -> i32;
    pub fn sbx_drainage_shade(h: *mut sbx_drainage, view: *const sbx_view) -> i32;
    pub fn sbx_drainage_step_host(h: *mut sbx_drainage, dt: f64, nsteps: i32, view: *const sbx_view, host_in: *const c_void,
                                  in_dtype: i32, host_out_h: *mut c_void, out_dtype: i32, host_out_rgb: *mut c_void) -> i32;
    pub fn sbx_drainage_download_rgb(h: *mut sbx_drainage, host: *mut c_void) -> i32;
    pub fn sbx_drainage_fronts(h: *mut sbx_drainage, out8: *mut sbx_tfe_front, count: *mut i32) -> i32;
    pub fn sbx_drainage_frame_stats(h: *mut sbx_drainage, bubble: i32, theta: f64, phi: f64, out: *mut sbx_frame_stats) -> i32;
    pub fn sbx_drainage_get_state(h: *mut sbx_drainage, out: *mut sbx_tfe_state) -> i32;
    pub fn sbx_drainage_set_state(h: *mut sbx_drainage, state: *const sbx_tfe_state) -> i32;
    pub fn sbx_drainage_params_get(h: *mut sbx_drainage, out: *mut sbx_drainage_params) -> i32;
    pub fn sbx_drainage_set_physics(h: *mut sbx_drainage, density: f64, gravity: f64, viscosity: f64, diffusion: f64,
                                    critical_thickness_m: f64, radius_m: f64) -> i32;
    pub fn sbx_drainage_set_lut(h: *mut sbx_drainage, rgba: *const u8, t: i32, a: i32, max_nm: f32) -> i32;
    pub fn sbx_drainage_set_spectrum(h: *mut sbx_drainage, wavelengths_nm: *const f32, n: i32) -> i32;
    pub fn sbx_drainage_launch_count(h: *mut sbx_drainage, out: *mut i64) -> i32;
    pub fn sbx_caustic_params_default(p: *mut sbx_caustic_params) -> i32;
    pub fn sbx_drainage_caustics(h: *mut sbx_drainage, p: *const sbx_caustic_params, host_out: *mut f32) -> i32;
    pub fn sbx_drainage_sample_uv(h: *mut sbx_drainage, bubble: i32, uv: *const f32, n: usize, out_h: *mut f32,
                                  out_grad: *mut f32) -> i32;
    pub fn sbx_drainage_sample_normal(h: *mut sbx_drainage, bubble: i32, normals: *const f32, n: usize, scale: f32,
                                      out_h: *mut f32) -> i32;
    pub fn sbx_lut_build_spectral(refractive_index: f32, intensity: f32, wavelengths_nm: *const f32, n_lambda: i32, t: i32,
                                  a: i32, max_nm: f32, out: *mut u8, device: i32) -> i32;
    pub fn sbx_frame_params_default(p: *mut sbx_frame_params) -> i32;
    pub fn sbx_render_bubble_frame(p: *const sbx_frame_params, lut_rgba: *const u8, t: i32, a: i32, rgba_out: *mut u8,
                                   device: i32) -> i32;
    pub fn sbx_drainage_render_frame(h: *mut sbx_drainage, bubble: i32, p: *const sbx_frame_params, n_film: f64,
                                     rgba_out: *mut u8) -> i32;
    pub fn sbx_export_frame_png(path: *const std::os::raw::c_char, width: u32, height: u32, rgba: *const u8,
                                len: usize) -> i32;
    pub fn sbx_lut_shade_points(lut_rgba: *const u8, t: i32, a: i32, max_nm: f32, d_nm: *const f32, cos_theta: *const f32,
                                n: usize, intensity: f32, rgb_out: *mut f32, device: i32) -> i32;
    pub fn sbx_interference_terms(n_film: f64, d_nm: *const f64, cos_i: *const f64, n: usize, wavelength_nm: f64,
                                  out_cos_t: *mut f64, out_fresnel: *mut f64, out_intensity: *mut f64, device: i32) -> i32;
    pub fn sbx_comm_unique_id(out128: *mut u8) -> i32;
    pub fn sbx_comm_init(id128: *const u8, nranks: i32, rank: i32, device: i32, out: *mut *mut sbx_comm) -> i32;
    pub fn sbx_comm_destroy(c: *mut sbx_comm) -> i32;
    pub fn sbx_drainage_attach_comm(h: *mut sbx_drainage, c: *mut sbx_comm) -> i32;
    pub fn sbx_drainage_ipc_export(h: *mut sbx_drainage, out: *mut u8) -> i32;
    pub fn sbx_drainage_ipc_attach(h: *mut sbx_drainage, up: *const u8, down: *const u8) -> i32;
    pub fn sbx_drainage_peer_status(h: *mut sbx_drainage, out: *mut i32) -> i32;
    pub fn sbx_drainage_minmax_global(h: *mut sbx_drainage, min_out: *mut f64, max_out: *mut f64) -> i32;
    pub fn sbx_interference_eval(n_film: f64, wavelengths_nm: *const f64, d_nm: *const f64, cos_i: *const f64,
                                 n: usize, gamma: i32, rgb_out: *mut f64, device: i32) -> i32;
}
