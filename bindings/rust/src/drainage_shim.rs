//! Safe wrappers with the exact signatures of the reference's hot-path types, so that
//! `src/physics/drainage.rs`, `src/render/gpu_drainage.rs` and `src/render/interference_lut.rs` can be
//! replaced file-for-file.  Source only -- not compiled in this repository's image.
use crate::config::SimulationConfig;
use crate::sbx_sys::*;
use std::cell::{Cell, RefCell};
use std::ffi::CStr;

fn check(rc: i32) {
    // the reference hot path is infallible; a CUDA failure is treated like the OOB-index panic it has today
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(sbx_last_error()) }.to_string_lossy().into_owned();
        panic!("libsbx error {rc}: {msg}");
    }
}

/// Host mirror of the device field (drainage.rs:96-204); refreshed lazily by `thickness_field()`.
#[derive(Debug, Clone)]
pub struct ThicknessField { num_theta: usize, num_phi: usize, data: Vec<f64>, delta_theta: f64, delta_phi: f64 }

impl ThicknessField {
    pub fn new(num_theta: usize, num_phi: usize, initial_thickness: f64) -> Self {
        Self { num_theta, num_phi, data: vec![initial_thickness; num_theta * num_phi],
               delta_theta: std::f64::consts::PI / (num_theta - 1) as f64,
               delta_phi: 2.0 * std::f64::consts::PI / num_phi as f64 }
    }
    #[inline] pub fn get(&self, t: usize, p: usize) -> f64 { self.data[t * self.num_phi + p] }
    #[inline] pub fn set(&mut self, t: usize, p: usize, v: f64) { self.data[t * self.num_phi + p] = v; }
    #[inline] pub fn theta_at(&self, t: usize) -> f64 { t as f64 * self.delta_theta }
    #[inline] pub fn phi_at(&self, p: usize) -> f64 { p as f64 * self.delta_phi }
    pub fn num_theta(&self) -> usize { self.num_theta }
    pub fn num_phi(&self) -> usize { self.num_phi }
    pub fn delta_theta(&self) -> f64 { self.delta_theta }
    pub fn delta_phi(&self) -> f64 { self.delta_phi }
    pub fn min_thickness(&self) -> f64 { self.data.iter().copied().fold(f64::INFINITY, f64::min) }
    pub fn max_thickness(&self) -> f64 { self.data.iter().copied().fold(f64::NEG_INFINITY, f64::max) }
    pub fn data(&self) -> &[f64] { &self.data }
    pub fn data_mut(&mut self) -> &mut [f64] { &mut self.data }
}

/// DrainageSimulator (drainage.rs:212-614) over libsbx.
pub struct DrainageSimulator {
    h: *mut sbx_drainage,
    mirror: RefCell<ThicknessField>,
    stale: Cell<bool>,
    /// per-frame statistics, fetched behind one synchronisation by the first query after a step (sbx_drainage_frame_stats)
    stats: Cell<Option<(sbx_frame_stats, (f64, f64))>>,
    last_query: Cell<(f64, f64)>,
}

impl DrainageSimulator {
    pub fn new(config: &SimulationConfig) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<sbx_drainage_cfg>() };
        check(unsafe { sbx_drainage_cfg_default(&mut cfg) });
        cfg.num_theta = config.resolution as i32;
        cfg.num_phi = config.resolution as i32 * 2;
        cfg.h0_m = config.film_thickness_meters();
        cfg.density = config.fluid.density;
        cfg.gravity = config.environment.gravity;
        cfg.viscosity = config.fluid.viscosity;
        cfg.diffusion = config.fluid.surfactant_diffusion;
        cfg.critical_thickness_m = config.bubble.critical_thickness_nm * 1e-9;
        cfg.radius_m = config.bubble_radius();
        cfg.precision = SBX_PREC_F32;
        cfg.h_ref_m = cfg.h0_m;
        let mut h = std::ptr::null_mut();
        check(unsafe { sbx_drainage_create(&cfg, 0, &mut h) });
        let mirror = ThicknessField::new(cfg.num_theta as usize, cfg.num_phi as usize, cfg.h0_m);
        Self { h, mirror: RefCell::new(mirror), stale: Cell::new(true), stats: Cell::new(None),
               last_query: Cell::new((std::f64::consts::FRAC_PI_2, 0.0)) }  // what pipeline.rs:1303-1306 asks for
    }
    /// The frame loop (pipeline.rs:1289-1322) calls step, min_thickness, max_thickness, get_thickness(pi/2, 0) and
    /// has_critical_region one after the other; the first of those queries after a step fetches all of them at once.
    fn frame(&self, query: Option<(f64, f64)>) -> sbx_frame_stats {
        if let Some((st, q)) = self.stats.get() {
            if query.map_or(true, |want| want == q) { return st; }
        }
        let q = query.unwrap_or(self.last_query.get());
        let mut st: sbx_frame_stats = unsafe { std::mem::zeroed() };
        check(unsafe { sbx_drainage_frame_stats(self.h, 0, q.0, q.1, &mut st) });
        self.stats.set(Some((st, q)));
        st
    }
    pub fn step(&mut self, dt: f64) { check(unsafe { sbx_drainage_step(self.h, dt, 1) }); self.stale.set(true); self.stats.set(None); }
    /// n applications of `step(dt)` in one call.  On grids that fit in shared memory (the default 128 x 256 does) the
    /// library takes all n steps inside one kernel launch: ~1.8 us per step instead of ~14 us.
    pub fn step_many(&mut self, dt: f64, n: usize) { check(unsafe { sbx_drainage_step(self.h, dt, n as i32) }); self.stale.set(true); self.stats.set(None); }
    pub fn get_thickness(&self, theta: f64, phi: f64) -> f64 {
        self.last_query.set((theta, phi));
        self.frame(Some((theta, phi))).thickness
    }
    /// `&ThicknessField` has to be host memory: download on demand (this is why sbx_drainage_download exists).
    pub fn thickness_field(&self) -> std::cell::Ref<'_, ThicknessField> {
        if self.stale.get() {
            let mut m = self.mirror.borrow_mut();
            check(unsafe { sbx_drainage_download(self.h, m.data_mut().as_mut_ptr() as *mut _, SBX_F64) });
            self.stale.set(false);
        }
        self.mirror.borrow()
    }
    /// device-side reductions for the two folds `RenderPipeline::update` does every frame (pipeline.rs:1298-1300)
    pub fn min_max_thickness(&self) -> (f64, f64) {
        let st = self.frame(None);
        (st.min_thickness, st.max_thickness)
    }
    pub fn current_time(&self) -> f64 { let mut t = 0.0; check(unsafe { sbx_drainage_current_time(self.h, &mut t) }); t }
    pub fn has_critical_region(&self) -> bool { self.frame(None).has_critical != 0 }
    pub fn critical_thickness(&self) -> f64 { let mut t = 0.0; check(unsafe { sbx_drainage_critical_thickness(self.h, &mut t) }); t }
    pub fn reset(&mut self, initial_thickness: f64) { check(unsafe { sbx_drainage_reset(self.h, initial_thickness) }); self.stale.set(true); self.stats.set(None); }
    pub fn set_marginal_regeneration(&mut self, enabled: bool) { check(unsafe { sbx_drainage_set_marginal_regeneration(self.h, enabled as i32) }); }
    pub fn marginal_regeneration_enabled(&self) -> bool { let mut f = 0; check(unsafe { sbx_drainage_marginal_regeneration(self.h, &mut f) }); f != 0 }
    pub fn set_tfe_thickness_ratio(&mut self, ratio: f64) { check(unsafe { sbx_drainage_set_tfe_ratio(self.h, ratio) }); }
    pub fn tfe_thickness_ratio(&self) -> f64 { let mut r = 0.0; check(unsafe { sbx_drainage_tfe_ratio(self.h, &mut r) }); r }
    pub fn active_tfe_count(&self) -> usize { let mut n = 0; check(unsafe { sbx_drainage_active_tfe_count(self.h, &mut n) }); n as usize }
    /// Not in the reference (its fields are private and it has no serialisation): current_time, next_tfe_spawn_time
    /// and tfe_fronts (drainage.rs:237-255) for checkpoint / resume; the field itself travels through thickness_field().
    pub fn tfe_state(&self) -> sbx_tfe_state {
        let mut st: sbx_tfe_state = unsafe { std::mem::zeroed() };
        check(unsafe { sbx_drainage_get_state(self.h, &mut st) });
        st
    }
    pub fn restore_tfe_state(&mut self, st: &sbx_tfe_state) { check(unsafe { sbx_drainage_set_state(self.h, st) }); }
}

impl Drop for DrainageSimulator { fn drop(&mut self) { unsafe { sbx_drainage_destroy(self.h) }; } }

/// GPUDrainageSimulator (src/render/gpu_drainage.rs:81-491) over libsbx: the same methods without the wgpu
/// `device` / `queue` / `encoder` arguments.  Buffers are CUDA device pointers (`current_thickness_buffer`), valid until
/// the next step.
pub struct GPUDrainageSimulator {
    h: *mut sbx_drainage,
    params: sbx_drainage_params,
    grid_width: u32,
    grid_height: u32,
    current_time: f32,
    pub enabled: bool,          // gpu_drainage.rs:107 (false by default)
    pub time_scale: f32,        // :109 (100.0)
    pub steps_per_frame: u32,   // :111 (10)
    pub marangoni_enabled: bool,
}

impl GPUDrainageSimulator {
    pub fn new(initial_thickness: f32, grid_width: u32, grid_height: u32) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<sbx_drainage_cfg>() };
        check(unsafe { sbx_drainage_cfg_default(&mut cfg) });
        cfg.semantics = SBX_SEM_WGSL;
        cfg.num_phi = grid_width as i32;
        cfg.num_theta = grid_height as i32;
        cfg.h0_m = initial_thickness as f64;
        let mut h = std::ptr::null_mut();
        check(unsafe { sbx_drainage_create(&cfg, 0, &mut h) });
        let mut params = cfg.wgsl;
        check(unsafe { sbx_drainage_params_get(h, &mut params) });
        Self { h, params, grid_width, grid_height, current_time: 0.0, enabled: false, time_scale: 100.0,
               steps_per_frame: 10, marangoni_enabled: false }
    }
    /// gpu_drainage.rs:337-370: `steps_per_frame` dispatches of drainage_step, one library call (resident in shared
    /// memory on the reference's 128 x 64 grid: 0.8 us per dispatch).
    pub fn step(&mut self, dt: f32) {
        if !self.enabled { return; }
        check(unsafe { sbx_drainage_step(self.h, self.params.dt as f64, self.steps_per_frame as i32) });
        self.current_time += dt * self.time_scale;
    }
    pub fn update_params(&mut self, params: sbx_drainage_params) {
        self.params = params;
        check(unsafe { sbx_drainage_update_params(self.h, &self.params) });
    }
    pub fn set_dt(&mut self, dt: f32) { self.params.dt = dt; self.push(); }
    pub fn set_physics(&mut self, viscosity: f32, density: f32, gravity: f32, diffusion: f32) {
        self.params.viscosity = viscosity; self.params.density = density;
        self.params.gravity = gravity; self.params.diffusion_coeff = diffusion; self.push();
    }
    pub fn set_bubble_radius(&mut self, radius: f32) { self.params.bubble_radius = radius; self.push(); }
    pub fn reset(&mut self, initial_thickness: f32) {
        check(unsafe { sbx_drainage_reset(self.h, initial_thickness as f64) });
        self.current_time = 0.0;
    }
    pub fn set_marangoni_enabled(&mut self, enabled: bool) {
        self.marangoni_enabled = enabled; self.params.marangoni_enabled = enabled as u32; self.push();
    }
    pub fn set_marangoni_params(&mut self, gamma_air: f32, gamma_reduction: f32, surfactant_diffusion: f32, marangoni_coeff: f32) {
        self.params.gamma_air = gamma_air; self.params.gamma_reduction = gamma_reduction;
        self.params.surfactant_diffusion = surfactant_diffusion; self.params.marangoni_coeff = marangoni_coeff; self.push();
    }
    pub fn concentration_buffer(&self) -> *mut std::os::raw::c_void { self.ptr(SBX_PTR_CONCENTRATION) }
    pub fn current_buffer(&self) -> *mut std::os::raw::c_void { self.ptr(SBX_PTR_THICKNESS) }
    pub fn current_thickness_buffer(&self) -> *mut std::os::raw::c_void { self.ptr(SBX_PTR_THICKNESS) }
    pub fn current_time(&self) -> f32 { self.current_time }
    pub fn grid_size(&self) -> (u32, u32) { (self.grid_width, self.grid_height) }
    pub fn params(&self) -> &sbx_drainage_params { &self.params }
    /// CausticRenderer::compute (caustics.rs:441-457) on the latest thickness buffer; the map stays on the device.
    pub fn compute_caustics(&mut self, p: &sbx_caustic_params) { check(unsafe { sbx_drainage_caustics(self.h, p, std::ptr::null_mut()) }); }
    fn push(&mut self) { check(unsafe { sbx_drainage_update_params(self.h, &self.params) }); }
    fn ptr(&self, kind: i32) -> *mut std::os::raw::c_void {
        let mut p = std::ptr::null_mut();
        check(unsafe { sbx_drainage_device_ptr(self.h, kind, &mut p) });
        p
    }
}

impl Drop for GPUDrainageSimulator { fn drop(&mut self) { unsafe { sbx_drainage_destroy(self.h) }; } }

/// generate_interference_lut (interference_lut.rs:18-41)
pub const LUT_THICKNESS_SAMPLES: u32 = 256;
pub const LUT_ANGLE_SAMPLES: u32 = 64;
pub const LUT_MAX_THICKNESS_NM: f32 = 2000.0;
pub fn generate_interference_lut(refractive_index: f32, intensity: f32) -> Vec<u8> {
    let mut data = vec![0u8; (LUT_THICKNESS_SAMPLES * LUT_ANGLE_SAMPLES * 4) as usize];
    check(unsafe { sbx_lut_build(refractive_index, intensity, data.as_mut_ptr(), 0) });
    data
}

/// HeadlessRenderPipeline for the bubble (src/render/headless.rs:27-552): same public fields and methods; `new` is no
/// longer async because no wgpu adapter is requested.  The frame is bubble.wgsl::fs_main evaluated per pixel by libsbx.
pub struct HeadlessRenderPipeline {
    width: u32, height: u32, lut: Vec<u8>,
    pub camera: crate::render::camera::Camera,
    pub bubble_uniform: crate::render::pipeline::BubbleUniform,
}

impl HeadlessRenderPipeline {
    pub fn new(width: u32, height: u32) -> Option<Self> {
        let bubble_uniform = crate::render::pipeline::BubbleUniform::default();
        let mut ndev = 0;
        if unsafe { sbx_device_count(&mut ndev) } != 0 || ndev == 0 { return None; }  // "no GPU available": tests skip
        Some(Self { width, height, lut: generate_interference_lut(bubble_uniform.refractive_index, 1.0),  // headless.rs:143-146
                    camera: crate::render::camera::Camera::new(width as f32 / height as f32), bubble_uniform })
    }
    pub fn render_to_buffer(&mut self) -> Vec<u8> {
        let b = &self.bubble_uniform;
        let p = sbx_frame_params {
            width: self.width, height: self.height,
            cam_distance: self.camera.distance, cam_yaw: self.camera.yaw, cam_pitch: self.camera.pitch,
            cam_fov_rad: self.camera.fov, radius: 0.025, base_thickness_nm: b.base_thickness_nm,
            film_time: b.film_time, swirl_intensity: b.swirl_intensity, drainage_speed: b.drainage_speed,
            pattern_scale: b.pattern_scale, interference_intensity: b.interference_intensity,
            base_alpha: b.base_alpha, edge_alpha: b.edge_alpha, edge_smoothing_mode: b.edge_smoothing_mode,
            background: [b.background_r, b.background_g, b.background_b], lut_max_nm: LUT_MAX_THICKNESS_NM,
        };
        let mut pixels = vec![0u8; (self.width * self.height * 4) as usize];
        check(unsafe { sbx_render_bubble_frame(&p, self.lut.as_ptr(), LUT_THICKNESS_SAMPLES as i32,
                                               LUT_ANGLE_SAMPLES as i32, pixels.as_mut_ptr(), 0) });
        pixels
    }
    pub fn size(&self) -> (u32, u32) { (self.width, self.height) }
    pub fn set_thickness(&mut self, thickness_nm: f32) { self.bubble_uniform.base_thickness_nm = thickness_nm; }
    pub fn set_refractive_index(&mut self, n: f32) { self.bubble_uniform.refractive_index = n; }
    pub fn orbit_camera(&mut self, delta_x: f32, delta_y: f32) { self.camera.orbit(delta_x, delta_y); }
    pub fn zoom_camera(&mut self, delta: f32) { self.camera.zoom(delta); }
    pub fn set_camera_distance(&mut self, distance: f32) { self.camera.distance = distance.clamp(0.05, 10.0); }
    pub fn set_time(&mut self, time: f32) { self.bubble_uniform.time = time; self.bubble_uniform.film_time = time; }
}

/// export_frame (src/export/image_export.rs:42-76) with the reference's error enum.
pub fn export_frame<P: AsRef<std::path::Path>>(path: P, width: u32, height: u32, data: &[u8])
    -> Result<(), crate::export::ExportError> {
    use crate::export::ExportError;
    if width == 0 || height == 0 { return Err(ExportError::InvalidDimensions { width, height }); }
    let expected_len = (width * height * 4) as usize;
    if data.len() != expected_len {
        return Err(ExportError::BufferCreation(format!("Data length {} doesn't match expected {} ({}x{}x4)",
                                                       data.len(), expected_len, width, height)));
    }
    let c = std::ffi::CString::new(path.as_ref().to_string_lossy().as_bytes())
        .map_err(|e| ExportError::SaveError(e.to_string()))?;
    match unsafe { sbx_export_frame_png(c.as_ptr(), width, height, data.as_ptr(), data.len()) } {
        0 => Ok(()),
        _ => Err(ExportError::SaveError(unsafe { CStr::from_ptr(sbx_last_error()) }.to_string_lossy().into_owned())),
    }
}
