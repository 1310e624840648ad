/*
 * sbx.h -- C ABI of libsbx.so: the B200-native drop-in for the reference's thin-film drainage
 * step + thin-film interference path (pjt222/soap_bubble_simulation).
 *
 * Every entry point cites the reference interface it replaces (file:line relative to the
 * reference repository root).  The reference has no FFI today (pure Rust + wgpu); the Rust binding a
 * maintainer would add is shown in INTEGRATION.md and shipped as source in bindings/rust/.
 *
 * Conventions
 *   - All functions return int32_t: SBX_OK (0) or a negative sbx_status; sbx_last_error() returns a
 *     thread-local, human-readable description of the last failure on the calling thread.
 *   - Fields are row-major [bubble][theta][phi], phi contiguous, index = theta*num_phi + phi
 *     (src/physics/drainage.rs:103,138; src/render/shaders/drainage.wgsl:42-44); thickness in METRES.
 *     theta_i = i*pi/(num_theta-1) includes both poles, phi_j = j*2pi/num_phi is periodic.
 *   - The library owns all device memory and streams behind opaque handles; callers own every host
 *     buffer they pass.  Handles are not thread-safe.  step*() calls are asynchronous on the handle's
 *     stream; download/minmax/get_thickness/has_critical synchronise that stream.
 *   - There is NO CPU fallback: every compute entry point fails with SBX_ERR_CUDA when no CUDA
 *     device is usable.
 */
#ifndef SBX_H
#define SBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBX_VERSION_MAJOR 0
#define SBX_VERSION_MINOR 1

typedef enum {
    SBX_OK = 0,
    SBX_ERR_INVALID = -1,     /* bad argument / configuration */
    SBX_ERR_CUDA = -2,        /* CUDA runtime error, or no device */
    SBX_ERR_NOMEM = -3,
    SBX_ERR_UNSUPPORTED = -4, /* valid request this build does not implement */
    SBX_ERR_COMM = -5,        /* NCCL / peer-memory failure */
    SBX_ERR_IO = -6           /* file could not be written (ExportError::SaveError, image_export.rs:11) */
} sbx_status;

/* Which reference implementation the step reproduces. */
typedef enum {
    SBX_SEM_CPU = 0, /* DrainageSimulator::step, src/physics/drainage.rs:298-392 (gravity + diffusion,
                        pole averaging, marginal-regeneration fronts) */
    SBX_SEM_WGSL = 1 /* drainage.wgsl::drainage_step, src/render/shaders/drainage.wgsl:72-202 (f32,
                        second concentration field, Marangoni coupling, poles frozen, no fronts) */
} sbx_semantics;

/* Storage/arithmetic type of the thickness field on the device. */
typedef enum {
    SBX_PREC_F32 = 0,       /* 4 B/cell; field stored as (h - h_ref_m), see sbx_drainage_cfg.h_ref_m */
    SBX_PREC_F64 = 1,       /* 8 B/cell; divisions folded into per-row reciprocals */
    SBX_PREC_F64_STRICT = 2 /* 8 B/cell; the oracle's exact operation order (no FMA contraction, true
                               divisions, left-to-right pole sums): bit-identical to the f64 CPU path */
} sbx_precision;

typedef enum { SBX_F32 = 0, SBX_F64 = 1 } sbx_dtype;

/* How thickness x view angle becomes colour in the fused step. */
typedef enum {
    SBX_RGB_NONE = 0,
    SBX_RGB_LUT = 1,            /* bubble.wgsl:425-439: bilinear fetch of the 256x64 RGBA8 LUT at
                                   (d/2000 nm, cos theta), times intensity, clamp -- the reference's
                                   production shading path */
    SBX_RGB_EXACT3 = 2,         /* InterferenceCalculator::calculate_interference_color,
                                   src/physics/interference.rs:380-400 (3 wavelengths, exact Fresnel,
                                   multi-bounce Airy, gamma 1/2.2) */
    SBX_RGB_EXACT3_LINEAR = 3,  /* calculate_interference_color_linear, interference.rs:413-426 */
    SBX_RGB_SCHLICK_WALL = 4,   /* wall.wgsl:165-207: Schlick Fresnel, transmitted Airy, 650/532/450 nm */
    SBX_RGB_SPECTRAL = 6,       /* direct per-cell spectral evaluation without a LUT: thin_film_interference of the
                                   fragment shaders (bubble_instanced.wgsl:190-218, bubble.wgsl:442-483 ==
                                   interference_lut.rs:45-93) with the wavelength list as a parameter
                                   (sbx_drainage_set_spectrum; default the reference's 7 wavelengths 400..700 nm):
                                   Snell, unpolarised Fresnel, Airy per wavelength, CIE Gaussians -> XYZ -> linear
                                   sRGB, times intensity, clamp.  SFU-bound by construction (BASELINE config 5's
                                   direct half); colour only -- the drainage step runs first, then this pass */
    SBX_RGB_SCHLICK_TWO_BEAM = 5 /* the documented original model (docs/key-insights.md:288-309): two-beam
                                   I = (1+cos phi)/2 with phi = 2 pi (2 n d cos theta_t)/lambda + pi, times Schlick
                                   Fresnel, times 4 -- what README/north_star describe; no compiled reference code
                                   implements it any more, so it is checked against the restated formula only */
} sbx_rgb_mode;

typedef enum {
    SBX_OUT_RGBA8 = 0,  /* 4 B/cell, A = 255, channel = round(255*x) (RgbColor::to_u8_tuple,
                           interference.rs:73-79) */
    SBX_OUT_RGB_F32 = 1 /* 12 B/cell, 3 floats per cell (parity mode) */
} sbx_rgb_format;

typedef enum {
    SBX_PTR_THICKNESS = 0,     /* latest thickness, first owned row (device pointer) */
    SBX_PTR_CONCENTRATION = 1, /* latest concentration (SBX_SEM_WGSL only) */
    SBX_PTR_RGB = 2,           /* last colour frame written by step_fused/shade */
    SBX_PTR_CAUSTICS = 3       /* last caustic map written by sbx_drainage_caustics */
} sbx_ptr_kind;

/* DrainageParams, src/render/gpu_drainage.rs:10-44 == src/render/shaders/drainage.wgsl:14-31
 * (64-byte #[repr(C)] POD, same field order). */
typedef struct {
    float dt, gravity, viscosity, density, diffusion_coeff, bubble_radius, critical_thickness;
    uint32_t grid_width;  /* phi points */
    uint32_t grid_height; /* theta points */
    uint32_t marangoni_enabled;
    float gamma_air, gamma_reduction, surfactant_diffusion, marangoni_coeff;
    uint32_t _padding1, _padding2;
} sbx_drainage_params;

/* What DrainageSimulator::new reads from SimulationConfig (src/physics/drainage.rs:263-288) plus the
 * device-side choices.  Fill with sbx_drainage_cfg_default() first. */
typedef struct {
    int32_t num_theta;            /* global theta points (config.resolution) */
    int32_t num_phi;              /* phi points (2*config.resolution in the reference) */
    int32_t batch;                /* independent bubbles sharing the grid shape (foam), >= 1 */
    double h0_m;                  /* initial uniform thickness (film_thickness_nm * 1e-9) */
    double density;               /* fluid.density */
    double gravity;               /* environment.gravity */
    double viscosity;             /* fluid.viscosity */
    double diffusion;             /* fluid.surfactant_diffusion */
    double critical_thickness_m;  /* bubble.critical_thickness_nm * 1e-9 */
    double radius_m;              /* bubble.diameter / 2 */
    int32_t semantics;            /* sbx_semantics */
    int32_t precision;            /* sbx_precision */
    double h_ref_m;               /* SBX_PREC_F32: the device field holds (h - h_ref_m) so that sub-ulp
                                     increments around a uniform start are kept; 0 = plain f32 like the
                                     WGSL path.  Ignored for f64. */
    int32_t tfe_enabled;          /* marginal_regeneration_enabled, drainage.rs:284 */
    double tfe_ratio;             /* tfe_thickness_ratio, drainage.rs:283 */
    /* latitude-band sharding: this handle owns global rows [row_begin, row_end); 0,0 = whole sphere */
    int32_t row_begin, row_end;
    /* optional per-bubble overrides for batch > 1 (NULL = use h0_m / radius_m for all) */
    const double *batch_h0_m;
    const double *batch_radius_m;
    /* SBX_SEM_WGSL: initial parameters (grid_width/grid_height are taken from num_phi/num_theta) */
    sbx_drainage_params wgsl;
} sbx_drainage_cfg;

/* Camera + optics for the colour stage.  cos(theta_i) per cell is |dot(normalize(cam - R n), n)| with
 * n = (sin th cos ph, cos th, sin th sin ph) (src/physics/geometry.rs:184-186, bubble.wgsl:487-494). */
typedef struct {
    double cam_pos[3]; /* Camera::position(), src/render/camera.rs:63-68 */
    double n_film;     /* bubble.refractive_index (f64 like InterferenceCalculator; the LUT path rounds it to f32) */
    float intensity;   /* bubble.interference_intensity (4.0 default, pipeline.rs:75); LUT mode only */
    int32_t mode;      /* sbx_rgb_mode */
    int32_t format;    /* sbx_rgb_format */
} sbx_view;

typedef struct {
    double phi_center, extent, width, velocity; /* TfeFront, src/physics/drainage.rs:23-33 */
} sbx_tfe_front;

typedef struct sbx_drainage sbx_drainage;
typedef struct sbx_comm sbx_comm;

/* ---------------------------------------------------------------- library */
const char *sbx_last_error(void);
int32_t sbx_version(int32_t *major, int32_t *minor);
int32_t sbx_device_count(int32_t *count);
/* page-locked host memory for the host-buffer entry points */
int32_t sbx_host_alloc(size_t bytes, void **out);
int32_t sbx_host_free(void *p);

/* ---------------------------------------------------------------- DrainageSimulator / GPUDrainageSimulator */
/* SimulationConfig::default + config/default.json (src/config.rs:26-35,56-66,93-103,305-316) and
 * DrainageParams::default (src/render/gpu_drainage.rs:46-68). */
int32_t sbx_drainage_cfg_default(sbx_drainage_cfg *cfg);
/* DrainageSimulator::new (drainage.rs:263-288) / GPUDrainageSimulator::new (gpu_drainage.rs:118-331) */
int32_t sbx_drainage_create(const sbx_drainage_cfg *cfg, int32_t device, sbx_drainage **out);
int32_t sbx_drainage_destroy(sbx_drainage *h);
/* run subsequent work on a caller-provided cudaStream_t (NULL restores the handle's own stream) */
int32_t sbx_drainage_set_stream(sbx_drainage *h, void *cuda_stream);
int32_t sbx_drainage_sync(sbx_drainage *h);

/* ThicknessField::data_mut / data (drainage.rs:196-203): whole owned field, host <-> device.
 * host holds batch*(row_end-row_begin)*num_phi elements of `dtype`. */
int32_t sbx_drainage_upload(sbx_drainage *h, const void *host, int32_t dtype);
int32_t sbx_drainage_download(sbx_drainage *h, void *host, int32_t dtype);
/* concentration buffers (gpu_drainage.rs:148-159, 459-461), SBX_SEM_WGSL only, f32 */
int32_t sbx_drainage_upload_concentration(sbx_drainage *h, const float *host);
int32_t sbx_drainage_download_concentration(sbx_drainage *h, float *host);

/* DrainageSimulator::step (drainage.rs:298-392) nsteps times; for SBX_SEM_WGSL one call is nsteps
 * dispatches of drainage_step with params.dt (dt is then only used to advance current_time, as in
 * GPUDrainageSimulator::step, gpu_drainage.rs:338-370). */
int32_t sbx_drainage_step(sbx_drainage *h, double dt, int32_t nsteps);
/* the fused step: every step also writes the colour frame of the updated field (one HBM round trip
 * of the thickness field per step). */
int32_t sbx_drainage_step_fused(sbx_drainage *h, double dt, int32_t nsteps, const sbx_view *view);
/* colour frame of the current field without stepping */
int32_t sbx_drainage_shade(sbx_drainage *h, const sbx_view *view);
/* host-buffer round trip used for end-to-end timing: upload `host_in` (dtype), nsteps fused steps,
 * download thickness into host_out_h (may be NULL) and the colour frame into host_out_rgb (may be
 * NULL).  Host buffers should be page-locked (sbx_host_alloc). */
int32_t sbx_drainage_step_host(sbx_drainage *h, double dt, int32_t nsteps, const sbx_view *view,
                               const void *host_in, int32_t in_dtype, void *host_out_h,
                               int32_t out_dtype, void *host_out_rgb);
int32_t sbx_drainage_download_rgb(sbx_drainage *h, void *host);

/* DrainageSimulator::get_thickness (drainage.rs:516-552) */
int32_t sbx_drainage_get_thickness(sbx_drainage *h, int32_t bubble, double theta, double phi, double *out);
/* ThicknessField::min_thickness / max_thickness (drainage.rs:180-193), over the owned rows */
int32_t sbx_drainage_minmax(sbx_drainage *h, double *min_out, double *max_out);
/* has_critical_region (drainage.rs:565-567) */
/* min / max as a by-product of the step kernel instead of a pass over the field: 0 = never, 1 = every step the TMA
 * kernel takes, 2 = automatic (default): switched on once queries follow steps, as in the reference's frame loop
 * (step, min_thickness, max_thickness, has_critical_region: pipeline.rs:1289-1312), and off again when they stop.
 * Results are identical either way; no reference equivalent (its folds are separate passes, drainage.rs:180-193). */
int32_t sbx_drainage_track_minmax(sbx_drainage *h, int32_t mode);
int32_t sbx_drainage_has_critical(sbx_drainage *h, int32_t *out);
/* Everything RenderPipeline::update reads after a step (src/render/pipeline.rs:1298-1312: min_thickness, max_thickness,
 * get_thickness(theta, phi), has_critical_region; plus current_time) in ONE call and one synchronisation of the handle's
 * stream instead of one per query.  Same values as the separate entry points.  A shim that keeps the reference's
 * separate methods calls this from the first query after a step (with the (theta, phi) its caller asked for last
 * frame) and serves the others from the result. */
typedef struct {
    double min_thickness, max_thickness; /* metres, over the owned rows of all bubbles */
    double thickness;                    /* get_thickness(theta, phi) of `bubble` */
    double time;                         /* current_time */
    int32_t has_critical, _pad;
} sbx_frame_stats;
int32_t sbx_drainage_frame_stats(sbx_drainage *h, int32_t bubble, double theta, double phi, sbx_frame_stats *out);
int32_t sbx_drainage_critical_thickness(sbx_drainage *h, double *out); /* drainage.rs:570-572 */
int32_t sbx_drainage_current_time(sbx_drainage *h, double *out);       /* drainage.rs:560-562; gpu_drainage.rs:464-466 */
/* reset (drainage.rs:575-583; gpu_drainage.rs:415-432: also concentration := 0.5) */
int32_t sbx_drainage_reset(sbx_drainage *h, double h0_m);
int32_t sbx_drainage_set_marginal_regeneration(sbx_drainage *h, int32_t enabled); /* drainage.rs:588-593 */
int32_t sbx_drainage_marginal_regeneration(sbx_drainage *h, int32_t *out);        /* drainage.rs:596-598 */
int32_t sbx_drainage_set_tfe_ratio(sbx_drainage *h, double ratio);                /* drainage.rs:601-603 */
int32_t sbx_drainage_tfe_ratio(sbx_drainage *h, double *out);                     /* drainage.rs:606-608 */
int32_t sbx_drainage_active_tfe_count(sbx_drainage *h, int32_t *out);             /* drainage.rs:611-613 */
int32_t sbx_drainage_fronts(sbx_drainage *h, sbx_tfe_front *out8, int32_t *count);
/* Checkpoint / resume of the host state machine: the fields DrainageSimulator keeps beside the thickness field
 * (current_time, tfe_fronts, next_tfe_spawn_time; drainage.rs:237-255).  Together with download / upload of the
 * field this restores a simulator exactly (the reference has no serialisation of its own; a shim would expose
 * these as plain getters / setters).  set_state rejects count > 8 and non-finite values. */
typedef struct {
    double time;       /* current_time */
    double next_spawn; /* next_tfe_spawn_time */
    int32_t count;     /* tfe_fronts.len(), <= 8 (drainage.rs:474) */
    int32_t _pad;
    sbx_tfe_front fronts[8];
} sbx_tfe_state;
int32_t sbx_drainage_get_state(sbx_drainage *h, sbx_tfe_state *out);
int32_t sbx_drainage_set_state(sbx_drainage *h, const sbx_tfe_state *state);

/* GPUDrainageSimulator setters (gpu_drainage.rs:377-456), SBX_SEM_WGSL only */
int32_t sbx_drainage_update_params(sbx_drainage *h, const sbx_drainage_params *p); /* :377-380 */
int32_t sbx_drainage_params_get(sbx_drainage *h, sbx_drainage_params *out);        /* :488-490 */
/* generic physics update for SBX_SEM_CPU handles (the Rust struct fields, drainage.rs:219-235) */
int32_t sbx_drainage_set_physics(sbx_drainage *h, double density, double gravity, double viscosity,
                                 double diffusion, double critical_thickness_m, double radius_m);

/* Borrowed device pointer to the latest buffer of `kind`; invalidated by the next step (ping-pong).
 * The reference's current_buffer() returns the second-to-last step's buffer (gpu_drainage.rs:471-475,
 * see SURVEY.md section 3C); this returns the latest. */
int32_t sbx_drainage_device_ptr(sbx_drainage *h, int32_t kind, void **out);

/* install a T x A RGBA8 LUT (row-major [angle][thickness]) for SBX_RGB_LUT; by default the handle
 * builds generate_interference_lut(view.n_film, 1.0) itself on first use */
int32_t sbx_drainage_set_lut(sbx_drainage *h, const uint8_t *rgba, int32_t T, int32_t A, float max_nm);

/* wavelength list of SBX_RGB_SPECTRAL (1 <= n <= 256, nanometres); the handle starts with the reference's
 * seven (400, 450, ..., 700 nm; bubble_instanced.wgsl:191) */
int32_t sbx_drainage_set_spectrum(sbx_drainage *h, const float *wavelengths_nm, int32_t n);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int32_t sbx_drainage_launch_count(sbx_drainage *h, int64_t *out);

/* ---------------------------------------------------------------- thickness-buffer consumer: caustic map */
/* CausticParams (src/render/caustics.rs:19-79 == src/render/shaders/caustics_compute.wgsl:5-22), 64-byte POD */
typedef struct {
    uint32_t grid_width, grid_height;
    float refractive_index, film_thickness_scale, light_dir_x, light_dir_y, light_dir_z, light_intensity;
    float focal_length, caustic_intensity, caustic_sharpness, branch_threshold, ground_y, bubble_radius, time;
    uint32_t _padding;
} sbx_caustic_params;
int32_t sbx_caustic_params_default(sbx_caustic_params *p);
/* CausticRenderer::compute (caustics.rs:441-457) = caustics_compute.wgsl::main (:114-126) on the handle's LATEST
 * thickness buffer (the reference reads the second-to-last one): gradient / Laplacian / Hessian
 * caustic intensity per cell, f32.  host_out (may be NULL) receives batch*rows*num_phi floats; the map also stays
 * on the device (SBX_PTR_CAUSTICS).  grid_width/grid_height in p are overridden by the handle's grid. */
int32_t sbx_drainage_caustics(sbx_drainage *h, const sbx_caustic_params *p, float *host_out);

/* sample_thickness_at_uv / thickness_gradient_uv of the branched-flow ray marcher
 * (src/render/shaders/branched_flow_compute.wgsl:187-210), generalised from its hard-coded 128x64 grid to the
 * handle's grid: nearest-texel thickness (metres, f32) of bubble `bubble` at n uv positions (u = phi/2pi,
 * v = theta/pi; host array of n x 2 floats) and its central-difference gradient with eps = 0.01 in uv space.
 * out_h (n floats) and out_grad (n x 2 floats) are host arrays; either may be NULL.  Whole-sphere handles only. */
int32_t sbx_drainage_sample_uv(sbx_drainage *h, int32_t bubble, const float *uv, size_t n, float *out_h, float *out_grad);

/* sample_thickness(normal) of the same shader (branched_flow_compute.wgsl:110-116, 294-300): thickness * scale at the
 * texel each unit normal (host array of n x 3 floats) points at.  out_h = n floats on the host. */
int32_t sbx_drainage_sample_normal(sbx_drainage *h, int32_t bubble, const float *normals, size_t n, float scale, float *out_h);

/* ---------------------------------------------------------------- headless bubble frame + PNG export (rows f3/f4) */
/* What one frame of HeadlessRenderPipeline shows (src/render/headless.rs:27-39): its Camera (camera.rs:49-60) and the
 * BubbleUniform fields the fragment shader's shading path reads (pipeline.rs:20-112; bubble.wgsl:12-57). */
typedef struct {
    uint32_t width, height;                       /* HeadlessRenderPipeline::new(width, height), headless.rs:50 */
    float cam_distance, cam_yaw, cam_pitch;       /* 0.2, 0, 0.3 (camera.rs:49-60); target = origin */
    float cam_fov_rad;                            /* 45 deg vertical; aspect = width / height */
    float radius;                                 /* sphere radius in metres, 0.025 */
    float base_thickness_nm;                      /* 500 */
    float film_time, swirl_intensity, drainage_speed, pattern_scale; /* 0, 1, 0.5, 1 */
    float interference_intensity;                 /* 4 */
    float base_alpha, edge_alpha;                 /* 0.3, 0.6 */
    uint32_t edge_smoothing_mode;                 /* 0 linear, 1 smoothstep (default), 2 power */
    float background[3];                          /* clear colour (0.1, 0.1, 0.15), linear */
    float lut_max_nm;                             /* LUT_MAX_THICKNESS_NM = 2000 (interference_lut.rs:13) */
} sbx_frame_params;
int32_t sbx_frame_params_default(sbx_frame_params *p);
/* HeadlessRenderPipeline::render_to_buffer (headless.rs:437-512) for the bubble: bubble.wgsl::fs_main (:487-541)
 * evaluated per pixel on `device` -- procedural simplex/FBM film thickness (:200-229), |view . normal|, LUT fetch
 * (linear, clamp-to-edge sampler) x interference_intensity (:425-439), black film below 30 nm (:517-519), edge
 * alpha (:501-513) -- alpha-blended over the background into an Rgba8UnormSrgb target.  The rasterised mesh is an
 * analytic ray-sphere hit here.  lut_rgba: host, A rows of T RGBA8 texels ([angle][thickness]); rgba_out: host,
 * width*height*4 bytes, row 0 at the top, alpha 255. */
int32_t sbx_render_bubble_frame(const sbx_frame_params *p, const uint8_t *lut_rgba, int32_t T, int32_t A,
                                uint8_t *rgba_out, int32_t device);
/* The same frame with the film thickness taken from bubble `bubble` of a drainage handle instead of the procedural
 * noise: the latest thickness field blended at (theta, phi) of the surface normal like get_thickness
 * (drainage.rs:516-552), coloured through the handle's LUT (built for n_film on first use; sbx_drainage_set_lut
 * overrides it).  p->base_thickness_nm and the film-dynamics fields are ignored.  Whole-sphere handles only. */
int32_t sbx_drainage_render_frame(sbx_drainage *h, int32_t bubble, const sbx_frame_params *p, double n_film,
                                  uint8_t *rgba_out);
/* export_frame(path, width, height, data) (src/export/image_export.rs:42-76): RGBA8 -> PNG file (host only, zlib).
 * SBX_ERR_INVALID for zero dimensions (ExportError::InvalidDimensions) or len != width*height*4
 * (ExportError::BufferCreation); SBX_ERR_IO when the file cannot be written (ExportError::SaveError). */
int32_t sbx_export_frame_png(const char *path, uint32_t width, uint32_t height, const uint8_t *rgba, size_t len);

/* ---------------------------------------------------------------- interference LUT / calculator */
/* generate_interference_lut(refractive_index, intensity) -> 256x64 RGBA8
 * (src/render/interference_lut.rs:18-41), built on `device`. */
int32_t sbx_lut_build(float refractive_index, float intensity, uint8_t *out65536, int32_t device);
/* same algorithm (interference_lut.rs:45-93) with n_lambda wavelengths normalised by n_lambda, T x A
 * texels over [0,max_nm] x [0,1] (BASELINE config 5) */
int32_t sbx_lut_build_spectral(float refractive_index, float intensity, const float *wavelengths_nm,
                               int32_t n_lambda, int32_t T, int32_t A, float max_nm, uint8_t *out,
                               int32_t device);
/* thin_film_interference() of the fragment shader (src/render/shaders/bubble.wgsl:425-439) for n arbitrary
 * (thickness_nm, cos_theta) pairs in host memory: bilinear, clamp-to-edge fetch of a T x A RGBA8 LUT through the
 * texture unit at uv = (d/max_nm, cos), times intensity, clamped; rgb_out = n x 3 floats. */
int32_t sbx_lut_shade_points(const uint8_t *lut_rgba, int32_t T, int32_t A, float max_nm, const float *d_nm,
                             const float *cos_theta, size_t n, float intensity, float *rgb_out, int32_t device);
/* InterferenceCalculator (src/physics/interference.rs:127-513) evaluated in f64 on the device for n
 * (thickness_nm, cos_theta_incident) pairs held in host memory; rgb_out = n x 3 doubles.
 * gamma != 0: calculate_interference_color (:380-400); gamma == 0: _linear (:413-426). */
int32_t sbx_interference_eval(double n_film, const double wavelengths_nm[3], const double *d_nm,
                              const double *cos_i, size_t n, int32_t gamma, double *rgb_out,
                              int32_t device);
/* per-wavelength pieces of the calculator for n host-resident inputs (interference.rs:197-353):
 * out_cos_t[n], out_fresnel[n][3] = (Rs,Rp,avg), out_intensity[n] at wavelength_nm; any may be NULL */
int32_t sbx_interference_terms(double n_film, const double *d_nm, const double *cos_i, size_t n,
                               double wavelength_nm, double *out_cos_t, double *out_fresnel,
                               double *out_intensity, int32_t device);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU) */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the caller (torch.distributed) */
int32_t sbx_comm_unique_id(uint8_t out128[128]);
int32_t sbx_comm_init(const uint8_t id128[128], int32_t nranks, int32_t rank, int32_t device, sbx_comm **out);
int32_t sbx_comm_destroy(sbx_comm *c);
/* attach a communicator to a sharded handle: every step then exchanges one halo row with the
 * latitude-band neighbours (rank-1 above, rank+1 below).  From then on the calls that change or step the field --
 * sbx_drainage_upload / _upload_concentration / _reset / _step / _step_fused / _step_host -- and the ones that read
 * ghost rows (_caustics) are collective over the ranks: every rank makes the same calls in the same order, as with
 * any NCCL program (they contain the neighbour exchanges).  Queries (_download, _get_thickness, _minmax, _shade,
 * _frame_stats, ...) are local. */
int32_t sbx_drainage_attach_comm(sbx_drainage *h, sbx_comm *c);
/* Peer-memory halo (all ranks on one NVLink/NVSwitch box).  Each rank exports a descriptor of its field
 * buffers (CUDA IPC handles), the caller moves the blobs between processes (torch.distributed), and attach maps
 * the neighbours' buffers.  From then on the ring kernels (f32 and f64 CPU semantics, WGSL semantics with both of
 * its fields) read the row next to their band straight out of the neighbour's field from inside the stencil
 * kernel (a bulk copy from the peer address over NVLink) and order neighbouring steps with two flags in device
 * memory -- no per-step collective and no remote stores.  Requires an attached communicator (used for the
 * exchanges around upload / reset / the host-buffer step and for the modes the ring kernels do not cover: strict
 * f64, ragged widths, float colour output).  A neighbour may read this band's boundary rows until its own step has
 * finished: keep the handle alive until every rank is done.  up/down = blob of rank-1 / rank+1 or NULL. */
#define SBX_IPC_BYTES 512
int32_t sbx_drainage_ipc_export(sbx_drainage *h, uint8_t out[SBX_IPC_BYTES]);
int32_t sbx_drainage_ipc_attach(sbx_drainage *h, const uint8_t *up, const uint8_t *down);
/* Device-side wait status since the last query, a bit mask: 0 = ok, bit 0 = a wait for a latitude neighbour's step
 * flag timed out, bit 1 = a wait for the other strips of a pole ring (the folded pole rows) timed out.  Every
 * in-kernel spin is bounded, so a lost neighbour or a failed launch shows up here instead of hanging the GPU. */
int32_t sbx_drainage_peer_status(sbx_drainage *h, int32_t *out);
/* global min/max over all ranks' bands */
int32_t sbx_drainage_minmax_global(sbx_drainage *h, double *min_out, double *max_out);

#ifdef __cplusplus
}
#endif
#endif /* SBX_H */
