// sbx.hpp -- header-only C++17 mirror of the reference's Rust API over the C ABI in sbx.h.
//
// Same names, argument meaning and error behaviour as src/physics/drainage.rs, src/render/gpu_drainage.rs,
// src/physics/interference.rs and src/render/interference_lut.rs (snake_case kept so that the reference's
// tests read the same).  The Rust hot path is infallible; here CUDA failures surface as sbx::Error.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "sbx.h"

namespace sbx {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error("libsbx error " + std::to_string(c) + ": " + m), code(c) {}
};

inline void check(int32_t rc)
{
    if (rc != SBX_OK) throw Error(rc, sbx_last_error());
}

// src/config.rs:283-371 -- the fields DrainageSimulator::new reads (drainage.rs:263-288)
struct SimulationConfig {
    double diameter = 0.05, film_thickness_nm = 500.0, critical_thickness_nm = 10.0, refractive_index = 1.33;
    double viscosity = 0.001, surface_tension = 0.025, density = 1000.0, surfactant_diffusion = 1e-9;
    double gravity = 9.81;
    double dt = 0.001;
    uint32_t resolution = 128;
    double bubble_radius() const { return diameter / 2.0; }                    // config.rs:357-359
    double film_thickness_meters() const { return film_thickness_nm * 1e-9; }  // config.rs:368-370
};

// ThicknessField (drainage.rs:96-204): host mirror of the device field, refreshed on demand
class ThicknessField {
public:
    ThicknessField(size_t num_theta, size_t num_phi, double initial_thickness)
        : num_theta_(num_theta), num_phi_(num_phi), data_(num_theta * num_phi, initial_thickness),
          delta_theta_(M_PI / double(num_theta - 1)), delta_phi_(2.0 * M_PI / double(num_phi)) {}
    double get(size_t ti, size_t pj) const { return data_[ti * num_phi_ + pj]; }
    void set(size_t ti, size_t pj, double v) { data_[ti * num_phi_ + pj] = v; }
    double theta_at(size_t ti) const { return double(ti) * delta_theta_; }
    double phi_at(size_t pj) const { return double(pj) * delta_phi_; }
    size_t num_theta() const { return num_theta_; }
    size_t num_phi() const { return num_phi_; }
    double delta_theta() const { return delta_theta_; }
    double delta_phi() const { return delta_phi_; }
    const std::vector<double> &data() const { return data_; }
    std::vector<double> &data_mut() { return data_; }

private:
    size_t num_theta_, num_phi_;
    std::vector<double> data_;
    double delta_theta_, delta_phi_;
};

// DrainageSimulator (drainage.rs:212-614)
class DrainageSimulator {
public:
    explicit DrainageSimulator(const SimulationConfig &c, int32_t precision = SBX_PREC_F32, int32_t device = 0)
        : mirror_(c.resolution, size_t(c.resolution) * 2, c.film_thickness_meters())
    {
        sbx_drainage_cfg cfg;
        check(sbx_drainage_cfg_default(&cfg));
        cfg.num_theta = int32_t(c.resolution);
        cfg.num_phi = int32_t(c.resolution) * 2;  // drainage.rs:265
        cfg.h0_m = c.film_thickness_meters();
        cfg.density = c.density;
        cfg.gravity = c.gravity;
        cfg.viscosity = c.viscosity;
        cfg.diffusion = c.surfactant_diffusion;
        cfg.critical_thickness_m = c.critical_thickness_nm * 1e-9;
        cfg.radius_m = c.bubble_radius();
        cfg.precision = precision;
        cfg.h_ref_m = cfg.h0_m;  // offset storage keeps the default configuration's sub-ulp increments
        check(sbx_drainage_create(&cfg, device, &h_));
    }
    ~DrainageSimulator() { sbx_drainage_destroy(h_); }
    DrainageSimulator(const DrainageSimulator &) = delete;
    DrainageSimulator &operator=(const DrainageSimulator &) = delete;

    void step(double dt) { check(sbx_drainage_step(h_, dt, 1)); stale_ = true; }                    // :298-392
    // n applications of step(dt) in one call: on grids that fit in shared memory (the default 128 x 256 does) all n
    // steps run inside one kernel launch
    void step(double dt, int n) { check(sbx_drainage_step(h_, dt, n)); stale_ = true; }
    // thickness-buffer consumers of the frame (pipeline.rs:1808-1818): caustic map (caustics_compute.wgsl:29-126) and
    // the branched-flow shader's thickness sampling (branched_flow_compute.wgsl:187-210), both on the latest buffer
    std::vector<float> caustic_map(const sbx_caustic_params *p = nullptr)
    {
        sbx_caustic_params q;
        check(sbx_caustic_params_default(&q));
        if (p) q = *p;
        std::vector<float> out(mirror_.num_theta() * mirror_.num_phi());
        check(sbx_drainage_caustics(h_, &q, out.data()));
        return out;
    }
    void sample_uv(const std::vector<float> &uv, std::vector<float> &h, std::vector<float> &grad) const
    {
        const size_t n = uv.size() / 2;
        h.resize(n);
        grad.resize(2 * n);
        check(sbx_drainage_sample_uv(h_, 0, uv.data(), n, h.data(), grad.data()));
    }
    void step_fused(double dt, const sbx_view &v) { check(sbx_drainage_step_fused(h_, dt, 1, &v)); stale_ = true; }
    double get_thickness(double theta, double phi) const                                            // :516-552
    {
        double out;
        check(sbx_drainage_get_thickness(h_, 0, theta, phi, &out));
        return out;
    }
    const ThicknessField &thickness_field()                                                         // :555-557
    {
        if (stale_) { check(sbx_drainage_download(h_, mirror_.data_mut().data(), SBX_F64)); stale_ = false; }
        return mirror_;
    }
    double min_thickness() const { double a, b; check(sbx_drainage_minmax(h_, &a, &b)); return a; } // :180-185 on device
    double max_thickness() const { double a, b; check(sbx_drainage_minmax(h_, &a, &b)); return b; } // :188-193 on device
    double current_time() const { double t; check(sbx_drainage_current_time(h_, &t)); return t; }   // :560-562
    bool has_critical_region() const { int32_t f; check(sbx_drainage_has_critical(h_, &f)); return f != 0; }  // :565-567
    double critical_thickness() const { double t; check(sbx_drainage_critical_thickness(h_, &t)); return t; }
    void reset(double h0) { check(sbx_drainage_reset(h_, h0)); stale_ = true; }                      // :575-583
    void set_marginal_regeneration(bool on) { check(sbx_drainage_set_marginal_regeneration(h_, on)); }
    bool marginal_regeneration_enabled() const { int32_t f; check(sbx_drainage_marginal_regeneration(h_, &f)); return f != 0; }
    void set_tfe_thickness_ratio(double r) { check(sbx_drainage_set_tfe_ratio(h_, r)); }
    double tfe_thickness_ratio() const { double r; check(sbx_drainage_tfe_ratio(h_, &r)); return r; }
    size_t active_tfe_count() const { int32_t n; check(sbx_drainage_active_tfe_count(h_, &n)); return size_t(n); }
    sbx_drainage *handle() { return h_; }

private:
    sbx_drainage *h_ = nullptr;
    ThicknessField mirror_;
    bool stale_ = true;
};

// GPUDrainageSimulator (gpu_drainage.rs:81-491) without the wgpu device/queue/encoder arguments
class GPUDrainageSimulator {
public:
    bool enabled = false;          // gpu_drainage.rs:326
    float time_scale = 100.0f;     // :327
    uint32_t steps_per_frame = 10; // :328
    bool marangoni_enabled = false;

    GPUDrainageSimulator(float initial_thickness, uint32_t grid_width, uint32_t grid_height, int32_t device = 0)
    {
        sbx_drainage_cfg cfg;
        check(sbx_drainage_cfg_default(&cfg));
        cfg.num_phi = int32_t(grid_width);
        cfg.num_theta = int32_t(grid_height);
        cfg.h0_m = initial_thickness;
        cfg.semantics = SBX_SEM_WGSL;
        cfg.h_ref_m = 0.0;
        params_ = cfg.wgsl;
        params_.grid_width = grid_width;
        params_.grid_height = grid_height;
        cfg.wgsl = params_;
        check(sbx_drainage_create(&cfg, device, &h_));
    }
    ~GPUDrainageSimulator() { sbx_drainage_destroy(h_); }
    GPUDrainageSimulator(const GPUDrainageSimulator &) = delete;
    GPUDrainageSimulator &operator=(const GPUDrainageSimulator &) = delete;

    void step(float dt)  // gpu_drainage.rs:338-370
    {
        if (!enabled) return;
        const float step_dt = dt * time_scale / float(steps_per_frame);
        check(sbx_drainage_step(h_, double(step_dt), int32_t(steps_per_frame)));
    }
    void update_params(const sbx_drainage_params &p) { params_ = p; check(sbx_drainage_update_params(h_, &params_)); }
    void set_dt(float dt) { params_.dt = dt; update_params(params_); }
    void set_physics(float viscosity, float density, float gravity, float diffusion)
    {
        params_.viscosity = viscosity; params_.density = density; params_.gravity = gravity; params_.diffusion_coeff = diffusion;
        update_params(params_);
    }
    void set_bubble_radius(float r) { params_.bubble_radius = r; update_params(params_); }
    void reset(float h0) { check(sbx_drainage_reset(h_, h0)); }
    void set_marangoni_enabled(bool on) { marangoni_enabled = on; params_.marangoni_enabled = on ? 1u : 0u; update_params(params_); }
    void set_marangoni_params(float gamma_air, float gamma_reduction, float surfactant_diffusion, float marangoni_coeff)
    {
        params_.gamma_air = gamma_air; params_.gamma_reduction = gamma_reduction;
        params_.surfactant_diffusion = surfactant_diffusion; params_.marangoni_coeff = marangoni_coeff;
        update_params(params_);
    }
    void *concentration_buffer() const { void *p; check(sbx_drainage_device_ptr(h_, SBX_PTR_CONCENTRATION, &p)); return p; }
    void *current_buffer() const { void *p; check(sbx_drainage_device_ptr(h_, SBX_PTR_THICKNESS, &p)); return p; }
    void *current_thickness_buffer() const { return current_buffer(); }
    double current_time() const { double t; check(sbx_drainage_current_time(h_, &t)); return t; }
    std::pair<uint32_t, uint32_t> grid_size() const { return {params_.grid_width, params_.grid_height}; }
    const sbx_drainage_params &params() const { return params_; }
    sbx_drainage *handle() { return h_; }

private:
    sbx_drainage *h_ = nullptr;
    sbx_drainage_params params_{};
};

// interference_lut.rs:9-41
constexpr uint32_t LUT_THICKNESS_SAMPLES = 256;
constexpr uint32_t LUT_ANGLE_SAMPLES = 64;
constexpr float LUT_MAX_THICKNESS_NM = 2000.0f;

inline std::vector<uint8_t> generate_interference_lut(float refractive_index, float intensity, int32_t device = 0)
{
    std::vector<uint8_t> out(size_t(LUT_THICKNESS_SAMPLES) * LUT_ANGLE_SAMPLES * 4);
    check(sbx_lut_build(refractive_index, intensity, out.data(), device));
    return out;
}

// ---------------------------------------------------------------- headless frame + PNG export (rows f3/f4)
// src/render/headless.rs:27-552 (bubble only), src/render/camera.rs:28-105, src/export/image_export.rs:42-76
struct Camera {
    float distance = 0.2f, yaw = 0.0f, pitch = 0.3f, fov = 45.0f * float(M_PI / 180.0), aspect = 1.0f;
    void orbit(float dx, float dy)  // camera.rs:71-78
    {
        yaw += dx * 0.01f;
        pitch += dy * 0.01f;
        const float max_pitch = float(M_PI) / 2.0f - 0.01f;
        pitch = std::fmin(std::fmax(pitch, -max_pitch), max_pitch);
    }
    void zoom(float delta)  // camera.rs:81-84
    {
        distance *= 1.0f - delta * 0.1f;
        distance = std::fmin(std::fmax(distance, 0.05f), 10.0f);
    }
};

class HeadlessRenderPipeline {
public:
    Camera camera;
    sbx_frame_params bubble_uniform{};  // base_thickness_nm, film_time, ... (BubbleUniform::default, pipeline.rs:69-112)

    HeadlessRenderPipeline(uint32_t width, uint32_t height, int32_t device = 0) : device_(device)
    {
        check(sbx_frame_params_default(&bubble_uniform));
        bubble_uniform.width = width;
        bubble_uniform.height = height;
        camera.aspect = float(width) / float(height);
        lut_ = generate_interference_lut(1.33f, 1.0f, device);  // headless.rs:143-146
    }
    std::array<uint32_t, 2> size() const { return {bubble_uniform.width, bubble_uniform.height}; }
    void set_thickness(float nm) { bubble_uniform.base_thickness_nm = nm; }
    void set_time(float t) { bubble_uniform.film_time = t; }
    void set_camera_distance(float d) { camera.distance = std::fmin(std::fmax(d, 0.05f), 10.0f); }
    void orbit_camera(float dx, float dy) { camera.orbit(dx, dy); }
    void zoom_camera(float delta) { camera.zoom(delta); }
    std::vector<uint8_t> render_to_buffer()  // headless.rs:437-512
    {
        sbx_frame_params p = bubble_uniform;
        p.cam_distance = camera.distance;
        p.cam_yaw = camera.yaw;
        p.cam_pitch = camera.pitch;
        p.cam_fov_rad = camera.fov;
        std::vector<uint8_t> out(size_t(p.width) * p.height * 4);
        check(sbx_render_bubble_frame(&p, lut_.data(), LUT_THICKNESS_SAMPLES, LUT_ANGLE_SAMPLES, out.data(), device_));
        return out;
    }

private:
    int32_t device_;
    std::vector<uint8_t> lut_;
};

// export_frame, image_export.rs:42-76 (throws sbx::Error: SBX_ERR_INVALID for InvalidDimensions / BufferCreation,
// SBX_ERR_IO for SaveError)
inline void export_frame(const std::string &path, uint32_t width, uint32_t height, const std::vector<uint8_t> &data)
{
    check(sbx_export_frame_png(path.c_str(), width, height, data.data(), data.size()));
}

struct RgbColor {  // interference.rs:43-95
    double red = 0, green = 0, blue = 0;
    static RgbColor make(double r, double g, double b)
    {
        auto c = [](double x) { return x < 0 ? 0.0 : (x > 1 ? 1.0 : x); };
        return {c(r), c(g), c(b)};
    }
    std::array<uint8_t, 3> to_u8_tuple() const
    {
        return {uint8_t(std::round(red * 255.0)), uint8_t(std::round(green * 255.0)), uint8_t(std::round(blue * 255.0))};
    }
};

// InterferenceCalculator (interference.rs:127-513), evaluated in f64 on the device
class InterferenceCalculator {
public:
    explicit InterferenceCalculator(double n_film = 1.33, int32_t device = 0) : n_(n_film), device_(device) {}
    double refractive_index_film() const { return n_; }
    double refractive_index_medium() const { return 1.0; }
    RgbColor calculate_interference_color(double d_nm, double cos_i) const { return eval(d_nm, cos_i, 1); }         // :380-400
    RgbColor calculate_interference_color_linear(double d_nm, double cos_i) const { return eval(d_nm, cos_i, 0); }  // :413-426
    double calculate_reflected_intensity(double d_nm, double cos_i, double wavelength_nm) const                      // :312-353
    {
        double out;
        check(sbx_interference_terms(n_, &d_nm, &cos_i, 1, wavelength_nm, nullptr, nullptr, &out, device_));
        return out;
    }
    double calculate_transmission_angle_cos(double cos_i) const  // :197-211
    {
        double d = 0.0, out;
        check(sbx_interference_terms(n_, &d, &cos_i, 1, 550.0, &out, nullptr, nullptr, device_));
        return out;
    }
    std::vector<RgbColor> generate_color_lookup_table(double lo, double hi, size_t n, double cos_theta) const  // :461-477
    {
        std::vector<double> d(n), c(n, cos_theta), rgb(3 * n);
        const double step = (hi - lo) / double(n - 1);
        for (size_t i = 0; i < n; ++i) d[i] = lo + double(i) * step;
        check(sbx_interference_eval(n_, wl_, d.data(), c.data(), n, 1, rgb.data(), device_));
        std::vector<RgbColor> out(n);
        for (size_t i = 0; i < n; ++i) out[i] = RgbColor::make(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        return out;
    }

private:
    RgbColor eval(double d_nm, double cos_i, int gamma) const
    {
        double rgb[3];
        check(sbx_interference_eval(n_, wl_, &d_nm, &cos_i, 1, gamma, rgb, device_));
        return RgbColor::make(rgb[0], rgb[1], rgb[2]);
    }
    double n_;
    int32_t device_;
    double wl_[3] = {650.0, 532.0, 450.0};  // interference.rs:26-30
};

}  // namespace sbx
