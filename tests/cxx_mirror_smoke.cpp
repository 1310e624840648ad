// Compiles the C++ mirror (include/sbx.hpp) against libsbx.so and, when a GPU is present, replays a few of
// the reference's own unit tests (drainage.rs:636-719, interference.rs:649-686) through it.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../include/sbx.hpp"

int main()
{
    int32_t ndev = 0;
    if (sbx_device_count(&ndev) != SBX_OK || ndev == 0) {
        try {
            sbx::DrainageSimulator sim{sbx::SimulationConfig{}};
            std::printf("FAIL: expected an error without a device\n");
            return 1;
        } catch (const sbx::Error &e) {
            std::printf("no device: %s\nCXX_MIRROR_NO_GPU_OK\n", e.what());
            return 0;
        }
    }
    sbx::SimulationConfig cfg;
    sbx::DrainageSimulator sim(cfg);
    if (std::fabs(sim.current_time()) > 1e-16 || sim.has_critical_region()) return 2;
    const double top0 = sim.get_thickness(0.1, 0.0);
    for (int i = 0; i < 100; ++i) sim.step(cfg.dt);
    if (!(sim.get_thickness(0.1, 0.0) < top0)) return 3;                        // drainage.rs:668-688
    if (std::fabs(sim.current_time() - 0.1) > 1e-12) return 4;
    if (std::fabs(sim.get_thickness(1.0, 0.0) - sim.get_thickness(1.0, 2.0 * M_PI)) > 1e-12) return 5;  // :691-702
    const sbx::ThicknessField &f = sim.thickness_field();
    if (f.num_theta() != 128 || f.num_phi() != 256 || !(f.get(64, 0) < 500e-9)) return 6;
    sim.reset(600e-9);
    if (std::fabs(sim.get_thickness(0.5, 0.5) - 600e-9) > 1e-15) return 7;      // :705-719
    sbx::InterferenceCalculator ic(1.33);
    if (std::fabs(ic.calculate_transmission_angle_cos(1.0) - 1.0) > 1e-6) return 8;  // interference.rs:649-654
    if (ic.calculate_reflected_intensity(0.0, 1.0, 550.0) != 0.0) return 9;          // :682-686
    auto c = ic.calculate_interference_color(1000.0, 0.0).to_u8_tuple();
    if (c[0] != 255 || c[1] != 255 || c[2] != 255) return 10;
    if (sbx::generate_interference_lut(1.33f, 1.0f).size() != 65536) return 11;      // interference_lut.rs:170-174
    sbx::GPUDrainageSimulator g(500e-9f, 128, 64);
    g.enabled = true;
    g.step(1.0f / 60.0f);
    if (std::fabs(g.current_time() - 100.0 / 60.0) > 1e-5) return 12;
    // one multi-step call equals the same number of single steps (both paths in f32 offset storage: pole means may
    // differ in the last bit of their f64 sums, nothing else)
    sbx::DrainageSimulator a(cfg), b(cfg);
    a.step(cfg.dt, 50);
    for (int i = 0; i < 50; ++i) b.step(cfg.dt);
    if (std::fabs(a.current_time() - b.current_time()) > 0.0) return 13;
    if (std::fabs(a.get_thickness(0.7, 1.3) / b.get_thickness(0.7, 1.3) - 1.0) > 1e-6) return 14;
    const std::vector<float> cm = a.caustic_map();
    if (cm.size() != 128u * 256u || !(std::fabs(cm[5000] - 2.0f) < 1e-3f)) return 15;  // flat film: intensity 2.0
    std::vector<float> hh, gg;
    a.sample_uv({0.25f, 0.5f, 0.75f, 0.1f}, hh, gg);
    if (hh.size() != 2 || !(hh[0] > 400e-9f && hh[0] < 500.1e-9f)) return 16;
    sbx::HeadlessRenderPipeline hp(64, 64);                                             // headless.rs:50
    hp.set_thickness(500.0f);
    hp.set_time(0.0f);
    const std::vector<uint8_t> px = hp.render_to_buffer();
    if (px.size() != 64u * 64u * 4u || px[0] != 89 || px[1] != 89 || px[2] != 108 || px[3] != 255) return 17;  // background
    const uint8_t *mid = &px[(32u * 64u + 32u) * 4u];
    if (mid[0] == 89 && mid[1] == 89 && mid[2] == 108) return 18;                       // the bubble covers the centre
    sbx::export_frame("/tmp/sbx_cxx_mirror.png", 64, 64, px);                           // image_export.rs:84-98
    try {
        sbx::export_frame("/tmp/sbx_cxx_mirror.png", 0, 100, {});                       // :100-104
        return 19;
    } catch (const sbx::Error &e) {
        if (e.code != SBX_ERR_INVALID) return 20;
    }
    std::printf("CXX_MIRROR_GPU_OK\n");
    return 0;
}
