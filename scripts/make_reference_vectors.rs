//! make_reference_vectors.rs -- golden vectors of `DrainageSimulator::step` and `InterferenceCalculator` written by
//! the REFERENCE ITSELF (pjt222/soap_bubble_simulation), for pinning libsbx's oracle and its strict f64 mode.
//!
//! The sbx build image has no Rust toolchain, so these vectors cannot be produced there.  On any machine with cargo:
//!
//!     cp scripts/make_reference_vectors.rs  <reference checkout>/tests/make_reference_vectors.rs
//!     cd <reference checkout> && cargo test --release --test make_reference_vectors -- --nocapture
//!     cp target/sbx_vectors/*.bin  <sbx checkout>/tests/golden/
//!
//! (debug and release builds give the same bits: rustc never contracts a*b+c into an FMA and has no fast-math.)
//! `tests/test_oracle_reference_vectors.py` (CPU) then checks the C oracle bit for bit against every checkpoint and
//! `tests/test_gpu_round2.py::test_strict_mode_reproduces_reference_vectors` (GPU) checks SBX_PREC_F64_STRICT.
//!
//! Only the reference's public API is used: `DrainageSimulator::{new, step, thickness_field, current_time,
//! active_tfe_count}` (src/physics/drainage.rs:263,298,555,560,611) and
//! `InterferenceCalculator::{new, calculate_interference_color, calculate_interference_color_linear}`
//! (src/physics/interference.rs:141,380,413).  The thickness field cannot be set from outside, so every drainage
//! case starts from the uniform field `new` builds; gravity and the TFE fronts make it non-uniform.
//!
//! File format (little endian), read by tests/reference_vectors.py:
//!   drainage_<name>.bin:  b"SBXVEC01" | u32 resolution | u32 num_theta | u32 num_phi | f64 film_thickness_nm
//!                         | f64 diameter | f64 critical_thickness_nm | f64 density | f64 viscosity
//!                         | f64 surfactant_diffusion | f64 gravity | f64 dt | u32 tfe_enabled | u32 n_checkpoints
//!                         | n_checkpoints x { u64 steps_done | f64 current_time | u32 active_tfe_count | u32 0
//!                                             | f64 x num_theta*num_phi }
//!   interference_<name>.bin:  b"SBXINT01" | f64 n_film | u32 n_points | u32 0
//!                         | n_points x { f64 d_nm | f64 cos_i | f64 r g b (gamma) | f64 r g b (linear) }
use std::fs::{create_dir_all, File};
use std::io::{BufWriter, Write};

use soap_bubble_sim::config::SimulationConfig;
use soap_bubble_sim::physics::drainage::DrainageSimulator;
use soap_bubble_sim::physics::interference::InterferenceCalculator;

fn out_dir() -> std::path::PathBuf {
    let d = std::path::Path::new(env!("CARGO_MANIFEST_DIR")).join("target").join("sbx_vectors");
    create_dir_all(&d).unwrap();
    d
}

struct Case {
    name: &'static str,
    resolution: u32,
    film_thickness_nm: f64,
    dt: f64,
    tfe: bool,
    checkpoints: &'static [u64],
}

fn write_drainage(case: &Case) {
    let mut config = SimulationConfig::default();
    config.resolution = case.resolution;
    config.bubble.film_thickness_nm = case.film_thickness_nm;
    let mut sim = DrainageSimulator::new(&config);
    sim.set_marginal_regeneration(case.tfe);
    let (nt, np) = (sim.thickness_field().num_theta(), sim.thickness_field().num_phi());
    let mut w = BufWriter::new(File::create(out_dir().join(format!("drainage_{}.bin", case.name))).unwrap());
    w.write_all(b"SBXVEC01").unwrap();
    for v in [case.resolution, nt as u32, np as u32] {
        w.write_all(&v.to_le_bytes()).unwrap();
    }
    for v in [
        config.bubble.film_thickness_nm,
        config.bubble.diameter,
        config.bubble.critical_thickness_nm,
        config.fluid.density,
        config.fluid.viscosity,
        config.fluid.surfactant_diffusion,
        config.environment.gravity,
        case.dt,
    ] {
        w.write_all(&v.to_le_bytes()).unwrap();
    }
    w.write_all(&(case.tfe as u32).to_le_bytes()).unwrap();
    w.write_all(&(case.checkpoints.len() as u32).to_le_bytes()).unwrap();
    let mut done: u64 = 0;
    for &target in case.checkpoints {
        while done < target {
            sim.step(case.dt);
            done += 1;
        }
        w.write_all(&done.to_le_bytes()).unwrap();
        w.write_all(&sim.current_time().to_le_bytes()).unwrap();
        w.write_all(&(sim.active_tfe_count() as u32).to_le_bytes()).unwrap();
        w.write_all(&0u32.to_le_bytes()).unwrap();
        for v in sim.thickness_field().data() {
            w.write_all(&v.to_le_bytes()).unwrap();
        }
    }
    println!("wrote drainage_{}.bin: {}x{}, {} checkpoints, {} steps", case.name, nt, np, case.checkpoints.len(), done);
}

#[test]
fn drainage_vectors() {
    let cases = [
        // config/default.json: 128 x 256, 500 nm, dt = 1e-3, TFE on (BASELINE config 1)
        Case { name: "default_128", resolution: 128, film_thickness_nm: 500.0, dt: 1e-3, tfe: true, checkpoints: &[1, 10, 100, 1000] },
        // fronts spawn, overlap and retire: 6000 steps of 5 ms on a small grid
        Case { name: "fronts_32", resolution: 32, film_thickness_nm: 500.0, dt: 5e-3, tfe: true, checkpoints: &[100, 1000, 3000, 6000] },
        // a thick film, where the gravity term moves the field visibly (increments far above f32 resolution)
        Case { name: "thick_64", resolution: 64, film_thickness_nm: 50000.0, dt: 1e-3, tfe: false, checkpoints: &[1, 50, 500] },
        // below-critical cells: starts at 12 nm (critical 10 nm), the fronts push cells under the threshold
        Case { name: "critical_48", resolution: 48, film_thickness_nm: 12.0, dt: 2e-2, tfe: true, checkpoints: &[50, 500, 2500] },
    ];
    for c in &cases {
        write_drainage(c);
    }
}

#[test]
fn interference_vectors() {
    for (name, n_film) in [("n133", 1.33f64), ("n150", 1.5f64)] {
        let calc = InterferenceCalculator::new(n_film);
        let mut pts: Vec<(f64, f64)> = Vec::new();
        for i in 0..=400 {
            let d = 5.0 * i as f64; // 0 .. 2000 nm
            for j in 0..=20 {
                pts.push((d, j as f64 / 20.0));
            }
        }
        pts.push((500.0, 1.0)); // exactly on a green zero at n = 1.33
        pts.push((1000.0, 0.0));
        let mut w = BufWriter::new(File::create(out_dir().join(format!("interference_{}.bin", name))).unwrap());
        w.write_all(b"SBXINT01").unwrap();
        w.write_all(&n_film.to_le_bytes()).unwrap();
        w.write_all(&(pts.len() as u32).to_le_bytes()).unwrap();
        w.write_all(&0u32.to_le_bytes()).unwrap();
        for (d, c) in &pts {
            let g = calc.calculate_interference_color(*d, *c);
            let l = calc.calculate_interference_color_linear(*d, *c);
            for v in [*d, *c, g.red, g.green, g.blue, l.red, l.green, l.blue] {
                w.write_all(&v.to_le_bytes()).unwrap();
            }
        }
        println!("wrote interference_{}.bin: {} points", name, pts.len());
    }
}
