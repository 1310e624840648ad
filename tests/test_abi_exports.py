"""The C-ABI library loads without a GPU and exports every symbol include/sbx.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "sbx.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from soap_bubble_simulation_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from soap_bubble_simulation_b200 import _lib

    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_rust_binding_source_covers_the_header():
    """bindings/rust/src/sbx_sys.rs (source only: no rustc in this image) declares every entry point of the header
    with the same number of arguments."""
    with open(os.path.join(ROOT, "bindings", "rust", "src", "sbx_sys.rs")) as f:
        rs = f.read()
    rust = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (sbx_[a-z0-9_]+)\s*\(([^)]*)\)", rs, flags=re.S)}
    assert sorted(rust) == declared_symbols()
    with open(os.path.join(ROOT, "include", "sbx.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    for name, args in re.findall(r"\b(sbx_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        n_c = 0 if args.strip() == "void" else args.count(",") + 1
        n_rs = 0 if not rust[name].strip() else rust[name].count(",") + 1
        assert n_c == n_rs, (name, n_c, n_rs)


def test_struct_layouts_match_the_reference_pod():
    """DrainageParams is a 64-byte #[repr(C)] POD (gpu_drainage.rs:10-44; pipeline.rs has the same style of
    size/offset checks for BubbleUniform at :2918-2948)."""
    from soap_bubble_simulation_b200 import _lib

    assert ctypes.sizeof(_lib.DrainageParamsC) == 64
    assert _lib.DrainageParamsC.grid_width.offset == 28
    assert _lib.DrainageParamsC.marangoni_coeff.offset == 52
    assert ctypes.sizeof(_lib.TfeFrontC) == 32


def test_no_device_is_a_loud_error_not_a_fallback():
    import pytest
    import torch

    import soap_bubble_simulation_b200 as S

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(S.SbxError) as ei:
        S.DrainageSimulator(S.SimulationConfig.default())
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(S.SbxError):
        S.generate_interference_lut(1.33, 1.0)
    with pytest.raises(S.SbxError):
        S.InterferenceCalculator(1.33).calculate_interference_color(500.0, 1.0)


def test_host_side_config_mirror(tmp_path):
    """SimulationConfig JSON round trip (config.rs:328-354) and derived values (config.rs:356-371)."""
    import soap_bubble_simulation_b200 as S

    c = S.SimulationConfig.default()
    assert c.resolution == 128 and c.dt == 0.001 and c.bubble.refractive_index == 1.33
    assert c.bubble_radius() == 0.025
    assert c.film_thickness_meters() == 500.0 * 1e-9
    assert abs(c.internal_pressure_difference() - 4.0) < 1e-12
    p = tmp_path / "cfg.json"
    c.bubble.film_thickness_nm = 321.0
    c.to_file(p)
    d = S.SimulationConfig.from_file(p)
    assert d.bubble.film_thickness_nm == 321.0 and d.fluid.viscosity == 0.001
    import pytest

    with pytest.raises(S.ConfigError):
        S.SimulationConfig.from_file(tmp_path / "missing.json")
    (tmp_path / "bad.json").write_text("{not json")
    with pytest.raises(S.ConfigError):
        S.SimulationConfig.from_file(tmp_path / "bad.json")


def test_rgb_color_and_camera_host_helpers():
    import soap_bubble_simulation_b200 as S

    assert S.RgbColor.new(1.5, -0.5, 0.5) == S.RgbColor(1.0, 0.0, 0.5)  # interference.rs:619-624
    assert S.RgbColor.new(1.0, 0.5, 0.0).to_u8_tuple() == (255, 128, 0)  # interference.rs:627-633
    x, y, z = S.camera_position(0.2, 0.0, 0.3)  # camera.rs:49-68
    assert abs(x) < 1e-12 and abs(y - 0.0591040) < 1e-6 and abs(z - 0.1910673) < 1e-6


def test_rust_build_script_lists_every_source_of_the_makefile():
    """bindings/rust/build.rs (source only) must compile the same .cu files csrc/Makefile links into libsbx.so."""
    with open(os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc", "Makefile")) as f:
        srcs = re.search(r"^SRCS\s*:=\s*(.+)$", f.read(), flags=re.M).group(1).split()
    with open(os.path.join(ROOT, "bindings", "rust", "build.rs")) as f:
        rs = f.read()
    assert srcs and all(f'"{s}"' in rs for s in srcs), [s for s in srcs if f'"{s}"' not in rs]
