"""export_frame (src/export/image_export.rs:42-76) behind sbx_export_frame_png: the reference's three unit tests
(:84-110) plus a decode round trip with an independent PNG reader.  Host-only code: runs without a GPU."""
import os

import numpy as np
import pytest
from PIL import Image

import soap_bubble_simulation_b200 as S


def test_export_frame(tmp_path):  # image_export.rs:84-98
    data = bytes([255, 0, 0, 255] * 4)
    path = tmp_path / "test_export.png"
    S.export_frame(path, 2, 2, data)
    img = Image.open(path)
    assert img.size == (2, 2) and img.mode == "RGBA"
    assert np.array_equal(np.array(img).reshape(-1), np.frombuffer(data, np.uint8))


def test_invalid_dimensions(tmp_path):  # image_export.rs:100-104
    with pytest.raises(S.ExportError) as ei:
        S.export_frame(tmp_path / "test.png", 0, 100, b"")
    assert ei.value.kind == "InvalidDimensions"
    assert not os.path.exists(tmp_path / "test.png")


def test_wrong_data_length(tmp_path):  # image_export.rs:106-110
    with pytest.raises(S.ExportError) as ei:
        S.export_frame(tmp_path / "test.png", 10, 10, bytes(100))
    assert ei.value.kind == "BufferCreation"
    assert "100" in str(ei.value) and "400" in str(ei.value)


def test_save_error_is_reported(tmp_path):
    with pytest.raises(S.ExportError) as ei:
        S.export_frame(tmp_path / "no_such_dir" / "x.png", 1, 1, bytes(4))
    assert ei.value.kind == "SaveError"


@pytest.mark.parametrize("shape", [(1, 1), (37, 53), (256, 256), (3, 1025)])
def test_round_trip_through_an_independent_decoder(tmp_path, shape):
    h, w = shape
    img = np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 4)).astype(np.uint8)
    path = tmp_path / "rt.png"
    S.export_frame(path, w, h, img)
    back = np.array(Image.open(path))
    assert back.shape == (h, w, 4) and np.array_equal(back, img)


def test_c_abi_argument_checks(tmp_path):
    from soap_bubble_simulation_b200 import _lib as L

    lib = L.lib()
    buf = bytes(16)
    assert lib.sbx_export_frame_png(None, 2, 2, buf, 16) == L.ERR_INVALID
    assert lib.sbx_export_frame_png(os.fsencode(tmp_path / "a.png"), 2, 2, None, 16) == L.ERR_INVALID
    assert lib.sbx_export_frame_png(os.fsencode(tmp_path / "a.png"), 2, 2, buf, 15) == L.ERR_INVALID
    assert b"doesn't match" in lib.sbx_last_error()
    assert lib.sbx_export_frame_png(os.fsencode(tmp_path / "a.png"), 2, 2, buf, 16) == L.SBX_OK


def test_frame_params_default_matches_the_reference_defaults():
    """Camera::new (camera.rs:49-60) and BubbleUniform::default (pipeline.rs:69-112)."""
    import ctypes as C
    import math

    from soap_bubble_simulation_b200 import _lib as L

    p = L.FrameParamsC()
    L.check(L.lib().sbx_frame_params_default(C.byref(p)))
    assert C.sizeof(L.FrameParamsC) == 80 and L.FrameParamsC.background.offset == 64  # 20 four-byte fields, no padding
    assert (p.width, p.height) == (256, 256)
    assert p.cam_distance == pytest.approx(0.2) and p.cam_yaw == 0.0 and p.cam_pitch == pytest.approx(0.3)
    assert p.cam_fov_rad == pytest.approx(math.radians(45.0))
    assert p.base_thickness_nm == 500.0 and p.interference_intensity == 4.0
    assert p.base_alpha == pytest.approx(0.3) and p.edge_alpha == pytest.approx(0.6) and p.edge_smoothing_mode == 1
    assert list(p.background) == pytest.approx([0.1, 0.1, 0.15]) and p.lut_max_nm == 2000.0
    assert (p.swirl_intensity, p.drainage_speed, p.pattern_scale, p.film_time) == (1.0, 0.5, 1.0, 0.0)


def test_camera_mirror():
    """Camera::orbit / zoom clamps (camera.rs:71-84) and HeadlessRenderPipeline::set_camera_distance (headless.rs:540-542)."""
    import math

    c = S.Camera(1.0)
    c.orbit(100.0, 50.0)
    assert c.yaw == pytest.approx(1.0) and c.pitch == pytest.approx(0.8)
    c.orbit(0.0, 1000.0)
    assert c.pitch == pytest.approx(math.pi / 2 - 0.01, abs=1e-6)
    c.zoom(100.0)
    assert c.distance == 0.05
    x, y, z = S.Camera(1.0).position()
    assert (x, y, z) == pytest.approx((0.0, 0.2 * math.sin(0.3), 0.2 * math.cos(0.3)), abs=1e-7)
