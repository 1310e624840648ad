"""The reference's host-side unit tests around the path, replayed on the Python mirrors (no GPU): camera.rs:121-330
(orbit camera the headless frame is rendered with), headless.rs:566-571 (BubbleUniform defaults), interference.rs:618-646
(RgbColor clamping, default calculator).  The two matrix-determinant tests (camera.rs:254-270) have no counterpart: the
mirror hands the camera's orbit parameters to libsbx, which builds the look-at basis itself (checked on the GPU against
the reference's golden screenshots)."""
import math

import numpy as np
import pytest

import soap_bubble_simulation_b200 as S


def _len(v):
    return float(np.linalg.norm(np.array(v)))


def _dist(a, b):
    return float(np.linalg.norm(np.array(a) - np.array(b)))


def test_camera_creation():  # camera.rs:122-127
    c = S.Camera(16.0 / 9.0)
    assert c.distance > 0.0 and c.fov > 0.0


def test_camera_position():  # camera.rs:129-135
    c = S.Camera(1.0)
    assert abs(_len(c.position()) - c.distance) < 0.001


def test_orbit_increases_and_decreases_yaw():  # camera.rs:145-165
    c = S.Camera(1.0)
    c.orbit(100.0, 0.0)
    assert c.yaw > 0.0
    c = S.Camera(1.0)
    c.orbit(-100.0, 0.0)
    assert c.yaw < 0.0


def test_orbit_changes_position():  # camera.rs:167-177
    c = S.Camera(1.0)
    p0 = c.position()
    c.orbit(50.0, 50.0)
    assert _dist(p0, c.position()) > 0.001


def test_pitch_clamping():  # camera.rs:179-201
    c = S.Camera(1.0)
    c.orbit(0.0, 10000.0)
    assert c.pitch <= math.pi / 2.0 - 0.01 + 1e-6
    c = S.Camera(1.0)
    c.orbit(0.0, -10000.0)
    assert c.pitch >= -(math.pi / 2.0 - 0.01) - 1e-6


def test_zoom():  # camera.rs:203-245
    c = S.Camera(1.0)
    d0 = c.distance
    c.zoom(2.0)
    assert c.distance < d0
    c = S.Camera(1.0)
    c.zoom(-2.0)
    assert c.distance > d0
    c = S.Camera(1.0)
    c.zoom(1000.0)
    assert c.distance >= 0.05
    c = S.Camera(1.0)
    c.zoom(-1000.0)
    assert c.distance <= 10.0


def test_aspect_ratio_update():  # camera.rs:247-252
    c = S.Camera(16.0 / 9.0)
    assert abs(c.aspect - 16.0 / 9.0) < 0.001
    c.set_aspect(4.0 / 3.0)
    assert abs(c.aspect - 4.0 / 3.0) < 0.001


def test_input_sequence_simulation():  # camera.rs:272-298
    c = S.Camera(1.0)
    p0 = c.position()
    for _ in range(10):
        c.orbit(10.0, 5.0)
    c.zoom(2.0)
    p1 = c.position()
    assert _dist(p0, p1) > 0.01
    assert abs(_len(p1) - c.distance) < 0.001


def test_orbit_360_degrees():  # camera.rs:300-321
    c = S.Camera(1.0)
    p0 = c.position()
    for _ in range(628):
        c.orbit(1.0, 0.0)
    assert _dist(p0, c.position()) < 0.1


def test_bubble_uniform_defaults():  # headless.rs:566-571, pipeline.rs:69-112
    u = S.BubbleUniform()
    assert abs(u.refractive_index - 1.33) < 0.01 and abs(u.base_thickness_nm - 500.0) < 1.0
    assert (u.interference_intensity, u.base_alpha, u.edge_alpha, u.edge_smoothing_mode) == (4.0, 0.3, 0.6, 1)
    assert (u.background_r, u.background_g, u.background_b) == (0.1, 0.1, 0.15) and u.branched_flow_enabled == 0


def test_headless_setters_follow_the_reference():  # headless.rs:515-548, without building the LUT (needs a GPU)
    p = S.HeadlessRenderPipeline.__new__(S.HeadlessRenderPipeline)
    p.width, p.height, p.camera, p.bubble_uniform = 256, 128, S.Camera(2.0), S.BubbleUniform()
    assert p.size() == (256, 128)
    p.set_thickness(800.0)
    p.set_refractive_index(1.4)
    p.set_time(2.5)
    p.set_camera_distance(100.0)
    assert p.bubble_uniform.base_thickness_nm == 800.0 and p.bubble_uniform.refractive_index == 1.4
    assert p.bubble_uniform.time == 2.5 and p.bubble_uniform.film_time == 2.5
    assert p.camera.distance == 10.0  # clamp(0.05, 10.0)
    p.set_camera_distance(0.0)
    assert p.camera.distance == 0.05
    p.zoom_camera(-2.0)
    p.orbit_camera(100.0, 50.0)
    assert p.camera.distance == pytest.approx(0.06) and p.camera.yaw == pytest.approx(1.0)


def test_rgb_color_clamping():  # interference.rs:618-624
    c = S.RgbColor.new(1.5, -0.5, 0.5)
    assert abs(c.red - 1.0) < 1e-10 and abs(c.green) < 1e-10 and abs(c.blue - 0.5) < 1e-10


def test_default_calculator():  # interference.rs:642-646 (REFRACTIVE_INDEX_SOAP_FILM = 1.33, :52-61)
    assert abs(S.InterferenceCalculator.default_soap_film().refractive_index_film() - 1.33) < 1e-10
    assert S.InterferenceCalculator.default_soap_film().refractive_index_medium() == 1.0
