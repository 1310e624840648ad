"""The reference's integration tests (tests/integration_harness.rs:195-396: TestHarness over HeadlessRenderPipeline,
scripted TestSteps, frame utilities) replayed twice: on the numpy restatement of the headless renderer (CPU, always
runs -- it also proves the test logic) and on the product's HeadlessRenderPipeline mirror, whose frames come from
k_bubble_frame on the GPU."""
import numpy as np
import pytest

import golden_render as G
import oracle_lib as O
import soap_bubble_simulation_b200 as S


class _RestatementPipeline:
    """Same surface as S.HeadlessRenderPipeline, frames from tests/golden_render.py with the oracle's LUT."""

    def __init__(self, width, height):
        assert width == height, "the restatement renders square frames"
        self.width, self.height = width, height
        self.camera, self.bubble_uniform = S.Camera(1.0), S.BubbleUniform()
        self._lut = O.generate_interference_lut(1.33, 1.0)

    size = S.HeadlessRenderPipeline.size
    set_thickness = S.HeadlessRenderPipeline.set_thickness
    set_time = S.HeadlessRenderPipeline.set_time
    set_camera_distance = S.HeadlessRenderPipeline.set_camera_distance
    orbit_camera = S.HeadlessRenderPipeline.orbit_camera
    zoom_camera = S.HeadlessRenderPipeline.zoom_camera

    def render_to_buffer(self):
        return G.render_bubble(self._lut, base_thickness_nm=self.bubble_uniform.base_thickness_nm,
                               distance=self.camera.distance, yaw=self.camera.yaw, pitch=self.camera.pitch,
                               size=self.width, film_time=self.bubble_uniform.film_time)


class TestHarness:  # integration_harness.rs:31-134
    __test__ = False

    def __init__(self, pipeline):
        self.pipeline, self.frames_rendered = pipeline, 0

    def size(self):
        return self.pipeline.size()

    def render_frame(self):
        self.frames_rendered += 1
        return self.pipeline.render_to_buffer().reshape(-1)

    def run_scenario(self, steps):  # :94-123
        frames = []
        for name, *args in steps:
            if name == "RenderFrame":
                frames.append(self.render_frame().copy())
            else:
                {"OrbitCamera": self.pipeline.orbit_camera, "ZoomCamera": self.pipeline.zoom_camera,
                 "SetThickness": self.pipeline.set_thickness, "SetTime": self.pipeline.set_time,
                 "SetCameraDistance": self.pipeline.set_camera_distance}[name](*args)
        return frames


def average_color(pixels):  # :137-163
    p = np.asarray(pixels, np.uint8).reshape(-1)
    if p.size == 0 or p.size % 4 != 0:
        return (0.0, 0.0, 0.0, 0.0)
    return tuple(float(v) for v in p.reshape(-1, 4).astype(np.float64).mean(axis=0))


def frame_diff_ratio(a, b, tolerance):  # :171-188
    a, b = np.asarray(a, np.uint8).reshape(-1), np.asarray(b, np.uint8).reshape(-1)
    if a.size != b.size or a.size == 0:
        return 1.0
    d = np.abs(a.reshape(-1, 4)[:, :3].astype(int) - b.reshape(-1, 4)[:, :3].astype(int))
    return float(np.count_nonzero((d > tolerance).any(axis=1))) / (a.size // 4)


@pytest.fixture(params=["restatement", pytest.param("product", marks=pytest.mark.gpu)])
def make_harness(request):
    if request.param == "restatement":
        return lambda w, h: TestHarness(_RestatementPipeline(w, h))
    return lambda w, h: TestHarness(S.HeadlessRenderPipeline.new(w, h))


def test_harness_creation(make_harness):  # :194-201
    assert make_harness(256, 256).size() == (256, 256)


def test_full_render_cycle_and_content(make_harness):  # :203-236
    frame = make_harness(256, 256).render_frame()
    assert frame.size == 256 * 256 * 4
    assert frame.any(), "Frame should have some content (not all black)"


def test_camera_orbit_changes_view(make_harness):  # :238-259
    h = make_harness(128, 128)
    f1 = h.render_frame().copy()
    h.pipeline.orbit_camera(100.0, 0.0)
    assert not np.array_equal(f1, h.render_frame()), "Camera orbit should change the rendered view"


def test_zoom_changes_view(make_harness):  # :261-282
    h = make_harness(128, 128)
    f1 = h.render_frame().copy()
    h.pipeline.zoom_camera(3.0)
    assert not np.array_equal(f1, h.render_frame()), "Zoom should change the rendered view"


def test_thickness_changes_appearance(make_harness):  # :284-307
    h = make_harness(128, 128)
    h.pipeline.set_thickness(500.0)
    f1 = h.render_frame().copy()
    h.pipeline.set_thickness(200.0)
    assert not np.array_equal(f1, h.render_frame()), "Different film thickness should produce different colors"


def test_scenario_execution(make_harness):  # :309-349
    h = make_harness(128, 128)
    frames = h.run_scenario([("SetThickness", 500.0), ("RenderFrame",), ("OrbitCamera", 50.0, 0.0), ("RenderFrame",),
                             ("ZoomCamera", 2.0), ("RenderFrame",), ("SetThickness", 300.0), ("RenderFrame",)])
    assert len(frames) == 4 and h.frames_rendered == 4
    for a, b in zip(frames, frames[1:]):
        assert frame_diff_ratio(a, b, 0) > 0.0
    r, g, b, a = average_color(frames[0])
    assert a == 255.0 and 60.0 < r < 140.0 and 60.0 < g < 140.0 and 80.0 < b < 160.0  # mostly the (89, 89, 108) background


def test_multiple_renders_deterministic(make_harness):  # :351-375
    h = make_harness(128, 128)
    h.pipeline.set_thickness(500.0)
    h.pipeline.set_time(0.0)
    h.pipeline.set_camera_distance(0.2)
    assert np.array_equal(h.render_frame(), h.render_frame()), "Rendering same state twice should be deterministic"


def test_frame_utilities():  # :377-396
    a = [255, 0, 0, 255, 0, 255, 0, 255]
    c = [0, 0, 255, 255, 255, 255, 0, 255]
    assert frame_diff_ratio(a, list(a), 0) == 0.0 and frame_diff_ratio(a, c, 0) == 1.0
    r, g, b, al = average_color([255, 0, 0, 255, 255, 0, 0, 255])
    assert abs(r - 255.0) < 0.001 and abs(g) < 0.001 and abs(b) < 0.001 and abs(al - 255.0) < 0.001
