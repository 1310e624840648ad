"""Rows f3/f4 on the GPU: the bubble fragment shader's shading path as a CUDA kernel (csrc/sbx_frame.cu) behind the
reference's HeadlessRenderPipeline surface, checked (a) against the reference's own golden screenshots with the
reference's comparator (tests/screenshot_tests.rs:120-145: channel tolerance 5/255, at most 5 % of pixels; we require
3 %), the scenes set up exactly as its tests do (:286-470), (b) pixel by pixel against the numpy restatement
(tests/golden_render.py), and (c) with the film thickness read from a drainage handle's device field."""
import os

import numpy as np
import pytest
from PIL import Image

import golden_render as G
from synth import synthetic_field

pytestmark = pytest.mark.gpu
PNG = os.path.join(os.path.dirname(__file__), "golden", "reference_png")
LIMIT = 0.03


def _ref(name):
    return np.array(Image.open(os.path.join(PNG, name + ".png")).convert("RGBA"))


def _pipeline(name):
    """The state each reference screenshot test puts the pipeline in (tests/screenshot_tests.rs:286-470)."""
    import soap_bubble_simulation_b200 as S

    p = S.HeadlessRenderPipeline.new(256, 256)
    p.set_time(0.0)
    p.set_thickness({"thickness_200nm": 200.0, "thickness_400nm": 400.0, "thickness_600nm": 600.0,
                     "thickness_800nm": 800.0}.get(name, 500.0))
    p.set_camera_distance({"zoomed_in": 0.1, "zoomed_out": 0.5}.get(name, 0.2))
    if name == "rotated_view":
        p.orbit_camera(100.0, 50.0)
    return p


def _off(a, b, lsb):
    """fraction of pixels with a colour channel off by more than `lsb`"""
    d = np.abs(a[..., :3].astype(int) - b[..., :3].astype(int))
    return float(np.count_nonzero((d > lsb).any(axis=-1))) / (a.shape[0] * a.shape[1])


def _changed(a, b, lsb=2):
    """fraction of the bubble's pixels (anything that is not background in either image) that differ by more than `lsb`"""
    bg = np.array([89, 89, 108])
    hit = (a[..., :3] != bg).any(axis=-1) | (b[..., :3] != bg).any(axis=-1)
    d = np.abs(a[..., :3].astype(int) - b[..., :3].astype(int))
    return float(np.count_nonzero((d > lsb).any(axis=-1) & hit)) / max(int(hit.sum()), 1)


@pytest.mark.parametrize("name", sorted(G.SCENES))
def test_gpu_frame_reproduces_reference_golden(name):
    img = _pipeline(name).render_to_buffer()
    assert img.shape == (256, 256, 4) and (img[..., 3] == 255).all()
    ratio = G.diff_ratio(_ref(name), img)
    assert ratio <= LIMIT, f"{name}: {100 * ratio:.2f} % of pixels differ by more than 5/255"
    assert tuple(img[2, 2, :3]) == (89, 89, 108)  # background = sRGB(0.1, 0.1, 0.15)


@pytest.mark.parametrize("name", sorted(G.SCENES))
def test_gpu_frame_matches_numpy_restatement(name):
    import soap_bubble_simulation_b200 as S

    lut = S.generate_interference_lut(1.33, 1.0).reshape(64, 256, 4)
    ref = G.render_bubble(lut, **G.SCENES[name])
    img = _pipeline(name).render_to_buffer()
    off2 = _off(ref, img, 2)
    assert off2 <= 0.005, f"{name}: {100 * off2:.3f} % of pixels off by more than 2 LSB"


def test_gpu_frame_time_and_size_parameters():
    """set_time drives film_time (headless.rs:545-548); non-square targets keep the vertical field of view."""
    import soap_bubble_simulation_b200 as S

    p = S.HeadlessRenderPipeline.new(320, 200)
    p.set_time(3.7)
    img = p.render_to_buffer()
    assert img.shape == (200, 320, 4)
    lut = S.generate_interference_lut(1.33, 1.0).reshape(64, 256, 4)
    # golden_render is square-only: the centre 200 x 200 crop has the same pixel rays as a 200 x 200 frame
    ref = G.render_bubble(lut, size=200, film_time=3.7)
    assert _off(ref, img[:, 60:260], 2) <= 0.005
    still = S.HeadlessRenderPipeline.new(320, 200).render_to_buffer()
    assert _changed(still, img) > 0.10  # the film pattern moved


def test_black_film_and_edge_modes():
    import soap_bubble_simulation_b200 as S

    p = S.HeadlessRenderPipeline.new(128, 128)
    p.set_thickness(20.0)  # everything below 30 nm: Newton's black film (bubble.wgsl:517-519)
    img = p.render_to_buffer()
    lut = S.generate_interference_lut(1.33, 1.0).reshape(64, 256, 4)
    ref = G.render_bubble(lut, base_thickness_nm=20.0, size=128)
    assert _off(ref, img, 1) <= 0.005
    assert (img[64, 64, :3].astype(int) < np.array([89, 89, 108])).all()  # darker than the background behind it
    base = S.HeadlessRenderPipeline.new(128, 128).render_to_buffer()
    for mode in (0, 2):
        p = S.HeadlessRenderPipeline.new(128, 128)
        p.bubble_uniform.edge_smoothing_mode = mode
        other = p.render_to_buffer()
        assert other.shape == (128, 128, 4) and _changed(base, other) > 0.05  # a different alpha ramp


def _bilinear_nm(field, n):
    """get_thickness (drainage.rs:516-552) at the normals' (theta, phi), vectorised, in nm."""
    nt, nphi = field.shape
    theta = np.arccos(np.clip(n[:, 1].astype(np.float64), -1, 1))
    phi = np.mod(np.arctan2(n[:, 2].astype(np.float64), n[:, 0].astype(np.float64)), 2 * np.pi)
    tc, pc = theta / (np.pi / (nt - 1)), phi / (2 * np.pi / nphi)
    ti = np.minimum(np.floor(tc).astype(int), nt - 2)
    pi0 = np.floor(pc).astype(int) % nphi
    pi1 = (pi0 + 1) % nphi
    ft, fp = tc - ti, pc - np.floor(pc)
    v0 = field[ti, pi0] * (1 - fp) + field[ti, pi1] * fp
    v1 = field[ti + 1, pi0] * (1 - fp) + field[ti + 1, pi1] * fp
    return (v0 * (1 - ft) + v1 * ft) * 1e9


@pytest.mark.parametrize("prec", ["f32", "f32_offset", "f64"])
def test_drainage_field_frame(prec):
    """sbx_drainage_render_frame: thickness from the handle's device field (after a few steps) instead of the noise."""
    import soap_bubble_simulation_b200 as S

    nt, nphi = 128, 256
    f = synthetic_field(nt, nphi)
    kw = {"f32": dict(precision=S.PREC_F32, h_ref=0.0), "f32_offset": dict(precision=S.PREC_F32, h_ref=500e-9),
          "f64": dict(precision=S.PREC_F64)}[prec]
    sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=nt, num_phi=nphi, **kw)
    sim.upload(f)
    sim.step(1e-4, 3)
    field = sim.download()[0].astype(np.float64)
    img = S.render_drainage_frame(sim, 256, 256)
    lut = S.generate_interference_lut(1.33, 1.0).reshape(64, 256, 4)
    ref = G.render_bubble(lut, thickness_fn=lambda n: _bilinear_nm(field, n))
    off2 = _off(ref, img, 2)
    assert off2 <= 0.005, f"{100 * off2:.3f} % of pixels off by more than 2 LSB"
    sim.reset(300e-9)  # a different field gives a different picture
    assert _changed(S.render_drainage_frame(sim, 256, 256), img) > 0.2
    sim.close()


def test_drainage_frame_argument_checks():
    import soap_bubble_simulation_b200 as S

    sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=64, num_phi=128, row_begin=0, row_end=32)
    with pytest.raises(S.SbxError):
        S.render_drainage_frame(sim, 64, 64)  # a latitude band, not a whole sphere
    sim.close()
    sim = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=64, num_phi=128)
    with pytest.raises(S.SbxError):
        S.render_drainage_frame(sim, 64, 64, bubble=1)
    cam = S.Camera(1.0)
    cam.distance = 0.01  # inside the bubble
    with pytest.raises(S.SbxError):
        S.render_drainage_frame(sim, 64, 64, camera=cam)
    sim.close()


def test_frame_to_png_round_trip(tmp_path):
    """render -> export_frame -> independent decoder: the f3 wire format end to end."""
    import soap_bubble_simulation_b200 as S

    img = _pipeline("default_bubble").render_to_buffer()
    path = tmp_path / "default_bubble.png"
    S.export_frame(path, 256, 256, img)
    back = np.array(Image.open(path).convert("RGBA"))
    assert np.array_equal(back, img)
    assert G.diff_ratio(_ref("default_bubble"), back) <= LIMIT
