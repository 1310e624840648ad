"""The reference's own golden screenshots (tests/golden/*.png, checked by tests/screenshot_tests.rs with a 5/255
per-channel tolerance and at most 5 % differing pixels) re-rendered with this repository's interference LUT.

These eight images are the only reference *outputs* that exist for the interference path.  The scene around the
LUT (camera, sphere, simplex-noise film thickness, alpha blending, sRGB target) is restated in tests/golden_render.py;
the LUT and its sampling come (a) from the CPU oracle -- which pins the oracle to the reference -- and (b), on a GPU
box, from the product (sbx_lut_build + the texture-unit LUT fetch).  Wrong refractive index, channel order,
intensity, wavelength set or angle axis all push the difference above the limit (see the negative test)."""
import os

import numpy as np
import pytest
from PIL import Image

import golden_render as G
import oracle_lib as O

PNG = os.path.join(os.path.dirname(__file__), "golden", "reference_png")
LIMIT = 0.03  # stricter than the reference's own 0.05


def _ref(name):
    return np.array(Image.open(os.path.join(PNG, name + ".png")).convert("RGBA"))


@pytest.mark.parametrize("name", sorted(G.SCENES))
def test_oracle_lut_reproduces_reference_golden(name):
    lut = O.generate_interference_lut(1.33, 1.0)
    img = G.render_bubble(lut, **G.SCENES[name])
    ratio = G.diff_ratio(_ref(name), img)
    assert ratio <= LIMIT, f"{name}: {100 * ratio:.2f} % of pixels differ by more than 5/255"
    assert tuple(img[2, 2, :3]) == (89, 89, 108)  # background = sRGB(0.1, 0.1, 0.15), SURVEY section 4


def test_golden_comparison_rejects_wrong_luts():
    """The comparison has teeth: each of these plausible mistakes fails on the close-up scene."""
    good = O.generate_interference_lut(1.33, 1.0)
    ref = _ref("zoomed_in")
    kw = G.SCENES["zoomed_in"]
    assert G.diff_ratio(ref, G.render_bubble(good, **kw)) <= LIMIT
    wrong = {
        "n=1.40": O.generate_interference_lut(1.40, 1.0),
        "swapped R/B": np.ascontiguousarray(good[..., [2, 1, 0, 3]]),
        "3 wavelengths instead of 7": O.generate_lut_spectral(1.33, 1.0, [650.0, 532.0, 450.0], 256, 64, 2000.0),
        "angle axis flipped": np.ascontiguousarray(good[::-1]),
    }
    for what, lut in wrong.items():
        assert G.diff_ratio(ref, G.render_bubble(lut, **kw)) > 0.2, what
    assert G.diff_ratio(ref, G.render_bubble(good, intensity=2.0, **kw)) > 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.SCENES))
def test_product_lut_and_texture_fetch_reproduce_reference_golden(name):
    import soap_bubble_simulation_b200 as S

    lut = S.generate_interference_lut(1.33, 1.0).reshape(64, 256, 4)
    img = G.render_bubble(lut, shade=lambda l, d, c, i: S.thin_film_interference(l, d, c, i), **G.SCENES[name])
    ratio = G.diff_ratio(_ref(name), img)
    assert ratio <= LIMIT, f"{name}: {100 * ratio:.2f} % of pixels differ by more than 5/255"
