"""Test-side software restatement of the reference's headless bubble renderer, used only to compare this
repository's interference LUT with the reference's golden screenshots (tests/golden/reference_png).

Follows src/render/headless.rs (256x256 Rgba8UnormSrgb target cleared to (0.1,0.1,0.15), alpha blending, depth
test Less so only the front surface survives for this camera, sphere radius 0.025 at the origin),
src/render/camera.rs:49-105 (orbit camera, look_at_rh, perspective_rh fov 45 deg) and
src/render/shaders/bubble.wgsl: simplex/FBM film thickness (:104-229), LUT sampling x interference_intensity
(:425-439), black film below 30 nm (:517-519), edge alpha (:501-513), `base*0.1 + interference` (:528-531).
The triangle mesh (128x64 segments) is replaced by an analytic ray-sphere hit; only silhouette pixels can differ.
Everything is float32 like the shader.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np

f32 = np.float32


def _mod(x, y):  # WGSL float %: truncated remainder
    return (x - y * np.trunc(x / y)).astype(f32)


def _permute(x):
    return _mod((x * f32(34.0) + f32(1.0)) * x, f32(289.0))


def _step(edge, x):
    return (x >= edge).astype(f32)


def simplex_noise_3d(v):
    """bubble.wgsl:113-176 (Gustavson's simplex noise); v: (n,3) float32 -> (n,)"""
    v = v.astype(f32)
    Cx, Cy = f32(1.0 / 6.0), f32(1.0 / 3.0)
    i = np.floor(v + (v @ np.array([Cy, Cy, Cy], f32))[:, None]).astype(f32)
    x0 = (v - i + (i @ np.array([Cx, Cx, Cx], f32))[:, None]).astype(f32)
    g = _step(x0[:, [1, 2, 0]], x0)            # step(x0.yzx, x0.xyz)
    l = f32(1.0) - g
    i1 = np.minimum(g, l[:, [2, 0, 1]])         # min(g.xyz, l.zxy)
    i2 = np.maximum(g, l[:, [2, 0, 1]])
    x1 = x0 - i1 + Cx
    x2 = x0 - i2 + Cy
    x3 = x0 - f32(0.5)
    i = _mod(i, f32(289.0))
    one = np.ones(len(v), f32)
    zero = np.zeros(len(v), f32)

    def col(a, b, c, d):
        return np.stack([a, b, c, d], axis=1)

    p = _permute(_permute(_permute(i[:, 2:3] + col(zero, i1[:, 2], i2[:, 2], one)) + i[:, 1:2]
                          + col(zero, i1[:, 1], i2[:, 1], one)) + i[:, 0:1] + col(zero, i1[:, 0], i2[:, 0], one))
    n_ = f32(0.142857142857)
    # ns = n_ * D.wyz - D.xzx with D = (0, 0.5, 1, 2)
    ns = np.array([n_ * f32(2.0) - f32(0.0), n_ * f32(0.5) - f32(1.0), n_ * f32(1.0) - f32(0.0)], f32)
    j = p - f32(49.0) * np.floor(p * ns[2] * ns[2])
    x_ = np.floor(j * ns[2])
    y_ = np.floor(j - f32(7.0) * x_)
    x = x_ * ns[0] + ns[1]
    y = y_ * ns[0] + ns[1]
    h = f32(1.0) - np.abs(x) - np.abs(y)
    b0 = np.stack([x[:, 0], x[:, 1], y[:, 0], y[:, 1]], axis=1)
    b1 = np.stack([x[:, 2], x[:, 3], y[:, 2], y[:, 3]], axis=1)
    s0 = np.floor(b0) * f32(2.0) + f32(1.0)
    s1 = np.floor(b1) * f32(2.0) + f32(1.0)
    sh = -_step(h, np.zeros_like(h))            # -step(h, 0): -(0 >= h)
    a0 = b0[:, [0, 2, 1, 3]] + s0[:, [0, 2, 1, 3]] * sh[:, [0, 0, 1, 1]]
    a1 = b1[:, [0, 2, 1, 3]] + s1[:, [0, 2, 1, 3]] * sh[:, [2, 2, 3, 3]]
    p0 = np.stack([a0[:, 0], a0[:, 1], h[:, 0]], axis=1)
    p1 = np.stack([a0[:, 2], a0[:, 3], h[:, 1]], axis=1)
    p2 = np.stack([a1[:, 0], a1[:, 1], h[:, 2]], axis=1)
    p3 = np.stack([a1[:, 2], a1[:, 3], h[:, 3]], axis=1)

    def dot(a, b):
        return np.sum(a * b, axis=1, dtype=f32)

    norm = f32(1.79284291400159) - f32(0.85373472095314) * np.stack([dot(p0, p0), dot(p1, p1), dot(p2, p2), dot(p3, p3)], 1)
    p0, p1, p2, p3 = p0 * norm[:, 0:1], p1 * norm[:, 1:2], p2 * norm[:, 2:3], p3 * norm[:, 3:4]
    m = np.maximum(f32(0.6) - np.stack([dot(x0, x0), dot(x1, x1), dot(x2, x2), dot(x3, x3)], 1), f32(0.0))
    m = m * m
    return (f32(42.0) * np.sum(m * m * np.stack([dot(p0, x0), dot(p1, x1), dot(p2, x2), dot(p3, x3)], 1), axis=1)).astype(f32)


def fbm_noise(p, octaves):  # bubble.wgsl:179-191
    value = np.zeros(len(p), f32)
    amplitude, frequency = f32(0.5), f32(1.0)
    for _ in range(octaves):
        value = value + amplitude * simplex_noise_3d(p * frequency)
        amplitude *= f32(0.5)
        frequency *= f32(2.0)
    return value


def film_thickness(normal, base_nm, film_time=0.0, scale=1.0, swirl=1.0, drainage_speed=0.5):
    """get_film_thickness, bubble.wgsl:200-229"""
    n = normal.astype(f32)
    t = f32(film_time)
    drain_progress = min(t * f32(drainage_speed) * f32(0.1), f32(1.0))
    drainage = f32(0.4) * (f32(1.0) - n[:, 1]) * f32(0.5) * (f32(1.0) + drain_progress)
    noise_time = t * f32(0.08)
    noise_coord = n * f32(scale) * f32(3.0) + np.array([noise_time, noise_time * f32(0.7), noise_time * f32(0.5)], f32)
    organic = fbm_noise(noise_coord, 4) * f32(swirl) * f32(0.12)
    ang = t * f32(0.5)
    sx = n[:, 0] * np.cos(ang) - n[:, 2] * np.sin(ang)
    sz = n[:, 0] * np.sin(ang) + n[:, 2] * np.cos(ang)
    swirl_noise = simplex_noise_3d(np.stack([sx, n[:, 1], sz], 1).astype(f32) * f32(5.0) + t * f32(0.2)) * f32(swirl) * f32(0.05)
    wave = f32(0.02) * f32(swirl) * np.sin(n[:, 1] * f32(scale) * f32(8.0) - t * f32(2.0)) * (f32(1.0) - np.abs(n[:, 1]))
    return (f32(base_nm) * (f32(1.0) - drainage + organic + swirl_noise - wave)).astype(f32)


def lut_sample(lut, u, v):
    """textureSample with a linear, clamp-to-edge sampler on an Rgba8Unorm texture (pipeline.rs:316-359)."""
    A, T = lut.shape[0], lut.shape[1]
    x = u.astype(f32) * f32(T) - f32(0.5)
    y = v.astype(f32) * f32(A) - f32(0.5)
    x0f, y0f = np.floor(x), np.floor(y)
    wx, wy = (x - x0f)[:, None], (y - y0f)[:, None]
    x0 = np.clip(x0f.astype(int), 0, T - 1)
    x1 = np.clip(x0f.astype(int) + 1, 0, T - 1)
    y0 = np.clip(y0f.astype(int), 0, A - 1)
    y1 = np.clip(y0f.astype(int) + 1, 0, A - 1)
    t = lut[..., :3].astype(f32) / f32(255.0)
    top = t[y0, x0] * (1 - wx) + t[y0, x1] * wx
    bot = t[y1, x0] * (1 - wx) + t[y1, x1] * wx
    return (top * (1 - wy) + bot * wy).astype(f32)


def srgb_encode(c):
    c = np.clip(c, 0.0, 1.0)
    return np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1.0 / 2.4) - 0.055)


def render_bubble(lut, base_thickness_nm=500.0, distance=0.2, yaw=0.0, pitch=0.3, size=256, intensity=4.0,
                  base_alpha=0.3, edge_alpha=0.6, background=(0.1, 0.1, 0.15), radius=0.025, shade=None,
                  thickness_fn=None, film_time=0.0):
    # shade(lut, thickness_nm, cos_theta, intensity) -> (n,3): the product's LUT shading; default: numpy bilinear
    # thickness_fn(normals (n,3) f32) -> thickness in nm: replaces the shader's procedural film thickness
    """(size,size,4) uint8 image of the default headless scene (BubbleUniform::default, pipeline.rs:69-112)."""
    d, yw, pt = f32(distance), f32(yaw), f32(pitch)
    cam = np.array([d * np.cos(pt) * np.sin(yw), d * np.sin(pt), d * np.cos(pt) * np.cos(yw)], f32)  # camera.rs:63-68
    fwd = -cam / np.linalg.norm(cam)                                   # look_at_rh(position, origin, Y)
    right = np.cross(fwd, np.array([0, 1, 0], f32))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    tan_half = np.tan(np.radians(45.0) / 2.0)
    px = (np.arange(size) + 0.5) / size * 2.0 - 1.0                    # NDC x at pixel centres
    py = 1.0 - (np.arange(size) + 0.5) / size * 2.0                    # NDC y (row 0 is the top)
    X, Y = np.meshgrid(px, py)
    dirs = fwd[None, None, :] + (X * tan_half)[..., None] * right + (Y * tan_half)[..., None] * up
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    b = dirs @ cam
    disc = b * b - (cam @ cam - radius * radius)
    hit = disc > 0
    tnear = -b - np.sqrt(np.where(hit, disc, 0.0))
    pos = cam + tnear[..., None] * dirs
    img = np.empty((size, size, 3), np.float64)
    img[:] = background
    idx = np.nonzero(hit)
    p = pos[idx].astype(f32)
    n = (p / np.linalg.norm(p, axis=1, keepdims=True)).astype(f32)
    view_dir = cam - p
    view_dir /= np.linalg.norm(view_dir, axis=1, keepdims=True)
    cos_theta = np.abs(np.sum(view_dir * n, axis=1)).astype(f32)        # bubble.wgsl:490-494
    thick = film_thickness(n, base_thickness_nm, film_time) if thickness_fn is None else thickness_fn(n).astype(f32)
    edge = f32(1.0) - cos_theta
    smooth = edge * edge * (f32(3.0) - f32(2.0) * edge)                 # edge_smoothing_mode 1
    alpha = (f32(base_alpha) + f32(edge_alpha) * smooth).astype(np.float64)
    u = np.clip(thick / f32(2000.0), 0, 1)
    v = np.clip(cos_theta, 0, 1)
    if shade is None:
        interf = np.clip(lut_sample(lut, u, v) * f32(intensity), 0, 1)  # bubble.wgsl:425-439
    else:
        interf = shade(lut, thick, cos_theta, intensity)
    color = (np.array([0.95, 0.97, 1.0], f32) * f32(0.1) + interf).astype(np.float64)
    black = thick < 30.0                                                # bubble.wgsl:517-519
    color[black] = 0.02
    alpha = np.where(black, alpha * 0.5, alpha)
    a = np.clip(alpha, 0, 1)[:, None]
    # BlendState::ALPHA_BLENDING on a unorm target: the fragment colour is clamped to [0,1] before the blend (this
    # also matches the golden screenshots better than an unclamped source: zoomed_in 2.1 % -> 0.9 % of pixels off)
    img[idx] = np.clip(color, 0, 1) * a + np.array(background) * (1 - a)
    out = np.empty((size, size, 4), np.uint8)
    out[..., :3] = np.floor(srgb_encode(img) * 255.0 + 0.5).astype(np.uint8)  # Rgba8UnormSrgb store
    out[..., 3] = 255
    return out


def diff_ratio(expected, actual, color_tolerance=5):
    """ScreenshotComparator::compute_diff (tests/screenshot_tests.rs:120-145): fraction of pixels with any of
    r,g,b off by more than `color_tolerance`."""
    d = np.abs(expected[..., :3].astype(int) - actual[..., :3].astype(int))
    return float(np.count_nonzero((d > color_tolerance).any(axis=-1))) / (expected.shape[0] * expected.shape[1])


SCENES = {  # tests/screenshot_tests.rs:286-470
    "default_bubble": dict(base_thickness_nm=500.0),
    "thickness_200nm": dict(base_thickness_nm=200.0),
    "thickness_400nm": dict(base_thickness_nm=400.0),
    "thickness_600nm": dict(base_thickness_nm=600.0),
    "thickness_800nm": dict(base_thickness_nm=800.0),
    "zoomed_in": dict(base_thickness_nm=500.0, distance=0.1),
    "zoomed_out": dict(base_thickness_nm=500.0, distance=0.5),
    "rotated_view": dict(base_thickness_nm=500.0, yaw=100.0 * 0.01, pitch=0.3 + 50.0 * 0.01),  # orbit_camera(100, 50)
}
