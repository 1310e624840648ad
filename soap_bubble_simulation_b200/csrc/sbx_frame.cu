// sbx_frame.cu -- the bubble fragment shader's shading path as a compute kernel.
//
// One thread per pixel of a headless frame (src/render/headless.rs: Rgba8UnormSrgb target cleared to the
// background colour, BlendState::ALPHA_BLENDING, depth test Less so only the front surface of the sphere survives
// for a camera outside the bubble).  The rasterised 128x64-segment sphere is replaced by an analytic ray-sphere
// hit; everything after the hit follows bubble.wgsl::fs_main (:487-541) in f32:
//   cos_theta = |view . n| (:490-494), film thickness (procedural simplex/FBM get_film_thickness :200-229, or the
//   drainage handle's thickness field blended like DrainageSimulator::get_thickness, drainage.rs:516-552),
//   edge alpha (:501-513), Newton's black film below 30 nm (:517-519), LUT fetch x interference_intensity with a
//   linear clamp-to-edge sampler (:425-439, pipeline.rs:350-359), base tint * 0.1 + interference (:528-531).
// Fixed-point targets clamp the fragment colour to [0,1] before blending; the store applies the sRGB transfer curve.
// Branched-flow filaments (:534-540) are off in BubbleUniform::default (pipeline.rs:91) and are not produced here.
#include "sbx_launch.h"

namespace sbx {

namespace {

struct F3 { float x, y, z; };

__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float stepf(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
__device__ __forceinline__ float mod289(float x) { return fmodf(x, 289.0f); }                   // WGSL float %: truncated
__device__ __forceinline__ float permute1(float x) { return mod289((x * 34.0f + 1.0f) * x); }  // bubble.wgsl:104-106

// simplex_noise_3d, bubble.wgsl:113-176 (Gustavson's 3-D simplex noise)
__device__ float simplex3(F3 v)
{
    const float Cx = 1.0f / 6.0f, Cy = 1.0f / 3.0f;
    const float s = (v.x + v.y + v.z) * Cy;
    F3 i = {floorf(v.x + s), floorf(v.y + s), floorf(v.z + s)};
    const float t = (i.x + i.y + i.z) * Cx;
    const F3 x0 = {v.x - i.x + t, v.y - i.y + t, v.z - i.z + t};
    // g = step(x0.yzx, x0.xyz), l = 1 - g, i1 = min(g, l.zxy), i2 = max(g, l.zxy)
    const F3 g = {stepf(x0.y, x0.x), stepf(x0.z, x0.y), stepf(x0.x, x0.z)};
    const F3 l = {1.0f - g.x, 1.0f - g.y, 1.0f - g.z};
    const F3 i1 = {fminf(g.x, l.z), fminf(g.y, l.x), fminf(g.z, l.y)};
    const F3 i2 = {fmaxf(g.x, l.z), fmaxf(g.y, l.x), fmaxf(g.z, l.y)};
    const F3 x1 = {x0.x - i1.x + Cx, x0.y - i1.y + Cx, x0.z - i1.z + Cx};
    const F3 x2 = {x0.x - i2.x + Cy, x0.y - i2.y + Cy, x0.z - i2.z + Cy};
    const F3 x3 = {x0.x - 0.5f, x0.y - 0.5f, x0.z - 0.5f};
    i.x = mod289(i.x); i.y = mod289(i.y); i.z = mod289(i.z);
    const float oz[4] = {0.0f, i1.z, i2.z, 1.0f}, oy[4] = {0.0f, i1.y, i2.y, 1.0f}, ox[4] = {0.0f, i1.x, i2.x, 1.0f};
    float p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = permute1(permute1(permute1(i.z + oz[k]) + i.y + oy[k]) + i.x + ox[k]);
    // ns = n_ * D.wyz - D.xzx with D = (0, 0.5, 1, 2)
    const float n_ = 0.142857142857f;
    const float nsx = n_ * 2.0f - 0.0f, nsy = n_ * 0.5f - 1.0f, nsz = n_ * 1.0f - 0.0f;
    float gx[4], gy[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float j = p[k] - 49.0f * floorf(p[k] * nsz * nsz);
        const float x_ = floorf(j * nsz);
        const float y_ = floorf(j - 7.0f * x_);
        gx[k] = x_ * nsx + nsy;
        gy[k] = y_ * nsx + nsy;
        h[k] = 1.0f - fabsf(gx[k]) - fabsf(gy[k]);
    }
    // b0 = (x.xy, y.xy), b1 = (x.zw, y.zw); s = floor(b)*2+1; sh = -step(h, 0); a0 = b0.xzyw + s0.xzyw*sh.xxyy, ...
    F3 pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float sh = -stepf(h[k], 0.0f);
        const float sx = floorf(gx[k]) * 2.0f + 1.0f, sy = floorf(gy[k]) * 2.0f + 1.0f;
        pk[k].x = gx[k] + sx * sh;
        pk[k].y = gy[k] + sy * sh;
        pk[k].z = h[k];
    }
    const F3 xs[4] = {x0, x1, x2, x3};
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float norm = 1.79284291400159f - 0.85373472095314f * dot3(pk[k], pk[k]);  // taylor_inv_sqrt
        const F3 q = {pk[k].x * norm, pk[k].y * norm, pk[k].z * norm};
        float m = fmaxf(0.6f - dot3(xs[k], xs[k]), 0.0f);
        m = m * m;
        acc += (m * m) * dot3(q, xs[k]);
    }
    return 42.0f * acc;
}

// fbm_noise, bubble.wgsl:179-191
__device__ float fbm4(F3 p)
{
    float value = 0.0f, amplitude = 0.5f, frequency = 1.0f;
#pragma unroll 1
    for (int o = 0; o < 4; ++o) {
        value += amplitude * simplex3({p.x * frequency, p.y * frequency, p.z * frequency});
        amplitude *= 0.5f;
        frequency *= 2.0f;
    }
    return value;
}

// get_film_thickness, bubble.wgsl:200-229
__device__ float film_thickness_procedural(const FrameArgs &a, F3 n)
{
    const float t = a.film_time, scale = a.pattern_scale, swirl = a.swirl;
    const float drain_progress = fminf(t * a.drainage_speed * 0.1f, 1.0f);
    const float drainage = 0.4f * (1.0f - n.y) * 0.5f * (1.0f + drain_progress);
    const float noise_time = t * 0.08f;
    const F3 nc = {n.x * scale * 3.0f + noise_time, n.y * scale * 3.0f + noise_time * 0.7f, n.z * scale * 3.0f + noise_time * 0.5f};
    const float organic = fbm4(nc) * swirl * 0.12f;
    const float ang = t * 0.5f;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    const float sx = n.x * cs - n.z * sn, sz = n.x * sn + n.z * cs;
    const float swirl_noise = simplex3({sx * 5.0f + t * 0.2f, n.y * 5.0f + t * 0.2f, sz * 5.0f + t * 0.2f}) * swirl * 0.05f;
    const float wave = 0.02f * swirl * sinf(n.y * scale * 8.0f - t * 2.0f) * (1.0f - fabsf(n.y));
    return a.base_nm * (1.0f - drainage + organic + swirl_noise - wave);
}

// Thickness of the drainage field at the surface normal: n = (sin th cos ph, cos th, sin th sin ph)
// (geometry.rs:184-186), blended over the four surrounding cells like get_thickness (drainage.rs:516-552):
// theta clamped to [0, pi] with the row index capped at N-2, phi wrapped.
__device__ float film_thickness_field(const FrameArgs &a, F3 n)
{
    const double PI = 3.14159265358979323846;
    const double theta = acos(fmin(fmax((double)n.y, -1.0), 1.0));
    double phi = atan2((double)n.z, (double)n.x);
    if (phi < 0.0) phi += 2.0 * PI;
    if (phi >= 2.0 * PI) phi -= 2.0 * PI;
    const double dth = PI / (double)(a.num_theta - 1), dph = 2.0 * PI / (double)a.num_phi;
    const double tf_ = theta / dth, pf_ = phi / dph;
    int ti = (int)floor(tf_);
    ti = min(max(ti, 0), a.num_theta - 2);
    int pi0 = (int)floor(pf_) % a.num_phi;
    const int pi1 = (pi0 + 1) % a.num_phi;
    const double ft = tf_ - (double)ti, fp = pf_ - floor(pf_);
    const long long r0 = (long long)ti * a.num_phi, r1 = r0 + a.num_phi;
    double v00, v01, v10, v11;
    if (a.field_f64) {
        const double *f = static_cast<const double *>(a.field);
        v00 = f[r0 + pi0]; v01 = f[r0 + pi1]; v10 = f[r1 + pi0]; v11 = f[r1 + pi1];
    } else {
        const float *f = static_cast<const float *>(a.field);
        v00 = f[r0 + pi0]; v01 = f[r0 + pi1]; v10 = f[r1 + pi0]; v11 = f[r1 + pi1];
    }
    const double v0 = v00 * (1.0 - fp) + v01 * fp, v1 = v10 * (1.0 - fp) + v11 * fp;
    return (float)((v0 * (1.0 - ft) + v1 * ft + a.h_ref) * 1e9);
}

__device__ __forceinline__ float srgb_encode(float c)
{
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}

__device__ __forceinline__ uint32_t pack_srgb(float r, float g, float b)
{
    const uint32_t R = (uint32_t)floorf(srgb_encode(r) * 255.0f + 0.5f), G = (uint32_t)floorf(srgb_encode(g) * 255.0f + 0.5f),
                   B = (uint32_t)floorf(srgb_encode(b) * 255.0f + 0.5f);
    return R | (G << 8) | (B << 16) | 0xff000000u;
}

template <bool FIELD>
__global__ void __launch_bounds__(256) k_bubble_frame(const FrameArgs a)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= a.width || py >= a.height) return;
    uint32_t *dst = a.out + (size_t)py * a.width + px;
    // pixel centre in NDC, row 0 at the top
    const float X = ((float)px + 0.5f) / (float)a.width * 2.0f - 1.0f;
    const float Y = 1.0f - ((float)py + 0.5f) / (float)a.height * 2.0f;
    const float kx = X * a.tan_half_x, ky = Y * a.tan_half_y;
    F3 d = {a.fwd[0] + kx * a.right[0] + ky * a.up[0], a.fwd[1] + kx * a.right[1] + ky * a.up[1],
            a.fwd[2] + kx * a.right[2] + ky * a.up[2]};
    const float dl = rsqrtf(dot3(d, d));
    d = {d.x * dl, d.y * dl, d.z * dl};
    const F3 cam = {a.cam[0], a.cam[1], a.cam[2]};
    const float b = dot3(d, cam);
    const float disc = b * b - (a.cam2 - a.radius * a.radius);
    if (!(disc > 0.0f)) {
        *dst = pack_srgb(a.bg[0], a.bg[1], a.bg[2]);
        return;
    }
    const float tn = -b - sqrtf(disc);
    const F3 p = {cam.x + tn * d.x, cam.y + tn * d.y, cam.z + tn * d.z};
    const float pl = rsqrtf(dot3(p, p));
    const F3 n = {p.x * pl, p.y * pl, p.z * pl};
    F3 vd = {cam.x - p.x, cam.y - p.y, cam.z - p.z};
    const float vl = rsqrtf(dot3(vd, vd));
    vd = {vd.x * vl, vd.y * vl, vd.z * vl};
    const float cos_theta = fabsf(dot3(vd, n));  // bubble.wgsl:490-494

    const float thickness = FIELD ? film_thickness_field(a, n) : film_thickness_procedural(a, n);

    const float edge = 1.0f - cos_theta;         // :501-513
    float smooth;
    if (a.edge_mode == 1) smooth = edge * edge * (3.0f - 2.0f * edge);
    else if (a.edge_mode == 2) smooth = powf(edge, 1.5f);
    else smooth = edge;
    float alpha = a.base_alpha + a.edge_alpha * smooth;
    float r, g, bl;
    if (thickness < 30.0f) {                     // :517-519
        r = g = bl = 0.02f;
        alpha *= 0.5f;
    } else {                                     // :425-439, :528-531
        const float u = fminf(fmaxf(thickness * a.inv_max_nm, 0.0f), 1.0f), v = fminf(fmaxf(cos_theta, 0.0f), 1.0f);
        const float4 t = tex2D<float4>(a.lut, u, v);
        r = 0.95f * 0.1f + fminf(fmaxf(t.x * a.intensity, 0.0f), 1.0f);
        g = 0.97f * 0.1f + fminf(fmaxf(t.y * a.intensity, 0.0f), 1.0f);
        bl = 1.0f * 0.1f + fminf(fmaxf(t.z * a.intensity, 0.0f), 1.0f);
    }
    // unorm target: source colour and alpha are clamped to [0,1] before src*a + dst*(1-a)
    alpha = fminf(fmaxf(alpha, 0.0f), 1.0f);
    r = fminf(fmaxf(r, 0.0f), 1.0f); g = fminf(fmaxf(g, 0.0f), 1.0f); bl = fminf(fmaxf(bl, 0.0f), 1.0f);
    const float ia = 1.0f - alpha;
    *dst = pack_srgb(r * alpha + a.bg[0] * ia, g * alpha + a.bg[1] * ia, bl * alpha + a.bg[2] * ia);
}

}  // namespace

cudaError_t launch_bubble_frame(const FrameArgs &a, cudaStream_t st)
{
    const dim3 block(32, 8), grid((unsigned)((a.width + 31) / 32), (unsigned)((a.height + 7) / 8));
    if (a.field) k_bubble_frame<true><<<grid, block, 0, st>>>(a);
    else k_bubble_frame<false><<<grid, block, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sbx
