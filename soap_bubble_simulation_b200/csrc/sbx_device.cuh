// sbx_device.cuh -- device-side building blocks: vector loads, the per-cell drainage update in its
// fast and strict forms, TFE membership, and the colour stage.
#pragma once

#include "sbx_internal.cuh"

namespace sbx {

// ------------------------------------------------------------------ 16-byte vectors
template <typename T, int VEC>
struct Vec {
    T v[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> ldv(const T *p)
{
    Vec<T, VEC> r;
    if constexpr (VEC == 1) {
        r.v[0] = __ldg(p);
    } else if constexpr (sizeof(T) == 4) {
        static_assert(VEC == 4, "float vectors are 4 wide");
        float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    } else {
        static_assert(VEC == 2, "double vectors are 2 wide");
        double2 q = __ldg(reinterpret_cast<const double2 *>(p));
        r.v[0] = q.x; r.v[1] = q.y;
    }
    return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void stv(T *p, const Vec<T, VEC> &r)
{
    if constexpr (VEC == 1) {
        *p = r.v[0];
    } else if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else {
        *reinterpret_cast<double2 *>(p) = make_double2(r.v[0], r.v[1]);
    }
}

// ------------------------------------------------------------------ drainage update, fast form
// dh/dt = -(rho g/3 eta) h^3 sin(theta) + D/R^2 [ d2h/dth2 + cot(theta) dh/dth + d2h/dph2 / sin^2 ]
// (src/physics/drainage.rs:347-372) with the divisions folded into per-row coefficients
//   G = (rho g / 3 eta) sin(theta),  A = D/(R^2 dth^2),  B = D cot(theta)/(R^2 2 dth),
//   E = D/(R^2 dph^2 sin^2(theta)).
// `h` is the stored value (h_abs - h_ref).  Differences are taken neighbour-minus-centre so that they
// are exact in the storage type (Sterbenz) and sub-ulp increments are not lost before the final add.
template <typename T>
__device__ __forceinline__ T cell_update(T h, T up, T dn, T lf, T rt, T G, T A, T B, T E, T dscale,
                                         T href, T dt, T crit)
{
    const T habs = h + href;
    const T dN = dn - h, dS = up - h;
    const T dE = rt - h, dW = lf - h;
    const T t1 = dN + dS, t2 = dN - dS, t3 = dE + dW;
    T diff = A * t1;
    diff = fma(B, t2, diff);
    diff = fma(E, t3, diff);
    const T rate = fma(diff, dscale, -G * (habs * habs * habs));
    T hn = fma(dt, rate, h);
    hn = hn > -href ? hn : -href;            // max(h_abs, 0), drainage.rs:378
    return habs < crit ? h : hn;             // below-critical cells keep their value, drainage.rs:331-333
}

// strict form: the reference's operation order with IEEE round-to-nearest ops that nvcc never
// contracts into FMAs (drainage.rs:347-379).
struct StrictRow {
    double sin_theta, cot, sin2;
};

__device__ __forceinline__ double cell_update_strict(double hc, double h_tm, double h_tp, double h_pm,
                                                     double h_pp, const StrictRow &r, double coef,
                                                     double dth2, double two_dth, double dph2, double r2,
                                                     double D, double dt, double crit)
{
    if (hc < crit) return hc;
    const double cubed = __dmul_rn(__dmul_rn(hc, hc), hc);
    const double drainage_term = __dmul_rn(__dmul_rn(-coef, cubed), r.sin_theta);
    const double two_h = __dmul_rn(2.0, hc);
    const double d2t = __ddiv_rn(__dadd_rn(__dsub_rn(h_tp, two_h), h_tm), dth2);
    const double dht = __ddiv_rn(__dsub_rn(h_tp, h_tm), two_dth);
    const double d2p = __ddiv_rn(__dadd_rn(__dsub_rn(h_pp, two_h), h_pm), dph2);
    const double lap =
        __ddiv_rn(__dadd_rn(__dadd_rn(d2t, __dmul_rn(r.cot, dht)), __ddiv_rn(d2p, r.sin2)), r2);
    const double dh_dt = __dadd_rn(drainage_term, __dmul_rn(D, lap));
    const double nh = __dadd_rn(hc, __dmul_rn(dt, dh_dt));
    return nh > 0.0 ? nh : 0.0;
}

// ------------------------------------------------------------------ TFE membership (integer)
__device__ __forceinline__ unsigned front_row_mask(const FrontTable &ft, int grow)
{
    unsigned m = 0;
    if (grow < ft.row_min || grow > ft.row_max) return 0;
#pragma unroll
    for (int k = 0; k < kMaxFronts; ++k)
        if (k < ft.count && grow >= ft.f[k].row_lo && grow <= ft.f[k].row_hi) m |= 1u << k;
    return m;
}

__device__ __forceinline__ bool front_hit(const FrontTable &ft, unsigned rowmask, int col)
{
    bool hit = false;
#pragma unroll
    for (int k = 0; k < kMaxFronts; ++k) {
        if (rowmask & (1u << k)) {
            const FrontCells &f = ft.f[k];
            hit |= (col >= f.c0_lo && col <= f.c0_hi) || (col >= f.c1_lo && col <= f.c1_hi);
        }
    }
    return hit;
}

// fronts whose phi interval contains column `col` (bit per front); independent of the row
__device__ __forceinline__ unsigned front_col_mask(const FrontTable &ft, int col)
{
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < kMaxFronts; ++k) {
        const FrontCells &f = ft.f[k];
        if (k < ft.count && ((col >= f.c0_lo && col <= f.c0_hi) || (col >= f.c1_lo && col <= f.c1_hi))) m |= 1u << k;
    }
    return m;
}

// ------------------------------------------------------------------ colour stage (f32)
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// 1/sqrt(x) on the SFU without the denormal pre/post-scaling rsqrtf() adds (|cam - R n|^2 is never denormal)
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// cos(theta_i) = |n.(c - R n)| / |c - R n| with n.c = sin(th) A_j + cos(th) c_y
__device__ __forceinline__ float view_cos(float sin_th, float cos_th_cy, float colA, float R, float cam2)
{
    const float nc = fmaf(sin_th, colA, cos_th_cy);
    const float len2 = fmaf(-2.f * R, nc, cam2 + R * R);
    return fabsf(nc - R) * rsqrt_approx(len2);
}

// approximate SFU ops (1-2 ulp), one MUFU each
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// sin(pi x) for |x| <= 0.5 from the Taylor series of sin(pi x)/x in x^2 up to x^10 (truncation error
// < 6e-8 at |x| = 0.5, relative accuracy kept near 0).  The fringes' zeros need *relative* accuracy:
// the 1/2.2 gamma turns an absolute 1e-8 intensity error into 2e-4 of colour.
__device__ __forceinline__ float sinpi_half(float x)
{
    const float x2 = x * x;
    float p = -7.3704310e-3f;             // -pi^11/11!
    p = fmaf(p, x2, 8.2145887e-2f);       //  pi^9/9!
    p = fmaf(p, x2, -5.9926453e-1f);      // -pi^7/7!
    p = fmaf(p, x2, 2.5501640f);          //  pi^5/5!
    p = fmaf(p, x2, -5.1677128f);         // -pi^3/3!
    p = fmaf(p, x2, 3.1415927f);          //  pi
    return p * x;
}

// Fringe phase in cycles, reduced to [-0.5, 0.5], for the three colour wavelengths:
//   cyc_k = 2 n d cos(theta_t) / lambda_k          (interference.rs:226-254 without the lambda/2 term)
// evaluated in f64.  Near grazing incidence R -> 1 and the Airy fringes become needle-sharp (finesse
// ~ 1/(1-R)): an fp32 phase (relative error 1e-7 on ~10 cycles) would move the colour by > 1e-4 there.
// cos(theta_t) is an fp32 SFU square root polished by one Newton step in f64.
struct Phase3 {
    float y[3];  // distance of the phase from the nearest fringe zero, in cycles: 1/2 - |frac(cyc)|
    float ct;    // cos(theta_t)
};

__device__ __forceinline__ Phase3 phase3(double d_nm, float ci, const RgbArgs &a)
{
    const double cid = (double)ci;
    const double x = fma(-(1.0 - cid * cid), a.inv_n2_d, 1.0);  // 1 - sin^2(theta_t), interference.rs:199-210
    const float ct0 = sqrt_approx(fmaxf((float)x, 1e-30f));
    const double c0 = (double)ct0;
    const double ctd = fma(x - c0 * c0, (double)(0.5f * rcp_approx(ct0)), c0);
    const double base = d_nm * ctd;
    Phase3 p;
    p.ct = (float)ctd;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double cyc = base * a.two_n_over_wl[k];
        p.y[k] = (float)(0.5 - fabs(cyc - rint(cyc)));  // formed in f64: keeps relative accuracy at the zeros
    }
    return p;
}

// The same phase without f64: float-float (hi + lo pairs of f32) arithmetic on the FP32 pipe.  On this part a cell's
// f64 phase costs ~25 half-rate DP operations plus ~10 f32<->f64 conversions on the quarter-rate pipe the SFU ops
// share; the pairs below are ~60 full-rate f32 operations, exact products through FMA residuals (Dekker / Knuth
// two-sum, two-product), ~2^-45 relative on the cycle count.  __fmul_rn / __fadd_rn keep nvcc from contracting the
// error-free transformations into FMAs.  Input: the STORED thickness h and its offset h_ref (f32 storage only).
__device__ __forceinline__ Phase3 phase3_ff(float h, float href, float ci, const RgbArgs &a)
{
    // d = (h + h_ref) * 1e9 nm as a pair (1e9 is exact in f32)
    const float s = __fadd_rn(h, href);
    const float bb = __fsub_rn(s, h);
    const float se = __fadd_rn(__fsub_rn(h, __fsub_rn(s, bb)), __fsub_rn(href, bb));
    const float d_hi = __fmul_rn(s, 1e9f);
    const float d_lo = fmaf(se, 1e9f, fmaf(s, 1e9f, -d_hi));
    // x = 1 - (1 - ci^2) / n^2   (1 - sin^2(theta_t), interference.rs:199-210)
    const float c2 = __fmul_rn(ci, ci), c2e = fmaf(ci, ci, -c2);
    const float t = __fsub_rn(1.0f, c2);
    const float tl = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, t), c2), c2e);  // 1 >= ci^2: fast two-sum
    const float m = __fmul_rn(t, a.inv_n2_hi);
    const float ml = fmaf(tl, a.inv_n2_hi, fmaf(t, a.inv_n2_lo, fmaf(t, a.inv_n2_hi, -m)));
    const float x = __fsub_rn(1.0f, m);
    const float xl = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, x), m), ml);
    // cos(theta_t) = sqrt(x): one SFU reciprocal square root polished by a Newton step on the pair
    const float rs = rsqrt_approx(fmaxf(x, 1e-30f));
    const float c0 = __fmul_rn(x, rs);
    const float q = __fmul_rn(c0, c0), qe = fmaf(c0, c0, -q);
    const float res = __fadd_rn(__fsub_rn(__fsub_rn(x, q), qe), xl);
    const float cl = __fmul_rn(res, 0.5f * rs);
    // base = d * cos(theta_t)
    const float b_hi = __fmul_rn(d_hi, c0);
    const float b_lo = fmaf(d_lo, c0, fmaf(d_hi, cl, fmaf(d_hi, c0, -b_hi)));
    Phase3 p;
    p.ct = __fadd_rn(c0, cl);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float P = __fmul_rn(b_hi, a.w_hi[k]);
        const float E = fmaf(b_lo, a.w_hi[k], fmaf(b_hi, a.w_lo[k], fmaf(b_hi, a.w_hi[k], -P)));
        const float r = __fsub_rn(__fadd_rn(P, 12582912.0f), 12582912.0f);  // rint(P) for |P| < 2^22 cycles
        const float g = __fsub_rn(P, r);                                      // exact
        const float y0 = __fsub_rn(0.5f, fabsf(g));                           // exact where it matters (|g| >= 1/4)
        // y = 1/2 - |g + E| = y0 - sign(g) E: the tail joins after the cancellation, so y keeps relative accuracy at the zeros
        const float Es = __uint_as_float(__float_as_uint(E) ^ (__float_as_uint(g) & 0x80000000u));
        p.y[k] = __fsub_rn(y0, Es);
    }
    return p;
}

// sin^2(pi (cyc + 1/2)) = cos^2(pi fr) = sin^2(pi (1/2 - |fr|)): the lambda/2 reflection shift
// (interference.rs:235-237) folded in without adding 0.5 to a rounded number
__device__ __forceinline__ float fringe_sin2(float y)
{
    const float s = sinpi_half(y);
    return s * s;
}

// InterferenceCalculator::calculate_interference_color[_linear] (src/physics/interference.rs:380-426):
// Snell (197-211), exact Fresnel average (266-298), OPD with lambda/2 (226-238), multi-bounce Airy
// (344-352) with 1 - cos(phi) = 2 sin^2(phi/2).  1 - R is formed without cancellation from the
// transmittances: 1 - r_s^2 = 4 ci n ct / (ci + n ct)^2, 1 - r_p^2 = 4 n ci ct / (n ci + ct)^2.
__device__ __forceinline__ void rgb_exact3_from_phase(const Phase3 &ph, bool positive, float ci, const RgbArgs &a, bool gamma,
                                                      float rgb[3]);
__device__ __forceinline__ void rgb_exact3(double d_nm, float ci, const RgbArgs &a, bool gamma, float rgb[3])
{
    ci = fminf(fmaxf(ci, -1.f), 1.f);
    rgb_exact3_from_phase(phase3(d_nm, ci, a), d_nm > 0.0, ci, a, gamma, rgb);
}
// f32 storage on the fast paths: the float-float phase from the stored value and its offset
__device__ __forceinline__ void rgb_exact3_ff(float h, float href, float ci, const RgbArgs &a, bool gamma, float rgb[3])
{
    ci = fminf(fmaxf(ci, -1.f), 1.f);
    rgb_exact3_from_phase(phase3_ff(h, href, ci, a), __fadd_rn(h, href) > 0.f, ci, a, gamma, rgb);
}
__device__ __forceinline__ void rgb_exact3_from_phase(const Phase3 &ph, bool positive, float ci, const RgbArgs &a, bool gamma,
                                                      float rgb[3])
{
    const float n = a.n_film, ct = ph.ct;
    const float ds = fmaf(n, ct, ci), dp = fmaf(n, ci, ct);
    const float inv = rcp_approx(ds * dp);
    const float P = 4.f * n * ci * ct;
    const float omr = 0.5f * P * inv * inv * fmaf(ds, ds, dp * dp);  // 1 - R
    const float R = 1.f - omr;
    const float omr2 = omr * omr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float num = 4.f * R * fringe_sin2(ph.y[k]);
        const float den = omr2 + num;
        float I = den < 1e-30f ? R : num * rcp_approx(den);  // |den| < eps -> R, interference.rs:348-350
        I = positive ? clamp01(I) : 0.f;                      // d <= 0 -> 0, interference.rs:319-321
        if (gamma) I = ex2_approx(lg2_approx(I) * (1.0f / 2.2f));  // powf(1/2.2); lg2(0) = -inf -> 0
        rgb[k] = clamp01(I);
    }
}

// wall.wgsl:165-207
__device__ __forceinline__ void rgb_schlick_wall(double d_nm, float ci, const RgbArgs &a, float rgb[3])
{
    const Phase3 ph = phase3(d_nm, ci, a);
    const float n = a.n_film;
    const float q = (1.f - n) / (1.f + n);
    const float r0 = q * q;
    const float m = 1.f - ci, m2 = m * m;
    const float R = r0 + (1.f - r0) * (m2 * m2 * m);
    const float omr = 1.f - R;
    const float F = 4.f * R * rcp_approx(omr * omr);
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[k] = rcp_approx(fmaf(F, fringe_sin2(ph.y[k]), 1.f));
}

// The documented original model (docs/key-insights.md:288-309): two-beam interference with the pi reflection shift,
// I = (1 + cos(2 pi cyc + pi))/2 = sin^2(pi cyc), cyc = 2 n d cos(theta_t)/lambda, times Schlick Fresnel, times 4.
// Low finesse, so f32 cycles and the SFU sine on the reduced argument are ample for the 1e-4 colour tolerance.
__device__ __forceinline__ void rgb_two_beam(float d_nm, float ci, const RgbArgs &a, float rgb[3])
{
    // I_k = 4 F (1 + cos(2 pi OPD / lambda_k + pi)) / 2 = 2 F (1 - cos(2 pi OPD / lambda_k)), docs/key-insights.md:288-309.
    // The SFU cosine takes its argument in revolutions and drops the integer part itself, so no explicit range
    // reduction is needed: the phase (< 80 rad) carries 5e-6 rad of fp32 rounding -- as in the f32 reference shader --
    // which is < 2e-5 of colour even at grazing incidence (tolerance 1e-4).
    const float s2 = (1.f - ci * ci) * a.inv_n2;
    const float ct = sqrt_approx(fmaxf(1.f - s2, 0.f));
    const float opd = a.two_n * d_nm * ct;
    const float m = 1.f - ci, m2 = m * m;
    const float f2 = 2.f * fmaf(a.omr0, m2 * m2 * m, a.r0);
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[k] = __saturatef(fmaf(-f2, __cosf(opd * a.two_pi_over_wl_f[k]), f2));
}

// bubble.wgsl:425-439 through the texture unit: linear filter, clamp-to-edge, normalised coordinates,
// unorm8 -> float (sampler at pipeline.rs:350-359, texture format Rgba8Unorm at pipeline.rs:325)
__device__ __forceinline__ void rgb_lut(float d_nm, float ci, const RgbArgs &a, float rgb[3])
{
    const float u = clamp01(d_nm * a.inv_max_nm);
    const float v = clamp01(ci);
    const float4 t = tex2D<float4>(a.lut, u, v);
    rgb[0] = clamp01(t.x * a.intensity);
    rgb[1] = clamp01(t.y * a.intensity);
    rgb[2] = clamp01(t.z * a.intensity);
}

// Direct spectral evaluation per cell (no LUT): thin_film_interference of the fragment shaders
// (bubble_instanced.wgsl:190-218 == interference_lut.rs:45-93) over a.n_spec wavelengths.  Per wavelength the shader's
// sin^2(phase / 2) with phase = 2 pi opd / lambda + pi is cos^2(pi opd / lambda) = (1 + cos(2 pi opd / lambda)) / 2, and
// the colour-matching weights are folded with the XYZ -> sRGB matrix and the 1 / n normalisation on the host, so a
// wavelength costs one SFU cosine, half an SFU reciprocal (two Airy denominators share one) and ~10 FP32 operations.
// f32 like the shader; the order of the (linear) accumulation differs from it by rounding only.
__device__ __forceinline__ void rgb_spectral(float d_nm, float ci, const RgbArgs &a, float rgb[3])
{
    const float sin_i = sqrt_approx(fmaxf(1.0f - ci * ci, 0.0f));          // snells_law, interference_lut.rs:96-100
    const float sin_t = sin_i * rcp_approx(a.n_film);
    const float ct = sqrt_approx(fmaxf(1.0f - sin_t * sin_t, 0.0f));
    const float n2ct = a.n_film * ct, n2ci = a.n_film * ci;                // fresnel_unpolarized(ci, ct, 1, n), :103-120
    const float rs = (ci - n2ct) * rcp_approx(fmaxf(ci + n2ct, 0.0001f));
    const float rp = (n2ci - ct) * rcp_approx(fmaxf(n2ci + ct, 0.0001f));
    const float R = (rs * rs + rp * rp) * 0.5f;
    const float omr = fmaxf(1.0f - R, 0.001f);                             // airy_interference, :123-133
    const float Fh = 2.0f * R * rcp_approx(omr * omr);                     // F / 2
    const float opd = a.two_n * d_nm * ct;                                 // :61
    float r = 0.f, g = 0.f, b = 0.f;
    const float4 *sp = a.spec;
    int k = 0;
    for (; k + 1 < a.n_spec; k += 2) {
        const float4 s0 = __ldg(sp + k), s1 = __ldg(sp + k + 1);
        const float f0 = fmaf(Fh, __cosf(opd * s0.x), Fh), f1 = fmaf(Fh, __cosf(opd * s1.x), Fh);  // F sin^2
        const float d0 = 1.0f + f0, d1 = 1.0f + f1;                        // >= 1: the shader's max(., 0.001) never binds
        const float inv = rcp_approx(d0 * d1);
        const float i0 = f0 * (inv * d1), i1 = f1 * (inv * d0);
        r = fmaf(i0, s0.y, r); g = fmaf(i0, s0.z, g); b = fmaf(i0, s0.w, b);
        r = fmaf(i1, s1.y, r); g = fmaf(i1, s1.z, g); b = fmaf(i1, s1.w, b);
    }
    if (k < a.n_spec) {
        const float4 s0 = __ldg(sp + k);
        const float f0 = fmaf(Fh, __cosf(opd * s0.x), Fh);
        const float i0 = f0 * rcp_approx(1.0f + f0);
        r = fmaf(i0, s0.y, r); g = fmaf(i0, s0.z, g); b = fmaf(i0, s0.w, b);
    }
    rgb[0] = clamp01(r * a.intensity);
    rgb[1] = clamp01(g * a.intensity);
    rgb[2] = clamp01(b * a.intensity);
}

// The same evaluation for the VEC cells a thread of the ring kernels owns: wavelengths in the OUTER loop, so a table entry
// is loaded once per VEC cells and the VEC independent Airy chains interleave (one cell at a time the loop is a single
// dependent SFU chain: 25 Gcell/s at 64 wavelengths; interleaved ~1.6x that).
template <int VEC>
__device__ __forceinline__ void rgb_spectral_vec(const float *d_nm, const float *ci, const RgbArgs &a, float (*rgb)[3])
{
    float Fh[VEC], opd[VEC], r[VEC], g[VEC], b[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        const float sin_i = sqrt_approx(fmaxf(1.0f - ci[c] * ci[c], 0.0f));
        const float sin_t = sin_i * rcp_approx(a.n_film);
        const float ct = sqrt_approx(fmaxf(1.0f - sin_t * sin_t, 0.0f));
        const float n2ct = a.n_film * ct, n2ci = a.n_film * ci[c];
        const float rs = (ci[c] - n2ct) * rcp_approx(fmaxf(ci[c] + n2ct, 0.0001f));
        const float rp = (n2ci - ct) * rcp_approx(fmaxf(n2ci + ct, 0.0001f));
        const float R = (rs * rs + rp * rp) * 0.5f;
        const float omr = fmaxf(1.0f - R, 0.001f);
        Fh[c] = 2.0f * R * rcp_approx(omr * omr);
        opd[c] = a.two_n * d_nm[c] * ct;
        r[c] = g[c] = b[c] = 0.f;
    }
    const float4 *sp = a.spec;
    int k = 0;
#pragma unroll 2
    for (; k + 1 < a.n_spec; k += 2) {
        const float4 s0 = __ldg(sp + k), s1 = __ldg(sp + k + 1);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            const float f0 = fmaf(Fh[c], __cosf(opd[c] * s0.x), Fh[c]), f1 = fmaf(Fh[c], __cosf(opd[c] * s1.x), Fh[c]);
            const float d0 = 1.0f + f0, d1 = 1.0f + f1;
            const float inv = rcp_approx(d0 * d1);
            const float i0 = f0 * (inv * d1), i1 = f1 * (inv * d0);
            r[c] = fmaf(i0, s0.y, r[c]); g[c] = fmaf(i0, s0.z, g[c]); b[c] = fmaf(i0, s0.w, b[c]);
            r[c] = fmaf(i1, s1.y, r[c]); g[c] = fmaf(i1, s1.z, g[c]); b[c] = fmaf(i1, s1.w, b[c]);
        }
    }
    if (k < a.n_spec) {
        const float4 s0 = __ldg(sp + k);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            const float f0 = fmaf(Fh[c], __cosf(opd[c] * s0.x), Fh[c]);
            const float i0 = f0 * rcp_approx(1.0f + f0);
            r[c] = fmaf(i0, s0.y, r[c]); g[c] = fmaf(i0, s0.z, g[c]); b[c] = fmaf(i0, s0.w, b[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        rgb[c][0] = clamp01(r[c] * a.intensity);
        rgb[c][1] = clamp01(g[c] * a.intensity);
        rgb[c][2] = clamp01(b[c] * a.intensity);
    }
}

// thickness in nm handed to the colour stage: f32 is enough for the LUT fetch, the analytic modes need
// the f64 phase (see phase3)
template <int MODE> struct Dnm { using type = double; };
template <> struct Dnm<SBX_RGB_LUT> { using type = float; };
template <> struct Dnm<SBX_RGB_NONE> { using type = float; };
template <> struct Dnm<SBX_RGB_SCHLICK_TWO_BEAM> { using type = float; };
template <> struct Dnm<SBX_RGB_SPECTRAL> { using type = float; };

template <int MODE>
__device__ __forceinline__ void rgb_eval(typename Dnm<MODE>::type d_nm, float ci, const RgbArgs &a, float rgb[3])
{
    if constexpr (MODE == SBX_RGB_LUT) rgb_lut(d_nm, ci, a, rgb);
    else if constexpr (MODE == SBX_RGB_EXACT3) rgb_exact3(d_nm, ci, a, true, rgb);
    else if constexpr (MODE == SBX_RGB_EXACT3_LINEAR) rgb_exact3(d_nm, ci, a, false, rgb);
    else if constexpr (MODE == SBX_RGB_SCHLICK_WALL) rgb_schlick_wall(d_nm, ci, a, rgb);
    else if constexpr (MODE == SBX_RGB_SCHLICK_TWO_BEAM) rgb_two_beam(d_nm, ci, a, rgb);
    else if constexpr (MODE == SBX_RGB_SPECTRAL) rgb_spectral(d_nm, ci, a, rgb);
}

// RgbColor::to_u8_tuple (interference.rs:73-79): round(255 x), alpha 255
__device__ __forceinline__ uint32_t pack_rgba8(const float rgb[3])
{
    // f64::round is half-away-from-zero; inputs are in [0,1] so floor(255 x + 0.5) matches it
    const uint32_t r = __float2uint_rz(fmaf(rgb[0], 255.f, 0.5f));
    const uint32_t g = __float2uint_rz(fmaf(rgb[1], 255.f, 0.5f));
    const uint32_t b = __float2uint_rz(fmaf(rgb[2], 255.f, 0.5f));
    return r | (g << 8) | (b << 16) | 0xff000000u;
}

// RGBA8 from three floats in [0,1]: the low mantissa byte of (255 x + 2^23) is round(255 x); the blue
// channel's magic constant also plants 0xff in byte 1, which becomes alpha.  Two PRMTs gather the bytes.
__device__ __forceinline__ uint32_t pack_rgba8_fast(float r, float g, float b)
{
    const uint32_t ur = __float_as_uint(fmaf(r, 255.f, 8388608.0f));
    const uint32_t ug = __float_as_uint(fmaf(g, 255.f, 8388608.0f));
    const uint32_t ub = __float_as_uint(fmaf(b, 255.f, 8453888.0f));  // 2^23 + 0xff00
    const uint32_t rg = __byte_perm(ur, ug, 0x0040);                  // bytes: r, g, -, -
    return __byte_perm(rg, ub, 0x5410);                               // bytes: r, g, b, 0xff
}

template <int MODE, typename T = float>
__device__ __forceinline__ uint32_t shade_px(T h, T href, float href_u, float colA, float sin_th, float ccy,
                                             float R, float camR2, const RgbArgs &rgb)
{
    const float nc = fmaf(sin_th, colA, ccy);      // n . cam
    const float len2 = fmaf(-2.f * R, nc, camR2);  // |cam - R n|^2
    if constexpr (MODE == SBX_RGB_LUT) {
        // u = clamp(d_nm / max_nm), v = clamp(cos theta_i): bubble.wgsl:429-433
        const float u = __saturatef(fmaf((float)h, rgb.nm_over_max, href_u));
        const float v = __saturatef(fabsf(nc - R) * rsqrt_approx(len2));
        const float4 t = tex2D<float4>(rgb.lut, u, v);
        return pack_rgba8_fast(__saturatef(t.x * rgb.intensity), __saturatef(t.y * rgb.intensity),
                               __saturatef(t.z * rgb.intensity));
    } else {
        const float ci = fabsf(nc - R) * rsqrt_approx(len2);
        float c[3];
#ifndef SBX_EXACT3_F64  // A/B builds: the f64 phase on the fast paths too
        if constexpr (sizeof(T) == 4 && (MODE == SBX_RGB_EXACT3 || MODE == SBX_RGB_EXACT3_LINEAR)) {
            rgb_exact3_ff((float)h, (float)href, ci, rgb, MODE == SBX_RGB_EXACT3, c);
            return pack_rgba8_fast(c[0], c[1], c[2]);
        }
#endif
        using DN = typename Dnm<MODE>::type;
        const DN d_nm = ((DN)h + (DN)href) * (DN)1e9;
        rgb_eval<MODE>(d_nm, ci, rgb, c);
        return pack_rgba8_fast(c[0], c[1], c[2]);
    }
}

// ------------------------------------------------------------------ caustic map (thickness-buffer consumer, row f2)
// caustics_compute.wgsl::main (src/render/shaders/caustics_compute.wgsl:29-126): 9-point stencil on the thickness
// buffer, phi wrapped, theta clamped at the poles (ghost rows supply the neighbours of a latitude band).
__device__ __forceinline__ float caustic_intensity(float center, float left, float right, float up, float down,
                                                   float ul, float ur, float dl, float dr, const sbx_caustic_params &p)
{
    // h = 1: the divisions by 2h, h^2 and 4h^2 are exact scalings (caustics_compute.wgsl:46-92)
    const float gx = (right - left) * 0.5f, gy = (down - up) * 0.5f;
    const float lr = left + right, ud = up + down;
    const float laplacian = ((lr + up) + down) - 4.0f * center;
    const float d2x = lr - 2.0f * center, d2y = ud - 2.0f * center;
    const float dxy = (((dr - dl) - ur) + ul) * 0.25f;
    const float hessian_det = __fsub_rn(__fmul_rn(d2x, d2y), __fmul_rn(dxy, dxy));
    const float curvature_factor = -laplacian * p.focal_length;                                   // :95-112
    const float focusing = rcp_approx(fmaxf(fabsf(1.0f - curvature_factor), 0.01f));
    const float branch_factor = hessian_det < -p.branch_threshold ? 1.5f : 1.0f;
    const float grad_mag = sqrt_approx(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    const float refraction_factor = fmaf(grad_mag, 0.5f, 1.0f);
    float intensity = focusing * branch_factor * refraction_factor;
    intensity = ex2_approx(lg2_approx(intensity) * p.caustic_sharpness);  // pow(x, sharpness), x >= 0.005
    intensity *= p.caustic_intensity;
    return fminf(fmaxf(intensity, 0.f), 10.f);
}


}  // namespace sbx
