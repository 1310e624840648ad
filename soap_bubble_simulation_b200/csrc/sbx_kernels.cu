// sbx_kernels.cu -- hand-written sm_100a kernels of the drainage + interference path.
//
// Main kernel: k_march.  A block owns a strip of `blockDim.x * VEC` consecutive phi columns and
// marches down `band_rows` theta rows keeping a three-row window in registers, so every thickness
// value is fetched from global memory once per band (plus one halo row each side of the band):
//   - loads/stores are 16-byte vectors, fully coalesced along phi;
//   - the phi neighbours of a thread's 16-byte packet come from the adjacent lanes by warp shuffle;
//     only the two edge lanes of a warp fetch one extra scalar, and the periodic seam (column 0 <->
//     num_phi-1) is the same scalar fetch with a wrapped column index;
//   - the colour stage runs on the freshly computed thickness in registers, so the field makes one
//     HBM round trip per step (read h, write h', write RGBA8): 12 B/cell in f32.
// Pole rows need the mean of the *updated* ring next to them (src/physics/drainage.rs:480-503): the
// blocks that update those rings emit fixed-slot partial sums and k_poles finishes in a fixed order.
#include <cfloat>

#include "sbx_device.cuh"
#include "sbx_launch.h"

namespace sbx {

int device_sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------ colour output helpers
template <int MODE, int VEC>
__device__ __forceinline__ void emit_rgb(const typename Dnm<MODE>::type *d_nm, const float *ci, const RgbArgs &rgb,
                                         long long cell0)
{
    if constexpr (MODE != SBX_RGB_NONE) {
        if (rgb.format == SBX_OUT_RGBA8) {
            uint32_t px[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float c[3];
                rgb_eval<MODE>(d_nm[k], ci[k], rgb, c);
                px[k] = pack_rgba8(c);
            }
            uint32_t *o = reinterpret_cast<uint32_t *>(rgb.out) + cell0;
            if constexpr (VEC == 4) *reinterpret_cast<uint4 *>(o) = make_uint4(px[0], px[1], px[2], px[3]);
            else if constexpr (VEC == 2) *reinterpret_cast<uint2 *>(o) = make_uint2(px[0], px[1]);
            else o[0] = px[0];
        } else {
            float *o = reinterpret_cast<float *>(rgb.out) + 3 * cell0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float c[3];
                rgb_eval<MODE>(d_nm[k], ci[k], rgb, c);
                o[3 * k + 0] = c[0];
                o[3 * k + 1] = c[1];
                o[3 * k + 2] = c[2];
            }
        }
    }
}

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double warp_part[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) warp_part[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) s += warp_part[i];  // fixed order
    }
    __syncthreads();
    return s;
}

// ------------------------------------------------------------------ the marching stencil kernel
template <typename T, int VEC, int MODE, bool TFE, bool STRICT>
__global__ void __launch_bounds__(256) k_march(const StepArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int N = a.num_phi;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const bool active = j0 < N;
    const int jc = active ? j0 : 0;
    const int b = blockIdx.z;
    const T *in = a.in + (long long)b * a.bubble_stride;
    T *out = a.out + (long long)b * a.bubble_stride;
    const int l0 = a.l_first + blockIdx.y * a.band_rows;
    const int l1 = min(l0 + a.band_rows - 1, a.l_last);  // inclusive
    if (l0 > l1) return;

    const bool edgeL = lane == 0;
    const bool edgeR = lane == 31 || jc + VEC >= N;
    const int jl = jc == 0 ? N - 1 : jc - 1;
    const int jr = jc + VEC >= N ? 0 : jc + VEC;

    const BubbleParam<T> bp = a.bub[b];
    const T href = bp.h_ref;

    float colA[VEC];
    if constexpr (MODE != SBX_RGB_NONE) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) colA[k] = __ldg(a.rgb.colA + jc + k);
    }
    const float Rv = bp.radius;
    unsigned colmask = 0;  // byte k: fronts whose phi interval covers column jc + k (does not depend on the row)
    if constexpr (TFE) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) colmask |= front_col_mask(a.fronts, jc + k) << (8 * k);
    }

    Vec<T, VEC> up = ldv<T, VEC>(in + (long long)(l0 - 1) * N + jc);
    Vec<T, VEC> cur = ldv<T, VEC>(in + (long long)l0 * N + jc);
    T eL = edgeL ? __ldg(in + (long long)l0 * N + jl) : T(0);
    T eR = edgeR ? __ldg(in + (long long)l0 * N + jr) : T(0);

    for (int l = l0; l <= l1; ++l) {
        const Vec<T, VEC> dn = ldv<T, VEC>(in + (long long)(l + 1) * N + jc);
        T eLn = T(0), eRn = T(0);
        if (l < l1) {
            if (edgeL) eLn = __ldg(in + (long long)(l + 1) * N + jl);
            if (edgeR) eRn = __ldg(in + (long long)(l + 1) * N + jr);
        }
        const int grow = a.row_begin + (l - 1);

        T left = __shfl_up_sync(0xffffffffu, cur.v[VEC - 1], 1);
        T right = __shfl_down_sync(0xffffffffu, cur.v[0], 1);
        if (edgeL) left = eL;
        if (edgeR) right = eR;

        Vec<T, VEC> nw;
        if constexpr (STRICT) {
            const double4 cx = a.coefX[grow];
            const StrictRow sr{cx.x, cx.y, cx.z};
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const T lf = k == 0 ? left : cur.v[k - 1];
                const T rt = k == VEC - 1 ? right : cur.v[k + 1];
                nw.v[k] = (T)cell_update_strict((double)cur.v[k], (double)up.v[k], (double)dn.v[k],
                                                (double)lf, (double)rt, sr, a.s_drain_coef, a.s_dth2,
                                                a.s_two_dth, a.s_dph2, a.s_r2, a.s_D, a.s_dt, a.s_crit);
            }
        } else {
            const T G = __ldg(a.coefS + 4 * grow + 0);
            const T B = __ldg(a.coefS + 4 * grow + 1);
            const T E = __ldg(a.coefS + 4 * grow + 2);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const T lf = k == 0 ? left : cur.v[k - 1];
                const T rt = k == VEC - 1 ? right : cur.v[k + 1];
                nw.v[k] = cell_update<T>(cur.v[k], up.v[k], dn.v[k], lf, rt, G, a.coefA, B, E,
                                         bp.diff_scale, href, a.dt, a.crit);
            }
        }

        // ring sums for the poles use the values before the TFE reduction (drainage.rs:384-389)
        if (grow == 1 || grow == a.num_theta - 2) {  // block-uniform
            double s = 0.0;
            if (active) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) s += (double)nw.v[k];
            }
            s = block_sum(s);
            if (threadIdx.x == 0) {
                double *pp = a.pole_partial + (long long)b * 2 * a.pole_slots;
                if (grow == 1) pp[blockIdx.x] = s;
                if (grow == a.num_theta - 2) pp[a.pole_slots + blockIdx.x] = s;
            }
        }

        if constexpr (TFE) {
            const unsigned rm = front_row_mask(a.fronts, grow);
            if (rm) {
                const unsigned m = (rm * 0x01010101u) & colmask;  // row mask x column mask, as in k_march_tma
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    if (m & (0xffu << (8 * k))) {
                        if constexpr (STRICT) nw.v[k] = (T)__dmul_rn((double)nw.v[k], a.s_tfe_factor);
                        else nw.v[k] = fma(-(nw.v[k] + href), a.tfe_rdt, nw.v[k]);  // h *= 1 - r dt, drainage.rs:453-455
                    }
                }
            }
        }

        if (active) {
            stv<T, VEC>(out + (long long)l * N + jc, nw);
            if constexpr (MODE != SBX_RGB_NONE) {
                const float2 sc = __ldg(a.rowSC + grow);
                const float ccy = sc.y * a.rgb.cam_y;
                using DN = typename Dnm<MODE>::type;
                DN d_nm[VEC];
                float ci[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    d_nm[k] = ((DN)nw.v[k] + (DN)href) * (DN)1e9;
                    ci[k] = view_cos(sc.x, ccy, colA[k], Rv, a.rgb.cam2);
                }
                emit_rgb<MODE, VEC>(d_nm, ci, a.rgb,
                                    (long long)b * a.rgb_bubble_stride + (long long)(l - 1) * N + jc);
            }
        }
        up = cur;
        cur = dn;
        eL = eLn;
        eR = eRn;
    }
}

// ------------------------------------------------------------------ poles
// thickness[0][*] = mean(updated ring 1), thickness[N-1][*] = mean(updated ring N-2)
// (src/physics/drainage.rs:480-503), then the TFE reduction if a front's row range reaches the pole
// row (drainage.rs:437-460), then the pole rows' colour.  grid = (2 * blocks_per_pole, batch); every
// block re-derives the (tiny) mean itself and writes a 1024-column slice of the row.
template <typename T, int MODE, bool STRICT>
__global__ void __launch_bounds__(256) k_poles(const StepArgs<T> a, int has_top, int has_bottom, int blocks_per_pole)
{
    const int which = blockIdx.x / blocks_per_pole;  // 0 top, 1 bottom
    const int col0 = (blockIdx.x % blocks_per_pole) * 1024;
    if ((which == 0 && !has_top) || (which == 1 && !has_bottom)) return;
    const int b = blockIdx.y;
    const int N = a.num_phi;
    T *out = a.out + (long long)b * a.bubble_stride;
    const int grow = which == 0 ? 0 : a.num_theta - 1;
    const int ring = which == 0 ? 1 : a.num_theta - 2;
    const int l = grow - a.row_begin + 1, lr = ring - a.row_begin + 1;
    const BubbleParam<T> bp = a.bub[b];

    __shared__ double s_avg;
    if (threadIdx.x == 0) {
        double sum = 0.0;
        if constexpr (STRICT) {
            // left-to-right f64 accumulation of the stored ring, as the reference does.  TFE fronts
            // never reach the pole-adjacent ring on grids this mode is meant for; the host falls back
            // to the separate TFE pass (k_tfe) when they would, so `out` still holds pre-TFE values.
            for (int j = 0; j < N; ++j) sum = __dadd_rn(sum, (double)out[(long long)lr * N + j]);
            s_avg = __ddiv_rn(sum, (double)N);
        } else {
            const double *pp = a.pole_partial + (long long)b * 2 * a.pole_slots + which * a.pole_slots;
            for (int i = 0; i < a.pole_slots; ++i) sum += pp[i];
            s_avg = sum / (double)N;
        }
    }
    __syncthreads();
    T v = (T)s_avg;
    const float2 sc = __ldg(a.rowSC + grow);
    const float ccy = sc.y * a.rgb.cam_y;
    unsigned rm = 0;
    if (a.fronts.count > 0) rm = front_row_mask(a.fronts, grow);
    for (int j = col0 + threadIdx.x; j < min(N, col0 + 1024); j += blockDim.x) {
        T w = v;
        if (rm && front_hit(a.fronts, rm, j)) {
            if constexpr (STRICT) w = (T)__dmul_rn((double)w, a.s_tfe_factor);
            else w = fma(-(w + bp.h_ref), a.tfe_rdt, w);
        }
        out[(long long)l * N + j] = w;
        if constexpr (MODE != SBX_RGB_NONE) {
            using DN = typename Dnm<MODE>::type;
            DN d_nm[1] = {((DN)w + (DN)bp.h_ref) * (DN)1e9};
            float ci[1] = {view_cos(sc.x, ccy, __ldg(a.rgb.colA + j), bp.radius, a.rgb.cam2)};
            emit_rgb<MODE, 1>(d_nm, ci, a.rgb, (long long)b * a.rgb_bubble_stride + (long long)(l - 1) * N + j);
        }
    }
}

// separate TFE pass in the reference's order (after the poles), used by the strict mode
template <typename T>
__global__ void k_tfe(T *buf, long long bubble_stride, int num_phi, int row_begin, int rows_local,
                      FrontTable ft, double factor)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int grow = ft.row_min + blockIdx.y;
    if (j >= num_phi || grow > ft.row_max) return;
    const int l = grow - row_begin + 1;
    if (l < 1 || l > rows_local) return;
    const unsigned rm = front_row_mask(ft, grow);
    if (rm && front_hit(ft, rm, j)) {
        T *p = buf + (long long)blockIdx.z * bubble_stride + (long long)l * num_phi + j;
        *p = (T)__dmul_rn((double)*p, factor);
    }
}

// ------------------------------------------------------------------ shade only (no stencil)
template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(256) k_shade(const StepArgs<T> a)
{
    const int N = a.num_phi;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (j0 >= N) return;
    const int b = blockIdx.z;
    const T *in = a.in + (long long)b * a.bubble_stride;
    const BubbleParam<T> bp = a.bub[b];
    float colA[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) colA[k] = __ldg(a.rgb.colA + j0 + k);
    const int l0 = a.l_first + blockIdx.y * a.band_rows;
    const int l1 = min(l0 + a.band_rows - 1, a.l_last);
    for (int l = l0; l <= l1; ++l) {
        const int grow = a.row_begin + (l - 1);
        const Vec<T, VEC> h = ldv<T, VEC>(in + (long long)l * N + j0);
        const float2 sc = __ldg(a.rowSC + grow);
        const float ccy = sc.y * a.rgb.cam_y;
        if constexpr (sizeof(T) == 4 && VEC == 4) {
            if (a.rgb.format == SBX_OUT_RGBA8) {  // streaming fast path: 4 pixels per thread, one 16-byte store
                const float camR2 = a.rgb.cam2 + bp.radius * bp.radius;
                const float href_u = bp.h_ref * a.rgb.nm_over_max;
                uint32_t px[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    px[k] = shade_px<MODE>(h.v[k], bp.h_ref, href_u, colA[k], sc.x, ccy, bp.radius, camR2, a.rgb);
                uint32_t *o = reinterpret_cast<uint32_t *>(a.rgb.out) + (long long)b * a.rgb_bubble_stride +
                              (long long)(l - 1) * N + j0;
                *reinterpret_cast<uint4 *>(o) = make_uint4(px[0], px[1], px[2], px[3]);
                continue;
            }
        }
        using DN = typename Dnm<MODE>::type;
        DN d_nm[VEC];
        float ci[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            d_nm[k] = ((DN)h.v[k] + (DN)bp.h_ref) * (DN)1e9;
            ci[k] = view_cos(sc.x, ccy, colA[k], bp.radius, a.rgb.cam2);
        }
        emit_rgb<MODE, VEC>(d_nm, ci, a.rgb, (long long)b * a.rgb_bubble_stride + (long long)(l - 1) * N + j0);
    }
}

// f32 -> RGBA8 streaming variant of k_shade: two rows per trip with both 16-byte loads issued before the first
// use and running pointers instead of per-row index arithmetic (the plain loop above has one dependent load per
// row; four rows per trip ran into tex_throttle, profiles/r01_k_shade_rgba8.txt).
template <int MODE>
__global__ void __launch_bounds__(256, 6) k_shade_rgba8(const StepArgs<float> a)
{
    const int N = a.num_phi;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= N) return;
    const int b = blockIdx.z;
    const BubbleParam<float> bp = a.bub[b];
    const float4 ca = __ldg(reinterpret_cast<const float4 *>(a.rgb.colA + j0));
    const float colA[4] = {ca.x, ca.y, ca.z, ca.w};
    const float camR2 = a.rgb.cam2 + bp.radius * bp.radius;
    const float href_u = bp.h_ref * a.rgb.nm_over_max;
    const int l0 = a.l_first + blockIdx.y * a.band_rows;
    const int l1 = min(l0 + a.band_rows - 1, a.l_last);
    const float *__restrict__ in_p = a.in + (long long)b * a.bubble_stride + (long long)l0 * N + j0;
    uint32_t *__restrict__ out_p = reinterpret_cast<uint32_t *>(a.rgb.out) + (long long)b * a.rgb_bubble_stride +
                                   (long long)(l0 - 1) * N + j0;
    const float2 *__restrict__ sc_p = a.rowSC + a.row_begin + (l0 - 1);
    auto shade_row = [&](const float4 &h, const float2 &sc, uint32_t *dst) {
        const float ccy = sc.y * a.rgb.cam_y;
        const float hv[4] = {h.x, h.y, h.z, h.w};
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            px[k] = shade_px<MODE>(hv[k], bp.h_ref, href_u, colA[k], sc.x, ccy, bp.radius, camR2, a.rgb);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(px[0], px[1], px[2], px[3]);
    };
    int left = l1 - l0 + 1;
    for (; left >= 2; left -= 2) {
        const float4 h0 = __ldg(reinterpret_cast<const float4 *>(in_p));
        const float4 h1 = __ldg(reinterpret_cast<const float4 *>(in_p + N));
        const float2 s0 = __ldg(sc_p), s1 = __ldg(sc_p + 1);
        shade_row(h0, s0, out_p);
        shade_row(h1, s1, out_p + N);
        in_p += 2 * (long long)N;
        out_p += 2 * (long long)N;
        sc_p += 2;
    }
    if (left) shade_row(__ldg(reinterpret_cast<const float4 *>(in_p)), __ldg(sc_p), out_p);
}

// ------------------------------------------------------------------ WGSL-semantics dispatch
// drainage.wgsl::drainage_step (src/render/shaders/drainage.wgsl:72-202): f32, thickness +
// concentration, Marangoni coupling, pole rows frozen (they are written once into both ping-pong
// buffers at upload/reset and never touched again).  Same marching structure with two windows.
template <int VEC, int MODE>
__global__ void __launch_bounds__(256) k_wgsl(const WgslArgs a)
{
    const int lane = threadIdx.x & 31;
    const int N = a.num_phi;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const bool active = j0 < N;
    const int jc = active ? j0 : 0;
    const int b = blockIdx.z;
    const float *hin = a.h_in + (long long)b * a.bubble_stride;
    const float *cin = a.c_in + (long long)b * a.bubble_stride;
    float *hout = a.h_out + (long long)b * a.bubble_stride;
    float *cout = a.c_out + (long long)b * a.bubble_stride;
    const int l0 = a.l_first + blockIdx.y * a.band_rows;
    const int l1 = min(l0 + a.band_rows - 1, a.l_last);
    if (l0 > l1) return;
    const bool edgeL = lane == 0;
    const bool edgeR = lane == 31 || jc + VEC >= N;
    const int jl = jc == 0 ? N - 1 : jc - 1;
    const int jr = jc + VEC >= N ? 0 : jc + VEC;

    const sbx_drainage_params &p = a.p;
    const float inv_dth2 = 1.f / (a.delta_theta * a.delta_theta);
    const float inv_2dth = 1.f / (2.f * a.delta_theta);
    const float inv_dph2 = 1.f / (a.delta_phi * a.delta_phi);
    const float inv_2dph = 1.f / (2.f * a.delta_phi);
    const float r2 = p.bubble_radius * p.bubble_radius;
    const float inv_r2 = 1.f / r2;
    const float inv_r = 1.f / p.bubble_radius;
    const float drainage_coeff = p.density * p.gravity / (3.0f * p.viscosity);
    const bool mar = p.marangoni_enabled != 0u;

    float colA[VEC];
    if constexpr (MODE != SBX_RGB_NONE) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) colA[k] = __ldg(a.rgb.colA + jc + k);
    }

    Vec<float, VEC> hu = ldv<float, VEC>(hin + (long long)(l0 - 1) * N + jc);
    Vec<float, VEC> hc = ldv<float, VEC>(hin + (long long)l0 * N + jc);
    Vec<float, VEC> cu = ldv<float, VEC>(cin + (long long)(l0 - 1) * N + jc);
    Vec<float, VEC> cc = ldv<float, VEC>(cin + (long long)l0 * N + jc);
    float heL = edgeL ? __ldg(hin + (long long)l0 * N + jl) : 0.f;
    float heR = edgeR ? __ldg(hin + (long long)l0 * N + jr) : 0.f;
    float ceL = edgeL ? __ldg(cin + (long long)l0 * N + jl) : 0.f;
    float ceR = edgeR ? __ldg(cin + (long long)l0 * N + jr) : 0.f;

    for (int l = l0; l <= l1; ++l) {
        const Vec<float, VEC> hd = ldv<float, VEC>(hin + (long long)(l + 1) * N + jc);
        const Vec<float, VEC> cd = ldv<float, VEC>(cin + (long long)(l + 1) * N + jc);
        float heLn = 0.f, heRn = 0.f, ceLn = 0.f, ceRn = 0.f;
        if (l < l1) {
            if (edgeL) { heLn = __ldg(hin + (long long)(l + 1) * N + jl); ceLn = __ldg(cin + (long long)(l + 1) * N + jl); }
            if (edgeR) { heRn = __ldg(hin + (long long)(l + 1) * N + jr); ceRn = __ldg(cin + (long long)(l + 1) * N + jr); }
        }
        const int grow = a.row_begin + (l - 1);
        const float4 rw = __ldg(reinterpret_cast<const float4 *>(a.rowW) + grow);
        const float sin_theta = rw.x, cot = rw.y, inv_sin2 = rw.z, inv_sin = rw.w;

        float hleft = __shfl_up_sync(0xffffffffu, hc.v[VEC - 1], 1);
        float hright = __shfl_down_sync(0xffffffffu, hc.v[0], 1);
        float cleft = __shfl_up_sync(0xffffffffu, cc.v[VEC - 1], 1);
        float cright = __shfl_down_sync(0xffffffffu, cc.v[0], 1);
        if (edgeL) { hleft = heL; cleft = ceL; }
        if (edgeR) { hright = heR; cright = ceR; }

        Vec<float, VEC> hn, cn;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float h = hc.v[k], conc = cc.v[k];
            const float h_tm = hu.v[k], h_tp = hd.v[k];
            const float h_pm = k == 0 ? hleft : hc.v[k - 1], h_pp = k == VEC - 1 ? hright : hc.v[k + 1];
            const float c_tm = cu.v[k], c_tp = cd.v[k];
            const float c_pm = k == 0 ? cleft : cc.v[k - 1], c_pp = k == VEC - 1 ? cright : cc.v[k + 1];

            const float drainage_term = -drainage_coeff * (h * h * h) * sin_theta;
            const float d2h_dtheta2 = (h_tp - 2.0f * h + h_tm) * inv_dth2;
            const float dh_dtheta = (h_tp - h_tm) * inv_2dth;
            const float d2h_dphi2 = (h_pp - 2.0f * h + h_pm) * inv_dph2;
            const float h_lap = (d2h_dtheta2 + cot * dh_dtheta + d2h_dphi2 * inv_sin2) * inv_r2;
            const float diffusion_term = p.diffusion_coeff * h_lap;
            float marangoni_term = 0.f;
            const float dc_dtheta = (c_tp - c_tm) * inv_2dth;
            if (mar) {
                const float dc_dphi = (c_pp - c_pm) * inv_2dph;
                const float gt = dc_dtheta * inv_r;
                const float gp = dc_dphi * inv_r * inv_sin;
                const float grad_c_mag = sqrtf(gt * gt + gp * gp);
                marangoni_term = p.marangoni_coeff * (p.gamma_reduction * grad_c_mag) * h * h;
            }
            const float new_h = fmaxf(h + p.dt * (drainage_term + diffusion_term + marangoni_term), 0.f);

            const float d2c_dtheta2 = (c_tp - 2.0f * conc + c_tm) * inv_dth2;
            const float d2c_dphi2 = (c_pp - 2.0f * conc + c_pm) * inv_dph2;
            const float c_lap = (d2c_dtheta2 + cot * dc_dtheta + d2c_dphi2 * inv_sin2) * inv_r2;
            float dc_dt = p.surfactant_diffusion * c_lap;
            if (mar) dc_dt += 0.001f * (h_lap * r2) * conc;
            const float new_c = fminf(fmaxf(conc + p.dt * dc_dt, 0.f), 1.f);
            const bool skip = h < p.critical_thickness;  // drainage.wgsl:108-112
            hn.v[k] = skip ? h : new_h;
            cn.v[k] = skip ? conc : new_c;
        }
        if (active) {
            stv<float, VEC>(hout + (long long)l * N + jc, hn);
            stv<float, VEC>(cout + (long long)l * N + jc, cn);
            if constexpr (MODE != SBX_RGB_NONE) {
                const float2 sc = __ldg(a.rowSC + grow);
                const float ccy = sc.y * a.rgb.cam_y;
                using DN = typename Dnm<MODE>::type;
                DN d_nm[VEC];
                float ci[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    d_nm[k] = (DN)hn.v[k] * (DN)1e9;
                    ci[k] = view_cos(sc.x, ccy, colA[k], a.radius_view, a.rgb.cam2);
                }
                emit_rgb<MODE, VEC>(d_nm, ci, a.rgb,
                                    (long long)b * a.rgb_bubble_stride + (long long)(l - 1) * N + jc);
            }
        }
        hu = hc; hc = hd; cu = cc; cc = cd;
        heL = heLn; heR = heRn; ceL = ceLn; ceR = ceRn;
    }
}

// ------------------------------------------------------------------ small utility kernels
// host dtype <-> storage type; F32 storage keeps (h - h_ref[bubble]), the subtraction is done in f64
template <typename TD, typename TS>
__global__ void k_convert_in(TD *dst, const TS *src, long long n, long long bubble_stride, const double *href)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (TD)((double)src[i] - (href ? href[i / bubble_stride] : 0.0));
}

template <typename TD, typename TS>
__global__ void k_convert_out(TD *dst, const TS *src, long long n, long long bubble_stride, const double *href)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (TD)((double)src[i] + (href ? href[i / bubble_stride] : 0.0));
}

template <typename T>
__global__ void k_fill(T *dst, long long n, T v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = v;
}

// per-block min/max partials of the owned rows of every bubble (ThicknessField::min/max_thickness,
// drainage.rs:180-193), one launch for the whole batch.  Work items are (bubble, chunk of kMmChunk packets); the
// reduction runs in the storage type (exact, NaN ignored like f64::min/max) and the per-bubble storage offset is added
// once per item -- (double)v + h_ref is monotone in v, so min/max commute with it.
constexpr int kMmChunk = 4096;  // 16-byte packets per work item
template <typename T, int VEC>
__global__ void __launch_bounds__(256) k_minmax(const T *buf, int batch, long long bubble_stride, long long first,
                                                long long n, const double *href, double *part)
{
    const long long nv = n / VEC;  // packets per bubble (VEC == 1: elements)
    const long long chunks = (nv + kMmChunk - 1) / kMmChunk;
    const long long items = chunks * batch;
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int b = (int)(it / chunks);
        const long long c0 = (it - (long long)b * chunks) * kMmChunk;
        const long long c1 = c0 + kMmChunk < nv ? c0 + kMmChunk : nv;
        const T *src = buf + (long long)b * bubble_stride + first;
        T lo = src[c0 * VEC], hi = lo;  // any element of the item: a neutral start for threads that load nothing
        for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            const Vec<T, VEC> q = ldv<T, VEC>(src + i * VEC);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if constexpr (sizeof(T) == 4) { lo = fminf(lo, q.v[k]); hi = fmaxf(hi, q.v[k]); }
                else { lo = fmin(lo, q.v[k]); hi = fmax(hi, q.v[k]); }
            }
        }
        const double off = href ? href[b] : 0.0;
        mn = fmin(mn, (double)lo + off);
        mx = fmax(mx, (double)hi + off);
    }
    __shared__ double smn[8], smx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { mn = fmin(mn, smn[i]); mx = fmax(mx, smx[i]); }
        part[2 * blockIdx.x] = mn;
        part[2 * blockIdx.x + 1] = mx;
    }
}

// DrainageSimulator::get_thickness (drainage.rs:543-551): bilinear blend of four cells in f64 with
// the reference's operation order; indices and fractions come from the host (index arithmetic only).
template <typename T>
__global__ void k_sample(const T *buf, long long i00, long long i01, long long i10, long long i11,
                         double tf, double pf, double h_ref, double *out)
{
    const double v00 = (double)buf[i00] + h_ref, v01 = (double)buf[i01] + h_ref;
    const double v10 = (double)buf[i10] + h_ref, v11 = (double)buf[i11] + h_ref;
    const double omp = __dsub_rn(1.0, pf), omt = __dsub_rn(1.0, tf);
    const double v0 = __dadd_rn(__dmul_rn(v00, omp), __dmul_rn(v01, pf));
    const double v1 = __dadd_rn(__dmul_rn(v10, omp), __dmul_rn(v11, pf));
    *out = __dadd_rn(__dmul_rn(v0, omt), __dmul_rn(v1, tf));
}

// ------------------------------------------------------------------ interference LUT
// generate_interference_lut / thin_film_interference (src/render/interference_lut.rs:18-163) with the
// wavelength list as a parameter; one thread per texel.
__device__ __forceinline__ float gaussianf_(float x, float mean, float sigma)
{
    const float t = (x - mean) / sigma;
    return expf(-0.5f * t * t);
}

__global__ void k_lut_build(uint8_t *out, int T, int A, float max_nm, float n_film, float intensity,
                            const float *wl, int n_lambda)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ai = blockIdx.y;
    if (t >= T) return;
    const float PI = 3.14159265358979323846f;
    const float d_nm = __fmul_rn(__fdiv_rn((float)t, (float)(T - 1)), max_nm);
    const float cos_theta = __fdiv_rn((float)ai, (float)(A - 1));
    // snells_law, :96-100
    const float sin_i = __fsqrt_rn(fmaxf(__fsub_rn(1.0f, __fmul_rn(cos_theta, cos_theta)), 0.0f));
    const float sin_t = __fdiv_rn(sin_i, n_film);
    const float cos_t = __fsqrt_rn(fmaxf(__fsub_rn(1.0f, __fmul_rn(sin_t, sin_t)), 0.0f));
    // fresnel_unpolarized(cos_i, cos_t, 1.0, n), :103-120
    const float n1ci = cos_theta, n2ct = __fmul_rn(n_film, cos_t);
    const float r_s = __fdiv_rn(__fsub_rn(n1ci, n2ct), fmaxf(__fadd_rn(n1ci, n2ct), 0.0001f));
    const float n2ci = __fmul_rn(n_film, cos_theta), n1ct = cos_t;
    const float r_p = __fdiv_rn(__fsub_rn(n2ci, n1ct), fmaxf(__fadd_rn(n2ci, n1ct), 0.0001f));
    const float fres = __fmul_rn(__fadd_rn(__fmul_rn(r_s, r_s), __fmul_rn(r_p, r_p)), 0.5f);
    const float optical_path = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, n_film), d_nm), cos_t);
    // airy constants, :123-126
    const float omr = fmaxf(__fsub_rn(1.0f, fres), 0.001f);
    const float F = __fdiv_rn(__fmul_rn(4.0f, fres), __fmul_rn(omr, omr));
    float X = 0.f, Y = 0.f, Z = 0.f;
    for (int k = 0; k < n_lambda; ++k) {
        const float w = wl[k];
        const float phase = __fadd_rn(__fdiv_rn(__fmul_rn(__fmul_rn(2.0f, PI), optical_path), w), PI);
        const float sh = sinf(__fmul_rn(phase, 0.5f));
        const float sin2 = __fmul_rn(sh, sh);
        const float num = __fmul_rn(F, sin2);
        const float den = __fadd_rn(1.0f, num);
        const float airy = __fdiv_rn(num, fmaxf(den, 0.001f));
        // cie_color_matching, :136-148
        const float cx = fmaxf(__fsub_rn(__fadd_rn(__fmul_rn(1.056f, gaussianf_(w, 599.8f, 37.9f)),
                                                   __fmul_rn(0.362f, gaussianf_(w, 442.0f, 16.0f))),
                                         __fmul_rn(0.065f, gaussianf_(w, 501.1f, 20.4f))), 0.f);
        const float cy = fmaxf(__fadd_rn(__fmul_rn(0.821f, gaussianf_(w, 568.8f, 46.9f)),
                                         __fmul_rn(0.286f, gaussianf_(w, 530.9f, 31.1f))), 0.f);
        const float cz = fmaxf(__fadd_rn(__fmul_rn(1.217f, gaussianf_(w, 437.0f, 11.8f)),
                                         __fmul_rn(0.681f, gaussianf_(w, 459.0f, 26.0f))), 0.f);
        X = __fadd_rn(X, __fmul_rn(cx, airy));
        Y = __fadd_rn(Y, __fmul_rn(cy, airy));
        Z = __fadd_rn(Z, __fmul_rn(cz, airy));
    }
    const float nl = (float)n_lambda;
    X = __fdiv_rn(X, nl); Y = __fdiv_rn(Y, nl); Z = __fdiv_rn(Z, nl);
    // xyz_to_rgb, :158-163
    float r = __fsub_rn(__fsub_rn(__fmul_rn(3.2404542f, X), __fmul_rn(1.5371385f, Y)), __fmul_rn(0.4985314f, Z));
    float g = __fadd_rn(__fadd_rn(__fmul_rn(-0.969266f, X), __fmul_rn(1.8760108f, Y)), __fmul_rn(0.0415560f, Z));
    float bl = __fadd_rn(__fsub_rn(__fmul_rn(0.0556434f, X), __fmul_rn(0.2040259f, Y)), __fmul_rn(1.0572252f, Z));
    r = __fmul_rn(r, intensity); g = __fmul_rn(g, intensity); bl = __fmul_rn(bl, intensity);
    // (x.clamp(0,1) * 255.0) as u8 : truncation, :33-35
    uint8_t *o = out + 4 * ((size_t)ai * (size_t)T + (size_t)t);
    o[0] = (uint8_t)__fmul_rn(clamp01(r), 255.0f);
    o[1] = (uint8_t)__fmul_rn(clamp01(g), 255.0f);
    o[2] = (uint8_t)__fmul_rn(clamp01(bl), 255.0f);
    o[3] = 255;
}

// ------------------------------------------------------------------ InterferenceCalculator in f64
struct InterfD {
    double n_film, n_medium, wl[3];
};

__device__ __forceinline__ double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

__device__ __forceinline__ double d_transmission_cos(const InterfD &c, double cos_i)
{  // interference.rs:197-211
    const double ci = clampd(cos_i, -1.0, 1.0);
    const double sin_i = sqrt(__dsub_rn(1.0, __dmul_rn(ci, ci)));
    const double sin_t = __dmul_rn(__ddiv_rn(c.n_medium, c.n_film), sin_i);
    if (fabs(sin_t) > 1.0) return 0.0;
    return sqrt(__dsub_rn(1.0, __dmul_rn(sin_t, sin_t)));
}

__device__ __forceinline__ void d_fresnel(const InterfD &c, double cos_i, double out[3])
{  // interference.rs:266-298
    const double ci = clampd(cos_i, -1.0, 1.0);
    const double ct = d_transmission_cos(c, ci);
    const double n1ci = __dmul_rn(c.n_medium, ci), n2ct = __dmul_rn(c.n_film, ct);
    const double n2ci = __dmul_rn(c.n_film, ci), n1ct = __dmul_rn(c.n_medium, ct);
    const double rs = __ddiv_rn(__dsub_rn(n1ci, n2ct), __dadd_rn(n1ci, n2ct));
    const double rp = __ddiv_rn(__dsub_rn(n2ci, n1ct), __dadd_rn(n2ci, n1ct));
    out[0] = __dmul_rn(rs, rs);
    out[1] = __dmul_rn(rp, rp);
    out[2] = __ddiv_rn(__dadd_rn(out[0], out[1]), 2.0);
}

__device__ __forceinline__ double d_reflected_intensity(const InterfD &c, double d_nm, double cos_i, double wl)
{  // interference.rs:312-353
    if (d_nm <= 0.0) return 0.0;
    const double ct = d_transmission_cos(c, cos_i);
    const double geometric = __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.n_film), d_nm), ct);
    const double opd = __dadd_rn(geometric, __ddiv_rn(wl, 2.0));
    const double phase = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, 3.14159265358979323846), opd), wl);
    double fr[3];
    d_fresnel(c, cos_i, fr);
    const double R = fr[2];
    const double cos_phase = cos(phase);
    const double numerator = __dmul_rn(__dmul_rn(2.0, R), __dsub_rn(1.0, cos_phase));
    const double denominator =
        __dsub_rn(__dadd_rn(1.0, __dmul_rn(R, R)), __dmul_rn(__dmul_rn(2.0, R), cos_phase));
    if (fabs(denominator) < 2.220446049250313e-16) return R;
    return clampd(__ddiv_rn(numerator, denominator), 0.0, 1.0);
}

__global__ void k_interference_eval(InterfD c, const double *d_nm, const double *cos_i, size_t n, int gamma,
                                    double *rgb)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < 3; ++k) {
        double I = d_reflected_intensity(c, d_nm[i], cos_i[i], c.wl[k]);
        if (gamma) I = pow(I, 1.0 / 2.2);  // interference.rs:394-397
        rgb[3 * i + k] = clampd(I, 0.0, 1.0);
    }
}

__global__ void k_interference_terms(InterfD c, const double *d_nm, const double *cos_i, size_t n, double wl,
                                     double *out_cos_t, double *out_fresnel, double *out_intensity)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (out_cos_t) out_cos_t[i] = d_transmission_cos(c, cos_i[i]);
    if (out_fresnel) d_fresnel(c, cos_i[i], out_fresnel + 3 * i);
    if (out_intensity) out_intensity[i] = d_reflected_intensity(c, d_nm[i], cos_i[i], wl);
}

// bubble.wgsl:425-439 for arbitrary (thickness, cos) pairs: same texture-unit fetch as the fused colour stage
__global__ void k_lut_shade_points(cudaTextureObject_t lut, float inv_max_nm, float intensity, const float *d_nm,
                                   const float *cos_theta, size_t n, float *rgb)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 t = tex2D<float4>(lut, clamp01(d_nm[i] * inv_max_nm), clamp01(cos_theta[i]));
    rgb[3 * i + 0] = clamp01(t.x * intensity);
    rgb[3 * i + 1] = clamp01(t.y * intensity);
    rgb[3 * i + 2] = clamp01(t.z * intensity);
}

cudaError_t launch_lut_shade_points(cudaTextureObject_t lut, float inv_max_nm, float intensity, const float *d_nm,
                                    const float *cos_theta, size_t n, float *rgb, cudaStream_t st)
{
    k_lut_shade_points<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lut, inv_max_nm, intensity, d_nm, cos_theta, n, rgb);
    return cudaGetLastError();
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) k_caustics(const T *buf, long long bubble_stride, float *out, int num_phi,
                                                  int num_theta, int row_begin, int rows_local, const double *href,
                                                  sbx_caustic_params p)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int l = blockIdx.y * blockDim.y + threadIdx.y + 1;  // local row, 1..rows_local
    const int b = blockIdx.z;
    if (x0 >= num_phi || l > rows_local) return;
    const T *in = buf + (long long)b * bubble_stride;
    const float hr = href ? (float)href[b] : 0.f;
    const int g = row_begin + l - 1;
    const int lu = g == 0 ? l : l - 1, ld = g == num_theta - 1 ? l : l + 1;  // clamp(y, 0, height-1)
    const int xl = x0 == 0 ? num_phi - 1 : x0 - 1, xr = x0 + VEC >= num_phi ? 0 : x0 + VEC;
    float r[3][VEC + 2];  // rows up / centre / down: [left edge, VEC cells, right edge]
    const int rows[3] = {lu, l, ld};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T *row = in + (long long)rows[k] * num_phi;
        const Vec<T, VEC> v = ldv<T, VEC>(row + x0);
        r[k][0] = (float)__ldg(row + xl) + hr;
#pragma unroll
        for (int j = 0; j < VEC; ++j) r[k][1 + j] = (float)v.v[j] + hr;
        r[k][VEC + 1] = (float)__ldg(row + xr) + hr;
    }
    float o[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        o[j] = caustic_intensity(r[1][1 + j], r[1][j], r[1][2 + j], r[0][1 + j], r[2][1 + j], r[0][j], r[0][2 + j],
                                 r[2][j], r[2][2 + j], p);
    float *dst = out + ((long long)b * rows_local + (l - 1)) * num_phi + x0;
    if constexpr (VEC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2 *>(dst) = make_float2(o[0], o[1]);
    else dst[0] = o[0];
}

template <typename T>
cudaError_t launch_caustics(const T *buf, long long bubble_stride, float *out, int num_phi, int num_theta, int row_begin,
                            int rows_local, int batch, const double *href, const sbx_caustic_params &p, cudaStream_t st,
                            bool use_tma)
{
    constexpr int VECW = 16 / (int)sizeof(T);
    const bool vec_ok = num_phi % VECW == 0;
    if (use_tma && vec_ok && num_phi >= 32 * VECW && rows_local >= 8)
        return launch_caustics_tma<T>(buf, bubble_stride, out, num_phi, num_theta, row_begin, rows_local, batch, href, p, st);
    const int packets = vec_ok ? num_phi / VECW : num_phi;
    const dim3 block(64, 4), grid((unsigned)((packets + 63) / 64), (unsigned)((rows_local + 3) / 4), (unsigned)batch);
    if (vec_ok) k_caustics<T, VECW><<<grid, block, 0, st>>>(buf, bubble_stride, out, num_phi, num_theta, row_begin, rows_local, href, p);
    else k_caustics<T, 1><<<grid, block, 0, st>>>(buf, bubble_stride, out, num_phi, num_theta, row_begin, rows_local, href, p);
    return cudaGetLastError();
}
template cudaError_t launch_caustics<float>(const float *, long long, float *, int, int, int, int, int, const double *, const sbx_caustic_params &, cudaStream_t, bool);
template cudaError_t launch_caustics<double>(const double *, long long, float *, int, int, int, int, int, const double *, const sbx_caustic_params &, cudaStream_t, bool);

// branched_flow_compute.wgsl:187-210 for n query points
template <typename T>
__global__ void k_sample_uv(const T *field /* first owned row of the bubble */, int W, int H, float href, const float *uv,
                            size_t n, float *out_h, float *out_grad)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto S = [&](float u, float v) {
        const float cu = fminf(fmaxf(u, 0.f), 1.f), cv = fminf(fmaxf(v, 0.f), 1.f);
        const unsigned x = (unsigned)__fmul_rn(cu, (float)(W - 1)), y = (unsigned)__fmul_rn(cv, (float)(H - 1));
        return (float)field[(size_t)y * W + x] + href;
    };
    const float eps = 0.01f;
    const float u = uv[2 * i], v = uv[2 * i + 1];
    if (out_h) out_h[i] = S(u, v);
    if (out_grad) {
        out_grad[2 * i] = __fdiv_rn(__fsub_rn(S(__fadd_rn(u, eps), v), S(__fsub_rn(u, eps), v)), __fmul_rn(2.0f, eps));
        out_grad[2 * i + 1] = __fdiv_rn(__fsub_rn(S(u, __fadd_rn(v, eps)), S(u, __fsub_rn(v, eps))), __fmul_rn(2.0f, eps));
    }
}

template <typename T>
__global__ void k_sample_normal(const T *field, int W, int H, float href, const float *nrm, size_t n, float scale, float *out_h)
{ // branched_flow_compute.wgsl:110-116, 294-300
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float PI_F = 3.14159265359f;
    const float phi = atan2f(nrm[3 * i + 2], nrm[3 * i]);
    const float theta = acosf(fminf(fmaxf(nrm[3 * i + 1], -1.f), 1.f));
    const float u = __fdiv_rn(__fadd_rn(phi, PI_F), __fmul_rn(2.0f, PI_F)), v = __fdiv_rn(theta, PI_F);
    const unsigned x = min((unsigned)__fmul_rn(u, (float)(W - 1)), (unsigned)(W - 1));
    const unsigned y = min((unsigned)__fmul_rn(v, (float)(H - 1)), (unsigned)(H - 1));
    out_h[i] = __fmul_rn((float)field[(size_t)y * W + x] + href, scale);
}

template <typename T>
cudaError_t launch_sample_normal(const T *field, int W, int H, float href, const float *nrm, size_t n, float scale,
                                 float *out_h, cudaStream_t st)
{
    k_sample_normal<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(field, W, H, href, nrm, n, scale, out_h);
    return cudaGetLastError();
}
template cudaError_t launch_sample_normal<float>(const float *, int, int, float, const float *, size_t, float, float *, cudaStream_t);
template cudaError_t launch_sample_normal<double>(const double *, int, int, float, const float *, size_t, float, float *, cudaStream_t);

template <typename T>
cudaError_t launch_sample_uv(const T *field, int W, int H, float href, const float *uv, size_t n, float *out_h,
                             float *out_grad, cudaStream_t st)
{
    k_sample_uv<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(field, W, H, href, uv, n, out_h, out_grad);
    return cudaGetLastError();
}
template cudaError_t launch_sample_uv<float>(const float *, int, int, float, const float *, size_t, float *, float *, cudaStream_t);
template cudaError_t launch_sample_uv<double>(const double *, int, int, float, const float *, size_t, float *, float *, cudaStream_t);

// per-column view term A_j = cos(phi_j) cam_x + sin(phi_j) cam_z, evaluated in f64
__global__ void k_col_view(float *colA, int num_phi, double cam_x, double cam_z)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_phi) return;
    const double phi = (double)j * (2.0 * 3.14159265358979323846 / (double)num_phi);
    double s, c;
    sincos(phi, &s, &c);
    colA[j] = (float)(c * cam_x + s * cam_z);
}

// ==================================================================== host-side launchers
static inline dim3 grid1d(long long n, int threads, int max_blocks = 148 * 16)
{
    long long b = (n + threads - 1) / threads;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return dim3((unsigned)b);
}

template <typename T, int VEC, bool TFE, bool STRICT>
static cudaError_t launch_march_mode(const StepArgs<T> &a, dim3 grid, dim3 block, cudaStream_t st)
{
    switch (a.rgb.mode) {
    case SBX_RGB_NONE: k_march<T, VEC, SBX_RGB_NONE, TFE, STRICT><<<grid, block, 0, st>>>(a); break;
    case SBX_RGB_LUT: k_march<T, VEC, SBX_RGB_LUT, TFE, STRICT><<<grid, block, 0, st>>>(a); break;
    case SBX_RGB_EXACT3: k_march<T, VEC, SBX_RGB_EXACT3, TFE, STRICT><<<grid, block, 0, st>>>(a); break;
    case SBX_RGB_EXACT3_LINEAR: k_march<T, VEC, SBX_RGB_EXACT3_LINEAR, TFE, STRICT><<<grid, block, 0, st>>>(a); break;
    case SBX_RGB_SCHLICK_WALL: k_march<T, VEC, SBX_RGB_SCHLICK_WALL, TFE, STRICT><<<grid, block, 0, st>>>(a); break;
    case SBX_RGB_SCHLICK_TWO_BEAM: k_march<T, VEC, SBX_RGB_SCHLICK_TWO_BEAM, TFE, STRICT><<<grid, block, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <typename T, int VEC>
static cudaError_t launch_march_vec(const StepArgs<T> &a, bool tfe, bool strict, dim3 grid, dim3 block,
                                    cudaStream_t st)
{
    if constexpr (sizeof(T) == 8) {
        if (strict) {
            // strict mode applies the TFE reduction in a separate pass after the poles
            return launch_march_mode<T, VEC, false, true>(a, grid, block, st);
        }
    }
    if (tfe) return launch_march_mode<T, VEC, true, false>(a, grid, block, st);
    return launch_march_mode<T, VEC, false, false>(a, grid, block, st);
}

static void march_geometry(int num_phi, int vec, int rows, int batch, int band_rows_req, dim3 &grid,
                           dim3 &block, int &band_rows)
{
    const int packets = (num_phi + vec - 1) / vec;
    int tx = ((packets + 31) / 32) * 32;
    if (tx > 256) tx = 256;
    const int gx = (packets + tx - 1) / tx;
    band_rows = band_rows_req > 0 ? band_rows_req : 32;
    if (band_rows > rows) band_rows = rows > 0 ? rows : 1;
    const int gy = rows > 0 ? (rows + band_rows - 1) / band_rows : 1;
    grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)batch);
    block = dim3((unsigned)tx);
}

template <typename T>
cudaError_t launch_step(StepArgs<T> a, int batch, bool tfe, bool strict, bool has_top, bool has_bottom,
                        int band_rows_req, bool use_tma, cudaStream_t st, int *launches)
{
    constexpr int VECW = 16 / (int)sizeof(T);
    const bool vec_ok = (a.num_phi % VECW) == 0;
    const int vec = vec_ok ? VECW : 1;
    const int rows = a.l_last - a.l_first + 1;
    dim3 grid, block;
    int band_rows;
    march_geometry(a.num_phi, vec, rows, batch, band_rows_req, grid, block, band_rows);
    a.band_rows = band_rows;
    a.pole_slots = (int)grid.x;
    cudaError_t e = cudaSuccess;
    if (rows > 0) {
        bool done = false;
        {
            if (!strict && use_tma && march_tma_supported<T>(a)) {
                int slots = 0;
                // a.fold_top / a.fold_bottom / a.pole_want were set by the caller iff march_tma_folds_poles(a)
                e = launch_march_tma<T>(a, batch, tfe, band_rows_req, st, &slots);
                a.pole_slots = slots;
                done = true;
                if (a.fold_top) has_top = false;  // the strips of the ring rows wrote the pole rows themselves
                if (a.fold_bottom) has_bottom = false;
            }
        }
        if (!done)
            e = vec_ok ? launch_march_vec<T, VECW>(a, tfe, strict, grid, block, st)
                       : launch_march_vec<T, 1>(a, tfe, strict, grid, block, st);
        if (e != cudaSuccess) return e;
        ++*launches;
    }
    if (has_top || has_bottom) {
        const int bpp = (a.num_phi + 1023) / 1024;
        const dim3 pg((unsigned)(2 * bpp), (unsigned)batch);
#define SBX_POLES(MODE)                                                                         \
    if (strict) { if constexpr (sizeof(T) == 8) k_poles<T, MODE, true><<<pg, 256, 0, st>>>(a, has_top, has_bottom, bpp); } \
    else k_poles<T, MODE, false><<<pg, 256, 0, st>>>(a, has_top, has_bottom, bpp)
        switch (a.rgb.mode) {
        case SBX_RGB_NONE: SBX_POLES(SBX_RGB_NONE); break;
        case SBX_RGB_LUT: SBX_POLES(SBX_RGB_LUT); break;
        case SBX_RGB_EXACT3: SBX_POLES(SBX_RGB_EXACT3); break;
        case SBX_RGB_EXACT3_LINEAR: SBX_POLES(SBX_RGB_EXACT3_LINEAR); break;
        case SBX_RGB_SCHLICK_WALL: SBX_POLES(SBX_RGB_SCHLICK_WALL); break;
        case SBX_RGB_SCHLICK_TWO_BEAM: SBX_POLES(SBX_RGB_SCHLICK_TWO_BEAM); break;
        default: return cudaErrorInvalidValue;
        }
#undef SBX_POLES
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        ++*launches;
    }
    return cudaSuccess;
}

template cudaError_t launch_step<float>(StepArgs<float>, int, bool, bool, bool, bool, int, bool, cudaStream_t, int *);
template cudaError_t launch_step<double>(StepArgs<double>, int, bool, bool, bool, bool, int, bool, cudaStream_t, int *);

template <typename T>
cudaError_t launch_tfe_pass(T *buf, long long bubble_stride, int num_phi, int row_begin, int rows_local,
                            int batch, const FrontTable &ft, double factor, cudaStream_t st, int *launches)
{
    if (ft.count == 0 || ft.row_max < ft.row_min) return cudaSuccess;
    const dim3 grid((unsigned)((num_phi + 127) / 128), (unsigned)(ft.row_max - ft.row_min + 1), (unsigned)batch);
    k_tfe<T><<<grid, 128, 0, st>>>(buf, bubble_stride, num_phi, row_begin, rows_local, ft, factor);
    ++*launches;
    return cudaGetLastError();
}
template cudaError_t launch_tfe_pass<float>(float *, long long, int, int, int, int, const FrontTable &, double, cudaStream_t, int *);
template cudaError_t launch_tfe_pass<double>(double *, long long, int, int, int, int, const FrontTable &, double, cudaStream_t, int *);

template <typename T>
cudaError_t launch_shade(StepArgs<T> a, int batch, int band_rows_req, cudaStream_t st, int *launches, bool use_tma)
{
    constexpr int VECW = 16 / (int)sizeof(T);
    const bool vec_ok = (a.num_phi % VECW) == 0;
    const int rows = a.l_last - a.l_first + 1;
    if (rows <= 0) return cudaSuccess;
    if (use_tma && rows >= 8 && shade_tma_supported<T>(a)) {
        ++*launches;
        return launch_shade_tma<T>(a, batch, band_rows_req, st);
    }
    dim3 grid, block;
    int band_rows;
    march_geometry(a.num_phi, vec_ok ? VECW : 1, rows, batch, band_rows_req, grid, block, band_rows);
    a.band_rows = band_rows;
#define SBX_SHADE(MODE)                                                       \
    if constexpr (sizeof(T) == 4) {                                           \
        if (vec_ok && a.rgb.format == SBX_OUT_RGBA8) { k_shade_rgba8<MODE><<<grid, block, 0, st>>>(a); break; } \
    }                                                                         \
    if (vec_ok) k_shade<T, VECW, MODE><<<grid, block, 0, st>>>(a);            \
    else k_shade<T, 1, MODE><<<grid, block, 0, st>>>(a)
    switch (a.rgb.mode) {
    case SBX_RGB_LUT: SBX_SHADE(SBX_RGB_LUT); break;
    case SBX_RGB_EXACT3: SBX_SHADE(SBX_RGB_EXACT3); break;
    case SBX_RGB_EXACT3_LINEAR: SBX_SHADE(SBX_RGB_EXACT3_LINEAR); break;
    case SBX_RGB_SCHLICK_WALL: SBX_SHADE(SBX_RGB_SCHLICK_WALL); break;
    case SBX_RGB_SCHLICK_TWO_BEAM: SBX_SHADE(SBX_RGB_SCHLICK_TWO_BEAM); break;
    case SBX_RGB_SPECTRAL: SBX_SHADE(SBX_RGB_SPECTRAL); break;
    default: return cudaErrorInvalidValue;
    }
#undef SBX_SHADE
    ++*launches;
    return cudaGetLastError();
}
template cudaError_t launch_shade<float>(StepArgs<float>, int, int, cudaStream_t, int *, bool);
template cudaError_t launch_shade<double>(StepArgs<double>, int, int, cudaStream_t, int *, bool);

cudaError_t launch_wgsl(WgslArgs a, int batch, int band_rows_req, bool use_tma, cudaStream_t st, int *launches)
{
    const bool vec_ok = (a.num_phi % 4) == 0;
    const int rows = a.l_last - a.l_first + 1;
    if (rows <= 0) return cudaSuccess;
    if (use_tma && wgsl_tma_supported(a)) {
        ++*launches;
        return launch_wgsl_tma(a, batch, band_rows_req, st);
    }
    dim3 grid, block;
    int band_rows;
    march_geometry(a.num_phi, vec_ok ? 4 : 1, rows, batch, band_rows_req, grid, block, band_rows);
    a.band_rows = band_rows;
#define SBX_WGSL(MODE)                                                  \
    if (vec_ok) k_wgsl<4, MODE><<<grid, block, 0, st>>>(a);             \
    else k_wgsl<1, MODE><<<grid, block, 0, st>>>(a)
    switch (a.rgb.mode) {
    case SBX_RGB_NONE: SBX_WGSL(SBX_RGB_NONE); break;
    case SBX_RGB_LUT: SBX_WGSL(SBX_RGB_LUT); break;
    case SBX_RGB_EXACT3: SBX_WGSL(SBX_RGB_EXACT3); break;
    case SBX_RGB_EXACT3_LINEAR: SBX_WGSL(SBX_RGB_EXACT3_LINEAR); break;
    case SBX_RGB_SCHLICK_WALL: SBX_WGSL(SBX_RGB_SCHLICK_WALL); break;
    case SBX_RGB_SCHLICK_TWO_BEAM: SBX_WGSL(SBX_RGB_SCHLICK_TWO_BEAM); break;
    default: return cudaErrorInvalidValue;
    }
#undef SBX_WGSL
    ++*launches;
    return cudaGetLastError();
}

template <typename TD, typename TS>
cudaError_t launch_convert_in(TD *dst, const TS *src, long long n, long long bubble_stride, const double *href, cudaStream_t st)
{
    k_convert_in<TD, TS><<<grid1d(n, 256), 256, 0, st>>>(dst, src, n, bubble_stride, href);
    return cudaGetLastError();
}
template <typename TD, typename TS>
cudaError_t launch_convert_out(TD *dst, const TS *src, long long n, long long bubble_stride, const double *href, cudaStream_t st)
{
    k_convert_out<TD, TS><<<grid1d(n, 256), 256, 0, st>>>(dst, src, n, bubble_stride, href);
    return cudaGetLastError();
}
#define SBX_INST_CONV(TD, TS)                                                                    \
    template cudaError_t launch_convert_in<TD, TS>(TD *, const TS *, long long, long long, const double *, cudaStream_t); \
    template cudaError_t launch_convert_out<TD, TS>(TD *, const TS *, long long, long long, const double *, cudaStream_t);
SBX_INST_CONV(float, float)
SBX_INST_CONV(float, double)
SBX_INST_CONV(double, float)
SBX_INST_CONV(double, double)
#undef SBX_INST_CONV

template <typename T>
cudaError_t launch_fill(T *dst, long long n, T v, cudaStream_t st)
{
    k_fill<T><<<grid1d(n, 256), 256, 0, st>>>(dst, n, v);
    return cudaGetLastError();
}
template cudaError_t launch_fill<float>(float *, long long, float, cudaStream_t);
template cudaError_t launch_fill<double>(double *, long long, double, cudaStream_t);

template <typename T>
cudaError_t launch_minmax(const T *buf, int batch, long long bubble_stride, long long first, long long n,
                          const double *href, double *part, int nblocks, cudaStream_t st)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const bool vec = n % VEC == 0 && first % VEC == 0 && bubble_stride % VEC == 0;
    if (vec) k_minmax<T, VEC><<<nblocks, 256, 0, st>>>(buf, batch, bubble_stride, first, n, href, part);
    else k_minmax<T, 1><<<nblocks, 256, 0, st>>>(buf, batch, bubble_stride, first, n, href, part);
    return cudaGetLastError();
}
template cudaError_t launch_minmax<float>(const float *, int, long long, long long, long long, const double *, double *, int, cudaStream_t);
template cudaError_t launch_minmax<double>(const double *, int, long long, long long, long long, const double *, double *, int, cudaStream_t);

template <typename T>
cudaError_t launch_sample(const T *buf, long long i00, long long i01, long long i10, long long i11, double tf,
                          double pf, double h_ref, double *out, cudaStream_t st)
{
    k_sample<T><<<1, 1, 0, st>>>(buf, i00, i01, i10, i11, tf, pf, h_ref, out);
    return cudaGetLastError();
}
template cudaError_t launch_sample<float>(const float *, long long, long long, long long, long long, double, double, double, double *, cudaStream_t);
template cudaError_t launch_sample<double>(const double *, long long, long long, long long, long long, double, double, double, double *, cudaStream_t);

cudaError_t launch_lut_build(uint8_t *out, int T, int A, float max_nm, float n_film, float intensity,
                             const float *wl_dev, int n_lambda, cudaStream_t st)
{
    const dim3 grid((unsigned)((T + 127) / 128), (unsigned)A);
    k_lut_build<<<grid, 128, 0, st>>>(out, T, A, max_nm, n_film, intensity, wl_dev, n_lambda);
    return cudaGetLastError();
}

cudaError_t launch_interference_eval(double n_film, const double wl[3], const double *d_nm, const double *cos_i,
                                     size_t n, int gamma, double *rgb, cudaStream_t st)
{
    InterfD c{n_film, 1.0, {wl[0], wl[1], wl[2]}};
    k_interference_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(c, d_nm, cos_i, n, gamma, rgb);
    return cudaGetLastError();
}

cudaError_t launch_interference_terms(double n_film, const double *d_nm, const double *cos_i, size_t n, double wl,
                                      double *out_cos_t, double *out_fresnel, double *out_intensity, cudaStream_t st)
{
    InterfD c{n_film, 1.0, {650.0, 532.0, 450.0}};
    k_interference_terms<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(c, d_nm, cos_i, n, wl, out_cos_t, out_fresnel, out_intensity);
    return cudaGetLastError();
}

cudaError_t launch_col_view(float *colA, int num_phi, double cam_x, double cam_z, cudaStream_t st)
{
    k_col_view<<<(num_phi + 255) / 256, 256, 0, st>>>(colA, num_phi, cam_x, cam_z);
    return cudaGetLastError();
}

}  // namespace sbx
