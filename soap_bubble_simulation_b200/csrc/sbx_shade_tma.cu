// sbx_shade_tma.cu -- colour of the current thickness field (sbx_drainage_shade; the LUT-shading half of the
// spectral sweep, BASELINE config 5) on the TMA ring skeleton of sbx_march_tma.cu.
//
// The direct-load shade kernel (k_shade_rgba8) is a load -> TEX -> store chain with nothing to overlap inside a
// thread and reaches 64 % of the HBM roofline at 8 B/cell.  Here a producer warp streams the strip's rows into a
// shared-memory ring with cp.async.bulk (rows need no halo: the colour stage has no stencil), so the consumers'
// instruction stream has no global load at all and many rows are in flight per SM at any time.
#include "sbx_device.cuh"
#include "sbx_launch.h"
#include "sbx_tma.cuh"

namespace sbx {

namespace {
constexpr int kSStages = 4;        // ring depth
constexpr int kSRowsPerStage = 2;  // theta rows per stage
constexpr int kSMaxBandRows = 64;  // rows a CTA shades at most
}  // namespace

template <typename T, int MODE>
__global__ void __launch_bounds__(288, 5) k_shade_tma(const StepArgs<T> a)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ float2 s_sc[kSMaxBandRows];  // {sin, cos} of the band's rows, staged once
    const int C = blockDim.x - 32;  // consumer threads
    const int N = a.num_phi;
    const int strip = VEC * C;
    const int c0 = blockIdx.x * strip;
    const int w = min(strip, N - c0);  // multiple of VEC
    T *ring = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kSStages * kSRowsPerStage * strip * sizeof(T));
    uint64_t *empty = full + kSStages;

    const int b = blockIdx.z;
    const T *in = a.in + (long long)b * a.bubble_stride;
    const int l0 = a.l_first + blockIdx.y * a.band_rows;
    const int l1 = min(l0 + a.band_rows - 1, a.l_last);
    if (l0 > l1) return;
    const int rows = l1 - l0 + 1;
    const int nchunks = (rows + kSRowsPerStage - 1) / kSRowsPerStage;
    const int nwarps_c = C >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarps_c);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= C) {
        if (threadIdx.x == C) {  // producer: one lane issues the bulk copies
            const uint32_t row_bytes = (uint32_t)w * (uint32_t)sizeof(T);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kSStages;
                if (c >= kSStages) mbar_wait(&empty[s], (uint32_t)(((c / kSStages) - 1) & 1));
                const int r_begin = c * kSRowsPerStage, r_end = min(r_begin + kSRowsPerStage, rows);
                mbar_expect_tx(&full[s], row_bytes * (uint32_t)(r_end - r_begin));
                for (int r = r_begin; r < r_end; ++r)
                    tma_bulk_g2s(ring + ((size_t)s * kSRowsPerStage + (r - r_begin)) * strip, in + (long long)(l0 + r) * N + c0,
                                 row_bytes, &full[s]);
            }
        }
        return;
    }

    const int tid = threadIdx.x, lane = tid & 31;
    const int jloc = tid * VEC;
    const bool active = jloc < w;
    const int jc = c0 + (active ? jloc : 0);
    const BubbleParam<T> bp = a.bub[b];
    const T href = bp.h_ref;
    float colA[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) colA[k] = __ldg(a.rgb.colA + jc + k);
    const float Rv = bp.radius;
    const float camR2 = a.rgb.cam2 + Rv * Rv;
    const float href_u = (float)href * a.rgb.nm_over_max;
    uint32_t *rgb_p = reinterpret_cast<uint32_t *>(a.rgb.out) + (long long)b * a.rgb_bubble_stride + (long long)(l0 - 1) * N + jc;
    for (int r = tid; r < rows; r += C) s_sc[r] = __ldg(a.rowSC + (a.row_begin + l0 - 1 + r));
    bar_consumers(C);
    const float2 *sc_p = s_sc;
    const T *my = ring + (active ? jloc : 0);

    for (int c = 0; c < nchunks; ++c) {
        const int s = c % kSStages;
        mbar_wait(&full[s], (uint32_t)((c / kSStages) & 1));
        const int nr = min(kSRowsPerStage, rows - c * kSRowsPerStage);
        T v[kSRowsPerStage][VEC];
#pragma unroll
        for (int r = 0; r < kSRowsPerStage; ++r) {
            if (r < nr) {
                const T *src = my + ((size_t)s * kSRowsPerStage + r) * strip;
                if constexpr (VEC == 4) {
                    const float4 q = *reinterpret_cast<const float4 *>(src);
                    v[r][0] = q.x; v[r][1] = q.y; v[r][2] = q.z; v[r][3] = q.w;
                } else {
                    const double2 q = *reinterpret_cast<const double2 *>(src);
                    v[r][0] = q.x; v[r][1] = q.y;
                }
            }
        }
        uint32_t bits = 0;  // the arrive must not overtake the loads above (see mbar_arrive_after)
#pragma unroll
        for (int r = 0; r < kSRowsPerStage; ++r) {
            if (r < nr) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) bits |= dep_bits(v[r][k]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_after(&empty[s], bits);  // the stage is in registers: hand it back before the colour work
#pragma unroll
        for (int r = 0; r < kSRowsPerStage; ++r) {
            if (r < nr) {
                if (active) {
                    const float2 sc = *sc_p;
                    const float ccy = sc.y * a.rgb.cam_y;
                    uint32_t px[VEC];
                    if constexpr (MODE == SBX_RGB_SPECTRAL) {
                        // the thread's VEC cells share one pass over the wavelength table (rgb_spectral_vec)
                        float dn[VEC], civ[VEC], c3[VEC][3];
#pragma unroll
                        for (int k = 0; k < VEC; ++k) {
                            const float nc = fmaf(sc.x, colA[k], ccy);
                            civ[k] = fabsf(nc - Rv) * rsqrt_approx(fmaf(-2.f * Rv, nc, camR2));
                            dn[k] = ((float)v[r][k] + (float)href) * 1e9f;
                        }
                        rgb_spectral_vec<VEC>(dn, civ, a.rgb, c3);
#pragma unroll
                        for (int k = 0; k < VEC; ++k) px[k] = pack_rgba8_fast(c3[k][0], c3[k][1], c3[k][2]);
                    } else {
#pragma unroll
                        for (int k = 0; k < VEC; ++k)
                            px[k] = shade_px<MODE, T>(v[r][k], href, href_u, colA[k], sc.x, ccy, Rv, camR2, a.rgb);
                    }
                    if constexpr (VEC == 4) *reinterpret_cast<uint4 *>(rgb_p) = make_uint4(px[0], px[1], px[2], px[3]);
                    else *reinterpret_cast<uint2 *>(rgb_p) = make_uint2(px[0], px[1]);
                }
                rgb_p += N;
                ++sc_p;
            }
        }
    }
}

template <typename T>
bool shade_tma_supported(const StepArgs<T> &a)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    return a.num_phi % VEC == 0 && a.num_phi >= 32 * VEC && a.rgb.mode != SBX_RGB_NONE && a.rgb.format == SBX_OUT_RGBA8;
}
template bool shade_tma_supported<float>(const StepArgs<float> &);
template bool shade_tma_supported<double>(const StepArgs<double> &);

template <typename T, int MODE>
static cudaError_t launch_shade_one(const StepArgs<T> &a, dim3 grid, dim3 block, size_t smem, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        // the carveout that just holds the resident rings leaves the rest of the SM's 256 KB to L1, where the LUT texture lives
        static const int carve = [] { const char *e = getenv("SBX_SHADE_CARVEOUT"); return e ? atoi(e) : (int)cudaSharedmemCarveoutMaxShared; }();
        cudaError_t e = cudaFuncSetAttribute(k_shade_tma<T, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    k_shade_tma<T, MODE><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_shade_tma(StepArgs<T> a, int batch, int band_rows_req, cudaStream_t st)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int rows = a.l_last - a.l_first + 1;
    if (rows <= 0) return cudaSuccess;
    int consumers = 256;
    while (consumers > 32 && consumers * VEC / 2 >= a.num_phi) consumers >>= 1;
    const int strip = consumers * VEC;
    const int gx = (a.num_phi + strip - 1) / strip;
    int band_rows = band_rows_req > 0 ? band_rows_req : auto_band_rows(rows, (long long)gx * batch, device_sm_count() * 4, 24);
    if (band_rows > rows) band_rows = rows;
    if (band_rows > kSMaxBandRows) band_rows = kSMaxBandRows;
    a.band_rows = band_rows;
    const dim3 grid((unsigned)gx, (unsigned)((rows + band_rows - 1) / band_rows), (unsigned)batch), block((unsigned)(consumers + 32));
    const size_t smem = (size_t)kSStages * kSRowsPerStage * strip * sizeof(T) + 2 * kSStages * sizeof(uint64_t);
    switch (a.rgb.mode) {
    case SBX_RGB_LUT: return launch_shade_one<T, SBX_RGB_LUT>(a, grid, block, smem, st);
    case SBX_RGB_EXACT3: return launch_shade_one<T, SBX_RGB_EXACT3>(a, grid, block, smem, st);
    case SBX_RGB_EXACT3_LINEAR: return launch_shade_one<T, SBX_RGB_EXACT3_LINEAR>(a, grid, block, smem, st);
    case SBX_RGB_SCHLICK_WALL: return launch_shade_one<T, SBX_RGB_SCHLICK_WALL>(a, grid, block, smem, st);
    case SBX_RGB_SCHLICK_TWO_BEAM: return launch_shade_one<T, SBX_RGB_SCHLICK_TWO_BEAM>(a, grid, block, smem, st);
    case SBX_RGB_SPECTRAL: return launch_shade_one<T, SBX_RGB_SPECTRAL>(a, grid, block, smem, st);
    default: return cudaErrorInvalidValue;
    }
}
template cudaError_t launch_shade_tma<float>(StepArgs<float>, int, int, cudaStream_t);
template cudaError_t launch_shade_tma<double>(StepArgs<double>, int, int, cudaStream_t);

// ------------------------------------------------------------------ caustic map on the ring (row f2)
// caustics_compute.wgsl:29-126: a 9-point stencil of the thickness buffer, phi wrapped, theta clamped at the poles.
// One row per ring slot (strip + a 16-byte halo each side, wrapped across the seam as in sbx_march_tma.cu); an output
// row reads slots l-1, l, l+1 and hands slot l-1 back.  The rows above the first / below the last pole row are the
// pole rows themselves (clamp), rows beyond a latitude band are the neighbours' ghost rows.
namespace {
constexpr int kCSlots = 8;
}

template <typename T>
__global__ void __launch_bounds__(288, 4)
k_caustics_tma(const T *buf, long long bubble_stride, float *out, int N, int NT, int row_begin, int rows_local,
               const double *href, sbx_caustic_params p, int band_rows)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int HALO = VEC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = blockDim.x - 32;
    const int strip = VEC * C;
    const int c0 = blockIdx.x * strip;
    const int w = min(strip, N - c0);
    const int rowf = strip + 2 * HALO;
    T *ring = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kCSlots * rowf * sizeof(T));
    uint64_t *empty = full + kCSlots;
    const int b = blockIdx.z;
    const T *in = buf + (long long)b * bubble_stride;
    const int l0 = 1 + blockIdx.y * band_rows;
    const int l1 = min(l0 + band_rows - 1, rows_local);
    if (l0 > l1) return;
    const int nout = l1 - l0 + 1, nload = nout + 2;
    const int nwarps_c = C >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCSlots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarps_c);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= C) {
        if (threadIdx.x == C) {
            const bool wrapL = c0 == 0, wrapR = c0 + w >= N;
            const uint32_t main_elems = (uint32_t)w + (wrapL ? 0 : HALO) + (wrapR ? 0 : HALO);
            const uint32_t row_bytes = ((uint32_t)w + 2 * HALO) * (uint32_t)sizeof(T);
            for (int r = 0; r < nload; ++r) {
                const int s = r % kCSlots;
                if (r >= kCSlots) mbar_wait(&empty[s], (uint32_t)(((r / kCSlots) - 1) & 1));
                int lr = l0 - 1 + r;                      // local row (0 = ghost above, rows_local + 1 = ghost below)
                const int g = row_begin + lr - 1;         // its global row
                if (g < 0) lr += 1;                       // clamp(y - 1, 0, height - 1)
                if (g > NT - 1) lr -= 1;                  // clamp(y + 1, 0, height - 1)
                mbar_expect_tx(&full[s], row_bytes);
                const T *grow_ptr = in + (long long)lr * N;
                T *dst = ring + (size_t)s * rowf;
                tma_bulk_g2s(dst + (wrapL ? HALO : 0), grow_ptr + c0 - (wrapL ? 0 : HALO), main_elems * (uint32_t)sizeof(T), &full[s]);
                if (wrapL) tma_bulk_g2s(dst, grow_ptr + (N - HALO), 16u, &full[s]);
                if (wrapR) tma_bulk_g2s(dst + HALO + w, grow_ptr, 16u, &full[s]);
            }
        }
        return;
    }

    const int tid = threadIdx.x, lane = tid & 31;
    const int jloc = tid * VEC;
    const bool active = jloc < w;
    const int soff = HALO + (active ? jloc : 0);
    const float hr = href ? (float)href[b] : 0.f;
    float *dst = out + ((long long)b * rows_local + (l0 - 1)) * N + c0 + (active ? jloc : 0);

    auto load_row = [&](float (&r)[VEC + 2], int slot) {
        const T *src = ring + (size_t)slot * rowf + soff;
        if constexpr (VEC == 4) {
            const float4 q = *reinterpret_cast<const float4 *>(src);
            r[1] = q.x + hr; r[2] = q.y + hr; r[3] = q.z + hr; r[4] = q.w + hr;
        } else {
            const double2 q = *reinterpret_cast<const double2 *>(src);
            r[1] = (float)q.x + hr; r[2] = (float)q.y + hr;
        }
        r[0] = (float)src[-1] + hr;
        r[VEC + 1] = (float)src[VEC] + hr;
    };

    float up[VEC + 2], cur[VEC + 2], dn[VEC + 2];
    mbar_wait(&full[0], 0);
    load_row(up, 0);
    mbar_wait(&full[1 % kCSlots], 0);
    load_row(cur, 1);
#pragma unroll 1
    for (int i = 0; i < nout; ++i, dst += N) {
        const int rl = i + 2, s = rl % kCSlots;
        mbar_wait(&full[s], (uint32_t)((rl / kCSlots) & 1));
        load_row(dn, s);
        __syncwarp();
        // row i (the `up` row) is in registers everywhere in this warp: its loads were consumed by the `+ hr` adds of
        // load_row two iterations ago, which precede this arrive in program order (in-order issue), so a plain arrive is safe
        if (lane == 0) mbar_arrive(&empty[i % kCSlots]);
        if (active) {
            float o[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                o[j] = caustic_intensity(cur[1 + j], cur[j], cur[2 + j], up[1 + j], dn[1 + j], up[j], up[2 + j], dn[j], dn[2 + j], p);
            if constexpr (VEC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            else *reinterpret_cast<float2 *>(dst) = make_float2(o[0], o[1]);
        }
#pragma unroll
        for (int j = 0; j < VEC + 2; ++j) { up[j] = cur[j]; cur[j] = dn[j]; }
    }
}

template <typename T>
cudaError_t launch_caustics_tma(const T *buf, long long bubble_stride, float *out, int num_phi, int num_theta, int row_begin,
                                int rows_local, int batch, const double *href, const sbx_caustic_params &p, cudaStream_t st)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    int consumers = 256;
    while (consumers > 32 && consumers * VEC / 2 >= num_phi) consumers >>= 1;
    const int strip = consumers * VEC;
    const int gx = (num_phi + strip - 1) / strip;
    int band_rows = auto_band_rows(rows_local, (long long)gx * batch, device_sm_count() * 4, 32);
    const dim3 grid((unsigned)gx, (unsigned)((rows_local + band_rows - 1) / band_rows), (unsigned)batch), block((unsigned)(consumers + 32));
    const size_t smem = (size_t)kCSlots * (strip + 2 * VEC) * sizeof(T) + 2 * kCSlots * sizeof(uint64_t);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_caustics_tma<T>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    k_caustics_tma<T><<<grid, block, smem, st>>>(buf, bubble_stride, out, num_phi, num_theta, row_begin, rows_local, href, p, band_rows);
    return cudaGetLastError();
}
template cudaError_t launch_caustics_tma<float>(const float *, long long, float *, int, int, int, int, int, const double *, const sbx_caustic_params &, cudaStream_t);
template cudaError_t launch_caustics_tma<double>(const double *, long long, float *, int, int, int, int, int, const double *, const sbx_caustic_params &, cudaStream_t);

}  // namespace sbx
