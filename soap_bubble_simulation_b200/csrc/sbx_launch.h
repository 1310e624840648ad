// sbx_launch.h -- host-callable launchers implemented in sbx_kernels.cu (internal).
#pragma once

#include "sbx_internal.cuh"

namespace sbx {

// Rows per latitude band for the marching kernels, from same-box scans on B200 (profiles/r02_ab_band_rows*.txt):
//  * launches of four waves of CTAs or more take the default height.  With the L1-friendly carveout of round 2, 24-row bands
//    measured best at 8192 x 16384 (fused 528 Gcell/s against 511-517 at the 56 rows the whole-wave model used to pick;
//    drainage only 820 against 788-807) and at 4096 x 16384 (511 against 492): a partial last wave is cheap when there are
//    several, shorter bands start more streams early and keep the band-edge rows in L2;
//  * two to four waves: the height in [default, default + 8] that fills its last wave best (2048 x 16384: 28 rows = exactly
//    two waves, 494 Gcell/s against 441 for one wave of 56-row bands and 474 at 24 rows);
//  * fewer: a band height that leaves the last wave half empty costs its full time -- 2048 x 4096 at 24 rows is 344 CTAs on
//    592 slots, at 14 rows 588: +30 %; 1024 x 16384: 28 rows = one full wave.  Cost model there: waves x (band rows + ring
//    fill / halo overhead).
inline int auto_band_rows(int rows, long long ctas_per_band, int slots, int dflt)
{
    if (rows <= 0) return 1;
    const long long at_dflt = ctas_per_band * ((rows + dflt - 1) / dflt);
    if (slots <= 0 || at_dflt >= 4LL * slots) return dflt < rows ? dflt : rows;
    if (at_dflt >= 2LL * slots) {
        int best = dflt;
        double best_fill = -1.0;
        for (int b = dflt; b <= dflt + 8 && b <= rows; ++b) {
            const long long ctas = ctas_per_band * ((rows + b - 1) / b);
            const long long waves = (ctas + slots - 1) / slots;
            const double fill = (double)ctas / (double)(waves * slots);
            if (fill > best_fill + 1e-9) { best_fill = fill; best = b; }
        }
        return best;
    }
    int best = dflt < rows ? dflt : rows;
    double best_cost = 1e300;
    for (int b = 6; b <= 64 && b <= rows; ++b) {
        const long long bands = (rows + b - 1) / b;
        const long long waves = (ctas_per_band * bands + slots - 1) / slots;
        const double cost = (double)waves * (b + 4.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = b; }
    }
    return best;
}
int device_sm_count();

// one drainage step in CPU semantics: marching stencil (+ fused colour) then the pole rows.
// `launches` is incremented once per kernel launched.
template <typename T>
cudaError_t launch_step(StepArgs<T> a, int batch, bool tfe, bool strict, bool has_top, bool has_bottom,
                        int band_rows_req, bool use_tma, cudaStream_t st, int *launches);

// fast path (sbx_march_tma.cu): TMA-staged ring, f32 or f64 (non-strict), RGBA8 or no colour, whole 16-byte packets
template <typename T>
bool march_tma_supported(const StepArgs<T> &a);
template <typename T>
bool march_tma_folds_poles(const StepArgs<T> &a);
template <typename T>
cudaError_t launch_march_tma(StepArgs<T> a, int batch, bool tfe, int band_rows_req, cudaStream_t st, int *pole_slots);

// shared-memory-resident multi-step kernels for small whole-sphere grids (sbx_resident.cu): cluster width that
// would be used, 0 = not applicable
template <typename T>
int resident_cluster(const StepArgs<T> &a, int batch, bool strict);
template <typename T>
cudaError_t launch_resident(StepArgs<T> a, int batch, bool strict, const FrontTable *tables_dev, int nsteps,
                            cudaStream_t st);
// colour stages the resident kernels' epilogue covers (RGBA8 frame of the final field, shaded from shared memory)
bool resident_shades(const RgbArgs &rgb);
int resident_cluster_wgsl(const WgslArgs &a, int batch);
cudaError_t launch_resident_wgsl(WgslArgs a, int batch, int nsteps, cudaStream_t st);

template <typename T>
cudaError_t launch_tfe_pass(T *buf, long long bubble_stride, int num_phi, int row_begin, int rows_local,
                            int batch, const FrontTable &ft, double factor, cudaStream_t st, int *launches);

template <typename T>
cudaError_t launch_shade(StepArgs<T> a, int batch, int band_rows_req, cudaStream_t st, int *launches, bool use_tma = false);
// TMA-staged ring variant (sbx_shade_tma.cu): whole 16-byte packets per row, RGBA8 output
template <typename T>
bool shade_tma_supported(const StepArgs<T> &a);
template <typename T>
cudaError_t launch_shade_tma(StepArgs<T> a, int batch, int band_rows_req, cudaStream_t st);

cudaError_t launch_wgsl(WgslArgs a, int batch, int band_rows_req, bool use_tma, cudaStream_t st, int *launches);
// TMA-staged ring variant (sbx_wgsl_tma.cu): num_phi % 4 == 0, RGBA8 or no colour
bool wgsl_tma_supported(const WgslArgs &a);
cudaError_t launch_wgsl_tma(WgslArgs a, int batch, int band_rows_req, cudaStream_t st);

template <typename TD, typename TS>
cudaError_t launch_convert_in(TD *dst, const TS *src, long long n, long long bubble_stride, const double *href, cudaStream_t st);
template <typename TD, typename TS>
cudaError_t launch_convert_out(TD *dst, const TS *src, long long n, long long bubble_stride, const double *href, cudaStream_t st);
template <typename T>
cudaError_t launch_fill(T *dst, long long n, T v, cudaStream_t st);
template <typename T>
cudaError_t launch_minmax(const T *buf, int batch, long long bubble_stride, long long first, long long n,
                          const double *href, double *part, int nblocks, cudaStream_t st);
template <typename T>
cudaError_t launch_sample(const T *buf, long long i00, long long i01, long long i10, long long i11, double tf,
                          double pf, double h_ref, double *out, cudaStream_t st);

cudaError_t launch_lut_build(uint8_t *out, int T, int A, float max_nm, float n_film, float intensity,
                             const float *wl_dev, int n_lambda, cudaStream_t st);
cudaError_t launch_interference_eval(double n_film, const double wl[3], const double *d_nm, const double *cos_i,
                                     size_t n, int gamma, double *rgb, cudaStream_t st);
cudaError_t launch_interference_terms(double n_film, const double *d_nm, const double *cos_i, size_t n, double wl,
                                      double *out_cos_t, double *out_fresnel, double *out_intensity, cudaStream_t st);
cudaError_t launch_lut_shade_points(cudaTextureObject_t lut, float inv_max_nm, float intensity, const float *d_nm,
                                    const float *cos_theta, size_t n, float *rgb, cudaStream_t st);
template <typename T>
cudaError_t launch_caustics(const T *buf, long long bubble_stride, float *out, int num_phi, int num_theta, int row_begin,
                            int rows_local, int batch, const double *href, const sbx_caustic_params &p, cudaStream_t st,
                            bool use_tma = false);
template <typename T>
cudaError_t launch_caustics_tma(const T *buf, long long bubble_stride, float *out, int num_phi, int num_theta, int row_begin,
                                int rows_local, int batch, const double *href, const sbx_caustic_params &p, cudaStream_t st);
template <typename T>
cudaError_t launch_sample_uv(const T *field, int W, int H, float href, const float *uv, size_t n, float *out_h,
                             float *out_grad, cudaStream_t st);
template <typename T>
cudaError_t launch_sample_normal(const T *field, int W, int H, float href, const float *nrm, size_t n, float scale,
                                 float *out_h, cudaStream_t st);
cudaError_t launch_col_view(float *colA, int num_phi, double cam_x, double cam_z, cudaStream_t st);

// the bubble fragment shader as a compute kernel (sbx_frame.cu): one thread per pixel of a headless frame
cudaError_t launch_bubble_frame(const FrameArgs &a, cudaStream_t st);

}  // namespace sbx
