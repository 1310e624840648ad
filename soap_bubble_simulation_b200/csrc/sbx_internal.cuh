// sbx_internal.cuh -- shared device/host structures of libsbx (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sbx.h"

namespace sbx {

constexpr int kMaxFronts = 8;  // src/physics/drainage.rs:474

// Integer description of the cells one TFE front covers this step.  Membership is decided on the
// host with the reference's exact f64 predicate (TfeFront::contains, drainage.rs:55-71) and handed to
// the kernels as closed global-row / column intervals, so the device test is pure integer compares.
struct FrontCells {
    int row_lo, row_hi;  // global theta rows, inclusive; row_lo > row_hi = empty
    int c0_lo, c0_hi;    // direct phi interval, inclusive (empty if lo > hi)
    int c1_lo, c1_hi;    // wrapped part of the interval (touches column 0 or num_phi-1)
};

struct FrontTable {
    int count;
    int row_min, row_max;  // union row range the reference iterates (drainage.rs:425-440)
    FrontCells f[kMaxFronts];
    // resident multi-step kernel only: do the row / column intervals differ from the previous step's table?  (Row
    // intervals move every few hundred steps per front, column intervals only when a front is born or dies.)
    int rows_changed, cols_changed;
};

// per-bubble constants (foam batches may vary radius and reference thickness per bubble)
template <typename T>
struct BubbleParam {
    T diff_scale;  // (radius_ref / radius_b)^2 applied to the diffusion terms
    T h_ref;       // F32 storage offset (0 for f64)
    float radius;  // for the view cosine
    float pad;
};

// colour-stage arguments
struct RgbArgs {
    int mode;    // sbx_rgb_mode
    int format;  // sbx_rgb_format
    void *out;   // RGBA8 (uint32 per cell) or 3 floats per cell, same [bubble][row][phi] order as the
                 // owned rows (no ghost rows)
    const float *colA;  // per column: cos(phi)*cam_x + sin(phi)*cam_z
    float cam_y, cam2;  // cam_y, |cam|^2
    float n_film, inv_n2;
    float intensity;
    float inv_max_nm;  // 1 / LUT max thickness
    float nm_over_max; // 1e9 / LUT max thickness: metres -> LUT u coordinate
    cudaTextureObject_t lut;
    float inv_wl_f[3];        // 1/lambda_k in f32
    float two_pi_over_wl_f[3];  // 2 pi / lambda_k in f32 (two-beam model)
    float two_n, r0, omr0;    // 2n, Schlick R0 = ((n-1)/(n+1))^2, 1 - R0
    double inv_n2_d;          // 1/n^2 in f64 for the phase
    double two_n_over_wl[3];  // 2 n / lambda_k at 650, 532, 450 nm (interference.rs:26-30)
    // the same two constants split into float pairs (hi + lo == the f64 value to 2^-48) for the float-float phase
    float inv_n2_hi, inv_n2_lo;
    float w_hi[3], w_lo[3];
    // SBX_RGB_SPECTRAL: per wavelength {2 pi / lambda, wr, wg, wb} with w = (XYZ -> linear sRGB)(cie(lambda)) / n
    const float4 *spec;
    int n_spec;
};

// Peer-memory halo (latitude-band sharding, one process per GPU): the band next to a neighbour reads the row it
// needs straight out of the neighbour's field over NVLink (a bulk copy from the peer address into the shared-memory
// ring), and the neighbours order their steps with two monotonic flags instead of a collective.  No kernel writes
// remote memory: each band publishes "finished step e" in its own device memory and the neighbour polls that word.
// (Round 1-2 pushed the boundary rows and the flag into the neighbour instead; remote stores made the end of every
// launch ~5 us longer at 1024 rows per GPU, profiles/r02_peer_protocol_cost.txt.)  See sbx_march_tma.cu.
struct PeerArgs {
    const void *up_row;                        // up neighbour's IN buffer of this step, its last owned row (nullptr: none); element = T
    const void *dn_row;                        // down neighbour's IN buffer of this step, its first owned row (nullptr: none)
    const void *up_row2, *dn_row2;             // WGSL semantics: the same rows of the concentration field
    const unsigned long long *flag_up;         // up neighbour's "my last band finished step e" word (peer memory, polled)
    const unsigned long long *flag_dn;         // down neighbour's "my first band finished step e" word (peer memory, polled)
    unsigned long long *my_first_done;         // mine, local: published when every CTA of my first band has finished step e
    unsigned long long *my_last_done;          // mine, local: the same for my last band
    unsigned int *done_first, *done_last;      // CTA counters of my first / last band
    unsigned int *status;                      // bit 0 set: a neighbour wait timed out (same word as StepArgs::dev_status)
    unsigned long long epoch;                  // number of this step, >= 1
    int enabled;
#ifdef SBX_PEER_DEBUG
    unsigned debug;  // timing experiments only (results are wrong): 1 no wait, 4 no signal, 8 no system fence
#endif
};

// arguments of one drainage step (CPU semantics)
template <typename T>
struct StepArgs {
    const T *in;   // local buffer incl. ghost row at local row 0
    T *out;
    const T *coefS;          // per global row: {G, B, E, 0} (see host: upload_tables_t)
    const T *coefF;          // TMA fast path, per global row: {A+B, A-B, E, G}
    const float2 *rowSC;     // per global row: {sin theta, cos theta}
    const double4 *coefX;    // strict mode per global row: {sin theta, cos/sin_safe, sin_safe^2, 0}
    const BubbleParam<T> *bub;
    double *pole_partial;    // [batch][2][pole_slots]
    unsigned long long *pole_count;  // [batch][2] monotonic arrival counters (TMA kernel's folded pole rows)
    unsigned long long pole_want[2]; // arrivals per strip count the folded launches: wait for pole_want * strips
    int fold_top, fold_bottom;       // TMA kernel: this launch also writes the top / bottom pole row
    unsigned int *dev_status;        // bit mask of in-kernel waits that timed out: 1 = latitude neighbour, 2 = pole ring
    int pole_slots;
    int num_phi, num_theta;  // global grid
    int row_begin;           // global row of local row 1
    int l_first, l_last;     // local rows to update, inclusive (global poles excluded)
    int band_rows;           // rows marched by one block
    // TMA kernel, latitude-band sharding with the in-kernel halo push: the band next to a neighbour is short (top_rows /
    // bot_rows rows; 0 = no special band) and, with boundary_first, dispatched before the interior bands, so that its
    // boundary row reaches the neighbour -- and the "step finished here" flag is released -- early in the launch
    int top_rows, bot_rows, boundary_first;
    int ring_stages;         // TMA kernel: stages of the shared-memory ring in this launch (2 or 3)
    int l2_keep_edges;       // TMA kernel: L2 evict_last policy on the rows neighbouring bands share (fields larger than L2)
    // TMA kernel: min / max of the stored values this launch writes, as a by-product (nullptr: off).  Two
    // order-preserving unsigned words {min, max} per bubble (32-bit for f32 storage, 64-bit for f64), merged with
    // atomicMin / atomicMax once per CTA; the launch also re-arms the other set for the next step.
    void *mm_cur, *mm_next;
    long long bubble_stride; // elements per bubble in in/out: (rows_local + 2) * num_phi
    long long rgb_bubble_stride;  // cells per bubble in rgb out: rows_local * num_phi
    T dt, crit, coefA;       // coefA = D / (R^2 dtheta^2)
    T tfe_rdt;               // (1 - ratio) * 0.5 * dt
    // strict-mode scalars, computed on the host exactly as the reference does (drainage.rs:305,353-369)
    double s_drain_coef, s_dth2, s_two_dth, s_dph2, s_r2, s_D, s_dt, s_crit, s_tfe_factor;
    FrontTable fronts;
    RgbArgs rgb;
    PeerArgs peer;
};

// arguments of one WGSL-semantics dispatch
struct WgslArgs {
    const float *h_in, *c_in;
    float *h_out, *c_out;
    const float *rowW;  // per global row float4: {sin, cos/sin_safe, 1/sin_safe^2, 1/sin_safe}
    const float2 *rowSC;
    int num_phi, num_theta, row_begin, l_first, l_last, band_rows;
    int rows_local;
    int l2_keep_edges;  // as in StepArgs
    long long bubble_stride, rgb_bubble_stride;
    sbx_drainage_params p;
    float delta_theta, delta_phi;
    float radius_view;
    RgbArgs rgb;
    PeerArgs peer;  // in-kernel halo push of both fields (k_wgsl_tma)
};

// arguments of one headless bubble frame (sbx_frame.cu; rows f3/f4)
struct FrameArgs {
    int width, height;
    float cam[3], fwd[3], right[3], up[3];
    float tan_half_x, tan_half_y;
    float radius, cam2;
    float base_nm, film_time, swirl, drainage_speed, pattern_scale;
    float intensity, base_alpha, edge_alpha;
    int edge_mode;
    float bg[3];
    float inv_max_nm;
    cudaTextureObject_t lut;
    const void *field;  // nullptr: procedural thickness; else [num_theta][num_phi] stored thickness (metres - h_ref)
    int field_f64;
    int num_theta, num_phi;
    double h_ref;
    uint32_t *out;
};

}  // namespace sbx
