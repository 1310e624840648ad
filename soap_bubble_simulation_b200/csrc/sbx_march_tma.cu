// sbx_march_tma.cu -- the f32 fast path of the fused drainage + interference step for sm_100a.
//
// Each CTA owns a strip of up to 4*C consecutive phi columns (C = consumer threads) and marches down a
// band of theta rows.  A dedicated producer warp streams the strip's rows -- main segment plus a 4-float
// halo on each side, wrapped across the periodic seam when the strip touches column 0 or num_phi-1 --
// into a shared-memory ring with TMA bulk copies (cp.async.bulk, SASS UBLKCP) that complete on
// mbarriers; the consumer warps pull each row out of the ring exactly once (one LDS.128 + two LDS.32
// per thread), keep a three-row window in registers, and write h' and RGBA8 with 16-byte stores.
// Compared with the direct-load kernel (k_march) this removes every global-load instruction, the warp
// shuffles and the predicated edge loads from the instruction stream, which is what bounds the fused
// step (profiles/: k_march is issue-slot bound at 77 instructions per cell).
#include <cstdlib>

#include "sbx_device.cuh"
#include "sbx_launch.h"
#include "sbx_tma.cuh"

namespace sbx {

namespace {

constexpr int kStages = 3;        // ring depth at most (barriers are laid out for it); StepArgs::ring_stages = 2 or 3 per launch
constexpr int kRowsPerStage = 3;  // theta rows per stage == the unroll of the three-row register window
constexpr int kMaxBandRows = 64;
  // rows a CTA marches at most (their coefficients are staged in shared memory)

template <typename T, int VEC>
struct Row {
    T v[VEC];
    T eL, eR;
};

template <typename T> __device__ __forceinline__ T tmax(T a, T b);
template <> __device__ __forceinline__ float tmax<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double tmax<double>(double a, double b) { return fmax(a, b); }

// order-preserving map of a floating-point value onto an unsigned word (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned int ordered_bits(float v)
{
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : b | 0x80000000u;
}
__device__ __forceinline__ unsigned long long ordered_bits(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : b | 0x8000000000000000ull;
}
template <typename T> struct OrderedWord { using type = unsigned int; };
template <> struct OrderedWord<double> { using type = unsigned long long; };
template <typename T> __device__ __forceinline__ T tmin(T a, T b);
template <> __device__ __forceinline__ float tmin<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double tmin<double>(double a, double b) { return fmin(a, b); }

template <typename T> struct MinBlocks { static constexpr int value = 4; };
template <> struct MinBlocks<double> { static constexpr int value = 3; };

}  // namespace

// T = float (VEC 4) or double (VEC 2): a row travels in 16-byte packets; a thread owns PK consecutive packets of a row
// (W = VEC * PK cells), a strip is 16 * PK * C bytes wide plus a 16-byte halo each side, so ring geometry and copy sizes
// are identical for both types.  PK = 2 (f32; A/B builds with -DSBX_ENABLE_PK2 and SBX_PK=2): half the consumer threads
// with twice the cells each, so the per-row bookkeeping is amortised over 8 cells and eight texture fetches per thread are
// in flight.  Measured and NOT adopted (profiles/r02_ab_pk_rejected.txt): 8192 x 16384 unchanged (506.7 vs 508.9 Gcell/s --
// the step is bound by the memory system there, not by issue slots), while every launch with short bands loses the
// latency hiding of the second set of warps (2048 x 4096: 400 vs 464, foam 4096 bubbles: 396 vs 498).
template <typename T, int MODE, bool TFE, int PK>
__global__ void __launch_bounds__(PK == 2 ? 160 : 288, MinBlocks<T>::value) k_march_tma(const StepArgs<T> a)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int W = VEC * PK;  // cells per thread and row
    constexpr int HALO = VEC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double warp_part[8];
    // per-row coefficients {A+B, A-B, E, G} and {sin, cos} of this band, staged once: a row then costs two LDS
    // instead of two global loads on the critical path of its first instruction
    __shared__ __align__(16) T s_coef[kMaxBandRows][4];
    __shared__ float2 s_sc[kMaxBandRows];
    __shared__ unsigned s_rm[kMaxBandRows];  // TFE: which fronts cover each row of the band (bit per front)
    const int C = blockDim.x - 32;  // consumer threads
    const int N = a.num_phi;
    const int strip_w_max = W * C;
    const int c0 = blockIdx.x * strip_w_max;
    const int w = min(strip_w_max, N - c0);   // multiple of VEC
    const int rowf = strip_w_max + 2 * HALO;  // elements per ring row
    const int stages = a.ring_stages;  // 2 or 3 (uniform)
    T *ring = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * kRowsPerStage * rowf * sizeof(T));
    uint64_t *empty = full + kStages;

    const int b = blockIdx.z;
    const T *in = a.in + (long long)b * a.bubble_stride;
    int l0, l1;
    {
        const int gy = (int)gridDim.y, ht = a.top_rows > 0 ? 1 : 0, hb = a.bot_rows > 0 ? 1 : 0;
        int band = (int)blockIdx.y;
        // boundary first: y = 0 is the band next to the up neighbour, y = 1 the one next to the down neighbour
        if (a.boundary_first && hb && gy > 2) band = band == 0 ? 0 : (band == 1 ? gy - 1 : band - 1);
        if (ht && band == 0) { l0 = a.l_first; l1 = l0 + a.top_rows - 1; }
        else if (hb && band == gy - 1) { l1 = a.l_last; l0 = l1 - a.bot_rows + 1; }
        else {
            l0 = a.l_first + a.top_rows + (band - ht) * a.band_rows;
            l1 = min(l0 + a.band_rows - 1, a.l_last - a.bot_rows);
        }
    }
    if (l0 > l1) return;
    const int rows_load = (l1 - l0 + 1) + 2;  // local rows l0-1 .. l1+1
    const int nchunks = (rows_load + kRowsPerStage - 1) / kRowsPerStage;
    const int nwarps_c = C >> 5;

    // Programmatic dependent launch: let the next step's grid start filling SMs as this one drains, and do our
    // own set-up (barrier init) before waiting for the previous step's grid to finish and flush its writes.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarps_c);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below may read the previous step's output

    if (threadIdx.x >= C) {
        // ------------------------------------------------------------ producer warp (one lane issues)
        if (threadIdx.x == C) {
            const bool touch_top = a.peer.enabled && a.peer.up_row && l0 == 1;
            const bool touch_bot = a.peer.enabled && a.peer.dn_row && l1 == a.l_last;
            const bool wrapL = c0 == 0, wrapR = c0 + w >= N;
            const uint32_t main_elems = (uint32_t)w + (wrapL ? 0 : HALO) + (wrapR ? 0 : HALO);
            const uint32_t row_bytes = ((uint32_t)w + 2 * HALO) * (uint32_t)sizeof(T);
            // Fields larger than L2: the band's two halo rows and its first / last owned rows are read again by the
            // neighbouring band about one band-time later, by when the streaming traffic has evicted them (ncu: 3.7 %
            // of the reads reached DRAM twice).  An evict_last policy on just those rows keeps them: +1.7 % at
            // 8192 x 16384; on an L2-resident field (2048 x 4096) the same hint costs 7 %, so the host decides.
            const bool keep_edges = a.l2_keep_edges != 0;
            const uint64_t pol = keep_edges ? l2_policy_evict_last() : 0ull;
            int s = 0, round = 0;  // stage of chunk c and how often the ring has wrapped
            for (int c = 0; c < nchunks; ++c, s = (s + 1 == stages ? 0 : s + 1), round += (s == 0)) {
                if (round > 0) mbar_wait(&empty[s], (uint32_t)((round - 1) & 1));
                const int r_begin = c * kRowsPerStage;
                const int r_end = min(r_begin + kRowsPerStage, rows_load);
                mbar_expect_tx(&full[s], row_bytes * (uint32_t)(r_end - r_begin));
                for (int r = r_begin; r < r_end; ++r) {
                    // A band next to a neighbour takes that row out of the neighbour's field (peer memory): wait until the
                    // neighbour's band has finished the previous step, then copy from there instead of the local ghost row
                    const T *grow_ptr = in + (long long)(l0 - 1 + r) * N;
                    bool remote = false;
                    if (r == 0 && touch_top) {
#ifdef SBX_PEER_DEBUG
                        if (!(a.peer.debug & 1u))
#endif
                        wait_neighbour(a.peer.flag_up, a.peer.epoch - 1, a.peer.status);
                        grow_ptr = reinterpret_cast<const T *>(a.peer.up_row);
                        remote = true;
                    }
                    if (r == rows_load - 1 && touch_bot) {
#ifdef SBX_PEER_DEBUG
                        if (!(a.peer.debug & 1u))
#endif
                        wait_neighbour(a.peer.flag_dn, a.peer.epoch - 1, a.peer.status);
                        grow_ptr = reinterpret_cast<const T *>(a.peer.dn_row);
                        remote = true;
                    }
                    T *dst = ring + ((size_t)s * kRowsPerStage + (r - r_begin)) * rowf;
                    // ring row layout: [0,HALO) left halo | [HALO, HALO+w) main | [HALO+w, 2 HALO+w) right halo
                    if (keep_edges && !remote && (r <= 1 || r >= rows_load - 2))
                        tma_bulk_g2s_hint(dst + (wrapL ? HALO : 0), grow_ptr + c0 - (wrapL ? 0 : HALO),
                                          main_elems * (uint32_t)sizeof(T), &full[s], pol);
                    else
                        tma_bulk_g2s(dst + (wrapL ? HALO : 0), grow_ptr + c0 - (wrapL ? 0 : HALO),
                                     main_elems * (uint32_t)sizeof(T), &full[s]);
                    if (wrapL) tma_bulk_g2s(dst, grow_ptr + (N - HALO), 16u, &full[s]);       // periodic seam
                    if (wrapR) tma_bulk_g2s(dst + HALO + w, grow_ptr, 16u, &full[s]);         // periodic seam
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // stage the band's row coefficients while the producer's first copies are in flight
    for (int r = tid; r < l1 - l0 + 1; r += C) {
        const long long g = a.row_begin + l0 - 1 + (long long)r;
        if constexpr (VEC == 4) {
            *reinterpret_cast<float4 *>(&s_coef[r][0]) = __ldg(reinterpret_cast<const float4 *>(a.coefF + 4 * g));
        } else {
            *reinterpret_cast<double2 *>(&s_coef[r][0]) = __ldg(reinterpret_cast<const double2 *>(a.coefF + 4 * g));
            *reinterpret_cast<double2 *>(&s_coef[r][2]) = __ldg(reinterpret_cast<const double2 *>(a.coefF + 4 * g) + 1);
        }
        s_sc[r] = __ldg(a.rowSC + g);
        if constexpr (TFE) s_rm[r] = front_row_mask(a.fronts, (int)g);
    }
    bar_consumers(C);
    const int jloc = tid * W;  // column offset inside the strip
    const bool active = jloc < w;
    const int jc = c0 + (active ? jloc : 0);
    const int soff = HALO + (active ? jloc : 0);  // element offset of this thread's packet in a ring row
    T *out = a.out + (long long)b * a.bubble_stride;
    const BubbleParam<T> bp = a.bub[b];
    const T href = bp.h_ref;
    const T neg_href = -href;
    const T dscale = bp.diff_scale;
    const T dt = a.dt;
    const T crit = a.crit;

    float colA[W];
#pragma unroll
    for (int k = 0; k < W; ++k) colA[k] = 0.f;
    if constexpr (MODE != SBX_RGB_NONE) {
#pragma unroll
        for (int k = 0; k < W; ++k) colA[k] = __ldg(a.rgb.colA + jc + k);
    }
    // TFE: a front's phi interval does not depend on the row, and a thread keeps its columns for the whole band, so
    // the column half of TfeFront::contains (drainage.rs:55-71) is taken once: byte k of colmask = fronts covering
    // column jc + k.  A cell is then in a front iff (row mask & its byte) != 0 -- one AND per row plus a test per cell
    // instead of 8 x 4 compares per cell (C3 with 8 fronts alive was 164 Gcell/s with the per-cell compares).
    unsigned colmask[PK];  // one byte per cell, one word per packet
#pragma unroll
    for (int p = 0; p < PK; ++p) colmask[p] = 0;
    if constexpr (TFE) {
#pragma unroll
        for (int k = 0; k < W; ++k) colmask[k / VEC] |= front_col_mask(a.fronts, jc + k) << (8 * (k % VEC));
    }
    const float Rv = bp.radius;
    const float camR2 = a.rgb.cam2 + Rv * Rv;
    const float href_u = (float)href * a.rgb.nm_over_max;

    // Ring bookkeeping is static: a stage holds exactly the three rows one trip of the unrolled loop pulls
    // (sub-step A takes slot 2 and releases the stage, B waits for the next stage and takes slot 0, C takes
    // slot 1), so no per-row counters or modulo arithmetic are needed.
    int stage = 0;
    uint32_t full_parity = 0;
    const T *ring_t = ring + soff;                 // this thread's packet in ring row 0
    const uint32_t stage_elems = kRowsPerStage * rowf;
    const T *stage_p = ring_t;                     // this thread's packet in slot 0 of the current stage
    using RowT = Row<T, W>;

    auto load_slot = [&](RowT &r, int slot) {
        const T *src = stage_p + slot * rowf;
#pragma unroll
        for (int p = 0; p < PK; ++p) {
            if constexpr (VEC == 4) {
                const float4 q = *reinterpret_cast<const float4 *>(src + p * VEC);
                r.v[p * VEC + 0] = q.x; r.v[p * VEC + 1] = q.y; r.v[p * VEC + 2] = q.z; r.v[p * VEC + 3] = q.w;
            } else {
                const double2 q = *reinterpret_cast<const double2 *>(src + p * VEC);
                r.v[p * VEC + 0] = q.x; r.v[p * VEC + 1] = q.y;
            }
        }
        r.eL = src[-1];
        r.eR = src[W];
    };
    // `bits`: OR of the registers whose loads have not been consumed by an instruction yet (see mbar_arrive_after)
    auto release_stage = [&](uint32_t bits) {
        __syncwarp();
        if (lane == 0) mbar_arrive_after(&empty[stage], bits);  // every lane of this warp has read the stage
        if (++stage == stages) { stage = 0; full_parity ^= 1u; stage_p = ring_t; }
        else stage_p += stage_elems;
    };
    auto row_bits = [&](const RowT &r) {
        uint32_t b = dep_bits(r.eL) | dep_bits(r.eR);
#pragma unroll
        for (int k = 0; k < W; ++k) b |= dep_bits(r.v[k]);
        return b;
    };
    auto store_px = [&](uint32_t *dst, const uint32_t *px) {
        // (streaming __stcs stores for the frame and/or h' were measured: -0.8 % / -1 %, so plain stores)
#pragma unroll
        for (int p = 0; p < PK; ++p) {
            if constexpr (VEC == 4) *reinterpret_cast<uint4 *>(dst + p * VEC) = make_uint4(px[p * VEC], px[p * VEC + 1], px[p * VEC + 2], px[p * VEC + 3]);
            else *reinterpret_cast<uint2 *>(dst + p * VEC) = make_uint2(px[p * VEC], px[p * VEC + 1]);
        }
    };
    auto store_h = [&](T *dst, const T *v) {  // PK 16-byte packets
#pragma unroll
        for (int p = 0; p < PK; ++p) {
            if constexpr (VEC == 4) *reinterpret_cast<float4 *>(dst + p * VEC) = make_float4(v[p * VEC], v[p * VEC + 1], v[p * VEC + 2], v[p * VEC + 3]);
            else *reinterpret_cast<double2 *>(dst + p * VEC) = make_double2(v[p * VEC], v[p * VEC + 1]);
        }
    };

    // running output pointers (bumped by one row per sub-step)
    T *out_p = out + (long long)l0 * N + jc;
    uint32_t *rgb_p = reinterpret_cast<uint32_t *>(a.rgb.out) + (long long)b * a.rgb_bubble_stride +
                      (long long)(l0 - 1) * N + jc;
    const T *cf_p = &s_coef[0][0];
    const float2 *sc_p = s_sc;
    const unsigned *rm_p = s_rm;
    int grow = a.row_begin + l0 - 1;
    const int ring_top = 1, ring_bot = a.num_theta - 2;
    const bool has_ring = (grow <= ring_top && ring_top <= a.row_begin + l1 - 1) ||
                          (grow <= ring_bot && ring_bot <= a.row_begin + l1 - 1);  // block-uniform

    // min / max by-product (a.mm_cur != nullptr): running extremes of what this thread stores, NaN ignored like f64::min/max
#ifdef SBX_NO_MM_TRACK  // A/B builds: the kernel without the by-product code
    constexpr bool track = false;
#else
    const bool track = a.mm_cur != nullptr;
#endif
    T run_lo = (T)INFINITY, run_hi = (T)-INFINITY;
    if (track && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {  // re-arm the other set for the next step
        using U = typename OrderedWord<T>::type;
        U *nx = static_cast<U *>(a.mm_next) + 2 * b;
        nx[0] = ~(U)0;
        nx[1] = (U)0;
    }

    auto compute_row = [&](const RowT &up, const RowT &cur, const RowT &dn) {
        T cApB, cAmB, cE, cG;  // {A+B, A-B, E, G}
        if constexpr (VEC == 4) {
            const float4 cf = *reinterpret_cast<const float4 *>(cf_p);
            cApB = cf.x; cAmB = cf.y; cE = cf.z; cG = cf.w;
        } else {
            const double2 c01 = *reinterpret_cast<const double2 *>(cf_p);
            const double2 c23 = *(reinterpret_cast<const double2 *>(cf_p) + 1);
            cApB = c01.x; cAmB = c01.y; cE = c23.x; cG = c23.y;
        }
        struct { T v[W]; } nw;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const T h = cur.v[k];
            const T lf = k == 0 ? cur.eL : cur.v[k - 1];
            const T rt = k == W - 1 ? cur.eR : cur.v[k + 1];
            const T habs = h + href;
            // A (dN + dS) + B (dN - dS) + E (dE + dW) with neighbour-minus-centre differences
            T diff = cApB * (dn.v[k] - h);
            diff = fma(cAmB, up.v[k] - h, diff);
            diff = fma(cE, (rt - h) + (lf - h), diff);
            const T g3 = (cG * habs) * (habs * habs);
            const T rate = fma(dscale, diff, -g3);
            T hn = fma(dt, rate, h);
            hn = tmax<T>(hn, neg_href);      // max(h_abs, 0), drainage.rs:378
            nw.v[k] = habs < crit ? h : hn;  // below-critical cells keep their value, drainage.rs:331-333
        }
        if (has_ring && (grow == ring_top || grow == ring_bot)) {  // ring sums for the poles (pre-TFE)
            double s = 0.0;
            if (active) {
#pragma unroll
                for (int k = 0; k < W; ++k) s += (double)nw.v[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) warp_part[tid >> 5] = s;
            bar_consumers(C);
            if (tid == 0) {
                double t = 0.0;
                for (int i = 0; i < nwarps_c; ++i) t += warp_part[i];
                double *pp = a.pole_partial + (long long)b * 2 * a.pole_slots;
                if (grow == ring_top) pp[blockIdx.x] = t;
                if (grow == ring_bot) pp[a.pole_slots + blockIdx.x] = t;
                __threadfence();  // partial visible before the arrival count
                if (grow == ring_top && a.fold_top) atomicAdd(a.pole_count + 2 * b, 1ull);
                if (grow == ring_bot && a.fold_bottom) atomicAdd(a.pole_count + 2 * b + 1, 1ull);
            }
            bar_consumers(C);
        }
        if constexpr (TFE) {
            const unsigned rm = *rm_p;
            if (rm) {  // uniform over the CTA
#pragma unroll
                for (int k = 0; k < W; ++k)
                    if ((rm * 0x01010101u) & colmask[k / VEC] & (0xffu << (8 * (k % VEC)))) nw.v[k] = fma(-(nw.v[k] + href), a.tfe_rdt, nw.v[k]);
            }
            ++rm_p;
        }
        if (active) {
            if (track) {
#pragma unroll
                for (int k = 0; k < W; ++k) { run_lo = tmin<T>(run_lo, nw.v[k]); run_hi = tmax<T>(run_hi, nw.v[k]); }
            }
            store_h(out_p, nw.v);
            if constexpr (MODE != SBX_RGB_NONE) {
                const float2 sc = *sc_p;
                const float ccy = sc.y * a.rgb.cam_y;
                uint32_t px[W];
#pragma unroll
                for (int k = 0; k < W; ++k)
                    px[k] = shade_px<MODE, T>(nw.v[k], href, href_u, colA[k], sc.x, ccy, Rv, camR2, a.rgb);
                store_px(rgb_p, px);
            }
        }
        out_p += N;
        rgb_p += N;
        cf_p += 4;
        ++sc_p;
        ++grow;
    };

    RowT r0, r1, r2;
    mbar_wait(&full[0], 0);
    load_slot(r0, 0);
    load_slot(r1, 1);
    int left = l1 - l0 + 1;  // rows still to compute
    uint32_t unconsumed = row_bits(r0) | row_bits(r1);  // first release only: nothing of stage 0 has been used yet
    while (true) {
        load_slot(r2, 2);
        // slot 2 was loaded just now; slot 1's edge values (r1.eL/eR) are first used by the row computed next
        release_stage(row_bits(r2) | dep_bits(r1.eL) | dep_bits(r1.eR) | unconsumed);
        unconsumed = 0;
        compute_row(r0, r1, r2);
        if (--left == 0) break;
        mbar_wait(&full[stage], full_parity);
        load_slot(r0, 0);
        compute_row(r1, r2, r0);
        if (--left == 0) break;
        load_slot(r1, 1);
        compute_row(r2, r0, r1);
        if (--left == 0) break;
    }
    // Pole rows (drainage.rs:480-503), folded into this kernel: the strips that updated a pole-adjacent ring wait
    // until all strips of that ring have published their partial sums (they were dispatched together, so this
    // is a short spin among co-resident CTAs), add them in a fixed order and write their own slice of the pole
    // row, its TFE reduction (drainage.rs:437-460) and its colour.
    if (has_ring && (a.fold_top || a.fold_bottom)) {
        const int g_first = a.row_begin + l0 - 1, g_last = a.row_begin + l1 - 1;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const int ring = which == 0 ? ring_top : ring_bot;
            if (!(which == 0 ? a.fold_top : a.fold_bottom) || ring < g_first || ring > g_last) continue;
            const unsigned long long want = a.pole_want[which] * (unsigned long long)gridDim.x;
            if (tid == 0) {
                // bounded like wait_neighbour: if a sibling strip never arrives (a failed launch, or strips that could
                // not be co-resident) the step finishes with a wrong pole row and says so instead of hanging the GPU
                const volatile unsigned long long *cnt = a.pole_count + 2 * b + which;
                const unsigned long long t0 = globaltimer_ns();
                while (*cnt < want) {
                    if (globaltimer_ns() - t0 > 2000000000ull) {  // 2 s
                        atomicOr(a.dev_status, 2u);
                        break;
                    }
                    __nanosleep(32);
                }
                __threadfence();
                const double *pp = a.pole_partial + (long long)b * 2 * a.pole_slots + which * a.pole_slots;
                double t = 0.0;
                for (int i = 0; i < (int)gridDim.x; ++i) t += __ldcg(pp + i);  // fixed order
                warp_part[0] = t / (double)N;
            }
            bar_consumers(C);
            const T v = (T)warp_part[0];
            bar_consumers(C);
            if (active) {
                const int prow = which == 0 ? 0 : a.num_theta - 1;
                const int pl = prow - a.row_begin + 1;
                struct { T v[W]; } pv;
#pragma unroll
                for (int k = 0; k < W; ++k) pv.v[k] = v;
                if constexpr (TFE) {
                    const unsigned prm = front_row_mask(a.fronts, prow) * 0x01010101u;
#pragma unroll
                    for (int k = 0; k < W; ++k)
                        if (prm & colmask[k / VEC] & (0xffu << (8 * (k % VEC)))) pv.v[k] = fma(-(pv.v[k] + href), a.tfe_rdt, pv.v[k]);
                }
                if (track) {
#pragma unroll
                    for (int k = 0; k < W; ++k) { run_lo = tmin<T>(run_lo, pv.v[k]); run_hi = tmax<T>(run_hi, pv.v[k]); }
                }
                store_h(out + (long long)pl * N + jc, pv.v);
                if constexpr (MODE != SBX_RGB_NONE) {
                    const float2 sc = __ldg(a.rowSC + prow);
                    const float ccy = sc.y * a.rgb.cam_y;
                    uint32_t px[W];
#pragma unroll
                    for (int k = 0; k < W; ++k)
                        px[k] = shade_px<MODE, T>(pv.v[k], href, href_u, colA[k], sc.x, ccy, Rv, camR2, a.rgb);
                    store_px(reinterpret_cast<uint32_t *>(a.rgb.out) + (long long)b * a.rgb_bubble_stride +
                                 (long long)(pl - 1) * N + jc, px);
                }
            }
        }
    }

    if (track) {  // one atomicMin / atomicMax per CTA
        using U = typename OrderedWord<T>::type;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            run_lo = tmin<T>(run_lo, __shfl_down_sync(0xffffffffu, run_lo, o));
            run_hi = tmax<T>(run_hi, __shfl_down_sync(0xffffffffu, run_hi, o));
        }
        __shared__ T s_lo[8], s_hi[8];
        if (lane == 0) { s_lo[tid >> 5] = run_lo; s_hi[tid >> 5] = run_hi; }
        bar_consumers(C);
        if (tid == 0) {
            for (int i = 1; i < nwarps_c; ++i) { run_lo = tmin<T>(run_lo, s_lo[i]); run_hi = tmax<T>(run_hi, s_hi[i]); }
            U *slot = static_cast<U *>(a.mm_cur) + 2 * b;
            if (run_lo <= run_hi) {  // at least one finite-or-infinite value was seen
                atomicMin(slot, ordered_bits(run_lo));
                atomicMax(slot + 1, ordered_bits(run_hi));
            }
        }
    }

    if (a.peer.enabled)
        peer_signal_done(a.peer, a.peer.up_row && l0 == 1, a.peer.dn_row && l1 == a.l_last, C, tid);
}

static size_t march_tma_smem_bytes(int packets, int stages)  // 16-byte packets per strip row
{
    return (size_t)stages * kRowsPerStage * (16 * packets + 32) + 2 * kStages * sizeof(uint64_t);
}

static bool use_pdl()
{
    static const bool on = [] { const char *e = getenv("SBX_NO_PDL"); return !(e && atoi(e) != 0); }();
    return on;
}

template <typename T, int MODE, bool TFE, int PK>
static cudaError_t launch_one(const StepArgs<T> &a, dim3 grid, dim3 block, size_t smem, cudaStream_t st)
{
    // Shared-memory carveout, once per device: four 37 KB rings (+ 3 KB of static and reserved shared memory each) need the
    // 164 KB configuration.  Asking for the MAXIMUM (228 KB, round 1) left 28 KB of L1 for the 64 KB LUT texture the colour
    // stage samples; 72 % of the maximum = 164 KB leaves 92 KB.  Same-box A/B (profiles/r02_ab_carveout.txt): 8192 x 16384
    // fused 500.6 / 483.4 -> 503.0 / 504.6 Gcell/s, the L2-resident 2048 x 4096 grid 430.5 / 435.6 -> 464.4 / 465.1 (+8 %).
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        static const int carve_env = [] { const char *e = getenv("SBX_SMEM_CARVEOUT"); return e ? atoi(e) : -1; }();
        const int carve = carve_env >= 0 ? carve_env : (MODE == SBX_RGB_NONE ? (int)cudaSharedmemCarveoutMaxShared : 72);
        cudaError_t e = cudaFuncSetAttribute(k_march_tma<T, MODE, TFE, PK>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // PDL pays when the grid has several waves (measured: +1 % at 8192x16384, -10 % at the single-wave 2048x4096)
    static const size_t pdl_min = [] { const char *e = getenv("SBX_PDL_MIN_CTAS"); return e ? (size_t)atoll(e) : (size_t)500; }();
    cfg.numAttrs = (use_pdl() && (size_t)grid.x * grid.y * grid.z >= pdl_min) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_march_tma<T, MODE, TFE, PK>, a);
}

template <typename T, bool TFE, int PK>
static cudaError_t launch_tma_mode(const StepArgs<T> &a, dim3 grid, dim3 block, size_t smem, cudaStream_t st)
{
    switch (a.rgb.mode) {
    case SBX_RGB_NONE: return launch_one<T, SBX_RGB_NONE, TFE, PK>(a, grid, block, smem, st);
    case SBX_RGB_LUT: return launch_one<T, SBX_RGB_LUT, TFE, PK>(a, grid, block, smem, st);
    default: break;
    }
    if constexpr (PK == 1) {  // the analytic colour modes are arithmetic-bound: one packet per thread only
        switch (a.rgb.mode) {
        case SBX_RGB_EXACT3: return launch_one<T, SBX_RGB_EXACT3, TFE, 1>(a, grid, block, smem, st);
        case SBX_RGB_EXACT3_LINEAR: return launch_one<T, SBX_RGB_EXACT3_LINEAR, TFE, 1>(a, grid, block, smem, st);
        case SBX_RGB_SCHLICK_WALL: return launch_one<T, SBX_RGB_SCHLICK_WALL, TFE, 1>(a, grid, block, smem, st);
        case SBX_RGB_SCHLICK_TWO_BEAM: return launch_one<T, SBX_RGB_SCHLICK_TWO_BEAM, TFE, 1>(a, grid, block, smem, st);
        default: break;
        }
    }
    return cudaErrorInvalidValue;
}

// packets per thread of this launch: 2 for f32 storage with the LUT (or no) colour stage on strips of at least 64 packets
template <typename T>
static int tma_packets_per_thread(const StepArgs<T> &a, int packets)
{
#ifdef SBX_ENABLE_PK2
    static const int env = [] { const char *e = getenv("SBX_PK"); return e ? atoi(e) : 1; }();
    constexpr int VEC = 16 / (int)sizeof(T);
    if (sizeof(T) != 4 || env < 2 || packets < 64 || a.num_phi % (2 * VEC) != 0) return 1;
    return (a.rgb.mode == SBX_RGB_NONE || a.rgb.mode == SBX_RGB_LUT) ? 2 : 1;
#else
    (void)a; (void)packets;
    return 1;
#endif
}

// The fast path handles: non-strict f32/f64 storage, RGBA8 output (or no colour), rows of whole 16-byte packets.
template <typename T>
bool march_tma_supported(const StepArgs<T> &a)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    if (a.num_phi % VEC != 0 || a.num_phi < 32 * VEC) return false;
    if (a.rgb.mode != SBX_RGB_NONE && a.rgb.format != SBX_OUT_RGBA8) return false;
    return true;
}

template <typename T>
static int tma_consumers(int num_phi)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    int consumers = 256;
    while (consumers > 32 && consumers * VEC / 2 >= num_phi) consumers >>= 1;  // narrow grids: fewer, fuller CTAs
    return consumers;
}

// The pole rows are folded into the kernel only when all strips of a ring are certainly co-resident (they
// wait for each other); otherwise k_poles finishes them.
template <typename T>
bool march_tma_folds_poles(const StepArgs<T> &a)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    if (!march_tma_supported<T>(a) || a.l_last < a.l_first) return false;
    // the strips of a ring are dispatched back to back and wait for each other at the end of their band: that
    // terminates as long as they all fit on the device at once, whatever else is resident (SMs x resident CTAs)
    const int strip = tma_consumers<T>(a.num_phi) * VEC;
    const int slots = device_sm_count() * MinBlocks<T>::value;
    return (a.num_phi + strip - 1) / strip <= (slots < 128 ? slots : 128);
}

template <typename T>
cudaError_t launch_march_tma(StepArgs<T> a, int batch, bool tfe, int band_rows_req, cudaStream_t st, int *pole_slots)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int rows = a.l_last - a.l_first + 1;
    const int packets = tma_consumers<T>(a.num_phi);  // 16-byte packets per strip row
    const int strip = packets * VEC;
    const int gx = (a.num_phi + strip - 1) / strip;
    // Ring depth: three stages (9 rows in flight per CTA).  A two-stage ring looked better while the bands were 56 rows high
    // (508 -> 516 Gcell/s); at the 24-row bands many-wave launches use now the deeper ring is as fast for the fused step and
    // faster for the colourless one (528.3 / 819.9 vs 528.9 / 798.0, profiles/r02_ab_band_rows.txt).  SBX_RING_STAGES=2
    // keeps the variant for A/B runs.
    static const int stages_env = [] { const char *e = getenv("SBX_RING_STAGES"); return e ? atoi(e) : 0; }();
    a.ring_stages = stages_env == 2 ? 2 : 3;
    const size_t smem = march_tma_smem_bytes(packets, a.ring_stages);
    const int pk = tma_packets_per_thread<T>(a, packets);
    const int consumers = packets / pk;
    // resident CTAs per SM of this launch shape (narrow strips have smaller CTAs and more of them fit), asked once
    static int occ_cache[2][9] = {{0}};  // by packets per thread, packets / 32
    int &occ = occ_cache[pk - 1][packets / 32];
    if (occ == 0) {
        int nb = 0;
        cudaError_t oe;
#ifdef SBX_ENABLE_PK2
        if constexpr (sizeof(T) == 4) {
            oe = pk == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_march_tma<T, SBX_RGB_LUT, false, 2>, consumers + 32, smem)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_march_tma<T, SBX_RGB_LUT, false, 1>, consumers + 32, smem);
        } else
#endif
        {
            oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_march_tma<T, SBX_RGB_LUT, false, 1>, consumers + 32, smem);
        }
        if (oe != cudaSuccess || nb <= 0) nb = MinBlocks<T>::value;
        occ = nb;
    }
    // Sharded with the in-kernel halo push: short bands next to the neighbours, dispatched first (see StepArgs).  The
    // neighbour's next step waits for this band's flag, and the next launch here cannot start before the band that
    // waited for the neighbour has finished; a short band makes both early instead of last.
    static const int edge_env = [] { const char *e = getenv("SBX_EDGE_ROWS"); return e ? atoi(e) : 8; }();
    static const int edge_min_waves = [] { const char *e = getenv("SBX_EDGE_MIN_WAVES"); return e ? atoi(e) : 3; }();
    a.top_rows = a.bot_rows = a.boundary_first = 0;
    // Only when the launch runs several waves of CTAs: in a single wave every CTA lives for the whole launch anyway and
    // uneven bands only unbalance it (measured at 1024 rows x 16384 per GPU: 40.2 us uniform, 43.1 us with short bands;
    // at 4096 and 8192 rows per GPU the short bands win 0.8 % and 1.2 %).
    const long long ctas_uniform = (long long)gx * ((rows + 23) / 24);
    if (a.peer.enabled && batch == 1 && edge_env > 0 && rows >= 8 * edge_env &&
        ctas_uniform >= (long long)edge_min_waves * device_sm_count() * occ) {
        if (a.peer.up_row) a.top_rows = edge_env;
        if (a.peer.dn_row) a.bot_rows = edge_env;
        a.boundary_first = 1;
    }
    const int n_edge = (a.top_rows > 0) + (a.bot_rows > 0);
    const int rows_mid = rows - a.top_rows - a.bot_rows;
    int slots = device_sm_count() * occ - n_edge * gx;  // the edge bands' CTAs take slots of the first wave
    if (slots < gx) slots = gx;
    int band_rows = band_rows_req > 0 ? band_rows_req : auto_band_rows(rows_mid, (long long)gx * batch, slots, 24);
    if (band_rows > rows_mid) band_rows = rows_mid;
    if (band_rows > kMaxBandRows) band_rows = kMaxBandRows;
    const int gy = (rows_mid + band_rows - 1) / band_rows + n_edge;
    a.band_rows = band_rows;
    a.pole_slots = gx;
    *pole_slots = gx;
    const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)batch), block((unsigned)(consumers + 32));
#ifdef SBX_ENABLE_PK2
    if constexpr (sizeof(T) == 4) {
        if (pk == 2) return tfe ? launch_tma_mode<T, true, 2>(a, grid, block, smem, st) : launch_tma_mode<T, false, 2>(a, grid, block, smem, st);
    }
#endif
    return tfe ? launch_tma_mode<T, true, 1>(a, grid, block, smem, st) : launch_tma_mode<T, false, 1>(a, grid, block, smem, st);
}

template bool march_tma_supported<float>(const StepArgs<float> &);
template bool march_tma_supported<double>(const StepArgs<double> &);
template bool march_tma_folds_poles<float>(const StepArgs<float> &);
template bool march_tma_folds_poles<double>(const StepArgs<double> &);
template cudaError_t launch_march_tma<float>(StepArgs<float>, int, bool, int, cudaStream_t, int *);
template cudaError_t launch_march_tma<double>(StepArgs<double>, int, bool, int, cudaStream_t, int *);

}  // namespace sbx
