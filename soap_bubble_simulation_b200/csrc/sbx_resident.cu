// sbx_resident.cu -- shared-memory-resident multi-step kernels for small grids.
//
// The reference runs its drainage on small grids (config/default.json: 128 x 256 on the CPU,
// src/physics/drainage.rs:264-265; 128 x 64 on the GPU, src/render/pipeline.rs:673-674) and takes many steps
// per frame (`steps_per_frame`, gpu_drainage.rs:337-369).  At those sizes one step of the HBM-streaming kernels is
// pure launch latency (about 12 us for 32 K cells), so `sbx_drainage_step(h, dt, n)` with n >= 2 takes this
// path instead when the grid fits: one thread-block cluster per bubble keeps the whole thickness field in the
// shared memory of its CTAs (a latitude band each, double-buffered), takes all n steps inside one launch and
// touches HBM twice -- load at entry, store at exit.
//
// What bounds a step is the one-way latency between neighbouring SMs (several hundred cycles from the remote store to
// the neighbour's wake-up), so the step is ordered to hide it.  Per step and CTA:
//   1. wait for the neighbours' boundary rows of the previous step (an mbarrier per neighbour and step parity);
//   2. update the band's first and last row and send every new value straight from the register that holds it into
//      the neighbour CTA's ghost slot with st.async -- a remote shared-memory store that counts its bytes on the
//      neighbour's mbarrier, so data and "row is there" signal are the same one-way DSMEM message and there is
//      neither a block barrier before the exchange nor a cluster-wide barrier anywhere in the loop;
//   3. update the other rows while those messages fly: a thread marches a run of rows down one column with the
//      column's three-row window in registers.  The step's TFE fronts (a per-step table the host state machine
//      precomputed, drainage.rs:387-460, prefetched two steps ahead with cp.async) thin the cells they cover on
//      the fly;
//   4. in the two CTAs that own a pole row the last warp is the pole warp: the others update the ring next to the
//      pole early, and the pole warp takes its mean (drainage.rs:480-503; a fixed-order f64 sum, or the reference's
//      left-to-right sum in strict mode), writes the pole row and applies fronts that reach ring or pole, all
//      beside step 3;
//   5. one __syncthreads ends the step.
// The per-cell arithmetic is the same cell_update / cell_update_strict the streaming kernels use, so strict-mode
// results stay bit-identical to the reference and the fast modes agree with the streaming path to the last bits of
// the pole means.  The loop body is kept small on purpose (rolled loops, out-of-line front tests): with one CTA per
// SM the instruction cache, not the issue rate, was the first limit.  WGSL semantics (thickness + concentration,
// frozen poles) has its own kernel below.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <type_traits>

#include "sbx_device.cuh"
#include "sbx_launch.h"
#include "sbx_tma.cuh"

namespace cg = cooperative_groups;

namespace sbx {

constexpr int kResThreads = 512;
constexpr size_t kResSmemMax = 227 * 1024;
constexpr size_t kResFixed = 3 * sizeof(FrontTable) + 160;  // three front tables, barriers, alignment slack

__host__ __device__ inline int res_row0(int r, int C, int NT) { return (int)((long long)r * NT / C); }

template <typename T>
struct alignas(16) ResRow {
    T v[4];  // G, B, E, -
};

__device__ __forceinline__ unsigned char *res_align16(void *p)
{
    return reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
}

// one item of the marching phase: a column and a run of rows marched top to bottom
struct ResItems {
    int L0, L1;  // local rows, inclusive
    int RC, TY;  // rows per run, runs per column
};

__device__ __forceinline__ ResItems res_items(int L0, int L1, int N, int nthreads)
{
    ResItems it;
    it.L0 = L0;
    it.L1 = L1;
    const int nrows = L1 - L0 + 1;
    if (nrows <= 0) { it.RC = 1; it.TY = 0; return it; }
    const int ty = min(max(nthreads / N, 1), nrows);
    it.RC = (nrows + ty - 1) / ty;
    it.TY = (nrows + it.RC - 1) / it.RC;
    return it;
}

// ------------------------------------------------------------------ PTX: cp.async, remote stores that signal an mbarrier
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async(uint32_t raddr, float v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async(uint32_t raddr, double v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}

// neighbour-to-neighbour row exchange state of one CTA: bar[dir][step parity], dir 0 = rows from the CTA above
struct ResLink {
    uint64_t *bar;   // 4 barriers: [0][0], [0][1], [1][0], [1][1]
    int r, C, nsteps;
    uint32_t bytes;  // per neighbour and step

    __device__ __forceinline__ bool pushes(int s) const { return s + 1 < nsteps; }  // the last step's rows are not needed
    // arm the barriers for the rows produced by step s (one arrival carrying the byte count)
    __device__ __forceinline__ void arm(int s) const
    {
        if (!pushes(s)) return;
        if (r > 0) mbar_expect_tx(bar + 0 + (s & 1), bytes);
        if (r < C - 1) mbar_expect_tx(bar + 2 + (s & 1), bytes);
    }
    // wait for the rows produced by step s (all threads), re-arm for step s + 2 (thread 0).  Nothing of step s + 2 can
    // arrive before every thread here has seen step s complete: the neighbour sends it only after it has received
    // this CTA's step s + 1 rows, which are sent after this wait.
    __device__ __forceinline__ void wait(int s) const
    {
        const uint32_t parity = (uint32_t)(s >> 1) & 1u;
        if (r > 0) mbar_wait(bar + 0 + (s & 1), parity);
        if (r < C - 1) mbar_wait(bar + 2 + (s & 1), parity);
        if (threadIdx.x == 0) arm(s + 2);
    }
};

__device__ __forceinline__ void res_link_init(ResLink &lk, uint64_t *bar, int r, int C, int nsteps, uint32_t bytes)
{
    lk.bar = bar; lk.r = r; lk.C = C; lk.nsteps = nsteps; lk.bytes = bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        lk.arm(0);
        lk.arm(1);
    }
}

// sum of one shared-memory row by one warp in a fixed order (four lane-strided partial sums, then a butterfly)
template <typename T>
__device__ __forceinline__ double res_row_sum(const T *row, int N)
{
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int j = threadIdx.x & 31;
#pragma unroll 1
    for (; j + 96 < N; j += 128) {
        v0 += (double)row[j];
        v1 += (double)row[j + 32];
        v2 += (double)row[j + 64];
        v3 += (double)row[j + 96];
    }
#pragma unroll 1
    for (; j < N; j += 32) v0 += (double)row[j];
    double v = (v0 + v1) + (v2 + v3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, bool STRICT>
__device__ __forceinline__ T res_tfe(T v, T href, const StepArgs<T> &a)
{
    if constexpr (STRICT) return (T)__dmul_rn((double)v, a.s_tfe_factor);
    else return fma(-(v + href), a.tfe_rdt, v);  // h *= 1 - r dt, drainage.rs:453-455
}

// Rolled, out-of-line front tests (the unrolled eight-way forms of sbx_device.cuh would triple the loop body)
__device__ __noinline__ unsigned res_row_mask(const FrontTable &ft, int grow)
{
    unsigned m = 0;
    if (grow < ft.row_min || grow > ft.row_max) return 0;
    const int n = ft.count;
#pragma unroll 1
    for (int k = 0; k < n; ++k)
        if (grow >= ft.f[k].row_lo && grow <= ft.f[k].row_hi) m |= 1u << k;
    return m;
}

// fronts whose phi intervals contain column j
__device__ __noinline__ unsigned res_col_mask(const FrontTable &ft, int j)
{
    unsigned m = 0;
    const int n = ft.count;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const FrontCells &f = ft.f[k];
        if ((j >= f.c0_lo && j <= f.c0_hi) || (j >= f.c1_lo && j <= f.c1_hi)) m |= 1u << k;
    }
    return m;
}

#ifdef SBX_RES_PROFILE
#define RES_T(x) long long x; asm volatile("mov.u64 %0, %%clock64;" : "=l"(x)::"memory")
#define RES_ACC(acc, t0, t1) acc += (t1) - (t0)
#else
#define RES_T(x)
#define RES_ACC(acc, t0, t1)
#endif

// Colour epilogue of a resident launch: the frame of the final field is shaded straight from shared memory while the
// band is stored, so `step_fused(dt, n, view)` -- the reference's steps_per_frame loop followed by the draw
// (gpu_drainage.rs:338-370) -- is ONE launch.  Out of line and switched at run time: the step loop's instruction
// footprint (the first limit of this kernel) does not grow.  RGBA8 only; `src` = the band's first owned row.
template <int MODE, typename T>
__device__ __noinline__ void res_shade_band(const T *src, int g0, int Rb, int N, int b, T href, float radius,
                                            const float2 *rowSC, const RgbArgs &rgb, long long rgb_bubble_stride, unsigned magic)
{
    uint32_t *out = reinterpret_cast<uint32_t *>(rgb.out) + (long long)b * rgb_bubble_stride + (long long)g0 * N;
    const float camR2 = rgb.cam2 + radius * radius;
    const float href_u = (float)href * rgb.nm_over_max;
#pragma unroll 1
    for (int i = threadIdx.x; i < Rb * N; i += kResThreads) {
        const int l = (int)__umulhi((unsigned)i, magic);  // i / N
        const int j = i - l * N;
        const float2 sc = __ldg(rowSC + g0 + l);
        out[i] = shade_px<MODE, T>(src[i], href, href_u, __ldg(rgb.colA + j), sc.x, sc.y * rgb.cam_y, radius, camR2, rgb);
    }
}
template <typename T>
__device__ __forceinline__ void res_shade_dispatch(const T *src, int g0, int Rb, int N, int b, T href, float radius,
                                                   const float2 *rowSC, const RgbArgs &rgb, long long stride, unsigned magic)
{
    switch (rgb.mode) {
    case SBX_RGB_LUT: res_shade_band<SBX_RGB_LUT, T>(src, g0, Rb, N, b, href, radius, rowSC, rgb, stride, magic); break;
    case SBX_RGB_EXACT3: res_shade_band<SBX_RGB_EXACT3, T>(src, g0, Rb, N, b, href, radius, rowSC, rgb, stride, magic); break;
    case SBX_RGB_EXACT3_LINEAR: res_shade_band<SBX_RGB_EXACT3_LINEAR, T>(src, g0, Rb, N, b, href, radius, rowSC, rgb, stride, magic); break;
    case SBX_RGB_SCHLICK_WALL: res_shade_band<SBX_RGB_SCHLICK_WALL, T>(src, g0, Rb, N, b, href, radius, rowSC, rgb, stride, magic); break;
    case SBX_RGB_SCHLICK_TWO_BEAM: res_shade_band<SBX_RGB_SCHLICK_TWO_BEAM, T>(src, g0, Rb, N, b, href, radius, rowSC, rgb, stride, magic); break;
    default: break;
    }
}
// which colour stages the epilogue covers (the others keep the streaming fused step for the last step)
bool resident_shades(const RgbArgs &rgb)
{
    return rgb.format == SBX_OUT_RGBA8 && rgb.mode >= SBX_RGB_LUT && rgb.mode <= SBX_RGB_SCHLICK_TWO_BEAM;
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(kResThreads, 1)
k_resident(const StepArgs<T> a, const FrontTable *__restrict__ tables, int nsteps, unsigned magic, int rows_max)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = gridDim.x, r = blockIdx.x, b = blockIdx.y;
    const int N = a.num_phi, NT = a.num_theta, tid = threadIdx.x;
    const int g0 = res_row0(r, C, NT), g1 = res_row0(r + 1, C, NT), Rb = g1 - g0;

    extern __shared__ __align__(16) unsigned char res_smem[];
    const size_t slab = (size_t)(rows_max + 2) * N;  // local row 0 = ghost above, 1..Rb owned, Rb+1 = ghost below
    T *cur = reinterpret_cast<T *>(res_smem);
    T *nxt = cur + slab;
    using RowC = typename std::conditional<STRICT, double4, ResRow<T>>::type;
    RowC *rowc = reinterpret_cast<RowC *>(res_align16(nxt + slab));
    FrontTable *ft = reinterpret_cast<FrontTable *>(rowc + rows_max);  // ring of three: ft[s % 3] = fronts of step s
    unsigned *rowmask = reinterpret_cast<unsigned *>(ft + 3);          // two halves of rows_max: fronts covering each owned row
    unsigned *colmask = rowmask + 2 * rows_max;                        // two halves of N: fronts covering each column
    uint64_t *bars = reinterpret_cast<uint64_t *>(res_align16(colmask + (tables != nullptr ? 2 * N : 0)));

    T *field = const_cast<T *>(a.in) + (long long)b * a.bubble_stride + N;  // global row 0
    const BubbleParam<T> bp = a.bub[b];
    const T href = bp.h_ref;
    const bool tfe = tables != nullptr;
    constexpr int kTabWords = (int)(sizeof(FrontTable) / 4);

    // ---- load the band and its ghost rows, the per-row coefficients and the first two steps' fronts
    ResLink lk;
    res_link_init(lk, bars, r, C, nsteps, (uint32_t)(N * sizeof(T)));
    {
        const int gl = max(g0 - 1, 0), gh = min(g1 + 1, NT);
        const int cnt = (gh - gl) * N;
        const T *src = field + (long long)gl * N;
        T *dst = cur + (size_t)(gl - g0 + 1) * N;
#pragma unroll 1
        for (int i = tid; i < cnt; i += kResThreads) dst[i] = src[i];
#pragma unroll 1
        for (int l = tid; l < Rb; l += kResThreads) {
            const int grow = g0 + l;
            if constexpr (STRICT) rowc[l] = a.coefX[grow];
            else {
                RowC c;
                c.v[0] = a.coefS[4 * grow + 0];  // G
                c.v[1] = a.coefS[4 * grow + 1];  // B
                c.v[2] = a.coefS[4 * grow + 2];  // E
                c.v[3] = T(0);
                rowc[l] = c;
            }
        }
        if (tfe) {
            const int words = kTabWords * min(nsteps, 2);
#pragma unroll 1
            for (int i = tid; i < words; i += kResThreads) reinterpret_cast<int *>(ft)[i] = reinterpret_cast<const int *>(tables)[i];
        }
    }
    __syncthreads();
    if (tfe && ft[0].count > 0) {
#pragma unroll 1
        for (int l = tid; l < Rb; l += kResThreads) rowmask[l] = res_row_mask(ft[0], g0 + l);
#pragma unroll 1
        for (int j = tid; j < N; j += kResThreads) colmask[j] = res_col_mask(ft[0], j);
    }
    cluster.sync();  // barriers initialised and every CTA of the cluster running before any remote store

    const bool pole_cta = r == 0 || r == C - 1;
    const int n_a = pole_cta ? kResThreads - 32 : kResThreads;  // threads that update cells
    const bool pole_warp = pole_cta && tid >= kResThreads - 32;
    const int ring_top = 1 - g0 + 1, ring_bot = NT - 2 - g0 + 1;  // local rows of the pole-adjacent rings (if owned)
    // rows of the marching phase: the band without pole rows, rings (updated early for the pole warp) and the two
    // boundary rows (updated first and sent to the neighbours)
    const int L0 = r == 0 ? 3 : 2, L1 = r == C - 1 ? Rb - 2 : Rb - 1;
    const ResItems it = res_items(L0, L1, N, n_a);
    const int nitems = N * it.TY;
    const int Rb_up = r > 0 ? g0 - res_row0(r - 1, C, NT) : 0;
    // shared::cluster addresses of the neighbours' ghost slots (same buffer layout in every CTA) and barriers
    const uint32_t up_slot0 = r > 0 ? mapa_u32(smem_u32(cur + (size_t)(Rb_up + 1) * N), r - 1) : 0u;
    const uint32_t up_slot1 = r > 0 ? mapa_u32(smem_u32(nxt + (size_t)(Rb_up + 1) * N), r - 1) : 0u;
    const uint32_t dn_slot0 = r < C - 1 ? mapa_u32(smem_u32(cur), r + 1) : 0u;
    const uint32_t dn_slot1 = r < C - 1 ? mapa_u32(smem_u32(nxt), r + 1) : 0u;
    const uint32_t up_bar = r > 0 ? mapa_u32(smem_u32(bars + 2), r - 1) : 0u;      // I am its lower neighbour: its dir 1
    const uint32_t dn_bar = r < C - 1 ? mapa_u32(smem_u32(bars + 0), r + 1) : 0u;  // I am its upper neighbour: its dir 0
    const double inv_n = 1.0 / (double)N;

    auto update = [&](T hc, T up, T dn, T lf, T rt, int l) -> T {
        if constexpr (STRICT) {
            const double4 cx = rowc[l - 1];
            const StrictRow sr{cx.x, cx.y, cx.z};
            return (T)cell_update_strict((double)hc, (double)up, (double)dn, (double)lf, (double)rt, sr, a.s_drain_coef,
                                         a.s_dth2, a.s_two_dth, a.s_dph2, a.s_r2, a.s_D, a.s_dt, a.s_crit);
        } else {
            const RowC c = rowc[l - 1];
            return cell_update<T>(hc, up, dn, lf, rt, c.v[0], a.coefA, c.v[1], c.v[2], bp.diff_scale, href, a.dt, a.crit);
        }
    };

    constexpr int kRunUnroll = STRICT ? 4 : 2;  // the strict update is a chain of IEEE divisions: overlap them
    int i_cur = 0;  // s % 3
    int rm_buf = 0, cm_buf = 0;  // which half of rowmask / colmask holds the current step's masks
#ifdef SBX_RES_PROFILE
    long long p_wait = 0, p_b = 0, p_a = 0, p_sync = 0;
#endif
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        RES_T(c0);
        const int i_next = i_cur == 2 ? 0 : i_cur + 1, i_pre = i_next == 2 ? 0 : i_next + 1;
        const FrontTable &ftc = ft[i_cur];
        const unsigned *rmc = rowmask + rm_buf * rows_max, *cmc = colmask + cm_buf * N;
        int rm_next = rm_buf, cm_next = cm_buf;
        const bool prefetch = tfe && s + 2 < nsteps;
        if (prefetch && tid < kTabWords)
            cp_async4(reinterpret_cast<int *>(ft + i_pre) + tid, reinterpret_cast<const int *>(tables + s + 2) + tid);
        if (s > 0) lk.wait(s - 1);
        RES_T(c1);
        const bool fronts_now = tfe && ftc.count > 0;

        if (pole_warp) {
            // ---- pole rows = mean of the updated neighbouring ring, values before the TFE reduction
            asm volatile("bar.sync 1, %0;" ::"n"(kResThreads) : "memory");  // the rings of this step are in nxt
            const int lane = tid & 31;
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                if ((which == 0 && r != 0) || (which == 1 && r != C - 1)) continue;
                const int lp = which == 0 ? 1 : Rb, lr = which == 0 ? ring_top : ring_bot;
                T *ring = nxt + (size_t)lr * N, *pole = nxt + (size_t)lp * N;
                T avg;
                if constexpr (STRICT) {  // the reference's left-to-right sum (drainage.rs:484-489)
                    double m = 0.0;
                    if (lane == 0) {
                        double sum = 0.0;
#pragma unroll 8
                        for (int j = 0; j < N; ++j) sum = __dadd_rn(sum, (double)ring[j]);  // loads run ahead of the add chain
                        m = __ddiv_rn(sum, (double)N);
                    }
                    avg = (T)__shfl_sync(0xffffffffu, m, 0);
                } else {
                    avg = (T)(res_row_sum(ring, N) * inv_n);  // 1 ulp of f64 from the quotient, far inside tolerance
                }
                const unsigned rm_pole = fronts_now ? rmc[lp - 1] : 0u, rm_ring = fronts_now ? rmc[lr - 1] : 0u;
                const T avg_thin = res_tfe<T, STRICT>(avg, href, a);
#pragma unroll 4
                for (int j = lane; j < N; j += 32) {
                    const unsigned cm = (rm_pole | rm_ring) != 0u ? cmc[j] : 0u;
                    pole[j] = (rm_pole & cm) != 0u ? avg_thin : avg;
                    if ((rm_ring & cm) != 0u) ring[j] = res_tfe<T, STRICT>(ring[j], href, a);
                }
            }
        } else {
            // ---- pole CTAs: the ring next to the pole before anything else, so that the pole warp can start (no fronts
            // yet: it applies them after taking the mean)
            if (pole_cta) {
#pragma unroll 1
                for (int which = 0; which < 2; ++which) {
                    if ((which == 0 && r != 0) || (which == 1 && r != C - 1)) continue;
                    const int l = which == 0 ? ring_top : ring_bot;
                    const T *row = cur + (size_t)l * N;
                    T *orow = nxt + (size_t)l * N;
#pragma unroll 1
                    for (int j = tid; j < N; j += n_a) {
                        const int jm = j == 0 ? N - 1 : j - 1, jp = j == N - 1 ? 0 : j + 1;
                        orow[j] = update(row[j], row[j - N], row[j + N], row[jm], row[jp], l);
                    }
                }
                asm volatile("bar.arrive 1, %0;" ::"n"(kResThreads) : "memory");
            }
            // ---- boundary rows first: every new value goes to the neighbour's ghost slot from its register
            const bool push = lk.pushes(s);
            const int o = (s + 1) & 1;  // nxt is buffer 1 in even steps
            if (r > 0 && r < C - 1) {  // both neighbours: the two cells of a column are independent work for one thread
                const uint32_t urow = o ? up_slot1 : up_slot0, drow = o ? dn_slot1 : dn_slot0;
                const uint32_t ubar = up_bar + 8u * (s & 1), dbar = dn_bar + 8u * (s & 1);
                const unsigned rm1 = fronts_now ? rmc[0] : 0u, rm2 = fronts_now ? rmc[Rb - 1] : 0u;
                const T *row1 = cur + (size_t)N, *row2 = cur + (size_t)Rb * N;
                T *orow1 = nxt + (size_t)N, *orow2 = nxt + (size_t)Rb * N;
#pragma unroll 1
                for (int j = tid; j < N; j += n_a) {
                    const int jm = j == 0 ? N - 1 : j - 1, jp = j == N - 1 ? 0 : j + 1;
                    T n1 = update(row1[j], row1[j - N], row1[j + N], row1[jm], row1[jp], 1);
                    T n2 = update(row2[j], row2[j - N], row2[j + N], row2[jm], row2[jp], Rb);
                    if ((rm1 | rm2) != 0u) {
                        const unsigned cm = cmc[j];
                        if ((rm1 & cm) != 0u) n1 = res_tfe<T, STRICT>(n1, href, a);
                        if ((rm2 & cm) != 0u) n2 = res_tfe<T, STRICT>(n2, href, a);
                    }
                    orow1[j] = n1;
                    orow2[j] = n2;
                    if (push) {
                        st_async(urow + (uint32_t)sizeof(T) * j, n1, ubar);
                        st_async(drow + (uint32_t)sizeof(T) * j, n2, dbar);
                    }
                }
            } else if (C > 1) {
                const int side = r == 0 ? 1 : 0;
                const int l = side == 0 ? 1 : Rb;
                const uint32_t rrow = side == 0 ? (o ? up_slot1 : up_slot0) : (o ? dn_slot1 : dn_slot0);
                const uint32_t rbar = (side == 0 ? up_bar : dn_bar) + 8u * (s & 1);
                const unsigned rm = fronts_now ? rmc[l - 1] : 0u;
                const T *row = cur + (size_t)l * N;
                T *orow = nxt + (size_t)l * N;
#pragma unroll 1
                for (int j = tid; j < N; j += n_a) {
                    const int jm = j == 0 ? N - 1 : j - 1, jp = j == N - 1 ? 0 : j + 1;
                    T nw = update(row[j], row[j - N], row[j + N], row[jm], row[jp], l);
                    if (rm != 0u && (rm & cmc[j]) != 0u) nw = res_tfe<T, STRICT>(nw, href, a);
                    orow[j] = nw;
                    if (push) st_async(rrow + (uint32_t)sizeof(T) * j, nw, rbar);
                }
            }
            RES_T(c2);
            // ---- the other rows, a run of rows per thread, fronts applied on the fly
#pragma unroll 1
            for (int w = tid; w < nitems; w += n_a) {
                const int ty = (int)__umulhi((unsigned)w, magic);  // w / N, exact for w * N < 2^32
                const int j = w - ty * N;
                const int jm = j == 0 ? N - 1 : j - 1, jp = j == N - 1 ? 0 : j + 1;
                const int la = it.L0 + ty * it.RC, lb = min(la + it.RC - 1, it.L1);
                const unsigned cm = fronts_now ? cmc[j] : 0u;
                const T *__restrict__ row = cur + (size_t)la * N;
                T *__restrict__ orow = nxt + (size_t)la * N;
                T up = row[j - N], hc = row[j];
#pragma unroll kRunUnroll
                for (int l = la; l <= lb; ++l, row += N, orow += N) {
                    const T dn = row[j + N], lf = row[jm], rt = row[jp];
                    T nw = update(hc, up, dn, lf, rt, l);
                    if (cm != 0u && (rmc[l - 1] & cm) != 0u) nw = res_tfe<T, STRICT>(nw, href, a);
                    orow[j] = nw;
                    up = hc;
                    hc = dn;
                }
            }
            if (tfe && s + 1 < nsteps && ft[i_next].count > 0) {
                // row and column masks of the next step's fronts (its table landed during the previous step), rebuilt
                // only when the host flagged a change, into the half nobody reads during this step
                if (ft[i_next].rows_changed) {
                    unsigned *rmn = rowmask + (rm_buf ^ 1) * rows_max;
#pragma unroll 1
                    for (int l = tid; l < Rb; l += n_a) rmn[l] = res_row_mask(ft[i_next], g0 + l);
                }
                if (ft[i_next].cols_changed) {
                    unsigned *cmn = colmask + (cm_buf ^ 1) * N;
#pragma unroll 1
                    for (int j = tid; j < N; j += n_a) cmn[j] = res_col_mask(ft[i_next], j);
                }
            }
            if (prefetch) cp_async_wait_all();
            RES_T(c3);
            RES_ACC(p_b, c1, c2); RES_ACC(p_a, c2, c3);
        }
        if (tfe && s + 1 < nsteps && ft[i_next].count > 0) {  // uniform: every thread flips together
            if (ft[i_next].rows_changed) rm_next = rm_buf ^ 1;
            if (ft[i_next].cols_changed) cm_next = cm_buf ^ 1;
        }
        RES_T(c4);
        __syncthreads();
        rm_buf = rm_next;
        cm_buf = cm_next;
        T *t = cur; cur = nxt; nxt = t;
        i_cur = i_next;
        RES_T(c5);
        RES_ACC(p_wait, c0, c1); RES_ACC(p_sync, c4, c5);
    }
#ifdef SBX_RES_PROFILE
    if ((tid == 0 || tid == 511) && b == 0 && (r == 0 || r == C / 2))
        printf("[res] cta %d/%d tid %d steps %d: wait %lld  boundary %lld  rest %lld  sync %lld  cycles/step\n", r, C, tid, nsteps,
               p_wait / nsteps, p_b / nsteps, p_a / nsteps, p_sync / nsteps);
#endif

    // ---- store the band
    {
        T *dst = field + (long long)g0 * N;
        const T *src = cur + (size_t)N;
#pragma unroll 1
        for (int i = tid; i < Rb * N; i += kResThreads) dst[i] = src[i];
        if (a.rgb.mode != SBX_RGB_NONE)
            res_shade_dispatch<T>(src, g0, Rb, N, b, href, bp.radius, a.rowSC, a.rgb, a.rgb_bubble_stride, magic);
    }
    cluster.sync();  // nobody leaves while a neighbour could still address its shared memory
}

// ------------------------------------------------------------------ WGSL semantics (drainage.wgsl:60-175)
// thickness and concentration both resident; pole rows are never written (drainage.wgsl:96-105).
struct WgslConsts {
    float inv_dth2, inv_2dth, inv_dph2, inv_2dph, r2, inv_r2, inv_r, drainage_coeff;
    float diffusion_coeff, marangoni_coeff, gamma_reduction, surfactant_diffusion, dt, critical;
    bool mar;
};

__device__ __forceinline__ void wgsl_cell(const WgslConsts &k, const float4 rw, float h, float h_tm, float h_tp, float h_pm,
                                          float h_pp, float conc, float c_tm, float c_tp, float c_pm, float c_pp,
                                          float &out_h, float &out_c)
{
    const float sin_theta = rw.x, cot = rw.y, inv_sin2 = rw.z, inv_sin = rw.w;
    const float drainage_term = -k.drainage_coeff * (h * h * h) * sin_theta;
    const float d2h_dtheta2 = (h_tp - 2.0f * h + h_tm) * k.inv_dth2;
    const float dh_dtheta = (h_tp - h_tm) * k.inv_2dth;
    const float d2h_dphi2 = (h_pp - 2.0f * h + h_pm) * k.inv_dph2;
    const float h_lap = (d2h_dtheta2 + cot * dh_dtheta + d2h_dphi2 * inv_sin2) * k.inv_r2;
    const float diffusion_term = k.diffusion_coeff * h_lap;
    float marangoni_term = 0.f;
    const float dc_dtheta = (c_tp - c_tm) * k.inv_2dth;
    if (k.mar) {
        const float dc_dphi = (c_pp - c_pm) * k.inv_2dph;
        const float gt = dc_dtheta * k.inv_r;
        const float gp = dc_dphi * k.inv_r * inv_sin;
        const float grad_c_mag = sqrtf(gt * gt + gp * gp);
        marangoni_term = k.marangoni_coeff * (k.gamma_reduction * grad_c_mag) * h * h;
    }
    const float new_h = fmaxf(h + k.dt * (drainage_term + diffusion_term + marangoni_term), 0.f);
    const float d2c_dtheta2 = (c_tp - 2.0f * conc + c_tm) * k.inv_dth2;
    const float d2c_dphi2 = (c_pp - 2.0f * conc + c_pm) * k.inv_dph2;
    const float c_lap = (d2c_dtheta2 + cot * dc_dtheta + d2c_dphi2 * inv_sin2) * k.inv_r2;
    float dc_dt = k.surfactant_diffusion * c_lap;
    if (k.mar) dc_dt += 0.001f * (h_lap * k.r2) * conc;
    const float new_c = fminf(fmaxf(conc + k.dt * dc_dt, 0.f), 1.f);
    const bool skip = h < k.critical;  // drainage.wgsl:108-112
    out_h = skip ? h : new_h;
    out_c = skip ? conc : new_c;
}

__global__ void __launch_bounds__(kResThreads, 1)
k_resident_wgsl(const WgslArgs a, int nsteps, unsigned magic, int rows_max)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = gridDim.x, r = blockIdx.x, b = blockIdx.y;
    const int N = a.num_phi, NT = a.num_theta, tid = threadIdx.x;
    const int g0 = res_row0(r, C, NT), g1 = res_row0(r + 1, C, NT), Rb = g1 - g0;

    extern __shared__ __align__(16) unsigned char res_smem[];
    const size_t slab = (size_t)(rows_max + 2) * N;
    float *hcur = reinterpret_cast<float *>(res_smem), *hnxt = hcur + slab;
    float *ccur = hnxt + slab, *cnxt = ccur + slab;
    float4 *roww = reinterpret_cast<float4 *>(res_align16(cnxt + slab));
    uint64_t *bars = reinterpret_cast<uint64_t *>(roww + rows_max);

    float *hfield = const_cast<float *>(a.h_in) + (long long)b * a.bubble_stride + N;
    float *cfield = const_cast<float *>(a.c_in) + (long long)b * a.bubble_stride + N;
    ResLink lk;
    res_link_init(lk, bars, r, C, nsteps, (uint32_t)(2 * N * sizeof(float)));
    {
        const int gl = max(g0 - 1, 0), gh = min(g1 + 1, NT);
        const int cnt = (gh - gl) * N;
        const size_t so = (size_t)gl * N, doff = (size_t)(gl - g0 + 1) * N;
#pragma unroll 1
        for (int i = tid; i < cnt; i += kResThreads) {
            const float hv = hfield[so + i], cv = cfield[so + i];
            hcur[doff + i] = hv; hnxt[doff + i] = hv;  // pole rows are frozen: keep them in both halves
            ccur[doff + i] = cv; cnxt[doff + i] = cv;
        }
#pragma unroll 1
        for (int l = tid; l < Rb; l += kResThreads) roww[l] = reinterpret_cast<const float4 *>(a.rowW)[g0 + l];
    }
    __syncthreads();
    cluster.sync();

    WgslConsts k;
    {
        const sbx_drainage_params &p = a.p;
        k.inv_dth2 = 1.f / (a.delta_theta * a.delta_theta);
        k.inv_2dth = 1.f / (2.f * a.delta_theta);
        k.inv_dph2 = 1.f / (a.delta_phi * a.delta_phi);
        k.inv_2dph = 1.f / (2.f * a.delta_phi);
        k.r2 = p.bubble_radius * p.bubble_radius;
        k.inv_r2 = 1.f / k.r2;
        k.inv_r = 1.f / p.bubble_radius;
        k.drainage_coeff = p.density * p.gravity / (3.0f * p.viscosity);
        k.diffusion_coeff = p.diffusion_coeff;
        k.marangoni_coeff = p.marangoni_coeff;
        k.gamma_reduction = p.gamma_reduction;
        k.surfactant_diffusion = p.surfactant_diffusion;
        k.dt = p.dt;
        k.critical = p.critical_thickness;
        k.mar = p.marangoni_enabled != 0u;
    }

    // marching rows: the band without the frozen pole rows and the two boundary rows
    const int L0 = 2, L1 = Rb - 1;
    const ResItems it = res_items(L0, L1, N, kResThreads);
    const int nitems = N * it.TY;
    const int Rb_up = r > 0 ? g0 - res_row0(r - 1, C, NT) : 0;
    const size_t up_off = (size_t)(Rb_up + 1) * N;
    const uint32_t up_h0 = r > 0 ? mapa_u32(smem_u32(hcur + up_off), r - 1) : 0u, up_h1 = r > 0 ? mapa_u32(smem_u32(hnxt + up_off), r - 1) : 0u;
    const uint32_t up_c0 = r > 0 ? mapa_u32(smem_u32(ccur + up_off), r - 1) : 0u, up_c1 = r > 0 ? mapa_u32(smem_u32(cnxt + up_off), r - 1) : 0u;
    const uint32_t dn_h0 = r < C - 1 ? mapa_u32(smem_u32(hcur), r + 1) : 0u, dn_h1 = r < C - 1 ? mapa_u32(smem_u32(hnxt), r + 1) : 0u;
    const uint32_t dn_c0 = r < C - 1 ? mapa_u32(smem_u32(ccur), r + 1) : 0u, dn_c1 = r < C - 1 ? mapa_u32(smem_u32(cnxt), r + 1) : 0u;
    const uint32_t up_bar = r > 0 ? mapa_u32(smem_u32(bars + 2), r - 1) : 0u;
    const uint32_t dn_bar = r < C - 1 ? mapa_u32(smem_u32(bars + 0), r + 1) : 0u;

#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        if (s > 0) lk.wait(s - 1);
        const bool push = lk.pushes(s);
        const int o = (s + 1) & 1;
        // boundary rows first (row 1 of the top band and row Rb of the bottom band are the frozen poles)
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int l = side == 0 ? 1 : Rb;
            if ((side == 0 && r == 0) || (side == 1 && r == C - 1) || (side == 1 && Rb == 1)) continue;
            const bool send = push && (side == 0 ? r > 0 : r < C - 1);
            const uint32_t rh = side == 0 ? (o ? up_h1 : up_h0) : (o ? dn_h1 : dn_h0);
            const uint32_t rc = side == 0 ? (o ? up_c1 : up_c0) : (o ? dn_c1 : dn_c0);
            const uint32_t rbar = (side == 0 ? up_bar : dn_bar) + 8u * (s & 1);
            const float *hr = hcur + (size_t)l * N, *cr = ccur + (size_t)l * N;
            const float4 rw = roww[l - 1];
#pragma unroll 1
            for (int j = tid; j < N; j += kResThreads) {
                const int jm = j == 0 ? N - 1 : j - 1, jp = j == N - 1 ? 0 : j + 1;
                float nh, nc;
                wgsl_cell(k, rw, hr[j], hr[j - N], hr[j + N], hr[jm], hr[jp], cr[j], cr[j - N], cr[j + N], cr[jm], cr[jp], nh, nc);
                hnxt[(size_t)l * N + j] = nh;
                cnxt[(size_t)l * N + j] = nc;
                if (send) {
                    st_async(rh + 4u * j, nh, rbar);
                    st_async(rc + 4u * j, nc, rbar);
                }
            }
        }
#pragma unroll 1
        for (int w = tid; w < nitems; w += kResThreads) {
            const int ty = (int)__umulhi((unsigned)w, magic);
            const int j = w - ty * N;
            const int jm = j == 0 ? N - 1 : j - 1, jp = j == N - 1 ? 0 : j + 1;
            const int la = it.L0 + ty * it.RC, lb = min(la + it.RC - 1, it.L1);
            const float *__restrict__ hr = hcur + (size_t)la * N, *__restrict__ cr = ccur + (size_t)la * N;
            float *__restrict__ ho = hnxt + (size_t)la * N, *__restrict__ co = cnxt + (size_t)la * N;
            float h_tm = hr[j - N], h = hr[j], c_tm = cr[j - N], conc = cr[j];
#pragma unroll 1
            for (int l = la; l <= lb; ++l, hr += N, cr += N, ho += N, co += N) {
                const float h_tp = hr[j + N], c_tp = cr[j + N];
                float nh, nc;
                wgsl_cell(k, roww[l - 1], h, h_tm, h_tp, hr[jm], hr[jp], conc, c_tm, c_tp, cr[jm], cr[jp], nh, nc);
                ho[j] = nh;
                co[j] = nc;
                h_tm = h; h = h_tp;
                c_tm = conc; conc = c_tp;
            }
        }
        __syncthreads();
        float *t = hcur; hcur = hnxt; hnxt = t;
        t = ccur; ccur = cnxt; cnxt = t;
    }
    {
        const size_t go = (size_t)g0 * N;
#pragma unroll 1
        for (int i = tid; i < Rb * N; i += kResThreads) {
            hfield[go + i] = hcur[N + i];
            cfield[go + i] = ccur[N + i];
        }
        if (a.rgb.mode != SBX_RGB_NONE)
            res_shade_dispatch<float>(hcur + N, g0, Rb, N, b, 0.f, a.radius_view, a.rowSC, a.rgb, a.rgb_bubble_stride, magic);
    }
    cluster.sync();
}

// ------------------------------------------------------------------ host side
static void cluster_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int C, int batch, size_t smem, cudaStream_t st)
{
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)C, (unsigned)batch, 1);
    cfg.blockDim = dim3(kResThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

__global__ void __launch_bounds__(kResThreads, 1) k_resident_probe() {}

// widest cluster this device schedules with one 227 KB CTA per SM: 16 (non-portable size, opt-in) or 8.
// SBX_RESIDENT_CLUSTER caps it (A/B measurements).
static int max_cluster_width()
{
    static int cached = 0;
    if (cached) return cached;
    int cap = 16;
    if (const char *e = getenv("SBX_RESIDENT_CLUSTER")) cap = std::max(1, atoi(e));
    int width = std::min(cap, 8);
    if (cap >= 16) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr[1];
        int n = 0;
        if (cudaFuncSetAttribute(k_resident_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
            cudaFuncSetAttribute(k_resident_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResSmemMax) == cudaSuccess) {
            cluster_config(cfg, attr, 16, 1, kResSmemMax, nullptr);
            if (cudaOccupancyMaxActiveClusters(&n, k_resident_probe, &cfg) == cudaSuccess && n >= 1) width = 16;
        }
        cudaGetLastError();
    }
    cached = width;
    return cached;
}

static int pick_cluster(int NT, int batch, size_t bytes_per_row, size_t extra_per_row, size_t fixed, int *rows_max_out,
                        size_t *smem_out)
{
    // widest cluster (most SMs per bubble) whose bands fit in shared memory and whose CTAs are all co-resident
    for (int C = max_cluster_width(); C >= 1; C >>= 1) {
        if (C > 1 && NT / C < 3) continue;  // every band keeps its pushed rows apart from the pole-adjacent ring
        if ((long long)batch * C > 128 && C > 1) continue;  // past one wave the streaming kernels win
        int rows_max = 0;
        for (int r = 0; r < C; ++r) rows_max = std::max(rows_max, res_row0(r + 1, C, NT) - res_row0(r, C, NT));
        const size_t smem = (size_t)(rows_max + 2) * bytes_per_row + (size_t)rows_max * extra_per_row + fixed;
        if (smem > kResSmemMax) continue;
        *rows_max_out = rows_max;
        *smem_out = smem;
        return C;
    }
    return 0;
}

template <typename T>
static size_t res_bytes_per_row(int N) { return 2 * (size_t)N * sizeof(T); }

template <typename T>
int resident_cluster(const StepArgs<T> &a, int batch, bool strict)
{
    if (a.row_begin != 0 || a.l_last - a.l_first + 1 != a.num_theta - 2) return 0;  // whole sphere only
    if (a.num_theta < 4 || a.num_phi < 2 || batch > 128) return 0;
    int rows_max;
    size_t smem;
    return pick_cluster(a.num_theta, batch, res_bytes_per_row<T>(a.num_phi), (strict ? sizeof(double4) : 4 * sizeof(T)) + 8,
                        kResFixed + 8 * (size_t)a.num_phi, &rows_max, &smem);
}
template int resident_cluster<float>(const StepArgs<float> &, int, bool);
template int resident_cluster<double>(const StepArgs<double> &, int, bool);

template <typename K, typename... Args>
static cudaError_t launch_cluster(K kernel, int C, int batch, size_t smem, cudaStream_t st, Args... args)
{
    // the attributes are sticky per kernel and device: set them when they change, not on every launch (a frame loop
    // launches this kernel thousands of times with the same shape)
    static thread_local struct { const void *k; int dev; size_t smem; int c; } last = {nullptr, -1, 0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaSuccess;
    if (last.k != (const void *)kernel || last.dev != dev || last.smem != smem || last.c != C) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (C > 8) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
        last = {(const void *)kernel, dev, smem, C};
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config(cfg, attr, C, batch, smem, st);
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <typename T>
cudaError_t launch_resident(StepArgs<T> a, int batch, bool strict, const FrontTable *tables_dev, int nsteps,
                            cudaStream_t st)
{
    int rows_max = 0;
    size_t smem = 0;
    const int C = pick_cluster(a.num_theta, batch, res_bytes_per_row<T>(a.num_phi),
                               (strict ? sizeof(double4) : 4 * sizeof(T)) + 8, kResFixed + 8 * (size_t)a.num_phi, &rows_max, &smem);
    if (C == 0) return cudaErrorInvalidConfiguration;
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned long long)a.num_phi - 1) / (unsigned long long)a.num_phi);
    if (strict) {
        if constexpr (sizeof(T) == 8) return launch_cluster(k_resident<T, true>, C, batch, smem, st, a, tables_dev, nsteps, magic, rows_max);
        return cudaErrorInvalidConfiguration;
    }
    return launch_cluster(k_resident<T, false>, C, batch, smem, st, a, tables_dev, nsteps, magic, rows_max);
}
template cudaError_t launch_resident<float>(StepArgs<float>, int, bool, const FrontTable *, int, cudaStream_t);
template cudaError_t launch_resident<double>(StepArgs<double>, int, bool, const FrontTable *, int, cudaStream_t);

int resident_cluster_wgsl(const WgslArgs &a, int batch)
{
    if (a.row_begin != 0 || a.rows_local != a.num_theta || a.num_theta < 4 || a.num_phi < 2 || batch > 128) return 0;
    int rows_max;
    size_t smem;
    return pick_cluster(a.num_theta, batch, 4 * (size_t)a.num_phi * sizeof(float), sizeof(float4), 96, &rows_max, &smem);
}

cudaError_t launch_resident_wgsl(WgslArgs a, int batch, int nsteps, cudaStream_t st)
{
    int rows_max = 0;
    size_t smem = 0;
    const int C = pick_cluster(a.num_theta, batch, 4 * (size_t)a.num_phi * sizeof(float), sizeof(float4), 96, &rows_max, &smem);
    if (C == 0) return cudaErrorInvalidConfiguration;
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned long long)a.num_phi - 1) / (unsigned long long)a.num_phi);
    return launch_cluster(k_resident_wgsl, C, batch, smem, st, a, nsteps, magic, rows_max);
}

}  // namespace sbx
