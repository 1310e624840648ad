// sbx_wgsl_tma.cu -- TMA-staged ring variant of the WGSL-semantics dispatch (drainage.wgsl::drainage_step,
// src/render/shaders/drainage.wgsl:72-202): thickness + surfactant concentration, Marangoni coupling, frozen poles.
//
// Same skeleton as k_march_tma (sbx_march_tma.cu): a producer warp streams strip rows of BOTH fields into a
// shared-memory ring with cp.async.bulk + mbarriers, consumers keep two three-row register windows and write the
// two updated fields (and optionally RGBA8) with 16-byte stores: 16 B/cell (+4) of HBM traffic and no global-load
// instructions, shuffles or predicated edge loads in the instruction stream.  Arithmetic is the one of k_wgsl.
#include <cstdlib>

#include "sbx_device.cuh"
#include "sbx_launch.h"
#include "sbx_tma.cuh"

namespace sbx {

namespace {

constexpr int kWStages = 2;        // two fields double the bytes per stage, so two stages keep 4 CTAs per SM
constexpr int kWRowsPerStage = 3;  // == unroll of the register window
constexpr int kWHalo = 4;
constexpr int kWMaxBandRows = 64;  // rows a CTA marches at most (their coefficients are staged in shared memory)

struct RowW {
    float v[4];
    float eL, eR;
};

}  // namespace

template <int MODE>
__global__ void __launch_bounds__(288, 3) k_wgsl_tma(const WgslArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // per-row {sin, cot, 1/sin^2, 1/sin} and {sin, cos} of this band, staged once (see sbx_march_tma.cu)
    __shared__ float4 s_rw[kWMaxBandRows];
    __shared__ float2 s_sc[kWMaxBandRows];
    const int C = blockDim.x - 32;
    const int N = a.num_phi;
    const int strip_w_max = 4 * C;
    const int c0 = blockIdx.x * strip_w_max;
    const int w = min(strip_w_max, N - c0);
    const int rowf = strip_w_max + 2 * kWHalo;  // floats per field row in the ring
    const int slotf = 2 * rowf;                 // a ring slot holds the thickness row then the concentration row
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kWStages * kWRowsPerStage * slotf * sizeof(float));
    uint64_t *empty = full + kWStages;

    const int b = blockIdx.z;
    const float *hin = a.h_in + (long long)b * a.bubble_stride;
    const float *cin = a.c_in + (long long)b * a.bubble_stride;
    const int l0 = a.l_first + blockIdx.y * a.band_rows;
    const int l1 = min(l0 + a.band_rows - 1, a.l_last);
    if (l0 > l1) return;
    const int rows_load = (l1 - l0 + 1) + 2;
    const int nchunks = (rows_load + kWRowsPerStage - 1) / kWRowsPerStage;
    const int nwarps_c = C >> 5;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarps_c);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (threadIdx.x >= C) {
        if (threadIdx.x == C) {  // producer: one lane issues all bulk copies
            const bool wrapL = c0 == 0, wrapR = c0 + w >= N;
            const uint32_t main_floats = (uint32_t)w + (wrapL ? 0 : kWHalo) + (wrapR ? 0 : kWHalo);
            const uint32_t row_bytes = 2u * ((uint32_t)w + 2 * kWHalo) * 4u;  // both fields
            const bool keep_edges = a.l2_keep_edges != 0;  // fields larger than L2, see k_march_tma
            const uint64_t pol = keep_edges ? l2_policy_evict_last() : 0ull;
            const bool touch_top = a.peer.enabled && a.peer.up_row && l0 == 1;
            const bool touch_bot = a.peer.enabled && a.peer.dn_row && l1 == a.l_last;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kWStages;
                if (c >= kWStages) mbar_wait(&empty[s], (uint32_t)(((c / kWStages) - 1) & 1));
                const int r_begin = c * kWRowsPerStage;
                const int r_end = min(r_begin + kWRowsPerStage, rows_load);
                mbar_expect_tx(&full[s], row_bytes * (uint32_t)(r_end - r_begin));
                for (int r = r_begin; r < r_end; ++r) {
                    // a band next to a neighbour reads that row of both fields out of the neighbour's buffers (peer memory),
                    // once the neighbour's band has finished the previous step (see PeerArgs)
                    const float *remote_h = nullptr, *remote_c = nullptr;
                    if (r == 0 && touch_top) {
                        wait_neighbour(a.peer.flag_up, a.peer.epoch - 1, a.peer.status);
                        remote_h = reinterpret_cast<const float *>(a.peer.up_row);
                        remote_c = reinterpret_cast<const float *>(a.peer.up_row2);
                    }
                    if (r == rows_load - 1 && touch_bot) {
                        wait_neighbour(a.peer.flag_dn, a.peer.epoch - 1, a.peer.status);
                        remote_h = reinterpret_cast<const float *>(a.peer.dn_row);
                        remote_c = reinterpret_cast<const float *>(a.peer.dn_row2);
                    }
                    float *slot = ring + ((size_t)s * kWRowsPerStage + (r - r_begin)) * slotf;
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const float *grow_ptr = remote_h ? (f == 0 ? remote_h : remote_c) : (f == 0 ? hin : cin) + (long long)(l0 - 1 + r) * N;
                        float *dst = slot + f * rowf;
                        if (keep_edges && !remote_h && (r <= 1 || r >= rows_load - 2))  // rows the neighbouring band reads again
                            tma_bulk_g2s_hint(dst + (wrapL ? kWHalo : 0), grow_ptr + c0 - (wrapL ? 0 : kWHalo), main_floats * 4u, &full[s], pol);
                        else
                            tma_bulk_g2s(dst + (wrapL ? kWHalo : 0), grow_ptr + c0 - (wrapL ? 0 : kWHalo), main_floats * 4u, &full[s]);
                        if (wrapL) tma_bulk_g2s(dst, grow_ptr + (N - kWHalo), kWHalo * 4u, &full[s]);
                        if (wrapR) tma_bulk_g2s(dst + kWHalo + w, grow_ptr, kWHalo * 4u, &full[s]);
                    }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int r = tid; r < l1 - l0 + 1; r += C) {
        s_rw[r] = __ldg(reinterpret_cast<const float4 *>(a.rowW) + (a.row_begin + l0 - 1 + r));
        s_sc[r] = __ldg(a.rowSC + (a.row_begin + l0 - 1 + r));
    }
    asm volatile("bar.sync 1, %0;" ::"r"(C) : "memory");
    const int jloc = tid * 4;
    const bool active = jloc < w;
    const int jc = c0 + (active ? jloc : 0);
    const int soff = kWHalo + (active ? jloc : 0);
    float *hout = a.h_out + (long long)b * a.bubble_stride;
    float *cout = a.c_out + (long long)b * a.bubble_stride;

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
    const float dt = p.dt, crit = p.critical_thickness;

    float colA[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (MODE != SBX_RGB_NONE) {
        const float4 ca = __ldg(reinterpret_cast<const float4 *>(a.rgb.colA + jc));
        colA[0] = ca.x; colA[1] = ca.y; colA[2] = ca.z; colA[3] = ca.w;
    }
    const float Rv = a.radius_view;
    const float camR2 = a.rgb.cam2 + Rv * Rv;

    int stage = 0;
    uint32_t full_parity = 0;
    const float *ring_t = ring + soff;
    const uint32_t stage_floats = kWRowsPerStage * slotf;
    const float *stage_p = ring_t;

    auto load_slot = [&](RowW &rh, RowW &rc, int slot) {
        const float *src = stage_p + slot * slotf;
        const float4 q = *reinterpret_cast<const float4 *>(src);
        rh.v[0] = q.x; rh.v[1] = q.y; rh.v[2] = q.z; rh.v[3] = q.w;
        rh.eL = src[-1];
        rh.eR = src[4];
        const float *src2 = src + rowf;
        const float4 q2 = *reinterpret_cast<const float4 *>(src2);
        rc.v[0] = q2.x; rc.v[1] = q2.y; rc.v[2] = q2.z; rc.v[3] = q2.w;
        rc.eL = src2[-1];
        rc.eR = src2[4];
    };
    auto row_bits = [&](const RowW &r) {
        return dep_bits(r.v[0]) | dep_bits(r.v[1]) | dep_bits(r.v[2]) | dep_bits(r.v[3]) | dep_bits(r.eL) | dep_bits(r.eR);
    };
    // `bits`: OR of the registers whose loads no instruction has consumed yet (see mbar_arrive_after)
    auto release_stage = [&](uint32_t bits) {
        __syncwarp();
        if (lane == 0) mbar_arrive_after(&empty[stage], bits);
        if (++stage == kWStages) { stage = 0; full_parity ^= 1u; stage_p = ring_t; }
        else stage_p += stage_floats;
    };

    float *hout_p = hout + (long long)l0 * N + jc;
    float *cout_p = cout + (long long)l0 * N + jc;
    uint32_t *rgb_p = reinterpret_cast<uint32_t *>(a.rgb.out) + (long long)b * a.rgb_bubble_stride +
                      (long long)(l0 - 1) * N + jc;
    const float4 *rw_p = s_rw;
    const float2 *sc_p = s_sc;
    int grow = a.row_begin + l0 - 1;  // global row of the row computed next

    auto compute_row = [&](const RowW &hu, const RowW &hc, const RowW &hd, const RowW &cu, const RowW &cc, const RowW &cd) {
        const float4 rw = *rw_p;  // {sin, cos/sin_safe, 1/sin_safe^2, 1/sin_safe}, drainage.wgsl:92-101
        const float sin_theta = rw.x, cot = rw.y, inv_sin2 = rw.z, inv_sin = rw.w;
        float hn[4], cn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float h = hc.v[k], conc = cc.v[k];
            const float h_tm = hu.v[k], h_tp = hd.v[k];
            const float h_pm = k == 0 ? hc.eL : hc.v[k - 1], h_pp = k == 3 ? hc.eR : hc.v[k + 1];
            const float c_tm = cu.v[k], c_tp = cd.v[k];
            const float c_pm = k == 0 ? cc.eL : cc.v[k - 1], c_pp = k == 3 ? cc.eR : cc.v[k + 1];
            // thickness: gravity + surface Laplacian (+ Marangoni), drainage.wgsl:126-176
            const float drainage_term = -drainage_coeff * (h * h * h) * sin_theta;
            const float d2h_dtheta2 = (h_tp - 2.0f * h + h_tm) * inv_dth2;
            const float dh_dtheta = (h_tp - h_tm) * inv_2dth;
            const float d2h_dphi2 = (h_pp - 2.0f * h + h_pm) * inv_dph2;
            const float h_lap = (d2h_dtheta2 + cot * dh_dtheta + d2h_dphi2 * inv_sin2) * inv_r2;
            const float diffusion_term = p.diffusion_coeff * h_lap;
            const float dc_dtheta = (c_tp - c_tm) * inv_2dth;
            float marangoni_term = 0.f;
            if (mar) {
                const float dc_dphi = (c_pp - c_pm) * inv_2dph;
                const float gt = dc_dtheta * inv_r;
                const float gp = dc_dphi * inv_r * inv_sin;
                const float grad_c_mag = sqrtf(gt * gt + gp * gp);
                marangoni_term = p.marangoni_coeff * (p.gamma_reduction * grad_c_mag) * h * h;
            }
            const float new_h = fmaxf(h + dt * (drainage_term + diffusion_term + marangoni_term), 0.f);
            // concentration: diffusion (+ coupling), drainage.wgsl:181-201
            const float d2c_dtheta2 = (c_tp - 2.0f * conc + c_tm) * inv_dth2;
            const float d2c_dphi2 = (c_pp - 2.0f * conc + c_pm) * inv_dph2;
            const float c_lap = (d2c_dtheta2 + cot * dc_dtheta + d2c_dphi2 * inv_sin2) * inv_r2;
            float dc_dt = p.surfactant_diffusion * c_lap;
            if (mar) dc_dt += 0.001f * (h_lap * r2) * conc;
            const float new_c = fminf(fmaxf(conc + dt * dc_dt, 0.f), 1.f);
            const bool skip = h < crit;  // drainage.wgsl:108-112
            hn[k] = skip ? h : new_h;
            cn[k] = skip ? conc : new_c;
        }
        if (active) {
            *reinterpret_cast<float4 *>(hout_p) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4 *>(cout_p) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            if constexpr (MODE != SBX_RGB_NONE) {
                const float2 sc = *sc_p;
                const float ccy = sc.y * a.rgb.cam_y;
                uint32_t px[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    px[k] = shade_px<MODE>(hn[k], 0.f, 0.f, colA[k], sc.x, ccy, Rv, camR2, a.rgb);
                *reinterpret_cast<uint4 *>(rgb_p) = make_uint4(px[0], px[1], px[2], px[3]);
            }
        }
        hout_p += N;
        cout_p += N;
        rgb_p += N;
        ++rw_p;
        ++sc_p;
        ++grow;
    };

    RowW h0, h1, h2, g0, g1, g2;
    mbar_wait(&full[0], 0);
    load_slot(h0, g0, 0);
    load_slot(h1, g1, 1);
    int left = l1 - l0 + 1;
    uint32_t unconsumed = row_bits(h0) | row_bits(g0) | row_bits(h1) | row_bits(g1);  // first release: stage 0 is untouched
    while (true) {
        load_slot(h2, g2, 2);
        // slot 2 (h2, g2) was loaded just now; slot 1's edge values are first used by the row computed next
        release_stage(row_bits(h2) | row_bits(g2) | dep_bits(h1.eL) | dep_bits(h1.eR) | dep_bits(g1.eL) | dep_bits(g1.eR) | unconsumed);
        unconsumed = 0;
        compute_row(h0, h1, h2, g0, g1, g2);
        if (--left == 0) break;
        mbar_wait(&full[stage], full_parity);
        load_slot(h0, g0, 0);
        compute_row(h1, h2, h0, g1, g2, g0);
        if (--left == 0) break;
        load_slot(h1, g1, 1);
        compute_row(h2, h0, h1, g2, g0, g1);
        if (--left == 0) break;
    }
    if (a.peer.enabled)
        peer_signal_done(a.peer, a.peer.up_row && l0 == 1, a.peer.dn_row && l1 == a.l_last, C, tid);
}

bool wgsl_tma_supported(const WgslArgs &a)
{
    if (a.num_phi % 4 != 0 || a.num_phi < 128) return false;
    if (a.rgb.mode != SBX_RGB_NONE && a.rgb.format != SBX_OUT_RGBA8) return false;
    return true;
}

template <int MODE>
static cudaError_t launch_wgsl_one(const WgslArgs &a, dim3 grid, dim3 block, size_t smem, size_t smem_max, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_wgsl_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        {
            // three 50 KB rings need the 164 KB configuration (72 % of the maximum); the maximum would leave the LUT texture
            // 28 KB of L1 instead of 92 KB: + 3.5 % on the Marangoni + LUT step (profiles/r02_ab_carveout.txt)
            static const int carve_env = [] { const char *e = getenv("SBX_WGSL_CARVEOUT"); return e ? atoi(e) : -1; }();
            const int carve = carve_env >= 0 ? carve_env : (MODE == SBX_RGB_NONE ? (int)cudaSharedmemCarveoutMaxShared : 72);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wgsl_tma<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        }
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
    cfg.numAttrs = ((size_t)grid.x * grid.y * grid.z >= 1184) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_wgsl_tma<MODE>, a);
}

cudaError_t launch_wgsl_tma(WgslArgs a, int batch, int band_rows_req, cudaStream_t st)
{
    const int rows = a.l_last - a.l_first + 1;
    int consumers = 256;
    while (consumers > 32 && consumers * 2 >= a.num_phi) consumers >>= 1;
    const int strip = consumers * 4;
    const int gx = (a.num_phi + strip - 1) / strip;
    int band_rows = band_rows_req > 0 ? band_rows_req : auto_band_rows(rows, (long long)gx * batch, device_sm_count() * 3, 24);
    if (band_rows > rows) band_rows = rows;
    if (band_rows > kWMaxBandRows) band_rows = kWMaxBandRows;
    const int gy = (rows + band_rows - 1) / band_rows;
    a.band_rows = band_rows;
    const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)batch), block((unsigned)(consumers + 32));
    // the smem attribute is set once with the largest (256-consumer) size so that narrower launches fit too
    const size_t smem_max = (size_t)kWStages * kWRowsPerStage * 2 * (4 * 256 + 2 * kWHalo) * sizeof(float) + 2 * kWStages * sizeof(uint64_t);
    const size_t smem = (size_t)kWStages * kWRowsPerStage * 2 * (4 * consumers + 2 * kWHalo) * sizeof(float) + 2 * kWStages * sizeof(uint64_t);
    switch (a.rgb.mode) {
    case SBX_RGB_NONE: return launch_wgsl_one<SBX_RGB_NONE>(a, grid, block, smem, smem_max, st);
    case SBX_RGB_LUT: return launch_wgsl_one<SBX_RGB_LUT>(a, grid, block, smem, smem_max, st);
    case SBX_RGB_EXACT3: return launch_wgsl_one<SBX_RGB_EXACT3>(a, grid, block, smem, smem_max, st);
    case SBX_RGB_EXACT3_LINEAR: return launch_wgsl_one<SBX_RGB_EXACT3_LINEAR>(a, grid, block, smem, smem_max, st);
    case SBX_RGB_SCHLICK_WALL: return launch_wgsl_one<SBX_RGB_SCHLICK_WALL>(a, grid, block, smem, smem_max, st);
    case SBX_RGB_SCHLICK_TWO_BEAM: return launch_wgsl_one<SBX_RGB_SCHLICK_TWO_BEAM>(a, grid, block, smem, smem_max, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace sbx
