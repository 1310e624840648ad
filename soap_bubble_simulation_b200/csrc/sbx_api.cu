// sbx_api.cu -- the C ABI of libsbx.so (include/sbx.h): handles, host-side state machines
// (time, TFE fronts, ping-pong), coefficient tables, and kernel sequencing.  All arithmetic on the
// thickness field happens in the kernels of sbx_kernels.cu; there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

#include "sbx_comm.h"
#include "sbx_launch.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

using namespace sbx;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int32_t fail(int32_t code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? SBX_ERR_NOMEM : SBX_ERR_CUDA,           \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define CHECK_H(h)                                             \
    if (!(h)) return fail(SBX_ERR_INVALID, "null handle");     \
    CU(cudaSetDevice((h)->device))

extern "C" const char *sbx_last_error(void) { return g_err.c_str(); }

extern "C" int32_t sbx_version(int32_t *major, int32_t *minor)
{
    if (major) *major = SBX_VERSION_MAJOR;
    if (minor) *minor = SBX_VERSION_MINOR;
    return SBX_OK;
}

extern "C" int32_t sbx_device_count(int32_t *count)
{
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (count) *count = n;
    return SBX_OK;
}

extern "C" int32_t sbx_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(SBX_ERR_INVALID, "out is null");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return SBX_OK;
}

extern "C" int32_t sbx_host_free(void *p)
{
    CU(cudaFreeHost(p));
    return SBX_OK;
}

// ------------------------------------------------------------------ handle
struct sbx_drainage {
    sbx_drainage_cfg cfg{};
    int device = 0;
    int num_theta = 0, num_phi = 0, batch = 1;
    int row_begin = 0, row_end = 0, rows_local = 0;
    bool f64 = false, strict = false, wgsl = false;
    size_t esz = 4;
    long long bubble_stride = 0;  // elements
    long long total_elems = 0;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host-pipelined step: H2D / D2H engines
    cudaStream_t halo_stream = nullptr;                  // NCCL halo path: boundary bands + exchange beside the interior
    cudaEvent_t ev_prev = nullptr, ev_halo = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_comp;
    cudaEvent_t ev_start = nullptr, ev_out = nullptr, ev_all = nullptr;
    bool ghost_stale = false;  // sharded: the ghost rows of the current buffer were not refreshed since the field changed
    bool stale_by_pull = false;  // ... and only because peer-halo steps read the neighbours' rows in place (flag-ordered)
    void *buf[2] = {nullptr, nullptr};   // thickness ping-pong
    float *conc[2] = {nullptr, nullptr}; // WGSL concentration ping-pong
    int cur = 0;

    void *coefS = nullptr;      // T x 4 per global row
    void *coefF = nullptr;      // TMA fast path: 4 T per global row
    bool use_tma = true;        // SBX_NO_TMA=1 forces the direct-load kernel (A/B measurements)
    float2 *rowSC = nullptr;
    double4 *coefX = nullptr;
    float *rowW = nullptr;      // WGSL per-row float4
    void *bub = nullptr;        // BubbleParam<T>[batch]
    double *pole_partial = nullptr;
    unsigned long long *pole_count = nullptr;
    unsigned long long fold_count[2] = {0, 0};  // folded pole launches so far (top, bottom)
    int pole_slots_cap = 0;
    double *mm_part = nullptr;  // minmax partials
    double *mm_host = nullptr;  // pinned
    double *scalar_dev = nullptr;
    double *scalar_host = nullptr;  // pinned
    static constexpr int kMinmaxBlocks = 592;
    bool mm_valid = false;      // mm_cached holds min / max of the current field (cleared by everything that changes it)
    double mm_cached[2] = {0.0, 0.0};
    // min / max as a by-product of the TMA step kernel (StepArgs::mm_cur): two sets of {min, max} ordered words per
    // bubble; a tracked launch accumulates into one set and re-arms the other
    void *mm_track[2] = {nullptr, nullptr};
    void *mm_track_host = nullptr;   // pinned, one set
    int mm_track_set = 0;            // set the last tracked launch wrote
    bool mm_kernel_valid = false;    // that set describes the current field
    int mm_track_mode = 2;           // 0 off, 1 on, 2 auto: on once queries follow steps (SBX_TRACK_MINMAX overrides)
    bool mm_auto_on = false, stepped_since_mm = false;
    int mm_auto_hits = 0, mm_auto_misses = 0;

    // physics (the Rust struct's fields, drainage.rs:219-235)
    double density, gravity, viscosity, diffusion, critical, radius;
    std::vector<double> b_h0, b_radius, b_href;  // per bubble
    double h_ref = 0.0;
    double *href_dev = nullptr;  // per-bubble storage offsets (f64) for the conversion kernels
    bool href_zero = true;
    double coefA = 0.0;
    bool tables_dirty = true;
    sbx_drainage_params wp{};  // WGSL params

    // time + TFE state machine (drainage.rs:237-255)
    double time = 0.0;
    std::vector<sbx_tfe_front> fronts;
    // phi intervals of a front depend only on its centre and width, which never change after the spawn: cached so
    // that the per-step table costs a row interval per front (the resident path builds thousands of tables per call)
    struct ColCache { double phi_center, width; FrontCells c; };
    mutable ColCache col_cache[kMaxFronts + 1];
    mutable int col_cache_n = 0, col_cache_next = 0;
    double tfe_ratio = 0.85;
    bool tfe_enabled = true;
    double next_spawn = 0.5, spawn_interval = 2.0;

    // colour stage
    void *rgb_buf = nullptr;
    size_t rgb_bytes = 0;
    int rgb_format = -1;
    float *colA = nullptr;
    double cam_cached[3] = {NAN, NAN, NAN};
    cudaArray_t lut_array = nullptr;
    cudaTextureObject_t lut_tex = 0;
    float lut_n = -1.f;
    float lut_max_nm = 2000.f;
    bool lut_user = false;

    float4 *spec_dev = nullptr;  // SBX_RGB_SPECTRAL wavelength table
    int n_spec = 0;
    float spec_n_film = -1.f;
    std::vector<float> spec_wl;
    float *caustic_buf = nullptr;
    sbx_comm *comm = nullptr;
    // peer-memory halo (the ring kernels read the neighbours' boundary rows in place)
    bool peer_enabled = false;
    unsigned long long *sync_dev = nullptr;  // [0] flag from up, [1] flag from down, [2],[3] band counters, [4] status
    void *peer_up_buf[2] = {nullptr, nullptr}, *peer_dn_buf[2] = {nullptr, nullptr};
    void *peer_up_conc[2] = {nullptr, nullptr}, *peer_dn_conc[2] = {nullptr, nullptr};  // WGSL: concentration buffers
    unsigned long long *peer_up_sync = nullptr, *peer_dn_sync = nullptr;
    int peer_up_rows = 0;
    unsigned long long epoch = 0;
    int band_rows = 0;
    bool l2_keep_edges = false;          // field (in + out) larger than L2: keep the rows neighbouring bands share (SBX_L2_KEEP=0/1 overrides)
    bool use_resident = true;            // SBX_NO_RESIDENT=1: always use the streaming kernels
    FrontTable *res_tabs_host[2] = {nullptr, nullptr};  // pinned: per-step TFE tables of a resident launch
    FrontTable *res_tabs_dev[2] = {nullptr, nullptr};
    cudaEvent_t res_tabs_ev[2] = {nullptr, nullptr}, res_copy_ev[2] = {nullptr, nullptr};
    cudaStream_t tab_stream = nullptr;  // uploads the next resident launch's TFE tables beside the running one
    int res_slot = 0;
    int64_t launches = 0;
};

template <typename T>
static T *tbuf(sbx_drainage *h, int i) { return reinterpret_cast<T *>(h->buf[i]); }

static void free_all(sbx_drainage *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < 2; ++i) {
        if (h->buf[i]) cudaFree(h->buf[i]);
        if (h->conc[i]) cudaFree(h->conc[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (h->res_tabs_host[i]) cudaFreeHost(h->res_tabs_host[i]);
        cudaFree(h->res_tabs_dev[i]);
        if (h->res_tabs_ev[i]) cudaEventDestroy(h->res_tabs_ev[i]);
        if (h->res_copy_ev[i]) cudaEventDestroy(h->res_copy_ev[i]);
    }
    cudaFree(h->coefS);
    cudaFree(h->coefF);
    cudaFree(h->rowSC);
    cudaFree(h->coefX);
    cudaFree(h->rowW);
    cudaFree(h->bub);
    cudaFree(h->pole_partial);
    cudaFree(h->pole_count);
    cudaFree(h->mm_part);
    cudaFree(h->mm_track[0]);
    cudaFree(h->mm_track[1]);
    if (h->mm_track_host) cudaFreeHost(h->mm_track_host);
    cudaFree(h->scalar_dev);
    cudaFree(h->rgb_buf);
    cudaFree(h->colA);
    cudaFree(h->href_dev);
    for (int i = 0; i < 2; ++i) {
        if (h->peer_up_buf[i]) cudaIpcCloseMemHandle(h->peer_up_buf[i]);
        if (h->peer_dn_buf[i]) cudaIpcCloseMemHandle(h->peer_dn_buf[i]);
        if (h->peer_up_conc[i]) cudaIpcCloseMemHandle(h->peer_up_conc[i]);
        if (h->peer_dn_conc[i]) cudaIpcCloseMemHandle(h->peer_dn_conc[i]);
    }
    if (h->peer_up_sync) cudaIpcCloseMemHandle(h->peer_up_sync);
    if (h->peer_dn_sync) cudaIpcCloseMemHandle(h->peer_dn_sync);
    cudaFree(h->sync_dev);
    cudaFree(h->caustic_buf);
    cudaFree(h->spec_dev);
    if (h->mm_host) cudaFreeHost(h->mm_host);
    if (h->scalar_host) cudaFreeHost(h->scalar_host);
    if (h->lut_tex) cudaDestroyTextureObject(h->lut_tex);
    if (h->lut_array) cudaFreeArray(h->lut_array);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    if (h->halo_stream) cudaStreamDestroy(h->halo_stream);
    if (h->tab_stream) cudaStreamDestroy(h->tab_stream);
    if (h->ev_prev) cudaEventDestroy(h->ev_prev);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    for (auto e : h->ev_in) cudaEventDestroy(e);
    for (auto e : h->ev_comp) cudaEventDestroy(e);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->ev_all) cudaEventDestroy(h->ev_all);
    delete h;
}

extern "C" int32_t sbx_drainage_cfg_default(sbx_drainage_cfg *c)
{
    if (!c) return fail(SBX_ERR_INVALID, "cfg is null");
    memset(c, 0, sizeof *c);
    c->num_theta = 128;  // resolution, src/config.rs:313
    c->num_phi = 256;    // 2 * resolution, src/physics/drainage.rs:265
    c->batch = 1;
    c->h0_m = 500.0 * 1e-9;               // film_thickness_nm * 1e-9, config.rs:30,369-371
    c->density = 1000.0;                  // config.rs:61
    c->gravity = 9.81;                    // config.rs:96
    c->viscosity = 0.001;                 // config.rs:59
    c->diffusion = 1e-9;                  // config.rs:62
    c->critical_thickness_m = 10.0 * 1e-9;  // config.rs:31, drainage.rs:278
    c->radius_m = 0.05 / 2.0;             // config.rs:29,356-358
    c->semantics = SBX_SEM_CPU;
    c->precision = SBX_PREC_F32;
    c->h_ref_m = 0.0;
    c->tfe_enabled = 1;  // drainage.rs:284
    c->tfe_ratio = 0.85; // drainage.rs:283
    sbx_drainage_params &p = c->wgsl;  // gpu_drainage.rs:46-68
    p.dt = 0.001f; p.gravity = 9.81f; p.viscosity = 0.001f; p.density = 1000.0f;
    p.diffusion_coeff = 1e-9f; p.bubble_radius = 0.025f; p.critical_thickness = 30e-9f;
    p.grid_width = 128; p.grid_height = 64; p.marangoni_enabled = 0;
    p.gamma_air = 0.072f; p.gamma_reduction = 0.045f; p.surfactant_diffusion = 1e-9f;
    p.marangoni_coeff = 0.01f;
    return SBX_OK;
}

// ------------------------------------------------------------------ coefficient tables
template <typename T>
static int32_t upload_bubbles(sbx_drainage *h)
{
    std::vector<BubbleParam<T>> bp((size_t)h->batch);
    for (int b = 0; b < h->batch; ++b) {
        const double rb = h->b_radius[b];
        bp[b].diff_scale = (T)((h->radius / rb) * (h->radius / rb));
        bp[b].h_ref = (T)h->b_href[b];
        bp[b].radius = (float)rb;
        bp[b].pad = 0.f;
    }
    CU(cudaMemcpyAsync(h->bub, bp.data(), bp.size() * sizeof(BubbleParam<T>), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

template <typename T>
static int32_t upload_tables_t(sbx_drainage *h)
{
    const int N = h->num_theta;
    const double dth = M_PI / (double)(N - 1);           // drainage.rs:123
    const double dph = 2.0 * M_PI / (double)h->num_phi;  // drainage.rs:124
    const double coef = h->density * h->gravity / (3.0 * h->viscosity);  // drainage.rs:305
    const double r2 = h->radius * h->radius;
    std::vector<T> cs((size_t)N * 4, T(0));
    std::vector<T> cf((size_t)N * 4, T(0));
    std::vector<float2> sc((size_t)N);
    std::vector<double4> cx((size_t)N);
    for (int g = 0; g < N; ++g) {
        const double theta = (double)g * dth;  // theta_at, drainage.rs:149-151
        const double s = sin(theta), c = cos(theta);
        const double s_safe = fabs(s) < 1e-10 ? copysign(1e-10, s) : s;  // drainage.rs:320-324
        sc[g] = make_float2((float)s, (float)c);
        cx[g] = make_double4(s, c / s_safe, s_safe * s_safe, 0.0);
        if (g > 0 && g < N - 1) {
            cs[4 * (size_t)g + 0] = (T)(coef * s);
            cs[4 * (size_t)g + 1] = (T)(h->diffusion * (c / s_safe) / (r2 * 2.0 * dth));
            cs[4 * (size_t)g + 2] = (T)(h->diffusion / (r2 * dph * dph * s_safe * s_safe));
            const double A = h->diffusion / (r2 * dth * dth);
            const double B = h->diffusion * (c / s_safe) / (r2 * 2.0 * dth);
            cf[4 * (size_t)g + 0] = (T)(A + B);
            cf[4 * (size_t)g + 1] = (T)(A - B);
            cf[4 * (size_t)g + 2] = (T)(h->diffusion / (r2 * dph * dph * s_safe * s_safe));
            cf[4 * (size_t)g + 3] = (T)(coef * s);
        }
    }
    h->coefA = h->diffusion / (r2 * dth * dth);
    CU(cudaMemcpyAsync(h->coefS, cs.data(), cs.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->coefF, cf.data(), cf.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->rowSC, sc.data(), sc.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->coefX, cx.data(), cx.size() * sizeof(double4), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));  // host vectors go out of scope
    return upload_bubbles<T>(h);
}

static int32_t upload_tables_wgsl(sbx_drainage *h)
{
    // per-row constants evaluated in f32 exactly as the shader does (drainage.wgsl:92-101)
    const int N = h->num_theta;
    const float PI_F = 3.14159265358979323846f;
    const float dth = PI_F / (float)(N - 1);
    std::vector<float4> rw((size_t)N);
    std::vector<float2> sc((size_t)N);
    for (int g = 0; g < N; ++g) {
        const float theta = (float)g * dth;
        const float s = sinf(theta), c = cosf(theta);
        const float sgn = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
        const float s_safe = fabsf(s) < 0.0001f ? sgn * 0.0001f : s;
        rw[g] = (g > 0 && g < N - 1) ? make_float4(s, c / s_safe, 1.0f / (s_safe * s_safe), 1.0f / s_safe)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        const double td = (double)g * (M_PI / (double)(N - 1));
        sc[g] = make_float2((float)sin(td), (float)cos(td));
    }
    CU(cudaMemcpyAsync(h->rowW, rw.data(), rw.size() * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->rowSC, sc.data(), sc.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return upload_bubbles<float>(h);
}

static int32_t ensure_tables(sbx_drainage *h)
{
    if (!h->tables_dirty) return SBX_OK;
    int32_t rc = h->wgsl ? upload_tables_wgsl(h) : (h->f64 ? upload_tables_t<double>(h) : upload_tables_t<float>(h));
    if (rc == SBX_OK) h->tables_dirty = false;
    return rc;
}

// ------------------------------------------------------------------ create / destroy
template <typename T>
static int32_t fill_initial(sbx_drainage *h)
{
    // uniform start per bubble (ThicknessField::new, drainage.rs:121-122), ghost rows included
    for (int b = 0; b < h->batch; ++b) {
        const T v = (T)(h->b_h0[b] - h->b_href[b]);
        for (int i = 0; i < 2; ++i)
            CU(launch_fill<T>(tbuf<T>(h, i) + (long long)b * h->bubble_stride, h->bubble_stride, v, h->stream));
    }
    h->launches += 2 * h->batch;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_create(const sbx_drainage_cfg *cfg, int32_t device, sbx_drainage **out)
{
    if (!cfg || !out) return fail(SBX_ERR_INVALID, "cfg/out is null");
    *out = nullptr;
    if (cfg->num_theta < 3) return fail(SBX_ERR_INVALID, "num_theta must be >= 3 (got %d)", cfg->num_theta);
    if (cfg->num_phi < 2) return fail(SBX_ERR_INVALID, "num_phi must be >= 2 (got %d)", cfg->num_phi);
    if (cfg->batch < 1 || cfg->batch > 65535) return fail(SBX_ERR_INVALID, "batch must be in [1,65535] (got %d)", cfg->batch);
    if (cfg->semantics != SBX_SEM_CPU && cfg->semantics != SBX_SEM_WGSL) return fail(SBX_ERR_INVALID, "bad semantics %d", cfg->semantics);
    if (cfg->precision < SBX_PREC_F32 || cfg->precision > SBX_PREC_F64_STRICT) return fail(SBX_ERR_INVALID, "bad precision %d", cfg->precision);
    if (cfg->semantics == SBX_SEM_WGSL && cfg->precision != SBX_PREC_F32)
        return fail(SBX_ERR_INVALID, "SBX_SEM_WGSL is an f32 path (drainage.wgsl); precision must be SBX_PREC_F32");
    if (!(cfg->radius_m > 0.0) || !(cfg->viscosity > 0.0)) return fail(SBX_ERR_INVALID, "radius and viscosity must be > 0");
    int rb = cfg->row_begin, re = cfg->row_end;
    if (rb == 0 && re == 0) re = cfg->num_theta;
    if (rb < 0 || re > cfg->num_theta || re - rb < 1) return fail(SBX_ERR_INVALID, "bad row range [%d,%d)", rb, re);
    if ((rb == 0 && re < 2 && cfg->num_theta > 2) || (re == cfg->num_theta && re - rb < 2))
        return fail(SBX_ERR_INVALID, "a band holding a pole row must also hold the ring next to it");
    if (rb == 1 || re == cfg->num_theta - 1) return fail(SBX_ERR_INVALID, "a pole row may not be split from its ring");
    if (cfg->precision == SBX_PREC_F64_STRICT && cfg->batch_radius_m)
        return fail(SBX_ERR_UNSUPPORTED, "strict mode does not take per-bubble radii");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SBX_ERR_CUDA, "no usable CUDA device (%s); libsbx has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SBX_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));

    sbx_drainage *h = new sbx_drainage();
    h->cfg = *cfg;
    h->cfg.batch_h0_m = nullptr;
    h->cfg.batch_radius_m = nullptr;
    h->device = device;
    h->num_theta = cfg->num_theta;
    h->num_phi = cfg->num_phi;
    h->batch = cfg->batch;
    h->row_begin = rb;
    h->row_end = re;
    h->rows_local = re - rb;
    h->wgsl = cfg->semantics == SBX_SEM_WGSL;
    h->f64 = cfg->precision != SBX_PREC_F32;
    h->strict = cfg->precision == SBX_PREC_F64_STRICT;
    h->esz = h->f64 ? 8 : 4;
    h->bubble_stride = (long long)(h->rows_local + 2) * h->num_phi;
    h->total_elems = h->bubble_stride * h->batch;
    h->density = cfg->density; h->gravity = cfg->gravity; h->viscosity = cfg->viscosity;
    h->diffusion = cfg->diffusion; h->critical = cfg->critical_thickness_m; h->radius = cfg->radius_m;
    h->h_ref = h->f64 ? 0.0 : cfg->h_ref_m;
    h->tfe_enabled = cfg->tfe_enabled != 0;
    h->tfe_ratio = cfg->tfe_ratio;
    h->wp = cfg->wgsl;
    h->wp.grid_width = (uint32_t)cfg->num_phi;
    h->wp.grid_height = (uint32_t)cfg->num_theta;
    h->b_h0.assign((size_t)h->batch, cfg->h0_m);
    h->b_radius.assign((size_t)h->batch, h->wgsl ? (double)h->wp.bubble_radius : cfg->radius_m);
    if (cfg->batch_h0_m) for (int b = 0; b < h->batch; ++b) h->b_h0[b] = cfg->batch_h0_m[b];
    if (cfg->batch_radius_m) for (int b = 0; b < h->batch; ++b) h->b_radius[b] = cfg->batch_radius_m[b];
    // F32 storage offset: one value for the handle, or (with per-bubble h0 and a non-zero h_ref_m) each
    // bubble's own initial thickness
    h->b_href.assign((size_t)h->batch, h->h_ref);
    if (!h->f64 && cfg->batch_h0_m && cfg->h_ref_m != 0.0) h->b_href = h->b_h0;
    h->href_zero = std::all_of(h->b_href.begin(), h->b_href.end(), [](double v) { return v == 0.0; });
    if (const char *s = getenv("SBX_BAND_ROWS")) h->band_rows = atoi(s);
    if (const char *s = getenv("SBX_NO_TMA")) h->use_tma = atoi(s) == 0;
    if (const char *s = getenv("SBX_NO_RESIDENT")) h->use_resident = atoi(s) == 0;
    {
        int l2 = 0;
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device);
        h->l2_keep_edges = l2 > 0 && 2.0 * (double)h->total_elems * (double)h->esz > (double)l2;
        if (const char *s = getenv("SBX_L2_KEEP")) h->l2_keep_edges = atoi(s) != 0;
    }

#define CU_H(expr)                                                                     \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            int32_t rc__ = fail(e__ == cudaErrorMemoryAllocation ? SBX_ERR_NOMEM : SBX_ERR_CUDA, \
                                "%s failed: %s", #expr, cudaGetErrorString(e__));      \
            free_all(h);                                                               \
            return rc__;                                                               \
        }                                                                              \
    } while (0)

    CU_H(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    const size_t bytes = (size_t)h->total_elems * h->esz;
    for (int i = 0; i < 2; ++i) CU_H(cudaMalloc(&h->buf[i], bytes));
    if (h->wgsl)
        for (int i = 0; i < 2; ++i) CU_H(cudaMalloc(&h->conc[i], (size_t)h->total_elems * 4));
    const int N = h->num_theta;
    CU_H(cudaMalloc(&h->coefS, (size_t)N * 4 * h->esz));
    CU_H(cudaMalloc(&h->coefF, (size_t)N * 4 * h->esz));
    CU_H(cudaMalloc(&h->rowSC, (size_t)N * sizeof(float2)));
    CU_H(cudaMalloc(&h->coefX, (size_t)N * sizeof(double4)));
    CU_H(cudaMalloc(&h->rowW, (size_t)N * sizeof(float4)));
    CU_H(cudaMalloc(&h->bub, (size_t)h->batch * (h->f64 ? sizeof(BubbleParam<double>) : sizeof(BubbleParam<float>))));
    h->pole_slots_cap = h->num_phi / 32 + 2;
    CU_H(cudaMalloc(&h->pole_partial, (size_t)h->batch * 2 * h->pole_slots_cap * sizeof(double)));
    CU_H(cudaMemsetAsync(h->pole_partial, 0, (size_t)h->batch * 2 * h->pole_slots_cap * sizeof(double), h->stream));
    CU_H(cudaMalloc(&h->pole_count, (size_t)h->batch * 2 * sizeof(unsigned long long)));
    CU_H(cudaMemsetAsync(h->pole_count, 0, (size_t)h->batch * 2 * sizeof(unsigned long long), h->stream));
    CU_H(cudaMalloc(&h->mm_part, 2 * sbx_drainage::kMinmaxBlocks * sizeof(double)));
    CU_H(cudaHostAlloc(&h->mm_host, 2 * sbx_drainage::kMinmaxBlocks * sizeof(double), cudaHostAllocDefault));
    {
        const size_t wb = h->f64 ? 8 : 4, set_bytes = (size_t)h->batch * 2 * wb;
        CU_H(cudaHostAlloc(&h->mm_track_host, set_bytes, cudaHostAllocDefault));
        for (int b = 0; b < h->batch; ++b) {  // armed: min = all ones, max = 0
            if (h->f64) { ((unsigned long long *)h->mm_track_host)[2 * b] = ~0ull; ((unsigned long long *)h->mm_track_host)[2 * b + 1] = 0ull; }
            else { ((unsigned int *)h->mm_track_host)[2 * b] = ~0u; ((unsigned int *)h->mm_track_host)[2 * b + 1] = 0u; }
        }
        for (int i = 0; i < 2; ++i) {
            CU_H(cudaMalloc(&h->mm_track[i], set_bytes));
            CU_H(cudaMemcpy(h->mm_track[i], h->mm_track_host, set_bytes, cudaMemcpyHostToDevice));
        }
        if (const char *e = getenv("SBX_TRACK_MINMAX")) h->mm_track_mode = atoi(e) != 0 ? 1 : 0;
    }
    CU_H(cudaMalloc(&h->scalar_dev, 8 * sizeof(double)));
    CU_H(cudaHostAlloc(&h->scalar_host, 8 * sizeof(double), cudaHostAllocDefault));
    CU_H(cudaMalloc(&h->colA, (size_t)h->num_phi * sizeof(float)));
    CU_H(cudaMalloc(&h->href_dev, (size_t)h->batch * sizeof(double)));
    CU_H(cudaMalloc(&h->sync_dev, 8 * sizeof(unsigned long long)));
    CU_H(cudaMemset(h->sync_dev, 0, 8 * sizeof(unsigned long long)));
    CU_H(cudaMemcpy(h->href_dev, h->b_href.data(), (size_t)h->batch * sizeof(double), cudaMemcpyHostToDevice));
#undef CU_H

    int32_t rc = h->f64 ? fill_initial<double>(h) : fill_initial<float>(h);
    if (rc == SBX_OK && h->wgsl) {
        for (int i = 0; i < 2 && rc == SBX_OK; ++i) {  // initial concentration 0.5, gpu_drainage.rs:130-131
            cudaError_t ce = launch_fill<float>(h->conc[i], h->total_elems, 0.5f, h->stream);
            if (ce != cudaSuccess) rc = fail(SBX_ERR_CUDA, "fill failed: %s", cudaGetErrorString(ce));
            h->launches++;
        }
    }
    if (rc == SBX_OK) rc = ensure_tables(h);
    if (rc != SBX_OK) {
        free_all(h);
        return rc;
    }
    *out = h;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_destroy(sbx_drainage *h)
{
    if (!h) return SBX_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_all(h);
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_set_stream(sbx_drainage *h, void *cuda_stream)
{
    CHECK_H(h);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_sync(sbx_drainage *h)
{
    CHECK_H(h);
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

// ------------------------------------------------------------------ upload / download
template <typename TD, typename TS>
static int32_t upload_conv(sbx_drainage *h, TD *dst, const TS *host)
{
    const size_t row_bytes_src = (size_t)h->rows_local * h->num_phi * sizeof(TS);
    if (sizeof(TD) == sizeof(TS) && h->href_zero) {  // same type: strided copy straight into place
        CU(cudaMemcpy2DAsync(dst + h->num_phi, (size_t)h->bubble_stride * sizeof(TD), host, row_bytes_src,
                             row_bytes_src, (size_t)h->batch, cudaMemcpyHostToDevice, h->stream));
        return SBX_OK;
    }
    TS *stage = nullptr;
    CU(cudaMalloc(&stage, (size_t)h->total_elems * sizeof(TS)));
    cudaError_t e = cudaMemcpy2DAsync(stage + h->num_phi, (size_t)h->bubble_stride * sizeof(TS), host, row_bytes_src,
                                      row_bytes_src, (size_t)h->batch, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = launch_convert_in<TD, TS>(dst, stage, h->total_elems, h->bubble_stride, h->href_zero ? nullptr : h->href_dev, h->stream);
    h->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(stage);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

template <typename TH, typename TS>
static int32_t download_conv(sbx_drainage *h, TH *host, const TS *src)
{
    const size_t row_bytes_dst = (size_t)h->rows_local * h->num_phi * sizeof(TH);
    if (sizeof(TH) == sizeof(TS) && h->href_zero) {
        CU(cudaMemcpy2DAsync(host, row_bytes_dst, src + h->num_phi, (size_t)h->bubble_stride * sizeof(TS),
                             row_bytes_dst, (size_t)h->batch, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return SBX_OK;
    }
    TH *stage = nullptr;
    CU(cudaMalloc(&stage, (size_t)h->total_elems * sizeof(TH)));
    cudaError_t e = launch_convert_out<TH, TS>(stage, src, h->total_elems, h->bubble_stride, h->href_zero ? nullptr : h->href_dev, h->stream);
    h->launches++;
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(host, row_bytes_dst, stage + h->num_phi, (size_t)h->bubble_stride * sizeof(TH),
                              row_bytes_dst, (size_t)h->batch, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(stage);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

static int32_t halo_exchange(sbx_drainage *h);
static int32_t ensure_ghosts(sbx_drainage *h);
// Peer halo: a neighbour's last ring-kernel step may still be reading this band's boundary rows in place.  Anything
// that overwrites the field outside such a step first meets the neighbours in a (stream-ordered) ghost-row exchange:
// its receive cannot complete before the neighbour's send, which is queued behind the neighbour's step.
static int32_t quiesce_peers(sbx_drainage *h);

extern "C" int32_t sbx_drainage_upload(sbx_drainage *h, const void *host, int32_t dtype)
{
    CHECK_H(h);
    h->mm_valid = h->mm_kernel_valid = false;
    if (!host) return fail(SBX_ERR_INVALID, "host is null");
    if (dtype != SBX_F32 && dtype != SBX_F64) return fail(SBX_ERR_INVALID, "bad dtype %d", dtype);
    int32_t rc = quiesce_peers(h);
    if (rc != SBX_OK) return rc;
    if (h->f64) {
        rc = dtype == SBX_F64 ? upload_conv<double, double>(h, tbuf<double>(h, h->cur), (const double *)host)
                              : upload_conv<double, float>(h, tbuf<double>(h, h->cur), (const float *)host);
    } else {
        rc = dtype == SBX_F64 ? upload_conv<float, double>(h, tbuf<float>(h, h->cur), (const double *)host)
                              : upload_conv<float, float>(h, tbuf<float>(h, h->cur), (const float *)host);
    }
    if (rc != SBX_OK) return rc;
    if (h->wgsl)  // frozen pole rows live in both ping-pong buffers (drainage.wgsl:85-89)
        CU(cudaMemcpyAsync(h->buf[1 - h->cur], h->buf[h->cur], (size_t)h->total_elems * h->esz, cudaMemcpyDeviceToDevice, h->stream));
    return halo_exchange(h);
}

extern "C" int32_t sbx_drainage_download(sbx_drainage *h, void *host, int32_t dtype)
{
    CHECK_H(h);
    if (!host) return fail(SBX_ERR_INVALID, "host is null");
    if (dtype != SBX_F32 && dtype != SBX_F64) return fail(SBX_ERR_INVALID, "bad dtype %d", dtype);
    if (h->f64)
        return dtype == SBX_F64 ? download_conv<double, double>(h, (double *)host, tbuf<double>(h, h->cur))
                                : download_conv<float, double>(h, (float *)host, tbuf<double>(h, h->cur));
    return dtype == SBX_F64 ? download_conv<double, float>(h, (double *)host, tbuf<float>(h, h->cur))
                            : download_conv<float, float>(h, (float *)host, tbuf<float>(h, h->cur));
}

extern "C" int32_t sbx_drainage_upload_concentration(sbx_drainage *h, const float *host)
{
    CHECK_H(h);
    if (!h->wgsl) return fail(SBX_ERR_INVALID, "concentration exists only with SBX_SEM_WGSL");
    if (!host) return fail(SBX_ERR_INVALID, "host is null");
    { int32_t rcq = quiesce_peers(h); if (rcq != SBX_OK) return rcq; }
    const size_t rb = (size_t)h->rows_local * h->num_phi * 4;
    CU(cudaMemcpy2DAsync(h->conc[h->cur] + h->num_phi, (size_t)h->bubble_stride * 4, host, rb, rb, (size_t)h->batch,
                         cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->conc[1 - h->cur], h->conc[h->cur], (size_t)h->total_elems * 4, cudaMemcpyDeviceToDevice, h->stream));
    return halo_exchange(h);
}

extern "C" int32_t sbx_drainage_download_concentration(sbx_drainage *h, float *host)
{
    CHECK_H(h);
    if (!h->wgsl) return fail(SBX_ERR_INVALID, "concentration exists only with SBX_SEM_WGSL");
    if (!host) return fail(SBX_ERR_INVALID, "host is null");
    const size_t rb = (size_t)h->rows_local * h->num_phi * 4;
    CU(cudaMemcpy2DAsync(host, rb, h->conc[h->cur] + h->num_phi, (size_t)h->bubble_stride * 4, rb, (size_t)h->batch,
                         cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

// ------------------------------------------------------------------ TFE state machine (host, f64)
// update_marginal_regeneration without the field pass (drainage.rs:395-411) ...
static void tfe_advance(sbx_drainage *h, double dt)
{
    h->next_spawn -= dt;
    if (h->next_spawn <= 0.0) {
        const double t = h->time;  // current_time before this step's increment
        sbx_tfe_front f;           // spawn_tfe_front, drainage.rs:464-477
        f.phi_center = fabs(sin(t * 7.89 + 1.23)) * 2.0 * M_PI;
        f.width = 0.3 + fabs(sin(t * 3.21)) * 0.4;
        f.velocity = 0.05 + fabs(cos(t * 5.67)) * 0.05;
        f.extent = 0.0;
        h->fronts.push_back(f);
        if (h->fronts.size() > (size_t)kMaxFronts) h->fronts.erase(h->fronts.begin());
        const double random_factor = 0.5 + fabs(sin(h->time * 12.345));
        h->next_spawn = h->spawn_interval * random_factor;
    }
    for (auto &f : h->fronts) {  // TfeFront::advance, drainage.rs:47-49
        const double e = f.extent + f.velocity * dt;
        f.extent = e < M_PI_2 ? e : M_PI_2;
    }
    h->fronts.erase(std::remove_if(h->fronts.begin(), h->fronts.end(),
                                   [](const sbx_tfe_front &f) { return !(f.extent < M_PI_2 * 0.95); }),
                    h->fronts.end());
}

// largest [lo,hi] within [a_est-2,b_est+2] ∩ [lo_bound,hi_bound] on which P holds at both ends;
// P is true on one contiguous run, and the estimates are within a cell of its ends.
template <typename Pred>
static void exact_interval(Pred P, long a_est, long b_est, long lo_bound, long hi_bound, int &lo, int &hi)
{
    long a = std::max(a_est - 2, lo_bound), a_stop = std::min(std::max(a_est, b_est) + 2, hi_bound);
    while (a <= a_stop && !P(a)) ++a;
    long b = std::min(b_est + 2, hi_bound), b_stop = std::max(std::min(a_est, b_est) - 2, lo_bound);
    while (b >= b_stop && !P(b)) --b;
    if (a > a_stop || b < b_stop || a > b) { lo = 1; hi = 0; return; }
    lo = (int)a;
    hi = (int)b;
}

// ... and the cells its field pass touches (drainage.rs:419-460, TfeFront::contains :55-71), as exact
// integer intervals.
static void tfe_table(const sbx_drainage *h, FrontTable &ft)
{
    memset(&ft, 0, sizeof ft);
    ft.row_min = 1; ft.row_max = 0;
    if (h->fronts.empty()) return;
    const int NT = h->num_theta, NP = h->num_phi;
    const double dth = M_PI / (double)(NT - 1), dph = 2.0 * M_PI / (double)NP;
    size_t min_idx = (size_t)NT, max_idx = 0;
    for (const auto &f : h->fronts) {  // drainage.rs:428-434
        const double lo = M_PI_2 - f.extent, hi = M_PI_2 + f.extent;
        const double flo = floor(lo / dth), fhi = ceil(hi / dth);
        size_t a = flo <= 0.0 ? 0 : (size_t)flo, b = fhi <= 0.0 ? 0 : (size_t)fhi;
        b = std::min(b, (size_t)(NT - 1));
        min_idx = std::min(min_idx, a);
        max_idx = std::max(max_idx, b);
    }
    min_idx = std::min(min_idx, (size_t)(NT - 1));
    ft.row_min = (int)min_idx;
    ft.row_max = (int)max_idx;
    ft.count = (int)h->fronts.size();
    for (int k = 0; k < ft.count; ++k) {
        const sbx_tfe_front &f = h->fronts[k];
        FrontCells &c = ft.f[k];
        auto Prow = [&](long i) { const double th = (double)i * dth; return !(fabs(th - M_PI_2) > f.extent); };
        exact_interval(Prow, (long)ceil((M_PI_2 - f.extent) / dth), (long)floor((M_PI_2 + f.extent) / dth),
                       (long)min_idx, (long)max_idx, c.row_lo, c.row_hi);
        const sbx_drainage::ColCache *hit = nullptr;
        for (int i = 0; i < h->col_cache_n; ++i)
            if (h->col_cache[i].phi_center == f.phi_center && h->col_cache[i].width == f.width) { hit = &h->col_cache[i]; break; }
        if (!hit) {
            FrontCells cc;
            memset(&cc, 0, sizeof cc);
            const double hw = f.width / 2.0;
            auto Pdir = [&](long j) { const double d = fabs((double)j * dph - f.phi_center); return !(d > M_PI) && d < hw; };
            auto Pwrap = [&](long j) { const double d = fabs((double)j * dph - f.phi_center); return d > M_PI && (2.0 * M_PI - d) < hw; };
            exact_interval(Pdir, (long)ceil((f.phi_center - hw) / dph), (long)floor((f.phi_center + hw) / dph), 0, NP - 1,
                           cc.c0_lo, cc.c0_hi);
            cc.c1_lo = 1; cc.c1_hi = 0;
            if (f.phi_center - hw < 0.0)
                exact_interval(Pwrap, (long)ceil((f.phi_center + 2.0 * M_PI - hw) / dph), NP - 1, 0, NP - 1, cc.c1_lo, cc.c1_hi);
            else if (f.phi_center + hw > 2.0 * M_PI)
                exact_interval(Pwrap, 0, (long)floor((f.phi_center - 2.0 * M_PI + hw) / dph), 0, NP - 1, cc.c1_lo, cc.c1_hi);
            sbx_drainage::ColCache &slot = h->col_cache[h->col_cache_next];
            h->col_cache_next = (h->col_cache_next + 1) % (kMaxFronts + 1);
            h->col_cache_n = std::min(h->col_cache_n + 1, kMaxFronts + 1);
            slot.phi_center = f.phi_center;
            slot.width = f.width;
            slot.c = cc;
            hit = &slot;
        }
        c.c0_lo = hit->c.c0_lo; c.c0_hi = hit->c.c0_hi;
        c.c1_lo = hit->c.c1_lo; c.c1_hi = hit->c.c1_hi;
    }
}

// ------------------------------------------------------------------ colour stage set-up
static const float kLutWl7[7] = {400.f, 450.f, 500.f, 550.f, 600.f, 650.f, 700.f};  // interference_lut.rs:52
static int32_t ensure_lut(sbx_drainage *h, float n_film);

// per-wavelength table of SBX_RGB_SPECTRAL: {2 pi / lambda, w} with w = M cie(lambda) / n -- the CIE Gaussian fit
// (interference_lut.rs:136-148), the XYZ -> linear sRGB matrix (:158-163) and the 1 / n normalisation (:81-83),
// evaluated in f32 like the reference and folded once on the host
static int32_t ensure_spectrum(sbx_drainage *h)
{
    if (h->spec_dev && h->n_spec == (int)h->spec_wl.size()) return SBX_OK;
    if (h->spec_wl.empty()) h->spec_wl.assign(kLutWl7, kLutWl7 + 7);
    const int n = (int)h->spec_wl.size();
    std::vector<float4> tab((size_t)n);
    auto gauss = [](float x, float mean, float sigma) { const float t = (x - mean) / sigma; return expf(-0.5f * t * t); };
    for (int k = 0; k < n; ++k) {
        const float wl = h->spec_wl[k];
        const float x = fmaxf(1.056f * gauss(wl, 599.8f, 37.9f) + 0.362f * gauss(wl, 442.0f, 16.0f) - 0.065f * gauss(wl, 501.1f, 20.4f), 0.0f);
        const float y = fmaxf(0.821f * gauss(wl, 568.8f, 46.9f) + 0.286f * gauss(wl, 530.9f, 31.1f), 0.0f);
        const float z = fmaxf(1.217f * gauss(wl, 437.0f, 11.8f) + 0.681f * gauss(wl, 459.0f, 26.0f), 0.0f);
        const float inv_n = 1.0f / (float)n;
        tab[k] = make_float4((float)(2.0 * M_PI / (double)wl),
                             (3.2404542f * x - 1.5371385f * y - 0.4985314f * z) * inv_n,
                             (-0.969266f * x + 1.8760108f * y + 0.0415560f * z) * inv_n,
                             (0.0556434f * x - 0.2040259f * y + 1.0572252f * z) * inv_n);
    }
    CU(cudaStreamSynchronize(h->stream));
    if (h->spec_dev) { CU(cudaFree(h->spec_dev)); h->spec_dev = nullptr; }
    CU(cudaMalloc(&h->spec_dev, (size_t)n * sizeof(float4)));
    CU(cudaMemcpy(h->spec_dev, tab.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    h->n_spec = n;
    return SBX_OK;
}

static int32_t prepare_rgb(sbx_drainage *h, const sbx_view *v, RgbArgs &r)
{
    memset(&r, 0, sizeof r);
    if (!v || v->mode == SBX_RGB_NONE) { r.mode = SBX_RGB_NONE; return SBX_OK; }
    if (v->mode < SBX_RGB_NONE || v->mode > SBX_RGB_SPECTRAL) return fail(SBX_ERR_INVALID, "bad rgb mode %d", v->mode);
    if (v->format != SBX_OUT_RGBA8 && v->format != SBX_OUT_RGB_F32) return fail(SBX_ERR_INVALID, "bad rgb format %d", v->format);
    if (!(v->n_film > 0.0)) return fail(SBX_ERR_INVALID, "n_film must be > 0");
    const size_t cells = (size_t)h->batch * h->rows_local * h->num_phi;
    const size_t need = cells * (v->format == SBX_OUT_RGBA8 ? 4 : 12);
    if (h->rgb_bytes < need) {
        if (h->rgb_buf) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(h->rgb_buf)); h->rgb_buf = nullptr; h->rgb_bytes = 0; }
        CU(cudaMalloc(&h->rgb_buf, need));
        h->rgb_bytes = need;
    }
    h->rgb_format = v->format;
    if (!(v->cam_pos[0] == h->cam_cached[0] && v->cam_pos[1] == h->cam_cached[1] && v->cam_pos[2] == h->cam_cached[2])) {
        CU(launch_col_view(h->colA, h->num_phi, v->cam_pos[0], v->cam_pos[2], h->stream));
        h->launches++;
        memcpy(h->cam_cached, v->cam_pos, sizeof h->cam_cached);
    }
    if (v->mode == SBX_RGB_LUT) {
        int32_t rc = ensure_lut(h, (float)v->n_film);
        if (rc != SBX_OK) return rc;
    }
    if (v->mode == SBX_RGB_SPECTRAL) {
        int32_t rc = ensure_spectrum(h);
        if (rc != SBX_OK) return rc;
        r.spec = h->spec_dev;
        r.n_spec = h->n_spec;
    }
    r.mode = v->mode;
    r.format = v->format;
    r.out = h->rgb_buf;
    r.colA = h->colA;
    r.cam_y = (float)v->cam_pos[1];
    r.cam2 = (float)(v->cam_pos[0] * v->cam_pos[0] + v->cam_pos[1] * v->cam_pos[1] + v->cam_pos[2] * v->cam_pos[2]);
    r.n_film = (float)v->n_film;
    r.inv_n2 = (float)(1.0 / (v->n_film * v->n_film));
    r.intensity = v->intensity;
    r.inv_max_nm = 1.0f / h->lut_max_nm;
    r.nm_over_max = 1e9f / h->lut_max_nm;
    r.lut = h->lut_tex;
    r.inv_n2_d = 1.0 / (v->n_film * v->n_film);
    r.two_n = (float)(2.0 * v->n_film);
    r.r0 = (float)(((v->n_film - 1.0) / (v->n_film + 1.0)) * ((v->n_film - 1.0) / (v->n_film + 1.0)));
    r.omr0 = 1.0f - r.r0;
    const double wl[3] = {650.0, 532.0, 450.0};  // interference.rs:26-30
    for (int k = 0; k < 3; ++k) { r.two_n_over_wl[k] = 2.0 * v->n_film / wl[k]; r.inv_wl_f[k] = (float)(1.0 / wl[k]); r.two_pi_over_wl_f[k] = (float)(2.0 * M_PI / wl[k]); }
    r.inv_n2_hi = (float)r.inv_n2_d;
    r.inv_n2_lo = (float)(r.inv_n2_d - (double)r.inv_n2_hi);
    for (int k = 0; k < 3; ++k) { r.w_hi[k] = (float)r.two_n_over_wl[k]; r.w_lo[k] = (float)(r.two_n_over_wl[k] - (double)r.w_hi[k]); }
    return SBX_OK;
}

static int32_t install_lut(sbx_drainage *h, const uint8_t *dev_or_host, bool on_device, int T, int A, float max_nm)
{
    CU(cudaStreamSynchronize(h->stream));
    if (h->lut_tex) { CU(cudaDestroyTextureObject(h->lut_tex)); h->lut_tex = 0; }
    if (h->lut_array) { CU(cudaFreeArray(h->lut_array)); h->lut_array = nullptr; }
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
    CU(cudaMallocArray(&h->lut_array, &ch, (size_t)T, (size_t)A));
    CU(cudaMemcpy2DToArrayAsync(h->lut_array, 0, 0, dev_or_host, (size_t)T * 4, (size_t)T * 4, (size_t)A,
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = h->lut_array;
    cudaTextureDesc td{};  // sampler at pipeline.rs:350-359: clamp-to-edge, linear, Rgba8Unorm
    td.addressMode[0] = cudaAddressModeClamp;
    td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    CU(cudaCreateTextureObject(&h->lut_tex, &rd, &td, nullptr));
    h->lut_max_nm = max_nm;
    return SBX_OK;
}


static int32_t build_lut_device(float n, float intensity, const float *wl, int nl, int T, int A, float max_nm,
                                uint8_t **dev_out, cudaStream_t st)
{
    float *wl_dev = nullptr;
    uint8_t *out = nullptr;
    CU(cudaMalloc(&wl_dev, (size_t)nl * 4));
    cudaError_t e = cudaMalloc(&out, (size_t)T * A * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(wl_dev, wl, (size_t)nl * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_lut_build(out, T, A, max_nm, n, intensity, wl_dev, nl, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(wl_dev);
    if (e != cudaSuccess) {
        cudaFree(out);
        return fail(SBX_ERR_CUDA, "LUT build failed: %s", cudaGetErrorString(e));
    }
    *dev_out = out;
    return SBX_OK;
}

static int32_t ensure_lut(sbx_drainage *h, float n_film)
{
    if (h->lut_tex && (h->lut_user || h->lut_n == n_film)) return SBX_OK;
    // generate_interference_lut(refractive_index, 1.0): intensity is applied at shading time
    // (pipeline.rs:312-315, bubble.wgsl:435-436)
    uint8_t *dev = nullptr;
    int32_t rc = build_lut_device(n_film, 1.0f, kLutWl7, 7, 256, 64, 2000.f, &dev, h->stream);
    if (rc != SBX_OK) return rc;
    h->launches++;
    rc = install_lut(h, dev, true, 256, 64, 2000.f);
    cudaFree(dev);
    if (rc == SBX_OK) { h->lut_n = n_film; h->lut_user = false; }
    return rc;
}

extern "C" int32_t sbx_drainage_set_lut(sbx_drainage *h, const uint8_t *rgba, int32_t T, int32_t A, float max_nm)
{
    CHECK_H(h);
    if (!rgba || T < 2 || A < 2 || !(max_nm > 0.f)) return fail(SBX_ERR_INVALID, "bad LUT arguments");
    int32_t rc = install_lut(h, rgba, false, T, A, max_nm);
    if (rc == SBX_OK) h->lut_user = true;
    return rc;
}

// ------------------------------------------------------------------ stepping
static void interior_range(const sbx_drainage *h, int &l_first, int &l_last)
{
    const int g_first = std::max(1, h->row_begin), g_last = std::min(h->num_theta - 2, h->row_end - 1);
    l_first = g_first - h->row_begin + 1;
    l_last = g_last - h->row_begin + 1;
}

template <typename T>
static void fill_step_args(sbx_drainage *h, StepArgs<T> &a, double dt, const RgbArgs &rgb)
{
    memset(&a, 0, sizeof a);
    a.in = tbuf<T>(h, h->cur);
    a.out = tbuf<T>(h, 1 - h->cur);
    a.coefS = reinterpret_cast<const T *>(h->coefS);
    a.coefF = reinterpret_cast<const T *>(h->coefF);
    a.rowSC = h->rowSC;
    a.coefX = h->coefX;
    a.bub = reinterpret_cast<const BubbleParam<T> *>(h->bub);
    a.pole_partial = h->pole_partial;
    a.pole_count = h->pole_count;
    a.dev_status = (unsigned int *)(h->sync_dev + 4);
    a.num_phi = h->num_phi;
    a.num_theta = h->num_theta;
    a.row_begin = h->row_begin;
    interior_range(h, a.l_first, a.l_last);
    a.bubble_stride = h->bubble_stride;
    a.rgb_bubble_stride = (long long)h->rows_local * h->num_phi;
    a.dt = (T)dt;
    a.crit = (T)h->critical;
    a.coefA = (T)h->coefA;
    const double reduction_rate = (1.0 - h->tfe_ratio) * 0.5;  // drainage.rs:453
    a.tfe_rdt = (T)(reduction_rate * dt);
    const double dth = M_PI / (double)(h->num_theta - 1), dph = 2.0 * M_PI / (double)h->num_phi;
    a.s_drain_coef = h->density * h->gravity / (3.0 * h->viscosity);
    a.s_dth2 = dth * dth;
    a.s_two_dth = 2.0 * dth;
    a.s_dph2 = dph * dph;
    a.s_r2 = h->radius * h->radius;
    a.s_D = h->diffusion;
    a.s_dt = dt;
    a.s_crit = h->critical;
    a.s_tfe_factor = 1.0 - reduction_rate * dt;  // drainage.rs:454
    a.rgb = rgb;
    a.l2_keep_edges = h->l2_keep_edges ? 1 : 0;
}

// arguments of the in-kernel halo read for the step that reads ping-pong half cur (every rank reads the same half in
// the same step); element size = the handle's storage type
static void fill_peer(sbx_drainage *h, PeerArgs &p)
{
    p.enabled = 1;
    p.epoch = ++h->epoch;
    const size_t N = (size_t)h->num_phi;
    const int c = h->cur;
    const size_t up_off = (size_t)h->peer_up_rows * N;  // the up neighbour's last owned row (its local row 0 is a ghost row)
    p.up_row = h->peer_up_buf[c] ? (char *)h->peer_up_buf[c] + up_off * h->esz : nullptr;
    p.dn_row = h->peer_dn_buf[c] ? (char *)h->peer_dn_buf[c] + N * h->esz : nullptr;  // its first owned row
    p.up_row2 = h->peer_up_conc[c] ? (char *)h->peer_up_conc[c] + up_off * 4 : nullptr;
    p.dn_row2 = h->peer_dn_conc[c] ? (char *)h->peer_dn_conc[c] + N * 4 : nullptr;
    p.flag_up = h->peer_up_sync ? h->peer_up_sync + 1 : nullptr;  // its last band
    p.flag_dn = h->peer_dn_sync ? h->peer_dn_sync + 0 : nullptr;  // its first band
    p.my_first_done = h->sync_dev + 0;
    p.my_last_done = h->sync_dev + 1;
    p.done_first = (unsigned int *)(h->sync_dev + 2);
    p.done_last = (unsigned int *)(h->sync_dev + 3);
    p.status = (unsigned int *)(h->sync_dev + 4);
#ifdef SBX_PEER_DEBUG
    static const unsigned dbg = [] { const char *e = getenv("SBX_PEER_DEBUG_BITS"); return e ? (unsigned)atoi(e) : 0u; }();
    p.debug = dbg;
#endif
    // the ghost rows of the local buffers are not written by these steps: a reader that wants them (a step that does
    // not go through the ring kernels, the host-buffer step, get_thickness) fetches them first (ensure_ghosts)
    if (!h->ghost_stale) h->stale_by_pull = true;
    h->ghost_stale = true;
}

// decide whether this launch writes the pole rows inside the TMA kernel, and which arrival count it waits for.
// The host-side count is committed only after the launch has been accepted (commit_pole_fold): a failed launch
// must not leave the device counters one round behind what every later step would wait for.
template <typename T>
static void set_pole_fold(sbx_drainage *h, StepArgs<T> &a, bool top, bool bottom)
{
    a.fold_top = a.fold_bottom = 0;
    if (!h->use_tma || h->strict || !march_tma_folds_poles<T>(a)) return;
    if (top) { a.fold_top = 1; a.pole_want[0] = h->fold_count[0] + 1; }
    if (bottom) { a.fold_bottom = 1; a.pole_want[1] = h->fold_count[1] + 1; }
}
template <typename T>
static void commit_pole_fold(sbx_drainage *h, const StepArgs<T> &a)
{
    if (a.fold_top) ++h->fold_count[0];
    if (a.fold_bottom) ++h->fold_count[1];
}

template <typename T>
static int32_t step_cpu_sem(sbx_drainage *h, double dt, const RgbArgs &rgb)
{
    FrontTable ft;
    memset(&ft, 0, sizeof ft);
    ft.row_min = 1; ft.row_max = 0;
    if (h->tfe_enabled) {  // drainage.rs:387-389
        tfe_advance(h, dt);
        tfe_table(h, ft);
    }
    StepArgs<T> a;
    fill_step_args<T>(h, a, dt, rgb);
    const bool has_top = h->row_begin == 0, has_bottom = h->row_end == h->num_theta;
    int n = 0;
    if (h->strict) {
        // the reference's pass order: interior, poles, TFE (drainage.rs:314-389); colour afterwards
        { int32_t rcg = ensure_ghosts(h); if (rcg != SBX_OK) return rcg; }
        RgbArgs none;
        memset(&none, 0, sizeof none);
        a.rgb = none;
        CU(launch_step<T>(a, h->batch, false, true, has_top, has_bottom, h->band_rows, false, h->stream, &n));
        CU(launch_tfe_pass<T>(tbuf<T>(h, 1 - h->cur), h->bubble_stride, h->num_phi, h->row_begin, h->rows_local,
                              h->batch, ft, a.s_tfe_factor, h->stream, &n));
        h->cur = 1 - h->cur;
        if (rgb.mode != SBX_RGB_NONE) {
            StepArgs<T> s;
            fill_step_args<T>(h, s, dt, rgb);
            s.l_first = 1;
            s.l_last = h->rows_local;
            CU(launch_shade<T>(s, h->batch, h->band_rows, h->stream, &n, h->use_tma));
        }
    } else {
        a.fronts = ft;
        bool pushed = false;
        set_pole_fold<T>(h, a, has_top, has_bottom);
        {
            if (h->peer_enabled && h->use_tma && march_tma_supported<T>(a)) {
                fill_peer(h, a.peer);
                pushed = true;
            }
        }
        if (!pushed) { int32_t rcg = ensure_ghosts(h); if (rcg != SBX_OK) return rcg; }
        const bool want_track = h->mm_track_mode == 1 || (h->mm_track_mode == 2 && h->mm_auto_on);
        const bool tracked = want_track && h->use_tma && march_tma_supported<T>(a) &&
                             ((!has_top || a.fold_top) && (!has_bottom || a.fold_bottom));  // the kernel writes every owned row
        if (tracked) {
            const int set = 1 - h->mm_track_set;
            a.mm_cur = h->mm_track[set];
            a.mm_next = h->mm_track[1 - set];
            h->mm_track_set = set;
        }
        // NCCL halo path (the modes the peer halo does not cover, or SBX halo = nccl): boundary first.  The bands next
        // to a neighbour run on a second, high-priority stream and their rows are exchanged there with ncclSend/ncclRecv
        // while the interior bands run on the handle's stream; the step ends when both are done.
        const int edge = 32;  // rows of a boundary launch
        const bool up_nb = h->comm && h->row_begin > 0, dn_nb = h->comm && h->row_end < h->num_theta;
        if (!pushed && (up_nb || dn_nb) && h->batch == 1 && !tracked && a.l_last - a.l_first + 1 >= 4 * edge &&
            !getenv("SBX_NO_BOUNDARY_FIRST")) {
            if (!h->halo_stream) {
                int lo = 0, hi = 0;
                CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CU(cudaStreamCreateWithPriority(&h->halo_stream, cudaStreamNonBlocking, hi));
                CU(cudaEventCreateWithFlags(&h->ev_prev, cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
            }
            CU(cudaEventRecord(h->ev_prev, h->stream));
            CU(cudaStreamWaitEvent(h->halo_stream, h->ev_prev, 0));
            const int lf = a.l_first, ll = a.l_last;
            int int_first = lf, int_last = ll;
            auto launch_part = [&](int p_first, int p_last, cudaStream_t st) -> int32_t {
                StepArgs<T> ap = a;
                ap.l_first = p_first;
                ap.l_last = p_last;
                // the pole rows are written by the launch that holds the ring next to them
                const bool top = has_top && p_first == lf, bottom = has_bottom && p_last == ll;
                set_pole_fold<T>(h, ap, top, bottom);
                CU(launch_step<T>(ap, 1, ft.count > 0, false, top, bottom, h->band_rows, h->use_tma, st, &n));
                commit_pole_fold<T>(h, ap);
                return SBX_OK;
            };
            int32_t rc = SBX_OK;
            if (up_nb) { rc = launch_part(lf, lf + edge - 1, h->halo_stream); int_first = lf + edge; }
            if (rc == SBX_OK && dn_nb) { rc = launch_part(ll - edge + 1, ll, h->halo_stream); int_last = ll - edge; }
            if (rc != SBX_OK) return rc;
            h->cur = 1 - h->cur;
            {   // exchange the new boundary rows on the halo stream
                cudaStream_t keep = h->stream;
                h->stream = h->halo_stream;
                rc = halo_exchange(h);
                h->stream = keep;
                if (rc != SBX_OK) return rc;
            }
            CU(cudaEventRecord(h->ev_halo, h->halo_stream));
            h->cur = 1 - h->cur;
            rc = launch_part(int_first, int_last, h->stream);
            if (rc != SBX_OK) return rc;
            CU(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
            h->cur = 1 - h->cur;
            h->mm_kernel_valid = false;
            h->launches += n;
            h->time += dt;
            return SBX_OK;
        }
        CU(launch_step<T>(a, h->batch, ft.count > 0, false, has_top, has_bottom, h->band_rows, h->use_tma, h->stream, &n));
        commit_pole_fold<T>(h, a);
        h->cur = 1 - h->cur;
        h->mm_kernel_valid = tracked;
        if (pushed) {
            h->launches += n;
            h->time += dt;
            return SBX_OK;  // the next step reads the neighbours' rows in place
        }
    }
    h->launches += n;
    h->time += dt;  // drainage.rs:391
    return halo_exchange(h);
}

static void fill_wgsl_args(sbx_drainage *h, WgslArgs &a, const RgbArgs &rgb)
{
    memset(&a, 0, sizeof a);
    a.h_in = tbuf<float>(h, h->cur);
    a.h_out = tbuf<float>(h, 1 - h->cur);
    a.c_in = h->conc[h->cur];
    a.c_out = h->conc[1 - h->cur];
    a.rowW = h->rowW;
    a.rowSC = h->rowSC;
    a.num_phi = h->num_phi;
    a.num_theta = h->num_theta;
    a.row_begin = h->row_begin;
    a.rows_local = h->rows_local;
    interior_range(h, a.l_first, a.l_last);
    a.bubble_stride = h->bubble_stride;
    a.rgb_bubble_stride = (long long)h->rows_local * h->num_phi;
    a.p = h->wp;
    const float PI_F = 3.14159265358979323846f;
    a.delta_theta = PI_F / (float)(h->wp.grid_height - 1u);  // drainage.wgsl:92
    a.delta_phi = 2.0f * PI_F / (float)h->wp.grid_width;     // drainage.wgsl:93
    a.radius_view = h->wp.bubble_radius;
    a.rgb = rgb;
}

static int32_t step_wgsl_sem(sbx_drainage *h, double dt, const RgbArgs &rgb)
{
    WgslArgs a;
    fill_wgsl_args(h, a, rgb);
    int n = 0;
    a.l2_keep_edges = h->l2_keep_edges ? 1 : 0;
    const bool pushed = h->peer_enabled && h->use_tma && wgsl_tma_supported(a);
    if (pushed) fill_peer(h, a.peer);
    else { int32_t rcg = ensure_ghosts(h); if (rcg != SBX_OK) return rcg; }
    CU(launch_wgsl(a, h->batch, h->band_rows, h->use_tma, h->stream, &n));
    if (rgb.mode != SBX_RGB_NONE && (h->row_begin == 0 || h->row_end == h->num_theta)) {
        // frozen pole rows still need their colour: shade them from the (unchanged) pole values
        StepArgs<float> s;
        fill_step_args<float>(h, s, dt, rgb);
        s.in = tbuf<float>(h, 1 - h->cur);
        if (h->row_begin == 0) { s.l_first = 1; s.l_last = 1; CU(launch_shade<float>(s, h->batch, 1, h->stream, &n)); }
        if (h->row_end == h->num_theta) { s.l_first = h->rows_local; s.l_last = h->rows_local; CU(launch_shade<float>(s, h->batch, 1, h->stream, &n)); }
    }
    h->cur = 1 - h->cur;
    h->launches += n;
    h->time += dt;  // gpu_drainage.rs:368
    if (pushed) return SBX_OK;  // the next step reads the neighbours' rows of both fields in place
    return halo_exchange(h);
}

// Small whole-sphere grids: take `nsteps` colourless steps inside one launch with the field resident in shared
// memory (sbx_resident.cu).  Returns the number of steps taken (0: not applicable, the caller streams them).
static bool resident_possible(const sbx_drainage *h)
{
    return h->use_resident && !h->comm && !h->peer_enabled && h->row_begin == 0 && h->rows_local == h->num_theta;
}

constexpr int kResidentChunk = 1024;

// `rgb` (mode != NONE): the last launch of the call also writes the colour frame of the final field
template <typename T>
static int32_t steps_resident(sbx_drainage *h, double dt, int nsteps, int *taken, const RgbArgs &rgb)
{
    *taken = 0;
    RgbArgs none;
    memset(&none, 0, sizeof none);
    StepArgs<T> a;
    fill_step_args<T>(h, a, dt, none);
    if (!resident_possible(h) || nsteps < 2 || resident_cluster<T>(a, h->batch, h->strict) == 0) return SBX_OK;
    // Steps per launch.  The host state machine of chunk k + 1 (0.14 us per step) runs while the kernel of chunk k does:
    // two pinned host / device table buffers alternate, an event per buffer says when its upload has been consumed.
    constexpr int kChunk = kResidentChunk;
    if (h->tfe_enabled && !h->res_tabs_host[0]) {
        for (int i = 0; i < 2; ++i) {
            CU(cudaMallocHost(&h->res_tabs_host[i], (size_t)kChunk * sizeof(FrontTable)));
            CU(cudaMalloc(&h->res_tabs_dev[i], (size_t)kChunk * sizeof(FrontTable)));
            CU(cudaEventCreateWithFlags(&h->res_tabs_ev[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->res_copy_ev[i], cudaEventDisableTiming));
        }
        CU(cudaStreamCreateWithFlags(&h->tab_stream, cudaStreamNonBlocking));
    }
    while (nsteps > 0) {
        const int n = std::min(nsteps, kChunk);
        const FrontTable *tabs = nullptr;
        int used_slot = -1;
        if (h->tfe_enabled) {
            const int slot = h->res_slot;
            h->res_slot ^= 1;
            CU(cudaEventSynchronize(h->res_tabs_ev[slot]));  // the upload that last used this host buffer is done
            FrontTable *host = h->res_tabs_host[slot];
            for (int s = 0; s < n; ++s) {  // the host state machine, one step at a time (drainage.rs:387-391)
                tfe_advance(h, dt);
                tfe_table(h, host[s]);
                h->time += dt;
                FrontTable &t = host[s];
                t.rows_changed = t.cols_changed = 1;
                if (s > 0 && host[s - 1].count == t.count) {
                    const FrontTable &p = host[s - 1];
                    int rows = p.row_min != t.row_min || p.row_max != t.row_max, cols = 0;
                    for (int k = 0; k < t.count; ++k) {
                        rows |= p.f[k].row_lo != t.f[k].row_lo || p.f[k].row_hi != t.f[k].row_hi;
                        cols |= p.f[k].c0_lo != t.f[k].c0_lo || p.f[k].c0_hi != t.f[k].c0_hi || p.f[k].c1_lo != t.f[k].c1_lo ||
                                p.f[k].c1_hi != t.f[k].c1_hi;
                    }
                    t.rows_changed = rows;
                    t.cols_changed = cols;
                }
            }
            bool any_front = false;
            for (int s = 0; s < n; ++s) any_front |= host[s].count > 0;
            if (!any_front) {
                // no front alive in any of these steps (e.g. the first 0.5 s of a simulator): the launch needs no tables
                tabs = nullptr;
            } else {
                // The tables travel on a stream of their own, so the copy for launch k + 1 runs while launch k computes
                // instead of sitting between two launches (a frame loop calls this every 10 steps: the serialised copy was
                // a third of the frame).  Safe: the host waited above for the launch that last read this slot, and the
                // launch below waits for the copy.  (Reading the tables straight from pinned host memory was measured
                // and rejected: the per-step prefetch then pays a PCIe round trip, 42 us per 10-step frame instead of 29.)
                CU(cudaMemcpyAsync(h->res_tabs_dev[slot], host, (size_t)n * sizeof(FrontTable), cudaMemcpyHostToDevice, h->tab_stream));
                CU(cudaEventRecord(h->res_copy_ev[slot], h->tab_stream));
                CU(cudaStreamWaitEvent(h->stream, h->res_copy_ev[slot], 0));
                tabs = h->res_tabs_dev[slot];
                used_slot = slot;
            }
        } else {
            for (int s = 0; s < n; ++s) h->time += dt;
        }
        if (n == nsteps) a.rgb = rgb;  // last launch: colour epilogue
        CU(launch_resident<T>(a, h->batch, h->strict, tabs, n, h->stream));
        if (used_slot >= 0) CU(cudaEventRecord(h->res_tabs_ev[used_slot], h->stream));  // slot free again when this launch is done
        h->launches++;
        *taken += n;
        nsteps -= n;
    }
    return SBX_OK;
}

static int32_t steps_resident_wgsl(sbx_drainage *h, double dt, int nsteps, int *taken, const RgbArgs &rgb)
{
    *taken = 0;
    WgslArgs a;
    fill_wgsl_args(h, a, rgb);
    if (!resident_possible(h) || nsteps < 2 || resident_cluster_wgsl(a, h->batch) == 0) return SBX_OK;
    CU(launch_resident_wgsl(a, h->batch, nsteps, h->stream));  // in place on the current buffers
    for (int s = 0; s < nsteps; ++s) h->time += dt;             // gpu_drainage.rs:368
    h->launches++;
    *taken = nsteps;
    return SBX_OK;
}

static int32_t shade_current(sbx_drainage *h, const RgbArgs &rgb);

static int32_t do_steps(sbx_drainage *h, double dt, int32_t nsteps, const sbx_view *view)
{
    if (nsteps < 0) return fail(SBX_ERR_INVALID, "nsteps < 0");
    if (nsteps > 0) {
        h->mm_valid = h->mm_kernel_valid = false;
        // auto mode: a caller that asks for min/max after its steps gets them from the kernel; one that stops asking loses it
        if (h->stepped_since_mm && ++h->mm_auto_misses >= 8) { h->mm_auto_on = false; h->mm_auto_hits = 0; }
        h->stepped_since_mm = true;
    }
    int32_t rc = ensure_tables(h);
    if (rc != SBX_OK) return rc;
    // (peer halo: the ring kernels read the neighbours' rows in place; the step functions fetch ghost rows only for
    //  the launches that do not)
    if (nsteps > 0 && !(h->peer_enabled && h->stale_by_pull) && (rc = ensure_ghosts(h)) != SBX_OK) return rc;
    RgbArgs rgb;
    rc = prepare_rgb(h, view, rgb);
    if (rc != SBX_OK) return rc;
    // SBX_RGB_SPECTRAL is a colour pass of its own (SFU-bound, ~10x the step): steps run without a colour stage
    const bool colour_pass_after = rgb.mode == SBX_RGB_SPECTRAL;
    const RgbArgs rgb_after = rgb;
    if (colour_pass_after) memset(&rgb, 0, sizeof rgb);
    // only the last step's colour frame is observable: the steps before it may run without the colour stage
    // ... unless the resident kernels can shade the final field themselves (one launch for the whole call)
    const bool epilogue = rgb.mode != SBX_RGB_NONE && resident_shades(rgb);
    const int plain = rgb.mode == SBX_RGB_NONE || epilogue ? nsteps : nsteps - 1;
    if (plain >= 2) {
        int taken = 0;
        RgbArgs none;
        memset(&none, 0, sizeof none);
        const RgbArgs &last = epilogue ? rgb : none;
        rc = h->wgsl ? steps_resident_wgsl(h, dt, plain, &taken, last)
                     : (h->f64 ? steps_resident<double>(h, dt, plain, &taken, last) : steps_resident<float>(h, dt, plain, &taken, last));
        if (rc != SBX_OK) return rc;
        nsteps -= taken;
    }
    for (int s = 0; s < nsteps; ++s) {
        rc = h->wgsl ? step_wgsl_sem(h, dt, rgb) : (h->f64 ? step_cpu_sem<double>(h, dt, rgb) : step_cpu_sem<float>(h, dt, rgb));
        if (rc != SBX_OK) return rc;
    }
    if (colour_pass_after) return shade_current(h, rgb_after);
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_step(sbx_drainage *h, double dt, int32_t nsteps)
{
    CHECK_H(h);
    return do_steps(h, dt, nsteps, nullptr);
}

extern "C" int32_t sbx_drainage_step_fused(sbx_drainage *h, double dt, int32_t nsteps, const sbx_view *view)
{
    CHECK_H(h);
    if (!view) return fail(SBX_ERR_INVALID, "view is null");
    return do_steps(h, dt, nsteps, view);
}

extern "C" int32_t sbx_drainage_shade(sbx_drainage *h, const sbx_view *view)
{
    CHECK_H(h);
    if (!view || view->mode == SBX_RGB_NONE) return fail(SBX_ERR_INVALID, "view is null or mode is NONE");
    int32_t rc = ensure_tables(h);
    if (rc != SBX_OK) return rc;
    RgbArgs rgb;
    rc = prepare_rgb(h, view, rgb);
    if (rc != SBX_OK) return rc;
    return shade_current(h, rgb);
}

// colour frame of the current field (all owned rows)
static int32_t shade_current(sbx_drainage *h, const RgbArgs &rgb)
{
    int n = 0;
    if (h->f64) {
        StepArgs<double> s;
        fill_step_args<double>(h, s, 0.0, rgb);
        s.l_first = 1; s.l_last = h->rows_local;
        CU(launch_shade<double>(s, h->batch, h->band_rows, h->stream, &n, h->use_tma));
    } else {
        StepArgs<float> s;
        fill_step_args<float>(h, s, 0.0, rgb);
        s.l_first = 1; s.l_last = h->rows_local;
        if (h->wgsl) {
            // WGSL handles carry radius in params, h_ref = 0; BubbleParam was filled from them at create
        }
        CU(launch_shade<float>(s, h->batch, h->band_rows, h->stream, &n, h->use_tma));
    }
    h->launches += n;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_download_rgb(sbx_drainage *h, void *host)
{
    CHECK_H(h);
    if (!host) return fail(SBX_ERR_INVALID, "host is null");
    if (!h->rgb_buf || h->rgb_format < 0) return fail(SBX_ERR_INVALID, "no colour frame has been produced yet");
    const size_t cells = (size_t)h->batch * h->rows_local * h->num_phi;
    CU(cudaMemcpyAsync(host, h->rgb_buf, cells * (h->rgb_format == SBX_OUT_RGBA8 ? 4 : 12), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

// One fused step with host buffers, pipelined over latitude chunks so that the H2D copy of chunk c+1, the
// kernels of chunk c and the D2H copies of chunk c-1 overlap (three streams, both copy engines busy: the step is
// PCIe-bound, so the round trip costs max(H2D, D2H) instead of their sum).  f32 in/out, CPU semantics, one
// bubble, whole sphere or one latitude band of a sharded sphere.  On a sharded handle every rank makes this call
// with its own band: the first and the last chunk are uploaded first, the neighbours' boundary rows of the INPUT
// field are exchanged once (sbx_comm_halo on the upload stream, 2 x num_phi x 4 bytes) while the other chunks
// are still crossing PCIe, and the two chunks that read a ghost row wait for that exchange only.
static int32_t ensure_ghosts(sbx_drainage *h)
{
    if (!h->comm || !h->ghost_stale) return SBX_OK;
    h->ghost_stale = false;
    return halo_exchange(h);
}

static int32_t quiesce_peers(sbx_drainage *h)
{
    return h->peer_enabled ? ensure_ghosts(h) : SBX_OK;
}

static int host_chunk_rows(const sbx_drainage *h)
{
    // ~32 MiB of f32 per chunk (504 rows at 16384 columns: a multiple of the kernel's band heights), >= 8 rows
    if (const char *e = getenv("SBX_HOST_CHUNK_ROWS")) {
        const int env = atoi(e);
        if (env >= 8) return env;
    }
    long long r = (32ll << 20) / ((long long)h->num_phi * 4);
    r = r / 56 * 56;
    if (r < 56) r = 56;
    return (int)r;
}

static int32_t step_host_pipelined(sbx_drainage *h, double dt, const sbx_view *view, const float *host_in,
                                   float *host_out_h, void *host_out_rgb)
{
    h->mm_valid = h->mm_kernel_valid = false;
    int32_t rc = ensure_tables(h);
    if (rc != SBX_OK) return rc;
    RgbArgs rgb;
    rc = prepare_rgb(h, view, rgb);
    if (rc != SBX_OK) return rc;
    const int N = h->num_phi, rows = h->rows_local;
    const int chunk_rows = host_chunk_rows(h);
    const int nchunks = (rows + chunk_rows - 1) / chunk_rows;
    const int L = nchunks - 1;
    if (!h->copy_in) {
        CU(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_all, cudaEventDisableTiming));
    }
    while ((int)h->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_in.push_back(a);
        h->ev_comp.push_back(b);
    }
    FrontTable ft;
    memset(&ft, 0, sizeof ft);
    ft.row_min = 1; ft.row_max = 0;
    if (h->tfe_enabled) { tfe_advance(h, dt); tfe_table(h, ft); }
    StepArgs<float> a;
    fill_step_args<float>(h, a, dt, rgb);
    a.fronts = ft;
    const int l_first = a.l_first, l_last = a.l_last;
    float *in = tbuf<float>(h, h->cur), *out = tbuf<float>(h, 1 - h->cur);
    const size_t px = rgb.mode == SBX_RGB_NONE ? 0 : (rgb.format == SBX_OUT_RGBA8 ? 4 : 12);
    // everything queued on the handle's stream so far must be visible to the copy streams
    CU(cudaEventRecord(h->ev_start, h->stream));
    CU(cudaStreamWaitEvent(h->copy_in, h->ev_start, 0));
    CU(cudaStreamWaitEvent(h->copy_out, h->ev_start, 0));
    auto upload_chunk = [&](int c) -> int32_t {
        const int r0 = c * chunk_rows, r1 = std::min(rows, r0 + chunk_rows);  // owned rows [r0, r1)
        CU(cudaMemcpyAsync(in + (size_t)(r0 + 1) * N, host_in + (size_t)r0 * N, (size_t)(r1 - r0) * N * 4,
                           cudaMemcpyHostToDevice, h->copy_in));
        CU(cudaEventRecord(h->ev_in[c], h->copy_in));
        return SBX_OK;
    };
    // upload order: 0, L, [halo rows of the input field], 1 .. L-1
    if ((rc = upload_chunk(0)) != SBX_OK) return rc;
    if (L > 0 && (rc = upload_chunk(L)) != SBX_OK) return rc;
    if (h->comm) {
        const char *err = nullptr;
        const size_t row_bytes = (size_t)N * 4;
        char *base = (char *)in;
        if (sbx_comm_halo(h->comm, base + row_bytes, base, base + (size_t)rows * row_bytes,
                          base + (size_t)(rows + 1) * row_bytes, row_bytes, h->copy_in, &err) != 0)
            return fail(SBX_ERR_COMM, "halo exchange failed: %s", err ? err : "?");
    }
    for (int c = 1; c < L; ++c)
        if ((rc = upload_chunk(c)) != SBX_OK) return rc;
    CU(cudaEventRecord(h->ev_all, h->copy_in));
    int n = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * chunk_rows, r1 = std::min(rows, r0 + chunk_rows);
        // chunk c reads the first row of chunk c + 1 too; in upload order ev_in[c + 1] (c + 1 < L) covers 0, L, the
        // halo rows and 1 .. c + 1; the last two chunks wait for everything
        CU(cudaStreamWaitEvent(h->stream, c + 1 < L ? h->ev_in[c + 1] : h->ev_all, 0));
        StepArgs<float> ac = a;
        ac.l_first = std::max(l_first, r0 + 1);
        ac.l_last = std::min(l_last, r1);
        // the pole rows are written by the launch that holds the ring next to them
        const bool top = c == 0 && h->row_begin == 0;
        const bool bottom = c == (rows - 2) / chunk_rows && h->row_end == h->num_theta;
        if (ac.l_first <= ac.l_last) {
            set_pole_fold<float>(h, ac, top, bottom);
            CU(launch_step<float>(ac, 1, ft.count > 0, false, top, bottom, h->band_rows, h->use_tma, h->stream, &n));
            commit_pole_fold<float>(h, ac);
        }
        CU(cudaEventRecord(h->ev_comp[c], h->stream));
        CU(cudaStreamWaitEvent(h->copy_out, h->ev_comp[c], 0));
        if (host_out_rgb && px)
            CU(cudaMemcpyAsync((char *)host_out_rgb + (size_t)r0 * N * px, (char *)h->rgb_buf + (size_t)r0 * N * px,
                               (size_t)(r1 - r0) * N * px, cudaMemcpyDeviceToHost, h->copy_out));
        if (host_out_h)
            CU(cudaMemcpyAsync(host_out_h + (size_t)r0 * N, out + (size_t)(r0 + 1) * N, (size_t)(r1 - r0) * N * 4,
                               cudaMemcpyDeviceToHost, h->copy_out));
    }
    h->launches += n;
    h->cur = 1 - h->cur;
    h->time += dt;
    h->ghost_stale = h->comm != nullptr;  // the new field's ghost rows are fetched when a device-resident step needs them
    h->stale_by_pull = false;             // (and a peer-halo step must not read the neighbour's rows before that exchange)
    CU(cudaEventRecord(h->ev_out, h->copy_out));
    CU(cudaStreamWaitEvent(h->stream, h->ev_out, 0));
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

// The same round trip for a foam batch (independent whole-sphere bubbles, BASELINE config 4): pipelined over groups of
// bubbles.  With offset storage (h_ref != 0) the uploaded absolute thickness is converted in place after it lands and
// the new field is converted back into the (now free) input half before it leaves, so the device keeps its
// sub-ulp-preserving representation and the host sees metres.
static int32_t step_host_pipelined_foam(sbx_drainage *h, double dt, const sbx_view *view, const float *host_in,
                                        float *host_out_h, void *host_out_rgb)
{
    h->mm_valid = h->mm_kernel_valid = false;
    int32_t rc = ensure_tables(h);
    if (rc != SBX_OK) return rc;
    RgbArgs rgb;
    rc = prepare_rgb(h, view, rgb);
    if (rc != SBX_OK) return rc;
    const int N = h->num_phi, rows = h->rows_local, B = h->batch;
    const size_t cells_b = (size_t)rows * N;  // owned cells per bubble
    int group = (int)std::max<size_t>(1, ((size_t)32 << 20) / (cells_b * 4));
    if (const char *e = getenv("SBX_HOST_CHUNK_BUBBLES")) group = std::max(1, atoi(e));
    const int nchunks = (B + group - 1) / group;
    if (!h->copy_in) {
        CU(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_all, cudaEventDisableTiming));
    }
    while ((int)h->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_in.push_back(a);
        h->ev_comp.push_back(b);
    }
    FrontTable ft;
    memset(&ft, 0, sizeof ft);
    ft.row_min = 1; ft.row_max = 0;
    if (h->tfe_enabled) { tfe_advance(h, dt); tfe_table(h, ft); }
    StepArgs<float> a;
    fill_step_args<float>(h, a, dt, rgb);
    a.fronts = ft;
    set_pole_fold<float>(h, a, true, true);  // every chunk of this step waits for the same arrival count (one per bubble)
    float *in = tbuf<float>(h, h->cur), *out = tbuf<float>(h, 1 - h->cur);
    const size_t px = rgb.mode == SBX_RGB_NONE ? 0 : (rgb.format == SBX_OUT_RGBA8 ? 4 : 12);
    const double *href = h->href_zero ? nullptr : h->href_dev;
    CU(cudaEventRecord(h->ev_start, h->stream));
    CU(cudaStreamWaitEvent(h->copy_in, h->ev_start, 0));
    CU(cudaStreamWaitEvent(h->copy_out, h->ev_start, 0));
    for (int c = 0; c < nchunks; ++c) {
        const int b0 = c * group, nb = std::min(group, B - b0);
        CU(cudaMemcpy2DAsync(in + (size_t)b0 * h->bubble_stride + N, (size_t)h->bubble_stride * 4, host_in + (size_t)b0 * cells_b,
                             cells_b * 4, cells_b * 4, (size_t)nb, cudaMemcpyHostToDevice, h->copy_in));
        CU(cudaEventRecord(h->ev_in[c], h->copy_in));
    }
    int n = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int b0 = c * group, nb = std::min(group, B - b0);
        const long long off = (long long)b0 * h->bubble_stride, cnt = (long long)nb * h->bubble_stride;
        CU(cudaStreamWaitEvent(h->stream, h->ev_in[c], 0));
        if (href) { CU((launch_convert_in<float, float>(in + off, in + off, cnt, h->bubble_stride, href + b0, h->stream))); ++n; }
        StepArgs<float> ac = a;
        ac.in = in + off;
        ac.out = out + off;
        ac.bub = a.bub + b0;
        ac.pole_count = a.pole_count + 2 * b0;  // per-bubble arrival counters; the partial-sum slots are reused chunk after chunk
        if (px) ac.rgb.out = (char *)h->rgb_buf + (size_t)b0 * cells_b * px;
        CU(launch_step<float>(ac, nb, ft.count > 0, false, true, true, h->band_rows, h->use_tma, h->stream, &n));
        const float *src_h = out + off;
        if (href && host_out_h) {  // metres for the host: into the input half, which this step no longer needs
            CU((launch_convert_out<float, float>(in + off, out + off, cnt, h->bubble_stride, href + b0, h->stream)));
            ++n;
            src_h = in + off;
        }
        CU(cudaEventRecord(h->ev_comp[c], h->stream));
        CU(cudaStreamWaitEvent(h->copy_out, h->ev_comp[c], 0));
        if (host_out_rgb && px)
            CU(cudaMemcpyAsync((char *)host_out_rgb + (size_t)b0 * cells_b * px, (char *)h->rgb_buf + (size_t)b0 * cells_b * px,
                               (size_t)nb * cells_b * px, cudaMemcpyDeviceToHost, h->copy_out));
        if (host_out_h)
            CU(cudaMemcpy2DAsync(host_out_h + (size_t)b0 * cells_b, cells_b * 4, src_h + N, (size_t)h->bubble_stride * 4, cells_b * 4,
                                 (size_t)nb, cudaMemcpyDeviceToHost, h->copy_out));
    }
    commit_pole_fold<float>(h, a);
    h->launches += n;
    h->cur = 1 - h->cur;
    h->time += dt;
    CU(cudaEventRecord(h->ev_out, h->copy_out));
    CU(cudaStreamWaitEvent(h->stream, h->ev_out, 0));
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_step_host(sbx_drainage *h, double dt, int32_t nsteps, const sbx_view *view,
                                          const void *host_in, int32_t in_dtype, void *host_out_h, int32_t out_dtype,
                                          void *host_out_rgb)
{
    CHECK_H(h);
    int32_t rc = host_in ? quiesce_peers(h) : SBX_OK;
    if (rc != SBX_OK) return rc;
    // (on a sharded handle the choice must be the same on every rank -- the two paths issue different NCCL calls -- so
    // it is made from the global grid, not from this rank's band height)
    const bool tall = h->comm ? h->num_theta >= 9 * sbx_comm_size(h->comm) : h->rows_local >= 8;
    if (host_in && nsteps == 1 && !h->wgsl && !h->f64 && h->batch == 1 && h->href_zero &&
        !(view && view->mode == SBX_RGB_SPECTRAL) &&
        in_dtype == SBX_F32 && (!host_out_h || out_dtype == SBX_F32) && h->num_phi % 4 == 0 && tall &&
        !getenv("SBX_NO_HOST_PIPELINE"))
        return step_host_pipelined(h, dt, view, (const float *)host_in, (float *)host_out_h, host_out_rgb);
    if (host_in && nsteps == 1 && !h->wgsl && !h->f64 && h->batch > 1 && !h->comm && h->rows_local == h->num_theta &&
        !(view && view->mode == SBX_RGB_SPECTRAL) && in_dtype == SBX_F32 && (!host_out_h || out_dtype == SBX_F32) &&
        h->num_phi % 4 == 0 && !getenv("SBX_NO_HOST_PIPELINE"))
        return step_host_pipelined_foam(h, dt, view, (const float *)host_in, (float *)host_out_h, host_out_rgb);
    if (host_in) rc = sbx_drainage_upload(h, host_in, in_dtype);
    if (rc != SBX_OK) return rc;
    rc = do_steps(h, dt, nsteps, view);
    if (rc != SBX_OK) return rc;
    if (host_out_rgb) {
        if (!view || view->mode == SBX_RGB_NONE) return fail(SBX_ERR_INVALID, "host_out_rgb needs a colour view");
        const size_t cells = (size_t)h->batch * h->rows_local * h->num_phi;
        CU(cudaMemcpyAsync(host_out_rgb, h->rgb_buf, cells * (h->rgb_format == SBX_OUT_RGBA8 ? 4 : 12), cudaMemcpyDeviceToHost, h->stream));
    }
    if (host_out_h) return sbx_drainage_download(h, host_out_h, out_dtype);
    CU(cudaStreamSynchronize(h->stream));
    return SBX_OK;
}

// ------------------------------------------------------------------ queries
// get_thickness (drainage.rs:516-552): index arithmetic on the host, the blend itself on the device.  Enqueues the
// sample kernel and the 8-byte copy into scalar_host[0]; the caller synchronises.
static int32_t enqueue_sample(sbx_drainage *h, int32_t bubble, double theta, double phi)
{
    if (bubble < 0 || bubble >= h->batch) return fail(SBX_ERR_INVALID, "bubble %d out of range", bubble);
    // index arithmetic of get_thickness (drainage.rs:516-540); the blend itself runs on the device
    const int NT = h->num_theta, NP = h->num_phi;
    const double dth = M_PI / (double)(NT - 1), dph = 2.0 * M_PI / (double)NP;
    const double theta_clamped = theta < 0.0 ? 0.0 : (theta > M_PI ? M_PI : theta);
    const double two_pi = 2.0 * M_PI;
    double phi_n = fmod(phi, two_pi);
    if (phi_n < 0.0) phi_n += two_pi;
    const double tc = theta_clamped / dth, pc = phi_n / dph;
    size_t ti = (size_t)floor(tc);
    if (ti > (size_t)(NT - 2)) ti = (size_t)(NT - 2);
    const size_t pj = (size_t)floor(pc) % (size_t)NP;
    const size_t pj_hi = (pj + 1) % (size_t)NP;
    const double tf = tc - (double)ti, pf = pc - (double)pj;
    // the lower of the two rows may be the bottom ghost row when it is kept current by the per-step NCCL exchange
    // (with the in-kernel peer halo the ghost rows are not maintained at all: the kernels read the neighbour's rows in place)
    const bool ghost_ok = h->comm && !h->peer_enabled && !h->ghost_stale && h->row_end < h->num_theta;
    if ((int)ti < h->row_begin || (int)ti + 1 > h->row_end || ((int)ti + 1 == h->row_end && !ghost_ok))
        return fail(SBX_ERR_INVALID, "theta rows %zu,%zu are not owned by this band [%d,%d)", ti, ti + 1, h->row_begin, h->row_end);
    const long long base = (long long)bubble * h->bubble_stride;
    const long long r0 = base + (long long)((int)ti - h->row_begin + 1) * NP, r1 = r0 + NP;
    if (h->f64) CU(launch_sample<double>(tbuf<double>(h, h->cur), r0 + pj, r0 + pj_hi, r1 + pj, r1 + pj_hi, tf, pf, 0.0, h->scalar_dev, h->stream));
    else CU(launch_sample<float>(tbuf<float>(h, h->cur), r0 + pj, r0 + pj_hi, r1 + pj, r1 + pj_hi, tf, pf, h->b_href[bubble], h->scalar_dev, h->stream));
    h->launches++;
    CU(cudaMemcpyAsync(h->scalar_host, h->scalar_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_get_thickness(sbx_drainage *h, int32_t bubble, double theta, double phi, double *out)
{
    CHECK_H(h);
    if (!out) return fail(SBX_ERR_INVALID, "out is null");
    int32_t rc = enqueue_sample(h, bubble, theta, phi);
    if (rc != SBX_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));
    *out = h->scalar_host[0];
    return SBX_OK;
}

// min / max of the current field in two halves around ONE synchronisation of the handle's stream:
// enqueue_minmax() queues whatever has to run (nothing if cached; a 16-byte-per-bubble fetch if the last step kernel
// left the extremes behind; else one reduction pass + the fetch of its partials), finish_minmax() decodes afterwards.
static int32_t enqueue_minmax(sbx_drainage *h, int *pending)
{
    // The reference's frame loop asks for min, max and has_critical_region one after the other (pipeline.rs:1298-1312,
    // three full passes on the CPU); here the first query after the field changed runs ONE pass over the owned rows
    // of all bubbles and the others reuse it.
    *pending = 0;
    if (!h->mm_valid && h->stepped_since_mm) {
        h->stepped_since_mm = false;
        h->mm_auto_misses = 0;
        if (++h->mm_auto_hits >= 2) h->mm_auto_on = true;
    }
    if (h->mm_valid) return SBX_OK;
    if (h->mm_kernel_valid) {  // the last step kernel left them behind: 16 bytes per bubble to fetch
        const size_t wb = h->f64 ? 8 : 4, set_bytes = (size_t)h->batch * 2 * wb;
        CU(cudaMemcpyAsync(h->mm_track_host, h->mm_track[h->mm_track_set], set_bytes, cudaMemcpyDeviceToHost, h->stream));
        *pending = 1;
        return SBX_OK;
    }
    const int nb = sbx_drainage::kMinmaxBlocks;
    const long long n = (long long)h->rows_local * h->num_phi;  // owned rows only (ghost rows are skipped)
    const double *href = h->href_zero || h->f64 ? nullptr : h->href_dev;
    if (h->f64) CU(launch_minmax<double>(tbuf<double>(h, h->cur), h->batch, h->bubble_stride, h->num_phi, n, href, h->mm_part, nb, h->stream));
    else CU(launch_minmax<float>(tbuf<float>(h, h->cur), h->batch, h->bubble_stride, h->num_phi, n, href, h->mm_part, nb, h->stream));
    h->launches++;
    CU(cudaMemcpyAsync(h->mm_host, h->mm_part, 2 * nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    *pending = 2;
    return SBX_OK;
}

static void finish_minmax(sbx_drainage *h, int pending)
{
    if (pending == 1) {
        double mn = INFINITY, mx = -INFINITY;
        for (int b = 0; b < h->batch; ++b) {
            double lo, hi;
            if (h->f64) {
                auto dec = [](unsigned long long u) { u = (u >> 63) ? u & 0x7fffffffffffffffull : ~u; double d; memcpy(&d, &u, 8); return d; };
                lo = dec(((unsigned long long *)h->mm_track_host)[2 * b]);
                hi = dec(((unsigned long long *)h->mm_track_host)[2 * b + 1]);
            } else {
                auto dec = [](unsigned int u) { u = (u >> 31) ? u & 0x7fffffffu : ~u; float f; memcpy(&f, &u, 4); return (double)f; };
                lo = dec(((unsigned int *)h->mm_track_host)[2 * b]) + h->b_href[b];
                hi = dec(((unsigned int *)h->mm_track_host)[2 * b + 1]) + h->b_href[b];
            }
            mn = fmin(mn, lo);
            mx = fmax(mx, hi);
        }
        h->mm_cached[0] = mn;
        h->mm_cached[1] = mx;
        h->mm_valid = true;
    } else if (pending == 2) {
        const int nb = sbx_drainage::kMinmaxBlocks;
        double mn = INFINITY, mx = -INFINITY;
        for (int i = 0; i < nb; ++i) { mn = fmin(mn, h->mm_host[2 * i]); mx = fmax(mx, h->mm_host[2 * i + 1]); }
        h->mm_cached[0] = mn;
        h->mm_cached[1] = mx;
        h->mm_valid = true;
    }
}

extern "C" int32_t sbx_drainage_minmax(sbx_drainage *h, double *min_out, double *max_out)
{
    CHECK_H(h);
    int pending = 0;
    int32_t rc = enqueue_minmax(h, &pending);
    if (rc != SBX_OK) return rc;
    if (pending) {
        CU(cudaStreamSynchronize(h->stream));
        finish_minmax(h, pending);
    }
    if (min_out) *min_out = h->mm_cached[0];
    if (max_out) *max_out = h->mm_cached[1];
    return SBX_OK;
}

// The values RenderPipeline::update reads after every step (pipeline.rs:1298-1312: min_thickness, max_thickness,
// get_thickness(pi/2, 0), has_critical_region, current_time) behind ONE synchronisation instead of one per query.
extern "C" int32_t sbx_drainage_frame_stats(sbx_drainage *h, int32_t bubble, double theta, double phi, sbx_frame_stats *out)
{
    CHECK_H(h);
    if (!out) return fail(SBX_ERR_INVALID, "out is null");
    int pending = 0;
    int32_t rc = enqueue_minmax(h, &pending);
    if (rc == SBX_OK) rc = enqueue_sample(h, bubble, theta, phi);
    if (rc != SBX_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));
    finish_minmax(h, pending);
    out->min_thickness = h->mm_cached[0];
    out->max_thickness = h->mm_cached[1];
    out->thickness = h->scalar_host[0];
    out->time = h->time;
    out->has_critical = h->mm_cached[0] < (h->wgsl ? (double)h->wp.critical_thickness : h->critical) ? 1 : 0;  // drainage.rs:565-567
    out->_pad = 0;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_track_minmax(sbx_drainage *h, int32_t mode)
{
    if (!h) return fail(SBX_ERR_INVALID, "null handle");
    if (mode < 0 || mode > 2) return fail(SBX_ERR_INVALID, "mode must be 0 (off), 1 (on) or 2 (auto)");
    h->mm_track_mode = mode;
    h->mm_auto_on = false;
    h->mm_auto_hits = h->mm_auto_misses = 0;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_has_critical(sbx_drainage *h, int32_t *out)
{
    double mn, mx;
    int32_t rc = sbx_drainage_minmax(h, &mn, &mx);
    if (rc != SBX_OK) return rc;
    if (out) *out = mn < (h->wgsl ? (double)h->wp.critical_thickness : h->critical) ? 1 : 0;  // drainage.rs:565-567
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_critical_thickness(sbx_drainage *h, double *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = h->wgsl ? (double)h->wp.critical_thickness : h->critical;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_current_time(sbx_drainage *h, double *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = h->time;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_reset(sbx_drainage *h, double h0_m)
{
    CHECK_H(h);
    h->mm_valid = h->mm_kernel_valid = false;
    { int32_t rcq = quiesce_peers(h); if (rcq != SBX_OK) return rcq; }  // a neighbour may still be reading this band's rows
    std::fill(h->b_h0.begin(), h->b_h0.end(), h0_m);
    if (!h->f64 && !h->href_zero) {
        // offset storage follows the new uniform level, so the reset field is exact (delta == 0)
        h->h_ref = h0_m;
        std::fill(h->b_href.begin(), h->b_href.end(), h0_m);
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaMemcpy(h->href_dev, h->b_href.data(), (size_t)h->batch * sizeof(double), cudaMemcpyHostToDevice));
        int32_t rcb = upload_bubbles<float>(h);
        if (rcb != SBX_OK) return rcb;
    }
    int32_t rc = h->f64 ? fill_initial<double>(h) : fill_initial<float>(h);
    if (rc != SBX_OK) return rc;
    if (h->wgsl)
        for (int i = 0; i < 2; ++i) { CU(launch_fill<float>(h->conc[i], h->total_elems, 0.5f, h->stream)); h->launches++; }
    h->cur = 0;
    h->time = 0.0;
    h->fronts.clear();
    h->next_spawn = 0.5;
    // sharded: every rank resets (like upload); the exchange orders the neighbours' next in-place reads after this fill
    return h->comm ? halo_exchange(h) : SBX_OK;
}

extern "C" int32_t sbx_drainage_set_marginal_regeneration(sbx_drainage *h, int32_t enabled)
{
    if (!h) return fail(SBX_ERR_INVALID, "null handle");
    h->tfe_enabled = enabled != 0;
    if (!enabled) h->fronts.clear();
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_marginal_regeneration(sbx_drainage *h, int32_t *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = h->tfe_enabled ? 1 : 0;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_set_tfe_ratio(sbx_drainage *h, double ratio)
{
    if (!h) return fail(SBX_ERR_INVALID, "null handle");
    h->tfe_ratio = ratio < 0.5 ? 0.5 : (ratio > 0.95 ? 0.95 : ratio);  // clamp(0.5, 0.95)
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_tfe_ratio(sbx_drainage *h, double *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = h->tfe_ratio;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_active_tfe_count(sbx_drainage *h, int32_t *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = (int32_t)h->fronts.size();
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_fronts(sbx_drainage *h, sbx_tfe_front *out8, int32_t *count)
{
    if (!h || !out8 || !count) return fail(SBX_ERR_INVALID, "null argument");
    *count = (int32_t)h->fronts.size();
    for (size_t i = 0; i < h->fronts.size(); ++i) out8[i] = h->fronts[i];
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_get_state(sbx_drainage *h, sbx_tfe_state *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    memset(out, 0, sizeof *out);
    out->time = h->time;
    out->next_spawn = h->next_spawn;
    out->count = (int32_t)h->fronts.size();
    for (size_t i = 0; i < h->fronts.size() && i < (size_t)kMaxFronts; ++i) out->fronts[i] = h->fronts[i];
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_set_state(sbx_drainage *h, const sbx_tfe_state *st)
{
    if (!h || !st) return fail(SBX_ERR_INVALID, "null argument");
    if (st->count < 0 || st->count > kMaxFronts) return fail(SBX_ERR_INVALID, "count %d out of range [0,%d]", st->count, kMaxFronts);
    if (!std::isfinite(st->time) || !std::isfinite(st->next_spawn)) return fail(SBX_ERR_INVALID, "time / next_spawn must be finite");
    for (int i = 0; i < st->count; ++i) {
        const sbx_tfe_front &f = st->fronts[i];
        if (!std::isfinite(f.phi_center) || !std::isfinite(f.extent) || !std::isfinite(f.width) || !std::isfinite(f.velocity))
            return fail(SBX_ERR_INVALID, "front %d has a non-finite field", i);
    }
    h->time = st->time;
    h->next_spawn = st->next_spawn;
    h->fronts.assign(st->fronts, st->fronts + st->count);
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_update_params(sbx_drainage *h, const sbx_drainage_params *p)
{
    if (!h || !p) return fail(SBX_ERR_INVALID, "null argument");
    if (!h->wgsl) return fail(SBX_ERR_INVALID, "update_params applies to SBX_SEM_WGSL handles");
    if (p->grid_width != (uint32_t)h->num_phi || p->grid_height != (uint32_t)h->num_theta)
        return fail(SBX_ERR_INVALID, "grid size is fixed at creation (%dx%d)", h->num_phi, h->num_theta);
    if (p->bubble_radius != h->wp.bubble_radius) h->tables_dirty = true;  // BubbleParam.radius feeds the view cosine
    h->wp = *p;
    std::fill(h->b_radius.begin(), h->b_radius.end(), (double)p->bubble_radius);
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_params_get(sbx_drainage *h, sbx_drainage_params *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = h->wp;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_set_physics(sbx_drainage *h, double density, double gravity, double viscosity,
                                            double diffusion, double critical_thickness_m, double radius_m)
{
    if (!h) return fail(SBX_ERR_INVALID, "null handle");
    if (h->wgsl) return fail(SBX_ERR_INVALID, "use sbx_drainage_update_params for SBX_SEM_WGSL handles");
    if (!(radius_m > 0.0) || !(viscosity > 0.0)) return fail(SBX_ERR_INVALID, "radius and viscosity must be > 0");
    const double old_r = h->radius;
    h->density = density; h->gravity = gravity; h->viscosity = viscosity;
    h->diffusion = diffusion; h->critical = critical_thickness_m; h->radius = radius_m;
    for (auto &r : h->b_radius) if (r == old_r) r = radius_m;
    h->tables_dirty = true;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_device_ptr(sbx_drainage *h, int32_t kind, void **out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    switch (kind) {
    case SBX_PTR_THICKNESS:
        h->mm_valid = h->mm_kernel_valid = false;  // the caller may write through the pointer
        *out = (char *)h->buf[h->cur] + (size_t)h->num_phi * h->esz;
        return SBX_OK;
    case SBX_PTR_CONCENTRATION:
        if (!h->wgsl) return fail(SBX_ERR_INVALID, "concentration exists only with SBX_SEM_WGSL");
        *out = h->conc[h->cur] + h->num_phi;
        return SBX_OK;
    case SBX_PTR_RGB:
        if (!h->rgb_buf) return fail(SBX_ERR_INVALID, "no colour frame has been produced yet");
        *out = h->rgb_buf;
        return SBX_OK;
    case SBX_PTR_CAUSTICS:
        if (!h->caustic_buf) return fail(SBX_ERR_INVALID, "no caustic map has been produced yet");
        *out = h->caustic_buf;
        return SBX_OK;
    }
    return fail(SBX_ERR_INVALID, "bad pointer kind %d", kind);
}

extern "C" int32_t sbx_drainage_set_spectrum(sbx_drainage *h, const float *wl, int32_t n)
{
    if (!h || !wl) return fail(SBX_ERR_INVALID, "null argument");
    if (n < 1 || n > 256) return fail(SBX_ERR_INVALID, "n must be in [1,256] (got %d)", n);
    for (int k = 0; k < n; ++k)
        if (!(wl[k] > 0.f) || !std::isfinite(wl[k])) return fail(SBX_ERR_INVALID, "wavelength %d is not a positive number", k);
    h->spec_wl.assign(wl, wl + n);
    h->n_spec = -1;  // rebuilt on next use
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_launch_count(sbx_drainage *h, int64_t *out)
{
    if (!h || !out) return fail(SBX_ERR_INVALID, "null argument");
    *out = h->launches;
    return SBX_OK;
}

// ------------------------------------------------------------------ caustic map (thickness-buffer consumer)
extern "C" int32_t sbx_caustic_params_default(sbx_caustic_params *p)
{
    if (!p) return fail(SBX_ERR_INVALID, "p is null");
    memset(p, 0, sizeof *p);  // src/render/caustics.rs:58-79
    p->grid_width = 128; p->grid_height = 64; p->refractive_index = 1.33f; p->film_thickness_scale = 1e6f;
    p->light_dir_y = -1.0f; p->light_intensity = 1.0f; p->focal_length = 0.1f; p->caustic_intensity = 2.0f;
    p->caustic_sharpness = 1.5f; p->branch_threshold = 0.001f; p->ground_y = -0.1f; p->bubble_radius = 0.025f;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_caustics(sbx_drainage *h, const sbx_caustic_params *p, float *host_out)
{
    CHECK_H(h);
    if (!p) return fail(SBX_ERR_INVALID, "p is null");
    { int32_t rcg = ensure_ghosts(h); if (rcg != SBX_OK) return rcg; }
    const size_t cells = (size_t)h->batch * h->rows_local * h->num_phi;
    if (!h->caustic_buf) CU(cudaMalloc(&h->caustic_buf, cells * sizeof(float)));
    sbx_caustic_params q = *p;
    q.grid_width = (uint32_t)h->num_phi;
    q.grid_height = (uint32_t)h->num_theta;
    const double *href = h->href_zero ? nullptr : h->href_dev;
    if (h->f64) CU(launch_caustics<double>(tbuf<double>(h, h->cur), h->bubble_stride, h->caustic_buf, h->num_phi, h->num_theta,
                                           h->row_begin, h->rows_local, h->batch, href, q, h->stream, h->use_tma));
    else CU(launch_caustics<float>(tbuf<float>(h, h->cur), h->bubble_stride, h->caustic_buf, h->num_phi, h->num_theta,
                                   h->row_begin, h->rows_local, h->batch, href, q, h->stream, h->use_tma));
    h->launches++;
    if (host_out) {
        CU(cudaMemcpyAsync(host_out, h->caustic_buf, cells * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_sample_uv(sbx_drainage *h, int32_t bubble, const float *uv, size_t n, float *out_h, float *out_grad)
{
    CHECK_H(h);
    if (!uv) return fail(SBX_ERR_INVALID, "uv is null");
    if (bubble < 0 || bubble >= h->batch) return fail(SBX_ERR_INVALID, "bubble %d out of range", bubble);
    if (h->rows_local != h->num_theta) return fail(SBX_ERR_UNSUPPORTED, "uv sampling needs a whole-sphere handle");
    if (n == 0) return SBX_OK;
    float *duv = nullptr, *dh = nullptr, *dg = nullptr;
    cudaError_t e = cudaMalloc(&duv, n * 8);
    if (e == cudaSuccess && out_h) e = cudaMalloc(&dh, n * 4);
    if (e == cudaSuccess && out_grad) e = cudaMalloc(&dg, n * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(duv, uv, n * 8, cudaMemcpyHostToDevice, h->stream);
    const size_t off = (size_t)bubble * h->bubble_stride + h->num_phi;
    if (e == cudaSuccess) {
        if (h->f64) e = launch_sample_uv<double>(tbuf<double>(h, h->cur) + off, h->num_phi, h->num_theta, 0.f, duv, n, dh, dg, h->stream);
        else e = launch_sample_uv<float>(tbuf<float>(h, h->cur) + off, h->num_phi, h->num_theta, (float)h->b_href[bubble], duv, n, dh, dg, h->stream);
        h->launches++;
    }
    if (e == cudaSuccess && dh) e = cudaMemcpyAsync(out_h, dh, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && dg) e = cudaMemcpyAsync(out_grad, dg, n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(duv); cudaFree(dh); cudaFree(dg);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "uv sampling failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_sample_normal(sbx_drainage *h, int32_t bubble, const float *normals, size_t n, float scale, float *out_h)
{
    CHECK_H(h);
    if (!normals || !out_h) return fail(SBX_ERR_INVALID, "normals/out_h is null");
    if (bubble < 0 || bubble >= h->batch) return fail(SBX_ERR_INVALID, "bubble %d out of range", bubble);
    if (h->rows_local != h->num_theta) return fail(SBX_ERR_UNSUPPORTED, "normal sampling needs a whole-sphere handle");
    if (n == 0) return SBX_OK;
    float *dn = nullptr, *dh = nullptr;
    cudaError_t e = cudaMalloc(&dn, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&dh, n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dn, normals, n * 12, cudaMemcpyHostToDevice, h->stream);
    const size_t off = (size_t)bubble * h->bubble_stride + h->num_phi;
    if (e == cudaSuccess) {
        if (h->f64) e = launch_sample_normal<double>(tbuf<double>(h, h->cur) + off, h->num_phi, h->num_theta, 0.f, dn, n, scale, dh, h->stream);
        else e = launch_sample_normal<float>(tbuf<float>(h, h->cur) + off, h->num_phi, h->num_theta, (float)h->b_href[bubble], dn, n, scale, dh, h->stream);
        h->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_h, dh, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dn); cudaFree(dh);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "normal sampling failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

// ------------------------------------------------------------------ LUT / calculator entry points
static int32_t pick_device(int32_t device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SBX_ERR_CUDA, "no usable CUDA device (%s); libsbx has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SBX_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    return SBX_OK;
}

extern "C" int32_t sbx_lut_build_spectral(float n, float intensity, const float *wl, int32_t nl, int32_t T, int32_t A,
                                          float max_nm, uint8_t *out, int32_t device)
{
    if (!wl || !out || nl < 1 || T < 2 || A < 2) return fail(SBX_ERR_INVALID, "bad LUT arguments");
    int32_t rc = pick_device(device);
    if (rc != SBX_OK) return rc;
    uint8_t *dev = nullptr;
    rc = build_lut_device(n, intensity, wl, nl, T, A, max_nm, &dev, 0);
    if (rc != SBX_OK) return rc;
    cudaError_t e = cudaMemcpy(out, dev, (size_t)T * A * 4, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "LUT download failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

extern "C" int32_t sbx_lut_build(float n, float intensity, uint8_t *out65536, int32_t device)
{
    return sbx_lut_build_spectral(n, intensity, kLutWl7, 7, 256, 64, 2000.f, out65536, device);
}

extern "C" int32_t sbx_lut_shade_points(const uint8_t *lut_rgba, int32_t T, int32_t A, float max_nm, const float *d_nm,
                                        const float *cos_theta, size_t n, float intensity, float *rgb_out, int32_t device)
{
    if (!lut_rgba || !d_nm || !cos_theta || !rgb_out || T < 2 || A < 2 || !(max_nm > 0.f))
        return fail(SBX_ERR_INVALID, "bad arguments");
    int32_t rc = pick_device(device);
    if (rc != SBX_OK) return rc;
    if (n == 0) return SBX_OK;
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    float *dd = nullptr, *dc = nullptr, *dr = nullptr;
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
    cudaError_t e = cudaMallocArray(&arr, &ch, (size_t)T, (size_t)A);
    if (e == cudaSuccess) e = cudaMemcpy2DToArray(arr, 0, 0, lut_rgba, (size_t)T * 4, (size_t)T * 4, (size_t)A, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    }
    if (e == cudaSuccess) e = cudaMalloc(&dd, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dc, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dr, n * 12);
    if (e == cudaSuccess) e = cudaMemcpy(dd, d_nm, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc, cos_theta, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_lut_shade_points(tex, 1.0f / max_nm, intensity, dd, dc, n, dr, 0);
    if (e == cudaSuccess) e = cudaMemcpy(rgb_out, dr, n * 12, cudaMemcpyDeviceToHost);
    cudaFree(dd); cudaFree(dc); cudaFree(dr);
    if (tex) cudaDestroyTextureObject(tex);
    if (arr) cudaFreeArray(arr);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "LUT shading failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

// ------------------------------------------------------------------ headless bubble frame + PNG export (rows f3/f4)
extern "C" int32_t sbx_frame_params_default(sbx_frame_params *p)
{
    if (!p) return fail(SBX_ERR_INVALID, "p is null");
    memset(p, 0, sizeof *p);
    p->width = p->height = 256;       // tests/screenshot_tests.rs: 256 x 256 frames
    p->cam_distance = 0.2f;           // Camera::new, camera.rs:49-60
    p->cam_yaw = 0.0f;
    p->cam_pitch = 0.3f;
    p->cam_fov_rad = 45.0f * (float)(M_PI / 180.0);
    p->radius = 0.025f;
    p->base_thickness_nm = 500.0f;    // BubbleUniform::default, pipeline.rs:69-112
    p->film_time = 0.0f;
    p->swirl_intensity = 1.0f;
    p->drainage_speed = 0.5f;
    p->pattern_scale = 1.0f;
    p->interference_intensity = 4.0f;
    p->base_alpha = 0.3f;
    p->edge_alpha = 0.6f;
    p->edge_smoothing_mode = 1;
    p->background[0] = 0.1f; p->background[1] = 0.1f; p->background[2] = 0.15f;
    p->lut_max_nm = 2000.0f;
    return SBX_OK;
}

static int32_t frame_args_from(const sbx_frame_params *p, FrameArgs &a)
{
    if (!p) return fail(SBX_ERR_INVALID, "frame params are null");
    if (p->width == 0 || p->height == 0 || p->width > 16384 || p->height > 16384)
        return fail(SBX_ERR_INVALID, "Invalid dimensions: %ux%u", p->width, p->height);
    if (!(p->radius > 0.f) || !(p->cam_distance > p->radius)) return fail(SBX_ERR_INVALID, "camera must be outside the bubble");
    if (!(p->lut_max_nm > 0.f) || !(p->cam_fov_rad > 0.f) || p->edge_smoothing_mode > 2u)
        return fail(SBX_ERR_INVALID, "bad frame parameters");
    memset(&a, 0, sizeof a);
    a.width = (int)p->width;
    a.height = (int)p->height;
    // Camera::position (camera.rs:63-68), Mat4::look_at_rh(position, origin, Y), perspective_rh(fov, aspect)
    const float d = p->cam_distance, cp = cosf(p->cam_pitch), sp = sinf(p->cam_pitch);
    const float cam[3] = {d * cp * sinf(p->cam_yaw), d * sp, d * cp * cosf(p->cam_yaw)};
    const float cl = sqrtf(cam[0] * cam[0] + cam[1] * cam[1] + cam[2] * cam[2]);
    const float f[3] = {-cam[0] / cl, -cam[1] / cl, -cam[2] / cl};
    float r[3] = {f[1] * 0.f - f[2] * 1.f, f[2] * 0.f - f[0] * 0.f, f[0] * 1.f - f[1] * 0.f};  // cross(f, Y)
    const float rl = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (!(rl > 0.f)) return fail(SBX_ERR_INVALID, "camera looks along the up axis");
    for (int k = 0; k < 3; ++k) r[k] /= rl;
    const float u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};  // cross(r, f)
    for (int k = 0; k < 3; ++k) { a.cam[k] = cam[k]; a.fwd[k] = f[k]; a.right[k] = r[k]; a.up[k] = u[k]; a.bg[k] = p->background[k]; }
    a.tan_half_y = tanf(0.5f * p->cam_fov_rad);
    a.tan_half_x = a.tan_half_y * ((float)p->width / (float)p->height);
    a.radius = p->radius;
    a.cam2 = cam[0] * cam[0] + cam[1] * cam[1] + cam[2] * cam[2];
    a.base_nm = p->base_thickness_nm;
    a.film_time = p->film_time;
    a.swirl = p->swirl_intensity;
    a.drainage_speed = p->drainage_speed;
    a.pattern_scale = p->pattern_scale;
    a.intensity = p->interference_intensity;
    a.base_alpha = p->base_alpha;
    a.edge_alpha = p->edge_alpha;
    a.edge_mode = (int)p->edge_smoothing_mode;
    a.inv_max_nm = 1.0f / p->lut_max_nm;
    return SBX_OK;
}

extern "C" int32_t sbx_render_bubble_frame(const sbx_frame_params *p, const uint8_t *lut_rgba, int32_t T, int32_t A,
                                           uint8_t *rgba_out, int32_t device)
{
    if (!lut_rgba || !rgba_out || T < 2 || A < 2) return fail(SBX_ERR_INVALID, "bad arguments");
    FrameArgs a;
    int32_t rc = frame_args_from(p, a);
    if (rc != SBX_OK) return rc;
    rc = pick_device(device);
    if (rc != SBX_OK) return rc;
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    uint32_t *out = nullptr;
    const size_t bytes = (size_t)a.width * a.height * 4;
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
    cudaError_t e = cudaMallocArray(&arr, &ch, (size_t)T, (size_t)A);
    if (e == cudaSuccess) e = cudaMemcpy2DToArray(arr, 0, 0, lut_rgba, (size_t)T * 4, (size_t)T * 4, (size_t)A, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td{};  // sampler at pipeline.rs:350-359
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    }
    if (e == cudaSuccess) e = cudaMalloc(&out, bytes);
    a.lut = tex;
    a.out = out;
    if (e == cudaSuccess) e = launch_bubble_frame(a, 0);
    if (e == cudaSuccess) e = cudaMemcpy(rgba_out, out, bytes, cudaMemcpyDeviceToHost);
    cudaFree(out);
    if (tex) cudaDestroyTextureObject(tex);
    if (arr) cudaFreeArray(arr);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "bubble frame failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_render_frame(sbx_drainage *h, int32_t bubble, const sbx_frame_params *p, double n_film,
                                             uint8_t *rgba_out)
{
    CHECK_H(h);
    if (!rgba_out) return fail(SBX_ERR_INVALID, "rgba_out is null");
    if (bubble < 0 || bubble >= h->batch) return fail(SBX_ERR_INVALID, "bubble %d out of range (batch %d)", bubble, h->batch);
    if (h->row_begin != 0 || h->rows_local != h->num_theta)
        return fail(SBX_ERR_UNSUPPORTED, "render_frame needs a whole-sphere handle (this one owns rows %d..%d of %d)",
                    h->row_begin, h->row_end, h->num_theta);
    if (!(n_film > 0.0)) return fail(SBX_ERR_INVALID, "n_film must be > 0");
    FrameArgs a;
    int32_t rc = frame_args_from(p, a);
    if (rc != SBX_OK) return rc;
    rc = ensure_lut(h, (float)n_film);
    if (rc != SBX_OK) return rc;
    a.inv_max_nm = 1.0f / h->lut_max_nm;
    a.lut = h->lut_tex;
    // local row 0 of the buffer is the ghost row: the field starts one row in
    const char *base = static_cast<const char *>(h->buf[h->cur]);
    a.field = base + ((size_t)bubble * h->bubble_stride + (size_t)h->num_phi) * h->esz;
    a.field_f64 = h->f64 ? 1 : 0;
    a.num_theta = h->num_theta;
    a.num_phi = h->num_phi;
    a.h_ref = h->b_href.empty() ? h->h_ref : h->b_href[bubble];
    const size_t bytes = (size_t)a.width * a.height * 4;
    uint32_t *out = nullptr;
    CU(cudaMalloc(&out, bytes));
    a.out = out;
    cudaError_t e = launch_bubble_frame(a, h->stream);
    h->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgba_out, out, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(out);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "bubble frame failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

// export_frame (src/export/image_export.rs:42-76): 8-bit RGBA PNG, one IDAT chunk, filter 0 on every scanline.
static void png_chunk(std::vector<uint8_t> &o, const char type[4], const uint8_t *data, size_t n)
{
    const uint32_t len = (uint32_t)n;
    const uint8_t be[4] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len};
    o.insert(o.end(), be, be + 4);
    const size_t at = o.size();
    o.insert(o.end(), type, type + 4);
    if (n) o.insert(o.end(), data, data + n);
    const uint32_t crc = (uint32_t)crc32(0L, o.data() + at, (uInt)(n + 4));
    const uint8_t cb[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    o.insert(o.end(), cb, cb + 4);
}

extern "C" int32_t sbx_export_frame_png(const char *path, uint32_t width, uint32_t height, const uint8_t *rgba, size_t len)
{
    if (!path || !rgba) return fail(SBX_ERR_INVALID, "null argument");
    if (width == 0 || height == 0) return fail(SBX_ERR_INVALID, "Invalid dimensions: %ux%u", width, height);
    const size_t expect = (size_t)width * height * 4;
    if (len != expect || expect > 0x7fffffffull)
        return fail(SBX_ERR_INVALID, "Data length %zu doesn't match expected %zu (%ux%ux4)", len, expect, width, height);
    std::vector<uint8_t> raw((size_t)height * ((size_t)width * 4 + 1));
    for (uint32_t y = 0; y < height; ++y) {
        uint8_t *row = raw.data() + (size_t)y * ((size_t)width * 4 + 1);
        row[0] = 0;  // filter type None
        memcpy(row + 1, rgba + (size_t)y * width * 4, (size_t)width * 4);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(SBX_ERR_IO, "Failed to save image: deflate failed");
    std::vector<uint8_t> o;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    o.insert(o.end(), sig, sig + 8);
    const uint8_t ihdr[13] = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                              (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                              8, 6, 0, 0, 0};  // 8 bits per channel, colour type 6 = RGBA, no interlace
    png_chunk(o, "IHDR", ihdr, sizeof ihdr);
    png_chunk(o, "IDAT", z.data(), (size_t)zlen);
    png_chunk(o, "IEND", nullptr, 0);
    FILE *f = fopen(path, "wb");
    if (!f) return fail(SBX_ERR_IO, "Failed to save image: cannot open %s", path);
    const bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
    if (fclose(f) != 0 || !ok) return fail(SBX_ERR_IO, "Failed to save image: write to %s failed", path);
    return SBX_OK;
}

extern "C" int32_t sbx_interference_eval(double n_film, const double wl[3], const double *d_nm, const double *cos_i,
                                         size_t n, int32_t gamma, double *rgb_out, int32_t device)
{
    if (!wl || !d_nm || !cos_i || !rgb_out) return fail(SBX_ERR_INVALID, "null argument");
    int32_t rc = pick_device(device);
    if (rc != SBX_OK) return rc;
    if (n == 0) return SBX_OK;
    double *dd = nullptr, *dc = nullptr, *dr = nullptr;
    CU(cudaMalloc(&dd, n * 8));
    cudaError_t e = cudaMalloc(&dc, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dr, n * 24);
    if (e == cudaSuccess) e = cudaMemcpy(dd, d_nm, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc, cos_i, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_interference_eval(n_film, wl, dd, dc, n, gamma, dr, 0);
    if (e == cudaSuccess) e = cudaMemcpy(rgb_out, dr, n * 24, cudaMemcpyDeviceToHost);
    cudaFree(dd); cudaFree(dc); cudaFree(dr);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "interference eval failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

extern "C" int32_t sbx_interference_terms(double n_film, const double *d_nm, const double *cos_i, size_t n, double wl,
                                          double *out_cos_t, double *out_fresnel, double *out_intensity, int32_t device)
{
    if (!d_nm || !cos_i) return fail(SBX_ERR_INVALID, "null argument");
    int32_t rc = pick_device(device);
    if (rc != SBX_OK) return rc;
    if (n == 0) return SBX_OK;
    double *dd = nullptr, *dc = nullptr, *o1 = nullptr, *o2 = nullptr, *o3 = nullptr;
    cudaError_t e = cudaMalloc(&dd, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dc, n * 8);
    if (e == cudaSuccess && out_cos_t) e = cudaMalloc(&o1, n * 8);
    if (e == cudaSuccess && out_fresnel) e = cudaMalloc(&o2, n * 24);
    if (e == cudaSuccess && out_intensity) e = cudaMalloc(&o3, n * 8);
    if (e == cudaSuccess) e = cudaMemcpy(dd, d_nm, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc, cos_i, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_interference_terms(n_film, dd, dc, n, wl, o1, o2, o3, 0);
    if (e == cudaSuccess && o1) e = cudaMemcpy(out_cos_t, o1, n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && o2) e = cudaMemcpy(out_fresnel, o2, n * 24, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && o3) e = cudaMemcpy(out_intensity, o3, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dd); cudaFree(dc); cudaFree(o1); cudaFree(o2); cudaFree(o3);
    if (e != cudaSuccess) return fail(SBX_ERR_CUDA, "interference terms failed: %s", cudaGetErrorString(e));
    return SBX_OK;
}

// ------------------------------------------------------------------ multi-GPU glue
static int32_t halo_exchange(sbx_drainage *h)
{
    h->ghost_stale = false;
    h->stale_by_pull = false;
    if (!h->comm) return SBX_OK;
    const char *err = nullptr;
    // my first owned row -> upper neighbour's bottom ghost; my last owned row -> lower neighbour's top ghost
    const size_t row_bytes = (size_t)h->num_phi * h->esz;
    for (int b = 0; b < h->batch; ++b) {
        char *base = (char *)h->buf[h->cur] + (size_t)b * h->bubble_stride * h->esz;
        int rc = sbx_comm_halo(h->comm, base + row_bytes, base, base + (size_t)h->rows_local * row_bytes,
                               base + (size_t)(h->rows_local + 1) * row_bytes, row_bytes, h->stream, &err);
        if (rc != 0) return fail(SBX_ERR_COMM, "halo exchange failed: %s", err ? err : "?");
        if (h->wgsl) {
            char *cb = (char *)h->conc[h->cur] + (size_t)b * h->bubble_stride * 4;
            const size_t rbc = (size_t)h->num_phi * 4;
            rc = sbx_comm_halo(h->comm, cb + rbc, cb, cb + (size_t)h->rows_local * rbc, cb + (size_t)(h->rows_local + 1) * rbc,
                               rbc, h->stream, &err);
            if (rc != 0) return fail(SBX_ERR_COMM, "halo exchange failed: %s", err ? err : "?");
        }
    }
    return SBX_OK;
}

extern "C" int32_t sbx_comm_unique_id(uint8_t out128[128])
{
    const char *err = nullptr;
    if (sbx_comm_unique_id_impl(out128, &err) != 0) return fail(SBX_ERR_COMM, "%s", err ? err : "nccl unique id failed");
    return SBX_OK;
}

extern "C" int32_t sbx_comm_init(const uint8_t id128[128], int32_t nranks, int32_t rank, int32_t device, sbx_comm **out)
{
    if (!id128 || !out) return fail(SBX_ERR_INVALID, "null argument");
    int32_t rc = pick_device(device);
    if (rc != SBX_OK) return rc;
    const char *err = nullptr;
    if (sbx_comm_init_impl(id128, nranks, rank, device, out, &err) != 0) return fail(SBX_ERR_COMM, "%s", err ? err : "nccl init failed");
    return SBX_OK;
}

extern "C" int32_t sbx_comm_destroy(sbx_comm *c)
{
    sbx_comm_destroy_impl(c);
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_attach_comm(sbx_drainage *h, sbx_comm *c)
{
    CHECK_H(h);
    h->comm = c;
    return halo_exchange(h);
}

struct IpcBlob {
    cudaIpcMemHandle_t buf[2];
    cudaIpcMemHandle_t conc[2];  // SBX_SEM_WGSL only
    cudaIpcMemHandle_t sync;
    int32_t rows_local, num_phi, esz, magic, wgsl, pad[3];
};
static_assert(sizeof(IpcBlob) <= SBX_IPC_BYTES, "IPC blob too large");

extern "C" int32_t sbx_drainage_ipc_export(sbx_drainage *h, uint8_t out[SBX_IPC_BYTES])
{
    CHECK_H(h);
    if (!out) return fail(SBX_ERR_INVALID, "out is null");
    IpcBlob b;
    memset(&b, 0, sizeof b);
    CU(cudaIpcGetMemHandle(&b.buf[0], h->buf[0]));
    CU(cudaIpcGetMemHandle(&b.buf[1], h->buf[1]));
    CU(cudaIpcGetMemHandle(&b.sync, h->sync_dev));
    if (h->wgsl) {
        CU(cudaIpcGetMemHandle(&b.conc[0], h->conc[0]));
        CU(cudaIpcGetMemHandle(&b.conc[1], h->conc[1]));
    }
    b.wgsl = h->wgsl ? 1 : 0;
    b.rows_local = h->rows_local;
    b.num_phi = h->num_phi;
    b.esz = (int32_t)h->esz;
    b.magic = 0x53425831;
    memset(out, 0, SBX_IPC_BYTES);
    memcpy(out, &b, sizeof b);
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_ipc_attach(sbx_drainage *h, const uint8_t *up, const uint8_t *down)
{
    CHECK_H(h);
    if (!h->comm) return fail(SBX_ERR_INVALID, "attach a communicator first (sbx_drainage_attach_comm)");
    if (h->strict || h->batch != 1)
        return fail(SBX_ERR_UNSUPPORTED, "the peer halo covers the TMA ring kernels (f32 / f64 CPU semantics, WGSL semantics), one bubble");
    if ((up == nullptr) != (h->row_begin == 0) || (down == nullptr) != (h->row_end == h->num_theta))
        return fail(SBX_ERR_INVALID, "neighbour blobs do not match this band's position");
    auto open_one = [&](const uint8_t *blob, void *bufs[2], void *concs[2], unsigned long long **sync, int *rows) -> int32_t {
        IpcBlob b;
        memcpy(&b, blob, sizeof b);
        if (b.magic != 0x53425831 || b.num_phi != h->num_phi || b.esz != (int32_t)h->esz || b.wgsl != (h->wgsl ? 1 : 0))
            return fail(SBX_ERR_INVALID, "neighbour blob is not a matching libsbx band");
        for (int i = 0; i < 2; ++i) CU(cudaIpcOpenMemHandle(&bufs[i], b.buf[i], cudaIpcMemLazyEnablePeerAccess));
        if (h->wgsl)
            for (int i = 0; i < 2; ++i) CU(cudaIpcOpenMemHandle(&concs[i], b.conc[i], cudaIpcMemLazyEnablePeerAccess));
        void *s = nullptr;
        CU(cudaIpcOpenMemHandle(&s, b.sync, cudaIpcMemLazyEnablePeerAccess));
        *sync = (unsigned long long *)s;
        if (rows) *rows = b.rows_local;
        return SBX_OK;
    };
    int32_t rc = SBX_OK;
    if (up) rc = open_one(up, h->peer_up_buf, h->peer_up_conc, &h->peer_up_sync, &h->peer_up_rows);
    if (rc == SBX_OK && down) rc = open_one(down, h->peer_dn_buf, h->peer_dn_conc, &h->peer_dn_sync, nullptr);
    if (rc != SBX_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));
    h->peer_enabled = true;
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_peer_status(sbx_drainage *h, int32_t *out)
{
    CHECK_H(h);
    if (!out) return fail(SBX_ERR_INVALID, "out is null");
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(h->scalar_host, h->sync_dev + 4, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(&v, h->scalar_host, sizeof v);
    *out = (int32_t)(v & 3ull);  // bit 0: neighbour wait timed out, bit 1: pole-ring wait timed out
    if (v) CU(cudaMemsetAsync(h->sync_dev + 4, 0, sizeof v, h->stream));
    return SBX_OK;
}

extern "C" int32_t sbx_drainage_minmax_global(sbx_drainage *h, double *min_out, double *max_out)
{
    double mn, mx;
    int32_t rc = sbx_drainage_minmax(h, &mn, &mx);
    if (rc != SBX_OK) return rc;
    if (h->comm) {
        const char *err = nullptr;
        double v[2] = {-mn, mx};  // one max-allreduce for both
        if (sbx_comm_allreduce_max2(h->comm, v, h->scalar_dev, h->stream, &err) != 0)
            return fail(SBX_ERR_COMM, "allreduce failed: %s", err ? err : "?");
        mn = -v[0];
        mx = v[1];
    }
    if (min_out) *min_out = mn;
    if (max_out) *max_out = mx;
    return SBX_OK;
}
