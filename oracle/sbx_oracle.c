/*
 * sbx_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see sbx_oracle.h for the rules and the parity-pinning statement).
 *
 * Every function cites the reference file:line it follows.  The restatement keeps the reference's
 * operation order (so that, compiled without FMA contraction against the same libm, results are
 * expected to be bit-identical to the Rust code), but not its structure: it is a flat C API over
 * plain arrays.
 */
#include "sbx_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

/* ================================================================== drainage (f64) */

struct orc_drainage {
    int num_theta, num_phi; /* rows held in h / scratch, columns */
    /* Window mode (orc_drainage_create_window): h holds global rows [row0, row0 + num_theta) of a sphere with
     * nt_global rows.  The whole-sphere simulator is the window row0 = 0, num_theta = nt_global. */
    int nt_global, row0;
    double *h, *scratch;
    double delta_theta, delta_phi;
    double density, gravity, viscosity, diffusion, critical, radius;
    double time;
    orc_tfe_front fronts[ORC_MAX_FRONTS + 1];
    int nfronts;
    double tfe_ratio;
    int tfe_enabled;
    double next_spawn, spawn_interval;
    int nthreads;
};

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

orc_drainage *orc_drainage_create(int num_theta, int num_phi, double h0, double density,
                                  double gravity, double viscosity, double diffusion,
                                  double critical, double radius)
{
    orc_drainage *s = (orc_drainage *)calloc(1, sizeof(*s));
    size_t n = (size_t)num_theta * (size_t)num_phi;
    s->num_theta = num_theta;
    s->nt_global = num_theta;
    s->row0 = 0;
    s->num_phi = num_phi;
    s->h = (double *)malloc(n * sizeof(double));
    s->scratch = (double *)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; ++i) s->h[i] = h0; /* ThicknessField::new, drainage.rs:121-122 */
    memcpy(s->scratch, s->h, n * sizeof(double));
    s->delta_theta = M_PI / (double)(num_theta - 1); /* drainage.rs:123 */
    s->delta_phi = 2.0 * M_PI / (double)num_phi;     /* drainage.rs:124 */
    s->density = density;
    s->gravity = gravity;
    s->viscosity = viscosity;
    s->diffusion = diffusion;
    s->critical = critical;
    s->radius = radius;
    s->time = 0.0;
    s->nfronts = 0;
    s->tfe_ratio = 0.85;     /* drainage.rs:283 */
    s->tfe_enabled = 1;      /* drainage.rs:284 */
    s->next_spawn = 0.5;     /* drainage.rs:285 */
    s->spawn_interval = 2.0; /* drainage.rs:286 */
    s->nthreads = 1;
    return s;
}

/* A latitude slab of a larger sphere: same simulator, but the field holds only global rows [row0, row0 + rows).
 * Rows whose theta neighbours lie outside the window keep their values, so after k steps the k rows next to a
 * cut edge are no longer those of the whole-sphere run; everything further inside is bit-identical to it (a cell
 * reads its four neighbours only, drainage.rs:336-343).  Pole rows are averaged only when the window holds them. */
orc_drainage *orc_drainage_create_window(int nt_global, int num_phi, int row0, int rows, double h0, double density,
                                         double gravity, double viscosity, double diffusion, double critical,
                                         double radius)
{
    if (row0 < 0 || rows < 2 || row0 + rows > nt_global) return NULL;
    orc_drainage *s = orc_drainage_create(rows, num_phi, h0, density, gravity, viscosity, diffusion, critical, radius);
    s->nt_global = nt_global;
    s->row0 = row0;
    s->delta_theta = M_PI / (double)(nt_global - 1); /* drainage.rs:123 on the global grid */
    return s;
}

void orc_drainage_destroy(orc_drainage *s)
{
    if (!s) return;
    free(s->h);
    free(s->scratch);
    free(s);
}

void orc_drainage_set_threads(orc_drainage *s, int nthreads) { s->nthreads = nthreads < 1 ? 1 : nthreads; }

/* TfeFront::contains, drainage.rs:55-71 */
static int front_contains(const orc_tfe_front *f, double theta, double phi)
{
    double dist = fabs(theta - M_PI_2);
    if (dist > f->extent) return 0;
    double phi_diff = fabs(phi - f->phi_center);
    if (phi_diff > M_PI) phi_diff = 2.0 * M_PI - phi_diff;
    return phi_diff < f->width / 2.0;
}

/* spawn_tfe_front, drainage.rs:464-477 (uses current_time before this step's increment) */
static void spawn_front(orc_drainage *s)
{
    double t = s->time;
    orc_tfe_front f;
    f.phi_center = fabs(sin(t * 7.89 + 1.23)) * 2.0 * M_PI;
    f.width = 0.3 + fabs(sin(t * 3.21)) * 0.4;
    f.velocity = 0.05 + fabs(cos(t * 5.67)) * 0.05;
    f.extent = 0.0;
    s->fronts[s->nfronts++] = f;
    if (s->nfronts > ORC_MAX_FRONTS) { /* remove(0) */
        memmove(&s->fronts[0], &s->fronts[1], (size_t)(s->nfronts - 1) * sizeof(orc_tfe_front));
        s->nfronts--;
    }
}

/* update_marginal_regeneration, drainage.rs:395-461 */
static void marginal_regeneration(orc_drainage *s, double dt)
{
    s->next_spawn -= dt;
    if (s->next_spawn <= 0.0) {
        spawn_front(s);
        double random_factor = 0.5 + fabs(sin(s->time * 12.345));
        s->next_spawn = s->spawn_interval * random_factor;
    }
    for (int k = 0; k < s->nfronts; ++k) { /* TfeFront::advance, drainage.rs:47-49 */
        double e = s->fronts[k].extent + s->fronts[k].velocity * dt;
        s->fronts[k].extent = e < M_PI_2 ? e : M_PI_2;
    }
    int w = 0; /* retain, drainage.rs:411 */
    for (int k = 0; k < s->nfronts; ++k)
        if (s->fronts[k].extent < M_PI_2 * 0.95) s->fronts[w++] = s->fronts[k];
    s->nfronts = w;
    if (s->nfronts == 0) return;

    const int num_theta = s->nt_global, num_phi = s->num_phi;
    const double dth = s->delta_theta, dph = s->delta_phi;
    size_t min_idx = (size_t)num_theta, max_idx = 0;
    for (int k = 0; k < s->nfronts; ++k) { /* drainage.rs:428-434 */
        double lo = M_PI_2 - s->fronts[k].extent, hi = M_PI_2 + s->fronts[k].extent;
        double flo = floor(lo / dth), fhi = ceil(hi / dth);
        size_t a = flo <= 0.0 ? 0 : (size_t)flo; /* `as usize` saturates */
        size_t b = fhi <= 0.0 ? 0 : (size_t)fhi;
        if (b > (size_t)(num_theta - 1)) b = (size_t)(num_theta - 1);
        if (a < min_idx) min_idx = a;
        if (b > max_idx) max_idx = b;
    }
    if (min_idx > (size_t)(num_theta - 1)) min_idx = (size_t)(num_theta - 1);
    for (size_t i = min_idx; i <= max_idx; ++i) { /* drainage.rs:440-460 */
        double theta = (double)i * dth;
        if ((long)i < (long)s->row0 || (long)i >= (long)s->row0 + s->num_theta) continue; /* window mode */
        const size_t li = i - (size_t)s->row0;
        for (int j = 0; j < num_phi; ++j) {
            double phi = (double)j * dph;
            for (int k = 0; k < s->nfronts; ++k) {
                if (front_contains(&s->fronts[k], theta, phi)) {
                    double current = s->h[li * (size_t)num_phi + (size_t)j];
                    double reduction_rate = (1.0 - s->tfe_ratio) * 0.5;
                    s->h[li * (size_t)num_phi + (size_t)j] = current * (1.0 - reduction_rate * dt);
                    break;
                }
            }
        }
    }
}

/* update_poles, drainage.rs:480-503 */
static void update_poles(orc_drainage *s)
{
    const int num_phi = s->num_phi;
    const size_t last = (size_t)s->num_theta - 1;
    if (s->row0 == 0) {
        double top = 0.0;
        for (int j = 0; j < num_phi; ++j) top += s->h[(size_t)num_phi + (size_t)j];
        double top_avg = top / (double)num_phi;
        for (int j = 0; j < num_phi; ++j) s->h[j] = top_avg;
    }
    if (s->row0 + s->num_theta == s->nt_global) {
        double bot = 0.0;
        for (int j = 0; j < num_phi; ++j) bot += s->h[(last - 1) * (size_t)num_phi + (size_t)j];
        double bot_avg = bot / (double)num_phi;
        for (int j = 0; j < num_phi; ++j) s->h[last * (size_t)num_phi + (size_t)j] = bot_avg;
    }
}

/* one interior row, drainage.rs:314-381 */
static void interior_row(orc_drainage *s, int i, double dt, double drainage_coefficient)
{
    const int num_phi = s->num_phi;
    const double delta_theta = s->delta_theta, delta_phi = s->delta_phi;
    const double *sc = s->scratch;
    double *out = s->h;
    double theta = (double)(i + s->row0) * delta_theta; /* i is the row inside the window */
    double sin_theta = sin(theta);
    double cos_theta = cos(theta);
    double sin_theta_safe = fabs(sin_theta) < 1e-10 ? copysign(1e-10, sin_theta) : sin_theta;
    const size_t row = (size_t)i * (size_t)num_phi;
    for (int j = 0; j < num_phi; ++j) {
        double hc = sc[row + (size_t)j];
        if (hc < s->critical) continue; /* drainage.rs:331-333 */
        double h_tm = sc[row - (size_t)num_phi + (size_t)j];
        double h_tp = sc[row + (size_t)num_phi + (size_t)j];
        int jm = (j + num_phi - 1) % num_phi;
        int jp = (j + 1) % num_phi;
        double h_pm = sc[row + (size_t)jm];
        double h_pp = sc[row + (size_t)jp];
        double cubed = hc * hc * hc;
        double drainage_term = -drainage_coefficient * cubed * sin_theta;
        double radius_squared = s->radius * s->radius;
        double d2h_dtheta2 = (h_tp - 2.0 * hc + h_tm) / (delta_theta * delta_theta);
        double dh_dtheta = (h_tp - h_tm) / (2.0 * delta_theta);
        double d2h_dphi2 = (h_pp - 2.0 * hc + h_pm) / (delta_phi * delta_phi);
        double laplacian = (d2h_dtheta2 + cos_theta / sin_theta_safe * dh_dtheta +
                            d2h_dphi2 / (sin_theta_safe * sin_theta_safe)) /
                           radius_squared;
        double diffusion_term = s->diffusion * laplacian;
        double dh_dt = drainage_term + diffusion_term;
        double nh = hc + dt * dh_dt;
        out[row + (size_t)j] = nh > 0.0 ? nh : 0.0; /* .max(0.0); NaN -> 0.0 like f64::max */
    }
}

void orc_drainage_step(orc_drainage *s, double dt)
{
    const int num_theta = s->num_theta;
    const size_t n = (size_t)num_theta * (size_t)s->num_phi;
    double drainage_coefficient = s->density * s->gravity / (3.0 * s->viscosity); /* :305 */
    memcpy(s->scratch, s->h, n * sizeof(double));                                  /* :311 */
    if (s->nthreads > 1) {
#pragma omp parallel for schedule(static) num_threads(s->nthreads)
        for (int i = 1; i < num_theta - 1; ++i) interior_row(s, i, dt, drainage_coefficient);
    } else {
        for (int i = 1; i < num_theta - 1; ++i) interior_row(s, i, dt, drainage_coefficient);
    }
    update_poles(s);                                  /* :384 */
    if (s->tfe_enabled) marginal_regeneration(s, dt); /* :387-389 */
    s->time += dt;                                    /* :391 */
}

double *orc_drainage_data(orc_drainage *s) { return s->h; }
int orc_drainage_num_theta(const orc_drainage *s) { return s->num_theta; }
int orc_drainage_num_phi(const orc_drainage *s) { return s->num_phi; }
double orc_drainage_delta_theta(const orc_drainage *s) { return s->delta_theta; }
double orc_drainage_delta_phi(const orc_drainage *s) { return s->delta_phi; }

/* get_thickness, drainage.rs:516-552 */
double orc_drainage_get_thickness(const orc_drainage *s, double theta, double phi)
{
    const int num_theta = s->num_theta, num_phi = s->num_phi;
    double theta_clamped = clampd(theta, 0.0, M_PI);
    double two_pi = 2.0 * M_PI;
    double phi_normalized = fmod(phi, two_pi); /* rem_euclid */
    if (phi_normalized < 0.0) phi_normalized += two_pi;
    double theta_continuous = theta_clamped / s->delta_theta;
    double phi_continuous = phi_normalized / s->delta_phi;
    size_t ti = (size_t)floor(theta_continuous);
    if (ti > (size_t)(num_theta - 2)) ti = (size_t)(num_theta - 2);
    size_t pj = (size_t)floor(phi_continuous) % (size_t)num_phi;
    size_t ti_hi = ti + 1;
    size_t pj_hi = (pj + 1) % (size_t)num_phi;
    double tf = theta_continuous - (double)ti;
    double pf = phi_continuous - (double)pj;
    const double *h = s->h;
    double v00 = h[ti * (size_t)num_phi + pj];
    double v01 = h[ti * (size_t)num_phi + pj_hi];
    double v10 = h[ti_hi * (size_t)num_phi + pj];
    double v11 = h[ti_hi * (size_t)num_phi + pj_hi];
    double v0 = v00 * (1.0 - pf) + v01 * pf;
    double v1 = v10 * (1.0 - pf) + v11 * pf;
    return v0 * (1.0 - tf) + v1 * tf;
}

double orc_drainage_min(const orc_drainage *s)
{
    double m = INFINITY;
    size_t n = (size_t)s->num_theta * (size_t)s->num_phi;
    for (size_t i = 0; i < n; ++i) m = fmin(m, s->h[i]);
    return m;
}

double orc_drainage_max(const orc_drainage *s)
{
    double m = -INFINITY;
    size_t n = (size_t)s->num_theta * (size_t)s->num_phi;
    for (size_t i = 0; i < n; ++i) m = fmax(m, s->h[i]);
    return m;
}

int orc_drainage_has_critical(const orc_drainage *s) { return orc_drainage_min(s) < s->critical; }
double orc_drainage_critical(const orc_drainage *s) { return s->critical; }
double orc_drainage_time(const orc_drainage *s) { return s->time; }

void orc_drainage_reset(orc_drainage *s, double h0)
{
    size_t n = (size_t)s->num_theta * (size_t)s->num_phi;
    for (size_t i = 0; i < n; ++i) s->h[i] = h0;
    s->time = 0.0;
    s->nfronts = 0;
    s->next_spawn = 0.5;
}

void orc_drainage_set_marginal_regeneration(orc_drainage *s, int enabled)
{
    s->tfe_enabled = enabled ? 1 : 0;
    if (!enabled) s->nfronts = 0;
}
int orc_drainage_marginal_regeneration(const orc_drainage *s) { return s->tfe_enabled; }
void orc_drainage_set_tfe_ratio(orc_drainage *s, double ratio) { s->tfe_ratio = clampd(ratio, 0.5, 0.95); }
double orc_drainage_tfe_ratio(const orc_drainage *s) { return s->tfe_ratio; }
int orc_drainage_active_tfe_count(const orc_drainage *s) { return s->nfronts; }
int orc_drainage_fronts(const orc_drainage *s, orc_tfe_front *out)
{
    memcpy(out, s->fronts, (size_t)s->nfronts * sizeof(orc_tfe_front));
    return s->nfronts;
}
double orc_drainage_next_spawn(const orc_drainage *s) { return s->next_spawn; }

/* restore current_time, next_tfe_spawn_time and tfe_fronts (drainage.rs:237-255): checkpoint / resume of a run */
int orc_drainage_set_state(orc_drainage *s, double time, double next_spawn, const orc_tfe_front *fronts, int count)
{
    if (count < 0 || count > ORC_MAX_FRONTS) return -1;
    s->time = time;
    s->next_spawn = next_spawn;
    s->nfronts = count;
    if (count) memcpy(s->fronts, fronts, (size_t)count * sizeof(orc_tfe_front));
    return 0;
}

/* ================================================================== WGSL drainage (f32) */

void orc_wgsl_params_default(orc_wgsl_params *p)
{
    memset(p, 0, sizeof(*p));
    p->dt = 0.001f;
    p->gravity = 9.81f;
    p->viscosity = 0.001f;
    p->density = 1000.0f;
    p->diffusion_coeff = 1e-9f;
    p->bubble_radius = 0.025f;
    p->critical_thickness = 30e-9f;
    p->grid_width = 128;
    p->grid_height = 64;
    p->marangoni_enabled = 0;
    p->gamma_air = 0.072f;
    p->gamma_reduction = 0.045f;
    p->surfactant_diffusion = 1e-9f;
    p->marangoni_coeff = 0.01f;
}

static const float PI_F = 3.14159265358979323846f; /* drainage.wgsl:39 */

static float wgsl_get(const float *a, const orc_wgsl_params *p, int ti, int pj)
{ /* drainage.wgsl:47-64 */
    int H = (int)p->grid_height, W = (int)p->grid_width;
    int t = ti < 0 ? 0 : (ti > H - 1 ? H - 1 : ti);
    int q = pj % W;
    if (q < 0) q += W;
    return a[(size_t)t * (size_t)W + (size_t)q];
}

void orc_wgsl_step(const orc_wgsl_params *p, const float *h_in, const float *c_in, float *h_out,
                   float *c_out)
{
    const int W = (int)p->grid_width, H = (int)p->grid_height;
    for (int ti = 0; ti < H; ++ti) {
        for (int pj = 0; pj < W; ++pj) {
            size_t idx = (size_t)ti * (size_t)W + (size_t)pj;
            if (ti == 0 || ti == H - 1) { /* drainage.wgsl:85-89 */
                h_out[idx] = h_in[idx];
                c_out[idx] = c_in[idx];
                continue;
            }
            float delta_theta = PI_F / (float)(p->grid_height - 1u);
            float delta_phi = 2.0f * PI_F / (float)p->grid_width;
            float theta = (float)ti * delta_theta;
            float sin_theta = sinf(theta);
            float cos_theta = cosf(theta);
            float sgn = sin_theta > 0.f ? 1.f : (sin_theta < 0.f ? -1.f : 0.f);
            float sin_theta_safe = fabsf(sin_theta) < 0.0001f ? sgn * 0.0001f : sin_theta; /* :101 */
            float h = h_in[idx];
            float conc = c_in[idx];
            if (h < p->critical_thickness) { /* :108-112 */
                h_out[idx] = h;
                c_out[idx] = conc;
                continue;
            }
            float h_tm = wgsl_get(h_in, p, ti - 1, pj), h_tp = wgsl_get(h_in, p, ti + 1, pj);
            float h_pm = wgsl_get(h_in, p, ti, pj - 1), h_pp = wgsl_get(h_in, p, ti, pj + 1);
            float c_tm = wgsl_get(c_in, p, ti - 1, pj), c_tp = wgsl_get(c_in, p, ti + 1, pj);
            float c_pm = wgsl_get(c_in, p, ti, pj - 1), c_pp = wgsl_get(c_in, p, ti, pj + 1);

            float drainage_coeff = p->density * p->gravity / (3.0f * p->viscosity);
            float h_cubed = h * h * h;
            float drainage_term = -drainage_coeff * h_cubed * sin_theta;
            float radius_squared = p->bubble_radius * p->bubble_radius;
            float d2h_dtheta2 = (h_tp - 2.0f * h + h_tm) / (delta_theta * delta_theta);
            float dh_dtheta = (h_tp - h_tm) / (2.0f * delta_theta);
            float d2h_dphi2 = (h_pp - 2.0f * h + h_pm) / (delta_phi * delta_phi);
            float h_laplacian = (d2h_dtheta2 + cos_theta / sin_theta_safe * dh_dtheta +
                                 d2h_dphi2 / (sin_theta_safe * sin_theta_safe)) /
                                radius_squared;
            float diffusion_term = p->diffusion_coeff * h_laplacian;
            float marangoni_term = 0.0f;
            if (p->marangoni_enabled != 0u) { /* :151-171 */
                float dc_dtheta = (c_tp - c_tm) / (2.0f * delta_theta);
                float dc_dphi = (c_pp - c_pm) / (2.0f * delta_phi);
                float grad_c_theta = dc_dtheta / p->bubble_radius;
                float grad_c_phi = dc_dphi / (p->bubble_radius * sin_theta_safe);
                float grad_c_mag = sqrtf(grad_c_theta * grad_c_theta + grad_c_phi * grad_c_phi);
                float marangoni_stress = p->gamma_reduction * grad_c_mag;
                marangoni_term = p->marangoni_coeff * marangoni_stress * h * h;
            }
            float dh_dt = drainage_term + diffusion_term + marangoni_term;
            float new_h = fmaxf(h + p->dt * dh_dt, 0.0f);
            h_out[idx] = new_h;

            float d2c_dtheta2 = (c_tp - 2.0f * conc + c_tm) / (delta_theta * delta_theta);
            float dc_dtheta2 = (c_tp - c_tm) / (2.0f * delta_theta);
            float d2c_dphi2 = (c_pp - 2.0f * conc + c_pm) / (delta_phi * delta_phi);
            float c_laplacian = (d2c_dtheta2 + cos_theta / sin_theta_safe * dc_dtheta2 +
                                 d2c_dphi2 / (sin_theta_safe * sin_theta_safe)) /
                                radius_squared;
            float dc_dt = p->surfactant_diffusion * c_laplacian;
            if (p->marangoni_enabled != 0u) { /* :193-198 */
                float coupling = 0.001f * (h_laplacian * p->bubble_radius * p->bubble_radius);
                dc_dt += coupling * conc;
            }
            float new_c = conc + p->dt * dc_dt;
            new_c = new_c < 0.f ? 0.f : (new_c > 1.f ? 1.f : new_c);
            c_out[idx] = new_c;
        }
    }
}

/* ================================================================== interference (f64) */

void orc_interference_new(orc_interference *c, double n_film)
{
    c->n_film = n_film;
    c->n_medium = 1.0;            /* interference.rs:36,144 */
    c->wavelengths_nm[0] = 650.0; /* interference.rs:26 */
    c->wavelengths_nm[1] = 532.0; /* interference.rs:28 */
    c->wavelengths_nm[2] = 450.0; /* interference.rs:30 */
}

double orc_transmission_angle_cos(const orc_interference *c, double cos_i)
{
    double ci = clampd(cos_i, -1.0, 1.0);
    double sin_i = sqrt(1.0 - ci * ci);
    double sin_t = (c->n_medium / c->n_film) * sin_i;
    if (fabs(sin_t) > 1.0) return 0.0;
    return sqrt(1.0 - sin_t * sin_t);
}

double orc_optical_path_difference(const orc_interference *c, double d_nm, double cos_t, double wl)
{
    double geometric_path = 2.0 * c->n_film * d_nm * cos_t;
    return geometric_path + wl / 2.0;
}

double orc_phase_difference(double opd_nm, double wl) { return 2.0 * M_PI * opd_nm / wl; }

void orc_fresnel_reflection(const orc_interference *c, double cos_i, double out[3])
{
    double ci = clampd(cos_i, -1.0, 1.0);
    double ct = orc_transmission_angle_cos(c, ci);
    double n1 = c->n_medium, n2 = c->n_film;
    double n1_cos_i = n1 * ci, n2_cos_t = n2 * ct, n2_cos_i = n2 * ci, n1_cos_t = n1 * ct;
    double rs = (n1_cos_i - n2_cos_t) / (n1_cos_i + n2_cos_t);
    double Rs = rs * rs;
    double rp = (n2_cos_i - n1_cos_t) / (n2_cos_i + n1_cos_t);
    double Rp = rp * rp;
    out[0] = Rs;
    out[1] = Rp;
    out[2] = (Rs + Rp) / 2.0;
}

double orc_reflected_intensity(const orc_interference *c, double d_nm, double cos_i, double wl)
{
    if (d_nm <= 0.0) return 0.0;
    double ct = orc_transmission_angle_cos(c, cos_i);
    double opd = orc_optical_path_difference(c, d_nm, ct, wl);
    double phase = orc_phase_difference(opd, wl);
    double fr[3];
    orc_fresnel_reflection(c, cos_i, fr);
    double R = fr[2];
    double cos_phase = cos(phase);
    double numerator = 2.0 * R * (1.0 - cos_phase);
    double denominator = 1.0 + R * R - 2.0 * R * cos_phase;
    if (fabs(denominator) < DBL_EPSILON) return R;
    return clampd(numerator / denominator, 0.0, 1.0);
}

void orc_interference_color_linear(const orc_interference *c, double d_nm, double cos_i, double rgb[3])
{
    for (int k = 0; k < 3; ++k)
        rgb[k] = clampd(orc_reflected_intensity(c, d_nm, cos_i, c->wavelengths_nm[k]), 0.0, 1.0);
}

void orc_interference_color(const orc_interference *c, double d_nm, double cos_i, double rgb[3])
{
    double gamma = 2.2;
    for (int k = 0; k < 3; ++k) {
        double I = orc_reflected_intensity(c, d_nm, cos_i, c->wavelengths_nm[k]);
        rgb[k] = clampd(pow(I, 1.0 / gamma), 0.0, 1.0);
    }
}

void orc_rgb_to_u8(const double rgb[3], uint8_t out[3])
{
    for (int k = 0; k < 3; ++k) {
        double v = round(rgb[k] * 255.0); /* f64::round: half away from zero */
        out[k] = (uint8_t)(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v));
    }
}

void orc_color_lut_1d(const orc_interference *c, double min_nm, double max_nm, int n, double cos_theta,
                      double *out)
{
    double step = (max_nm - min_nm) / (double)(n - 1);
    for (int i = 0; i < n; ++i) orc_interference_color(c, min_nm + (double)i * step, cos_theta, out + 3 * i);
}

void orc_color_lut_2d(const orc_interference *c, double min_nm, double max_nm, int T, int A, double *out)
{
    int tden = T - 1 > 1 ? T - 1 : 1, aden = A - 1 > 1 ? A - 1 : 1;
    double tstep = (max_nm - min_nm) / (double)tden;
    double astep = 1.0 / (double)aden;
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < A; ++j)
            orc_interference_color(c, min_nm + (double)i * tstep, (double)j * astep,
                                   out + 3 * ((size_t)i * (size_t)A + (size_t)j));
}

void orc_interference_batch(const orc_interference *c, const double *d_nm, const double *cos_i, size_t n,
                            int gamma, double *rgb)
{
    for (size_t i = 0; i < n; ++i) {
        if (gamma) orc_interference_color(c, d_nm[i], cos_i[i], rgb + 3 * i);
        else orc_interference_color_linear(c, d_nm[i], cos_i[i], rgb + 3 * i);
    }
}

/* wall.wgsl:165-207 */
void orc_schlick_wall_rgb(float d_nm, float cos_i, float n_film, float rgb[3])
{
    static const float WL[3] = {650.0f, 532.0f, 450.0f}; /* wall.wgsl:157-159 */
    for (int k = 0; k < 3; ++k) {
        float n_ratio = 1.0f / n_film;
        float sin_i_sq = 1.0f - cos_i * cos_i;
        float sin_t_sq = sin_i_sq * n_ratio * n_ratio;
        float cos_t = sin_t_sq > 1.0f ? 0.0f : sqrtf(1.0f - sin_t_sq);
        float optical_path = 2.0f * n_film * d_nm * cos_t;
        float delta = 2.0f * PI_F * optical_path / WL[k] + PI_F;
        float q = (1.0f - n_film) / (1.0f + n_film);
        float r0 = q * q;
        float R = r0 + (1.0f - r0) * powf(1.0f - cos_i, 5.0f);
        float F = 4.0f * R / ((1.0f - R) * (1.0f - R));
        float sh = sinf(delta / 2.0f);
        rgb[k] = 1.0f / (1.0f + F * sh * sh);
    }
}

/* docs/key-insights.md:288-309 (== soap_bubble_simulation_roadmap.md:371-399) */
void orc_schlick_two_beam_rgb(float d_nm, float cos_i, float n_film, float rgb[3])
{
    static const float WL[3] = {650.0f, 532.0f, 450.0f};
    float sin_i = sqrtf(fmaxf(0.0f, 1.0f - cos_i * cos_i)); /* snells_law, bubble.wgsl:336-341 */
    float sin_t = sin_i * 1.0f / n_film;
    float cos_t = sqrtf(fmaxf(0.0f, 1.0f - sin_t * sin_t));
    float optical_path = 2.0f * n_film * d_nm * cos_t;
    float q = (n_film - 1.0f) / (n_film + 1.0f);
    float r0 = q * q;
    float fresnel = r0 + (1.0f - r0) * powf(1.0f - cos_i, 5.0f);
    for (int k = 0; k < 3; ++k) {
        float phase = 2.0f * PI_F * optical_path / WL[k] + PI_F;
        float intensity = 0.5f * (1.0f + cosf(phase));
        rgb[k] = intensity * fresnel * 4.0f;
    }
}

/* ================================================================== interference LUT (f32) */

float orc_snells_law(float cos_i, float n_film)
{
    float sin_i = sqrtf(fmaxf(1.0f - cos_i * cos_i, 0.0f));
    float sin_t = sin_i / n_film;
    return sqrtf(fmaxf(1.0f - sin_t * sin_t, 0.0f));
}

float orc_fresnel_unpolarized(float cos_i, float cos_t, float n1, float n2)
{
    float n1_cos_i = n1 * cos_i, n2_cos_t = n2 * cos_t;
    float r_s = (n1_cos_i - n2_cos_t) / fmaxf(n1_cos_i + n2_cos_t, 0.0001f);
    float n2_cos_i = n2 * cos_i, n1_cos_t = n1 * cos_t;
    float r_p = (n2_cos_i - n1_cos_t) / fmaxf(n2_cos_i + n1_cos_t, 0.0001f);
    return (r_s * r_s + r_p * r_p) * 0.5f;
}

static float airy_interference(float phase, float reflectance)
{ /* interference_lut.rs:123-133 */
    float one_minus_r = fmaxf(1.0f - reflectance, 0.001f);
    float f = 4.0f * reflectance / (one_minus_r * one_minus_r);
    float sh = sinf(phase * 0.5f);
    float sin2 = sh * sh;
    float numerator = f * sin2;
    float denominator = 1.0f + f * sin2;
    return numerator / fmaxf(denominator, 0.001f);
}

static float gaussianf(float x, float mean, float sigma)
{ /* interference_lut.rs:152-155 */
    float t = (x - mean) / sigma;
    return expf(-0.5f * t * t);
}

static void cie_color_matching(float wl, float out[3])
{ /* interference_lut.rs:136-148 */
    float x = 1.056f * gaussianf(wl, 599.8f, 37.9f) + 0.362f * gaussianf(wl, 442.0f, 16.0f) -
              0.065f * gaussianf(wl, 501.1f, 20.4f);
    float y = 0.821f * gaussianf(wl, 568.8f, 46.9f) + 0.286f * gaussianf(wl, 530.9f, 31.1f);
    float z = 1.217f * gaussianf(wl, 437.0f, 11.8f) + 0.681f * gaussianf(wl, 459.0f, 26.0f);
    out[0] = fmaxf(x, 0.0f);
    out[1] = fmaxf(y, 0.0f);
    out[2] = fmaxf(z, 0.0f);
}

static void thin_film_spectral(float d_nm, float cos_theta, float n_film, float intensity,
                               const float *wl, int n_lambda, float rgb[3])
{ /* interference_lut.rs:45-93 with the wavelength list as a parameter */
    float cos_t = orc_snells_law(cos_theta, n_film);
    float fresnel = orc_fresnel_unpolarized(cos_theta, cos_t, 1.0f, n_film);
    float optical_path = 2.0f * n_film * d_nm * cos_t;
    float xyz[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < n_lambda; ++k) {
        float phase = 2.0f * PI_F * optical_path / wl[k] + PI_F;
        float a = airy_interference(phase, fresnel);
        float cie[3];
        cie_color_matching(wl[k], cie);
        xyz[0] += cie[0] * a;
        xyz[1] += cie[1] * a;
        xyz[2] += cie[2] * a;
    }
    float nl = (float)n_lambda;
    xyz[0] /= nl;
    xyz[1] /= nl;
    xyz[2] /= nl;
    /* xyz_to_rgb, interference_lut.rs:158-163 */
    float r = 3.2404542f * xyz[0] - 1.5371385f * xyz[1] - 0.4985314f * xyz[2];
    float g = -0.969266f * xyz[0] + 1.8760108f * xyz[1] + 0.0415560f * xyz[2];
    float b = 0.0556434f * xyz[0] - 0.2040259f * xyz[1] + 1.0572252f * xyz[2];
    rgb[0] = r * intensity;
    rgb[1] = g * intensity;
    rgb[2] = b * intensity;
}

static const float LUT_WL7[7] = {400.f, 450.f, 500.f, 550.f, 600.f, 650.f, 700.f}; /* :52 */

/* thin_film_interference of the fragment shaders (bubble_instanced.wgsl:190-218, bubble.wgsl:442-483): the same f32
 * algorithm per pixel with interference_intensity applied and the result clamped; wavelength list as a parameter */
void orc_thin_film_spectral(float d_nm, float cos_theta, float n_film, float intensity, const float *wl, int n_lambda,
                            float rgb[3])
{
    thin_film_spectral(d_nm, cos_theta, n_film, intensity, wl, n_lambda, rgb);
    for (int k = 0; k < 3; ++k) rgb[k] = rgb[k] < 0.f ? 0.f : (rgb[k] > 1.f ? 1.f : rgb[k]);
}

void orc_thin_film_interference(float d_nm, float cos_theta, float n_film, float intensity, float rgb[3])
{
    thin_film_spectral(d_nm, cos_theta, n_film, intensity, LUT_WL7, 7, rgb);
}

static uint8_t unorm_trunc(float x)
{ /* (x.clamp(0,1) * 255.0) as u8 -- truncation, interference_lut.rs:33-35 */
    float c = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
    if (c != c) return 0; /* NaN as u8 == 0 */
    return (uint8_t)(c * 255.0f);
}

void orc_generate_lut_spectral(float n_film, float intensity, const float *wl, int n_lambda, int T, int A,
                               float max_nm, uint8_t *out)
{
    size_t o = 0;
    for (int a = 0; a < A; ++a) {
        for (int t = 0; t < T; ++t) {
            float d_nm = ((float)t / (float)(T - 1)) * max_nm;
            float cos_theta = (float)a / (float)(A - 1);
            float rgb[3];
            thin_film_spectral(d_nm, cos_theta, n_film, intensity, wl, n_lambda, rgb);
            out[o++] = unorm_trunc(rgb[0]);
            out[o++] = unorm_trunc(rgb[1]);
            out[o++] = unorm_trunc(rgb[2]);
            out[o++] = 255;
        }
    }
}

void orc_generate_interference_lut(float n_film, float intensity, uint8_t *out)
{
    orc_generate_lut_spectral(n_film, intensity, LUT_WL7, 7, ORC_LUT_THICKNESS_SAMPLES,
                              ORC_LUT_ANGLE_SAMPLES, ORC_LUT_MAX_THICKNESS_NM, out);
}

/* bubble.wgsl:425-439 + sampler at pipeline.rs:350-359 (linear, clamp-to-edge, Rgba8Unorm).
 * Texel-centre addressing: texel i covers [i, i+1)/N, its centre is (i+0.5)/N.
 * weight_bits = 0: ideal bilinear weights in f32.  weight_bits = 8: the weights are rounded to 8 fractional
 * bits first, which is what GPU texture units do (CUDA Programming Guide, "Texture Fetching": 9-bit fixed
 * point with 8 bits of fractional value; Vulkan only guarantees subTexelPrecisionBits) -- i.e. what the
 * reference's textureSample actually evaluates on hardware. */
void orc_lut_shade_w(const uint8_t *lut, int T, int A, float max_nm, float d_nm, float cos_theta,
                     float intensity, int weight_bits, float rgb[3])
{
    float u = d_nm / max_nm;
    u = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);
    float v = cos_theta < 0.f ? 0.f : (cos_theta > 1.f ? 1.f : cos_theta);
    float x = u * (float)T - 0.5f, y = v * (float)A - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float wx = x - fx0, wy = y - fy0;
    if (weight_bits > 0) {
        const float q = (float)(1 << weight_bits);
        wx = floorf(wx * q + 0.5f) / q;
        wy = floorf(wy * q + 0.5f) / q;
    }
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > T - 1) x1 = T - 1;
    if (y1 > A - 1) y1 = A - 1;
    for (int k = 0; k < 3; ++k) {
        float t00 = (float)lut[4 * ((size_t)y0 * (size_t)T + (size_t)x0) + (size_t)k] / 255.0f;
        float t10 = (float)lut[4 * ((size_t)y0 * (size_t)T + (size_t)x1) + (size_t)k] / 255.0f;
        float t01 = (float)lut[4 * ((size_t)y1 * (size_t)T + (size_t)x0) + (size_t)k] / 255.0f;
        float t11 = (float)lut[4 * ((size_t)y1 * (size_t)T + (size_t)x1) + (size_t)k] / 255.0f;
        float c;
        if (weight_bits > 0) {
            c = (1.f - wx) * (1.f - wy) * t00 + wx * (1.f - wy) * t10 + (1.f - wx) * wy * t01 + wx * wy * t11;
        } else {
            float top = t00 + (t10 - t00) * wx;
            float bot = t01 + (t11 - t01) * wx;
            c = top + (bot - top) * wy;
        }
        c *= intensity;
        rgb[k] = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
    }
}

void orc_lut_shade(const uint8_t *lut, int T, int A, float max_nm, float d_nm, float cos_theta,
                   float intensity, float rgb[3])
{
    orc_lut_shade_w(lut, T, A, max_nm, d_nm, cos_theta, intensity, 0, rgb);
}

double orc_view_cos(double theta, double phi, double radius, const double cam[3])
{
    double n[3] = {sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi)};
    double v[3] = {cam[0] - radius * n[0], cam[1] - radius * n[1], cam[2] - radius * n[2]};
    double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double d = (v[0] * n[0] + v[1] * n[1] + v[2] * n[2]) / len;
    return fabs(d);
}

/* per-cell spectral colour of a field band (the view cosine as in orc_shade_field): RGBA8 by round(255 x) or 3 floats */
void orc_shade_field_spectral(const double *h, int num_theta, int num_phi, int row0, int num_rows, double radius,
                              const double cam[3], float n_film, float intensity, const float *wl, int n_lambda,
                              uint8_t *out_rgba, float *out_rgb, int nthreads)
{
    const double dth = M_PI / (double)(num_theta - 1), dph = 2.0 * M_PI / (double)num_phi;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int r = 0; r < num_rows; ++r) {
        const double theta = (double)(row0 + r) * dth;
        for (int j = 0; j < num_phi; ++j) {
            const size_t idx = (size_t)r * (size_t)num_phi + (size_t)j;
            const double c = orc_view_cos(theta, (double)j * dph, radius, cam);
            float f[3];
            orc_thin_film_spectral((float)(h[idx] * 1e9), (float)c, n_film, intensity, wl, n_lambda, f);
            if (out_rgb) { out_rgb[3 * idx] = f[0]; out_rgb[3 * idx + 1] = f[1]; out_rgb[3 * idx + 2] = f[2]; }
            if (out_rgba) {
                double d[3] = {f[0], f[1], f[2]};
                uint8_t px[3];
                orc_rgb_to_u8(d, px);
                out_rgba[4 * idx] = px[0]; out_rgba[4 * idx + 1] = px[1]; out_rgba[4 * idx + 2] = px[2]; out_rgba[4 * idx + 3] = 255;
            }
        }
    }
}

void orc_shade_field(const double *h, int num_theta, int num_phi, int row0, int num_rows, double radius,
                     const double cam[3], double n_film, const uint8_t *lut, int T, int A, float max_nm,
                     float intensity, int mode, uint8_t *out, int nthreads)
{
    orc_interference ic;
    orc_interference_new(&ic, n_film);
    const double dth = M_PI / (double)(num_theta - 1), dph = 2.0 * M_PI / (double)num_phi;
    if (nthreads < 1) nthreads = 1;
    /* per-column trig hoisted out of the cell loop; the values (and so the results) are those of
     * orc_view_cos evaluated per cell */
    double *sp = (double *)malloc((size_t)num_phi * sizeof(double));
    double *cp = (double *)malloc((size_t)num_phi * sizeof(double));
    for (int j = 0; j < num_phi; ++j) {
        sp[j] = sin((double)j * dph);
        cp[j] = cos((double)j * dph);
    }
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int r = 0; r < num_rows; ++r) {
        const double theta = (double)(row0 + r) * dth;
        const double st = sin(theta), ct = cos(theta);
        for (int j = 0; j < num_phi; ++j) {
            const size_t idx = (size_t)r * (size_t)num_phi + (size_t)j;
            const double n[3] = {st * cp[j], ct, st * sp[j]};
            const double v[3] = {cam[0] - radius * n[0], cam[1] - radius * n[1], cam[2] - radius * n[2]};
            const double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            const double c = fabs((v[0] * n[0] + v[1] * n[1] + v[2] * n[2]) / len);
            const double d_nm = h[idx] * 1e9;
            double rgb[3];
            if (mode == 1 || mode == 4) {
                float f[3];
                orc_lut_shade_w(lut, T, A, max_nm, (float)d_nm, (float)c, intensity, mode == 4 ? 8 : 0, f);
                rgb[0] = f[0]; rgb[1] = f[1]; rgb[2] = f[2];
            } else if (mode == 2) {
                orc_interference_color(&ic, d_nm, c, rgb);
            } else {
                orc_interference_color_linear(&ic, d_nm, c, rgb);
            }
            uint8_t px[3];
            orc_rgb_to_u8(rgb, px);
            out[4 * idx + 0] = px[0];
            out[4 * idx + 1] = px[1];
            out[4 * idx + 2] = px[2];
            out[4 * idx + 3] = 255;
        }
    }
    free(sp);
    free(cp);
}

/* ================================================================== caustic intensity map (f32) */

void orc_caustic_params_default(orc_caustic_params *p)
{
    memset(p, 0, sizeof(*p));
    p->grid_width = 128;
    p->grid_height = 64;
    p->refractive_index = 1.33f;
    p->film_thickness_scale = 1e6f;
    p->light_dir_y = -1.0f;
    p->light_intensity = 1.0f;
    p->focal_length = 0.1f;
    p->caustic_intensity = 2.0f;
    p->caustic_sharpness = 1.5f;
    p->branch_threshold = 0.001f;
    p->ground_y = -0.1f;
    p->bubble_radius = 0.025f;
}

static float caustic_sample(const orc_caustic_params *p, const float *t, int x, int y)
{ /* caustics_compute.wgsl:29-43 */
    int W = (int)p->grid_width, H = (int)p->grid_height;
    int sx = x % W;
    if (sx < 0) sx += W;
    int sy = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    return t[(size_t)sy * (size_t)W + (size_t)sx];
}

void orc_caustic_map(const orc_caustic_params *p, const float *t, float *out)
{
    const int W = (int)p->grid_width, H = (int)p->grid_height;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const float h = 1.0f, h2 = h * h;
            const float center = caustic_sample(p, t, x, y);
            const float left = caustic_sample(p, t, x - 1, y), right = caustic_sample(p, t, x + 1, y);
            const float up = caustic_sample(p, t, x, y - 1), down = caustic_sample(p, t, x, y + 1);
            const float gx = (right - left) / (2.0f * h), gy = (down - up) / (2.0f * h);          /* :46-51 */
            const float laplacian = (left + right + up + down - 4.0f * center) / h2;              /* :54-65 */
            const float d2x = (left + right - 2.0f * center) / h2;                                /* :68-92 */
            const float d2y = (up + down - 2.0f * center) / h2;
            const float ul = caustic_sample(p, t, x - 1, y - 1), ur = caustic_sample(p, t, x + 1, y - 1);
            const float dl = caustic_sample(p, t, x - 1, y + 1), dr = caustic_sample(p, t, x + 1, y + 1);
            const float dxy = (dr - dl - ur + ul) / (4.0f * h2);
            const float hessian_det = d2x * d2y - dxy * dxy;
            const float curvature_factor = -laplacian * p->focal_length;                          /* :95-112 */
            const float focusing = 1.0f / fmaxf(fabsf(1.0f - curvature_factor), 0.01f);
            const float branch_factor = hessian_det < -p->branch_threshold ? 1.5f : 1.0f;
            const float grad_mag = sqrtf(gx * gx + gy * gy);
            const float refraction_factor = 1.0f + grad_mag * 0.5f;
            float intensity = focusing * branch_factor * refraction_factor;
            intensity = powf(intensity, p->caustic_sharpness);
            intensity *= p->caustic_intensity;
            out[(size_t)y * (size_t)W + (size_t)x] = intensity < 0.f ? 0.f : (intensity > 10.f ? 10.f : intensity);
        }
    }
}

/* ================================================================== branched-flow sampling side (f32) */

static float sample_uv(const float *t, int W, int H, float u, float v)
{ /* branched_flow_compute.wgsl:187-192 */
    float cu = u < 0.f ? 0.f : (u > 1.f ? 1.f : u), cv = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    uint32_t x = (uint32_t)(cu * (float)(W - 1)), y = (uint32_t)(cv * (float)(H - 1));
    return t[(size_t)y * (size_t)W + (size_t)x];
}

void orc_sample_thickness_uv(const float *t, int W, int H, const float *uv, size_t n, float *out_h, float *out_grad)
{
    const float eps = 0.01f;
    for (size_t i = 0; i < n; ++i) {
        const float u = uv[2 * i], v = uv[2 * i + 1];
        if (out_h) out_h[i] = sample_uv(t, W, H, u, v);
        if (out_grad) { /* :196-210 */
            const float h_right = sample_uv(t, W, H, u + eps, v), h_left = sample_uv(t, W, H, u - eps, v);
            const float h_up = sample_uv(t, W, H, u, v + eps), h_down = sample_uv(t, W, H, u, v - eps);
            out_grad[2 * i] = (h_right - h_left) / (2.0f * eps);
            out_grad[2 * i + 1] = (h_up - h_down) / (2.0f * eps);
        }
    }
}

void orc_sample_thickness_normal(const float *t, int W, int H, const float *nrm, size_t n, float scale, float *out_h,
                                 float *out_uv)
{
    const float PI_F = 3.14159265359f; /* branched_flow_compute.wgsl PI constant */
    for (size_t i = 0; i < n; ++i) {
        const float nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2];
        const float phi = atan2f(nz, nx);
        const float cy = ny < -1.f ? -1.f : (ny > 1.f ? 1.f : ny);
        const float theta = acosf(cy);
        const float u = (phi + PI_F) / (2.0f * PI_F), v = theta / PI_F;
        uint32_t x = (uint32_t)(u * (float)(W - 1)), y = (uint32_t)(v * (float)(H - 1));
        if (x > (uint32_t)(W - 1)) x = (uint32_t)(W - 1); /* u can round to 1 + 1 ulp; the shader would read OOB */
        if (y > (uint32_t)(H - 1)) y = (uint32_t)(H - 1);
        out_h[i] = t[(size_t)y * (size_t)W + x] * scale;
        if (out_uv) { out_uv[2 * i] = u; out_uv[2 * i + 1] = v; }
    }
}
