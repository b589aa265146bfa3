/*
 * CPU oracle for the TinySPH compute-rank timestep -- see sph_oracle.h.
 * TEST INFRASTRUCTURE ONLY.  Build: gcc -std=c99 -O2 -ffp-contract=off -fno-fast-math.
 *
 * File:line citations are to AdamSimpson/SPH `src/`.
 */
#include "sph_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================
 *  shared per-particle pieces
 * ===================================================================== */

/* hash.c:35-47 -- fp32 divide, floor in double, truncation to unsigned, row-major */
unsigned orc_hash_val(float x, float y, float spacing, unsigned size_x)
{
    unsigned gx = (unsigned)floor(x / spacing);
    unsigned gy = (unsigned)floor(y / spacing);
    return gy * size_x + gx;
}

/* fluid.c:613-625 */
void orc_check_velocity(float *vx, float *vy)
{
    const float vmax = 5.0f;
    if (*vx > vmax) *vx = vmax; else if (*vx < -vmax) *vx = -vmax;
    if (*vy > vmax) *vy = vmax; else if (*vy < -vmax) *vy = -vmax;
}

/* fluid.c:656-744 -- mover push-out, then tank clamp (min -> 0 exactly, max -> max-0.001f) */
void orc_boundary(float *px, float *py, float tank_w, float tank_h, const sph_tunable *t)
{
    float cx = t->mover_center_x, cy = t->mover_center_y;
    float x = *px, y = *py;
    if (t->mover_type == SPH_SPHERE_MOVER) {                       /* :663-685 */
        float radius = t->mover_width * 0.5f;
        float d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
        if (d2 <= radius * radius && d2 > 0.0f) {
            float d = sqrt(d2);
            float nx = (cx - x) / d;
            float ny = (cy - y) / d;
            float pen = radius - d;
            x -= pen * nx;
            y -= pen * ny;
        }
    } else if (t->mover_type == SPH_RECTANGLE_MOVER) {             /* :688-727 */
        float hw = t->mover_width * 0.5;
        float hh = t->mover_height * 0.5;
        float rx = x - cx, ry = y - cy;
        float ax = fabs(rx), ay = fabs(ry);
        if (ax < hw && ay < hh) {
            float penx = hw - ax, peny = hh - ay;
            if (penx < peny) { if (rx < 0.0f) x -= penx; else x += penx; }
            else             { if (ry < 0.0f) y -= peny; else y += peny; }
        }
    }
    if (x < 0.0f) x = 0.0f; else if (x > tank_w) x = tank_w - 0.001f;   /* :732-743 */
    if (y < 0.0f) y = 0.0f; else if (y > tank_h) y = tank_h - 0.001f;
    *px = x; *py = y;
}

/* fluid.c:144 */
float orc_spacing(float water_w, float water_h, int n_request)
{
    float area = water_w * water_h;
    float s = pow(area / n_request, 1.0 / 2.0);
    return s;
}

/* =====================================================================
 *  orc_seq: the reference algorithm as written
 * ===================================================================== */

orc_seq *orc_seq_create(int cap, float tank_w, float tank_h, const sph_tunable *t)
{
    orc_seq *s = (orc_seq *)calloc(1, sizeof *s);
    s->cap = cap; s->tank_w = tank_w; s->tank_h = tank_h; s->t = *t;
    float **f[] = { &s->x, &s->y, &s->xp, &s->yp, &s->vx, &s->vy, &s->dens, &s->densn, &s->press, &s->pressn };
    for (unsigned i = 0; i < sizeof f / sizeof f[0]; i++) *f[i] = (float *)calloc(cap, sizeof(float));
    s->spacing = t->smoothing_radius;                                  /* fluid.c:176 */
    s->size_x = ceil((tank_w - 0.0f) / s->spacing);                    /* fluid.c:214-215 */
    s->size_y = ceil((tank_h - 0.0f) / s->spacing);
    s->max_bucket = SPH_REF_MAX_BUCKET;                                /* fluid.c:174-175 */
    s->max_nbr = SPH_REF_MAX_NEIGHBORS;
    size_t cells = (size_t)s->size_x * s->size_y;
    s->bcount = (int *)calloc(cells, sizeof(int));
    s->bitems = (int *)calloc(cells * s->max_bucket, sizeof(int));
    s->ncount = (int *)calloc(cap, sizeof(int));
    s->nitems = (int *)calloc((size_t)cap * s->max_nbr, sizeof(int));
    return s;
}

void orc_seq_destroy(orc_seq *s)
{
    if (!s) return;
    free(s->x); free(s->y); free(s->xp); free(s->yp); free(s->vx); free(s->vy);
    free(s->dens); free(s->densn); free(s->press); free(s->pressn);
    free(s->bcount); free(s->bitems); free(s->ncount); free(s->nitems);
    free(s);
}

void orc_seq_load(orc_seq *s, const sph_particle *a, int n)
{
    s->n = n;
    for (int i = 0; i < n; i++) {
        s->xp[i] = a[i].x_prev; s->yp[i] = a[i].y_prev; s->x[i] = a[i].x; s->y[i] = a[i].y;
        s->vx[i] = a[i].v_x; s->vy[i] = a[i].v_y; s->dens[i] = a[i].density; s->densn[i] = a[i].density_near;
        s->press[i] = a[i].pressure; s->pressn[i] = a[i].pressure_near;
        s->ncount[i] = 0;
    }
}

void orc_seq_store(const orc_seq *s, sph_particle *a)
{
    for (int i = 0; i < s->n; i++) {
        a[i].x_prev = s->xp[i]; a[i].y_prev = s->yp[i]; a[i].x = s->x[i]; a[i].y = s->y[i];
        a[i].v_x = s->vx[i]; a[i].v_y = s->vy[i]; a[i].a_x = 0.0f; a[i].a_y = 0.0f;
        a[i].density = s->dens[i]; a[i].density_near = s->densn[i];
        a[i].pressure = s->press[i]; a[i].pressure_near = s->pressn[i]; a[i].id = i;
    }
}

/* fluid.c:398-413 (single rank: no halo) */
void orc_seq_apply_gravity(orc_seq *s)
{
    float dt = s->t.time_step, g = -s->t.g;
    for (int i = 0; i < s->n; i++) {
        s->vy[i] += g * dt;
        s->dens[i] = 0.0f;
        s->densn[i] = 0.0f;
    }
}

/* fluid.c:416-478: owners N-1..0, in place, listed q in list order */
void orc_seq_viscosity(orc_seq *s)
{
    float h_recip = 1.0f / s->t.smoothing_radius;
    float sigma = s->t.sigma, beta = s->t.beta, dt = s->t.time_step;
    for (int i = s->n; i-- > 0;) {
        float px = s->x[i], py = s->y[i];
        const int *list = &s->nitems[(size_t)i * s->max_nbr];
        for (int j = 0; j < s->ncount[i]; j++) {
            int q = list[j];
            float dx = s->x[q] - px, dy = s->y[q] - py;
            float r = sqrt(dx * dx + dy * dy);
            float r_recip = 1.0f / r;
            float ratio = r * h_recip;
            float u = ((s->vx[i] - s->vx[q]) * dx + (s->vy[i] - s->vy[q]) * dy) * r_recip;
            if (u > 0.0f) {
                float imp = dt * (1 - ratio) * (sigma * u + beta * u * u);
                float ix = imp * dx * r_recip, iy = imp * dy * r_recip;
                orc_check_velocity(&ix, &iy);
                s->vx[i] -= ix * 0.5f; s->vy[i] -= iy * 0.5f;
                s->vx[q] += ix * 0.5f; s->vy[q] += iy * 0.5f;   /* single rank: q is never a halo (:464) */
            }
        }
    }
}

/* fluid.c:507-523 */
void orc_seq_predict(orc_seq *s)
{
    float dt = s->t.time_step;
    for (int i = 0; i < s->n; i++) {
        s->xp[i] = s->x[i]; s->yp[i] = s->y[i];
        s->x[i] += s->vx[i] * dt;
        s->y[i] += s->vy[i] * dt;
        orc_boundary(&s->x[i], &s->y[i], s->tank_w, s->tank_h, &s->t);
    }
}

/* fluid.c:527-539 */
static void seq_density_pair(orc_seq *s, int p, int q, float ratio)
{
    float omr2 = (1.0f - ratio) * (1.0f - ratio);
    if (ratio < 1.0f) {
        s->dens[p] += omr2; s->densn[p] += omr2 * (1.0f - ratio);
        s->dens[q] += omr2; s->densn[q] += omr2 * (1.0f - ratio);
    }
}

/* list append + optional density for one candidate pair; `owner` is the particle whose
 * forward list receives `other` (hash.c:183-196 same cell, :219-231 forward cells).  The
 * squared distance is sign-symmetric and the two density accumulators are distinct
 * variables, so only the ORDER OF PAIRS matters for bit-exactness, not argument order. */
static void seq_try_pair(orc_seq *s, int owner, int other, float h2, float h_recip, int compute_density)
{
    float dx = s->x[owner] - s->x[other], dy = s->y[owner] - s->y[other];
    float r2 = dx * dx + dy * dy;
    if (r2 > h2) return;
    if (s->ncount[owner] < s->max_nbr) {
        s->nitems[(size_t)owner * s->max_nbr + s->ncount[owner]++] = other;
        if (compute_density) {
            float r = sqrt(r2);
            float ratio = r * h_recip;
            seq_density_pair(s, owner, other, ratio);
        }
    }
}

/* hash.c:127-242 */
void orc_seq_hash(orc_seq *s, int compute_density)
{
    float h = s->t.smoothing_radius, h_recip = 1.0f / h, h2 = h * h;
    size_t cells = (size_t)s->size_x * s->size_y;
    for (size_t c = 0; c < cells; c++) s->bcount[c] = 0;                       /* :148-150 */
    for (int i = 0; i < s->n; i++) {                                          /* :153-166 */
        s->ncount[i] = 0;
        unsigned c = orc_hash_val(s->x[i], s->y[i], s->spacing, s->size_x);
        if (s->bcount[c] < s->max_bucket) s->bitems[(size_t)c * s->max_bucket + s->bcount[c]++] = i;
    }
    for (int j = 0; j < (int)s->size_y; j++)                                   /* :169-240 */
        for (int i = 0; i < (int)s->size_x; i++) {
            size_t c = (size_t)j * s->size_x + i;
            int nc = s->bcount[c];
            if (nc == 0) continue;
            const int *items = &s->bitems[c * s->max_bucket];
            for (int a = 0; a < nc; a++)                                       /* :178-199 same cell, a<b */
                for (int b = a + 1; b < nc; b++)
                    seq_try_pair(s, items[a], items[b], h2, h_recip, compute_density);
            for (int dx = 0; dx <= 1; dx++)                                    /* :203-237 forward cells */
                for (int dy = (dx ? -1 : 1); dy <= 1; dy++) {
                    if (j + dy < 0 || i + dx < 0 || i + dx >= (int)s->size_x || j + dy >= (int)s->size_y) continue;
                    size_t nb = (size_t)(j + dy) * s->size_x + (i + dx);
                    const int *nitems = &s->bitems[nb * s->max_bucket];
                    for (int a = 0; a < nc; a++)
                        for (int b = 0; b < s->bcount[nb]; b++)
                            seq_try_pair(s, items[a], nitems[b], h2, h_recip, compute_density);
                }
        }
}

/* fluid.c:541-611 */
void orc_seq_relax(orc_seq *s)
{
    float k = s->t.k, k_near = s->t.k_near, k_spring = s->t.k_spring;
    float h = s->t.smoothing_radius, h_recip = 1.0f / h, dt = s->t.time_step, rest = s->t.rest_density;
    for (int i = 0; i < s->n; i++) {                                          /* :560-565 */
        s->press[i] = k * (s->dens[i] - rest);
        s->pressn[i] = k_near * s->densn[i];
    }
    for (int i = s->n; i-- > 0;) {                                            /* :568-610 */
        float pp = s->press[i], ppn = s->pressn[i];
        const int *list = &s->nitems[(size_t)i * s->max_nbr];
        for (int j = 0; j < s->ncount[i]; j++) {
            int q = list[j];
            float r = sqrt((s->x[i] - s->x[q]) * (s->x[i] - s->x[q]) + (s->y[i] - s->y[q]) * (s->y[i] - s->y[q]));
            float r_recip = 1.0f / r;
            float ratio = r * h_recip;
            float omr = 1.0f - ratio;
            if (r <= 0.000001f) { s->x[i] += 0.000001f; s->y[i] += 0.000001f; }   /* :583-586 */
            if (ratio < 1.0f && r > 0.0f) {
                /* :591 -- the 0.5 literal is double: last add and the dt*dt product run in fp64 */
                float D = dt * dt * ((pp + s->press[q]) * omr + (ppn + s->pressn[q]) * omr * omr + k_spring * (h - r) * 0.5);
                float Dx = D * (s->x[q] - s->x[i]) * r_recip;
                float Dy = D * (s->y[q] - s->y[i]) * r_recip;
                s->x[q] += Dx; s->y[q] += Dy;
                s->x[i] -= Dx; s->y[i] -= Dy;
            }
        }
    }
}

/* fluid.c:627-653 */
void orc_seq_update_velocities(orc_seq *s)
{
    float dt = s->t.time_step;
    for (int i = 0; i < s->n; i++) {
        orc_boundary(&s->x[i], &s->y[i], s->tank_w, s->tank_h, &s->t);
        float vx = (s->x[i] - s->xp[i]) / dt, vy = (s->y[i] - s->yp[i]) / dt;
        orc_check_velocity(&vx, &vy);
        s->vx[i] = vx; s->vy[i] = vy;
    }
}

/* fluid.c:270-348 on one rank (halo/OOB calls are no-ops there) */
void orc_seq_step(orc_seq *s, const sph_tunable *queued)
{
    orc_seq_apply_gravity(s);
    orc_seq_viscosity(s);
    orc_seq_predict(s);
    if (queued) s->t = *queued;            /* MPI_Scatterv lands here at sub_step 3 (:293-294) */
    orc_seq_hash(s, 1);
    orc_seq_relax(s);
    orc_seq_update_velocities(s);
    orc_seq_hash(s, 0);
}

/* =====================================================================
 *  host-side geometry / partition / balancing
 * ===================================================================== */

/* geometry.c:101-160 evaluated for all ranks */
int orc_partition(float tank_w, float water_min_x, float water_max_x, float water_min_y, float water_max_y,
                  float spacing, int nranks, int *start_col, int *ncols, float *start_x, float *end_x)
{
    int cols = floor((water_max_x - water_min_x) / spacing) + 1;
    int equal = floor(cols / nranks);
    int remaining = cols - equal * nranks;
    int left = 0, total = 0;
    for (int r = 0; r < nranks; r++) {
        int len = equal + (r < remaining ? 1 : 0);
        start_col[r] = left; ncols[r] = len;
        start_x[r] = water_min_x + ((left - 1) * spacing);
        end_x[r] = start_x[r] + (len * spacing);
        if (r == 0) start_x[r] = 0.0f;
        if (r == nranks - 1) end_x[r] = tank_w;
        left += len; total += len;
    }
    int num_y = floor((water_max_y - water_min_y) / spacing);
    return total * num_y;
}

/* geometry.c:29-59, fluid.c:762-767 */
int orc_lattice(float water_min_x, float water_min_y, float water_max_y, float spacing,
                int start_col, int ncols, int total_cols, sph_particle *out, uint32_t *uid)
{
    int num_y = floor((water_max_y - water_min_y) / spacing);
    int i = 0;
    for (int ny = 0; ny < num_y; ny++) {
        float y = water_min_y + ny * spacing;
        for (int nx = 0; nx < ncols; nx++) {
            float x = water_min_x + (start_col + nx) * spacing;
            memset(&out[i], 0, sizeof out[i]);
            out[i].x = x; out[i].y = y; out[i].id = i;
            if (uid) uid[i] = (uint32_t)(ny * total_cols + start_col + nx);
            i++;
        }
    }
    return i;
}

/* renderer.c:427-477 */
void orc_balance(sph_tunable *m, int nactive, const int *counts, int total)
{
    int even = total / nactive;
    int max_diff = even / 15.0f;
    float h = m[0].smoothing_radius;
    float dx = h * 0.125;
    for (int r = nactive; r-- > 1;) {
        float len = m[r].node_end_x - m[r].node_start_x;
        float len_left = m[r - 1].node_end_x - m[r - 1].node_start_x;
        int diff = counts[r] - even;
        if (diff > max_diff && len > 2 * h) { m[r].node_start_x += dx; m[r - 1].node_end_x = m[r].node_start_x; }
        else if (diff < -max_diff && len_left > 2 * h) { m[r].node_start_x -= dx; m[r - 1].node_end_x = m[r].node_start_x; }
    }
    if (nactive > 1) {
        float len = m[0].node_end_x - m[0].node_start_x;
        float len_right = m[1].node_end_x - m[1].node_start_x;
        int diff = counts[0] - even;
        if (diff > max_diff && len > 2 * h) { m[0].node_end_x -= dx; m[1].node_start_x = m[0].node_end_x; }
        else if (diff < -max_diff && len_right > 2 * h) { m[0].node_end_x += dx; m[1].node_start_x = m[0].node_end_x; }
    }
}

/* =====================================================================
 *  orc_g: gather (Jacobi) form over cell-sorted SoA, slab by slab.
 *  Mirrors include/sph_b200.h; SURVEY.md Appendix B gives the formulas.
 * ===================================================================== */

#define DIV SPH_CELL_DIV          /* sort-grid refinement shared with the CUDA library (include/sph_b200.h) */
#define HALO_BIT 0x80000000u
#define UID_MASK 0x7fffffffu
#define KEY_DROP (-1)

enum { ST_READY = 0, ST_ADVECTED, ST_SORTED1, ST_DENSITY, ST_RELAXED };

struct orc_g {
    sph_config cfg;
    sph_tunable t, queued;
    int have_queued;
    float edge_start, edge_end;
    int has_left, has_right;
    int size_x, size_y;          /* the reference's grid (cells of side h) */
    int sort_rows;               /* rows of the sort grid = DIV * size_y */
    int gx0, wx;                 /* window of SORT-grid columns covered by this slab */
    int cap, msg_cap;
    /* A: cell-sorted resident state; q = velocity (ST_READY) or previous position (after sort 1) */
    float *ax, *ay, *aqx, *aqy; uint32_t *auid;
    int n_tot, n_local;
    /* T: staging in source order */
    float *tx, *ty, *tqx, *tqy; uint32_t *tuid; int *tkey;
    int n_src;
    float *dens, *densn;
    float *coupling;             /* per entry: sum of its pairs' viscosity coefficients (stabilised mode only) */
    float visc_gamma;            /* 0 = the plain Jacobi gather; > 0 = stabilised, see orc_g_set_viscosity_stabilisation */
    float visc_min_dt_sigma;     /* the stabilised pass engages for parameter blocks with dt*sigma at or above this */
    int *cell_start;
    int stage;
    unsigned char *send[2], *recv[2];   /* [0]=left, [1]=right */
    size_t msg_bytes;
    sph_status st;
};

/* message layout (shared with the CUDA library): 16-byte header, then SoA sections */
static size_t msg_bytes_for(int m) { return 16 + (size_t)m * 32; }
static int *msg_hdr(unsigned char *b) { return (int *)b; }
static float *msg_mig_pos(unsigned char *b, int m) { (void)m; return (float *)(b + 16); }
static float *msg_mig_q(unsigned char *b, int m) { return (float *)(b + 16 + (size_t)m * 8); }
static uint32_t *msg_mig_uid(unsigned char *b, int m) { return (uint32_t *)(b + 16 + (size_t)m * 16); }
static float *msg_halo_pos(unsigned char *b, int m) { return (float *)(b + 16 + (size_t)m * 20); }
static uint32_t *msg_halo_uid(unsigned char *b, int m) { return (uint32_t *)(b + 16 + (size_t)m * 28); }
/* exchange 1 reuses: halo pos at mig_pos, halo vel at mig_q, halo uid at mig_uid */

orc_g *orc_g_create(const sph_config *cfg)
{
    orc_g *g = (orc_g *)calloc(1, sizeof *g);
    g->cfg = *cfg;
    if (g->cfg.halo_width <= 0.0f) g->cfg.halo_width = 2.0f;
    g->cap = cfg->capacity; g->msg_cap = cfg->msg_capacity > 0 ? cfg->msg_capacity : 1;
    g->size_x = (int)ceil((cfg->tank_w - 0.0f) / cfg->h);
    g->size_y = (int)ceil((cfg->tank_h - 0.0f) / cfg->h);
    float **f[] = { &g->ax, &g->ay, &g->aqx, &g->aqy, &g->tx, &g->ty, &g->tqx, &g->tqy, &g->dens, &g->densn };
    for (unsigned i = 0; i < sizeof f / sizeof f[0]; i++) *f[i] = (float *)calloc(g->cap, sizeof(float));
    g->auid = (uint32_t *)calloc(g->cap, 4); g->tuid = (uint32_t *)calloc(g->cap, 4);
    g->tkey = (int *)calloc(g->cap, sizeof(int));
    g->coupling = (float *)calloc(g->cap, sizeof(float));
    g->visc_gamma = 0.5f; g->visc_min_dt_sigma = 0.5f;       /* the product library's defaults (sph_create) */
    g->sort_rows = g->size_y * DIV;
    g->cell_start = (int *)calloc((size_t)g->size_x * g->size_y * DIV * DIV + 1, sizeof(int));
    g->msg_bytes = msg_bytes_for(g->msg_cap);
    for (int s = 0; s < 2; s++) {
        g->send[s] = (unsigned char *)calloc(1, g->msg_bytes);
        g->recv[s] = (unsigned char *)calloc(1, g->msg_bytes);
    }
    g->edge_start = 0.0f; g->edge_end = cfg->tank_w;
    g->has_left = cfg->rank > 0; g->has_right = cfg->rank < cfg->nranks - 1;
    g->gx0 = 0; g->wx = g->size_x * DIV;
    g->stage = ST_READY;
    return g;
}

void orc_g_destroy(orc_g *g)
{
    if (!g) return;
    free(g->ax); free(g->ay); free(g->aqx); free(g->aqy); free(g->auid);
    free(g->tx); free(g->ty); free(g->tqx); free(g->tqy); free(g->tuid); free(g->tkey);
    free(g->dens); free(g->densn); free(g->cell_start); free(g->coupling);
    for (int s = 0; s < 2; s++) { free(g->send[s]); free(g->recv[s]); }
    free(g);
}

void orc_g_set_params(orc_g *g, const sph_tunable *t)
{
    g->t = *t; g->edge_start = t->node_start_x; g->edge_end = t->node_end_x;
}
void orc_g_queue_params(orc_g *g, const sph_tunable *t) { g->queued = *t; g->have_queued = 1; }
void orc_g_set_edges(orc_g *g, float s, float e) { g->edge_start = s; g->edge_end = e; }
void orc_g_set_neighbors(orc_g *g, int l, int r) { g->has_left = l; g->has_right = r; }

void orc_g_exchange_buffers(orc_g *g, int which, void **sl, void **rl, void **sr, void **rr, size_t *bytes)
{
    *sl = g->send[0]; *rl = g->recv[0]; *sr = g->send[1]; *rr = g->recv[1];
    *bytes = which == 0 ? g->msg_bytes : 16 + (size_t)g->msg_cap * 20;
}

/* window of grid columns this slab can touch: slab + ghost layer + one spare column */
static void g_window(orc_g *g)
{
    if (g->cfg.nranks <= 1) { g->gx0 = 0; g->wx = g->size_x * DIV; return; }
    float w = g->cfg.halo_width * g->cfg.h;
    int lo = (int)floor((g->edge_start - w) / g->cfg.h) - 1;
    int hi = (int)floor((g->edge_end + w) / g->cfg.h) + 1;
    if (lo < 0) lo = 0;
    if (hi > g->size_x - 1) hi = g->size_x - 1;
    if (hi < lo) hi = lo;
    g->gx0 = lo * DIV; g->wx = (hi - lo + 1) * DIV;
}

/* coordinate in the sort grid: the reference's x/h quotient scaled by DIV (exact for a power of two) */
static int g_sort_coord(const orc_g *g, float v)
{
    float q = v / g->cfg.h;
    return (int)floor(q * (float)DIV);
}

/* key of a position inside the current window, or KEY_DROP when outside it */
static int g_key(const orc_g *g, float x, float y)
{
    int gx = g_sort_coord(g, x);
    int gy = g_sort_coord(g, y);
    int wxi = gx - g->gx0;
    if (wxi < 0 || wxi >= g->wx || gy < 0 || gy >= g->sort_rows) return KEY_DROP;
    return gy * g->wx + wxi;
}

typedef struct { int key; uint32_t uid; int src; } sort_rec;
static int cmp_rec(const void *a, const void *b)
{
    const sort_rec *p = (const sort_rec *)a, *q = (const sort_rec *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    uint32_t up = p->uid & UID_MASK, uq = q->uid & UID_MASK;
    if (up != uq) return up < uq ? -1 : 1;
    return 0;
}

/* T (n_src entries with keys) -> A, ordered by (cell key, uid); builds cell_start */
static void g_sort_T_into_A(orc_g *g)
{
    sort_rec *rec = (sort_rec *)malloc(sizeof(sort_rec) * (size_t)(g->n_src > 0 ? g->n_src : 1));
    int m = 0;
    for (int i = 0; i < g->n_src; i++)
        if (g->tkey[i] != KEY_DROP) { rec[m].key = g->tkey[i]; rec[m].uid = g->tuid[i]; rec[m].src = i; m++; }
    qsort(rec, m, sizeof(sort_rec), cmp_rec);
    int ncell = g->wx * g->sort_rows;
    int c = 0, nl = 0, maxb = 0, over = 0;
    g->cell_start[0] = 0;
    for (int d = 0; d < m; d++) {
        while (c < rec[d].key) g->cell_start[++c] = d;
        int s = rec[d].src;
        g->ax[d] = g->tx[s]; g->ay[d] = g->ty[s]; g->aqx[d] = g->tqx[s]; g->aqy[d] = g->tqy[s];
        g->auid[d] = g->tuid[s];
        if (!(g->tuid[s] & HALO_BIT)) nl++;
    }
    while (c < ncell) g->cell_start[++c] = m;
    /* statistics of the REFERENCE's buckets (cells of side h = DIV x DIV sort cells), hash.c:160-165 */
    for (int R = 0; R < g->sort_rows / DIV; R++)
        for (int C = 0; C < g->wx / DIV; C++) {
            int cnt = 0;
            for (int r = 0; r < DIV; r++) {
                int row = R * DIV + r;
                cnt += g->cell_start[row * g->wx + (C + 1) * DIV] - g->cell_start[row * g->wx + C * DIV];
            }
            if (cnt > maxb) maxb = cnt;
            if (cnt > SPH_REF_MAX_BUCKET) over++;
        }
    g->n_tot = m; g->n_local = nl;
    g->st.max_bucket = maxb; g->st.bucket_overflow += over;
    free(rec);
}

static int g_append_T(orc_g *g, float x, float y, float qx, float qy, uint32_t uid)
{
    if (g->n_src >= g->cap) { g->st.capacity_overflow++; return -1; }
    int i = g->n_src++;
    g->tx[i] = x; g->ty[i] = y; g->tqx[i] = qx; g->tqy[i] = qy; g->tuid[i] = uid;
    g->tkey[i] = g_key(g, x, y);
    /* a ghost outside this slab's window is simply not needed (parked slab, controls.c:405-426) */
    if (g->tkey[i] == KEY_DROP && !(uid & HALO_BIT)) g->st.capacity_overflow++;
    return i;
}

int orc_g_upload(orc_g *g, const sph_particle *a, const uint32_t *uid, int n)
{
    if (n > g->cap) return -SPH_ERR_CAPACITY;
    g_window(g);
    g->n_src = 0;
    for (int i = 0; i < n; i++)
        g_append_T(g, a[i].x, a[i].y, a[i].v_x, a[i].v_y, uid ? (uid[i] & UID_MASK) : (uint32_t)i);
    g_sort_T_into_A(g);
    g->stage = ST_READY;
    return SPH_OK;
}

/* iterate the (2 DIV + 1)^2 sort-cell neighbourhood of window cell (wxi, gy): one contiguous index range per row */
#define FOR_EACH_CANDIDATE(g, wxi, gy, j, BODY)                                          \
    for (int _dy = -DIV; _dy <= DIV; _dy++) {                                            \
        int _row = (gy) + _dy;                                                           \
        if (_row < 0 || _row >= (g)->sort_rows) continue;                                \
        int _c0 = (wxi) > DIV ? (wxi) - DIV : 0;                                         \
        int _c1 = (wxi) < (g)->wx - DIV ? (wxi) + DIV : (g)->wx - 1;                     \
        int _b = (g)->cell_start[_row * (g)->wx + _c0];                                  \
        int _e = (g)->cell_start[_row * (g)->wx + _c1 + 1];                              \
        for (int j = _b; j < _e; j++) { BODY }                                           \
    }

static void g_cell_of(const orc_g *g, int i, int *wxi, int *gy)
{
    int key = g_key(g, g->ax[i], g->ay[i]);
    *wxi = key % g->wx; *gy = key / g->wx;
}

static void g_pack_halo0(orc_g *g, int side, float x, float y, uint32_t uid)
{
    int *hdr = msg_hdr(g->send[side]);
    if (hdr[1] >= g->msg_cap) { g->st.msg_overflow++; return; }
    int k = hdr[1]++;
    float *p = msg_halo_pos(g->send[side], g->msg_cap);
    p[2 * k] = x; p[2 * k + 1] = y;
    msg_halo_uid(g->send[side], g->msg_cap)[k] = uid;
}

/* Stabilised viscosity gather (PROPOSAL, off by default; the CUDA path does not implement it yet).
 *
 * The reference applies the viscosity impulses pair by pair IN PLACE (fluid.c:442-472): every pair sees the
 * velocities the pairs before it left behind, which damps a pair's approach speed by the factor
 * 1 - dt (1-q)(sigma + beta u) and never overshoots.  The plain gather sums all of a particle's impulses from
 * FROZEN velocities; with C_i = sum_j dt (1-q_ij)(sigma + beta u_ij) over its approaching pairs, the particle's
 * velocity changes by about C_i / 2 times its approach speed, and for C_i >> 1 that overshoots and feeds
 * itself: with the "goo" preset (sigma 100, beta 10, controls.c:359-371: dt sigma = 0.83 per pair) the fluid
 * never settles.  Here every pair's impulse is scaled by
 *     s_ij = 1 / max(1, gamma * max(C_i, C_j)),
 * which is symmetric in (i, j) -- momentum is still exchanged pairwise -- and equal to 1 wherever the gather
 * was stable anyway (the default fluid: C_i ~ 0.3), so those results do not change by a bit.
 * gamma = 0.5 reproduced the reference's long-run statistics for all four presets within the reference's
 * own sensitivity to the particle order. */
void orc_g_set_viscosity_stabilisation(orc_g *g, float gamma) { g->visc_gamma = gamma; g->visc_min_dt_sigma = 0.0f; }
/* with the threshold of sph_set_viscosity_stabilisation; the default of both libraries is (0.5, 0.5): the pass
 * engages by itself for the goo preset (dt*sigma = 0.83, controls.c:359-371) and for no other preset */
void orc_g_set_viscosity_stabilisation_ex(orc_g *g, float gamma, float min_dt_sigma)
{
    g->visc_gamma = gamma; g->visc_min_dt_sigma = min_dt_sigma;
}

static void g_viscosity_coupling(orc_g *g)
{
    const sph_tunable *t = &g->t;
    float dt = t->time_step, gdt = (-t->g) * dt;
    float h = t->smoothing_radius, h_recip = 1.0f / h, h2 = h * h;
    float sigma = t->sigma, beta = t->beta;
    for (int i = 0; i < g->n_tot; i++) {
        float px = g->ax[i], py = g->ay[i];
        float vix = g->aqx[i], viy = g->aqy[i] + gdt;
        float c = 0.0f;
        int wxi, gy; g_cell_of(g, i, &wxi, &gy);
        FOR_EACH_CANDIDATE(g, wxi, gy, j,
            if (j == i) continue;
            float dx = g->ax[j] - px; float dy = g->ay[j] - py;
            float r2 = dx * dx + dy * dy;
            if (r2 > h2) continue;
            float r = sqrtf(r2);
            float u = ((vix - g->aqx[j]) * dx + (viy - (g->aqy[j] + gdt)) * dy) * (1.0f / r);
            if (u > 0.0f) c += dt * (1 - r * h_recip) * (sigma + beta * u);
        )
        g->coupling[i] = c;
    }
}

/* gravity + viscosity gather + predict + boundary + migration/halo classification */
void orc_g_advect(orc_g *g)
{
    const sph_tunable *t = &g->t;
    float dt = t->time_step, gdt = (-t->g) * dt;
    float h = t->smoothing_radius, h_recip = 1.0f / h, h2 = h * h;
    float sigma = t->sigma, beta = t->beta;
    g->n_src = g->n_tot;
    for (int s = 0; s < 2; s++) { int *hd = msg_hdr(g->send[s]); hd[0] = hd[1] = hd[2] = hd[3] = 0; }
    g->st.migrated_left = g->st.migrated_right = 0;
    const float gamma = (g->visc_gamma > 0.0f && dt * t->sigma >= g->visc_min_dt_sigma) ? g->visc_gamma : 0.0f;
    if (gamma > 0.0f) g_viscosity_coupling(g);

    for (int i = 0; i < g->n_tot; i++) {
        g->tuid[i] = g->auid[i];
        g->tqx[i] = g->ax[i]; g->tqy[i] = g->ay[i];          /* x_prev (fluid.c:515-516) */
        if (g->auid[i] & HALO_BIT) { g->tkey[i] = KEY_DROP; g->tx[i] = g->ax[i]; g->ty[i] = g->ay[i]; continue; }
        float px = g->ax[i], py = g->ay[i];
        float vix = g->aqx[i], viy = g->aqy[i] + gdt;        /* apply_gravity (fluid.c:407) */
        float vx = vix, vy = viy;
        int wxi, gy; g_cell_of(g, i, &wxi, &gy);
        FOR_EACH_CANDIDATE(g, wxi, gy, j,
            if (j == i) continue;
            float dx = g->ax[j] - px; float dy = g->ay[j] - py;
            float r2 = dx * dx + dy * dy;
            if (r2 > h2) continue;                           /* list membership (hash.c:185,221,99) */
            float r = sqrtf(r2);
            float r_recip = 1.0f / r;
            float ratio = r * h_recip;
            float vjx = g->aqx[j]; float vjy = g->aqy[j] + gdt;
            float u = ((vix - vjx) * dx + (viy - vjy) * dy) * r_recip;
            if (u > 0.0f) {                                  /* fluid.c:451-462 */
                float imp = dt * (1 - ratio) * (sigma * u + beta * u * u);
                if (gamma > 0.0f) {
                    float cmax = g->coupling[i] > g->coupling[j] ? g->coupling[i] : g->coupling[j];
                    if (gamma * cmax > 1.0f) imp = imp / (gamma * cmax);
                }
                float ix = imp * dx * r_recip; float iy = imp * dy * r_recip;
                orc_check_velocity(&ix, &iy);
                vx -= ix * 0.5f; vy -= iy * 0.5f;
            }
        )
        float nx = px + vx * dt, ny = py + vy * dt;          /* fluid.c:517-518 */
        orc_boundary(&nx, &ny, g->cfg.tank_w, g->cfg.tank_h, t);
        g->tx[i] = nx; g->ty[i] = ny;
    }

    /* the render rank's parameter scatter lands here (fluid.c:293-294) */
    if (g->have_queued) { orc_g_set_params(g, &g->queued); g->have_queued = 0; }
    g_window(g);

    /* identify_oob_particles (fluid.c:481-498) + ghost-layer selection (communication.c:134-141,
     * widened to halo_width*h and tested independently per side) */
    float w = g->cfg.halo_width * g->cfg.h;
    for (int i = 0; i < g->n_tot; i++) {
        if (g->auid[i] & HALO_BIT) continue;
        float x = g->tx[i], y = g->ty[i];
        int side = -1;
        if (x < g->edge_start && g->has_left) side = 0;
        else if (x > g->edge_end && g->has_right) side = 1;
        if (side >= 0) {
            int *hdr = msg_hdr(g->send[side]);
            if (hdr[0] >= g->msg_cap) { g->st.msg_overflow++; }
            else {
                int k = hdr[0]++;
                float *p = msg_mig_pos(g->send[side], g->msg_cap), *q = msg_mig_q(g->send[side], g->msg_cap);
                p[2 * k] = x; p[2 * k + 1] = y; q[2 * k] = g->tqx[i]; q[2 * k + 1] = g->tqy[i];
                msg_mig_uid(g->send[side], g->msg_cap)[k] = g->tuid[i];
                if (side == 0) g->st.migrated_left++; else g->st.migrated_right++;
                g->tuid[i] |= HALO_BIT;       /* stays resident as a ghost of its new owner */
            }
        } else {
            if (g->has_left && x - g->edge_start <= w) g_pack_halo0(g, 0, x, y, g->tuid[i]);
            if (g->has_right && g->edge_end - x <= w) g_pack_halo0(g, 1, x, y, g->tuid[i]);
        }
        g->tkey[i] = g_key(g, x, y);
        /* an emigrant pushed out of this slab's window (mover) is simply not kept as a ghost */
        if (g->tkey[i] == KEY_DROP && !(g->tuid[i] & HALO_BIT)) g->st.capacity_overflow++;
    }
    g->stage = ST_ADVECTED;
}

/* unpack neighbour messages (if any) into T, then counting sort T -> A */
void orc_g_sort(orc_g *g)
{
    int which = g->stage == ST_ADVECTED ? 0 : 1;
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? g->has_left : g->has_right)) continue;
        unsigned char *b = g->recv[side];
        int *hdr = msg_hdr(b);
        if (which == 0) {
            float *p = msg_mig_pos(b, g->msg_cap), *q = msg_mig_q(b, g->msg_cap);
            uint32_t *u = msg_mig_uid(b, g->msg_cap);
            for (int k = 0; k < hdr[0]; k++) g_append_T(g, p[2 * k], p[2 * k + 1], q[2 * k], q[2 * k + 1], u[k] & UID_MASK);
            float *hp = msg_halo_pos(b, g->msg_cap); uint32_t *hu = msg_halo_uid(b, g->msg_cap);
            for (int k = 0; k < hdr[1]; k++) g_append_T(g, hp[2 * k], hp[2 * k + 1], 0.0f, 0.0f, hu[k] | HALO_BIT);
        } else {
            float *p = msg_mig_pos(b, g->msg_cap), *q = msg_mig_q(b, g->msg_cap);
            uint32_t *u = msg_mig_uid(b, g->msg_cap);
            for (int k = 0; k < hdr[1]; k++) g_append_T(g, p[2 * k], p[2 * k + 1], q[2 * k], q[2 * k + 1], u[k] | HALO_BIT);
        }
    }
    g_sort_T_into_A(g);
    g->stage = which == 0 ? ST_SORTED1 : ST_READY;
    if (g->stage == ST_READY) g->st.steps++;
}

/* fluid.c:527-539 as a gather: every resident particle sums over its 3x3 neighbourhood */
void orc_g_density(orc_g *g)
{
    float h = g->t.smoothing_radius, h_recip = 1.0f / h, h2 = h * h;
    for (int i = 0; i < g->n_tot; i++) {
        float px = g->ax[i], py = g->ay[i];
        float d = 0.0f, dn = 0.0f;
        int wxi, gy; g_cell_of(g, i, &wxi, &gy);
        FOR_EACH_CANDIDATE(g, wxi, gy, j,
            if (j == i) continue;
            float dx = g->ax[j] - px; float dy = g->ay[j] - py;
            float r2 = dx * dx + dy * dy;
            if (r2 > h2) continue;
            float ratio = sqrtf(r2) * h_recip;
            if (ratio < 1.0f) {
                float omr = 1.0f - ratio;
                float omr2 = omr * omr;
                d += omr2; dn += omr2 * omr;
            }
        )
        g->dens[i] = d; g->densn[i] = dn;
    }
    g->stage = ST_DENSITY;
}

/* ownership rule for the coincident-particle nudge (hash.c:178-224), in REFERENCE cells: same cell ->
 * earlier bucket slot (= lower uid on one rank); otherwise the particle whose forward stencil
 * (0,+1),(1,-1),(1,0),(1,+1) holds the other */
static int g_owns(const orc_g *g, int i, int j)
{
    int xi = (int)floor(g->ax[i] / g->cfg.h), yi = (int)floor(g->ay[i] / g->cfg.h);
    int xj = (int)floor(g->ax[j] / g->cfg.h), yj = (int)floor(g->ay[j] / g->cfg.h);
    if (xi == xj && yi == yj) return (g->auid[i] & UID_MASK) < (g->auid[j] & UID_MASK);
    if (xi != xj) return xi < xj;
    return yi < yj;
}

/* double_density_relaxation (fluid.c:541-611) as a gather + updateVelocities (fluid.c:642-653) */
void orc_g_relax(orc_g *g)
{
    const sph_tunable *t = &g->t;
    float k = t->k, k_near = t->k_near, k_spring = t->k_spring, rest = t->rest_density;
    float h = t->smoothing_radius, h_recip = 1.0f / h, h2 = h * h, dt = t->time_step;
    float w = g->cfg.halo_width * g->cfg.h;
    g->n_src = g->n_tot;
    for (int s = 0; s < 2; s++) { int *hd = msg_hdr(g->send[s]); hd[0] = hd[1] = hd[2] = hd[3] = 0; }
    for (int i = 0; i < g->n_tot; i++) {
        g->tuid[i] = g->auid[i];
        if (g->auid[i] & HALO_BIT) { g->tkey[i] = KEY_DROP; continue; }
        float px = g->ax[i], py = g->ay[i];
        float pp = k * (g->dens[i] - rest), ppn = k_near * g->densn[i];     /* :563-564 */
        float x = px, y = py;
        int wxi, gy; g_cell_of(g, i, &wxi, &gy);
        FOR_EACH_CANDIDATE(g, wxi, gy, j,
            if (j == i) continue;
            float dx = g->ax[j] - px; float dy = g->ay[j] - py;
            float r2 = dx * dx + dy * dy;
            if (r2 > h2) continue;
            float r = sqrtf(r2);
            float r_recip = 1.0f / r;
            float ratio = r * h_recip;
            float omr = 1.0f - ratio;
            if (r <= 0.000001f && g_owns(g, i, j)) { x += 0.000001f; y += 0.000001f; }   /* :583-586 */
            if (ratio < 1.0f && r > 0.0f) {
                float pq = k * (g->dens[j] - rest); float pqn = k_near * g->densn[j];
                float D = dt * dt * ((pp + pq) * omr + (ppn + pqn) * omr * omr + k_spring * (h - r) * 0.5);  /* :591 */
                x -= D * dx * r_recip; y -= D * dy * r_recip;
            }
        )
        orc_boundary(&x, &y, g->cfg.tank_w, g->cfg.tank_h, t);               /* :649 */
        float vx = (x - g->aqx[i]) / dt, vy = (y - g->aqy[i]) / dt;          /* :632-633 */
        orc_check_velocity(&vx, &vy);
        g->tx[i] = x; g->ty[i] = y; g->tqx[i] = vx; g->tqy[i] = vy;
        g->tkey[i] = g_key(g, x, y);
        if (g->tkey[i] == KEY_DROP) g->st.capacity_overflow++;
        /* second ghost exchange (fluid.c:337): relaxed position + new velocity */
        for (int side = 0; side < 2; side++) {
            int on = side == 0 ? (g->has_left && x - g->edge_start <= w) : (g->has_right && g->edge_end - x <= w);
            if (!on) continue;
            int *hdr = msg_hdr(g->send[side]);
            if (hdr[1] >= g->msg_cap) { g->st.msg_overflow++; continue; }
            int m = hdr[1]++;
            float *p = msg_mig_pos(g->send[side], g->msg_cap), *q = msg_mig_q(g->send[side], g->msg_cap);
            p[2 * m] = x; p[2 * m + 1] = y; q[2 * m] = vx; q[2 * m + 1] = vy;
            msg_mig_uid(g->send[side], g->msg_cap)[m] = g->tuid[i];
        }
    }
    g->stage = ST_RELAXED;
}

void orc_g_step(orc_g *g, int n)
{
    for (int s = 0; s < n; s++) {
        orc_g_advect(g); orc_g_sort(g); orc_g_density(g); orc_g_relax(g); orc_g_sort(g);
    }
}

void orc_g_get_status(orc_g *g, sph_status *out)
{
    g->st.n_local = g->n_local; g->st.n_halo = g->n_tot - g->n_local;
    *out = g->st;
}

typedef struct { uint32_t uid; int idx; } uid_idx;
static int cmp_uid(const void *a, const void *b)
{
    uint32_t p = ((const uid_idx *)a)->uid & UID_MASK, q = ((const uid_idx *)b)->uid & UID_MASK;
    return p < q ? -1 : (p > q ? 1 : 0);
}

int orc_g_download(orc_g *g, sph_particle *a, uint32_t *uid, int order, int include_halo)
{
    uid_idx *v = (uid_idx *)malloc(sizeof(uid_idx) * (size_t)(g->n_tot > 0 ? g->n_tot : 1));
    int m = 0;
    for (int i = 0; i < g->n_tot; i++)
        if (include_halo || !(g->auid[i] & HALO_BIT)) { v[m].uid = g->auid[i]; v[m].idx = i; m++; }
    if (order == SPH_ORDER_UID) qsort(v, m, sizeof(uid_idx), cmp_uid);
    int has_prev = g->stage == ST_SORTED1 || g->stage == ST_DENSITY;
    int has_dens = g->stage == ST_DENSITY;
    for (int k = 0; k < m; k++) {
        int i = v[k].idx;
        memset(&a[k], 0, sizeof a[k]);
        a[k].x = g->ax[i]; a[k].y = g->ay[i];
        if (has_prev) { a[k].x_prev = g->aqx[i]; a[k].y_prev = g->aqy[i]; }
        else { a[k].x_prev = g->ax[i]; a[k].y_prev = g->ay[i]; a[k].v_x = g->aqx[i]; a[k].v_y = g->aqy[i]; }
        if (has_dens) {
            a[k].density = g->dens[i]; a[k].density_near = g->densn[i];
            a[k].pressure = g->t.k * (g->dens[i] - g->t.rest_density);
            a[k].pressure_near = g->t.k_near * g->densn[i];
        }
        a[k].id = k;
        if (uid) uid[k] = g->auid[i];
    }
    free(v);
    return m;
}

int orc_g_get_cells(orc_g *g, uint32_t *uid, uint32_t *cell, int cap)
{
    int m = 0;
    for (int i = 0; i < g->n_tot; i++) {
        if (g->auid[i] & HALO_BIT) continue;
        if (m < cap) { uid[m] = g->auid[i]; cell[m] = orc_hash_val(g->ax[i], g->ay[i], g->cfg.h, (unsigned)g->size_x); }
        m++;
    }
    return m;
}

long long orc_g_get_pairs(orc_g *g, uint64_t *pairs, long long cap)
{
    float h = g->t.smoothing_radius, h2 = h * h;
    long long m = 0;
    for (int i = 0; i < g->n_tot; i++) {
        int wxi, gy; g_cell_of(g, i, &wxi, &gy);
        uint32_t ui = g->auid[i] & UID_MASK;
        FOR_EACH_CANDIDATE(g, wxi, gy, j,
            uint32_t uj = g->auid[j] & UID_MASK;
            if (uj <= ui) continue;
            float dx = g->ax[i] - g->ax[j]; float dy = g->ay[i] - g->ay[j];
            if (dx * dx + dy * dy > h2) continue;
            if (m < cap) pairs[m] = ((uint64_t)ui << 32) | uj;
            m++;
        )
    }
    return m;
}

int orc_g_get_forward_counts(orc_g *g, uint32_t *uid, int *count, int cap)
{
    float h = g->t.smoothing_radius, h2 = h * h;
    int m = 0;
    for (int i = 0; i < g->n_tot; i++) {
        if (g->auid[i] & HALO_BIT) continue;
        int wxi, gy; g_cell_of(g, i, &wxi, &gy);
        int c = 0;
        FOR_EACH_CANDIDATE(g, wxi, gy, j,
            if (j == i) continue;
            float dx = g->ax[i] - g->ax[j]; float dy = g->ay[i] - g->ay[j];
            if (dx * dx + dy * dy > h2) continue;
            if ((g->auid[j] & HALO_BIT) || g_owns(g, i, j)) c++;
        )
        if (c > SPH_REF_MAX_NEIGHBORS) g->st.neighbor_overflow++;
        if (m < cap) { uid[m] = g->auid[i]; count[m] = c; }
        m++;
    }
    return m;
}

/* fluid.c:358-361 */
int orc_g_pack_coords(orc_g *g, int16_t *xy, int cap)
{
    int m = 0;
    for (int i = 0; i < g->n_tot; i++) {
        if (g->auid[i] & HALO_BIT) continue;
        if (m < cap) {
            xy[2 * m] = (2.0f * g->ax[i] / g->cfg.tank_w - 1.0f) * SHRT_MAX;
            xy[2 * m + 1] = (2.0f * g->ay[i] / g->cfg.tank_h - 1.0f) * SHRT_MAX;
        }
        m++;
    }
    return m;
}
