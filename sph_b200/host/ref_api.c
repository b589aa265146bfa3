/*
 * The reference's entry points (include/sph_ref_api.h) on top of the C ABI (include/sph_b200.h).
 *
 * The reference's step is eleven calls (fluid.c:273-348); the GPU path is five stages.  The mapping:
 *
 *   apply_gravity, viscosity_impluses   noted; they run fused with the prediction
 *   predict_positions                   -> sph_advect   (gravity + viscosity + predict + boundary)
 *   identify_oob_particles              one rank: nothing leaves the slab
 *   hash_fluid(compute_density)         -> sph_sort [+ sph_density]
 *   start/finishHaloExchange, hash_halo one rank: no neighbours
 *   double_density_relaxation           noted; runs fused with the velocity update
 *   updateVelocities                    -> sph_relax    (relaxation + boundary + velocity)
 *
 * Parameter changes need no special path: every call carries `param *`, and a block that differs
 * from the last one pushed is sent to the device before the stage is enqueued, which reproduces
 * where the render rank's scatter lands (between predict_positions and the later calls,
 * fluid.c:279-310).
 */
#include "sph_ref_api.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static struct {
    sph_ctx *ctx;
    sph_tunable pushed;
    int have_pushed;
    int pending_gravity, pending_viscosity, pending_relax;
    int n;
    float tank_w, tank_h;
    char err[256];
} G;

static void note(const char *where, int rc)
{
    if (rc == SPH_OK) return;
    snprintf(G.err, sizeof G.err, "%s: error %d: %s", where, rc, G.ctx ? sph_last_error(G.ctx) : "not attached");
    fprintf(stderr, "sph_ref_api: %s\n", G.err);
}

const char *sph_ref_last_error(void) { return G.err; }
sph_ctx *sph_ref_context(void) { return G.ctx; }

static void sync_params(const char *where, const param *params)
{
    if (!G.ctx) { note(where, SPH_ERR_STATE); return; }
    if (G.have_pushed && memcmp(&G.pushed, &params->tunable_params, sizeof(sph_tunable)) == 0) return;
    G.pushed = params->tunable_params;
    G.have_pushed = 1;
    note(where, sph_set_params(G.ctx, &G.pushed));
}

int sph_ref_attach(fluid_particle **pointers, param *params, AABB_t *boundary, neighbor_grid_t *grid, int device)
{
    sph_ref_detach();
    memset(&G, 0, sizeof G);
    const int n = params->number_fluid_particles_local;
    sph_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.tank_w = boundary->max_x; cfg.tank_h = boundary->max_y;
    cfg.h = grid->spacing;
    /* the reference sizes every rank for the whole problem (fluid.c:156) */
    cfg.capacity = params->number_fluid_particles_global > n ? params->number_fluid_particles_global : n;
    if (cfg.capacity < 1) cfg.capacity = 1;
    cfg.msg_capacity = 1;
    cfg.device = device; cfg.rank = 0; cfg.nranks = 1;
    int rc = sph_create(&cfg, &G.ctx);
    if (rc) { note("sph_ref_attach", rc); if (G.ctx) { sph_destroy(G.ctx); G.ctx = NULL; } return rc; }
    G.tank_w = cfg.tank_w; G.tank_h = cfg.tank_h;
    sync_params("sph_ref_attach", params);
    sph_particle *flat = (sph_particle *)malloc((size_t)(n > 0 ? n : 1) * sizeof(sph_particle));
    for (int i = 0; i < n; i++) flat[i] = *pointers[i];
    rc = sph_upload(G.ctx, flat, NULL, n);           /* uid = pointer index */
    free(flat);
    G.n = n;
    note("sph_ref_attach", rc);
    return rc;
}

void sph_ref_detach(void)
{
    if (G.ctx) sph_destroy(G.ctx);
    G.ctx = NULL;
}

int sph_ref_sync_to_host(fluid_particle **pointers, param *params)
{
    if (!G.ctx) return SPH_ERR_STATE;
    sph_particle *flat = (sph_particle *)malloc((size_t)(G.n > 0 ? G.n : 1) * sizeof(sph_particle));
    int n = sph_download(G.ctx, flat, NULL, SPH_ORDER_UID, 0);
    if (n < 0) { free(flat); note("sph_ref_sync_to_host", -n); return -n; }
    for (int i = 0; i < n; i++) { *pointers[i] = flat[i]; pointers[i]->id = i; }
    free(flat);
    params->number_fluid_particles_local = n;
    params->number_halo_particles = 0;
    return SPH_OK;
}

int sph_ref_pack_coords(short *coords, int max_pairs)
{
    if (!G.ctx) return -SPH_ERR_STATE;
    return sph_pack_coords(G.ctx, (int16_t *)coords, max_pairs);
}

/* ------------------------------------------------------------------ fluid.h */

void apply_gravity(fluid_particle **pointers, param *params)
{
    (void)pointers;
    sync_params("apply_gravity", params);
    G.pending_gravity = 1;
}

void viscosity_impluses(fluid_particle **pointers, neighbor *neighbors, param *params)
{
    (void)pointers; (void)neighbors;
    sync_params("viscosity_impluses", params);
    G.pending_viscosity = 1;
}

void predict_positions(fluid_particle **pointers, AABB_t *boundary_global, param *params)
{
    (void)pointers; (void)boundary_global;
    sync_params("predict_positions", params);
    if (!G.pending_gravity || !G.pending_viscosity) {
        snprintf(G.err, sizeof G.err, "predict_positions: the fused stage needs apply_gravity and "
                 "viscosity_impluses first (fluid.c:273-279 order)");
        fprintf(stderr, "sph_ref_api: %s\n", G.err);
        return;
    }
    G.pending_gravity = G.pending_viscosity = 0;
    note("predict_positions", sph_advect(G.ctx));
}

void identify_oob_particles(fluid_particle **pointers, fluid_particle *particles, oob_t *oob, AABB_t *b, param *params)
{
    (void)pointers; (void)particles; (void)oob; (void)b;
    sync_params("identify_oob_particles", params);     /* the scatter has landed by now (fluid.c:293-310) */
}

void double_density_relaxation(fluid_particle **pointers, neighbor *neighbors, param *params)
{
    (void)pointers; (void)neighbors;
    sync_params("double_density_relaxation", params);
    G.pending_relax = 1;
}

void updateVelocities(fluid_particle **pointers, edge_t *edges, AABB_t *boundary_global, param *params)
{
    (void)pointers; (void)edges; (void)boundary_global;
    sync_params("updateVelocities", params);
    if (!G.pending_relax) {
        snprintf(G.err, sizeof G.err, "updateVelocities: the fused stage needs double_density_relaxation first");
        fprintf(stderr, "sph_ref_api: %s\n", G.err);
        return;
    }
    G.pending_relax = 0;
    note("updateVelocities", sph_relax(G.ctx));
}

/* ------------------------------------------------------------------ hash.h */

void hash_fluid(fluid_particle **pointers, neighbor_grid_t *grid, param *params, bool compute_density)
{
    (void)pointers; (void)grid;
    sync_params("hash_fluid", params);
    int rc = sph_sort(G.ctx);
    note("hash_fluid", rc);
    if (rc == SPH_OK && compute_density) note("hash_fluid", sph_density(G.ctx));
}

void hash_halo(fluid_particle **pointers, neighbor_grid_t *grid, param *params, bool compute_density)
{
    (void)pointers; (void)grid; (void)params; (void)compute_density;   /* ghosts are binned by the sort */
}

unsigned int hash_val(float x, float y, neighbor_grid_t *grid, param *params)
{
    (void)params;
    const float cell = grid->spacing;                  /* hash.c:37-46 */
    const unsigned int col = (unsigned int)floor(x / cell);
    const unsigned int row = (unsigned int)floor(y / cell);
    return row * grid->size_x + col;
}

/* ------------------------------------------------------------------ communication.h (one rank) */

void startHaloExchange(fluid_particle **pointers, fluid_particle *particles, edge_t *edges, param *params)
{ (void)pointers; (void)particles; (void)edges; (void)params; }
void finishHaloExchange(fluid_particle **pointers, fluid_particle *particles, edge_t *edges, param *params)
{ (void)pointers; (void)particles; (void)edges; params->number_halo_particles = 0; }
void transferOOBParticles(fluid_particle **pointers, fluid_particle *particles, oob_t *oob, param *params)
{ (void)pointers; (void)particles; (void)oob; (void)params; }

/* ------------------------------------------------------------------ per-particle helpers (host) */

void checkVelocity(float *v_x, float *v_y)
{
    const float limit = 5.0f;                          /* fluid.c:615 */
    *v_x = *v_x > limit ? limit : (*v_x < -limit ? -limit : *v_x);
    *v_y = *v_y > limit ? limit : (*v_y < -limit ? -limit : *v_y);
}

void updateVelocity(fluid_particle *p, param *params)
{
    const float dt = params->tunable_params.time_step; /* fluid.c:629-638 */
    float vx = (p->x - p->x_prev) / dt, vy = (p->y - p->y_prev) / dt;
    checkVelocity(&vx, &vy);
    p->v_x = vx; p->v_y = vy;
}

void calculate_density(fluid_particle *p, fluid_particle *q, float ratio)
{
    if (!(ratio < 1.0f)) return;                       /* fluid.c:530-537 */
    const float w = 1.0f - ratio, w2 = w * w, w3 = w2 * w;
    p->density += w2; p->density_near += w3;
    q->density += w2; q->density_near += w3;
}

void boundaryConditions(fluid_particle *p, AABB_t *boundary, param *params)
{
    const sph_tunable *t = &params->tunable_params;
    const float cx = t->mover_center_x, cy = t->mover_center_y;
    if (t->mover_type == SPH_SPHERE_MOVER) {           /* fluid.c:663-685 */
        const float radius = t->mover_width * 0.5f;
        const float ox = p->x - cx, oy = p->y - cy;
        const float d2 = ox * ox + oy * oy;
        if (d2 <= radius * radius && d2 > 0.0f) {
            const float d = (float)sqrt(d2);
            const float depth = radius - d;
            const float nx = (cx - p->x) / d, ny = (cy - p->y) / d;
            p->x -= depth * nx;
            p->y -= depth * ny;
        }
    } else if (t->mover_type == SPH_RECTANGLE_MOVER) { /* fluid.c:688-727 */
        const float half_w = (float)(t->mover_width * 0.5), half_h = (float)(t->mover_height * 0.5);
        const float ox = p->x - cx, oy = p->y - cy;
        const float ax = (float)fabs(ox), ay = (float)fabs(oy);
        if (ax < half_w && ay < half_h) {
            const float depth_x = half_w - ax, depth_y = half_h - ay;
            if (depth_x < depth_y) p->x += ox < 0.0f ? -depth_x : depth_x;
            else p->y += oy < 0.0f ? -depth_y : depth_y;
        }
    }
    if (p->x < boundary->min_x) p->x = boundary->min_x;             /* fluid.c:732-743 */
    else if (p->x > boundary->max_x) p->x = boundary->max_x - 0.001f;
    if (p->y < boundary->min_y) p->y = boundary->min_y;
    else if (p->y > boundary->max_y) p->y = boundary->max_y - 0.001f;
}
