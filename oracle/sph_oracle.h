/*
 * CPU oracle for the TinySPH compute-rank timestep.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call anything in oracle/.  The product (sph_b200/, include/)
 * never links it and has no CPU path.
 *
 * Two restatements, both plain C99, both built with -O2 -ffp-contract=off (the flags
 * that define "the reference" for parity):
 *
 *  orc_seq_*   the reference's algorithm as written: in-place, pair-by-pair
 *              (Gauss-Seidel) sweeps over forward-half neighbour lists, owners N-1..0,
 *              fixed-capacity buckets (100) and lists (400).  Pinned BIT-EXACT against
 *              the unmodified reference compiled into oracle/_ref/libsph_ref.so
 *              (tests/test_oracle_pin.py) and against tests/golden/ made from it.
 *
 *  orc_g_*     the same pair physics as a race-free gather (Jacobi) over cell-sorted
 *              SoA arrays, slab by slab, mirroring include/sph_b200.h call for call.
 *              This is what the CUDA kernels are checked against at rounding-level
 *              tolerance; its distance from orc_seq_* is the algorithmic
 *              (order-of-update) difference the north star allows for, measured and
 *              bounded in tests/test_oracle_gather.py.
 *
 * Parity status: PINNED (reference has no golden vectors of its own; the
 * pin is the reference itself, compiled from /root/reference/src by oracle/ref_build).
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stdint.h>
#include "../include/sph_b200.h"   /* record layouts only */

/* ------------------------------------------------------------------ orc_seq */

typedef struct orc_seq {
    int n, cap;
    float tank_w, tank_h;
    sph_tunable t;
    float *x, *y, *xp, *yp, *vx, *vy, *dens, *densn, *press, *pressn;
    /* hash grid (hash.h:42-48) */
    float spacing;
    unsigned size_x, size_y;
    int max_bucket, max_nbr;
    int *bcount, *bitems;     /* cells x max_bucket, insertion order */
    int *ncount, *nitems;     /* n x max_nbr forward lists */
} orc_seq;

orc_seq *orc_seq_create(int cap, float tank_w, float tank_h, const sph_tunable *t);
void orc_seq_destroy(orc_seq *s);
void orc_seq_load(orc_seq *s, const sph_particle *aos, int n);
void orc_seq_store(const orc_seq *s, sph_particle *aos);

unsigned orc_hash_val(float x, float y, float spacing, unsigned size_x);       /* hash.c:35-47 */
void orc_boundary(float *x, float *y, float tank_w, float tank_h, const sph_tunable *t); /* fluid.c:656-744 */
void orc_check_velocity(float *vx, float *vy);                                /* fluid.c:613-625 */

void orc_seq_apply_gravity(orc_seq *s);            /* fluid.c:398-413 */
void orc_seq_viscosity(orc_seq *s);                /* fluid.c:416-478 */
void orc_seq_predict(orc_seq *s);                  /* fluid.c:507-523 */
void orc_seq_hash(orc_seq *s, int compute_density);/* hash.c:127-242 */
void orc_seq_relax(orc_seq *s);                    /* fluid.c:541-611 */
void orc_seq_update_velocities(orc_seq *s);        /* fluid.c:642-653 */
void orc_seq_step(orc_seq *s, const sph_tunable *queued); /* fluid.c:270-348, one rank */

/* ---- host-side geometry / partition / balancing restatements ---- */
/* geometry.c:101-160 for every rank at once; returns N_global actually used */
int orc_partition(float tank_w, float water_min_x, float water_max_x, float water_min_y, float water_max_y,
                  float spacing, int nranks, int *start_col, int *ncols, float *start_x, float *end_x);
/* geometry.c:29-59 + fluid.c:762-767: lattice fill of one rank's columns; returns count */
int orc_lattice(float water_min_x, float water_min_y, float water_max_y, float spacing,
                int start_col, int ncols, int total_cols, sph_particle *out, uint32_t *uid);
/* renderer.c:427-477 */
void orc_balance(sph_tunable *master, int nactive, const int *coord_counts, int total);
/* fluid.c:144 */
float orc_spacing(float water_w, float water_h, int n_request);

/* ------------------------------------------------------------------ orc_g (mirrors sph_b200.h) */

typedef struct orc_g orc_g;

orc_g *orc_g_create(const sph_config *cfg);
void orc_g_destroy(orc_g *g);
void orc_g_set_params(orc_g *g, const sph_tunable *t);
void orc_g_queue_params(orc_g *g, const sph_tunable *t);
void orc_g_set_edges(orc_g *g, float start_x, float end_x);
/* PROPOSAL, off by default (gamma = 0): symmetric damping of the viscosity gather where the plain Jacobi
 * sum overshoots (stiff presets); see the comment at the definition. */
void orc_g_set_viscosity_stabilisation(orc_g *g, float gamma);
void orc_g_set_viscosity_stabilisation_ex(orc_g *g, float gamma, float min_dt_sigma);
int orc_g_upload(orc_g *g, const sph_particle *aos, const uint32_t *uid, int n);
int orc_g_download(orc_g *g, sph_particle *aos, uint32_t *uid, int order, int include_halo);
void orc_g_advect(orc_g *g);
void orc_g_sort(orc_g *g);
void orc_g_density(orc_g *g);
void orc_g_relax(orc_g *g);
void orc_g_step(orc_g *g, int n);
void orc_g_exchange_buffers(orc_g *g, int which, void **send_left, void **recv_left,
                            void **send_right, void **recv_right, size_t *bytes);
void orc_g_set_neighbors(orc_g *g, int has_left, int has_right);
void orc_g_get_status(orc_g *g, sph_status *out);
int orc_g_get_cells(orc_g *g, uint32_t *uid, uint32_t *cell, int cap);
long long orc_g_get_pairs(orc_g *g, uint64_t *pairs, long long cap);
int orc_g_get_forward_counts(orc_g *g, uint32_t *uid, int *count, int cap);
int orc_g_pack_coords(orc_g *g, int16_t *xy, int cap);

#endif
