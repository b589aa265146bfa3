/*
 * Harness around the UNMODIFIED TinySPH reference (TEST INFRASTRUCTURE ONLY).
 *
 * The reference's fluid.c / hash.c / geometry.c / communication.c are compiled
 * where they lie (/root/reference/src) by oracle/ref_build/Makefile and linked
 * with this file and the mini-MPI shim into
 *     oracle/_ref/libsph_ref.so   (ctypes: pins oracle/sph_oracle.c, makes tests/golden/)
 *     oracle/_ref/sph_ref_run     (multi-rank CPU baseline + multi-rank fixtures)
 *
 * Why a harness instead of the reference's own start_simulation(): the problem
 * size (1500) and tank width (15.0f) are source literals there
 * (/root/reference/src/fluid.c:111-119) and the function needs a render rank.
 * This file re-creates ONLY the set-up and loop ORDER of start_simulation with
 * the size/tank as arguments; every physics, hashing, geometry and exchange
 * call goes to the reference's own functions, in the reference's order
 * (fluid.c:270-348).  The particle-count load balancer lives in the render rank
 * (renderer.c:427-477, which needs GL headers to compile), so its arithmetic is
 * restated here in refh_balance().
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include "mpi.h"
#include "fluid.h"
#include "hash.h"
#include "geometry.h"
#include "communication.h"

typedef struct {
    int n_request;          /* requested global particle count (fluid.c:113) */
    float tank_w, tank_h;   /* boundary_global.max_x / max_y (fluid.c:119,127) */
    float water_min_x, water_max_x, water_min_y, water_max_y; /* fluid.c:130-133 */
    float mover_cx, mover_cy, mover_w, mover_h;
    int mover_type;
    int steps_per_frame;    /* 4 (fluid.c:105) */
    int cap_factor;         /* particle array capacity = cap_factor * N_global (2 in fluid.c:156) */
} refh_config;

typedef struct {
    param params;
    AABB_t boundary, water;
    edge_t edges;
    oob_t oob;
    neighbor_grid_t grid;
    fluid_particle *particles;
    fluid_particle **pointers;
    neighbor *neighbors;
    fluid_particle **neighbor_store;
    bucket_t *buckets;
    fluid_particle **bucket_store;
    size_t capacity;
    float spacing;
    int start_x, number_particles_x, total_x;
    int steps_per_frame, sub_step;
    int have_pending;
    tunable_parameters pending;
    long step_count;
} refh_sim;

static int g_types_ready = 0;

refh_sim *refh_create(const refh_config *cfg)
{
    refh_sim *s = (refh_sim *)calloc(1, sizeof *s);
    if (!g_types_ready) { createMpiTypes(); g_types_ready = 1; }
    MPI_COMM_COMPUTE = MPI_COMM_WORLD;

    /* fluid.c:88-107 */
    tunable_parameters *t = &s->params.tunable_params;
    t->kill_sim = 0;
    t->active = 1;
    t->g = 6.0f;
    t->time_step = 1.0f / 30.0f;
    t->k = 0.2f;
    t->k_near = 6.0f;
    t->k_spring = 10.0f;
    t->sigma = 5.0f;
    t->beta = 0.5f;
    t->rest_density = 30.0f;
    t->mover_width = cfg->mover_w;
    t->mover_height = cfg->mover_h;
    t->mover_type = (char)cfg->mover_type;
    t->mover_center_x = cfg->mover_cx;   /* left uninitialised by the reference */
    t->mover_center_y = cfg->mover_cy;
    s->steps_per_frame = cfg->steps_per_frame;
    t->time_step /= (float)s->steps_per_frame;
    s->params.number_fluid_particles_global = cfg->n_request;

    /* fluid.c:116-135 */
    s->boundary.min_x = 0.0f; s->boundary.max_x = cfg->tank_w;
    s->boundary.min_y = 0.0f; s->boundary.max_y = cfg->tank_h;
    s->water.min_x = cfg->water_min_x; s->water.max_x = cfg->water_max_x;
    s->water.min_y = cfg->water_min_y; s->water.max_y = cfg->water_max_y;
    s->params.number_halo_particles = 0;

    /* fluid.c:141-144 */
    float area = (s->water.max_x - s->water.min_x) * (s->water.max_y - s->water.min_y);
    float spacing = pow(area / s->params.number_fluid_particles_global, 1.0 / 2.0);
    s->spacing = spacing;

    /* fluid.c:147-159 */
    partitionProblem(&s->boundary, &s->water, &s->start_x, &s->number_particles_x, spacing, &s->params);
    setParticleNumbers(&s->boundary, &s->water, &s->edges, &s->oob, s->number_particles_x, spacing, &s->params);
    s->total_x = (int)floor((s->water.max_x - s->water.min_x) / spacing) + 1;
    size_t cap = (size_t)cfg->cap_factor * (size_t)s->params.number_fluid_particles_global;
    s->capacity = cap;
    t->smoothing_radius = 2.0f * spacing;

    /* fluid.c:173-225 (sizes in size_t: the reference's unsigned products overflow above ~5.3M) */
    s->grid.max_bucket_size = 100;
    s->grid.max_neighbors = s->grid.max_bucket_size * 4;
    s->grid.spacing = t->smoothing_radius;
    s->particles = (fluid_particle *)malloc(cap * sizeof(fluid_particle));
    s->pointers = (fluid_particle **)malloc(cap * sizeof(fluid_particle *));
    s->neighbors = (neighbor *)calloc(cap, sizeof(neighbor));
    s->neighbor_store = (fluid_particle **)calloc(cap * s->grid.max_neighbors, sizeof(fluid_particle *));
    for (size_t i = 0; i < cap; i++)
        s->neighbors[i].fluid_neighbors = &s->neighbor_store[i * s->grid.max_neighbors];
    s->grid.neighbors = s->neighbors;
    s->grid.size_x = ceil((s->boundary.max_x - s->boundary.min_x) / s->grid.spacing);
    s->grid.size_y = ceil((s->boundary.max_y - s->boundary.min_y) / s->grid.spacing);
    size_t cells = (size_t)s->grid.size_x * s->grid.size_y;
    s->buckets = (bucket_t *)calloc(cells, sizeof(bucket_t));
    s->bucket_store = (fluid_particle **)calloc(cells * s->grid.max_bucket_size, sizeof(fluid_particle *));
    for (size_t i = 0; i < cells; i++)
        s->buckets[i].fluid_particles = &s->bucket_store[i * s->grid.max_bucket_size];
    s->grid.grid_buckets = s->buckets;
    if (!s->particles || !s->pointers || !s->neighbors || !s->neighbor_store || !s->buckets || !s->bucket_store) {
        fprintf(stderr, "refh_create: allocation failed\n");
        return NULL;
    }

    /* fluid.c:228-233 */
    s->edges.edge_pointers_left = (fluid_particle **)malloc(s->edges.max_edge_particles * sizeof(fluid_particle *));
    s->edges.edge_pointers_right = (fluid_particle **)malloc(s->edges.max_edge_particles * sizeof(fluid_particle *));
    s->oob.oob_pointer_indicies_left = (int *)malloc(s->oob.max_oob_particles * sizeof(int));
    s->oob.oob_pointer_indicies_right = (int *)malloc(s->oob.max_oob_particles * sizeof(int));
    s->oob.vacant_indicies = (int *)malloc(2 * (size_t)s->oob.max_oob_particles * sizeof(int));

    /* fluid.c:238 */
    initParticles(s->pointers, s->particles, &s->water, s->start_x, s->number_particles_x,
                  &s->edges, (int)cap, spacing, &s->params);

    /* persistent uid stashed in the dead a_x field (carried by Particletype, never read) */
    int nx = s->number_particles_x;
    for (int i = 0; i < s->params.number_fluid_particles_local; i++) {
        int32_t uid = (i / nx) * s->total_x + s->start_x + (i % nx);
        memcpy(&s->particles[i].a_x, &uid, 4);
    }
    s->sub_step = 0;
    return s;
}

void refh_destroy(refh_sim *s)
{
    if (!s) return;
    free(s->particles); free(s->pointers); free(s->neighbors); free(s->neighbor_store);
    free(s->buckets); free(s->bucket_store);
    free(s->edges.edge_pointers_left); free(s->edges.edge_pointers_right);
    free(s->oob.oob_pointer_indicies_left); free(s->oob.oob_pointer_indicies_right);
    free(s->oob.vacant_indicies);
    free(s);
}

/* Parameters delivered at sub_step == steps_per_frame-1, between predict and OOB (fluid.c:293-294). */
void refh_queue_params(refh_sim *s, const tunable_parameters *t) { s->pending = *t; s->have_pending = 1; }

/* One iteration of the loop at fluid.c:270-372 (non-RASPI build), render-rank messages removed. */
void refh_step(refh_sim *s)
{
    apply_gravity(s->pointers, &s->params);
    viscosity_impluses(s->pointers, s->neighbors, &s->params);
    predict_positions(s->pointers, &s->boundary, &s->params);
    if (s->sub_step == s->steps_per_frame - 1 && s->have_pending) {
        s->params.tunable_params = s->pending;
        s->have_pending = 0;
    }
    identify_oob_particles(s->pointers, s->particles, &s->oob, &s->boundary, &s->params);
    hash_fluid(s->pointers, &s->grid, &s->params, true);
    startHaloExchange(s->pointers, s->particles, &s->edges, &s->params);
    finishHaloExchange(s->pointers, s->particles, &s->edges, &s->params);
    hash_halo(s->pointers, &s->grid, &s->params, true);
    double_density_relaxation(s->pointers, s->neighbors, &s->params);
    updateVelocities(s->pointers, &s->edges, &s->boundary, &s->params);
    startHaloExchange(s->pointers, s->particles, &s->edges, &s->params);
    hash_fluid(s->pointers, &s->grid, &s->params, false);
    finishHaloExchange(s->pointers, s->particles, &s->edges, &s->params);
    hash_halo(s->pointers, &s->grid, &s->params, false);
    s->sub_step = (s->sub_step == s->steps_per_frame - 1) ? 0 : s->sub_step + 1;
    s->step_count++;
}

/* ---- accessors (pointer order == sweep order == bucket insertion order) ---- */
int refh_n_local(refh_sim *s) { return s->params.number_fluid_particles_local; }
int refh_n_halo(refh_sim *s) { return s->params.number_halo_particles; }
int refh_n_global(refh_sim *s) { return s->params.number_fluid_particles_global; }
float refh_spacing(refh_sim *s) { return s->spacing; }
param *refh_params(refh_sim *s) { return &s->params; }
AABB_t *refh_boundary(refh_sim *s) { return &s->boundary; }
neighbor_grid_t *refh_grid(refh_sim *s) { return &s->grid; }
fluid_particle **refh_pointers(refh_sim *s) { return s->pointers; }
fluid_particle *refh_particles(refh_sim *s) { return s->particles; }
neighbor *refh_neighbors(refh_sim *s) { return s->neighbors; }
edge_t *refh_edges(refh_sim *s) { return &s->edges; }
oob_t *refh_oob(refh_sim *s) { return &s->oob; }

/* out: n x 13 32-bit words (the 52-byte record), in pointer order */
void refh_get_state(refh_sim *s, void *out, int include_halo)
{
    int n = s->params.number_fluid_particles_local + (include_halo ? s->params.number_halo_particles : 0);
    for (int i = 0; i < n; i++)
        memcpy((char *)out + (size_t)i * sizeof(fluid_particle), s->pointers[i], sizeof(fluid_particle));
}

/* Replace the local particle set (single-rank use): in: n x 13 words; id is rewritten to the index. */
void refh_set_state(refh_sim *s, const void *in, int n)
{
    for (int i = 0; i < n; i++) {
        memcpy(&s->particles[i], (const char *)in + (size_t)i * sizeof(fluid_particle), sizeof(fluid_particle));
        s->particles[i].id = i;
        s->pointers[i] = &s->particles[i];
        s->neighbors[i].number_fluid_neighbors = 0;
    }
    for (size_t i = (size_t)n; i < s->capacity; i++) s->pointers[i] = NULL;
    s->params.number_fluid_particles_local = n;
    s->params.max_fluid_particle_index = n - 1;
    s->params.number_halo_particles = 0;
    s->oob.number_vacancies = 0;
}

/* forward neighbour lists as local indices (q->id); counts[i], flat lists back to back */
long refh_get_neighbor_lists(refh_sim *s, int *counts, int *flat, long flat_cap)
{
    long k = 0;
    int n = s->params.number_fluid_particles_local;
    for (int i = 0; i < n; i++) {
        neighbor *ne = &s->neighbors[i];
        counts[i] = ne->number_fluid_neighbors;
        for (int j = 0; j < ne->number_fluid_neighbors; j++) {
            if (flat && k < flat_cap) flat[k] = ne->fluid_neighbors[j]->id;
            k++;
        }
    }
    return k;
}

/* bucket contents as local indices, in insertion order */
long refh_get_buckets(refh_sim *s, int *counts, int *flat, long flat_cap)
{
    long k = 0;
    size_t cells = (size_t)s->grid.size_x * s->grid.size_y;
    for (size_t c = 0; c < cells; c++) {
        counts[c] = (int)s->buckets[c].number_fluid;
        for (unsigned j = 0; j < s->buckets[c].number_fluid; j++) {
            if (flat && k < flat_cap) flat[k] = s->buckets[c].fluid_particles[j]->id;
            k++;
        }
    }
    return k;
}

/*
 * Particle-count load balancer: arithmetic of check_partition_left
 * (/root/reference/src/renderer.c:427-477) on an array of per-rank parameter
 * blocks.  `counts` follow the reference in being coordinate counts
 * (2 x particles, renderer.c:280,290).
 */
void refh_balance(tunable_parameters *master, int nactive, const int *counts, int total)
{
    int rank, diff;
    float h, dx, length, length_left, length_right;
    int even_particles = total / nactive;
    int max_diff = even_particles / 15.0f;
    h = master[0].smoothing_radius;
    dx = h * 0.125;
    for (rank = nactive; rank-- > 1;) {
        length = master[rank].node_end_x - master[rank].node_start_x;
        length_left = master[rank - 1].node_end_x - master[rank - 1].node_start_x;
        diff = counts[rank] - even_particles;
        if (diff > max_diff && length > 2 * h) {
            master[rank].node_start_x += dx;
            master[rank - 1].node_end_x = master[rank].node_start_x;
        } else if (diff < -max_diff && length_left > 2 * h) {
            master[rank].node_start_x -= dx;
            master[rank - 1].node_end_x = master[rank].node_start_x;
        }
    }
    if (nactive > 1) {
        length = master[0].node_end_x - master[0].node_start_x;
        length_right = master[1].node_end_x - master[1].node_start_x;
        diff = counts[0] - even_particles;
        if (diff > max_diff && length > 2 * h) {
            master[0].node_end_x -= dx;
            master[1].node_start_x = master[0].node_end_x;
        } else if (diff < -max_diff && length_right > 2 * h) {
            master[0].node_end_x += dx;
            master[1].node_start_x = master[0].node_end_x;
        }
    }
}

/* ------------------------------------------------------------------ binary */
#ifdef REFH_MAIN

typedef struct {
    int ranks, n, steps, warmup, balance, quiet;
    float tank_w, tank_h, water_frac, mover_x_frac, mover_y_frac;
    const char *dump;
} cli_t;

static double arg_f(int argc, char **argv, const char *name, double dflt)
{
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return atof(argv[i + 1]);
    return dflt;
}
static const char *arg_s(int argc, char **argv, const char *name, const char *dflt)
{
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
    return dflt;
}

/* one compute rank: run, balance once per frame, time, dump */
static int run_rank(int rank, const cli_t *c)
{
    mini_mpi_bind(rank);
    refh_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_request = c->n;
    cfg.tank_w = c->tank_w; cfg.tank_h = c->tank_h;
    cfg.water_min_x = 0.0f; cfg.water_max_x = c->tank_w * c->water_frac;
    cfg.water_min_y = 0.0f; cfg.water_max_y = c->tank_h;
    cfg.mover_w = cfg.mover_h = 2.0f * c->tank_w / 15.0f;
    cfg.mover_cx = c->mover_x_frac * c->tank_w; cfg.mover_cy = c->mover_y_frac * c->tank_h;
    cfg.mover_type = SPHERE_MOVER;
    cfg.steps_per_frame = 4;
    cfg.cap_factor = 2;
    int out_fd = dup(1);   /* result line goes to the real stdout; the reference's printf chatter does not */
    if (c->quiet) { fflush(stdout); if (!freopen("/dev/null", "w", stdout)) return 2; }
    refh_sim *s = refh_create(&cfg);
    if (!s) return 3;

    int K = c->ranks;
    double *shared = mini_mpi_shared_doubles();   /* [0..K) counts, [K..3K) edges, [3K..4K) loop time */
    tunable_parameters *master = (tunable_parameters *)calloc(K, sizeof *master);
    int *counts = (int *)calloc(K, sizeof(int));
    shared[K + 2 * rank] = s->params.tunable_params.node_start_x;
    shared[K + 2 * rank + 1] = s->params.tunable_params.node_end_x;
    mini_mpi_barrier();
    for (int r = 0; r < K; r++) {
        master[r] = s->params.tunable_params;
        master[r].node_start_x = (float)shared[K + 2 * r];
        master[r].node_end_x = (float)shared[K + 2 * r + 1];
    }
    mini_mpi_barrier();

    double t0 = 0.0, t1 = 0.0;
    long total_steps = (long)c->warmup + c->steps;
    for (long it = 0; it < total_steps; it++) {
        if (it == c->warmup) { mini_mpi_barrier(); t0 = MPI_Wtime(); }
        /* frame boundary: render rank would gather counts and scatter edges (renderer.c:268-290) */
        if (c->balance && K > 1 && s->sub_step == s->steps_per_frame - 1) {
            shared[rank] = 2.0 * refh_n_local(s);
            mini_mpi_barrier();
            int total = 0;
            for (int r = 0; r < K; r++) { counts[r] = (int)shared[r]; total += counts[r]; }
            refh_balance(master, K, counts, total);
            mini_mpi_barrier();
            tunable_parameters mine = s->params.tunable_params;
            mine.node_start_x = master[rank].node_start_x;
            mine.node_end_x = master[rank].node_end_x;
            refh_queue_params(s, &mine);
        }
        refh_step(s);
    }
    mini_mpi_barrier();
    t1 = MPI_Wtime();

    if (c->dump) {
        char path[1024];
        snprintf(path, sizeof path, "%s.rank%d.bin", c->dump, rank);
        FILE *f = fopen(path, "wb");
        int n = refh_n_local(s);
        void *buf = malloc((size_t)n * sizeof(fluid_particle));
        refh_get_state(s, buf, 0);
        int hdr[4] = { n, refh_n_global(s), rank, K };
        float edges[2] = { s->params.tunable_params.node_start_x, s->params.tunable_params.node_end_x };
        fwrite(hdr, sizeof hdr, 1, f);
        fwrite(edges, sizeof edges, 1, f);
        fwrite(buf, sizeof(fluid_particle), n, f);
        fclose(f);
        free(buf);
    }
    if (rank == 0) {
        double secs = t1 - t0;
        fflush(stdout);
        dprintf(out_fd, "{\"impl\": \"reference\", \"ranks\": %d, \"n_global\": %d, \"steps\": %d, \"warmup\": %d, "
                     "\"seconds\": %.6f, \"particle_steps_per_s\": %.6e, \"h\": %.9g, \"balance\": %d}\n",
                K, refh_n_global(s), c->steps, c->warmup, secs,
                (double)refh_n_global(s) * c->steps / (secs > 0 ? secs : 1e-9),
                s->params.tunable_params.smoothing_radius, c->balance);
    }
    refh_destroy(s);
    free(master); free(counts);
    return 0;
}

int main(int argc, char **argv)
{
    cli_t c;
    c.ranks = (int)arg_f(argc, argv, "--ranks", 1);
    c.n = (int)arg_f(argc, argv, "--n", 1500);
    c.steps = (int)arg_f(argc, argv, "--steps", 100);
    c.warmup = (int)arg_f(argc, argv, "--warmup", 0);
    c.balance = (int)arg_f(argc, argv, "--balance", 1);
    c.quiet = (int)arg_f(argc, argv, "--quiet", 1);
    c.tank_w = (float)arg_f(argc, argv, "--tank-w", 15.0);
    c.tank_h = (float)arg_f(argc, argv, "--tank-h", 15.0 * 9.0 / 16.0);
    c.water_frac = (float)arg_f(argc, argv, "--water-frac", 1.0);
    c.mover_x_frac = (float)arg_f(argc, argv, "--mover-x-frac", 0.5);
    c.mover_y_frac = (float)arg_f(argc, argv, "--mover-y-frac", 0.35);
    c.dump = arg_s(argc, argv, "--dump", NULL);
    size_t ring = (size_t)arg_f(argc, argv, "--ring-mb", 64) << 20;
    if (mini_mpi_world_create(c.ranks, ring) != 0) { fprintf(stderr, "world create failed\n"); return 1; }
    fflush(NULL);
    pid_t *pids = (pid_t *)calloc(c.ranks, sizeof(pid_t));
    for (int r = 1; r < c.ranks; r++) {
        pids[r] = fork();
        if (pids[r] == 0) _exit(run_rank(r, &c));
    }
    int rc = run_rank(0, &c);
    for (int r = 1; r < c.ranks; r++) {
        int st = 0;
        waitpid(pids[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 4;
    }
    return rc;
}
#endif
