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
 *
 * More than one compute rank (one slab per rank, geometry.c:101-160).  The library does not link MPI; the
 * host's transport reaches it through sph_ref_set_transport() or, for a driver that is not touched at all,
 * through the weak hook sph_ref_host_mpi() that the glue object sph_b200/host/glue/sph_ref_mpi_glue.c
 * defines (compiled with the host's own mpi.h, added to the link line).  Then:
 *
 *   identify_oob_particles              -> exchange 0: migrants + ghost layer with predicted positions, one message
 *                                          per neighbour (transferOOBParticles, communication.c:249-450, and the
 *                                          first start/finishHaloExchange, :120-246, in one meeting)
 *   first start/finishHaloExchange      nothing left to do (hash_halo likewise: the sort bins the ghosts)
 *   startHaloExchange after updateVelocities -> exchange 1: relaxed position + velocity of the ghost layer
 *   hash_fluid(false)                   -> sph_sort, unpacks them
 *
 * Both go through sph_exchange_via_host, i.e. two send/receive pairs in the reference's own order
 * (communication.c:158-161).  New slab edges in a scattered block take effect at the next predict_positions:
 * the device has binned the predicted positions into the old window by the time the scatter lands.  Results
 * do not depend on where the edges are, so this shifts the balancer by one sub-step and nothing else.
 * The host mirror then also maintains the pointer array: slot i of the particle array holds the i-th local
 * particle in ascending uid, pointers past the local count are NULL (fluid.c:758-759).
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
    int mirror;                 /* keep the host AoS current (an unmodified driver reads it, fluid.c:358-362) */
    int mirror_every;           /* ... every N completed steps (1 = every step) */
    long steps_done;
    int rank, nranks;           /* what MPI_Comm_rank/size(MPI_COMM_COMPUTE) would say (geometry.c:105-108) */
    float edge_start, edge_end; /* several ranks: the slab edges in force on the device (they change at step boundaries only) */
    int relaxed;                /* several ranks: updateVelocities has run, the next startHaloExchange is exchange 1 */
    int capacity;
    char err[256];
} G = { .nranks = 1 };

/* what the host told us; survives attach / detach */
static struct {
    sph_sendrecv_fn fn;
    void *user;
    int asked;
    int x_start, len_x, total_x;    /* this rank's share of the lattice columns (partitionProblem) */
    fluid_particle *base;           /* the particle array behind the pointer array (identify_oob_particles carries it) */
} H;

/* Defined by the host's glue object, if there is one (sph_b200/host/glue/sph_ref_mpi_glue.c; declared in sph_ref_api.h). */
#pragma weak sph_ref_host_mpi

static void ask_host(void)
{
    if (H.asked) return;
    H.asked = 1;
    if (!sph_ref_host_mpi) return;
    int rank = 0, nranks = 1;
    sph_sendrecv_fn fn = NULL;
    void *user = NULL;
    if (sph_ref_host_mpi(&rank, &nranks, &fn, &user) == 0 && nranks >= 1 && rank >= 0 && rank < nranks) {
        G.rank = rank; G.nranks = nranks;
        H.fn = fn; H.user = user;
    }
}

void sph_ref_set_transport(sph_sendrecv_fn fn, void *user) { H.fn = fn; H.user = user; H.asked = 1; }

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
    sph_tunable t = params->tunable_params;
    if (G.nranks > 1) { t.node_start_x = G.edge_start; t.node_end_x = G.edge_end; }    /* new edges wait for predict_positions */
    if (G.have_pushed && memcmp(&G.pushed, &t, sizeof(sph_tunable)) == 0) return;
    G.pushed = t;
    G.have_pushed = 1;
    note(where, sph_set_params(G.ctx, &G.pushed));
}

void sph_ref_set_rank(int rank, int nranks) { G.rank = rank; G.nranks = nranks > 0 ? nranks : 1; }
void sph_ref_set_mirror(int every_n_steps) { G.mirror = every_n_steps > 0; G.mirror_every = every_n_steps > 0 ? every_n_steps : 1; }

int sph_ref_attach(fluid_particle **pointers, param *params, AABB_t *boundary, neighbor_grid_t *grid, int device)
{
    sph_ref_detach();
    ask_host();
    if (G.nranks > 1 && !H.fn) {
        /* without a transport start/finishHaloExchange and transferOOBParticles could do nothing: carrying on would
         * simulate every slab as if it were alone */
        snprintf(G.err, sizeof G.err, "sph_ref_attach: without a transport the reference-named entry points serve ONE compute rank "
                 "(got %d): link host/glue/sph_ref_mpi_glue.c, call sph_ref_set_transport, or use the handle API", G.nranks);
        fprintf(stderr, "sph_ref_api: %s\n", G.err);
        return SPH_ERR_STATE;
    }
    const int rank = G.rank, nranks = G.nranks, mirror = G.mirror, mirror_every = G.mirror_every;
    const int armed_g = G.pending_gravity, armed_v = G.pending_viscosity;    /* a lazy attach happens between arming and firing */
    memset(&G, 0, sizeof G);
    G.rank = rank; G.nranks = nranks; G.mirror = mirror; G.mirror_every = mirror_every ? mirror_every : 1;
    G.pending_gravity = armed_g; G.pending_viscosity = armed_v;
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
    if (nranks > 1) {
        /* A message carries the ghost layer (2 h of a column of fluid) plus the migrants of one step.  The reference
         * sizes its out-of-bounds buffers for the GLOBAL particle count (geometry.c:87), which is what lets
         * remove_partition drain a whole slab in one step (controls.c:405-426); so does this path while that stays
         * affordable (40 bytes per entry and buffer), else a slab drains at one capacity per step -- nobody is lost,
         * sph_status.msg_overflow says so.  SPH_REF_MSG_CAPACITY overrides. */
        const int rows = (int)ceilf(cfg.tank_h / cfg.h);
        cfg.msg_capacity = 30 * rows > 4096 ? 30 * rows : 4096;
        if (params->number_fluid_particles_global > cfg.msg_capacity)
            cfg.msg_capacity = params->number_fluid_particles_global < (4 << 20) ? params->number_fluid_particles_global : (4 << 20);
        if (getenv("SPH_REF_MSG_CAPACITY") && atoi(getenv("SPH_REF_MSG_CAPACITY")) > 0) cfg.msg_capacity = atoi(getenv("SPH_REF_MSG_CAPACITY"));
        cfg.capacity += 2 * cfg.msg_capacity;
        cfg.rank = rank; cfg.nranks = nranks;
        G.edge_start = params->tunable_params.node_start_x;
        G.edge_end = params->tunable_params.node_end_x;
    }
    G.capacity = cfg.capacity;
    int rc = sph_create(&cfg, &G.ctx);
    if (rc) { note("sph_ref_attach", rc); if (G.ctx) { sph_destroy(G.ctx); G.ctx = NULL; } return rc; }
    G.tank_w = cfg.tank_w; G.tank_h = cfg.tank_h;
    {   /* the stabilised viscosity gather is on by default (0.5, 0.5: engages for the goo preset only, sph_create);
         * hosts driven through the reference's own entry points, which have no call for it, can override it:
         * SPH_VISC_STAB=gamma[,min_dt_sigma] ("0" = the plain gather for every preset) */
        const char *vs = getenv("SPH_VISC_STAB");
        if (vs) {
            float gamma = 0.0f, thr = 0.0f;
            if (sscanf(vs, "%f,%f", &gamma, &thr) >= 1) note("sph_set_viscosity_stabilisation", sph_set_viscosity_stabilisation(G.ctx, gamma, thr));
        }
    }
    sync_params("sph_ref_attach", params);
    sph_particle *flat = (sph_particle *)malloc((size_t)(n > 0 ? n : 1) * sizeof(sph_particle));
    for (int i = 0; i < n; i++) flat[i] = *pointers[i];
    uint32_t *uid = NULL;
    if (nranks > 1) {
        /* one numbering for all ranks: the pointer index the particle has in a one-rank run, row-major over the whole
         * lattice (geometry.c:44-62) -- the order inside a cell, hence every bit of the result, is then the one-rank
         * run's.  A host that did not come through partitionProblem gets disjoint blocks per rank. */
        uid = (uint32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(uint32_t));
        const int lattice = H.len_x > 0 && H.total_x >= H.len_x && n % H.len_x == 0;
        for (int i = 0; i < n; i++)
            uid[i] = lattice ? (uint32_t)((i / H.len_x) * H.total_x + H.x_start + i % H.len_x)
                             : (uint32_t)rank * (uint32_t)params->number_fluid_particles_global + (uint32_t)i;
    }
    rc = sph_upload(G.ctx, flat, uid, n);            /* one rank: uid = pointer index */
    free(uid);
    free(flat);
    if (rc == SPH_OK && nranks > 1) {
        /* a slab uploads its own particles only; if they are already moving (a restart) the first viscosity pass needs
         * the neighbours across the edges: one ghost exchange, all ranks together (harmless from rest, where the
         * reference starts: fluid.c:762-767) */
        if ((rc = sph_refresh_ghosts(G.ctx)) == SPH_OK && (rc = sph_exchange_via_host(G.ctx, 1, H.fn, H.user)) == SPH_OK)
            rc = sph_sort(G.ctx);
    }
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
    const int room = G.nranks > 1 ? G.capacity : G.n;
    sph_particle *flat = (sph_particle *)malloc((size_t)(room > 0 ? room : 1) * sizeof(sph_particle));
    int n = sph_download(G.ctx, flat, NULL, SPH_ORDER_UID, 0);
    if (n < 0) { free(flat); note("sph_ref_sync_to_host", -n); return -n; }
    if (G.nranks > 1) {
        /* the population changes with every migration (communication.c:370-431 rewires the pointer array there) */
        if (!H.base && n > G.n) { free(flat); note("sph_ref_sync_to_host", SPH_ERR_STATE); return SPH_ERR_STATE; }
        for (int i = 0; i < n && H.base; i++) pointers[i] = H.base + i;
        for (int i = n; i < G.n; i++) pointers[i] = NULL;
        G.n = n;
        params->max_fluid_particle_index = n - 1;
    }
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
    if (G.ctx) sync_params("apply_gravity", params);     /* (an unmodified driver attaches at predict_positions) */
    G.pending_gravity = 1;
}

void viscosity_impluses(fluid_particle **pointers, neighbor *neighbors, param *params)
{
    (void)pointers; (void)neighbors;
    if (G.ctx) sync_params("viscosity_impluses", params);
    G.pending_viscosity = 1;
}

/* An UNMODIFIED start_simulation never calls sph_ref_attach.  Its first call that has to run on the
 * device is predict_positions, and by then everything attach needs has been passed in: the pointer
 * array and counts (apply_gravity), the tank (this call), and the hash spacing, which the driver sets
 * to the smoothing radius (fluid.c:176).  Such a driver also reads the host AoS when it packs the
 * frame (fluid.c:358-362), so the mirror is switched on (SPH_REF_MIRROR_EVERY=N relaxes it to every
 * N-th step; the driver packs every 4th, fluid.c:105).  Without a usable GPU this ABORTS: the
 * reference's functions are void, and carrying on would hand the renderer a frozen fluid. */
static void lazy_attach(fluid_particle **pointers, AABB_t *boundary, param *params)
{
    neighbor_grid_t grid;
    memset(&grid, 0, sizeof grid);
    grid.spacing = params->tunable_params.smoothing_radius;
    const char *dev = getenv("SPH_B200_DEVICE"), *ndev = getenv("SPH_B200_DEVICES"), *every = getenv("SPH_REF_MIRROR_EVERY");
    if (!G.mirror) sph_ref_set_mirror(every && atoi(every) > 0 ? atoi(every) : 1);
    ask_host();
    /* several ranks on one box: SPH_B200_DEVICES=8 spreads them, rank r on device r mod 8 */
    const int device = dev ? atoi(dev) : (ndev && atoi(ndev) > 0 ? G.rank % atoi(ndev) : 0);
    const int rc = sph_ref_attach(pointers, params, boundary, &grid, device);
    if (rc != SPH_OK) {
        fprintf(stderr, "sph_ref_api: cannot put the simulation on the GPU (%s); there is no CPU path\n", G.err);
        abort();
    }
}

void predict_positions(fluid_particle **pointers, AABB_t *boundary_global, param *params)
{
    if (!G.ctx) lazy_attach(pointers, boundary_global, params);
    sync_params("predict_positions", params);
    if (!G.pending_gravity || !G.pending_viscosity) {
        snprintf(G.err, sizeof G.err, "predict_positions: the fused stage needs apply_gravity and "
                 "viscosity_impluses first (fluid.c:273-279 order)");
        fprintf(stderr, "sph_ref_api: %s\n", G.err);
        return;
    }
    G.pending_gravity = G.pending_viscosity = 0;
    if (G.nranks > 1 && (params->tunable_params.node_start_x != G.edge_start || params->tunable_params.node_end_x != G.edge_end)) {
        /* the balancer moved this slab (renderer.c:427-477): the step boundary is where the device can follow */
        G.edge_start = params->tunable_params.node_start_x;
        G.edge_end = params->tunable_params.node_end_x;
        G.pushed = params->tunable_params;
        note("predict_positions", sph_queue_params(G.ctx, &G.pushed));
    }
    note("predict_positions", sph_advect(G.ctx));
}

void identify_oob_particles(fluid_particle **pointers, fluid_particle *particles, oob_t *oob, AABB_t *b, param *params)
{
    (void)pointers; (void)oob; (void)b;
    H.base = particles;
    sync_params("identify_oob_particles", params);     /* the scatter has landed by now (fluid.c:293-310) */
    if (G.nranks > 1) note("identify_oob_particles", sph_exchange_via_host(G.ctx, 0, H.fn, H.user));
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
    G.relaxed = 1;
}

/* ------------------------------------------------------------------ hash.h */

void hash_fluid(fluid_particle **pointers, neighbor_grid_t *grid, param *params, bool compute_density)
{
    (void)grid;
    sync_params("hash_fluid", params);
    int rc = sph_sort(G.ctx);
    note("hash_fluid", rc);
    if (rc == SPH_OK && compute_density) note("hash_fluid", sph_density(G.ctx));
    if (!compute_density) G.relaxed = 0;
    if (rc == SPH_OK && !compute_density) {
        /* the re-hash of fluid.c:341 ends the step: positions and velocities are final */
        G.steps_done++;
        if (G.mirror && G.steps_done % G.mirror_every == 0) {
            note("hash_fluid", sph_ref_sync_to_host(pointers, params));
            {
                /* the reference's functions are void: what a slab could not hold or send is reported here (stderr and
                 * sph_ref_last_error), where the host is synchronised anyway.  Particles that exceeded the capacity are
                 * GONE: like the failed attach, that ends the program unless SPH_REF_STRICT=0 asks to carry on. */
                sph_status st;
                if (sph_get_status(G.ctx, &st) == SPH_OK && (st.msg_overflow || st.capacity_overflow || st.exchange_timeouts)) {
                    snprintf(G.err, sizeof G.err, "rank %d: %d entries did not fit a neighbour message, %d particles exceeded the slab's capacity, "
                             "%d neighbour messages never arrived", G.rank, st.msg_overflow, st.capacity_overflow, st.exchange_timeouts);
                    fprintf(stderr, "sph_ref_api: %s\n", G.err);
                    const char *strict = getenv("SPH_REF_STRICT");
                    if ((st.capacity_overflow || st.exchange_timeouts) && !(strict && strict[0] == '0')) {
                        fprintf(stderr, "sph_ref_api: particles were lost; stopping (SPH_REF_STRICT=0 to carry on)\n");
                        abort();
                    }
                }
            }
        }
    }
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

/* ------------------------------------------------------------------ communication.h */

void startHaloExchange(fluid_particle **pointers, fluid_particle *particles, edge_t *edges, param *params)
{
    (void)pointers; (void)edges; (void)params;
    H.base = particles;
    /* the call after updateVelocities (fluid.c:337) carries the relaxed ghost layer; the one before the relaxation
     * (fluid.c:318) has nothing left to move, its ghosts came with the migrants.  (A one-exchange build of the library
     * relaxes its ghosts itself and skips this meeting.) */
    if (G.ctx && G.nranks > 1 && G.relaxed && sph_exchanges_per_step() == 2) {
        G.relaxed = 0;
        note("startHaloExchange", sph_exchange_via_host(G.ctx, 1, H.fn, H.user));
    }
}
void finishHaloExchange(fluid_particle **pointers, fluid_particle *particles, edge_t *edges, param *params)
{ (void)pointers; (void)particles; (void)edges; params->number_halo_particles = 0; }
void transferOOBParticles(fluid_particle **pointers, fluid_particle *particles, oob_t *oob, param *params)
{ (void)pointers; (void)particles; (void)oob; (void)params; }

/* ------------------------------------------------------------------ geometry.h / fluid.h start-up (host) */

/* The leading members of edge_t / oob_t (communication.h:45-63).  edge_t ends in MPI_Request reqs[4],
 * whose size depends on the MPI ABI; nothing here goes that far. */
struct edge_head { int max_edge_particles; fluid_particle **left, **right; int n_left, n_right; };
struct oob_head { int max_oob_particles; int *left, *right; int n_left, n_right; int *vacant; int number_vacancies; };

void constructFluidVolume(fluid_particle **pointers, fluid_particle *particles, AABB_t *fluid, int start_x,
                          int number_particles_x, edge_t *edges, float spacing, param *params)
{
    const int num_y = (int)floor((fluid->max_y - fluid->min_y) / spacing);          /* geometry.c:35 */
    struct edge_head *e = (struct edge_head *)edges;
    e->n_left = 0; e->n_right = 0;                                                  /* :38-39 */
    int i = 0;
    for (int ny = 0; ny < num_y; ny++) {
        const float y = fluid->min_y + ny * spacing;                                /* :47 */
        for (int nx = 0; nx < number_particles_x; nx++) {
            fluid_particle *p = particles + i;
            p->x = fluid->min_x + (start_x + nx) * spacing;                         /* :49 */
            p->y = y;
            pointers[i] = p;
            p->id = i;
            i++;
        }
    }
    params->number_fluid_particles_local = i;                                       /* :65-66 */
    params->max_fluid_particle_index = i - 1;
}

void setParticleNumbers(AABB_t *boundary_global, AABB_t *fluid_global, edge_t *edges, oob_t *out_of_bounds,
                        int number_particles_x, float spacing, param *params)
{
    (void)boundary_global; (void)fluid_global; (void)number_particles_x; (void)spacing;
    ((struct edge_head *)edges)->max_edge_particles = params->number_fluid_particles_global;       /* geometry.c:82 */
    ((struct oob_head *)out_of_bounds)->max_oob_particles = params->number_fluid_particles_global; /* :87 */
    ((struct oob_head *)out_of_bounds)->number_vacancies = 0;                                      /* :96 */
}

void partitionProblem(AABB_t *boundary_global, AABB_t *fluid_global, int *x_start, int *length_x, float spacing, param *params)
{
    ask_host();
    const int nprocs = G.nranks, rank = G.rank;                                     /* geometry.c:105-108 */
    const int fluid_particles_x = (int)floor((fluid_global->max_x - fluid_global->min_x) / spacing) + 1;   /* :112 */
    const int equal = fluid_particles_x / nprocs, remaining = fluid_particles_x - equal * nprocs;          /* :118-125 */
    int number_to_left = 0, mine = 0, total_x = 0;
    for (int i = 0; i < nprocs; i++) {
        const int len = equal + (i < remaining ? 1 : 0);                            /* :128-129 */
        if (i < rank) number_to_left += len;
        if (i == rank) mine = len;
        total_x += len;
    }
    *x_start = number_to_left;                                                      /* :137-139 */
    *length_x = mine;
    H.x_start = number_to_left; H.len_x = mine; H.total_x = total_x;
    sph_tunable *t = &params->tunable_params;
    t->node_start_x = fluid_global->min_x + ((number_to_left - 1) * spacing);       /* :142-143 */
    t->node_end_x = t->node_start_x + (mine * spacing);
    if (rank == 0) t->node_start_x = boundary_global->min_x;                        /* :145-148 */
    if (rank == nprocs - 1) t->node_end_x = boundary_global->max_x;
    const int num_y = (int)floor((fluid_global->max_y - fluid_global->min_y) / spacing);   /* :153-157 */
    params->number_fluid_particles_global = total_x * num_y;
}

void initParticles(fluid_particle **pointers, fluid_particle *particles, AABB_t *water, int start_x,
                   int number_particles_x, edge_t *edges, int max_fluid_particles_local, float spacing, param *params)
{
    constructFluidVolume(pointers, particles, water, start_x, number_particles_x, edges, spacing, params);
    for (int i = params->number_fluid_particles_local; i < max_fluid_particles_local; i++) pointers[i] = NULL;   /* fluid.c:758-759 */
    for (int i = 0; i < params->number_fluid_particles_local; i++) {                /* :762-767 */
        pointers[i]->a_x = 0.0f; pointers[i]->a_y = 0.0f;
        pointers[i]->v_x = 0.0f; pointers[i]->v_y = 0.0f;
    }
}

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
