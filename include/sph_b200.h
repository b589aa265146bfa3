/*
 * sph_b200 -- C ABI of the B200-native TinySPH compute-rank timestep.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  The library behind
 * it (sph_b200/csrc -> libsph_b200.so) is hand-written CUDA for sm_100a and has
 * NO CPU fallback: every entry point returns SPH_ERR_CUDA when no device is
 * usable.
 *
 * Each entry point cites the reference interface it replaces
 * (paths under AdamSimpson/SPH `src/`).  The reference-named wrappers
 * (apply_gravity, hash_fluid, ... with the reference's own signatures) are in
 * include/sph_ref_api.h and are implemented on top of this header.
 *
 * Data model: the reference passes a host AoS (`fluid_particle`, 52 bytes) plus
 * a pointer-array view into every call (fluid.h:112-126).  Here the state is
 * resident on the device as cell-sorted SoA float2 arrays; the host AoS is a
 * mirror that is filled only on request (sph_download / sph_pack_coords).
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- records, bit-for-bit the reference's (fluid.h:56-106) ---- */

/* == struct FLUID_PARTICLE (fluid.h:56-70): 52 bytes, id @48 */
typedef struct sph_particle {
    float x_prev, y_prev;
    float x, y;
    float v_x, v_y;
    float a_x, a_y;               /* dead fields in the reference (fluid.c:763-764) */
    float density, density_near;
    float pressure, pressure_near;
    int id;                       /* local index in the pointer array */
} sph_particle;

/* == struct TUNABLE_PARAMETERS (fluid.h:78-97): 64 bytes */
typedef struct sph_tunable {
    float rest_density;
    float smoothing_radius;
    float g;
    float k;
    float k_near;
    float k_spring;
    float sigma;
    float beta;
    float time_step;
    float node_start_x;
    float node_end_x;
    float mover_center_x;
    float mover_center_y;
    float mover_width;
    float mover_height;
    char mover_type;              /* SPH_SPHERE_MOVER / SPH_RECTANGLE_MOVER (fluid.h:48-49) */
    char kill_sim;
    char active;
} sph_tunable;

/* == struct PARAM (fluid.h:100-106): 80 bytes, counts @64.. */
typedef struct sph_param {
    sph_tunable tunable_params;
    int number_fluid_particles_global;
    int number_fluid_particles_local;
    int max_fluid_particle_index;
    int number_halo_particles;
} sph_param;

#define SPH_SPHERE_MOVER 0
#define SPH_RECTANGLE_MOVER 1

/* Sort-grid refinement.  Particles are binned into the reference's hash cells (side h, hash.c:35-47)
 * subdivided SPH_CELL_DIV times per axis, and the neighbour search visits (2*DIV+1)^2 sub-cells: 5x5
 * half-size cells cover 6.25 h^2 instead of the 9 h^2 of 3x3 full cells, i.e. 30% fewer distance tests.
 * The reference cell id is exactly recoverable: floor(DIV * (x/h)) / DIV == floor(x/h) for DIV a power of
 * two, because scaling an fp32 quotient by a power of two is exact.  The oracle's gather form uses the
 * same constant so that both sum in the same order. */
#define SPH_CELL_DIV 2

/* reference capacities (fluid.c:174-175): detected, see sph_status */
#define SPH_REF_MAX_BUCKET 100
#define SPH_REF_MAX_NEIGHBORS 400

/* ---- errors ---- */
enum {
    SPH_OK = 0,
    SPH_ERR_CUDA = 1,        /* CUDA runtime error or no device: never falls back to the CPU */
    SPH_ERR_ARG = 2,
    SPH_ERR_CAPACITY = 3,    /* particle / message capacity exceeded */
    SPH_ERR_STATE = 4
};

typedef struct sph_ctx sph_ctx;

typedef struct sph_config {
    float tank_w, tank_h;    /* boundary_global.max_x / max_y; min is 0 (fluid.c:117-127) */
    float h;                 /* smoothing radius == hash grid spacing (fluid.c:159,176) */
    int capacity;            /* max resident particles on this device (local + halo) */
    int msg_capacity;        /* max particles in one neighbour message (halo or migrants) */
    int device;              /* CUDA device ordinal */
    int rank, nranks;        /* slab index / number of slabs; nranks == 1: no exchange */
    float halo_width;        /* ghost-layer width in units of h; 0 -> default 2.0 (3.5 with one exchange per step).  With two
                              * exchanges per step, the stabilised viscosity gather (goo) and a mover that moves, ask for 3:
                              * its coupling sums are one more pair pass, which leaves the 2 h layer no margin (below) */
    void *stream;            /* cudaStream_t to run on, or NULL for a private stream */
    int exchanges_per_step;  /* slabs: 2 = neighbours meet after the prediction and after the relaxation (ghost layer 2 h);
                              * 1 = once, the ghosts are relaxed redundantly (layer >= 3.5 h; sph_set_exchange_period makes it
                              * once every E steps); 0 = the build's default (2) */
} sph_config;

/* what the reference silently drops (hash.c:160-165, :188-197, :223-232) is counted here */
typedef struct sph_status {
    int n_local, n_halo;
    int max_bucket;              /* largest population of a reference bucket (cell of side h) in the current state */
    int bucket_overflow;         /* buckets above SPH_REF_MAX_BUCKET now, or sub-cells above it in any earlier sort
                                  * (the reference would have dropped particles, hash.c:160-165) */
    int neighbor_overflow;       /* particles whose reference forward list (owner rule of hash.c:178-224) would exceed
                                  * SPH_REF_MAX_NEIGHBORS in the current state; exact.  (Between a producing kernel and its
                                  * sort: the density kernels' conservative running count instead.) */
    int capacity_overflow;       /* particles dropped because `capacity` was exceeded (fatal) */
    int msg_overflow;            /* entries that did not fit a neighbour message: a ghost is missing on the other side, or
                                  * an emigrant had to wait for the next exchange (it stays resident: nobody is lost) --
                                  * either way particles were stepped without their true neighbours: size msg_capacity up */
    int migrated_left, migrated_right;   /* last step */
    int exchange_timeouts;       /* neighbour messages that never arrived (device-side waits that gave up): the run is invalid */
    long long steps;
} sph_status;

/* ---- lifecycle (the reference allocates everything in start_simulation, fluid.c:178-233) ---- */
int sph_create(const sph_config *cfg, sph_ctx **out);
void sph_destroy(sph_ctx *ctx);
const char *sph_last_error(const sph_ctx *ctx);
int sph_synchronize(sph_ctx *ctx);
int sph_get_status(sph_ctx *ctx, sph_status *out);

/* Stream-ordered copy of the local particle count (one int) to DEVICE memory, without synchronising:
 * what a multi-rank driver all-gathers once per frame for the edge balancer (renderer.c:280,290). */
int sph_copy_n_local(sph_ctx *ctx, void *device_dst);
/* Same, two ints: {local particle count, work estimate of the last completed step}.  The estimate is the
 * sum over resident entries (ghosts included: their density is computed too) of 14 + neighbours, formed
 * by the density kernel; it is the input of the OPTIONAL cost-based edge policy (sph_host_balance_ex fed
 * with costs instead of counts).  Results do not depend on where the edges are, so the
 * policy changes the schedule only. */
int sph_copy_load(sph_ctx *ctx, void *device_dst);
/* Same, four ints: {local particle count, work estimate, this slab's OWN device time since the last call [us], the time
 * it spent waiting for its neighbours over the same span [us]} (peer-memory transport; 0, 0 otherwise).  The own time
 * runs from the end of one wait to the end of the next send -- everything the slab did between two meetings, however
 * long it then waited -- and feeds the OPTIONAL time-based edge policy: measured cost instead of a model of it. */
int sph_copy_work(sph_ctx *ctx, void *device_dst);

/* ---- parameters ---- */
/* Full tunable block, as the render rank scatters it (fluid.c:293-294). Stream-ordered.
 * On a slab (nranks > 1) in the MIDDLE of a step (between sph_advect and the step's last sph_sort) only the physics
 * takes effect at once; the edges are queued and land with the next sph_advect, like sph_queue_params -- the window
 * of grid columns the running step sorts into cannot move under it. */
int sph_set_params(sph_ctx *ctx, const sph_tunable *t);
/* Parameters that take effect between position prediction and migration of the NEXT
 * advect stage, which is where the reference's MPI_Scatterv lands (fluid.c:279-310). */
int sph_queue_params(sph_ctx *ctx, const sph_tunable *t);
/* Slab edges only (node_start_x / node_end_x); used by migration and halo selection.  On a slab only at a step
 * boundary (SPH_ERR_STATE otherwise; see sph_set_params).  Where the edges are never changes a result, with one
 * condition on how far they MOVE at once: a slab must keep a ghost layer (halo_width) of the extent it had before the
 * move, so that the strip its neighbour needs as ghosts in that step is already its own (DESIGN.md 6;
 * sph_host_balance_time enforces it, the reference's h/8 per frame cannot violate it).  Applies to the edges carried
 * by sph_set_params / sph_queue_params as well. */
int sph_set_edges(sph_ctx *ctx, float node_start_x, float node_end_x);

/* Stabilised viscosity gather (not in the reference).  DEFAULT: gamma = 0.5, min_dt_sigma = 0.5, i.e. it
 * engages by itself for the reference's goo preset (dt*sigma = 0.83) and for none of the others (<= 0.17), whose
 * results it would not change by a bit anyway; gamma = 0 switches it off (the plain gather for every block).
 * The reference applies viscosity_impluses pair by pair in place (fluid.c:442-472), which never overshoots.
 * A gather sums a particle's impulses from frozen velocities, and once C_i = sum_j dt (1-q)(sigma + beta u)
 * over its approaching pairs exceeds ~2 (the "goo" preset, controls.c:359-371) it overshoots and never
 * settles.  With gamma > 0 every pair's impulse is scaled by s_ij = 1 / max(1, gamma * max(C_i, C_j)):
 * symmetric (momentum still exchanged pairwise), independent of the slab decomposition, and exactly 1
 * wherever the plain gather is stable.  It costs one more gather pass (C), so it only runs for parameter
 * blocks with dt * sigma >= min_dt_sigma (0 = whenever gamma > 0).  gamma = 0.5 reproduced the
 * reference's long-run statistics for the goo preset within the reference's own order sensitivity.
 * This is the library's one algorithmic deviation from the reference beyond gather-vs-sweep order. */
int sph_set_viscosity_stabilisation(sph_ctx *ctx, float gamma, float min_dt_sigma);

/* ---- state ---- */
/* Snapshot of the whole resident state at a step boundary -- particles, cell tables, counters AND the parameter block
 * in force (a block set after the snapshot is undone by the restore) -- in DEVICE memory (one slot), and its restoration: the
 * same steps can be run again (a benchmark timing identical work repeatedly; a host that rewinds).  On slabs all
 * ranks save and restore together; the first step after a restore is an exchange step. */
int sph_state_save(sph_ctx *ctx);
int sph_state_restore(sph_ctx *ctx);
/* Host AoS -> device SoA, then bins by cell so the first viscosity pass has its
 * neighbour structure (the reference starts with empty lists, fluid.c:202; velocities are
 * zero there so both give no impulse).  uid may be NULL (uid = index). */
int sph_upload(sph_ctx *ctx, const sph_particle *aos, const uint32_t *uid, int n);
/* constructFluidVolume + initParticles (geometry.c:29-59, fluid.c:747-768) on the device: the lattice of
 * columns [start_col, start_col+ncols) of the water block, uid = row * total_cols + column, velocities
 * zero; then binned like an upload.  Large problems never need a host AoS (the reference's own
 * allocation overflows 32 bits above ~5.3 M particles, fluid.c:203).  Returns the count or <0. */
int sph_init_lattice(sph_ctx *ctx, float water_min_x, float water_min_y, float water_max_y, float spacing,
                     int start_col, int ncols, int total_cols);
#define SPH_ORDER_UID 0      /* ascending uid == the reference's pointer order on one rank */
#define SPH_ORDER_CELL 1     /* device order: row-major cell, then uid == bucket order */
/* Device SoA -> host AoS (local particles; halo too if include_halo). Returns count or <0. */
int sph_download(sph_ctx *ctx, sph_particle *aos, uint32_t *uid, int order, int include_halo);

/* ---- stages: each is the gather form of the named reference functions ---- */
/* apply_gravity (fluid.c:398) + viscosity_impluses (:416) + predict_positions (:507) incl.
 * boundaryConditions (:656) + identify_oob_particles (:481): one fused kernel. */
int sph_advect(sph_ctx *ctx);
/* hash_fluid pass 1 (hash.c:148-166) + hash_halo insertion (hash.c:51): counting sort by
 * hash_val (hash.c:35); completes the binning started inside advect/relax. */
int sph_sort(sph_ctx *ctx);
/* calculate_density over hash_fluid/hash_halo pairs (fluid.c:527, hash.c:190-228, :106-110) */
int sph_density(sph_ctx *ctx);
/* double_density_relaxation (fluid.c:541) + updateVelocities (:642) incl. boundaryConditions */
int sph_relax(sph_ctx *ctx);
/* n iterations of the loop at fluid.c:270-348 on one slab with no neighbours (nranks == 1);
 * replayed from a CUDA graph. */
int sph_step(sph_ctx *ctx, int n);

/* ---- slab exchange (communication.c:120-450): device-side message buffers ---- */
/* Message layout: 16-byte header {int n_migrants, n_halo, 0, 0} then records.  The
 * pack is fused into advect/relax, the unpack into sort; a transport only moves bytes.
 * which = 0: after advect (migrants + predicted-position halo), 1: after relax (pos+vel halo).
 * Results equal the one-slab run's bit for bit as long as no particle is displaced further than (halo_width - 1) h past
 * its slab's edge within one step ((halo_width - 2) h while the stabilised viscosity gather is engaged): ordinary
 * motion is clamped to 0.07 h per step; only the mover's push-out can do that -- a mover teleported into the fluid
 * beside an edge, or, with the stabilised gather on the default 2 h layer, any moving mover on an edge
 * ("The condition"; a valid but decomposition-dependent step follows). */
int sph_exchange_buffers(sph_ctx *ctx, int which, void **send_left, void **recv_left,
                         void **send_right, void **recv_right, size_t *bytes);
/* The same exchange for a host whose transport moves HOST memory (plain MPI_Sendrecv, sockets): the library stages
 * the message buffers through pinned memory around two calls of `fn`, ordered like the reference's own pair of
 * MPI_Sendrecv (communication.c:158-161, :340-344): first (send to the right, receive from the left), then (send to the left,
 * receive from the right); side 0 = left neighbour (rank - 1), 1 = right (rank + 1).  An absent neighbour appears
 * as (NULL, 0), the reference's MPI_PROC_NULL.  Call it where the buffers of sph_exchange_buffers(which) would be
 * moved.  Blocks until the outgoing messages are in host memory; the incoming ones are copied stream-ordered. */
typedef void (*sph_sendrecv_fn)(const void *send, size_t send_bytes, int to_side,
                                void *recv, size_t recv_bytes, int from_side, void *user);
int sph_exchange_via_host(sph_ctx *ctx, int which, sph_sendrecv_fn fn, void *user);

/* Restart from a moving snapshot on several slabs.  sph_upload places a slab's own particles only, so the viscosity
 * pass of the very first step would miss the neighbours across the edges (the reference, and sph_init_lattice,
 * start at rest, where that pass does nothing).  After every rank has uploaded: sph_refresh_ghosts (packs the
 * ghost layer, position + velocity, like the end of a step) -> move the which = 1 buffers of
 * sph_exchange_buffers (nothing to do with sph_p2p_connect) -> sph_sort.  All ranks must do this together: it
 * consumes one message sequence number.  No-op for nranks == 1. */
int sph_refresh_ghosts(sph_ctx *ctx);
/* Exchanges per step of this build: 2 (which = 0 after sph_advect, which = 1 after sph_relax), or 1 for the
 * one-exchange build variant (-DSPH_ONE_EXCHANGE=1), whose ghosts are relaxed redundantly and which needs
 * halo_width >= 3 (4 with the stabilised viscosity gather); the driver then skips the which = 1 transfer. */
int sph_exchanges_per_step(void);                       /* the build's default */
int sph_ctx_exchanges_per_step(const sph_ctx *ctx);    /* this context's (sph_config.exchanges_per_step) */
/* Exchange period (one-exchange build): neighbours meet every `period` steps instead of every step; in between a
 * slab advances its ghosts itself, redundantly and bit for bit as their owner does.  Every pair pass invalidates
 * one h of the layer from the outside, so a period of E steps needs halo_width >= 3.5 * E (4.5 * E while the
 * stabilised viscosity gather is engaged); with a narrower layer the library meets as often as the layer allows.
 * A queued parameter block (sph_queue_params) always makes its own step an exchange step -- it lands exactly where
 * the exchange is, fluid.c:293-310 -- and so does anything set outside the queue.  Migration happens at exchange
 * steps only; between them a particle that crossed an edge stays with its owner.  The result is still the
 * single-slab result bit for bit.  Replaces 2 meetings per step of the reference (fluid.c:310-348) by 1 / E. */
int sph_set_exchange_period(sph_ctx *ctx, int period);
/* Peer-memory transport: microseconds the exchange kernel spent {sending, waiting for the neighbours, unpacking},
 * summed over *meetings exchanges since the last reset (one block's view).  Synchronises the stream. */
int sph_get_exchange_times(sph_ctx *ctx, double us[3], int *meetings, int reset);
/* 1 if the step in progress (between sph_advect and the end of the step) is an exchange step, or, at a step
 * boundary, if the coming step will be one: a host-side transport moves the which = 0 buffers only then. */
int sph_exchange_due(sph_ctx *ctx);
/* Mark a neighbour as absent for the coming sort (edge slabs): its recv buffer is ignored. */
int sph_set_neighbors(sph_ctx *ctx, int has_left, int has_right);

/* Peer-memory transport (one process per GPU on one NVLink/NVSwitch box).  Each rank publishes the
 * 64-byte cudaIpc handle of its exchange block; after sph_p2p_connect the pack code of advect/relax
 * stores outgoing records straight into the neighbour's block and releases an arrival flag, and the
 * sort waits for the neighbours' flags on the device.  No transport calls, no host synchronisation:
 * sph_step() then works for slabs too (one CUDA graph per step).  Pass NULL for an absent side. */
int sph_p2p_local_handle(sph_ctx *ctx, void *handle64);
int sph_p2p_connect(sph_ctx *ctx, const void *left_handle64, const void *right_handle64);

/* ---- parity / inspection ---- */
/* per local particle: uid and hash_val cell id in the reference's GLOBAL grid numbering */
int sph_get_cells(sph_ctx *ctx, uint32_t *uid, uint32_t *cell, int cap);
/* every pair (uid_a < uid_b) with unfused r2 <= h2 among resident particles: the symmetric
 * closure of the reference's forward lists (hash.c:185,221,99). Returns pair count or <0. */
long long sph_get_pairs(sph_ctx *ctx, uint64_t *pairs, long long cap);
/* forward-neighbour count per local particle under the reference's ownership rule */
int sph_get_forward_counts(sph_ctx *ctx, uint32_t *uid, int *count, int cap);

/* ---- render feed (fluid.c:354-365): int16 pixel-range coordinates of local particles ---- */
int sph_pack_coords(sph_ctx *ctx, int16_t *xy_pairs, int cap);

/* ---- one render frame of the compute rank (the loop body at fluid.c:270-372, `steps` times) ----
 * `t` (may be NULL) is the block the render rank scatters; it lands in the LAST sub-step between
 * position prediction and migration (fluid.c:293-294).  Afterwards the int16 coordinate feed
 * (fluid.c:354-365) is copied into the HOST buffer xy_pairs (may be NULL).  nranks == 1 only.
 * Returns the local particle count or a negative error. */
int sph_run_frame(sph_ctx *ctx, const sph_tunable *t, int steps, int16_t *xy_pairs, int cap);

/* ---- the same feed without stalling the compute rank ----
 * The reference MPI_Isends its frame and starts the next one; it waits for that send only before it overwrites
 * the buffer (fluid.c:283-287, :354-365).  sph_pack_coords_async packs on the compute stream, lets a second
 * stream copy the frame into xy_pairs (pinned host memory; untouched until collected) and returns a ticket
 * (0 or 1, alternating; at most two frames in flight) or a negative error.  sph_coords_wait(ticket) blocks until
 * that frame has arrived and returns its particle count, like sph_pack_coords.  A single slab copies exactly its
 * particles; a slab among others, whose count is not known on the host without a synchronisation, copies the
 * population of the last frame it collected plus an eighth (the whole buffer until it has collected one), and
 * sph_coords_wait fetches the remainder in the rare frame in which the slab grew by more.  Each ticket has its own
 * device-side frame: the compute stream never waits for a copy.  sph_coords_copied(ticket): entries of that frame
 * that crossed to the host (for D2H accounting).  sph_run_frame_async = sph_run_frame ending in
 * sph_pack_coords_async. */
int sph_pack_coords_async(sph_ctx *ctx, int16_t *xy_pairs, int cap);
int sph_coords_wait(sph_ctx *ctx, int ticket);
int sph_coords_copied(sph_ctx *ctx, int ticket);
int sph_run_frame_async(sph_ctx *ctx, const sph_tunable *t, int steps, int16_t *xy_pairs, int cap);

/* kernels launched since the context was created (bench bookkeeping) */
long long sph_launch_count(const sph_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
