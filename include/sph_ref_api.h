/*
 * sph_ref_api -- the reference's own entry points, served by the B200 library.
 *
 * These are the functions the TinySPH compute-rank driver (start_simulation, fluid.c:270-348)
 * calls, with the reference's exact names, argument lists and record layouts
 * (fluid.h:112-126, hash.h:50-52, communication.h:65-70), so that driver links against
 * libsph_b200.so instead of its own fluid.c/hash.c/communication.c bodies.
 *
 * Data model.  The reference passes host arrays into every call; here the state lives on the
 * device between sph_ref_attach() and sph_ref_detach() and the calls enqueue GPU work.  The host
 * AoS is only a mirror: it is read once at attach and written back by sph_ref_sync_to_host().
 * The reference has no such hooks (its callers own plain malloc'd arrays, fluid.c:181-233), so
 * these three are the only additions a maintainer makes to the driver.
 *
 * Error convention.  The reference's functions are void and report nothing (fluid.c:184-185,
 * hash.c:160-165); the same here, with an out-of-band query: sph_ref_last_error().
 *
 * If the reference's fluid.h was included first, its type definitions are used; otherwise the
 * identical layouts below are.
 */
#ifndef SPH_REF_API_H
#define SPH_REF_API_H

#include <stdbool.h>
#include "sph_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef fluid_fluid_h   /* the reference's own include guard (fluid.h:25) */
typedef struct sph_particle fluid_particle;          /* struct FLUID_PARTICLE, fluid.h:56-70 */
typedef struct sph_tunable tunable_parameters;       /* struct TUNABLE_PARAMETERS, fluid.h:78-97 */
typedef struct sph_param param;                      /* struct PARAM, fluid.h:100-106 */
typedef struct NEIGHBOR neighbor;                    /* fluid.h:72-75 -- never dereferenced here */
typedef struct BUCKET_T bucket_t;                    /* hash.h:35-38 -- never dereferenced here */
typedef struct EDGE_T edge_t;                        /* communication.h:45-52 -- opaque */
typedef struct OOB_T oob_t;                          /* communication.h:55-63 -- opaque */
typedef struct NEIGHBOR_GRID_T {                     /* hash.h:40-48 */
    float spacing;
    unsigned int size_x, size_y;
    neighbor *neighbors;
    bucket_t *grid_buckets;
    unsigned int max_neighbors, max_bucket_size;
} neighbor_grid_t;
typedef struct AABB_T {                              /* geometry.h:36-43 */
    float min_x, max_x, min_y, max_y, min_z, max_z;
} AABB_t;
#endif

/* ---- the three hooks the reference lacks ---- */
/* Upload the particles the pointer array refers to (pointer order becomes the uid) and create the
 * device context for this tank / grid.  Returns SPH_OK or an SPH_ERR_* code. */
int sph_ref_attach(fluid_particle **fluid_particle_pointers, param *params, AABB_t *boundary_global,
                   neighbor_grid_t *grid, int device);
/* Materialise the device state into the host AoS (pointer order), refresh params' counts. */
int sph_ref_sync_to_host(fluid_particle **fluid_particle_pointers, param *params);
void sph_ref_detach(void);
/* fluid.c:354-365: the int16 coordinate feed, straight from the device. Returns the count. */
int sph_ref_pack_coords(short *fluid_particle_coords, int max_pairs);
const char *sph_ref_last_error(void);
sph_ctx *sph_ref_context(void);
/* An UNMODIFIED start_simulation (no attach call at all) also works: the first predict_positions attaches
 * by itself from what the calls carry (device from SPH_B200_DEVICE, default 0) and switches the host
 * mirror on, because such a driver packs its frame from the host AoS (fluid.c:358-362).  If the GPU
 * cannot be used that path aborts -- there is no CPU fallback.
 * sph_ref_set_mirror(n): write positions/velocities back into the host AoS after every n-th step
 * (0 = never, the default after an explicit attach; SPH_REF_MIRROR_EVERY overrides the lazy default 1). */
void sph_ref_set_mirror(int every_n_steps);
/* what MPI_Comm_rank / MPI_Comm_size(MPI_COMM_COMPUTE) say on this rank (partitionProblem asks them,
 * geometry.c:105-108; the library itself does not link MPI).  Default 0 of 1. */
void sph_ref_set_rank(int rank, int nranks);
/* Several compute ranks, one slab each: how this rank's neighbour messages travel (sph_exchange_via_host's callback;
 * side 0 = rank - 1, 1 = rank + 1).  Without a transport an attach with nranks > 1 is refused.  An unmodified driver
 * cannot make these two calls: it links sph_b200/host/glue/sph_ref_mpi_glue.c instead, whose sph_ref_host_mpi() the
 * library finds by itself and which answers both questions from MPI_COMM_COMPUTE. */
void sph_ref_set_transport(sph_sendrecv_fn fn, void *user);
/* NOT defined by the library: the host's glue object defines it (sph_b200/host/glue/sph_ref_mpi_glue.c); the library
 * refers to it weakly and calls it once, before it needs to know its slab.  Returns 0 and fills all four. */
int sph_ref_host_mpi(int *rank, int *nranks, sph_sendrecv_fn *fn, void **user);

/* ---- fluid.h:112-126 ---- */
void apply_gravity(fluid_particle **fluid_particle_pointers, param *params);
void viscosity_impluses(fluid_particle **fluid_particle_pointers, neighbor *neighbors, param *params);
void predict_positions(fluid_particle **fluid_particle_pointers, AABB_t *boundary_global, param *params);
void double_density_relaxation(fluid_particle **fluid_particle_pointers, neighbor *neighbors, param *params);
void updateVelocities(fluid_particle **fluid_particle_pointers, edge_t *edges, AABB_t *boundary_global, param *params);
void identify_oob_particles(fluid_particle **fluid_particle_pointers, fluid_particle *fluid_particles,
                            oob_t *out_of_bounds, AABB_t *boundary_global, param *params);
/* per-particle helpers, evaluated on the host with the reference's arithmetic */
void boundaryConditions(fluid_particle *p, AABB_t *boundary, param *params);
void calculate_density(fluid_particle *p, fluid_particle *q, float ratio);
void updateVelocity(fluid_particle *p, param *params);
void checkVelocity(float *v_x, float *v_y);

/* ---- start-up: geometry.h:45-49, fluid.h:121 (host arithmetic identical to the reference's; silent) ---- */
void constructFluidVolume(fluid_particle **fluid_particle_pointers, fluid_particle *fluid_particles, AABB_t *fluid,
                          int start_x, int number_particles_x, edge_t *edges, float spacing, param *params);
void setParticleNumbers(AABB_t *boundary_global, AABB_t *fluid_global, edge_t *edges, oob_t *out_of_bounds,
                        int number_particles_x, float spacing, param *params);
void partitionProblem(AABB_t *boundary_global, AABB_t *fluid_global, int *x_start, int *length_x, float spacing, param *params);
void initParticles(fluid_particle **fluid_particle_pointers, fluid_particle *fluid_particles, AABB_t *water, int start_x,
                   int number_particles_x, edge_t *edges, int max_fluid_particles_local, float spacing, param *params);

/* ---- hash.h:50-52 ---- */
unsigned int hash_val(float x, float y, neighbor_grid_t *grid, param *params);
void hash_fluid(fluid_particle **fluid_particle_pointers, neighbor_grid_t *grid, param *params, bool compute_density);
void hash_halo(fluid_particle **fluid_particle_pointers, neighbor_grid_t *grid, param *params, bool compute_density);

/* ---- communication.h:65-70 (one rank: nothing to exchange; several ranks: identify_oob_particles moves migrants +
 *      ghost layer, the startHaloExchange after updateVelocities moves the relaxed ghost layer, through the transport) ---- */
void startHaloExchange(fluid_particle **fluid_particle_pointers, fluid_particle *fluid_particles, edge_t *edges, param *params);
void finishHaloExchange(fluid_particle **fluid_particle_pointers, fluid_particle *fluid_particles, edge_t *edges, param *params);
void transferOOBParticles(fluid_particle **fluid_particle_pointers, fluid_particle *fluid_particles, oob_t *out_of_bounds, param *params);

#ifdef __cplusplus
}
#endif
#endif
