/*
 * sph_host -- C host layer of sph_b200: the start-up geometry, parameter model and slab
 * load balancer of the TinySPH compute path, in plain C99 (no CUDA types).  Linked into
 * libsph_b200.so next to the CUDA library; it prepares inputs for, and steers, the entry points
 * of sph_b200.h.  Citations: AdamSimpson/SPH `src/`.
 */
#ifndef SPH_HOST_H
#define SPH_HOST_H

#include "sph_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* initial particle spacing: sqrt(water area / N) as fluid.c:141-144 computes it */
float sph_host_spacing(float water_w, float water_h, int n_request);

/* hard-coded start-up parameters of start_simulation (fluid.c:88-107, :159), mover parked at
 * (0.5 W, 0.35 H) with diameter 2/15 W (the reference leaves the centre uninitialised) */
void sph_host_default_params(sph_tunable *t, float h, float tank_w, float tank_h);

/* fluid presets of the game-pad buttons: 'x' water, 'y' goo, 'a' zero-g, 'b' spring gas
 * (controls.c:344-401). Returns 0 on success. */
int sph_host_preset(sph_tunable *t, char which);

/* partitionProblem (geometry.c:101-160) for all ranks at once: equal lattice columns, remainder
 * to the left ranks, first/last slab clamped to the tank. Returns N_global actually used
 * (geometry.c:152-156). */
int sph_host_partition(float tank_w, float water_min_x, float water_max_x, float water_min_y,
                       float water_max_y, float spacing, int nranks,
                       int *start_col, int *ncols, float *start_x, float *end_x);

/* constructFluidVolume (geometry.c:29-59) + initParticles (fluid.c:747-768) for one slab's
 * columns; uid = row * total_cols + column. Returns the number of particles written. */
int sph_host_lattice(float water_min_x, float water_min_y, float water_max_y, float spacing,
                     int start_col, int ncols, int total_cols, sph_particle *out, uint32_t *uid);

/* check_partition_left (renderer.c:427-477): nudge interior slab edges by h/8 toward equal
 * particle counts; dead band even/15, minimum slab width 2h. `counts` are whatever the caller
 * uses consistently on both sides of the ratio (the reference passes coordinate counts,
 * renderer.c:280,290). */
void sph_host_balance(sph_tunable *master, int nactive, const int *counts, int total);
/* The same edge arithmetic with the dead band as a parameter (even / band_divisor; the reference's is 15):
 * for callers that balance on a work estimate (sph_copy_load) and want it tighter than +-6.7 %. */
void sph_host_balance_ex(sph_tunable *master, int nactive, const int *counts, int total, float band_divisor);

/* OPTIONAL edge policy on MEASURED slab times (sph_copy_work; not the reference's): every interior edge moves towards
 * the slower of its two slabs by `gain` x the shift that would equalise their times, at most max_shift_h smoothing
 * radii per call, keeping every slab at least min_width_h radii wide AND at least min_width_h radii of its extent
 * before the call (a slab whose two edges move the same way must still own, in the step the edges land, the ghosts its
 * neighbour needs: DESIGN.md 6); dead band 0.5 %.  The particles do not depend on
 * where the edges are, so this changes the schedule only. */
void sph_host_balance_time(sph_tunable *master, int nactive, const int *busy, float gain, float max_shift_h, float min_width_h);

/* The render rank's idle "autopilot" for the mover (renderer.c:513-531): per frame gl_x += 0.01 * dir,
 * direction flips outside [-1, 1], gl_y = sinf(3.14 * 5 * gl_x) / 10 - 0.6, then opengl_to_sim
 * (renderer.c:396-404).  Updates t->mover_center_{x,y}; *gl_x / *direction carry the state.
 * The reference does not keep gl_x: every frame it forms it again from the mover's centre (sim_to_opengl,
 * renderer.c:494: x / (tank_w / 2) - 1).  A caller that does the same before each call gets the compiled reference's
 * trajectory BIT FOR BIT (tests/test_host.py, 1200 frames against update_inactive_state of the unmodified renderer.c);
 * with *gl_x carried over, the rounding of that round trip is absent and a reversal at gl_x = +-1 can fall one frame
 * later: the same path, shifted by two frames from there on. */
void sph_host_mover_autopilot(sph_tunable *t, float tank_w, float tank_h, float *gl_x, int *direction);
/* The same path with the per-frame step as a parameter.  The reference's 0.01 GL units are 0.075 simulation units
 * per frame in ITS tank (width 15, fluid.c:119), 2.25 units/s against the +-5 velocity clamp (fluid.c:613-625).  In a
 * tank scaled with the particle count the same GL step is a teleport (4 M particles: 164 units/s), and whatever the
 * mover meets is piled onto its surface beyond the reference's bucket capacity (hash.c:160-165).  Scaled problems
 * (BASELINE config 4) therefore pass dx_gl = 0.01 * 15 / tank_w: the reference's step in simulation units. */
void sph_host_mover_autopilot_ex(sph_tunable *t, float tank_w, float tank_h, float *gl_x, int *direction, float dx_gl);

/* remove_partition / add_partition (controls.c:405-455): park the last active slab outside the
 * tank / split the last active slab in half. Return the new number of active slabs. */
int sph_host_remove_partition(sph_tunable *master, int nactive);
int sph_host_add_partition(sph_tunable *master, int nactive, int nranks);

#ifdef __cplusplus
}
#endif
#endif
