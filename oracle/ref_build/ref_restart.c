/*
 * ref_restart -- a host of the "edited anyway" kind on the reference-named entry points (include/sph_ref_api.h): explicit
 * sph_ref_set_rank / sph_ref_set_transport / sph_ref_attach, the reference's call order per step (fluid.c:273-348),
 * sph_ref_sync_to_host at the end, no host mirror -- starting from a MOVING fluid, which the reference's own driver never
 * does (fluid.c:762-767).  K forked ranks over the mini-MPI, one slab each.  TEST INFRASTRUCTURE ONLY; links the product
 * library, no reference code.
 *
 *   ref_restart --ranks K --steps N --out FILE     ->  FILE.r<rank>: int32 n, n x (x, y, v_x, v_y) float32, ascending uid,
 *                                                      then n x 2 int16: the slab's frame from sph_ref_pack_coords (the
 *                                                      device-side feed of INTEGRATION.md 2, device order)
 */
#include <math.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include "mpi.h"
#include "sph_host.h"
#include "sph_ref_api.h"

static int g_me, g_ranks;

static void xfer(const void *send, size_t ns, int to_side, void *recv, size_t nr, int from_side, void *user)
{
    (void)user;
    const int to = !send ? MPI_PROC_NULL : to_side == 0 ? g_me - 1 : g_me + 1;
    const int from = !recv ? MPI_PROC_NULL : from_side == 0 ? g_me - 1 : g_me + 1;
    MPI_Sendrecv((void *)send, (int)ns, MPI_CHAR, to, 41, recv, (int)nr, MPI_CHAR, from, 41, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
}

int main(int argc, char **argv)
{
    int ranks = 1, steps = 8;
    const char *out = "ref_restart.bin";
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "--ranks")) ranks = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--steps")) steps = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--out")) out = argv[i + 1];
    }
    if (mini_mpi_world_create(ranks, (size_t)16 << 20)) return 2;
    int me = -1;
    pid_t pids[256];
    for (int r = 0; r < ranks; r++) {
        pid_t pid = fork();
        if (pid < 0) return 2;
        if (pid == 0) { me = r; break; }
        pids[r] = pid;
    }
    if (me < 0) {
        int worst = 0, st;
        while (wait(&st) > 0)
            if ((!WIFEXITED(st) || WEXITSTATUS(st)) && !worst) {
                worst = WIFEXITED(st) ? WEXITSTATUS(st) : 5;
                for (int i = 0; i < ranks; i++) kill(pids[i], SIGKILL);
            }
        return worst;
    }
    prctl(PR_SET_PDEATHSIG, SIGKILL);
    alarm(240);
    mini_mpi_bind(me);
    g_me = me; g_ranks = ranks;

    /* the reference's default tank and lattice (fluid.c:117-160 with a 1920 x 1080 screen) */
    AABB_t boundary = { 0.0f, 15.0f, 0.0f, 8.4375f, 0.0f, 0.0f }, water = boundary;
    const float spacing = sph_host_spacing(15.0f, 8.4375f, 1500);
    param params;
    memset(&params, 0, sizeof params);
    sph_host_default_params(&params.tunable_params, 2.0f * spacing, boundary.max_x, boundary.max_y);
    sph_ref_set_transport(xfer, NULL);
    sph_ref_set_rank(me, ranks);
    int x_start = 0, len_x = 0;
    partitionProblem(&boundary, &water, &x_start, &len_x, spacing, &params);
    const int cap = 2 * params.number_fluid_particles_global;
    fluid_particle *particles = calloc((size_t)cap, sizeof *particles);
    fluid_particle **pointers = calloc((size_t)cap, sizeof *pointers);
    char edges[512] = {0}, oob[512] = {0};
    setParticleNumbers(&boundary, &water, (edge_t *)edges, (oob_t *)oob, len_x, spacing, &params);
    initParticles(pointers, particles, &water, x_start, len_x, (edge_t *)edges, cap, spacing, &params);
    /* ... already in motion: a shear plus a swirl, a function of the position only, so every decomposition starts alike */
    for (int i = 0; i < params.number_fluid_particles_local; i++) {
        fluid_particle *p = pointers[i];
        p->v_x = 1.5f * sinf(0.7f * p->x + 0.3f * p->y);
        p->v_y = 1.0f * cosf(0.5f * p->x) - 0.4f;
        p->x_prev = p->x; p->y_prev = p->y;
    }
    neighbor_grid_t grid;
    memset(&grid, 0, sizeof grid);
    grid.spacing = params.tunable_params.smoothing_radius;
    grid.size_x = (unsigned)ceilf(boundary.max_x / grid.spacing); grid.size_y = (unsigned)ceilf(boundary.max_y / grid.spacing);
    if (sph_ref_attach(pointers, &params, &boundary, &grid, 0) != SPH_OK) { fprintf(stderr, "ref_restart: %s\n", sph_ref_last_error()); return 3; }
    for (int s = 0; s < steps; s++) {
        apply_gravity(pointers, &params);
        viscosity_impluses(pointers, NULL, &params);
        predict_positions(pointers, &boundary, &params);
        identify_oob_particles(pointers, particles, (oob_t *)oob, &boundary, &params);
        hash_fluid(pointers, &grid, &params, true);
        startHaloExchange(pointers, particles, (edge_t *)edges, &params);
        finishHaloExchange(pointers, particles, (edge_t *)edges, &params);
        hash_halo(pointers, &grid, &params, true);
        double_density_relaxation(pointers, NULL, &params);
        updateVelocities(pointers, (edge_t *)edges, &boundary, &params);
        startHaloExchange(pointers, particles, (edge_t *)edges, &params);
        hash_fluid(pointers, &grid, &params, false);
        finishHaloExchange(pointers, particles, (edge_t *)edges, &params);
        hash_halo(pointers, &grid, &params, false);
    }
    if (sph_ref_sync_to_host(pointers, &params) != SPH_OK || sph_ref_last_error()[0]) { fprintf(stderr, "ref_restart: %s\n", sph_ref_last_error()); return 4; }
    char path[4096];
    snprintf(path, sizeof path, "%s.r%d", out, me);
    FILE *f = fopen(path, "wb");
    if (!f) return 2;
    const int n = params.number_fluid_particles_local;
    fwrite(&n, 4, 1, f);
    for (int i = 0; i < n; i++) { float v[4] = { pointers[i]->x, pointers[i]->y, pointers[i]->v_x, pointers[i]->v_y }; fwrite(v, 4, 4, f); }
    short *coords = calloc((size_t)2 * cap, sizeof *coords);
    if (sph_ref_pack_coords(coords, cap) != n) { fprintf(stderr, "ref_restart: sph_ref_pack_coords disagrees about the population\n"); return 5; }
    fwrite(coords, 4, (size_t)n, f);
    fclose(f);
    sph_ref_detach();
    return 0;
}
