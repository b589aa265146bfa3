/*
 * ref_world -- the reference's whole program, UNMODIFIED: its own main() (fluid.c:46-68), render rank
 * (renderer.c + controls.c: parameter scatter, coordinate gather, load balancer) and K compute ranks
 * (start_simulation), as K + 1 forked processes over the mini-MPI -- "mpirun -n K+1" without an MPI installation and
 * without a display (render_stubs.c).  TEST INFRASTRUCTURE ONLY.
 *
 *   oracle/_ref/sph_ref_world_cpu   the pure reference
 *   oracle/_ref/sph_ref_world_gpu   the same objects with the MPI glue and libsph_b200.so in front: the compute ranks'
 *                                   hot path runs in the product library, one slab per rank, and
 *                                   the reference's own renderer balances them and consumes their frames
 *
 *   sph_ref_world_* --ranks K --frames F --out FILE      (record format: render_stubs.c)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <signal.h>
#include <string.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include "mpi.h"

int ref_main(int argc, char **argv);      /* fluid.c's main(), renamed by -Dmain=ref_main */

static void report_binding(const char *name)
{
    void *sym = dlsym(RTLD_DEFAULT, name);
    Dl_info info;
    if (sym && dladdr(sym, &info) && info.dli_fname) printf("binding: %s -> %s\n", name, info.dli_fname);
    else printf("binding: %s -> ?\n", name);
}

int main(int argc, char **argv)
{
    const char *probe[] = { "start_renderer", "check_partition_left", "set_mover_gl_center", "start_simulation",
                            "apply_gravity", "viscosity_impluses", "predict_positions", "identify_oob_particles",
                            "hash_fluid", "hash_halo", "startHaloExchange", "finishHaloExchange",
                            "double_density_relaxation", "updateVelocities", "partitionProblem", "setParticleNumbers",
                            "initParticles" };
    for (size_t i = 0; i < sizeof probe / sizeof *probe; i++) report_binding(probe[i]);
    int ranks = 3;
    const char *frames = "4", *out = "ref_world.bin";
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "--ranks")) ranks = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--frames")) frames = argv[i + 1];
        if (!strcmp(argv[i], "--out")) out = argv[i + 1];
    }
    setenv("SPH_RENDER_FRAMES", frames, 1);
    setenv("SPH_RENDER_OUT", out, 1);
    if (mini_mpi_world_create_render(ranks, (size_t)16 << 20)) { fprintf(stderr, "ref_world: cannot create %d ranks\n", ranks); return 2; }
    fflush(stdout);
    int me = -2;
    pid_t pids[257];
    for (int r = -1; r < ranks; r++) {           /* -1: the render rank */
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 2; }
        if (pid == 0) { me = r; break; }
        pids[r + 1] = pid;
    }
    if (me == -2) {
        int worst = 0, st;
        /* compute ranks return an uninitialised code (fluid.c:48,68): only a rank that was KILLED counts, and then
         * the others, which would wait for its messages for ever, go too */
        while (wait(&st) > 0)
            if (!WIFEXITED(st) && !worst) { worst = 5; for (int i = 0; i <= ranks; i++) kill(pids[i], SIGKILL); }
        return worst;
    }
    prctl(PR_SET_PDEATHSIG, SIGKILL);            /* nobody outlives the launcher (a test's time-out kills only that) */
    alarm(getenv("SPH_WORLD_TIMEOUT") ? (unsigned)atoi(getenv("SPH_WORLD_TIMEOUT")) : 240);
    if (me < 0) mini_mpi_bind_render(); else mini_mpi_bind(me);
    ref_main(argc, argv);
    fflush(stdout);
    _exit(0);
}
