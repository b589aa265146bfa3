/*
 * sph_ref_mpi_glue -- the one file a TinySPH build adds to run SEVERAL compute ranks on libsph_b200.so without
 * touching the reference's sources: compile it with the host's own MPI (mpicc -c, next to fluid.c) and put it on
 * the link line together with -lsph_b200.  Not part of libsph_b200.so, which does not link MPI.
 *
 * The library finds sph_ref_host_mpi() as a weak symbol (sph_b200/host/ref_api.c) the first time it needs to know
 * which slab it is -- partitionProblem, geometry.c:105-108 -- and moves its neighbour messages with the callback:
 * plain bytes between rank - 1 / rank + 1 of MPI_COMM_COMPUTE, the communicator and the neighbour arithmetic of
 * the reference's own exchange (communication.c:36-47, :144-161).
 */
#include <stddef.h>
#include <mpi.h>

#include "communication.h"      /* the reference's: MPI_COMM_COMPUTE (brings fluid.h, whose types sph_ref_api.h then reuses) */
#include "sph_ref_api.h"

#define SPH_GLUE_TAG 23         /* the reference uses 7, 8, 9 and 17 */

static void glue_sendrecv(const void *send, size_t send_bytes, int to_side, void *recv, size_t recv_bytes, int from_side,
                          void *user)
{
    (void)user;
    int rank, nranks;
    MPI_Comm_rank(MPI_COMM_COMPUTE, &rank);
    MPI_Comm_size(MPI_COMM_COMPUTE, &nranks);
    int to = to_side == 0 ? rank - 1 : rank + 1, from = from_side == 0 ? rank - 1 : rank + 1;
    if (!send || to < 0 || to >= nranks) to = MPI_PROC_NULL;
    if (!recv || from < 0 || from >= nranks) from = MPI_PROC_NULL;
    MPI_Sendrecv((void *)send, (int)send_bytes, MPI_CHAR, to, SPH_GLUE_TAG, recv, (int)recv_bytes, MPI_CHAR, from,
                 SPH_GLUE_TAG, MPI_COMM_COMPUTE, MPI_STATUS_IGNORE);
}

int sph_ref_host_mpi(int *rank, int *nranks, sph_sendrecv_fn *fn, void **user)
{
    MPI_Comm_rank(MPI_COMM_COMPUTE, rank);
    MPI_Comm_size(MPI_COMM_COMPUTE, nranks);
    *fn = glue_sendrecv;
    *user = NULL;
    return 0;
}
