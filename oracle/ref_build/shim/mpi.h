/*
 * mini-MPI: the 26 MPI calls TinySPH's compute path references, implemented
 * from scratch for a box without mpicc/mpirun (see oracle/ref_build/mpi_shim.c).
 *
 * TEST INFRASTRUCTURE ONLY.  This header exists so the reference's own
 * fluid.c / hash.c / geometry.c / communication.c compile UNMODIFIED from
 * /root/reference/src into oracle/_ref/ (the parity oracle and the CPU
 * baseline).  Nothing in the product path (sph_b200/, include/) includes it.
 *
 * Semantics the reference relies on:
 *   - MPI_Request is a scalar handle comparable with MPI_REQUEST_NULL
 *   - MPI_Get_count counts in units of the datatype passed
 *   - sends to MPI_PROC_NULL are no-ops, receives from it deliver 0 elements
 *   - MPI_Type_indexed is created/freed every call (handles are recycled)
 */
#ifndef SPH_MINI_MPI_H
#define SPH_MINI_MPI_H

#include <stddef.h>

typedef int MPI_Datatype;
typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Request;
typedef ptrdiff_t MPI_Aint;

typedef struct {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    int _nbytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0

#define MPI_CHAR 1
#define MPI_SHORT 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DATATYPE_NULL 0

#define MPI_PROC_NULL (-2)
#define MPI_ANY_SOURCE (-1)
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
double MPI_Wtime(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_excl(MPI_Group group, int n, const int ranks[], MPI_Group *newgroup);
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group *newgroup);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm);
int MPI_Group_free(MPI_Group *group);

int MPI_Type_create_struct(int count, const int blocklens[], const MPI_Aint disps[],
                           const MPI_Datatype types[], MPI_Datatype *newtype);
int MPI_Type_indexed(int count, const int blocklens[], const int disps[],
                     MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Type_free(MPI_Datatype *type);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count);

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]);
int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int src, int recvtag,
                 MPI_Comm comm, MPI_Status *status);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                const int recvcounts[], const int displs[], MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatterv(const void *sendbuf, const int sendcounts[], const int displs[], MPI_Datatype sendtype,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);

/* ---- shim control (not MPI): used by oracle/ref_build/ref_harness.c ---- */
/* Create the shared mailboxes for `nranks` ranks; call BEFORE fork(). */
int mini_mpi_world_create(int nranks, size_t ring_bytes);
/* The same plus world rank 0 as a process of its own (the reference's render rank, oracle/ref_build/ref_world.c):
 * Bcast / Recv / Gatherv / Scatterv / Probe / Irecv between it and the compute ranks, as renderer.c uses them. */
int mini_mpi_world_create_render(int nranks, size_t ring_bytes);
void mini_mpi_bind_render(void);
/* Bind the calling process to `rank` (call in each child after fork). */
void mini_mpi_bind(int rank);
/* Spin barrier across all ranks of the world. */
void mini_mpi_barrier(void);
/* Small shared scratch array of doubles (nranks*8 slots) for harness reductions. */
double *mini_mpi_shared_doubles(void);


/* ---- headless render rank living in the same process (oracle/ref_build/ref_drive.c) ----
 * The reference's start_simulation talks to rank 0 of MPI_COMM_WORLD: Bcast of the pixel size
 * (fluid.c:122-124), Sends of the world size and particle count (:167-171), Gatherv of the first
 * parameter block (:238), one Scatterv of parameters (:293-294) and one Isend of int16 coordinates
 * (:365) per frame.  With these hooks installed the shim serves exactly that and nothing else. */
#define MINI_MPI_TAG_GATHER (-100)
typedef struct {
    void (*bcast)(void *buf, size_t bytes);
    void (*from_compute)(const void *buf, size_t bytes, int tag);   /* Send, Isend, Gatherv (MINI_MPI_TAG_GATHER) */
    void (*scatter)(void *buf, size_t bytes);
} mini_mpi_render_t;
void mini_mpi_set_render(const mini_mpi_render_t *r);

#endif
