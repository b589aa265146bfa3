/*
 * mini-MPI implementation (TEST INFRASTRUCTURE ONLY; see shim/mpi.h).
 *
 * Transport: ranks are processes created by fork() from the harness launcher.
 * TinySPH's compute ranks only ever talk to rank-1 / rank+1
 * (/root/reference/src/communication.c:144-145, :262-263), so the world is a
 * line of ranks with one FIFO byte ring per directed neighbour pair, living in
 * a MAP_SHARED anonymous mapping made before the fork.  Sends are eager
 * (copied into the ring at MPI_Isend/MPI_Send time), receives drain the ring
 * in order and check the tag.  Both sides of every exchange in the reference
 * run the same program order, so FIFO matching is exact.
 *
 * Datatypes are reduced to (element bytes, list of element displacements):
 * basic types, one "struct" type treated as a contiguous record of its extent,
 * and MPI_Type_indexed over a struct type (block length 1 everywhere in the
 * reference: communication.c:174-185, :287-299, :317-325).
 *
 * The render-rank protocol (Bcast/Gatherv/Scatterv/Probe/Irecv between world
 * rank 0 and the compute ranks) exists in two forms: hooks for a render stub
 * living inside the compute process (mini_mpi_set_render, ref_drive.c), and a
 * render rank that is a process of its own with one ring to and one from every
 * compute rank (mini_mpi_world_create_render, ref_world.c: the reference's own
 * renderer.c).  Without either, those calls abort.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>

/* ------------------------------------------------------------------ world */

typedef struct {
    volatile uint64_t head;   /* bytes consumed (receiver-owned) */
    char pad0[56];
    volatile uint64_t tail;   /* bytes produced (sender-owned)   */
    char pad1[56];
} ring_ctl_t;

typedef struct {
    int nranks;
    int render_proc;          /* world rank 0 is a process of its own (mini_mpi_world_create_render) */
    size_t ring_bytes;
    volatile int barrier_count;
    volatile int barrier_sense;
    double scratch[8 * 256];
} world_hdr_t;

static world_hdr_t *g_world = NULL;
static ring_ctl_t *g_ctl = NULL;   /* 2*nranks rings */
static char *g_data = NULL;
static int g_rank = 0;
static int g_nranks = 1;

static void die(const char *msg)
{
    fprintf(stderr, "mini-mpi[rank %d]: %s\n", g_rank, msg);
    abort();
}

static int g_render_proc = 0;      /* the world has a render process ... */
static int g_is_render = 0;        /* ... and this is it */

static int world_create(int nranks, size_t ring_bytes, int render_proc);
int mini_mpi_world_create(int nranks, size_t ring_bytes) { return world_create(nranks, ring_bytes, 0); }
/* K compute ranks plus the render rank as a process of its own: two more rings per compute rank (to / from it) */
int mini_mpi_world_create_render(int nranks, size_t ring_bytes) { return world_create(nranks, ring_bytes, 1); }
void mini_mpi_bind_render(void) { g_is_render = 1; g_rank = -1; }

static int world_create(int nranks, size_t ring_bytes, int render_proc)
{
    if (nranks < 1 || nranks > 256) return -1;
    size_t nrings = (render_proc ? 4 : 2) * (size_t)nranks;
    g_render_proc = render_proc;
    size_t bytes = sizeof(world_hdr_t) + nrings * sizeof(ring_ctl_t) + nrings * ring_bytes;
    void *m = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) return -1;
    g_world = (world_hdr_t *)m;
    g_world->nranks = nranks;
    g_world->ring_bytes = ring_bytes;
    g_world->render_proc = render_proc;
    g_world->barrier_count = 0;
    g_world->barrier_sense = 0;
    g_ctl = (ring_ctl_t *)((char *)m + sizeof(world_hdr_t));
    g_data = (char *)(g_ctl + nrings);
    for (size_t i = 0; i < nrings; i++) { g_ctl[i].head = 0; g_ctl[i].tail = 0; }
    g_nranks = nranks;
    g_rank = 0;
    return 0;
}

void mini_mpi_bind(int rank) { g_rank = rank; }

double *mini_mpi_shared_doubles(void)
{
    static double local[8 * 256];
    return g_world ? g_world->scratch : local;
}

void mini_mpi_barrier(void)
{
    if (!g_world || g_nranks == 1) return;
    int sense = g_world->barrier_sense;
    if (__sync_add_and_fetch(&g_world->barrier_count, 1) == g_nranks) {
        g_world->barrier_count = 0;
        __sync_synchronize();
        g_world->barrier_sense = !sense;
    } else {
        while (g_world->barrier_sense == sense) sched_yield();
    }
    __sync_synchronize();
}

/* ring for src -> dst (neighbours only) */
static int ring_id(int src, int dst)
{
    if (dst == src + 1) return 2 * src;
    if (dst == src - 1) return 2 * src + 1;
    die("non-neighbour message");
    return -1;
}

static void ring_write(int id, const void *p, size_t n)
{
    ring_ctl_t *c = &g_ctl[id];
    size_t cap = g_world->ring_bytes;
    if (n > cap) die("message larger than ring; raise ring_bytes");
    while (c->tail + n - c->head > cap) sched_yield();
    char *base = g_data + (size_t)id * cap;
    size_t off = (size_t)(c->tail % cap);
    size_t first = n < cap - off ? n : cap - off;
    memcpy(base + off, p, first);
    if (first < n) memcpy(base, (const char *)p + first, n - first);
    __sync_synchronize();
    c->tail += n;
}

static void ring_read(int id, void *p, size_t n)
{
    ring_ctl_t *c = &g_ctl[id];
    size_t cap = g_world->ring_bytes;
    while (c->tail - c->head < n) sched_yield();
    __sync_synchronize();
    char *base = g_data + (size_t)id * cap;
    size_t off = (size_t)(c->head % cap);
    size_t first = n < cap - off ? n : cap - off;
    memcpy(p, base + off, first);
    if (first < n) memcpy((char *)p + first, base, n - first);
    __sync_synchronize();
    c->head += n;
}

/* -------------------------------------------------------------- datatypes */

#define MAX_TYPES 64
typedef struct {
    int used;
    size_t elem;     /* bytes of one block */
    int nblocks;     /* number of blocks in ONE element of this type */
    int *disp;       /* displacement of each block, in units of elem (NULL: contiguous, 1 block) */
} dtype_t;

static dtype_t g_types[MAX_TYPES] = {
    {1, 0, 0, NULL},
    {1, 1, 1, NULL},  /* MPI_CHAR  */
    {1, 2, 1, NULL},  /* MPI_SHORT */
    {1, 4, 1, NULL},  /* MPI_INT   */
    {1, 4, 1, NULL},  /* MPI_FLOAT */
};

static dtype_t *T(MPI_Datatype t)
{
    if (t <= 0 || t >= MAX_TYPES || !g_types[t].used) die("bad datatype handle");
    return &g_types[t];
}

static int new_type(void)
{
    for (int i = 5; i < MAX_TYPES; i++)
        if (!g_types[i].used) { g_types[i].used = 1; return i; }
    die("out of datatype handles");
    return 0;
}

int MPI_Type_create_struct(int count, const int blocklens[], const MPI_Aint disps[],
                           const MPI_Datatype types[], MPI_Datatype *newtype)
{
    /* extent = end of the last member, rounded up to 4 (all members here are <= 4 bytes) */
    size_t end = 0;
    for (int i = 0; i < count; i++) {
        size_t e = (size_t)disps[i] + (size_t)blocklens[i] * T(types[i])->elem;
        if (e > end) end = e;
    }
    end = (end + 3) & ~(size_t)3;
    int h = new_type();
    g_types[h].elem = end;
    g_types[h].nblocks = 1;
    g_types[h].disp = NULL;
    *newtype = h;
    return MPI_SUCCESS;
}

int MPI_Type_indexed(int count, const int blocklens[], const int disps[],
                     MPI_Datatype oldtype, MPI_Datatype *newtype)
{
    dtype_t *o = T(oldtype);
    if (o->disp) die("nested indexed types unsupported");
    int h = new_type();
    g_types[h].elem = o->elem;
    g_types[h].nblocks = count;
    g_types[h].disp = (int *)malloc((count > 0 ? count : 1) * sizeof(int));
    for (int i = 0; i < count; i++) {
        if (blocklens[i] != 1) die("indexed block length != 1 unsupported");
        g_types[h].disp[i] = disps[i];
    }
    *newtype = h;
    return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype *type) { (void)type; return MPI_SUCCESS; }

int MPI_Type_free(MPI_Datatype *type)
{
    dtype_t *t = T(*type);
    if (*type < 5) die("freeing basic type");
    free(t->disp);
    t->disp = NULL;
    t->used = 0;
    *type = MPI_DATATYPE_NULL;
    return MPI_SUCCESS;
}

int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count)
{
    *count = (int)((size_t)status->_nbytes / T(type)->elem);
    return MPI_SUCCESS;
}

static size_t type_bytes(MPI_Datatype type, int count)
{
    dtype_t *t = T(type);
    return (size_t)count * (size_t)t->nblocks * t->elem;
}

/* ---------------------------------------------------------- point to point */

typedef struct { int tag; int nbytes; } msg_hdr_t;

static void do_send(const void *buf, int count, MPI_Datatype type, int dest, int tag)
{
    if (dest == MPI_PROC_NULL) return;
    if (dest < 0 || dest >= g_nranks) die("send to bad rank");
    dtype_t *t = T(type);
    int id = ring_id(g_rank, dest);
    msg_hdr_t h = { tag, (int)type_bytes(type, count) };
    ring_write(id, &h, sizeof h);
    if (!t->disp) {
        ring_write(id, buf, (size_t)h.nbytes);
    } else {
        for (int c = 0; c < count; c++)
            for (int b = 0; b < t->nblocks; b++)
                ring_write(id, (const char *)buf + (size_t)t->disp[b] * t->elem, t->elem);
    }
}

static void do_recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Status *status)
{
    MPI_Status local;
    MPI_Status *st = status ? status : &local;
    st->MPI_SOURCE = src;
    st->MPI_TAG = tag;
    st->MPI_ERROR = MPI_SUCCESS;
    st->_nbytes = 0;
    if (src == MPI_PROC_NULL) return;
    if (src < 0 || src >= g_nranks) die("recv from bad rank");
    dtype_t *t = T(type);
    int id = ring_id(src, g_rank);
    msg_hdr_t h;
    ring_read(id, &h, sizeof h);
    if (h.tag != tag) {
        fprintf(stderr, "mini-mpi[rank %d]: tag mismatch from %d: got %d want %d\n", g_rank, src, h.tag, tag);
        abort();
    }
    if ((size_t)h.nbytes > type_bytes(type, count)) die("message truncated");
    st->_nbytes = h.nbytes;
    if (!t->disp) {
        ring_read(id, buf, (size_t)h.nbytes);
    } else {
        size_t n = (size_t)h.nbytes / t->elem;
        for (size_t b = 0; b < n; b++)
            ring_read(id, (char *)buf + (size_t)t->disp[b] * t->elem, t->elem);
    }
}

static int to_render(MPI_Comm comm, int rank);
static int send_to_render(const void *buf, int count, MPI_Datatype type, int tag);

/* ---- the render rank as a process of its own: rings 2K + 2r (compute r -> render) and 2K + 2r + 1 (render -> r) ---- */
static int ring_up(int r) { return 2 * g_nranks + 2 * r; }
static int ring_down(int r) { return 2 * g_nranks + 2 * r + 1; }
static int compute_to_render_proc(MPI_Comm comm, int rank) { return g_render_proc && !g_is_render && comm == MPI_COMM_WORLD && rank == 0; }

static void raw_send(int id, const void *buf, size_t bytes, int tag)
{
    msg_hdr_t h = { tag, (int)bytes };
    ring_write(id, &h, sizeof h);
    if (bytes) ring_write(id, buf, bytes);
}

static size_t raw_recv(int id, void *buf, size_t room, int tag)
{
    msg_hdr_t h;
    ring_read(id, &h, sizeof h);
    if (h.tag != tag) { fprintf(stderr, "mini-mpi: render protocol: got tag %d, want %d\n", h.tag, tag); abort(); }
    if ((size_t)h.nbytes > room) die("render protocol: message truncated");
    if (h.nbytes) ring_read(id, buf, (size_t)h.nbytes);
    return (size_t)h.nbytes;
}

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
    if (compute_to_render_proc(comm, dest)) { raw_send(ring_up(g_rank), buf, type_bytes(type, count), tag); return MPI_SUCCESS; }
    if (to_render(comm, dest)) return send_to_render(buf, count, type, tag);
    do_send(buf, count, type, dest, tag);
    return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status *status)
{
    if (g_is_render) {      /* renderer.c:133,138: from world rank src = compute rank src - 1 */
        size_t n = raw_recv(ring_up(src - 1), buf, type_bytes(type, count), tag);
        if (status) { status->MPI_SOURCE = src; status->MPI_TAG = tag; status->MPI_ERROR = MPI_SUCCESS; status->_nbytes = (int)n; }
        return MPI_SUCCESS;
    }
    (void)comm; do_recv(buf, count, type, src, tag, status); return MPI_SUCCESS;
}

int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int src, int recvtag,
                 MPI_Comm comm, MPI_Status *status)
{
    (void)comm;
    do_send(sendbuf, sendcount, sendtype, dest, sendtag);
    do_recv(recvbuf, recvcount, recvtype, src, recvtag, status);
    return MPI_SUCCESS;
}

/* requests: sends complete eagerly; receives are deferred to Wait/Waitall */
#define MAX_REQS 64
typedef struct { int used; int is_recv; void *buf; int count; MPI_Datatype type; int src; int tag; } req_t;
static req_t g_reqs[MAX_REQS];

static int new_req(void)
{
    for (int i = 0; i < MAX_REQS; i++)
        if (!g_reqs[i].used) { g_reqs[i].used = 1; return i; }
    die("out of request handles");
    return -1;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
    if (compute_to_render_proc(comm, dest)) raw_send(ring_up(g_rank), buf, type_bytes(type, count), tag);
    else if (to_render(comm, dest)) send_to_render(buf, count, type, tag);
    else do_send(buf, count, type, dest, tag);
    int r = new_req();
    g_reqs[r].is_recv = 0;
    *req = r;
    return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Request *req)
{
    (void)comm;
    if (g_is_render) {
        /* renderer.c:276-282 posts this right after MPI_Probe(MPI_ANY_SOURCE): the message must be CONSUMED here, or
         * the next Probe would report the same rank again */
        raw_recv(ring_up(src - 1), buf, type_bytes(type, count), tag);
        *req = MPI_REQUEST_NULL;
        return MPI_SUCCESS;
    }
    int r = new_req();
    g_reqs[r].is_recv = 1;
    g_reqs[r].buf = buf; g_reqs[r].count = count; g_reqs[r].type = type;
    g_reqs[r].src = src; g_reqs[r].tag = tag;
    *req = r;
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{
    if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    req_t *r = &g_reqs[*req];
    if (!r->used) die("wait on stale request");
    if (r->is_recv) do_recv(r->buf, r->count, r->type, r->src, r->tag, status);
    else if (status) { status->_nbytes = 0; status->MPI_ERROR = MPI_SUCCESS; }
    r->used = 0;
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[])
{
    for (int i = 0; i < count; i++) MPI_Wait(&reqs[i], statuses ? &statuses[i] : NULL);
    return MPI_SUCCESS;
}

/* ------------------------------------------------------------- bookkeeping */

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Communicators.  Without a render stub every communicator is "all compute ranks" (the harness sets
 * MPI_COMM_COMPUTE = MPI_COMM_WORLD).  With one (mini_mpi_set_render, used by ref_drive.c to run the
 * reference's own start_simulation) the world has one more rank in front, the render rank, which lives
 * in the hooks: MPI_COMM_WORLD ranks are shifted by one and MPI_Comm_create hands out the compute
 * communicator (communication.c:36-47). */
#define COMM_COMPUTE 2
static const mini_mpi_render_t *g_render = NULL;
void mini_mpi_set_render(const mini_mpi_render_t *r) { g_render = r; }
static int to_render(MPI_Comm comm, int rank) { return g_render && comm == MPI_COMM_WORLD && rank == 0; }

int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = g_rank + ((g_render || g_render_proc) && comm == MPI_COMM_WORLD); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = g_nranks + ((g_render || g_render_proc) && comm == MPI_COMM_WORLD); return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) { (void)comm; *group = 1; return MPI_SUCCESS; }
int MPI_Group_excl(MPI_Group g, int n, const int r[], MPI_Group *ng) { (void)g; (void)n; (void)r; *ng = 2; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group g, int n, const int r[], MPI_Group *ng) { (void)g; (void)n; (void)r; *ng = 3; return MPI_SUCCESS; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *nc)
{ (void)c; *nc = ((g_render || g_render_proc) && g == 2) ? COMM_COMPUTE : MPI_COMM_WORLD; return MPI_SUCCESS; }
int MPI_Group_free(MPI_Group *g) { *g = 0; return MPI_SUCCESS; }

/* Render-rank protocol (fluid.c:122-124, :167-171, :238, :293-294, :365): served by the hooks when a
 * render stub is installed, otherwise deliberately unimplemented (see file header). */
#define UNIMPL(name) do { die(name " is part of the render-rank protocol, not provided by this shim"); return -1; } while (0)
int MPI_Probe(int s, int t, MPI_Comm c, MPI_Status *st)
{
    (void)s; (void)c;
    if (!g_is_render) UNIMPL("MPI_Probe");
    for (;;) {      /* renderer.c:276: any compute rank whose next message has at least its header in the ring */
        for (int r = 0; r < g_nranks; r++) {
            ring_ctl_t *ctl = &g_ctl[ring_up(r)];
            if (ctl->tail - ctl->head < sizeof(msg_hdr_t)) continue;
            __sync_synchronize();
            msg_hdr_t h;
            size_t cap = g_world->ring_bytes, off = (size_t)(ctl->head % cap);
            char *base = g_data + (size_t)ring_up(r) * cap;
            size_t first = sizeof h < cap - off ? sizeof h : cap - off;
            memcpy(&h, base + off, first);
            if (first < sizeof h) memcpy((char *)&h + first, base, sizeof h - first);
            if (h.tag != t) die("MPI_Probe: unexpected tag at the head of a ring");
            st->MPI_SOURCE = r + 1; st->MPI_TAG = h.tag; st->MPI_ERROR = MPI_SUCCESS; st->_nbytes = h.nbytes;
            return MPI_SUCCESS;
        }
        sched_yield();
    }
}
#define TAG_BCAST (-101)
#define TAG_SCATTER (-102)
int MPI_Bcast(void *b, int n, MPI_Datatype t, int r, MPI_Comm c)
{
    if (g_render_proc && c == MPI_COMM_WORLD && r == 0) {
        if (g_is_render) for (int k = 0; k < g_nranks; k++) raw_send(ring_down(k), b, type_bytes(t, n), TAG_BCAST);
        else raw_recv(ring_down(g_rank), b, type_bytes(t, n), TAG_BCAST);
        return MPI_SUCCESS;
    }
    if (!to_render(c, r)) UNIMPL("MPI_Bcast");
    g_render->bcast(b, type_bytes(t, n));
    return MPI_SUCCESS;
}
int MPI_Gatherv(const void *sb, int sc, MPI_Datatype st, void *rb, const int rc[], const int d[], MPI_Datatype rt, int r, MPI_Comm c)
{
    if (g_render_proc && c == MPI_COMM_WORLD && r == 0) {
        if (g_is_render) {      /* renderer.c:151: one block from every compute rank, world rank k lands at displacement d[k] */
            for (int k = 1; k <= g_nranks; k++)
                raw_recv(ring_up(k - 1), (char *)rb + (size_t)d[k] * type_bytes(rt, 1), type_bytes(rt, rc[k]), MINI_MPI_TAG_GATHER);
        } else raw_send(ring_up(g_rank), sb, type_bytes(st, sc), MINI_MPI_TAG_GATHER);
        return MPI_SUCCESS;
    }
    (void)rb; (void)rc; (void)d; (void)rt;
    if (!to_render(c, r)) UNIMPL("MPI_Gatherv");
    g_render->from_compute(sb, type_bytes(st, sc), MINI_MPI_TAG_GATHER);
    return MPI_SUCCESS;
}
int MPI_Scatterv(const void *sb, const int sc[], const int d[], MPI_Datatype st, void *rb, int rc, MPI_Datatype rt, int r, MPI_Comm c)
{
    if (g_render_proc && c == MPI_COMM_WORLD && r == 0) {
        if (g_is_render) {      /* renderer.c:248,268 */
            for (int k = 1; k <= g_nranks; k++)
                raw_send(ring_down(k - 1), (const char *)sb + (size_t)d[k] * type_bytes(st, 1), type_bytes(st, sc[k]), TAG_SCATTER);
        } else raw_recv(ring_down(g_rank), rb, type_bytes(rt, rc), TAG_SCATTER);
        return MPI_SUCCESS;
    }
    (void)sb; (void)sc; (void)d; (void)st;
    if (!to_render(c, r)) UNIMPL("MPI_Scatterv");
    g_render->scatter(rb, type_bytes(rt, rc));
    return MPI_SUCCESS;
}

static int send_to_render(const void *buf, int count, MPI_Datatype type, int tag)
{
    if (T(type)->disp) die("indexed type sent to the render rank");
    g_render->from_compute(buf, type_bytes(type, count), tag);
    return MPI_SUCCESS;
}

/* fluid.c's main() (renamed by -Dmain=ref_main) references the render rank. */
#ifndef MINI_MPI_WITH_RENDERER   /* (the full-program build links the reference's own renderer.c) */
int start_renderer(void) { die("start_renderer: no render rank in the oracle build"); return -1; }
#endif
