// C-ABI of the B200-native TinySPH compute-rank timestep (include/sph_b200.h).
// Host side of the CUDA library: context, buffers, stage launches, CUDA-graph step loop.
// There is no CPU path in this file: every failure to reach the GPU is SPH_ERR_CUDA.
#include "../../include/sph_b200.h"
#include "sph_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// every kernel launch of the library: SPH_THREADS threads per block, no dynamic shared memory
#if !defined(SPH_EMU) && !SPH_PDL
#define SPH_LAUNCH(kernel, grid, stream) kernel<<<(grid), SPH_THREADS, 0, (stream)>>>
#elif !defined(SPH_EMU)
// -DSPH_PDL=1: every launch may be scheduled while the kernel before it in the stream drains (pdl_enter() in
// sph_device.cuh is the other half); stream capture records these as programmatic edges of the step graph
template <class... A> struct PdlLauncher {
    void (*kernel)(A...);
    int grid;
    cudaStream_t stream;
    template <class... B> void operator()(B... args) const
    {
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(SPH_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kernel, static_cast<A>(args)...);     // errors: cudaGetLastError() after the launch
    }
};
template <class... A> static PdlLauncher<A...> pdl_launcher(void (*k)(A...), int grid, cudaStream_t s) { return PdlLauncher<A...>{k, grid, s}; }
#define SPH_LAUNCH(kernel, grid, stream) pdl_launcher(kernel, (grid), (stream))
#else       // tests/emu: this file compiled by g++ against a fake runtime (test infrastructure, see sph_device.cuh)
#define SPH_LAUNCH(kernel, grid, stream) emu::make_launcher(kernel, (grid), SPH_THREADS, (stream))
#endif

#ifndef SPH_GRID_MULT
#define SPH_GRID_MULT 8          // blocks per SM of the gather kernels (0: one block per 256 entries of capacity)
#endif
#ifndef SPH_GRID_ADVECT
#define SPH_GRID_ADVECT (SPH_GRID_MULT > 0 ? 2 * SPH_GRID_MULT : 1 << 20)
#endif
#ifndef SPH_GRID_DENSITY
#define SPH_GRID_DENSITY (SPH_GRID_MULT > 0 ? 2 * SPH_GRID_MULT : 1 << 20)
#endif
#ifndef SPH_GRID_RELAX
#define SPH_GRID_RELAX (SPH_GRID_MULT > 0 ? 3 * SPH_GRID_MULT / 2 : 1 << 20)
#endif
#ifndef SPH_GRID_MULT_SORT
#define SPH_GRID_MULT_SORT 4     // blocks per SM of the sort's scatter and reorder kernels (4 entries per thread and trip)
#endif
#ifndef SPH_GRID_MULT_SCAN
#define SPH_GRID_MULT_SCAN 8     // blocks per SM of the scan (one 1024-cell tile per block and trip)
#endif

enum { ST_READY = 0, ST_ADVECTED, ST_SORTED1, ST_DENSITY, ST_RELAXED, ST_REQUEUED };

struct sph_ctx {
    sph_config cfg;
    cudaStream_t stream;
    bool own_stream;
    DevParams hp;            // host shadow of *dp, always describing the state after all launched work
    DevParams *dp;
    sph_tunable tun, queued;
    bool have_queued;
    int *counters;
    // rotating SoA buffers: see the role table in launch_* below
    float2 *P[4];
    float2 *Q[3];
    uint32_t *U[2];
    float2 *dens;
    sph_mask_t *nmask;               // SPH_NROWS x capacity: per-row acceptance masks from k_density for k_relax
#if SPH_RELAX_PD4
    float4 *pd;                      // (x, y, density, density_near) per entry: k_density -> k_relax
#endif
    float *coupling;                 // per entry: sum of its pairs' viscosity coefficients (stabilised viscosity gather only)
    DevOptions *dopt;                // device copy of the optional-path parameters
    float visc_gamma, visc_min_dt_sigma;
    int *cnt, *cell_start, *t_key, *t_slot, *ord_src, *ord_key;
    uint32_t *ord_uid;
    float4 *pv;                      // SPH_ADVECT_PV4: (x, y, vx, vy) of the resident entries after sort 2
    int *tile_total;                 // one population total per scan tile
    int ntiles_max;
    unsigned char *send[2], *recv[2];
    unsigned char *xchg;             // exchange block for peer-memory mode (flags + 8 message buffers)
    void *peer[2];                   // neighbours' exchange blocks mapped with cudaIpcOpenMemHandle
    int scan_grid;                   // tiles of the widest possible window
    long long *xt;                   // k_unpack's time sums [ns]: sending, waiting, unpacking, meetings, ... (see the kernel)
    DevParams *stage_dp;             // pinned: the two parameter blocks a captured queued-parameter step copies in
    cudaEvent_t stage_free;          // ... and the point in the stream after which they may be rewritten
    bool stage_busy;
    bool dens_seen = false;          // a density pass has run since the last upload (its running count primes sph_get_status)
    int unpack_grid;                 // k_unpack waits on the neighbour inside the kernel: grid must be fully co-resident
    short2 *coords;                  // device-side frame of the synchronous feed and of ticket 0
    short2 *coords1;                 // ... of ticket 1 (the asynchronous feed packs frame f while frame f-1 still drains)
    // asynchronous coordinate feed (sph_pack_coords_async): the copy of frame f drains on its own stream while the
    // steps of frame f+1 run, like the reference's MPI_Isend of its frame (fluid.c:283-287, :354-365)
    cudaStream_t copy_stream;
    struct { cudaEvent_t packed, copied; int cap; bool pending; int entries; int16_t *xy; } feed[2];
    int feed_last_n;                 // a slab's population as of the last collected frame (0: none yet)
    int feed_margin;                 // entries copied beyond feed_last_n * 9/8 (SPH_FEED_MARGIN_ENTRIES; tests make it negative)
    int *feed_cnt_dev;               // 2 x CN_COUNT: counters as they stood when the frame was packed
    int *feed_cnt_host;              // the same, in pinned host memory
    unsigned feed_seq;
    int n_uploaded;                  // single slab: the particle count never changes after an upload
    unsigned char *stage_host;       // sph_exchange_via_host: 4 pinned message buffers (send l/r, recv l/r), allocated on first use
    int stage;
    int grid;                        // per-particle gather kernels
    int sort_grid;                   // the sort's streaming kernels
    int grid_advect, grid_density, grid_relax;
    int size_x, size_y;
    cudaGraphExec_t graph[8];        // whole step: index = stabilised viscosity gather + 2 * exchange step + 4 * queued block
    bool graph_ready[8];
    // exchange period (one-exchange build, sph_set_exchange_period): neighbours meet every `xperiod` steps
    int launch_per_graph[8];         // kernel launches inside each captured step
    int xperiod;                     // requested
    int since_x;                     // steps since the last exchange step
    bool force_x;                    // the coming step must exchange (fresh upload, parameters set outside the queue, ...)
    bool cur_x;                      // the step in progress is an exchange step
    bool one_x;                      // one exchange per step (sph_config.exchanges_per_step)
    long long launches;
    long long steps;
    // device-memory snapshot of the state at a step boundary (sph_state_save / sph_state_restore)
    struct {
        bool valid;
        float2 *P, *Q; uint32_t *U; int *cell_start, *ord_key, *counters;
        DevParams hp; sph_tunable tun; long long steps;
    } snap;
    char err[256];
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof ctx->err, "%s:%d %s: %s", __FILE__, __LINE__, #call,         \
                     cudaGetErrorString(e_));                                                      \
            return SPH_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static int fail(sph_ctx *ctx, int code, const char *msg)
{
    snprintf(ctx->err, sizeof ctx->err, "%s", msg);
    return code;
}

// window of grid columns a slab can touch: slab + ghost layer + one spare column
static void compute_window(const sph_ctx *ctx, float edge_start, float edge_end, int *gx0, int *wx)
{
    // in reference cells first, then refined to sort-grid columns
    if (ctx->cfg.nranks <= 1) { *gx0 = 0; *wx = ctx->size_x * SPH_CELL_DIV; return; }
    float w = ctx->cfg.halo_width * ctx->cfg.h;
    int lo = (int)floorf((edge_start - w) / ctx->cfg.h) - 1;
    int hi = (int)floorf((edge_end + w) / ctx->cfg.h) + 1;
    lo = std::max(lo, 0);
    hi = std::min(hi, ctx->size_x - 1);
    hi = std::max(hi, lo);
    *gx0 = lo * SPH_CELL_DIV;
    *wx = (hi - lo + 1) * SPH_CELL_DIV;
}

static void fill_phys(DevParams &P, const sph_tunable &t)
{
    P.rest_density = t.rest_density; P.h = t.smoothing_radius; P.g = t.g; P.k = t.k; P.k_near = t.k_near;
    P.k_spring = t.k_spring; P.sigma = t.sigma; P.beta = t.beta; P.dt = t.time_step;
    P.mover_cx = t.mover_center_x; P.mover_cy = t.mover_center_y; P.mover_w = t.mover_width;
    P.mover_h = t.mover_height; P.mover_type = (int)t.mover_type;
}

static void fill_edges(sph_ctx *ctx, float s, float e)
{
    ctx->hp.edge_start = s; ctx->hp.edge_end = e;
    compute_window(ctx, s, e, &ctx->hp.gx0_new, &ctx->hp.wx_new);
}

static int push_params(sph_ctx *ctx)
{
    // pageable source: the runtime stages the small block before returning, so hp may change right away
    CK(cudaMemcpyAsync(ctx->dp, &ctx->hp, sizeof(DevParams), cudaMemcpyHostToDevice, ctx->stream));
    return SPH_OK;
}

extern "C" int sph_create(const sph_config *cfg, sph_ctx **out)
{
    if (!cfg || !out || cfg->capacity <= 0 || cfg->h <= 0.0f || cfg->nranks < 1) return SPH_ERR_ARG;
    sph_ctx *ctx = new sph_ctx();
    memset(ctx, 0, sizeof *ctx);
    *out = ctx;
    ctx->cfg = *cfg;
    if (cfg->exchanges_per_step < 0 || cfg->exchanges_per_step > 2) return SPH_ERR_ARG;
    ctx->one_x = cfg->exchanges_per_step == 1 || (cfg->exchanges_per_step == 0 && SPH_ONE_EXCHANGE);
    if (ctx->cfg.halo_width <= 0.0f) ctx->cfg.halo_width = ctx->one_x ? 3.5f : 2.0f;
    if (ctx->one_x && cfg->nranks > 1 && ctx->cfg.halo_width < 3.0f)
        return fail(ctx, SPH_ERR_ARG, "one-exchange build: the ghost layer must be at least 3 h wide (4 h with the stabilised viscosity gather)");
    if (ctx->cfg.msg_capacity <= 0) ctx->cfg.msg_capacity = 1;
    ctx->cfg.msg_capacity = (ctx->cfg.msg_capacity + 3) & ~3;       // message sections on 16-byte boundaries (copy_words)
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(ctx, SPH_ERR_CUDA, "no CUDA device: sph_b200 has no CPU path");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (cfg->stream) { ctx->stream = (cudaStream_t)cfg->stream; ctx->own_stream = false; }
    else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }

    ctx->size_x = (int)ceil((cfg->tank_w - 0.0f) / cfg->h);      // fluid.c:214-215
    ctx->size_y = (int)ceil((cfg->tank_h - 0.0f) / cfg->h);
    const size_t cap = (size_t)cfg->capacity;
    const size_t ncell_max = (size_t)ctx->size_x * ctx->size_y * SPH_CELL_DIV * SPH_CELL_DIV;
    // Blocks of the per-particle kernels.  SPH_GRID_MULT > 0: persistent grid-stride blocks, that many per SM;
    // 0: one block per 256 entries of CAPACITY (the population lives on the device and a captured graph must
    // survive its changes; blocks beyond it return at once), so the block scheduler balances the chunks dynamically.
#if SPH_GRID_MULT > 0
    ctx->grid = std::min<int>((int)((cap + SPH_THREADS - 1) / SPH_THREADS), prop.multiProcessorCount * SPH_GRID_MULT);
#else
    ctx->grid = (int)((cap + SPH_THREADS - 1) / SPH_THREADS);
#endif
    // per-kernel block counts (A/B of round 2: k_advect and k_density like 16 blocks per SM, k_relax 12)
    ctx->grid_advect = std::min<int>((int)((cap + SPH_THREADS - 1) / SPH_THREADS), prop.multiProcessorCount * SPH_GRID_ADVECT);
    ctx->grid_density = std::min<int>((int)((cap + SPH_THREADS - 1) / SPH_THREADS), prop.multiProcessorCount * SPH_GRID_DENSITY);
    ctx->grid_relax = std::min<int>((int)((cap + SPH_THREADS - 1) / SPH_THREADS), prop.multiProcessorCount * SPH_GRID_RELAX);
    ctx->sort_grid = std::min<int>((int)((cap + SPH_THREADS - 1) / SPH_THREADS), prop.multiProcessorCount * SPH_GRID_MULT_SORT);

    // (SPH_PAIRMASK: a masked pair trip reads one entry past a candidate range and discards it; the entry must exist and
    //  be finite, so the arrays the gathers read are padded and start out as zeros)
    const size_t pad = 4;
    for (int i = 0; i < 4; i++) { CK(cudaMalloc(&ctx->P[i], (cap + pad) * sizeof(float2))); CK(cudaMemset(ctx->P[i], 0, (cap + pad) * sizeof(float2))); }
    for (int i = 0; i < 3; i++) { CK(cudaMalloc(&ctx->Q[i], (cap + pad) * sizeof(float2))); CK(cudaMemset(ctx->Q[i], 0, (cap + pad) * sizeof(float2))); }
    for (int i = 0; i < 2; i++) CK(cudaMalloc(&ctx->U[i], cap * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->dens, cap * sizeof(float2)));
    CK(cudaMalloc(&ctx->nmask, SPH_NROWS * cap * sizeof(sph_mask_t)));
#if SPH_RELAX_PD4
    CK(cudaMalloc(&ctx->pd, cap * sizeof(float4)));
#endif
#if SPH_ADVECT_PV4
    CK(cudaMalloc(&ctx->pv, (cap + pad) * sizeof(float4)));
    CK(cudaMemset(ctx->pv, 0, (cap + pad) * sizeof(float4)));
#endif
    CK(cudaMalloc(&ctx->coupling, (cap + pad) * sizeof(float)));
    CK(cudaMemset(ctx->coupling, 0, (cap + pad) * sizeof(float)));
    CK(cudaMalloc(&ctx->dopt, sizeof(DevOptions)));
    {
        // Default: the stabilised viscosity gather engages by itself for parameter blocks with dt*sigma >= 0.5, i.e.
        // for the reference's goo preset (controls.c:359-371: 0.83) and for none of its other presets (<= 0.17),
        // whose results it would not change by a bit anyway.  The plain gather does not settle that preset
        // (DESIGN.md 5b); this is the library's one algorithmic deviation and needs no call and no environment.
        ctx->visc_gamma = 0.5f; ctx->visc_min_dt_sigma = 0.5f;
        DevOptions o; o.visc_gamma = ctx->visc_gamma;
        CK(cudaMemcpy(ctx->dopt, &o, sizeof o, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&ctx->cnt, (ncell_max + 1) * sizeof(int)));
    CK(cudaMalloc(&ctx->cell_start, (ncell_max + 1) * sizeof(int)));
    CK(cudaMalloc(&ctx->t_key, cap * sizeof(int)));
    CK(cudaMalloc(&ctx->t_slot, cap * sizeof(int)));
    CK(cudaMalloc(&ctx->ord_src, cap * sizeof(int)));
    CK(cudaMalloc(&ctx->ord_key, cap * sizeof(int)));
    CK(cudaMalloc(&ctx->ord_uid, cap * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->coords, cap * sizeof(short2)));
    CK(cudaMalloc(&ctx->coords1, cap * sizeof(short2)));
    ctx->feed_margin = 4096;
    if (const char *fm = getenv("SPH_FEED_MARGIN_ENTRIES")) ctx->feed_margin = atoi(fm);
    const size_t ntiles_max = (ncell_max + SCAN_TILE - 1) / SCAN_TILE + 1;
    ctx->ntiles_max = (int)ntiles_max;
    CK(cudaMalloc(&ctx->tile_total, ntiles_max * sizeof(int)));
    CK(cudaMemset(ctx->tile_total, 0, ntiles_max * sizeof(int)));
    CK(cudaMemset(ctx->cnt, 0, (ncell_max + 1) * sizeof(int)));
    CK(cudaMemset(ctx->cell_start, 0, (ncell_max + 1) * sizeof(int)));
    CK(cudaMalloc(&ctx->counters, CN_COUNT * sizeof(int)));
    CK(cudaMemset(ctx->counters, 0, CN_COUNT * sizeof(int)));
    const size_t mb = msg_bytes_full(ctx->cfg.msg_capacity);
    for (int s = 0; s < 2; s++) {
        CK(cudaMalloc(&ctx->send[s], mb));
        CK(cudaMalloc(&ctx->recv[s], mb));
        CK(cudaMemset(ctx->send[s], 0, mb));
        CK(cudaMemset(ctx->recv[s], 0, mb));
    }
    CK(cudaMalloc(&ctx->xt, 8 * sizeof(long long)));
    CK(cudaMemset(ctx->xt, 0, 8 * sizeof(long long)));
    CK(cudaMallocHost(&ctx->stage_dp, 2 * sizeof(DevParams)));
    CK(cudaEventCreateWithFlags(&ctx->stage_free, cudaEventDisableTiming));
    CK(cudaMalloc(&ctx->xchg, xchg_bytes(ctx->cfg.msg_capacity)));
    CK(cudaMemset(ctx->xchg, 0, xchg_bytes(ctx->cfg.msg_capacity)));
    ctx->scan_grid = (int)((ncell_max + SCAN_TILE - 1) / SCAN_TILE);
    {
        // k_unpack sends, then spins until the neighbour's message has arrived, and a message is only
        // released once EVERY block of the sender's grid has run.  If the grid did not fit on the device at
        // once, the resident blocks of both neighbours would wait for each other's unscheduled blocks
        // forever.  So its grid is bounded by what the occupancy calculator says is co-resident.
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_unpack, SPH_THREADS, 0));
        if (per_sm < 1) return fail(ctx, SPH_ERR_CUDA, "k_unpack cannot be resident");
        // (and kept small: the messages are a few hundred KB, and every block pays for a fence and a flag poll)
        ctx->unpack_grid = std::min(std::min(ctx->grid, per_sm * prop.multiProcessorCount), 2 * prop.multiProcessorCount);
    }
    CK(cudaMalloc(&ctx->dp, sizeof(DevParams)));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreateWithFlags(&ctx->feed[k].packed, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->feed[k].copied, cudaEventDisableTiming));
    }
    CK(cudaMalloc(&ctx->feed_cnt_dev, 2 * CN_COUNT * sizeof(int)));
    CK(cudaMallocHost(&ctx->feed_cnt_host, 2 * CN_COUNT * sizeof(int)));

    DevParams &P = ctx->hp;
    P.tank_w = cfg->tank_w; P.tank_h = cfg->tank_h; P.cell_h = cfg->h;
    P.size_x = ctx->size_x; P.size_y = ctx->size_y; P.sort_rows = ctx->size_y * SPH_CELL_DIV;
    P.halo_w = ctx->cfg.halo_width * cfg->h;
    P.has_left = cfg->rank > 0; P.has_right = cfg->rank < cfg->nranks - 1; P.nranks = cfg->nranks;
    P.cap = cfg->capacity; P.msg_cap = ctx->cfg.msg_capacity;
    P.one_x = ctx->one_x ? 1 : 0;
    P.p2p = 0; P.xchg_base = (unsigned long long)ctx->xchg; P.remote_base[0] = P.remote_base[1] = 0;
    {   // device-side waits give up after this long (default 10 s) instead of hanging the GPU
        const char *ms = getenv("SPH_SPIN_TIMEOUT_MS");
        const double khz = prop.clockRate > 0 ? (double)prop.clockRate : 1.9e6;
        P.spin_timeout = (long long)((ms ? atof(ms) : 10000.0) * khz);
    }
    P.h = cfg->h; P.dt = 1.0f / 120.0f; P.mover_type = -1;
    fill_edges(ctx, 0.0f, cfg->tank_w);
    P.gx0 = P.gx0_new; P.wx = P.wx_new;
    int rc = push_params(ctx);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());       // the memsets above ran on the legacy stream, ctx->stream does not wait for it
    ctx->stage = ST_READY;
    ctx->xperiod = 1; ctx->since_x = 0; ctx->force_x = true; ctx->cur_x = true;
    return SPH_OK;
}

extern "C" void sph_destroy(sph_ctx *ctx)
{
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    for (int m = 0; m < 8; m++) if (ctx->graph_ready[m]) cudaGraphExecDestroy(ctx->graph[m]);
    cudaFree(ctx->coupling); cudaFree(ctx->dopt);
#if SPH_RELAX_PD4
    cudaFree(ctx->pd);
#endif
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (int k = 0; k < 2; k++) {
        if (ctx->feed[k].packed) cudaEventDestroy(ctx->feed[k].packed);
        if (ctx->feed[k].copied) cudaEventDestroy(ctx->feed[k].copied);
    }
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    cudaFree(ctx->feed_cnt_dev);
    if (ctx->feed_cnt_host) cudaFreeHost(ctx->feed_cnt_host);
    for (int i = 0; i < 4; i++) cudaFree(ctx->P[i]);
    for (int i = 0; i < 3; i++) cudaFree(ctx->Q[i]);
    for (int i = 0; i < 2; i++) cudaFree(ctx->U[i]);
    cudaFree(ctx->dens); cudaFree(ctx->nmask); cudaFree(ctx->cnt); cudaFree(ctx->cell_start); cudaFree(ctx->t_key);
    cudaFree(ctx->t_slot); cudaFree(ctx->ord_src); cudaFree(ctx->ord_key); cudaFree(ctx->ord_uid); cudaFree(ctx->coords); cudaFree(ctx->coords1);
    cudaFree(ctx->pv); cudaFree(ctx->tile_total); cudaFree(ctx->counters); cudaFree(ctx->dp);
    for (int s = 0; s < 2; s++) { cudaFree(ctx->send[s]); cudaFree(ctx->recv[s]); }
    for (int s = 0; s < 2; s++) if (ctx->peer[s]) cudaIpcCloseMemHandle(ctx->peer[s]);
    cudaFree(ctx->xchg); cudaFree(ctx->xt);
    if (ctx->stage_dp) cudaFreeHost(ctx->stage_dp);
    if (ctx->stage_free) cudaEventDestroy(ctx->stage_free);
    cudaFree(ctx->snap.P); cudaFree(ctx->snap.Q); cudaFree(ctx->snap.U); cudaFree(ctx->snap.cell_start);
    cudaFree(ctx->snap.ord_key); cudaFree(ctx->snap.counters);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *sph_last_error(const sph_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" long long sph_launch_count(const sph_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int sph_synchronize(sph_ctx *ctx)
{
    CK(cudaStreamSynchronize(ctx->stream));
    return SPH_OK;
}

static int apply_params(sph_ctx *ctx, const sph_tunable *t)
{
    ctx->tun = *t;
    fill_phys(ctx->hp, *t);
    fill_edges(ctx, t->node_start_x, t->node_end_x);
    return push_params(ctx);
}

extern "C" int sph_set_params(sph_ctx *ctx, const sph_tunable *t)
{
    if (!ctx || !t) return SPH_ERR_ARG;
    ctx->force_x = true;      // physics or edges changed outside the queue: the ghosts' validity budget starts over
    if (ctx->cfg.nranks > 1 && ctx->stage != ST_READY) {
        // Inside a step of a slab the window of grid columns is in use: the producing kernel has binned its particles
        // for the edges in force when it ran, and the sort that follows must use the same ones (edges moved here would
        // silently misplace or lose particles).  The physics takes effect now, as for a single slab; the edges are
        // queued and land with the next sph_advect, which is where the reference's scatter puts them (fluid.c:293-310).
        ctx->tun = *t;
        ctx->tun.node_start_x = ctx->hp.edge_start; ctx->tun.node_end_x = ctx->hp.edge_end;
        fill_phys(ctx->hp, *t);
        ctx->queued = *t;
        ctx->have_queued = true;
        return push_params(ctx);
    }
    return apply_params(ctx, t);
}

extern "C" int sph_queue_params(sph_ctx *ctx, const sph_tunable *t)
{
    if (!ctx || !t) return SPH_ERR_ARG;
    ctx->queued = *t;
    ctx->have_queued = true;
    return SPH_OK;
}

extern "C" int sph_set_viscosity_stabilisation(sph_ctx *ctx, float gamma, float min_dt_sigma)
{
    if (!ctx || !(gamma >= 0.0f) || !(min_dt_sigma >= 0.0f)) return SPH_ERR_ARG;
    ctx->visc_gamma = gamma;
    ctx->visc_min_dt_sigma = min_dt_sigma;
    ctx->force_x = true;
    DevOptions o;
    o.visc_gamma = gamma;
    CK(cudaMemcpyAsync(ctx->dopt, &o, sizeof o, cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before returning
    return SPH_OK;
}

extern "C" int sph_set_edges(sph_ctx *ctx, float s, float e)
{
    if (!ctx) return SPH_ERR_ARG;
    if (ctx->cfg.nranks > 1 && ctx->stage != ST_READY)
        return fail(ctx, SPH_ERR_STATE, "sph_set_edges: a slab's edges can only move at a step boundary (or through sph_queue_params)");
    ctx->tun.node_start_x = s; ctx->tun.node_end_x = e;
    ctx->force_x = true;
    fill_edges(ctx, s, e);
    return push_params(ctx);
}

extern "C" int sph_set_neighbors(sph_ctx *ctx, int has_left, int has_right)
{
    if (!ctx) return SPH_ERR_ARG;
    ctx->hp.has_left = has_left; ctx->hp.has_right = has_right;
    ctx->force_x = true;
    return push_params(ctx);
}

extern "C" int sph_exchange_buffers(sph_ctx *ctx, int which, void **sl, void **rl, void **sr, void **rr, size_t *bytes)
{
    if (!ctx) return SPH_ERR_ARG;
    if (sl) *sl = ctx->send[0];
    if (rl) *rl = ctx->recv[0];
    if (sr) *sr = ctx->send[1];
    if (rr) *rr = ctx->recv[1];
    if (bytes) *bytes = which == 0 ? msg_bytes_full(ctx->cfg.msg_capacity) : msg_bytes_halo1(ctx->cfg.msg_capacity);
    return SPH_OK;
}

// The exchange for a host whose transport moves HOST memory (plain MPI_Sendrecv, sockets): the message buffers are
// staged through pinned memory around two calls of the host's sendrecv, ordered like the reference's own pair of
// MPI_Sendrecv (communication.c:158-161, :340-344: to the right / from the left, then to the left / from the right), so a
// blocking transport cannot deadlock.  An absent neighbour is (NULL, 0), the reference's MPI_PROC_NULL.
extern "C" int sph_exchange_via_host(sph_ctx *ctx, int which, sph_sendrecv_fn fn, void *user)
{
    if (!ctx || !fn || (which != 0 && which != 1)) return SPH_ERR_ARG;
    if (ctx->cfg.nranks <= 1) return SPH_OK;
    if (which == 0 && ctx->stage == ST_ADVECTED && !ctx->cur_x) return SPH_OK;      // not an exchange step (sph_set_exchange_period)
    const size_t full = msg_bytes_full(ctx->cfg.msg_capacity);
    const size_t nb = which == 0 ? full : msg_bytes_halo1(ctx->cfg.msg_capacity);
    if (!ctx->stage_host) CK(cudaMallocHost(&ctx->stage_host, 4 * full));
    unsigned char *hs[2] = {ctx->stage_host, ctx->stage_host + full}, *hr[2] = {ctx->stage_host + 2 * full, ctx->stage_host + 3 * full};
    const int present[2] = {ctx->hp.has_left, ctx->hp.has_right};
    for (int s = 0; s < 2; s++)
        if (present[s]) CK(cudaMemcpyAsync(hs[s], ctx->send[s], nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // to the right / from the left, then to the left / from the right
    fn(present[1] ? hs[1] : nullptr, present[1] ? nb : 0, 1, present[0] ? hr[0] : nullptr, present[0] ? nb : 0, 0, user);
    fn(present[0] ? hs[0] : nullptr, present[0] ? nb : 0, 0, present[1] ? hr[1] : nullptr, present[1] ? nb : 0, 1, user);
    for (int s = 0; s < 2; s++)
        if (present[s]) CK(cudaMemcpyAsync(ctx->recv[s], hr[s], nb, cudaMemcpyHostToDevice, ctx->stream));
    return SPH_OK;
}

// how often neighbours meet per step in this build: 2 (after the prediction and after the relaxation), or 1 in the
// one-exchange build, where a multi-rank driver skips the second transfer
extern "C" int sph_exchanges_per_step(void) { return SPH_ONE_EXCHANGE ? 1 : 2; }        // the build's default
extern "C" int sph_ctx_exchanges_per_step(const sph_ctx *ctx) { return ctx && ctx->one_x ? 1 : 2; }


// Where the meetings' time goes (peer-memory transport): microseconds spent by the exchange kernel sending its
// messages, waiting for the neighbours' and unpacking them, summed over `*meetings` exchanges since the last reset
// (block 0's view; the wait is the sum of the protocol's flight time and of how much later the neighbour arrived).
extern "C" int sph_get_exchange_times(sph_ctx *ctx, double us[3], int *meetings, int reset)
{
    if (!ctx || !us || !meetings) return SPH_ERR_ARG;
    long long h[4];
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(h, ctx->xt, sizeof h, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) us[k] = (double)h[k] * 1e-3;
    *meetings = (int)h[3];
    if (reset) CK(cudaMemset(ctx->xt, 0, sizeof h));
    return SPH_OK;
}

// ------------------------------------------------------------------------------------------
// peer-memory exchange: neighbours map each other's exchange block (cudaIpc) and the pack code in
// k_advect / k_relax stores outgoing records straight into it over NVLink
// ------------------------------------------------------------------------------------------
extern "C" int sph_p2p_local_handle(sph_ctx *ctx, void *handle64)
{
    if (!ctx || !handle64) return SPH_ERR_ARG;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->xchg));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return SPH_OK;
}

extern "C" int sph_p2p_connect(sph_ctx *ctx, const void *left_handle64, const void *right_handle64)
{
    if (!ctx) return SPH_ERR_ARG;
    const void *hs[2] = {left_handle64, right_handle64};
    for (int s = 0; s < 2; s++) {
        if (!hs[s]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[s], 64);
        CK(cudaIpcOpenMemHandle(&ctx->peer[s], h, cudaIpcMemLazyEnablePeerAccess));
        ctx->hp.remote_base[s] = (unsigned long long)ctx->peer[s];
    }
    ctx->hp.p2p = 1;
    for (int m = 0; m < 8; m++) if (ctx->graph_ready[m]) { cudaGraphExecDestroy(ctx->graph[m]); ctx->graph_ready[m] = false; }
    ctx->force_x = true;
    return push_params(ctx);
}

// ------------------------------------------------------------------------------------------
// stage launches.  Buffer roles (period-1 rotation, so one captured graph is a whole step):
//   READY    : pos P0, vel Q0, uid U0 (cell-sorted)
//   advect   : reads P0,Q0,U0            writes predicted pos P1 (source order)
//   sort 1   : src (P1, prev=P0, U0)     dst (P2, prev Q1, U1)
//   density  : reads P2                  writes dens
//   relax    : reads P2,Q1,U1,dens       writes relaxed pos P3, vel Q2 (source order)
//   sort 2   : src (P3, Q2, U1)          dst (P0, Q0, U0)
// ------------------------------------------------------------------------------------------
static int launch_sort(sph_ctx *ctx, int which, bool with_unpack = true, bool refresh = false)
{
    float2 *sp = which == 0 ? ctx->P[1] : ctx->P[3];
    float2 *sq = which == 0 ? ctx->P[0] : ctx->Q[2];
    uint32_t *su = which == 0 ? ctx->U[0] : ctx->U[1];
    float2 *dp = which == 0 ? ctx->P[2] : ctx->P[0];
    float2 *dq = which == 0 ? ctx->Q[1] : ctx->Q[0];
    uint32_t *du = which == 0 ? ctx->U[1] : ctx->U[0];
    // (one-exchange build: nothing arrives after the relaxation, the ghosts were relaxed here)
    // (refresh: sph_refresh_ghosts always brings its ghosts in the which = 1 format)
    if (ctx->cfg.nranks > 1 && with_unpack && (refresh || (!(ctx->one_x && which == 1) && ctx->cur_x))) {
        SPH_LAUNCH(k_unpack, ctx->unpack_grid, ctx->stream)(ctx->dp, ctx->counters, which, ctx->send[0], ctx->send[1],
                                                             ctx->recv[0], ctx->recv[1],
                                                             sp, sq, su, ctx->cnt, ctx->t_key, ctx->t_slot, ctx->xt, ctx->tile_total);
        ctx->launches++;
    }
    // grids sized for the widest window (tile loops inside): a captured graph survives moving slab edges
    const int sgrid = std::max(1, std::min(ctx->scan_grid, SPH_GRID_MULT_SCAN * 148));
#if !SPH_TILE_ATOMICS
    SPH_LAUNCH(k_scan_totals, sgrid, ctx->stream)(ctx->dp, ctx->counters, ctx->cnt, ctx->tile_total);
    ctx->launches++;
#endif
#if SPH_SCAN_FAST && SPH_TILE_ATOMICS
#define SPH_K_SCAN k_scan_apply_fast
#else
#define SPH_K_SCAN k_scan_apply
#endif
    SPH_LAUNCH(SPH_K_SCAN, sgrid, ctx->stream)(ctx->dp, ctx->counters, ctx->cnt, ctx->cell_start, ctx->tile_total,
                                                         ctx->cfg.nranks > 1 ? ctx->send[0] : nullptr,
                                                         ctx->cfg.nranks > 1 ? ctx->send[1] : nullptr,
                                                         (which == 1 && with_unpack && (refresh || ctx->cur_x)) ? 1 : 0);
#if SPH_SORT_SRC
    SPH_LAUNCH(k_scatter_uid, ctx->sort_grid, ctx->stream)(ctx->counters, ctx->cell_start, ctx->t_key, ctx->t_slot, su,
                                                              ctx->ord_uid, ctx->tile_total, ctx->ntiles_max);
    SPH_LAUNCH(k_reorder_src, ctx->sort_grid, ctx->stream)(ctx->dp, ctx->counters, ctx->cell_start, ctx->t_key,
                                                              ctx->ord_uid, su, sp, sq, dp, dq, du, ctx->ord_key
#if SPH_ADVECT_PV4
                                                              , which == 1 ? ctx->pv : nullptr
#endif
                                                              );
#else
    SPH_LAUNCH(k_scatter, ctx->sort_grid, ctx->stream)(ctx->counters, ctx->cell_start, ctx->t_key, ctx->t_slot, su,
                                                          ctx->ord_uid, ctx->ord_src, ctx->ord_key, ctx->tile_total, ctx->ntiles_max);
    SPH_LAUNCH(k_reorder, ctx->sort_grid, ctx->stream)(ctx->dp, ctx->counters, ctx->cell_start, ctx->ord_key,
                                                          ctx->ord_uid, ctx->ord_src, sp, sq, dp, dq, du);
#endif
    ctx->launches += 3;
    ctx->hp.gx0 = ctx->hp.gx0_new;     // the scan kernel did the same on the device
    ctx->hp.wx = ctx->hp.wx_new;
    CK(cudaGetLastError());
    return SPH_OK;
}

// The stabilised viscosity gather runs for the parameter blocks that need it: gamma > 0 and a per-pair
// coefficient dt*sigma at or above the caller's threshold (the goo preset: 0.83; the default fluid: 0.17).
static bool stabilised(const sph_ctx *ctx)
{
    return ctx->visc_gamma > 0.0f && ctx->hp.dt * ctx->hp.sigma >= ctx->visc_min_dt_sigma;
}

// ---- exchange period (one-exchange build) ------------------------------------------------------------------
// Between two exchanges a slab advances its ghosts itself, redundantly: a ghost comes out exactly as on its owner
// as long as everything within reach of it was itself exact, and every pair pass eats one h of the layer from the
// outside (k_advect, k_density, k_relax: 3 h per step; 4 h with the stabilised viscosity gather, whose coupling
// sums are one more pair pass), plus half an h per step for the motion of the particles themselves.  So a layer
// of `unit * E` h carries E steps, unit = 3.5 (4.5 stabilised): E = 1 is the one-exchange build of round 1.
static float layer_unit(const sph_ctx *ctx, float dt, float sigma)
{
    return (ctx->visc_gamma > 0.0f && dt * sigma >= ctx->visc_min_dt_sigma) ? 4.5f : 3.5f;
}

// the period the ghost layer of this context can carry with the parameter block that will be in force
static int effective_period(const sph_ctx *ctx, const sph_tunable *coming)
{
    const float unit = layer_unit(ctx, coming ? coming->time_step : ctx->hp.dt, coming ? coming->sigma : ctx->hp.sigma);
    const int fits = (int)floorf(ctx->cfg.halo_width / unit + 1e-3f);
    return std::max(1, std::min(ctx->xperiod, fits));
}

// Decide whether the step that starts now is an exchange step, and set the width of the layer it sends.
// Every rank takes the same decision from the same call sequence (the reference's parameter scatter is collective).
static int begin_step(sph_ctx *ctx)
{
    if (ctx->cfg.nranks <= 1) { ctx->cur_x = false; return SPH_OK; }
    if (!ctx->one_x) { ctx->cur_x = true; return SPH_OK; }
    const sph_tunable *coming = ctx->have_queued ? &ctx->queued : nullptr;
    const int period = effective_period(ctx, coming);
    // a queued block lands between this step's prediction and its exchange (fluid.c:279-310): meeting in that same
    // step means the whole coming period runs under one block, whose viscosity decides the layer it needs
    ctx->cur_x = ctx->force_x || ctx->have_queued || ctx->since_x + 1 >= period;
    if (ctx->cur_x) {
        const float unit = layer_unit(ctx, coming ? coming->time_step : ctx->hp.dt, coming ? coming->sigma : ctx->hp.sigma);
        // period 1 keeps the layer the caller configured (round 1's one-exchange build); longer periods send what they need
        const float w = (ctx->xperiod <= 1 ? ctx->cfg.halo_width : unit * (float)period) * ctx->cfg.h;
        if (w != ctx->hp.halo_w) { ctx->hp.halo_w = w; int rc = push_params(ctx); if (rc) return rc; }
    }
    return SPH_OK;
}

static void end_step(sph_ctx *ctx)
{
    if (ctx->cur_x) { ctx->since_x = 0; ctx->force_x = false; }
    else ctx->since_x++;
}

// One-exchange build only: neighbours meet every `period` steps instead of every step (see begin_step above).
extern "C" int sph_set_exchange_period(sph_ctx *ctx, int period)
{
    if (!ctx || period < 1) return SPH_ERR_ARG;
    if (period > 1 && !ctx->one_x)
        return fail(ctx, SPH_ERR_STATE, "sph_set_exchange_period: this context exchanges twice per step (create it with exchanges_per_step = 1)");
    ctx->xperiod = period;
    ctx->force_x = true;
    return SPH_OK;
}

// 1 if the step in progress (after sph_advect) exchanges -- or, at a step boundary, if the coming step would
extern "C" int sph_exchange_due(sph_ctx *ctx)
{
    if (!ctx || ctx->cfg.nranks <= 1) return 0;
    if (ctx->stage != ST_READY) return ctx->cur_x ? 1 : 0;
    if (!ctx->one_x) return 1;
    return (ctx->force_x || ctx->have_queued ||
            ctx->since_x + 1 >= effective_period(ctx, ctx->have_queued ? &ctx->queued : nullptr)) ? 1 : 0;
}


#if SPH_ADVECT_PV4
#define SPH_PV4_ARG , ctx->pv
#else
#define SPH_PV4_ARG
#endif
// (function pointers: a template argument list with a comma cannot pass through the SPH_LAUNCH macro)
static constexpr auto k_advect_plain = k_advect<false>, k_advect_stab = k_advect<true>;
static constexpr auto k_advect_plain_hold = k_advect<false, true>, k_advect_stab_hold = k_advect<true, true>;

static int launch_advect(sph_ctx *ctx)
{
    // a slab's step between two exchanges: the instantiation that keeps waiting emigrants resident (k_advect, HOLD)
    const bool hold = ctx->cfg.nranks > 1 && !ctx->cur_x;
    if (stabilised(ctx)) {
        SPH_LAUNCH(k_coupling, ctx->grid_advect, ctx->stream)(ctx->dp, ctx->counters, ctx->P[0], ctx->Q[0], ctx->cell_start, ctx->coupling, ctx->ord_key);
        const auto advect_stab = hold ? k_advect_stab_hold : k_advect_stab;
        SPH_LAUNCH(advect_stab, ctx->grid_advect, ctx->stream)(ctx->dp, ctx->counters, ctx->P[0], ctx->Q[0], ctx->U[0],
                                                            ctx->cell_start, ctx->P[1], ctx->cnt, ctx->t_key, ctx->t_slot,
                                                            ctx->send[0], ctx->send[1], ctx->coupling, ctx->dopt, ctx->cur_x ? 1 : 0, ctx->ord_key, ctx->tile_total SPH_PV4_ARG);
        ctx->launches += 2;
        CK(cudaGetLastError());
        return SPH_OK;
    }
    const auto advect_plain = hold ? k_advect_plain_hold : k_advect_plain;
    SPH_LAUNCH(advect_plain, ctx->grid_advect, ctx->stream)(ctx->dp, ctx->counters, ctx->P[0], ctx->Q[0], ctx->U[0],
                                                         ctx->cell_start, ctx->P[1], ctx->cnt, ctx->t_key, ctx->t_slot,
                                                         ctx->send[0], ctx->send[1], nullptr, nullptr, ctx->cur_x ? 1 : 0, ctx->ord_key, ctx->tile_total SPH_PV4_ARG);
    ctx->launches++;
    CK(cudaGetLastError());
    return SPH_OK;
}

static int launch_density(sph_ctx *ctx)
{
    ctx->dens_seen = true;
    SPH_LAUNCH(k_density, ctx->grid_density, ctx->stream)(ctx->dp, ctx->counters, ctx->P[2], ctx->cell_start, ctx->dens, ctx->nmask, ctx->ord_key
#if SPH_RELAX_PD4
                                                   , ctx->pd
#endif
                                                   );
    ctx->launches++;
    CK(cudaGetLastError());
    return SPH_OK;
}

static int launch_relax(sph_ctx *ctx)
{
    SPH_LAUNCH(k_relax, ctx->grid_relax, ctx->stream)(ctx->dp, ctx->counters, ctx->P[2], ctx->Q[1], ctx->U[1], ctx->dens,
                                                        ctx->cell_start, ctx->nmask, ctx->P[3], ctx->Q[2], ctx->cnt, ctx->t_key,
                                                        ctx->t_slot, ctx->send[0], ctx->send[1], ctx->ord_key, ctx->tile_total
#if SPH_RELAX_PD4
                                                        , ctx->pd
#endif
                                                        );
    ctx->launches++;
    CK(cudaGetLastError());
    return SPH_OK;
}

// One-exchange build: an interior slab narrower than the ghost layer would leave its neighbours' layers incomplete
// (a rank only sends its own particles).  The reference's balancer keeps slabs >= 2 h wide, which covers the 2 h
// layer of the two-exchange build but not this one's.
static int check_layer(sph_ctx *ctx)
{
    if (ctx->one_x && ctx->hp.has_left && ctx->hp.has_right && ctx->hp.edge_end > ctx->hp.edge_start &&
        ctx->hp.edge_end - ctx->hp.edge_start < ctx->hp.halo_w)
        return fail(ctx, SPH_ERR_STATE, "one-exchange mode: interior slab narrower than the ghost layer");
    return SPH_OK;
}

extern "C" int sph_advect(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    if (ctx->stage != ST_READY) return fail(ctx, SPH_ERR_STATE, "sph_advect: state is not at a step boundary");
    int rc;
    if ((rc = begin_step(ctx))) return rc;
    if (ctx->have_queued) {
        // the scatter from the render rank lands between prediction and migration (fluid.c:279-310):
        // new slab edges first (only the classification at the end of the kernel reads them) ...
        fill_edges(ctx, ctx->queued.node_start_x, ctx->queued.node_end_x);
        if ((rc = push_params(ctx))) return rc;
    }
    if ((rc = check_layer(ctx))) return rc;
    if ((rc = launch_advect(ctx))) return rc;
    if (ctx->have_queued) {
        // ... then everything else, for the stages after the prediction
        ctx->have_queued = false;
        if ((rc = apply_params(ctx, &ctx->queued))) return rc;
    }
    ctx->stage = ST_ADVECTED;
    return SPH_OK;
}

extern "C" int sph_sort(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    int rc;
    if (ctx->stage == ST_ADVECTED) { if ((rc = launch_sort(ctx, 0))) return rc; ctx->stage = ST_SORTED1; }
    else if (ctx->stage == ST_RELAXED) { if ((rc = launch_sort(ctx, 1))) return rc; ctx->stage = ST_READY; ctx->steps++; end_step(ctx); }
    else if (ctx->stage == ST_REQUEUED) { if ((rc = launch_sort(ctx, 1, true, true))) return rc; ctx->stage = ST_READY; }
    else return fail(ctx, SPH_ERR_STATE, "sph_sort: nothing to sort");
    return SPH_OK;
}

// Ghosts for a slab that has just been uploaded (a restart from a moving snapshot): see include/sph_b200.h
extern "C" int sph_refresh_ghosts(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    if (ctx->stage != ST_READY) return fail(ctx, SPH_ERR_STATE, "sph_refresh_ghosts: state is not at a step boundary");
    if (ctx->cfg.nranks <= 1) return SPH_OK;
    SPH_LAUNCH(k_requeue, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, ctx->P[0], ctx->Q[0], ctx->U[0], ctx->P[3], ctx->Q[2],
                                                  ctx->U[1], ctx->cnt, ctx->t_key, ctx->t_slot, ctx->send[0], ctx->send[1], ctx->tile_total);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->stage = ST_REQUEUED;
    return SPH_OK;
}

extern "C" int sph_density(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    if (ctx->stage != ST_SORTED1) return fail(ctx, SPH_ERR_STATE, "sph_density: call after the first sort");
    int rc = launch_density(ctx);
    if (rc) return rc;
    ctx->stage = ST_DENSITY;
    return SPH_OK;
}

extern "C" int sph_relax(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    if (ctx->stage != ST_DENSITY) return fail(ctx, SPH_ERR_STATE, "sph_relax: call after sph_density");
    int rc = launch_relax(ctx);
    if (rc) return rc;
    ctx->stage = ST_RELAXED;
    return SPH_OK;
}

static int launch_step(sph_ctx *ctx, bool queued)
{
    int rc;
    // a step in which a queued parameter block lands copies it in from pinned staging, in the two halves of
    // sph_advect: the new edges before the prediction, everything after it (fluid.c:279-310)
    if (queued) CK(cudaMemcpyAsync(ctx->dp, &ctx->stage_dp[0], sizeof(DevParams), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_advect(ctx))) return rc;
    if (queued) CK(cudaMemcpyAsync(ctx->dp, &ctx->stage_dp[1], sizeof(DevParams), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_sort(ctx, 0))) return rc;
    if ((rc = launch_density(ctx))) return rc;
    if ((rc = launch_relax(ctx))) return rc;
    if ((rc = launch_sort(ctx, 1))) return rc;
    return SPH_OK;
}

extern "C" int sph_step(sph_ctx *ctx, int n)
{
    if (!ctx || n < 0) return SPH_ERR_ARG;
    if (ctx->cfg.nranks != 1 && !ctx->hp.p2p)
        return fail(ctx, SPH_ERR_STATE, "sph_step: slabs need sph_p2p_connect (or step stage by stage around a transport)");
    if (ctx->stage != ST_READY) return fail(ctx, SPH_ERR_STATE, "sph_step: state is not at a step boundary");
    int rc;
    for (int s = 0; s < n; s++) {
        const bool queued = ctx->have_queued;
        if ((rc = begin_step(ctx))) return rc;
        if (queued) {
            // Parameter change inside this step (round 1 ran such a step stage by stage, eleven separate launches once
            // per frame): the captured step reads the two blocks from pinned staging when it RUNS.  The staging may
            // only be rewritten once the previous queued step has read it.
            if (ctx->stage_busy) CK(cudaEventSynchronize(ctx->stage_free));
            fill_edges(ctx, ctx->queued.node_start_x, ctx->queued.node_end_x);      // the prediction runs on the old physics
            ctx->stage_dp[0] = ctx->hp;
        }
        if ((rc = check_layer(ctx))) return rc;
        // one captured step per variant: viscosity gather (plain / stabilised) x exchange step or not x queued block or not
        const int m = (stabilised(ctx) ? 1 : 0) + (ctx->cur_x ? 2 : 0) + (queued ? 4 : 0);
        if (!ctx->graph_ready[m]) {
            cudaGraph_t g;
            long long before = ctx->launches;
            CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            rc = launch_step(ctx, queued);
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            ctx->launch_per_graph[m] = (int)(ctx->launches - before);
            ctx->launches = before;
            if (rc) return rc;
            CK(e);
            CK(cudaGraphInstantiate(&ctx->graph[m], g, 0));
            CK(cudaGraphDestroy(g));
            ctx->graph_ready[m] = true;
        }
        if (queued) {
            ctx->have_queued = false;
            ctx->tun = ctx->queued;
            fill_phys(ctx->hp, ctx->queued);
            ctx->stage_dp[1] = ctx->hp;
        }
        CK(cudaGraphLaunch(ctx->graph[m], ctx->stream));
        if (queued) { CK(cudaEventRecord(ctx->stage_free, ctx->stream)); ctx->stage_busy = true; }
        ctx->hp.gx0 = ctx->hp.gx0_new;          // the step's first scan did the same on the device
        ctx->hp.wx = ctx->hp.wx_new;
        ctx->launches += ctx->launch_per_graph[m];
        ctx->steps++;
        end_step(ctx);
    }
    return SPH_OK;
}

// ------------------------------------------------------------------------------------------
// state transfer
// ------------------------------------------------------------------------------------------
// n particles sit in the sort-2 source arrays (P3 / Q2 / U1): reset the counters, bin and sort them
static int ingest(sph_ctx *ctx, int n)
{
    int zero[CN_COUNT] = {0};
    zero[CN_NTOT] = n;
    CK(cudaMemcpyAsync(ctx->counters, zero, sizeof zero, cudaMemcpyHostToDevice, ctx->stream));
    const size_t ncell_max = (size_t)ctx->size_x * ctx->size_y * SPH_CELL_DIV * SPH_CELL_DIV;
    CK(cudaMemsetAsync(ctx->cnt, 0, (ncell_max + 1) * sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->xchg, 0, SPH_XCHG_HDR, ctx->stream));      // message sequence numbers restart
    for (int s = 0; s < 2; s++) CK(cudaMemsetAsync(ctx->send[s], 0, 16, ctx->stream));
    int rc = push_params(ctx);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->tile_total, 0, (size_t)ctx->ntiles_max * sizeof(int), ctx->stream));
    SPH_LAUNCH(k_bin_upload, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, ctx->P[3], ctx->cnt, ctx->t_key, ctx->t_slot, ctx->tile_total);
    ctx->launches++;
    // no neighbour messages belong to an upload: skip the unpack kernel
    if ((rc = launch_sort(ctx, 1, false))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stage = ST_READY;
    ctx->dens_seen = false;
    ctx->n_uploaded = n;
    ctx->force_x = true;          // a fresh upload has no ghosts
    return SPH_OK;
}

// ---- snapshot in device memory -------------------------------------------------------------------------------
// The sorted state at a step boundary (positions, velocities, uids, ghosts included; cell table; counters; the
// parameter block with its window) copied aside, and put back later: a benchmark can time the SAME steps again and
// again (the dam-break keeps compressing, so blocks timed one after the other are not the same work), a host can
// rewind.  On a slab every rank restores together; message sequence numbers keep counting, and the first step after
// a restore is an exchange step.
extern "C" int sph_state_save(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    if (ctx->stage != ST_READY) return fail(ctx, SPH_ERR_STATE, "sph_state_save: state is not at a step boundary");
    if (ctx->have_queued) return fail(ctx, SPH_ERR_STATE, "sph_state_save: a queued parameter block is pending");
    const size_t cap = (size_t)ctx->cfg.capacity;
    const size_t ncell = (size_t)ctx->size_x * ctx->size_y * SPH_CELL_DIV * SPH_CELL_DIV + 1;
    if (!ctx->snap.P) {
        CK(cudaMalloc(&ctx->snap.P, cap * sizeof(float2)));
        CK(cudaMalloc(&ctx->snap.Q, cap * sizeof(float2)));
        CK(cudaMalloc(&ctx->snap.U, cap * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->snap.ord_key, cap * sizeof(int)));
        CK(cudaMalloc(&ctx->snap.cell_start, ncell * sizeof(int)));
        CK(cudaMalloc(&ctx->snap.counters, CN_COUNT * sizeof(int)));
    }
    const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(ctx->snap.P, ctx->P[0], cap * sizeof(float2), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->snap.Q, ctx->Q[0], cap * sizeof(float2), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->snap.U, ctx->U[0], cap * sizeof(uint32_t), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->snap.ord_key, ctx->ord_key, cap * sizeof(int), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->snap.cell_start, ctx->cell_start, ncell * sizeof(int), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->snap.counters, ctx->counters, CN_COUNT * sizeof(int), dd, ctx->stream));
    ctx->snap.hp = ctx->hp; ctx->snap.tun = ctx->tun; ctx->snap.steps = ctx->steps;
    ctx->snap.valid = true;
    return SPH_OK;
}

extern "C" int sph_state_restore(sph_ctx *ctx)
{
    if (!ctx) return SPH_ERR_ARG;
    if (!ctx->snap.valid) return fail(ctx, SPH_ERR_STATE, "sph_state_restore: nothing saved");
    if (ctx->stage != ST_READY) return fail(ctx, SPH_ERR_STATE, "sph_state_restore: state is not at a step boundary");
    const size_t cap = (size_t)ctx->cfg.capacity;
    const size_t ncell = (size_t)ctx->size_x * ctx->size_y * SPH_CELL_DIV * SPH_CELL_DIV + 1;
    const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(ctx->P[0], ctx->snap.P, cap * sizeof(float2), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->Q[0], ctx->snap.Q, cap * sizeof(float2), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->U[0], ctx->snap.U, cap * sizeof(uint32_t), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->ord_key, ctx->snap.ord_key, cap * sizeof(int), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->cell_start, ctx->snap.cell_start, ncell * sizeof(int), dd, ctx->stream));
    // counters: everything but the message sequence number (the neighbours' arrival flags only ever grow) and the
    // cumulative error counters
    CK(cudaMemcpyAsync(ctx->counters, ctx->snap.counters, CN_MAX_BUCKET * sizeof(int), dd, ctx->stream));
    CK(cudaMemcpyAsync(ctx->counters + CN_COST, ctx->snap.counters + CN_COST, sizeof(int), dd, ctx->stream));
    const DevParams now = ctx->hp;
    ctx->hp = ctx->snap.hp;
    ctx->hp.p2p = now.p2p; ctx->hp.xchg_base = now.xchg_base;
    ctx->hp.remote_base[0] = now.remote_base[0]; ctx->hp.remote_base[1] = now.remote_base[1];
    ctx->tun = ctx->snap.tun; ctx->steps = ctx->snap.steps;
    ctx->have_queued = false;
    ctx->force_x = true;
    for (int s = 0; s < 2; s++) CK(cudaMemsetAsync(ctx->send[s], 0, 16, ctx->stream));
#if SPH_ADVECT_PV4
    SPH_LAUNCH(k_interleave, ctx->sort_grid, ctx->stream)(ctx->counters, ctx->P[0], ctx->Q[0], ctx->pv);
#endif
    return push_params(ctx);
}

extern "C" int sph_upload(sph_ctx *ctx, const sph_particle *aos, const uint32_t *uid, int n)
{
    if (!ctx || (!aos && n > 0) || n < 0) return SPH_ERR_ARG;
    if (n > ctx->cfg.capacity) return fail(ctx, SPH_ERR_CAPACITY, "sph_upload: more particles than capacity");
    std::vector<float2> p(n), v(n);
    std::vector<uint32_t> u(n);
    for (int i = 0; i < n; i++) {
        p[i] = make_float2(aos[i].x, aos[i].y);
        v[i] = make_float2(aos[i].v_x, aos[i].v_y);
        u[i] = uid ? (uid[i] & SPH_UID_MASK) : (uint32_t)i;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    // enter the pipeline where sort 2 does: source arrays P3 / Q2 / U1
    CK(cudaMemcpyAsync(ctx->P[3], p.data(), n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->Q[2], v.data(), n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->U[1], u.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));       // the host vectors go out of scope
    return ingest(ctx, n);
}

extern "C" int sph_init_lattice(sph_ctx *ctx, float water_min_x, float water_min_y, float water_max_y, float spacing,
                                int start_col, int ncols, int total_cols)
{
    if (!ctx || spacing <= 0.0f || ncols < 0) return -SPH_ERR_ARG;
    const int rows = (int)floor((water_max_y - water_min_y) / spacing);           // geometry.c:35
    const long long n = (long long)rows * ncols;
    if (n > ctx->cfg.capacity) { fail(ctx, SPH_ERR_CAPACITY, "sph_init_lattice: more particles than capacity"); return -SPH_ERR_CAPACITY; }
    SPH_LAUNCH(k_init_lattice, ctx->grid, ctx->stream)(water_min_x, water_min_y, spacing, start_col, ncols, rows,
                                                               total_cols, ctx->P[3], ctx->Q[2], ctx->U[1]);
    ctx->launches++;
    int rc = ingest(ctx, (int)n);
    return rc ? -rc : (int)n;
}

static int read_counters(sph_ctx *ctx, int *c)
{
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(c, ctx->counters, CN_COUNT * sizeof(int), cudaMemcpyDeviceToHost));
    return SPH_OK;
}

// which sorted arrays hold the resident state at this stage
static int current_arrays(sph_ctx *ctx, float2 **pos, float2 **q, uint32_t **uid, bool *q_is_prev)
{
    if (ctx->stage == ST_READY) { *pos = ctx->P[0]; *q = ctx->Q[0]; *uid = ctx->U[0]; *q_is_prev = false; return SPH_OK; }
    if (ctx->stage == ST_SORTED1 || ctx->stage == ST_DENSITY) { *pos = ctx->P[2]; *q = ctx->Q[1]; *uid = ctx->U[1]; *q_is_prev = true; return SPH_OK; }
    return fail(ctx, SPH_ERR_STATE, "state is between a producing kernel and its sort: call sph_sort first");
}

extern "C" int sph_download(sph_ctx *ctx, sph_particle *aos, uint32_t *uid_out, int order, int include_halo)
{
    if (!ctx || !aos) return -SPH_ERR_ARG;
    int c[CN_COUNT];
    if (read_counters(ctx, c)) return -SPH_ERR_CUDA;
    float2 *dpos, *dq; uint32_t *duid; bool q_is_prev;
    if (current_arrays(ctx, &dpos, &dq, &duid, &q_is_prev)) return -SPH_ERR_STATE;
    const int n = c[CN_NTOT];
    std::vector<float2> p(n), q(n), d(n);
    std::vector<uint32_t> u(n);
    auto cp = [&](void *dst, const void *src, size_t bytes) { return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
    if (n > 0) {
        if (cp(p.data(), dpos, n * sizeof(float2)) || cp(q.data(), dq, n * sizeof(float2)) ||
            cp(u.data(), duid, n * sizeof(uint32_t))) { fail(ctx, SPH_ERR_CUDA, "download copy failed"); return -SPH_ERR_CUDA; }
        if (ctx->stage == ST_DENSITY && cp(d.data(), ctx->dens, n * sizeof(float2))) { fail(ctx, SPH_ERR_CUDA, "download copy failed"); return -SPH_ERR_CUDA; }
    }
    std::vector<int> idx;
    idx.reserve(n);
    for (int i = 0; i < n; i++) if (include_halo || !(u[i] & SPH_HALO_BIT)) idx.push_back(i);
    if (order == SPH_ORDER_UID)
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (u[a] & SPH_UID_MASK) < (u[b] & SPH_UID_MASK); });
    for (size_t k = 0; k < idx.size(); k++) {
        const int i = idx[k];
        sph_particle &o = aos[k];
        memset(&o, 0, sizeof o);
        o.x = p[i].x; o.y = p[i].y;
        if (q_is_prev) { o.x_prev = q[i].x; o.y_prev = q[i].y; }
        else { o.x_prev = p[i].x; o.y_prev = p[i].y; o.v_x = q[i].x; o.v_y = q[i].y; }
        if (ctx->stage == ST_DENSITY) {
            o.density = d[i].x; o.density_near = d[i].y;
            o.pressure = ctx->tun.k * (d[i].x - ctx->tun.rest_density);     // fluid.c:563-564
            o.pressure_near = ctx->tun.k_near * d[i].y;
        }
        o.id = (int)k;
        if (uid_out) uid_out[k] = u[i];
    }
    return (int)idx.size();
}

extern "C" int sph_copy_n_local(sph_ctx *ctx, void *device_dst)
{
    if (!ctx || !device_dst) return SPH_ERR_ARG;
    CK(cudaMemcpyAsync(device_dst, ctx->counters + CN_NLOCAL, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    return SPH_OK;
}

extern "C" int sph_copy_load(sph_ctx *ctx, void *device_dst)
{
    if (!ctx || !device_dst) return SPH_ERR_ARG;
    CK(cudaMemcpyAsync(device_dst, ctx->counters + CN_NLOCAL, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync((int *)device_dst + 1, ctx->counters + CN_COST, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    return SPH_OK;
}

// {local particles, work estimate, this slab's own time since the last call [us], its waits over the same span [us]}
// as four ints into DEVICE memory, stream-ordered, without synchronising; restarts the two time accumulators
extern "C" int sph_copy_work(sph_ctx *ctx, void *device_dst)
{
    if (!ctx || !device_dst) return SPH_ERR_ARG;
    SPH_LAUNCH(k_pack_work, 1, ctx->stream)(ctx->counters, ctx->xt, (int *)device_dst);
    ctx->launches++;
    CK(cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_get_status(sph_ctx *ctx, sph_status *out)
{
    if (!ctx || !out) return SPH_ERR_ARG;
    int c[CN_COUNT];
    int rc = read_counters(ctx, c);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->n_local = c[CN_NLOCAL];
    out->n_halo = c[CN_NTOT] - c[CN_NLOCAL];
    {
        // exact statistics of the reference's buckets for the current sorted state (the hot path only keeps a
        // conservative per-sub-cell count in CN_BUCKET_OVER)
        int *dstat = ctx->counters + CN_SPARE0;          // two spare counters as scratch
        int hz[2] = {0, 0};
        CK(cudaMemcpyAsync(dstat, hz, sizeof hz, cudaMemcpyHostToDevice, ctx->stream));
        SPH_LAUNCH(k_bucket_stats, ctx->grid, ctx->stream)(ctx->dp, ctx->cell_start, dstat);
        ctx->launches++;
        CK(cudaMemcpyAsync(hz, dstat, sizeof hz, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        out->max_bucket = hz[0];
        out->bucket_overflow = std::max(hz[1], c[CN_BUCKET_OVER]);
    }
    out->neighbor_overflow = c[CN_NEIGH_OVER];
    // (a freshly uploaded state has not been through a density pass: nothing has primed the running count yet)
    if (c[CN_NEIGH_OVER] > 0 || !ctx->dens_seen) {
        // the hot path's count is conservative (full neighbour count, candidates past a row's mask taken as accepted) and
        // cumulative; where the state is sorted, count the reference's forward lists of the CURRENT state exactly
        float2 *dpos, *dq; uint32_t *duid; bool q_is_prev;
        if (ctx->stage == ST_READY || ctx->stage == ST_SORTED1 || ctx->stage == ST_DENSITY) {
            current_arrays(ctx, &dpos, &dq, &duid, &q_is_prev);
            int *dstat = ctx->counters + CN_SPARE0;
            int z = 0;
            CK(cudaMemcpyAsync(dstat, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
            SPH_LAUNCH(k_forward_overflow, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, dpos, duid, ctx->cell_start, dstat);
            ctx->launches++;
            CK(cudaMemcpyAsync(&z, dstat, sizeof z, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            out->neighbor_overflow = z;
        }
    }
    out->capacity_overflow = c[CN_CAP_OVER];
    out->msg_overflow = c[CN_MSG_OVER];
    out->exchange_timeouts = c[CN_TIMEOUT_MSG];
    if (c[CN_TIMEOUT_MSG])
        snprintf(ctx->err, sizeof ctx->err, "device-side waits timed out: %d neighbour messages never arrived", c[CN_TIMEOUT_MSG]);
    if (ctx->cfg.nranks > 1) {
        int hl[2] = {0, 0}, hr[2] = {0, 0};
        CK(cudaMemcpy(hl, ctx->send[0], sizeof hl, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hr, ctx->send[1], sizeof hr, cudaMemcpyDeviceToHost));
        out->migrated_left = hl[0]; out->migrated_right = hr[0];
    }
    out->steps = ctx->steps;
    return SPH_OK;
}

// ---- asynchronous coordinate feed --------------------------------------------------------------------------
// The reference's compute rank MPI_Isends its frame and goes on with the next one; it only waits for that send
// before it overwrites the buffer (fluid.c:283-287, :354-365).  Same here: sph_pack_coords_async packs on the
// compute stream, hands the copy to a second stream and returns a ticket; the caller launches the next frame's
// steps and then collects the ticket.  At most two frames are in flight (tickets 0 and 1 alternate); `xy` should
// be pinned memory and must stay untouched until the ticket is collected.
extern "C" int sph_pack_coords_async(sph_ctx *ctx, int16_t *xy, int cap)
{
    if (!ctx || !xy || cap < 0) return -SPH_ERR_ARG;
    float2 *dpos, *dq; uint32_t *duid; bool q_is_prev;
    if (current_arrays(ctx, &dpos, &dq, &duid, &q_is_prev)) return -SPH_ERR_STATE;
    const int k = (int)(ctx->feed_seq & 1);
    if (ctx->feed[k].pending) { fail(ctx, SPH_ERR_STATE, "sph_pack_coords_async: two frames in flight, collect one with sph_coords_wait"); return -SPH_ERR_STATE; }
    auto bail = [&](cudaError_t e) { snprintf(ctx->err, sizeof ctx->err, "pack_coords_async: %s", cudaGetErrorString(e)); return -SPH_ERR_CUDA; };
    cudaError_t e;
    // one device-side frame per ticket: the compute stream never waits for a copy.  (Frame f-2's copy out of this
    // buffer has drained: its ticket was collected, or this call would have been refused above.  Round 2's first build
    // had ONE buffer and made the pack of frame f wait for the copy of frame f-1 -- on 8 GPUs, whose 12 MB frames share
    // the host's memory, that wait was part of every frame.)
    short2 *dcoords = k == 0 ? ctx->coords : ctx->coords1;
    if ((e = cudaMemsetAsync(ctx->counters + CN_COORDS, 0, sizeof(int), ctx->stream))) return bail(e);
    SPH_LAUNCH(k_pack_coords, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, dpos, duid, dcoords, ctx->cfg.capacity);
    ctx->launches++;
    // the counters as they stand now (the next frame's sorts will rewrite them while the copy is still running)
    if ((e = cudaMemcpyAsync(ctx->feed_cnt_dev + k * CN_COUNT, ctx->counters, CN_COUNT * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream))) return bail(e);
    if ((e = cudaEventRecord(ctx->feed[k].packed, ctx->stream))) return bail(e);
    if ((e = cudaStreamWaitEvent(ctx->copy_stream, ctx->feed[k].packed, 0))) return bail(e);
    // how many entries to bring over is not known on the host without a synchronisation: a single slab keeps the
    // count of its upload; a slab among others copies the population of the last frame it collected plus an eighth
    // (plus a margin), and sph_coords_wait fetches the rest in the rare frame in which the slab grew by more than that
    // (until a frame has been collected the whole buffer travels)
    int entries = ctx->cfg.capacity;
    if (ctx->cfg.nranks == 1) entries = ctx->n_uploaded;
    else if (ctx->feed_last_n > 0)
        entries = (int)std::min<long long>(ctx->cfg.capacity, std::max<long long>(0, (long long)ctx->feed_last_n + ctx->feed_last_n / 8 + ctx->feed_margin));
    entries = std::min(cap, entries);
    if ((e = cudaMemcpyAsync(ctx->feed_cnt_host + k * CN_COUNT, ctx->feed_cnt_dev + k * CN_COUNT, CN_COUNT * sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream))) return bail(e);
    if (entries > 0 && (e = cudaMemcpyAsync(xy, dcoords, (size_t)entries * sizeof(short2), cudaMemcpyDeviceToHost, ctx->copy_stream))) return bail(e);
    if ((e = cudaEventRecord(ctx->feed[k].copied, ctx->copy_stream))) return bail(e);
    ctx->feed[k].entries = entries;
    ctx->feed[k].xy = xy;
    ctx->feed[k].cap = cap;
    ctx->feed[k].pending = true;
    ctx->feed_seq++;
    return k;
}

// Collect a ticket: blocks until that frame's coordinates are in `xy`; returns the number of particles (like
// sph_pack_coords: the count, of which min(count, cap) were written) or <0.
extern "C" int sph_coords_wait(sph_ctx *ctx, int ticket)
{
    if (!ctx || ticket < 0 || ticket > 1) return -SPH_ERR_ARG;
    if (!ctx->feed[ticket].pending) { fail(ctx, SPH_ERR_STATE, "sph_coords_wait: no such frame in flight"); return -SPH_ERR_STATE; }
    cudaError_t e = cudaEventSynchronize(ctx->feed[ticket].copied);
    ctx->feed[ticket].pending = false;
    if (e) { snprintf(ctx->err, sizeof ctx->err, "coords_wait: %s", cudaGetErrorString(e)); return -SPH_ERR_CUDA; }
    const int n = ctx->feed_cnt_host[ticket * CN_COUNT + CN_NLOCAL];
    const int want = std::min(n, ctx->feed[ticket].cap), have = ctx->feed[ticket].entries;
    if (want > have) {
        // the slab grew by more than the estimate allowed for: the rest of the frame is still in this ticket's device
        // buffer (nothing packs into it before the ticket is collected)
        const short2 *dcoords = ticket == 0 ? ctx->coords : ctx->coords1;
        e = cudaMemcpy(ctx->feed[ticket].xy + 2 * (size_t)have, dcoords + have, (size_t)(want - have) * sizeof(short2), cudaMemcpyDeviceToHost);
        if (e) { snprintf(ctx->err, sizeof ctx->err, "coords_wait (remainder): %s", cudaGetErrorString(e)); return -SPH_ERR_CUDA; }
        ctx->feed[ticket].entries = want;
    }
    ctx->feed_last_n = n;
    return n;
}

// entries of the frame of `ticket` that crossed to the host (after sph_coords_wait: what the asynchronous copy brought
// plus what the wait had to fetch); for the D2H accounting of a bench
extern "C" int sph_coords_copied(sph_ctx *ctx, int ticket)
{
    if (!ctx || ticket < 0 || ticket > 1) return -SPH_ERR_ARG;
    return ctx->feed[ticket].entries;
}

extern "C" int sph_pack_coords(sph_ctx *ctx, int16_t *xy, int cap)
{
    if (!ctx || !xy) return -SPH_ERR_ARG;
    float2 *dpos, *dq; uint32_t *duid; bool q_is_prev;
    if (current_arrays(ctx, &dpos, &dq, &duid, &q_is_prev)) return -SPH_ERR_STATE;
    auto bail = [&](cudaError_t e) { snprintf(ctx->err, sizeof ctx->err, "pack_coords: %s", cudaGetErrorString(e)); return -SPH_ERR_CUDA; };
    cudaError_t e;
    // frames of the asynchronous feed still draining read the same device buffer
    for (int k = 0; k < 2; k++)
        if (ctx->feed[k].pending && (e = cudaEventSynchronize(ctx->feed[k].copied))) return bail(e);
    if ((e = cudaMemsetAsync(ctx->counters + CN_COORDS, 0, sizeof(int), ctx->stream))) return bail(e);
    SPH_LAUNCH(k_pack_coords, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, dpos, duid, ctx->coords, ctx->cfg.capacity);
    ctx->launches++;
    int c[CN_COUNT];
    if ((e = cudaMemcpyAsync(c, ctx->counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream))) return bail(e);
    if ((e = cudaStreamSynchronize(ctx->stream))) return bail(e);
    const int n = c[CN_NLOCAL];
    const int m = std::min(n, cap);
    if (m > 0) {
        if ((e = cudaMemcpyAsync(xy, ctx->coords, (size_t)m * sizeof(short2), cudaMemcpyDeviceToHost, ctx->stream))) return bail(e);
        if ((e = cudaStreamSynchronize(ctx->stream))) return bail(e);
    }
    return n;
}

extern "C" int sph_run_frame(sph_ctx *ctx, const sph_tunable *t, int steps, int16_t *xy, int cap)
{
    if (!ctx || steps < 1) return -SPH_ERR_ARG;
    int rc;
    if ((rc = sph_step(ctx, steps - 1))) return -rc;
    if (t && (rc = sph_queue_params(ctx, t))) return -rc;
    if ((rc = sph_step(ctx, 1))) return -rc;
    if (xy) return sph_pack_coords(ctx, xy, cap);
    int c[CN_COUNT];
    if ((rc = read_counters(ctx, c))) return -rc;
    return c[CN_NLOCAL];
}

// sph_run_frame without the wait at its end: returns a ticket for sph_coords_wait
extern "C" int sph_run_frame_async(sph_ctx *ctx, const sph_tunable *t, int steps, int16_t *xy, int cap)
{
    if (!ctx || steps < 1 || !xy) return -SPH_ERR_ARG;
    int rc;
    if ((rc = sph_step(ctx, steps - 1))) return -rc;
    if (t && (rc = sph_queue_params(ctx, t))) return -rc;
    if ((rc = sph_step(ctx, 1))) return -rc;
    return sph_pack_coords_async(ctx, xy, cap);
}

extern "C" int sph_get_cells(sph_ctx *ctx, uint32_t *uid, uint32_t *cell, int cap)
{
    if (!ctx || !uid || !cell) return -SPH_ERR_ARG;
    float2 *dpos, *dq; uint32_t *duid; bool q_is_prev;
    if (current_arrays(ctx, &dpos, &dq, &duid, &q_is_prev)) return -SPH_ERR_STATE;
    int c[CN_COUNT];
    if (read_counters(ctx, c)) return -SPH_ERR_CUDA;
    const int n = c[CN_NTOT];
    uint32_t *dcell = (uint32_t *)ctx->ord_uid;      // scratch: free between sorts
    SPH_LAUNCH(k_export_cells, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, dpos, dcell);
    ctx->launches++;
    std::vector<uint32_t> hu(n), hc(n);
    if (n > 0) {
        if (cudaMemcpyAsync(hc.data(), dcell, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) ||
            cudaMemcpyAsync(hu.data(), duid, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) ||
            cudaStreamSynchronize(ctx->stream)) { fail(ctx, SPH_ERR_CUDA, "get_cells copy failed"); return -SPH_ERR_CUDA; }
    }
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (hu[i] & SPH_HALO_BIT) continue;
        if (m < cap) { uid[m] = hu[i]; cell[m] = hc[i]; }
        m++;
    }
    return m;
}

static long long export_pairs(sph_ctx *ctx, uint64_t *pairs, long long cap, uint32_t *uid, int *count, int count_cap)
{
    float2 *dpos, *dq; uint32_t *duid; bool q_is_prev;
    if (current_arrays(ctx, &dpos, &dq, &duid, &q_is_prev)) return -SPH_ERR_STATE;
    int c[CN_COUNT];
    if (read_counters(ctx, c)) return -SPH_ERR_CUDA;
    const int n = c[CN_NTOT];
    unsigned long long *dpairs = nullptr, *dn = nullptr;
    int *dfwd = nullptr;
    long long result = -SPH_ERR_CUDA;
    unsigned long long np = 0;
    do {
        if (cudaMalloc(&dn, sizeof(unsigned long long)) || cudaMemset(dn, 0, sizeof(unsigned long long))) break;
        if (pairs && cap > 0 && cudaMalloc(&dpairs, (size_t)cap * sizeof(unsigned long long))) break;
        if (count && cudaMalloc(&dfwd, (size_t)std::max(n, 1) * sizeof(int))) break;
        // count-only calls still need a non-null marker so the kernel counts pairs
        unsigned long long *pairs_arg = (pairs || !count) ? (dpairs ? dpairs : (unsigned long long *)dn) : nullptr;
        SPH_LAUNCH(k_export_pairs, ctx->grid, ctx->stream)(ctx->dp, ctx->counters, dpos, duid, ctx->cell_start,
                                                                   pairs_arg, dpairs ? (unsigned long long)cap : 0ull, dn, dfwd);
        ctx->launches++;
        if (cudaStreamSynchronize(ctx->stream)) break;
        if (cudaMemcpy(&np, dn, sizeof np, cudaMemcpyDeviceToHost)) break;
        if (dpairs && cudaMemcpy(pairs, dpairs, (size_t)std::min<unsigned long long>(np, (unsigned long long)cap) * sizeof(uint64_t), cudaMemcpyDeviceToHost)) break;
        if (count) {
            std::vector<int> hf(n);
            std::vector<uint32_t> hu(n);
            if (n > 0 && (cudaMemcpy(hf.data(), dfwd, n * sizeof(int), cudaMemcpyDeviceToHost) ||
                          cudaMemcpy(hu.data(), duid, n * sizeof(uint32_t), cudaMemcpyDeviceToHost))) break;
            int m = 0;
            for (int i = 0; i < n; i++) {
                if (hu[i] & SPH_HALO_BIT) continue;
                if (m < count_cap) { uid[m] = hu[i]; count[m] = hf[i]; }
                m++;
            }
            result = m;
        } else {
            result = (long long)np;
        }
    } while (0);
    if (result < 0) fail(ctx, SPH_ERR_CUDA, "export_pairs: CUDA failure");
    cudaFree(dpairs); cudaFree(dn); cudaFree(dfwd);
    return result;
}

extern "C" long long sph_get_pairs(sph_ctx *ctx, uint64_t *pairs, long long cap)
{
    if (!ctx) return -SPH_ERR_ARG;
    return export_pairs(ctx, pairs, pairs ? cap : 0, nullptr, nullptr, 0);
}

extern "C" int sph_get_forward_counts(sph_ctx *ctx, uint32_t *uid, int *count, int cap)
{
    if (!ctx || !uid || !count) return -SPH_ERR_ARG;
    return (int)export_pairs(ctx, nullptr, 0, uid, count, cap);
}
