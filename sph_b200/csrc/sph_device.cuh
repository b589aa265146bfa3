// Device-side definitions shared by the kernels and the C-ABI layer.
// sm_100a only.  Citations are to AdamSimpson/SPH `src/`.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sph_b200.h"

#define SPH_NROWS (2 * SPH_CELL_DIV + 1)     // sort-grid rows (and columns) a neighbourhood spans
static_assert(SPH_CELL_DIV == 1 || SPH_CELL_DIV == 2 || SPH_CELL_DIV == 4, "power of two: keeps the reference cell id exact");

#define SPH_HALO_BIT 0x80000000u
#define SPH_UID_MASK 0x7fffffffu
#define SPH_KEY_DROP (-1)
#define SPH_KEY_EMIG (1 << 30)            // entry stays resident but becomes a ghost of its new owner
#define SPH_KEY_MASK (SPH_KEY_EMIG - 1)

// Everything a kernel needs to know, resident in device memory so that a captured CUDA
// graph keeps working when the render rank changes parameters (fluid.c:293-294).
struct DevParams {
    // physics: struct TUNABLE_PARAMETERS (fluid.h:78-97)
    float rest_density, h, g, k, k_near, k_spring, sigma, beta, dt;
    float mover_cx, mover_cy, mover_w, mover_h;
    int mover_type;
    // tank (AABB min is 0: fluid.c:117) and hash grid (fluid.c:176,214-215)
    float tank_w, tank_h, cell_h;
    int size_x, size_y;          // the reference's grid (cells of side h): exported hash ids
    int sort_rows;               // rows of the sort grid = SPH_CELL_DIV * size_y
    // slab (communication.c): edges, ghost-layer width, neighbours present
    float edge_start, edge_end, halo_w;
    int has_left, has_right, nranks;
    // window of SORT-grid columns the sorted arrays cover now / will cover after the next sort
    int gx0, wx;
    int gx0_new, wx_new;
    int cap, msg_cap;
    // peer-memory exchange (NVLink): this rank's exchange block and the neighbours' blocks as mapped
    // into this process (cudaIpc); 0 = use the local send/recv buffers and an external transport
    int p2p;
    int one_x;                       // one-exchange mode: ghosts travel with x_prev and are relaxed (and advanced) redundantly
    unsigned long long xchg_base;
    unsigned long long remote_base[2];
    long long spin_timeout;          // clock64 cycles a device-side wait may last before it gives up and flags an error
};

// Parameters of the optional paths, resident like DevParams (a captured graph keeps working when they change)
// but kept out of it, so that the kernels measured in round 1 keep their machine code.
struct DevOptions {
    float visc_gamma;                // stabilised viscosity gather: s_ij = 1 / max(1, gamma max(C_i, C_j))
};

// device-side counters (ints); indices below
enum {
    CN_NTOT = 0,        // resident entries in the sorted arrays
    CN_EXTRA,           // entries appended by the unpack kernel for the coming sort
    CN_NSRC,            // entries the scatter kernel must visit (written by the scan)
    CN_NLOCAL,          // non-ghost entries after the last sort (accumulated by reorder)
    CN_MAX_BUCKET, CN_BUCKET_OVER, CN_NEIGH_OVER, CN_CAP_OVER, CN_MSG_OVER,
    CN_SPARE0, CN_SPARE1,
    CN_COORDS,          // pack_coords compaction cursor
    CN_SPARE2,
    CN_STEP,            // completed steps (message sequence numbers and buffer parity)
    CN_PUB,             // blocks that finished packing (last one publishes the message)
    CN_TIMEOUT_MSG,     // neighbour messages that never arrived (peer-memory waits that timed out)
    CN_COST,            // work estimate of the last step: sum over resident entries of (SPH_COST_BASE + neighbours)
    CN_COUNT = 24
};

// Work estimate of a slab for the optional cost-based edge policy (sph_copy_load): the three gathers cost
// a fixed part per entry plus a part per neighbour; 1 M particles took 113 us + 8.2 us per mean neighbour
// over the dam-break's states, i.e. about 14 neighbour-equivalents per entry.
#define SPH_COST_BASE 14

// One-exchange mode (sph_config.exchanges_per_step = 1; -DSPH_ONE_EXCHANGE=1 makes it the build's default):
// neighbours meet ONCE per step (or every E steps, sph_set_exchange_period).  The ghosts
// of exchange 0 carry x_prev as well, the ghost layer is >= 3h wide, and k_relax relaxes the ghosts redundantly
// (those within layer - 2h of the edge come out exactly as on their owner: same neighbours, same order), so the
// next k_advect has its neighbours' velocities without the second message.
#ifndef SPH_ONE_EXCHANGE
#define SPH_ONE_EXCHANGE 0
#endif

// neighbour message: 16-byte header {n_migrants, n_halo, 0, 0} then SoA sections sized by msg_cap
__host__ __device__ inline size_t msg_bytes_full(int m) { return 16 + (size_t)m * 40; }     // (the last 8 bytes per slot: one-exchange mode)
__host__ __device__ inline size_t msg_bytes_halo1(int m) { return 16 + (size_t)m * 20; }
__device__ __forceinline__ int *msg_hdr(unsigned char *b) { return (int *)b; }
__device__ __forceinline__ float2 *msg_a(unsigned char *b) { return (float2 *)(b + 16); }                       // migrant pos / halo-1 pos
__device__ __forceinline__ float2 *msg_b(unsigned char *b, int m) { return (float2 *)(b + 16 + (size_t)m * 8); } // migrant prev / halo-1 vel
__device__ __forceinline__ uint32_t *msg_u(unsigned char *b, int m) { return (uint32_t *)(b + 16 + (size_t)m * 16); }
__device__ __forceinline__ float2 *msg_hp(unsigned char *b, int m) { return (float2 *)(b + 16 + (size_t)m * 20); } // halo-0 pos
__device__ __forceinline__ uint32_t *msg_hu(unsigned char *b, int m) { return (uint32_t *)(b + 16 + (size_t)m * 28); }
__device__ __forceinline__ float2 *msg_hq(unsigned char *b, int m) { return (float2 *)(b + 16 + (size_t)m * 32); } // halo-0 x_prev (one-exchange build)

// Exchange block of one rank (peer-memory mode): 4 arrival flags, then one message buffer per
// (side it arrives from, which exchange, step parity).  Double buffering by step parity is enough:
// a neighbour can only be one exchange ahead, because each of its sorts waits for this rank's message.
#define SPH_XCHG_HDR 256
__host__ __device__ inline size_t xchg_msg_stride(int m) { return (msg_bytes_full(m) + 255) & ~(size_t)255; }
__host__ __device__ inline size_t xchg_offset(int side, int which, int parity, int m)
{
    return SPH_XCHG_HDR + (size_t)((side * 2 + which) * 2 + parity) * xchg_msg_stride(m);
}
__host__ __device__ inline size_t xchg_bytes(int m) { return SPH_XCHG_HDR + 8 * xchg_msg_stride(m); }
__device__ __forceinline__ int *xchg_flag(unsigned long long base, int side, int which)
{
    return (int *)(base + (size_t)(side * 2 + which) * 16);
}
// The only PTX in the library.  (SPH_EMU: tests/emu compiles this source for the host to exercise the kernel
// and C-ABI logic without a GPU -- test infrastructure; the product is built by nvcc without it.)
#ifndef SPH_EMU
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bare MUFU.RSQ / MUFU.SQRT (relative error ~2^-22; rsqrt(0) = +inf, sqrt(0) = 0): rsqrtf() wraps the
// instruction in a denormal rescue (4 more instructions per pair); squared distances below 1.2e-38 are
// flushed and behave like coincident particles
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// a line towards L1 without occupying a register: the per-particle inputs of a thread's NEXT particle, and inputs
// of this one that are only read after the candidate loops (SPH_PIPE, sph_kernels.cuh)
__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
// nanoseconds on a clock all SMs share (clock64 is per SM)
__device__ __forceinline__ long long global_timer_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one more entry in scan tile `tile`: one atomic per warp and tile (the lanes of a warp bin neighbouring particles,
// nearly always into the same tile).  Called from divergent code: the lanes that arrive together form the group.
__device__ __forceinline__ void tile_count_add(int *tile_total, int tile)
{
    const unsigned act = __activemask();
    const unsigned same = __match_any_sync(act, tile);
    if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&tile_total[tile], __popc(same));
}
// FMNMX.NAN: a NaN operand wins, so that a running maximum also reports "not finite"
__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// Packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2, two IEEE round-to-nearest fp32 operations per issued instruction;
// the gathers are bound by instruction issue, not by the FMA pipe).  A pair lives in a 64-bit register; a float2
// loaded from memory is already one.  Every operation is the same rn operation as its scalar counterpart, so
// packing changes no result bit.  Used by the SPH_PACKED build of the gather kernels.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpk2(f32x2 v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 ld2(const float2 *p) { return *reinterpret_cast<const f32x2 *>(p); }
// Asynchronous global -> shared copies (LDGSTS): the per-particle inputs of a thread's NEXT particle travel into a
// shared-memory slot of its own while the thread works through the current one (SPH_ASYNC, sph_kernels.cuh).  No register
// is held and no warp waits for them; cp_async_wait<1>() returns when all but the most recently committed group of this
// thread have landed, and a thread only ever reads the slots it filled itself, so no barrier is involved.
__device__ __forceinline__ void cp_async4(void *smem, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Programmatic dependent launch (build variant -DSPH_PDL=1; default off, then this is empty and the machine code
// of every kernel is unchanged).  Every kernel of the library starts with pdl_enter(): it waits until the grid
// before it in the stream has completed and its memory is visible -- nothing a predecessor wrote, *Pp and the
// counters included, is read earlier, so the stream order is the one of plain launches -- and then lets the grid
// after it be scheduled.  What overlaps is only the launch of grid n+1 (block scheduling, parameter loads)
// with the tail of grid n: a step is 11 launches of 4-100 us each (13 on a slab), all dependent.
#ifndef SPH_PDL
#define SPH_PDL 0
#endif
__device__ __forceinline__ void pdl_enter()
{
#if SPH_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#if SPH_PDL == 1
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
#endif
}
// SPH_PDL=2: the trigger at the END of a block's work instead (round 2: with the trigger at the top the blocks of grid
// n+1 took the block slots that grid n's second wave of blocks was waiting for, and the step got slower, 263 us
// against 253; profiles/r2_variants.md).  The dependents then start while the last blocks of grid n drain.
__device__ __forceinline__ void pdl_done()
{
#if SPH_PDL == 2
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#else
#include "sph_emu_ptx.h"      // tests/emu/fake/: host stand-ins for the helpers above (test infrastructure, not product code)
#endif

// hash_val (hash.c:35-47): IEEE fp32 divide, floor; kept as two coordinates
__device__ __forceinline__ int cell_coord(float v, float cell_h) { return (int)floorf(__fdiv_rn(v, cell_h)); }
// coordinate in the sort grid (cells of side h / SPH_CELL_DIV); sort_coord / DIV == cell_coord exactly
__device__ __forceinline__ int sort_coord(float v, float cell_h)
{
    return (int)floorf(__fmul_rn(__fdiv_rn(v, cell_h), (float)SPH_CELL_DIV));
}

// unfused squared distance: the list cut-off r2 <= h2 must match hash.c:185,221,99 bit for bit
__device__ __forceinline__ float dist2(float dx, float dy) { return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)); }

// checkVelocity (fluid.c:613-625)
// (two FMNMX; differs from the reference's compare chain only for NaN, which cannot reach it: the
//  impulse is gated by u > 0 and the velocity by finite positions)
__device__ __forceinline__ float clamp5(float v) { return fminf(fmaxf(v, -5.0f), 5.0f); }

// boundaryConditions (fluid.c:656-744)
__device__ __forceinline__ float2 boundary(float2 p, const DevParams &P)
{
    float x = p.x, y = p.y;
    if (P.mover_type == 0) {                                        // sphere: :663-685
        float radius = P.mover_w * 0.5f;
        float ddx = x - P.mover_cx, ddy = y - P.mover_cy;
        float d2 = dist2(ddx, ddy);
        if (d2 <= __fmul_rn(radius, radius) && d2 > 0.0f) {
            float d = __fsqrt_rn(d2);
            float nx = __fdiv_rn(P.mover_cx - x, d);
            float ny = __fdiv_rn(P.mover_cy - y, d);
            float pen = radius - d;
            x = __fsub_rn(x, __fmul_rn(pen, nx));
            y = __fsub_rn(y, __fmul_rn(pen, ny));
        }
    } else if (P.mover_type == 1) {                                 // rectangle: :688-727
        float hw = P.mover_w * 0.5f, hh = P.mover_h * 0.5f;
        float rx = x - P.mover_cx, ry = y - P.mover_cy;
        float ax = fabsf(rx), ay = fabsf(ry);
        if (ax < hw && ay < hh) {
            float penx = hw - ax, peny = hh - ay;
            if (penx < peny) { if (rx < 0.0f) x -= penx; else x += penx; }
            else             { if (ry < 0.0f) y -= peny; else y += peny; }
        }
    }
    if (x < 0.0f) x = 0.0f; else if (x > P.tank_w) x = P.tank_w - 0.001f;   // :732-743
    if (y < 0.0f) y = 0.0f; else if (y > P.tank_h) y = P.tank_h - 0.001f;
    return make_float2(x, y);
}

// the nearest cell of the NEW window (an emigrant that could not be sent stays resident there)
__device__ __forceinline__ int window_key_clamped(float2 p, const DevParams &P)
{
    int gx = sort_coord(p.x, P.cell_h) - P.gx0_new;
    int gy = sort_coord(p.y, P.cell_h);
    gx = min(max(gx, 0), P.wx_new - 1);
    gy = min(max(gy, 0), P.sort_rows - 1);
    return gy * P.wx_new + gx;
}

// key of a position in the NEW window (the one the coming sort will use), or DROP
__device__ __forceinline__ int window_key_new(float2 p, const DevParams &P)
{
    int gx = sort_coord(p.x, P.cell_h) - P.gx0_new;
    int gy = sort_coord(p.y, P.cell_h);
    if (gx < 0 || gx >= P.wx_new || gy < 0 || gy >= P.sort_rows) return SPH_KEY_DROP;
    return gy * P.wx_new + gx;
}
