// Kernels of the B200-native TinySPH compute-rank timestep (sm_100a).
//
// State is cell-sorted SoA (float2 position, float2 velocity-or-previous-position, u32 uid with
// a ghost flag).  Cells are the reference's hash cells (hash.c:35-47) split SPH_CELL_DIV times per
// axis, numbered row-major inside the slab's window of grid columns, so the neighbourhood of a
// particle is 2*DIV+1 CONTIGUOUS INDEX RANGES (one per sort-grid row; the columns gx-DIV..gx+DIV of a
// row are adjacent in memory).
//
// The reference's symmetric pair scatter (fluid.c:461-471, :598-607) is a gather here: every
// particle sums the contributions of all particles within h, in a fixed order (row, then sorted
// index == ascending uid inside a cell), so there are no atomics on the pair loops and results
// are bit-reproducible and independent of the slab decomposition.
#pragma once

#include "sph_device.cuh"

#define SPH_THREADS 256
#ifndef SCAN_ITEMS
#define SCAN_ITEMS 4           // (a multiple of 4: the tile is loaded and stored as int4)
#endif
#define SCAN_TILE (SPH_THREADS * SCAN_ITEMS)      // cells per tile of the prefix sum (k_scan_apply)
// SPH_TILE_ATOMICS=1 (round 2): the kernels that bin a position also add it to its scan tile's total (one atomic per
// warp and tile), so the prefix sum needs no pass of its own over the cell populations to form the tile totals:
// k_scan_totals (5.5 us per sort at 1 M particles, latency-bound) disappears, the sort is three kernels instead of four.
#ifndef SPH_TILE_ATOMICS
#define SPH_TILE_ATOMICS 1
#endif
// SPH_SORT_SRC=1 (default since round 2's second half; measured, profiles/r2_variants.md: 35 -> 30.6 us per sort, 27.7
// with the grids and trip sizes below): the sort's last two kernels walk the SOURCE order.  k_scatter_uid stores only a
// uid into its cell's range (arrival order); k_reorder_src then reads key, uid and payload of a source entry coalesced,
// ranks the uid inside its cell (cells of population 1 -- most of them -- need no look at all) and stores the entry at
// its final place.  The (uid, source, key) triple no longer round-trips through memory: 16 bytes per entry less, and the
// payload loads no longer hang behind the load of their own index.  SPH_SORT_ITEMS entries per thread and trip.
#ifndef SPH_SORT_SRC
#define SPH_SORT_SRC 1
#endif
// SPH_SCAN_FAST=1 (default, same series): k_scan_apply asks for its tile's populations BEFORE it forms the tile's offset (the DRAM latency
// overlaps the reduction), forms offset, warp totals and bucket statistics behind ONE barrier instead of four, issues
// one statistics atomic per tile instead of one per warp, and neither reads nor clears the populations of a tile
// whose total is zero (the air above the fluid: half the table in the dam-break).
#ifndef SPH_SCAN_FAST
#define SPH_SCAN_FAST 1
#endif

#ifndef SPH_UNROLL
#define SPH_UNROLL 4          // candidates per trip of the gather loops (loads issued together)
#endif
constexpr int kGatherUnroll = SPH_UNROLL;     // (#pragma unroll takes a constant expression, not a macro)
// SPH_PACKED=1 (default since round 2): the candidate loops of k_advect, k_coupling and k_density process two
// candidates per trip with packed FP32 instructions (FADD2 / FMUL2 / FFMA2, sph_device.cuh): the same rn operations
// in the same order, so the results do not change by a bit (tests/test_emu_variants.py).  Measured on the B200
// (profiles/r2_variants.md): 26 % fewer issued instructions in k_advect but only 4-5 % less time (90 -> 86 us) --
// a packed instruction holds the FMA pipe for two passes, so the pipe cycles stay and only issue slots are freed.
#ifndef SPH_PACKED
#define SPH_PACKED 1
#endif
// SPH_PACKED_RELAX=1 (default since round 2): the same for k_relax's pair physics ((x, y) accumulate in one
// register).  Measured: within noise of the scalar pair (66.1 vs 66.9 us).
#ifndef SPH_PACKED_RELAX
#define SPH_PACKED_RELAX 1
#endif
// SPH_TRIM=1 (default since round 2): k_advect's candidate loop applies a pair's half impulse with two FFMAs and
// moves the reference's per-component +-5 clamp (fluid.c:459) into an exact redo of the rare rows in which it can
// bind (see the kernel).  Measured: k_advect 86 -> 82 us on top of the packed loop.
#ifndef SPH_TRIM
#define SPH_TRIM 1
#endif
// SPH_PIPE=1 (from the per-instruction stall samples of round 2's call 2): the per-particle prologue and epilogue of
// the gathers were a chain of dependent long-latency accesses -- uid, then position, then an IEEE division for the
// cell, then the ten range bounds, (k_relax: then one mask per row), ..., then the atomic that hands out the arrival
// slot and the store that waits for it -- and 18 % (k_advect) to 40 % (k_relax) of all stall samples sat on them.
// With the flag: the sort-grid cell comes from the key the sort left behind (ord_key is constant inside a cell, so
// ord_key[i] IS the key of sorted entry i: no division, no dependence on the position), the inputs of the thread's
// next particle and the late inputs of this one are prefetched into L1, and the arrival slot is stored one iteration
// later, when its atomic has long returned.  Same arithmetic, same results.  Measured on the B200 and REJECTED as a
// default (0): k_advect 88 -> 84 us, but k_relax 66 -> 72 us (the stall samples on the prologue fell by a quarter, yet
// the added prefetch traffic and code size cost more).
#ifndef SPH_PIPE
#define SPH_PIPE 0
#endif
// SPH_RELAX_PD4=1: k_density also writes (x, y, density, density_near) as one 16-byte record per entry and
// k_relax's neighbour walk reads that record: one load and one address per neighbour instead of two.
// Measured on the B200 in round 2 and REJECTED: k_relax 74 us against 67 (the extra 16 bytes per particle that
// k_density writes and the larger L1 footprint cost more than the saved address arithmetic).
#ifndef SPH_RELAX_PD4
#define SPH_RELAX_PD4 0
#endif
// SPH_RELAX_BF=1: k_relax's neighbour loop branch-free in candidate order with mask-predicated loads (see the kernel).
// Measured on the B200 in round 2 and REJECTED: 88 us against 66 -- a warp runs to the highest set bit of its 32
// lanes, 53 M issued warp instructions against the walk's 37 M.
#ifndef SPH_RELAX_BF
#define SPH_RELAX_BF 0
#endif
// SPH_ASYNC (bit mask: 1 k_advect, 2 k_relax, 4 k_density): the per-particle inputs of a thread's next particle are
// copied global -> shared asynchronously (cp.async, LDGSTS) while it works on the current one.  The stall samples of the
// r2_final capture (profiles/r2_final_source_regions.md) put 19 % (k_advect) and 38 % (k_relax) of all samples on a chain
// of dependent per-particle loads -- uid, then position, then one acceptance mask per row, then the previous position
// -- each a full memory round trip that only other warps could hide; prefetching them into L1 (SPH_PIPE) did not help.
// SPH_DEFER=1: the store of a particle's arrival slot waits for its atomic until the thread's next particle is done
// (5 % of the samples sat on that store).
#ifndef SPH_ASYNC
#define SPH_ASYNC 2         // (measured: staging pays in k_relax only, profiles/r2_variants.md)
#endif
#ifndef SPH_DEFER
#define SPH_DEFER 0
#endif
// SPH_PREFETCH (same bit mask; needs the matching SPH_ASYNC bit): when a thread has finished a particle it forms the
// candidate rows of its NEXT one (whose position is already in its staging slot) and asks for the first and last line
// of each row towards L1.  The candidate loads of the gathers hit L1 73-83 % of the time; what misses is the first
// touch of a row by a block, and a warp that meets one in a trip waits an L2 round trip with 8 warps per scheduler to
// cover it.
#ifndef SPH_PREFETCH
#define SPH_PREFETCH 0
#endif
#if (SPH_PREFETCH & ~SPH_ASYNC) != 0
#error "SPH_PREFETCH needs the same bits in SPH_ASYNC (the next particle's position comes from its staging slot)"
#endif
#define SPH_DEFER_ADVECT (SPH_PIPE || (SPH_DEFER & 1))      // (SPH_DEFER is a bit mask like SPH_ASYNC: 1 k_advect, 2 k_relax)
#define SPH_DEFER_RELAX (SPH_PIPE || (SPH_DEFER & 2))
// SPH_KEYROWS=1: a particle's candidate rows from the cell key its sort assigned (one more 4-byte load per particle)
// instead of from its position: no IEEE divisions in the prologue of the three gathers (about 60 of its 100-110
// instructions), and the ten cell_start loads no longer wait for them.  This is the first part of SPH_PIPE on its own.
#ifndef SPH_KEYROWS
#define SPH_KEYROWS 0
#endif
#define SPH_ROWS_FROM_KEY (SPH_PIPE || SPH_KEYROWS)
// SPH_ADVECT_PV4=1 (needs SPH_SORT_SRC): the reorder of sort 2 also writes (x, y, vx, vy) records, and k_advect's candidate
// loop reads one 16-byte record per candidate instead of a position and a velocity from two arrays: half the load
// instructions and half the L1 tag look-ups of the loop (l1tex throughput, 45 % of peak, was its busiest unit after the
// issue slots).  Same values, same arithmetic.
#ifndef SPH_ADVECT_PV4
#define SPH_ADVECT_PV4 0
#endif
#if SPH_ADVECT_PV4 && !(SPH_SORT_SRC && SPH_PACKED && SPH_TRIM)
#error "SPH_ADVECT_PV4 is written for the source-order sort and the packed, trimmed loop"
#endif
#if SPH_ADVECT_PV4
#define SPH_PV4_PARAM , float4 *__restrict__ pv
#define SPH_PV4_CPARAM , const float4 *__restrict__ pv
#else
#define SPH_PV4_PARAM
#define SPH_PV4_CPARAM
#endif
// SPH_PAIRMASK=1: the packed candidate loops of k_advect (trimmed loop) / k_coupling / k_density take a row's LAST odd
// candidate in a masked pair trip instead of a separate scalar loop: the pair's second slot then reads the entry just past
// the range (allocated, finite: the arrays are zero-filled and padded at creation) and is excluded by a validity
// predicate folded into the membership test, so it adds an exact zero.  The per-instruction counts of round 2's capture
// showed every warp paying the 28-40 instructions of the left-over loop in every row (5 rows per particle: 12-13 % of
// all instructions of k_advect / k_density).  Same operations in the same order for every real candidate.
#ifndef SPH_PAIRMASK
#define SPH_PAIRMASK 0
#endif
#if SPH_PAIRMASK && !(SPH_PACKED && SPH_TRIM)
#error "SPH_PAIRMASK is written for the packed, trimmed loops"
#endif
// SPH_RELAX_RARE=1: k_relax's mask walk without the coincident-pair rules in its loop.  The walk's pair physics is
// straight-line (no branch per neighbour, the neighbours of a trip interleave); a pair with r2 <= 1e-12 -- the only case
// fluid.c:583-588 treats differently -- raises a flag, and a flagged particle is redone from its entry position by the
// exact walk.  Same operations in the same order for every other particle.  SPH_RELAX_TRIP: neighbours per trip.
#ifndef SPH_RELAX_RARE
#define SPH_RELAX_RARE 1
#endif
#ifndef SPH_RELAX_TRIP
#define SPH_RELAX_TRIP 3
#endif
#if SPH_RELAX_PD4
#define SPH_PD4_PARAM , float4 *__restrict__ pd
#define SPH_PD4_CPARAM , const float4 *__restrict__ pd
#else
#define SPH_PD4_PARAM
#define SPH_PD4_CPARAM
#endif
constexpr int kPackedUnroll = SPH_UNROLL / 2 > 0 ? SPH_UNROLL / 2 : 1;
// resident blocks per SM the register allocation of each gather is held to (measured, DESIGN.md 8: 64 / 40 / 64
// registers; 5 blocks of k_advect or k_relax spill and lose, profiles/r2_variants.md)
#ifndef SPH_BLOCKS_ADVECT
#define SPH_BLOCKS_ADVECT 4
#endif
#ifndef SPH_BLOCKS_DENSITY
#define SPH_BLOCKS_DENSITY 6
#endif
#ifndef SPH_BLOCKS_RELAX
#define SPH_BLOCKS_RELAX 4
#endif

// -------------------------------------------------------------------------------------------
// neighbourhood iteration: sort-grid rows gy-DIV..gy+DIV, columns gx-DIV..gx+DIV of the CURRENT window
// -------------------------------------------------------------------------------------------
struct Rows { int b[SPH_NROWS], e[SPH_NROWS]; };

// per-row acceptance masks handed from k_density to k_relax: 3 rows x 64 bits for full-size cells,
// 5 rows x 32 bits for half-size cells (a row then holds ~8 candidates, rarely more than 32)
#if SPH_CELL_DIV == 1
typedef unsigned long long sph_mask_t;
#define SPH_MASK_BITS 64
__device__ __forceinline__ int sph_mask_ffs(sph_mask_t m) { return __ffsll((long long)m); }
__device__ __forceinline__ int sph_mask_popc(sph_mask_t m) { return __popcll(m); }
#else
typedef unsigned int sph_mask_t;
#define SPH_MASK_BITS 32
__device__ __forceinline__ int sph_mask_ffs(sph_mask_t m) { return __ffs((int)m); }
__device__ __forceinline__ int sph_mask_popc(sph_mask_t m) { return __popc(m); }
#endif

// r and 1/r from the (exact, unfused) squared distance with one MUFU.RSQ instead of the IEEE sqrt +
// divide sequences (~18 instructions).  Only the pair PHYSICS uses these; list membership is decided
// on r2 itself, so neighbour sets stay bit-exact.  Relative error ~2^-22: orders of magnitude below
// the stated parity tolerance.  r2 == 0 (coincident particles) gives r = 0, 1/r = +inf like the
// reference's 1.0f/r.
__device__ __forceinline__ void r_and_recip(float r2, float &r, float &r_recip)
{
    r_recip = rsqrt_approx(r2);
    r = r2 > 0.0f ? r2 * r_recip : 0.0f;
}

// Candidate ranges of a particle: the columns gx-DIV..gx+DIV of the rows gy-DIV..gy+DIV.
// (Cutting each row's columns to what its distance in y leaves reachable -- about 14 % fewer candidates
//  per particle for DIV = 2 -- was measured and gained nothing: a warp runs the LONGEST range of its 32
//  lanes, and those sit at different places inside their cells; the per-lane ranges also stop coinciding,
//  which costs L1 wavefronts.)
__device__ __forceinline__ Rows candidate_rows(float2 p, const DevParams &P, const int *__restrict__ cell_start)
{
    Rows r;
    int gx = sort_coord(p.x, P.cell_h) - P.gx0;
    int gy = sort_coord(p.y, P.cell_h);
    gx = min(max(gx, 0), P.wx - 1);
    gy = min(max(gy, 0), P.sort_rows - 1);
    int c0 = max(gx - SPH_CELL_DIV, 0), c1 = min(gx + SPH_CELL_DIV, P.wx - 1);
#pragma unroll
    for (int d = 0; d < SPH_NROWS; d++) {
        int row = gy + d - SPH_CELL_DIV;
        if (row < 0 || row >= P.sort_rows) { r.b[d] = 0; r.e[d] = 0; continue; }
        r.b[d] = __ldg(&cell_start[row * P.wx + c0]);
        r.e[d] = __ldg(&cell_start[row * P.wx + c1 + 1]);
    }
    return r;
}


// The same ranges from the cell key the last sort assigned to this entry (key = gy * wx + gx in the current window;
// an entry in the sorted arrays is inside the window, so no clamping of the cell itself is needed).
__device__ __forceinline__ Rows candidate_rows_key(int key, const DevParams &P, float inv_wx, const int *__restrict__ cell_start)
{
    Rows r;
    // key / wx by reciprocal and one correction step (quotient < 2^15 rows: the float estimate is off by at most 1)
    int gy = (int)((float)key * inv_wx);
    int gx = key - gy * P.wx;
    if (gx < 0) { gy--; gx += P.wx; } else if (gx >= P.wx) { gy++; gx -= P.wx; }
    const int c0 = max(gx - SPH_CELL_DIV, 0), c1 = min(gx + SPH_CELL_DIV, P.wx - 1);
    const int base = key - gx;
#pragma unroll
    for (int d = 0; d < SPH_NROWS; d++) {
        const int row = gy + d - SPH_CELL_DIV;
        if (row < 0 || row >= P.sort_rows) { r.b[d] = 0; r.e[d] = 0; continue; }
        const int rb = base + (d - SPH_CELL_DIV) * P.wx;
        r.b[d] = __ldg(&cell_start[rb + c0]);
        r.e[d] = __ldg(&cell_start[rb + c1 + 1]);
    }
    return r;
}

// first and last line of each candidate row towards L1 (rows span 1-2 lines of 128 bytes)
template <class T>
__device__ __forceinline__ void prefetch_rows(const Rows &R, const T *__restrict__ arr)
{
#pragma unroll
    for (int d = 0; d < SPH_NROWS; d++)
        if (R.e[d] > R.b[d]) { prefetch_l1(arr + R.b[d]); prefetch_l1(arr + R.e[d] - 1); }
}

// -------------------------------------------------------------------------------------------
// Peer-memory exchange.  The compute kernels (k_advect, k_relax) append outgoing records to a compact
// LOCAL message; the first kernel of the following sort (k_unpack) then (1) copies that message into
// the neighbour GPU's exchange block over NVLink with coalesced stores spread over the whole grid,
// (2) the last block to finish releases the neighbour's arrival flag with system scope, and (3) every
// block waits for this rank's own arrival flags before unpacking.  Both neighbours send before they
// wait; the host launches this kernel with a grid that is fully co-resident (occupancy API), because a
// message is released only after EVERY block of the sender has run -- with unscheduled blocks on both
// sides the resident ones would wait for each other forever.  Nothing goes through the host or a
// collective library.
// (Storing each record remotely from inside the compute kernels was tried first: the scattered 8-byte
// NVLink stores lengthened k_advect/k_relax by 30-45 us.)
// -------------------------------------------------------------------------------------------
// (16 bytes per store: the message capacity is a multiple of 4, so every section starts on a 16-byte boundary and
//  is a multiple of 16 bytes long; the last store of a section may carry up to 12 stale bytes of the same section)
__device__ __forceinline__ void copy_words(unsigned char *dst, const unsigned char *src, size_t off, int nwords,
                                           int gtid, int gstride, bool &wrote)
{
    const int4 *s = (const int4 *)(src + off);
    int4 *d = (int4 *)(dst + off);
    const int nq = (nwords + 3) >> 2;
    for (int w = gtid; w < nq; w += gstride) { d[w] = s[w]; wrote = true; }
}

__device__ __forceinline__ void send_messages(const DevParams &P, int *counters, int which, int step,
                                              unsigned char *send_l, unsigned char *send_r)
{
    unsigned char *local[2] = {send_l, send_r};
    const int present[2] = {P.has_left, P.has_right};
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    const int m = P.msg_cap;
    bool wrote = false;
    for (int s = 0; s < 2; s++) {
        if (!present[s]) continue;
        // it arrives at the neighbour from the opposite side
        unsigned char *remote = (unsigned char *)(P.remote_base[s] + xchg_offset(1 - s, which, step & 1, m));
        const int n_mig = which == 0 ? min(msg_hdr(local[s])[0], m) : 0;
        const int n_halo = min(msg_hdr(local[s])[1], m);
        if (which == 0) {
            copy_words(remote, local[s], 16, n_mig * 2, gtid, gstride, wrote);                      // migrant pos
            copy_words(remote, local[s], 16 + (size_t)m * 8, n_mig * 2, gtid, gstride, wrote);      // migrant x_prev
            copy_words(remote, local[s], 16 + (size_t)m * 16, n_mig, gtid, gstride, wrote);         // migrant uid
            copy_words(remote, local[s], 16 + (size_t)m * 20, n_halo * 2, gtid, gstride, wrote);    // ghost pos
            copy_words(remote, local[s], 16 + (size_t)m * 28, n_halo, gtid, gstride, wrote);        // ghost uid
            if (P.one_x) copy_words(remote, local[s], 16 + (size_t)m * 32, n_halo * 2, gtid, gstride, wrote);    // ghost x_prev
        } else {
            copy_words(remote, local[s], 16, n_halo * 2, gtid, gstride, wrote);                     // ghost pos
            copy_words(remote, local[s], 16 + (size_t)m * 8, n_halo * 2, gtid, gstride, wrote);     // ghost vel
            copy_words(remote, local[s], 16 + (size_t)m * 16, n_halo, gtid, gstride, wrote);        // ghost uid
        }
    }
    // only blocks that stored remotely pay for a system-scope fence (they run in parallel)
    const int any = __syncthreads_or(wrote ? 1 : 0);
    if (threadIdx.x == 0) {
        if (any) __threadfence_system(); else __threadfence();
        if (atomicAdd(&counters[CN_PUB], 1) == (int)gridDim.x - 1) {
            counters[CN_PUB] = 0;
            // last block: counts into the remote headers, ONE system-scope fence, then the two arrival
            // flags as plain stores (fence + store = release; every other block fenced before its
            // atomicAdd, which this thread has observed).  The first version issued four system fences
            // back to back here; at several microseconds each they were most of the exchange cost.
            for (int s = 0; s < 2; s++) {
                if (!present[s]) continue;
                int *rh = msg_hdr((unsigned char *)(P.remote_base[s] + xchg_offset(1 - s, which, step & 1, m)));
                rh[0] = which == 0 ? min(msg_hdr(local[s])[0], m) : 0;
                rh[1] = min(msg_hdr(local[s])[1], m);
            }
            __threadfence_system();
            for (int s = 0; s < 2; s++)
                if (present[s]) *(volatile int *)xchg_flag(P.remote_base[s], 1 - s, which) = 2 * step + which + 1;
        }
    }
}

// bin a freshly produced position for the coming sort: key, arrival slot, cell population
__device__ __forceinline__ void bin_position(int i, float2 p, int extra_bits, const DevParams &P,
                                             int *__restrict__ cnt, int *__restrict__ t_key,
                                             int *__restrict__ t_slot, int *__restrict__ counters, int *__restrict__ tile_total,
                                             bool keep = false)
{
    int key = window_key_new(p, P);
    if (key == SPH_KEY_DROP && keep) key = window_key_clamped(p, P);
    if (key == SPH_KEY_DROP) {
        // Outside this slab's window.  For an emigrant that is fine: it has been handed to its new owner
        // and is too far away to matter as a ghost (a mover can push a particle several cells at once).
        // For anything else the particle would be lost: counted, and fatal for the caller.
        if (!(extra_bits & SPH_KEY_EMIG)) atomicAdd(&counters[CN_CAP_OVER], 1);
        t_key[i] = SPH_KEY_DROP;
        return;
    }
    t_slot[i] = atomicAdd(&cnt[key], 1);
    t_key[i] = key | extra_bits;
#if SPH_TILE_ATOMICS
    tile_count_add(tile_total, key / SCAN_TILE);
#endif
}

// The same with the store of the arrival slot left to the caller (SPH_PIPE): returns the slot, or -1 when the entry
// was dropped.  The caller stores it an iteration later, so that no instruction waits for the atomic's round trip.
__device__ __forceinline__ bool bin_position_deferred(int i, float2 p, int extra_bits, const DevParams &P,
                                                      int *__restrict__ cnt, int *__restrict__ t_key, int *__restrict__ counters,
                                                      int *__restrict__ tile_total, int &slot, bool keep = false)
{
    int key = window_key_new(p, P);
    if (key == SPH_KEY_DROP && keep) key = window_key_clamped(p, P);
    if (key == SPH_KEY_DROP) {
        if (!(extra_bits & SPH_KEY_EMIG)) atomicAdd(&counters[CN_CAP_OVER], 1);
        t_key[i] = SPH_KEY_DROP;
        return false;
    }
    t_key[i] = key | extra_bits;
    slot = atomicAdd(&cnt[key], 1);        // (the caller must not look at it before the next iteration)
#if SPH_TILE_ATOMICS
    tile_count_add(tile_total, key / SCAN_TILE);
#endif
    return true;
}

// -------------------------------------------------------------------------------------------
// K1  apply_gravity (fluid.c:398) + viscosity_impluses (:416) + predict_positions (:507)
//     + boundaryConditions (:656) + identify_oob_particles (:481) + ghost selection
//     (communication.c:134-141) + first half of hash_fluid pass 1 (hash.c:153-166)
// -------------------------------------------------------------------------------------------
// s_ij = 1 / max(1, gamma max(C_i, C_j)): symmetric, and exactly 1 wherever the plain gather is stable
template <bool STAB>
__device__ __forceinline__ float stab_scale(float t, float gci, float gamma, const float *__restrict__ coupling, int j)
{
    if constexpr (STAB) {
        const float gc = fmaxf(gci, gamma * coupling[j]);
        return t * (gc > 1.0f ? rcp_approx(gc) : 1.0f);
    } else {
        return t;
    }
}

// STAB: the stabilised viscosity gather (k_coupling below); false = the plain gather, the default.
// HOLD: the instantiation launched for a slab's steps BETWEEN two exchanges (exchange period > 1): nobody can be handed
// over in such a step, so a local outside the slab's window -- an emigrant still waiting for room in a message: a parked
// slab, or one that has just given half of itself away, holds more of them than one message takes -- stays resident
// (binned into the nearest window cell) instead of being dropped.  Found by tests/fuzz/fuzz_slabs.py: a slab parked
// under an exchange period of 2 lost whatever it had not sent yet in the first step without an exchange.  A separate
// instantiation so that the machine code of every other step (all of a single slab's, and every exchange step) is
// exactly the code that was measured and profiled (profiles/r2b_sass_hashes.json).
template <bool STAB, bool HOLD = false>
__global__ void __launch_bounds__(SPH_THREADS, SPH_BLOCKS_ADVECT)
k_advect(const DevParams *__restrict__ Pp, int *__restrict__ counters,
         const float2 *__restrict__ pos, const float2 *__restrict__ vel, const uint32_t *__restrict__ uid,
         const int *__restrict__ cell_start,
         float2 *__restrict__ pos_pred, int *__restrict__ cnt, int *__restrict__ t_key, int *__restrict__ t_slot,
         unsigned char *send_l, unsigned char *send_r, const float *__restrict__ coupling, const DevOptions *__restrict__ Op,
         int xstep, const int *__restrict__ ckey, int *__restrict__ tile_total SPH_PV4_CPARAM)
{
    // xstep: neighbours exchange after this prediction (always, except in the one-exchange build with an exchange
    // period > 1, where between exchanges the ghosts are advanced here like everything else; see sph_set_exchange_period)
    pdl_enter();
    const DevParams P = *Pp;
    const float gamma = STAB ? Op->visc_gamma : 0.0f;
    const int n = counters[CN_NTOT];
    const float dt = P.dt;
    const float gdt = (-P.g) * dt;
    const float hdt = 0.5f * dt;
    const float h_recip = __fdiv_rn(1.0f, P.h);
    const float h2 = __fmul_rn(P.h, P.h);
    if (blockIdx.x == 0 && threadIdx.x == 0) { counters[CN_MAX_BUCKET] = 0; counters[CN_COST] = 0; }
#if SPH_ROWS_FROM_KEY
    const float inv_wx = 1.0f / (float)P.wx;
#endif
#if SPH_DEFER_ADVECT
    int slot_i = -1, slot_v = 0;                                        // arrival slot whose store is still owed
#endif
#if SPH_ASYNC & 1
    // staging slots of this thread: [buffer][thread]
    __shared__ uint32_t s_uid[2][SPH_THREADS];
    __shared__ float2 s_pos[2][SPH_THREADS], s_vel[2][SPH_THREADS];
    const int gstride = gridDim.x * blockDim.x;
    int buf = 0;
    {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < n) { cp_async4(&s_uid[0][threadIdx.x], uid + i0); cp_async8(&s_pos[0][threadIdx.x], pos + i0); cp_async8(&s_vel[0][threadIdx.x], vel + i0); }
        cp_async_commit();
    }
#endif

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#if SPH_ASYNC & 1
        {
            const int nx = i + gstride;
            if (nx < n) { cp_async4(&s_uid[buf ^ 1][threadIdx.x], uid + nx); cp_async8(&s_pos[buf ^ 1][threadIdx.x], pos + nx); cp_async8(&s_vel[buf ^ 1][threadIdx.x], vel + nx); }
            cp_async_commit();
            cp_async_wait<1>();                                         // this particle's copies (the group before) have landed
        }
        const uint32_t u = s_uid[buf][threadIdx.x];
        const float2 p = s_pos[buf][threadIdx.x];
        const float2 v0 = s_vel[buf][threadIdx.x];
        buf ^= 1;
#else
        const uint32_t u = uid[i];
#endif
#if SPH_PIPE
        // all of this particle's inputs are requested before the first of them is looked at
        const float2 p = pos[i];
        const float2 v0 = vel[i];
        const int key_i = ckey[i];
        {
            const int nx = i + gridDim.x * blockDim.x;
            if (nx < n) { prefetch_l1(uid + nx); prefetch_l1(pos + nx); prefetch_l1(vel + nx); prefetch_l1(ckey + nx); }
        }
#endif
        // ghosts are replaced at every exchange; in the one-exchange mode there are steps without one (exchange period > 1),
        // in which a ghost is advanced here like everything else
#if SPH_KEYROWS && !SPH_PIPE
        const int key_i = ckey[i];
#endif
        const bool ghost = (u & SPH_HALO_BIT) != 0;
        if (ghost && (xstep || !P.one_x)) { t_key[i] = SPH_KEY_DROP; continue; }
#if !SPH_PIPE && !(SPH_ASYNC & 1)
        const float2 p = pos[i];
        const float2 v0 = vel[i];
#endif
        float vx = v0.x, vy = v0.y + gdt;                               // apply_gravity
#if SPH_ROWS_FROM_KEY
        const Rows R = candidate_rows_key(key_i, P, inv_wx, cell_start);
#else
        const Rows R = candidate_rows(p, P, cell_start);
#endif
        const float gci = STAB ? gamma * coupling[i] : 0.0f;
#if SPH_TRIM
        // SPH_TRIM=1 (round 2): the candidate loop without the per-component clamp.  The +-2.5 clamp of a pair's
        // half impulse (fluid.c:459 on the whole impulse) can only bind when |t| h > 2.5, because a neighbour has
        // |dx|, |dy| <= h; the loop applies v -= t d as two FFMAs, tracks max |t| over the row (one FMNMX.NAN), and
        // only a row in which the clamp COULD bind (never, in a settled fluid) is redone from its saved entry
        // velocity with the exact body below.  26 instead of 33 issued instructions per candidate; -w hdt is formed
        // as fma(r, hdt/h, -hdt).  Rounding differs from the default build's (fused accumulate), order does not.
        const float hdt_over_h = hdt * h_recip, tlim = 2.5f * h_recip;
        // one candidate, exactly as in the default build: used for the rows that need the clamp
        auto cand_exact = [&](int j) {
            const float2 q = pos[j];
            const float dx = q.x - p.x, dy = q.y - p.y;
            const float r2 = dist2(dx, dy);
            const bool in = r2 <= h2;
            const float2 vq = vel[j];
            const float rs = rsqrt_approx(r2);
            const float u_in = ((v0.x - vq.x) * dx + (v0.y - vq.y) * dy) * rs;
            const bool hit = in && u_in > 0.0f;
            const float whdt = fmaf(-r2 * rs, h_recip, 1.0f) * hdt;
            const float t = stab_scale<STAB>(u_in * fmaf(P.beta, u_in, P.sigma) * whdt * rs, gci, gamma, coupling, j);
            const float ix = fminf(fmaxf(t * dx, -2.5f), 2.5f);
            const float iy = fminf(fmaxf(t * dy, -2.5f), 2.5f);
            vx -= hit ? ix : 0.0f;
            vy -= hit ? iy : 0.0f;
        };
#if SPH_PACKED
        const f32x2 pp = pk2(p.x, p.y), v0p = pk2(v0.x, v0.y);
        const f32x2 hh2 = pk2(hdt_over_h, hdt_over_h), nhdt2 = pk2(-hdt, -hdt);
        const f32x2 beta2 = pk2(P.beta, P.beta), sigma2 = pk2(P.sigma, P.sigma);
#endif
#pragma unroll
        for (int d = 0; d < SPH_NROWS; d++) {
            const float vx_row = vx, vy_row = vy;
            float tmax = 0.0f;
            int j = R.b[d];
            const int je = R.e[d];
#if SPH_PACKED
            f32x2 vv = pk2(vx, vy);
#pragma unroll kPackedUnroll
#if SPH_PAIRMASK
            for (; j < je; j += 2) {
                const bool v1 = j + 1 < je;                 // the row's last trip may hold one candidate only
#else
            for (; j + 1 < je; j += 2) {
                const bool v1 = true;
#endif
#if SPH_ADVECT_PV4
                const float4 c0 = pv[j], c1 = pv[j + 1];
                const f32x2 d0 = sub2(pk2(c0.x, c0.y), pp), d1 = sub2(pk2(c1.x, c1.y), pp);
                const f32x2 w0 = sub2(v0p, pk2(c0.z, c0.w)), w1 = sub2(v0p, pk2(c1.z, c1.w));
#else
                const f32x2 d0 = sub2(ld2(pos + j), pp), d1 = sub2(ld2(pos + j + 1), pp);
                const f32x2 w0 = sub2(v0p, ld2(vel + j)), w1 = sub2(v0p, ld2(vel + j + 1));
#endif
                const float2 s0 = unpk2(mul2(d0, d0)), s1 = unpk2(mul2(d1, d1));
                const float r20 = __fadd_rn(s0.x, s0.y), r21 = __fadd_rn(s1.x, s1.y);
                const float2 m0 = unpk2(mul2(w0, d0)), m1 = unpk2(mul2(w1, d1));
                const f32x2 rs = pk2(rsqrt_approx(r20), rsqrt_approx(r21));
                const f32x2 u = mul2(pk2(__fadd_rn(m0.x, m0.y), __fadd_rn(m1.x, m1.y)), rs);
                const float2 uu = unpk2(u);
                const bool hit0 = (r20 <= h2) & (uu.x > 0.0f), hit1 = v1 & (r21 <= h2) & (uu.y > 0.0f);
                const f32x2 nw = fma2(mul2(pk2(r20, r21), rs), hh2, nhdt2);          // -(1 - r/h) dt/2
                f32x2 t = mul2(mul2(mul2(u, fma2(beta2, u, sigma2)), nw), rs);
                if (STAB) {
                    const float g0 = fmaxf(gci, gamma * coupling[j]), g1 = fmaxf(gci, gamma * coupling[j + 1]);
                    t = mul2(t, pk2(g0 > 1.0f ? rcp_approx(g0) : 1.0f, g1 > 1.0f ? rcp_approx(g1) : 1.0f));
                }
                const float2 tt = unpk2(t);
                const float t0 = hit0 ? tt.x : 0.0f, t1 = hit1 ? tt.y : 0.0f;
                tmax = max_nan(tmax, fabsf(t0));
                tmax = max_nan(tmax, fabsf(t1));
                vv = fma2(pk2(t0, t0), d0, vv);
                vv = fma2(pk2(t1, t1), d1, vv);
            }
            { const float2 v2 = unpk2(vv); vx = v2.x; vy = v2.y; }
#endif
#if !SPH_PAIRMASK
#if SPH_PACKED
#pragma unroll 1
            for (; j < je; j++) {      // at most one left over
#else
#pragma unroll kGatherUnroll
            for (; j < je; j++) {
#endif
                const float2 q = pos[j];
                const float dx = q.x - p.x, dy = q.y - p.y;
                const float r2 = dist2(dx, dy);
                const bool in = r2 <= h2;
                const float2 vq = vel[j];
                const float rs = rsqrt_approx(r2);
                const float u_in = ((v0.x - vq.x) * dx + (v0.y - vq.y) * dy) * rs;
                const bool hit = in && u_in > 0.0f;
                const float nw = fmaf(r2 * rs, hdt_over_h, -hdt);
                float t = stab_scale<STAB>(u_in * fmaf(P.beta, u_in, P.sigma) * nw * rs, gci, gamma, coupling, j);
                t = hit ? t : 0.0f;
                tmax = max_nan(tmax, fabsf(t));
                vx = fmaf(t, dx, vx);
                vy = fmaf(t, dy, vy);
            }
#endif
            if (!(tmax <= tlim)) {      // the clamp may bind in this row (or something was not finite): exact body
                vx = vx_row; vy = vy_row;
#pragma unroll 1
                for (int k = R.b[d]; k < je; k++) cand_exact(k);
            }
        }
#else
#if SPH_PACKED
        const f32x2 pp = pk2(p.x, p.y), v0p = pk2(v0.x, v0.y);
        const f32x2 nh2 = pk2(-h_recip, -h_recip), one2 = pk2(1.0f, 1.0f), hdt2 = pk2(hdt, hdt);
        const f32x2 beta2 = pk2(P.beta, P.beta), sigma2 = pk2(P.sigma, P.sigma);
#endif
#pragma unroll
        for (int d = 0; d < SPH_NROWS; d++) {
            // Branch-free body (see k_density): a candidate that is not a neighbour, or a neighbour that
            // is not approaching, subtracts nothing, so values and order of the sums are those of the gated
            // loop.  With the branches the warp ran the impulse path on nearly every trip with half its
            // lanes idle.
#if SPH_PACKED
            int j = R.b[d];
            const int je = R.e[d];
            f32x2 vv = pk2(vx, vy);
#pragma unroll kPackedUnroll
            for (; j + 1 < je; j += 2) {
                // two candidates per trip; per candidate the operations and their order are those of the scalar
                // loop below (the dot product is formed unfused, like the oracle's)
                const f32x2 d0 = sub2(ld2(pos + j), pp), d1 = sub2(ld2(pos + j + 1), pp);
                const float2 s0 = unpk2(mul2(d0, d0)), s1 = unpk2(mul2(d1, d1));
                const float r20 = __fadd_rn(s0.x, s0.y), r21 = __fadd_rn(s1.x, s1.y);
                const float2 m0 = unpk2(mul2(sub2(v0p, ld2(vel + j)), d0)), m1 = unpk2(mul2(sub2(v0p, ld2(vel + j + 1)), d1));
                const f32x2 rs = pk2(rsqrt_approx(r20), rsqrt_approx(r21));
                const f32x2 u = mul2(pk2(__fadd_rn(m0.x, m0.y), __fadd_rn(m1.x, m1.y)), rs);
                const float2 uu = unpk2(u);
                const bool hit0 = (r20 <= h2) & (uu.x > 0.0f), hit1 = (r21 <= h2) & (uu.y > 0.0f);
                const f32x2 whdt = mul2(fma2(mul2(pk2(r20, r21), rs), nh2, one2), hdt2);
                f32x2 t = mul2(mul2(mul2(u, fma2(beta2, u, sigma2)), whdt), rs);
                if (STAB) {
                    const float g0 = fmaxf(gci, gamma * coupling[j]), g1 = fmaxf(gci, gamma * coupling[j + 1]);
                    t = mul2(t, pk2(g0 > 1.0f ? rcp_approx(g0) : 1.0f, g1 > 1.0f ? rcp_approx(g1) : 1.0f));
                }
                // a candidate that does not contribute gets t = 0: its clamped components are +-0, and subtracting
                // those changes nothing (one select per candidate instead of one per component)
                const float2 tt = unpk2(t);
                const float t0 = hit0 ? tt.x : 0.0f, t1 = hit1 ? tt.y : 0.0f;
                const float2 i0 = unpk2(mul2(d0, pk2(t0, t0))), i1 = unpk2(mul2(d1, pk2(t1, t1)));
                vv = sub2(vv, pk2(fminf(fmaxf(i0.x, -2.5f), 2.5f), fminf(fmaxf(i0.y, -2.5f), 2.5f)));
                vv = sub2(vv, pk2(fminf(fmaxf(i1.x, -2.5f), 2.5f), fminf(fmaxf(i1.y, -2.5f), 2.5f)));
            }
            { const float2 v2 = unpk2(vv); vx = v2.x; vy = v2.y; }
#pragma unroll 1
            for (; j < je; j++) {      // at most one left over
#else
#pragma unroll kGatherUnroll
            for (int j = R.b[d]; j < R.e[d]; j++) {
#endif
                const float2 q = pos[j];
                const float dx = q.x - p.x, dy = q.y - p.y;
                const float r2 = dist2(dx, dy);
                const bool in = r2 <= h2;                               // list membership
                const float2 vq = vel[j];                               // unconditional: cheaper than a predicated address
                const float rs = rsqrt_approx(r2);
                // gravity is the same on both sides of the difference (fluid.c:398-412, :446-449)
                const float u_in = ((v0.x - vq.x) * dx + (v0.y - vq.y) * dy) * rs;
                // fluid.c:451-462; the particle itself and any coincident neighbour have r2 == 0, rs = inf,
                // u = NaN, which fails u > 0 here exactly as the reference's 0/0 does
                const bool hit = in && u_in > 0.0f;
                // half the impulse, so that the +-5 clamp of checkVelocity becomes +-2.5 (exact scaling)
                const float whdt = fmaf(-r2 * rs, h_recip, 1.0f) * hdt;
                const float t = stab_scale<STAB>(u_in * fmaf(P.beta, u_in, P.sigma) * whdt * rs, gci, gamma, coupling, j);
                const float ix = fminf(fmaxf(t * dx, -2.5f), 2.5f);
                const float iy = fminf(fmaxf(t * dy, -2.5f), 2.5f);
                vx -= hit ? ix : 0.0f;
                vy -= hit ? iy : 0.0f;
            }
        }
#endif
        float2 np = make_float2(p.x + vx * dt, p.y + vy * dt);          // fluid.c:517-518
        np = boundary(np, P);
        pos_pred[i] = np;

        int extra = ghost ? SPH_KEY_EMIG : 0;                           // a ghost advanced between exchanges stays a ghost
        bool unsent = false;
        if (P.nranks > 1 && xstep && !ghost) {
            // identify_oob_particles: strict < start / > end (fluid.c:494-497)
            unsigned char *dst = nullptr;
            if (np.x < P.edge_start && P.has_left) dst = send_l;
            else if (np.x > P.edge_end && P.has_right) dst = send_r;
            if (dst) {
                int k = atomicAdd(&msg_hdr(dst)[0], 1);
                if (k < P.msg_cap) {
                    msg_a(dst)[k] = np;
                    msg_b(dst, P.msg_cap)[k] = p;                       // x_prev travels with the migrant
                    msg_u(dst, P.msg_cap)[k] = u;
                    extra = SPH_KEY_EMIG;
                } else {
                    // The message is full: this emigrant stays with this slab for now and tries again in the next
                    // exchange step.  Its position may lie outside the slab's new window (a slab parked outside the
                    // tank, controls.c:405-426, drains ALL its particles this way, one message capacity per exchange):
                    // it is binned into the nearest window cell instead of being dropped.  Counted: a particle that
                    // waits is stepped without its true neighbours, so the run is flagged, but nobody is lost.
                    atomicAdd(&counters[CN_MSG_OVER], 1);
                    unsent = true;
                }
            } else {
                // ghost layer: widened to halo_w and tested per side (communication.c:137-140 uses h, else-if)
                if (P.has_left && np.x - P.edge_start <= P.halo_w) {
                    int k = atomicAdd(&msg_hdr(send_l)[1], 1);
                    if (k < P.msg_cap) {
                        msg_hp(send_l, P.msg_cap)[k] = np; msg_hu(send_l, P.msg_cap)[k] = u;
                        if (P.one_x) msg_hq(send_l, P.msg_cap)[k] = p;
                    }
                    else atomicAdd(&counters[CN_MSG_OVER], 1);
                }
                if (P.has_right && P.edge_end - np.x <= P.halo_w) {
                    int k = atomicAdd(&msg_hdr(send_r)[1], 1);
                    if (k < P.msg_cap) {
                        msg_hp(send_r, P.msg_cap)[k] = np; msg_hu(send_r, P.msg_cap)[k] = u;
                        if (P.one_x) msg_hq(send_r, P.msg_cap)[k] = p;
                    }
                    else atomicAdd(&counters[CN_MSG_OVER], 1);
                }
            }
        }
        const bool keep = HOLD ? !ghost : unsent;                       // (HOLD steps never exchange: `unsent` stays false)
#if SPH_DEFER_ADVECT
        if (slot_i >= 0) t_slot[slot_i] = slot_v;                       // the previous particle's: its atomic is back by now
        slot_i = bin_position_deferred(i, np, extra, P, cnt, t_key, counters, tile_total, slot_v, keep) ? i : -1;
#else
        bin_position(i, np, extra, P, cnt, t_key, t_slot, counters, tile_total, keep);
#endif
#if SPH_PREFETCH & 1
        if (i + gstride < n) {
            cp_async_wait<0>();                                         // (issued a whole particle ago)
            const Rows Rn = candidate_rows(s_pos[buf][threadIdx.x], P, cell_start);
            prefetch_rows(Rn, pos);
            prefetch_rows(Rn, vel);
        }
#endif
    }
#if SPH_DEFER_ADVECT
    if (slot_i >= 0) t_slot[slot_i] = slot_v;
#endif
    pdl_done();
}

// -------------------------------------------------------------------------------------------
// K1s Stabilised viscosity gather (optional, sph_set_viscosity_stabilisation).
//     The reference applies its impulses pair by pair IN PLACE (fluid.c:442-472), so every pair sees the
//     velocities the pairs before it left behind and a pair's approach speed is damped by the factor
//     1 - dt (1-q)(sigma + beta u) without overshoot.  A gather sums a particle's impulses from FROZEN
//     velocities; with C_i = sum_j dt (1-q_ij)(sigma + beta u_ij) over its approaching pairs the velocity
//     changes by about C_i / 2 times the approach speed, which overshoots and feeds itself once C_i >> 1:
//     the "goo" preset (sigma 100, beta 10, controls.c:359-371; dt sigma = 0.83 per pair) never settles.
//     k_coupling forms C for every resident entry (ghosts included: the 2h ghost layer holds all their
//     neighbours), and k_advect_stab scales each pair's impulse by s_ij = 1 / max(1, gamma max(C_i, C_j)):
//     symmetric, so momentum is still exchanged pairwise; independent of the decomposition; exactly 1
//     wherever the plain gather is stable, so those results do not change.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS, SPH_BLOCKS_DENSITY)
k_coupling(const DevParams *__restrict__ Pp, const int *__restrict__ counters,
           const float2 *__restrict__ pos, const float2 *__restrict__ vel, const int *__restrict__ cell_start,
           float *__restrict__ coupling, const int *__restrict__ ckey)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    const float h_recip = __fdiv_rn(1.0f, P.h);
    const float h2 = __fmul_rn(P.h, P.h);
#if SPH_ROWS_FROM_KEY
    const float inv_wx = 1.0f / (float)P.wx;
#endif
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 p = pos[i];
        const float2 v0 = vel[i];
        float c = 0.0f;
#if SPH_ROWS_FROM_KEY
        const Rows R = candidate_rows_key(ckey[i], P, inv_wx, cell_start);
#if SPH_PIPE
        { const int nx = i + gridDim.x * blockDim.x; if (nx < n) { prefetch_l1(pos + nx); prefetch_l1(vel + nx); prefetch_l1(ckey + nx); } }
#endif
#else
        const Rows R = candidate_rows(p, P, cell_start);
#endif
#if SPH_PACKED
        const f32x2 pp = pk2(p.x, p.y), v0p = pk2(v0.x, v0.y);
        const f32x2 nh2 = pk2(-h_recip, -h_recip), one2 = pk2(1.0f, 1.0f), dt2 = pk2(P.dt, P.dt);
        const f32x2 beta2 = pk2(P.beta, P.beta), sigma2 = pk2(P.sigma, P.sigma);
#endif
#pragma unroll
        for (int d = 0; d < SPH_NROWS; d++) {
#if SPH_PACKED
            int j = R.b[d];
            const int je = R.e[d];
#pragma unroll kPackedUnroll
#if SPH_PAIRMASK
            for (; j < je; j += 2) {
                const bool v1 = j + 1 < je;
#else
            for (; j + 1 < je; j += 2) {
                const bool v1 = true;
#endif
                const f32x2 d0 = sub2(ld2(pos + j), pp), d1 = sub2(ld2(pos + j + 1), pp);
                const float2 s0 = unpk2(mul2(d0, d0)), s1 = unpk2(mul2(d1, d1));
                const float r20 = __fadd_rn(s0.x, s0.y), r21 = __fadd_rn(s1.x, s1.y);
                const float2 m0 = unpk2(mul2(sub2(v0p, ld2(vel + j)), d0)), m1 = unpk2(mul2(sub2(v0p, ld2(vel + j + 1)), d1));
                const f32x2 rs = pk2(rsqrt_approx(r20), rsqrt_approx(r21));
                const f32x2 u = mul2(pk2(__fadd_rn(m0.x, m0.y), __fadd_rn(m1.x, m1.y)), rs);
                const float2 uu = unpk2(u);
                const float2 cj = unpk2(mul2(mul2(fma2(mul2(pk2(r20, r21), rs), nh2, one2), dt2), fma2(beta2, u, sigma2)));
                c += (r20 <= h2 && uu.x > 0.0f) ? cj.x : 0.0f;
                c += (v1 && r21 <= h2 && uu.y > 0.0f) ? cj.y : 0.0f;
            }
#endif
#if !SPH_PAIRMASK
#if SPH_PACKED
#pragma unroll 1
            for (; j < je; j++) {      // at most one left over
#else
#pragma unroll kGatherUnroll
            for (int j = R.b[d]; j < R.e[d]; j++) {
#endif
                const float2 q = pos[j];
                const float dx = q.x - p.x, dy = q.y - p.y;
                const float r2 = dist2(dx, dy);
                const float2 vq = vel[j];
                const float rs = rsqrt_approx(r2);
                // as in k_advect: the particle itself (r2 == 0) gives u = NaN and fails u > 0
                const float u_in = ((v0.x - vq.x) * dx + (v0.y - vq.y) * dy) * rs;
                const bool hit = r2 <= h2 && u_in > 0.0f;
                const float wdt = fmaf(-r2 * rs, h_recip, 1.0f) * P.dt;
                const float cj = wdt * fmaf(P.beta, u_in, P.sigma);
                c += hit ? cj : 0.0f;
            }
#endif
        }
        coupling[i] = c;
    }
    pdl_done();
}

// -------------------------------------------------------------------------------------------
// K2  unpack neighbour messages into the source arrays and bin them (hash_halo, hash.c:51;
//     the receive side of transferOOBParticles, communication.c:340-358).  Slab mode only.
//     which == 0: migrants (pos, prev) become locals, ghosts carry a position only.
//     which == 1: ghosts carry position + velocity.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_unpack(const DevParams *__restrict__ Pp, int *__restrict__ counters, int which,
         unsigned char *send_l, unsigned char *send_r, unsigned char *recv_l, unsigned char *recv_r,
         float2 *__restrict__ src_pos, float2 *__restrict__ src_q, uint32_t *__restrict__ src_uid,
         int *__restrict__ cnt, int *__restrict__ t_key, int *__restrict__ t_slot, long long *__restrict__ xt,
         int *__restrict__ tile_total)
{
    pdl_enter();
    const DevParams P = *Pp;
    // block 0's view of the meeting, in ns of the global timer: xt[0] sending, [1] waiting for the neighbours,
    // [2] unpacking, [3] meetings (cumulative, sph_get_exchange_times); [4] when the last wait ended, [5] this slab's OWN
    // time since sph_copy_work last asked -- from the end of one wait to the end of the next send, i.e. everything it
    // did between two meetings, independent of how long it then waited -- and [6] its waits over the same span
    const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
    long long tc0 = 0, tc1 = 0, tc2 = 0;
    if (timer) tc0 = global_timer_ns();
    const int base = counters[CN_NTOT];
    int n_mig[2] = {0, 0}, n_halo[2] = {0, 0};
    unsigned char *buf[2] = {recv_l, recv_r};
    const int present[2] = {P.has_left, P.has_right};
    if (P.p2p) {
        const int step = counters[CN_STEP];
        send_messages(P, counters, which, step, send_l, send_r);
        if (timer) tc1 = global_timer_ns();
        // wait for the neighbours' messages of this exchange: they were stored into this rank's
        // exchange block by the neighbours' k_unpack; the flag is released after the payload
        __shared__ int s_ok[2];
        if (threadIdx.x < 2) {
            const int s = threadIdx.x;
            int ok = 1;
            if (present[s]) {
                const int *flag = xchg_flag(P.xchg_base, s, which);
                const long long t0 = clock64();
                while (ld_acquire_sys(flag) < 2 * step + which + 1) {
                    if (clock64() - t0 > P.spin_timeout) { ok = 0; break; }    // neighbour lost
                    __nanosleep(200);
                }
            }
            s_ok[s] = ok;
        }
        __syncthreads();
        if (timer) tc2 = global_timer_ns();
        for (int s = 0; s < 2; s++) {
            buf[s] = (unsigned char *)(P.xchg_base + xchg_offset(s, which, step & 1, P.msg_cap));
            if (!s_ok[s]) {      // treat the missing message as empty; the run is invalid and says so
                buf[s] = nullptr;
                if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&counters[CN_TIMEOUT_MSG], 1); atomicAdd(&counters[CN_MSG_OVER], 1); }
            }
        }
    }
    for (int s = 0; s < 2; s++) {
        if (!present[s] || !buf[s]) continue;
        // (plain L2 loads: the thread that saw the flag did the system-scope acquire, the barrier after it
        //  orders the rest of the block; a sys-scope acquire per thread here cost microseconds)
        n_mig[s] = which == 0 ? min(__ldcg(&msg_hdr(buf[s])[0]), P.msg_cap) : 0;
        n_halo[s] = min(__ldcg(&msg_hdr(buf[s])[1]), P.msg_cap);
    }
    const int total = n_mig[0] + n_halo[0] + n_mig[1] + n_halo[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int fit = min(total, max(P.cap - base, 0));
        counters[CN_EXTRA] = fit;
        if (fit < total) atomicAdd(&counters[CN_CAP_OVER], total - fit);
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int idx = base + t;
        if (idx >= P.cap) break;
        int k = t, s = 0, is_mig = 0;
        if (k < n_mig[0]) { s = 0; is_mig = 1; }
        else if ((k -= n_mig[0]) < n_halo[0]) { s = 0; }
        else if ((k -= n_halo[0]) < n_mig[1]) { s = 1; is_mig = 1; }
        else { k -= n_mig[1]; s = 1; }
        float2 p, q;
        uint32_t u;
        if (which == 0 && !is_mig) {
            p = msg_hp(buf[s], P.msg_cap)[k];
            q = P.one_x ? msg_hq(buf[s], P.msg_cap)[k]                  // the ghost will be relaxed here too
                        : make_float2(0.0f, 0.0f);
            u = msg_hu(buf[s], P.msg_cap)[k] | SPH_HALO_BIT;
        } else {
            p = msg_a(buf[s])[k];
            q = msg_b(buf[s], P.msg_cap)[k];
            u = msg_u(buf[s], P.msg_cap)[k];
            u = is_mig ? (u & SPH_UID_MASK) : (u | SPH_HALO_BIT);
        }
        src_pos[idx] = p;
        src_q[idx] = q;
        src_uid[idx] = u;
        // a ghost outside this slab's window is simply not needed (a slab parked outside the tank,
        // controls.c:405-426, still receives its neighbour's edge particles): same flag as an emigrant
        bin_position(idx, p, (u & SPH_HALO_BIT) ? SPH_KEY_EMIG : 0, P, cnt, t_key, t_slot, counters, tile_total);
    }
    if (timer && P.p2p) {
        const long long tc3 = global_timer_ns();
        xt[0] += tc1 - tc0; xt[1] += tc2 - tc1; xt[2] += tc3 - tc2; xt[3] += 1;
        if (xt[4] != 0) { xt[5] += tc1 - xt[4]; xt[6] += tc2 - tc1; }
        xt[4] = tc2;
    }
    pdl_done();
}

// -------------------------------------------------------------------------------------------
// K3  exclusive prefix sum of the cell populations -> cell_start, in two plain streaming kernels:
//       k_scan_totals  one total per tile of 8192 cells (also the bucket statistics)
//       k_scan_apply   tile offset = sum of the preceding tile totals (read coalesced, block-reduced),
//                      then the tile's own scan; clears the populations for the next sort and
//                      publishes the new entry counts.
//     The first versions did this in one kernel with inter-block waiting (chained look-back, then
//     "sum all predecessors"); both were latency-bound at ~18-20 us for 8 MB because every tile sat in
//     a dependent chain of L2 round trips and barriers (profiles/r1_div2_full.csv).  Two independent
//     passes have no waiting at all and the second read of the populations hits L2.
//     The reference silently drops particles above 100 per bucket (hash.c:160-165): the largest
//     population seen is recorded so that this can be detected.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void scan_load_tile(const int *__restrict__ cnt, int base, int ncell, int (&v)[SCAN_ITEMS])
{
    if (base + SCAN_ITEMS <= ncell) {
        const int4 *src = reinterpret_cast<const int4 *>(cnt + base);     // 32 consecutive cells = 128 bytes per thread
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            const int4 q = src[k];
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = base + k < ncell ? cnt[base + k] : 0;
    }
}

#if !SPH_TILE_ATOMICS
__global__ void __launch_bounds__(SPH_THREADS)
k_scan_totals(const DevParams *__restrict__ Pp, int *__restrict__ counters, const int *__restrict__ cnt,
              int *__restrict__ tile_total)
{
    pdl_enter();
    __shared__ int s_sum[SPH_THREADS / 32], s_max[SPH_THREADS / 32], s_over[SPH_THREADS / 32];
    const int ncell = Pp->wx_new * Pp->sort_rows;
    const int ntiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int v[SCAN_ITEMS];
        scan_load_tile(cnt, tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS, ncell, v);
        int sum = 0, mx = 0, over = 0;
#pragma unroll
        // (per sort-grid sub-cell: a sub-cell above 100 implies its reference bucket is; the exact bucket
        //  statistics are computed on demand by k_bucket_stats)
        for (int k = 0; k < SCAN_ITEMS; k++) { sum += v[k]; mx = max(mx, v[k]); over += v[k] > 100; }
        sum = __reduce_add_sync(0xffffffffu, sum);
        mx = __reduce_max_sync(0xffffffffu, mx);
        over = __reduce_add_sync(0xffffffffu, over);
        if (lane == 0) { s_sum[warp] = sum; s_max[warp] = mx; s_over[warp] = over; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0, m = 0, o = 0;
#pragma unroll
            for (int w = 0; w < SPH_THREADS / 32; w++) { t += s_sum[w]; m = max(m, s_max[w]); o += s_over[w]; }
            tile_total[tile] = t;
            if (m > 0) atomicMax(&counters[CN_MAX_BUCKET], m);
            if (o > 0) atomicAdd(&counters[CN_BUCKET_OVER], o);
        }
        __syncthreads();
    }
    pdl_done();
}

#endif

#if !(SPH_SCAN_FAST && SPH_TILE_ATOMICS)
__global__ void __launch_bounds__(SPH_THREADS)
k_scan_apply(DevParams *__restrict__ Pp, int *__restrict__ counters, int *__restrict__ cnt, int *__restrict__ cell_start,
             const int *__restrict__ tile_total, unsigned char *send_l, unsigned char *send_r, int end_of_step)
{
    pdl_enter();
    __shared__ int s_warp[SPH_THREADS / 32];
    __shared__ int s_prefix;
    const int ncell = Pp->wx_new * Pp->sort_rows;
    const int ntiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // offset of this tile: all preceding totals, coalesced, block-reduced
        int part = 0;
        for (int idx = threadIdx.x; idx < tile; idx += SPH_THREADS) part += tile_total[idx];
        part = __reduce_add_sync(0xffffffffu, part);
        if (lane == 0) s_warp[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < SPH_THREADS / 32; w++) t += s_warp[w];
            s_prefix = t;
        }
        __syncthreads();
        const int prefix = s_prefix;
        __syncthreads();                                   // s_warp is reused below

        const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        int v[SCAN_ITEMS];
        scan_load_tile(cnt, base, ncell, v);
        int sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) sum += v[k];
#if SPH_TILE_ATOMICS
        {
            // the bucket statistics k_scan_totals used to form (per sort-grid sub-cell: a sub-cell above 100 implies its
            // reference bucket is; the exact bucket statistics are computed on demand by k_bucket_stats)
            int mx = 0, over = 0;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; k++) { mx = max(mx, v[k]); over += v[k] > 100; }
            mx = __reduce_max_sync(0xffffffffu, mx);
            over = __reduce_add_sync(0xffffffffu, over);
            if (lane == 0 && mx > 0) atomicMax(&counters[CN_MAX_BUCKET], mx);
            if (lane == 0 && over > 0) atomicAdd(&counters[CN_BUCKET_OVER], over);
        }
#endif
        int inc = sum;                                     // warp inclusive scan of the thread sums
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int warp_off = 0, block_total = 0;
#pragma unroll
        for (int w = 0; w < SPH_THREADS / 32; w++) {
            const int sw = s_warp[w];
            if (w < warp) warp_off += sw;
            block_total += sw;
        }
        int run = prefix + warp_off + (inc - sum);
        if (base + SCAN_ITEMS <= ncell) {
            int4 *dst = reinterpret_cast<int4 *>(cell_start + base);
            int4 *zero = reinterpret_cast<int4 *>(cnt + base);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS / 4; k++) {
                int4 o;
                o.x = run; run += v[4 * k];
                o.y = run; run += v[4 * k + 1];
                o.z = run; run += v[4 * k + 2];
                o.w = run; run += v[4 * k + 3];
                dst[k] = o;
                zero[k] = make_int4(0, 0, 0, 0);
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; k++) {
                const int idx = base + k;
                if (idx < ncell) { cell_start[idx] = run; cnt[idx] = 0; }
                run += v[k];
            }
        }
        if (tile == ntiles - 1 && threadIdx.x == SPH_THREADS - 1) {
            const int total = prefix + block_total;
            cell_start[ncell] = total;
            counters[CN_NSRC] = counters[CN_NTOT] + counters[CN_EXTRA];
            counters[CN_EXTRA] = 0;
            counters[CN_NTOT] = total;
            counters[CN_NLOCAL] = 0;
            // the sorted arrays are about to be rebuilt for the new window
            Pp->gx0 = Pp->gx0_new;
            Pp->wx = Pp->wx_new;
            // send buffers are free again: the messages they held were consumed before this sort
            if (send_l) { msg_hdr(send_l)[0] = 0; msg_hdr(send_l)[1] = 0; }
            if (send_r) { msg_hdr(send_r)[0] = 0; msg_hdr(send_r)[1] = 0; }
            if (end_of_step) counters[CN_STEP] += 1;
        }
        __syncthreads();
    }
    pdl_done();
}

#endif

#if SPH_SCAN_FAST
// K3' (SPH_SCAN_FAST): see the flag's comment at the top.  Same arguments and results as k_scan_apply.
__global__ void __launch_bounds__(SPH_THREADS)
k_scan_apply_fast(DevParams *__restrict__ Pp, int *__restrict__ counters, int *__restrict__ cnt, int *__restrict__ cell_start,
                  const int *__restrict__ tile_total, unsigned char *send_l, unsigned char *send_r, int end_of_step)
{
    pdl_enter();
    constexpr int NW = SPH_THREADS / 32;
    __shared__ int s_inc[NW], s_part[NW], s_max[NW], s_over[NW];
    const int ncell = Pp->wx_new * Pp->sort_rows;
    const int ntiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        // the tile's own total counts exactly the increments its cells received (bin_position adds to both): zero means
        // every population of the tile is zero already
        const bool occupied = tile_total[tile] != 0;
        int v[SCAN_ITEMS];
        if (occupied) scan_load_tile(cnt, base, ncell, v);
        else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; k++) v[k] = 0;
        }
        int part = 0;                                      // all preceding totals, coalesced
        for (int idx = threadIdx.x; idx < tile; idx += SPH_THREADS) part += tile_total[idx];
        int sum = 0, mx = 0, over = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { sum += v[k]; mx = max(mx, v[k]); over += v[k] > 100; }
        int inc = sum;                                     // warp inclusive scan of the thread sums
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        part = __reduce_add_sync(0xffffffffu, part);
        mx = __reduce_max_sync(0xffffffffu, mx);
        over = __reduce_add_sync(0xffffffffu, over);
        if (lane == 31) s_inc[warp] = inc;
        if (lane == 0) { s_part[warp] = part; s_max[warp] = mx; s_over[warp] = over; }
        __syncthreads();
        int prefix = 0, warp_off = 0, block_total = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            prefix += s_part[w];
            const int sw = s_inc[w];
            if (w < warp) warp_off += sw;
            block_total += sw;
        }
        if (threadIdx.x == 0 && occupied) {
            int m = 0, o = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) { m = max(m, s_max[w]); o += s_over[w]; }
            if (m > 0) atomicMax(&counters[CN_MAX_BUCKET], m);
            if (o > 0) atomicAdd(&counters[CN_BUCKET_OVER], o);
        }
        int run = prefix + warp_off + (inc - sum);
        if (base + SCAN_ITEMS <= ncell) {
            int4 *dst = reinterpret_cast<int4 *>(cell_start + base);
            int4 *zero = reinterpret_cast<int4 *>(cnt + base);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS / 4; k++) {
                int4 o;
                o.x = run; run += v[4 * k];
                o.y = run; run += v[4 * k + 1];
                o.z = run; run += v[4 * k + 2];
                o.w = run; run += v[4 * k + 3];
                dst[k] = o;
                if (occupied) zero[k] = make_int4(0, 0, 0, 0);
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; k++) {
                const int idx = base + k;
                if (idx < ncell) { cell_start[idx] = run; if (occupied) cnt[idx] = 0; }
                run += v[k];
            }
        }
        if (tile == ntiles - 1 && threadIdx.x == SPH_THREADS - 1) {
            const int total = prefix + block_total;
            cell_start[ncell] = total;
            counters[CN_NSRC] = counters[CN_NTOT] + counters[CN_EXTRA];
            counters[CN_EXTRA] = 0;
            counters[CN_NTOT] = total;
            counters[CN_NLOCAL] = 0;
            Pp->gx0 = Pp->gx0_new;
            Pp->wx = Pp->wx_new;
            if (send_l) { msg_hdr(send_l)[0] = 0; msg_hdr(send_l)[1] = 0; }
            if (send_r) { msg_hdr(send_r)[0] = 0; msg_hdr(send_r)[1] = 0; }
            if (end_of_step) counters[CN_STEP] += 1;
        }
        __syncthreads();                                   // the shared arrays are reused by the block's next tile
    }
    pdl_done();
}
#endif

#if SPH_SORT_SRC
// entries per thread and trip of the two kernels below: every load of a level is issued for all of them before the
// first is looked at (the kernels are bound by the latency of three dependent levels, not by bytes)
#ifndef SPH_SORT_ITEMS
#define SPH_SORT_ITEMS 4          // (1 / 2 / 4 / 8 measured: profiles/r2_variants.md)
#endif
// -------------------------------------------------------------------------------------------
// K4' (SPH_SORT_SRC) a source entry's uid into its cell's range, at its arrival slot.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_scatter_uid(const int *__restrict__ counters, const int *__restrict__ cell_start,
              const int *__restrict__ t_key, const int *__restrict__ t_slot, const uint32_t *__restrict__ src_uid,
              uint32_t *__restrict__ ord_uid, int *__restrict__ tile_total, int ntiles_max)
{
    pdl_enter();
    constexpr int K = SPH_SORT_ITEMS;
    const int n = counters[CN_NSRC];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
#if SPH_TILE_ATOMICS
    for (int t = gtid; t < ntiles_max; t += gstride) tile_total[t] = 0;
#endif
    for (int s0 = gtid; s0 < n; s0 += K * gstride) {
        // level 1: 3 K independent coalesced loads; level 2: K cell offsets; then K stores nobody waits for
        int key[K], slot[K], b[K];
        uint32_t u[K];
        bool on[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int s = s0 + k * gstride;
            key[k] = s < n ? t_key[s] : SPH_KEY_DROP;
            u[k] = s < n ? src_uid[s] : 0u;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            on[k] = key[k] != SPH_KEY_DROP;
            slot[k] = on[k] ? t_slot[s0 + k * gstride] : 0;
        }
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = on[k] ? cell_start[key[k] & SPH_KEY_MASK] : 0;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (on[k]) ord_uid[b[k] + slot[k]] = u[k] & SPH_UID_MASK;
    }
    pdl_done();
}

// -------------------------------------------------------------------------------------------
// K5' (SPH_SORT_SRC) canonical order inside each cell (ascending uid, as K5) and the physical reorder, walking the
//     source order: key, uid and payload of an entry are coalesced first-level loads.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ int rank_in_cell(const uint32_t *__restrict__ ord_uid, int b, int e, uint32_t um)
{
    int rank = 0;
    if (e - b > 1)                                         // alone in its cell: nothing to look at
        for (int k = b; k < e; k++) rank += ord_uid[k] < um;
    return rank;
}

__global__ void __launch_bounds__(SPH_THREADS)
k_reorder_src(const DevParams *__restrict__ Pp, int *__restrict__ counters, const int *__restrict__ cell_start,
              const int *__restrict__ t_key, const uint32_t *__restrict__ ord_uid, const uint32_t *__restrict__ src_uid,
              const float2 *__restrict__ src_pos, const float2 *__restrict__ src_q,
              float2 *__restrict__ dst_pos, float2 *__restrict__ dst_q, uint32_t *__restrict__ dst_uid, int *__restrict__ dst_key
              SPH_PV4_PARAM)
{
    pdl_enter();
    constexpr int K = SPH_SORT_ITEMS;
    const int n = counters[CN_NSRC];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    int locals = 0;
    for (int s0 = gtid; s0 < n; s0 += K * gstride) {
        int key[K], b[K], e[K];
        uint32_t u[K];
        float2 p[K], q[K];
        bool on[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int s = s0 + k * gstride;
            key[k] = s < n ? t_key[s] : SPH_KEY_DROP;
            u[k] = s < n ? src_uid[s] : 0u;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int s = s0 + k * gstride;
            on[k] = key[k] != SPH_KEY_DROP;
            p[k] = q[k] = make_float2(0.0f, 0.0f);
            b[k] = e[k] = 0;
            if (on[k]) {
                p[k] = src_pos[s]; q[k] = src_q[s];
                const int c = key[k] & SPH_KEY_MASK;
                b[k] = cell_start[c]; e[k] = cell_start[c + 1];
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (!on[k]) continue;
            if (key[k] & SPH_KEY_EMIG) u[k] |= SPH_HALO_BIT;
            const int dst = b[k] + rank_in_cell(ord_uid, b[k], e[k], u[k] & SPH_UID_MASK);
            dst_pos[dst] = p[k]; dst_q[dst] = q[k]; dst_uid[dst] = u[k];
#if SPH_ADVECT_PV4
            if (pv) pv[dst] = make_float4(p[k].x, p[k].y, q[k].x, q[k].y);
#endif
#if SPH_ROWS_FROM_KEY
            dst_key[dst] = key[k] & SPH_KEY_MASK;
#endif
            locals += !(u[k] & SPH_HALO_BIT);
        }
    }
    pdl_done();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) locals += __shfl_xor_sync(0xffffffffu, locals, o);
    if ((threadIdx.x & 31) == 0 && locals) atomicAdd(&counters[CN_NLOCAL], locals);
}
#endif

#if !SPH_SORT_SRC
// -------------------------------------------------------------------------------------------
// K4  scatter (uid, source index) to cell_start[key] + arrival slot.  Arrival order inside a
//     cell is whatever the atomics produced; K5 makes it canonical.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_scatter(const int *__restrict__ counters, const int *__restrict__ cell_start,
          const int *__restrict__ t_key, const int *__restrict__ t_slot, const uint32_t *__restrict__ src_uid,
          uint32_t *__restrict__ ord_uid, int *__restrict__ ord_src, int *__restrict__ ord_key,
          int *__restrict__ tile_total, int ntiles_max)
{
    pdl_enter();
    const int n = counters[CN_NSRC];
#if SPH_TILE_ATOMICS
    // the scan before this kernel has consumed the tile totals: clear them for the producers of the next sort
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles_max; t += gridDim.x * blockDim.x) tile_total[t] = 0;
#endif
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const int key = t_key[s];
        if (key == SPH_KEY_DROP) continue;
        const int d = cell_start[key & SPH_KEY_MASK] + t_slot[s];
        uint32_t u = src_uid[s];
        if (key & SPH_KEY_EMIG) u |= SPH_HALO_BIT;
        ord_uid[d] = u;
        ord_src[d] = s;
        ord_key[d] = key & SPH_KEY_MASK;
    }
    pdl_done();
}

// -------------------------------------------------------------------------------------------
// K5  canonical order inside each cell (ascending uid == the reference's bucket order on one
//     rank, hash.c:160-163 inserts in pointer order) and the physical reorder of the payload.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_reorder(const DevParams *__restrict__ Pp, int *__restrict__ counters, const int *__restrict__ cell_start,
          const int *__restrict__ ord_key, const uint32_t *__restrict__ ord_uid, const int *__restrict__ ord_src,
          const float2 *__restrict__ src_pos, const float2 *__restrict__ src_q,
          float2 *__restrict__ dst_pos, float2 *__restrict__ dst_q, uint32_t *__restrict__ dst_uid)
{
    pdl_enter();
    const int n = counters[CN_NTOT];
    int locals = 0;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
        // three levels of dependent loads instead of five: (src, uid, key) | (payload, cell bounds) | cell uids
        const int s = ord_src[d];
        const uint32_t u = ord_uid[d];
        const int key = ord_key[d];
        const float2 pp = src_pos[s], qq = src_q[s];
        const int b = cell_start[key], e = cell_start[key + 1];
        const uint32_t um = u & SPH_UID_MASK;
        int rank = 0;
        for (int k = b; k < e; k++) rank += (ord_uid[k] & SPH_UID_MASK) < um;
        const int dst = b + rank;
        dst_pos[dst] = pp;
        dst_q[dst] = qq;
        dst_uid[dst] = u;
        locals += !(u & SPH_HALO_BIT);
    }
    pdl_done();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) locals += __shfl_xor_sync(0xffffffffu, locals, o);
    if ((threadIdx.x & 31) == 0 && locals) atomicAdd(&counters[CN_NLOCAL], locals);
}

#endif      // !SPH_SORT_SRC

// -------------------------------------------------------------------------------------------
// K6  calculate_density (fluid.c:527-539) over every pair within h, as a gather.
//     Output: (density, density_near) per resident entry, ghosts included (their pressure is
//     needed by the relaxation of the locals next to them, fluid.c:560-565).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS, SPH_BLOCKS_DENSITY)
k_density(const DevParams *__restrict__ Pp, int *__restrict__ counters,
          const float2 *__restrict__ pos, const int *__restrict__ cell_start, float2 *__restrict__ dens,
          sph_mask_t *__restrict__ nmask, const int *__restrict__ ckey SPH_PD4_PARAM)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    const float h_recip = __fdiv_rn(1.0f, P.h);
    const float h2 = __fmul_rn(P.h, P.h);
    int cost = 0;
#if SPH_ROWS_FROM_KEY
    const float inv_wx = 1.0f / (float)P.wx;
#endif
#if SPH_ASYNC & 4
    __shared__ float2 s_pos[2][SPH_THREADS];
    const int gstride = gridDim.x * blockDim.x;
    int buf = 0;
    {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < n) cp_async8(&s_pos[0][threadIdx.x], pos + i0);
        cp_async_commit();
    }
#endif
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#if SPH_ASYNC & 4
        {
            const int nx = i + gstride;
            if (nx < n) cp_async8(&s_pos[buf ^ 1][threadIdx.x], pos + nx);
            cp_async_commit();
            cp_async_wait<1>();
        }
        const float2 p = s_pos[buf][threadIdx.x];
        buf ^= 1;
#else
        const float2 p = pos[i];
#endif
        float d = 0.0f, dn = 0.0f;
        int nn = 0;
#if SPH_ROWS_FROM_KEY
        const Rows R = candidate_rows_key(ckey[i], P, inv_wx, cell_start);
#if SPH_PIPE
        { const int nx = i + gridDim.x * blockDim.x; if (nx < n) { prefetch_l1(pos + nx); prefetch_l1(ckey + nx); } }
#endif
#else
        const Rows R = candidate_rows(p, P, cell_start);
#endif
#if SPH_PACKED
        const f32x2 pp = pk2(p.x, p.y), nh2 = pk2(-h_recip, -h_recip), one2 = pk2(1.0f, 1.0f);
#endif
#pragma unroll
        for (int dd = 0; dd < SPH_NROWS; dd++) {
            // acceptance mask of this row's first SPH_MASK_BITS candidates: k_relax works on the same
            // positions and the same ranges, so it walks these bits instead of repeating the distance tests
            sph_mask_t m = 0;
            const int b = R.b[dd], e = R.e[dd];
            // the particle itself sits in its own row's range: walk [b, i) and (i, e) instead of paying a
            // self test on every candidate (coincident OTHER particles do count: ratio 0)
            const int self = (dd == SPH_CELL_DIV && i >= b && i < e) ? i : e;
            for (int seg = 0; seg < 2; seg++) {
                const int jb = seg == 0 ? b : self + 1, je = seg == 0 ? self : e;
                sph_mask_t bit = jb - b < SPH_MASK_BITS ? (sph_mask_t)1 << (jb - b) : 0;   // shifts out after the last mask bit
                // Branch-free body: a candidate outside h adds an exact +0 (the weight is selected to 0),
                // so the value and the order of the sums are those of the gated loop, but no lane ever
                // waits for another lane's accept path.  With the branch the warp ran the accept path
                // on nearly every trip with half its lanes idle (profiles/r1_final_full.csv: 23 of 32).
#if SPH_PACKED
                int j = jb;
#pragma unroll kPackedUnroll
#if SPH_PAIRMASK
                for (; j < je; j += 2) {
                    const bool v1 = j + 1 < je;
#else
                for (; j + 1 < je; j += 2) {
                    const bool v1 = true;
#endif
                    const f32x2 d0 = sub2(ld2(pos + j), pp), d1 = sub2(ld2(pos + j + 1), pp);
                    const float2 s0 = unpk2(mul2(d0, d0)), s1 = unpk2(mul2(d1, d1));
                    const float r20 = __fadd_rn(s0.x, s0.y), r21 = __fadd_rn(s1.x, s1.y);
                    const bool in0 = r20 <= h2, in1 = v1 & (r21 <= h2);
                    m |= in0 ? bit : (sph_mask_t)0;
                    bit += bit;
                    m |= in1 ? bit : (sph_mask_t)0;
                    bit += bit;
                    const float2 w = unpk2(fma2(pk2(sqrt_approx(r20), sqrt_approx(r21)), nh2, one2));
                    const float w0 = in0 ? fmaxf(w.x, 0.0f) : 0.0f, w1 = in1 ? fmaxf(w.y, 0.0f) : 0.0f;
                    const float2 ww = unpk2(mul2(pk2(w0, w1), pk2(w0, w1)));
                    d += ww.x;
                    dn = fmaf(ww.x, w0, dn);
                    d += ww.y;
                    dn = fmaf(ww.y, w1, dn);
                }
#endif
#if !SPH_PAIRMASK
#if SPH_PACKED
#pragma unroll 1
                for (; j < je; j++) {      // at most one left over
#else
#pragma unroll kGatherUnroll
                for (int j = jb; j < je; j++) {
#endif
                    const float2 q = pos[j];
                    const float dx = q.x - p.x, dy = q.y - p.y;
                    const float r2 = dist2(dx, dy);
                    const bool in = r2 <= h2;                           // list membership (hash.c:185,221)
                    m |= in ? bit : (sph_mask_t)0;
                    bit += bit;
                    // fluid.c:527-539 with r from one MUFU.SQRT (only r is needed here; sqrt(0) = 0, so a
                    // coincident neighbour counts with ratio 0 like in the reference)
                    float w = fmaxf(fmaf(-sqrt_approx(r2), h_recip, 1.0f), 0.0f);   // ratio < 1 gate: (1 - ratio)+
                    w = in ? w : 0.0f;
                    const float w2 = w * w;
                    d += w2;
                    dn = fmaf(w2, w, dn);
                }
#endif
            }
            nmask[(size_t)dd * P.cap + i] = m;
            nn += sph_mask_popc(m) + max(e - b - SPH_MASK_BITS, 0);     // candidates past the mask count as accepted
        }
        dens[i] = make_float2(d, dn);
#if SPH_RELAX_PD4
        pd[i] = make_float4(p.x, p.y, d, dn);
#endif
        // The reference's 400-entry cap is on a particle's FORWARD list (hash.c:188,223).  A forward list cannot exceed
        // the full neighbour count, and candidates past a row's mask were counted as accepted: only a particle that
        // fails this cheap, conservative test is counted again exactly, with the reference's owner rule (rare path).
        if (nn > SPH_REF_MAX_NEIGHBORS) atomicAdd(&counters[CN_NEIGH_OVER], 1);      // (sph_get_status then counts exactly)
        cost += SPH_COST_BASE + nn;
#if SPH_PREFETCH & 4
        if (i + gstride < n) {
            cp_async_wait<0>();
            prefetch_rows(candidate_rows(s_pos[buf][threadIdx.x], P, cell_start), pos);
        }
#endif
    }
    pdl_done();
    // work estimate of this slab (one atomic per warp per launch): input of the cost-based edge policy
    cost = __reduce_add_sync(0xffffffffu, cost);
    if ((threadIdx.x & 31) == 0 && cost) atomicAdd(&counters[CN_COST], cost);
}

// -------------------------------------------------------------------------------------------
// K7  double_density_relaxation (fluid.c:541-611) as a gather + updateVelocities (:642-653)
//     incl. boundaryConditions + second ghost selection (fluid.c:337) + binning for the re-hash
//     (fluid.c:341).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS, SPH_BLOCKS_RELAX)
k_relax(const DevParams *__restrict__ Pp, int *__restrict__ counters,
        const float2 *__restrict__ pos, const float2 *__restrict__ prev, const uint32_t *__restrict__ uid,
        const float2 *__restrict__ dens, const int *__restrict__ cell_start,
        const sph_mask_t *__restrict__ nmask,
        float2 *__restrict__ pos_out, float2 *__restrict__ vel_out,
        int *__restrict__ cnt, int *__restrict__ t_key, int *__restrict__ t_slot,
        unsigned char *send_l, unsigned char *send_r, const int *__restrict__ ckey, int *__restrict__ tile_total SPH_PD4_CPARAM)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
#if SPH_ROWS_FROM_KEY
    const float inv_wx = 1.0f / (float)P.wx;
#endif
#if SPH_DEFER_RELAX
    int slot_i = -1, slot_v = 0;                                        // arrival slot whose store is still owed
#endif
    const float dt = P.dt, dt2 = dt * dt;
    const float h = P.h;
    const float h_recip = __fdiv_rn(1.0f, h);
    const float h2 = __fmul_rn(h, h);
    const float K1 = dt2 * P.k, K2 = dt2 * P.k_near, Cs = dt2 * P.k_spring * h * 0.5f;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CN_MAX_BUCKET] = 0;

#if SPH_ASYNC & 2
    // staging slots of this thread: [buffer][thread]; everything k_relax reads once per particle
    __shared__ uint32_t s_uid[2][SPH_THREADS];
    __shared__ float2 s_pos[2][SPH_THREADS], s_dens[2][SPH_THREADS], s_prev[2][SPH_THREADS];
    __shared__ sph_mask_t s_mask[2][SPH_NROWS][SPH_THREADS];
#if SPH_KEYROWS
    __shared__ int s_key[2][SPH_THREADS];
#endif
    static_assert(sizeof(sph_mask_t) == 4, "the staging copies move 4-byte masks");
    const int gstride = gridDim.x * blockDim.x;
    int buf = 0;
    auto stage = [&](int b, int k) {
        cp_async4(&s_uid[b][threadIdx.x], uid + k);
        cp_async8(&s_pos[b][threadIdx.x], pos + k);
        cp_async8(&s_dens[b][threadIdx.x], dens + k);
        cp_async8(&s_prev[b][threadIdx.x], prev + k);
#if SPH_KEYROWS
        cp_async4(&s_key[b][threadIdx.x], ckey + k);
#endif
#pragma unroll
        for (int d = 0; d < SPH_NROWS; d++) cp_async4(&s_mask[b][d][threadIdx.x], nmask + (size_t)d * P.cap + k);
    };
    {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < n) stage(0, i0);
        cp_async_commit();
    }
#endif
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#if SPH_ASYNC & 2
        {
            const int nx = i + gstride;
            if (nx < n) stage(buf ^ 1, nx);
            cp_async_commit();
            cp_async_wait<1>();                                         // this particle's copies (the group before) have landed
        }
        const int cur = buf;
        buf ^= 1;
        const uint32_t u = s_uid[cur][threadIdx.x];
        const float2 p = s_pos[cur][threadIdx.x];
        const float2 di = s_dens[cur][threadIdx.x];
#if SPH_KEYROWS && !SPH_PIPE
        const int key_i = s_key[cur][threadIdx.x];
#endif
#define SPH_ROW_MASK(d) s_mask[cur][d][threadIdx.x]
#define SPH_PREV_I s_prev[cur][threadIdx.x]
#else
        const uint32_t u = uid[i];
#if SPH_KEYROWS && !SPH_PIPE
        const int key_i = ckey[i];
#endif
#define SPH_ROW_MASK(d) nmask[(size_t)(d) * P.cap + i]
#define SPH_PREV_I prev[i]
#endif
#if SPH_PIPE
        const float2 p = pos[i];
        const float2 di = dens[i];
        const int key_i = ckey[i];
#endif
        // one-exchange mode: a ghost is relaxed redundantly and kept as a ghost for the coming k_advect
        const bool ghost = (u & SPH_HALO_BIT) != 0;
        if (ghost && !P.one_x) { t_key[i] = SPH_KEY_DROP; continue; }
#if !SPH_PIPE
#if !(SPH_ASYNC & 2)
        const float2 p = pos[i];
        const float2 di = dens[i];
#endif
#else
        {
            // this particle's late inputs (the masks of the rows after the first, its previous position) and the
            // next particle's early ones, all towards L1 now
#pragma unroll
            for (int d = 1; d < SPH_NROWS; d++) prefetch_l1(nmask + (size_t)d * P.cap + i);
            prefetch_l1(prev + i);
            const int nx = i + gridDim.x * blockDim.x;
            if (nx < n) {
                prefetch_l1(uid + nx); prefetch_l1(pos + nx); prefetch_l1(dens + nx); prefetch_l1(ckey + nx);
#pragma unroll
                for (int d = 0; d < SPH_NROWS; d++) prefetch_l1(nmask + (size_t)d * P.cap + nx);
            }
        }
#endif
        // fluid.c:563-564 + :591 with the constants folded.  With w = 1 - r/h the spring term is
        // k_spring*(h - r)/2 = k_spring*h*w/2, so
        //   D = dt^2 ((P_i+P_j) w + (Pn_i+Pn_j) w^2 + k_spring (h-r)/2) = w (A + B w),
        //   A = K1 (rho_i + rho_j - 2 rho0) + C,  B = K2 (rhon_i + rhon_j),
        //   K1 = dt^2 k,  K2 = dt^2 k_near,  C = dt^2 k_spring h / 2
        // and a pair costs two FFMAs for A and B instead of forming four pressures.
        const float Ai = fmaf(K1, di.x - 2.0f * P.rest_density, Cs);
        const float Bi = K2 * di.y;
        float x = p.x, y = p.y;
#if SPH_ROWS_FROM_KEY
        // (the reference cell of this particle is only needed by the coincident-pair rule: formed there)
#define SPH_GXI cell_coord(p.x, P.cell_h)
#define SPH_GYI cell_coord(p.y, P.cell_h)
        const Rows R = candidate_rows_key(key_i, P, inv_wx, cell_start);
#else
        const int gxi = cell_coord(p.x, P.cell_h), gyi = cell_coord(p.y, P.cell_h);
#define SPH_GXI gxi
#define SPH_GYI gyi
        const Rows R = candidate_rows(p, P, cell_start);
#endif
        // pair physics for one listed neighbour (membership r2 <= h2, j != i already established)
#if SPH_PACKED_RELAX
        // packed build: (x, y) accumulate in one 64-bit register; the same rn operations in the same order
        const f32x2 pp = pk2(p.x, p.y), K12 = pk2(K1, K2), AB0 = pk2(Ai, Bi);
        f32x2 xy = pp;
        auto pair = [&](int j, float2 q, float2 dj) {
            const f32x2 dd = sub2(pk2(q.x, q.y), pp);
            const float2 sq = unpk2(mul2(dd, dd));
            const float r2 = __fadd_rn(sq.x, sq.y);
            const float2 ab = unpk2(fma2(pk2(dj.x, dj.y), K12, AB0));
            const float A = ab.x, B = ab.y;
            if (r2 <= 1.0001e-12f) {
                // (nearly) coincident particles, rare: as in the scalar build below
                const float2 d1 = unpk2(dd);
                float2 a = unpk2(xy);
                const float r = __fsqrt_rn(r2);
                if (r <= 0.000001f) {
                    const int gxj = cell_coord(q.x, P.cell_h), gyj = cell_coord(q.y, P.cell_h);
                    const int cxi = SPH_GXI, cyi = SPH_GYI;
                    const bool owner = (cxi == gxj && cyi == gyj) ? ((u & SPH_UID_MASK) < (uid[j] & SPH_UID_MASK))
                                                                  : (cxi != gxj ? cxi < gxj : cyi < gyj);
                    if (owner) { a.x += 0.000001f; a.y += 0.000001f; }
                }
                const float ratio = r * h_recip;
                if (ratio < 1.0f && r > 0.0f) {
                    const float w = 1.0f - ratio;
                    const float s = __fdiv_rn(fmaf(B, w, A) * w, r);
                    a.x = fmaf(-s, d1.x, a.x);
                    a.y = fmaf(-s, d1.y, a.y);
                }
                xy = pk2(a.x, a.y);
                return;
            }
            const float rs = rsqrt_approx(r2);
            const float w = fmaxf(fmaf(-r2 * rs, h_recip, 1.0f), 0.0f);
            const float s = fmaf(B, w, A) * w * rs;
            xy = fma2(pk2(-s, -s), dd, xy);
        };
#else
        auto pair = [&](int j, float2 q, float2 dj) {
            const float dx = q.x - p.x, dy = q.y - p.y;
            const float r2 = dist2(dx, dy);
            const float A = fmaf(dj.x, K1, Ai), B = fmaf(dj.y, K2, Bi);
            if (r2 <= 1.0001e-12f) {
                // (nearly) coincident particles, rare: the reference's tests as written, with an IEEE r
                const float r = __fsqrt_rn(r2);
                if (r <= 0.000001f) {
                    // only the list owner is nudged (fluid.c:583-586); owner = earlier bucket slot in
                    // the same cell, else the cell whose forward stencil (0,+1),(1,-1),(1,0),(1,+1)
                    // holds the other (hash.c:178-224)
                    const int gxj = cell_coord(q.x, P.cell_h), gyj = cell_coord(q.y, P.cell_h);
                    const int cxi = SPH_GXI, cyi = SPH_GYI;
                    const bool owner = (cxi == gxj && cyi == gyj) ? ((u & SPH_UID_MASK) < (uid[j] & SPH_UID_MASK))
                                                                  : (cxi != gxj ? cxi < gxj : cyi < gyj);
                    if (owner) { x += 0.000001f; y += 0.000001f; }
                }
                const float ratio = r * h_recip;
                if (ratio < 1.0f && r > 0.0f) {                        // fluid.c:588
                    const float w = 1.0f - ratio;
                    const float s = __fdiv_rn(fmaf(B, w, A) * w, r);
                    x = fmaf(-s, dx, x);
                    y = fmaf(-s, dy, y);
                }
                return;
            }
            // r and 1/r from one MUFU.RSQ (r2 > 1e-12: finite); the ratio < 1 gate as (1 - ratio)+,
            // branch-free: a listed neighbour has r2 <= h2, so the gate can only fail within rounding
            // of r == h, where the displacement vanishes anyway
            const float rs = rsqrt_approx(r2);
            const float w = fmaxf(fmaf(-r2 * rs, h_recip, 1.0f), 0.0f);
            const float s = fmaf(B, w, A) * w * rs;
            x = fmaf(-s, dx, x);
            y = fmaf(-s, dy, y);
        };
#endif
        // the exact walk over k_density's acceptance bits, two neighbours per trip, with the coincident-pair rules
        auto walk_all = [&]() {
#pragma unroll
        for (int d = 0; d < SPH_NROWS; d++) {
            // the lists were built by k_density on these same positions: walk its acceptance bits
            // (candidate order, so the order of summation is unchanged) ...
            // (A branch-free loop over ALL candidates like k_density's was measured here too: 121 us
            //  against 68 -- it loads position AND density of ~50 candidates instead of ~22 neighbours,
            //  and L1 wavefronts, not instruction issue, then bound the kernel.)
            const int b = R.b[d];
            sph_mask_t m = SPH_ROW_MASK(d);
            while (m) {
                // two neighbours per trip, all four loads (position + density of each) issued up front:
                // load latency was this kernel's top stall (profiles/r1_div2_full.csv, long_scoreboard);
                // four per trip (eight loads) was measured slower, 83 us against 68
                const int j0 = b + sph_mask_ffs(m) - 1;
                m &= m - 1;
                const bool two = m != 0;
                const int j1 = two ? b + sph_mask_ffs(m) - 1 : j0;
                m &= m - 1;
#if SPH_RELAX_PD4
                const float4 r0 = pd[j0], r1 = pd[j1];
                const float2 q0 = make_float2(r0.x, r0.y), d0 = make_float2(r0.z, r0.w);
                const float2 q1 = make_float2(r1.x, r1.y), d1 = make_float2(r1.z, r1.w);
#else
                const float2 q0 = pos[j0], d0 = dens[j0], q1 = pos[j1], d1 = dens[j1];
#endif
                pair(j0, q0, d0);
                if (two) pair(j1, q1, d1);
            }
            // ... and test the rare candidates beyond those a mask covers
            for (int j = b + SPH_MASK_BITS; j < R.e[d]; j++) {
                const float2 q = pos[j];
                if (dist2(q.x - p.x, q.y - p.y) > h2 || j == i) continue;
                pair(j, q, dens[j]);
            }
        }
        };
#if SPH_RELAX_RARE && !SPH_RELAX_BF
        {
            float r2min = 1.0f;                      // smallest squared distance met (one FMNMX per neighbour)
#if !SPH_PACKED_RELAX
            const float x_in = x, y_in = y;
#endif
            // one listed neighbour, no special cases; `on` = false for the filler slots of a row's last trip
            auto pair_fast = [&](float2 q, float2 dj, bool on) {
#if SPH_PACKED_RELAX
                const f32x2 dd = sub2(pk2(q.x, q.y), pp);
                const float2 sq = unpk2(mul2(dd, dd));
                const float r2 = __fadd_rn(sq.x, sq.y);
                r2min = fminf(r2min, r2);
                const float2 ab = unpk2(fma2(pk2(dj.x, dj.y), K12, AB0));
                const float rs = rsqrt_approx(r2);
                const float w = fmaxf(fmaf(-r2 * rs, h_recip, 1.0f), 0.0f);
                float s = fmaf(ab.y, w, ab.x) * w * rs;
                s = on ? s : 0.0f;
                xy = fma2(pk2(-s, -s), dd, xy);
#else
                const float dx = q.x - p.x, dy = q.y - p.y;
                const float r2 = dist2(dx, dy);
                r2min = fminf(r2min, r2);
                const float A = fmaf(dj.x, K1, Ai), B = fmaf(dj.y, K2, Bi);
                const float rs = rsqrt_approx(r2);
                const float w = fmaxf(fmaf(-r2 * rs, h_recip, 1.0f), 0.0f);
                float s = fmaf(B, w, A) * w * rs;
                s = on ? s : 0.0f;
                x = fmaf(-s, dx, x);
                y = fmaf(-s, dy, y);
#endif
            };
#pragma unroll
            for (int d = 0; d < SPH_NROWS; d++) {
                const int b = R.b[d];
                sph_mask_t m = SPH_ROW_MASK(d);
                while (m) {
                    int jj[SPH_RELAX_TRIP];
                    bool on[SPH_RELAX_TRIP];
                    on[0] = true;
                    jj[0] = b + sph_mask_ffs(m) - 1;
                    m &= m - 1;
#pragma unroll
                    for (int t = 1; t < SPH_RELAX_TRIP; t++) {
                        on[t] = m != 0;
                        jj[t] = on[t] ? b + sph_mask_ffs(m) - 1 : jj[0];      // (a filler repeats the trip's first neighbour)
                        m &= m - 1;
                    }
                    float2 qq[SPH_RELAX_TRIP], dj[SPH_RELAX_TRIP];
#pragma unroll
                    for (int t = 0; t < SPH_RELAX_TRIP; t++) {
#if SPH_RELAX_PD4
                        const float4 r4 = pd[jj[t]];
                        qq[t] = make_float2(r4.x, r4.y); dj[t] = make_float2(r4.z, r4.w);
#else
                        qq[t] = pos[jj[t]]; dj[t] = dens[jj[t]];
#endif
                    }
#pragma unroll
                    for (int t = 0; t < SPH_RELAX_TRIP; t++) pair_fast(qq[t], dj[t], on[t]);
                }
                // the rare candidates beyond those a mask covers
                for (int j = b + SPH_MASK_BITS; j < R.e[d]; j++) {
                    const float2 q = pos[j];
                    if (dist2(q.x - p.x, q.y - p.y) > h2 || j == i) continue;
                    pair(j, q, dens[j]);
                }
            }
            if (!(r2min > 1.0001e-12f)) {            // a (nearly) coincident pair: redo this particle with the reference's rules
#if SPH_PACKED_RELAX
                xy = pp;
#else
                x = x_in; y = y_in;
#endif
                walk_all();
            }
        }
#elif SPH_RELAX_BF
        // SPH_RELAX_BF=1 (round 2): the row's first SPH_MASK_BITS candidates in CANDIDATE order, branch-free: a
        // candidate whose acceptance bit is off loads nothing (predicated loads) and contributes an exact zero, so
        // values and order of the sum are the walk's, but the index of a neighbour is a loop counter instead of
        // the end of a dependent chain mask -> find-first-set -> address -> load (this kernel sat at 50 % issue
        // utilisation on exactly that chain, profiles/r2_call2_*).  The loop stops after the row's last set bit.
        // A (nearly) coincident pair -- rare -- raises a flag and the particle is redone by the exact walk.
        {
            bool rare = false;
#if SPH_PACKED_RELAX
            const f32x2 di2 = pk2(di.x, di.y);
#endif
#pragma unroll
            for (int d = 0; d < SPH_NROWS; d++) {
                const int b = R.b[d];
                sph_mask_t m = SPH_ROW_MASK(d);
#pragma unroll 4
                for (int j = b; m != 0; j++, m >>= 1) {
                    const bool on = (m & 1) != 0;
#if SPH_PACKED_RELAX
                    f32x2 qv = pp, dv = di2;
#if SPH_RELAX_PD4
                    if (on) { const float4 r4 = pd[j]; qv = pk2(r4.x, r4.y); dv = pk2(r4.z, r4.w); }
#else
                    if (on) { qv = ld2(pos + j); dv = ld2(dens + j); }
#endif
                    const f32x2 dd = sub2(qv, pp);
                    const float2 sq = unpk2(mul2(dd, dd));
                    const float r2 = __fadd_rn(sq.x, sq.y);
                    rare |= on & (r2 <= 1.0001e-12f);
                    const float2 ab = unpk2(fma2(dv, K12, AB0));
                    const float rs = rsqrt_approx(r2);
                    const float w = fmaxf(fmaf(-r2 * rs, h_recip, 1.0f), 0.0f);
                    float s = fmaf(ab.y, w, ab.x) * w * rs;
                    s = on ? s : 0.0f;
                    xy = fma2(pk2(-s, -s), dd, xy);
#else
                    float2 q = p, dj = di;
#if SPH_RELAX_PD4
                    if (on) { const float4 r4 = pd[j]; q = make_float2(r4.x, r4.y); dj = make_float2(r4.z, r4.w); }
#else
                    if (on) { q = pos[j]; dj = dens[j]; }
#endif
                    const float dx = q.x - p.x, dy = q.y - p.y;
                    const float r2 = dist2(dx, dy);
                    rare |= on & (r2 <= 1.0001e-12f);
                    const float A = fmaf(dj.x, K1, Ai), B = fmaf(dj.y, K2, Bi);
                    const float rs = rsqrt_approx(r2);
                    const float w = fmaxf(fmaf(-r2 * rs, h_recip, 1.0f), 0.0f);
                    float s = fmaf(B, w, A) * w * rs;
                    s = on ? s : 0.0f;
                    x = fmaf(-s, dx, x);
                    y = fmaf(-s, dy, y);
#endif
                }
                // the rare candidates beyond those a mask covers
                for (int j = b + SPH_MASK_BITS; j < R.e[d]; j++) {
                    const float2 q = pos[j];
                    if (dist2(q.x - p.x, q.y - p.y) > h2 || j == i) continue;
                    pair(j, q, dens[j]);
                }
            }
            if (rare) {
#if SPH_PACKED_RELAX
                xy = pp;
#else
                x = p.x; y = p.y;
#endif
                walk_all();
            }
        }
#else
        walk_all();
#endif
#if SPH_PACKED_RELAX
        { const float2 a = unpk2(xy); x = a.x; y = a.y; }
#endif
        float2 np = boundary(make_float2(x, y), P);                     // fluid.c:649
        const float2 pv = SPH_PREV_I;
        const float2 v = make_float2(clamp5(__fdiv_rn(np.x - pv.x, dt)), clamp5(__fdiv_rn(np.y - pv.y, dt)));
        pos_out[i] = np;
        vel_out[i] = v;
        if (P.one_x) {
#if SPH_DEFER_RELAX
            if (slot_i >= 0) t_slot[slot_i] = slot_v;
            slot_i = bin_position_deferred(i, np, ghost ? SPH_KEY_EMIG : 0, P, cnt, t_key, counters, tile_total, slot_v, !ghost) ? i : -1;
#else
            bin_position(i, np, ghost ? SPH_KEY_EMIG : 0, P, cnt, t_key, t_slot, counters, tile_total, !ghost);
#endif
            continue;
        }
        if (P.nranks > 1) {
            if (P.has_left && np.x - P.edge_start <= P.halo_w) {
                int k = atomicAdd(&msg_hdr(send_l)[1], 1);
                if (k < P.msg_cap) { msg_a(send_l)[k] = np; msg_b(send_l, P.msg_cap)[k] = v; msg_u(send_l, P.msg_cap)[k] = u; }
                else atomicAdd(&counters[CN_MSG_OVER], 1);
            }
            if (P.has_right && P.edge_end - np.x <= P.halo_w) {
                int k = atomicAdd(&msg_hdr(send_r)[1], 1);
                if (k < P.msg_cap) { msg_a(send_r)[k] = np; msg_b(send_r, P.msg_cap)[k] = v; msg_u(send_r, P.msg_cap)[k] = u; }
                else atomicAdd(&counters[CN_MSG_OVER], 1);
            }
        }
#if SPH_DEFER_RELAX
        if (slot_i >= 0) t_slot[slot_i] = slot_v;                       // the previous particle's: its atomic is back by now
        // (a local outside the window can only be an emigrant that is still waiting for room in a message: kept)
        slot_i = bin_position_deferred(i, np, 0, P, cnt, t_key, counters, tile_total, slot_v, true) ? i : -1;
#else
        bin_position(i, np, 0, P, cnt, t_key, t_slot, counters, tile_total, true);
#endif
#if SPH_PREFETCH & 2
        if (i + gstride < n) {
            cp_async_wait<0>();
            const Rows Rn = candidate_rows(s_pos[buf][threadIdx.x], P, cell_start);
            prefetch_rows(Rn, pos);
            prefetch_rows(Rn, dens);
        }
#endif
    }
#if SPH_DEFER_RELAX
    if (slot_i >= 0) t_slot[slot_i] = slot_v;
#endif
#undef SPH_GXI
#undef SPH_GYI
#undef SPH_ROW_MASK
#undef SPH_PREV_I
    pdl_done();
}

#if SPH_ADVECT_PV4
// (x, y, vx, vy) records of the resident entries rebuilt from the two arrays (sph_state_restore)
__global__ void __launch_bounds__(SPH_THREADS)
k_interleave(const int *__restrict__ counters, const float2 *__restrict__ pos, const float2 *__restrict__ vel, float4 *__restrict__ pv)
{
    const int n = counters[CN_NTOT];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 p = pos[i], v = vel[i];
        pv[i] = make_float4(p.x, p.y, v.x, v.y);
    }
}
#endif

// -------------------------------------------------------------------------------------------
// upload helper: bin freshly uploaded particles (sph_upload)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_bin_upload(const DevParams *__restrict__ Pp, int *__restrict__ counters, const float2 *__restrict__ pos,
             int *__restrict__ cnt, int *__restrict__ t_key, int *__restrict__ t_slot, int *__restrict__ tile_total)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        bin_position(i, pos[i], 0, P, cnt, t_key, t_slot, counters, tile_total);
}

// -------------------------------------------------------------------------------------------
// restart helper (sph_refresh_ghosts): a slab that was uploaded has no ghosts yet.  Put the resident LOCALS
// back into the sort's source arrays, bin them, and pack the ghost message a k_relax would have packed
// (position + velocity, which = 1); the following sort exchanges and unpacks it like any other.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_requeue(const DevParams *__restrict__ Pp, int *__restrict__ counters,
          const float2 *__restrict__ pos, const float2 *__restrict__ vel, const uint32_t *__restrict__ uid,
          float2 *__restrict__ pos_out, float2 *__restrict__ vel_out, uint32_t *__restrict__ uid_out,
          int *__restrict__ cnt, int *__restrict__ t_key, int *__restrict__ t_slot,
          unsigned char *send_l, unsigned char *send_r, int *__restrict__ tile_total)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t u = uid[i];
        uid_out[i] = u;
        if (u & SPH_HALO_BIT) { t_key[i] = SPH_KEY_DROP; continue; }
        const float2 p = pos[i], v = vel[i];
        pos_out[i] = p;
        vel_out[i] = v;
        if (P.nranks > 1) {
            if (P.has_left && p.x - P.edge_start <= P.halo_w) {
                int k = atomicAdd(&msg_hdr(send_l)[1], 1);
                if (k < P.msg_cap) { msg_a(send_l)[k] = p; msg_b(send_l, P.msg_cap)[k] = v; msg_u(send_l, P.msg_cap)[k] = u; }
                else atomicAdd(&counters[CN_MSG_OVER], 1);
            }
            if (P.has_right && P.edge_end - p.x <= P.halo_w) {
                int k = atomicAdd(&msg_hdr(send_r)[1], 1);
                if (k < P.msg_cap) { msg_a(send_r)[k] = p; msg_b(send_r, P.msg_cap)[k] = v; msg_u(send_r, P.msg_cap)[k] = u; }
                else atomicAdd(&counters[CN_MSG_OVER], 1);
            }
        }
        bin_position(i, p, 0, P, cnt, t_key, t_slot, counters, tile_total);
    }
}

// -------------------------------------------------------------------------------------------
// device-side constructFluidVolume + initParticles (geometry.c:29-59, fluid.c:747-768): the lattice of
// one slab's columns written straight into the sort's source arrays, so large problems never need a
// host AoS.  x = min_x + (start_col + col) * spacing, y = min_y + row * spacing, unfused as on the host.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_init_lattice(float min_x, float min_y, float spacing, int start_col, int ncols, int rows, int total_cols,
               float2 *__restrict__ pos, float2 *__restrict__ vel, uint32_t *__restrict__ uid)
{
    pdl_enter();
    const long long n = (long long)ncols * rows;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / ncols), col = (int)(i % ncols);
        pos[i] = make_float2(__fadd_rn(min_x, __fmul_rn((float)(start_col + col), spacing)),
                             __fadd_rn(min_y, __fmul_rn((float)row, spacing)));
        vel[i] = make_float2(0.0f, 0.0f);
        uid[i] = (uint32_t)(row * total_cols + start_col + col);
    }
}

// {local particles, work estimate, own time since the last call [us], waits over the same span [us]} for the edge
// policies (sph_copy_work); the two time accumulators start over
__global__ void k_pack_work(const int *__restrict__ counters, long long *__restrict__ xt, int *__restrict__ out)
{
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = counters[CN_NLOCAL];
        out[1] = counters[CN_COST];
        const long long busy_us = xt[5] / 1000, wait_us = xt[6] / 1000;
        out[2] = busy_us < 2000000000ll ? (int)busy_us : 2000000000;
        out[3] = wait_us < 2000000000ll ? (int)wait_us : 2000000000;
        xt[5] = 0; xt[6] = 0;
    }
}

// -------------------------------------------------------------------------------------------
// render feed (fluid.c:358-361): (2x/max_x - 1) * SHRT_MAX, truncated to int16
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPH_THREADS)
k_pack_coords(const DevParams *__restrict__ Pp, int *__restrict__ counters, const float2 *__restrict__ pos,
              const uint32_t *__restrict__ uid, short2 *__restrict__ out, int cap)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    const int lane = threadIdx.x & 31;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool mine = i < n && !(uid[i] & SPH_HALO_BIT);
        int k = i;
        if (P.nranks > 1) {
            // a slab's locals are compacted: ONE atomic per warp on the cursor (a million atomics on a single address
            // made the multi-GPU frame 25 % longer than its four steps)
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            int first = 0;
            if (lane == 0 && m) first = atomicAdd(&counters[CN_COORDS], __popc(m));
            first = __shfl_sync(0xffffffffu, first, 0);
            k = first + __popc(m & ((1u << lane) - 1u));
        }
        if (!mine || k >= cap) continue;
        const float2 p = pos[i];
        float fx = __fmul_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, p.x), P.tank_w), 1.0f), 32767.0f);
        float fy = __fmul_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, p.y), P.tank_h), 1.0f), 32767.0f);
        out[k] = make_short2((short)fx, (short)fy);
    }
}

// -------------------------------------------------------------------------------------------
// parity / inspection kernels (not on the hot path)
// -------------------------------------------------------------------------------------------
// Populations of the REFERENCE's buckets (cells of side h) from the sort grid's cell_start: the largest one
// and how many exceed the reference's capacity of 100 (hash.c:160-165 drops the excess silently).
__global__ void k_bucket_stats(const DevParams *__restrict__ Pp, const int *__restrict__ cell_start, int *__restrict__ out)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int cols = P.wx / SPH_CELL_DIV, rows = P.sort_rows / SPH_CELL_DIV;
    int mx = 0, over = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols * rows; c += gridDim.x * blockDim.x) {
        const int C = c % cols, R = c / cols;
        int n = 0;
        for (int r = 0; r < SPH_CELL_DIV; r++) {
            const int row = R * SPH_CELL_DIV + r;
            n += cell_start[row * P.wx + (C + 1) * SPH_CELL_DIV] - cell_start[row * P.wx + C * SPH_CELL_DIV];
        }
        mx = max(mx, n);
        over += n > SPH_REF_MAX_BUCKET;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    over = __reduce_add_sync(0xffffffffu, over);
    if ((threadIdx.x & 31) == 0) { if (mx) atomicMax(&out[0], mx); if (over) atomicAdd(&out[1], over); }
}

// Particles whose REFERENCE forward list (hash.c:178-224: owner = earlier bucket slot in the same cell, else the cell
// whose forward stencil holds the other; ghosts are appended to the local's list) would exceed its 400 entries, in the
// current sorted state: the exact figure behind sph_status.neighbor_overflow (the density kernel only keeps a cheap
// conservative count).
__global__ void k_forward_overflow(const DevParams *__restrict__ Pp, const int *__restrict__ counters,
                                   const float2 *__restrict__ pos, const uint32_t *__restrict__ uid,
                                   const int *__restrict__ cell_start, int *__restrict__ out)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    const float h2 = __fmul_rn(P.h, P.h);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (uid[i] & SPH_HALO_BIT) continue;
        const float2 p = pos[i];
        const uint32_t ui = uid[i] & SPH_UID_MASK;
        const int gxi = cell_coord(p.x, P.cell_h), gyi = cell_coord(p.y, P.cell_h);
        const Rows R = candidate_rows(p, P, cell_start);
        int total = 0;
        for (int d = 0; d < SPH_NROWS; d++) total += R.e[d] - R.b[d];
        if (total <= SPH_REF_MAX_NEIGHBORS) continue;                  // cannot have more forward neighbours than candidates
        int fwd = 0;
        for (int d = 0; d < SPH_NROWS; d++)
            for (int j = R.b[d]; j < R.e[d]; j++) {
                if (j == i) continue;
                const float2 q = pos[j];
                if (dist2(p.x - q.x, p.y - q.y) > h2) continue;
                const uint32_t uj = uid[j] & SPH_UID_MASK;
                const int gxj = cell_coord(q.x, P.cell_h), gyj = cell_coord(q.y, P.cell_h);
                const bool owner = (gxi == gxj && gyi == gyj) ? (ui < uj) : (gxi != gxj ? gxi < gxj : gyi < gyj);
                if ((uid[j] & SPH_HALO_BIT) || owner) fwd++;
            }
        if (fwd > SPH_REF_MAX_NEIGHBORS) atomicAdd(out, 1);
    }
}

__global__ void k_export_cells(const DevParams *__restrict__ Pp, const int *__restrict__ counters,
                               const float2 *__restrict__ pos, uint32_t *__restrict__ cell)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned gx = (unsigned)cell_coord(pos[i].x, P.cell_h), gy = (unsigned)cell_coord(pos[i].y, P.cell_h);
        cell[i] = gy * (unsigned)P.size_x + gx;             // global numbering of hash_val (hash.c:44)
    }
}

__global__ void k_export_pairs(const DevParams *__restrict__ Pp, const int *__restrict__ counters,
                               const float2 *__restrict__ pos, const uint32_t *__restrict__ uid,
                               const int *__restrict__ cell_start, unsigned long long *__restrict__ pairs,
                               unsigned long long cap, unsigned long long *__restrict__ n_pairs,
                               int *__restrict__ fwd_count)
{
    pdl_enter();
    const DevParams P = *Pp;
    const int n = counters[CN_NTOT];
    const float h2 = __fmul_rn(P.h, P.h);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 p = pos[i];
        const uint32_t ui = uid[i] & SPH_UID_MASK;
        const int gxi = cell_coord(p.x, P.cell_h), gyi = cell_coord(p.y, P.cell_h);
        int fwd = 0;
        const Rows R = candidate_rows(p, P, cell_start);
        for (int d = 0; d < SPH_NROWS; d++)
            for (int j = R.b[d]; j < R.e[d]; j++) {
                if (j == i) continue;
                const float2 q = pos[j];
                if (dist2(p.x - q.x, p.y - q.y) > h2) continue;
                const uint32_t uj = uid[j] & SPH_UID_MASK;
                if (uj > ui && pairs) {
                    unsigned long long k = atomicAdd(n_pairs, 1ull);
                    if (k < cap) pairs[k] = ((unsigned long long)ui << 32) | uj;
                }
                const int gxj = cell_coord(q.x, P.cell_h), gyj = cell_coord(q.y, P.cell_h);
                const bool owner = (gxi == gxj && gyi == gyj) ? (ui < uj) : (gxi != gxj ? gxi < gxj : gyi < gyj);
                if ((uid[j] & SPH_HALO_BIT) || owner) fwd++;
            }
        if (fwd_count) fwd_count[i] = fwd;
    }
}
