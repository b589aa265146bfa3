/*
 * C host layer of sph_b200 (include/sph_host.h): start-up geometry, parameter model, slab
 * load balancer.  Plain C99, no CUDA; linked into libsph_b200.so.
 * Citations: AdamSimpson/SPH `src/`.
 */
#include "sph_host.h"

#include <math.h>
#include <string.h>

float sph_host_spacing(float water_w, float water_h, int n_request)
{
    /* fluid.c:141-144: float area, float quotient, pow() in double, back to float */
    float area = water_w * water_h;
    float per_particle = area / (float)n_request;
    return (float)pow((double)per_particle, 0.5);
}

int sph_host_preset(sph_tunable *t, char which)
{
    /* controls.c:344-401: g, k, k_near, k_spring, sigma, beta, rest_density */
    static const struct { char key; float v[7]; } table[] = {
        { 'x', { 6.0f, 0.2f, 6.0f, 10.0f, 5.0f, 0.5f, 30.0f } },
        { 'y', { 6.0f, 0.1f, 3.0f, -30.0f, 100.0f, 10.0f, 30.0f } },
        { 'a', { 0.0f, 0.2f, 6.0f, 10.0f, 20.0f, 2.0f, 55.0f } },
        { 'b', { 6.0f, 0.0f, 0.0f, 115.0f, 20.0f, 2.0f, 0.0f } },
    };
    for (unsigned i = 0; i < sizeof table / sizeof table[0]; i++) {
        if (table[i].key != which) continue;
        const float *v = table[i].v;
        t->g = v[0]; t->k = v[1]; t->k_near = v[2]; t->k_spring = v[3];
        t->sigma = v[4]; t->beta = v[5]; t->rest_density = v[6];
        return 0;
    }
    return -1;
}

void sph_host_default_params(sph_tunable *t, float h, float tank_w, float tank_h)
{
    memset(t, 0, sizeof *t);
    sph_host_preset(t, 'x');                            /* fluid.c:90-97 are the 'x' values */
    t->smoothing_radius = h;                            /* fluid.c:159 */
    t->time_step = (1.0f / 30.0f) / 4.0f;               /* fluid.c:91,105-106 */
    t->node_start_x = 0.0f;
    t->node_end_x = tank_w;
    t->mover_center_x = 0.5f * tank_w;
    t->mover_center_y = 0.35f * tank_h;
    t->mover_width = 2.0f * tank_w / 15.0f;             /* 2.0 in the 15-wide default tank (fluid.c:98-99) */
    t->mover_height = t->mover_width;
    t->mover_type = SPH_SPHERE_MOVER;
    t->kill_sim = 0;
    t->active = 1;
}

int sph_host_partition(float tank_w, float water_min_x, float water_max_x, float water_min_y,
                       float water_max_y, float spacing, int nranks,
                       int *start_col, int *ncols, float *start_x, float *end_x)
{
    /* geometry.c:111-127: columns incl. the zeroth, split evenly, remainder to the left ranks */
    const int columns = (int)floor((water_max_x - water_min_x) / spacing) + 1;
    const int base = columns / nranks;
    const int extra = columns - base * nranks;
    int before = 0;
    for (int r = 0; r < nranks; r++) {
        const int mine = base + (r < extra ? 1 : 0);
        start_col[r] = before;
        ncols[r] = mine;
        /* geometry.c:141-147 */
        float s = water_min_x + ((before - 1) * spacing);
        float e = s + (mine * spacing);
        start_x[r] = (r == 0) ? 0.0f : s;
        end_x[r] = (r == nranks - 1) ? tank_w : e;
        before += mine;
    }
    const int rows = (int)floor((water_max_y - water_min_y) / spacing);      /* geometry.c:152-156 */
    return before * rows;
}

int sph_host_lattice(float water_min_x, float water_min_y, float water_max_y, float spacing,
                     int start_col, int ncols, int total_cols, sph_particle *out, uint32_t *uid)
{
    const int rows = (int)floor((water_max_y - water_min_y) / spacing);      /* geometry.c:35 */
    int n = 0;
    for (int row = 0; row < rows; row++) {                                   /* geometry.c:46-59: y outer, x inner */
        const float y = water_min_y + row * spacing;
        for (int col = 0; col < ncols; col++, n++) {
            sph_particle *p = &out[n];
            memset(p, 0, sizeof *p);                                          /* fluid.c:762-767: v = a = 0 */
            p->x = water_min_x + (start_col + col) * spacing;
            p->y = y;
            p->id = n;
            if (uid) uid[n] = (uint32_t)(row * total_cols + start_col + col);
        }
    }
    return n;
}

static float slab_len(const sph_tunable *t) { return t->node_end_x - t->node_start_x; }

/* move the edge shared by slabs `left` and `left+1` */
static void shift_edge(sph_tunable *m, int left, float delta)
{
    m[left + 1].node_start_x += delta;
    m[left].node_end_x = m[left + 1].node_start_x;
}

void sph_host_balance(sph_tunable *m, int nactive, const int *counts, int total)
{
    sph_host_balance_ex(m, nactive, counts, total, 15.0f);
}

void sph_host_balance_ex(sph_tunable *m, int nactive, const int *counts, int total, float band_divisor)
{
    /* renderer.c:433-440 */
    const int even = total / nactive;
    const int band = (int)(even / band_divisor);
    const float h = m[0].smoothing_radius;
    const float dx = (float)(h * 0.125);
    /* renderer.c:444-458: right to left, each slab looks at its own left edge */
    for (int r = nactive - 1; r >= 1; r--) {
        const int diff = counts[r] - even;
        if (diff > band && slab_len(&m[r]) > 2 * h) shift_edge(m, r - 1, dx);
        else if (diff < -band && slab_len(&m[r - 1]) > 2 * h) shift_edge(m, r - 1, -dx);
    }
    /* renderer.c:461-476: the leftmost slab is tested once more through its right edge */
    if (nactive > 1) {
        const int diff = counts[0] - even;
        if (diff > band && slab_len(&m[0]) > 2 * h) {
            m[0].node_end_x -= dx;
            m[1].node_start_x = m[0].node_end_x;
        } else if (diff < -band && slab_len(&m[1]) > 2 * h) {
            m[0].node_end_x += dx;
            m[1].node_start_x = m[0].node_end_x;
        }
    }
}

void sph_host_balance_time(sph_tunable *m, int nactive, const int *busy, float gain, float max_shift_h, float min_width_h)
{
    /* Not the reference's policy (that is sph_host_balance above).  Every interior edge moves towards the slower of
     * its two slabs by `gain` times the shift that would equalise their measured times if a slab's time were
     * proportional to its width: delta = gain (T_right - T_left) / (T_left / w_left + T_right / w_right),
     * bounded by max_shift_h smoothing radii per call, never leaving a slab narrower than min_width_h radii, and left
     * alone inside a dead band of 0.5 % of the pair's time.  Left to right, on the edges as they were on entry. */
    const float h = m[0].smoothing_radius;
    float len[64], shift[64];
    if (nactive < 2 || nactive > 64) return;
    for (int r = 0; r < nactive; r++) len[r] = slab_len(&m[r]);
    for (int e = 0; e + 1 < nactive; e++) {          /* edge e: between slab e and slab e + 1 */
        shift[e] = 0.0f;
        const float tl = (float)busy[e], tr = (float)busy[e + 1];
        if (tl <= 0.0f || tr <= 0.0f || len[e] <= 0.0f || len[e + 1] <= 0.0f) continue;
        if (fabsf(tr - tl) <= 0.005f * (tl + tr)) continue;
        float d = gain * (tr - tl) / (tl / len[e] + tr / len[e + 1]);     /* > 0: the right slab is slower, it shrinks */
        const float cap = max_shift_h * h;
        if (d > cap) d = cap;
        if (d < -cap) d = -cap;
        shift[e] = d;
    }
    /* widths after this and the neighbouring edges' moves must stay above the minimum.  Withdrawing one edge's move
     * changes what its neighbours' moves leave of the slab between them, so the test is repeated until nothing is
     * withdrawn any more (a single left-to-right pass accepted edge e against a move of edge e + 1 that was then
     * withdrawn: slabs came out up to max_shift_h narrower than the minimum; found by a randomised property check).
     * No move at all is always admissible, so this ends after at most nactive passes. */
    for (int pass = 0; pass < nactive; pass++) {
        int withdrawn = 0;
        for (int e = 0; e + 1 < nactive; e++) {
            const float d = shift[e];
            if (d == 0.0f) continue;
            const float left_after = len[e] + d - (e > 0 ? shift[e - 1] : 0.0f);
            const float right_after = len[e + 1] - d + (e + 2 < nactive ? shift[e + 1] : 0.0f);
            /* ... and so must what the move leaves of the OLD extent of the slab it eats into.  In the step in which
             * the new edges land, the slab on the other side of that slab still holds what is being handed over: a slab
             * whose two edges move the same way keeps its width, but if its old and new extents overlap by less than
             * the ghost layer, the strip its neighbour needs as ghosts is owned, for that one step, by the neighbour's
             * neighbour (soak run 93072 of tests/fuzz/fuzz_slabs.py: a slab 2.75 h wide, both edges 2 h to the left, 23
             * particles beside the new edge differ by ulps from the one-slab run). */
            const float left_kept = len[e] + (d < 0.0f ? d : 0.0f);
            const float right_kept = len[e + 1] - (d > 0.0f ? d : 0.0f);
            if (left_after < min_width_h * h || right_after < min_width_h * h ||
                left_kept < min_width_h * h || right_kept < min_width_h * h) { shift[e] = 0.0f; withdrawn = 1; }
        }
        if (!withdrawn) break;
    }
    for (int e = 0; e + 1 < nactive; e++)
        if (shift[e] != 0.0f) shift_edge(m, e, shift[e]);
}

void sph_host_mover_autopilot(sph_tunable *t, float tank_w, float tank_h, float *gl_x, int *direction)
{
    sph_host_mover_autopilot_ex(t, tank_w, tank_h, gl_x, direction, 0.01f);
}

void sph_host_mover_autopilot_ex(sph_tunable *t, float tank_w, float tank_h, float *gl_x, int *direction, float dx_gl)
{
    /* renderer.c:513-531 */
    float x = *gl_x + dx_gl * (float)(*direction);
    if (x > 1.0f || x < -1.0f) *direction = -*direction;
    const float y = sinf(3.14f * 5.0f * x) / 10.0f - 0.6f;
    *gl_x = x;
    /* opengl_to_sim, renderer.c:396-404 */
    const float half_w = tank_w * 0.5f, half_h = tank_h * 0.5f;
    t->mover_center_x = x * half_w + half_w;
    t->mover_center_y = y * half_h + half_h;
}

int sph_host_remove_partition(sph_tunable *m, int nactive)
{
    /* controls.c:405-426 */
    if (nactive == 1) return nactive;
    const int gone = nactive - 1;
    m[gone - 1].node_end_x = m[gone].node_end_x;
    const float parked = (float)(m[gone].node_end_x + 1.0);
    m[gone].node_start_x = parked;
    m[gone].node_end_x = parked;
    m[gone].active = 0;
    return nactive - 1;
}

int sph_host_add_partition(sph_tunable *m, int nactive, int nranks)
{
    /* controls.c:429-455 */
    if (nactive == nranks) return nactive;
    const float len = slab_len(&m[nactive - 1]);
    const float h = m[0].smoothing_radius;
    if (len < 2.5 * h) return nactive;
    m[nactive].node_end_x = m[nactive - 1].node_end_x;
    const float mid = (float)(m[nactive - 1].node_start_x + len * 0.5);
    m[nactive - 1].node_end_x = mid;
    m[nactive].node_start_x = mid;
    m[nactive].active = 1;
    return nactive + 1;
}
