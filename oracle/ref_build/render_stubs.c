/*
 * render_stubs -- a headless platform for the reference's UNMODIFIED renderer.c / controls.c.  TEST INFRASTRUCTURE ONLY.
 *
 * renderer.c (start_renderer, the parameter scatter, the coordinate gather, check_partition_left) and controls.c
 * (presets, mover, add/remove_partition) are compiled where they lie against the fake GL headers of fake_gl/.
 * What they call in the GL modules (particles_gl.c, liquid_gl.c, ... glfw_utils.c) is provided here as no-ops,
 * except for three hooks:
 *
 *   init_ogl             1920 x 1080 "screen" (renderer.c:127-129 broadcasts it: a 16:9 tank)
 *   check_user_input     the "user": drags the mover along a fixed path through the reference's own
 *                        set_mover_gl_center (controls.c:227-237), one position per frame
 *                        and presses the keys listed in $SPH_RENDER_SCRIPT ("4:remove 9:add 6:b 12:x": at frame 4
 *                        the reference's own remove_partition, controls.c:405-426; add_partition, :429-455; the
 *                        fluid presets set_fluid_x/y/a/b, :344-401)
 *   render_liquid /      records the frame the renderer would have drawn
 *   render_particles
 *   swap_ogl             end of frame: appends it to $SPH_RENDER_OUT
 *   window_should_close  true after $SPH_RENDER_FRAMES frames (then renderer.c:241-250 scatters kill_sim)
 *
 * Record:  "SPHR", int32 K, float world_w, world_h;  per frame: int32 n, K x (float start, end) slab edges as
 * scattered for this frame, float mover_x, mover_y, n x (float gl_x, gl_y) in arrival order.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpi.h"
#include "particles_gl.h"
#include "liquid_gl.h"
#include "mover_gl.h"
#include "background_gl.h"
#include "font_gl.h"
#include "dividers_gl.h"
#include "exit_menu_gl.h"
#include "renderer.h"
#include "controls.h"

static struct {
    render_t *rs;
    FILE *out;
    int frames, wanted, header;
    const float *points;
    int n, stride;
} S;

void init_ogl(gl_t *state, render_t *render_state)
{
    state->screen_width = 1920; state->screen_height = 1080; state->window = NULL;
    S.rs = render_state;
    const char *f = getenv("SPH_RENDER_FRAMES"), *o = getenv("SPH_RENDER_OUT");
    S.wanted = f ? atoi(f) : 4;
    S.out = fopen(o ? o : "ref_world.bin", "wb");
    if (!S.out) { perror("SPH_RENDER_OUT"); exit(2); }
}

bool window_should_close(gl_t *state) { (void)state; return S.frames >= S.wanted; }

void check_user_input(gl_t *state)
{
    (void)state;
    set_mover_gl_center(S.rs, -0.5f + 0.06f * (float)S.frames, -0.4f);
    const char *script = getenv("SPH_RENDER_SCRIPT");
    for (const char *p = script; p && *p; ) {
        int frame = 0, used = 0;
        char key[16];
        if (sscanf(p, " %d:%15s%n", &frame, key, &used) < 2) break;
        p += used;
        if (frame != S.frames) continue;
        if (!strcmp(key, "remove")) remove_partition(S.rs);
        else if (!strcmp(key, "add")) add_partition(S.rs);
        else if (!strcmp(key, "x")) set_fluid_x(S.rs);
        else if (!strcmp(key, "y")) set_fluid_y(S.rs);
        else if (!strcmp(key, "a")) set_fluid_a(S.rs);
        else if (!strcmp(key, "b")) set_fluid_b(S.rs);
        else { fprintf(stderr, "render_stubs: unknown key '%s' in SPH_RENDER_SCRIPT\n", key); exit(2); }
    }
}

void render_liquid(float *points, float diameter_pixels, int num_points, liquid_t *state)
{ (void)diameter_pixels; (void)state; S.points = points; S.n = num_points; S.stride = 2; }
void render_particles(float *points, float diameter_pixels, int num_points, particles_t *state)
{ (void)diameter_pixels; (void)state; S.points = points; S.n = num_points; S.stride = 5; }

void swap_ogl(gl_t *state)
{
    (void)state;
    const int K = S.rs->num_compute_procs;
    if (!S.header) {
        fwrite("SPHR", 1, 4, S.out);
        fwrite(&K, 4, 1, S.out);
        fwrite(&S.rs->sim_width, 4, 1, S.out);
        fwrite(&S.rs->sim_height, 4, 1, S.out);
        S.header = 1;
    }
    fwrite(&S.n, 4, 1, S.out);
    for (int i = 0; i < K; i++) {
        fwrite(&S.rs->node_params[i].node_start_x, 4, 1, S.out);
        fwrite(&S.rs->node_params[i].node_end_x, 4, 1, S.out);
    }
    fwrite(&S.rs->node_params[0].mover_center_x, 4, 1, S.out);
    fwrite(&S.rs->node_params[0].mover_center_y, 4, 1, S.out);
    for (int j = 0; j < S.n; j++) fwrite(S.points + (size_t)j * S.stride, 4, 2, S.out);
    S.frames++;
}

void exit_ogl(gl_t *state) { (void)state; if (S.out) fclose(S.out); S.out = NULL; }

void init_particles(particles_t *state, int screen_width, int screen_height) { (void)state; (void)screen_width; (void)screen_height; }
void init_liquid(liquid_t *state, int screen_width, int screen_height) { (void)state; (void)screen_width; (void)screen_height; }
void init_mover(mover_t *state) { (void)state; }
void init_font(font_t *state, int screen_width, int screen_height) { (void)state; (void)screen_width; (void)screen_height; }
void init_background(background_t *state, int screen_width, int screen_height) { (void)state; (void)screen_width; (void)screen_height; }
void init_dividers(dividers_t *state, int screen_width, int screen_height) { (void)state; (void)screen_width; (void)screen_height; }
void init_exit_menu(exit_menu_t *state, gl_t *gl_state) { (void)state; (void)gl_state; }
void exit_exit_menu(exit_menu_t *state) { (void)state; }
void render_mover(float *center, float *gl_dims, float *color, mover_t *state) { (void)center; (void)gl_dims; (void)color; (void)state; }
void render_all_text(font_t *state, render_t *render_state, double fps) { (void)state; (void)render_state; (void)fps; }
void draw_background(background_t *state) { (void)state; }
void render_dividers(dividers_t *state, float *node_edges, float *colors_by_rank, int num_nodes)
{ (void)state; (void)node_edges; (void)colors_by_rank; (void)num_nodes; }
void render_exit_menu(exit_menu_t *state, float cursor_x, float cursor_y) { (void)state; (void)cursor_x; (void)cursor_y; }
