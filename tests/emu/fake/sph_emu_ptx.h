// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  See cuda_runtime.h in this directory.
// Host stand-ins for the PTX helpers of sph_b200/csrc/sph_device.cuh (included from there under SPH_EMU only):
// IEEE operations in place of the MUFU approximations, a struct of two floats in place of a packed register pair,
// GCC atomics in place of the system-scope acquire / release.
#pragma once
static inline float rcp_approx(float x) { return 1.0f / x; }
static inline void prefetch_l1(const void *) {}
static inline void tile_count_add(int *tile_total, int tile) { tile_total[tile] += 1; }   // (fibers of one OS thread: a plain add)
static inline long long global_timer_ns() { return clock64(); }
static inline float max_nan(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
struct f32x2 { float lo, hi; };
static inline f32x2 pk2(float lo, float hi) { return f32x2{lo, hi}; }
static inline float2 unpk2(f32x2 v) { return float2{v.lo, v.hi}; }
static inline f32x2 add2(f32x2 a, f32x2 b) { return f32x2{a.lo + b.lo, a.hi + b.hi}; }
static inline f32x2 sub2(f32x2 a, f32x2 b) { return f32x2{a.lo - b.lo, a.hi - b.hi}; }
static inline f32x2 mul2(f32x2 a, f32x2 b) { return f32x2{a.lo * b.lo, a.hi * b.hi}; }
static inline f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { return f32x2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
static inline f32x2 ld2(const float2 *p) { return f32x2{p->x, p->y}; }
static inline int ld_acquire_sys(const int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_release_sys(int *p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
static inline float sqrt_approx(float x) { return sqrtf(x); }
static inline void pdl_enter() {}   // launches of the emulator are already serialised
static inline void pdl_done() {}
// asynchronous copies land at once (a fiber only reads the slots it filled itself)
static inline void cp_async4(void *smem, const void *g) { memcpy(smem, g, 4); }
static inline void cp_async8(void *smem, const void *g) { memcpy(smem, g, 8); }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
