// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile sph_b200/csrc/*.cu(h) UNCHANGED for the host, so that
// the logic of the kernels and of the C-ABI layer (stage order, buffer rotation, graph capture, counters,
// message packing) can be exercised by `pytest -m "not gpu"` in a container without a GPU.  Only
// tests/emu/build_emu.py puts this directory on an include path, and only tests/test_emu_*.py load the
// resulting tests/emu/_build/libsph_emu.so.  libsph_b200.so (nvcc, sm_100a) never sees any of this and has
// no CPU path; the GPU parity tests (-m gpu) remain the parity tests proper.
//
// Execution model: the host thread that calls the library is the device.  A kernel launch runs its blocks one
// after another; the SPH_THREADS threads of a block are fibers (ucontext) run round-robin, switching only at __syncthreads and warp collectives.
// Data races between threads of a launch are therefore NOT detected (atomics are plain read-modify-writes),
// and floating point differs from the GPU where nvcc contracts a*b+c into FFMA and where MUFU approximations
// are used (the IEEE results here are closer to the oracle, not further).  The _rn intrinsics are exact.
#pragma once
#ifndef SPH_EMU
#error "tests/emu/fake/cuda_runtime.h is only for the kernel-source emulator build (-DSPH_EMU)"
#endif

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <functional>
#include <tuple>
#include <utility>
#include <vector>

// ---------------------------------------------------------------------------------------------
// language
// ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local   // a device's blocks run one after another on its OS thread: one instance IS the block's copy

struct float2 { float x, y; };
struct int4 { int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct short2 { short x, y; };
struct uint3 { unsigned x, y, z; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline short2 make_short2(short x, short y) { return short2{x, y}; }

namespace emu {
struct Warp { long long slot[32]; long long result[32]; int arrived; unsigned gen; };
struct Block {
    int nthreads, arrived, or_acc, or_result;
    unsigned gen;
    Warp warp[32];
};
struct Fiber { ucontext_t ctx; uint3 tid; bool done; };
// one emulated device per OS thread: two slabs of a peer-memory test run as two threads of one process
extern thread_local Fiber *cur;
extern thread_local Block blk;
extern thread_local uint3 block_idx, block_dim, grid_dim;
void yield();
void launch(int grid, int threads, std::function<void()> body);     // runs now, or records into a capturing stream
}

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::block_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)

// block barrier
static inline void emu_block_arrive_and_wait()
{
    emu::Block &b = emu::blk;
    if (++b.arrived == b.nthreads) { b.arrived = 0; b.or_result = b.or_acc; b.or_acc = 0; b.gen++; return; }
    const unsigned g = b.gen;
    while (b.gen == g) emu::yield();
}
static inline void __syncthreads() { emu_block_arrive_and_wait(); }
static inline int __syncthreads_or(int p)
{
    emu::blk.or_acc |= p != 0;
    emu_block_arrive_and_wait();
    // every thread reads the result before any thread can complete the NEXT barrier (that needs all of them)
    return emu::blk.or_result;
}

// warp collectives (full masks only: every kernel here runs whole warps through them)
template <class F> static inline long long emu_warp_collective(long long v, F combine)
{
    const int lane = threadIdx.x & 31;
    emu::Warp &w = emu::blk.warp[threadIdx.x >> 5];
    w.slot[lane] = v;
    if (++w.arrived == 32) { combine(w.slot, w.result); w.arrived = 0; w.gen++; }
    else { const unsigned g = w.gen; while (w.gen == g) emu::yield(); }
    return w.result[lane];
}
static inline int __reduce_add_sync(unsigned, int v)
{
    return (int)emu_warp_collective(v, [](long long *s, long long *r) { long long t = 0; for (int i = 0; i < 32; i++) t += s[i]; for (int i = 0; i < 32; i++) r[i] = (int)t; });
}
static inline int __reduce_max_sync(unsigned, int v)
{
    return (int)emu_warp_collective(v, [](long long *s, long long *r) { long long t = s[0]; for (int i = 1; i < 32; i++) t = s[i] > t ? s[i] : t; for (int i = 0; i < 32; i++) r[i] = t; });
}
static inline int __shfl_up_sync(unsigned, int v, int o)
{
    return (int)emu_warp_collective(v, [o](long long *s, long long *r) { for (int i = 0; i < 32; i++) r[i] = i >= o ? s[i - o] : s[i]; });
}
static inline int __shfl_sync(unsigned, int v, int src)
{
    return (int)emu_warp_collective(v, [src](long long *s, long long *r) { for (int i = 0; i < 32; i++) r[i] = s[src & 31]; });
}
static inline int __shfl_xor_sync(unsigned, int v, int o)
{
    return (int)emu_warp_collective(v, [o](long long *s, long long *r) { for (int i = 0; i < 32; i++) r[i] = s[i ^ o]; });
}
static inline unsigned __ballot_sync(unsigned, int p)
{
    return (unsigned)emu_warp_collective(p != 0, [](long long *s, long long *r) { unsigned m = 0; for (int i = 0; i < 32; i++) m |= (unsigned)(s[i] != 0) << i; for (int i = 0; i < 32; i++) r[i] = m; });
}

// atomics: within a launch one OS thread runs all fibers and switches only at barriers, so plain read-modify-writes
// suffice; memory shared with ANOTHER emulated device (peer-memory exchange blocks) is ordered by the fences
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu::yield(); }
static inline long long clock64() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (long long)t.tv_sec * 1000000000ll + t.tv_nsec; }

// loads
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }

// arithmetic: the build uses -ffp-contract=off, so plain expressions are IEEE single operations
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// runtime API (the subset sph_capi.cu uses)
// ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNoDevice = 100, cudaErrorNotSupported = 801,
       cudaErrorStreamCaptureUnsupported = 900 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
struct emu_stream;
struct emu_graph;
typedef emu_stream *cudaStream_t;
typedef emu_graph *cudaGraph_t;
typedef emu_graph *cudaGraphExec_t;
struct cudaDeviceProp { int multiProcessorCount; int clockRate; char name[64]; };
struct cudaIpcMemHandle_t { char reserved[64]; };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, int mode);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s);
// events: every stream of the emulator is synchronous, so an event has nothing to remember
struct emu_event;
typedef emu_event *cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t emu_malloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return emu_malloc((void **)p, bytes); }
cudaError_t cudaFreeHost(void *p);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return emu_malloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t s);
cudaError_t cudaMemcpy(void *d, const void *s, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t bytes, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
// an emulated device runs ONE block at a time: that is its co-resident capacity (k_unpack's grid is bounded by it,
// because its blocks wait for the neighbour device inside the kernel)
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }

// kernel launch: SPH_LAUNCH(kernel, grid, stream)(args...)  (sph_capi.cu; k<<<grid, SPH_THREADS, 0, stream>>> under nvcc)
namespace emu {
template <class... A> struct Launcher {
    void (*k)(A...);
    int grid, threads;
    cudaStream_t stream;
    template <class... B> void operator()(B &&...b) const
    {
        std::tuple<A...> args(static_cast<A>(b)...);
        void (*kk)(A...) = k;
        launch_on(stream, grid, threads, [kk, args]() { std::apply(kk, args); });
    }
    static void launch_on(cudaStream_t s, int grid, int threads, std::function<void()> body);
};
void launch_on_stream(cudaStream_t s, int grid, int threads, std::function<void()> body);
template <class... A> void Launcher<A...>::launch_on(cudaStream_t s, int grid, int threads, std::function<void()> body)
{
    launch_on_stream(s, grid, threads, std::move(body));
}
template <class... A> Launcher<A...> make_launcher(void (*k)(A...), int grid, int threads, cudaStream_t s)
{
    return Launcher<A...>{k, grid, threads, s};
}
}
