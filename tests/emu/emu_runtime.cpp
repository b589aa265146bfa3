// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  See fake/cuda_runtime.h.
// The fiber scheduler and the fake runtime API behind the kernel-source emulator.
#include <cuda_runtime.h>

#include <stdio.h>

#include <algorithm>

namespace emu {

thread_local Fiber *cur = nullptr;
thread_local Block blk;
thread_local uint3 block_idx, block_dim, grid_dim;

static const size_t kStack = 256 * 1024;
static thread_local std::vector<Fiber> fibers;
static thread_local std::vector<char *> stacks;
static thread_local ucontext_t main_ctx;
static thread_local const std::function<void()> *body_now = nullptr;
static thread_local int last_error = cudaSuccess;

// Fiber switch.  x86-64: callee-saved registers + stack pointer, no system call (swapcontext saves and restores
// the signal mask with two of them per switch, which was a quarter of the suite's run time); elsewhere ucontext.
#if defined(__x86_64__)
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch, .-emu_switch\n");
static thread_local void *main_sp = nullptr;
static thread_local std::vector<void *> fiber_sp;
static void to_main() { emu_switch(&fiber_sp[cur->tid.x], main_sp); }
static void to_fiber(int t) { emu_switch(&main_sp, fiber_sp[t]); }
static void trampoline();
static void prepare(int t)
{
    // a frame that emu_switch "returns" into: six zeroed registers, then the entry point; the stack pointer is
    // 8 modulo 16 when trampoline starts, as after a call
    void **top = (void **)(((uintptr_t)stacks[t] + kStack) & ~(uintptr_t)15);
    top[-2] = (void *)trampoline;
    for (int k = 3; k <= 8; k++) top[-k] = nullptr;
    fiber_sp[t] = (void *)(top - 8);
}
#else
static void to_main() { swapcontext(&cur->ctx, &main_ctx); }
static void to_fiber(int t) { swapcontext(&main_ctx, &fibers[t].ctx); }
static void trampoline();
static void prepare(int t)
{
    Fiber &f = fibers[t];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks[t];
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, trampoline, 0);
}
#endif

void yield() { to_main(); }

static void trampoline()
{
    (*body_now)();
    cur->done = true;
    to_main();
    abort();        // a finished fiber is never resumed
}

static void run_grid(int grid, int threads, const std::function<void()> &body)
{
    if (grid <= 0 || threads <= 0 || threads > 1024 || (threads & 31)) { last_error = cudaErrorInvalidValue; return; }
    if ((int)fibers.size() < threads) {
        fibers.resize(threads);
        while ((int)stacks.size() < threads) stacks.push_back((char *)malloc(kStack));
#if defined(__x86_64__)
        fiber_sp.resize(threads);
#endif
    }
    body_now = &body;
    grid_dim = uint3{(unsigned)grid, 1, 1};
    block_dim = uint3{(unsigned)threads, 1, 1};
    // SPH_EMU_ORDER=reverse | shuffle: blocks (and the threads inside them) run in another order, so atomics hand out
    // their slots differently -- what a result must not depend on (tests/test_emu_parity.py)
    const char *order = getenv("SPH_EMU_ORDER");
    const int mode = !order ? 0 : (order[0] == 'r' ? 1 : 2);
    static thread_local unsigned long long lcg = 88172645463325252ull;
    std::vector<int> border(grid), torder(threads);
    for (int b = 0; b < grid; b++) border[b] = mode == 1 ? grid - 1 - b : b;
    for (int t = 0; t < threads; t++) torder[t] = mode == 1 ? threads - 1 - t : t;
    auto shuffle = [&](std::vector<int> &v) {
        for (int k = (int)v.size() - 1; k > 0; k--) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(v[k], v[(int)((lcg >> 33) % (unsigned)(k + 1))]);
        }
    };
    if (mode == 2) shuffle(border);
    for (int bi = 0; bi < grid; bi++) {
        const int b = border[bi];
        if (mode == 2) shuffle(torder);
        block_idx = uint3{(unsigned)b, 0, 0};
        memset(&blk, 0, sizeof blk);
        blk.nthreads = threads;
        for (int t = 0; t < threads; t++) {
            fibers[t].tid = uint3{(unsigned)t, 0, 0};
            fibers[t].done = false;
            prepare(t);
        }
        int alive = threads;
        long long rounds = 0;
        while (alive > 0) {
            alive = 0;
            for (int ti = 0; ti < threads; ti++) {
                const int t = torder[ti];
                if (fibers[t].done) continue;
                cur = &fibers[t];
                to_fiber(t);
                alive += !fibers[t].done;
            }
            if (++rounds > 100000000ll) { fprintf(stderr, "emu: block %d never finished (barrier not reached by all threads?)\n", b); abort(); }
        }
    }
    cur = nullptr;
    body_now = nullptr;
}

}  // namespace emu

struct emu_graph { std::vector<std::tuple<int, int, std::function<void()>>> nodes; };
struct emu_stream { emu_graph *capturing; };

// emulated devices are just host threads; SPH_EMU_DEVICES=0 plays a box without a GPU
static int device_count()
{
    const char *e = getenv("SPH_EMU_DEVICES");
    return e ? atoi(e) : 8;
}

void emu::launch_on_stream(cudaStream_t s, int grid, int threads, std::function<void()> body)
{
    if (s && s->capturing) { s->capturing->nodes.emplace_back(grid, threads, std::move(body)); return; }
    run_grid(grid, threads, body);
}

const char *cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid value";
    case cudaErrorNoDevice: return "no device";
    case cudaErrorNotSupported: return "not supported by the emulator";
    case cudaErrorStreamCaptureUnsupported: return "operation not permitted while capturing";
    default: return "error";
    }
}
cudaError_t cudaGetLastError() { int e = emu::last_error; emu::last_error = cudaSuccess; return e; }
cudaError_t cudaGetDeviceCount(int *n) { *n = device_count(); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d >= 0 && d < device_count() ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    p->multiProcessorCount = 1;                     // one block at a time; grids of at most 8 blocks keep the fiber count down
    p->clockRate = 1000000;                         // kHz; clock64() counts nanoseconds here
    strcpy(p->name, "kernel-source emulator (host)");
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream{nullptr}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { return s && s->capturing ? cudaErrorStreamCaptureUnsupported : cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t s, int)
{
    if (!s || s->capturing) return cudaErrorInvalidValue;
    s->capturing = new emu_graph();
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g)
{
    if (!s || !s->capturing) return cudaErrorInvalidValue;
    *g = s->capturing;
    s->capturing = nullptr;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long) { *e = new emu_graph(*g); return cudaSuccess; }
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s)
{
    if (s && s->capturing) return cudaErrorStreamCaptureUnsupported;
    for (auto &n : e->nodes) emu::run_grid(std::get<0>(n), std::get<1>(n), std::get<2>(n));
    return cudaSuccess;
}

struct emu_event { int unused; };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { return !e || (s && s->capturing) ? cudaErrorInvalidValue : cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { return e ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t e, unsigned) { return e ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }

// device memory: filled with a poison pattern, because cudaMalloc does not zero either
cudaError_t emu_malloc(void **p, size_t bytes)
{
    *p = malloc(bytes ? bytes : 1);
    if (!*p) return cudaErrorInvalidValue;
    memset(*p, 0xCD, bytes);
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
static bool capturing(cudaStream_t s) { return s && s->capturing; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t s)
{
    if (capturing(s)) return cudaErrorStreamCaptureUnsupported;
    memset(p, v, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t bytes, cudaMemcpyKind) { memmove(d, s, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t bytes, cudaMemcpyKind, cudaStream_t st)
{
    // captured: a memcpy node, which reads its source when the graph RUNS (the library's queued-parameter step)
    if (capturing(st)) { st->capturing->nodes.emplace_back(1, 32, [=]() { if (threadIdx.x == 0) memmove(d, s, bytes); }); return cudaSuccess; }
    memmove(d, s, bytes);
    return cudaSuccess;
}
// "peer memory": every emulated device lives in this process, so a handle is just the pointer
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
