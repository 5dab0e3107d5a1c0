// TEST INFRASTRUCTURE (CPU suite only; never shipped, never loaded by the product package).
//
// A host-memory stand-in for the handful of CUDA runtime entry points the host operator (csrc/dfd_exec.cu) calls, so
// that the operator's HOST logic — chunk coalescing, bitmap / offset staging, list and view conversion, slicing of the
// destination-sorted chunk into Arrow batches, error fan-out, the pinned pool and its cache — can be exercised by
// `pytest -m "not gpu"` with the very object file nvcc built for the product.  "Device" memory is host memory, streams
// and events complete immediately (every copy is synchronous), and the partition kernels are replaced by the CPU
// oracle in harness_dfd.cu.  Nothing here is a CPU fallback of the product: it is linked only into the test library
// tests/test_exec_cpu_harness.py builds under a temporary directory.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

// A native backtrace on SIGSEGV (HARNESS_BACKTRACE=1): there is no debugger in the build image.
static void harness_segv(int) {
    void* frames[64];
    const int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
__attribute__((constructor)) static void harness_install_segv() {
    if (getenv("HARNESS_BACKTRACE")) signal(SIGSEGV, harness_segv);
}

extern "C" {
typedef int cudaError_t_;  // cudaError_t is a C enum: int-sized in the ABI
struct FakeHandle { int dummy; };

// ---- bookkeeping for the tests: live allocations (leak check) and one-shot fault injection -----------------------------
static std::atomic<long> g_live{0};          // device + pinned allocations not yet freed
static std::atomic<long> g_fail_in[3];       // > 0: fail the n-th call from now of {cudaMalloc, cudaHostAlloc, cudaMemcpyAsync}
static bool inject(int what) {
    long v = g_fail_in[what].load();
    while (v > 0)
        if (g_fail_in[what].compare_exchange_weak(v, v - 1)) return v == 1;
    return false;
}
static std::atomic<unsigned long long> g_copy_ns{0}, g_copy_bytes{0}, g_copy_calls{0};  // what the copies cost on the host (profiling aid)
long harness_live_allocations(void) { return g_live.load(); }
unsigned long long harness_copy_ns(void) { return g_copy_ns.load(); }
unsigned long long harness_copy_bytes(void) { return g_copy_bytes.load(); }
unsigned long long harness_copy_calls(void) { return g_copy_calls.load(); }
static std::atomic<unsigned long long> g_alloc_ns{0};
unsigned long long harness_alloc_ns(void) { return g_alloc_ns.load(); }
void harness_fail_nth(int what, long n) { g_fail_in[what].store(n); }

static void* fake_alloc(size_t n) {
    struct T { std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
               ~T() { g_alloc_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); } } timed;
    void* p = nullptr;
    if (posix_memalign(&p, 256, n ? n : 1) != 0) return nullptr;
    memset(p, 0xCD, n);  // poison: the operator must not rely on zeroed allocations
    g_live.fetch_add(1);
    return p;
}
static void fake_free(void* p) {
    if (p) g_live.fetch_sub(1);
    free(p);
}

cudaError_t_ cudaSetDevice(int) { return 0; }
cudaError_t_ cudaMalloc(void** p, size_t n) {
    *p = nullptr;
    if (inject(0)) return 2;  // cudaErrorMemoryAllocation
    *p = fake_alloc(n);
    return *p ? 0 : 2;
}
cudaError_t_ cudaFree(void* p) { fake_free(p); return 0; }
cudaError_t_ cudaHostAlloc(void** p, size_t n, unsigned) {
    *p = nullptr;
    if (inject(1)) return 2;
    *p = fake_alloc(n);
    return *p ? 0 : 2;
}
cudaError_t_ cudaFreeHost(void* p) { fake_free(p); return 0; }
cudaError_t_ cudaMemcpyAsync(void* dst, const void* src, size_t n, int, void*) {
    if (inject(2)) return 719;  // cudaErrorLaunchFailure: a sticky device error
    const auto t0 = std::chrono::steady_clock::now();
    if (n) memmove(dst, src, n);
    g_copy_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    g_copy_bytes += n;
    g_copy_calls += 1;
    return 0;
}
cudaError_t_ cudaMemsetAsync(void* dst, int v, size_t n, void*) { if (n) memset(dst, v, n); return 0; }
cudaError_t_ cudaStreamCreateWithFlags(void** s, unsigned) { *s = new FakeHandle(); return 0; }
cudaError_t_ cudaStreamDestroy(void* s) { delete (FakeHandle*)s; return 0; }
cudaError_t_ cudaStreamSynchronize(void*) { return 0; }
cudaError_t_ cudaStreamWaitEvent(void*, void*, unsigned) { return 0; }
cudaError_t_ cudaEventCreateWithFlags(void** e, unsigned) { *e = new FakeHandle(); return 0; }
cudaError_t_ cudaEventDestroy(void* e) { delete (FakeHandle*)e; return 0; }
cudaError_t_ cudaEventRecord(void*, void*) { return 0; }
cudaError_t_ cudaEventQuery(void*) { return 0; }
cudaError_t_ cudaEventSynchronize(void*) { return 0; }
const char* cudaGetErrorString(cudaError_t_) { return "fake CUDA runtime (CPU test harness)"; }
// ---- what the exchange (csrc/dfd_exchange.cu) needs on top: IPC handles, synchronous copies, kernel launches ------------
struct FakeIpcHandle { char reserved[64]; };  // cudaIpcMemHandle_t: the "handle" is the pointer (all workers are threads of one process)
cudaError_t_ cudaIpcGetMemHandle(FakeIpcHandle* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return 0; }
cudaError_t_ cudaIpcOpenMemHandle(void** p, FakeIpcHandle h, unsigned) { memcpy(p, h.reserved, sizeof *p); return 0; }
cudaError_t_ cudaIpcCloseMemHandle(void*) { return 0; }
cudaError_t_ cudaMemcpy(void* dst, const void* src, size_t n, int) { if (n) memmove(dst, src, n); return 0; }
cudaError_t_ cudaMemset(void* dst, int v, size_t n) { if (n) memset(dst, v, n); return 0; }
cudaError_t_ cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, int, void*) {
    for (size_t r = 0; r < height; ++r) memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    return 0;
}
cudaError_t_ cudaEventCreate(void** e) { *e = new FakeHandle(); return 0; }
cudaError_t_ cudaEventElapsedTime(float* ms, void*, void*) { *ms = 0.f; return 0; }

// Kernel launches: nvcc's host stubs call __cudaPushCallConfiguration / __cudaPopCallConfiguration / cudaLaunchKernel with the
// host-side function pointer it registered under the kernel's (mangled) device name.  The harness registers CPU
// emulations by name fragment (harness_register_kernel); a launch of a kernel without one fails with
// cudaErrorNotSupported, which the library sees through cudaGetLastError like any launch failure.
struct FakeDim3 { unsigned x, y, z; };
typedef int (*HarnessKernel)(FakeDim3 grid, FakeDim3 block, void** args);
struct KernelEntry { const void* host_fun; char name[256]; HarnessKernel cpu; };
static KernelEntry g_kernels[256];
static std::atomic<int> g_n_kernels{0};
struct CpuKernel { char fragment[64]; HarnessKernel fn; };
static CpuKernel g_cpu_kernels[32];
static std::atomic<int> g_n_cpu_kernels{0};
static thread_local cudaError_t_ t_last_error = 0;
static thread_local struct { FakeDim3 grid, block; size_t smem; void* stream; } t_config;

void harness_register_kernel(const char* name_fragment, HarnessKernel fn) {
    const int i = g_n_cpu_kernels.fetch_add(1);
    strncpy(g_cpu_kernels[i].fragment, name_fragment, sizeof g_cpu_kernels[i].fragment - 1);
    g_cpu_kernels[i].fn = fn;
    for (int k = 0; k < g_n_kernels.load(); ++k)  // (fat binaries may have registered their kernels before this constructor ran)
        if (!g_kernels[k].cpu && strstr(g_kernels[k].name, name_fragment)) g_kernels[k].cpu = fn;
}
void __cudaRegisterFunction(void**, const char* host_fun, char*, const char* device_name, int, void*, void*, void*, void*, int*) {
    const int i = g_n_kernels.fetch_add(1);
    if (i >= 256) return;
    g_kernels[i].host_fun = host_fun;
    strncpy(g_kernels[i].name, device_name, sizeof g_kernels[i].name - 1);
    g_kernels[i].cpu = nullptr;
    for (int k = 0; k < g_n_cpu_kernels.load(); ++k)
        if (strstr(device_name, g_cpu_kernels[k].fragment)) g_kernels[i].cpu = g_cpu_kernels[k].fn;
}
unsigned __cudaPushCallConfiguration(FakeDim3 grid, FakeDim3 block, size_t smem, void* stream) {
    t_config.grid = grid; t_config.block = block; t_config.smem = smem; t_config.stream = stream;
    return 0;
}
cudaError_t_ __cudaPopCallConfiguration(FakeDim3* grid, FakeDim3* block, size_t* smem, void** stream) {
    *grid = t_config.grid; *block = t_config.block; *smem = t_config.smem; *stream = t_config.stream;
    return 0;
}
cudaError_t_ cudaLaunchKernel(const void* func, FakeDim3 grid, FakeDim3 block, void** args, size_t, void*) {
    for (int k = 0; k < g_n_kernels.load(); ++k)
        if (g_kernels[k].host_fun == func) {
            if (!g_kernels[k].cpu) break;
            const int rc = g_kernels[k].cpu(grid, block, args);
            if (rc) t_last_error = rc;
            return rc;
        }
    t_last_error = 801;  // cudaErrorNotSupported: no CPU emulation of this kernel in the harness
    return 801;
}
cudaError_t_ cudaGetLastError(void) { const cudaError_t_ e = t_last_error; t_last_error = 0; return e; }

// fat-binary registration emitted by nvcc for every .cu object: nothing to register
void** __cudaRegisterFatBinary(void*) { static void* h = nullptr; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
}
