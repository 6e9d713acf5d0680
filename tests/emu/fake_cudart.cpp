// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the handful of CUDA runtime entry points the host side of the library
// uses, implemented on plain host memory, so that the WHOLE library (C ABI, host orchestration, kernels) can be built
// by g++ into tests/emu/_build/libmpm_b200_emu.so and driven by the GPU parity tests on small scenes without a GPU.
// "Device" memory is host memory, streams and events are no-ops (emulated launches are synchronous), graphs are refused.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof *p);
    std::strcpy(p->name, "host emulation (tests only)");
    p->multiProcessorCount = 1;          // persistent grids are sized from this: keep the emulated launches small
    p->major = 10; p->minor = 0;
    return cudaSuccess;
}
// With EMU_SHM_IPC=1 (the multi-process test of the peer-memory halo) "device" allocations of 64 KB and more live in POSIX
// shared memory, so that another emulated rank can map them: the IPC handle is the object's name.
struct ShmBlock { void* ptr; size_t bytes; char name[48]; };
static std::vector<ShmBlock>& shm_blocks() { static std::vector<ShmBlock> v; return v; }
static std::mutex& shm_mutex() { static std::mutex m; return m; }
static bool shm_ipc() { static const bool on = std::getenv("EMU_SHM_IPC") != nullptr; return on; }
cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (shm_ipc() && bytes >= (64u << 10)) {
        static int counter = 0;
        std::lock_guard<std::mutex> lk(shm_mutex());
        ShmBlock b;
        b.bytes = (bytes + 4095) / 4096 * 4096;
        std::snprintf(b.name, sizeof b.name, "/mpm_emu_%d_%d", (int)getpid(), counter++);
        const int fd = shm_open(b.name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)b.bytes) != 0) { if (fd >= 0) close(fd); return cudaErrorMemoryAllocation; }
        b.ptr = mmap(nullptr, b.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (b.ptr == MAP_FAILED) { shm_unlink(b.name); return cudaErrorMemoryAllocation; }
        std::memset(b.ptr, 0xcd, bytes);
        shm_blocks().push_back(b);
        *p = b.ptr;
        return cudaSuccess;
    }
    *p = std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
    if (*p) std::memset(*p, 0xcd, bytes);          // poison like the emulated shared memory
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
    if (p && shm_ipc()) {
        std::lock_guard<std::mutex> lk(shm_mutex());
        for (size_t i = 0; i < shm_blocks().size(); ++i)
            if (shm_blocks()[i].ptr == p) {
                munmap(p, shm_blocks()[i].bytes); shm_unlink(shm_blocks()[i].name);
                shm_blocks().erase(shm_blocks().begin() + (long)i);
                return cudaSuccess;
            }
    }
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t bytes) { *p = std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free((void*)s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free((void*)e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
// "IPC": with EMU_SHM_IPC=1 the handle is the name of the shared-memory object behind the allocation (other processes map
// it); otherwise it carries the pointer itself (one process).
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    std::memset(h, 0, sizeof *h);
    if (shm_ipc()) {
        std::lock_guard<std::mutex> lk(shm_mutex());
        for (const ShmBlock& b : shm_blocks())
            if (b.ptr == p) { std::memcpy(h, b.name, sizeof b.name); return cudaSuccess; }
        return cudaErrorInvalidValue;             // like the real call for a pointer that is not the base of an allocation
    }
    std::memcpy(h, &p, sizeof p);
    return cudaSuccess;
}
struct ShmMapping { void* ptr; size_t bytes; };
static std::vector<ShmMapping>& shm_mappings() { static std::vector<ShmMapping> v; return v; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    if (shm_ipc()) {
        char name[sizeof(cudaIpcMemHandle_t) + 1] = { 0 };
        std::memcpy(name, &h, sizeof h);
        const int fd = shm_open(name, O_RDWR, 0600);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) close(fd); return cudaErrorInvalidValue; }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
        std::lock_guard<std::mutex> lk(shm_mutex());
        shm_mappings().push_back(ShmMapping{ m, (size_t)st.st_size });
        *p = m;
        return cudaSuccess;
    }
    std::memcpy(p, &h, sizeof *p);
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
    if (shm_ipc()) {
        std::lock_guard<std::mutex> lk(shm_mutex());
        for (size_t i = 0; i < shm_mappings().size(); ++i)
            if (shm_mappings()[i].ptr == p) { munmap(p, shm_mappings()[i].bytes); shm_mappings().erase(shm_mappings().begin() + (long)i); break; }
    }
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { if (g) *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
}
