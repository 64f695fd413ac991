// cuda_runtime.h -- a HOST stand-in for the CUDA runtime and the handful of device built-ins the reference-order path uses.
// TEST INFRASTRUCTURE (tests/hostcheck): lets g++ compile engine.cu, kernel_reference.cu, mcpe.cu, stepgen.cu and tabulate.cu
// UNMODIFIED (one syntax rewrite: the three `kernel<<<grid, block, shmem, stream>>>(args)` launches become
// HOSTCHECK_LAUNCH(kernel, grid, block, args), see tests/hostcheck/__init__.py), so that the SOURCE of everything but the fast
// kernel can be run against the oracle on a machine without a GPU.  This is a test of the source text, not a way to run
// the product: nothing under clsim_b200/ knows of it, libclsimcuda.so has no CPU path, and the fast kernel (inline PTX, warp
// collectives, shared memory) is not part of it -- its launcher here reports an error.
//
// Semantics: every stream operation completes before the call returns (streams are program order, which the engine's
// event waits already respect); a kernel runs its threads one after the other, block by block (the kernels compiled here
// have no barrier and no shared memory; a warp collective sees a warp of one lane); device memory is host memory.
#ifndef CLSIM_HOSTCHECK_CUDA_RUNTIME_H
#define CLSIM_HOSTCHECK_CUDA_RUNTIME_H
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

// ---- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__ __restrict

// ---- runtime types
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef struct HostcheckStream *cudaStream_t;
typedef struct HostcheckEvent { std::chrono::steady_clock::time_point at; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned X = 1, unsigned Y = 1, unsigned Z = 1) : x(X), y(Y), z(Z) {}
};

// ---- runtime calls: memory is memory, streams are program order
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "hostcheck error"; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = static_cast<T *>(std::calloc(std::max<size_t>(n, 1), 1)); return *p ? cudaSuccess : cudaErrorUnknown; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { *p = static_cast<T *>(std::calloc(std::max<size_t>(n, 1), 1)); return *p ? cudaSuccess : cudaErrorUnknown; }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaHostAlloc(p, n, 0); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = reinterpret_cast<cudaStream_t>(std::malloc(1)); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, 0); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new HostcheckEvent(); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->at = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->at - a->at).count();
    return cudaSuccess;
}

// ---- the executing thread's coordinates, and a launch = a loop
struct HostcheckIdx { unsigned x, y, z; };
extern thread_local HostcheckIdx threadIdx, blockIdx, blockDim, gridDim;
#define HOSTCHECK_LAUNCH(kernel, grid, block, ...)                                   \
    do {                                                                              \
        const dim3 hc_grid(grid), hc_block(block);                                    \
        gridDim = HostcheckIdx{hc_grid.x, 1, 1};                                      \
        blockDim = HostcheckIdx{hc_block.x, 1, 1};                                    \
        for (unsigned hc_b = 0; hc_b < hc_grid.x; ++hc_b)                             \
            for (unsigned hc_t = 0; hc_t < hc_block.x; ++hc_t) {                      \
                blockIdx = HostcheckIdx{hc_b, 0, 0};                                  \
                threadIdx = HostcheckIdx{hc_t, 0, 0};                                 \
                kernel(__VA_ARGS__);                                                  \
            }                                                                         \
    } while (0)

// ---- device built-ins
template <class T> inline T __ldg(const T *p) { return *p; }
inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float __uint2float_rz(unsigned u)
{
    int lead = 0;
    for (unsigned t = u; t >>= 1;) ++lead;
    if (lead > 23) u &= ~0u << (lead - 23);
    return static_cast<float>(u);
}
inline int __float2int_rd(float v)
{
    const float f = std::floor(v);
    if (f != f) return 0;
    if (f >= 2147483648.f) return 2147483647;
    if (f <= -2147483648.f) return -2147483647 - 1;
    return static_cast<int>(f);
}
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return static_cast<unsigned long long>((static_cast<unsigned __int128>(a) * b) >> 64); }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<unsigned long long>(a) * b) >> 32); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float v) { return std::exp(v); }
inline float __logf(float v) { return std::log(v); }
inline float rsqrtf(float v) { return 1.0f / std::sqrt(v); }
inline float __saturatef(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }
inline void sincosf_hc(float a, float *s, float *c) { *s = std::sin(a); *c = std::cos(a); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
// a warp of one lane (the lane's own number decides its bit, as on the device)
inline unsigned __ballot_sync(unsigned, int pred) { return pred ? (1u << (threadIdx.x & 31u)) : 0u; }
template <class T> inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
inline int __any_sync(unsigned, int pred) { return pred != 0; }
inline int __all_sync(unsigned, int pred) { return pred != 0; }
inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }
inline void __syncwarp(unsigned = 0xffffffffu) {}
template <class T> inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float *p, float v)
{
    float old = *p, want;
    do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
inline double atomicAdd(double *p, double v)
{
    double old = *p, want;
    do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
template <class T> inline T atomicMax(T *p, T v) { T old = *p; while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return old;
}
using std::max;
using std::min;
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif
