// TEST INFRASTRUCTURE ONLY — a tiny single-threaded emulation of the CUDA execution model so that the
// product's .cu sources (kernels AND host driver) can be compiled with g++ and exercised on a machine
// without a GPU (tests/test_emu_*.py).  It is NOT a fallback: the shipped library is built by nvcc only
// and refuses to run without a device.  Each CUDA thread of a block is a ucontext fiber; __syncthreads
// and warp shuffles are cooperative barriers; blocks run one after another.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ucontext.h>
#include <vector>
#include <functional>
#include <algorithm>

#define ZP_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __constant__ static

struct uint3_emu { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace emu {
extern uint3_emu threadIdx_, blockIdx_;
extern dim3 blockDim_, gridDim_;
extern char* dyn_smem;
void sync_block();
void sync_warp();
uint32_t shfl(uint32_t v, int src_lane);
uint32_t ballot(int pred);
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emu

#define threadIdx emu::threadIdx_
#define blockIdx emu::blockIdx_
#define blockDim emu::blockDim_
#define gridDim emu::gridDim_
#define warpSize 32

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int lane) { return emu::shfl(v, lane & 31); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) { return emu::shfl(v, (threadIdx.x & 31) ^ m); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) {
    int l = threadIdx.x & 31;
    return emu::shfl(v, l >= (int)d ? l - d : l);
}
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) {
    int l = threadIdx.x & 31;
    return emu::shfl(v, l + d < 32 ? l + d : l);
}
static inline uint32_t __ballot_sync(unsigned, int p) { return emu::ballot(p); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(uint32_t x) { return __builtin_ffs(x); }
static inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
static inline void __threadfence() {}
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }

// ---- runtime API subset ------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct EmuEvent { double t; }* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return 0; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = 0; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -1; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
struct cudaDeviceProp { int multiProcessorCount; char name[64]; int major, minor; size_t totalGlobalMem; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    p->multiProcessorCount = 4;
    strcpy(p->name, "zp-emu");
    p->major = 10;
    p->minor = 0;
    p->totalGlobalMem = (size_t)8 << 30;
    return 0;
}
