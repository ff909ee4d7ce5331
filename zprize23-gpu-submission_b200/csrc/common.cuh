// Shared plumbing for the sm_100a prover library: error handling, launch macro, device buffers.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <atomic>
#ifndef ZP_EMU
#include <cuda_runtime.h>
#endif
#include "field.cuh"

#ifdef ZP_EMU
#define ZP_LAUNCH(kern, grid, block, smem, stream, ...) do { emu::launch(grid, block, smem, [&]() { kern(__VA_ARGS__); }); zp::g_launch_count++; } while (0)
#define ZP_DYN_SMEM(T, name) T* name = (T*)emu::dyn_smem
#else
#define ZP_LAUNCH(kern, grid, block, smem, stream, ...)          \
    do {                                                          \
        kern<<<grid, block, smem, stream>>>(__VA_ARGS__);         \
        zp::check_cuda(cudaGetLastError(), #kern, __FILE__, __LINE__); \
        zp::g_launch_count++;                                     \
    } while (0)
#define ZP_DYN_SMEM(T, name)                                      \
    extern __shared__ __align__(16) unsigned char name##_raw[];   \
    T* name = reinterpret_cast<T*>(name##_raw)
#endif

#define ZP_CUDA(expr) zp::check_cuda((expr), #expr, __FILE__, __LINE__)

namespace zp {

extern std::atomic<unsigned long long> g_launch_count;  // kernels launched by this library (bench.py "gpu_launches")

// The reference exits the process on a CUDA failure (lib/caffe/common.hpp:23-30).  We throw; the C-ABI
// layer turns the exception into an error code / message (and, for the by-value `gen_proof` symbol that
// has no error channel, into the reference's behaviour: message + exit(1)).
inline void check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
        throw std::runtime_error(buf);
    }
}

// 32-byte / 96-byte device elements are moved as 16-byte vectors.
ZP_HD fr_t load_fr(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
ZP_HD void store_fr(fr_t* p, const fr_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
ZP_HD fq_t load_fq(const fq_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    fq_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
ZP_HD void store_fq(fq_t* p, const fq_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}

// Owning device buffer.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) ZP_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

static inline int ilog2(size_t x) {
    int l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

}  // namespace zp
