// 32-bit carry-chain primitives for the Montgomery arithmetic.
// On the device every primitive is ONE PTX instruction using the condition-code carry flag
// (add.cc / addc.cc / mad.lo.cc / madc.hi.cc …); ptxas fuses adjacent mad.lo.cc + madc.hi.cc pairs
// into IMAD.WIDE with carry.  On the host (product host code that needs a few field operations per
// proof — challenges, the final window combination of an MSM — and the CPU emulation build used by
// the unit tests) the same primitives are emulated with an explicit carry variable, so the SAME limb
// algorithms run and are testable without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define ZP_DEVICE_CODE 1
#else
#define ZP_DEVICE_CODE 0
#endif

#if defined(__CUDACC__)
#define ZP_HD __host__ __device__ __forceinline__
#define ZP_D __device__ __forceinline__
#else
#define ZP_HD inline __attribute__((always_inline))
#define ZP_D inline __attribute__((always_inline))
#endif

namespace zp {

#if !ZP_DEVICE_CODE
// Host emulation of the PTX condition-code register. One per OS thread; a carry chain never spans a
// yield point of the emulator.
static thread_local uint32_t zp_cc_flag = 0;
#endif

ZP_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a + b;
    zp_cc_flag = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a + b + zp_cc_flag;
    zp_cc_flag = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t addc(uint32_t a, uint32_t b) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a + b + zp_cc_flag;
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    // PTX: CC.CF = borrow
    uint64_t s = (uint64_t)a - b;
    zp_cc_flag = (uint32_t)(s >> 63);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a - b - zp_cc_flag;
    zp_cc_flag = (uint32_t)(s >> 63);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t subc(uint32_t a, uint32_t b) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a - b - zp_cc_flag;
    return (uint32_t)s;
#endif
}
// r = lo(a*b) + c, sets carry
ZP_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c;
    zp_cc_flag = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + zp_cc_flag;
    zp_cc_flag = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (((uint64_t)a * b) >> 32) + c;
    zp_cc_flag = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (((uint64_t)a * b) >> 32) + c + zp_cc_flag;
    zp_cc_flag = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#if ZP_DEVICE_CODE
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (((uint64_t)a * b) >> 32) + c + zp_cc_flag;
    return (uint32_t)s;
#endif
}
ZP_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }

}  // namespace zp
