// ptx.cuh -- 32-bit carry-chain primitives.
//
// Device build (nvcc, sm_100a): one inline-PTX instruction per wrapper.  ptxas fuses each
// `mad(c).lo.cc` / `madc.hi.cc` pair on the same operands into ONE `IMAD.WIDE.U32(.X)` with the
// carry travelling in a predicate register, which is what the field multiplier is built from.
//
// Host-sim build (g++ -DPLUME_HOSTSIM, tests only): the same wrappers emulate the carry flag in
// a thread-local variable so that the *identical* limb-level code paths can be checked on a
// machine without a GPU.  The host-sim build is never linked into libplume_b200.so.
#pragma once
#include <stdint.h>

#ifdef PLUME_HOSTSIM
#define PLUME_DEV static inline
#define PLUME_DEV_NOINLINE static
#define PLUME_DEV_MEMBER inline
static thread_local uint32_t plume_cc_ = 0;
PLUME_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; plume_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
PLUME_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + plume_cc_; plume_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
PLUME_DEV uint32_t addc(uint32_t a, uint32_t b) { return a + b + plume_cc_; }
PLUME_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; plume_cc_ = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
PLUME_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - plume_cc_; plume_cc_ = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
PLUME_DEV uint32_t subc(uint32_t a, uint32_t b) { return a - b - plume_cc_; }
PLUME_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
PLUME_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
PLUME_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c; plume_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
PLUME_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c + plume_cc_; plume_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
PLUME_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c; plume_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
PLUME_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + plume_cc_; plume_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
PLUME_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + plume_cc_; }
PLUME_DEV uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }
PLUME_DEV uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
#else
#define PLUME_DEV __device__ __forceinline__
#define PLUME_DEV_NOINLINE static __device__ __noinline__
#define PLUME_DEV_MEMBER __device__ __forceinline__
PLUME_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLUME_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLUME_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLUME_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLUME_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLUME_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLUME_DEV uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
PLUME_DEV uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
#endif
