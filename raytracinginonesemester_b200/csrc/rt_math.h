// rt_math.h — exactly-rounded float3 algebra shared by every kernel.
//
// The reference CPU build (x86-64, no FMA contraction) fixes the rounding of every
// intermediate: dot = (x*x' + y*y') + z*z', cross as three a*b - c*d pairs, three
// divides in unit_vector (HW1/include/vec3.h:46-56, GPUandCPU/include/vec3.h:45-58).
// nvcc fuses a*b+c into FFMA by default, which changes closest-hit ids on edge pixels
// (SURVEY §7 H1), so everything on the parity-critical path goes through the X* macros:
// on the device they are the round-to-nearest intrinsics ptxas never contracts; in a
// host pass they are plain operators (host objects are built with -ffp-contract=off).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define XADD(a, b) __fadd_rn((a), (b))
#define XSUB(a, b) __fsub_rn((a), (b))
#define XMUL(a, b) __fmul_rn((a), (b))
#define XDIV(a, b) __fdiv_rn((a), (b))
#define XSQRT(a)   __fsqrt_rn((a))
#define XRCP(a)    __fdiv_rn(1.0f, (a))  // (__frcp_rn gives the same correctly rounded value but measured 2 % slower on the frame kernel: range check + branch first)
// glibc powf is computed in double and is correctly rounded in all but ~1e-8 of cases;
// an fp64 pow narrowed to float reproduces it (CUDA powf is 4+ ulp off).
#define XPOW(a, b) ((float)pow((double)(a), (double)(b)))
#else
#define XADD(a, b) ((a) + (b))
#define XSUB(a, b) ((a) - (b))
#define XMUL(a, b) ((a) * (b))
#define XDIV(a, b) ((a) / (b))
#define XSQRT(a)   sqrtf((a))
#define XRCP(a)    (1.0f / (a))
#define XPOW(a, b) powf((a), (b))
#endif

#if defined(__CUDA_ARCH__)
#define RT_LDG(p) __ldg(p)
#define RT_F2I(f) __float_as_int(f)
#define RT_I2F(i) __int_as_float(i)
#define RT_CLZ32(x) __clz((int)(x))
#define RT_CLZ64(x) __clzll((long long)(x))
#define RT_U2F(u) __uint2float_rn(u)
#define RT_D2F_UP(d) __double2float_ru(d)
#else
#include <string.h>
#define RT_LDG(p) (*(p))
static inline int rt_f2i_host(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float rt_i2f_host(int i) { float f; memcpy(&f, &i, 4); return f; }
#define RT_F2I(f) rt_f2i_host(f)
#define RT_I2F(i) rt_i2f_host(i)
#define RT_CLZ32(x) ((x) == 0 ? 32 : __builtin_clz((unsigned)(x)))
#define RT_CLZ64(x) ((x) == 0 ? 64 : __builtin_clzll((unsigned long long)(x)))
#define RT_U2F(u) ((float)(u))
static inline float rt_d2f_up_host(double d) { float f = (float)d; return ((double)f < d) ? nextafterf(f, INFINITY) : f; }
#define RT_D2F_UP(d) rt_d2f_up_host(d)
#endif

struct f3 { float x, y, z; };
#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }
#endif

RT_HD f3 mk3(float x, float y, float z) { f3 v; v.x = x; v.y = y; v.z = z; return v; }
RT_HD f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
RT_HD f3 xadd3(f3 a, f3 b) { return mk3(XADD(a.x, b.x), XADD(a.y, b.y), XADD(a.z, b.z)); }
RT_HD f3 xsub3(f3 a, f3 b) { return mk3(XSUB(a.x, b.x), XSUB(a.y, b.y), XSUB(a.z, b.z)); }
RT_HD f3 xneg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_HD f3 xmulv(f3 a, f3 b) { return mk3(XMUL(a.x, b.x), XMUL(a.y, b.y), XMUL(a.z, b.z)); }
RT_HD f3 xmuls(f3 v, float t) { return mk3(XMUL(v.x, t), XMUL(v.y, t), XMUL(v.z, t)); }
// operator/(Vec3,double) with a float divisor: fp64 divide then narrow == fp32 divide
// (53 >= 2*24+2 bits, double rounding innocuous).
RT_HD f3 xdivs(f3 v, float t) { return mk3(XDIV(v.x, t), XDIV(v.y, t), XDIV(v.z, t)); }
RT_HD float xdot(f3 u, f3 v) { return XADD(XADD(XMUL(u.x, v.x), XMUL(u.y, v.y)), XMUL(u.z, v.z)); }
RT_HD f3 xcross(f3 u, f3 v) {
    return mk3(XSUB(XMUL(u.y, v.z), XMUL(u.z, v.y)),
               XSUB(XMUL(u.z, v.x), XMUL(u.x, v.z)),
               XSUB(XMUL(u.x, v.y), XMUL(u.y, v.x)));
}
RT_HD float xlen3(f3 v) { return XSQRT(xdot(v, v)); }
// The three unit-vector routines are NOT inlined in device code: each is an IEEE square root and three IEEE divides
// (~45 SASS instructions with their range checks), shading calls them eight times per hit, and the frame kernel is
// instruction-cache bound outside its traversal loop (ncu r2: 19 % of the warp stalls were instruction fetches, almost
// all of them in straight-line shading code that every packet streams through once).  One copy, eight calls.
#ifndef RT_X_NOINLINE_UNIT
#define RT_X_NOINLINE_UNIT 0
#endif
#if defined(__CUDA_ARCH__) && RT_X_NOINLINE_UNIT
#define RT_UNIT_FN static __device__ __noinline__
#else
#define RT_UNIT_FN RT_HD
#endif
// global unit_vector(): vec3.h:53-56
RT_UNIT_FN f3 xunit(f3 v) { return xdivs(v, xlen3(v)); }
// unit_vector() of the CPUOnly renderer: zero vector below 1e-12 (HW2/HW2/CPUOnly/include/vec3.h:53-57)
RT_UNIT_FN f3 xunit_c(f3 v) {
    float len = xlen3(v);
    if (len < 1e-12f) return mk3(0.0f, 0.0f, 0.0f);
    return xdivs(v, len);
}
// Camera::unit_vector(): GPUandCPU/include/camera.h:64-69 (fallback (0,0,1) below 1e-12)
RT_UNIT_FN f3 xunit_cam(f3 v) {
    float len = xlen3(v);
    if ((double)len < 1e-12) return mk3(0.0f, 0.0f, 1.0f);
    return xdivs(v, len);
}
