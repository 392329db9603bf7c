// Correctly-rounded-in-practice fp64 sin/cos via double-double arithmetic.
//
// Why it exists: the reference computes beam endpoints with CPython's math.cos/math.sin,
// i.e. glibc's fp64 cos/sin (server_nodes/dual_bot_mapper.py:890-891, 901-902), and then
// truncates (wx - ox) / res to a cell index (:123-124).  glibc's results are correctly
// rounded except in rare near-midpoint cases (stated bound 0.55 ULP); CUDA's sincos() is a
// <=2 ULP routine.  A last-bit difference flips a cell index only when the quotient lies
// within ~1e-11 of an integer, so the integrate kernels run a screened fast path (CUDA's
// sincos(), see beam_expand.cuh) and re-evaluate with this routine only the beams whose
// quotient is that close to a cell boundary.  Everything stays on the device.
//
// The file is plain C++ when compiled without nvcc so the CPU test-suite can check it
// against mpmath (tests/test_host_expand.py builds tests/host_harness.cpp with g++).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OCC_HD __host__ __device__ __forceinline__
#define OCC_HD_MEMBER __host__ __device__ __forceinline__
#else
#define OCC_HD static inline
#define OCC_HD_MEMBER inline
#endif

namespace occ {

// Individually rounded fp64 ops.  On the device the _rn intrinsics are never contracted
// into FMAs; on the host the harness is built with -ffp-contract=off.
#if defined(__CUDA_ARCH__)
#define OCC_DADD(a, b) __dadd_rn((a), (b))
#define OCC_DMUL(a, b) __dmul_rn((a), (b))
#define OCC_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define OCC_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define OCC_DADD(a, b) ((a) + (b))
#define OCC_DMUL(a, b) ((a) * (b))
#define OCC_DFMA(a, b, c) fma((a), (b), (c))
#define OCC_DDIV(a, b) ((a) / (b))
#endif

struct dd { double hi, lo; };

OCC_HD dd two_sum(double a, double b) {           // exact: a + b = hi + lo
    double s = OCC_DADD(a, b);
    double bb = OCC_DADD(s, -a);
    double e = OCC_DADD(OCC_DADD(a, -OCC_DADD(s, -bb)), OCC_DADD(b, -bb));
    return dd{s, e};
}
OCC_HD dd quick_two_sum(double a, double b) {     // requires |a| >= |b|
    double s = OCC_DADD(a, b);
    return dd{s, OCC_DADD(b, -OCC_DADD(s, -a))};
}
OCC_HD dd two_prod(double a, double b) {          // exact: a * b = hi + lo
    double p = OCC_DMUL(a, b);
    return dd{p, OCC_DFMA(a, b, -p)};
}
OCC_HD dd dd_add(dd a, dd b) {                    // accurate (IEEE-style) dd + dd
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = OCC_DADD(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = OCC_DADD(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
OCC_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = OCC_DADD(s.lo, a.lo);
    return quick_two_sum(s.hi, s.lo);
}
OCC_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = OCC_DADD(p.lo, OCC_DADD(OCC_DMUL(a.hi, b.lo), OCC_DMUL(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}
OCC_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }

// 1/(2k+1)!, k = 1..14 and 1/(2k)!, k = 1..15 as double-doubles (generated with mpmath).
#if defined(__CUDA_ARCH__)
#define OCC_TABLE __device__ const
#else
#define OCC_TABLE static const
#endif
OCC_TABLE double kSinC[14][2] = {
    {0x1.5555555555555p-3, 0x1.5555555555555p-57},   {0x1.1111111111111p-7, 0x1.1111111111111p-63},
    {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73},  {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
    {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80}, {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
    {0x1.ae7f3e733b81fp-41, 0x1.1d8656b0ee8cbp-97},  {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
    {0x1.2f49b46814157p-57, 0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
    {0x1.761b41316381ap-75, -0x1.3423c7d91404fp-130}, {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},
    {0x1.d1ab1c2dccea3p-94, 0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157}};
OCC_TABLE double kCosC[15][2] = {
    {0x1.0000000000000p-1, 0.0},                     {0x1.5555555555555p-5, 0x1.5555555555555p-59},
    {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
    {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76},  {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
    {0x1.93974a8c07c9dp-37, 0x1.05d6f8a2efd1fp-92},  {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
    {0x1.6827863b97d97p-53, 0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
    {0x1.0ce396db7f853p-70, -0x1.aebcdbd20331cp-124}, {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},
    {0x1.88e85fc6a4e5ap-89, -0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153},
    {0x1.3932c5047d60ep-108, 0x1.832b7b530a627p-162}};

// Largest |x| the double-double reduction handles (k = rint(x * 2/pi) < 2^20 keeps the
// absolute reduction error below 2^-85).  Beyond it the caller keeps the library result.
#define OCC_SINCOS_DD_MAX 1.0e6

// sin(x), cos(x) rounded to nearest from a ~100-bit evaluation.  Returns false (outputs
// untouched) when |x| > OCC_SINCOS_DD_MAX or x is not finite.
OCC_HD bool sincos_dd(double x, double* s_out, double* c_out) {
    if (!(fabs(x) <= OCC_SINCOS_DD_MAX)) return false;
    // pi/2 as a triple-double
    const double P1 = 0x1.921fb54442d18p+0, P2 = 0x1.1a62633145c07p-54, P3 = -0x1.f1976b7ed8fbcp-110;
    double k = rint(OCC_DMUL(x, 0x1.45f306dc9c883p-1));
    dd r;
    if (k == 0.0) {
        r = dd{x, 0.0};
    } else {
        dd p1 = two_prod(k, P1);
        double s = OCC_DADD(x, -p1.hi);            // exact (Sterbenz: x/2 <= k*P1 <= 2x)
        dd t = two_sum(s, -p1.lo);
        dd p2 = two_prod(k, P2);
        t = dd_add(t, dd_neg(p2));
        r = dd_add_d(t, -OCC_DMUL(k, P3));
    }
    dd r2 = dd_mul(r, r);
    // sin r = r * (1 - r2/3! + r2^2/5! - ...): Horner from the highest term
    dd ps = dd{kSinC[13][0], kSinC[13][1]};
    for (int i = 12; i >= 0; --i) {
        ps = dd_mul(ps, r2);
        ps = dd_add(dd{kSinC[i][0], kSinC[i][1]}, dd_neg(ps));
    }
    // ps = 1/3! - r2/5! + ...  ->  sin = r - r*r2*ps
    dd sn = dd_add(r, dd_neg(dd_mul(dd_mul(r, r2), ps)));
    dd pc = dd{kCosC[14][0], kCosC[14][1]};
    for (int i = 13; i >= 0; --i) {
        pc = dd_mul(pc, r2);
        pc = dd_add(dd{kCosC[i][0], kCosC[i][1]}, dd_neg(pc));
    }
    // pc = 1/2! - r2/4! + ...  ->  cos = 1 - r2*pc
    dd cs = dd_add(dd{1.0, 0.0}, dd_neg(dd_mul(r2, pc)));
    double sv = OCC_DADD(sn.hi, sn.lo), cv = OCC_DADD(cs.hi, cs.lo);
    long long q = (long long)k & 3;
    double so, co;
    switch (q) {
        case 0: so = sv;  co = cv;  break;
        case 1: so = cv;  co = -sv; break;
        case 2: so = -sv; co = -cv; break;
        default: so = -cv; co = sv; break;
    }
    *s_out = so;
    *c_out = co;
    return true;
}

}  // namespace occ
