// exact_div.cuh -- correctly rounded fp32 division by a divisor whose reciprocal is known in advance.
//
// Both shaders divide by launch-uniform values in their hot loops ((p - b0) / (b1 - b0), x / baseScaling,
// x / maxExtinction, x / sigma, x / float(z + 1); PathTracer.comp:117,338,351, ShadowVolume.comp:46,89,92).
// div.rn.f32 costs ~10 instructions + one MUFU.RCP + a slow-path branch.  With r = RN(1/d) computed once
// (on the host, IEEE division) the quotient is obtained with one multiply and four FMAs:
//
//     q0 = RN(n * r)                      within ~2 ulp of n/d
//     q1 = RN(q0 + (n - q0*d) * r)        remainder is exact in an FMA; q1 is a faithful rounding of n/d
//     q  = RN(q1 + (n - q1*d) * r)        Markstein's theorem: faithful q1 + correctly rounded r  =>  q = RN(n/d)
//
// so the result is bit-identical to IEEE division (and to the CPU oracle) for every finite n/d that neither
// overflows nor falls into the subnormal range.  tests/test_exact_div.py checks the sequence against `/` on
// the CPU (same operations through fmaf) for the divisor families used here.
// Differences outside that range, all harmless at the call sites (each one is commented there):
//   n = -0       -> +0 instead of -0
//   n = +-inf    -> NaN instead of +-inf
#pragma once

struct ExactDivisor {
    float d;   // the divisor
    float r;   // RN(1 / d), computed on the host
};

__device__ __forceinline__ float div_rn_by(float n, float d, float r)
{
    float q = __fmul_rn(n, r);
    float e = __fmaf_rn(-q, d, n);
    q = __fmaf_rn(e, r, q);
    e = __fmaf_rn(-q, d, n);
    return __fmaf_rn(e, r, q);
}

__device__ __forceinline__ float div_rn_by(float n, const ExactDivisor& x) { return div_rn_by(n, x.d, x.r); }
