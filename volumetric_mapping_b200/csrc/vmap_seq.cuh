// vmap_seq.cuh -- exact jump-ahead for the DDA's tMax recurrence.
//
// computeRayKeys advances each axis with   tMax[i] += tDelta[i]   (octomap OcTreeBaseImpl.hxx,
// SURVEY 3.3 step 6): a SEQUENTIAL double accumulation, one rounding per step, whose values are
// not tMax0 + k*tDelta.  The three axes are independent sequences
//        t(0) = tMax0,   t(k+1) = fl(t(k) + dt),
// and the walk consumes their elements in increasing order.  Splitting a ray between threads
// therefore needs, for a threshold tau, the number of elements below tau and the first element
// at or above it -- exactly, without performing every addition.
//
// Inside one binade [2^e, 2^(e+1)) every double is a multiple of u = 2^(e-52).  For t = m*u the
// exact sum m*u + dt rounds to (m + q)*u with q = round(dt/u); q does not depend on m unless
// dt/u lies exactly half-way between two integers, and in that case round-half-to-even makes
// every result even, after which q is constant as well.  So after two in-binade additions the
// increment is a fixed integer number of ulps and the remaining elements of the binade follow by
// integer arithmetic.  seq_advance does a few real additions at each binade entry (which also
// carries it across binade boundaries, where the grid changes) and one integer jump per binade.
// tests/test_seq_closed_form.py checks it against the plain loop on millions of cases,
// including constructed ties and binade crossings.
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VMAP_HD __host__ __device__ __forceinline__
#else
#define VMAP_HD static inline
#endif

namespace vmapseq {

VMAP_HD double seq_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;   // one IEEE addition, never contracted
  return r;
#endif
}
VMAP_HD uint64_t seq_bits(double v) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(v);
#else
  uint64_t b;
  memcpy(&b, &v, 8);
  return b;
#endif
}
VMAP_HD double seq_from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

// Advance the sequence while its current element is < tau and fewer than max_steps elements
// have been consumed.  On entry *t is the current (not yet consumed) element; on return *t is
// the first element that was not consumed.  Returns the number of consumed elements.
// Requires dt > 0 and finite, and a strictly increasing sequence (dt >= ulp(t)), which holds
// for the DDA (dt >= resolution, t bounded by the ray length).
VMAP_HD uint64_t seq_advance(double* t_io, double dt, double tau, uint64_t max_steps) {
  double t = *t_io;
  uint64_t n = 0;
  while (n < max_steps && t < tau) {
    // (1) real additions: t1 is the successor of t, t2 of t1, t3 of t2
    const double t1 = seq_add(t, dt);
    if (n + 1 >= max_steps || !(t1 < tau)) { t = t1; n += 1; break; }
    const double t2 = seq_add(t1, dt);
    if (n + 2 >= max_steps || !(t2 < tau)) { t = t2; n += 2; break; }
    const double t3 = seq_add(t2, dt);
    if (n + 3 >= max_steps || !(t3 < tau)) { t = t3; n += 3; break; }
    // consumed so far: t, t1, t2 (t3 is current).  Integer jump only when t1, t2, t3 are
    // positive normal numbers of one binade: then t2 came from an in-binade addition (so it is
    // even in the tie case) and t3 - t2 is the constant increment.
    const uint64_t b1 = seq_bits(t1), b2 = seq_bits(t2), b3 = seq_bits(t3);
    const uint64_t e1 = b1 >> 52, e2 = b2 >> 52, e3 = b3 >> 52;   // sign bit included: 0 for t > 0
    n += 3;
    t = t3;
    if (e1 != e2 || e2 != e3 || e3 == 0 || e3 >= 0x7FE) continue;
    const uint64_t q = b3 - b2;                 // increment in ulps (same exponent: plain integer difference)
    if (q == 0) continue;
    // Elements b3 + j*q (j >= 1) stay in the binade while <= top and are below tau while
    // <= bits(tau) - 1 (when tau lies in this binade).  Take the largest j with b3 + j*q <= lim,
    // plus one when that lands on the first element >= tau that is still inside the binade.
    const uint64_t top = (b3 | 0x000FFFFFFFFFFFFFull);
    const uint64_t bt = seq_bits(tau);
    const bool tau_here = (tau > 0 && bt <= top);     // tau > t3 here, so it is in this binade
    const uint64_t lim = tau_here ? bt - 1 : top;
    // one division: only an ESTIMATE of num / q is needed -- the two loops below correct it exactly.
    // Quotients below 2^20 (every DDA: j is a step count) take a single-precision estimate (one
    // MUFU.RCP + one multiply on the device instead of the ~40-instruction IEEE double division:
    // relative error < 2^-21, so the estimate is off by at most one); anything larger takes the
    // double division (both operands < 2^52 are exact).
    const uint64_t num = lim - b3;
    uint64_t j;
    if ((num >> 20) < q) {
#if defined(__CUDA_ARCH__)
      j = (uint64_t)__float2uint_rz(__fdividef(__ull2float_rz(num), __ull2float_rn(q)));
#else
      j = (uint64_t)(uint32_t)((float)num / (float)q);
#endif
    } else {
      j = (uint64_t)((double)num / (double)q);
    }
    while (j * q > num) --j;
    while ((j + 1) * q <= num) ++j;
    if (tau_here && b3 + (j + 1) * q <= top) ++j;     // step onto the first element >= tau
    if (j == 0) continue;
    if (j > max_steps - n) j = max_steps - n;
    t = seq_from_bits(b3 + j * q);
    n += j;
  }
  *t_io = t;
  return n;
}

// (Measured and dropped, round 2: plain additions for the first 40 elements before the jump -- two
// instructions per element, but a DEPENDENT chain of them: k_checkpoints 30 -> 41 us.)

}  // namespace vmapseq
