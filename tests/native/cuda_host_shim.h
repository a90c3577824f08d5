// Host stand-ins for the CUDA qualifiers and intrinsics used by the per-thread device code in
// volumetric_mapping_b200/csrc/{vmap_device,vmap_seq,vmap_walk}.cuh, so that the CPU test-suite
// can run that code (one "thread" at a time) against the plain DDA and the oracle.
// TEST INFRASTRUCTURE ONLY: nothing in the product includes this file.
// Build with  g++ -O2 -ffp-contract=off -DVMAP_HOST_EMULATION -include cuda_host_shim.h
// (x86-64 SSE2 arithmetic: every + - * / sqrt below is one correctly rounded IEEE operation).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)

using std::max;
using std::min;

static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { volatile double r = sqrt(a); return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
template <class T> static inline T __ldcg(const T* p) { return *p; }

// single "thread": atomics are plain read-modify-writes
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; *p = o < v ? o : v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
