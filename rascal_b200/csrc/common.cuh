// Shared device helpers for the rascal_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

#include "../../include/rascal_b200.h"

#define RB2_CHECK(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return rb2_set_cuda_error(e__); \
  } while (0)

int rb2_set_cuda_error(cudaError_t e);

namespace rb2 {

// Unfused FP64 arithmetic: the reference computes every product and sum as a
// separate NumPy ufunc (two roundings); FMA contraction would change last bits
// and flip mask / bin decisions at edges.
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// np.linspace(start, stop, num)[i] with numpy's exact arithmetic
// (numpy/_core/function_base.py: y = arange(num)*step + start, y[-1] = stop).
struct Linspace {
  double start, stop, step, delta, div;
  int num;
  bool step_zero;
  __host__ __device__ Linspace() {}
  __host__ __device__ Linspace(double a, double b, int n) : start(a), stop(b), num(n) {
    div = (double)(n - 1);
    delta = b - a;
    step = (n > 1) ? delta / div : 0.0;
    step_zero = (n > 1) && (step == 0.0);
  }
  __device__ __forceinline__ double at(int i) const {
    if (num > 1 && i == num - 1) return stop;
    if (num <= 1) return start;  // 0*delta + start
    double y = (double)i;
    y = step_zero ? mul(y / div, delta) : mul(y, step);
    return add(y, start);
  }
};

// largest k in [0, n-1] with edges[k] <= v, where v == edges[n] lands in n-1:
// np.histogramdd's searchsorted(edges, v, 'right') - 1 with the right-edge rule.
// `k` is a starting guess; edges has n+1 entries.
__device__ __forceinline__ int fix_bin(const double* __restrict__ edges, int n, double v, int k) {
  k = max(0, min(n - 1, k));
  while (k > 0 && v < edges[k]) --k;
  while (k < n - 1 && v >= edges[k + 1]) ++k;
  return k;
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// numpy.polynomial.polynomial.polyval: Horner with separate multiply and add.
template <int N>
__device__ __forceinline__ double horner_unfused(const double* c, double x) {
  double r = c[N - 1];
#pragma unroll
  for (int i = N - 2; i >= 0; --i) r = add(c[i], mul(r, x));
  return r;
}
__device__ __forceinline__ double horner_unfused(const double* c, int n, double x) {
  double r = c[n - 1];
  for (int i = n - 2; i >= 0; --i) r = add(c[i], mul(r, x));
  return r;
}

// self.polyval of the fit (calibrator.py:1621-1631) on n coefficients: numpy's polyval (Horner), legval or
// chebval (Clenshaw recurrences), each with numpy's own sequence of separately rounded operations.
__device__ __forceinline__ double basis_val(int basis, const double* c, int n, double x) {
  if (basis == 0 || n <= 0) return (n <= 0) ? 0.0 : horner_unfused(c, n, x);
  double c0, c1;
  if (n == 1) { c0 = c[0]; c1 = 0.0; }
  else if (n == 2) { c0 = c[0]; c1 = c[1]; }
  else if (basis == 1) {  // numpy.polynomial.legendre.legval
    double nd = (double)n;
    c0 = c[n - 2]; c1 = c[n - 1];
    for (int i = 3; i <= n; ++i) {
      const double tmp = c0;
      nd = nd - 1.0;
      c0 = sub(c[n - i], mul(c1, nd - 1.0) / nd);
      c1 = add(tmp, mul(mul(c1, x), 2.0 * nd - 1.0) / nd);
    }
  } else {                // numpy.polynomial.chebyshev.chebval
    const double x2 = mul(2.0, x);
    c0 = c[n - 2]; c1 = c[n - 1];
    for (int i = 3; i <= n; ++i) {
      const double tmp = c0;
      c0 = sub(c[n - i], c1);
      c1 = add(tmp, mul(c1, x2));
    }
  }
  return add(c0, mul(c1, x));
}

}  // namespace rb2
