// mbm_device.cuh — device-side building blocks of the batched time-stepping kernels (sm_100a, FP64).
//
// Two arithmetic regimes:
//  * the INTEGER-DETERMINING CHAIN (dt, t, Ri, node counts, mesh generation, Newton, pchip interval
//    search) is written with the never-contracted intrinsics __dmul_rn/__dadd_rn/__dsub_rn and IEEE
//    division, operation for operation as oracle/mbm_oracle.c does, so step counts, node counts and
//    regrid indices are bit-identical to the oracle;
//  * VALUE code (assembly, tridiagonal solve, pchip evaluation) uses plain operators (FMA contraction
//    allowed) and its own elimination order; it is held to 1e-10 relative against the oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mbm {

// ---- exact (never contracted) arithmetic ---------------------------------------------------------
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xinv(double a) { return __ddiv_rn(1.0, a); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// Base.Math.pow_body(x::Float64, n::Integer) (Julia 1.10 base/math.jl), as in oracle jl_pow_int
__device__ __noinline__ double jl_pow_int(double x, long long n) {
  if (n == 0) return 1.0;
  double y = 1.0, xnlo = 0.0, ynlo = 0.0;
  if (n == 3) return xmul(xmul(x, x), x);
  if (n < 0) {
    double rx = xinv(x);
    if (n == -2) return xmul(rx, rx);
    if (isfinite(x)) xnlo = xmul(-fma(x, rx, -1.0), rx);
    x = rx;
    n = -n;
  }
  int m = (int)n;                      // node counts: 32-bit loop control (n < 2^31 is guaranteed by the capacities)
  if (n > 0x7fffffffll) m = 0x7fffffff;
  while (m > 1) {
    if (m & 1) {
      double err = fma(y, xnlo, xmul(x, ynlo));
      double hi = xmul(x, y);
      double lo = fma(x, y, -hi);
      y = hi;
      ynlo = xadd(lo, err);
    }
    double err2 = xmul(xmul(x, 2.0), xnlo);
    double hi2 = xmul(x, x);
    double lo2 = fma(x, x, -hi2);
    x = hi2;
    xnlo = xadd(lo2, err2);
    m >>= 1;
  }
  double err = fma(y, xnlo, xmul(x, ynlo));
  if (isfinite(x) && isfinite(err)) return fma(x, y, err);
  return xmul(x, y);
}

__device__ __forceinline__ double pow_ndim(double x, int ndim) {
  if (ndim == 1) return x;
  if (ndim == 2) return x * x;
  return x * x * x;
}

// LinRange(a,b,len)[i], 0-based (Base.lerpi)
// (node counts are 32-bit: the int -> double conversions are single instructions and exact)
__device__ __forceinline__ double linrange_at_inl(double a, double b, int len, int i) {
  const int d = (len - 1 > 1) ? (len - 1) : 1;
  const double t = xdiv((double)i, (double)d);
  return xadd(xmul(xsub(1.0, t), a), xmul(t, b));
}
__device__ __noinline__ double linrange_at(double a, double b, int len, int i) { return linrange_at_inl(a, b, len, i); }

// x^n and x^(n-1) in one pass: the two pow_body calls of a Newton iteration square the same x, so the (x, xnlo)
// sequence is computed once and only the two y accumulations differ.  Every operation of either call is executed
// with the same operands in the same order as in jl_pow_int, so both results are bit-identical to separate calls.
// Requires 4 < n < 2^31 and n not a power of two (then n and n-1 have the same bit length and both loops run the same
// number of rounds); other exponents use jl_pow_int (pow_body's special cases / different round counts).
__device__ __noinline__ void jl_pow_pair(double x, int n, double *pn, double *pn1) {
  double ya = 1.0, yalo = 0.0, yb = 1.0, yblo = 0.0, xnlo = 0.0;
  int ma = n, mb = n - 1;
  while (ma > 1) {
    if (ma & 1) {
      const double err = fma(ya, xnlo, xmul(x, yalo));
      const double hi = xmul(x, ya);
      const double lo = fma(x, ya, -hi);
      ya = hi; yalo = xadd(lo, err);
    }
    if (mb & 1) {
      const double err = fma(yb, xnlo, xmul(x, yblo));
      const double hi = xmul(x, yb);
      const double lo = fma(x, yb, -hi);
      yb = hi; yblo = xadd(lo, err);
    }
    const double err2 = xmul(xmul(x, 2.0), xnlo);
    const double hi2 = xmul(x, x);
    const double lo2 = fma(x, x, -hi2);
    x = hi2; xnlo = xadd(lo2, err2);
    ma >>= 1; mb >>= 1;
  }
  const double ea = fma(ya, xnlo, xmul(x, yalo)), eb = fma(yb, xnlo, xmul(x, yblo));
  const bool fx = isfinite(x);
  *pn = (fx && isfinite(ea)) ? fma(x, ya, ea) : xmul(x, ya);
  *pn1 = (fx && isfinite(eb)) ? fma(x, yb, eb) : xmul(x, yb);
}

// newton_solver, reference src/functions.jl:732-757 (operation for operation as the oracle)
__device__ __noinline__ int newton_solver(double S, double d, long long n, double x0, double *root, int *warn) {
  double x = x0, Res = 1e23;
  const double Sd = xmul(S, xinv(d));
  for (int i = 1; i <= 200; ++i) {
    double xn, xn1;
    if (n > 4 && n < 0x7fffffffll && (n & (n - 1)) != 0) jl_pow_pair(x, (int)n, &xn, &xn1);
    else { xn = jl_pow_int(x, n); xn1 = jl_pow_int(x, n - 1); }
    double omx = xsub(1.0, x);
    double Fx = xsub(xmul(xsub(1.0, xn), xinv(omx)), Sd);
    double dF = xmul(fma(xmul(-(double)n, xn1), omx, xsub(1.0, xn)), xinv(xmul(omx, omx)));
    double idF = xinv(dF);
    Res = fabs(Fx);
    if (Res < 1e-8) { *root = x; return 0; }
    x = xsub(x, xmul(xmul(1.0, Fx), idF));
  }
  *root = x;
  if (Res > 1e-3) return 3;  // MBM_ERR_NEWTON
  *warn |= 1;                // MBM_WARN_NEWTON_RES
  return 0;
}

// order-independent checksum term of one regrid decision (same as oracle idx_term)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long idx_term(long long it, int phase, long long j, long long k, int branch) {
  unsigned long long key = ((unsigned long long)it << 40) ^ ((unsigned long long)phase << 39) ^
                           ((unsigned long long)j << 18) ^ ((unsigned long long)k << 2) ^ (unsigned long long)branch;
  return mix64(key);
}

// fast reciprocal for value code (~1 ulp): MUFU.RCP64H seed, then r (1 + e + e^2)
__device__ __forceinline__ double frcp_(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  // seed error e ~ 2^-20: r (1 + e + e^2) leaves e^3 ~ 2^-60, one dependent FMA less than two Newton steps
  const double e = fma(-x, r, 1.0);
  return fma(fma(e, e, e), r, r);
}

// ---- element matrices, reference src/functions.jl:525-540 (value code) ----------------------------
// returns A11 = M11/dt, A12 = M12/dt, A22 = M22/dt and kk = K12 (K = [-kk kk; kk -kk])
__device__ __forceinline__ void elem_mats(double x1, double x2, double D, double idt, int ndim, double &A11,
                                          double &A12, double &A22, double &kk) {
  const double df = x1 - x2;
  if (ndim == 3) {
    const double x1s = x1 * x1, x2s = x2 * x2, x12 = x1 * x2;
    const double c30 = idt * (1.0 / 30.0), c60 = idt * (1.0 / 60.0);
    A11 = -(df * (6.0 * x1s + 3.0 * x12 + x2s)) * c30;
    A12 = -(df * (3.0 * x1s + 4.0 * x12 + 3.0 * x2s)) * c60;
    A22 = -(df * (x1s + 3.0 * x12 + 6.0 * x2s)) * c30;
    kk = (D * (x1s + x12 + x2s)) * frcp_(3.0 * df);
  } else if (ndim == 2) {
    const double c12 = idt * (1.0 / 12.0);
    A11 = -(df * (3.0 * x1 + x2)) * c12;
    A12 = (x2 * x2 - x1 * x1) * c12;
    A22 = -(df * (x1 + 3.0 * x2)) * c12;
    kk = (D * (x1 + x2)) * frcp_(2.0 * df);
  } else {
    A11 = -df * (idt * (1.0 / 3.0));
    A12 = -df * (idt * (1.0 / 6.0));
    A22 = A11;
    kk = D * frcp_(df);
  }
}

// ---- warp / block reductions ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mbm
