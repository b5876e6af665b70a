// mbm_validate.cuh — validation layer (include/mbm_b200.h, mbm_validate_batch): the known-answer gates of the
// reference's own tests evaluated per member from the ensemble's state, so that a sweep is checked where it lives and
// only [M] results cross the host link.  Not on the hot path (one pass over the final state): one thread per member,
// the operations in the order of the reference (test/test_C1.jl, test/test_B5.jl; src/Benchmarks.jl:239-260, :284-301;
// src/functions.jl:183-187, :205-212, :565-585, :1310-1323), value code.
#pragma once
#include "mbm_kernels.cuh"

namespace mbm {

struct VArgs {
  int M, ndim, capL, capR, kind;
  const double *x_left, *C_left, *x_right, *C_right;
  const long long *res;
  const double *Ri, *rho, *Mass0, *Ri0, *C0_r, *KD0, *Di_r, *V_ip, *t_tot, *zeta, *C0_left, *C0_right;
  double *out, *relerr;
};

__device__ __forceinline__ double v_pown(double x, int ndim) { return ndim == 1 ? x : (ndim == 2 ? x * x : x * x * x); }

__global__ void mbm_validate_kernel(const VArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  const double *xl = a.x_left + (size_t)m * a.capL, *cl = a.C_left + (size_t)m * a.capL;
  const double *xr = a.x_right + (size_t)m * a.capR, *cr = a.C_right + (size_t)m * a.capR;
  const int n1 = (int)a.res[2 * (size_t)m], n2 = (int)a.res[2 * (size_t)m + 1];
  double o0 = nan(""), o1 = nan(""), e = nan("");
  if (n1 >= 2 && n2 >= 2 && n1 <= a.capL && n2 <= a.capR) {
    if (a.kind == 0) {
      // calc_mass_vol (functions.jl:205-212): sequential trapezoid sums as the reference, then calc_mass_err (:184)
      const double r1 = a.rho[2 * (size_t)m], r2 = a.rho[2 * (size_t)m + 1];
      double Ml = 0.0, Mr = 0.0;
      for (int i = 0; i + 1 < n1; ++i) {
        const double dx = xsub(xmul(r1, pow_ndim(xl[i + 1], a.ndim)), xmul(r1, pow_ndim(xl[i], a.ndim)));
        const double v = xmul(xmul(0.5, dx), xadd(cl[i], cl[i + 1]));
        Ml = (i == 0) ? v : xadd(Ml, v);
      }
      for (int i = 0; i + 1 < n2; ++i) {
        const double dx = xsub(xmul(r2, pow_ndim(xr[i + 1], a.ndim)), xmul(r2, pow_ndim(xr[i], a.ndim)));
        const double v = xmul(xmul(0.5, dx), xadd(cr[i], cr[i + 1]));
        Mr = (i == 0) ? v : xadd(Mr, v);
      }
      o0 = xadd(Ml, Mr);
      o1 = a.Mass0[m];
      e = xmul(xsub(o0, o1), xinv(o1));
    } else if (a.kind == 1) {
      // rayleigh_fractionation (Benchmarks.jl:245-248) integrated as test_C1.jl:10 does
      const double R01 = a.Ri0[2 * (size_t)m], R02 = a.Ri0[2 * (size_t)m + 1], Ri1 = a.Ri[2 * (size_t)m];
      const double C0r = a.C0_r[m], KD0 = a.KD0[m];
      const double iR = 1.0 / R02;
      const double fa = R01 * iR, fb = Ri1 * iR;
      double acc = 0.0, Fp = 0.0, Cp = 0.0;
      for (int i = 0; i < 1000; ++i) {
        const double t = (double)i / 999.0;
        const double Fs = v_pown((1.0 - t) * fa + t * fb, a.ndim);
        const double Cs = pow(1.0 - Fs, KD0 - 1.0) * C0r * KD0;
        if (i > 0) { const double v = 0.5 * (Fs - Fp) * (Cp + Cs); acc = (i == 1) ? v : acc + v; }
        Fp = Fs; Cp = Cs;
      }
      o0 = acc;
      // post-processing :250-258 as shipped (the first length(indi) nodes), prepended by the value interpolated at Ri0[1]
      int k = 0;
      for (int i = 0; i < n1; ++i) if (xl[i] > R01) ++k;
      double Cbc = nan("");
      {
        double mn = xl[0], mx = xl[0];
        for (int i = 1; i < n1; ++i) { mn = fmin(mn, xl[i]); mx = fmax(mx, xl[i]); }
        if (!(R01 < mn || R01 > mx))
          for (int i = 0; i + 1 < n1; ++i)
            if (R01 >= xl[i] && R01 <= xl[i + 1]) {
              const double x0 = xl[i], x1 = xl[i + 1];
              const double N1 = 1.0 - (R01 - x0) * (1.0 / (x1 - x0)), N2 = (R01 - x0) * (1.0 / (x1 - x0));
              Cbc = N1 * cl[i] + N2 * cl[i + 1];
              break;
            }
      }
      double xp = v_pown(R01 * iR, a.ndim), yp = Cbc;
      acc = 0.0;
      for (int i = 0; i < k; ++i) {
        const double xq = v_pown(xl[i] * iR, a.ndim), yq = cl[i];
        const double v = 0.5 * (xq - xp) * (yp + yq);
        acc = (i == 0) ? v : acc + v;
        xp = xq; yp = yq;
      }
      o1 = acc;
      const double den = fmax(fabs(o0), fabs(o1));
      e = den > 0.0 ? fabs(o0 - o1) / den : 0.0;
    } else if (a.kind == 3) {
      // calc_sinus_sphere (Benchmarks.jl:82-99) / crank_time_transformation3 (:192-212) with zeta = time-integrated
      // diffusivity, isapprox of test_B1.jl / test_B2.jl (norm-wise); the series stops where its exponential has underflowed
      // to zero (every further term is an exact zero)
      const double *c0l = a.C0_left + (size_t)m * a.capL, *c0r = a.C0_right + (size_t)m * a.capR;
      const double zeta = a.zeta[m];
      const int N = n1 + n2;
      const double PI = 3.14159265358979323846;
      const double L = xr[n2 - 1], iL2 = 1.0 / (L * L);
      double sd = 0.0, sa = 0.0, sb = 0.0;
      for (int i = 0; i < N; ++i) {
        const double x = (i < n1) ? xl[i] : xr[i - n1];
        const double c0 = (i < n1) ? c0l[i] : c0r[i - n1];
        const double c = (i < n1) ? cl[i] : cr[i - n1];
        double Can;
        if (i == N - 1) Can = c0;
        else {
          double acc = c0;
          for (int nt = 1; nt <= 1000; ++nt) {
            const double ex = exp(-zeta * (double)(nt * nt) * (PI * PI) * iL2);
            if (ex == 0.0) break;
            const double sg = (nt & 1) ? -1.0 : 1.0;
            if (i == 0) acc = acc + 2.0 * (sg * ex);
            else acc = acc + 2.0 * L * (1.0 / (PI * x)) * sg * (1.0 / (double)nt) * sin((double)nt * PI * x * (1.0 / L)) * ex;
          }
          Can = acc + 1.0;
        }
        sd += (Can - c) * (Can - c); sa += Can * Can; sb += c * c;
      }
      o0 = sqrt(sd);
      o1 = sqrt(fmax(sa, sb));
      e = o1 > 0.0 ? o0 / o1 : 0.0;
    } else {
      // smith (Benchmarks.jl:284-301), isapprox of test_B5.jl (norm-wise)
      const double Ri1 = a.Ri[2 * (size_t)m], Dan = a.Di_r[m], tan_ = a.t_tot[m], Kan = a.KD0[m], R = a.V_ip[m];
      const double ksi = 0.5 * pow(Dan * tan_, -0.5);
      const double Can0 = cr[n2 - 1];
      double sd = 0.0, sa = 0.0, sb = 0.0;
      for (int i = 0; i < n2; ++i) {
        const double xan = xr[i] - Ri1;
        double Can = 1.0 + ((1.0 - Kan) * (1.0 / (2.0 * Kan))) * exp(-R * (1.0 / Dan) * xan) * erfc(ksi * (xan - R * tan_))
                     - 0.5 * erfc(ksi * (xan + R * tan_))
                     + ((1.0 - Kan) * 0.5) * (1.0 / (1.0 - Kan) - 1.0 / Kan)
                       * exp(-(1.0 - Kan) * (R * (1.0 / Dan)) * (xan + Kan * R * tan_))
                       * erfc(ksi * (xan + (2.0 * Kan - 1.0) * R * tan_));
        Can = Can * Can0;
        sd += (Can - cr[i]) * (Can - cr[i]); sa += Can * Can; sb += cr[i] * cr[i];
      }
      o0 = sqrt(sd);
      o1 = sqrt(fmax(sa, sb));
      e = o1 > 0.0 ? o0 / o1 : 0.0;
    }
  }
  a.out[2 * (size_t)m] = o0; a.out[2 * (size_t)m + 1] = o1; a.relerr[m] = e;
}

}  // namespace mbm
