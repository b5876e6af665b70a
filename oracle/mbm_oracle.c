/*
 * mbm_oracle.c — CPU restatement of the time-stepping hot path of
 * AnStroh/MovingBoundaryMinerals.jl.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libmbm_b200.so) never
 * links, calls or falls back to anything in oracle/.
 *
 * PARITY PINNING.  The reference is Julia; there is no Julia in this image, so the reference
 * itself cannot be executed here and it ships no golden vectors.  The oracle is pinned
 * against every known-answer test the reference ships (test/test_A1.jl ... test_C1.jl,
 * test_Diff_couple_no_interaction.jl) with the shipped tolerances, see
 * tests/test_oracle_kat.py.  Those tolerances are 1e-1 ... 1e-5 against analytical
 * solutions: at the 1e-10 level PARITY IS UNPINNED BY THE REFERENCE; it is defined against
 * this restatement.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Julia semantics that a C reading would miss are restated explicitly:
 * `a * inv(b)` is two roundings, integer literal powers are multiplications, `round` is
 * ties-to-even, LinRange is lerp-indexed.  Build with -ffp-contract=off: Julia never
 * contracts a*b+c unless fma/muladd is written.
 *
 * Third-party arithmetic that is NOT under /root/reference (Julia stdlib, pinned only by
 * `julia = "1.10, 1.11"`, Project.toml:35):
 *   - dense `\`  -> LAPACK getrf/getrs (partial-pivot LU): restated as unblocked LU (solve_dense);
 *   - x^n, n::Int -> Base.Math.pow_body (compensated power by squaring): restated (jl_pow_int);
 *   - cumsum      -> pairwise, leaves < 128 (Base._accumulate_pairwise!): restated (cumsum_julia);
 *   - sum         -> pairwise + SIMD re-association (not reproducible; sequential here).
 *
 * Two orders of summation are provided for the grid generator:
 *   canonical (default): the order the CUDA kernels use (tiles of 9, Hillis-Steele over 32
 *     tiles, sequential over groups) so that node counts, step counts and regrid indices are
 *     bit-identical between this oracle and the GPU; make_dx_right's sequential products use
 *     the same scan shape.
 *   MBM_FLAG_SUM_JULIA: Julia's pairwise cumsum and the literal sequential recurrence of
 *     make_dx_right (functions.jl:707-714).  Differences are ulp-level (tests quantify them).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/mbm_b200.h"

/* ------------------------------------------------------------------------------------------ */
/* Julia arithmetic helpers                                                                    */
/* ------------------------------------------------------------------------------------------ */
static inline double jl_inv(double x) { return 1.0 / x; }

/* set_inner_bc_Lasaga!, src/functions.jl:1143-1206: the numerical part (sol1, sol2, selection :1189-1193, BC_left :1194),
 * term by term and left to right as Julia evaluates the expression (x^2 = x*x, Base.literal_pow; ^(1*inv(2)) restated
 * as sqrt); cl = C_left[end-1], cr = C_right[2] (old values), d1, d2 = interface spacings, r1, r2 = rho[1], rho[2] */
static void lasaga_bc(double cl, double cr, double Dl, double Dr, double K, double d1, double d2, double r1, double r2,
                      double *bc_left, double *bc_right) {
  const double cl2 = cl * cl, cr2 = cr * cr, Dl2 = Dl * Dl, Dr2 = Dr * Dr, K2 = K * K, d12 = d1 * d1, d22 = d2 * d2,
               r12 = r1 * r1, r22 = r2 * r2;
  double rad = cl2 * Dl2 * K2 * d22 * r12;
  rad = rad - 2 * cl2 * Dl2 * K * d22 * r12;
  rad = rad + cl2 * Dl2 * d22 * r12;
  rad = rad + 2 * cl * cr * Dl * Dr * K2 * d1 * d2 * r1 * r2;
  rad = rad - 4 * cl * cr * Dl * Dr * K * d1 * d2 * r1 * r2;
  rad = rad + 2 * cl * cr * Dl * Dr * d1 * d2 * r1 * r2;
  rad = rad - 2 * cl * Dl2 * K2 * d22 * r12;
  rad = rad + 2 * cl * Dl2 * K * d22 * r12;
  rad = rad + 2 * cl * Dl * Dr * K * d1 * d2 * r1 * r2;
  rad = rad - 2 * cl * Dl * Dr * d1 * d2 * r1 * r2;
  rad = rad + cr2 * Dr2 * K2 * d12 * r22;
  rad = rad - 2 * cr2 * Dr2 * K * d12 * r22;
  rad = rad + cr2 * Dr2 * d12 * r22;
  rad = rad - 2 * cr * Dl * Dr * K2 * d1 * d2 * r1 * r2;
  rad = rad + 2 * cr * Dl * Dr * K * d1 * d2 * r1 * r2;
  rad = rad + 2 * cr * Dr2 * K * d12 * r22;
  rad = rad - 2 * cr * Dr2 * d12 * r22;
  rad = rad + Dl2 * K2 * d22 * r12;
  rad = rad + 2 * Dl * Dr * K * d1 * d2 * r1 * r2;
  rad = rad + Dr2 * d12 * r22;
  const double root = sqrt(rad);
  const double den = 2 * (Dr * d1 * r2 - Dr * K * d1 * r2);
  double n1 = Dr * d1 * r2 - root;
  n1 = n1 + cl * Dl * d2 * r1; n1 = n1 + cr * Dr * d1 * r2; n1 = n1 + Dl * K * d2 * r1;
  n1 = n1 - cl * Dl * K * d2 * r1; n1 = n1 - cr * Dr * K * d1 * r2;
  const double sol1 = n1 * jl_inv(den);
  double n2 = root + Dr * d1 * r2;
  n2 = n2 + cl * Dl * d2 * r1; n2 = n2 + cr * Dr * d1 * r2; n2 = n2 + Dl * K * d2 * r1;
  n2 = n2 - cl * Dl * K * d2 * r1; n2 = n2 - cr * Dr * K * d1 * r2;
  const double sol2 = n2 * den;                                                            /* :1187: multiplies */
  const double br = (sol2 > 1.0 || sol2 < 0.0) ? sol1 : sol2;                              /* :1189-1193 */
  *bc_right = br;
  *bc_left = cl - Dr * d1 * jl_inv(Dl * d2) * (br - cr);                                   /* :1194 */
}
static inline double jl_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : x); }

/* Base.Math.pow_body(x::Float64, n::Integer), Julia 1.10 base/math.jl (published algorithm). */
static double jl_pow_int(double x, int64_t n) {
  if (n == 0) return 1.0;
  double y = 1.0, xnlo = 0.0, ynlo = 0.0;
  if (n == 3) return x * x * x;
  if (n < 0) {
    double rx = 1.0 / x;
    if (n == -2) return rx * rx;
    if (isfinite(x)) xnlo = -fma(x, rx, -1.0) * rx;
    x = rx;
    n = -n;
  }
  while (n > 1) {
    if (n & 1) {
      double err = fma(y, xnlo, x * ynlo);
      double hi = x * y;
      double lo = fma(x, y, -hi);
      y = hi;
      ynlo = lo + err;
    }
    double err2 = x * 2.0 * xnlo;
    double hi2 = x * x;
    double lo2 = fma(x, x, -hi2);
    x = hi2;
    xnlo = lo2 + err2;
    n >>= 1;
  }
  double err = fma(y, xnlo, x * ynlo);
  if (isfinite(x) && isfinite(err)) return fma(x, y, err);
  return x * y;
}

/* x ^ ndim with ndim in {1,2,3}: calc_volume functions.jl:285,288 (runtime Int exponent ->
 * pow_body: n==1 -> loop not entered -> muladd(x, 1.0, 0.0) = x; n==2 -> x*x; n==3 -> x*x*x) */
static inline double pow_ndim(double x, int ndim) {
  if (ndim == 1) return x;
  if (ndim == 2) return x * x;
  return x * x * x;
}

/* LinRange(a,b,len)[i] (0-based i): Base.lerpi, t = j/d, (1-t)*a + t*b, d = max(len-1,1) */
static inline double linrange_at(double a, double b, int64_t len, int64_t i) {
  int64_t d = (len - 1 > 1) ? (len - 1) : 1;
  double t = (double)i / (double)d;
  return (1.0 - t) * a + t * b;
}

/* ------------------------------------------------------------------------------------------ */
/* Scans                                                                                       */
/* ------------------------------------------------------------------------------------------ */
#define CSO_TILE 9
#define CSO_GROUP 32

/* canonical inclusive scan shared with the CUDA kernels (op: 0 = +, 1 = *).
 * out[i] = O_g (op) (E_q (op) s_k), where s_k is the sequential scan inside the tile of 9 (odd: bank-conflict-free on the GPU),
 * E_q the exclusive Hillis-Steele scan of the 32 tile totals of the group (identity for q = 0)
 * and O_g the sequential scan of the group totals (identity for g = 0).                       */
static void scan_canonical(const double *v, int64_t n, double *out, int op) {
  const double ident = op ? 1.0 : 0.0;
  int64_t ntiles = (n + CSO_TILE - 1) / CSO_TILE;
  int64_t ngroups = (ntiles + CSO_GROUP - 1) / CSO_GROUP;
  double Og = ident;
  for (int64_t g = 0; g < ngroups; ++g) {
    double inc[CSO_GROUP];
    for (int q = 0; q < CSO_GROUP; ++q) {
      int64_t base = (g * CSO_GROUP + q) * CSO_TILE;
      double s = ident;
      int first = 1;
      for (int k = 0; k < CSO_TILE; ++k) {
        int64_t i = base + k;
        if (i >= n) break;
        if (first) { s = v[i]; first = 0; }
        else s = op ? (s * v[i]) : (s + v[i]);
        out[i] = s; /* local value for now */
      }
      inc[q] = s; /* identity for empty tiles */
    }
    for (int off = 1; off < CSO_GROUP; off <<= 1) {
      double nxt[CSO_GROUP];
      for (int q = 0; q < CSO_GROUP; ++q) {
        if (q >= off) nxt[q] = op ? (inc[q - off] * inc[q]) : (inc[q - off] + inc[q]);
        else nxt[q] = inc[q];
      }
      memcpy(inc, nxt, sizeof(inc));
    }
    for (int q = 0; q < CSO_GROUP; ++q) {
      int64_t base = (g * CSO_GROUP + q) * CSO_TILE;
      double Eq = (q == 0) ? ident : inc[q - 1];
      for (int k = 0; k < CSO_TILE; ++k) {
        int64_t i = base + k;
        if (i >= n) break;
        double p = (q == 0) ? out[i] : (op ? (Eq * out[i]) : (Eq + out[i]));
        out[i] = (g == 0) ? p : (op ? (Og * p) : (Og + p));
      }
    }
    Og = (g == 0) ? inc[CSO_GROUP - 1] : (op ? (Og * inc[CSO_GROUP - 1]) : (Og + inc[CSO_GROUP - 1]));
  }
}

/* Base._accumulate_pairwise!(+, c, v, s, i1, n) (Julia base/accumulate.jl, published) */
static double cumsum_julia_rec(double *c, const double *v, double s, int has_s, int64_t i1, int64_t n) {
  if (n < 128) {
    double s_ = v[i1];
    c[i1] = has_s ? (s + s_) : s_;
    for (int64_t i = i1 + 1; i < i1 + n; ++i) {
      s_ = s_ + v[i];
      c[i] = has_s ? (s + s_) : s_;
    }
    return s_;
  }
  int64_t n2 = n >> 1;
  double sl = cumsum_julia_rec(c, v, s, has_s, i1, n2);
  double snew = has_s ? (s + sl) : sl;
  double sr = cumsum_julia_rec(c, v, snew, 1, i1 + n2, n - n2);
  return sl + sr;
}
static void cumsum_julia(const double *v, int64_t n, double *out) {
  if (n <= 0) return;
  cumsum_julia_rec(out, v, 0.0, 0, 0, n);
}

/* ------------------------------------------------------------------------------------------ */
/* Member state                                                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  /* configuration */
  int ndim, inner_bc, dt_mode;
  uint32_t flags;
  int64_t capL, capR;
  /* parameters */
  double Di[2], D0[2], Ea[2], rho[2], V_ip, t_tot, CFL, MRefin, R;
  int BCout[2];
  int64_t resmin[2];
  const double *t_ar, *T_ar, *KD_ar;
  int64_t hist_len;
  const double *coeff_up, *coeff_do;
  double C_leftB, C_rightB;
  int nd;
  const double *dens_X, *dens_T, *dens_left, *dens_right;
  int law;
  double law_par[8];
  /* state (arrays have capacity cap+1 so that the transient inserted node always fits) */
  double *xl, *cl, *xr, *cr;
  int64_t n1, n2;
  double Ri[2];
  double t, dt;
  int64_t it;
  double dx1, dx2, Mass0, newton_R;
  double D_l, D_r, KD, T;
  int Fl_regrid;
  int status, warn;
  uint64_t hash;
  int64_t node_steps;
} member_t;

/* order-independent checksum term of one regrid decision (shared with the CUDA kernels) */
static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t idx_term(int64_t it, int phase, int64_t j, int64_t k, int branch) {
  uint64_t key = ((uint64_t)it << 40) ^ ((uint64_t)phase << 39) ^ ((uint64_t)j << 18) ^
                 ((uint64_t)k << 2) ^ (uint64_t)branch;
  return mix64(key);
}

/* ------------------------------------------------------------------------------------------ */
/* linear_interpolation_1D, functions.jl:565-585                                               */
/* ------------------------------------------------------------------------------------------ */
static int lin_interp_1d(const double *x, const double *y, int64_t n, double xi, double *out) {
  double mn = x[0], mx = x[0];
  for (int64_t i = 1; i < n; ++i) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
  if (xi < mn || xi > mx) return MBM_ERR_INTERP_BOUNDS;       /* :569-571 */
  for (int64_t i = 0; i + 1 < n; ++i) {
    if (xi >= x[i] && xi <= x[i + 1]) {                       /* :575 */
      double x0 = x[i], y0 = y[i], x1 = x[i + 1], y1 = y[i + 1];
      double N1 = 1.0 - (xi - x0) * jl_inv(x1 - x0);          /* :579 */
      double N2 = (xi - x0) * jl_inv(x1 - x0);                /* :580 */
      *out = N1 * y0 + N2 * y1;                               /* :581 */
      return 0;
    }
  }
  *out = NAN; /* Julia would return `nothing`; unreachable for sorted x */
  return MBM_ERR_INTERP_BOUNDS;
}

static void minmax(const double *a, int64_t n, double *mn, double *mx) {
  *mn = a[0]; *mx = a[0];
  for (int64_t i = 1; i < n; ++i) { if (a[i] < *mn) *mn = a[i]; if (a[i] > *mx) *mx = a[i]; }
}

/* update_t_dependent_param!, functions.jl:1350-1378 */
static int update_t_dependent_param(member_t *m) {
  double KD, T;
  int rc = lin_interp_1d(m->t_ar, m->KD_ar, m->hist_len, m->t, &KD);  /* :1352 */
  if (rc) return rc;
  rc = lin_interp_1d(m->t_ar, m->T_ar, m->hist_len, m->t, &T);        /* :1353 */
  if (rc) return rc;
  const double tol = 1e-12;
  if (m->t == m->t_tot) {                                             /* :1356-1359 */
    T = m->T_ar[m->hist_len - 1];
    KD = m->KD_ar[m->hist_len - 1];
  } else {
    double mn, mx;
    minmax(m->KD_ar, m->hist_len, &mn, &mx);
    if (KD < mn - tol || KD > mx + tol) return MBM_ERR_KD_RANGE;      /* :1360-1362 */
    minmax(m->T_ar, m->hist_len, &mn, &mx);
    if (T < mn - tol || T > mx + tol) return MBM_ERR_T_RANGE;         /* :1363-1365 */
  }
  if (m->Di[0] == -1.0 && m->Di[1] == -1.0) {                         /* :1368-1370 */
    m->D_l = m->D0[0] * exp(-m->Ea[0] * jl_inv(m->R * T));
    m->D_r = m->D0[1] * exp(-m->Ea[1] * jl_inv(m->R * T));
  } else {                                                            /* :1371-1373 */
    m->D_l = m->Di[0];
    m->D_r = m->Di[1];
  }
  m->KD = KD;
  m->T = T;
  return 0;
}

/* find_dt, functions.jl:465-490 */
static double find_dt(double dx1, double dx2, double V, double D_l, double D_r, double CFL) {
  const double dt_drop = 0.99;
  double mn = (dx1 < dx2) ? dx1 : dx2;
  double mxD = (D_l > D_r) ? D_l : D_r;
  double dtV = mn * jl_inv(fabs(V));                /* :468 */
  double dtD = (mn * mn) * jl_inv(mxD) * 50.0;      /* :469 */
  dtV = dtV * dt_drop;                              /* :470 */
  if (V == 0.0) return dtD * CFL;                   /* :472-473 */
  if (dtV < dtD) return dtV * CFL;                  /* :475-476 */
  return dtD * CFL;                                 /* :478 */
}

/* update_time!, functions.jl:1442-1450 */
static void update_time(double *t, double *dt, int64_t *it, double t_tot) {
  *t = *t + *dt;
  *it = *it + 1;
  if (*t > t_tot) {
    *dt = *dt - (*t - t_tot);
    *t = t_tot;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Grid generation                                                                             */
/* ------------------------------------------------------------------------------------------ */
/* newton_solver, functions.jl:732-757; x0 = 1.1 in the reference (:734) */
static int newton_solver(double S, double d, int64_t n, double tol, int max_iter, double x0,
                         double *root, int *warn) {
  double x = x0, Res = 1e23;
  for (int i = 1; i <= max_iter; ++i) {
    double xn = jl_pow_int(x, n);
    double xn1 = jl_pow_int(x, n - 1);
    double Fx = (1.0 - xn) * jl_inv(1.0 - x) - (S * jl_inv(d));                          /* :737 */
    double omx = 1.0 - x;
    double dF = fma(-(double)n * xn1, omx, (1.0 - xn)) * jl_inv(omx * omx);             /* :738 */
    double _dFx = jl_inv(dF);
    Res = fabs(Fx);                                                                      /* :739 */
    if (Res < tol) { *root = x; return 0; }                                              /* :740-744 */
    x = x - 1.0 * Fx * _dFx;                                                             /* :746 */
  }
  *root = x;
  if (Res > 1e-3) return MBM_ERR_NEWTON;                                                 /* :751-752 */
  /* `NaN > 1e-3` is false in Julia: a NaN residual falls through to the warning branch */
  *warn |= MBM_WARN_NEWTON_RES;                                                          /* :754 */
  return 0;
}

/* define_new_grid, functions.jl:395-441 (+ make_dx_right :707-714).  X_left has nr[0] nodes,
 * X_right nr[1] nodes.  newton_R: in = warm-start value (<= 0: cold start), out = root.      */
static int define_new_grid(const double Ri[2], const int64_t nr[2], double Rfact, uint32_t flags,
                           double *X_left, double *X_right, double *newton_R, int *warn,
                           double *work /* >= 2*max(nr) */) {
  const int julia = (flags & MBM_FLAG_SUM_JULIA) != 0;
  int64_t n1 = nr[0], n2 = nr[1];
  if (Rfact == 1.0) {                                                                    /* :396-400 */
    for (int64_t i = 0; i < n1; ++i) X_left[i] = linrange_at(0.0, Ri[0], n1, i);
    for (int64_t i = 0; i < n2; ++i) X_right[i] = linrange_at(Ri[0], Ri[1], n2, i);
    return 0;
  }
  if (!(Rfact > 1.0)) return MBM_ERR_MREFIN;                                             /* :412-426 unsupported */
  /* :402 dx_left = LinRange(1.0, 1.0*inv(Rfact), nr[1]-1) */
  int64_t nl = n1 - 1;
  double *dxl = work, *cs = work + (nl > n2 ? nl : n2);
  double b = 1.0 * jl_inv(Rfact);
  for (int64_t i = 0; i < nl; ++i) dxl[i] = linrange_at(1.0, b, nl, i);
  /* :403 x_left = [0; cumsum(dx_left)] * Ri[1] * inv(sum(dx_left)) */
  double total;
  if (julia) {
    cumsum_julia(dxl, nl, cs);
    total = 0.0;
    for (int64_t i = 0; i < nl; ++i) total += dxl[i]; /* Julia's sum is pairwise+SIMD: order not reproducible */
  } else {
    scan_canonical(dxl, nl, cs, 0);
    total = cs[nl - 1];
  }
  double isum = jl_inv(total);
  X_left[0] = 0.0 * Ri[0] * isum;
  for (int64_t i = 0; i < nl; ++i) X_left[i + 1] = cs[i] * Ri[0] * isum;
  /* :406-408 */
  double S = Ri[1] - Ri[0];
  double d = X_left[n1 - 1] - X_left[n1 - 2];   /* :404,407 dx_left = diff(x_left); d = dx_left[end] */
  int64_t n = n2 - 1;
  double x0 = 1.1;                              /* :734 */
  if ((flags & MBM_FLAG_NEWTON_WARM) && newton_R && *newton_R > 0.0) x0 = *newton_R;
  double Rt;
  int rc = newton_solver(S, d, n, 1e-8, 200, x0, &Rt, warn);
  if (rc && x0 != 1.1) rc = newton_solver(S, d, n, 1e-8, 200, 1.1, &Rt, warn); /* warm start failed: retry cold */
  if (rc) return rc;
  if (newton_R) *newton_R = Rt;
  /* :409 make_dx_right(R, d, n): dx[1] = d; dx[i] = dx[i-1]*R */
  double *dxr = work;
  if (julia) {
    dxr[0] = d;
    for (int64_t i = 1; i < n; ++i) dxr[i] = dxr[i - 1] * Rt;
    cumsum_julia(dxr, n, cs);
  } else {
    double *u = cs;
    u[0] = d;
    for (int64_t i = 1; i < n; ++i) u[i] = Rt;
    scan_canonical(u, n, dxr, 1);
    scan_canonical(dxr, n, cs, 0);
  }
  /* :410 x_right = [Ri[1]; Ri[1] .+ cumsum(dx_right)] */
  X_right[0] = Ri[0];
  for (int64_t i = 0; i < n; ++i) X_right[i + 1] = Ri[0] + cs[i];
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pchip, functions.jl:775-839.  Returns 0 or MBM_ERR_PCHIP_INDEX.  d: work array (n).         */
/* ------------------------------------------------------------------------------------------ */
static void pchip_slopes(const double *x, const double *y, int64_t n, double *d) {
  const double msl = 3.0;
#define H(i) (x[(i) + 1] - x[(i)])
#define MS(i) ((y[(i) + 1] - y[(i)]) * jl_inv(H(i)))
  for (int64_t i = 0; i < n; ++i) {
    if (i == 0) {                                                                         /* :787-791 */
      double h1 = H(0), h2 = H(1), m1 = MS(0), m2 = MS(1);
      d[i] = fma(fma(2.0, h1, h2), m1, -(h1 * m2)) * jl_inv(h1 + h2);
      if (m1 * m2 < 0.0 || m1 == 0.0 || m2 == 0.0) d[i] = 0.0;
    } else if (i == n - 1) {                                                              /* :792-796 */
      double hn1 = H(n - 2), hn2 = H(n - 3), mn1 = MS(n - 2), mn2 = MS(n - 3);
      d[i] = fma(fma(2.0, hn1, hn2), mn1, -(hn1 * mn2)) * jl_inv(hn1 + hn2);
      if (mn1 * mn2 < 0.0 || mn1 == 0.0 || mn2 == 0.0) d[i] = 0.0;
    } else {                                                                              /* :797-805 */
      double wim1 = 2.0 * jl_inv(x[i] - x[i - 1]);
      double wi = 2.0 * jl_inv(x[i + 1] - x[i]);
      double mm = MS(i - 1), mp = MS(i);
      double di = (wim1 * mm + wi * mp) * jl_inv(wim1 + wi);
      if (mm * mp < 0.0 || mm == 0.0 || mp == 0.0) di = 0.0;
      double lim = msl * fmin(fabs(mm), fabs(mp));
      di = fmin(fabs(di), lim) * jl_sign(di);
      d[i] = di;
    }
  }
#undef H
#undef MS
}

static int pchip_eval(const double *x, const double *y, const double *d, int64_t n, const double *X,
                      int64_t nX, double *P, member_t *mb, int phase) {
  for (int64_t j = 0; j < nX; ++j) {
    double Xj = X[j];
    /* c = number of old nodes strictly below Xj (x is strictly increasing) */
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (x[mid] < Xj) lo = mid + 1; else hi = mid; }
    int64_t c = lo, k;
    int branch;
    if (c < n && x[c] == Xj) { k = c; branch = 0; P[j] = y[k]; }                          /* :810-812 */
    else if (c == n) { k = n - 1; branch = 1; P[j] = y[k]; }                              /* :813-815 */
    else if (c == 0) { k = 0; branch = 2; P[j] = y[k]; }                                  /* :816-818 */
    else {
      k = c - 1; branch = 3;                                                              /* :820 */
      if (k > n - 2) return MBM_ERR_PCHIP_INDEX;                                          /* :821-823 */
      double x1 = x[k], x2 = x[k + 1], y1 = y[k], y2 = y[k + 1], d1 = d[k], d2 = d[k + 1];
      double s = Xj - x1, h = x2 - x1;                                                    /* :830-831 */
      double s2 = s * s, s3 = s * s * s, h3 = h * h * h, h2 = h * h;
      double smh = s - h;
      double t1 = (3.0 * h * s2 - 2.0 * s3) * jl_inv(h3) * y2;                            /* :832 */
      double t2 = (h3 - 3.0 * h * s2 + 2.0 * s3) * jl_inv(h3) * y1;                       /* :833 */
      double t3 = (s2 * smh) * jl_inv(h2) * d2;                                           /* :834 */
      double t4 = (s * (smh * smh)) * jl_inv(h2) * d1;                                    /* :835 */
      P[j] = t1 + t2 + t3 + t4;
    }
    mb->hash += idx_term(mb->it, phase, j, k, branch);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Volumes / mass                                                                              */
/* ------------------------------------------------------------------------------------------ */
/* trapezoidal_integration, functions.jl:1310-1323 (cumsum(Int, dims=1) on an n x 1 Matrix is sequential) */
static double trapz(const double *x, double scale, int ndim_pow, const double *fx, int64_t n) {
  /* integrates fx over (scale * x^ndim_pow); ndim_pow = 0 means x itself (scale ignored) */
  double acc = 0.0;
  int first = 1;
  for (int64_t i = 0; i + 1 < n; ++i) {
    double a = ndim_pow ? scale * pow_ndim(x[i], ndim_pow) : x[i];
    double b = ndim_pow ? scale * pow_ndim(x[i + 1], ndim_pow) : x[i + 1];
    double dx = b - a;
    double v = 0.5 * dx * (fx[i] + fx[i + 1]);
    if (first) { acc = v; first = 0; } else acc = acc + v;
  }
  return acc;
}

/* calc_mass_vol, functions.jl:205-212: rho[k]*V_k integrated against C_k */
static double calc_mass_vol(const double *xl, const double *xr, const double *cl, const double *cr,
                            int64_t n1, int64_t n2, int ndim, const double rho[2]) {
  double Ml = trapz(xl, rho[0], ndim, cl, n1);
  double Mr = trapz(xr, rho[1], ndim, cr, n2);
  return Ml + Mr;
}

/* calc_volume, functions.jl:273-298: dVolC (length n1+n2) */
static void calc_dvolc(const double *xl, const double *xr, int64_t n1, int64_t n2, int ndim, double *w) {
  int64_t N = n1 + n2;
  /* V = [V1; V2]; dV[i] = V[i+1]-V[i]; dVC = [dV;0]*0.5 + [0;dV]*0.5 */
  for (int64_t i = 0; i < N; ++i) {
    double Vi = pow_ndim(i < n1 ? xl[i] : xr[i - n1], ndim);
    double dVf = 0.0, dVb = 0.0;
    if (i + 1 < N) { double Vn = pow_ndim((i + 1) < n1 ? xl[i + 1] : xr[i + 1 - n1], ndim); dVf = Vn - Vi; }
    if (i > 0) { double Vp = pow_ndim((i - 1) < n1 ? xl[i - 1] : xr[i - 1 - n1], ndim); dVb = Vi - Vp; }
    w[i] = dVf * 0.5 + dVb * 0.5;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* FEM assembly: fill_matrix!, functions.jl:510-549 — tridiagonal storage.                     */
/* sub[i] = L[i,i-1], dia[i] = L[i,i], sup[i] = L[i,i+1], rhs[i] = R[i].                        */
/* ------------------------------------------------------------------------------------------ */
static void element_matrices(double x_1, double x_2, double D, int ndim, double M[4], double K[4]) {
  if (ndim == 3) {                                                                        /* :525-529 */
    double x1s = x_1 * x_1, x2s = x_2 * x_2, df = x_1 - x_2;
    double i30 = jl_inv(30.0), i60 = jl_inv(60.0);
    M[0] = -(df * (6.0 * x1s + 3.0 * x_1 * x_2 + 1.0 * x2s)) * i30;
    M[1] = -(df * (3.0 * x1s + 4.0 * x_1 * x_2 + 3.0 * x2s)) * i60;
    M[2] = M[1];
    M[3] = -(df * (1.0 * x1s + 3.0 * x_1 * x_2 + 6.0 * x2s)) * i30;
    double kk = (D * (x1s + x_1 * x_2 + x2s)) * jl_inv(3.0 * df);
    K[0] = -kk; K[1] = kk; K[2] = kk; K[3] = -kk;   /* -(a)*inv(b) == -(a*inv(b)) exactly */
  } else if (ndim == 2) {                                                                 /* :530-534 */
    double df = x_1 - x_2, i12 = jl_inv(12.0);
    M[0] = -(df * (3.0 * x_1 + x_2)) * i12;
    M[1] = (x_2 * x_2) * i12 - (x_1 * x_1) * i12;
    M[2] = M[1];
    M[3] = -(df * (x_1 + 3.0 * x_2)) * i12;
    double kk = (D * (x_1 + x_2)) * jl_inv(2.0 * df);
    K[0] = -kk; K[1] = kk; K[2] = kk; K[3] = -kk;
  } else {                                                                                /* :535-539 */
    double i3 = jl_inv(3.0), i6 = jl_inv(6.0);
    M[0] = x_2 * i3 - x_1 * i3;
    M[1] = x_2 * i6 - x_1 * i6;
    M[2] = M[1];
    M[3] = M[0];
    double kk = D * jl_inv(x_1 - x_2);
    K[0] = -kk; K[1] = kk; K[2] = kk; K[3] = -kk;
  }
}

static void fill_matrix(const double *C, const double *x, double D, double dt, int ndim, int64_t n,
                        double *sub, double *dia, double *sup, double *rhs) {
  for (int64_t i = 0; i < n; ++i) { sub[i] = 0.0; dia[i] = 0.0; sup[i] = 0.0; rhs[i] = 0.0; }
  double _dt = jl_inv(dt);                                                                /* :519 */
  for (int64_t iel = 0; iel + 1 < n; ++iel) {
    double M[4], K[4];
    element_matrices(x[iel], x[iel + 1], D, ndim, M, K);
    double A0 = M[0] * _dt, A1 = M[1] * _dt, A2 = M[2] * _dt, A3 = M[3] * _dt;            /* Mloc .* _dt */
    double L0 = A0 + K[0], L1 = A1 + K[1], L2 = A2 + K[2], L3 = A3 + K[3];                /* :542 */
    double R0 = A0 * C[iel] + A1 * C[iel + 1];                                            /* :543 */
    double R1 = A2 * C[iel] + A3 * C[iel + 1];
    dia[iel] += L0; sup[iel] += L1; sub[iel + 1] += L2; dia[iel + 1] += L3;               /* :545 */
    rhs[iel] += R0; rhs[iel + 1] += R1;                                                   /* :546 */
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Linear solves                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* `L_g \ R_g` on a dense matrix: LU with partial pivoting (LAPACK getrf/getrs family),        */
/* functions.jl:1292.  A is N x N row-major, destroyed.                                        */
static int solve_dense(double *A, double *b, int64_t N) {
  for (int64_t k = 0; k < N; ++k) {
    int64_t p = k;
    double best = fabs(A[k * N + k]);
    for (int64_t i = k + 1; i < N; ++i) { double v = fabs(A[i * N + k]); if (v > best) { best = v; p = i; } }
    if (best == 0.0) return -1;
    if (p != k) {
      for (int64_t j = 0; j < N; ++j) { double tmp = A[k * N + j]; A[k * N + j] = A[p * N + j]; A[p * N + j] = tmp; }
      double tb = b[k]; b[k] = b[p]; b[p] = tb;
    }
    double ipiv = 1.0 / A[k * N + k];
    for (int64_t i = k + 1; i < N; ++i) {
      double l = A[i * N + k] * ipiv;
      if (l != 0.0) {
        A[i * N + k] = l;
        double *ri = A + i * N; const double *rk = A + k * N;
        for (int64_t j = k + 1; j < N; ++j) ri[j] -= l * rk[j];
        b[i] -= l * b[k];
      }
    }
  }
  for (int64_t i = N - 1; i >= 0; --i) {
    double s = b[i];
    const double *ri = A + i * N;
    for (int64_t j = i + 1; j < N; ++j) s -= ri[j] * b[j];
    b[i] = s / ri[i];
  }
  return 0;
}

/* Thomas algorithm for one tridiagonal system with two right-hand sides (in place). */
static void thomas2(const double *sub, const double *dia, const double *sup, double *r1, double *r2,
                    int64_t n, double *cp) {
  double bet = dia[0];
  cp[0] = sup[0] / bet;
  r1[0] = r1[0] / bet;
  r2[0] = r2[0] / bet;
  for (int64_t i = 1; i < n; ++i) {
    bet = dia[i] - sub[i] * cp[i - 1];
    cp[i] = sup[i] / bet;
    r1[i] = (r1[i] - sub[i] * r1[i - 1]) / bet;
    r2[i] = (r2[i] - sub[i] * r2[i - 1]) / bet;
  }
  for (int64_t i = n - 2; i >= 0; --i) {
    r1[i] = r1[i] - cp[i] * r1[i + 1];
    r2[i] = r2[i] - cp[i] * r2[i + 1];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* One couple step: loop body of main_codes/Diff_couple_1D_general_{MB,Flux}.jl:87-132,         */
/* Chemical_Stefan_problem_XT.jl:100-149, examples/Diff_couple_no_interaction.jl:96-124        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double *buf; size_t n; } scratch_t;
static double *scratch_get(scratch_t *s, size_t n) {
  if (s->n < n) { free(s->buf); s->buf = (double *)malloc(n * sizeof(double)); s->n = n; }
  return s->buf;
}

/* interp2, functionsStefan.jl:93-112 (scalar query) */
static int interp2(const double *X1d, const double *Y1d, const double *Z, int nd, double xi, double yi, double *zi) {
  if (!(xi >= X1d[0] && xi <= X1d[nd - 1])) return MBM_ERR_DENSITY_BOUNDS;                /* :97 */
  if (!(yi >= Y1d[0] && yi <= Y1d[nd - 1])) return MBM_ERR_DENSITY_BOUNDS;                /* :98 */
  int ix = -1, iy = -1;
  for (int i = 0; i < nd; ++i) { if (X1d[i] < xi) ix = i; if (Y1d[i] < yi) iy = i; }      /* findlast :99-100 */
  if (ix < 0 || iy < 0 || ix + 1 >= nd || iy + 1 >= nd) return MBM_ERR_DENSITY_BOUNDS;   /* `nothing` index -> error in Julia */
  double dx = X1d[ix + 1] - X1d[ix], dy = Y1d[iy + 1] - Y1d[iy];
  double idd = jl_inv(dx * dy);
  double w11 = (X1d[ix + 1] - xi) * (Y1d[iy + 1] - yi) * idd;
  double w12 = (X1d[ix + 1] - xi) * (yi - Y1d[iy]) * idd;
  double w21 = (xi - X1d[ix]) * (Y1d[iy + 1] - yi) * idd;
  double w22 = (xi - X1d[ix]) * (yi - Y1d[iy]) * idd;
#define ZZ(i, j) Z[(i) + (size_t)(j) * nd]
  *zi = w11 * ZZ(ix, iy) + w12 * ZZ(ix, iy + 1) + w21 * ZZ(ix + 1, iy) + w22 * ZZ(ix + 1, iy + 1);
#undef ZZ
  return 0;
}

/* Capture hook of the conditioning report (tools/conditioning_report.py): when armed, the next step
 * solved by either mode leaves its dense system (as the reference would hand it to `\`, functions.jl:1292), its
 * right-hand side and the solution it computed in the caller's buffers.  Single-threaded use only (one member).     */
static struct { double *A, *rhs, *sol; int64_t cap, *N; int armed; } g_capture;
void mbm_oracle_capture_next_system(double *A, double *rhs, double *sol, int64_t cap, int64_t *N) {
  g_capture.A = A; g_capture.rhs = rhs; g_capture.sol = sol; g_capture.cap = cap; g_capture.N = N;
  g_capture.armed = (A && rhs && sol && N) ? 1 : 0;
}

static int couple_step(member_t *m, scratch_t *sc) {
  int rc;
  double V_ip = m->V_ip;
  double rho[2] = { m->rho[0], m->rho[1] };
  /* ---- dt, time, time-dependent parameters ------------------------------------------------ */
  if (m->inner_bc == MBM_BC_STEFAN) {
    update_time(&m->t, &m->dt, &m->it, m->t_tot);                                         /* CSP :102 */
    if (m->t <= m->t_tot) {                                                               /* :103-105 */
      rc = lin_interp_1d(m->t_ar, m->T_ar, m->hist_len, m->t, &m->T);
      if (rc) return rc;
    }
    double T = m->T;
    /* composition, functionsStefan.jl:67-71 */
    double CLe = m->coeff_do[0] + m->coeff_do[1] * T + m->coeff_do[2] * (T * T);
    double CR1 = m->coeff_up[0] + m->coeff_up[1] * T + m->coeff_up[2] * (T * T);
    m->cl[m->n1 - 1] = CLe;                                                               /* :107 */
    m->cr[0] = CR1;
    double dC = m->cr[0] - m->cl[m->n1 - 1];                                              /* :108 */
    if (m->nd > 0) {                                                                      /* calculate_density :23-31 */
      double rL, rR;
      rc = interp2(m->dens_X, m->dens_T, m->dens_left, m->nd, m->C_leftB, T, &rL); if (rc) return rc;
      rc = interp2(m->dens_X, m->dens_T, m->dens_right, m->nd, m->C_rightB, T, &rR); if (rc) return rc;
      rho[0] = rL * jl_inv(rL);
      rho[1] = rR * jl_inv(rL);
    }
    m->rho[0] = rho[0]; m->rho[1] = rho[1];
    if (m->law == MBM_LAW_OLIVINE_MELT_D1) {                                              /* examples/D1.jl:126-132 */
      /* mean(C_left) after C_left[end] was set: canonical sum / length (Julia's pairwise `sum` is not reproducible
       * between builds; the canonical order is shared with the CUDA kernels)                                       */
      double *cs = scratch_get(sc, (size_t)m->n1 + 8);
      scan_canonical(m->cl, m->n1, cs, 0);
      double meanC = cs[m->n1 - 1] / (double)m->n1;
      double l001 = (m->law_par[0] - m->law_par[1] / ((2.303 * m->R) * T)) + 3.0 * (meanC - 0.14);       /* :126 */
      double loth = l001 - 0.7781512503836436 /* log10(6.0) */;                                                                 /* :127 */
      double D001 = pow(10.0, l001), Doth = pow(10.0, loth);                                             /* :128-130 */
      m->D_l = (D001 * m->law_par[2] + Doth * m->law_par[3]) + Doth * m->law_par[4];                     /* :131 */
      m->D_r = exp(-7.92 - 26222.0 / T);                                                                 /* :132 */
    }
    double JL = -m->D_l * rho[0] * (m->cl[m->n1 - 1] - m->cl[m->n1 - 2]) * jl_inv(m->dx1); /* :111 */
    double JR = -m->D_r * rho[1] * (m->cr[1] - m->cr[0]) * jl_inv(m->dx2);                 /* :112 */
    V_ip = (JR - JL) * jl_inv(dC);                                                         /* :113 */
    m->V_ip = V_ip;
    double R_left = CLe * jl_inv(1.0 - CLe), R_right = CR1 * jl_inv(1.0 - CR1);
    m->KD = R_left * jl_inv(R_right);
  } else {
    if (m->dt_mode == MBM_DT_ADVECTIVE) {
      double mn = (m->dx1 < m->dx2) ? m->dx1 : m->dx2;
      m->dt = mn / V_ip * m->CFL;                                                         /* C1.jl:93 */
    } else {
      m->dt = find_dt(m->dx1, m->dx2, V_ip, m->D_l, m->D_r, m->CFL);                      /* :89 */
    }
    update_time(&m->t, &m->dt, &m->it, m->t_tot);                                         /* :91 */
    rc = update_t_dependent_param(m);                                                     /* :93 */
    if (rc) return rc;
  }
  double dt = m->dt;
  /* ---- advect_interface_regrid!, functions.jl:29-64 -------------------------------------- */
  int Fl_regrid = 0;
  if (m->inner_bc != MBM_BC_DIRICHLET_HOLD) {
    double Rio = m->Ri[0];
    double Rnew = Rio + V_ip * dt;                                                        /* :31 */
    if (V_ip > 0.0) {
      if (m->n1 + 1 > m->capL) return MBM_ERR_CAPACITY;
      m->Ri[0] = Rnew;
      m->xl[m->n1] = Rnew; m->cl[m->n1] = m->cl[m->n1 - 1];                               /* :33-35 */
      if (Rnew > m->xr[1]) return MBM_ERR_IFACE_FAST_POS;                                 /* :37-40 */
      m->n1 += 1;
      m->xr[0] = Rnew;                                                                    /* :41 */
      Fl_regrid = 1;                                                                      /* :43-44 */
    } else if (V_ip < 0.0) {
      if (m->n2 + 1 > m->capR) return MBM_ERR_CAPACITY;
      m->Ri[0] = Rnew;
      memmove(m->xr + 1, m->xr, (size_t)m->n2 * sizeof(double));                          /* :46-48 */
      memmove(m->cr + 1, m->cr, (size_t)m->n2 * sizeof(double));
      m->n2 += 1;
      m->xr[0] = Rnew;
      if (Rnew < m->xl[m->n1 - 2]) return MBM_ERR_IFACE_FAST_NEG;                         /* :50-53 */
      m->xl[m->n1 - 1] = Rnew;                                                            /* :54 */
      Fl_regrid = 1;                                                                      /* :56-57 */
    } else {
      m->Ri[0] = Rnew;                                                                    /* V == 0 (or NaN): :58-61 */
    }
  }
  m->Fl_regrid = Fl_regrid;
  const int64_t n1 = m->n1, n2 = m->n2, N = n1 + n2;
  m->node_steps += N;
  /* ---- construct_matrix_fem, functions.jl:323-343 ---------------------------------------- */
  double *w = scratch_get(sc, (size_t)(12 * (N + 4)));
  double *sub = w, *dia = w + (N + 4), *sup = w + 2 * (N + 4), *rhs = w + 3 * (N + 4);
  double *rhs2 = w + 4 * (N + 4), *cp = w + 5 * (N + 4), *dvol = w + 6 * (N + 4);
  double *tmpA = w + 7 * (N + 4), *tmpB = w + 8 * (N + 4), *wk = w + 9 * (N + 4); /* wk: 3*(N+4) */
  fill_matrix(m->cl, m->xl, m->D_l, dt, m->ndim, n1, sub, dia, sup, rhs);                 /* :332 */
  fill_matrix(m->cr, m->xr, m->D_r, dt, m->ndim, n2, sub + n1, dia + n1, sup + n1, rhs + n1); /* :335 */
  const double Co_l1 = m->cl[0], Co_rend = m->cr[n2 - 1];
  const int64_t p = n1 - 1, q = n1;   /* 0-based interface rows: reference rows nr[1], nr[1]+1 */
  if (m->inner_bc == MBM_BC_MB) calc_dvolc(m->xl, m->xr, n1, n2, m->ndim, dvol);          /* MB :97 */
  double KD = m->KD;
  double dxl_if = m->xl[n1 - 1] - m->xl[n1 - 2], dxr_if = m->xr[1] - m->xr[0];

  const int capturing = g_capture.armed && N <= g_capture.cap;
  if ((m->flags & MBM_FLAG_SOLVER_FAITHFUL) || capturing) {
    /* dense N x N, rows overwritten exactly as the reference does, then LU */
    double *A = (double *)calloc((size_t)N * (size_t)N, sizeof(double));
    if (!A) return MBM_ERR_CAPACITY;
    double *rhs_struct = NULL;
    if (!(m->flags & MBM_FLAG_SOLVER_FAITHFUL)) {            /* capture beside the structured solve: work on a copy */
      rhs_struct = rhs;
      rhs = (double *)malloc((size_t)N * sizeof(double));
      if (!rhs) { free(A); return MBM_ERR_CAPACITY; }
      memcpy(rhs, rhs_struct, (size_t)N * sizeof(double));
    }
    for (int64_t i = 0; i < N; ++i) {
      A[i * N + i] = dia[i];
      if (i > 0 && i != n1) A[i * N + i - 1] = sub[i];
      if (i + 1 < N && i != n1 - 1) A[i * N + i + 1] = sup[i];
    }
    for (int64_t j = 0; j < N; ++j) { A[p * N + j] = 0.0; A[q * N + j] = 0.0; }
    const double ScF = 1.0;
    if (m->inner_bc == MBM_BC_FLUX) {                                                     /* :1054-1072 */
      A[p * N + p] = 1.0 * ScF; A[p * N + q] = -KD * ScF; rhs[p] = 0.0;
      A[q * N + q] = (-V_ip * rho[1] + rho[1] * m->D_r * jl_inv(dxr_if)) * ScF;
      A[q * N + q + 1] = (-rho[1] * m->D_r * jl_inv(dxr_if)) * ScF;
      A[q * N + p] = (+V_ip * rho[0] + rho[0] * m->D_l * jl_inv(dxl_if)) * ScF;
      A[q * N + p - 1] = (-rho[0] * m->D_l * jl_inv(dxl_if)) * ScF;
      rhs[q] = 0.0;
    } else if (m->inner_bc == MBM_BC_MB) {                                                /* :1094-1109 */
      for (int64_t j = 0; j < N; ++j) A[p * N + j] = dvol[j];
      rhs[p] = m->Mass0;
      A[q * N + p] = -1.0 * ScF; A[q * N + q] = +KD * ScF; rhs[q] = 0.0;
    } else if (m->inner_bc == MBM_BC_LASAGA) {                                            /* :1196-1203 */
      double bl, br;
      lasaga_bc(m->cl[n1 - 2], m->cr[1], m->D_l, m->D_r, KD, dxl_if, dxr_if, rho[0], rho[1], &bl, &br);
      A[p * N + p] = 1.0 * ScF; rhs[p] = bl * ScF;
      A[q * N + q] = 1.0 * ScF; rhs[q] = br * ScF;
    } else {                                                                              /* Stefan :131-144, DCNI :108-118 */
      A[p * N + p] = 1.0 * ScF; rhs[p] = m->cl[n1 - 1] * ScF;
      A[q * N + q] = 1.0 * ScF; rhs[q] = m->cr[0] * ScF;
    }
    if (m->BCout[0] == 1) { for (int64_t j = 0; j < N; ++j) A[j] = 0.0; A[0] = 1.0 * ScF; rhs[0] = Co_l1 * ScF; }   /* :1232-1236 */
    if (m->BCout[1] == 1) { for (int64_t j = 0; j < N; ++j) A[(N - 1) * N + j] = 0.0; A[(N - 1) * N + N - 1] = 1.0 * ScF; rhs[N - 1] = Co_rend * ScF; } /* :1237-1241 */
    if (capturing) {
      memcpy(g_capture.A, A, (size_t)N * (size_t)N * sizeof(double));
      memcpy(g_capture.rhs, rhs, (size_t)N * sizeof(double));
      *g_capture.N = N;
    }
    if (rhs_struct) { free(A); free(rhs); rhs = rhs_struct; }
    else {
      int lrc = solve_dense(A, rhs, N);                                                   /* :1292 */
      free(A);
      if (lrc) return MBM_ERR_NONFINITE;
      memcpy(m->cl, rhs, (size_t)n1 * sizeof(double));                                    /* :1293-1294 */
      memcpy(m->cr, rhs + n1, (size_t)n2 * sizeof(double));
    }
  }
  if (!(m->flags & MBM_FLAG_SOLVER_FAITHFUL)) {
    /* structured O(N) solve of the SAME linear system: both interface rows become identity rows
     * (which decouples the blocks), two right-hand sides U (FEM rhs, interface values 0) and
     * V (zero rhs, interface values 1); the interface condition is a 2x2 for (a,b) =
     * (C[p], C[q]); C = U + V*a on the left, U + V*b on the right. */
    if (m->BCout[0] == 1) { sub[0] = 0.0; dia[0] = 1.0; sup[0] = 0.0; rhs[0] = Co_l1; }
    if (m->BCout[1] == 1) { sub[N - 1] = 0.0; dia[N - 1] = 1.0; sup[N - 1] = 0.0; rhs[N - 1] = Co_rend; }
    sub[p] = 0.0; dia[p] = 1.0; sup[p] = 0.0; rhs[p] = 0.0;
    sub[q] = 0.0; dia[q] = 1.0; sup[q] = 0.0; rhs[q] = 0.0;
    sub[0] = 0.0; sup[N - 1] = 0.0;
    for (int64_t i = 0; i < N; ++i) rhs2[i] = 0.0;
    rhs2[p] = 1.0; rhs2[q] = 1.0;
    thomas2(sub, dia, sup, rhs, rhs2, N, cp);
    const double *U = rhs, *Vv = rhs2;
    double a, b;
    if (m->inner_bc == MBM_BC_FLUX) {
      double al = -rho[0] * m->D_l * jl_inv(dxl_if);
      double be = +V_ip * rho[0] + rho[0] * m->D_l * jl_inv(dxl_if);
      double ga = -V_ip * rho[1] + rho[1] * m->D_r * jl_inv(dxr_if);
      double de = -rho[1] * m->D_r * jl_inv(dxr_if);
      /* al*(U[p-1]+V[p-1]a) + be*a + ga*b + de*(U[q+1]+V[q+1]b) = 0 ;  a = KD*b */
      double ca = be + al * Vv[p - 1], cb = ga + de * Vv[q + 1];
      double r0 = -(al * U[p - 1] + de * U[q + 1]);
      b = r0 / (ca * KD + cb);
      a = KD * b;
    } else if (m->inner_bc == MBM_BC_MB) {
      double sU = 0.0, sVl = 0.0, sVr = 0.0;
      for (int64_t i = 0; i < N; ++i) sU += dvol[i] * U[i];
      for (int64_t i = 0; i < n1; ++i) sVl += dvol[i] * Vv[i];
      for (int64_t i = n1; i < N; ++i) sVr += dvol[i] * Vv[i];
      /* sU + sVl*a + sVr*b = Mass0 ; -a + KD*b = 0 */
      b = (m->Mass0 - sU) / (sVl * KD + sVr);
      a = KD * b;
    } else if (m->inner_bc == MBM_BC_LASAGA) {
      lasaga_bc(m->cl[n1 - 2], m->cr[1], m->D_l, m->D_r, KD, dxl_if, dxr_if, rho[0], rho[1], &a, &b);
    } else {
      a = m->cl[n1 - 1];
      b = m->cr[0];
    }
    for (int64_t i = 0; i < n1; ++i) m->cl[i] = U[i] + Vv[i] * a;
    for (int64_t i = 0; i < n2; ++i) m->cr[i] = U[n1 + i] + Vv[n1 + i] * b;
    m->cl[n1 - 1] = a;
    m->cr[0] = b;
  }
  if (capturing) {
    memcpy(g_capture.sol, m->cl, (size_t)n1 * sizeof(double));
    memcpy(g_capture.sol + n1, m->cr, (size_t)n2 * sizeof(double));
    g_capture.armed = 0;
  }
  /* ---- regrid!, functions.jl:869-905 ------------------------------------------------------ */
  if (Fl_regrid == 1) {
    int64_t nr[2] = { n1, n2 };
    if (V_ip > 0.0) {                                                                     /* :873-880 */
      nr[0] = (int64_t)nearbyint(m->Ri[0] * jl_inv(m->Ri[1] - m->Ri[0]) * (double)nr[1]);
    } else if (V_ip < 0.0) {                                                              /* :881-888 */
      nr[1] = (int64_t)nearbyint(m->Ri[1] * jl_inv(m->Ri[1] - m->Ri[0]) * (double)nr[0]);
    }
    if (nr[0] < m->resmin[0]) nr[0] = m->resmin[0];
    if (nr[1] < m->resmin[1]) nr[1] = m->resmin[1];
    if (nr[0] > m->capL || nr[1] > m->capR) return MBM_ERR_CAPACITY;
    if (nr[0] < 3 || nr[1] < 3) return MBM_ERR_CAPACITY;
    double *XL = tmpA, *XR = tmpB;
    /* tmpA/tmpB sized N+4 each: need nr[0], nr[1] <= N+4?  not guaranteed -> dedicated buffers */
    double *big = (double *)malloc((size_t)(3 * (nr[0] + nr[1]) + 16) * sizeof(double));
    if (!big) return MBM_ERR_CAPACITY;
    XL = big; XR = big + nr[0];
    double *work2 = big + nr[0] + nr[1];
    rc = define_new_grid(m->Ri, nr, m->MRefin, m->flags, XL, XR, &m->newton_R, &m->warn, work2); /* :892 */
    if (rc) { free(big); return rc; }
    double *PL = (double *)malloc((size_t)(nr[0] + nr[1]) * sizeof(double));
    double *PR = PL + nr[0];
    pchip_slopes(m->xl, m->cl, n1, wk);
    rc = pchip_eval(m->xl, m->cl, wk, n1, XL, nr[0], PL, m, 0);                           /* :895 */
    if (!rc) {
      pchip_slopes(m->xr, m->cr, n2, wk);
      rc = pchip_eval(m->xr, m->cr, wk, n2, XR, nr[1], PR, m, 1);                         /* :896 */
    }
    if (rc) { free(big); free(PL); return rc; }
    memcpy(m->xl, XL, (size_t)nr[0] * sizeof(double)); memcpy(m->cl, PL, (size_t)nr[0] * sizeof(double));
    memcpy(m->xr, XR, (size_t)nr[1] * sizeof(double)); memcpy(m->cr, PR, (size_t)nr[1] * sizeof(double));
    m->n1 = nr[0]; m->n2 = nr[1];
    free(big); free(PL);
  }
  m->dx1 = m->xl[m->n1 - 1] - m->xl[m->n1 - 2];                                           /* :893/:901 */
  m->dx2 = m->xr[1] - m->xr[0];                                                           /* :894/:902 */
  /* ---- Stefan: dt of the next step, Chemical_Stefan_problem_XT.jl:147-149 ----------------- */
  if (m->inner_bc == MBM_BC_STEFAN) {
    double mn = (m->dx1 < m->dx2) ? m->dx1 : m->dx2;
    double mxD = (m->D_l > m->D_r) ? m->D_l : m->D_r;
    double dtV = mn * jl_inv(fabs(V_ip));
    double dtD = (mn * mn) * jl_inv(mxD);
    m->dt = ((dtD < dtV) ? dtD : dtV) * m->CFL;
  }
  /* non-finite guard (the reference would carry NaN silently; the ABI reports it) */
  if (!isfinite(m->cl[0]) || !isfinite(m->cr[m->n2 - 1]) || !isfinite(m->cl[m->n1 - 1]) || !isfinite(m->Ri[0]))
    return MBM_ERR_NONFINITE;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* ABI: couple ensembles                                                                       */
/* ------------------------------------------------------------------------------------------ */
/* Threads of the OpenMP loops over members: launchers such as torchrun export OMP_NUM_THREADS=1 to their ranks, which
 * libgomp reads when it is loaded - bench.py's CPU legs ask for the host's cores explicitly and report what they got. */
int mbm_oracle_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

int mbm_version(void) { return MBM_ABI_VERSION; }
int mbm_device_count(void) { return 0; }
const char *mbm_backend(void) { return "oracle-cpu"; }
double mbm_last_kernel_seconds(void) { return 0.0; }
int64_t mbm_last_kernel_launches(void) { return 0; }
int mbm_set_stream(void *s) { (void)s; return 0; }
double mbm_measure_dfma_tflops(int device) { (void)device; return 0.0; }
int mbm_last_tuned_capacity(void) { return 0; }
int mbm_stream_pass_counts(int device, uint64_t *fused_steps, uint64_t *twopass_steps, uint64_t *rows_without_fill_in) {
  (void)device; if (fused_steps) *fused_steps = 0; if (twopass_steps) *twopass_steps = 0;
  if (rows_without_fill_in) *rows_without_fill_in = 0;
  return 0;
}

const char *mbm_status_string(int32_t s) {
  switch (s) {
    case MBM_OK: return "ok";
    case MBM_ERR_IFACE_FAST_POS: return "Interface moved too fast. Ri is larger than x_right[2].";
    case MBM_ERR_IFACE_FAST_NEG: return "Interface moved too fast. Ri is smaller than x_left[end-1].";
    case MBM_ERR_NEWTON: return "Cannot proceed (Newton Failure) - increase Ri[1] or increase grid resolution.";
    case MBM_ERR_KD_RANGE: return "KD exceeds the values of KD_ar.";
    case MBM_ERR_T_RANGE: return "Temperature exceeds the values of T_ar.";
    case MBM_ERR_INTERP_BOUNDS: return "x_interp is out of bounds of x.";
    case MBM_ERR_NDIM: return "Error in geometry (n must be 1,2 or 3).";
    case MBM_ERR_PCHIP_INDEX: return "k out of bounds";
    case MBM_ERR_DENSITY_BOUNDS: return "Value should be within bounds to interpolate.";
    case MBM_ERR_CAPACITY: return "node capacity exceeded";
    case MBM_ERR_NONFINITE: return "non-finite state";
    case MBM_ERR_MREFIN: return "MRefin < 1 is unsupported";
    default: return "unknown status";
  }
}

static int check_common(const mbm_common *c) {
  if (!c) return MBM_RC_BAD_ARG;
  if (c->abi_version != MBM_ABI_VERSION) return MBM_RC_BAD_ARG;
  if (c->n_members < 0 || c->cap_left < 3) return MBM_RC_BAD_ARG;
  return 0;
}

int mbm_couple_batch(const mbm_common *c, const mbm_couple_params *p, mbm_couple_io *io) {
  int rc = check_common(c);
  if (rc) return rc;
  if (!p || !io || c->cap_right < 3) return MBM_RC_BAD_ARG;
  if (c->inner_bc < MBM_BC_FLUX || c->inner_bc > MBM_BC_LASAGA) return MBM_RC_BAD_ARG;
  /* the same argument checks as the CUDA library's entry point (error behaviour is part of the ABI) */
  if (!p->Di || !p->rho || !p->t_tot || !p->CFL || !p->MRefin || !p->BCout || !p->resmin || !p->t_ar || !p->T_ar)
    return MBM_RC_BAD_ARG;
  if (c->inner_bc != MBM_BC_STEFAN && !p->KD_ar) return MBM_RC_BAD_ARG;
  if (!io->x_left || !io->C_left || !io->x_right || !io->C_right || !io->res || !io->Ri || !io->t || !io->dt ||
      !io->it || !io->status)
    return MBM_RC_BAD_ARG;
  if (c->dt_mode < MBM_DT_FIND_DT || c->dt_mode > MBM_DT_SIMPLE) return MBM_RC_BAD_ARG;
  if (c->inner_bc == MBM_BC_STEFAN) {
    if (c->dt_mode != MBM_DT_STEFAN) return MBM_RC_BAD_ARG;
    if (!p->coeff_up || !p->coeff_do || !p->C_leftB || !p->C_rightB) return MBM_RC_BAD_ARG;
  } else if (c->dt_mode == MBM_DT_STEFAN || c->dt_mode == MBM_DT_SIMPLE) return MBM_RC_BAD_ARG;
  if (p->nd < 0) return MBM_RC_BAD_ARG;
  if (p->nd > 0 && (p->nd < 2 || !p->dens_X || !p->dens_T || !p->dens_left || !p->dens_right)) return MBM_RC_BAD_ARG;
  if (p->law != MBM_LAW_NONE && p->law != MBM_LAW_OLIVINE_MELT_D1) return MBM_RC_BAD_ARG;
  if (p->law != MBM_LAW_NONE && (!p->law_par || c->inner_bc != MBM_BC_STEFAN)) return MBM_RC_BAD_ARG;
  const int64_t M = c->n_members, capL = c->cap_left, capR = c->cap_right;
#pragma omp parallel
  {
    scratch_t sc = { 0, 0 };
    double *xl = (double *)malloc((size_t)(capL + 2) * sizeof(double));
    double *cl = (double *)malloc((size_t)(capL + 2) * sizeof(double));
    double *xr = (double *)malloc((size_t)(capR + 2) * sizeof(double));
    double *cr = (double *)malloc((size_t)(capR + 2) * sizeof(double));
#pragma omp for schedule(dynamic, 1)
    for (int64_t k = 0; k < M; ++k) {
      if (io->status[k] != 0) continue;
      member_t m;
      memset(&m, 0, sizeof(m));
      m.ndim = c->ndim; m.inner_bc = c->inner_bc; m.dt_mode = c->dt_mode; m.flags = c->flags;
      m.capL = capL; m.capR = capR;
      for (int j = 0; j < 2; ++j) {
        m.Di[j] = p->Di[2 * k + j];
        m.D0[j] = p->D0 ? p->D0[2 * k + j] : NAN;
        m.Ea[j] = p->Ea ? p->Ea[2 * k + j] : NAN;
        m.rho[j] = p->rho[2 * k + j];
        m.BCout[j] = p->BCout[2 * k + j];
        m.resmin[j] = p->resmin[2 * k + j];
      }
      m.V_ip = p->V_ip ? p->V_ip[k] : 0.0;
      m.t_tot = p->t_tot[k]; m.CFL = p->CFL[k]; m.MRefin = p->MRefin[k]; m.R = p->R;
      m.hist_len = c->hist_len;
      m.t_ar = p->t_ar + k * p->t_stride;
      m.T_ar = p->T_ar + k * p->T_stride;
      m.KD_ar = p->KD_ar ? p->KD_ar + k * p->KD_stride : NULL;
      if (c->inner_bc == MBM_BC_STEFAN) {
        m.coeff_up = p->coeff_up + k * p->coeff_stride;
        m.coeff_do = p->coeff_do + k * p->coeff_stride;
        m.C_leftB = p->C_leftB[k]; m.C_rightB = p->C_rightB[k];
        m.nd = p->nd; m.dens_X = p->dens_X; m.dens_T = p->dens_T;
        m.dens_left = p->dens_left; m.dens_right = p->dens_right;
        m.law = p->law_par ? p->law : MBM_LAW_NONE;
        if (m.law) for (int j = 0; j < 8; ++j) m.law_par[j] = p->law_par[8 * k + j];
      }
      m.xl = xl; m.cl = cl; m.xr = xr; m.cr = cr;
      m.n1 = io->res[2 * k]; m.n2 = io->res[2 * k + 1];
      if (m.ndim < 1 || m.ndim > 3) { io->status[k] = MBM_ERR_NDIM; continue; }           /* functions.jl:512-514 */
      if (m.n1 < 3 || m.n2 < 3 || m.n1 > capL || m.n2 > capR) { io->status[k] = MBM_ERR_CAPACITY; continue; }
      memcpy(xl, io->x_left + k * capL, (size_t)m.n1 * sizeof(double));
      memcpy(cl, io->C_left + k * capL, (size_t)m.n1 * sizeof(double));
      memcpy(xr, io->x_right + k * capR, (size_t)m.n2 * sizeof(double));
      memcpy(cr, io->C_right + k * capR, (size_t)m.n2 * sizeof(double));
      m.Ri[0] = io->Ri[2 * k]; m.Ri[1] = io->Ri[2 * k + 1];
      m.t = io->t[k]; m.dt = io->dt[k]; m.it = io->it[k];
      m.Mass0 = io->Mass0 ? io->Mass0[k] : 0.0;
      m.newton_R = io->newton_R ? io->newton_R[k] : 0.0;
      m.hash = io->idx_hash ? io->idx_hash[k] : 0;
      m.node_steps = io->node_steps ? io->node_steps[k] : 0;
      m.dx1 = xl[m.n1 - 1] - xl[m.n1 - 2];                                                /* create_grid! :366 */
      m.dx2 = xr[1] - xr[0];                                                              /* :367 */
      int st = 0;
      if (c->inner_bc == MBM_BC_STEFAN) {
        m.D_l = m.Di[0]; m.D_r = m.Di[1];                                                 /* CSP :63-64 */
        m.T = m.T_ar[0]; m.KD = NAN;
      } else {
        st = update_t_dependent_param(&m);                                                /* pre-loop call, ..._MB.jl:75 */
      }
      int64_t steps = 0;
      while (!st && m.t < m.t_tot && (c->max_steps == 0 || steps < c->max_steps)) {
        st = couple_step(&m, &sc);
        if (st) break;
        if ((c->flags & MBM_FLAG_RECORD) && io->rec && m.it - 1 < c->rec_cap && m.it >= 1) {
          double *r = io->rec + ((size_t)k * (size_t)c->rec_cap + (size_t)(m.it - 1)) * 4;
          r[0] = calc_mass_vol(xl, xr, cl, cr, m.n1, m.n2, m.ndim, m.rho);
          r[1] = m.T; r[2] = m.KD; r[3] = m.Ri[0];
        }
        ++steps;
      }
      io->status[k] = st;
      if (io->warn) io->warn[k] = m.warn;
      /* the state is written back even for a failed member (it is the state at the throw site) */
      memcpy(io->x_left + k * capL, xl, (size_t)(m.n1 < capL ? m.n1 : capL) * sizeof(double));
      memcpy(io->C_left + k * capL, cl, (size_t)(m.n1 < capL ? m.n1 : capL) * sizeof(double));
      memcpy(io->x_right + k * capR, xr, (size_t)(m.n2 < capR ? m.n2 : capR) * sizeof(double));
      memcpy(io->C_right + k * capR, cr, (size_t)(m.n2 < capR ? m.n2 : capR) * sizeof(double));
      io->res[2 * k] = m.n1; io->res[2 * k + 1] = m.n2;
      io->Ri[2 * k] = m.Ri[0]; io->Ri[2 * k + 1] = m.Ri[1];
      io->t[k] = m.t; io->dt[k] = m.dt; io->it[k] = m.it;
      if (io->dx1) io->dx1[k] = m.dx1;
      if (io->dx2) io->dx2[k] = m.dx2;
      if (io->newton_R) io->newton_R[k] = m.newton_R;
      if (io->idx_hash) io->idx_hash[k] = m.hash;
      if (io->node_steps) io->node_steps[k] = m.node_steps;
      if (io->Mass && (c->flags & MBM_FLAG_MASS_FINAL))
        io->Mass[k] = calc_mass_vol(xl, xr, cl, cr, m.n1, m.n2, m.ndim, m.rho);
      if (io->last) {
        double *l = io->last + 8 * k;
        l[0] = m.D_l; l[1] = m.D_r; l[2] = m.KD; l[3] = m.T; l[4] = m.V_ip;
        l[5] = m.rho[0]; l[6] = m.rho[1]; l[7] = (double)m.Fl_regrid;
      }
    }
    free(xl); free(cl); free(xr); free(cr); free(sc.buf);
  }
  return 0;
}

int mbm_couple_flux_batch(const mbm_common *c, const mbm_couple_params *p, mbm_couple_io *io) {
  if (!c || (c->inner_bc != MBM_BC_FLUX && c->inner_bc != MBM_BC_DIRICHLET_HOLD && c->inner_bc != MBM_BC_LASAGA)) return MBM_RC_BAD_ARG;
  return mbm_couple_batch(c, p, io);
}
int mbm_couple_mb_batch(const mbm_common *c, const mbm_couple_params *p, mbm_couple_io *io) {
  if (!c || c->inner_bc != MBM_BC_MB) return MBM_RC_BAD_ARG;
  return mbm_couple_batch(c, p, io);
}
int mbm_stefan_xt_batch(const mbm_common *c, const mbm_couple_params *p, mbm_couple_io *io) {
  if (!c || c->inner_bc != MBM_BC_STEFAN || c->dt_mode != MBM_DT_STEFAN) return MBM_RC_BAD_ARG;
  if (!p || !p->coeff_up || !p->coeff_do || !p->C_leftB || !p->C_rightB) return MBM_RC_BAD_ARG;
  return mbm_couple_batch(c, p, io);
}

/* ------------------------------------------------------------------------------------------ */
/* ABI: single phase — loop body of main_codes/Simple_Diff_1D_general.jl:67-99                  */
/* ------------------------------------------------------------------------------------------ */
int mbm_simple_diff_batch(const mbm_common *c, const mbm_simple_params *p, mbm_simple_io *io) {
  int rc = check_common(c);
  if (rc) return rc;
  if (!p || !io) return MBM_RC_BAD_ARG;
  const int64_t M = c->n_members, cap = c->cap_left;
#pragma omp parallel
  {
    double *w = (double *)malloc((size_t)(8 * (cap + 4)) * sizeof(double));
#pragma omp for schedule(dynamic, 1)
    for (int64_t k = 0; k < M; ++k) {
      if (io->status[k] != 0) continue;
      const int64_t n = io->res[k];
      if (c->ndim < 1 || c->ndim > 3) { io->status[k] = MBM_ERR_NDIM; continue; }
      if (n < 3 || n > cap) { io->status[k] = MBM_ERR_CAPACITY; continue; }
      double *x = io->x + k * cap, *C = io->C + k * cap;
      double *sub = w, *dia = w + (cap + 4), *sup = w + 2 * (cap + 4), *rhs = w + 3 * (cap + 4);
      double *r2 = w + 4 * (cap + 4), *cp = w + 5 * (cap + 4);
      const double *t_ar = p->t_ar + k * p->t_stride, *T_ar = p->T_ar + k * p->T_stride;
      double t = io->t[k], dt = io->dt[k], t_tot = p->t_tot[k];
      int64_t it = io->it[k];
      double D = 0.0, T = 0.0;
      int st = 0;
      /* update_t_dependent_param_simple!, functions.jl:1402-1423 */
#define UPDATE_SIMPLE()                                                                        \
  do {                                                                                         \
    st = lin_interp_1d(t_ar, T_ar, c->hist_len, t, &T);                                        \
    if (!st) {                                                                                 \
      if (t == t_tot) T = T_ar[c->hist_len - 1];                                               \
      else { double mn, mx; minmax(T_ar, c->hist_len, &mn, &mx);                               \
             if (T < mn - 1e-12 || T > mx + 1e-12) st = MBM_ERR_T_RANGE; }                     \
      if (p->Di[k] == -1.0) D = p->D0[k] * exp(-p->Ea1[k] * jl_inv(p->R * T));                 \
      else D = p->Di[k];                                                                       \
    }                                                                                          \
  } while (0)
      UPDATE_SIMPLE();                                                                    /* SDG :52 */
      int64_t steps = 0, nsteps_nodes = io->node_steps ? io->node_steps[k] : 0;
      while (!st && t < t_tot && (c->max_steps == 0 || steps < c->max_steps)) {
        double Co1 = C[0], Coend = C[n - 1];                                              /* :69 Co = copy(C) */
        dt = p->CFL[k] * (p->dx[k] * p->dx[k]) * jl_inv(D);                               /* :71, functions.jl:166 */
        update_time(&t, &dt, &it, t_tot);                                                 /* :73 */
        UPDATE_SIMPLE();                                                                  /* :75 */
        if (st) break;
        fill_matrix(C, x, D, dt, c->ndim, n, sub, dia, sup, rhs);                         /* :78 */
        double sdiag = 0.0;
        for (int64_t i = 0; i < n; ++i) sdiag += dia[i];
        double ScF = sdiag * jl_inv((double)n);                                           /* :80 */
        if (p->BCout[2 * k] == 1) { sub[0] = 0.0; dia[0] = 1.0 * ScF; sup[0] = 0.0; rhs[0] = Co1 * ScF; }          /* :82 */
        if (p->BCout[2 * k + 1] == 1) { sub[n - 1] = 0.0; dia[n - 1] = 1.0 * ScF; sup[n - 1] = 0.0; rhs[n - 1] = Coend * ScF; }
        if (c->flags & MBM_FLAG_SOLVER_FAITHFUL) {
          /* sparse `\` = UMFPACK LU with partial pivoting; on a tridiagonal matrix that is banded LU
           * with row pivoting; restated with the dense routine (N <= a few hundred in the KATs) */
          double *A = (double *)calloc((size_t)n * (size_t)n, sizeof(double));
          for (int64_t i = 0; i < n; ++i) {
            A[i * n + i] = dia[i];
            if (i > 0) A[i * n + i - 1] = sub[i];
            if (i + 1 < n) A[i * n + i + 1] = sup[i];
          }
          if (solve_dense(A, rhs, n)) st = MBM_ERR_NONFINITE;
          free(A);
        } else {
          for (int64_t i = 0; i < n; ++i) r2[i] = 0.0;
          sub[0] = 0.0; sup[n - 1] = 0.0;
          thomas2(sub, dia, sup, rhs, r2, n, cp);                                         /* :84 */
        }
        memcpy(C, rhs, (size_t)n * sizeof(double));
        nsteps_nodes += n;
        ++steps;
        if (!isfinite(C[0]) || !isfinite(C[n - 1])) st = MBM_ERR_NONFINITE;
      }
#undef UPDATE_SIMPLE
      io->status[k] = st;
      io->t[k] = t; io->dt[k] = dt; io->it[k] = it;
      if (io->D) io->D[k] = D;
      if (io->node_steps) io->node_steps[k] = nsteps_nodes;
      if (io->Mass && (c->flags & MBM_FLAG_MASS_FINAL))                                   /* calc_mass_vol_simple_diff :237-254 */
        io->Mass[k] = trapz(x, p->rho[k], c->ndim, C, n);
    }
    free(w);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* ABI: pre-loop helpers                                                                       */
/* ------------------------------------------------------------------------------------------ */
int mbm_create_grid(int32_t M, int32_t capL, int32_t capR, const double *Ri, const int64_t *res,
                    const double *MRefin, double *x_left, double *x_right, double *dx1, double *dx2,
                    double *newton_R, int32_t *status) {
  for (int64_t k = 0; k < M; ++k) {
    int64_t nr[2] = { res[2 * k], res[2 * k + 1] };
    if (nr[0] < 3 || nr[1] < 3 || nr[0] > capL || nr[1] > capR) { status[k] = MBM_ERR_CAPACITY; continue; }
    double *work = (double *)malloc((size_t)(3 * (nr[0] + nr[1]) + 16) * sizeof(double));
    int warn = 0;
    double nR = 0.0;
    double ri[2] = { Ri[2 * k], Ri[2 * k + 1] };
    int rc = define_new_grid(ri, nr, MRefin[k], 0u, x_left + (size_t)k * capL, x_right + (size_t)k * capR,
                             &nR, &warn, work);
    free(work);
    status[k] = rc;
    if (rc) continue;
    const double *xl = x_left + (size_t)k * capL, *xr = x_right + (size_t)k * capR;
    if (dx1) dx1[k] = xl[nr[0] - 1] - xl[nr[0] - 2];                                      /* :366 */
    if (dx2) dx2[k] = xr[1] - xr[0];                                                      /* :367 */
    if (newton_R) newton_R[k] = nR;
  }
  return 0;
}

int mbm_calc_mass_vol(int32_t M, int32_t ndim, int32_t capL, int32_t capR, const double *x_left,
                      const double *x_right, const double *C_left, const double *C_right,
                      const int64_t *res, const double *rho, double *Mass) {
  for (int64_t k = 0; k < M; ++k) {
    double r[2] = { rho[2 * k], rho[2 * k + 1] };
    Mass[k] = calc_mass_vol(x_left + (size_t)k * capL, x_right + (size_t)k * capR, C_left + (size_t)k * capL,
                            C_right + (size_t)k * capR, res[2 * k], res[2 * k + 1], ndim, r);
  }
  return 0;
}

/* oracle-only: create_grid with explicit flags (to exercise MBM_FLAG_SUM_JULIA in tests) */
int mbm_oracle_create_grid_flags(int32_t M, int32_t capL, int32_t capR, const double *Ri, const int64_t *res,
                                 const double *MRefin, uint32_t flags, double *x_left, double *x_right,
                                 int32_t *status) {
  for (int64_t k = 0; k < M; ++k) {
    int64_t nr[2] = { res[2 * k], res[2 * k + 1] };
    double *work = (double *)malloc((size_t)(3 * (nr[0] + nr[1]) + 16) * sizeof(double));
    int warn = 0;
    double nR = 0.0, ri[2] = { Ri[2 * k], Ri[2 * k + 1] };
    status[k] = define_new_grid(ri, nr, MRefin[k], flags, x_left + (size_t)k * capL, x_right + (size_t)k * capR,
                                &nR, &warn, work);
    free(work);
  }
  return 0;
}

/* oracle-only: expose pchip for unit tests (functions.jl:775-839) */
int mbm_oracle_pchip(const double *x, const double *y, int64_t n, const double *X, int64_t nX, double *P) {
  member_t m; memset(&m, 0, sizeof(m));
  double *d = (double *)malloc((size_t)n * sizeof(double));
  pchip_slopes(x, y, n, d);
  int rc = pchip_eval(x, y, d, n, X, nX, P, &m, 0);
  free(d);
  return rc;
}
/* oracle-only: newton_solver (functions.jl:732-757) -> root, iterations not reported */
int mbm_oracle_newton(double S, double d, int64_t n, double x0, double *root, int *warn) {
  return newton_solver(S, d, n, 1e-8, 200, x0, root, warn);
}
double mbm_oracle_pow_int(double x, int64_t n) { return jl_pow_int(x, n); }
void mbm_oracle_scan(const double *v, int64_t n, double *out, int op, int julia) {
  if (julia && op == 0) cumsum_julia(v, n, out); else scan_canonical(v, n, out, op);
}

/* ------------------------------------------------------------------------------------------ */
/* Validation layer: the known-answer gates of the reference's tests,   */
/* per member, from the ensemble's state.                                                      */
/* ------------------------------------------------------------------------------------------ */
/* LinRange(a, b, n)[i] (0-based): Julia's lerpi, (1 - t)*a + t*b with t = i/(n - 1) */
static double linrange_val(double a, double b, int64_t n, int64_t i) {
  double t = (double)i / (double)(n - 1);
  return (1.0 - t) * a + t * b;
}

/* test/test_C1.jl:9-12 with rayleigh_fractionation, Benchmarks.jl:239-260 */
static void validate_rayleigh(const double *xl, const double *cl, int64_t n1, const double Ri0[2], double Ri1,
                              double C0r, double KD0, int ndim, double *int_Ray, double *int_model) {
  /* Ray_Fs = LinRange(Ri0[1]/Ri0[2], Ri[1]/Ri0[2], 1000).^n; Ray_Cs = (1 - Fs)^(KD0 - 1)*C0_r[end]*KD0   :245-248 */
  const double a = Ri0[0] * jl_inv(Ri0[1]), b = Ri1 * jl_inv(Ri0[1]);
  double acc = 0.0, Fp = 0.0, Cp = 0.0;
  for (int64_t i = 0; i < 1000; ++i) {
    double Fs = pow_ndim(linrange_val(a, b, 1000, i), ndim);
    double Cs = pow(1.0 - Fs, KD0 - 1.0) * C0r * KD0;
    if (i > 0) { double v = 0.5 * (Fs - Fp) * (Cp + Cs); acc = (i == 1) ? v : acc + v; }   /* trapezoidal_integration */
    Fp = Fs; Cp = Cs;
  }
  *int_Ray = acc;
  /* post-processing :250-258 as shipped: indi = findall(x -> x > Ri0[1], x_left) and x_left[LinearIndices(indi)],
   * i.e. the FIRST length(indi) nodes; prepended by (Ri0[1], C interpolated at Ri0[1])                          */
  int64_t k = 0;
  for (int64_t i = 0; i < n1; ++i) if (xl[i] > Ri0[0]) ++k;
  double Cbc = NAN;
  lin_interp_1d(xl, cl, n1, Ri0[0], &Cbc);
  const double iR = jl_inv(Ri0[1]);
  double xp = pow_ndim(Ri0[0] * iR, ndim), yp = Cbc;
  acc = 0.0;
  for (int64_t i = 0; i < k; ++i) {
    double xq = pow_ndim(xl[i] * iR, ndim), yq = cl[i];
    double v = 0.5 * (xq - xp) * (yp + yq);
    acc = (i == 0) ? v : acc + v;
    xp = xq; yp = yq;
  }
  *int_model = acc;
}

/* test/test_B5.jl with smith, Benchmarks.jl:284-301 (n = 1) */
static void validate_smith(const double *xr, const double *cr, int64_t n2, double Ri1, double Dan, double tan_, double Kan,
                           double R, double *ndiff, double *nref) {
  const double ksi = 0.5 * pow(Dan * tan_, -0.5);
  const double Can0 = cr[n2 - 1];
  double sd = 0.0, sa = 0.0, sb = 0.0;
  for (int64_t i = 0; i < n2; ++i) {
    double xan = xr[i] - Ri1;
    double Can = 1.0 + ((1.0 - Kan) * jl_inv(2.0 * Kan)) * exp(-R * jl_inv(Dan) * xan) * erfc(ksi * (xan - R * tan_))
                 - 0.5 * erfc(ksi * (xan + R * tan_))
                 + ((1.0 - Kan) * jl_inv(2.0)) * (1.0 * jl_inv(1.0 - Kan) - 1.0 * jl_inv(Kan))
                   * exp(-(1.0 - Kan) * (R * jl_inv(Dan)) * (xan + Kan * R * tan_))
                   * erfc(ksi * (xan + (2.0 * Kan - 1.0) * R * tan_));
    Can = Can * Can0;
    sd += (Can - cr[i]) * (Can - cr[i]); sa += Can * Can; sb += cr[i] * cr[i];
  }
  *ndiff = sqrt(sd);
  *nref = sqrt(sa > sb ? sa : sb);
}


/* calc_sinus_sphere (src/Benchmarks.jl:82-99) / crank_time_transformation3 (:192-212) with zeta = time-integrated
 * diffusivity, then isapprox as test/test_B1.jl / test_B2.jl (norm-wise).  Per node the terms are added in the reference's
 * order nt = 1..1000; once exp(-zeta nt^2 pi^2 / L^2) has underflowed to zero every further term is an exact zero, so the
 * loop stops there (bit-identical sum). */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
static void validate_sphere(const double *xl, const double *cl, const double *c0l, int64_t n1, const double *xr,
                            const double *cr, const double *c0r, int64_t n2, double zeta, double *o0, double *o1) {
  const int64_t N = n1 + n2;
  const double L = xr[n2 - 1], iL2 = jl_inv(L * L);
  double sd = 0.0, sa = 0.0, sb = 0.0;
  for (int64_t i = 0; i < N; ++i) {
    const double x = (i < n1) ? xl[i] : xr[i - n1];
    const double c0 = (i < n1) ? c0l[i] : c0r[i - n1];
    const double c = (i < n1) ? cl[i] : cr[i - n1];
    double Can;
    if (i == N - 1) Can = c0;                                                              /* :97  C[end] = Cs */
    else {
      double acc = c0;
      for (int nt = 1; nt <= 1000; ++nt) {
        const double e = exp(-zeta * (double)(nt * nt) * (M_PI * M_PI) * iL2);
        if (e == 0.0) break;
        const double sg = (nt & 1) ? -1.0 : 1.0;
        if (i == 0) acc = acc + 2.0 * (sg * e);                                            /* :94  C1 */
        else acc = acc + 2.0 * L * jl_inv(M_PI * x) * sg * jl_inv((double)nt) * sin((double)nt * M_PI * x * jl_inv(L)) * e;   /* :93 */
      }
      Can = acc + 1.0;                                                                     /* :95-96 */
    }
    sd += (Can - c) * (Can - c); sa += Can * Can; sb += c * c;
  }
  *o0 = sqrt(sd);
  *o1 = sqrt(sa > sb ? sa : sb);
}

int mbm_validate_batch(const mbm_validate_args *a) {
  if (!a || a->abi_version != MBM_ABI_VERSION) return MBM_RC_BAD_ARG;
  if (a->n_members < 0 || a->cap_left < 3 || a->cap_right < 3 || a->ndim < 1 || a->ndim > 3) return MBM_RC_BAD_ARG;
  if (a->kind < MBM_VAL_MASS_ERR || a->kind > MBM_VAL_SPHERE) return MBM_RC_BAD_ARG;
  if (a->flags & MBM_FLAG_DEVICE_POINTERS) return MBM_RC_UNSUPPORTED;
  if (!a->x_left || !a->C_left || !a->x_right || !a->C_right || !a->res || !a->Ri || !a->out || !a->relerr) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_MASS_ERR && (!a->rho || !a->Mass0)) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_RAYLEIGH && (!a->Ri0 || !a->C0_r || !a->KD0)) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_SMITH && (!a->KD0 || !a->Di_r || !a->V_ip || !a->t_tot || a->ndim != 1)) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_SPHERE && (!a->zeta || !a->C0_left || !a->C0_right)) return MBM_RC_BAD_ARG;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t m = 0; m < a->n_members; ++m) {
    const double *xl = a->x_left + m * a->cap_left, *cl = a->C_left + m * a->cap_left;
    const double *xr = a->x_right + m * a->cap_right, *cr = a->C_right + m * a->cap_right;
    const int64_t n1 = a->res[2 * m], n2 = a->res[2 * m + 1];
    double o0 = NAN, o1 = NAN, e = NAN;
    if (n1 >= 2 && n2 >= 2 && n1 <= a->cap_left && n2 <= a->cap_right) {
      if (a->kind == MBM_VAL_MASS_ERR) {
        const double rho[2] = { a->rho[2 * m], a->rho[2 * m + 1] };
        o0 = calc_mass_vol(xl, xr, cl, cr, n1, n2, a->ndim, rho);
        o1 = a->Mass0[m];
        e = (o0 - o1) * jl_inv(o1);                                                       /* functions.jl:184 */
      } else if (a->kind == MBM_VAL_RAYLEIGH) {
        validate_rayleigh(xl, cl, n1, a->Ri0 + 2 * m, a->Ri[2 * m], a->C0_r[m], a->KD0[m], a->ndim, &o0, &o1);
        const double den = fabs(o0) > fabs(o1) ? fabs(o0) : fabs(o1);
        e = den > 0.0 ? fabs(o0 - o1) / den : 0.0;
      } else if (a->kind == MBM_VAL_SPHERE) {
        validate_sphere(xl, cl, a->C0_left + m * a->cap_left, n1, xr, cr, a->C0_right + m * a->cap_right, n2, a->zeta[m], &o0, &o1);
        e = o1 > 0.0 ? o0 / o1 : 0.0;
      } else {
        validate_smith(xr, cr, n2, a->Ri[2 * m], a->Di_r[m], a->t_tot[m], a->KD0[m], a->V_ip[m], &o0, &o1);
        e = o1 > 0.0 ? o0 / o1 : 0.0;
      }
    }
    a->out[2 * m] = o0; a->out[2 * m + 1] = o1; a->relerr[m] = e;
  }
  return 0;
}
