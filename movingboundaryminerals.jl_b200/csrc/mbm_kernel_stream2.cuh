// mbm_kernel_stream2.cuh — second streaming kernel for long profiles on uniform meshes (BASELINE.json configs[4]:
// n = 65,536 nodes per phase, MRefin = 1).  Same chunk-interleaved layout, same row staging by bulk copy and the same
// arithmetic per row as mbm_kernel_stream.cuh, but only the CONCENTRATIONS cross HBM:
//
//  * MESHES ARE CLOSED-FORM.  With MRefin == 1 both meshes of a step are LinRanges (define_new_grid, functions.jl:396-400)
//    of the interface position they were generated with, except the one node the interface advection moves or inserts
//    (advect_interface_regrid!, :29-64).  Every x the step needs is recomputed from (a, b, len, i) with the reference's own
//    operations (Base.lerpi: t = i/d, (1-t)*a + t*b; bit-identical) instead of being stored, loaded and re-stored.  The
//    division is i * RN(1/d) corrected by one FMA (exactly RN(i/d) for integers 0 <= i <= d <= 300,000:
//    tools/linrange_div_check.c checks every pair; beyond that range the kernel is not selected).  At load the member's
//    x arrays are compared bit for bit with the closed form; a member whose mesh is anything else is handed to the first
//    streaming kernel (status -1), which keeps x in memory.
//  * BACK-SUBSTITUTION AND REGRID IN ONE PASS.  On an ordinary step the node counts and the layout do not change and new
//    node p lies in old interval p + o, o in {-1, 0} per phase.  The solved value of a row is then used where it is
//    produced: each column walks its rows towards the outer boundary, back-substitutes, and evaluates the new node whose
//    pchip stencil (four old nodes, functions.jl:775-839) has just been completed - the solved profile itself is never
//    written.  The two new nodes at either end of a column need solved values of the neighbouring columns: every
//    column publishes its first and last four solved rows in shared memory and those nodes are evaluated after the pass
//    (general evaluation, reference node order: equal / clamp branches, three-point end slopes).  The alignment is
//    verified node by node (exact comparisons on the closed-form meshes); a step that breaks it - node counts changed by
//    regrid!, a layout change - is redone by the two-pass path (back-substitution, then the interval-driven walk of the
//    first streaming kernel on closed-form meshes), which is also the path of short columns.
//
//  * ROWS OF THE ELIMINATION interleaved row by row (one row pointer, immediate offsets, one bulk copy per batch), batches
//    of interior rows as straight-line code, and a FILL-IN CUTOFF: beyond the decay length of a chunk's fill-in column the
//    rows are stored as (R, G) pairs (see S2_A_DROP below).  Fixed-interface members (V_ip == 0) run the same kernel
//    without the regrid.
//
// Traffic per node-step: forward sweep reads C, writes (R, A, G) or (R, G); second pass reads them back, writes the new C:
// ~54 B (64 B without the cutoff) against the 104 B of the first streaming kernel (algorithmic figure, SURVEY.md §8(d):
// 32 B with x, 16 B for C alone).
#pragma once
#include <type_traits>
#include "mbm_kernel_stream.cuh"

namespace mbm {

constexpr int S2_MAX_D = 300000;          // largest LinRange divisor the corrected reciprocal division is checked for
constexpr int S2_MIN_ROWS = 16;           // shortest column the fused pass takes
constexpr int S2_MAX_FIX = 10;            // most deferred nodes per column the fused pass accepts
// Fill-in cutoff.  The fill-in column A_k of a chunk (its rows' dependence on the previous separator) decays geometrically
// along the chunk (|A_k| = |sub inv| |A_k-1|, FEM rows are strictly diagonally dominant): after ~40..170 of a column's
// ~500 rows A_k s_prev is below the last bit of C_k.  Rows from Kc on (a multiple of the batch length, per member and
// step) are therefore stored as (R, G) pairs: every |A_k| dropped is checked against S2_A_DROP in the sweep itself (a
// violation repeats the sweep without cutoff), Kc of the next step comes from where the columns fell below S2_A_ZERO.
constexpr double S2_A_DROP = 1e-19, S2_A_ZERO = 1e-21;

// diagnostics (tests): steps taken by the fused pass / by the two-pass path since the library was loaded
__device__ unsigned long long g_s2_fused_steps = 0ull, g_s2_twopass_steps = 0ull, g_s2_cut_rows = 0ull;

// RN(i/d) for integers 0 <= i <= d <= S2_MAX_D, r = RN(1/d): q0 = i*r is within an ulp of the quotient, the remainder
// i - q0*d is exact in one FMA, and the corrected quotient rounds correctly (checked exhaustively)
__device__ __forceinline__ double s2_div(double fi, double fd, double r) {
  const double q0 = __dmul_rn(fi, r);
  const double rem = __fma_rn(-q0, fd, fi);
  return __fma_rn(rem, r, q0);
}
// a LinRange(a, b, len) addressed by the phase-ordered index p: node i = i0 + s*p (left: i = p; right: i = len-1-p)
struct Mesh1 { double a, b, fd, r, fi0, s; };
__device__ __forceinline__ Mesh1 s2_mesh(double a, double b, int len, bool reversed) {
  Mesh1 M;
  const int d = (len - 1 > 1) ? (len - 1) : 1;
  M.a = a; M.b = b; M.fd = (double)d; M.r = xinv(M.fd);
  M.fi0 = reversed ? (double)(len - 1) : 0.0; M.s = reversed ? -1.0 : 1.0;
  return M;
}
// LinRange(a,b,len)[i] exactly as linrange_at_inl (Base.lerpi), i given as a double
__device__ __forceinline__ double s2_lerp(const Mesh1 &M, double fi) {
  const double t = s2_div(fi, M.fd, M.r);
  return xadd(xmul(xsub(1.0, t), M.a), xmul(t, M.b));
}
__device__ __forceinline__ double s2_at(const Mesh1 &M, int p) { return s2_lerp(M, M.fi0 + M.s * (double)p); }

// the step's geometry (thread 0 writes it in the scalar phase, every thread reads it)
struct Geo2 {
  Mesh1 MO[2];                // meshes the state was generated on (before the advection), per phase
  Mesh1 MN[2];                // meshes of the regrid
  double Rnew;                // interface position of the step: the x of each phase's last node (phase order) in the solve
  int n[2], nr[2];            // node counts of the solve / after the regrid
  int o[2];                   // fused pass: new node p lies in old interval [p + o, p + o + 1] (phase order)
  int fast;
};
// x of the solve's mesh at phase-ordered index p (the interface node of either phase is the one node off the LinRange)
__device__ __forceinline__ double s2_xs(const Geo2 &G, int ph, int p) {
  return (p == G.n[ph] - 1) ? G.Rnew : s2_at(G.MO[ph], p);
}

// three-point end slope (functions.jl:786-806) and general evaluation of one new node (:809-837) on accessors, in the
// reference's node order: xf(i), yf(i), i = 0..n-1.  c0: guess of the number of old nodes strictly below Xj; the
// search gives up after 6 steps (*bad).  Same formulas as pchip_eval_otf / pchip_slope_v2.
template <class XF, class YF>
__device__ __forceinline__ double s2_end_slope(XF xf, YF yf, int n, int i) {
  const int a = (i == 0) ? 0 : n - 1, b = (i == 0) ? 1 : n - 2, c = (i == 0) ? 2 : n - 3;
  const double sg = (i == 0) ? 1.0 : -1.0;
  const double xa = xf(a), xb = xf(b), xc = xf(c), ya = yf(a), yb = yf(b), yc = yf(c);
  const double h1 = sg * (xb - xa), h2 = sg * (xc - xb);
  const double m1 = sg * (yb - ya) * frcp_(h1), m2 = sg * (yc - yb) * frcp_(h2);
  double d = fma(fma(2.0, h1, h2), m1, -(h1 * m2)) * frcp_(h1 + h2);
  if (m1 * m2 < 0.0 || m1 == 0.0 || m2 == 0.0) d = 0.0;
  return d;
}
template <class XF, class YF>
__device__ __noinline__ double s2_eval_general(XF xf, YF yf, const int n, const double Xj, const int c0, int *kout,
                                               int *branch, bool *bad) {
  int c = min(max(c0, 0), n);
  int steps = 0;
  while (c < n && xf(c) < Xj) { ++c; if (++steps > 6) { *bad = true; break; } }
  while (c > 0 && !(xf(c - 1) < Xj)) { --c; if (++steps > 6) { *bad = true; break; } }
  if (c < n && xf(c) == Xj) { *kout = c; *branch = 0; return yf(c); }
  if (c == n) { *kout = n - 1; *branch = 1; return yf(n - 1); }
  if (c == 0) { *kout = 0; *branch = 2; return yf(0); }
  const int k = c - 1;
  *kout = k; *branch = 3;
  const double x1 = xf(k), x2 = xf(k + 1), y1 = yf(k), y2 = yf(k + 1);
  const double h = x2 - x1;
  const double ih = frcp_(h), m1 = (y2 - y1) * ih;
  double d1, d2;
  if (k == 0) d1 = s2_end_slope(xf, yf, n, 0);
  else { const double ihm = frcp_(x1 - xf(k - 1)); d1 = slope_interior(ihm, ih, (y1 - yf(k - 1)) * ihm, m1); }
  if (k + 1 == n - 1) d2 = s2_end_slope(xf, yf, n, n - 1);
  else { const double ihp = frcp_(xf(k + 2) - x2); d2 = slope_interior(ih, ihp, m1, (yf(k + 2) - y2) * ihp); }
  const double s = Xj - x1, u = s * ih, smh = s - h;
  const double u2 = u * u, w2 = u2 * (3.0 - 2.0 * u);
  return w2 * y2 + (1.0 - w2) * y1 + (u2 * smh) * d2 + (u * smh) * (smh * ih) * d1;
}

// ---- two-pass path, second pass: the interval-driven walk of pchip_run_stream (mbm_kernel_stream.cuh) with the old x
// from the closed form (xo(p), phase order) and no mesh output.  Same walk, same formulas, same checksum terms.
template <int T, bool do_hash, class XO, class XN>
__device__ __forceinline__ void pchip_run_stream2(XO xo, XN xn, const double *C, const StreamLayout &L, const int phase,
                                                  const int n, const int nn, double *Cn, const int idx0, const int cnt,
                                                  const long long it, unsigned long long &hash) {
  const int tid = threadIdx.x;
  const bool MIR = (phase != 0);
  const double sg = MIR ? -1.0 : 1.0;
  auto opt = [&](int i) { PPt q; q.x = sg * xo(i); q.y = C[paddr(T, L, phase, i)]; return q; };
  auto oxs = [&](int i) { return sg * xo(i); };
  auto emit = [&](int k, double val, int kk, int br) {
    Cn[saddr(T, tid, k)] = val;
    if (do_hash) {
      if (MIR) {
        if (br == 0) kk = n - 1 - kk;
        else if (br == 1) { kk = 0; br = 2; }
        else if (br == 2) { kk = n - 1; br = 1; }
        else kk = n - 2 - kk;
      }
      hash += idx_term(it, phase, MIR ? (nn - 1 - (idx0 + k)) : (idx0 + k), kk, br);
    }
  };
  double Xj = sg * xn(idx0);
  int c = min(max((int)(((long long)idx0 * n) / nn), 1), n - 1);
  {
    double xa = oxs(c - 1), xb0 = oxs(c);
    int steps = 0;
    while (xb0 < Xj && c < n) {
      ++c; xa = xb0; xb0 = (c < n) ? oxs(c) : xb0;
      if (++steps > 64) break;
    }
    while (!(xa < Xj) && c > 0 && steps <= 64) {
      --c; xa = (c > 0) ? oxs(c - 1) : xa;
      ++steps;
    }
    if (steps > 64) {
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (oxs(mid) < Xj) lo = mid + 1; else hi = mid; }
      c = lo;
    }
  }
  int j = 0;
  while (c == 0) {
    const PPt q0 = opt(0);
    if (!(Xj <= q0.x)) { c = 1; break; }
    emit(j, q0.y, 0, (Xj == q0.x) ? 0 : 2);
    if (++j == cnt) return;
    Xj = sg * xn(idx0 + j);
  }
  int k = c - 1;
  if (k <= n - 2) {
    PPt Q0 = opt(k), Q1 = opt(k + 1), Q2 = Q1, Q3 = Q1;
    double IHa = frcp_(Q1.x - Q0.x), Ma = (Q1.y - Q0.y) * IHa, IHb = 0.0, Mb = 0.0, Da, Db;
    if (k == 0) Da = end_slope3(Q0, Q1, opt(2));
    else { const PPt Qm = opt(k - 1); const double ihm = frcp_(Q0.x - Qm.x); Da = slope_interior(ihm, IHa, (Q0.y - Qm.y) * ihm, Ma); }
    if (k + 1 == n - 1) Db = end_slope3(Q1, Q0, opt(n - 3));
    else { Q2 = opt(k + 2); IHb = frcp_(Q2.x - Q1.x); Mb = (Q2.y - Q1.y) * IHb; Db = slope_interior(IHa, IHb, Ma, Mb); }
    int wrow = 0; size_t wa = 0;
    if (k + 3 <= n - 1) { const int i3 = k + 3; wrow = i3 % L.m; wa = (size_t)wrow * T + (size_t)((phase ? L.colR0 : 0) + i3 / L.m); }
#define MBM_S2STEP(P0, P1, P2, P3, IH0, M0, D0, IH1, M1, D1)                                                  \
    {                                                                                                          \
      const double h = P1.x - P0.x;                                                                            \
      while (Xj <= P1.x) {                                                                                     \
        const double sx = Xj - P0.x;                                                                           \
        const double u = sx * IH0, smh = sx - h;                                                               \
        const double u2 = u * u;                                                                               \
        const double w2 = u2 * (3.0 - 2.0 * u);                                                                \
        double val = w2 * P1.y + (1.0 - w2) * P0.y + (u2 * smh) * D1 + (u * smh) * (smh * IH0) * D0;           \
        const bool eq = (Xj == P1.x);                                                                          \
        if (eq) val = P1.y;                                                                                    \
        emit(j, val, eq ? k + 1 : k, eq ? 0 : 3);                                                              \
        if (++j == cnt) return;                                                                                \
        Xj = sg * xn(idx0 + j);                                                                                \
      }                                                                                                        \
      if (k + 3 <= n - 1) {                                                                                    \
        if (wrow + SPF < L.m) asm volatile("prefetch.global.L1 [%0];" ::"l"(C + wa + (size_t)SPF * T));        \
        P3.x = sg * xo(k + 3); P3.y = C[wa];                                                                   \
        if (++wrow == L.m) { wrow = 0; wa = wa + 1 - (size_t)(L.m - 1) * T; } else wa += T;                    \
        IH0 = frcp_(P3.x - P2.x); M0 = (P3.y - P2.y) * IH0;                                                    \
        D0 = slope_interior(IH1, IH0, M1, M0);                                                                 \
      } else if (k + 3 == n) D0 = end_slope3(P2, P1, P0);                                                      \
      else break;                                                                                              \
      ++k;                                                                                                     \
    }
    for (;;) {
      MBM_S2STEP(Q0, Q1, Q2, Q3, IHa, Ma, Da, IHb, Mb, Db)
      MBM_S2STEP(Q1, Q2, Q3, Q0, IHb, Mb, Db, IHa, Ma, Da)
      MBM_S2STEP(Q2, Q3, Q0, Q1, IHa, Ma, Da, IHb, Mb, Db)
      MBM_S2STEP(Q3, Q0, Q1, Q2, IHb, Mb, Db, IHa, Ma, Da)
    }
#undef MBM_S2STEP
  }
  {
    const double yl = opt(n - 1).y;
    for (;;) {
      emit(j, yl, n - 1, 1);
      if (++j == cnt) return;
    }
  }
}

template <int T, int NDIM>
__global__ void __launch_bounds__(T, MBM_STREAM_MINB) mbm_stream2_kernel(const KArgs A, const int member0) {
  __shared__ Scal S;
  // ten rows of T doubles: the reduced system and the chunk functionals during the solve; rows 0..7 are re-used by the
  // fused pass for the solved values each column publishes (first four rows: 0..3, last four rows: 4..7)
  __shared__ double sh[10][T];
  double *const redS = sh[0], *const redD = sh[1], *const redU = sh[2], *const exA = sh[3], *const exG = sh[4],
               *const exR = sh[5], *const red0 = sh[8], *const red1 = sh[9];
  __shared__ StreamLayout Lcur, Lnew;
  __shared__ Geo2 G;
  __shared__ int s2_kz, s2_Kc;                 // fill-in cutoff: where this step's columns fell below S2_A_ZERO / cutoff row in use
  extern __shared__ __align__(128) double stg[];
  __shared__ __align__(8) unsigned long long fullb[STG_NBF];
  unsigned stg_par = 0;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int mem = member0 + blockIdx.x;
  if (mem >= A.M) return;
  if (tid == 0) {
    for (int b = 0; b < STG_NBF; ++b) mbar_init(&fullb[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int MC = A.kcapL;
  const size_t ASZ = (size_t)MC * T;
  // the first streaming kernel's workspace (6 arrays per member): the state C, the regrid's output Cn, and the rows of
  // the elimination interleaved row by row, RAG[(3 k + a) T + t], a = 0: R, 1: A (fill-in), 2: G - one row pointer and
  // immediate offsets for the stores of the forward sweep, one bulk copy per batch for the second pass
  double *base = A.workspace + (size_t)blockIdx.x * 6 * ASZ;
  double *C = base, *Cn = base + ASZ, *const RAG = base + 2 * ASZ;
  const int bc = A.inner_bc;
  const bool mb = (bc == BC_MB);
  MemberPtrs P;
  P.t_ar = A.t_ar + (size_t)mem * A.t_stride;
  P.T_ar = A.T_ar + (size_t)mem * A.T_stride;
  P.KD_ar = A.KD_ar + (size_t)mem * A.KD_stride;

  // ---------------- load member ----------------
  if (tid == 0) {
    S.status = A.status[mem];
    S.warn = 0;
    S.t = A.t[mem]; S.dt = A.dt[mem]; S.it = A.it[mem]; S.steps = 0;
    S.t_tot = A.t_tot[mem]; S.CFL = A.CFL[mem];
    S.Di0 = A.Di[2 * (size_t)mem]; S.Di1 = A.Di[2 * (size_t)mem + 1];
    S.D00 = A.D0 ? A.D0[2 * (size_t)mem] : 0.0; S.D01 = A.D0 ? A.D0[2 * (size_t)mem + 1] : 0.0;
    S.Ea0 = A.Ea ? A.Ea[2 * (size_t)mem] : 0.0; S.Ea1 = A.Ea ? A.Ea[2 * (size_t)mem + 1] : 0.0;
    S.R = A.R;
    S.arrhenius = (S.Di0 == -1.0 && S.Di1 == -1.0);
    S.bc0 = A.BCout[2 * (size_t)mem]; S.bc1 = A.BCout[2 * (size_t)mem + 1];
    S.hash = A.idx_hash ? A.idx_hash[mem] : 0ull;
    S.node_steps = A.node_steps ? A.node_steps[mem] : 0;
    S.hidx = 0; S.Fl = 0; S.KD = 0.0; S.T = 0.0; S.D_l = 0.0; S.D_r = 0.0;
    S.n1 = (int)A.res[2 * (size_t)mem]; S.n2 = (int)A.res[2 * (size_t)mem + 1];
    S.rho1 = A.rho[2 * (size_t)mem]; S.rho2 = A.rho[2 * (size_t)mem + 1];
    S.V = A.V_ip ? A.V_ip[mem] : 0.0;
    S.MRefin = A.MRefin[mem];
    S.resmin1 = A.resmin[2 * (size_t)mem]; S.resmin2 = A.resmin[2 * (size_t)mem + 1];
    S.Ri1 = A.Ri[2 * (size_t)mem]; S.Ri2 = A.Ri[2 * (size_t)mem + 1];
    S.Mass0 = A.Mass0 ? A.Mass0[mem] : 0.0;
    S.newton_R = A.newton_R ? A.newton_R[mem] : 0.0;
    bool mine = A.only_deferred ? (S.status == -1) : (S.status == 0);
    const int mneed = (S.n1 + S.n2 + 2 + (T - 3)) / (T - 2);
    // moving interface: uniform regrid (MRefin == 1); fixed interface (V_ip == 0: no regrid, any mesh that is a LinRange pair)
    const bool ok = ((S.V != 0.0 && S.MRefin == 1.0) || (S.V == 0.0 && A.dt_mode != DT_ADV)) && S.n1 >= 3 && S.n2 >= 3 && mneed <= MC &&
                    A.capL + 1 <= S2_MAX_D && A.capR + 1 <= S2_MAX_D;
    if (mine && !ok) { if (!A.only_deferred) A.status[mem] = -1; mine = false; }
    S.use_chain = mine ? 1 : 0;
    if (mine) { Lcur.m = mneed; Lcur.colR0 = (S.n1 + 1 + mneed - 1) / mneed; }
  }
  __syncthreads();
  if (!S.use_chain) return;
  {
    // the member's meshes must BE the closed form (bit for bit): LinRange(0, Ri1, n1), LinRange(Ri1, Ri2, n2)
    const int n1 = S.n1, n2 = S.n2;
    const double *gxl = A.x_left + (size_t)mem * A.capL, *gxr = A.x_right + (size_t)mem * A.capR;
    const Mesh1 ML = s2_mesh(0.0, S.Ri1, n1, false), MR = s2_mesh(S.Ri1, S.Ri2, n2, false);
    int differs = 0;
    for (int i = tid; i < n1; i += T) differs |= (gxl[i] != s2_at(ML, i));
    for (int i = tid; i < n2; i += T) differs |= (gxr[i] != s2_at(MR, i));
    if (__syncthreads_or(differs)) {
      // not a LinRange pair: the first streaming kernel (x in memory) takes the member
      if (tid == 0 && !A.only_deferred) A.status[mem] = -1;
      return;
    }
  }
  {
    const int n1 = S.n1, n2 = S.n2;
    const StreamLayout L = Lcur;
    const double *gcl = A.C_left + (size_t)mem * A.capL, *gcr = A.C_right + (size_t)mem * A.capR;
    const double *gxl = A.x_left + (size_t)mem * A.capL, *gxr = A.x_right + (size_t)mem * A.capR;
    for (int i = tid; i < n1; i += T) C[paddr(T, L, 0, i)] = gcl[i];
    for (int i = tid; i < n2; i += T) C[paddr(T, L, 1, n2 - 1 - i)] = gcr[i];
    double v[4] = {1e300, -1e300, 1e300, -1e300}, tm[2] = {1e300, -1e300};
    for (int i = tid; i < A.hist_len; i += T) {
      const double tt = P.t_ar[i], TT = P.T_ar[i], kd = P.KD_ar[i];
      tm[0] = fmin(tm[0], tt); tm[1] = fmax(tm[1], tt);
      v[2] = fmin(v[2], TT); v[3] = fmax(v[3], TT);
      v[0] = fmin(v[0], kd); v[1] = fmax(v[1], kd);
    }
    v[0] = warp_min(v[0]); v[1] = warp_max(v[1]); v[2] = warp_min(v[2]); v[3] = warp_max(v[3]);
    tm[0] = warp_min(tm[0]); tm[1] = warp_max(tm[1]);
    if (lane == 0) { redS[wid] = v[0]; redD[wid] = v[1]; redU[wid] = v[2]; red0[wid] = v[3]; red1[wid] = tm[0]; exA[wid] = tm[1]; }
    __syncthreads();
    if (tid == 0) {
      double a0 = 1e300, a1 = -1e300, a2 = 1e300, a3 = -1e300, a4 = 1e300, a5 = -1e300;
#pragma unroll 1
      for (int k = 0; k < T / 32; ++k) {
        a0 = fmin(a0, redS[k]); a1 = fmax(a1, redD[k]); a2 = fmin(a2, redU[k]); a3 = fmax(a3, red0[k]);
        a4 = fmin(a4, red1[k]); a5 = fmax(a5, exA[k]);
      }
      S.KDmin = a0; S.KDmax = a1; S.Tmin = a2; S.Tmax = a3; S.tmin = a4; S.tmax = a5;
      S.dx1 = xsub(gxl[n1 - 1], gxl[n1 - 2]);                                // create_grid!, functions.jl:366-367
      S.dx2 = xsub(gxr[1], gxr[0]);
      const int rc = update_t_dependent(S, A, P, false);
      if (rc) S.status = rc;
    }
  }
  __syncthreads();

  unsigned long long hash_local = 0ull;
  const bool do_hash = (A.idx_hash != nullptr);
  if (tid == 0) { s2_Kc = 0x3fffffff; s2_kz = 0; }        // first step: no cutoff (Kc is read after the scalar phase's barrier)
  const bool fixedV = (S.V == 0.0);            // fixed interface: advect_interface_regrid! / regrid! leave the meshes (:57-61, :900-903)
  HistCache hc;
  if (tid == 0) hist_load(hc, P, S.hidx);

  while (true) {
    if (S.status != 0 || !(S.t < S.t_tot) || (A.max_steps > 0 && S.steps >= A.max_steps)) break;
    __syncthreads();
    // ---------------- scalar phase (thread 0): dt, time, parameters, interface advection, new node counts, geometry -----
    if (tid == 0) {
      double dt;
      if (A.dt_mode == DT_ADV) dt = xmul(xdiv(fmin(S.dx1, S.dx2), S.V), S.CFL);                 // C1.jl:93
      else {                                                                                      // find_dt :465-490
        const double mn = fmin(S.dx1, S.dx2), mxD = fmax(S.D_l, S.D_r);
        double dtV = xmul(mn, xinv(fabs(S.V)));
        const double dtD = xmul(xmul(xmul(mn, mn), xinv(mxD)), 50.0);
        dtV = xmul(dtV, 0.99);
        dt = (dtV < dtD) ? xmul(dtV, S.CFL) : xmul(dtD, S.CFL);
      }
      S.dt = dt;
      S.t = xadd(S.t, S.dt);
      S.it += 1;
      if (S.t > S.t_tot) { S.dt = xsub(S.dt, xsub(S.t, S.t_tot)); S.t = S.t_tot; }
      double KD, Tt, Dl, Dr;
      int rc = chain_update_params(S, A, P, S.t, hc, &KD, &Tt, &Dl, &Dr);
      // S.Fl: what the member's meshes are if the step stops here - 0: the LinRanges of (S.Ri1, S.n1, S.n2); 1: the
      // interface moved too fast (:39 / :52: Ri1 is advanced, the arrays are untouched); 2: stopped after the advection
      // (capacity of the regrid: the arrays hold the inserted / moved nodes)
      S.Fl = 0;
      if (!rc) {
        S.KD = KD; S.T = Tt; S.D_l = Dl; S.D_r = Dr;
        const int n1 = S.n1, n2 = S.n2;
        // the meshes of the state: generated with the interface position before this step's advection
        G.MO[0] = s2_mesh(0.0, S.Ri1, n1, false);
        G.MO[1] = s2_mesh(S.Ri1, S.Ri2, n2, true);
        G.n[0] = n1; G.n[1] = n2;
        const double Rnew = xadd(S.Ri1, xmul(S.V, S.dt));                                        // functions.jl:31
        G.Rnew = Rnew;
        if (fixedV) { G.Rnew = S.Ri1; G.fast = 0; S.nr1 = n1; S.nr2 = n2; Lnew = Lcur; }         // both interface nodes ARE Ri1
        else if (S.V > 0.0) {
          if (n1 + 1 > A.capL) rc = 10;
          else {
            S.Ri1 = Rnew;
            if (Rnew > s2_lerp(G.MO[1], 1.0)) { rc = 1; S.Fl = 1; }                               // x_right[2], :37-40
            else S.n1 = n1 + 1;                       // :33-35, :41: new last left node, first right node moves (both = Rnew)
          }
        } else {
          if (n2 + 1 > A.capR) rc = 10;
          else {
            S.Ri1 = Rnew;
            if (Rnew < s2_lerp(G.MO[0], (double)(n1 - 2))) { rc = 2; S.Fl = 1; }                  // x_left[end-1], :50-53
            else S.n2 = n2 + 1;                       // :46-48, :54: new first right node, last left node moves
          }
        }
        G.n[0] = S.n1; G.n[1] = S.n2;
      }
      if (!rc && !fixedV) {                                                                       // regrid! :873-889
        long long nr1 = S.n1, nr2 = S.n2;
        if (S.V > 0.0) nr1 = (long long)rint(xmul(xmul(S.Ri1, xinv(xsub(S.Ri2, S.Ri1))), (double)nr2));
        else nr2 = (long long)rint(xmul(xmul(S.Ri2, xinv(xsub(S.Ri2, S.Ri1))), (double)nr1));
        if (nr1 < S.resmin1) nr1 = S.resmin1;
        if (nr2 < S.resmin2) nr2 = S.resmin2;
        if (nr1 > A.capL || nr2 > A.capR || nr1 < 3 || nr2 < 3) rc = 10;
        else {
          const int mneed = (int)((nr1 + nr2 + 2 + (T - 3)) / (T - 2));
          if (mneed > MC) rc = 10;
          else { Lnew.m = mneed; Lnew.colR0 = (int)((nr1 + 1 + mneed - 1) / mneed); S.nr1 = (int)nr1; S.nr2 = (int)nr2; }
        }
        if (rc) S.Fl = 2;
      }
      if (!rc && !fixedV) {
        G.nr[0] = S.nr1; G.nr[1] = S.nr2;
        G.MN[0] = s2_mesh(0.0, S.Ri1, S.nr1, false);
        G.MN[1] = s2_mesh(S.Ri1, S.Ri2, S.nr2, true);
        // fused pass: same layout, long columns, and both phases aligned with one offset (probed near both ends here,
        // verified node by node in the pass)
        int fast = (Lnew.m == Lcur.m && Lnew.colR0 == Lcur.colR0 && Lcur.m >= S2_MIN_ROWS) ? 1 : 0;
#ifdef MBM_S2_NO_FUSED
        fast = 0;
#endif
        for (int ph = 0; ph < 2 && fast; ++ph) {
          const int no = G.n[ph], nn = G.nr[ph];
          if (nn < 16 || no < 16 || nn - no > 2 || no - nn > 2) { fast = 0; break; }
          int oo[2];
          for (int q = 0; q < 2; ++q) {
            const int p = q ? (min(nn, no) - 5) : 3;
            const double X = s2_at(G.MN[ph], p);
            oo[q] = 99;
            for (int o = 0; o >= -1; --o) {
              const double xa = s2_xs(G, ph, p + o), xb = s2_xs(G, ph, p + o + 1);
              if (ph ? (xa > X && X > xb) : (xa < X && X < xb)) { oo[q] = o; break; }
            }
          }
          if (oo[0] != oo[1] || oo[0] == 99) fast = 0;
          G.o[ph] = oo[0];
        }
        G.fast = fast;
      }
      if (rc) S.status = rc; else S.node_steps += S.n1 + S.n2;
    }
    __syncthreads();
    if (S.status != 0) break;

    // ---------------- solve ------------------------------------------------------------------------------------------------
    const StreamLayout L = Lcur;
    const int n1 = S.n1, n2 = S.n2, m = L.m;
    const double idt = frcp_(S.dt);
    const int KL = (n1 - 1 + m - 1) / m, KR = (n2 - 1 + m - 1) / m;
    const int blk = (tid >= L.colR0) ? 1 : 0;
    const int jc = blk ? (tid - L.colR0) : tid;
    const int nph = blk ? n2 : n1;
    const bool active = blk ? (jc < KR) : (jc < KL);
    const int f = jc * m;
    const int len = active ? (min(f + m, nph - 1) - f) : 0;
    const bool lastChunk = active && (jc == (blk ? KR : KL) - 1);
    const double Dph = blk ? S.D_r : S.D_l;
    const bool dir0 = (blk ? (S.bc1 == 1) : (S.bc0 == 1)) && f == 0;
    const double Co_out = C[paddr(T, L, blk, 0)];
    double Ap = 0.0, Gp = 0.0, Rp = 0.0, P0 = 0.0, Qa = 0.0, Qc = 0.0;
    double bA = 0.0, bG = -1.0, bR = 0.0;
    const Mesh1 MO = G.MO[blk];
    const double Rnew = G.Rnew;
    // halo values must be read before any thread overwrites C in place
    double c_halo_prev = 0.0, c_halo_next = 0.0, x_halo_prev = 0.0, x_halo_next = 0.0;
    if (active) {
      if (f > 0) { x_halo_prev = s2_at(MO, f - 1); c_halo_prev = C[paddr(T, L, blk, f - 1)]; }
      x_halo_next = (f + len == nph - 1) ? Rnew : s2_at(MO, f + len);
      c_halo_next = C[paddr(T, L, blk, f + len)];
      // the node the advection inserts takes the value of its neighbour (functions.jl:34 / :47): it is the interface node
      // of the growing phase, i.e. the halo of that phase's last chunk (the state holds no value for it yet)
      if (lastChunk && ((S.V > 0.0 && !blk) || (S.V < 0.0 && blk))) c_halo_next = C[paddr(T, L, blk, f + len - 1)];
    }
    fence_proxy_async();
    __syncthreads();
    // rows k < Kc: (R, A, G) at 3 k T; rows k >= Kc: (R, G) at 3 Kc T + 2 (k - Kc) T
    int Kc = s2_Kc;
    if (A.flags & F_DEV_S2_NO_CUTOFF) Kc = 0x3fffffff;
    auto row_ptr = [&](const int k) { return RAG + ((k < Kc) ? (size_t)k * (3 * T) : (size_t)Kc * (3 * T) + (size_t)(k - Kc) * (2 * T)) + tid; };
    for (int attempt = 0;; ++attempt) {
      Ap = 0.0; Gp = 0.0; Rp = 0.0; P0 = 0.0; Qa = 0.0; Qc = 0.0;
      int kz = 0x3fffffff;                     // first row of this column whose fill-in is below S2_A_ZERO
      bool viol = false;                       // a fill-in value above S2_A_DROP was not stored
      // forward sweep: as in the first streaming kernel, x of every row from the closed form (fi: node index of the
      // NEXT row's node in its LinRange, stepped exactly by +-1)
      const int sd = blk ? -1 : 1;
      const double Dsd = (sd > 0) ? Dph : -Dph, idts = (sd > 0) ? idt : -idt;
      double x_self = s2_at(MO, f), c_self = C[saddr(T, tid, 0)];
      double fi = MO.fi0 + MO.s * (double)(f + 1);
      double dcar = 0.0, rcar = 0.0, sub = 0.0, hprev = 0.0, Vs = 0.0, Qb = 0.0, pi = 1.0;
      if (mb) Vs = powd<NDIM>(x_self);
      if (f > 0) {
        double Apv, A12, Aown, kk;
        elem_mats_dir<NDIM>(x_halo_prev, x_self, Dsd, idts, Apv, A12, Aown, kk);
        dcar = Aown - kk; rcar = fma(A12, c_halo_prev, Aown * c_self); sub = A12 + kk;
        if (mb) hprev = 0.5 * fabs(Vs - powd<NDIM>(x_halo_prev));
      }
      bA = 0.0; bR = 0.0;
      // one row of the sweep (k: row of the column), given the next node's x and C; rows go to pr[0] (R), pr[T] (A), pr[2T] (G).
      // INTERIOR rows (neither the first nor the last row of the column) carry no tests of k.
      auto fwd_row = [&](auto interior, auto with_a, const int k, double *pr, const double x_next, const double c_next) {
        constexpr bool INT = decltype(interior)::value, HASA = decltype(with_a)::value;
        const bool first = INT ? false : (k == 0), last = INT ? false : (k == len - 1);
        double Aself, A12, Anext, kk;
        elem_mats_dir<NDIM>(x_self, x_next, Dsd, idts, Aself, A12, Anext, kk);
        double diag = (Aself - kk) + dcar;
        double rhs = fma(Aself, c_self, A12 * c_next) + rcar;
        const double sup_true = A12 + kk;
        double sup = sup_true;
        if (first && dir0) { diag = 1.0; rhs = Co_out; sup = 0.0; sub = 0.0; }
        const double inv = frcp_(diag - sub * Gp);
        const double An = first ? (sub * inv) : (-sub * Ap * inv);
        const double Gn = sup * inv;
        const double Rn = (rhs - sub * Rp) * inv;
        if (HASA) { pr[0] = Rn; pr[T] = An; pr[2 * T] = Gn; if (fabs(An) < S2_A_ZERO) kz = min(kz, k); }
        else { pr[0] = Rn; pr[T] = Gn; viol |= !(fabs(An) < S2_A_DROP); }
        if (!last) { bR = fma(pi, Rn, bR); bA = fma(pi, An, bA); pi = -pi * Gn; }
        if (mb) {
          const double Vn = powd<NDIM>(x_next);
          const double hn = 0.5 * fabs(Vn - Vs);
          const double wgt = hn + hprev;
          if (!last) { const double tq = Qb + wgt; P0 = fma(tq, Rn, P0); Qa = fma(-tq, An, Qa); Qb = -tq * Gn; }
          else Qc = Qb + wgt;
          Vs = Vn; hprev = hn;
        }
        Ap = An; Gp = Gn; Rp = Rn;
        dcar = Anext - kk; rcar = fma(A12, c_self, Anext * c_next); sub = sup_true;
        x_self = x_next; c_self = c_next;
      };
      {
        const int nrows = m - 1, nbatch = (nrows + STG_R - 1) / STG_R;
        auto issue = [&](const int b) {       // thread 0: rows 1 + b*R .. of C into buffer b % NBF
          const int r0 = 1 + b * STG_R, nr = min(STG_R, nrows - b * STG_R), buf = b % STG_NBF;
          const unsigned bytes = (unsigned)nr * T * 8u;
          mbar_expect_tx(&fullb[buf], bytes);
          bulk_g2s(stg + (size_t)buf * (STG_R * T), C + (size_t)r0 * T, bytes, &fullb[buf]);
        };
        if (tid == 0) for (int b = 0; b < STG_NBF - 1 && b < nbatch; ++b) issue(b);
        for (int b = 0; b < nbatch; ++b) {
          const int buf = b % STG_NBF;
          if (tid == 0 && b + STG_NBF - 1 < nbatch) issue(b + STG_NBF - 1);
          mbar_wait(&fullb[buf], (stg_par >> buf) & 1u);
          stg_par ^= 1u << buf;
          const double *cs_ = stg + (size_t)buf * (STG_R * T) + tid;
          const int k0 = b * STG_R, nr = min(STG_R, nrows - k0);
          const bool hasA = (k0 < Kc);                       // Kc is a multiple of the batch length: whole batches
          double *const pr0 = row_ptr(k0);
          if (nr == STG_R && k0 >= 1 && k0 + STG_R <= len - 1) {
            // a batch of interior rows: straight-line code, immediate offsets
            if (hasA) {
#pragma unroll
              for (int rr = 0; rr < STG_R; ++rr) {
                const double xn_ = s2_lerp(MO, fi);
                fi += MO.s;
                fwd_row(std::true_type{}, std::true_type{}, k0 + rr, pr0 + rr * (3 * T), xn_, cs_[rr * T]);
              }
            } else {
#pragma unroll
              for (int rr = 0; rr < STG_R; ++rr) {
                const double xn_ = s2_lerp(MO, fi);
                fi += MO.s;
                fwd_row(std::true_type{}, std::false_type{}, k0 + rr, pr0 + rr * (2 * T), xn_, cs_[rr * T]);
              }
            }
          } else {
#pragma unroll 1
            for (int rr = 0; rr < nr; ++rr) {
              const int k = k0 + rr;
              if (k < len) {
                const bool last = (k == len - 1);
                const double xn_ = s2_lerp(MO, fi);
                fi += MO.s;
                if (hasA) fwd_row(std::false_type{}, std::true_type{}, k, pr0 + rr * (3 * T), last ? x_halo_next : xn_, last ? c_halo_next : cs_[rr * T]);
                else fwd_row(std::false_type{}, std::false_type{}, k, pr0 + rr * (2 * T), last ? x_halo_next : xn_, last ? c_halo_next : cs_[rr * T]);
              }
            }
          }
          __syncthreads();
        }
        if (len == m) {
          if (m - 1 < Kc) fwd_row(std::false_type{}, std::true_type{}, m - 1, row_ptr(m - 1), x_halo_next, c_halo_next);
          else fwd_row(std::false_type{}, std::false_type{}, m - 1, row_ptr(m - 1), x_halo_next, c_halo_next);
        }
      }
      bG = -pi;
      // cutoff bookkeeping: a dropped fill-in value that was not negligible repeats the sweep without cutoff (the state C is
      // untouched by the sweep); otherwise the cutoff of the next step follows where this step's columns became negligible
      if (__syncthreads_or(viol ? 1 : 0)) {
        Kc = 0x3fffffff;
        if (tid == 0) s2_Kc = Kc;
        continue;
      }
      if (tid == 0 && Kc < m) atomicAdd(&g_s2_cut_rows, (unsigned long long)(m - Kc));
      kz = min(kz, len);                       // columns shorter than their cutoff do not count
      for (int o_ = 16; o_ > 0; o_ >>= 1) kz = max(kz, __shfl_xor_sync(0xffffffffu, kz, o_));
      if (lane == 0) atomicMax(&s2_kz, kz);
      break;
    }
    fence_proxy_async();
    exA[tid] = bA; exG[tid] = bG; exR[tid] = bR;
    __syncthreads();
    if (active) {
      if (!lastChunk) {
        const double nA = exA[tid + 1], nG = exG[tid + 1], nR = exR[tid + 1];
        redS[tid] = Ap; redD[tid] = 1.0 - Gp * nA; redU[tid] = -Gp * nG; red0[tid] = Rp - Gp * nR; red1[tid] = 0.0;
      } else { redS[tid] = Ap; redD[tid] = 1.0; redU[tid] = 0.0; red0[tid] = Rp; red1[tid] = -Gp; }
    } else { redS[tid] = 0.0; redD[tid] = 1.0; redU[tid] = 0.0; red0[tid] = 0.0; red1[tid] = 0.0; }
    __syncthreads();
    if (tid < 32) reduced_solve<T>(redS, redD, redU, red0, red1);
    __syncthreads();
    // interface condition (nodes next to the interface: phase-ordered n-2)
    double aif, bif;
    {
      const double xp_ = Rnew, xq_ = Rnew;
      const double xpm = s2_at(G.MO[0], n1 - 2), xqp = s2_at(G.MO[1], n2 - 2);
      const int rL = KL - 1, rR = L.colR0 + KR - 1;
      if (bc == BC_FLUX) {
        const double idxl = frcp_(xp_ - xpm), idxr = frcp_(xqp - xq_);
        const double al = -S.rho1 * S.D_l * idxl, be = S.V * S.rho1 + S.rho1 * S.D_l * idxl;
        const double ga = -S.V * S.rho2 + S.rho2 * S.D_r * idxr, de = -S.rho2 * S.D_r * idxr;
        const double ca = be + al * red1[rL], cb = ga + de * red1[rR];
        bif = -(al * red0[rL] + de * red0[rR]) / (ca * S.KD + cb);
        aif = S.KD * bif;
      } else {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (active) {
          const double s0v = red0[tid], s1v = red1[tid];
          const double q0 = (jc > 0) ? red0[tid - 1] : 0.0, q1 = (jc > 0) ? red1[tid - 1] : 0.0;
          const double c0 = P0 + Qa * q0 + Qc * s0v, c1 = Qa * q1 + Qc * s1v;
          if (blk == 0) { v[0] = c0; v[1] = c1; } else { v[2] = c0; v[3] = c1; }
        }
        block_sum4<T>(v, S.red);
        const double Vp_ = powd<NDIM>(xp_), Vq_ = powd<NDIM>(xq_);
        const double wp = 0.5 * (Vq_ - Vp_) + 0.5 * (Vp_ - powd<NDIM>(xpm));
        const double wq = 0.5 * (powd<NDIM>(xqp) - Vq_) + 0.5 * (Vq_ - Vp_);
        bif = (S.Mass0 - (v[0] + v[2])) / ((v[1] + wp) * S.KD + (v[3] + wq));
        aif = S.KD * bif;
      }
    }
    __syncthreads();                            // every thread is past its reads of rows 0..7 of sh (reduced system)

    const StreamLayout LN = Lnew;
    const int nr1 = S.nr1, nr2 = S.nr2;
    const long long it_ = S.it;
    const double beta = blk ? bif : aif;
    const double sj = active ? (red0[tid] + red1[tid] * beta) : 0.0;
    const double spv = (active && jc > 0) ? (red0[tid - 1] + red1[tid - 1] * beta) : 0.0;
    // rows m-2 .. 0 of the elimination staged through shared memory, batch j = rows [j R, j R + R - 1] (the boundaries of
    // the forward sweep's batches, so a batch lies on one side of the cutoff), taken from the top batch down: the b-th
    // batch issued / consumed is j = nbatch_b - 1 - b.  Both variants of the second pass read the rows this way.
    const int nrows_b = m - 1, nbatch_b = (nrows_b + STG_R - 1) / STG_R;
    auto issue_b = [&](const int b) {
      const int j = nbatch_b - 1 - b, lo = j * STG_R, hi = min(lo + STG_R - 1, m - 2), nr = hi - lo + 1, buf = b % STG_NBB;
      const unsigned bytes = (unsigned)nr * ((lo < Kc) ? 3u : 2u) * T * 8u;
      mbar_expect_tx(&fullb[buf], bytes);
      bulk_g2s(stg + (size_t)buf * (3 * STG_R * T), row_ptr(lo) - tid, bytes, &fullb[buf]);
    };

    bool fused_done = false;
    if (G.fast) {
      // ================= fused pass: back-substitution + regrid =================
      const int no = nph, nn = blk ? nr2 : nr1, o = G.o[blk];
      const Mesh1 MN = G.MN[blk];
      const int cntN = max(0, min(f + m, nn) - f);
      const int mcl = min(cntN, len);
      const int pa = f + 2, pb = f + mcl - 2;                // the walk evaluates new nodes [pa, pb)
      unsigned long long hstep = 0ull;
      bool ok = true;
      double cn = sj;
      double x1 = 0.0, x2 = 0.0, y1 = sj, ih12 = 1.0, M12 = 0.0, D2 = 0.0;
      double fio = 0.0, fin = 0.0;                           // LinRange indices of the next row's old node / new node
      if (active) {
        const int k = len - 1;
        if (k < 4) sh[k][tid] = cn;
        sh[4 + (k - (len - 4))][tid] = cn;                   // k >= len - 4
        x1 = s2_at(MO, f + len - 1);
        x2 = (f + len == no - 1) ? Rnew : s2_at(MO, f + len);
        ih12 = frcp_(x2 - x1);
        fio = MO.fi0 + MO.s * (double)(f + len - 2);
        fin = MN.fi0 + MN.s * (double)(f + len - 1 - o);     // new node of row len - 2
      }
      // one row (k) of the walk: back-substitute old node q = f + k, complete the stencil of old interval [q + 1, q + 2]
      // and evaluate its new node pn = q + 1 - o into *pc.  rp: the row's (R, A, G) in the staging buffer.  INTERIOR rows
      // (4 <= k <= kI_hi) publish nothing and always evaluate.
      auto frow = [&](auto interior, auto with_a, const int k, const double *rp, double *pc) {
        constexpr bool INT = decltype(interior)::value, HASA = decltype(with_a)::value;
        if (HASA) cn = rp[0] - rp[T] * spv - rp[2 * T] * cn;  // solved value of old node q
        else cn = rp[0] - rp[T] * cn;                         // (beyond the cutoff: (R, G) rows)
        if (!INT) {
          if (k < 4) sh[k][tid] = cn;
          if (k >= len - 4) sh[4 + (k - (len - 4))][tid] = cn;
        }
        const double x0 = s2_lerp(MO, fio);
        fio -= MO.s;
        const double Xj = s2_lerp(MN, fin);
        fin -= MN.s;
        const double ih01 = frcp_(x1 - x0), M01 = (y1 - cn) * ih01;
        const double D1 = slope_interior(ih01, ih12, M01, M12);          // slope at old node q + 1
        const int pn = f + k + 1 - o;
        if (INT || (pn >= pa && pn < pb)) {
          const double sx = Xj - x1, dx2_ = x2 - Xj;
          ok &= (sx != 0.0) & (dx2_ != 0.0) & ((sx > 0.0) == (dx2_ > 0.0)) & ((sx > 0.0) != (blk != 0));
          const double u = sx * ih12, dd = D1 + D2;
          const double c3 = fma(-2.0, M12, dd), c2 = fma(3.0, M12, -dd) - D1;
          *pc = fma(sx, fma(u, fma(u, c3, c2), D1), y1);
          if (do_hash) hstep += idx_term(it_, blk, blk ? (nn - 1 - pn) : pn, blk ? (no - 2 - (pn + o)) : (pn + o), 3);
        }
        x2 = x1; x1 = x0; y1 = cn; ih12 = ih01; M12 = M01; D2 = D1;
      };
      const int kI_hi = mcl - 5;                             // with mcl <= len: not among the last four rows, pn < pb
      if (tid == 0) for (int b = 0; b < STG_NBB - 1 && b < nbatch_b; ++b) issue_b(b);
      for (int b = 0; b < nbatch_b; ++b) {
        const int buf = b % STG_NBB;
        if (tid == 0 && b + STG_NBB - 1 < nbatch_b) issue_b(b + STG_NBB - 1);
        mbar_wait(&fullb[buf], (stg_par >> buf) & 1u);
        stg_par ^= 1u << buf;
        const int lo = (nbatch_b - 1 - b) * STG_R, hi = min(lo + STG_R - 1, m - 2);
        const double *const rp0 = stg + (size_t)buf * (3 * STG_R * T) + tid;            // row lo
        double *const pc0 = Cn + (size_t)(lo + 1 - o) * T + tid;                         // new node of row lo
        const bool hasA = (lo < Kc);
        if (hi - lo + 1 == STG_R && lo >= 4 && hi <= kI_hi) {
          if (hasA) {
#pragma unroll
            for (int q = STG_R - 1; q >= 0; --q) frow(std::true_type{}, std::true_type{}, lo + q, rp0 + q * (3 * T), pc0 + q * T);
          } else {
#pragma unroll
            for (int q = STG_R - 1; q >= 0; --q) frow(std::true_type{}, std::false_type{}, lo + q, rp0 + q * (2 * T), pc0 + q * T);
          }
        } else {
#pragma unroll 1
          for (int k = hi; k >= lo; --k)
            if (k <= len - 2) {
              if (hasA) frow(std::false_type{}, std::true_type{}, k, rp0 + (k - lo) * (3 * T), pc0 + (k - lo) * T);
              else frow(std::false_type{}, std::false_type{}, k, rp0 + (k - lo) * (2 * T), pc0 + (k - lo) * T);
            }
        }
        __syncthreads();
      }
      // ---- the nodes at the column ends (and whatever a short last column leaves): general evaluation in the
      //      reference's node order on the published solved values
      bool bad = false;
      {
        auto cval = [&](const int p) -> double {             // solved old value at phase-ordered index p of this phase
          if (p == no - 1) return beta;
          const int jq = p / m, row = p - jq * m;
          const int col = (blk ? L.colR0 : 0) + jq;
          const int lenc = min((jq + 1) * m, no - 1) - jq * m;
          if (row < 4) return sh[row][col];
          if (row >= lenc - 4) return sh[4 + (row - (lenc - 4))][col];
          bad = true;
          return 0.0;
        };
        auto xf = [&](const int i) { return s2_xs(G, blk, blk ? (no - 1 - i) : i); };
        auto yf = [&](const int i) { return cval(blk ? (no - 1 - i) : i); };
        const int nfix = cntN - max(0, pb - pa);
        if (nfix > S2_MAX_FIX) bad = true;
        else {
          for (int q = 0; q < cntN; ++q) {
            const int pn = f + q;
            if (pn >= pa && pn < pb) { q = pb - f - 1; continue; }
            const double Xj = s2_at(MN, pn);
            const int kph = pn + o;                                          // aligned old interval, phase order
            const int c0 = blk ? (no - 1 - kph) : (kph + 1);
            int kk = 0, br = 0;
            const double val = s2_eval_general(xf, yf, no, Xj, c0, &kk, &br, &bad);
            Cn[(size_t)q * T + tid] = val;
            if (do_hash) hstep += idx_term(it_, blk, blk ? (nn - 1 - pn) : pn, kk, br);
          }
        }
      }
      if (!__syncthreads_or((bad || !ok) ? 1 : 0)) {
        fused_done = true;
        hash_local += hstep;
        if (tid == 0) atomicAdd(&g_s2_fused_steps, 1ull);
      }
    }
    if (!fused_done) {
      // ================= two-pass path: back-substitution, then the interval-driven walk =================
      {
        double cn = sj;
        if (active) C[saddr(T, tid, len - 1)] = cn;
        if (tid == 0) for (int b = 0; b < STG_NBB - 1 && b < nbatch_b; ++b) issue_b(b);
        for (int b = 0; b < nbatch_b; ++b) {
          const int buf = b % STG_NBB;
          if (tid == 0 && b + STG_NBB - 1 < nbatch_b) issue_b(b + STG_NBB - 1);
          mbar_wait(&fullb[buf], (stg_par >> buf) & 1u);
          stg_par ^= 1u << buf;
          const int lo = (nbatch_b - 1 - b) * STG_R, hi = min(lo + STG_R - 1, m - 2);
          const double *const rp0 = stg + (size_t)buf * (3 * STG_R * T) + tid;
          const bool hasA = (lo < Kc);
#pragma unroll 1
          for (int k = hi; k >= lo; --k) {
            if (k <= len - 2) {
              const double *rp = rp0 + (k - lo) * (hasA ? 3 * T : 2 * T);
              cn = hasA ? (rp[0] - rp[T] * spv - rp[2 * T] * cn) : (rp[0] - rp[T] * cn);
              C[(size_t)k * T + tid] = cn;
            }
          }
          __syncthreads();
        }
      }
      if (tid == 0) { C[paddr(T, L, 0, n1 - 1)] = aif; C[paddr(T, L, 1, n2 - 1)] = bif; }
      __threadfence_block();
      __syncthreads();
      if (!fixedV) {
        const int nblk = (tid >= LN.colR0) ? 1 : 0;
        const int ncol = nblk ? (tid - LN.colR0) : tid;
        const int nn = nblk ? nr2 : nr1;
        const int no = nblk ? n2 : n1;
        const int idx0 = ncol * LN.m;
        const int cnt = max(0, min(idx0 + LN.m, nn) - idx0);
        if (cnt > 0) {
          const Mesh1 MNq = G.MN[nblk];
          auto xo = [&](const int p) { return s2_xs(G, nblk, p); };
          auto xn = [&](const int p) { return s2_at(MNq, p); };
          if (do_hash) pchip_run_stream2<T, true>(xo, xn, C, L, nblk, no, nn, Cn, idx0, cnt, it_, hash_local);
          else pchip_run_stream2<T, false>(xo, xn, C, L, nblk, no, nn, Cn, idx0, cnt, 0, hash_local);
        }
      }
      if (tid == 0) atomicAdd(&g_s2_twopass_steps, 1ull);          // (fixed-interface steps are counted here too)
    }
    __threadfence_block();
    __syncthreads();
    if (!fixedV) { double *tp = C; C = Cn; Cn = tp; }
    if (tid == 0) {
      {                                            // cutoff of the next step: past the last column that was still above zero
        int kc = ((s2_kz + 8 + STG_R - 1) / STG_R) * STG_R;
        if (kc + STG_R >= m - 1) kc = 0x3fffffff;
        s2_Kc = kc; s2_kz = 0;
      }
      if (!fixedV) {
        S.n1 = S.nr1; S.n2 = S.nr2;
        Lcur = Lnew;
        S.dx1 = xsub(linrange_at(0.0, S.Ri1, S.n1, S.n1 - 1), linrange_at(0.0, S.Ri1, S.n1, S.n1 - 2));   // regrid! :893-894
        S.dx2 = xsub(linrange_at(S.Ri1, S.Ri2, S.n2, 1), linrange_at(S.Ri1, S.Ri2, S.n2, 0));
      }
      S.steps += 1;
      const double cf = C[paddr(T, Lnew, 0, 0)], cl = C[paddr(T, Lnew, 1, 0)], ci = C[paddr(T, Lnew, 0, S.n1 - 1)];
      if (!isfinite(cf) || !isfinite(cl) || !isfinite(ci) || !isfinite(S.Ri1)) S.status = 11;
    }
    __syncthreads();
  }
  __syncthreads();
  // ---------------- store member ----------------
  {
    const StreamLayout L = Lcur;
    const int n1 = S.n1, n2 = S.n2;
    double *gxl = A.x_left + (size_t)mem * A.capL, *gcl = A.C_left + (size_t)mem * A.capL;
    double *gxr = A.x_right + (size_t)mem * A.capR, *gcr = A.C_right + (size_t)mem * A.capR;
    // the member's meshes and profiles in the reference's node order.  Fl = 1 (too fast, :39 / :52): Ri1 is advanced, the
    // arrays are those of the last completed step; Fl = 2 (stopped after the advection): the solve's meshes, the inserted
    // node carrying its neighbour's value (:34 / :47)
    const int Fl = S.Fl;
    const double RiM = (Fl == 1) ? G.MO[0].b : S.Ri1;
    const Mesh1 ML = s2_mesh(0.0, RiM, n1, false), MR = s2_mesh(RiM, S.Ri2, n2, false);
    auto xl = [&](const int i) { return (Fl == 2) ? s2_xs(G, 0, i) : s2_at(ML, i); };
    auto xr = [&](const int i) { return (Fl == 2) ? s2_xs(G, 1, n2 - 1 - i) : s2_at(MR, i); };
    auto cl = [&](const int i) { return C[paddr(T, L, 0, (Fl == 2 && S.V > 0.0 && i == n1 - 1) ? (n1 - 2) : i)]; };
    auto cr = [&](const int i) { return C[paddr(T, L, 1, (Fl == 2 && S.V < 0.0 && i == 0) ? (n2 - 2) : (n2 - 1 - i))]; };
    for (int i = tid; i < n1 && i < A.capL; i += T) { gxl[i] = xl(i); gcl[i] = cl(i); }
    for (int i = tid; i < n2 && i < A.capR; i += T) { gxr[i] = xr(i); gcr[i] = cr(i); }
    if ((A.flags & F_MASS_FINAL) && A.Mass) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n1 - 1; i += T)
        v[0] += 0.5 * (S.rho1 * powd<NDIM>(xl(i + 1)) - S.rho1 * powd<NDIM>(xl(i))) * (cl(i) + cl(i + 1));
      for (int i = tid; i < n2 - 1; i += T)
        v[1] += 0.5 * (S.rho2 * powd<NDIM>(xr(i + 1)) - S.rho2 * powd<NDIM>(xr(i))) * (cr(i) + cr(i + 1));
      block_sum4<T>(v, S.red);
      if (tid == 0) A.Mass[mem] = v[0] + v[1];
    }
    if (A.idx_hash) {
      unsigned long long h = hash_local;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
      if (lane == 0) atomicAdd(&S.hash, h);
      __syncthreads();
    }
    if (tid == 0) {
      A.status[mem] = S.status;
      if (A.steps_done) A.steps_done[mem] = S.steps;
      if (A.warn) A.warn[mem] = S.warn;
      A.t[mem] = S.t; A.dt[mem] = S.dt; A.it[mem] = S.it;
      A.res[2 * (size_t)mem] = n1; A.res[2 * (size_t)mem + 1] = n2;
      A.Ri[2 * (size_t)mem] = S.Ri1; A.Ri[2 * (size_t)mem + 1] = S.Ri2;
      if (A.dx1) A.dx1[mem] = S.dx1;
      if (A.dx2) A.dx2[mem] = S.dx2;
      if (A.newton_R) A.newton_R[mem] = S.newton_R;
      if (A.last) {
        double *Lq = A.last + 8 * (size_t)mem;
        Lq[0] = S.D_l; Lq[1] = S.D_r; Lq[2] = S.KD; Lq[3] = S.T; Lq[4] = S.V; Lq[5] = S.rho1; Lq[6] = S.rho2; Lq[7] = 1.0;
      }
      if (A.idx_hash) A.idx_hash[mem] = S.hash;
      if (A.node_steps) A.node_steps[mem] = S.node_steps;
    }
  }
}

}  // namespace mbm
