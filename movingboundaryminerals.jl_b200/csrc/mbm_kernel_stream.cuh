// mbm_kernel_stream.cuh — global-memory ("streaming") kernel for long profiles: BASELINE.json configs[4]
// (n = 65,536 nodes per phase, MRefin = 1: the only mesh the reference can generate at that size).
// Covers prescribed-velocity couples (flux / mb, V_ip != 0) on uniform meshes; everything else that does not fit
// shared memory still runs in the general kernel on a node-major global workspace.
//
// One CTA per member, but the member's arrays live in global memory in a CHUNK-INTERLEAVED layout: thread t owns
// column t (m consecutive nodes of one phase, ordered towards the interface) and element k of column t is stored at
// k*T + t, so that every sweep of the solve and of the regrid reads and writes fully coalesced 8-byte streams:
//   forward sweep   read x, C           write A, G, R (R in place of C)
//   backward accumulation / back-substitution   read A, G, R   write C
//   regrid          read x, C (walking window)   write x_new, C_new (in the next step's layout)
// Both phases look the same in this layout: the left phase is stored by node index, the right phase by REVERSED node
// index, so "towards the interface" is "k increasing" for both and the node the interface advection inserts
// (functions.jl:33-35 / :46-48) is an append in either case.  A CTA's row k of an array is one contiguous 2 KB block:
// the forward sweep and the back-substitution stream their rows through shared memory with 1-D bulk copies
// (cp.async.bulk + mbarrier, batches of STG_R rows issued one batch ahead by thread 0), the regrid walk reads through
// L1 with a software prefetch.  The roofline that bounds this kernel is HBM: algorithmic 32 B per node-step
// (SURVEY.md §8(d)); this implementation moves 104 B per node-step.
#pragma once
#include "mbm_kernel_fixed.cuh"

namespace mbm {

#ifndef MBM_SPF
#define MBM_SPF 4
#endif
#ifndef MBM_STREAM_MINB
#define MBM_STREAM_MINB 2
#endif
constexpr int SPF = MBM_SPF;   // software prefetch distance of the regrid walk, in rows (one row of a CTA = 2 KB per array)

// ---- bulk-copy (TMA) row staging: a CTA's row k of an array (element k of all T columns) is one contiguous 2 KB
// block, so the column sweeps stream their rows through shared memory with 1-D bulk copies (cp.async.bulk, completion on
// an mbarrier) instead of per-thread loads: batches of STG_R rows, STG_NBF / STG_NBB batches resident for the forward
// sweep (2 arrays) / the back-substitution (3 arrays), issued by thread 0 one batch ahead of the one being consumed.
#ifndef MBM_STG_R
#define MBM_STG_R 6
#endif
#ifndef MBM_STG_NBF
#define MBM_STG_NBF 3
#endif
#ifndef MBM_STG_NBB
#define MBM_STG_NBB 2
#endif
constexpr int STG_R = MBM_STG_R, STG_NBF = MBM_STG_NBF, STG_NBB = MBM_STG_NBB;
constexpr int STG_BYTES_F = STG_R * STG_NBF * 2 * 256 * 8, STG_BYTES_B = STG_R * STG_NBB * 3 * 256 * 8;
constexpr int STG_BYTES = STG_BYTES_F > STG_BYTES_B ? STG_BYTES_F : STG_BYTES_B;   // one allocation for both passes (86016 by default)
static_assert(STG_NBF <= 8 && STG_NBB <= STG_NBF, "staging barriers: fullb[STG_NBF]");
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct StreamLayout {
  int m;        // nodes per column
  int colR0;    // first column of the right phase
};
__device__ __forceinline__ size_t saddr(int T, int col, int k) { return (size_t)k * T + col; }
// phase-ordered index (left: node index; right: reversed node index) -> address
__device__ __forceinline__ size_t paddr(int T, const StreamLayout &L, int phase, int idx) {
  const int col = (phase ? L.colR0 : 0) + idx / L.m;
  return (size_t)(idx % L.m) * T + col;
}

// pchip slope at phase-ordered index i of a phase with n nodes stored in layout L (functions.jl:786-806).  The slope
// formulas are symmetric under reversal of the node order except for the sign of the secants, which cancels.
template <int T>
__device__ __forceinline__ double stream_slope(const double *X, const double *C, const StreamLayout &L, int phase, int n,
                                               int i) {
  if (i == 0 || i == n - 1) {
    const int a = (i == 0) ? 0 : n - 1, b = (i == 0) ? 1 : n - 2, c = (i == 0) ? 2 : n - 3;
    const double xa = X[paddr(T, L, phase, a)], xb = X[paddr(T, L, phase, b)], xc = X[paddr(T, L, phase, c)];
    const double ya = C[paddr(T, L, phase, a)], yb = C[paddr(T, L, phase, b)], yc = C[paddr(T, L, phase, c)];
    // end formula on (h1, h2, m1, m2) measured from the end node inwards; orientation-independent
    const double h1 = fabs(xb - xa), h2 = fabs(xc - xb);
    const double sg = (xb > xa) ? 1.0 : -1.0;
    const double m1 = sg * (yb - ya) * (1.0 / h1), m2 = sg * (yc - yb) * (1.0 / h2);
    double d = fma(fma(2.0, h1, h2), m1, -(h1 * m2)) * (1.0 / (h1 + h2));
    if (m1 * m2 < 0.0 || m1 == 0.0 || m2 == 0.0) d = 0.0;
    return d;
  }
  const double xm = X[paddr(T, L, phase, i - 1)], x0 = X[paddr(T, L, phase, i)], xp = X[paddr(T, L, phase, i + 1)];
  const double ym = C[paddr(T, L, phase, i - 1)], y0 = C[paddr(T, L, phase, i)], yp = C[paddr(T, L, phase, i + 1)];
  const double ihm = frcp_(x0 - xm), ihp = frcp_(xp - x0);           // both negative for the reversed phase: cancels
  return slope_interior(fabs(ihm), fabs(ihp), (y0 - ym) * ihm, (yp - y0) * ihp);
}

// ---- regrid of one new column: the tuned kernel's interval-driven pchip walk (pchip_run, mbm_kernel_v2.cuh) over the
// chunk-interleaved layout.  Old nodes are visited in phase order; thread t's new nodes fall into old column ~t, so
// the walk of all threads is one coalesced stream with incremental addressing (no index divisions on the way).
// Right phase (stored by reversed node index): the walk runs in mirrored coordinates x' = -x, which are
// increasing in phase order; the interpolant is invariant and indices / branches are mapped back for the checksum.
template <int T>
__device__ __forceinline__ PPt spt(const double *X, const double *C, const StreamLayout &L, int phase, double sg, int i) {
  const size_t a = paddr(T, L, phase, i);
  PPt q; q.x = sg * X[a]; q.y = C[a];
  return q;
}
// three-point end slope (functions.jl:786-806) at end node Pe with inward neighbours Pa, Pb (x increasing or decreasing)
__device__ __noinline__ double end_slope3(PPt Pe, PPt Pa, PPt Pb) {
  const double h1 = fabs(Pa.x - Pe.x), h2 = fabs(Pb.x - Pa.x);
  const double sg = (Pa.x > Pe.x) ? 1.0 : -1.0;
  const double m1 = sg * (Pa.y - Pe.y) * (1.0 / h1), m2 = sg * (Pb.y - Pa.y) * (1.0 / h2);
  double d = fma(fma(2.0, h1, h2), m1, -(h1 * m2)) * (1.0 / (h1 + h2));
  if (m1 * m2 < 0.0 || m1 == 0.0 || m2 == 0.0) d = 0.0;
  return d;
}

template <int T, bool do_hash>
__device__ __forceinline__ void pchip_run_stream(const double *X, const double *C, const StreamLayout &L, const int phase,
                                                 const int n, const int nn, const double lr_a, const double lr_b,
                                                 double *Xn, double *Cn, const int idx0, const int cnt,
                                                 const long long it, unsigned long long &hash) {
  const int tid = threadIdx.x;
  const bool MIR = (phase != 0);                      // runtime: both phases run one instruction stream
  const double sg = MIR ? -1.0 : 1.0;
  auto newx = [&](int pi) { return linrange_at_inl(lr_a, lr_b, nn, MIR ? (nn - 1 - pi) : pi); };     // reference node order
  auto emit = [&](int k, double xj, double val, int kk, int br) {
    const size_t an = saddr(T, tid, k);
    Xn[an] = xj; Cn[an] = val;
    if (do_hash) {
      if (MIR) {                                      // back to the reference's (forward) node order and branch names
        if (br == 0) kk = n - 1 - kk;
        else if (br == 1) { kk = 0; br = 2; }
        else if (br == 2) { kk = n - 1; br = 1; }
        else kk = n - 2 - kk;
      }
      hash += idx_term(it, phase, MIR ? (nn - 1 - (idx0 + k)) : (idx0 + k), kk, br);
    }
  };
  double xj = newx(idx0);
  double Xj = sg * xj;
  // c = number of old nodes (phase order) strictly below the first new node, by walking from the proportional guess
  int c = min(max((int)(((long long)idx0 * n) / nn), 1), n - 1);
  {
    double xa = spt<T>(X, C, L, phase, sg, c - 1).x, xb0 = spt<T>(X, C, L, phase, sg, c).x;
    int steps = 0;
    while (xb0 < Xj && c < n) {
      ++c; xa = xb0; xb0 = (c < n) ? spt<T>(X, C, L, phase, sg, c).x : xb0;
      if (++steps > 64) break;
    }
    while (!(xa < Xj) && c > 0 && steps <= 64) {
      --c; xa = (c > 0) ? spt<T>(X, C, L, phase, sg, c - 1).x : xa;
      ++steps;
    }
    if (steps > 64) {
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (spt<T>(X, C, L, phase, sg, mid).x < Xj) lo = mid + 1; else hi = mid; }
      c = lo;
    }
  }
  int j = 0;
  while (c == 0) {                                    // at or below the first old node: branch 0 / 2
    const PPt q0 = spt<T>(X, C, L, phase, sg, 0);
    if (!(Xj <= q0.x)) { c = 1; break; }
    emit(j, xj, q0.y, 0, (Xj == q0.x) ? 0 : 2);
    if (++j == cnt) return;
    xj = newx(idx0 + j); Xj = sg * xj;
  }
  int k = c - 1;                                      // current old interval [k, k+1]
  if (k <= n - 2) {
    PPt Q0 = spt<T>(X, C, L, phase, sg, k), Q1 = spt<T>(X, C, L, phase, sg, k + 1), Q2 = Q1, Q3 = Q1;
    double IHa = frcp_(Q1.x - Q0.x), Ma = (Q1.y - Q0.y) * IHa, IHb = 0.0, Mb = 0.0, Da, Db;
    if (k == 0) Da = end_slope3(Q0, Q1, spt<T>(X, C, L, phase, sg, 2));
    else { const PPt Qm = spt<T>(X, C, L, phase, sg, k - 1); const double ihm = frcp_(Q0.x - Qm.x); Da = slope_interior(ihm, IHa, (Q0.y - Qm.y) * ihm, Ma); }
    if (k + 1 == n - 1) Db = end_slope3(Q1, Q0, spt<T>(X, C, L, phase, sg, n - 3));
    else { Q2 = spt<T>(X, C, L, phase, sg, k + 2); IHb = frcp_(Q2.x - Q1.x); Mb = (Q2.y - Q1.y) * IHb; Db = slope_interior(IHa, IHb, Ma, Mb); }
    // sequential reader positioned at old index k + 3
    int wrow = 0; size_t wa = 0;
    if (k + 3 <= n - 1) { const int i3 = k + 3; wrow = i3 % L.m; wa = (size_t)wrow * T + (size_t)((phase ? L.colR0 : 0) + i3 / L.m); }
#define MBM_SSTEP(P0, P1, P2, P3, IH0, M0, D0, IH1, M1, D1)                                                   \
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
        emit(j, xj, val, eq ? k + 1 : k, eq ? 0 : 3);                                                          \
        if (++j == cnt) return;                                                                                \
        xj = newx(idx0 + j); Xj = sg * xj;                                                              \
      }                                                                                                        \
      if (k + 3 <= n - 1) {                                                                                    \
        if (wrow + SPF < L.m) {                                                                                \
          asm volatile("prefetch.global.L1 [%0];" ::"l"(X + wa + (size_t)SPF * T));                            \
          asm volatile("prefetch.global.L1 [%0];" ::"l"(C + wa + (size_t)SPF * T));                            \
        }                                                                                                      \
        P3.x = sg * X[wa]; P3.y = C[wa];                                                             \
        if (++wrow == L.m) { wrow = 0; wa = wa + 1 - (size_t)(L.m - 1) * T; } else wa += T;                    \
        IH0 = frcp_(P3.x - P2.x); M0 = (P3.y - P2.y) * IH0;                                                    \
        D0 = slope_interior(IH1, IH0, M1, M0);                                                                 \
      } else if (k + 3 == n) D0 = end_slope3(P2, P1, P0);                                                      \
      else break;                                                                                              \
      ++k;                                                                                                     \
    }
    for (;;) {
      MBM_SSTEP(Q0, Q1, Q2, Q3, IHa, Ma, Da, IHb, Mb, Db)
      MBM_SSTEP(Q1, Q2, Q3, Q0, IHb, Mb, Db, IHa, Ma, Da)
      MBM_SSTEP(Q2, Q3, Q0, Q1, IHa, Ma, Da, IHb, Mb, Db)
      MBM_SSTEP(Q3, Q0, Q1, Q2, IHb, Mb, Db, IHa, Ma, Da)
    }
#undef MBM_SSTEP
  }
  {                                                   // above the last old node: branch 1
    const double yl = spt<T>(X, C, L, phase, sg, n - 1).y;
    for (;;) {
      emit(j, xj, yl, n - 1, 1);
      if (++j == cnt) return;
      xj = newx(idx0 + j);
    }
  }
}

template <int T, int NDIM>
__global__ void __launch_bounds__(T, MBM_STREAM_MINB) mbm_stream_kernel(const KArgs A, const int member0) {
  __shared__ Scal S;
  __shared__ double redS[T], redD[T], redU[T], red0[T], red1[T];
  __shared__ double exA[T], exG[T], exR[T];
  __shared__ StreamLayout Lcur, Lnew;
  extern __shared__ __align__(128) double stg[];                 // STG_BYTES of row staging
  __shared__ __align__(8) unsigned long long fullb[STG_NBF];
  unsigned stg_par = 0;                                          // parity bit of every staging barrier (uniform)

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int mem = member0 + blockIdx.x;
  if (mem >= A.M) return;
  if (tid == 0) {
    for (int b = 0; b < STG_NBF; ++b) mbar_init(&fullb[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int MC = A.kcapL;                     // rows per column capacity of the workspace
  const size_t ASZ = (size_t)MC * T;
  double *base = A.workspace + (size_t)blockIdx.x * 6 * ASZ;
  double *X = base, *C = base + ASZ, *FA = base + 2 * ASZ, *FG = base + 3 * ASZ, *Xn = base + 4 * ASZ, *Cn = base + 5 * ASZ;
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
    // columns: m nodes each; both phases plus one appended node must fit T columns of MC rows
    const int mneed = (S.n1 + S.n2 + 2 + (T - 3)) / (T - 2);
    const bool ok = (S.V != 0.0) && (S.MRefin == 1.0) && S.n1 >= 3 && S.n2 >= 3 && mneed <= MC;
    if (mine && !ok) { if (!A.only_deferred) A.status[mem] = -1; mine = false; }
    if (mine) S.status = 0;
    S.use_chain = mine ? 1 : 0;
    if (mine && A.only_deferred && A.steps_done) S.steps = A.steps_done[mem];
    if (mine) { Lcur.m = mneed; Lcur.colR0 = (S.n1 + 1 + mneed - 1) / mneed; }
  }
  __syncthreads();
  if (!S.use_chain) return;
  {
    const int n1 = S.n1, n2 = S.n2;
    const StreamLayout L = Lcur;
    const double *gxl = A.x_left + (size_t)mem * A.capL, *gcl = A.C_left + (size_t)mem * A.capL;
    const double *gxr = A.x_right + (size_t)mem * A.capR, *gcr = A.C_right + (size_t)mem * A.capR;
    for (int i = tid; i < n1; i += T) { const size_t a = paddr(T, L, 0, i); X[a] = gxl[i]; C[a] = gcl[i]; }
    for (int i = tid; i < n2; i += T) { const size_t a = paddr(T, L, 1, n2 - 1 - i); X[a] = gxr[i]; C[a] = gcr[i]; }
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
  HistCache hc;
  if (tid == 0) hist_load(hc, P, S.hidx);

  while (true) {
    if (S.status != 0 || !(S.t < S.t_tot) || (A.max_steps > 0 && S.steps >= A.max_steps)) break;
    __syncthreads();
    // ---------------- scalar phase (thread 0): dt, time, parameters, interface advection, new node counts ---------------
    if (tid == 0) {
      const StreamLayout L = Lcur;
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
      if (!rc) {
        S.KD = KD; S.T = Tt; S.D_l = Dl; S.D_r = Dr;
        const int n1 = S.n1, n2 = S.n2;
        const double Rnew = xadd(S.Ri1, xmul(S.V, S.dt));                                        // functions.jl:31
        if (S.V > 0.0) {
          if (n1 + 1 > A.capL) rc = 10;
          else {
            S.Ri1 = Rnew;
            if (Rnew > X[paddr(T, L, 1, n2 - 2)]) rc = 1;                                         // x_right[2], :37-40
            else {
              X[paddr(T, L, 0, n1)] = Rnew; C[paddr(T, L, 0, n1)] = C[paddr(T, L, 0, n1 - 1)];    // :33-35
              X[paddr(T, L, 1, n2 - 1)] = Rnew;                                                   // :41
              S.n1 = n1 + 1;
            }
          }
        } else {
          if (n2 + 1 > A.capR) rc = 10;
          else {
            S.Ri1 = Rnew;
            if (Rnew < X[paddr(T, L, 0, n1 - 2)]) rc = 2;                                         // x_left[end-1], :50-53
            else {
              X[paddr(T, L, 1, n2)] = Rnew; C[paddr(T, L, 1, n2)] = C[paddr(T, L, 1, n2 - 1)];    // :46-48 (append in reversed order)
              X[paddr(T, L, 0, n1 - 1)] = Rnew;                                                   // :54
              S.n2 = n2 + 1;
            }
          }
        }
      }
      if (!rc) {                                                                                  // regrid! :873-889
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
      }
      if (rc) S.status = rc; else S.node_steps += S.n1 + S.n2;
    }
    __syncthreads();
    if (S.status != 0) break;

    // ---------------- solve ------------------------------------------------------------------------------------------------
    const StreamLayout L = Lcur;
    const int n1 = S.n1, n2 = S.n2, m = L.m;
    const double idt = frcp_(S.dt);
    // thread t owns column t; its sweep rows are its nodes except the phase's last node (the interface node)
    const int KL = (n1 - 1 + m - 1) / m, KR = (n2 - 1 + m - 1) / m;       // columns with sweep rows
    const int blk = (tid >= L.colR0) ? 1 : 0;
    const int jc = blk ? (tid - L.colR0) : tid;
    const int nph = blk ? n2 : n1;
    const bool active = blk ? (jc < KR) : (jc < KL);
    const int f = jc * m;                                                   // first phase-ordered index of the column
    const int len = active ? (min(f + m, nph - 1) - f) : 0;
    const bool lastChunk = active && (jc == (blk ? KR : KL) - 1);
    const double Dph = blk ? S.D_r : S.D_l;
    const bool dir0 = (blk ? (S.bc1 == 1) : (S.bc0 == 1)) && f == 0;
    const double Co_out = X == nullptr ? 0.0 : C[paddr(T, L, blk, 0)];      // old value at the outer boundary of this phase
    const double a_old = C[paddr(T, L, 0, n1 - 1)], b_old = C[paddr(T, L, 1, n2 - 1)];
    (void)a_old; (void)b_old;
    double Ap = 0.0, Gp = 0.0, Rp = 0.0, P0 = 0.0, Qa = 0.0, Qc = 0.0;
    double bA = 0.0, bG = -1.0, bR = 0.0;
    // halo values must be read before any thread overwrites C in place
    double c_halo_prev = 0.0, c_halo_next = 0.0, x_halo_prev = 0.0, x_halo_next = 0.0;
    if (active) {
      if (f > 0) { x_halo_prev = X[paddr(T, L, blk, f - 1)]; c_halo_prev = C[paddr(T, L, blk, f - 1)]; }
      x_halo_next = X[paddr(T, L, blk, f + len)]; c_halo_next = C[paddr(T, L, blk, f + len)];
    }
    fence_proxy_async();                      // the rows were written with ordinary stores (regrid, node insertion)
    __syncthreads();
    {                                         // all threads (a column without rows has len == 0)
      // forward sweep over the column (coalesced: element k of column t at k*T + t), same row arithmetic as the tuned
      // resident kernel (direction-agnostic element matrices, combined element carry).  The first row of the column
      // in terms of the two separators, C_0 = bR - bA*s_prev - bG*s_own, is accumulated on the way
      // (bR = sum pi_k R_k, bA = sum pi_k A_k, pi_k = prod_{j<k}(-G_j), bG = -pi_{len-1}): no second pass over A, G, R.
      const int sd = blk ? -1 : 1;         // orientation of x along the sweep: left increasing, right (reversed) decreasing
      const double Dsd = (sd > 0) ? Dph : -Dph, idts = (sd > 0) ? idt : -idt;
      size_t a = saddr(T, tid, 0);
      double x_self = X[a], c_self = C[a];
      double dcar = 0.0, rcar = 0.0, sub = 0.0, hprev = 0.0, Vs = 0.0, Qb = 0.0, pi = 1.0;
      if (mb) Vs = powd<NDIM>(x_self);
      if (f > 0) {
        double Apv, A12, Aown, kk;
        elem_mats_dir<NDIM>(x_halo_prev, x_self, Dsd, idts, Apv, A12, Aown, kk);
        dcar = Aown - kk; rcar = fma(A12, c_halo_prev, Aown * c_self); sub = A12 + kk;
        if (mb) hprev = 0.5 * fabs(Vs - powd<NDIM>(x_halo_prev));
      }
      bA = 0.0; bR = 0.0;
      // one row of the sweep (k: row of the column, a = k*T + tid), given the next node's x and C
      auto fwd_row = [&](const int k, const size_t a, const double x_next, const double c_next) {
        const bool last = (k == len - 1);
        double Aself, A12, Anext, kk;
        elem_mats_dir<NDIM>(x_self, x_next, Dsd, idts, Aself, A12, Anext, kk);
        double diag = (Aself - kk) + dcar;
        double rhs = fma(Aself, c_self, A12 * c_next) + rcar;
        const double sup_true = A12 + kk;
        double sup = sup_true;
        if (k == 0 && dir0) { diag = 1.0; rhs = Co_out; sup = 0.0; sub = 0.0; }
        const double inv = frcp_(diag - sub * Gp);
        const double An = (k == 0) ? (sub * inv) : (-sub * Ap * inv);
        const double Gn = sup * inv;
        const double Rn = (rhs - sub * Rp) * inv;
        FA[a] = An; FG[a] = Gn; C[a] = Rn;
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
      // rows 1 .. m-1 (the "next" rows of iterations 0 .. m-2) come through shared memory in batches of STG_R rows;
      // iteration len-1 of every column uses its halo instead
      {
        const int nrows = m - 1, nbatch = (nrows + STG_R - 1) / STG_R;
        auto issue = [&](const int b) {       // thread 0: rows 1 + b*R .. of x and C into buffer b % NBF
          const int r0 = 1 + b * STG_R, nr = min(STG_R, nrows - b * STG_R), buf = b % STG_NBF;
          const unsigned bytes = (unsigned)nr * T * 8u;
          double *dst = stg + (size_t)buf * (2 * STG_R * T);
          mbar_expect_tx(&fullb[buf], 2u * bytes);
          bulk_g2s(dst, X + (size_t)r0 * T, bytes, &fullb[buf]);
          bulk_g2s(dst + STG_R * T, C + (size_t)r0 * T, bytes, &fullb[buf]);
        };
        if (tid == 0) for (int b = 0; b < STG_NBF - 1 && b < nbatch; ++b) issue(b);
        for (int b = 0; b < nbatch; ++b) {
          const int buf = b % STG_NBF;
          if (tid == 0 && b + STG_NBF - 1 < nbatch) issue(b + STG_NBF - 1);     // into the buffer freed by batch b-1
          mbar_wait(&fullb[buf], (stg_par >> buf) & 1u);
          stg_par ^= 1u << buf;
          const double *xs_ = stg + (size_t)buf * (2 * STG_R * T) + tid, *cs_ = xs_ + STG_R * T;
          const int k0 = b * STG_R, nr = min(STG_R, nrows - k0);
#pragma unroll 1
          for (int rr = 0; rr < nr; ++rr) {
            const int k = k0 + rr;
            if (k < len) {
              const bool last = (k == len - 1);
              fwd_row(k, (size_t)k * T + tid, last ? x_halo_next : xs_[rr * T], last ? c_halo_next : cs_[rr * T]);
            }
          }
          __syncthreads();                      // everybody is done with this buffer
        }
        if (len == m) fwd_row(m - 1, (size_t)(m - 1) * T + tid, x_halo_next, c_halo_next);
      }
      bG = -pi;
    }
    fence_proxy_async();                      // A, G, R rows go back through the bulk-copy path in the back-substitution
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
    // interface condition (nodes next to the interface: phase-ordered n-2 and n-3)
    double aif, bif;
    {
      const double xp_ = X[paddr(T, L, 0, n1 - 1)], xq_ = X[paddr(T, L, 1, n2 - 1)];
      const double xpm = X[paddr(T, L, 0, n1 - 2)], xqp = X[paddr(T, L, 1, n2 - 2)];
      const int rL = KL - 1, rR = L.colR0 + KR - 1;                         // reduced rows of the last chunks
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
    {
      // back-substitution, rows m-2 .. 0 of (R, A, G) staged through shared memory in batches of STG_R rows (descending)
      const double beta = blk ? bif : aif;
      double sp = 0.0, cn = 0.0;
      if (active) {
        cn = red0[tid] + red1[tid] * beta;
        sp = (jc > 0) ? (red0[tid - 1] + red1[tid - 1] * beta) : 0.0;
        C[saddr(T, tid, len - 1)] = cn;
      }
      const int nrows = m - 1, nbatch = (nrows + STG_R - 1) / STG_R;
      auto issue = [&](const int b) {
        const int hi = m - 2 - b * STG_R, lo = max(hi - STG_R + 1, 0), nr = hi - lo + 1, buf = b % STG_NBB;
        const unsigned bytes = (unsigned)nr * T * 8u;
        double *dst = stg + (size_t)buf * (3 * STG_R * T);
        mbar_expect_tx(&fullb[buf], 3u * bytes);
        bulk_g2s(dst, C + (size_t)lo * T, bytes, &fullb[buf]);
        bulk_g2s(dst + STG_R * T, FA + (size_t)lo * T, bytes, &fullb[buf]);
        bulk_g2s(dst + 2 * STG_R * T, FG + (size_t)lo * T, bytes, &fullb[buf]);
      };
      if (tid == 0) for (int b = 0; b < STG_NBB - 1 && b < nbatch; ++b) issue(b);
      for (int b = 0; b < nbatch; ++b) {
        const int buf = b % STG_NBB;
        if (tid == 0 && b + STG_NBB - 1 < nbatch) issue(b + STG_NBB - 1);
        mbar_wait(&fullb[buf], (stg_par >> buf) & 1u);
        stg_par ^= 1u << buf;
        const int hi = m - 2 - b * STG_R, lo = max(hi - STG_R + 1, 0);
        const double *rs_ = stg + (size_t)buf * (3 * STG_R * T) + tid, *as_ = rs_ + STG_R * T, *gs_ = as_ + STG_R * T;
#pragma unroll 1
        for (int k = hi; k >= lo; --k) {
          if (k <= len - 2) {
            const int o = (k - lo) * T;
            cn = rs_[o] - as_[o] * sp - gs_[o] * cn;
            C[(size_t)k * T + tid] = cn;
          }
        }
        __syncthreads();
      }
    }
    if (tid == 0) { C[paddr(T, L, 0, n1 - 1)] = aif; C[paddr(T, L, 1, n2 - 1)] = bif; }
    __threadfence_block();
    __syncthreads();

    // ---------------- regrid (MRefin == 1: two LinRanges, functions.jl:396-400) + pchip ------------------------------------
    {
      const StreamLayout LN = Lnew;
      const int nr1 = S.nr1, nr2 = S.nr2;
      const double Ri1 = S.Ri1, Ri2 = S.Ri2;
      // thread t fills column t of the NEW layout: nodes idx0..idx0+cnt-1 of one phase in phase order
      const int nblk = (tid >= LN.colR0) ? 1 : 0;
      const int ncol = nblk ? (tid - LN.colR0) : tid;
      const int nn = nblk ? nr2 : nr1;          // new node count of the phase
      const int no = nblk ? n2 : n1;            // old node count of the phase
      const int idx0 = ncol * LN.m;
      const int cnt = max(0, min(idx0 + LN.m, nn) - idx0);
      if (cnt > 0) {
        if (do_hash) pchip_run_stream<T, true>(X, C, L, nblk, no, nn, nblk ? Ri1 : 0.0, nblk ? Ri2 : Ri1, Xn, Cn, idx0, cnt, S.it, hash_local);
        else pchip_run_stream<T, false>(X, C, L, nblk, no, nn, nblk ? Ri1 : 0.0, nblk ? Ri2 : Ri1, Xn, Cn, idx0, cnt, 0, hash_local);
      }
    }
    __threadfence_block();
    __syncthreads();
    { double *tp = X; X = Xn; Xn = tp; tp = C; C = Cn; Cn = tp; }
    if (tid == 0) {
      S.n1 = S.nr1; S.n2 = S.nr2;
      Lcur = Lnew;
      S.dx1 = xsub(linrange_at(0.0, S.Ri1, S.n1, S.n1 - 1), linrange_at(0.0, S.Ri1, S.n1, S.n1 - 2));   // regrid! :893-894
      S.dx2 = xsub(linrange_at(S.Ri1, S.Ri2, S.n2, 1), linrange_at(S.Ri1, S.Ri2, S.n2, 0));
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
    for (int i = tid; i < n1 && i < A.capL; i += T) { const size_t a = paddr(T, L, 0, i); gxl[i] = X[a]; gcl[i] = C[a]; }
    for (int i = tid; i < n2 && i < A.capR; i += T) { const size_t a = paddr(T, L, 1, n2 - 1 - i); gxr[i] = X[a]; gcr[i] = C[a]; }
    if ((A.flags & F_MASS_FINAL) && A.Mass) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n1 - 1; i += T) {
        const size_t a = paddr(T, L, 0, i), b = paddr(T, L, 0, i + 1);
        v[0] += 0.5 * (S.rho1 * powd<NDIM>(X[b]) - S.rho1 * powd<NDIM>(X[a])) * (C[a] + C[b]);
      }
      for (int i = tid; i < n2 - 1; i += T) {
        const size_t a = paddr(T, L, 1, n2 - 1 - i), b = paddr(T, L, 1, n2 - 2 - i);
        v[1] += 0.5 * (S.rho2 * powd<NDIM>(X[b]) - S.rho2 * powd<NDIM>(X[a])) * (C[a] + C[b]);
      }
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
