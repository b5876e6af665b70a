// mbm_kernel_v2.cuh — the tuned resident kernel for the mainstream case of the hot path: prescribed-
// velocity couples with a moving interface on graded meshes (inner_bc = flux | mb, V_ip != 0, MRefin > 1),
// i.e. BASELINE.json configs[2..3] and the B4/B5/C1/MB_growth fixtures.  Everything else (fixed interface,
// Stefan, single phase, MRefin == 1, profiles longer than the shared-memory capacity) runs in the general
// kernel of mbm_kernels.cuh; members this kernel cannot hold are handed over with status -1.
//
// Per member: three shared-memory arrays (x, C, work) sized for two CTAs per SM; one forward elimination
// pass per step with the fill-in column held in registers; fast reciprocals (MUFU.RCP64H + 2 Newton steps)
// in value code; the new mesh is generated and the profile re-interpolated in registers, nine consecutive
// nodes per thread (the canonical scan tile), with an interval search that walks from the previous node.
#pragma once
#include "mbm_kernels.cuh"

namespace mbm {

__device__ __forceinline__ double frcp(double x) { return frcp_(x); }

// element matrices with the fast reciprocal (reference src/functions.jl:525-540)
template <int NDIM>
__device__ __forceinline__ double powd(double x) {
  if (NDIM == 1) return x;
  if (NDIM == 2) return x * x;
  return x * x * x;
}

template <int NDIM>
__device__ __forceinline__ void elem_mats_fast(double x1, double x2, double D, double idt, double &A11,
                                               double &A12, double &A22, double &kk) {
  const double df = x1 - x2;
  const double idf = frcp(df);
  if (NDIM == 3) {
    const double x1s = x1 * x1, x2s = x2 * x2, x12 = x1 * x2;
    const double c30 = -df * (idt * (1.0 / 30.0)), c60 = -df * (idt * (1.0 / 60.0));
    A11 = (6.0 * x1s + 3.0 * x12 + x2s) * c30;
    A12 = (3.0 * x1s + 4.0 * x12 + 3.0 * x2s) * c60;
    A22 = (x1s + 3.0 * x12 + 6.0 * x2s) * c30;
    kk = (D * (1.0 / 3.0)) * (x1s + x12 + x2s) * idf;
  } else if (NDIM == 2) {
    const double c12 = -df * (idt * (1.0 / 12.0));
    A11 = (3.0 * x1 + x2) * c12;
    A12 = (x1 + x2) * c12;          // (x2^2 - x1^2)/12 = -(x1-x2)(x1+x2)/12
    A22 = (x1 + 3.0 * x2) * c12;
    kk = (0.5 * D) * (x1 + x2) * idf;
  } else {
    A11 = -df * (idt * (1.0 / 3.0));
    A12 = -df * (idt * (1.0 / 6.0));
    A22 = A11;
    kk = D * idf;
  }
}

template <int T, int MAXM, int NDIM>
__global__ void __launch_bounds__(T, 2) mbm_step_kernel_v2(const KArgs A, const int member0) {
  extern __shared__ double smem[];
  __shared__ Scal S;
  __shared__ double redS[T], redD[T], redU[T], red0[T], red1[T];
  __shared__ double exA[T], exG[T], exR[T];
  __shared__ double gtotA[16], gtotB[16], gtotC[16];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int mem = member0 + blockIdx.x;
  if (mem >= A.M) return;
  const int kcapL = A.kcapL, kcapR = A.kcapR;
  const int SZ = kcapL + 1 + kcapR;
  const int RB0 = kcapL + 1;                 // storage index of x_right[1] after a regrid
  double *X = smem, *C = smem + SZ, *W = smem + 2 * (size_t)SZ;
  const int bc = A.inner_bc;
  constexpr int ndim = NDIM;
  MemberPtrs P;
  P.t_ar = A.t_ar + (size_t)mem * A.t_stride;
  P.T_ar = A.T_ar + (size_t)mem * A.T_stride;
  P.KD_ar = A.KD_ar + (size_t)mem * A.KD_stride;
  double *recs = A.recs + (size_t)blockIdx.x * T * REC_N;

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
    S.gap = RB0;                                   // here: storage index of x_right[1]
    S.rho1 = A.rho[2 * (size_t)mem]; S.rho2 = A.rho[2 * (size_t)mem + 1];
    S.V = A.V_ip ? A.V_ip[mem] : 0.0;
    S.MRefin = A.MRefin[mem];
    S.resmin1 = A.resmin[2 * (size_t)mem]; S.resmin2 = A.resmin[2 * (size_t)mem + 1];
    S.Ri1 = A.Ri[2 * (size_t)mem]; S.Ri2 = A.Ri[2 * (size_t)mem + 1];
    S.Mass0 = A.Mass0 ? A.Mass0[mem] : 0.0;
    S.newton_R = A.newton_R ? A.newton_R[mem] : 0.0;
    S.rec_status = 0; S.nbL = -1; S.nbR = -1; S.m = 1; S.KL = 0; S.KR = 0;
    // eligibility: anything this kernel does not cover is handed to the general kernel (status -1)
    S.use_chain = 1;
    if (S.status == 0) {
      const bool ok = (S.V != 0.0) && (S.MRefin > 1.0) && (ndim >= 1 && ndim <= 3) && S.n1 >= 3 && S.n2 >= 3 &&
                      S.n1 + 1 <= kcapL && S.n2 + 1 <= kcapR && S.resmin1 <= kcapL && S.resmin2 <= kcapR;
      if (!ok) { A.status[mem] = -1; if (A.steps_done) A.steps_done[mem] = 0; S.use_chain = 0; }
    } else S.use_chain = 0;
  }
  __syncthreads();
  if (!S.use_chain) return;
  {
    const int n1 = S.n1, n2 = S.n2;
    const double *gxl = A.x_left + (size_t)mem * A.capL, *gcl = A.C_left + (size_t)mem * A.capL;
    const double *gxr = A.x_right + (size_t)mem * A.capR, *gcr = A.C_right + (size_t)mem * A.capR;
    for (int i = tid; i < n1; i += T) { X[i] = gxl[i]; C[i] = gcl[i]; }
    for (int i = tid; i < n2; i += T) { X[RB0 + i] = gxr[i]; C[RB0 + i] = gcr[i]; }
    double v[4] = {1e300, -1e300, 1e300, -1e300}, tm[2] = {1e300, -1e300};
    for (int i = tid; i < A.hist_len; i += T) {
      const double tt = P.t_ar[i], TT = P.T_ar[i], kd = P.KD_ar[i];
      tm[0] = fmin(tm[0], tt); tm[1] = fmax(tm[1], tt);
      v[2] = fmin(v[2], TT); v[3] = fmax(v[3], TT);
      v[0] = fmin(v[0], kd); v[1] = fmax(v[1], kd);
    }
    redS[tid] = v[0]; redD[tid] = v[1]; redU[tid] = v[2]; red0[tid] = v[3]; red1[tid] = tm[0]; exA[tid] = tm[1];
    __syncthreads();
    if (tid == 0) {
      double a0 = 1e300, a1 = -1e300, a2 = 1e300, a3 = -1e300, a4 = 1e300, a5 = -1e300;
      for (int k = 0; k < T; ++k) {
        a0 = fmin(a0, redS[k]); a1 = fmax(a1, redD[k]); a2 = fmin(a2, redU[k]); a3 = fmax(a3, red0[k]);
        a4 = fmin(a4, red1[k]); a5 = fmax(a5, exA[k]);
      }
      S.KDmin = a0; S.KDmax = a1; S.Tmin = a2; S.Tmax = a3; S.tmin = a4; S.tmax = a5;
      S.dx1 = xsub(X[n1 - 1], X[n1 - 2]);
      S.dx2 = xsub(X[RB0 + 1], X[RB0]);
      const int rc = update_t_dependent(S, A, P, false);                   // pre-loop call (..._MB.jl:75)
      if (rc) S.status = rc;
      Chain &c = S.ch;
      c.t = S.t; c.dt = S.dt; c.it = S.it; c.steps = 0; c.Ri1 = S.Ri1; c.dx1 = S.dx1; c.dx2 = S.dx2;
      c.D_l = S.D_l; c.D_r = S.D_r; c.n1 = S.n1; c.n2 = S.n2; c.hidx = S.hidx;
      c.xl_m2 = X[n1 - 2]; c.xr_1 = X[RB0 + 1];
      c.csA = 0.0; c.csB = 0.0; c.cs_nl = -1; c.need_scan = 0; c.scan_nl = 0; c.done = 0; c.status = 0;
      c.count = 0; c.pos = 0;
    }
  }
  __syncthreads();

  unsigned long long hash_local = 0ull;
  const bool do_hash = (A.idx_hash != nullptr);
  const bool mb = (bc == BC_MB);
  long long steps_done = 0;
  int lpos = 0, lcount = 0;

  while (true) {
    // ---------- refill the batch of step records (run-ahead scalar chain + parallel Newton) ----------
    if (lpos >= lcount) {
      if (S.ch.done) break;
      __syncthreads();
      if (S.ch.need_scan) {
        const int nl = S.ch.scan_nl;
        const double bR = xmul(1.0, xinv(S.MRefin));
        scan_canonical<T, 0>([&](int i) { return linrange_at(1.0, bR, nl, i); }, nl, W, gtotA);
        if (tid == 0) { S.ch.csA = W[nl - 1]; S.ch.csB = W[nl - 2]; S.ch.cs_nl = nl; S.ch.need_scan = 0; }
        __syncthreads();
      }
      if (tid == 0) chain_build(S, A, P, recs, T, kcapL, kcapR);
      __syncthreads();
      const int cnt = S.ch.count;
      lcount = cnt; lpos = 0;
      if (tid < cnt) {
        double *r = recs + (size_t)tid * REC_N;
        if (r[REC_STATUS] == 0.0) {
          double Rt; int wn = 0;
          const int rc = newton_solver(xsub(S.Ri2, r[REC_RI1]), r[REC_D], (long long)r[REC_NR2] - 1, 1.1, &Rt, &wn);
          r[REC_R] = Rt; r[REC_WARN] = (double)wn;
          if (rc) r[REC_STATUS] = -(double)rc;
        }
      }
      __syncthreads();
      if (cnt == 0) { if (S.ch.done) break; continue; }
    }
    // ---------- consume one record: scalar phase of this step + node insertion ----------
    if (tid == 0 && S.status == 0) {
      const double *r = recs + (size_t)lpos * REC_N;
      const int st = (int)r[REC_STATUS];
      S.rec_status = 0;
      if (st > 0) {
        const bool soft = (st == 10) && (kcapL < A.capL || kcapR < A.capR);
        if (soft) S.status = -1;                    // does not fit this kernel: general kernel continues
        else {
          S.status = st; S.dt = r[REC_DT]; S.t = r[REC_T]; S.it = (long long)r[REC_IT];
          if (st == 1 || st == 2) S.Ri1 = r[REC_RI1];
        }
      } else {
        S.dt = r[REC_DT]; S.t = r[REC_T]; S.KD = r[REC_KD]; S.T = r[REC_TEMP]; S.D_l = r[REC_DL]; S.D_r = r[REC_DR];
        S.it = (long long)r[REC_IT];
        const int n1o = S.n1, rb = S.gap;
        const double Rnew = r[REC_RI1];
        S.Ri1 = Rnew;
        if (S.V > 0.0) { X[n1o] = Rnew; C[n1o] = C[n1o - 1]; X[rb] = Rnew; S.n1 = n1o + 1; }
        else { X[rb - 1] = Rnew; C[rb - 1] = C[rb]; S.gap = rb - 1; S.n2 += 1; X[n1o - 1] = Rnew; }
        S.nr1 = (int)r[REC_NR1]; S.nr2 = (int)r[REC_NR2];
        S.newton_R = r[REC_R]; S.grid_d = r[REC_D];
        S.warn |= (int)r[REC_WARN];
        if (st < 0) S.rec_status = -st;
        S.node_steps += S.n1 + S.n2;
        if (S.n1 != S.nbL + 1 || S.n2 != S.nbR + 1) {   // rows per thread (odd: conflict-free), chunks per block
          const int nbL = S.n1 - 1, nbR = S.n2 - 1;
          int m = (nbL + nbR + T - 1) / T;
          int KL = (nbL + m - 1) / m, KR = (nbR + m - 1) / m;
          while (KL + KR > T) { ++m; KL = (nbL + m - 1) / m; KR = (nbR + m - 1) / m; }
          if (!(m & 1)) { m += 1; KL = (nbL + m - 1) / m; KR = (nbR + m - 1) / m; }
          S.m = m; S.KL = KL; S.KR = KR; S.nbL = nbL; S.nbR = nbR;   // m <= MAXM: host guarantees kcapL+kcapR <= (T-2)*MAXM
        }
      }
    }
    lpos += 1;
    __syncthreads();
    if (S.status != 0) break;

    // =========================== solve ===========================
    const int n1 = S.n1, n2 = S.n2, rb = S.gap;
    const double idt = frcp(S.dt);
    const int nbL = n1 - 1, nbR = n2 - 1;
    const int m = S.m, KL = S.KL, KR = S.KR;     // partition computed by thread 0 when the record was consumed
    const bool active = tid < KL + KR;
    const int blk = (tid < KL) ? 0 : 1;
    const int jc = blk ? (tid - KL) : tid;
    const int nb = blk ? nbR : nbL;
    const int f = jc * m;
    const int len = active ? (min(f + m, nb) - f) : 0;
    const int l = f + len - 1;
    const int sbase = blk ? (rb + n2 - 1) : 0, sdir = blk ? -1 : 1;
#define ST(ell) (sbase + sdir * (ell))
    const double Dph = blk ? S.D_r : S.D_l;
    const bool dir0 = blk ? (S.bc1 == 1) : (S.bc0 == 1);
    const double Co_first = C[0], Co_last = C[rb + n2 - 1];
    const double a_old = C[n1 - 1], b_old = C[rb];
    double c_halo_prev = 0.0, c_halo_next = 0.0;
    if (active) {
      if (f > 0) c_halo_prev = C[ST(f - 1)];
      c_halo_next = C[ST(l + 1)];
    }
    __syncthreads();

    double Areg[MAXM];
    double P0 = 0.0, Qa = 0.0, Qc = 0.0;
    double Ap = 0.0, Gp = 0.0, Rp = 0.0;
    if (active) {
      // ---- forward pass: assemble row, eliminate, store (A: registers, G: W, R: C) ----
      double x_self = X[ST(f)], c_self = C[ST(f)];
      double x_prev = 0.0, c_prev = c_halo_prev;
      double pAc = 0.0, pAnext = 0.0, pkk = 0.0;
      bool haveP = f > 0;
      if (haveP) {
        x_prev = X[ST(f - 1)];
        double A11, A12, A22, kk;
        if (sdir > 0) { elem_mats_fast<NDIM>(x_prev, x_self, Dph, idt, A11, A12, A22, kk); pAnext = A22; }
        else { elem_mats_fast<NDIM>(x_self, x_prev, Dph, idt, A11, A12, A22, kk); pAnext = A11; }
        pAc = A12; pkk = kk;
      }
      double Qb = 0.0, Vm = 0.0, Vs = 0.0;
      if (mb) { Vs = powd<NDIM>(x_self); Vm = haveP ? powd<NDIM>(x_prev) : Vs; }
      const int s0 = ST(f);
#pragma unroll
      for (int k = 0; k < MAXM; ++k) {
        if (k < len) {
          const int ell = f + k;
          const int s = s0 + ((sdir > 0) ? k : -k);
          const int sn = s + sdir;
          const double x_next = X[sn];
          const double c_next = (k == len - 1) ? c_halo_next : C[sn];
          double A11, A12, A22, kk, nAself, nAnext;
          if (sdir > 0) { elem_mats_fast<NDIM>(x_self, x_next, Dph, idt, A11, A12, A22, kk); nAself = A11; nAnext = A22; }
          else { elem_mats_fast<NDIM>(x_next, x_self, Dph, idt, A11, A12, A22, kk); nAself = A22; nAnext = A11; }
          double diag = (nAself - kk) + (haveP ? (pAnext - pkk) : 0.0);
          double rhs = (nAself * c_self + A12 * c_next) + (haveP ? (pAc * c_prev + pAnext * c_self) : 0.0);
          double sup = A12 + kk;
          double sub = haveP ? (pAc + pkk) : 0.0;
          if (ell == 0 && dir0) { diag = 1.0; rhs = blk ? Co_last : Co_first; sup = 0.0; sub = 0.0; }
          const double inv = frcp(diag - sub * Gp);
          const double An = (k == 0) ? (sub * inv) : (-sub * Ap * inv);
          const double Gn = sup * inv;
          const double Rn = (rhs - sub * Rp) * inv;
          Areg[k] = An;
          W[s] = Gn; C[s] = Rn;
          if (mb) {
            const double Vn = powd<NDIM>(x_next);
            // trapezoid weight 0.5*(V[g+1]-V[g]) + 0.5*(V[g]-V[g-1]) in global order (functions.jl:291-296)
            const double w = (sdir > 0) ? (0.5 * (Vn - Vs) + (haveP ? 0.5 * (Vs - Vm) : 0.0))
                                        : ((haveP ? 0.5 * (Vm - Vs) : 0.0) + 0.5 * (Vs - Vn));
            if (k < len - 1) {
              const double tq = Qb + w;
              P0 = fma(tq, Rn, P0);
              Qa = fma(-tq, An, Qa);
              Qb = -tq * Gn;
            } else Qc = Qb + w;
            Vm = Vs; Vs = Vn;
          }
          Ap = An; Gp = Gn; Rp = Rn;
          x_prev = x_self; c_prev = c_self; x_self = x_next; c_self = c_next;
          pAc = A12; pAnext = nAnext; pkk = kk; haveP = true;
        }
      }
      // ---- backward accumulate: first row of the chunk in terms of the two separators ----
      double bA = 0.0, bG = -1.0, bR = 0.0;
      if (len > 1) {
        bool first = true;
#pragma unroll
        for (int k = MAXM - 2; k >= 0; --k) {
          if (k < len - 1) {
            const int s = s0 + ((sdir > 0) ? k : -k);
            const double Gk = W[s], Rk = C[s], Ak = Areg[k];
            if (first) { bA = Ak; bG = Gk; bR = Rk; first = false; }
            else { bR = Rk - Gk * bR; bA = Ak - Gk * bA; bG = -Gk * bG; }
          }
        }
      }
      exA[tid] = bA; exG[tid] = bG; exR[tid] = bR;
    }
    __syncthreads();
    if (active) {
      const bool lastChunk = (jc == (blk ? KR : KL) - 1);
      if (!lastChunk) {
        const double nA = exA[tid + 1], nG = exG[tid + 1], nR = exR[tid + 1];
        redS[tid] = Ap; redD[tid] = 1.0 - Gp * nA; redU[tid] = -Gp * nG; red0[tid] = Rp - Gp * nR; red1[tid] = 0.0;
      } else {
        redS[tid] = Ap; redD[tid] = 1.0; redU[tid] = 0.0; red0[tid] = Rp; red1[tid] = -Gp;
      }
    } else {
      redS[tid] = 0.0; redD[tid] = 1.0; redU[tid] = 0.0; red0[tid] = 0.0; red1[tid] = 0.0;
    }
    __syncthreads();
    if (tid < 32) reduced_solve<T>(redS, redD, redU, red0, red1);
    __syncthreads();

    // ---- interface condition ----
    double aif, bif;
    if (bc == BC_FLUX) {
      const double idxl = frcp(X[n1 - 1] - X[n1 - 2]), idxr = frcp(X[rb + 1] - X[rb]);
      const double al = -S.rho1 * S.D_l * idxl, be = S.V * S.rho1 + S.rho1 * S.D_l * idxl;
      const double ga = -S.V * S.rho2 + S.rho2 * S.D_r * idxr, de = -S.rho2 * S.D_r * idxr;
      const double U0 = red0[KL - 1], U1 = red1[KL - 1], W0 = red0[KL + KR - 1], W1 = red1[KL + KR - 1];
      const double ca = be + al * U1, cb = ga + de * W1;
      const double r = -(al * U0 + de * W0);
      bif = r / (ca * S.KD + cb);
      aif = S.KD * bif;
    } else {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      if (active) {
        const double s0 = red0[tid], s1 = red1[tid];
        const double q0 = (jc > 0) ? red0[tid - 1] : 0.0, q1 = (jc > 0) ? red1[tid - 1] : 0.0;
        const double c0 = P0 + Qa * q0 + Qc * s0, c1 = Qa * q1 + Qc * s1;
        if (blk == 0) { v[0] = c0; v[1] = c1; } else { v[2] = c0; v[3] = c1; }
      }
      block_sum4<T>(v, S.red);
      const double Vp_ = powd<NDIM>(X[n1 - 1]), Vq_ = powd<NDIM>(X[rb]);
      const double wp = 0.5 * (Vq_ - Vp_) + 0.5 * (Vp_ - powd<NDIM>(X[n1 - 2]));
      const double wq = 0.5 * (powd<NDIM>(X[rb + 1]) - Vq_) + 0.5 * (Vq_ - Vp_);
      bif = (S.Mass0 - (v[0] + v[2])) / ((v[1] + wp) * S.KD + (v[3] + wq));
      aif = S.KD * bif;
    }
    (void)a_old; (void)b_old;
    // ---- back-substitution ----
    if (active) {
      const double beta = blk ? bif : aif;
      const double sj = red0[tid] + red1[tid] * beta;
      const double sp = (jc > 0) ? (red0[tid - 1] + red1[tid - 1] * beta) : 0.0;
      double cn = sj;
      C[ST(l)] = cn;
#pragma unroll
      for (int k = MAXM - 2; k >= 0; --k) {
        if (k < len - 1) {
          const int s = ST(f + k);
          cn = C[s] - Areg[k] * sp - W[s] * cn;
          C[s] = cn;
        }
      }
    }
    if (tid == 0) { C[n1 - 1] = aif; C[rb] = bif; }
    __syncthreads();
#undef ST

    // =========================== regrid: new meshes + pchip ===========================
    {
      const int nr1 = S.nr1, nr2 = S.nr2;
      const double Ri1 = S.Ri1;
      const double *xl_o = X, *yl_o = C, *xr_o = X + rb, *yr_o = C + rb;
      // ---- phase A: new meshes by the canonical scans (one tile of 9 elements per thread and phase), staged in W
      //      (W is free here: the slopes of the old profiles are computed on the fly during the evaluation)
      const int nl = nr1 - 1, nrr = nr2 - 1;
      const int e0 = tid * CSO_TILE;
      double tl[CSO_TILE], tr[CSO_TILE];
      double incL, incR;
      {
        const double bR = xmul(1.0, xinv(S.MRefin));
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < CSO_TILE; ++k) {
          const int i = e0 + k;
          if (i < nl) { const double v = linrange_at(1.0, bR, nl, i); s = (k == 0) ? v : xadd(s, v); }
          tl[k] = s;
        }
        incL = s;
        const double Rt = S.newton_R, d = S.grid_d;
        double pr = 1.0;
#pragma unroll
        for (int k = 0; k < CSO_TILE; ++k) {
          const int i = e0 + k;
          if (i < nrr) { const double v = (i == 0) ? d : Rt; pr = (k == 0) ? v : xmul(pr, v); }
          tr[k] = pr;
        }
        incR = pr;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const double oL = __shfl_up_sync(0xffffffffu, incL, off), oR = __shfl_up_sync(0xffffffffu, incR, off);
          if (lane >= off) { incL = xadd(oL, incL); incR = xmul(oR, incR); }
        }
        if (lane == 31) { gtotA[wid] = incL; gtotB[wid] = incR; }
      }
      if (tid == 0 && S.rec_status) S.status = S.rec_status;      // Newton failure raised inside regrid!
      __syncthreads();
      if (S.status != 0) break;
      double incS;
      {
        const double EqL = __shfl_up_sync(0xffffffffu, incL, 1), EqR = __shfl_up_sync(0xffffffffu, incR, 1);
        double OgL = 0.0, OgR = 1.0;
        if (wid > 0) { OgL = gtotA[0]; OgR = gtotB[0]; for (int g = 1; g < wid; ++g) { OgL = xadd(OgL, gtotA[g]); OgR = xmul(OgR, gtotB[g]); } }
        const double isum = xinv(S.ch.csA);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < CSO_TILE; ++k) {
          double cs = (lane == 0) ? tl[k] : xadd(EqL, tl[k]);
          if (wid > 0) cs = xadd(OgL, cs);
          if (e0 + k < nl) W[e0 + k + 1] = xmul(xmul(cs, Ri1), isum);             // X_left[e0+k+1]
          double dx = (lane == 0) ? tr[k] : xmul(EqR, tr[k]);
          if (wid > 0) dx = xmul(OgR, dx);
          if (e0 + k < nrr) s = (k == 0) ? dx : xadd(s, dx);
          tr[k] = s;
        }
        if (tid == 0) W[0] = xmul(xmul(0.0, Ri1), isum);
        incS = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, incS, off);
          if (lane >= off) incS = xadd(o, incS);
        }
        if (lane == 31) gtotC[wid] = incS;
      }
      __syncthreads();
      {
        const double EqS = __shfl_up_sync(0xffffffffu, incS, 1);
        double OgS = 0.0;
        if (wid > 0) { OgS = gtotC[0]; for (int g = 1; g < wid; ++g) OgS = xadd(OgS, gtotC[g]); }
#pragma unroll
        for (int k = 0; k < CSO_TILE; ++k) {
          double cs = (lane == 0) ? tr[k] : xadd(EqS, tr[k]);
          if (wid > 0) cs = xadd(OgS, cs);
          if (e0 + k < nrr) W[RB0 + e0 + k + 1] = xadd(Ri1, cs);                   // X_right[e0+k+1]
        }
        if (tid == 0) W[RB0] = Ri1;
      }
      __syncthreads();
      // ---- phase B: evaluate the old profiles on the new meshes, nodes interleaved over all threads
      const int ntot = nr1 + nr2;
      const float ratL = (float)n1 / (float)nr1, ratR = (float)n2 / (float)nr2;
      double Pq[MAXM];
#pragma unroll
      for (int q = 0; q < MAXM; ++q) {
        const int j = tid + q * T;
        if (j < ntot) {
          const bool left = j < nr1;
          const int jj = left ? j : (j - nr1);
          const double Xj = W[left ? jj : (RB0 + jj)];
          int kk, br;
          Pq[q] = pchip_eval_otf(left ? xl_o : xr_o, left ? yl_o : yr_o, left ? n1 : n2, Xj,
                                 __float2int_rd((float)jj * (left ? ratL : ratR)), &kk, &br);
          if (do_hash) hash_local += idx_term(S.it, left ? 0 : 1, jj, kk, br);
        }
      }
      __syncthreads();
      // ---- phase C: commit the new meshes and profiles
#pragma unroll
      for (int q = 0; q < MAXM; ++q) {
        const int j = tid + q * T;
        if (j < ntot) {
          const int sidx = (j < nr1) ? j : (RB0 + j - nr1);
          X[sidx] = W[sidx]; C[sidx] = Pq[q];
        }
      }
      if (tid == 0) { S.n1 = nr1; S.n2 = nr2; S.gap = RB0; }
      steps_done += 1;
      __syncthreads();
      if (tid == 0) {
        if (!isfinite(C[0]) || !isfinite(C[RB0 + nr2 - 1]) || !isfinite(C[nr1 - 1]) || !isfinite(S.Ri1)) S.status = 11;
      }
      if ((A.flags & F_RECORD) && A.rec) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < nr1 - 1; i += T)
          v[0] += 0.5 * (S.rho1 * powd<NDIM>(X[i + 1]) - S.rho1 * powd<NDIM>(X[i])) * (C[i] + C[i + 1]);
        for (int i = tid; i < nr2 - 1; i += T)
          v[1] += 0.5 * (S.rho2 * powd<NDIM>(X[RB0 + i + 1]) - S.rho2 * powd<NDIM>(X[RB0 + i])) * (C[RB0 + i] + C[RB0 + i + 1]);
        block_sum4<T>(v, S.red);
        if (tid == 0 && S.it >= 1 && S.it - 1 < A.rec_cap) {
          double *r = A.rec + ((size_t)mem * (size_t)A.rec_cap + (size_t)(S.it - 1)) * 4;
          r[0] = v[0] + v[1]; r[1] = S.T; r[2] = S.KD; r[3] = S.Ri1;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---------------- store member ----------------
  {
    const int n1 = S.n1, n2 = S.n2, rb = S.gap;
    double *gxl = A.x_left + (size_t)mem * A.capL, *gcl = A.C_left + (size_t)mem * A.capL;
    double *gxr = A.x_right + (size_t)mem * A.capR, *gcr = A.C_right + (size_t)mem * A.capR;
    for (int i = tid; i < n1 && i < A.capL; i += T) { gxl[i] = X[i]; gcl[i] = C[i]; }
    for (int i = tid; i < n2 && i < A.capR; i += T) { gxr[i] = X[rb + i]; gcr[i] = C[rb + i]; }
    if ((A.flags & F_MASS_FINAL) && A.Mass) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n1 - 1; i += T)
        v[0] += 0.5 * (S.rho1 * powd<NDIM>(X[i + 1]) - S.rho1 * powd<NDIM>(X[i])) * (C[i] + C[i + 1]);
      for (int i = tid; i < n2 - 1; i += T)
        v[1] += 0.5 * (S.rho2 * powd<NDIM>(X[rb + i + 1]) - S.rho2 * powd<NDIM>(X[rb + i])) * (C[rb + i] + C[rb + i + 1]);
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
      if (A.steps_done) A.steps_done[mem] = steps_done;
      if (A.warn) A.warn[mem] = S.warn;
      A.t[mem] = S.t; A.dt[mem] = S.dt; A.it[mem] = S.it;
      A.res[2 * (size_t)mem] = n1; A.res[2 * (size_t)mem + 1] = n2;
      A.Ri[2 * (size_t)mem] = S.Ri1; A.Ri[2 * (size_t)mem + 1] = S.Ri2;
      if (A.dx1) A.dx1[mem] = xsub(X[n1 - 1], X[n1 - 2]);
      if (A.dx2) A.dx2[mem] = xsub(X[rb + 1], X[rb]);
      if (A.newton_R) A.newton_R[mem] = S.newton_R;
      if (A.last) {
        double *L = A.last + 8 * (size_t)mem;
        L[0] = S.D_l; L[1] = S.D_r; L[2] = S.KD; L[3] = S.T; L[4] = S.V; L[5] = S.rho1; L[6] = S.rho2; L[7] = 1.0;
      }
      if (A.idx_hash) A.idx_hash[mem] = S.hash;
      if (A.node_steps) A.node_steps[mem] = S.node_steps;
    }
  }
}

}  // namespace mbm
