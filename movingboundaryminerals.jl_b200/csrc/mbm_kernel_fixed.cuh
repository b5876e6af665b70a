// mbm_kernel_fixed.cuh — resident kernel for couples whose interface does not move (V_ip == 0): BASELINE.json
// configs[2] (B1-style flux ensembles) and the fixtures B1, B2, Diff_couple_Flux/MB (fixed), Diff_couple_no_interaction.
//
// The mesh never changes, so the per-step system matrix (M/dt + K with the interface rows) only changes when dt or the
// diffusivities change (last clipped step; Arrhenius D(T(t))).  The kernel keeps the *factorisation* of the partitioned
// elimination resident in shared memory (per row: fill-in A, G, sub*inv and the mass-matrix row pre-multiplied by inv;
// per chunk: the constants of the reduced system and of the mass-balance functional) and per step only runs
//   R_l = mli*C_{l-1} + mdi*C_l + mui*C_{l+1} - fs*R_{l-1}      (right-hand side + forward substitution, 4 FMA/row)
//   backward accumulation of the chunk's first row, reduced solve (warp 0), 2x2 interface condition, back-substitution.
// A step whose (dt, D_l, D_r) differ from the factorised ones first re-runs the coefficient sweep (same arithmetic as
// the tuned moving-interface kernel).  The right phase is stored by REVERSED node index (right node j at N-1-j), so
// both blocks sweep with increasing storage index and every access of the row loops is base + immediate.  Loop body: reference main_codes/Diff_couple_1D_general_Flux.jl:87-130 with
// Fl_regrid == 0 (functions.jl:58-61), examples/Diff_couple_MB.jl:87-109, Diff_couple_no_interaction.jl:96-124.
#pragma once
#include <type_traits>
#include "mbm_kernel_v2.cuh"

namespace mbm {

// resident warps per SM the register allocation is sized for: 16 = 128 registers per thread (the nine arrays of config 3
// leave room for four 128-thread CTAs anyway; with 24 warps = 80 registers the compiler re-derived the array
// addresses inside every row loop)
#ifndef MBM_FIXED_WARPS
#define MBM_FIXED_WARPS 16
#endif
template <int T, int NDIM>
__global__ void __launch_bounds__(T, MBM_FIXED_WARPS * 32 / T) mbm_fixed_kernel(const KArgs A, const int member0) {
  extern __shared__ double smem[];
  __shared__ Scal S;
  __shared__ double redSD[2 * T], redU[T], red0[T + 32], red1[T];       // red0: per-step right-hand side, skewed (red_ix)
  double *const redS = redSD, *const redD = redSD + T;
  __shared__ double exA[T], exG[T], exR[T];
  // step-time views of arrays the factorisation only needs while it runs: redB (the neighbour chunk's part of each reduced
  // right-hand side, skewed like red0) and the solved first / last row of every chunk, which the neighbours' next forward
  // sweep reads as halos (the sweep then touches only its own rows of C: no barrier between the steps' reads and writes)
  double *const redB = redSD, *const hFirst = redU, *const hLast = exR;
  static_assert(2 * T >= T + 32, "redB (T + 32 skewed entries) must fit redS | redD");
  __shared__ double red1c[T];                // solution of the reduced system for the interface-unknown column (matrix-only)
  __shared__ double rfac[RFAC_N(T) * 32];    // matrix part of the reduced solve (reduced_factor), redone per factorisation

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int mem = member0 + blockIdx.x;
  if (mem >= A.M) return;
  const int SZ = A.kcapR;                    // capacity of this kernel's arrays (n1 + n2 <= SZ)
  // x, C, then the seven per-row quantities of the factorisation interleaved row by row, CO[7 s + j] (j = 0..6: mass-matrix
  // row times inv (ml, md, mu), sub * inv, G, fill-in A, mass-balance weight): one row pointer and immediate offsets for
  // every load of the row loops; seven is odd, so chunks of an odd number of rows still start in distinct banks
  double *const X = smem, *const C = smem + SZ, *const CO = smem + 2 * (size_t)SZ;
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
    S.node_steps = A.node_steps ? A.node_steps[mem] : 0;
    S.hidx = 0; S.Fl = 0; S.KD = 0.0; S.T = 0.0; S.D_l = 0.0; S.D_r = 0.0;
    S.n1 = (int)A.res[2 * (size_t)mem]; S.n2 = (int)A.res[2 * (size_t)mem + 1];
    S.gap = 0;
    S.rho1 = A.rho[2 * (size_t)mem]; S.rho2 = A.rho[2 * (size_t)mem + 1];
    S.V = A.V_ip ? A.V_ip[mem] : 0.0;
    S.MRefin = A.MRefin[mem];
    S.Ri1 = A.Ri[2 * (size_t)mem]; S.Ri2 = A.Ri[2 * (size_t)mem + 1];
    S.Mass0 = A.Mass0 ? A.Mass0[mem] : 0.0;
    S.newton_R = A.newton_R ? A.newton_R[mem] : 0.0;
    S.grid_d = 0.0; S.grid_S = 0.0; S.grid_total = 0.0;          // factorised (dt, D_l, D_r); 0 = none yet
    // this kernel takes members whose interface is fixed; as a later stage it only looks at handed-over members
    bool mine = A.only_deferred ? (S.status == -1) : (S.status == 0);
    const bool ok = (S.V == 0.0) && S.n1 >= 3 && S.n2 >= 3 && S.n1 + S.n2 <= SZ;
    if (mine && !ok) { if (!A.only_deferred) A.status[mem] = -1; mine = false; }
    if (mine) S.status = 0;
    S.use_chain = mine ? 1 : 0;
    if (mine && A.only_deferred && A.steps_done) S.steps = A.steps_done[mem];
  }
  __syncthreads();
  if (!S.use_chain) return;
  const int n1 = S.n1, n2 = S.n2, N = n1 + n2, rb = n1;
  {
    const double *gxl = A.x_left + (size_t)mem * A.capL, *gcl = A.C_left + (size_t)mem * A.capL;
    const double *gxr = A.x_right + (size_t)mem * A.capR, *gcr = A.C_right + (size_t)mem * A.capR;
    for (int i = tid; i < n1; i += T) { X[i] = gxl[i]; C[i] = gcl[i]; }
    for (int i = tid; i < n2; i += T) { X[N - 1 - i] = gxr[i]; C[N - 1 - i] = gcr[i]; }     // reversed storage
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
      S.dx1 = xsub(X[n1 - 1], X[n1 - 2]);                                   // create_grid!, functions.jl:366-367
      S.dx2 = xsub(X[N - 2], X[N - 1]);                                     // x_right[2] - x_right[1]
      const int rc = update_t_dependent(S, A, P, false);                   // pre-loop call (e.g. B1.jl:72)
      if (rc) S.status = rc;
    }
  }
  __syncthreads();

  // ---- partition (fixed for the whole call): block L = rows 0..n1-2 (towards the interface), block R = rows N-1..n1+1
  const int nbL = n1 - 1, nbR = n2 - 1;
  int m = (nbL + nbR + T - 1) / T;
  int KL = (nbL + m - 1) / m, KR = (nbR + m - 1) / m;
  while (KL + KR > T) { ++m; KL = (nbL + m - 1) / m; KR = (nbR + m - 1) / m; }
  if (!(m & 1)) { m += 1; KL = (nbL + m - 1) / m; KR = (nbR + m - 1) / m; }
  const bool active = tid < KL + KR;
  const int blk = (tid < KL) ? 0 : 1;
  const int jc = blk ? (tid - KL) : tid;
  const int nb = blk ? nbR : nbL;
  const int f = jc * m;
  const int len = active ? (min(f + m, nb) - f) : 0;
  const bool xdec = (blk != 0);              // x decreases along the sweep of the right block (it starts at the outer end)
  const int s0 = blk ? (rb + f) : f;         // storage index of local row f: both blocks sweep upwards
  const bool lastChunk = active && (jc == (blk ? KR : KL) - 1);
  const bool dir0 = blk ? (S.bc1 == 1) : (S.bc0 == 1);
  // per-chunk constants of the factorisation (registers)
  double cQa = 0.0, cQc = 0.0, cGp = 0.0, cAp = 0.0, cbA = 0.0, cbG = -1.0;
  double r_s = 0.0, r_u = 0.0, r_inv = 1.0, r_Gp = 0.0;     // reduced row: sub, sup, 1/diag, G of the chunk's last row
  double gi_prev = 0.0;                                     // r_Gp * r_inv of the previous chunk (0: none / it is a last chunk)
  const double *const co0 = CO + 7 * (size_t)s0;            // this chunk's rows of the factorisation and of the profile
  double *const c0p = C + s0;
  // Chunks of 3 / 5 full rows keep their rows of the profile (cst) and of G (gk: read twice per step) in registers
  // across the steps: the shared-memory pipe is what bounds this kernel (68 % busy at config 3's size with every row
  // re-read from shared memory).  C in shared memory is brought up to date at the exit (and per step for the history).
  const int fastM = (active && len == m && (m == 3 || m == 5)) ? m : 0;
  const bool rec_on = (A.flags & F_RECORD) && A.rec;
  double cst[5], gk[4];
#pragma unroll
  for (int k = 0; k < 5; ++k) cst[k] = (k < fastM) ? c0p[k] : 0.0;
  // thread 0's cached find_dt and history interval live in fields of the member scalars this kernel does not use otherwise
  // (registers are what bounds the chunks' resident rows below)
  double &dt_full = S.lp0, &dtD_l = S.lp1, &dtD_r = S.lp2;
  auto hc_get = [&]() { HistCache h; h.i = S.nqL; h.x0 = S.lp3; h.x1 = S.lp4; h.iv = S.a_if; h.k0 = S.b_if; h.k1 = S.Co_first; h.T0 = S.Co_last; h.T1 = S.sdx; return h; };
  auto hc_put = [&](const HistCache &h) { S.nqL = h.i; S.lp3 = h.x0; S.lp4 = h.x1; S.a_if = h.iv; S.b_if = h.k0; S.Co_first = h.k1; S.Co_last = h.T0; S.sdx = h.T1; };
  if (tid == 0) { dt_full = 0.0; dtD_l = -1.0; dtD_r = -1.0; HistCache h0; hist_load(h0, P, S.hidx); hc_put(h0); }

  // ---- scalar batches (constant diffusivities): with V_ip == 0 and no Arrhenius law the per-step scalars do not depend
  // on the solution and dt only changes on the clipped last step, so thread 0 runs the time chain (t += dt, exact adds)
  // for SB steps ahead and the lanes of warp 0 interpolate KD(t), T(t) of one step each (same arithmetic and cursor
  // semantics as update_t_dependent_param!, functions.jl:1350-1378).  A step then reads its record; there is no scalar
  // phase of thread 0 and no barrier pair around it.  Arrhenius members keep the per-step scalar phase (D feeds dt).
  // The records live in the unused tail of the block-sum scratch S.red (block_sum4 touches its first 4 * T / 32 <= 32
  // doubles): a kilobyte more static shared memory per CTA costs the fourth resident CTA at config 3's size.
  constexpr int SB = 16;
  double *const srec = S.red + 32;            // dt, t, KD, T of the batch (SB * 4 doubles)
  int *const srst = reinterpret_cast<int *>(S.red + 32 + SB * 4);   // status per record (SB ints)
  int &srcount = S.Fl;                        // records in the batch (S.Fl is not used by this kernel otherwise)
  static_assert(32 + SB * 4 + SB / 2 <= 128, "records must fit S.red");
  const bool batched = !S.arrhenius;
  int rpos = 0, rcount = 0;
  double fact_dt = 0.0;                       // dt of the resident factorisation (batched mode; 0 = none yet)

  while (true) {
    bool refactor;
    double dt_k, KD_k;
    if (batched) {
      if (S.status != 0) break;
      if (rpos == rcount) {
        __syncthreads();
        if (tid == 0) {
          if (!(S.D_l == dtD_l && S.D_r == dtD_r)) {
            const double mn = fmin(S.dx1, S.dx2), mxD = fmax(S.D_l, S.D_r);
            dt_full = xmul(xmul(xmul(xmul(mn, mn), xinv(mxD)), 50.0), S.CFL);
            dtD_l = S.D_l; dtD_r = S.D_r;
          }
          int cnt = 0;
          double t = S.t;
          long long steps = S.steps;
          while (cnt < SB && t < S.t_tot && !(A.max_steps > 0 && steps >= A.max_steps)) {
            double dt = dt_full;
            t = xadd(t, dt);
            if (t > S.t_tot) { dt = xsub(dt, xsub(t, S.t_tot)); t = S.t_tot; }
            srec[4 * cnt] = dt; srec[4 * cnt + 1] = t;
            ++cnt; ++steps;
          }
          srcount = cnt;
        }
        __syncthreads();
        rcount = srcount; rpos = 0;
        if (tid < 32) {                         // warp 0: one step per lane
          int hi_new = 0;
          if (tid < rcount) {
            HistCache h;
            hist_load(h, P, S.hidx);
            double KD, Tt, Dl, Dr;
            const int rc = chain_update_params(S, A, P, srec[4 * tid + 1], h, &KD, &Tt, &Dl, &Dr);
            srec[4 * tid + 2] = KD; srec[4 * tid + 3] = Tt; srst[tid] = rc;
            hi_new = h.i;
          }
          __syncwarp();
          if (rcount > 0 && tid == rcount - 1) S.hidx = hi_new;      // cursor of the batch's last step
        }
        __syncthreads();
        if (rcount == 0) break;
      }
      dt_k = srec[4 * rpos]; KD_k = srec[4 * rpos + 2];
      const int st_k = srst[rpos];
      if (tid == 0) {
        S.dt = dt_k; S.t = srec[4 * rpos + 1]; S.it += 1;
        if (st_k) S.status = st_k;
        else { S.KD = KD_k; S.T = srec[4 * rpos + 3]; S.node_steps += N; }
      }
      rpos += 1;
      if (st_k) break;
      refactor = (dt_k != fact_dt);
      fact_dt = dt_k;
    } else {
    if (S.status != 0 || !(S.t < S.t_tot) || (A.max_steps > 0 && S.steps >= A.max_steps)) break;
    __syncthreads();
    if (tid == 0) {
      // scalar phase (V_ip == 0: no advection).  find_dt (functions.jl:465-490) only changes with D; KD(t), T(t) by
      // linear_interpolation_1D with a cached interval (same arithmetic as update_t_dependent_param!, :1350-1378)
      if (!(S.D_l == dtD_l && S.D_r == dtD_r)) {
        const double mn = fmin(S.dx1, S.dx2), mxD = fmax(S.D_l, S.D_r);
        dt_full = xmul(xmul(xmul(xmul(mn, mn), xinv(mxD)), 50.0), S.CFL);
        dtD_l = S.D_l; dtD_r = S.D_r;
      }
      S.dt = dt_full;
      S.t = xadd(S.t, S.dt);
      S.it += 1;
      if (S.t > S.t_tot) { S.dt = xsub(S.dt, xsub(S.t, S.t_tot)); S.t = S.t_tot; }
      double KD, Tt, Dl, Dr;
      HistCache hc = hc_get();
      const int rc = chain_update_params(S, A, P, S.t, hc, &KD, &Tt, &Dl, &Dr);
      hc_put(hc);
      if (rc) S.status = rc;
      else {
        S.KD = KD; S.T = Tt; S.D_l = Dl; S.D_r = Dr;
        S.node_steps += N;
        S.rec_status = (S.dt != S.grid_d || S.D_l != S.grid_S || S.D_r != S.grid_total) ? 1 : 0;   // refactor?
        if (S.rec_status) { S.grid_d = S.dt; S.grid_S = S.D_l; S.grid_total = S.D_r; }
      }
    }
    __syncthreads();
    if (S.status != 0) break;
    refactor = S.rec_status != 0; dt_k = S.dt; KD_k = S.KD;
    }

    if (refactor) {
      // ---------------- coefficient sweep: factorise the step matrix (fill_matrix!, functions.jl:510-549) -------------
      const double idt = frcp(dt_k);
      const double Dph = blk ? S.D_r : S.D_l;
      if (active) {
        double x_self = X[s0], x_prev = 0.0;
        double pAc = 0.0, pAnext = 0.0, pkk = 0.0;
        const bool haveP0 = f > 0;
        if (haveP0) {
          x_prev = X[s0 - 1];
          double A11, A12, A22, kk;
          if (!xdec) elem_mats_fast<NDIM>(x_prev, x_self, Dph, idt, A11, A12, A22, kk);
          else elem_mats_fast<NDIM>(x_self, x_prev, Dph, idt, A11, A12, A22, kk);
          pAnext = (!xdec) ? A22 : A11; pAc = A12; pkk = kk;
        }
        double Ap = 0.0, Gp = 0.0, Qb = 0.0, Qa = 0.0, Qc = 0.0;
        double Vs = 0.0, Vm = 0.0;
        if (mb) { Vs = powd<NDIM>(x_self); Vm = haveP0 ? powd<NDIM>(x_prev) : Vs; }
        int s = s0;
#pragma unroll 1
        for (int k = 0; k < len; ++k, ++s) {
          const double x_next = X[s + 1];
          double A11, A12, A22, kk;
          if (!xdec) elem_mats_fast<NDIM>(x_self, x_next, Dph, idt, A11, A12, A22, kk);
          else elem_mats_fast<NDIM>(x_next, x_self, Dph, idt, A11, A12, A22, kk);
          const double nAself = (!xdec) ? A11 : A22, nAnext = (!xdec) ? A22 : A11;
          double diag = (nAself - kk) + (pAnext - pkk);
          double ml = pAc, md = nAself + pAnext, mu = A12;            // mass-matrix row / dt: rhs = ml*C- + md*C + mu*C+
          double sup = A12 + kk, sub = pAc + pkk;
          if (dir0 && f + k == 0) { diag = 1.0; sup = 0.0; sub = 0.0; ml = 0.0; md = 1.0; mu = 0.0; }   // set_outer_bc!
          const double inv = frcp(diag - sub * Gp);
          const double An = (k == 0) ? (sub * inv) : (-sub * Ap * inv);
          const double Gn = sup * inv;
          CO[7 * (s) + 5] = An; CO[7 * (s) + 4] = Gn; CO[7 * (s) + 3] = sub * inv; CO[7 * (s) + 0] = ml * inv; CO[7 * (s) + 1] = md * inv; CO[7 * (s) + 2] = mu * inv;
          if (mb) {
            const double Vn = powd<NDIM>(x_next);
            const double wgt = 0.5 * fabs(Vn - Vs) + 0.5 * fabs(Vs - Vm);
            if (k < len - 1) { const double tq = Qb + wgt; CO[7 * (s) + 6] = tq; Qa = fma(-tq, An, Qa); Qb = -tq * Gn; }
            else { Qc = Qb + wgt; CO[7 * (s) + 6] = 0.0; }
            Vm = Vs; Vs = Vn;
          }
          Ap = An; Gp = Gn;
          x_prev = x_self; x_self = x_next; pAc = A12; pAnext = nAnext; pkk = kk;
        }
        cQa = Qa; cQc = Qc; cGp = Gp; cAp = Ap;
        // backward accumulation of the matrix part (bA, bG): first row of the chunk in terms of the two separators
        cbA = 0.0; cbG = -1.0;
        if (len > 1) {
          int sb = s0 + (len - 2);
          cbA = CO[7 * (sb) + 5]; cbG = CO[7 * (sb) + 4];
#pragma unroll 1
          for (int k = len - 3; k >= 0; --k) { --sb; const double Gk = CO[7 * (sb) + 4]; cbA = CO[7 * (sb) + 5] - Gk * cbA; cbG = -Gk * cbG; }
        }
        exA[tid] = cbA; exG[tid] = cbG;
      }
      __syncthreads();
      if (active) {
        if (!lastChunk) {
          const double nA = exA[tid + 1], nG = exG[tid + 1];
          r_inv = frcp(1.0 - cGp * nA); r_s = cAp * r_inv; r_u = -cGp * nG * r_inv;
        } else { r_inv = 1.0; r_s = cAp; r_u = 0.0; }
        r_Gp = cGp;
      }
      // the interface-unknown column of the reduced system depends on the matrix only: solve it once per factorisation
      if (active) { redS[tid] = r_s; redD[tid] = 1.0; redU[tid] = lastChunk ? 0.0 : r_u; red0[tid] = 0.0; red1[tid] = lastChunk ? -r_Gp : 0.0; }
      else { redS[tid] = 0.0; redD[tid] = 1.0; redU[tid] = 0.0; red0[tid] = 0.0; red1[tid] = 0.0; }
      __syncthreads();
      // (exA / exG are free from here on: exA carries each chunk's r_Gp * r_inv to its right-hand neighbour)
      exA[tid] = (active && !lastChunk) ? r_Gp * r_inv : 0.0;
      if (tid < 32) { reduced_solve<T, 2>(redS, redD, redU, red0, red1); reduced_factor<T>(redS, redD, redU, rfac); }
      __syncthreads();
      red1c[tid] = red1[tid];
      gi_prev = (tid > 0) ? exA[tid - 1] : 0.0;
      __syncthreads();                         // redS | redD, redU, exR change roles: step-time arrays from here
      if (fastM) {
        hFirst[tid] = cst[0]; hLast[tid] = (fastM == 3) ? cst[2] : cst[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) gk[k] = (k < fastM - 1) ? co0[7 * k + 4] : 0.0;
      } else if (active) { hFirst[tid] = C[s0]; hLast[tid] = C[s0 + len - 1]; }
      redB[red_ix<T>(tid)] = 0.0;
      __syncthreads();
      if (bc == BC_FLUX) {
        // the flux row's coefficients only change with the factorisation: al*U + (be + al*U1)*a + (ga + de*W1)*b + de*W = 0
        __syncthreads();
        if (tid == 0) {
          const int p_ = n1 - 1, q_ = N - 1;
          const double idxl = frcp(X[p_] - X[p_ - 1]), idxr = frcp(X[q_ - 1] - X[q_]);
          const double al = -S.rho1 * S.D_l * idxl, be = S.V * S.rho1 + S.rho1 * S.D_l * idxl;
          const double ga = -S.V * S.rho2 + S.rho2 * S.D_r * idxr, de = -S.rho2 * S.D_r * idxr;
          S.cu0 = al; S.cu1 = de; S.cu2 = be + al * red1c[KL - 1]; S.cd0 = ga + de * red1c[KL + KR - 1];
        }
      }
    }

    // ---------------- right-hand side + forward substitution with the resident factors ----------------------------------
    // Two block barriers per step (flux / hold; the mass-balance row adds its block sum): chunks exchange their end values
    // through hFirst / hLast, so the sweep reads and writes only its own rows of C; the reduced right-hand side is
    // assembled from two contributions (the chunk's own: red0, its right-hand neighbour's: redB) without a barrier between
    // them.  Chunks of 3 / 5 full rows run straight-line code with the rows of the elimination in registers.
    const int p = n1 - 1, q = N - 1;         // interface nodes (the right one is the last element of the reversed block)
    const double a_old = C[p], b_old = C[q];
    double c_halo_prev = 0.0, c_halo_next = 0.0;
    if (active) {
      if (f > 0) c_halo_prev = hLast[tid - 1];
      c_halo_next = lastChunk ? (blk ? b_old : a_old) : hFirst[tid + 1];
    }
    const int rix = red_ix<T>(tid), rixm = red_ix<T>(tid > 0 ? tid - 1 : 0);
    double P0 = 0.0, Rp = 0.0, bR = 0.0;
    double Rr[5];                             // rows of the elimination of a 3 / 5-row chunk
    auto fwd_fast = [&](auto mtag) {
      constexpr int M = decltype(mtag)::value;
      double c[M + 2];
      c[0] = c_halo_prev; c[M + 1] = c_halo_next;
#pragma unroll
      for (int k = 0; k < M; ++k) c[k + 1] = cst[k];
      double rp = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        rp = fma(co0[7 * k + 0], c[k], fma(co0[7 * k + 1], c[k + 1], fma(co0[7 * k + 2], c[k + 2], -co0[7 * k + 3] * rp)));
        Rr[k] = rp;
        if (mb) P0 = fma(co0[7 * k + 6], rp, P0);
      }
      Rp = rp;
      double b = Rr[M - 2];
#pragma unroll
      for (int k = M - 3; k >= 0; --k) b = Rr[k] - gk[k] * b;
      bR = b;
    };
    if (fastM == 5) fwd_fast(std::integral_constant<int, 5>{});
    else if (fastM == 3) fwd_fast(std::integral_constant<int, 3>{});
    else if (active) {
      double c_prev = c_halo_prev, c_self = C[s0];
      int s = s0;
      // (two copies of the loop: the mass-balance functional is only accumulated for inner_bc = mb, and a test of a
      //  kernel parameter inside the loop is re-read from the constant bank every row)
      if (mb) {
#pragma unroll 2
        for (int k = 0; k < len; ++k, ++s) {
          const double c_next = (k == len - 1) ? c_halo_next : C[s + 1];
          const double Rn = fma(CO[7 * (s) + 0], c_prev, fma(CO[7 * (s) + 1], c_self, fma(CO[7 * (s) + 2], c_next, -CO[7 * (s) + 3] * Rp)));
          C[s] = Rn;
          P0 = fma(CO[7 * (s) + 6], Rn, P0);
          Rp = Rn; c_prev = c_self; c_self = c_next;
        }
      } else {
#pragma unroll 2
        for (int k = 0; k < len; ++k, ++s) {
          const double c_next = (k == len - 1) ? c_halo_next : C[s + 1];
          const double Rn = fma(CO[7 * (s) + 0], c_prev, fma(CO[7 * (s) + 1], c_self, fma(CO[7 * (s) + 2], c_next, -CO[7 * (s) + 3] * Rp)));
          C[s] = Rn;
          Rp = Rn; c_prev = c_self; c_self = c_next;
        }
      }
      if (len > 1) {
        int sb = s0 + (len - 2);
        bR = C[sb];
#pragma unroll 1
        for (int k = len - 3; k >= 0; --k) { --sb; bR = C[sb] - CO[7 * (sb) + 4] * bR; }
      }
    }
    // right-hand side of the reduced system, row t = (Rp_t - r_Gp_t bR_{t+1}) r_inv_t in two parts: red0[t] = Rp_t r_inv_t by
    // chunk t, redB[t] = -(r_Gp r_inv)_{t-1} bR_t by chunk t for row t - 1; its matrix part was factorised with the step
    // matrix (reduced_factor)
    red0[rix] = active ? Rp * r_inv : 0.0;
    redB[rix] = -gi_prev * bR;
    __syncthreads();
    if (tid < 32) reduced_apply<T, true, true>(rfac, red0, redB);
    __syncthreads();

    // ---------------- interface condition (set_inner_bc_flux! / _mb! / Dirichlet hold) ----------------------------------
    double aif, bif;
    if (bc == BC_FLUX) {
      const double al = S.cu0, de = S.cu1, ca = S.cu2, cb = S.cd0;
      bif = -(al * red0[red_ix<T>(KL - 1)] + de * red0[red_ix<T>(KL + KR - 1)]) * frcp(ca * KD_k + cb);
      aif = KD_k * bif;
    } else if (mb) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      if (active) {
        const double s0v = red0[rix], s1v = red1c[tid];
        const double q0 = (jc > 0) ? red0[rixm] : 0.0, q1 = (jc > 0) ? red1c[tid - 1] : 0.0;
        const double c0 = P0 + cQa * q0 + cQc * s0v, c1 = cQa * q1 + cQc * s1v;
        if (blk == 0) { v[0] = c0; v[1] = c1; } else { v[2] = c0; v[3] = c1; }
      }
      block_sum4<T>(v, S.red);
      const double Vp_ = powd<NDIM>(X[p]), Vq_ = powd<NDIM>(X[q]);
      const double wp = 0.5 * (Vq_ - Vp_) + 0.5 * (Vp_ - powd<NDIM>(X[p - 1]));
      const double wq = 0.5 * (powd<NDIM>(X[q - 1]) - Vq_) + 0.5 * (Vq_ - Vp_);
      bif = (S.Mass0 - (v[0] + v[2])) * frcp((v[1] + wp) * KD_k + (v[3] + wq));
      aif = KD_k * bif;
    } else { aif = a_old; bif = b_old; }
    // ---------------- back-substitution -----------------------------------------------------------------------------------
    if (active) {
      const double beta = blk ? bif : aif;
      const double sj = red0[rix] + red1c[tid] * beta;
      const double sp = (jc > 0) ? (red0[rixm] + red1c[tid - 1] * beta) : 0.0;
      double cn = sj;
      auto back_fast = [&](auto mtag) {
        constexpr int M = decltype(mtag)::value;
        cst[M - 1] = cn;
#pragma unroll
        for (int k = M - 2; k >= 0; --k) { cn = Rr[k] - co0[7 * k + 5] * sp - gk[k] * cn; cst[k] = cn; }
        if (rec_on) {
#pragma unroll
          for (int k = 0; k < M; ++k) c0p[k] = cst[k];
        }
      };
      if (fastM == 5) back_fast(std::integral_constant<int, 5>{});
      else if (fastM == 3) back_fast(std::integral_constant<int, 3>{});
      else {
        int sb = s0 + (len - 1);
        C[sb] = cn;
#pragma unroll 2
        for (int k = len - 2; k >= 0; --k) { --sb; cn = C[sb] - CO[7 * (sb) + 5] * sp - CO[7 * (sb) + 4] * cn; C[sb] = cn; }
      }
      hLast[tid] = sj; hFirst[tid] = cn;
      // finite check of the step (status 11) on the outer ends, by the chunks that own them
      if (jc == 0 && !isfinite(cn)) S.status = 11;
    }
    if (tid == 0) {
      C[p] = aif; C[q] = bif;
      if (!isfinite(aif)) S.status = 11;
      S.steps += 1;
    }
    __syncthreads();
    if (rec_on) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n1 - 1; i += T)
        v[0] += 0.5 * (S.rho1 * powd<NDIM>(X[i + 1]) - S.rho1 * powd<NDIM>(X[i])) * (C[i] + C[i + 1]);
      for (int i = tid; i < n2 - 1; i += T)
        v[1] += 0.5 * (S.rho2 * powd<NDIM>(X[N - 2 - i]) - S.rho2 * powd<NDIM>(X[N - 1 - i])) * (C[N - 1 - i] + C[N - 2 - i]);
      block_sum4<T>(v, S.red);
      if (tid == 0 && S.it >= 1 && S.it - 1 < A.rec_cap) {
        double *r = A.rec + ((size_t)mem * (size_t)A.rec_cap + (size_t)(S.it - 1)) * 4;
        r[0] = v[0] + v[1]; r[1] = S.T; r[2] = S.KD; r[3] = S.Ri1;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) if (k < fastM) c0p[k] = cst[k];
  __syncthreads();
  // ---------------- store member ----------------
  {
    double *gcl = A.C_left + (size_t)mem * A.capL, *gcr = A.C_right + (size_t)mem * A.capR;
    for (int i = tid; i < n1; i += T) gcl[i] = C[i];
    for (int i = tid; i < n2; i += T) gcr[i] = C[N - 1 - i];
    if ((A.flags & F_MASS_FINAL) && A.Mass) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n1 - 1; i += T)
        v[0] += 0.5 * (S.rho1 * powd<NDIM>(X[i + 1]) - S.rho1 * powd<NDIM>(X[i])) * (C[i] + C[i + 1]);
      for (int i = tid; i < n2 - 1; i += T)
        v[1] += 0.5 * (S.rho2 * powd<NDIM>(X[N - 2 - i]) - S.rho2 * powd<NDIM>(X[N - 1 - i])) * (C[N - 1 - i] + C[N - 2 - i]);
      block_sum4<T>(v, S.red);
      if (tid == 0) A.Mass[mem] = v[0] + v[1];
    }
    if (tid == 0) {
      A.status[mem] = S.status;
      if (A.steps_done) A.steps_done[mem] = S.steps;
      if (A.warn) A.warn[mem] = S.warn;
      A.t[mem] = S.t; A.dt[mem] = S.dt; A.it[mem] = S.it;
      if (A.dx1) A.dx1[mem] = S.dx1;
      if (A.dx2) A.dx2[mem] = S.dx2;
      if (A.last) {
        double *L = A.last + 8 * (size_t)mem;
        L[0] = S.D_l; L[1] = S.D_r; L[2] = S.KD; L[3] = S.T; L[4] = S.V; L[5] = S.rho1; L[6] = S.rho2; L[7] = 0.0;
      }
      if (A.node_steps) A.node_steps[mem] = S.node_steps;
    }
  }
}

}  // namespace mbm
