/*
 * mbm_b200.h — C ABI of libmbm_b200.so (CUDA, sm_100a) and of oracle/libmbm_oracle.so
 * (CPU restatement, test infrastructure only).  Both libraries export exactly the
 * symbols declared here, with identical struct layouts, so every test can drive either.
 *
 * The reference (AnStroh/MovingBoundaryMinerals.jl) has no FFI seam; the seam is cut at
 * the level of the driver main functions, whose `while t < t_tot ... end` loops
 * (reference main_codes/<driver>.jl) are swallowed whole for a *batch* ("ensemble") of
 * independent members.  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - all arrays are caller-owned, column-major, member index = LAST dimension
 *     (x_left[cap_left, M]: member m starts at x_left + m*cap_left);
 *   - arrays are HOST memory unless MBM_FLAG_DEVICE_POINTERS is set, in which case every
 *     pointer in params/io is device memory on device_ids[0] (n_devices must be 1);
 *   - inputs are already non-dimensionalised by the host shim with the reference's own
 *     `scaling` convention (reference src/functions.jl:988-1008);
 *   - no exception crosses the boundary: return value 0 = call executed (inspect
 *     status[]), <0 = library-level failure; per-member status mirrors the reference's
 *     throw sites, a failed member is frozen and the others continue;
 *   - re-entrant: concurrent calls from different host threads are independent.
 */
#ifndef MBM_B200_H
#define MBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBM_ABI_VERSION 4

/* ---- interface condition (which `set_inner_bc_*` the step uses) -------------------- */
enum {
  MBM_BC_NONE           = 0, /* single phase: Simple_Diff_1D_general.jl:67-99                 */
  MBM_BC_FLUX           = 1, /* set_inner_bc_flux!   src/functions.jl:1054-1072               */
  MBM_BC_MB             = 2, /* set_inner_bc_mb!     src/functions.jl:1094-1109               */
  MBM_BC_STEFAN         = 3, /* set_inner_bc_stefan! src/functionsStefan.jl:131-144           */
  MBM_BC_DIRICHLET_HOLD = 4, /* inline rows, examples/Diff_couple_no_interaction.jl:108-118   */
  MBM_BC_LASAGA = 5          /* set_inner_bc_Lasaga!, src/functions.jl:1143-1206 (major-element exchange after Lasaga 1983):
                                both interface values prescribed from the step's OLD neighbour values, KD, D, rho and the
                                two interface spacings (closed-form roots sol1 / sol2 with the reference's own selection rule,
                                including sol2's product where sol1 divides, :1187); the reference exports the function but
                                ships no driver for it: the loop is DCGF's with this row function.  pow(.,1/2) is sqrt.   */
};

/* ---- time-step rule ----------------------------------------------------------------- */
enum {
  MBM_DT_FIND_DT   = 0, /* find_dt, src/functions.jl:465-490                                  */
  MBM_DT_ADVECTIVE = 1, /* dt = minimum([dx1,dx2])/V_ip*CFL, examples/C1.jl:93, B4.jl:93      */
  MBM_DT_STEFAN    = 2, /* Chemical_Stefan_problem_XT.jl:79,147-149 (dt carried in io.dt)     */
  MBM_DT_SIMPLE    = 3  /* calculate_dt, src/functions.jl:164-168                             */
};

/* ---- in-loop material laws (Stefan driver), mbm_couple_params.law ------------------------ */
enum {
  MBM_LAW_NONE = 0,            /* D_l, D_r stay params.Di (Chemical_Stefan_problem_XT.jl:63-64)        */
  /* examples/D1.jl:126-132: every step, after the interface compositions are set,
   *   log10D = law_par[0] - law_par[1]/(2.303*R*T) + 3*(mean(C_left) - 0.14)        (Dohmen & Chakraborty 2007)
   *   D_l    = 10^log10D*law_par[2] + 10^(log10D - log10(6))*law_par[3] + 10^(log10D - log10(6))*law_par[4]
   *   D_r    = exp(-7.92 - 26222/T)                                                  (Zhang & Cherniak 2010)
   * law_par = { log10(D0), Ea + (P-1e5)*deltaV, cos(alpha)^2, cos(beta)^2, cos(gamma)^2, 0, 0, 0 };
   * mean(C_left) = canonical sum / node count.                                      */
  MBM_LAW_OLIVINE_MELT_D1 = 1
};

/* ---- flags ---------------------------------------------------------------------------- */
#define MBM_FLAG_NEWTON_WARM      0x0001u /* start newton_solver from the previous root instead of 1.1
                                             (reference cold-starts every regrid: functions.jl:734)      */
#define MBM_FLAG_RECORD           0x0002u /* fill io.rec (per-step Mass/T/KD/Ri history)                 */
#define MBM_FLAG_SOLVER_FAITHFUL  0x0004u /* ORACLE ONLY: dense partial-pivot LU like `L_g \ R_g`
                                             (functions.jl:1292) instead of the O(N) structured solve   */
#define MBM_FLAG_SUM_JULIA        0x0008u /* ORACLE ONLY: Julia-like pairwise cumsum + sequential
                                             make_dx_right instead of the canonical order shared with
                                             the CUDA kernels ("canonical summation order")   */
#define MBM_FLAG_DEVICE_POINTERS  0x0010u /* CUDA ONLY: params/io arrays are device pointers             */
#define MBM_FLAG_FORCE_RESIDENT   0x0020u /* CUDA ONLY: force the shared-memory resident kernel          */
#define MBM_FLAG_FORCE_STREAMING  0x0040u /* CUDA ONLY: force the global-memory kernel                   */
#define MBM_FLAG_MASS_FINAL       0x0080u /* compute io.Mass after the last executed step
                                             (calc_mass_vol, functions.jl:205-212)                       */

/* ---- per-member status (mirrors the reference's error() sites) ---------------------- */
enum {
  MBM_OK                    = 0,
  MBM_ERR_IFACE_FAST_POS    = 1,  /* functions.jl:39  "Ri is larger than x_right[2]"          */
  MBM_ERR_IFACE_FAST_NEG    = 2,  /* functions.jl:52  "Ri is smaller than x_left[end-1]"      */
  MBM_ERR_NEWTON            = 3,  /* functions.jl:752 "Newton Failure"                        */
  MBM_ERR_KD_RANGE          = 4,  /* functions.jl:1362                                        */
  MBM_ERR_T_RANGE           = 5,  /* functions.jl:1365, :1412                                 */
  MBM_ERR_INTERP_BOUNDS     = 6,  /* functions.jl:571                                         */
  MBM_ERR_NDIM              = 7,  /* functions.jl:513                                         */
  MBM_ERR_PCHIP_INDEX       = 8,  /* functions.jl:822                                         */
  MBM_ERR_DENSITY_BOUNDS    = 9,  /* functionsStefan.jl:97-98                                 */
  MBM_ERR_CAPACITY          = 10, /* node count would exceed cap_left / cap_right             */
  MBM_ERR_NONFINITE         = 11, /* NaN/Inf in the state                                     */
  MBM_ERR_MREFIN            = 12  /* MRefin < 1 branch (functions.jl:412-426) is unsupported  */
};
#define MBM_WARN_NEWTON_RES 0x1   /* Newton residual in (1e-8, 1e-3] after 200 its, functions.jl:754 */

/* library-level return codes */
enum {
  MBM_RC_OK = 0, MBM_RC_BAD_ARG = -1, MBM_RC_CUDA = -2, MBM_RC_OOM = -3, MBM_RC_UNSUPPORTED = -4,
  MBM_RC_NO_DEVICE = -5
};

/* ---- argument blocks -------------------------------------------------------------------- */

/* replaces the "Numerics"/"Physics" literals at the top of every driver main function,
 * e.g. main_codes/Diff_couple_1D_general_MB.jl:7-33                                          */
typedef struct mbm_common {
  int32_t abi_version;   /* MBM_ABI_VERSION                                                   */
  int32_t n_members;     /* M                                                                 */
  int32_t ndim;          /* `n`: 1 planar, 2 cylindrical, 3 spherical                         */
  int32_t hist_len;      /* length of t_ar / T_ar / KD_ar (Stefan: tpath / Tpath)             */
  int32_t cap_left;      /* leading dimension of x_left/C_left (>= res[1]+1)                  */
  int32_t cap_right;     /* leading dimension of x_right/C_right (>= res[2]+1); 0 for simple  */
  int32_t inner_bc;      /* MBM_BC_*                                                          */
  int32_t dt_mode;       /* MBM_DT_*                                                          */
  uint32_t flags;        /* MBM_FLAG_*                                                        */
  int32_t n_devices;     /* CUDA: number of GPUs to shard members over (contiguous blocks)    */
  int64_t max_steps;     /* 0 = run every member to t_tot; else at most this many steps/call  */
  const int32_t *device_ids; /* [n_devices] or NULL = 0..n_devices-1                          */
  int64_t rec_cap;       /* leading dimension of io.rec (steps recorded per member)           */
} mbm_common;

/* Diff_couple_1D_general_{MB,Flux}.jl:7-33 and Chemical_Stefan_problem_XT.jl:8-26, one
 * column per member.  A member stride of 0 means "one array shared by all members".        */
typedef struct mbm_couple_params {
  const double *Di;      /* [2,M] scaled; Di == [-1 -1] selects Arrhenius (functions.jl:1368) */
  const double *D0;      /* [2,M]                                                             */
  const double *Ea;      /* [2,M] Ea1, Ea2                                                    */
  const double *rho;     /* [2,M]                                                             */
  const double *V_ip;    /* [M]   (ignored for MBM_BC_STEFAN: computed from the fluxes)       */
  const double *t_tot;   /* [M]                                                               */
  const double *CFL;     /* [M]                                                               */
  const double *MRefin;  /* [M]                                                               */
  const int32_t *BCout;  /* [2,M] 1 = Dirichlet, 0 = Neumann                                  */
  const int64_t *resmin; /* [2,M]                                                             */
  const double *t_ar;    /* [hist_len, M or 1]                                                */
  const double *T_ar;    /* [hist_len, M or 1]                                                */
  const double *KD_ar;   /* [hist_len, M or 1] (unused for Stefan)                            */
  int64_t t_stride, T_stride, KD_stride; /* elements between members; 0 = shared             */
  double R;              /* gas constant                                                      */
  /* --- Stefan only (NULL otherwise) --- */
  const double *coeff_up;   /* [3,M or 1] composition(): X_right(T), functionsStefan.jl:69    */
  const double *coeff_do;   /* [3,M or 1] composition(): X_left(T),  functionsStefan.jl:68    */
  int64_t coeff_stride;     /* 0 = shared                                                     */
  const double *C_leftB;    /* [M] initial equilibrium compositions used by calculate_density */
  const double *C_rightB;   /* [M]                                                            */
  int32_t nd;               /* density table is nd x nd; 0 = no table, rho stays params.rho   */
  int32_t law;              /* MBM_LAW_*: in-loop material law of the Stefan driver (0 = none)   */
  const double *dens_X;     /* [nd]     Xwm[:,1]                                              */
  const double *dens_T;     /* [nd]     Twm[1,:]                                              */
  const double *dens_left;  /* [nd,nd]  column-major                                          */
  const double *dens_right; /* [nd,nd]                                                        */
  const double *law_par;    /* [8,M] parameters of `law` (NULL for MBM_LAW_NONE), see MBM_LAW_*  */
} mbm_couple_params;

/* in/out state of a couple ensemble; calling again with the returned state resumes the run
 * (the reference keeps this state in local variables of the driver function).               */
typedef struct mbm_couple_io {
  double  *x_left;   /* [cap_left,  M]                                                        */
  double  *C_left;   /* [cap_left,  M]                                                        */
  double  *x_right;  /* [cap_right, M]                                                        */
  double  *C_right;  /* [cap_right, M]                                                        */
  int64_t *res;      /* [2,M] node counts (reference mutates `res`, functions.jl:44,57,874)   */
  double  *Ri;       /* [2,M] (reference mutates Ri[1], functions.jl:31)                      */
  double  *t;        /* [M]                                                                   */
  double  *dt;       /* [M] last dt; for MBM_DT_STEFAN the dt of the next step (in/out)       */
  int64_t *it;       /* [M] step counter                                                      */
  double  *dx1;      /* [M] out                                                               */
  double  *dx2;      /* [M] out                                                               */
  double  *Mass0;    /* [M] in  (calc_mass_vol at t = 0, e.g. ..._MB.jl:62)                   */
  double  *Mass;     /* [M] out (MBM_FLAG_MASS_FINAL) or NULL                                 */
  double  *newton_R; /* [M] in/out last geometric ratio (0 = none yet) or NULL                */
  double  *last;     /* [8,M] out: D_l, D_r, KD, T, V_ip, rho1, rho2, Fl_regrid; or NULL      */
  int32_t *status;   /* [M] in/out: members with status != 0 on entry are skipped             */
  int32_t *warn;     /* [M] out bitmask                                                       */
  uint64_t *idx_hash;/* [M] in/out order-independent checksum of every regrid interval index
                              and pchip branch taken so far, or NULL                          */
  int64_t *node_steps;/*[M] in/out sum over executed steps of the node count solved           */
  double  *rec;      /* [4, rec_cap, M] per-step (Mass, T, KD, Ri[1]) if MBM_FLAG_RECORD      */
} mbm_couple_io;

/* Simple_Diff_1D_general.jl:6-31 */
typedef struct mbm_simple_params {
  const double *Di;     /* [M]; -1 selects Arrhenius (functions.jl:1415)                      */
  const double *D0;     /* [M]                                                                */
  const double *Ea1;    /* [M]                                                                */
  const double *rho;    /* [M]                                                                */
  const double *dx;     /* [M] spacing used by calculate_dt (L*inv(res-1.0))                  */
  const double *t_tot;  /* [M]                                                                */
  const double *CFL;    /* [M]                                                                */
  const int32_t *BCout; /* [2,M]                                                              */
  const double *t_ar;   /* [hist_len, M or 1]                                                 */
  const double *T_ar;   /* [hist_len, M or 1]                                                 */
  int64_t t_stride, T_stride;
  double R;
} mbm_simple_params;

typedef struct mbm_simple_io {
  double  *x;        /* [cap_left, M] (fixed mesh)                                            */
  double  *C;        /* [cap_left, M]                                                         */
  int64_t *res;      /* [M]                                                                   */
  double  *t;        /* [M]                                                                   */
  double  *dt;       /* [M]                                                                   */
  int64_t *it;       /* [M]                                                                   */
  double  *D;        /* [M] out: last diffusivity                                             */
  double  *Mass;     /* [M] out (MBM_FLAG_MASS_FINAL; calc_mass_vol_simple_diff) or NULL      */
  int32_t *status;   /* [M]                                                                   */
  int64_t *node_steps;/*[M]                                                                   */
} mbm_simple_io;

/* ---- entry points ------------------------------------------------------------------------ */
int mbm_version(void);
int mbm_device_count(void);          /* oracle: 0                                             */
const char *mbm_backend(void);       /* "cuda-sm_100a" or "oracle-cpu"                        */
const char *mbm_status_string(int32_t status);

/* SDG(): main_codes/Simple_Diff_1D_general.jl:4-103 (also examples/A1.jl, A2.jl)             */
int mbm_simple_diff_batch(const mbm_common *, const mbm_simple_params *, mbm_simple_io *);
/* DCGF(): main_codes/Diff_couple_1D_general_Flux.jl:4-138 (B1..B7);
 * with inner_bc = MBM_BC_DIRICHLET_HOLD also DCNI(), examples/Diff_couple_no_interaction.jl  */
int mbm_couple_flux_batch(const mbm_common *, const mbm_couple_params *, mbm_couple_io *);
/* DCGMB(): main_codes/Diff_couple_1D_general_MB.jl:4-140 (C1, C2, Diff_couple_MB_growth)     */
int mbm_couple_mb_batch(const mbm_common *, const mbm_couple_params *, mbm_couple_io *);
/* CSP(): main_codes/Chemical_Stefan_problem_XT.jl:4-208 (D1)                                 */
int mbm_stefan_xt_batch(const mbm_common *, const mbm_couple_params *, mbm_couple_io *);
/* generic form of the three couple entry points: inner_bc taken from common                 */
int mbm_couple_batch(const mbm_common *, const mbm_couple_params *, mbm_couple_io *);

/* pre-loop helpers of the drivers, batched (host side of both libraries is plain CPU code;
 * they run once per run and are not on the hot path):
 * create_grid!: functions.jl:364-370 -> define_new_grid :395-441                             */
int mbm_create_grid(int32_t n_members, int32_t cap_left, int32_t cap_right,
                    const double *Ri /*[2,M]*/, const int64_t *res /*[2,M]*/,
                    const double *MRefin /*[M]*/, double *x_left, double *x_right,
                    double *dx1, double *dx2, double *newton_R /*[M] or NULL*/,
                    int32_t *status /*[M]*/);
/* calc_mass_vol: functions.jl:205-212 */
int mbm_calc_mass_vol(int32_t n_members, int32_t ndim, int32_t cap_left, int32_t cap_right,
                      const double *x_left, const double *x_right, const double *C_left,
                      const double *C_right, const int64_t *res, const double *rho,
                      double *Mass /*[M]*/);

/* ---- validation layer ---------------------------------------------
 * The known-answer gates of the reference's own tests, evaluated per member from the ensemble's
 * state so that a sweep of thousands of members is checked where it lives (MBM_FLAG_DEVICE_POINTERS:
 * nothing but [M] results crosses the host link).  Quantities are in the units of the state arrays
 * (the gates are ratios, invariant under the `scaling` of functions.jl:988-1008).
 *   MBM_VAL_MASS_ERR  calc_mass_err (functions.jl:183-187) on calc_mass_vol (:205-212) of the state:
 *                     out = (Mass, Mass0), relerr = (Mass - Mass0) * inv(Mass0)   [signed, as the reference]
 *   MBM_VAL_RAYLEIGH  test/test_C1.jl: rayleigh_fractionation (Benchmarks.jl:239-260, with the as-shipped
 *                     LinearIndices post-processing of :250-258) and the two trapezoidal integrals;
 *                     out = (int_Ray, int_model), relerr = |a - b| / max(|a|, |b|)  [isapprox, rtol]
 *   MBM_VAL_SMITH     test/test_B5.jl: smith (Benchmarks.jl:284-301, n = 1) against C_right;
 *                     out = (||Can - C_right||_2, max(||Can||_2, ||C_right||_2)), relerr = their ratio
 *   MBM_VAL_SPHERE    test/test_B1.jl and test/test_B2.jl: the eigenfunction series of diffusion in a sphere (Crank 1975,
 *                     ch. 6.3) summed over 1000 terms as calc_sinus_sphere (Benchmarks.jl:82-99) and
 *                     crank_time_transformation3 (:192-212) do, on x0 = [x_left; x_right] with the initial profile
 *                     C0 = [C0_left; C0_right], against C = [C_left; C_right].  zeta[m] is the time-integrated
 *                     diffusivity: D[1] * t for B1 (:94), trapezoidal_integration(tan, Dan) for B2 (:194-197; a
 *                     1000-point scalar integral the host evaluates);
 *                     out = (||Can - C||_2, max(||Can||_2, ||C||_2)), relerr = their ratio   [isapprox, rtol 1e-2] */
enum { MBM_VAL_MASS_ERR = 0, MBM_VAL_RAYLEIGH = 1, MBM_VAL_SMITH = 2, MBM_VAL_SPHERE = 3 };
typedef struct mbm_validate_args {
  int32_t abi_version;   /* MBM_ABI_VERSION                                                   */
  int32_t n_members, ndim, cap_left, cap_right, kind;
  uint32_t flags;        /* MBM_FLAG_DEVICE_POINTERS: every pointer below is device memory     */
  int32_t device;        /* CUDA device to run on                                             */
  const double *x_left, *C_left, *x_right, *C_right;   /* state, as in mbm_couple_io          */
  const int64_t *res;    /* [2,M]                                                             */
  const double *Ri;      /* [2,M]                                                             */
  const double *rho;     /* [2,M]  MASS_ERR                                                   */
  const double *Mass0;   /* [M]    MASS_ERR                                                   */
  const double *Ri0;     /* [2,M]  RAYLEIGH: radii at t = 0                                   */
  const double *C0_r;    /* [M]    RAYLEIGH: C0_r[end], initial composition of the right phase */
  const double *KD0;     /* [M]    RAYLEIGH, SMITH                                            */
  const double *Di_r;    /* [M]    SMITH: Di[2]                                               */
  const double *V_ip;    /* [M]    SMITH                                                      */
  const double *t_tot;   /* [M]    SMITH: the time reached                                    */
  double *out;           /* [2,M]  see the kind                                               */
  double *relerr;        /* [M]                                                               */
  const double *zeta;    /* [M]    SPHERE: time-integrated diffusivity (ABI 4)                */
  const double *C0_left, *C0_right;   /* [cap_left,M], [cap_right,M]  SPHERE: initial profiles */
} mbm_validate_args;
int mbm_validate_batch(const mbm_validate_args *);

/* CUDA library only: seconds spent in the last call's kernels on device_ids[0], measured
 * with CUDA events on the launching stream (0 for the oracle), and the number of kernel
 * launches of the last call.                                                               */
double mbm_last_kernel_seconds(void);
int64_t mbm_last_kernel_launches(void);

/* CUDA library only (oracle: no-ops returning 0).
 * mbm_set_stream: launch the kernels of subsequent calls made by THIS host thread on the caller's
 * cudaStream_t (e.g. torch's current stream) instead of a library-owned stream.  With
 * MBM_FLAG_DEVICE_POINTERS the call then returns without synchronising, so the caller can bracket
 * it with its own CUDA events.  NULL restores the default.
 * mbm_measure_dfma_tflops: FP64 FMA micro-benchmark on `device` (the denominator of the FP64
 * roofline; MEASURED_PEAKS.json has no FP64 entry).                                           */
int mbm_set_stream(void *cuda_stream);
double mbm_measure_dfma_tflops(int device);

/* CUDA library only (oracle: zeros).  Diagnostics of the streaming kernel for long uniform meshes
 * (MRefin == 1, closed-form meshes), counted on `device` since the library was loaded: member-steps that took the fused
 * back-substitution + regrid pass / the two-pass path (node counts or layout changed by regrid!,
 * functions.jl:873-889; fixed-interface steps), and the chunk rows (summed over member-steps) that were stored
 * without their fill-in value (below 1e-19).  Tests use it to check which path ran.  Any pointer may be NULL.
 * Returns 0, or MBM_RC_CUDA.                                                                          */
int mbm_stream_pass_counts(int device, uint64_t *fused_steps, uint64_t *twopass_steps, uint64_t *rows_without_fill_in);

/* CUDA library only (oracle: 0).  Combined node capacity (left | free slot | right) the tuned resident kernel was launched
 * with in the last couple call of this process, 0 if that call did not use it.  Two CTAs per SM hold up to ~2 944 nodes
 * each, one CTA per SM up to ~6 500: tests use it to check which plan a shard took.                              */
int mbm_last_tuned_capacity(void);

#ifdef __cplusplus
}
#endif
#endif /* MBM_B200_H */
