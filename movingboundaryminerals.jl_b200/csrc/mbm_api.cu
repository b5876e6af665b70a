// mbm_api.cu — C ABI of libmbm_b200.so (include/mbm_b200.h): argument checking, host<->device staging,
// member sharding over GPUs (contiguous blocks, no collective: members never communicate), kernel
// dispatch.  There is no CPU compute path in this library: without a CUDA device every entry point
// returns MBM_RC_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mbm_b200.h"
#include "mbm_kernel_stream.cuh"
#include "mbm_kernel_stream2.cuh"
#include "mbm_validate.cuh"

namespace {

thread_local cudaStream_t g_user_stream = nullptr;
thread_local bool g_have_user_stream = false;
thread_local double g_last_kernel_seconds = 0.0;
thread_local long long g_last_kernel_launches = 0;
// last-call metrics readable from any host thread (bench / smoke): atomics, written once per call
std::atomic<double> g_last_kernel_seconds_any{0.0};
std::atomic<long long> g_last_kernel_launches_any{0};
std::atomic<int> g_last_tuned_capacity{0};        // kcapL of the last launch of the tuned kernel (0: not used)

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      fprintf(stderr, "[mbm_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__,   \
              __LINE__, cudaGetErrorString(e_));                                                   \
      return (e_ == cudaErrorMemoryAllocation) ? MBM_RC_OOM : MBM_RC_CUDA;                         \
    }                                                                                              \
  } while (0)

// launch-site check that says what failed (no silent MBM_RC_CUDA)
int launch_rc(const char *what, cudaError_t e) {
  if (e == cudaSuccess) return 0;
  fprintf(stderr, "[mbm_b200] %s failed: %s (%s)\n", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return MBM_RC_CUDA;
}

// make `device` current for the lifetime of the object, then restore the caller's device
struct DeviceScope {
  int prev = -1, rc = 0;
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (cudaSetDevice(device) != cudaSuccess) rc = MBM_RC_CUDA;
  }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

// staging of one shard: device copies of host arrays, copied back on request
// Host-buffer calls are cut into pieces that run on their own streams (one host thread each).  Left alone, all pieces
// start their host-to-device copies at once, share the link, finish together and only then run their kernels: nothing
// overlaps.  PieceOrder staggers them: the copies of piece q wait (on the device) for the copies of piece q-1, so piece
// 0's kernel starts after 1/P of the input has arrived and every later copy - and every earlier piece's copy back -
// runs beside a kernel.
struct PieceOrder {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<int> turn;                 // per device group: the piece whose turn it is to enqueue its input copies
  std::vector<cudaEvent_t> ev;           // per job: recorded behind the job's input copies
  std::vector<int> group, index;         // per job: device group and position inside it
  void begin(size_t job, cudaStream_t s) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return turn[group[job]] == index[job]; });
    if (index[job] > 0 && ev[job - 1]) cudaStreamWaitEvent(s, ev[job - 1], 0);
  }
  void end(size_t job, cudaStream_t s) {
    if (ev[job]) cudaEventRecord(ev[job], s);
    { std::lock_guard<std::mutex> lk(mu); turn[group[job]] = index[job] + 1; }
    cv.notify_all();
  }
  // after the shard returned: pass the turn on if the shard never took it (error before its copies, or a shard kind
  // that does not stagger), so that later pieces cannot wait for ever
  void finish(size_t job) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return turn[group[job]] >= index[job]; });
    if (turn[group[job]] == index[job]) { turn[group[job]] = index[job] + 1; lk.unlock(); cv.notify_all(); }
  }
};
thread_local PieceOrder *tl_order = nullptr;
thread_local size_t tl_job = 0;
struct PieceTurn {                       // scope of a shard's input copies
  cudaStream_t s; bool open;
  explicit PieceTurn(cudaStream_t st) : s(st), open(tl_order != nullptr) { if (open) tl_order->begin(tl_job, s); }
  void done() { if (open) { tl_order->end(tl_job, s); open = false; } }
  ~PieceTurn() { done(); }
};

struct Stage {
  struct Item { void *h; void *d; size_t bytes; bool out; };
  std::vector<Item> items;
  std::vector<void *> owned;
  cudaStream_t stream = nullptr;
  bool device_ptrs = false;
  int rc = 0;

  template <typename Tp>
  Tp *in(const Tp *h, size_t count, bool out = false) {
    if (!h || count == 0) return nullptr;
    if (device_ptrs) return const_cast<Tp *>(h);
    void *d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, count * sizeof(Tp), stream);
    if (e != cudaSuccess) { rc = (e == cudaErrorMemoryAllocation) ? MBM_RC_OOM : MBM_RC_CUDA; return nullptr; }
    owned.push_back(d);
    e = cudaMemcpyAsync(d, h, count * sizeof(Tp), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { rc = MBM_RC_CUDA; fprintf(stderr, "[mbm_b200] host-to-device copy failed: %s\n", cudaGetErrorString(e)); return nullptr; }
    items.push_back({(void *)h, d, count * sizeof(Tp), out});
    return (Tp *)d;
  }
  template <typename Tp>
  Tp *inout(Tp *h, size_t count) { return in<Tp>(h, count, true); }
  template <typename Tp>
  Tp *outonly(Tp *h, size_t count) {
    if (!h || count == 0) return nullptr;
    if (device_ptrs) return h;
    void *d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, count * sizeof(Tp), stream);
    if (e != cudaSuccess) { rc = (e == cudaErrorMemoryAllocation) ? MBM_RC_OOM : MBM_RC_CUDA; return nullptr; }
    owned.push_back(d);
    e = cudaMemsetAsync(d, 0, count * sizeof(Tp), stream);
    if (e != cudaSuccess) { rc = MBM_RC_CUDA; fprintf(stderr, "[mbm_b200] device memset failed: %s\n", cudaGetErrorString(e)); return nullptr; }
    items.push_back({(void *)h, d, count * sizeof(Tp), true});
    return (Tp *)d;
  }
  void *scratch(size_t bytes) {
    void *d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, bytes, stream);
    if (e != cudaSuccess) { rc = (e == cudaErrorMemoryAllocation) ? MBM_RC_OOM : MBM_RC_CUDA; return nullptr; }
    owned.push_back(d);
    return d;
  }
  int copy_back() {
    for (auto &it : items)
      if (it.out) {
        cudaError_t e = cudaMemcpyAsync(it.h, it.d, it.bytes, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return launch_rc("device-to-host copy", e);
      }
    return 0;
  }
  void release() {
    for (void *p : owned) cudaFreeAsync(p, stream);
    owned.clear();
    items.clear();
  }
};

int threads_for(int capS) {
  if (const char *e = getenv("MBM_THREADS")) {
    int t = atoi(e);
    if (t == 32 || t == 64 || t == 128 || t == 256 || t == 512) return t;
  }
  if (capS <= 160) return 32;
  if (capS <= 640) return 64;
  if (capS <= 1600) return 128;
  return 256;
}

template <int T>
int launch_T(const mbm::KArgs &A, int member0, int count, bool use_global, size_t smem_bytes, cudaStream_t s) {
  auto kern = mbm::mbm_step_kernel<T>;
  if (T <= 64 && !use_global && A.inner_bc == mbm::BC_STEFAN && !getenv("MBM_NO_SPEC")) kern = mbm::mbm_step_kernel<T, (T <= 64) ? mbm::BC_STEFAN : 0>;
  if (!use_global) {
    if (int rc = launch_rc("general kernel: shared-memory opt-in", cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes))) return rc;
  }
  kern<<<count, T, use_global ? 0 : smem_bytes, s>>>(A, member0);
  return launch_rc("general kernel launch", cudaGetLastError());
}

#ifndef MBM_DEV_FIXED_T
#define MBM_DEV_FIXED_T 256
#define MBM_DEV_FIXED_N 2
#endif
#ifndef MBM_V2_T
#define MBM_V2_T 256
#endif
template <int T, int NDIM>
int launch_v2_TN(const mbm::KArgs &A, int member0, int count, size_t smem_bytes, cudaStream_t s) {
  auto kern = mbm::mbm_step_kernel_v2<T, NDIM>;
  if (int rc = launch_rc("tuned kernel: shared-memory opt-in", cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes))) return rc;
  kern<<<count, T, smem_bytes, s>>>(A, member0);
  return launch_rc("tuned kernel launch", cudaGetLastError());
}
template <int T>
int launch_v2_T(const mbm::KArgs &A, int member0, int count, size_t smem_bytes, cudaStream_t s) {
#ifdef MBM_DEV_MIN   // A/B builds: only the bench instantiation (T = 256 unless MBM_V2_T says otherwise, ndim = 2)
  if (T == 256 && A.ndim == 2) return launch_v2_TN<MBM_V2_T, 2>(A, member0, count, smem_bytes, s);
  return MBM_RC_UNSUPPORTED;
#else
  switch (A.ndim) {
    case 1: return launch_v2_TN<T, 1>(A, member0, count, smem_bytes, s);
    case 2: return launch_v2_TN<T, 2>(A, member0, count, smem_bytes, s);
    default: return launch_v2_TN<T, 3>(A, member0, count, smem_bytes, s);
  }
#endif
}

template <int T, int NDIM>
int launch_fixed_TN(const mbm::KArgs &A, int member0, int count, size_t smem_bytes, cudaStream_t s) {
  auto kern = mbm::mbm_fixed_kernel<T, NDIM>;
  if (int rc = launch_rc("fixed-interface kernel: shared-memory opt-in", cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes))) return rc;
  kern<<<count, T, smem_bytes, s>>>(A, member0);
  return launch_rc("fixed-interface kernel launch", cudaGetLastError());
}
template <int T>
int launch_fixed_T(const mbm::KArgs &A, int member0, int count, size_t smem_bytes, cudaStream_t s) {
#ifdef MBM_DEV_MIN   // A/B builds: one instantiation (MBM_DEV_FIXED_T threads, ndim MBM_DEV_FIXED_N; default 256, 2)
  if (T == MBM_DEV_FIXED_T && A.ndim == MBM_DEV_FIXED_N) return launch_fixed_TN<MBM_DEV_FIXED_T, MBM_DEV_FIXED_N>(A, member0, count, smem_bytes, s);
  return MBM_RC_UNSUPPORTED;
#else
  switch (A.ndim) {
    case 1: return launch_fixed_TN<T, 1>(A, member0, count, smem_bytes, s);
    case 2: return launch_fixed_TN<T, 2>(A, member0, count, smem_bytes, s);
    default: return launch_fixed_TN<T, 3>(A, member0, count, smem_bytes, s);
  }
#endif
}
template <int NDIM>
int launch_stream_N(const mbm::KArgs &A, int member0, int count, cudaStream_t s) {
  auto kern = mbm::mbm_stream_kernel<256, NDIM>;
  if (int rc = launch_rc("streaming kernel: shared-memory opt-in",
                         cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mbm::STG_BYTES))) return rc;
  kern<<<count, 256, mbm::STG_BYTES, s>>>(A, member0);
  return launch_rc("streaming kernel launch", cudaGetLastError());
}
template <int NDIM>
int launch_stream2_N(const mbm::KArgs &A, int member0, int count, cudaStream_t s) {
  auto kern = mbm::mbm_stream2_kernel<256, NDIM>;
  if (int rc = launch_rc("streaming kernel (closed-form meshes): shared-memory opt-in",
                         cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mbm::STG_BYTES))) return rc;
  kern<<<count, 256, mbm::STG_BYTES, s>>>(A, member0);
  return launch_rc("streaming kernel (closed-form meshes) launch", cudaGetLastError());
}
int launch_stream2(const mbm::KArgs &A, int member0, int count, cudaStream_t s) {
#ifdef MBM_DEV_MIN
  return launch_stream2_N<2>(A, member0, count, s);
#else
  switch (A.ndim) {
    case 1: return launch_stream2_N<1>(A, member0, count, s);
    case 2: return launch_stream2_N<2>(A, member0, count, s);
    default: return launch_stream2_N<3>(A, member0, count, s);
  }
#endif
}
int launch_stream(const mbm::KArgs &A, int member0, int count, cudaStream_t s) {
#ifdef MBM_DEV_MIN
  return launch_stream_N<2>(A, member0, count, s);
#else
  switch (A.ndim) {
    case 1: return launch_stream_N<1>(A, member0, count, s);
    case 2: return launch_stream_N<2>(A, member0, count, s);
    default: return launch_stream_N<3>(A, member0, count, s);
  }
#endif
}

// static shared memory of one instantiation, asked from the runtime (an estimate that is a few KB short makes the
// shared-memory opt-in fail at the largest capacities)
template <typename K>
int static_smem_of(K kern) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return 24 * 1024;
  return (int)fa.sharedSizeBytes;
}
template <int T>
int fixed_static_T(int ndim) {
#ifdef MBM_DEV_MIN
  return static_smem_of(mbm::mbm_fixed_kernel<MBM_DEV_FIXED_T, MBM_DEV_FIXED_N>);
#else
  switch (ndim) {
    case 1: return static_smem_of(mbm::mbm_fixed_kernel<T, 1>);
    case 2: return static_smem_of(mbm::mbm_fixed_kernel<T, 2>);
    default: return static_smem_of(mbm::mbm_fixed_kernel<T, 3>);
  }
#endif
}

// fixed-interface kernel: nine arrays of kF doubles per member (n1 + n2 <= kF)
void fixed_plan(int capL, int capR, int nmax_hint, int smem_optin, int ndim, int *Tf, int *kF) {
  int kf = capL + capR;
  if (nmax_hint > 0 && nmax_hint + 8 < kf) kf = (nmax_hint + 8) - (nmax_hint % 8);   // host buffers: size by the data
  if (const char *e = getenv("MBM_KCAP_FIXED")) { const int a = atoi(e); if (a >= 8 && a < kf) kf = a; }
  int T = 256;
  if (kf <= 320) T = 64;
  else if (kf <= 1100) T = 128;
  const int stat = (T == 64) ? fixed_static_T<64>(ndim) : (T == 128) ? fixed_static_T<128>(ndim) : fixed_static_T<256>(ndim);
  int kmax = (smem_optin - stat - 64) / (9 * 8);
  kmax -= kmax % 8;
  if (kf > kmax) kf = kmax;
  *Tf = T; *kF = kf;
}

// combined shared-memory capacity (left | free slot | right) of the tuned kernel: as large as the caps need, within
// two CTAs per SM (four arrays of kS doubles) and one scan tile per thread and phase (9*T nodes per phase)
template <int T>
int v2_static_T(int ndim) {
#ifdef MBM_DEV_MIN
  return static_smem_of(mbm::mbm_step_kernel_v2<MBM_V2_T, 2>);
#else
  switch (ndim) {
    case 1: return static_smem_of(mbm::mbm_step_kernel_v2<T, 1>);
    case 2: return static_smem_of(mbm::mbm_step_kernel_v2<T, 2>);
    default: return static_smem_of(mbm::mbm_step_kernel_v2<T, 3>);
  }
#endif
}
// largest combined capacity of the tuned kernel with two CTAs per SM / with one CTA per SM (opt-in limit per block)
void v2_capacities(int smem_per_sm, int smem_optin, int ndim, int *ks2, int *ks1) {
  const int stat = v2_static_T<256>(ndim) + 64;          // the widest instantiation bounds the others
  int a = ((smem_per_sm - 2 * 1024) / 2 - stat) / (4 * 8), b = (smem_optin - stat) / (4 * 8);
  a -= a % 8; b -= b % 8;
  *ks2 = a; *ks1 = b;
}
// members the host knows to be larger than two-per-SM allows (host buffers: largest n1 + n2 of the shard) but that fit one
// CTA per SM with 9 * 256 nodes per phase still take the tuned kernel, at half the occupancy, instead of the general
// kernel on a global workspace (an order of magnitude between the two)
bool v2_one_cta_plan(int nmax_hint, int ks2, int ks1) {
  if (nmax_hint <= 0 || getenv("MBM_V2_NO_ONE_CTA")) return false;
  // (the two-per-SM plan is only given up when the members as they ARE do not fit it: config 4's resorption members reach
  //  2 887 of the 2 944 nodes two CTAs per SM hold and must stay two per SM)
  const int need = nmax_hint + nmax_hint / 128 + 2;
  return need > ks2 && need <= ks1;
}
void v2_plan(int capL, int capR, int nmax_hint, int smem_per_sm, int smem_optin, int ndim, int *T2, int *kS) {
  int ks = capL + capR + 1;
  if (const char *e = getenv("MBM_KCAP")) { const int a = atoi(e); if (a >= 8 && a < ks) ks = a; }
  int ks_max = 0, ks_one = 0;
  v2_capacities(smem_per_sm, smem_optin, ndim, &ks_max, &ks_one);
  if (v2_one_cta_plan(nmax_hint, ks_max, ks_one)) ks_max = ks_one;
  if (ks > ks_max) ks = ks_max;
  // threads per member from the node counts when the host knows them (host buffers: largest n1 + n2 of the shard, with
  // room to grow by a quarter), else from the capacities; a member that outgrows 9*T nodes per phase is handed over
  const int nsz = (nmax_hint > 0 && nmax_hint + nmax_hint / 4 + 2 < ks) ? (nmax_hint + nmax_hint / 4 + 2) : ks;
  int T = 256;
  if (nsz <= 9 * 32) T = 32;               // one warp per member: every block barrier is a warp barrier
  else if (nsz <= 9 * 64) T = 64;
  else if (nsz <= 9 * 128) T = 128;
  if (const char *e = getenv("MBM_V2_THREADS")) {
    const int t = atoi(e);
    if (t == 32 || t == 64 || t == 128 || t == 256) T = t;
  }
#if defined(MBM_DEV_MIN) && MBM_V2_T != 256
  if (T == 256) { *T2 = MBM_V2_T; *kS = ks; return; }      // the launch switch maps "256" to the MBM_V2_T instantiation
#endif
  *T2 = T; *kS = ks;
}

size_t static_smem_for(int T) { return sizeof(mbm::Scal) + 8 * (size_t)T * sizeof(double) + 256; }

// run one shard [m0, m1) of an already device-resident KArgs
int launch_shard(mbm::KArgs A, int m0, int m1, int device, unsigned flags, cudaStream_t s, Stage *st,
                 double *kernel_seconds) {
  const int capS = A.capS;
  const size_t per_cta = 5 * (size_t)capS + (size_t)(capS / 256 + 8);
  const size_t dyn = per_cta * sizeof(double);
  // threads per member from the shard's largest node count when the host knows it (with room to grow), else from the
  // capacities: the capacities of small members are several times their node counts
  const int nsz = (A.pad3_ > 0 && 2 * A.pad3_ + 8 < capS) ? (2 * A.pad3_ + 8) : capS;
  int T = threads_for(nsz);
  int maxOptin = 0;
  CK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  bool use_global = (dyn + static_smem_for(T) > (size_t)maxOptin);
  if (flags & MBM_FLAG_FORCE_STREAMING) use_global = true;
  if ((flags & MBM_FLAG_FORCE_RESIDENT) && use_global && !(flags & MBM_FLAG_FORCE_STREAMING)) return MBM_RC_UNSUPPORTED;
  const int count = m1 - m0;
  if (count <= 0) return 0;
  int chunk = count;
  if (use_global) {
    if (!getenv("MBM_THREADS")) T = 512;
    // bound the workspace: at most 8 GiB, members processed in waves
    const size_t maxws = (size_t)8 << 30;
    size_t nmax = maxws / (per_cta * sizeof(double));
    if (nmax < 1) nmax = 1;
    if ((size_t)chunk > nmax) chunk = (int)nmax;
    A.workspace = (double *)st->scratch((size_t)chunk * per_cta * sizeof(double));
    if (!A.workspace) return st->rc ? st->rc : MBM_RC_OOM;
  }
  // streaming kernel (long profiles on uniform meshes): chunk-interleaved workspace of 6 arrays x MC rows x 256 columns
  // (the streaming kernels keep no per-step history: MBM_FLAG_RECORD calls stay with the general kernel, which does)
  const bool wants_history = (flags & MBM_FLAG_RECORD) && A.rec != nullptr && A.rec_cap > 0;
  const bool use_stream = use_global && (A.inner_bc == MBM_BC_FLUX || A.inner_bc == MBM_BC_MB) && !getenv("MBM_NO_STREAM") &&
                          A.ndim >= 1 && A.ndim <= 3 && !wants_history;
  mbm::KArgs AS = A;
  if (use_stream) {
    const int MC = (A.capL + A.capR + 2 + 253) / 254 + 1;
    AS.kcapL = MC;
    AS.workspace = (double *)st->scratch((size_t)chunk * 6 * (size_t)MC * 256 * sizeof(double));
    if (!AS.workspace) return st->rc ? st->rc : MBM_RC_OOM;
  }
  A.use_global = use_global ? 1 : 0;
  const bool couple_pv = (A.inner_bc == MBM_BC_FLUX || A.inner_bc == MBM_BC_MB);
  // MBM_FLAG_NEWTON_WARM: the run-ahead chain (tuned kernel, and the general kernel's batched Newton) always solves the
  // cold-start problems of the reference (functions.jl:734), so warm-start calls take the general kernel's serial path,
  // the one that starts newton_solver from the member's previous root like the oracle does (mbm_oracle.c, regrid)
  const bool newton_warm = (flags & MBM_FLAG_NEWTON_WARM) != 0;
  const bool no_chain = newton_warm || getenv("MBM_NO_CHAIN");
  bool use_v2 = couple_pv && !no_chain && !getenv("MBM_NO_V2") && A.ndim >= 1 && A.ndim <= 3;
  int T2 = 256;
  if (use_v2) {
    int smem_sm = 0;
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    if (use_global) {
      // capacities beyond the general kernel's shared-memory layout: the tuned kernel still takes the shard if the host
      // knows that its members fit one CTA per SM
      int k2 = 0, k1 = 0;
      v2_capacities(smem_sm, maxOptin, A.ndim, &k2, &k1);
      const int need = A.pad3_ + A.pad3_ / 128 + 2;
      use_v2 = A.pad3_ > 0 && need <= k1 && !getenv("MBM_V2_NO_ONE_CTA") && !(flags & MBM_FLAG_FORCE_STREAMING);
    }
  }
  if (use_v2) {
    int smem_sm = 0;
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    v2_plan(A.capL, A.capR, A.pad3_, smem_sm, maxOptin, A.ndim, &T2, &A.kcapL);
    if (A.kcapL < 8) use_v2 = false;
  }
  g_last_tuned_capacity.store(use_v2 ? A.kcapL : 0);
  if (couple_pv && !no_chain) {
    const int Tr = (use_v2 && T2 > T) ? T2 : T;
    A.recs = (double *)st->scratch((size_t)chunk * (size_t)Tr * mbm::REC_N * sizeof(double));
    if (!A.recs) return st->rc ? st->rc : MBM_RC_OOM;
  }
  const bool use_fixed = (couple_pv || A.inner_bc == MBM_BC_DIRICHLET_HOLD) && !use_global && !getenv("MBM_NO_FIXED") &&
                         A.ndim >= 1 && A.ndim <= 3;
  int Tf = 128;
  if (use_fixed) fixed_plan(A.capL, A.capR, A.pad3_, maxOptin, A.ndim, &Tf, &A.kcapR);
  if ((use_fixed || use_stream) && !use_v2) {
    A.steps_done = (long long *)st->scratch((size_t)count * sizeof(long long));
    if (!A.steps_done) return st->rc ? st->rc : MBM_RC_OOM;
    CK(cudaMemsetAsync(A.steps_done, 0, (size_t)count * sizeof(long long), s));
  }
  if (use_v2) {
    A.csbuf = (double *)st->scratch((size_t)chunk * (size_t)(9 * T2 + 8) * sizeof(double));
    if (!A.csbuf) return st->rc ? st->rc : MBM_RC_OOM;
    A.steps_done = (long long *)st->scratch((size_t)count * sizeof(long long));
    if (!A.steps_done) return st->rc ? st->rc : MBM_RC_OOM;
    CK(cudaMemsetAsync(A.steps_done, 0, (size_t)count * sizeof(long long), s));
  }
  const bool async = g_have_user_stream && st->device_ptrs;
  struct EventPair {                             // destroyed on every exit path of this function
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } ev;
  cudaEvent_t &e0 = ev.e0, &e1 = ev.e1;
  if (!async) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, s));
  }
  long long launches = 0;
  for (int off = 0; off < count; off += chunk) {
    const int n = (count - off < chunk) ? (count - off) : chunk;
    int rc;
    if (use_v2) {
      // tuned kernel first; the general kernel then only continues members handed over with status -1
      const size_t dyn2 = 4 * (size_t)A.kcapL * sizeof(double);
      mbm::KArgs A2 = A;
      A2.only_deferred = 0;
      switch (T2) {
        case 32: rc = launch_v2_T<32>(A2, m0 + off, n, dyn2, s); break;
        case 64: rc = launch_v2_T<64>(A2, m0 + off, n, dyn2, s); break;
        case 128: rc = launch_v2_T<128>(A2, m0 + off, n, dyn2, s); break;
        default: rc = launch_v2_T<256>(A2, m0 + off, n, dyn2, s); break;
      }
      if (rc) return rc;
      ++launches;
      A.only_deferred = 1;
    }
    if (use_fixed) {
      // fixed-interface members (V_ip == 0): resident kernel that keeps the factorisation
      const size_t dynf = 9 * (size_t)A.kcapR * sizeof(double);
      switch (Tf) {
        case 64: rc = launch_fixed_T<64>(A, m0 + off, n, dynf, s); break;
        case 128: rc = launch_fixed_T<128>(A, m0 + off, n, dynf, s); break;
        default: rc = launch_fixed_T<256>(A, m0 + off, n, dynf, s); break;
      }
      if (rc) return rc;
      ++launches;
      A.only_deferred = 1;
    }
    if (use_stream) {
      AS.steps_done = A.steps_done; AS.only_deferred = A.only_deferred; AS.recs = A.recs;
      // closed-form meshes first (only the concentrations cross HBM); members whose x arrays are not LinRanges are
      // handed on (status -1) to the kernel that keeps x in memory
      if (!getenv("MBM_NO_STREAM2")) {
        if (getenv("MBM_S2_NO_CUTOFF")) AS.flags |= mbm::F_DEV_S2_NO_CUTOFF;      // (tests: every row keeps its fill-in)
        rc = launch_stream2(AS, m0 + off, n, s);
        if (rc) return rc;
        ++launches;
        AS.only_deferred = 1;
      }
      rc = launch_stream(AS, m0 + off, n, s);
      if (rc) return rc;
      ++launches;
      A.only_deferred = 1;
    }
#ifdef MBM_DEV_MIN
    rc = launch_T<256>(A, m0 + off, n, use_global, dyn, s);
    if (rc) return rc;
    ++launches;
    continue;
#else
    switch (T) {
      case 32: rc = launch_T<32>(A, m0 + off, n, use_global, dyn, s); break;
      case 64: rc = launch_T<64>(A, m0 + off, n, use_global, dyn, s); break;
      case 128: rc = launch_T<128>(A, m0 + off, n, use_global, dyn, s); break;
      case 512: rc = launch_T<512>(A, m0 + off, n, use_global, dyn, s); break;
      default: rc = launch_T<256>(A, m0 + off, n, use_global, dyn, s); break;
    }
#endif
    if (rc) return rc;
    ++launches;
  }
  g_last_kernel_launches = launches;
  if (async) return 0;
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (kernel_seconds) *kernel_seconds = ms * 1e-3;
  return 0;
}

// keep the stream-ordered allocator's memory cached between calls (default: released at the next synchronisation,
// which re-maps hundreds of MB of scratch on every call)
void keep_pool_cached(int device) {
  static std::once_flag once[64];                // shard threads of concurrent calls race here: once per device
  if (device < 0 || device >= 64) return;
  std::call_once(once[device], [device] {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  });
}

// argument checks first (no CUDA call), the device check last: a bad argument is reported as such on any machine, and
// a well-formed call on a machine without a GPU fails loudly with MBM_RC_NO_DEVICE (there is no CPU path)
int check_common(const mbm_common *c) {
  if (!c || c->abi_version != MBM_ABI_VERSION) return MBM_RC_BAD_ARG;
  if (c->n_members < 0 || c->cap_left < 3 || c->hist_len < 2) return MBM_RC_BAD_ARG;
  if (c->n_devices < 1) return MBM_RC_BAD_ARG;
  if ((c->flags & MBM_FLAG_DEVICE_POINTERS) && c->n_devices != 1) return MBM_RC_BAD_ARG;
  return 0;
}
int check_devices(const mbm_common *c) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return MBM_RC_NO_DEVICE;
  for (int i = 0; i < c->n_devices; ++i) {
    const int d = c->device_ids ? c->device_ids[i] : i;
    if (d < 0 || d >= ndev) return MBM_RC_NO_DEVICE;
  }
  return 0;
}

// ---- couple shard ------------------------------------------------------------------------------------
int couple_shard(const mbm_common *c, const mbm_couple_params *p, mbm_couple_io *io, int m0, int m1, int device,
                 double *ksec) {
  CK(cudaSetDevice(device));
  keep_pool_cached(device);
  Stage st;
  st.device_ptrs = (c->flags & MBM_FLAG_DEVICE_POINTERS) != 0;
  const bool own_stream = !g_have_user_stream;
  if (own_stream) CK(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
  else st.stream = g_user_stream;
  const size_t n = (size_t)(m1 - m0), capL = c->cap_left, capR = c->cap_right, H = c->hist_len;
  PieceTurn turn(st.stream);
  mbm::KArgs A;
  memset(&A, 0, sizeof(A));
  A.M = (int)n; A.ndim = c->ndim; A.hist_len = c->hist_len; A.capL = c->cap_left; A.capR = c->cap_right;
  A.capS = c->cap_left + c->cap_right + 1; A.inner_bc = c->inner_bc; A.dt_mode = c->dt_mode; A.flags = c->flags;
  A.pstride = 2; A.max_steps = c->max_steps; A.rec_cap = c->rec_cap; A.R = p->R;
  // all arrays are re-based to the shard: member index inside the kernel is 0..n-1
  A.Di = st.in(p->Di + 2 * (size_t)m0, 2 * n);
  A.D0 = st.in(p->D0 ? p->D0 + 2 * (size_t)m0 : nullptr, 2 * n);
  A.Ea = st.in(p->Ea ? p->Ea + 2 * (size_t)m0 : nullptr, 2 * n);
  A.rho = st.in(p->rho + 2 * (size_t)m0, 2 * n);
  A.V_ip = st.in(p->V_ip ? p->V_ip + m0 : nullptr, n);
  A.t_tot = st.in(p->t_tot + m0, n);
  A.CFL = st.in(p->CFL + m0, n);
  A.MRefin = st.in(p->MRefin + m0, n);
  A.BCout = st.in(p->BCout + 2 * (size_t)m0, 2 * n);
  A.resmin = (const long long *)st.in(p->resmin + 2 * (size_t)m0, 2 * n);
  A.t_stride = p->t_stride; A.T_stride = p->T_stride; A.KD_stride = p->KD_stride;
  A.t_ar = st.in(p->t_ar + (size_t)m0 * p->t_stride, p->t_stride ? n * H : H);
  A.T_ar = st.in(p->T_ar + (size_t)m0 * p->T_stride, p->T_stride ? n * H : H);
  A.KD_ar = p->KD_ar ? st.in(p->KD_ar + (size_t)m0 * p->KD_stride, p->KD_stride ? n * H : H) : nullptr;
  if (c->inner_bc == MBM_BC_STEFAN) {
    A.coeff_stride = p->coeff_stride;
    A.coeff_up = st.in(p->coeff_up + (size_t)m0 * p->coeff_stride, p->coeff_stride ? 3 * n : 3);
    A.coeff_do = st.in(p->coeff_do + (size_t)m0 * p->coeff_stride, p->coeff_stride ? 3 * n : 3);
    A.C_leftB = st.in(p->C_leftB + m0, n);
    A.C_rightB = st.in(p->C_rightB + m0, n);
    A.nd = p->nd;
    A.law = p->law_par ? p->law : MBM_LAW_NONE;
    if (A.law) A.law_par = st.in(p->law_par + 8 * (size_t)m0, 8 * n);
    if (p->nd > 0) {
      A.dens_X = st.in(p->dens_X, (size_t)p->nd);
      A.dens_T = st.in(p->dens_T, (size_t)p->nd);
      A.dens_left = st.in(p->dens_left, (size_t)p->nd * p->nd);
      A.dens_right = st.in(p->dens_right, (size_t)p->nd * p->nd);
    }
  }
  if (!st.device_ptrs) {                       // largest node count of the shard (sizes the fixed-interface kernel)
    long long nmax = 0;
    for (size_t k = 0; k < n; ++k) { const long long v = io->res[2 * (m0 + k)] + io->res[2 * (m0 + k) + 1]; if (v > nmax) nmax = v; }
    A.pad3_ = (int)nmax;
  }
  A.x_left = st.inout(io->x_left + (size_t)m0 * capL, n * capL);
  A.C_left = st.inout(io->C_left + (size_t)m0 * capL, n * capL);
  A.x_right = st.inout(io->x_right + (size_t)m0 * capR, n * capR);
  A.C_right = st.inout(io->C_right + (size_t)m0 * capR, n * capR);
  A.res = (long long *)st.inout(io->res + 2 * (size_t)m0, 2 * n);
  A.Ri = st.inout(io->Ri + 2 * (size_t)m0, 2 * n);
  A.t = st.inout(io->t + m0, n);
  A.dt = st.inout(io->dt + m0, n);
  A.it = (long long *)st.inout(io->it + m0, n);
  A.dx1 = io->dx1 ? st.outonly(io->dx1 + m0, n) : nullptr;
  A.dx2 = io->dx2 ? st.outonly(io->dx2 + m0, n) : nullptr;
  A.Mass0 = io->Mass0 ? st.in(io->Mass0 + m0, n) : nullptr;
  A.Mass = io->Mass ? st.inout(io->Mass + m0, n) : nullptr;
  A.newton_R = io->newton_R ? st.inout(io->newton_R + m0, n) : nullptr;
  A.last = io->last ? st.inout(io->last + 8 * (size_t)m0, 8 * n) : nullptr;
  A.status = st.inout(io->status + m0, n);
  A.warn = io->warn ? st.inout(io->warn + m0, n) : nullptr;
  A.idx_hash = io->idx_hash ? (unsigned long long *)st.inout(io->idx_hash + m0, n) : nullptr;
  A.node_steps = io->node_steps ? (long long *)st.inout(io->node_steps + m0, n) : nullptr;
  if ((c->flags & MBM_FLAG_RECORD) && io->rec && c->rec_cap > 0)
    A.rec = st.inout(io->rec + 4 * (size_t)c->rec_cap * (size_t)m0, 4 * (size_t)c->rec_cap * n);
  turn.done();
  int rc = st.rc;
  if (!rc) rc = launch_shard(A, 0, (int)n, device, c->flags, st.stream, &st, ksec);
  const bool async = g_have_user_stream && st.device_ptrs;
  if (!rc) rc = st.copy_back();
  if (!async && !rc) rc = launch_rc("stream synchronisation after the kernels", cudaStreamSynchronize(st.stream));
  st.release();
  if (!async) cudaStreamSynchronize(st.stream);
  if (own_stream) cudaStreamDestroy(st.stream);
  return rc;
}

int simple_shard(const mbm_common *c, const mbm_simple_params *p, mbm_simple_io *io, int m0, int m1, int device,
                 double *ksec) {
  CK(cudaSetDevice(device));
  keep_pool_cached(device);
  Stage st;
  st.device_ptrs = (c->flags & MBM_FLAG_DEVICE_POINTERS) != 0;
  const bool own_stream = !g_have_user_stream;
  if (own_stream) CK(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
  else st.stream = g_user_stream;
  const size_t n = (size_t)(m1 - m0), cap = c->cap_left, H = c->hist_len;
  mbm::KArgs A;
  memset(&A, 0, sizeof(A));
  A.M = (int)n; A.ndim = c->ndim; A.hist_len = c->hist_len; A.capL = c->cap_left; A.capR = 0;
  A.capS = c->cap_left + 1; A.inner_bc = MBM_BC_NONE; A.dt_mode = MBM_DT_SIMPLE; A.flags = c->flags;
  A.pstride = 1; A.max_steps = c->max_steps; A.R = p->R;
  A.Di = st.in(p->Di + m0, n);
  A.D0 = st.in(p->D0 ? p->D0 + m0 : nullptr, n);
  A.Ea = st.in(p->Ea1 ? p->Ea1 + m0 : nullptr, n);
  A.rho = st.in(p->rho + m0, n);
  A.s_dx = st.in(p->dx + m0, n);
  A.t_tot = st.in(p->t_tot + m0, n);
  A.CFL = st.in(p->CFL + m0, n);
  A.BCout = st.in(p->BCout + 2 * (size_t)m0, 2 * n);
  A.t_stride = p->t_stride; A.T_stride = p->T_stride;
  A.t_ar = st.in(p->t_ar + (size_t)m0 * p->t_stride, p->t_stride ? n * H : H);
  A.T_ar = st.in(p->T_ar + (size_t)m0 * p->T_stride, p->T_stride ? n * H : H);
  A.x_left = st.inout(io->x + (size_t)m0 * cap, n * cap);
  A.C_left = st.inout(io->C + (size_t)m0 * cap, n * cap);
  A.res = (long long *)st.inout(io->res + m0, n);
  A.t = st.inout(io->t + m0, n);
  A.dt = st.inout(io->dt + m0, n);
  A.it = (long long *)st.inout(io->it + m0, n);
  A.s_D = io->D ? st.outonly(io->D + m0, n) : nullptr;
  A.Mass = io->Mass ? st.inout(io->Mass + m0, n) : nullptr;
  A.status = st.inout(io->status + m0, n);
  A.node_steps = io->node_steps ? (long long *)st.inout(io->node_steps + m0, n) : nullptr;
  int rc = st.rc;
  if (!rc) rc = launch_shard(A, 0, (int)n, device, c->flags, st.stream, &st, ksec);
  const bool async = g_have_user_stream && st.device_ptrs;
  if (!rc) rc = st.copy_back();
  if (!async && !rc) rc = launch_rc("stream synchronisation after the kernels", cudaStreamSynchronize(st.stream));
  st.release();
  if (!async) cudaStreamSynchronize(st.stream);
  if (own_stream) cudaStreamDestroy(st.stream);
  return rc;
}

// contiguous blocks of ceil(M/G) members per device, one host thread per device
// contiguous blocks of ceil(M/G) members per device, one host thread (and stream) per block.  With host buffers a
// device's block is further cut into pieces that run on their own streams, so that the H2D copy of one piece, the
// kernel of another and the D2H copy of a third overlap (members never communicate: no ordering is needed).
template <typename F>
int run_sharded(const mbm_common *c, F shard_fn) {
  const int G = c->n_devices, M = c->n_members;
  if (M == 0) return 0;
  const int per = (M + G - 1) / G;
  int pieces = 1;
  const bool host_buffers = !(c->flags & MBM_FLAG_DEVICE_POINTERS) && !g_have_user_stream;
  if (host_buffers && per >= 2048) pieces = (per >= 4096) ? 8 : 4;
  if (const char *e = getenv("MBM_PIECES")) { const int p = atoi(e); if (p >= 1 && p <= 16) pieces = p; }
  if (!host_buffers) pieces = 1;
  struct Job { int m0, m1, dev, g; };
  std::vector<Job> jobs;
  for (int g = 0; g < G; ++g) {
    const int m0 = g * per, m1 = (m0 + per < M) ? (m0 + per) : M;
    if (m0 >= m1) continue;
    const int dev = c->device_ids ? c->device_ids[g] : g;
    const int n = m1 - m0, pp = (n + pieces - 1) / pieces;
    for (int q = 0; q < pieces; ++q) {
      const int a = m0 + q * pp, b = (a + pp < m1) ? (a + pp) : m1;
      if (a < b) jobs.push_back({a, b, dev, g});
    }
  }
  const size_t J = jobs.size();
  std::vector<int> rcs(J, 0);
  std::vector<double> ks(J, 0.0);
  std::vector<long long> nl(J, 0);
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  if (J == 1) {
    rcs[0] = shard_fn(jobs[0].m0, jobs[0].m1, jobs[0].dev, &ks[0]);
    nl[0] = g_last_kernel_launches;
  } else {
    PieceOrder order;
    const bool staggered = pieces > 1 && !getenv("MBM_NO_STAGGER");
    if (staggered) {
      order.turn.assign(G, 0); order.ev.assign(J, nullptr); order.group.resize(J); order.index.resize(J);
      for (size_t j = 0; j < J; ++j) {
        order.group[j] = jobs[j].g;
        order.index[j] = (j > 0 && jobs[j - 1].g == jobs[j].g) ? order.index[j - 1] + 1 : 0;
        if (cudaSetDevice(jobs[j].dev) == cudaSuccess) cudaEventCreateWithFlags(&order.ev[j], cudaEventDisableTiming);
      }
    }
    std::vector<std::thread> th;
    for (size_t j = 0; j < J; ++j)
      th.emplace_back([&, j]() {
        if (staggered) { tl_order = &order; tl_job = j; }
        rcs[j] = shard_fn(jobs[j].m0, jobs[j].m1, jobs[j].dev, &ks[j]);
        nl[j] = g_last_kernel_launches;
        tl_order = nullptr;
        if (staggered) order.finish(j);
      });
    for (auto &t : th) t.join();
    for (cudaEvent_t e : order.ev) if (e) cudaEventDestroy(e);
  }
  cudaSetDevice(prev_dev);
  double k0 = 0.0;
  long long l0 = 0;
  for (size_t j = 0; j < J; ++j)
    if (jobs[j].g == jobs[0].g) { k0 += ks[j]; l0 += nl[j]; }
  g_last_kernel_seconds = k0;
  g_last_kernel_seconds_any = k0;
  g_last_kernel_launches_any = l0 ? l0 : 1;
  for (size_t j = 0; j < J; ++j)
    if (rcs[j]) return rcs[j];
  return 0;
}

}  // namespace

extern "C" {

int mbm_version(void) { return MBM_ABI_VERSION; }

int mbm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char *mbm_backend(void) { return "cuda-sm_100a"; }

double mbm_last_kernel_seconds(void) { return g_last_kernel_seconds_any; }

int mbm_set_stream(void *cuda_stream) {
  g_user_stream = (cudaStream_t)cuda_stream;
  g_have_user_stream = (cuda_stream != nullptr);
  return 0;
}

__global__ void mbm_dfma_kernel(double *out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double b = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double mbm_measure_dfma_tflops(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return 0.0;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 20000;
  double *out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) { cudaSetDevice(prev); return 0.0; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    mbm_dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  cudaSetDevice(prev);
  return best;
}
int64_t mbm_last_kernel_launches(void) { return g_last_kernel_launches_any; }
int mbm_last_tuned_capacity(void) { return g_last_tuned_capacity.load(); }
int mbm_stream_pass_counts(int device, uint64_t *fused_steps, uint64_t *twopass_steps, uint64_t *rows_without_fill_in) {
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return MBM_RC_CUDA;
  unsigned long long f = 0, t = 0, r = 0;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(&f, mbm::g_s2_fused_steps, sizeof(f));
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(&t, mbm::g_s2_twopass_steps, sizeof(t));
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(&r, mbm::g_s2_cut_rows, sizeof(r));
  cudaSetDevice(prev);
  if (e != cudaSuccess) return MBM_RC_CUDA;
  if (fused_steps) *fused_steps = f;
  if (twopass_steps) *twopass_steps = t;
  if (rows_without_fill_in) *rows_without_fill_in = r;
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

int mbm_couple_batch(const mbm_common *c, const mbm_couple_params *p, mbm_couple_io *io) {
  int rc = check_common(c);
  if (rc) return rc;
  if (!p || !io || c->cap_right < 3) return MBM_RC_BAD_ARG;
  if (c->inner_bc < MBM_BC_FLUX || c->inner_bc > MBM_BC_LASAGA) return MBM_RC_BAD_ARG;
  if (!p->Di || !p->rho || !p->t_tot || !p->CFL || !p->MRefin || !p->BCout || !p->resmin || !p->t_ar || !p->T_ar)
    return MBM_RC_BAD_ARG;
  if (c->inner_bc != MBM_BC_STEFAN && !p->KD_ar) return MBM_RC_BAD_ARG;
  if (!io->x_left || !io->C_left || !io->x_right || !io->C_right || !io->res || !io->Ri || !io->t || !io->dt ||
      !io->it || !io->status)
    return MBM_RC_BAD_ARG;
  if (c->flags & (MBM_FLAG_SOLVER_FAITHFUL | MBM_FLAG_SUM_JULIA)) return MBM_RC_UNSUPPORTED;  // oracle-only modes
  // every pointer a kernel would dereference for this interface kind is checked here, before any launch: a NULL that
  // reaches a kernel is an illegal-address fault and a sticky context error for the whole process
  if (c->dt_mode < MBM_DT_FIND_DT || c->dt_mode > MBM_DT_SIMPLE) return MBM_RC_BAD_ARG;
  if (c->inner_bc == MBM_BC_STEFAN) {
    if (c->dt_mode != MBM_DT_STEFAN) return MBM_RC_BAD_ARG;
    if (!p->coeff_up || !p->coeff_do || !p->C_leftB || !p->C_rightB) return MBM_RC_BAD_ARG;
  } else {
    if (c->dt_mode == MBM_DT_STEFAN || c->dt_mode == MBM_DT_SIMPLE) return MBM_RC_BAD_ARG;
  }
  if (p->nd < 0) return MBM_RC_BAD_ARG;
  if (p->nd > 0 && (p->nd < 2 || !p->dens_X || !p->dens_T || !p->dens_left || !p->dens_right)) return MBM_RC_BAD_ARG;
  if (p->law != MBM_LAW_NONE && p->law != MBM_LAW_OLIVINE_MELT_D1) return MBM_RC_BAD_ARG;
  if (p->law != MBM_LAW_NONE && (!p->law_par || c->inner_bc != MBM_BC_STEFAN)) return MBM_RC_BAD_ARG;
  if ((rc = check_devices(c))) return rc;
  return run_sharded(c, [&](int m0, int m1, int dev, double *ks) { return couple_shard(c, p, io, m0, m1, dev, ks); });
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

int mbm_simple_diff_batch(const mbm_common *c, const mbm_simple_params *p, mbm_simple_io *io) {
  int rc = check_common(c);
  if (rc) return rc;
  if (!p || !io || !p->Di || !p->rho || !p->dx || !p->t_tot || !p->CFL || !p->BCout || !p->t_ar || !p->T_ar)
    return MBM_RC_BAD_ARG;
  if (!io->x || !io->C || !io->res || !io->t || !io->dt || !io->it || !io->status) return MBM_RC_BAD_ARG;
  if (c->flags & (MBM_FLAG_SOLVER_FAITHFUL | MBM_FLAG_SUM_JULIA)) return MBM_RC_UNSUPPORTED;
  if ((rc = check_devices(c))) return rc;
  return run_sharded(c, [&](int m0, int m1, int dev, double *ks) { return simple_shard(c, p, io, m0, m1, dev, ks); });
}

int mbm_create_grid(int32_t M, int32_t capL, int32_t capR, const double *Ri, const int64_t *res, const double *MRefin,
                    double *x_left, double *x_right, double *dx1, double *dx2, double *newton_R, int32_t *status) {
  if (M < 0 || capL < 3 || capR < 3) return MBM_RC_BAD_ARG;
  if (M > 0 && (!Ri || !res || !MRefin || !x_left || !x_right || !status)) return MBM_RC_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return MBM_RC_NO_DEVICE;
  if (M == 0) return 0;
  // the pre-loop helpers run on device 0 whatever device the calling thread had current, and restore it
  DeviceScope dev_scope(0);
  if (dev_scope.rc) return dev_scope.rc;
  Stage st;
  CK(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
  const size_t n = (size_t)M;
  const double *dRi = st.in(Ri, 2 * n);
  const long long *dres = (const long long *)st.in(res, 2 * n);
  const double *dMR = st.in(MRefin, n);
  double *dxl = st.outonly(x_left, n * capL), *dxr = st.outonly(x_right, n * capR);
  double *d1 = dx1 ? st.outonly(dx1, n) : nullptr, *d2 = dx2 ? st.outonly(dx2, n) : nullptr;
  double *dnR = newton_R ? st.outonly(newton_R, n) : nullptr;
  int *dst = st.outonly(status, n);
  double *work = (double *)st.scratch(n * (size_t)capR * sizeof(double));
  int rc = st.rc;
  if (!rc) {
    mbm::mbm_create_grid_kernel<256><<<M, 256, 0, st.stream>>>(M, capL, capR, dRi, dres, dMR, dxl, dxr, d1, d2, dnR, dst, work);
    if (cudaGetLastError() != cudaSuccess) rc = MBM_RC_CUDA;
  }
  if (!rc) rc = st.copy_back();
  if (!rc && cudaStreamSynchronize(st.stream) != cudaSuccess) rc = MBM_RC_CUDA;
  st.release();
  cudaStreamSynchronize(st.stream);
  cudaStreamDestroy(st.stream);
  return rc;
}

int mbm_calc_mass_vol(int32_t M, int32_t ndim, int32_t capL, int32_t capR, const double *x_left, const double *x_right,
                      const double *C_left, const double *C_right, const int64_t *res, const double *rho, double *Mass) {
  if (M < 0 || capL < 1 || capR < 1 || ndim < 1 || ndim > 3) return MBM_RC_BAD_ARG;
  if (M > 0 && (!x_left || !x_right || !C_left || !C_right || !res || !rho || !Mass)) return MBM_RC_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return MBM_RC_NO_DEVICE;
  if (M == 0) return 0;
  DeviceScope dev_scope(0);
  if (dev_scope.rc) return dev_scope.rc;
  Stage st;
  CK(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
  const size_t n = (size_t)M;
  const double *dxl = st.in(x_left, n * capL), *dxr = st.in(x_right, n * capR);
  const double *dcl = st.in(C_left, n * capL), *dcr = st.in(C_right, n * capR);
  const long long *dres = (const long long *)st.in(res, 2 * n);
  const double *drho = st.in(rho, 2 * n);
  double *dM = st.outonly(Mass, n);
  int rc = st.rc;
  if (!rc) {
    mbm::mbm_calc_mass_kernel<<<(M + 63) / 64, 64, 0, st.stream>>>(M, ndim, capL, capR, dxl, dxr, dcl, dcr, dres, drho, dM);
    if (cudaGetLastError() != cudaSuccess) rc = MBM_RC_CUDA;
  }
  if (!rc) rc = st.copy_back();
  if (!rc && cudaStreamSynchronize(st.stream) != cudaSuccess) rc = MBM_RC_CUDA;
  st.release();
  cudaStreamSynchronize(st.stream);
  cudaStreamDestroy(st.stream);
  return rc;
}

// validation layer: host or device pointers, one thread per member (not a hot path)
int mbm_validate_batch(const mbm_validate_args *a) {
  if (!a || a->abi_version != MBM_ABI_VERSION) return MBM_RC_BAD_ARG;
  if (a->n_members < 0 || a->cap_left < 3 || a->cap_right < 3 || a->ndim < 1 || a->ndim > 3) return MBM_RC_BAD_ARG;
  if (a->kind < MBM_VAL_MASS_ERR || a->kind > MBM_VAL_SPHERE) return MBM_RC_BAD_ARG;
  if (!a->x_left || !a->C_left || !a->x_right || !a->C_right || !a->res || !a->Ri || !a->out || !a->relerr) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_MASS_ERR && (!a->rho || !a->Mass0)) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_RAYLEIGH && (!a->Ri0 || !a->C0_r || !a->KD0)) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_SMITH && (!a->KD0 || !a->Di_r || !a->V_ip || !a->t_tot || a->ndim != 1)) return MBM_RC_BAD_ARG;
  if (a->kind == MBM_VAL_SPHERE && (!a->zeta || !a->C0_left || !a->C0_right)) return MBM_RC_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return MBM_RC_NO_DEVICE;
  if (a->device < 0 || a->device >= ndev) return MBM_RC_NO_DEVICE;
  if (a->n_members == 0) return 0;
  DeviceScope dev_scope(a->device);
  if (dev_scope.rc) return dev_scope.rc;
  Stage st;
  st.device_ptrs = (a->flags & MBM_FLAG_DEVICE_POINTERS) != 0;
  const bool own_stream = !g_have_user_stream;
  if (own_stream) CK(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
  else st.stream = g_user_stream;
  const size_t n = (size_t)a->n_members, capL = a->cap_left, capR = a->cap_right;
  mbm::VArgs V;
  memset(&V, 0, sizeof(V));
  V.M = a->n_members; V.ndim = a->ndim; V.capL = a->cap_left; V.capR = a->cap_right; V.kind = a->kind;
  V.x_left = st.in(a->x_left, n * capL); V.C_left = st.in(a->C_left, n * capL);
  V.x_right = st.in(a->x_right, n * capR); V.C_right = st.in(a->C_right, n * capR);
  V.res = (const long long *)st.in(a->res, 2 * n);
  V.Ri = st.in(a->Ri, 2 * n);
  V.rho = st.in(a->rho, 2 * n); V.Mass0 = st.in(a->Mass0, n);
  V.Ri0 = st.in(a->Ri0, 2 * n); V.C0_r = st.in(a->C0_r, n); V.KD0 = st.in(a->KD0, n);
  V.Di_r = st.in(a->Di_r, n); V.V_ip = st.in(a->V_ip, n); V.t_tot = st.in(a->t_tot, n);
  if (a->kind == MBM_VAL_SPHERE) {
    V.zeta = st.in(a->zeta, n); V.C0_left = st.in(a->C0_left, n * capL); V.C0_right = st.in(a->C0_right, n * capR);
  }
  V.out = st.outonly(a->out, 2 * n); V.relerr = st.outonly(a->relerr, n);
  int rc = st.rc;
  if (!rc) {
    mbm::mbm_validate_kernel<<<(a->n_members + 63) / 64, 64, 0, st.stream>>>(V);
    rc = launch_rc("validation kernel", cudaGetLastError());
  }
  if (!rc) rc = st.copy_back();
  const bool async = g_have_user_stream && st.device_ptrs;
  if (!rc && !async && cudaStreamSynchronize(st.stream) != cudaSuccess) rc = MBM_RC_CUDA;
  st.release();
  if (own_stream) { cudaStreamSynchronize(st.stream); cudaStreamDestroy(st.stream); }
  return rc;
}

}  // extern "C"
