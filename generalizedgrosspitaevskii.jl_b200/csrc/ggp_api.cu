// C ABI of libggp.so (include/ggp.h): plan management, table upload, step scheduling.
// Replaces init / step! / the inner loop of solve! of the reference (src/strang_splitting.jl:32-90,
// src/fixed_time_stepping.jl:38-50).  No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <complex>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ggp.h"
#include <type_traits>
#include "kernels.cuh"
#ifdef GGP_TMA
#include "str_tma.cuh"
#endif
#ifdef GGP_PACKED
#include "packed.cuh"
#endif
#include "sizes_gen.h"

#ifdef GGP_WITH_NCCL
#include <nccl.h>
#endif

namespace ggp {

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define GGP_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(GGP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));            \
  } while (0)

#define GGP_LAUNCH(expr, what)                                                                  \
  do {                                                                                          \
    int e__ = (expr);                                                                           \
    if (e__ != 0)                                                                               \
      return fail(GGP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString((cudaError_t)e__)); \
  } while (0)

// ---- size dispatch ---------------------------------------------------------------------------
template <typename T>
static int dispatch_row(int N, int M, int pwv, const RowParams<T>& p, cudaStream_t st) {
  switch (N) {
#define X(n) \
  case n:    \
    return launch_row<T, n>(M, pwv, p, st);
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}
template <typename T>
static int dispatch_str(int N, int M, const StrParams<T>& p, long long nfast, long long nother, cudaStream_t st) {
  switch (N) {
#define X(n) \
  case n:    \
    return launch_str<T, n>(M, p, nfast, nother, st);
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}
template <typename T>
static int dispatch_oned(int N, int M, int pwv, const OneDParams<T>& p, cudaStream_t st) {
  switch (N) {
#define X(n) \
  case n:    \
    return launch_oned<T, n>(M, pwv, p, st);
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}
#ifdef GGP_TMA
template <typename T>
static int dispatch_str_tma(int N, int M, const StrTmaParams<T>& p, long long nfast, long long nother, int sms,
                            cudaStream_t st) {
  switch (N) {
#define X(n) \
  case n:    \
    return launch_str_tma<T, n>(M, p, nfast, nother, sms, st);
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}
#endif
template <typename T>
static void dispatch_str_query(int N, int M, int ax, long long nfast, int* W, int* LS, int* threads, int* us) {
  *W = 0;
  switch (N) {
#define X(n)                                   \
  case n:                                      \
    str_query<T, n>(M, ax, 0, nfast, W, LS, threads, us); \
    return;
    GGP_SIZES(X)
#undef X
  }
}
static bool size_supported(long long n) {
  switch (n) {
#define X(n_) \
  case n_:    \
    return true;
    GGP_SIZES(X)
#undef X
  }
  return false;
}

#ifdef GGP_PACKED
// packed two-line fp32 kernels (packed.cuh): line lengths >= GGP_PACKED_MIN_N of the built sizes
static bool packed_size(long long n) { return n >= GGP_PACKED_MIN_N && size_supported(n); }
static int dispatch_row2(int N, const RowParams<float>& p, cudaStream_t st) {
  switch (N) {
#define X(n)                                                         \
  case n:                                                            \
    if constexpr (n >= GGP_PACKED_MIN_N) return launch_row2<n>(p, st); \
    break;
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}
static int dispatch_str2(int N, const StrParams<float>& p, long long nfast, long long nother, cudaStream_t st) {
  switch (N) {
#define X(n)                                                                        \
  case n:                                                                           \
    if constexpr (n >= GGP_PACKED_MIN_N) return launch_str2<n>(p, nfast, nother, st); \
    break;
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}
#endif

// ---- observables -------------------------------------------------------------------------------
template <typename T>
__global__ void density_kernel(const cpx<T>* __restrict__ u, double* __restrict__ out, long long nspatial,
                               long long nbatch, double scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nspatial) return;
  double acc = 0;
  for (long long b = 0; b < nbatch; ++b) {
    const cpx<T> z = u[b * nspatial + i];
    acc += (double)z.x * (double)z.x + (double)z.y * (double)z.y;
  }
  out[i] = acc * scale;
}

__global__ void sum_kernel(const double* __restrict__ in, double* __restrict__ out, long long n) {
  __shared__ double sh[256];
  double acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += in[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, sh[0]);
}

// Second moment of the momentum occupation over the trajectories of a 1-D ensemble (the double loop `G2` of
// examples/truncated_wigner.jl:143-154 without its O(N^2 B) host pass):
//   out[m*N + n] += scale * sum_{t in this z-slice} |u~_t[m]|^2 |u~_t[n]|^2
// u~ is the transformed state (natural order).  16x16 output tile per CTA, trajectories in chunks of 16 through
// shared memory, blockIdx.z splits the trajectories; fp64 accumulation, one atomicAdd per output and z-slice.
template <typename T>
__global__ void g2_kernel(const cpx<T>* __restrict__ u, double* __restrict__ out, int N, long long nbatch, double scale) {
  __shared__ double Im[16][17], In[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m0 = blockIdx.y * 16, n0 = blockIdx.x * 16;
  const long long per = (nbatch + gridDim.z - 1) / gridDim.z;
  const long long t0 = (long long)blockIdx.z * per, t1 = t0 + per < nbatch ? t0 + per : nbatch;
  double acc = 0;
  for (long long tb = t0; tb < t1; tb += 16) {
    const long long t = tb + ty;
    double im = 0, in = 0;
    if (t < t1) {
      if (m0 + tx < N) {
        const cpx<T> z = u[t * N + m0 + tx];
        im = (double)z.x * (double)z.x + (double)z.y * (double)z.y;
      }
      if (n0 + tx < N) {
        const cpx<T> z = u[t * N + n0 + tx];
        in = (double)z.x * (double)z.x + (double)z.y * (double)z.y;
      }
    }
    Im[ty][tx] = im;
    In[ty][tx] = in;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += Im[k][ty] * In[k][tx];
    __syncthreads();
  }
  if (m0 + ty < N && n0 + tx < N) atomicAdd(out + (size_t)(m0 + ty) * N + n0 + tx, acc * scale);
}

// u[t][p] <- src[t][p] * w[p]   (window of the windowed Fourier transform, test/windowed_ft.jl:32-33)
template <typename T>
__global__ void window_kernel(cpx<T>* __restrict__ u, const cpx<T>* __restrict__ src, const cpx<T>* __restrict__ w, int N,
                              long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  u[i] = cmul(src[i], w[i % N]);
}

// First-order coherence between two transformed copies of a 1-D ensemble (test/windowed_ft.jl:38-46):
//   gram[k][l] = sum_t X1_t[k] conj(X2_t[l]),   written to out[i][j] (re, im), i = (k + N/2) % N, j = (l + N/2) % N,
//   with the sign (-1)^(k+l): fft(fftshift(x))[k] = (-1)^k fft(x)[k] and ifftshift is that index shift (N even).
template <typename T>
__global__ void g1_kernel(const cpx<T>* __restrict__ x1, const cpx<T>* __restrict__ x2, double* __restrict__ out, int N,
                          long long nbatch) {
  __shared__ double Ar[16][17], Ai[16][17], Br[16][17], Bi[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = blockIdx.y * 16, l0 = blockIdx.x * 16;
  const long long per = (nbatch + gridDim.z - 1) / gridDim.z;
  const long long t0 = (long long)blockIdx.z * per, t1 = t0 + per < nbatch ? t0 + per : nbatch;
  double accr = 0, acci = 0;
  for (long long tb = t0; tb < t1; tb += 16) {
    const long long t = tb + ty;
    cpx<T> a = mk<T>((T)0, (T)0), b = mk<T>((T)0, (T)0);
    if (t < t1) {
      if (k0 + tx < N) a = x1[t * N + k0 + tx];
      if (l0 + tx < N) b = x2[t * N + l0 + tx];
    }
    Ar[ty][tx] = (double)a.x;
    Ai[ty][tx] = (double)a.y;
    Br[ty][tx] = (double)b.x;
    Bi[ty][tx] = (double)b.y;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {  // a * conj(b)
      accr += Ar[q][ty] * Br[q][tx] + Ai[q][ty] * Bi[q][tx];
      acci += Ai[q][ty] * Br[q][tx] - Ar[q][ty] * Bi[q][tx];
    }
    __syncthreads();
  }
  const int k = k0 + ty, l = l0 + tx;
  if (k < N && l < N) {
    const double sg = ((k + l) & 1) ? -1.0 : 1.0;
    const size_t o = ((size_t)((k + N / 2) % N) * N + (size_t)((l + N / 2) % N)) * 2;
    atomicAdd(out + o, sg * accr);
    atomicAdd(out + o + 1, sg * acci);
  }
}

// Cross-GPU barrier of the slab decomposition (one CTA, lane q talks to rank q): publish this rank's epoch in
// every peer's flag array (a release store at system scope; the peer stores of the preceding kernel are complete
// at its end), then wait until every peer has published the same epoch in ours.  Bounded spin: a rank that died
// must not hang the others -- after ~20 s the error word is set instead.
struct SlabBarrierParams {
  unsigned* peer_flags[GGP_MAX_PEERS];
  unsigned* my_flags;
  int* err;
  int P, me;
  unsigned epoch;
};
__global__ void slab_barrier_kernel(const SlabBarrierParams p) {
  const int q = threadIdx.x;
  if (q >= p.P || q == p.me) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_flags[q] + p.me), "r"(p.epoch) : "memory");
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.my_flags + q) : "memory");
    if ((int)(v - p.epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) {
      *p.err = 1;
      break;
    }
    __nanosleep(200);
  }
}

// ---- plan ----------------------------------------------------------------------------------------
enum { KC_ROW = 0, KC_STR_D = 1, KC_STR_FI = 2, KC_ONED = 3, KC_COUNT = 4 };

struct PlanBase {
  virtual ~PlanBase() {}
  virtual int create(const ggp_desc& d) = 0;
  virtual int set_state(const void* const* u) = 0;
  virtual int get_state(void* const* u) = 0;
  virtual int step(int64_t nsteps, const double* amp, const void* const* noise, const void* const* profiles = nullptr) = 0;
  virtual int observe(int kind, double* out) = 0;
  virtual int observe_windowed(const double* w1, const double* w2, double* out) = 0;
  virtual void* state_ptr(int c) = 0;
  virtual int ipc_export(void* blob) = 0;
  virtual int ipc_attach(const void* blobs) = 0;
  virtual int save_async(void* const* u) = 0;
  virtual int save_wait() = 0;
  virtual int64_t checkpoint_bytes() = 0;
  virtual int checkpoint_save(void* blob, uint64_t capacity) = 0;
  virtual int checkpoint_load(const void* blob, uint64_t size) = 0;

  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  int64_t dev_bytes = 0;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> pev;      // pairs
  std::vector<int> pev_class;
  double prof_ms[KC_COUNT] = {0, 0, 0, 0};
  int64_t prof_n[KC_COUNT] = {0, 0, 0, 0};
#ifdef GGP_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  int nranks = 1;
  // measurement aid: evict L2 after every kernel (timing hygiene for working sets smaller than L2)
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // step-level windows (ggp_profile_steps_*): one event pair per steady-state step, optional flush BETWEEN windows
  bool step_windows = false;
  void* sflush_buf = nullptr;
  size_t sflush_bytes = 0;
  std::vector<cudaEvent_t> sev;
  std::vector<int64_t> sev_steps;   // steps covered by each window (1 for 2-D/3-D, the chunk length for 1-D)
  double steps_ms = 0;
  int64_t steps_n = 0;
  bool window_open = false;

  int window_begin(int64_t nsteps_covered = 1) {
    if (!step_windows || window_open) return 0;
    cudaEvent_t a;
    GGP_CUDA(cudaEventCreate(&a));
    sev.push_back(a);
    sev_steps.push_back(nsteps_covered);
    GGP_CUDA(cudaEventRecord(a, stream));
    window_open = true;
    return 0;
  }
  int window_end() {
    if (!step_windows || !window_open) return 0;
    cudaEvent_t b;
    GGP_CUDA(cudaEventCreate(&b));
    sev.push_back(b);
    GGP_CUDA(cudaEventRecord(b, stream));
    window_open = false;
    if (sflush_buf) GGP_CUDA(cudaMemsetAsync(sflush_buf, 0, sflush_bytes, stream));
    return 0;
  }
  int windows_collect() {
    GGP_CUDA(cudaStreamSynchronize(stream));
    for (size_t i = 0; i + 1 < sev.size(); i += 2) {
      float ms = 0;
      GGP_CUDA(cudaEventElapsedTime(&ms, sev[i], sev[i + 1]));
      steps_ms += ms;
      steps_n += sev_steps[i / 2];
    }
    for (cudaEvent_t e : sev) cudaEventDestroy(e);
    sev.clear();
    sev_steps.clear();
    window_open = false;
    return 0;
  }

  int prof_begin(int cls) {
    if (!profiling) return 0;
    cudaEvent_t a, b;
    GGP_CUDA(cudaEventCreate(&a));
    GGP_CUDA(cudaEventCreate(&b));
    pev.push_back(a);
    pev.push_back(b);
    pev_class.push_back(cls);
    GGP_CUDA(cudaEventRecord(a, stream));
    return 0;
  }
  int prof_end() {
    if (profiling) GGP_CUDA(cudaEventRecord(pev.back(), stream));
    if (flush_buf) GGP_CUDA(cudaMemsetAsync(flush_buf, 0, flush_bytes, stream));
    return 0;
  }
  int prof_collect() {
    GGP_CUDA(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < pev_class.size(); ++i) {
      float ms = 0;
      GGP_CUDA(cudaEventElapsedTime(&ms, pev[2 * i], pev[2 * i + 1]));
      prof_ms[pev_class[i]] += ms;
      prof_n[pev_class[i]] += 1;
      cudaEventDestroy(pev[2 * i]);
      cudaEventDestroy(pev[2 * i + 1]);
    }
    pev.clear();
    pev_class.clear();
    return 0;
  }
};

}  // namespace ggp
#include "generic_plan.cuh"
namespace ggp {

template <typename T>
struct PlanT : PlanBase {
  int ndim = 0, M = 0;
  long long n[3] = {1, 1, 1};
  long long nspatial = 0, nbatch = 0, batch_offset = 0;
  double dt = 0;
  cpx<T>* u[2] = {nullptr, nullptr};
  cpx<T>* D[4] = {nullptr, nullptr, nullptr, nullptr};
  cpx<T>* Daos = nullptr;    // two-component diagonal / matrix exp_D, point-major (StrParams::Daos)
  int dcols = 0;
  cpx<T>* V[4] = {nullptr, nullptr, nullptr, nullptr};
  cpx<T>* S[2] = {nullptr, nullptr};
  typename TwT<T>::type* tw[3] = {nullptr, nullptr, nullptr};
  typename TwT<T>::type* tw_row = nullptr;   // twiddles of the row kernel where its radix differs from the default (row_E)
  void* tw2[3] = {nullptr, nullptr, nullptr};  // duplicated twiddles of the packed two-line kernels (fp32 plans)
  bool packed_ok = false;
  int dkind = 0;
  // separable scalar dispersion: exp_D = Dperp[all but the last axis] * Dline[last axis]
  bool sep = false;
  cpx<T>* Dperp = nullptr;
  cpx<T>* Dline = nullptr;
  typename TwT<T>::type* Dsp[2] = {nullptr, nullptr};  // the same factors as hi + lo pairs (fp32 plans)
  double2* Dq[2] = {nullptr, nullptr};                  // ... and in Float64 (quirk Q6, mixed_precision_tables)
  bool q6 = false;
  PointwiseParams<T> pw;
  bool has_pointwise = false;
  int pump_kind = 0, noise_kind = 0, noise_real = 0;
  std::complex<double> amp_prev = 0;
  uint64_t half_ctr = 0;
  // dense time-dependent pump: ring of three profile buffers (F_{k-1}, F_k, F_{k+1}: the fused contiguous-axis
  // kernel applies two half-steps), SoA planes of T; pd_k = index of the latest profile (F_0 = primed at init)
  cpx<T>* pdense[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  int pump_ncomp = 0, table_prec = GGP_C128;
  uint64_t pd_k = 0;
  cpx<T>* pd_pin[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // page-locked staging, one per ring slot
  cudaEvent_t pd_ev[3] = {nullptr, nullptr, nullptr};                                   // H2D copy of that slot done
  HalfStep<T>* hs_dev = nullptr;
  size_t hs_cap = 0;
  void* xi_dev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  double* obs_dev = nullptr;
  double* g1_dev = nullptr;  // M * N * N complex doubles (ggp_observe_windowed)
  cpx<T>* win_dev = nullptr;
  cpx<T>* xbuf1 = nullptr;
  double* g2_dev = nullptr;  // M * N * N doubles, allocated on first use (GGP_OBS_G2_MOMENTUM)
  cpx<T>* scratch[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  // 3-D slab decomposition (one process per GPU): this rank holds z-planes [prank*n3loc, (prank+1)*n3loc) of the
  // global (n1, n2, n3g) grid; for the z pass the data is transposed (NCCL all-to-all) into y-slabs
  // (n1, n2loc, n3g) held in xbuf.  sendbuf stages the strided pack / unpack.
  bool slab = false;
  int P = 1, prank = 0;
  long long n2g = 1, n3g = 1, n2loc = 1, n3loc = 1;
  cpx<T>* xbuf[2] = {nullptr, nullptr};
  cpx<T>* sendbuf[2] = {nullptr, nullptr};   // NCCL path only, allocated on first use
  // fused transpose over peer memory (CUDA IPC): every rank maps every other rank's u, xbuf and barrier flags;
  // the strided kernels then store their results directly into the owner's slab (StrParams::scatter)
  bool p2p = false;
  cpx<T>* peer_u[GGP_MAX_PEERS][2];
  cpx<T>* peer_x[GGP_MAX_PEERS][2];
  unsigned* peer_flags[GGP_MAX_PEERS];
  unsigned* flags = nullptr;
  int* bar_err = nullptr;
  unsigned epoch = 0;
  std::vector<void*> ipc_opened;
  cudaStream_t aux_stream = nullptr;      // second stream of the chunked local phase (slab_iry)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // copy-engine transposes (GGP_SLAB_CE): the scatter passes write a destination-ordered LOCAL staging buffer and the
  // copy engines push it to the peers while the next chunk computes -- the NVLink time leaves the SMs
  bool l2_persist = false;   // exp_D table pinned in L2 (setup_l2_persistence)
  bool slab_ce = false;
  cpx<T>* stage[2] = {nullptr, nullptr};
  cudaStream_t ce_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_chunk[16] = {}, ev_ce[2] = {nullptr, nullptr};
  int slab_chunks = 4;
  bool slab_yocc1 = false;   // GGP_SLAB_YOCC1=1: the chunked scatter pass runs at one CTA per SM (room for the other stream)
  // TMA path of the strided kernel, per strided axis (1, 2)
  bool tma_ok[3] = {false, false, false};
  bool tma_d[3] = {false, false, false};  // exp_D staged by TMA as well
  CUtensorMap tmap[3][2];
  CUtensorMap dmap[3][4];
  bool tma_persistent = false;  // opt-in: the older persistent TMA kernel (make TMA=1, GGP_TMA_PERSISTENT=1)
  int sm_count = 148;
  // streaming saves: snapshot on the compute stream, device -> host on a copy stream
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t snap_ready = nullptr, copy_done = nullptr;
  cpx<T>* snap[2] = {nullptr, nullptr};
  bool copy_pending = false;

  ~PlanT() override {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (copy_stream) {
      cudaStreamSynchronize(copy_stream);
      cudaStreamDestroy(copy_stream);
    }
    if (snap_ready) cudaEventDestroy(snap_ready);
    if (copy_done) cudaEventDestroy(copy_done);
    if (aux_stream) {
      cudaStreamSynchronize(aux_stream);
      cudaStreamDestroy(aux_stream);
    }
    for (int i = 0; i < 2; ++i) {
      if (ce_stream[i]) {
        cudaStreamSynchronize(ce_stream[i]);
        cudaStreamDestroy(ce_stream[i]);
      }
      if (ev_ce[i]) cudaEventDestroy(ev_ce[i]);
    }
    for (int i = 0; i < 16; ++i)
      if (ev_chunk[i]) cudaEventDestroy(ev_chunk[i]);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (int r = 0; r < 3; ++r) {
      if (pd_ev[r]) cudaEventDestroy(pd_ev[r]);
      for (int c = 0; c < 2; ++c)
        if (pd_pin[r][c]) cudaFreeHost(pd_pin[r][c]);
    }
    for (void* q : ipc_opened) cudaIpcCloseMemHandle(q);
    for (void* p : allocs) cudaFree(p);
    if (l2_persist) {
      cudaCtxResetPersistingL2Cache();
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    if (flush_buf) cudaFree(flush_buf);
    if (sflush_buf) cudaFree(sflush_buf);
    for (cudaEvent_t e : sev) cudaEventDestroy(e);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (cudaEvent_t e : pev) cudaEventDestroy(e);
#ifdef GGP_WITH_NCCL
    if (comm) ncclCommDestroy(comm);
#endif
    if (own_stream && stream) cudaStreamDestroy(stream);
  }

  int dalloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess)
      return fail(GGP_ERR_ALLOC, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    allocs.push_back(*p);
    dev_bytes += (int64_t)bytes;
    return 0;
  }

  // host AoS table (npts x ncols, `prec`) -> device SoA planes of T, scaled
  int upload_table(const void* host, int prec, int ncols, double scale, cpx<T>** planes) {
    std::vector<cpx<T>> tmp((size_t)nspatial);
    for (int c = 0; c < ncols; ++c) {
      if (prec == GGP_C128) {
        const std::complex<double>* h = (const std::complex<double>*)host;
        for (long long i = 0; i < nspatial; ++i) {
          const std::complex<double> z = h[i * ncols + c] * scale;
          tmp[(size_t)i] = mk<T>((T)z.real(), (T)z.imag());
        }
      } else {
        const std::complex<float>* h = (const std::complex<float>*)host;
        for (long long i = 0; i < nspatial; ++i) {
          const std::complex<float> z = h[i * ncols + c];
          tmp[(size_t)i] = mk<T>((T)((double)z.real() * scale), (T)((double)z.imag() * scale));
        }
      }
      int rc = dalloc((void**)&planes[c], sizeof(cpx<T>) * (size_t)nspatial);
      if (rc) return rc;
      GGP_CUDA(cudaMemcpy(planes[c], tmp.data(), sizeof(cpx<T>) * (size_t)nspatial, cudaMemcpyHostToDevice));
    }
    return 0;
  }

  // exp_D(k) = Dperp(k_1..k_{d-1}) * Dline(k_d) holds whenever the dispersion is a sum over axes.
  // Checked numerically on the host table; if it holds the strided kernel never reads the full table.
  int detect_separable(const ggp_desc& d) {
    if (dkind != GGP_TABLE_SCALAR || ndim < 2 || getenv("GGP_NO_SEP")) return 0;
    const long long nl = slab ? n3g : n[ndim - 1], np = nspatial / nl;
    std::vector<std::complex<double>> tab((size_t)nspatial);
    if (d.table_precision == GGP_C128) {
      memcpy(tab.data(), d.disp_table, sizeof(std::complex<double>) * (size_t)nspatial);
    } else {
      const std::complex<float>* h = (const std::complex<float>*)d.disp_table;
      for (long long i = 0; i < nspatial; ++i) tab[(size_t)i] = std::complex<double>(h[i].real(), h[i].imag());
    }
    const std::complex<double> d0 = tab[0];
    if (std::abs(d0) == 0 || !std::isfinite(std::abs(d0))) return 0;
    double dmax = 0, emax = 0;
    std::vector<std::complex<double>> line((size_t)nl), perp((size_t)np);
    for (long long l = 0; l < nl; ++l) line[(size_t)l] = tab[(size_t)(l * np)] / d0;
    for (long long q = 0; q < np; ++q) perp[(size_t)q] = tab[(size_t)q];
    for (long long l = 0; l < nl; ++l)
      for (long long q = 0; q < np; ++q) {
        const std::complex<double> z = tab[(size_t)(l * np + q)];
        dmax = std::max(dmax, std::abs(z));
        emax = std::max(emax, std::abs(z - perp[(size_t)q] * line[(size_t)l]));
      }
    // Tolerance.  fp64 plans: 1e-13 (keeps the accumulated deviation far below the 1e-10 parity gate).
    // fp32 plans: 4e-6 (~32 eps).  A ComplexF32 problem has Float32 grids, so the reference's table is
    // cis(-dt * fl32(D(k))) -- separable only up to eps32 * |phase| (C2: 1.2e-6 at the highest, unpopulated
    // modes, 3e-10 at the populated ones); the factorised product stays inside the table's own rounding.
    // GGP_NO_SEP=1 disables the fast path, GGP_SEP_TOL overrides the tolerance.
    double tol = (d.table_precision == GGP_C128 && sizeof(T) == 8) ? 1e-13 : 4e-6;
    if (d.disp_sep_tol > 0) tol = d.disp_sep_tol;  // the host verified in Float64 that D is a sum over axes (ggp.h)
    if (const char* e = getenv("GGP_SEP_TOL")) tol = atof(e);
    if (!(emax <= tol * dmax)) return 0;
    return upload_sep(perp, line);
  }

  // device copies of the two factors of a separable exp_D; 1/prod(n) (the reference's ScaledPlan, src/misc.jl:56)
  // is folded into D_perp
  int upload_sep(const std::vector<std::complex<double>>& perp, const std::vector<std::complex<double>>& line) {
    const long long np = (long long)perp.size(), nl = (long long)line.size();
    std::vector<cpx<T>> hp((size_t)np), hl((size_t)nl);
    const double sc = 1.0 / ((double)nspatial * P);
    for (long long q = 0; q < np; ++q) hp[(size_t)q] = mk<T>((T)(perp[(size_t)q].real() * sc), (T)(perp[(size_t)q].imag() * sc));
    for (long long l = 0; l < nl; ++l) hl[(size_t)l] = mk<T>((T)line[(size_t)l].real(), (T)line[(size_t)l].imag());
    int rc;
    if ((rc = dalloc((void**)&Dperp, sizeof(cpx<T>) * (size_t)np))) return rc;
    if ((rc = dalloc((void**)&Dline, sizeof(cpx<T>) * (size_t)nl))) return rc;
    GGP_CUDA(cudaMemcpy(Dperp, hp.data(), sizeof(cpx<T>) * (size_t)np, cudaMemcpyHostToDevice));
    GGP_CUDA(cudaMemcpy(Dline, hl.data(), sizeof(cpx<T>) * (size_t)nl, cudaMemcpyHostToDevice));
    if (q6 && sizeof(T) == 4) {
      std::vector<double2> qp((size_t)np), ql((size_t)nl);
      for (long long q = 0; q < np; ++q) qp[(size_t)q] = make_double2(perp[(size_t)q].real() * sc, perp[(size_t)q].imag() * sc);
      for (long long l = 0; l < nl; ++l) ql[(size_t)l] = make_double2(line[(size_t)l].real(), line[(size_t)l].imag());
      if ((rc = dalloc((void**)&Dq[0], sizeof(double2) * (size_t)np))) return rc;
      if ((rc = dalloc((void**)&Dq[1], sizeof(double2) * (size_t)nl))) return rc;
      GGP_CUDA(cudaMemcpy(Dq[0], qp.data(), sizeof(double2) * (size_t)np, cudaMemcpyHostToDevice));
      GGP_CUDA(cudaMemcpy(Dq[1], ql.data(), sizeof(double2) * (size_t)nl, cudaMemcpyHostToDevice));
    }
    if constexpr (TwT<T>::split) {
      std::vector<typename TwT<T>::type> sp((size_t)np), sl((size_t)nl);
      for (long long q = 0; q < np; ++q)
        sp[(size_t)q] = TwT<T>::make((long double)perp[(size_t)q].real() * sc, (long double)perp[(size_t)q].imag() * sc);
      for (long long l = 0; l < nl; ++l) sl[(size_t)l] = TwT<T>::make(line[(size_t)l].real(), line[(size_t)l].imag());
      if ((rc = dalloc((void**)&Dsp[0], sizeof(sp[0]) * (size_t)np))) return rc;
      if ((rc = dalloc((void**)&Dsp[1], sizeof(sl[0]) * (size_t)nl))) return rc;
      GGP_CUDA(cudaMemcpy(Dsp[0], sp.data(), sizeof(sp[0]) * (size_t)np, cudaMemcpyHostToDevice));
      GGP_CUDA(cudaMemcpy(Dsp[1], sl.data(), sizeof(sl[0]) * (size_t)nl, cudaMemcpyHostToDevice));
    }
    sep = true;
    return 0;
  }

  // GGP_TABLE_SEP_AXES: exp_D(k) = prod_a disp_axes[a][k_a] handed over as d short vectors (include/ggp.h)
  int setup_sep_axes(const ggp_desc& d) {
    for (int a = 0; a < ndim; ++a)
      if (!d.disp_axes[a]) return fail(GGP_ERR_INVALID, "disp_axes[a] is NULL (GGP_TABLE_SEP_AXES)");
    auto get = [&](int a, long long i) -> std::complex<double> {
      if (d.table_precision == GGP_C128) return ((const std::complex<double>*)d.disp_axes[a])[i];
      const std::complex<float> z = ((const std::complex<float>*)d.disp_axes[a])[i];
      return std::complex<double>(z.real(), z.imag());
    };
    if (ndim == 1) {
      std::vector<cpx<T>> h((size_t)n[0]);
      const double sc = 1.0 / (double)nspatial;
      for (long long i = 0; i < n[0]; ++i) {
        const std::complex<double> z = get(0, i) * sc;
        h[(size_t)i] = mk<T>((T)z.real(), (T)z.imag());
      }
      int rc = dalloc((void**)&D[0], sizeof(cpx<T>) * h.size());
      if (rc) return rc;
      GGP_CUDA(cudaMemcpy(D[0], h.data(), sizeof(cpx<T>) * h.size(), cudaMemcpyHostToDevice));
      return 0;
    }
    const long long nl = slab ? n3g : n[ndim - 1], np = nspatial / nl;
    std::vector<std::complex<double>> perp((size_t)np), line((size_t)nl);
    for (long long l = 0; l < nl; ++l) line[(size_t)l] = get(ndim - 1, l);
    if (ndim == 2) {
      for (long long q = 0; q < np; ++q) perp[(size_t)q] = get(0, q);
    } else {
      const long long m2 = slab ? n2loc : n[1], off2 = slab ? (long long)prank * n2loc : 0;
      for (long long j = 0; j < m2; ++j) {
        const std::complex<double> dj = get(1, off2 + j);
        for (long long i = 0; i < n[0]; ++i) perp[(size_t)(i + n[0] * j)] = get(0, i) * dj;
      }
    }
    return upload_sep(perp, line);
  }

  // twiddle table of one axis for a radix schedule of E elements per thread: per-pass coalesced blocks (fft_line.cuh),
  // then, for long lines, the compact tables of the factorised twiddles
  static void build_twiddles(long long N, long long E, std::vector<typename TwT<T>::type>& h) {
    const long double twopi = 2.0L * 3.14159265358979323846264338327950288L;
    for (long long NS = 1; NS < N;) {
      const long long R = (N / NS >= E) ? E : N / NS;
      if (NS > 1)
        for (long long r = 1; r < R; ++r)
          for (long long k = 0; k < NS; ++k) {
            const long double ang = -twopi * (long double)(r * k) / (long double)(NS * R);
            h.push_back(TwT<T>::make(cosl(ang), sinl(ang)));
          }
      NS *= R;
    }
    if (h.empty()) h.push_back(TwT<T>::make(1.0L, 0.0L));
    // long lines: compact tables of the factorised twiddles behind the pass blocks (fft_line.cuh, FACT):
    // [pad to an even count | B[j] = w_N^j, j < 64 | A[j] = w_N^(64 j), j < N/64]
    if (N >= 4096 && !TwT<T>::split) {
      if (h.size() % 2) h.push_back(TwT<T>::make(1.0L, 0.0L));
      for (long long j = 0; j < TWF_LO; ++j) {
        const long double ang = -twopi * (long double)j / (long double)N;
        h.push_back(TwT<T>::make(cosl(ang), sinl(ang)));
      }
      for (long long j = 0; j < N / TWF_LO; ++j) {
        const long double ang = -twopi * (long double)(j * TWF_LO) / (long double)N;
        h.push_back(TwT<T>::make(cosl(ang), sinl(ang)));
      }
    }
  }

  static int ncols_of(int kind, int M) {
    return kind == GGP_TABLE_SCALAR ? 1 : kind == GGP_TABLE_DIAG ? M : kind == GGP_TABLE_FULL ? M * M : 0;
  }

  int create(const ggp_desc& d) override {
    ndim = d.ndim;
    M = d.ncomp;
    nspatial = 1;
    for (int i = 0; i < 3; ++i) n[i] = 1;
    for (int i = 0; i < ndim; ++i) {
      n[i] = d.n[i];
      nspatial *= n[i];
    }
    nbatch = d.nbatch;
    batch_offset = d.batch_offset;
    dt = d.dt;
    dkind = d.disp_kind;
    q6 = d.mixed_precision_tables != 0 && d.table_precision == GGP_C128 && sizeof(T) == 4;
    // (more components / matrix-valued nonlinearities run on the generic plan, which has no slab decomposition)
    if (M > 2) return fail(GGP_ERR_UNSUPPORTED, "more than two components with a slab decomposition are not supported");
    if (d.nl_kind != GGP_NL_NONE && d.nl_kind != GGP_NL_DIAG)
      return fail(GGP_ERR_UNSUPPORTED, "matrix-valued nonlinearities with a slab decomposition are not supported");
    if (d.slab_nranks > 1) {
      if (ndim != 3 || nbatch != 1) return fail(GGP_ERR_UNSUPPORTED, "slab decomposition needs a 3-D grid without batch dims");
      if (d.noise_kind == GGP_NOISE_FIELD)
        return fail(GGP_ERR_UNSUPPORTED, "slab decomposition with a field- / position-dependent noise amplitude is not supported");
      P = d.slab_nranks;
      prank = d.slab_rank;
      if (prank < 0 || prank >= P || n[1] % P || n[2] % P) return fail(GGP_ERR_INVALID, "slab: n2 and n3 must be divisible by the number of ranks");
      slab = true;
      n2g = n[1];
      n3g = n[2];
      n2loc = n2g / P;
      n3loc = n3g / P;
      for (int i = 0; i < 3; ++i)
        if (!size_supported(n[i])) return fail(GGP_ERR_UNSUPPORTED, "slab: axis lengths must be supported powers of two");
      n[2] = n3loc;            // the resident layout is the z-slab (n1, n2, n3loc)
      nspatial = n[0] * n[1] * n3loc;
    }
    if (dkind != GGP_TABLE_NONE)
      for (int i = 0; i < ndim; ++i)
        if (!size_supported(n[i]))
          return fail(GGP_ERR_UNSUPPORTED, "FFT axis length " + std::to_string(n[i]) +
                                               " is not a supported power of two (SURVEY §8f N4)");
    if (d.pot_kind == GGP_TABLE_FULL && d.nl_kind != GGP_NL_NONE && !d.nl_scalar && M > 1)
      return fail(GGP_ERR_INVALID,
                  "SVector nonlinearity with SMatrix potential is a DimensionMismatch in the reference (src/kernels.jl:9)");

    if (d.stream) {
      stream = (cudaStream_t)d.stream;
    } else {
      GGP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      own_stream = true;
    }
    GGP_CUDA(cudaEventCreate(&ev0));
    GGP_CUDA(cudaEventCreate(&ev1));

    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) {
      int rc = dalloc((void**)&u[c], bytes);
      if (rc) return rc;
      GGP_CUDA(cudaMemsetAsync(u[c], 0, bytes, stream));
    }
    if (slab) {
      for (int c = 0; c < M; ++c) {
        int rc2;
        if ((rc2 = dalloc((void**)&xbuf[c], bytes))) return rc2;
      }
      int rc3;
      if ((rc3 = dalloc((void**)&flags, 256))) return rc3;
      GGP_CUDA(cudaMemsetAsync(flags, 0, 256, stream));
      bar_err = (int*)(flags + GGP_MAX_PEERS);
    }
    // tables.  The inverse transform is unnormalised on the device; the reference's 1/prod(n)
    // (ScaledPlan, src/misc.jl:56) is folded into exp_D -- exact for power-of-two sizes.
    int rc;
    if (dkind == GGP_TABLE_SEP_AXES) {
      if ((rc = setup_sep_axes(d))) return rc;
      dkind = GGP_TABLE_SCALAR;   // from here on: a scalar table that (for ndim >= 2) exists only as its two factors
    } else if (dkind != GGP_TABLE_NONE) {
      if (!d.disp_table) return fail(GGP_ERR_INVALID, "disp_table is NULL");
      // (slab: the table arrives in the y-slab layout (n1, n2loc, n3g); same number of points as the z-slab)
      if ((rc = upload_table(d.disp_table, d.table_precision, ncols_of(dkind, M), 1.0 / ((double)nspatial * P), D))) return rc;
      if ((rc = detect_separable(d))) return rc;
      if (ndim >= 2 && M == 2 && (dkind == GGP_TABLE_DIAG || dkind == GGP_TABLE_FULL) && !getenv("GGP_NO_DAOS")) {
        // the same table point-major for the strided kernels (see StrParams::Daos)
        dcols = ncols_of(dkind, M);
        const double sc = 1.0 / ((double)nspatial * P);
        std::vector<cpx<T>> tmp((size_t)nspatial * dcols);
        for (size_t i = 0; i < tmp.size(); ++i) {
          std::complex<double> z;
          if (d.table_precision == GGP_C128) z = ((const std::complex<double>*)d.disp_table)[i];
          else { const std::complex<float> f = ((const std::complex<float>*)d.disp_table)[i]; z = std::complex<double>(f.real(), f.imag()); }
          z *= sc;
          tmp[i] = mk<T>((T)z.real(), (T)z.imag());
        }
        if ((rc = dalloc((void**)&Daos, sizeof(cpx<T>) * tmp.size()))) return rc;
        GGP_CUDA(cudaMemcpy(Daos, tmp.data(), sizeof(cpx<T>) * tmp.size(), cudaMemcpyHostToDevice));
      }
    }
    {
      for (int a = 0; a < ndim; ++a) {
        const long long na = (slab && a == 2) ? n3g : n[a];
        for (int b = 0; b < a; ++b)
          if ((b == 2 && slab ? n3g : n[b]) == na) {
            tw[a] = tw[b];
            tw2[a] = tw2[b];
          }
        if (tw[a]) continue;
        // per-pass coalesced layout, see fft_line.cuh
        std::vector<typename TwT<T>::type> h;
        {
          const long long N = na;
          const long long E = default_E<T>((int)N);
          build_twiddles(N, E, h);
        }
        if ((rc = dalloc((void**)&tw[a], sizeof(h[0]) * h.size()))) return rc;
        GGP_CUDA(cudaMemcpy(tw[a], h.data(), sizeof(h[0]) * h.size(), cudaMemcpyHostToDevice));
#ifdef GGP_PACKED
        if constexpr (std::is_same<T, float>::value) {
          // packed kernels: radix schedule of default_E<f2>, every entry duplicated into both lanes
          if (M == 1 && packed_size(na) && !getenv("GGP_NO_PACKED")) {
            std::vector<float> h2;
            const long long N = na;
            const long long E = default_E<f2>((int)N);
            const long double twopi = 2.0L * 3.14159265358979323846264338327950288L;
            for (long long NS = 1; NS < N;) {
              const long long R = (N / NS >= E) ? E : N / NS;
              if (NS > 1)
                for (long long r = 1; r < R; ++r)
                  for (long long k = 0; k < NS; ++k) {
                    const long double ang = -twopi * (long double)(r * k) / (long double)(NS * R);
                    const float c = (float)cosl(ang), sn = (float)sinl(ang);
                    h2.push_back(c); h2.push_back(c); h2.push_back(sn); h2.push_back(sn);
                  }
              NS *= R;
            }
            if (h2.empty()) h2.assign(4, 0.f);
            if ((rc = dalloc(&tw2[a], sizeof(float) * h2.size()))) return rc;
            GGP_CUDA(cudaMemcpy(tw2[a], h2.data(), sizeof(float) * h2.size(), cudaMemcpyHostToDevice));
          }
        }
#endif
      }
    }
    memset(&pw, 0, sizeof(pw));
    pw.dt = (T)(dt / 2);
    pw.sqrt_dt = (T)std::sqrt(dt / 2);
    pw.vkind = d.pot_kind;
    if (d.pot_kind != GGP_TABLE_NONE) {
      if (!d.pot_table) return fail(GGP_ERR_INVALID, "pot_table is NULL");
      if ((rc = upload_table(d.pot_table, d.table_precision, ncols_of(d.pot_kind, M), 1.0, V))) return rc;
      for (int i = 0; i < 4; ++i) pw.expV[i] = V[i];
    }
    pump_kind = d.pump_kind;
    pump_ncomp = d.pump_ncomp;
    table_prec = d.table_precision;
    if (pump_kind == GGP_PUMP_DENSE) {
      if (!d.pump_table) return fail(GGP_ERR_INVALID, "pump_table is NULL");
      if (d.pump_ncomp != 1 && d.pump_ncomp != M) return fail(GGP_ERR_INVALID, "pump_ncomp must be 1 or ncomp");
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < pump_ncomp; ++c)
          if ((rc = dalloc((void**)&pdense[r][c], sizeof(cpx<T>) * (size_t)nspatial))) return rc;
      pw.pump = pump_ncomp == 1 ? 1 : 2;
      pw.pump_dense = 1;
      pd_k = 0;
      if ((rc = upload_profile(d.pump_table, 0))) return rc;   // F_0 = pump at tspan[1] (src/strang_splitting.jl:58)
    } else if (pump_kind != GGP_PUMP_NONE) {
      if (!d.pump_table) return fail(GGP_ERR_INVALID, "pump_table is NULL");
      if (d.pump_ncomp != 1 && d.pump_ncomp != M) return fail(GGP_ERR_INVALID, "pump_ncomp must be 1 or ncomp");
      if ((rc = upload_table(d.pump_table, d.table_precision, d.pump_ncomp, 1.0, S))) return rc;
      pw.pump = d.pump_ncomp == 1 ? 1 : 2;
      {  // spatially constant profile: keep the value in the kernel parameters
        bool cst = true;
        std::complex<double> v0[2];
        for (int c = 0; c < d.pump_ncomp && cst; ++c)
          for (long long i = 0; i < nspatial && cst; ++i) {
            std::complex<double> z;
            if (d.table_precision == GGP_C128) z = ((const std::complex<double>*)d.pump_table)[i * d.pump_ncomp + c];
            else { const std::complex<float> f = ((const std::complex<float>*)d.pump_table)[i * d.pump_ncomp + c]; z = std::complex<double>(f.real(), f.imag()); }
            if (i == 0) v0[c] = z;
            else if (z != v0[c]) cst = false;
          }
        if (cst && !getenv("GGP_NO_PUMP_CONST")) {
          pw.pump_const = 1;
          for (int c = 0; c < d.pump_ncomp; ++c) pw.S_const[c] = mk<T>((T)v0[c].real(), (T)v0[c].imag());
        }
      }
      if (d.pump_ncomp == 2 && !pw.pump_const) {
        // a component whose profile is identically zero (the exciton component of examples/exciton_polariton.jl:55-59)
        // is never read by the kernels
        for (int c = 0; c < 2; ++c) {
          bool zero = true;
          for (long long i = 0; i < nspatial && zero; ++i) {
            if (d.table_precision == GGP_C128) zero = ((const std::complex<double>*)d.pump_table)[i * 2 + c] == std::complex<double>(0, 0);
            else zero = ((const std::complex<float>*)d.pump_table)[i * 2 + c] == std::complex<float>(0, 0);
          }
          pw.pump_zero[c] = zero ? 1 : 0;
        }
      }
      pw.S[0] = S[0];
      pw.S[1] = S[1];
      amp_prev = std::complex<double>(d.pump_amp0[0], d.pump_amp0[1]);
    }
    if (d.nl_kind == GGP_NL_DIAG) {
      bool cplx = false;
      for (int i = 0; i < M; ++i) {
        const int src = d.nl_scalar ? 0 : i;
        pw.nl_c_re[i] = (T)d.nl_c[src][0];
        pw.nl_c_im[i] = (T)d.nl_c[src][1];
        if (d.nl_c[src][1] != 0) cplx = true;
        for (int j = 0; j < M; ++j) {
          pw.nl_g_re[i][j] = (T)d.nl_g[src][j][0];
          pw.nl_g_im[i][j] = (T)d.nl_g[src][j][1];
          if (d.nl_g[src][j][1] != 0) cplx = true;
        }
      }
      pw.nl = cplx ? 2 : 1;
    }
    noise_kind = d.noise_kind;
    noise_real = d.noise_real;
    if (noise_kind != GGP_NOISE_NONE) {
      pw.noise = NOISE_PHILOX;
      pw.noise_real = d.noise_real;
      for (int i = 0; i < M; ++i) pw.eta[i] = mk<T>((T)d.noise_eta[i][0], (T)d.noise_eta[i][1]);
      if (noise_kind == GGP_NOISE_FIELD) {
        pw.noise_field = 1;
        pw.n1 = (int)n[0];
        for (int i = 0; i < M; ++i)
          for (int j = 0; j < M; ++j) pw.alpha[i][j] = mk<T>((T)d.noise_alpha[i][j][0], (T)d.noise_alpha[i][j][1]);
        if (d.noise_profile) {
          std::vector<cpx<T>> hp((size_t)n[0]);
          const double* src = (const double*)d.noise_profile;
          for (long long k = 0; k < n[0]; ++k) hp[(size_t)k] = mk<T>((T)src[2 * k], (T)src[2 * k + 1]);
          cpx<T>* dp = nullptr;
          if ((rc = dalloc((void**)&dp, sizeof(cpx<T>) * hp.size()))) return rc;
          GGP_CUDA(cudaMemcpy(dp, hp.data(), sizeof(cpx<T>) * hp.size(), cudaMemcpyHostToDevice));
          pw.nprof = dp;
        }
      } else if (noise_kind != GGP_NOISE_CONST) {
        return fail(GGP_ERR_INVALID, "unknown noise_kind");
      }
      pw.seed_lo = (uint32_t)d.seed;
      pw.seed_hi = (uint32_t)(d.seed >> 32);
      // global element index of this plan's element 0: trajectory shard, or z-slab (contiguous in the global array),
      // so the Philox stream does not depend on the decomposition
      pw.elem_offset = slab ? (long long)prank * nspatial : batch_offset * nspatial;
    }
    // products of plan constants for the Truncated-Wigner variant of the half-step (PW_TW, pointwise.cuh)
    for (int i = 0; i < M; ++i) {
      const int src = d.nl_scalar ? 0 : i;
      const bool has_nl = d.nl_kind == GGP_NL_DIAG;
      pw.nl_cd[i] = (T)(has_nl ? -(dt / 2) * d.nl_c[src][0] : 0.0);
      for (int j = 0; j < M; ++j) pw.nl_gd[i][j] = (T)(has_nl ? -(dt / 2) * d.nl_g[src][j][0] : 0.0);
      const double sq = std::sqrt(dt / 2);
      pw.eta_s[i] = noise_kind != GGP_NOISE_NONE ? mk<T>((T)(sq * d.noise_eta[i][1]), (T)(-sq * d.noise_eta[i][0]))
                                                 : mk<T>((T)0, (T)0);
    }
    has_pointwise = pw.vkind || pw.pump || pw.nl || pw.noise;
    {
      const int eo = size_supported(n[0]) ? row_E_of<T>((int)n[0], M, pw_variant()) : 0;
      if (eo > 0 && eo != default_E<T>((int)n[0])) {
        std::vector<typename TwT<T>::type> h;
        build_twiddles(n[0], eo, h);
        if ((rc = dalloc((void**)&tw_row, sizeof(h[0]) * h.size()))) return rc;
        GGP_CUDA(cudaMemcpy(tw_row, h.data(), sizeof(h[0]) * h.size(), cudaMemcpyHostToDevice));
      }
    }
    if ((rc = setup_tma())) return rc;
    if ((rc = setup_l2_persistence())) return rc;
    if ((rc = dalloc((void**)&obs_dev, sizeof(double) * (size_t)(nspatial * M + 8)))) return rc;
    GGP_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }

  // Keep the point-major exp_D table of a two-component plan resident in L2 (persisting access-policy window on the
  // plan's stream).  The table is a constant that every step re-reads in full (C3: 64 MB of 112 MB touched per step,
  // L2 hit rate 43 % without this, ncu r02q); the state streams through the rest of the 126 MB.  Only for plans that
  // own their stream.  MEASURED (r02r, one gpurun call): C3 c128 (64 MB table) 105.6 -> 111.9 us/step -- the pinned
  // table takes the L2 the state needs; c64 (32 MB table) strided pass 36.9 -> 34.9 us, contiguous-axis pass 36.2 -> 37.4,
  // step unchanged.  OFF by default; GGP_L2_PERSIST=1 switches it on.
  int setup_l2_persistence() {
    if (!Daos || !own_stream) return 0;
    const char* e = getenv("GGP_L2_PERSIST");
    if (!e || atoi(e) == 0) return 0;
    cudaDeviceProp prop;
    GGP_CUDA(cudaGetDeviceProperties(&prop, device));
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)dcols;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return 0;
    const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = (void*)Daos;
    attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    l2_persist = true;
    return 0;
  }

  // Tensor maps for the TMA-fed strided kernel: the field as a rank-4 tensor of reals
  // [2*n1, n2, n3, batch]; box = [2*W reals] x [min(N,256) rows along the strided axis].
  int setup_tma() {
    cudaDeviceProp prop;
    GGP_CUDA(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    // Tensor maps for the TMA staging of the strided kernel (StrParams::tma).  GGP_NO_TMA=1: plain LDG/STG.
    // (The older persistent TMA kernel, str_tma.cuh, needs make TMA=1 and GGP_TMA_PERSISTENT=1.)
    if (getenv("GGP_NO_TMA") || slab) return 0;
#ifdef GGP_TMA
    tma_persistent = getenv("GGP_TMA_PERSISTENT") != nullptr;
#endif
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      cudaGetLastError();
      return 0;  // no TMA descriptors available: the LDG version of the kernel is used
    }
    EncodeFn encode = (EncodeFn)fn;
    CUtensorMapL2promotion l2promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char* e = getenv("GGP_TMA_L2PROMO")) {  // tuning knob: 0 / 64 / 128 / 256
      const int v = atoi(e);
      l2promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }
    for (int ax = 1; ax < ndim; ++ax) {
      int W = 0, LS = 0, threads = 0, us = 0;
      dispatch_str_query<T>((int)n[ax], M, ax, n[0], &W, &LS, &threads, &us);
      if (!W || !us) continue;
      const size_t esz = sizeof(cpx<T>);
      if ((size_t)W * esz < 16 || (n[0] * esz) % 16 != 0) continue;
      const size_t smem = esz * ((tma_persistent ? (size_t)M * n[ax] * W : 0) + (size_t)W * M * LS) + 16;
      if (smem > (size_t)prop.sharedMemPerBlockOptin) continue;
      const cuuint32_t rows = (cuuint32_t)(n[ax] < 256 ? n[ax] : 256);
      bool ok = true;
      for (int c = 0; c < M && ok; ++c) {
        cuuint64_t gdim[4] = {(cuuint64_t)(2 * n[0]), (cuuint64_t)n[1], (cuuint64_t)n[2], (cuuint64_t)nbatch};
        cuuint64_t gstr[3] = {(cuuint64_t)(n[0] * esz), (cuuint64_t)(n[0] * n[1] * esz), (cuuint64_t)(nspatial * esz)};
        cuuint32_t box[4] = {(cuuint32_t)(2 * W), ax == 1 ? rows : 1u, ax == 2 ? rows : 1u, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmap[ax][c], sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                            4, (void*)u[c], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, l2promo,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) ok = false;
      }
      tma_ok[ax] = ok;
      // exp_D planes through TMA as well when the shared memory budget allows (only the axis that multiplies)
      const int nplanes = ncols_of(dkind, M);
      const bool last_axis = ax == ndim - 1;
      if (ok && tma_persistent && last_axis && nplanes > 0 && !getenv("GGP_NO_TMA_D")) {
        const size_t smem_d = smem + esz * (size_t)nplanes * n[ax] * W;
        if (smem_d <= (size_t)prop.sharedMemPerBlockOptin) {
          bool okd = true;
          for (int pl = 0; pl < nplanes && okd; ++pl) {
            cuuint64_t gdim[4] = {(cuuint64_t)(2 * n[0]), (cuuint64_t)n[1], (cuuint64_t)n[2], 1};
            cuuint64_t gstr[3] = {(cuuint64_t)(n[0] * esz), (cuuint64_t)(n[0] * n[1] * esz), (cuuint64_t)(nspatial * esz)};
            cuuint32_t box[4] = {(cuuint32_t)(2 * W), ax == 1 ? rows : 1u, ax == 2 ? rows : 1u, 1u};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(&dmap[ax][pl], sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                4, (void*)D[pl], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) okd = false;
          }
          tma_d[ax] = okd;
        }
      }
    }
    return 0;
  }

  int set_state(const void* const* uh) override {
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) GGP_CUDA(cudaMemcpyAsync(u[c], uh[c], bytes, cudaMemcpyHostToDevice, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int get_state(void* const* uh) override {
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) GGP_CUDA(cudaMemcpyAsync(uh[c], u[c], bytes, cudaMemcpyDeviceToHost, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    if (p2p) return barrier_status();
    return 0;
  }
  void* state_ptr(int c) override { return (c >= 0 && c < M) ? (void*)u[c] : nullptr; }

  // ---- streaming saves (replaces `map(copy!, slice, iter.u)`, src/fixed_time_stepping.jl:48, without stalling the
  // step chain): the state is snapshotted device-to-device on the compute stream (tens of microseconds), the slow
  // PCIe transfer of the snapshot runs on a second stream while the next save interval is already stepping.
  // One snapshot buffer: a new save waits (on the device, not on the host) for the previous transfer to drain.
  int save_async(void* const* uh) override {
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
    int rc;
    if (!copy_stream) {
      GGP_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      GGP_CUDA(cudaEventCreateWithFlags(&snap_ready, cudaEventDisableTiming));
      GGP_CUDA(cudaEventCreateWithFlags(&copy_done, cudaEventDisableTiming));
    }
    for (int c = 0; c < M; ++c)
      if (!snap[c] && (rc = dalloc((void**)&snap[c], bytes))) return rc;
    if (copy_pending) GGP_CUDA(cudaStreamWaitEvent(stream, copy_done, 0));
    for (int c = 0; c < M; ++c) GGP_CUDA(cudaMemcpyAsync(snap[c], u[c], bytes, cudaMemcpyDeviceToDevice, stream));
    GGP_CUDA(cudaEventRecord(snap_ready, stream));
    GGP_CUDA(cudaStreamWaitEvent(copy_stream, snap_ready, 0));
    for (int c = 0; c < M; ++c) GGP_CUDA(cudaMemcpyAsync(uh[c], snap[c], bytes, cudaMemcpyDeviceToHost, copy_stream));
    GGP_CUDA(cudaEventRecord(copy_done, copy_stream));
    copy_pending = true;
    return 0;
  }
  int save_wait() override {
    if (copy_pending) {
      GGP_CUDA(cudaEventSynchronize(copy_done));
      copy_pending = false;
    }
    if (p2p) return barrier_status();
    return 0;
  }

  // ---- checkpoint / resume (absent in the reference).  Everything a bit-identical continuation
  // needs that is not in the descriptor: the fields, the half-step counter (Philox counter word and pump phase)
  // and F_now's amplitude.  A restart from a saved slice alone is NOT identical because of the one-dt-late pump
  // (quirk Q1) and the noise counters.
  struct CkptHeader {
    uint64_t magic;
    uint32_t version, precision;
    int32_t ndim, ncomp;
    int64_t n[3], nbatch, batch_offset;
    uint64_t half_ctr;
    double amp_prev[2];
    uint64_t bytes_per_comp;
  };
  static constexpr uint64_t CKPT_MAGIC = 0x54504b4350474721ull;  // "!GGPCKPT"
  int64_t checkpoint_bytes() override {
    return (int64_t)(sizeof(CkptHeader) + (size_t)M * sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch);
  }
  int checkpoint_save(void* blob, uint64_t capacity) override {
    if (pump_kind == GGP_PUMP_DENSE)
      return fail(GGP_ERR_UNSUPPORTED, "checkpoints of dense-pump plans are not supported (F_now is a full profile, not an amplitude)");
    if ((int64_t)capacity < checkpoint_bytes()) return fail(GGP_ERR_INVALID, "checkpoint buffer too small (ggp_checkpoint_bytes)");
    CkptHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = CKPT_MAGIC;
    h.version = 1;
    h.precision = sizeof(T) == 4 ? GGP_C64 : GGP_C128;
    h.ndim = ndim;
    h.ncomp = M;
    for (int i = 0; i < 3; ++i) h.n[i] = n[i];
    h.nbatch = nbatch;
    h.batch_offset = batch_offset;
    h.half_ctr = half_ctr;
    h.amp_prev[0] = amp_prev.real();
    h.amp_prev[1] = amp_prev.imag();
    h.bytes_per_comp = sizeof(cpx<T>) * (uint64_t)nspatial * (uint64_t)nbatch;
    memcpy(blob, &h, sizeof(h));
    char* q = (char*)blob + sizeof(h);
    for (int c = 0; c < M; ++c)
      GGP_CUDA(cudaMemcpyAsync(q + (size_t)c * h.bytes_per_comp, u[c], h.bytes_per_comp, cudaMemcpyDeviceToHost, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int checkpoint_load(const void* blob, uint64_t size) override {
    if (size < sizeof(CkptHeader)) return fail(GGP_ERR_INVALID, "checkpoint truncated");
    CkptHeader h;
    memcpy(&h, blob, sizeof(h));
    if (h.magic != CKPT_MAGIC || h.version != 1) return fail(GGP_ERR_INVALID, "not a ggp checkpoint (magic/version)");
    const bool same = h.precision == (uint32_t)(sizeof(T) == 4 ? GGP_C64 : GGP_C128) && h.ndim == ndim && h.ncomp == M &&
                      h.n[0] == n[0] && h.n[1] == n[1] && h.n[2] == n[2] && h.nbatch == nbatch &&
                      h.batch_offset == batch_offset &&
                      h.bytes_per_comp == sizeof(cpx<T>) * (uint64_t)nspatial * (uint64_t)nbatch;
    if (!same) return fail(GGP_ERR_INVALID, "checkpoint was written by a plan of a different shape / precision / shard");
    if (size < sizeof(h) + (uint64_t)M * h.bytes_per_comp) return fail(GGP_ERR_INVALID, "checkpoint truncated");
    const char* q = (const char*)blob + sizeof(h);
    for (int c = 0; c < M; ++c)
      GGP_CUDA(cudaMemcpyAsync(u[c], q + (size_t)c * h.bytes_per_comp, h.bytes_per_comp, cudaMemcpyHostToDevice, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    half_ctr = h.half_ctr;
    amp_prev = std::complex<double>(h.amp_prev[0], h.amp_prev[1]);
    return 0;
  }

  // dense pump: host profile (point-major, then component; table_precision) -> ring slot, SoA planes of T
  int upload_profile(const void* host, int slot) {
    if (!host) return fail(GGP_ERR_INVALID, "pump profile pointer is NULL");
    if (!pd_ev[slot]) {
      GGP_CUDA(cudaEventCreateWithFlags(&pd_ev[slot], cudaEventDisableTiming));
      for (int c = 0; c < pump_ncomp; ++c) GGP_CUDA(cudaMallocHost((void**)&pd_pin[slot][c], sizeof(cpx<T>) * (size_t)nspatial));
    } else {
      GGP_CUDA(cudaEventSynchronize(pd_ev[slot]));   // the previous transfer out of this staging buffer has finished
    }
    for (int c = 0; c < pump_ncomp; ++c) {
      cpx<T>* pd_stage = pd_pin[slot][c];
      if (table_prec == GGP_C128) {
        const std::complex<double>* h = (const std::complex<double>*)host;
        for (long long i = 0; i < nspatial; ++i) {
          const std::complex<double> z = h[i * pump_ncomp + c];
          pd_stage[(size_t)i] = mk<T>((T)z.real(), (T)z.imag());
        }
      } else {
        const std::complex<float>* h = (const std::complex<float>*)host;
        for (long long i = 0; i < nspatial; ++i) {
          const std::complex<float> z = h[i * pump_ncomp + c];
          pd_stage[(size_t)i] = mk<T>((T)z.real(), (T)z.imag());
        }
      }
      GGP_CUDA(cudaMemcpyAsync(pdense[slot][c], pd_stage, sizeof(cpx<T>) * (size_t)nspatial, cudaMemcpyHostToDevice, stream));
    }
    GGP_CUDA(cudaEventRecord(pd_ev[slot], stream));
    return 0;
  }

  // half-step `half` (0/1) of the step whose pump amplitudes are (a_now, a_next)
  HalfStep<T> make_half(std::complex<double> a_now, std::complex<double> a_next, int slot) {
    HalfStep<T> h;
    memset(&h, 0, sizeof(h));
    const double q = dt / 4;  // (dt/2)/2, src/kernels.jl:45-46 with δt = dt/2
    h.fnow = mk<T>((T)(q * a_now.real()), (T)(q * a_now.imag()));
    h.fnext = mk<T>((T)(q * a_next.real()), (T)(q * a_next.imag()));
    if (pw.pump && pw.pump_const) {   // PW_TW: the pump terms of a spatially constant profile, ready-made
      for (int i = 0; i < M; ++i) {
        const cpx<T> sc = pw.S_const[pw.pump == 1 ? 0 : i];
        const std::complex<double> sv((double)sc.x, (double)sc.y);
        const std::complex<double> a = q * a_now * sv, b = q * a_next * sv;
        h.aS[i] = mk<T>((T)a.real(), (T)a.imag());
        h.bS[i] = mk<T>((T)b.real(), (T)b.imag());
      }
    }
    h.ctr = (uint32_t)half_ctr;
    h.ctr_hi = (uint32_t)(half_ctr >> 32);
    h.apply = has_pointwise ? 1 : 0;
    h.xi[0] = xi_dev[slot][0];
    h.xi[1] = xi_dev[slot][1];
    ++half_ctr;
    return h;
  }

  int upload_noise(const void* const* noise, int64_t s, int half, int slot) {
    const size_t esz = noise_real ? sizeof(T) : sizeof(cpx<T>);
    const size_t bytes = esz * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) {
      if (!xi_dev[slot][c]) {
        int rc = dalloc(&xi_dev[slot][c], bytes);
        if (rc) return rc;
      }
      GGP_CUDA(cudaMemcpyAsync(xi_dev[slot][c], noise[(s * 2 + half) * M + c], bytes, cudaMemcpyHostToDevice, stream));
    }
    return 0;
  }

  // zc0 / zcn: chunk of the local z-planes (slab plans, see slab_iry); zcn < 0: everything.  st: launch stream.
  int run_row(bool pre, bool post, const HalfStep<T>& hA, const HalfStep<T>& hB, long long zc0 = 0, long long zcn = -1,
              cudaStream_t st = nullptr) {
    if (!st) st = stream;
    RowParams<T> p;
    memset(&p, 0, sizeof(p));
    p.u[0] = u[0];
    p.u[1] = u[1];
    p.tw = tw_row ? tw_row : tw[0];
    p.nlines = (nspatial / n[0]) * nbatch;
    p.lines_per_image = nspatial / n[0];
    if (zcn >= 0) {
      p.line0 = zc0 * n[1];
      p.nlines = zcn * n[1];
    }
    p.pw = pw;
    p.hs[0] = hA;
    p.hs[1] = hB;
    p.flags = (pre ? 1 : 0) | (post ? 2 : 0);
    int rc = prof_begin(KC_ROW);
    if (rc) return rc;
#ifdef GGP_PACKED
    if constexpr (std::is_same<T, float>::value) {
      // two rows per thread group with packed fp32x2 arithmetic (packed.cuh)
      const bool any = hA.apply || hB.apply;
      if (tw2[0] && M == 1 && (p.nlines % 2 == 0) && (!any || pw_variant() == PW_KERR)) {
        p.tw2 = tw2[0];
        GGP_LAUNCH(dispatch_row2((int)n[0], p, stream), "row2_kernel");
        ++launches;
        return prof_end();
      }
    }
#endif
    GGP_LAUNCH(dispatch_row<T>((int)n[0], M, pw_variant(), p, st), "row_kernel");
    ++launches;
    return prof_end();
  }

  // strided pass along axis `ax` (1 or 2).  yslab: operate on xbuf in the transposed (n1, n2loc, n3g) layout.
  // scatter: 0 in place; 1 = y pass of the z-slab, results into the y-slabs (xbuf) of their owners; 2 = z pass of
  // the y-slab, results into the z-slabs (u) of their owners (fused all-to-all transpose over peer memory).
  // to_stage: the scatter pass writes the local staging buffer (copy-engine transposes) instead of peer memory
  int run_str(int ax, int mode, bool yslab = false, int scatter = 0, long long zc0 = 0, long long zcn = -1,
              cudaStream_t st = nullptr, bool to_stage = false) {
    if (!st) st = stream;
    StrParams<T> p;
    memset(&p, 0, sizeof(p));
    const long long g0 = n[0], g1 = yslab ? n2loc : n[1], g2 = yslab ? n3g : n[2];
    p.u[0] = yslab ? xbuf[0] : u[0];
    p.u[1] = yslab ? xbuf[1] : u[1];
    p.tw = tw[ax];
    for (int i = 0; i < 4; ++i) p.D[i] = D[i];
    p.Daos = Daos;
    p.dcols = dcols;
    p.dkind = dkind;
    if (sep && mode == 1) {
      p.D[0] = Dperp;
      p.D[1] = Dline;
      p.Dsp[0] = Dsp[0];
      p.Dsp[1] = Dsp[1];
      p.Dq[0] = Dq[0];
      p.Dq[1] = Dq[1];
      p.dkind = KIND_SEP;
    }
    p.mode = mode;
    p.ax = ax;
    p.slab = slab ? 1 : 0;
    long long nother;
    if (ax == 1) {
      p.ls = g0;
      p.no1 = g2;
      p.s1 = g0 * g1;
      p.s2 = nspatial;
      p.ts1 = g0 * g1;
      nother = g2 * nbatch;
    } else {
      p.ls = g0 * g1;
      p.no1 = g1;
      p.s1 = g0;
      p.s2 = nspatial;
      p.ts1 = g0;
      nother = g1 * nbatch;
    }
    const int N = (int)(ax == 1 ? g1 : g2);
    if (scatter) {
      p.scatter = scatter;
      for (int q = 0; q < P; ++q)
        for (int c = 0; c < M; ++c) p.dst[q][c] = scatter == 1 ? peer_x[q][c] : peer_u[q][c];
      if (scatter == 1) {          // (x, j, zl) of u  ->  rank j / n2loc, (x, j % n2loc, prank*n3loc + zl) of its xbuf
        p.dst_shift = ilog2((int)n2loc);
        p.dst_ls = n[0];
        p.dst_s1 = n[0] * n2loc;
        p.dst_base = n[0] * n2loc * (long long)prank * n3loc;
      } else {                     // (x, yl, j) of xbuf  ->  rank j / n3loc, (x, prank*n2loc + yl, j % n3loc) of its u
        p.dst_shift = ilog2((int)n3loc);
        p.dst_ls = n[0] * n2g;
        p.dst_s1 = n[0];
        p.dst_base = n[0] * (long long)prank * n2loc;
      }
      p.dst_s2 = 0;
      if (to_stage) {
        // destination-ordered staging: block q (n1 * n2loc * n3loc elements) is what rank q receives
        const long long blk = n[0] * n2loc * n3loc;
        if (scatter == 1) {
          // [zl][jl][x], the order of the peer's y-slab restricted to my z-range: same strides, contiguous per peer;
          // my own block still goes straight into my own xbuf
          for (int q = 0; q < P; ++q)
            for (int c = 0; c < M; ++c)
              if (q != prank) p.dst[q][c] = stage[c] + (long long)q * blk - p.dst_base;
        } else {
          // [zl][yl][x] per peer (own block included: it is copied like the others)
          p.dst_ls = n[0] * n2loc;
          p.dst_s1 = n[0];
          p.dst_base = 0;
          for (int q = 0; q < P; ++q)
            for (int c = 0; c < M; ++c) p.dst[q][c] = stage[c] + (long long)q * blk;
        }
      }
    }
    if (zcn >= 0) {
      // chunk of the slowest remaining axis: z-planes of the z-slab (y passes), y-rows of the y-slab (z pass)
      if ((ax == 1) == yslab) return fail(GGP_ERR_INVALID, "chunks apply to the y passes of the z-slab and the z pass of the y-slab");
      for (int c = 0; c < M; ++c) p.u[c] += zc0 * p.s1;
      p.tbase = zc0 * p.ts1;
      p.no1 = zcn;
      nother = zcn * nbatch;
      if (scatter) p.dst_base += zc0 * p.dst_s1;
      if (scatter && slab_yocc1 && !to_stage) p.occ1 = 1;
    }
    int rc = prof_begin(mode == 1 ? KC_STR_D : KC_STR_FI);
    if (rc) return rc;
    if (tma_ok[ax] && !slab && !yslab && !scatter && !tma_persistent) {
      p.tma = 1;
      p.ax = ax;
      for (int c = 0; c < M; ++c) p.map[c] = tmap[ax][c];
    }
#ifdef GGP_TMA
    if (tma_ok[ax] && !slab && tma_persistent) {
      StrTmaParams<T> q;
      memset(&q, 0, sizeof(q));
      for (int c = 0; c < M; ++c) q.map[c] = tmap[ax][c];
      q.u[0] = u[0];
      q.u[1] = u[1];
      q.tw = p.tw;
      q.ls = p.ls;
      q.no1 = p.no1;
      q.s1 = p.s1;
      q.s2 = p.s2;
      q.ts1 = p.ts1;
      for (int i = 0; i < 4; ++i) q.D[i] = p.D[i];
      q.dkind = p.dkind;
      q.mode = mode;
      q.ax = ax;
      q.nplanes = ncols_of(dkind, M);
      q.stage_d = (mode == 1 && tma_d[ax] && p.dkind != KIND_SEP) ? 1 : 0;
      for (int i = 0; i < 4; ++i) q.dmap[i] = dmap[ax][i];
      GGP_LAUNCH(dispatch_str_tma<T>(N, M, q, g0, nother, sm_count, stream), "str_tma_kernel");
      ++launches;
      return prof_end();
    }
#endif
#ifdef GGP_PACKED
    if constexpr (std::is_same<T, float>::value) {
      const bool dk_ok = mode != 1 || p.dkind == KIND_NONE || p.dkind == KIND_SCALAR || p.dkind == KIND_SEP;
      if (tw2[ax] && M == 1 && g0 >= 4 && dk_ok) {
        p.tw2 = tw2[ax];
        GGP_LAUNCH(dispatch_str2(N, p, g0, nother, stream), "str2_kernel");
        ++launches;
        return prof_end();
      }
    }
#endif
    GGP_LAUNCH(dispatch_str<T>(N, M, p, g0, nother, st), "str_kernel");
    ++launches;
    return prof_end();
  }

  // Slab plans, fused transposes: the local part of a step -- [inverse y pass of step s-1] -> contiguous-axis kernel ->
  // [forward y pass of step s, its results stored straight into the peers' y-slabs over NVLink] -- touches every
  // z-plane independently, so it runs in `slab_chunks` chunks of planes alternating between two streams: while one
  // chunk's forward pass is bound by its NVLink stores, the next chunk's local kernels (HBM-bound) run beside it.
  int ce_setup() {
    if (ce_stream[0]) return 0;
    int rc;
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial;
    for (int c = 0; c < M; ++c)
      if ((rc = dalloc((void**)&stage[c], bytes))) return rc;
    for (int i = 0; i < 2; ++i) {
      GGP_CUDA(cudaStreamCreateWithFlags(&ce_stream[i], cudaStreamNonBlocking));
      GGP_CUDA(cudaEventCreateWithFlags(&ev_ce[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 16; ++i) GGP_CUDA(cudaEventCreateWithFlags(&ev_chunk[i], cudaEventDisableTiming));
    return 0;
  }
  // the plan's stream waits for every copy issued so far on the copy-engine streams
  int ce_join() {
    for (int i = 0; i < 2; ++i) {
      GGP_CUDA(cudaEventRecord(ev_ce[i], ce_stream[i]));
      GGP_CUDA(cudaStreamWaitEvent(stream, ev_ce[i], 0));
    }
    return 0;
  }
  // copies of one chunk of the forward transpose: z-planes [z0, z1) of every remote block, contiguous on both sides
  int ce_push_y(int chunk, cudaStream_t produced_on, long long z0, long long z1) {
    cudaStream_t cs = ce_stream[chunk & 1];
    GGP_CUDA(cudaEventRecord(ev_chunk[chunk & 15], produced_on));
    GGP_CUDA(cudaStreamWaitEvent(cs, ev_chunk[chunk & 15], 0));
    const long long blk = n[0] * n2loc * n3loc, plane = n[0] * n2loc;
    const long long dbase = plane * (long long)prank * n3loc;
    for (int i = 1; i < P; ++i) {
      const int q = (prank + i) % P;          // every rank starts with a different peer
      for (int c = 0; c < M; ++c)
        GGP_CUDA(cudaMemcpyAsync(peer_x[q][c] + dbase + z0 * plane, stage[c] + (long long)q * blk + z0 * plane,
                                 sizeof(cpx<T>) * (size_t)((z1 - z0) * plane), cudaMemcpyDefault, cs));
    }
    return 0;
  }
  // copies of one chunk of the backward transpose: rows [y0, y1) of every block, n3loc pieces of n1 * (y1 - y0)
  int ce_push_z(int chunk, cudaStream_t produced_on, long long y0, long long y1) {
    cudaStream_t cs = ce_stream[chunk & 1];
    GGP_CUDA(cudaEventRecord(ev_chunk[chunk & 15], produced_on));
    GGP_CUDA(cudaStreamWaitEvent(cs, ev_chunk[chunk & 15], 0));
    const long long blk = n[0] * n2loc * n3loc;
    const size_t esz = sizeof(cpx<T>);
    for (int i = 0; i < P; ++i) {
      const int q = (prank + i) % P;
      for (int c = 0; c < M; ++c)
        GGP_CUDA(cudaMemcpy2DAsync(peer_u[q][c] + n[0] * ((long long)prank * n2loc + y0), esz * (size_t)(n[0] * n2g),
                                   stage[c] + (long long)q * blk + n[0] * y0, esz * (size_t)(n[0] * n2loc),
                                   esz * (size_t)(n[0] * (y1 - y0)), (size_t)n3loc, cudaMemcpyDefault, cs));
    }
    return 0;
  }
  // z pass of the y-slab (forward FFT_z x exp_D x inverse FFT_z) with its results sent back to the z-slabs
  int slab_z() {
    int rc;
    if (!slab_ce) return run_str(2, 1, true, 2);
    int C = (profiling || flush_buf) ? 1 : slab_chunks;
    if (C > n2loc) C = (int)n2loc;
    if (C < 1) C = 1;
    for (int c = 0; c < C; ++c) {
      const long long y0 = (long long)c * n2loc / C, y1 = (long long)(c + 1) * n2loc / C;
      if ((rc = run_str(2, 1, true, 2, y0, y1 - y0, stream, true))) return rc;
      if ((rc = ce_push_z(c, stream, y0, y1))) return rc;
    }
    return ce_join();
  }

  int slab_iry(bool do_inv, bool pre, bool post, const HalfStep<T>& hA, const HalfStep<T>& hB, bool do_y) {
    int rc;
    int C = (profiling || flush_buf) ? 1 : slab_chunks;
    if (C > n3loc) C = (int)n3loc;
    if (C < 1) C = 1;
    if (slab_ce && (rc = ce_setup())) return rc;
    if (C > 1) {
      if (!aux_stream) {
        GGP_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
        GGP_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        GGP_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      }
      GGP_CUDA(cudaEventRecord(ev_fork, stream));
      GGP_CUDA(cudaStreamWaitEvent(aux_stream, ev_fork, 0));
    }
    for (int c = 0; c < C; ++c) {
      cudaStream_t st = (c & 1) ? aux_stream : stream;
      const long long z0 = (long long)c * n3loc / C, z1 = (long long)(c + 1) * n3loc / C;
      if (do_inv && (rc = run_str(1, 2, false, 0, z0, z1 - z0, st))) return rc;
      if ((rc = run_row(pre, post, hA, hB, z0, z1 - z0, st))) return rc;
      if (do_y && (rc = run_str(1, 0, false, 1, z0, z1 - z0, st, slab_ce))) return rc;
      if (do_y && slab_ce && (rc = ce_push_y(c, st, z0, z1))) return rc;
    }
    if (C > 1) {
      GGP_CUDA(cudaEventRecord(ev_join, aux_stream));
      GGP_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    }
    if (do_y && slab_ce) return ce_join();
    return 0;
  }

  // All-to-all transposes of the slab decomposition (ncclSend/ncclRecv group on the plan's stream).
  // forward: z-slabs u (n1, n2g, n3loc) -> y-slabs xbuf (n1, n2loc, n3g);  backward: the inverse.
  int transpose(bool forward) {
#ifdef GGP_WITH_NCCL
    if (!comm) return fail(GGP_ERR_NCCL, "slab plan without communicator: call ggp_comm_init or ggp_slab_ipc_attach first");
    const size_t esz = sizeof(cpx<T>);
    for (int c = 0; c < M; ++c)
      if (!sendbuf[c]) {
        int rc2 = dalloc((void**)&sendbuf[c], esz * (size_t)nspatial);
        if (rc2) return rc2;
      }
    const long long blk = n[0] * n2loc * n3loc;                  // elements exchanged with each peer
    const size_t width = (size_t)(n[0] * n2loc) * esz;           // one (x, y-range) chunk
    const size_t zpitch = (size_t)(n[0] * n2g) * esz;            // one z-plane of the z-slab
    const ncclDataType_t ty = sizeof(T) == 4 ? ncclFloat : ncclDouble;
    for (int c = 0; c < M; ++c) {
      if (forward) {
        // my own block never touches NCCL: one strided copy u -> xbuf
        GGP_CUDA(cudaMemcpy2DAsync(xbuf[c] + (size_t)prank * blk, width, u[c] + (size_t)prank * n[0] * n2loc, zpitch, width,
                                   (size_t)n3loc, cudaMemcpyDeviceToDevice, stream));
        for (int q = 0; q < P; ++q)  // pack: peer q gets my z-planes restricted to its y-range
          if (q != prank)
            GGP_CUDA(cudaMemcpy2DAsync(sendbuf[c] + (size_t)q * blk, width, u[c] + (size_t)q * n[0] * n2loc, zpitch, width,
                                       (size_t)n3loc, cudaMemcpyDeviceToDevice, stream));
        ncclGroupStart();
        for (int q = 0; q < P; ++q) {
          if (q == prank) continue;
          ncclSend(sendbuf[c] + (size_t)q * blk, (size_t)blk * 2, ty, q, comm, stream);
          ncclRecv(xbuf[c] + (size_t)q * blk, (size_t)blk * 2, ty, q, comm, stream);   // lands in place: z slowest
        }
        ncclResult_t r = ncclGroupEnd();
        if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("all-to-all: ") + ncclGetErrorString(r));
      } else {
        GGP_CUDA(cudaMemcpy2DAsync(u[c] + (size_t)prank * n[0] * n2loc, zpitch, xbuf[c] + (size_t)prank * blk, width, width,
                                   (size_t)n3loc, cudaMemcpyDeviceToDevice, stream));
        ncclGroupStart();
        for (int q = 0; q < P; ++q) {
          if (q == prank) continue;
          ncclSend(xbuf[c] + (size_t)q * blk, (size_t)blk * 2, ty, q, comm, stream);    // contiguous: peer q's z-range
          ncclRecv(sendbuf[c] + (size_t)q * blk, (size_t)blk * 2, ty, q, comm, stream);
        }
        ncclResult_t r = ncclGroupEnd();
        if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("all-to-all: ") + ncclGetErrorString(r));
        for (int q = 0; q < P; ++q)  // unpack: peer q's y-range of my z-planes
          if (q != prank)
            GGP_CUDA(cudaMemcpy2DAsync(u[c] + (size_t)q * n[0] * n2loc, zpitch, sendbuf[c] + (size_t)q * blk, width, width,
                                       (size_t)n3loc, cudaMemcpyDeviceToDevice, stream));
      }
    }
    launches += 0;
    return 0;
#else
    (void)forward;
    return fail(GGP_ERR_NCCL, "libggp was built without NCCL");
#endif
  }

  int slab_barrier() {
    SlabBarrierParams b;
    memset(&b, 0, sizeof(b));
    for (int q = 0; q < P; ++q) b.peer_flags[q] = peer_flags[q];
    b.my_flags = flags;
    b.err = bar_err;
    b.P = P;
    b.me = prank;
    b.epoch = ++epoch;
    slab_barrier_kernel<<<1, 32, 0, stream>>>(b);
    GGP_CUDA(cudaGetLastError());
    ++launches;
    return 0;
  }

  // CUDA IPC handles of this rank's u, xbuf and flags (see ggp_slab_ipc_export)
  int ipc_export(void* blob) override {
    if (!slab) return fail(GGP_ERR_INVALID, "ggp_slab_ipc_export: not a slab plan");
    unsigned char* o = (unsigned char*)blob;
    memset(o, 0, GGP_IPC_BLOB_BYTES);
    int32_t hdr[4] = {0x47475031, M, P, prank};
    memcpy(o, hdr, sizeof(hdr));
    cudaIpcMemHandle_t h;
    size_t off = 64;
    void* ptrs[5] = {u[0], M > 1 ? u[1] : nullptr, xbuf[0], M > 1 ? xbuf[1] : nullptr, flags};
    for (int i = 0; i < 5; ++i, off += sizeof(h)) {
      if (!ptrs[i]) continue;
      GGP_CUDA(cudaIpcGetMemHandle(&h, ptrs[i]));
      memcpy(o + off, &h, sizeof(h));
    }
    return 0;
  }
  int ipc_attach(const void* blobs) override {
    if (!slab) return fail(GGP_ERR_INVALID, "ggp_slab_ipc_attach: not a slab plan");
    if (P > GGP_MAX_PEERS) return fail(GGP_ERR_UNSUPPORTED, "peer exchange supports at most 8 ranks");
    const unsigned char* b = (const unsigned char*)blobs;
    for (int q = 0; q < P; ++q) {
      const unsigned char* o = b + (size_t)q * GGP_IPC_BLOB_BYTES;
      int32_t hdr[4];
      memcpy(hdr, o, sizeof(hdr));
      if (hdr[0] != 0x47475031 || hdr[1] != M || hdr[2] != P || hdr[3] != q)
        return fail(GGP_ERR_INVALID, "ggp_slab_ipc_attach: blob " + std::to_string(q) + " does not belong to rank " + std::to_string(q));
      void* ptrs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
      if (q == prank) {
        ptrs[0] = u[0]; ptrs[1] = u[1]; ptrs[2] = xbuf[0]; ptrs[3] = xbuf[1]; ptrs[4] = flags;
      } else {
        size_t off = 64;
        for (int i = 0; i < 5; ++i, off += sizeof(cudaIpcMemHandle_t)) {
          if ((i == 1 || i == 3) && M < 2) continue;
          cudaIpcMemHandle_t h;
          memcpy(&h, o + off, sizeof(h));
          cudaError_t e = cudaIpcOpenMemHandle(&ptrs[i], h, cudaIpcMemLazyEnablePeerAccess);
          if (e != cudaSuccess)
            return fail(GGP_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(q) + "): " + cudaGetErrorString(e));
          ipc_opened.push_back(ptrs[i]);
        }
      }
      peer_u[q][0] = (cpx<T>*)ptrs[0]; peer_u[q][1] = (cpx<T>*)ptrs[1];
      peer_x[q][0] = (cpx<T>*)ptrs[2]; peer_x[q][1] = (cpx<T>*)ptrs[3];
      peer_flags[q] = (unsigned*)ptrs[4];
    }
    p2p = getenv("GGP_SLAB_NCCL") == nullptr;
    if (const char* e = getenv("GGP_SLAB_CHUNKS")) slab_chunks = atoi(e) > 0 ? atoi(e) : 1;
    slab_yocc1 = getenv("GGP_SLAB_YOCC1") != nullptr;
    slab_ce = getenv("GGP_SLAB_CE") != nullptr && atoi(getenv("GGP_SLAB_CE")) != 0;
    return 0;
  }
  int barrier_status() {
    int e = 0;
    if (bar_err) GGP_CUDA(cudaMemcpy(&e, bar_err, sizeof(int), cudaMemcpyDeviceToHost));
    return e ? fail(GGP_ERR_NCCL, "slab barrier timed out: a peer rank did not arrive") : 0;
  }

  // which compile-time variant of the half-step covers this problem (pointwise.cuh)
  int pw_variant() const {
    if (pw.noise) {
      if (pw.noise_field || pw.pump_dense) return PW_FIELD;
      // the Truncated-Wigner shape: everything the general stochastic variant decides per point is fixed (pointwise.cuh)
      static const bool no_tw = getenv("GGP_NO_TW") != nullptr;
      if (!no_tw && pw.noise == NOISE_PHILOX && pw.vkind == KIND_NONE && pw.nl != 2 && (!pw.pump || pw.pump_const)) return PW_TW;
      return PW_STOCH;
    }
    if (pw.pump_dense) return PW_DENSE;
    if (pw.vkind || pw.pump || pw.nl != 1) return PW_DET;
    return PW_KERR;
  }

  int ensure_hs(size_t count) {
    if (count <= hs_cap) return 0;
    int rc = dalloc((void**)&hs_dev, sizeof(HalfStep<T>) * count);
    if (rc) return rc;
    hs_cap = count;
    return 0;
  }

  int step(int64_t nsteps, const double* amp, const void* const* noise, const void* const* profiles) override {
    if (nsteps <= 0) return 0;
    if (noise && noise_kind == GGP_NOISE_NONE) return fail(GGP_ERR_INVALID, "noise_host given but the plan has no noise term");
    const bool dense = pump_kind == GGP_PUMP_DENSE;
    if (dense && !profiles) return fail(GGP_ERR_INVALID, "GGP_PUMP_DENSE plans are stepped with ggp_step_dense (pump profiles per half-step)");
    if (!dense && profiles) return fail(GGP_ERR_INVALID, "ggp_step_dense on a plan without a dense pump");
    pw.noise = noise_kind == GGP_NOISE_NONE ? NOISE_OFF : (noise ? NOISE_HOST : NOISE_PHILOX);
    auto amps = [&](int64_t s, int half) -> std::complex<double> {
      if (pump_kind == GGP_PUMP_NONE) return 0.0;
      if (!amp) return amp_prev;  // static pump
      return std::complex<double>(amp[2 * (2 * s + half)], amp[2 * (2 * s + half) + 1]);
    };
    int rc;
    // the next half-step in the reference's order: pump pair (F_now, F_next) = (previous, this one's)  (src/misc.jl:39-42)
    auto next_half = [&](int64_t s, int half, int slot, HalfStep<T>* out) -> int {
      if (dense) {
        const int ring = (int)((pd_k + 1) % 3);
        int e = upload_profile(profiles[2 * s + half], ring);
        if (e) return e;
        ++pd_k;
        *out = make_half(1.0, 1.0, slot);
        for (int c = 0; c < pump_ncomp; ++c) {
          out->pd_now[c] = pdense[(pd_k + 2) % 3][c];   // F_{k-1}
          out->pd_next[c] = pdense[pd_k % 3][c];        // F_k
        }
        return 0;
      }
      const std::complex<double> a = amps(s, half);
      *out = make_half(amp_prev, a, slot);
      amp_prev = a;
      return 0;
    };
    if (ndim == 1) {
      const int64_t chunk_max = (noise || dense) ? 1 : 4096;
      std::vector<HalfStep<T>> hs;
      for (int64_t s0 = 0; s0 < nsteps; s0 += chunk_max) {
        const int64_t cn = std::min<int64_t>(chunk_max, nsteps - s0);
        hs.resize((size_t)(2 * cn));
        for (int64_t s = 0; s < cn; ++s) {
          if (noise) {
            if ((rc = upload_noise(noise, s0 + s, 0, 0))) return rc;
            if ((rc = upload_noise(noise, s0 + s, 1, 1))) return rc;
          }
          if ((rc = next_half(s0 + s, 0, 0, &hs[(size_t)(2 * s)]))) return rc;
          if ((rc = next_half(s0 + s, 1, 1, &hs[(size_t)(2 * s + 1)]))) return rc;
        }
        if ((rc = ensure_hs((size_t)(2 * chunk_max)))) return rc;
        GGP_CUDA(cudaMemcpyAsync(hs_dev, hs.data(), sizeof(HalfStep<T>) * hs.size(), cudaMemcpyHostToDevice, stream));
        OneDParams<T> p;
        memset(&p, 0, sizeof(p));
        p.u[0] = u[0];
        p.u[1] = u[1];
        p.tw = tw[0];
        p.nlines = nbatch;
        p.pw = pw;
        p.hs = hs_dev;
        p.nsteps = (int)cn;
        for (int i = 0; i < 4; ++i) p.D[i] = D[i];
        p.dkind = dkind;
        if ((rc = window_begin(cn))) return rc;
        if ((rc = prof_begin(KC_ONED))) return rc;
        GGP_LAUNCH(dispatch_oned<T>((int)n[0], M, pw_variant(), p, stream), "oned_kernel");
        ++launches;
        if ((rc = prof_end())) return rc;
        if ((rc = window_end())) return rc;
      }
      return 0;
    }
    // 2-D / 3-D
    HalfStep<T> none;
    memset(&none, 0, sizeof(none));
    HalfStep<T> prev2 = none;
    for (int64_t s = 0; s < nsteps; ++s) {
      if (noise && (rc = upload_noise(noise, s, 0, 0))) return rc;
      HalfStep<T> h1;
      if ((rc = next_half(s, 0, 0, &h1))) return rc;
      if (dkind == GGP_TABLE_NONE) {
        // no dispersion: both half-steps of step s in one point-wise launch.  (Philox: the two then draw from the
        // (z,w) and (x,y) words of ONE call, whereas with a dispersion table a call is shared by the trailing
        // half-step of step s and the leading one of step s+1 -- distinct draws either way, different streams.)
        if (noise && (rc = upload_noise(noise, s, 1, 1))) return rc;
        HalfStep<T> h2;
        if ((rc = next_half(s, 1, 1, &h2))) return rc;
        if (has_pointwise) {
          if ((rc = window_begin())) return rc;
          if ((rc = run_row(false, false, h1, h2))) return rc;
          if ((rc = window_end())) return rc;
        }
        continue;
      }
      if (slab && p2p) {
        // [inverse y pass of step s-1] -> contiguous-axis kernel -> forward y pass + scatter, chunked over z-planes
        if ((rc = slab_iry(s > 0, s > 0, true, prev2, h1, true))) return rc;
        if (s > 0 && (rc = window_end())) return rc;
        if (noise && (rc = upload_noise(noise, s, 1, 1))) return rc;
        if ((rc = next_half(s, 1, 1, &prev2))) return rc;
        if ((rc = window_begin())) return rc;
        if ((rc = slab_barrier())) return rc;
        if ((rc = slab_z())) return rc;
        if ((rc = slab_barrier())) return rc;
        continue;
      }
      if ((rc = run_row(s > 0, true, prev2, h1))) return rc;
      if (s > 0 && (rc = window_end())) return rc;     // closes the window of step s-1
      if (noise && (rc = upload_noise(noise, s, 1, 1))) return rc;
      if ((rc = next_half(s, 1, 1, &prev2))) return rc;
      if ((rc = window_begin())) return rc;            // step s: strided pass(es) + the contiguous-axis kernel that closes it
      if (ndim == 2) {
        if ((rc = run_str(1, 1))) return rc;
      } else if (!slab) {
        if ((rc = run_str(1, 0))) return rc;
        if ((rc = run_str(2, 1))) return rc;
        if ((rc = run_str(1, 2))) return rc;
      } else {
        if ((rc = run_str(1, 0))) return rc;
        if ((rc = transpose(true))) return rc;
        if ((rc = run_str(2, 1, true))) return rc;
        if ((rc = transpose(false))) return rc;
        if ((rc = run_str(1, 2))) return rc;
      }
    }
    if (dkind != GGP_TABLE_NONE) {
      if (slab && p2p) {
        if ((rc = slab_iry(true, true, false, prev2, none, false))) return rc;
      } else if ((rc = run_row(true, false, prev2, none))) return rc;
      if ((rc = window_end())) return rc;
    }
    return 0;
  }

  // g1[i][j] = sum_traj conj(F2[j]) F1[i],  F_a = ifftshift(fft(fftshift(u .* w_a)))  (test/windowed_ft.jl:31-49), per
  // component; the caller divides by length(sol).  1-D ensembles, even N.
  int observe_windowed(const double* w1, const double* w2, double* out) override {
    if (ndim != 1 || slab) return fail(GGP_ERR_UNSUPPORTED, "windowed correlations are defined for 1-D ensembles");
    if (!size_supported(n[0]) || n[0] < 2) return fail(GGP_ERR_UNSUPPORTED, "windowed correlations need a power-of-two axis");
    const int N = (int)n[0];
    const long long total = nspatial * nbatch;
    const size_t bytes = sizeof(cpx<T>) * (size_t)total;
    int rc;
    const size_t cnt = (size_t)M * N * N * 2;
    if (!g1_dev && (rc = dalloc((void**)&g1_dev, sizeof(double) * cnt))) return rc;
    if (!win_dev && (rc = dalloc((void**)&win_dev, sizeof(cpx<T>) * 2 * (size_t)N))) return rc;
    if (!xbuf1 && (rc = dalloc((void**)&xbuf1, bytes))) return rc;
    std::vector<cpx<T>> hw(2 * (size_t)N);
    for (int k = 0; k < N; ++k) {
      hw[(size_t)k] = mk<T>((T)w1[2 * k], (T)w1[2 * k + 1]);
      hw[(size_t)N + k] = mk<T>((T)w2[2 * k], (T)w2[2 * k + 1]);
    }
    GGP_CUDA(cudaMemcpyAsync(win_dev, hw.data(), sizeof(cpx<T>) * hw.size(), cudaMemcpyHostToDevice, stream));
    GGP_CUDA(cudaMemsetAsync(g1_dev, 0, sizeof(double) * cnt, stream));
    HalfStep<T> none;
    memset(&none, 0, sizeof(none));
    for (int c = 0; c < M; ++c) {
      if (!scratch[c] && (rc = dalloc((void**)&scratch[c], bytes))) return rc;
      GGP_CUDA(cudaMemcpyAsync(scratch[c], u[c], bytes, cudaMemcpyDeviceToDevice, stream));
    }
    const unsigned wb = (unsigned)((total + 255) / 256);
    const unsigned tiles = (unsigned)((N + 15) / 16);
    unsigned zs = (unsigned)std::min<long long>(64, (nbatch + 255) / 256);
    if (zs < 1) zs = 1;
    // both windowed transforms of every component with the step's own FFT kernel (all components at once)
    for (int a = 0; a < 2; ++a) {
      for (int c = 0; c < M; ++c)
        window_kernel<T><<<wb, 256, 0, stream>>>(u[c], scratch[c], win_dev + (size_t)a * N, N, total);
      launches += M;
      if ((rc = run_row(false, true, none, none))) return rc;
      if (a == 0) {
        // keep X1 of component 0; a second component's X1 is recomputed below (M = 2 is rare here)
        GGP_CUDA(cudaMemcpyAsync(xbuf1, u[0], bytes, cudaMemcpyDeviceToDevice, stream));
      }
    }
    g1_kernel<T><<<dim3(tiles, tiles, zs), dim3(16, 16), 0, stream>>>(xbuf1, u[0], g1_dev, N, nbatch);
    ++launches;
    if (M == 2) {
      GGP_CUDA(cudaMemcpyAsync(xbuf1, u[1], bytes, cudaMemcpyDeviceToDevice, stream));  // X2 of component 1
      window_kernel<T><<<wb, 256, 0, stream>>>(u[1], scratch[1], win_dev, N, total);
      window_kernel<T><<<wb, 256, 0, stream>>>(u[0], scratch[0], win_dev, N, total);
      launches += 2;
      if ((rc = run_row(false, true, none, none))) return rc;                            // X1 of component 1 in u[1]
      g1_kernel<T><<<dim3(tiles, tiles, zs), dim3(16, 16), 0, stream>>>(u[1], xbuf1, g1_dev + (size_t)N * N * 2, N, nbatch);
      ++launches;
    }
    for (int c = 0; c < M; ++c) GGP_CUDA(cudaMemcpyAsync(u[c], scratch[c], bytes, cudaMemcpyDeviceToDevice, stream));
    GGP_CUDA(cudaGetLastError());
#ifdef GGP_WITH_NCCL
    if (comm && nranks > 1) {
      ncclResult_t r = ncclAllReduce(g1_dev, g1_dev, cnt, ncclDouble, ncclSum, comm, stream);
      if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    }
#endif
    GGP_CUDA(cudaMemcpyAsync(out, g1_dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }

  int observe(int kind, double* out) override {
    const int threads = 256;
    const unsigned blocks = (unsigned)((nspatial + threads - 1) / threads);
    size_t count = 0;
    double* result = obs_dev;
    if (kind == GGP_OBS_DENSITY) {
      for (int c = 0; c < M; ++c) {
        density_kernel<T><<<blocks, threads, 0, stream>>>(u[c], obs_dev + c * nspatial, nspatial, nbatch, 1.0);
        ++launches;
      }
      count = (size_t)(nspatial * M);
    } else if (kind == GGP_OBS_NORM) {
      GGP_CUDA(cudaMemsetAsync(obs_dev + nspatial * M, 0, sizeof(double) * 8, stream));
      for (int c = 0; c < M; ++c) {
        density_kernel<T><<<blocks, threads, 0, stream>>>(u[c], obs_dev, nspatial, nbatch, 1.0);
        sum_kernel<<<148, 256, 0, stream>>>(obs_dev, obs_dev + nspatial * M + c, nspatial);
        launches += 2;
      }
      GGP_CUDA(cudaMemcpyAsync(obs_dev, obs_dev + nspatial * M, sizeof(double) * M, cudaMemcpyDeviceToDevice, stream));
      count = (size_t)M;
    } else if (kind == GGP_OBS_MOMENTUM) {
      // n(k) = sum_traj |fft(u)(k)|^2 / N^2 (examples/truncated_wigner.jl:110-113): transform the state in
      // place with the step's own FFT kernels, accumulate, then restore the saved copy.
      if (slab) return fail(GGP_ERR_UNSUPPORTED, "momentum observable is not wired for slab-decomposed plans yet");
      for (int i = 0; i < ndim; ++i)
        if (!size_supported(n[i])) return fail(GGP_ERR_UNSUPPORTED, "momentum observable needs power-of-two axes");
      const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
      int rc;
      for (int c = 0; c < M; ++c) {
        if (!scratch[c] && (rc = dalloc((void**)&scratch[c], bytes))) return rc;
        GGP_CUDA(cudaMemcpyAsync(scratch[c], u[c], bytes, cudaMemcpyDeviceToDevice, stream));
      }
      HalfStep<T> none;
      memset(&none, 0, sizeof(none));
      if ((rc = run_row(false, true, none, none))) return rc;
      if (ndim >= 2 && (rc = run_str(1, 0))) return rc;
      if (ndim == 3 && (rc = run_str(2, 0))) return rc;
      const double sc = 1.0 / ((double)nspatial * (double)nspatial);
      for (int c = 0; c < M; ++c) {
        density_kernel<T><<<blocks, threads, 0, stream>>>(u[c], obs_dev + c * nspatial, nspatial, nbatch, sc);
        ++launches;
        GGP_CUDA(cudaMemcpyAsync(u[c], scratch[c], bytes, cudaMemcpyDeviceToDevice, stream));
      }
      count = (size_t)(nspatial * M);
    } else if (kind == GGP_OBS_G2_MOMENTUM) {
      if (ndim != 1 || slab) return fail(GGP_ERR_UNSUPPORTED, "the G2 observable is defined for 1-D ensembles (N x N output)");
      if (!size_supported(n[0])) return fail(GGP_ERR_UNSUPPORTED, "momentum observables need power-of-two axes");
      const int N = (int)n[0];
      const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
      int rc;
      if (!g2_dev && (rc = dalloc((void**)&g2_dev, sizeof(double) * (size_t)M * N * N))) return rc;
      GGP_CUDA(cudaMemsetAsync(g2_dev, 0, sizeof(double) * (size_t)M * N * N, stream));
      for (int c = 0; c < M; ++c) {
        if (!scratch[c] && (rc = dalloc((void**)&scratch[c], bytes))) return rc;
        GGP_CUDA(cudaMemcpyAsync(scratch[c], u[c], bytes, cudaMemcpyDeviceToDevice, stream));
      }
      HalfStep<T> none;
      memset(&none, 0, sizeof(none));
      if ((rc = run_row(false, true, none, none))) return rc;
      const double sc = 1.0 / ((double)N * N * (double)N * N);   // (fft / N) as in examples/truncated_wigner.jl:110
      const unsigned tiles = (unsigned)((N + 15) / 16);
      unsigned zs = (unsigned)std::min<long long>(64, (nbatch + 255) / 256);
      if (zs < 1) zs = 1;
      for (int c = 0; c < M; ++c) {
        g2_kernel<T><<<dim3(tiles, tiles, zs), dim3(16, 16), 0, stream>>>(u[c], g2_dev + (size_t)c * N * N, N, nbatch, sc);
        ++launches;
        GGP_CUDA(cudaMemcpyAsync(u[c], scratch[c], bytes, cudaMemcpyDeviceToDevice, stream));
      }
      count = (size_t)M * N * N;
      result = g2_dev;
    } else {
      return fail(GGP_ERR_INVALID, "unknown observable kind");
    }
    GGP_CUDA(cudaGetLastError());
#ifdef GGP_WITH_NCCL
    if (comm && nranks > 1) {
      ncclResult_t r = ncclAllReduce(result, result, count, ncclDouble, ncclSum, comm, stream);
      if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    }
#endif
    GGP_CUDA(cudaMemcpyAsync(out, result, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
};

}  // namespace ggp

struct ggp_plan {
  ggp::PlanBase* impl;
};

using namespace ggp;

extern "C" {

int ggp_version(void) { return 200; }

int ggp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* ggp_last_error(void) { return g_err.c_str(); }

int ggp_plan_create(const ggp_desc* d, ggp_plan** out) {
  if (!d || !out) return fail(GGP_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->abi_version != GGP_ABI_VERSION) return fail(GGP_ERR_INVALID, "abi_version mismatch");
  if (d->slab_nranks < 0 || d->slab_nranks == 1) { /* 0 or 1: no slab decomposition */ }
  if (d->struct_size != sizeof(ggp_desc)) return fail(GGP_ERR_INVALID, "struct_size mismatch");
  if (d->ndim < 1 || d->ndim > 3) return fail(GGP_ERR_INVALID, "ndim must be 1..3");
  if (d->ncomp < 1 || d->ncomp > GGP_MAX_COMPONENTS)
    return fail(GGP_ERR_UNSUPPORTED, "ncomp must be 1.." + std::to_string(GGP_MAX_COMPONENTS));
  for (int i = 0; i < d->ndim; ++i)
    if (!gen_axis_supported(d->n[i]))
      return fail(GGP_ERR_UNSUPPORTED, "axis length " + std::to_string(d->n[i]) +
                                           " is beyond the built transform sizes (powers of two up to 8192, any n up to 4096)");
  if (d->nbatch < 1) return fail(GGP_ERR_INVALID, "nbatch must be >= 1");
  for (int i = 0; i < d->ndim; ++i)
    if (d->n[i] < 1) return fail(GGP_ERR_INVALID, "n[i] must be >= 1");
  if (d->precision != GGP_C64 && d->precision != GGP_C128) return fail(GGP_ERR_INVALID, "bad precision");
  if (!(d->dt == d->dt)) return fail(GGP_ERR_INVALID, "dt is NaN");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(GGP_ERR_CUDA, "no CUDA device available (this backend has no CPU fallback)");
  }
  int dev = d->device;
  if (dev < 0) {
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(GGP_ERR_CUDA, "cudaGetDevice failed");
  }
  if (dev >= ndev) return fail(GGP_ERR_INVALID, "device ordinal out of range");
  GGP_CUDA(cudaSetDevice(dev));
  PlanBase* impl;
  if (gen_wanted(*d))   // any axis length, ncomp > 2, matrix-valued nonlinearity (generic_plan.cuh)
    impl = d->precision == GGP_C64 ? (PlanBase*)new GenPlanT<float>() : (PlanBase*)new GenPlanT<double>();
  else
    impl = d->precision == GGP_C64 ? (PlanBase*)new PlanT<float>() : (PlanBase*)new PlanT<double>();
  impl->device = dev;
  int rc = impl->create(*d);
  if (rc) {
    delete impl;
    return rc;
  }
  *out = new ggp_plan{impl};
  return 0;
}

int ggp_plan_destroy(ggp_plan* p) {
  if (!p) return 0;
  delete p->impl;
  delete p;
  return 0;
}

#define GGP_ENTER(p)                                        \
  if (!(p) || !(p)->impl) return fail(GGP_ERR_INVALID, "null plan"); \
  GGP_CUDA(cudaSetDevice((p)->impl->device));

int ggp_set_state(ggp_plan* p, const void* const* u) {
  GGP_ENTER(p);
  if (!u) return fail(GGP_ERR_INVALID, "null state");
  return p->impl->set_state(u);
}
int ggp_get_state(ggp_plan* p, void* const* u) {
  GGP_ENTER(p);
  if (!u) return fail(GGP_ERR_INVALID, "null state");
  return p->impl->get_state(u);
}
int ggp_save_async(ggp_plan* p, void* const* u) {
  GGP_ENTER(p);
  if (!u) return fail(GGP_ERR_INVALID, "null state");
  return p->impl->save_async(u);
}
int ggp_save_wait(ggp_plan* p) {
  GGP_ENTER(p);
  return p->impl->save_wait();
}
int64_t ggp_checkpoint_bytes(ggp_plan* p) { return (p && p->impl) ? p->impl->checkpoint_bytes() : (int64_t)GGP_ERR_INVALID; }
int ggp_checkpoint_save(ggp_plan* p, void* blob, uint64_t capacity) {
  GGP_ENTER(p);
  if (!blob) return fail(GGP_ERR_INVALID, "null blob");
  return p->impl->checkpoint_save(blob, capacity);
}
int ggp_checkpoint_load(ggp_plan* p, const void* blob, uint64_t size) {
  GGP_ENTER(p);
  if (!blob) return fail(GGP_ERR_INVALID, "null blob");
  return p->impl->checkpoint_load(blob, size);
}
int ggp_step(ggp_plan* p, int64_t nsteps, const double* amp, const void* const* noise) {
  GGP_ENTER(p);
  return p->impl->step(nsteps, amp, noise);
}
int ggp_step_dense(ggp_plan* p, int64_t nsteps, const void* const* profiles, const void* const* noise) {
  GGP_ENTER(p);
  if (nsteps > 0 && !profiles) return fail(GGP_ERR_INVALID, "null pump profiles");
  return p->impl->step(nsteps, nullptr, noise, profiles);
}
int ggp_synchronize(ggp_plan* p) {
  GGP_ENTER(p);
  GGP_CUDA(cudaStreamSynchronize(p->impl->stream));
  return 0;
}
int ggp_observe(ggp_plan* p, int kind, double* out) {
  GGP_ENTER(p);
  if (!out) return fail(GGP_ERR_INVALID, "null output");
  return p->impl->observe(kind, out);
}

int ggp_observe_windowed(ggp_plan* p, const double* w1, const double* w2, double* out) {
  GGP_ENTER(p);
  if (!w1 || !w2 || !out) return fail(GGP_ERR_INVALID, "null argument");
  return p->impl->observe_windowed(w1, w2, out);
}

int ggp_comm_unique_id(void* id) {
#ifdef GGP_WITH_NCCL
  ncclUniqueId uid;
  ncclResult_t r = ncclGetUniqueId(&uid);
  if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
  static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id, &uid, 128);
  return 0;
#else
  (void)id;
  return fail(GGP_ERR_NCCL, "libggp was built without NCCL");
#endif
}
int ggp_comm_init(ggp_plan* p, int nranks, int rank, const void* id) {
  GGP_ENTER(p);
#ifdef GGP_WITH_NCCL
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclResult_t r = ncclCommInitRank(&p->impl->comm, nranks, uid, rank);
  if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
  p->impl->nranks = nranks;
  return 0;
#else
  (void)nranks; (void)rank; (void)id;
  return fail(GGP_ERR_NCCL, "libggp was built without NCCL");
#endif
}

int ggp_slab_ipc_export(ggp_plan* p, void* blob) {
  GGP_ENTER(p);
  if (!blob) return fail(GGP_ERR_INVALID, "blob is NULL");
  return p->impl->ipc_export(blob);
}
int ggp_slab_ipc_attach(ggp_plan* p, const void* blobs) {
  GGP_ENTER(p);
  if (!blobs) return fail(GGP_ERR_INVALID, "blobs is NULL");
  return p->impl->ipc_attach(blobs);
}

void* ggp_state_device_ptr(ggp_plan* p, int c) { return (p && p->impl) ? p->impl->state_ptr(c) : nullptr; }

int ggp_timer_begin(ggp_plan* p) {
  GGP_ENTER(p);
  GGP_CUDA(cudaEventRecord(p->impl->ev0, p->impl->stream));
  return 0;
}
int ggp_timer_end(ggp_plan* p, float* ms) {
  GGP_ENTER(p);
  GGP_CUDA(cudaEventRecord(p->impl->ev1, p->impl->stream));
  GGP_CUDA(cudaEventSynchronize(p->impl->ev1));
  GGP_CUDA(cudaEventElapsedTime(ms, p->impl->ev0, p->impl->ev1));
  return 0;
}
int64_t ggp_launch_count(ggp_plan* p) { return (p && p->impl) ? p->impl->launches : -1; }
int64_t ggp_device_bytes(ggp_plan* p) { return (p && p->impl) ? p->impl->dev_bytes : -1; }

void* ggp_host_alloc(uint64_t bytes) {
  void* q = nullptr;
  if (cudaMallocHost(&q, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    g_err = "cudaMallocHost failed";
    return nullptr;
  }
  return q;
}
int ggp_host_free(void* q) {
  if (q) GGP_CUDA(cudaFreeHost(q));
  return 0;
}

int ggp_host_register(void* q, uint64_t bytes) {
  if (!q) return fail(GGP_ERR_INVALID, "null pointer");
  cudaError_t e = cudaHostRegister(q, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GGP_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  }
  return 0;
}
int ggp_host_unregister(void* q) {
  if (!q) return 0;
  cudaError_t e = cudaHostUnregister(q);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GGP_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
  }
  return 0;
}

int ggp_profile_steps_enable(ggp_plan* p, int on, uint64_t flush_bytes) {
  GGP_ENTER(p);
  PlanBase* b = p->impl;
  GGP_CUDA(cudaStreamSynchronize(b->stream));
  for (cudaEvent_t e : b->sev) cudaEventDestroy(e);
  b->sev.clear();
  b->sev_steps.clear();
  b->window_open = false;
  if (b->sflush_buf) {
    cudaFree(b->sflush_buf);
    b->sflush_buf = nullptr;
    b->sflush_bytes = 0;
  }
  b->step_windows = on != 0;
  if (on) {
    b->steps_ms = 0;
    b->steps_n = 0;
    if (flush_bytes) {
      GGP_CUDA(cudaMalloc(&b->sflush_buf, flush_bytes));
      b->sflush_bytes = flush_bytes;
    }
  }
  return 0;
}
int ggp_profile_steps_read(ggp_plan* p, double* ms_total, int64_t* windows) {
  GGP_ENTER(p);
  int rc = p->impl->windows_collect();
  if (rc) return rc;
  if (ms_total) *ms_total = p->impl->steps_ms;
  if (windows) *windows = p->impl->steps_n;
  return 0;
}

int ggp_debug_l2_flush(ggp_plan* p, uint64_t bytes) {
  GGP_ENTER(p);
  GGP_CUDA(cudaStreamSynchronize(p->impl->stream));
  if (p->impl->flush_buf) {
    cudaFree(p->impl->flush_buf);
    p->impl->flush_buf = nullptr;
    p->impl->flush_bytes = 0;
  }
  if (bytes) {
    GGP_CUDA(cudaMalloc(&p->impl->flush_buf, bytes));
    p->impl->flush_bytes = bytes;
  }
  return 0;
}

int ggp_debug_flush_only(ggp_plan* p, int64_t count, float* ms) {
  GGP_ENTER(p);
  if (!p->impl->flush_buf) return fail(GGP_ERR_INVALID, "ggp_debug_l2_flush is not enabled");
  GGP_CUDA(cudaEventRecord(p->impl->ev0, p->impl->stream));
  for (int64_t i = 0; i < count; ++i)
    GGP_CUDA(cudaMemsetAsync(p->impl->flush_buf, 0, p->impl->flush_bytes, p->impl->stream));
  GGP_CUDA(cudaEventRecord(p->impl->ev1, p->impl->stream));
  GGP_CUDA(cudaEventSynchronize(p->impl->ev1));
  GGP_CUDA(cudaEventElapsedTime(ms, p->impl->ev0, p->impl->ev1));
  return 0;
}

int ggp_profile_enable(ggp_plan* p, int on) {
  GGP_ENTER(p);
  p->impl->profiling = on != 0;
  if (on) {
    for (int i = 0; i < KC_COUNT; ++i) {
      p->impl->prof_ms[i] = 0;
      p->impl->prof_n[i] = 0;
    }
  }
  return 0;
}
int ggp_profile_read(ggp_plan* p, double* ms_total, int64_t* launches) {
  GGP_ENTER(p);
  int rc = p->impl->prof_collect();
  if (rc) return rc;
  for (int i = 0; i < KC_COUNT; ++i) {
    ms_total[i] = p->impl->prof_ms[i];
    launches[i] = p->impl->prof_n[i];
  }
  return 0;
}

}  // extern "C"
