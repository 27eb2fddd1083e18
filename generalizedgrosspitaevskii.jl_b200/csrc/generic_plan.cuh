// The generic plan: everything the reference's Strang step accepts that the fused kernels do not cover --
//   * FFT axes of any length (src/misc.jl:53-58: plan_fft is FFTW, any n) -- Bluestein on the in-register line FFT,
//   * more than two field components (NTuple{M}, src/kernels.jl:31-35), here M <= GGP_MAX_COMPONENTS,
//   * matrix-valued nonlinearities (src/kernels.jl:22-25,44: cis of an SMatrix is the matrix exponential).
// It runs the reference's own sequence (src/strang_splitting.jl:69-90) with one kernel per stage:
//   potential_pump_step!  -> gen_pw_kernel      (muladd_kernel! in real space, src/kernels.jl:37-54)
//   perform_ft!           -> gen_fft_kernel per component and axis (generic.cuh)
//   k-space muladd        -> gen_disp_kernel
//   inverse perform_ft!   -> gen_fft_kernel
//   potential_pump_step!  -> gen_pw_kernel
// Same C ABI, same noise stream (Philox counters as in the fused path), same pump schedule.  GGP_FORCE_GENERIC=1 routes
// every (non-slab) plan through it: the tests use that to check the two paths against each other.
// Included by ggp_api.cu after PlanBase.
#pragma once
#include "generic.cuh"

namespace ggp {

constexpr int GMAXM = GGP_MAX_COMPONENTS;

template <typename T>
struct GenHalf {
  cpx<T> fnow, fnext;        // (dt/4) a_now, (dt/4) a_next   (src/kernels.jl:45-46 with the half step dt/2)
  const void* xi[GMAXM];     // host-fed noise (test mode)
  const cpx<T>* pd_now;      // dense pump: F_now / F_next on the grid, point-major then component
  const cpx<T>* pd_next;
  uint32_t ctr, ctr_hi;      // global half-step counter (Philox)
};

template <typename T>
struct GenPwParams {
  cpx<T>* u[GMAXM];
  long long nspatial, total;
  T dt, sqrt_dt;             // half step, sqrt(half step)
  const cpx<T>* expV;        // point-major [point][ncols]; FULL: column-major entries
  int vkind;
  const cpx<T>* S;           // separable pump profile, point-major [point][pump_ncomp]
  int pump;                  // 0 none, 1 one profile added to every component, 2 one per component
  int pump_dense;
  int nlkind;                // 0 none, 1 Number, 2 SVector, 3 SMatrix
  const cpx<T>* nlc;         // Number: 1; SVector: M; SMatrix: M*M [i][j]
  const cpx<T>* nlg;         // Number: M; SVector: M*M [i][j]; SMatrix: M^3 [i][j][k]
  int nl_real;               // every coefficient is real (pure phases)
  int noise, noise_real, noise_field;
  const cpx<T>* eta;         // M
  const cpx<T>* alpha;       // M*M
  const cpx<T>* nprof;
  int n1;
  uint32_t seed_lo, seed_hi;
  long long elem_offset;
  GenHalf<T> h;
};

// exp(i A) of a small complex matrix: scaling and squaring around a Taylor polynomial (|A| / 2^s <= 1/2:
// the remainder of degree K is below the plan's rounding).  Agrees with the closed form / Pade of StaticArrays to
// rounding; the reference's own choice of algorithm is third-party (StaticArrays `exp`, not under /root/reference).
template <typename T, int M>
__device__ __forceinline__ void cis_matrix(const cpx<T> (&A)[M][M], cpx<T> (&E)[M][M]) {
  constexpr int K = sizeof(T) == 4 ? 9 : 16;
  T nrm = (T)0;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    T col = (T)0;
#pragma unroll
    for (int i = 0; i < M; ++i) col += fabs(A[i][j].x) + fabs(A[i][j].y);
    nrm = col > nrm ? col : nrm;
  }
  int s = 0;
  T sc = (T)1;
  while (nrm * sc > (T)0.5 && s < 60) {
    sc *= (T)0.5;
    ++s;
  }
  cpx<T> B[M][M];   // i A / 2^s
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) B[i][j] = mk<T>(-A[i][j].y * sc, A[i][j].x * sc);
  // Horner: R = I + B/1 (I + B/2 (I + ... (I + B/K)))
  cpx<T> R[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) R[i][j] = mk<T>(i == j ? (T)1 : (T)0, (T)0);
#pragma unroll 1
  for (int k = K; k >= 1; --k) {
    const T inv = (T)1 / (T)k;
    cpx<T> P[M][M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
        for (int q = 0; q < M; ++q) acc = acc + cmul(B[i][q], R[q][j]);
        P[i][j] = mk<T>(acc.x * inv + (i == j ? (T)1 : (T)0), acc.y * inv);
      }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) R[i][j] = P[i][j];
  }
#pragma unroll 1
  for (int q = 0; q < s; ++q) {
    cpx<T> P[M][M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
        for (int r = 0; r < M; ++r) acc = acc + cmul(R[i][r], R[r][j]);
        P[i][j] = acc;
      }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) R[i][j] = P[i][j];
  }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) E[i][j] = R[i][j];
}

// one real-space half-step at every (point, trajectory): the registered forms of muladd_kernel! (src/kernels.jl:37-54)
// with the reference's `_mul` algebra (src/kernels.jl:9-15) for every legal nonlinearity (x) potential combination
template <typename T, int M>
__global__ void __launch_bounds__(128) gen_pw_kernel(const GenPwParams<T> p) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.total) return;
  const long long sidx = g % p.nspatial;
  cpx<T> f[M];
#pragma unroll
  for (int c = 0; c < M; ++c) f[c] = p.u[c][g];
  T n2[M];
#pragma unroll
  for (int j = 0; j < M; ++j) n2[j] = cabs2(f[j]);

  // noise amplitude on the PRE-update field (src/kernels.jl:40-42)
  cpx<T> etav[M];
  if (p.noise) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      etav[i] = p.eta[i];
      if (p.noise_field) {
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const T a = sqrt(n2[j]);
          etav[i] = mk<T>(fma_(p.alpha[i * M + j].x, a, etav[i].x), fma_(p.alpha[i * M + j].y, a, etav[i].y));
        }
        if (p.nprof) etav[i] = cmul(p.nprof[sidx % p.n1], etav[i]);
      }
    }
  }
  // w = fields + (dt/2) F_now   (src/kernels.jl:46,48)
  cpx<T> w[M];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    w[j] = f[j];
    if (p.pump) {
      const int pc = p.pump == 1 ? 0 : j, np = p.pump == 1 ? 1 : M;
      const cpx<T> sv = p.pump_dense ? p.h.pd_now[sidx * np + pc] : p.S[sidx * np + pc];
      w[j] = w[j] + cmul(p.h.fnow, sv);
    }
  }
  const int ncv = p.vkind == KIND_SCALAR ? 1 : (p.vkind == KIND_DIAG ? M : (p.vkind == KIND_FULL ? M * M : 0));
  const cpx<T>* Vp = p.expV + sidx * ncv;
  cpx<T> res[M];
  if (p.nlkind == 3) {
    // SMatrix nonlinearity: E = exp(-i dt G) is a matrix; E (x) exp_V follows `_mul`: Number -> scaled matrix,
    // SVector -> the matrix-vector product E * exp_V (a VECTOR, applied elementwise afterwards -- what the reference
    // computes, src/kernels.jl:9,44,48), SMatrix -> matrix product
    cpx<T> A[M][M], E[M][M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        cpx<T> gij = p.nlc[i * M + j];
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const cpx<T> c = p.nlg[(i * M + j) * M + k];
          gij = mk<T>(fma_(c.x, n2[k], gij.x), fma_(c.y, n2[k], gij.y));
        }
        A[i][j] = mk<T>(-p.dt * gij.x, -p.dt * gij.y);
      }
    cis_matrix<T, M>(A, E);
    if (p.vkind == KIND_DIAG) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        cpx<T> e = mk<T>((T)0, (T)0);
#pragma unroll
        for (int j = 0; j < M; ++j) e = e + cmul(E[i][j], Vp[j]);
        res[i] = cmul(e, w[i]);
      }
    } else {
      cpx<T> x[M];
      if (p.vkind == KIND_FULL) {
#pragma unroll
        for (int i = 0; i < M; ++i) {
          cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
          for (int j = 0; j < M; ++j) acc = acc + cmul(Vp[j * M + i], w[j]);
          x[i] = acc;
        }
      } else {
#pragma unroll
        for (int i = 0; i < M; ++i) x[i] = p.vkind == KIND_SCALAR ? cmul(Vp[0], w[i]) : w[i];
      }
#pragma unroll
      for (int i = 0; i < M; ++i) {
        cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
        for (int j = 0; j < M; ++j) acc = acc + cmul(E[i][j], x[j]);
        res[i] = acc;
      }
    }
  } else {
    // Number / SVector nonlinearity: phases, kept as ph = cis(-dt G) - 1 (see sincosm1_t in pointwise.cuh)
    cpx<T> ph[M];
    if (p.nlkind) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const int src = p.nlkind == 1 ? 0 : i;
        cpx<T> gi = p.nlc[src];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const cpx<T> c = p.nlg[src * M + j];
          gi = mk<T>(fma_(c.x, n2[j], gi.x), fma_(c.y, n2[j], gi.y));
        }
        T s, cm1;
        sincosm1_t(-p.dt * gi.x, &s, &cm1);
        ph[i] = mk<T>(cm1, s);
        if (!p.nl_real) {
          const T em1 = expm1_t(p.dt * gi.y);
          ph[i] = mk<T>(fma_(em1, cm1, cm1 + em1), fma_(em1, s, s));
        }
      }
    }
    if (p.vkind == KIND_FULL) {
      // only a Number-valued nonlinearity (or none) can meet an SMatrix potential (src/kernels.jl:9)
#pragma unroll
      for (int i = 0; i < M; ++i) {
        cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
        for (int j = 0; j < M; ++j) acc = acc + cmul(Vp[j * M + i], w[j]);
        res[i] = p.nlkind ? rotate_m1(acc, ph[0].x, ph[0].y) : acc;
      }
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        cpx<T> e = w[i];
        if (p.vkind == KIND_SCALAR) e = cmul(Vp[0], e);
        if (p.vkind == KIND_DIAG) e = cmul(Vp[i], e);
        res[i] = p.nlkind ? rotate_m1(e, ph[i].x, ph[i].y) : e;
      }
    }
  }
  if (p.pump) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const int pc = p.pump == 1 ? 0 : i, np = p.pump == 1 ? 1 : M;
      const cpx<T> sv = p.pump_dense ? p.h.pd_next[sidx * np + pc] : p.S[sidx * np + pc];
      res[i] = res[i] + cmul(p.h.fnext, sv);
    }
  }
  if (p.noise) {
    const unsigned long long pair = ((((unsigned long long)p.h.ctr_hi << 32) | p.h.ctr) + 1ull) >> 1;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      cpx<T> xi;
      if (p.noise == NOISE_HOST) {
        xi = p.noise_real ? mk<T>(((const T*)p.h.xi[i])[g], (T)0) : ((const cpx<T>*)p.h.xi[i])[g];
      } else {
        const uint4 r = philox_for<T>(g + p.elem_offset, (uint32_t)pair, (uint32_t)(pair >> 32), i, p.seed_lo, p.seed_hi);
        xi = normal_from<T>(r, p.h.ctr, p.noise_real);
      }
      const cpx<T> ex = cmul(etav[i], xi);   // -i sqrt(dt) eta xi   (src/kernels.jl:42)
      res[i] = res[i] + mk<T>(p.sqrt_dt * ex.y, -p.sqrt_dt * ex.x);
    }
  }
#pragma unroll
  for (int c = 0; c < M; ++c) p.u[c][g] = res[c];
}

template <typename T>
struct GenDispParams {
  cpx<T>* u[GMAXM];
  const cpx<T>* D;    // point-major [k][ncols], 1/prod(n) folded in
  long long nspatial, total;
  int dkind;
};

// k-space muladd: u~ <- exp_D[k] (x) u~   (src/strang_splitting.jl:73-74)
template <typename T, int M>
__global__ void __launch_bounds__(128) gen_disp_kernel(const GenDispParams<T> p) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.total) return;
  const long long sidx = g % p.nspatial;
  cpx<T> f[M];
#pragma unroll
  for (int c = 0; c < M; ++c) f[c] = p.u[c][g];
  if (p.dkind == KIND_SCALAR) {
    const cpx<T> d = p.D[sidx];
#pragma unroll
    for (int c = 0; c < M; ++c) f[c] = cmul(d, f[c]);
  } else if (p.dkind == KIND_DIAG) {
#pragma unroll
    for (int c = 0; c < M; ++c) f[c] = cmul(p.D[sidx * M + c], f[c]);
  } else {
    cpx<T> r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
      for (int j = 0; j < M; ++j) acc = acc + cmul(p.D[sidx * M * M + j * M + i], f[j]);
      r[i] = acc;
    }
#pragma unroll
    for (int c = 0; c < M; ++c) f[c] = r[c];
  }
#pragma unroll
  for (int c = 0; c < M; ++c) p.u[c][g] = f[c];
}

template <typename T>
static int dispatch_gen_fft(int L, const GenFftParams<T>& p, cudaStream_t st) {
  switch (L) {
#define X(n) \
  case n:    \
    return launch_gen_fft<T, n>(p, st);
    GGP_SIZES(X)
#undef X
  }
  return (int)cudaErrorNotSupported;
}

// smallest built power of two >= v (0 if none)
static long long gen_pow2_at_least(long long v) {
  long long L = 2;
  while (L < v) L <<= 1;
  return size_supported(L) ? L : 0;
}
// can an FFT axis of n points run on the generic transform?
static bool gen_axis_supported(long long n) {
  if (n == 1) return true;
  if (size_supported(n)) return true;
  return n >= 2 && gen_pow2_at_least(2 * n - 1) != 0;
}

template <typename T>
struct GenPlanT : PlanBase {
  using Tw = typename TwT<T>::type;
  int ndim = 0, M = 0;
  long long n[3] = {1, 1, 1};
  long long nspatial = 0, nbatch = 0, batch_offset = 0;
  double dt = 0;
  cpx<T>* u[GMAXM] = {};
  cpx<T>* D = nullptr;
  int dkind = 0;
  cpx<T>* V = nullptr;
  cpx<T>* S = nullptr;
  cpx<T>* pd[2] = {nullptr, nullptr};   // dense pump: F_now / F_next, swapped per half-step (src/misc.jl:39-42)
  int pd_next = 0;
  std::vector<cpx<T>> pd_stage;
  int pump_kind = 0, pump_ncomp = 0, table_prec = GGP_C128;
  int noise_kind = 0, noise_real = 0;
  bool has_pointwise = false;
  std::complex<double> amp_prev = 0;
  uint64_t half_ctr = 0;
  GenPwParams<T> pw;
  // per axis: transform length L (== n: plain; > n: Bluestein), twiddles, chirp tables
  long long L[3] = {0, 0, 0};
  Tw* tw[3] = {nullptr, nullptr, nullptr};
  cpx<T>* chirp[3] = {nullptr, nullptr, nullptr};
  cpx<T>* bhat[3] = {nullptr, nullptr, nullptr};
  void* xi_dev[GMAXM] = {};
  double* obs_dev = nullptr;
  cpx<T>* scratch = nullptr;
  cudaEvent_t save_done = nullptr;
  bool save_pending = false;
  std::vector<void*> allocs;

  ~GenPlanT() override {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (void* q : allocs) cudaFree(q);
    if (save_done) cudaEventDestroy(save_done);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (cudaEvent_t e : pev) cudaEventDestroy(e);
    for (cudaEvent_t e : sev) cudaEventDestroy(e);
    if (flush_buf) cudaFree(flush_buf);
    if (sflush_buf) cudaFree(sflush_buf);
#ifdef GGP_WITH_NCCL
    if (comm) ncclCommDestroy(comm);
#endif
    if (own_stream && stream) cudaStreamDestroy(stream);
  }

  int dalloc(void** q, size_t bytes) {
    cudaError_t e = cudaMalloc(q, bytes ? bytes : 1);
    if (e != cudaSuccess)
      return fail(GGP_ERR_ALLOC, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    allocs.push_back(*q);
    dev_bytes += (int64_t)bytes;
    return 0;
  }
  static std::complex<double> host_entry(const void* host, int prec, size_t i) {
    if (prec == GGP_C128) return ((const std::complex<double>*)host)[i];
    const std::complex<float> f = ((const std::complex<float>*)host)[i];
    return std::complex<double>(f.real(), f.imag());
  }
  // host table (count entries of `prec`) -> device array of T in the same (point-major) order, scaled
  int upload_aos(const void* host, int prec, size_t count, double scale, cpx<T>** dst) {
    std::vector<cpx<T>> tmp(count);
    for (size_t i = 0; i < count; ++i) {
      const std::complex<double> z = host_entry(host, prec, i) * scale;
      tmp[i] = mk<T>((T)z.real(), (T)z.imag());
    }
    int rc = dalloc((void**)dst, sizeof(cpx<T>) * count);
    if (rc) return rc;
    GGP_CUDA(cudaMemcpy(*dst, tmp.data(), sizeof(cpx<T>) * count, cudaMemcpyHostToDevice));
    return 0;
  }
  int upload_coeffs(const std::vector<std::complex<double>>& h, const cpx<T>** dst) {
    std::vector<cpx<T>> tmp(h.size() ? h.size() : 1);
    for (size_t i = 0; i < h.size(); ++i) tmp[i] = mk<T>((T)h[i].real(), (T)h[i].imag());
    cpx<T>* q = nullptr;
    int rc = dalloc((void**)&q, sizeof(cpx<T>) * tmp.size());
    if (rc) return rc;
    GGP_CUDA(cudaMemcpy(q, tmp.data(), sizeof(cpx<T>) * tmp.size(), cudaMemcpyHostToDevice));
    *dst = q;
    return 0;
  }
  static int ncols_of(int kind, int M) {
    return kind == GGP_TABLE_SCALAR ? 1 : kind == GGP_TABLE_DIAG ? M : kind == GGP_TABLE_FULL ? M * M : 0;
  }
  static void build_twiddles(long long N, long long E, std::vector<Tw>& h) {
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
  }
  // in-place radix-2 transform in extended precision (host, plan creation only): the chirp filter of Bluestein
  static void host_fft(std::vector<std::complex<long double>>& a) {
    const size_t N = a.size();
    for (size_t i = 1, j = 0; i < N; ++i) {
      size_t bit = N >> 1;
      for (; j & bit; bit >>= 1) j ^= bit;
      j ^= bit;
      if (i < j) std::swap(a[i], a[j]);
    }
    const long double pi = 3.14159265358979323846264338327950288L;
    for (size_t len = 2; len <= N; len <<= 1) {
      std::vector<std::complex<long double>> w(len / 2);
      for (size_t k = 0; k < len / 2; ++k) {
        const long double ang = -2.0L * pi * (long double)k / (long double)len;
        w[k] = std::complex<long double>(cosl(ang), sinl(ang));
      }
      for (size_t i = 0; i < N; i += len)
        for (size_t k = 0; k < len / 2; ++k) {
          const std::complex<long double> x = a[i + k], y = a[i + k + len / 2] * w[k];
          a[i + k] = x + y;
          a[i + k + len / 2] = x - y;
        }
    }
  }
  int setup_axis(int a) {
    const long long na = n[a];
    if (na == 1) return 0;
    for (int b = 0; b < a; ++b)
      if (n[b] == na) {
        L[a] = L[b];
        tw[a] = tw[b];
        chirp[a] = chirp[b];
        bhat[a] = bhat[b];
        return 0;
      }
    const bool plain = size_supported(na);
    L[a] = plain ? na : gen_pow2_at_least(2 * na - 1);
    if (!L[a])
      return fail(GGP_ERR_UNSUPPORTED, "FFT axis length " + std::to_string(na) + " is beyond the built transform sizes");
    int rc;
    {
      std::vector<Tw> h;
      build_twiddles(L[a], default_E<T>((int)L[a]), h);
      if ((rc = dalloc((void**)&tw[a], sizeof(Tw) * h.size()))) return rc;
      GGP_CUDA(cudaMemcpy(tw[a], h.data(), sizeof(Tw) * h.size(), cudaMemcpyHostToDevice));
    }
    if (plain) return 0;
    // w_j = exp(-i pi j^2 / n), angle reduced exactly: j^2 mod 2n
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<std::complex<long double>> w((size_t)na);
    for (long long j = 0; j < na; ++j) {
      const long long r = (j * j) % (2 * na);
      const long double ang = -pi * (long double)r / (long double)na;
      w[(size_t)j] = std::complex<long double>(cosl(ang), sinl(ang));
    }
    std::vector<std::complex<long double>> b((size_t)L[a], std::complex<long double>(0, 0));
    b[0] = std::conj(w[0]);
    for (long long j = 1; j < na; ++j) b[(size_t)j] = b[(size_t)(L[a] - j)] = std::conj(w[(size_t)j]);
    host_fft(b);
    std::vector<cpx<T>> hc((size_t)na), hb((size_t)L[a]);
    for (long long j = 0; j < na; ++j) hc[(size_t)j] = mk<T>((T)w[(size_t)j].real(), (T)w[(size_t)j].imag());
    for (long long j = 0; j < L[a]; ++j) {
      const std::complex<long double> z = b[(size_t)j] / (long double)L[a];
      hb[(size_t)j] = mk<T>((T)z.real(), (T)z.imag());
    }
    if ((rc = dalloc((void**)&chirp[a], sizeof(cpx<T>) * hc.size()))) return rc;
    if ((rc = dalloc((void**)&bhat[a], sizeof(cpx<T>) * hb.size()))) return rc;
    GGP_CUDA(cudaMemcpy(chirp[a], hc.data(), sizeof(cpx<T>) * hc.size(), cudaMemcpyHostToDevice));
    GGP_CUDA(cudaMemcpy(bhat[a], hb.data(), sizeof(cpx<T>) * hb.size(), cudaMemcpyHostToDevice));
    return 0;
  }

  int create(const ggp_desc& d) override {
    ndim = d.ndim;
    M = d.ncomp;
    nspatial = 1;
    for (int i = 0; i < ndim; ++i) {
      n[i] = d.n[i];
      nspatial *= n[i];
    }
    nbatch = d.nbatch;
    batch_offset = d.batch_offset;
    dt = d.dt;
    dkind = d.disp_kind;
    if (d.slab_nranks > 1) return fail(GGP_ERR_UNSUPPORTED, "slab decomposition is not available on the generic plan");
    if (dkind == GGP_TABLE_SEP_AXES && !d.disp_table)
      return fail(GGP_ERR_UNSUPPORTED, "the generic plan needs the full dispersion table (disp_table), not per-axis factors");
    if (dkind == GGP_TABLE_SEP_AXES) dkind = GGP_TABLE_SCALAR;
    const bool ext = d.nl_c_ext != nullptr;
    if ((M > 2 || d.nl_kind == GGP_NL_MATRIX) && d.nl_kind != GGP_NL_NONE && (!d.nl_c_ext || !d.nl_g_ext))
      return fail(GGP_ERR_INVALID, "ncomp > 2 / GGP_NL_MATRIX: nl_c_ext and nl_g_ext are required");
    if (M > 2 && d.noise_kind != GGP_NOISE_NONE && !d.noise_eta_ext)
      return fail(GGP_ERR_INVALID, "ncomp > 2 with noise: noise_eta_ext is required");
    if (d.pot_kind == GGP_TABLE_FULL && d.nl_kind == GGP_NL_DIAG && !d.nl_scalar && M > 1)
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
    GGP_CUDA(cudaEventCreateWithFlags(&save_done, cudaEventDisableTiming));
    int rc;
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) {
      if ((rc = dalloc((void**)&u[c], bytes))) return rc;
      GGP_CUDA(cudaMemsetAsync(u[c], 0, bytes, stream));
    }
    if (dkind != GGP_TABLE_NONE) {
      if (!d.disp_table) return fail(GGP_ERR_INVALID, "disp_table is NULL");
      for (int a = 0; a < ndim; ++a)
        if ((rc = setup_axis(a))) return rc;
      // the inverse transforms are unnormalised; 1/prod(n) (ScaledPlan, src/misc.jl:56) rides in exp_D
      if ((rc = upload_aos(d.disp_table, d.table_precision, (size_t)nspatial * ncols_of(dkind, M), 1.0 / (double)nspatial, &D)))
        return rc;
    }
    memset(&pw, 0, sizeof(pw));
    pw.nspatial = nspatial;
    pw.total = nspatial * nbatch;
    pw.dt = (T)(dt / 2);
    pw.sqrt_dt = (T)std::sqrt(dt / 2);
    pw.vkind = d.pot_kind;
    if (d.pot_kind != GGP_TABLE_NONE) {
      if (!d.pot_table) return fail(GGP_ERR_INVALID, "pot_table is NULL");
      if ((rc = upload_aos(d.pot_table, d.table_precision, (size_t)nspatial * ncols_of(d.pot_kind, M), 1.0, &V))) return rc;
      pw.expV = V;
    }
    pump_kind = d.pump_kind;
    pump_ncomp = d.pump_ncomp;
    table_prec = d.table_precision;
    if (pump_kind != GGP_PUMP_NONE) {
      if (!d.pump_table) return fail(GGP_ERR_INVALID, "pump_table is NULL");
      if (d.pump_ncomp != 1 && d.pump_ncomp != M) return fail(GGP_ERR_INVALID, "pump_ncomp must be 1 or ncomp");
      pw.pump = pump_ncomp == 1 ? 1 : 2;
      const size_t cnt = (size_t)nspatial * pump_ncomp;
      if (pump_kind == GGP_PUMP_DENSE) {
        pw.pump_dense = 1;
        if ((rc = upload_aos(d.pump_table, d.table_precision, cnt, 1.0, &pd[0]))) return rc;   // primed at tspan[1]
        if ((rc = dalloc((void**)&pd[1], sizeof(cpx<T>) * cnt))) return rc;
        pd_next = 0;
        pd_stage.resize(cnt);
      } else {
        if ((rc = upload_aos(d.pump_table, d.table_precision, cnt, 1.0, &S))) return rc;
        pw.S = S;
        amp_prev = std::complex<double>(d.pump_amp0[0], d.pump_amp0[1]);
      }
    }
    auto cd = [](const double* q, size_t i) { return std::complex<double>(q[2 * i], q[2 * i + 1]); };
    if (d.nl_kind == GGP_NL_DIAG) {
      std::vector<std::complex<double>> c, g;
      const int rows = d.nl_scalar ? 1 : M;
      for (int i = 0; i < rows; ++i) {
        c.push_back(ext ? cd(d.nl_c_ext, (size_t)i) : std::complex<double>(d.nl_c[i][0], d.nl_c[i][1]));
        for (int j = 0; j < M; ++j)
          g.push_back(ext ? cd(d.nl_g_ext, (size_t)(i * M + j)) : std::complex<double>(d.nl_g[i][j][0], d.nl_g[i][j][1]));
      }
      bool real = true;
      for (auto& z : c) real = real && z.imag() == 0;
      for (auto& z : g) real = real && z.imag() == 0;
      pw.nlkind = d.nl_scalar ? 1 : 2;
      pw.nl_real = real ? 1 : 0;
      if ((rc = upload_coeffs(c, &pw.nlc))) return rc;
      if ((rc = upload_coeffs(g, &pw.nlg))) return rc;
    } else if (d.nl_kind == GGP_NL_MATRIX) {
      std::vector<std::complex<double>> c((size_t)M * M), g((size_t)M * M * M);
      for (size_t i = 0; i < c.size(); ++i) c[i] = cd(d.nl_c_ext, i);
      for (size_t i = 0; i < g.size(); ++i) g[i] = cd(d.nl_g_ext, i);
      pw.nlkind = 3;
      if ((rc = upload_coeffs(c, &pw.nlc))) return rc;
      if ((rc = upload_coeffs(g, &pw.nlg))) return rc;
    } else if (d.nl_kind != GGP_NL_NONE) {
      return fail(GGP_ERR_INVALID, "unknown nl_kind");
    }
    noise_kind = d.noise_kind;
    noise_real = d.noise_real;
    if (noise_kind != GGP_NOISE_NONE) {
      if (noise_kind != GGP_NOISE_CONST && noise_kind != GGP_NOISE_FIELD) return fail(GGP_ERR_INVALID, "unknown noise_kind");
      pw.noise = NOISE_PHILOX;
      pw.noise_real = d.noise_real;
      std::vector<std::complex<double>> eta, alpha;
      for (int i = 0; i < M; ++i)
        eta.push_back(d.noise_eta_ext ? cd(d.noise_eta_ext, (size_t)i) : std::complex<double>(d.noise_eta[i][0], d.noise_eta[i][1]));
      if ((rc = upload_coeffs(eta, &pw.eta))) return rc;
      if (noise_kind == GGP_NOISE_FIELD) {
        if (M > 2 && !d.noise_alpha_ext) return fail(GGP_ERR_INVALID, "ncomp > 2 with GGP_NOISE_FIELD: noise_alpha_ext is required");
        pw.noise_field = 1;
        pw.n1 = (int)n[0];
        for (int i = 0; i < M; ++i)
          for (int j = 0; j < M; ++j)
            alpha.push_back(d.noise_alpha_ext ? cd(d.noise_alpha_ext, (size_t)(i * M + j))
                                              : std::complex<double>(d.noise_alpha[i][j][0], d.noise_alpha[i][j][1]));
        if ((rc = upload_coeffs(alpha, &pw.alpha))) return rc;
        if (d.noise_profile) {
          std::vector<std::complex<double>> hp((size_t)n[0]);
          for (long long k = 0; k < n[0]; ++k) hp[(size_t)k] = cd((const double*)d.noise_profile, (size_t)k);
          if ((rc = upload_coeffs(hp, &pw.nprof))) return rc;
        }
      }
      pw.seed_lo = (uint32_t)d.seed;
      pw.seed_hi = (uint32_t)(d.seed >> 32);
      pw.elem_offset = batch_offset * nspatial;
    }
    has_pointwise = pw.vkind || pw.pump || pw.nlkind || pw.noise;
    if ((rc = dalloc((void**)&obs_dev, sizeof(double) * (size_t)(nspatial * M + 8)))) return rc;
    GGP_CUDA(cudaStreamSynchronize(stream));
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
    return 0;
  }
  void* state_ptr(int c) override { return (c >= 0 && c < M) ? (void*)u[c] : nullptr; }
  int ipc_export(void*) override { return fail(GGP_ERR_UNSUPPORTED, "slab decomposition is not available on the generic plan"); }
  int ipc_attach(const void*) override { return fail(GGP_ERR_UNSUPPORTED, "slab decomposition is not available on the generic plan"); }

  // saves: stream-ordered copy straight into the caller's (page-locked) arrays; the next step waits for it
  int save_async(void* const* uh) override {
    const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) GGP_CUDA(cudaMemcpyAsync(uh[c], u[c], bytes, cudaMemcpyDeviceToHost, stream));
    GGP_CUDA(cudaEventRecord(save_done, stream));
    save_pending = true;
    return 0;
  }
  int save_wait() override {
    if (save_pending) {
      GGP_CUDA(cudaEventSynchronize(save_done));
      save_pending = false;
    }
    return 0;
  }

  struct CkptHeader {
    uint64_t magic;
    uint32_t version, precision;
    int32_t ndim, ncomp;
    int64_t n[3], nbatch, batch_offset;
    uint64_t half_ctr;
    double amp_prev[2];
    uint64_t bytes_per_comp;
  };
  static constexpr uint64_t CKPT_MAGIC = 0x4e45474350474721ull;  // "!GGPCGEN"
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
    if (h.magic != CKPT_MAGIC || h.version != 1) return fail(GGP_ERR_INVALID, "not a checkpoint of a generic plan (magic/version)");
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

  // ---- stages ----
  template <int MM>
  int launch_pw_m(const GenPwParams<T>& q) {
    const unsigned blocks = (unsigned)((q.total + 127) / 128);
    gen_pw_kernel<T, MM><<<blocks, 128, 0, stream>>>(q);
    return (int)cudaGetLastError();
  }
  int run_pw(const GenHalf<T>& h) {
    if (!has_pointwise) return 0;
    GenPwParams<T> q = pw;
    for (int c = 0; c < M; ++c) q.u[c] = u[c];
    q.h = h;
    int rc = prof_begin(KC_ROW);
    if (rc) return rc;
    int e = 0;
    switch (M) {
      case 1: e = launch_pw_m<1>(q); break;
      case 2: e = launch_pw_m<2>(q); break;
      case 3: e = launch_pw_m<3>(q); break;
      case 4: e = launch_pw_m<4>(q); break;
      default: return fail(GGP_ERR_UNSUPPORTED, "ncomp beyond GGP_MAX_COMPONENTS");
    }
    GGP_LAUNCH(e, "gen_pw_kernel");
    ++launches;
    return prof_end();
  }
  template <int MM>
  int launch_disp_m(const GenDispParams<T>& q) {
    const unsigned blocks = (unsigned)((q.total + 127) / 128);
    gen_disp_kernel<T, MM><<<blocks, 128, 0, stream>>>(q);
    return (int)cudaGetLastError();
  }
  int run_disp() {
    GenDispParams<T> q;
    memset(&q, 0, sizeof(q));
    for (int c = 0; c < M; ++c) q.u[c] = u[c];
    q.D = D;
    q.nspatial = nspatial;
    q.total = nspatial * nbatch;
    q.dkind = dkind;
    int rc = prof_begin(KC_STR_D);
    if (rc) return rc;
    int e = 0;
    switch (M) {
      case 1: e = launch_disp_m<1>(q); break;
      case 2: e = launch_disp_m<2>(q); break;
      case 3: e = launch_disp_m<3>(q); break;
      case 4: e = launch_disp_m<4>(q); break;
      default: return fail(GGP_ERR_UNSUPPORTED, "ncomp beyond GGP_MAX_COMPONENTS");
    }
    GGP_LAUNCH(e, "gen_disp_kernel");
    ++launches;
    return prof_end();
  }
  // transform of axis a of one array (nspatial * nbatch elements); dcol >= 0: forward -> x exp_D column dcol -> inverse
  int run_fft(cpx<T>* arr, int a, bool inverse, int dcol = -1) {
    if (n[a] == 1 && dcol < 0) return 0;
    GenFftParams<T> q;
    memset(&q, 0, sizeof(q));
    if (dcol >= 0) {
      q.mode = 1;
      q.D = D;
      q.dcols = ncols_of(dkind, M);
      q.dcol = dcol;
      q.nspatial = nspatial;
    }
    q.u = arr;
    q.tw = tw[a];
    q.chirp = chirp[a];
    q.bhat = bhat[a];
    q.n = (int)n[a];
    q.sa = 1;
    for (int b = 0; b < a; ++b) q.sa *= n[b];
    q.nlines = nspatial * nbatch / n[a];
    q.inverse = inverse ? 1 : 0;
    int rc = prof_begin(KC_STR_FI);
    if (rc) return rc;
    GGP_LAUNCH(dispatch_gen_fft<T>((int)L[a], q, stream), "gen_fft_kernel");
    ++launches;
    return prof_end();
  }
  int run_ft_all(bool inverse) {
    int rc;
    for (int c = 0; c < M; ++c) {
      if (!inverse) {
        for (int a = 0; a < ndim; ++a)
          if ((rc = run_fft(u[c], a, false))) return rc;
      } else {
        for (int a = ndim - 1; a >= 0; --a)
          if ((rc = run_fft(u[c], a, true))) return rc;
      }
    }
    return 0;
  }

  int upload_noise(const void* const* noise, int64_t s, int half) {
    const size_t esz = noise_real ? sizeof(T) : sizeof(cpx<T>);
    const size_t bytes = esz * (size_t)nspatial * (size_t)nbatch;
    for (int c = 0; c < M; ++c) {
      if (!xi_dev[c]) {
        int rc = dalloc(&xi_dev[c], bytes);
        if (rc) return rc;
      }
      const void* src = noise[(size_t)((2 * s + half) * M + c)];
      if (!src) return fail(GGP_ERR_INVALID, "noise_host entry is NULL");
      GGP_CUDA(cudaMemcpyAsync(xi_dev[c], src, bytes, cudaMemcpyHostToDevice, stream));
    }
    return 0;
  }
  int upload_profile(const void* host) {
    if (!host) return fail(GGP_ERR_INVALID, "pump profile is NULL");
    const size_t cnt = (size_t)nspatial * pump_ncomp;
    // the staging vector is reused: wait until the previous copy out of it has been issued AND has completed
    GGP_CUDA(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < cnt; ++i) {
      const std::complex<double> z = host_entry(host, table_prec, i);
      pd_stage[i] = mk<T>((T)z.real(), (T)z.imag());
    }
    pd_next ^= 1;
    GGP_CUDA(cudaMemcpyAsync(pd[pd_next], pd_stage.data(), sizeof(cpx<T>) * cnt, cudaMemcpyHostToDevice, stream));
    return 0;
  }

  int step(int64_t nsteps, const double* amp, const void* const* noise, const void* const* profiles) override {
    if (nsteps <= 0) return 0;
    if (noise && noise_kind == GGP_NOISE_NONE) return fail(GGP_ERR_INVALID, "noise_host given but the plan has no noise term");
    const bool dense = pump_kind == GGP_PUMP_DENSE;
    if (dense && !profiles) return fail(GGP_ERR_INVALID, "GGP_PUMP_DENSE plans are stepped with ggp_step_dense (pump profiles per half-step)");
    if (!dense && profiles) return fail(GGP_ERR_INVALID, "ggp_step_dense on a plan without a dense pump");
    pw.noise = noise_kind == GGP_NOISE_NONE ? NOISE_OFF : (noise ? NOISE_HOST : NOISE_PHILOX);
    int rc;
    for (int64_t s = 0; s < nsteps; ++s) {
      if ((rc = window_begin())) return rc;
      for (int half = 0; half < 2; ++half) {
        // potential_pump_step! (src/strang_splitting.jl:78-84): new noise, pump pair (F_now, F_next), muladd
        GenHalf<T> h;
        memset(&h, 0, sizeof(h));
        if (noise && (rc = upload_noise(noise, s, half))) return rc;
        for (int c = 0; c < M; ++c) h.xi[c] = xi_dev[c];
        const double q = dt / 4;
        if (dense) {
          if ((rc = upload_profile(profiles[2 * s + half]))) return rc;
          h.pd_now = pd[pd_next ^ 1];
          h.pd_next = pd[pd_next];
          h.fnow = h.fnext = mk<T>((T)q, (T)0);
        } else if (pump_kind != GGP_PUMP_NONE) {
          const std::complex<double> a =
              amp ? std::complex<double>(amp[2 * (2 * s + half)], amp[2 * (2 * s + half) + 1]) : amp_prev;
          h.fnow = mk<T>((T)(q * amp_prev.real()), (T)(q * amp_prev.imag()));
          h.fnext = mk<T>((T)(q * a.real()), (T)(q * a.imag()));
          amp_prev = a;
        }
        h.ctr = (uint32_t)half_ctr;
        h.ctr_hi = (uint32_t)(half_ctr >> 32);
        ++half_ctr;
        if ((rc = run_pw(h))) return rc;
        if (half == 0 && dkind != GGP_TABLE_NONE) {
          // diffusion_step! (src/strang_splitting.jl:69-76)
          static const bool nofuse = getenv("GGP_GEN_NOFUSE") != nullptr;
          const int last = ndim - 1;
          if (dkind != GGP_TABLE_FULL && n[last] > 1 && !nofuse) {
            // Number / SVector exp_D: the multiply rides in the transform of the last axis, component by component
            for (int c = 0; c < M; ++c) {
              for (int a = 0; a < last; ++a)
                if ((rc = run_fft(u[c], a, false))) return rc;
              if ((rc = run_fft(u[c], last, false, dkind == GGP_TABLE_DIAG ? c : 0))) return rc;
              for (int a = last - 1; a >= 0; --a)
                if ((rc = run_fft(u[c], a, true))) return rc;
            }
          } else {
            if ((rc = run_ft_all(false))) return rc;
            if ((rc = run_disp())) return rc;
            if ((rc = run_ft_all(true))) return rc;
          }
        }
      }
      if ((rc = window_end())) return rc;
    }
    return 0;
  }

  int observe_windowed(const double*, const double*, double*) override {
    return fail(GGP_ERR_UNSUPPORTED, "windowed correlations are not available on the generic plan");
  }
  int observe(int kind, double* out) override {
    const int threads = 256;
    const unsigned blocks = (unsigned)((nspatial + threads - 1) / threads);
    size_t count = 0;
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
      // n(k) = sum_traj |fft(u)(k)|^2 / N^2 (examples/truncated_wigner.jl:110-113) on a scratch copy
      const size_t bytes = sizeof(cpx<T>) * (size_t)nspatial * (size_t)nbatch;
      int rc;
      if (!scratch && (rc = dalloc((void**)&scratch, bytes))) return rc;
      for (int a = 0; a < ndim; ++a)
        if (!L[a] && n[a] > 1 && (rc = setup_axis(a))) return rc;
      const double sc = 1.0 / ((double)nspatial * (double)nspatial);
      for (int c = 0; c < M; ++c) {
        GGP_CUDA(cudaMemcpyAsync(scratch, u[c], bytes, cudaMemcpyDeviceToDevice, stream));
        for (int a = 0; a < ndim; ++a)
          if ((rc = run_fft(scratch, a, false))) return rc;
        density_kernel<T><<<blocks, threads, 0, stream>>>(scratch, obs_dev + c * nspatial, nspatial, nbatch, sc);
        ++launches;
      }
      count = (size_t)(nspatial * M);
    } else {
      return fail(GGP_ERR_UNSUPPORTED, "this observable is not available on the generic plan");
    }
    GGP_CUDA(cudaGetLastError());
#ifdef GGP_WITH_NCCL
    if (comm && nranks > 1) {
      ncclResult_t r = ncclAllReduce(obs_dev, obs_dev, count, ncclDouble, ncclSum, comm, stream);
      if (r != ncclSuccess) return fail(GGP_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    }
#endif
    GGP_CUDA(cudaMemcpyAsync(out, obs_dev, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
    GGP_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
};

// does this descriptor need (or ask for) the generic plan?
static bool gen_wanted(const ggp_desc& d) {
  if (d.slab_nranks > 1) return false;
  if (d.ncomp > 2 || d.nl_kind == GGP_NL_MATRIX) return true;
  for (int i = 0; i < d.ndim; ++i)
    if (!size_supported(d.n[i])) return true;   // axes the fused kernels are not instantiated for (any n: Bluestein)
  // two ComplexF64 components on an 8192-point axis: the fused kernels would need 128 data registers in each of 512
  // threads per line (more than an SM's register file), the per-component transforms of the generic plan do not
  if (d.ncomp == 2 && d.precision == GGP_C128)
    for (int i = 0; i < d.ndim; ++i)
      if (d.n[i] >= 8192) return true;
  const char* e = getenv("GGP_FORCE_GENERIC");
  return e && atoi(e) != 0;
}

}  // namespace ggp
