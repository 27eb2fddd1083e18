// The three hot kernels of the Strang step (replacing src/strang_splitting.jl:69-90 + src/kernels.jl):
//
//   row_kernel   contiguous-axis lines:  [inverse FFT_x] -> V/2 of step n -> V/2 of step n+1 -> [forward FFT_x]
//                (the trailing half-step of one step and the leading half-step of the next are two
//                 sequential in-register applications, never one dt application -- SURVEY Q7)
//   str_kernel   strided-axis lines, W adjacent fast-axis positions per CTA for coalescing:
//                forward FFT -> x exp_D[k] -> inverse FFT  (MODE 1), or forward only / inverse only
//                for the middle axis of a 3-D grid
//   oned_kernel  1-D problems: the whole step, looped over many steps with the line resident on chip
//
// All kernels are in place: HBM holds one copy of the state (no ft_buffer, cf. src/strang_splitting.jl:49).
#pragma once
#include <type_traits>
#include "fft_line.cuh"
#include "pointwise.cuh"

namespace ggp {

template <typename T>
struct RowParams {
  cpx<T>* u[2];
  const cpx<T>* tw;
  long long nlines;           // (nspatial / N) * nbatch
  long long lines_per_image;  // nspatial / N ; spatial table line = line % lines_per_image
  PointwiseParams<T> pw;
  HalfStep<T> hA, hB;
};

template <typename T>
struct StrParams {
  cpx<T>* u[2];
  const cpx<T>* tw;
  long long ls;      // stride (elements) between consecutive points of a line
  long long ntx;     // tiles of W along the fast axis
  long long no1, s1, s2;  // remaining dims: offset = (o % no1) * s1 + (o / no1) * s2
  long long ts1;     // table stride of remaining dim 1 (tables do not have the batch dim)
  const cpx<T>* D[4];
  int dkind;
  int W, logW, LS;
};

template <typename T>
struct OneDParams {
  cpx<T>* u[2];
  const cpx<T>* tw;
  long long nlines;
  PointwiseParams<T> pw;
  const HalfStep<T>* hs;  // 2 * nsteps entries (device)
  int nsteps;
  const cpx<T>* D[4];
  int dkind;
};

template <typename T, int N>
struct KCfg {
  using L = LineCfg<T, N>;
  static constexpr int E = L::E;
  static constexpr int TPL = L::TPL;
  static constexpr int G = 128 / (int)sizeof(cpx<T>);  // lanes served by one shared-memory wavefront
  static constexpr int ROW_THREADS = TPL >= 128 ? TPL : 128;
  static constexpr int LPC = ROW_THREADS / TPL;  // lines per CTA (row / 1-D kernels)
  __host__ __device__ static constexpr int row_ls() {
    int ls = L::PADN;
    if (TPL < G)
      while (ls % G != TPL % G) ++ls;
    return ls;
  }
  __host__ static int str_ls(int W) {
    int lpg = W >= G ? 1 : G / W;
    int ls = L::PADN;
    while (ls % G != lpg % G) ++ls;
    return ls;
  }
  static constexpr bool USES_SMEM = E < N;
  // strided kernel: W adjacent fast-axis positions per CTA (coalescing width)
  static constexpr int WMIN = sizeof(T) == 4 ? 4 : 2;
  static constexpr int WDEF = (256 / TPL) < WMIN ? WMIN : ((256 / TPL) > 32 ? 32 : (256 / TPL));
  static constexpr int STR_THREADS = WDEF * TPL;
  using RowSync = typename std::conditional<(TPL <= 32), SyncWarp, SyncBlock>::type;
};

template <typename T, int N, int M, bool PRE, bool POST>
__global__ void __launch_bounds__(KCfg<T, N>::ROW_THREADS) row_kernel(const RowParams<T> p) {
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL, LPC = K::LPC, LS = K::row_ls();
  using SYNC = typename K::RowSync;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_raw);

  const int grp = threadIdx.x / TPL, t = threadIdx.x % TPL;
  const long long line = (long long)blockIdx.x * LPC + grp;
  const bool active = line < p.nlines;
  const long long goff = line * N + t;
  const long long soff = (line % p.lines_per_image) * N + t;
  cpx<T>* sl = smem + (size_t)grp * M * LS;

  cpx<T> v[M][E];
#pragma unroll
  for (int c = 0; c < M; ++c)
#pragma unroll
    for (int m = 0; m < E; ++m) v[c][m] = active ? p.u[c][goff + m * TPL] : mk<T>((T)0, (T)0);

  if (PRE) {
#pragma unroll
    for (int c = 0; c < M; ++c) fft_line<T, N, +1, SYNC, false>(v[c], t, sl + c * LS, p.tw);
  }
  if (p.hA.apply && active) {
#pragma unroll
    for (int m = 0; m < E; ++m) {
      cpx<T> f[M];
#pragma unroll
      for (int c = 0; c < M; ++c) f[c] = v[c][m];
      half_step_point<T, M>(f, p.pw, p.hA, soff + m * TPL, goff + m * TPL);
#pragma unroll
      for (int c = 0; c < M; ++c) v[c][m] = f[c];
    }
  }
  if (p.hB.apply && active) {
#pragma unroll
    for (int m = 0; m < E; ++m) {
      cpx<T> f[M];
#pragma unroll
      for (int c = 0; c < M; ++c) f[c] = v[c][m];
      half_step_point<T, M>(f, p.pw, p.hB, soff + m * TPL, goff + m * TPL);
#pragma unroll
      for (int c = 0; c < M; ++c) v[c][m] = f[c];
    }
  }
  if (POST) {
#pragma unroll
    for (int c = 0; c < M; ++c) fft_line<T, N, -1, SYNC, PRE>(v[c], t, sl + c * LS, p.tw);
  }
  if (active) {
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) p.u[c][goff + m * TPL] = v[c][m];
  }
}

// MODE 0: forward only, 1: forward -> x D -> inverse, 2: inverse only
template <typename T, int N, int M, int MODE>
__global__ void __launch_bounds__(KCfg<T, N>::STR_THREADS) str_kernel(const StrParams<T> p) {
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_raw);

  const int xw = threadIdx.x & (p.W - 1), t = threadIdx.x >> p.logW;
  const long long g = blockIdx.x;
  const long long xt = g % p.ntx, o = g / p.ntx;
  const long long o1 = o % p.no1, o2 = o / p.no1;
  const long long off = xt * p.W + xw + o1 * p.s1 + o2 * p.s2 + (long long)t * p.ls;
  const long long toff = xt * p.W + xw + o1 * p.ts1 + (long long)t * p.ls;
  const long long mstride = (long long)TPL * p.ls;
  cpx<T>* sl = smem + (size_t)xw * M * p.LS;

  cpx<T> v[M][E];
#pragma unroll
  for (int c = 0; c < M; ++c)
#pragma unroll
    for (int m = 0; m < E; ++m) v[c][m] = p.u[c][off + m * mstride];

  if (MODE != 2) {
#pragma unroll
    for (int c = 0; c < M; ++c) fft_line<T, N, -1, SyncBlock, false>(v[c], t, sl + c * p.LS, p.tw);
  }
  if (MODE == 1) {
#pragma unroll
    for (int m = 0; m < E; ++m) {
      cpx<T> f[M];
#pragma unroll
      for (int c = 0; c < M; ++c) f[c] = v[c][m];
      disp_point<T, M>(f, p.D, p.dkind, toff + m * mstride);
#pragma unroll
      for (int c = 0; c < M; ++c) v[c][m] = f[c];
    }
  }
  if (MODE != 0) {
#pragma unroll
    for (int c = 0; c < M; ++c) fft_line<T, N, +1, SyncBlock, MODE == 1>(v[c], t, sl + c * p.LS, p.tw);
  }
#pragma unroll
  for (int c = 0; c < M; ++c)
#pragma unroll
    for (int m = 0; m < E; ++m) p.u[c][off + m * mstride] = v[c][m];
}

template <typename T, int N, int M>
__global__ void __launch_bounds__(KCfg<T, N>::ROW_THREADS) oned_kernel(const OneDParams<T> p) {
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL, LPC = K::LPC, LS = K::row_ls();
  using SYNC = typename K::RowSync;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_raw);

  const int grp = threadIdx.x / TPL, t = threadIdx.x % TPL;
  const long long line = (long long)blockIdx.x * LPC + grp;
  const bool active = line < p.nlines;
  const long long goff = line * N + t;
  cpx<T>* sl = smem + (size_t)grp * M * LS;

  cpx<T> v[M][E];
#pragma unroll
  for (int c = 0; c < M; ++c)
#pragma unroll
    for (int m = 0; m < E; ++m) v[c][m] = active ? p.u[c][goff + m * TPL] : mk<T>((T)0, (T)0);

  for (int s = 0; s < p.nsteps; ++s) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const HalfStep<T> h = p.hs[2 * s + half];
      if (h.apply && active) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          cpx<T> f[M];
#pragma unroll
          for (int c = 0; c < M; ++c) f[c] = v[c][m];
          half_step_point<T, M>(f, p.pw, h, t + m * TPL, goff + m * TPL);
#pragma unroll
          for (int c = 0; c < M; ++c) v[c][m] = f[c];
        }
      }
      if (half == 0 && p.dkind != KIND_NONE) {
#pragma unroll
        for (int c = 0; c < M; ++c) fft_line<T, N, -1, SYNC, true>(v[c], t, sl + c * LS, p.tw);
#pragma unroll
        for (int m = 0; m < E; ++m) {
          cpx<T> f[M];
#pragma unroll
          for (int c = 0; c < M; ++c) f[c] = v[c][m];
          disp_point<T, M>(f, p.D, p.dkind, t + m * TPL);
#pragma unroll
          for (int c = 0; c < M; ++c) v[c][m] = f[c];
        }
#pragma unroll
        for (int c = 0; c < M; ++c) fft_line<T, N, +1, SYNC, true>(v[c], t, sl + c * LS, p.tw);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) p.u[c][goff + m * TPL] = v[c][m];
  }
}

// ---- launchers (explicitly instantiated per (T, N) in inst.cu) --------------------------------
template <typename T, int N>
int launch_row(int M, bool pre, bool post, const RowParams<T>& p, cudaStream_t st);
template <typename T, int N>
int launch_str(int M, int mode, StrParams<T> p, long long nfast, long long ngroups_other, cudaStream_t st);
template <typename T, int N>
int launch_oned(int M, const OneDParams<T>& p, cudaStream_t st);

}  // namespace ggp
