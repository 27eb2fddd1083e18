// Generic (unfused) line transform: one DFT of ANY length n <= L along ANY axis of the state.
//
// The fused kernels of kernels.cuh cover power-of-two axes and one or two components.  Everything else the
// reference accepts -- axes of any length (`plan_fft`, src/misc.jl:53-58, is FFTW: any n), more than two components
// (`NTuple{M}`, src/kernels.jl:31-35), matrix-valued nonlinearities (src/kernels.jl:22-25,44) -- runs through the
// generic plan (generic_plan.cuh): the reference's own sequence  muladd -> fft -> muladd -> ifft -> muladd  with one
// kernel per stage.  This file is the transform stage.
//
//   n == L (power of two)   plain Stockham transform of the line (fft_line.cuh)
//   n <  L                  Bluestein: X_k = w_k * sum_j (x_j w_j) conj(w_{k-j}),  w_j = exp(-i pi j^2 / n): a circular
//                           convolution of length L >= 2n - 1 -- chirp multiply, FFT_L, multiply by the transformed
//                           chirp filter (1/L folded in), inverse FFT_L, chirp multiply -- with the line resident in
//                           registers / shared memory from its load to its store.
// The inverse transform is conj(forward(conj(x))), unnormalised; 1/prod(n) rides in exp_D as in the fused path.
#pragma once
#include "fft_line.cuh"

namespace ggp {

template <typename T>
struct GenFftParams {
  cpx<T>* u;                            // one component: nspatial * nbatch elements
  const typename TwT<T>::type* tw;      // twiddles of the length-L transform (build_twiddles(L, default_E))
  const cpx<T>* chirp;                  // Bluestein: w_j, j < n;  nullptr: plain transform (n == L)
  const cpx<T>* bhat;                   // Bluestein: FFT_L(conj-chirp filter) / L, L entries
  long long nlines;                     // lines of this launch = total elements / n
  long long sa;                         // stride of the axis = number of adjacent lines that are contiguous in memory
  int n;                                // points per line
  int W;                                // lines per CTA
  int LS;                               // shared-memory stride between the lines of a CTA (elements)
  int inverse;
  // mode 1 (last FFT axis, scalar / SVector exp_D): forward -> x exp_D[k] -> inverse in one launch, the line resident
  // throughout (the k-space muladd of src/strang_splitting.jl:73-74 fused into the transform; two sweeps fewer per step)
  int mode;
  const cpx<T>* D;                      // point-major [k][dcols]
  int dcols, dcol;
  long long nspatial;
};

// thread (xw, t): line blockIdx.x * W + xw, elements t + m * TPL.  Strided axes (sa > 1): xw fastest, so that the
// W lines of a CTA are adjacent in memory; contiguous axis (sa == 1): t fastest.
template <typename T, int L>
__global__ void __launch_bounds__(512, 1) gen_fft_kernel(const GenFftParams<T> p) {
  using Cfg = LineCfg<T, L>;
  constexpr int E = Cfg::E, TPL = Cfg::TPL;
  extern __shared__ __align__(16) unsigned char gen_smem_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(gen_smem_raw);
  int xw, t;
  if (p.sa > 1) {
    xw = threadIdx.x % p.W;
    t = threadIdx.x / p.W;
  } else {
    t = threadIdx.x % TPL;
    xw = threadIdx.x / TPL;
  }
  const long long l = (long long)blockIdx.x * p.W + xw;
  const bool active = l < p.nlines;
  const long long inner = l % p.sa, outer = l / p.sa;
  cpx<T>* const base = p.u + inner + outer * p.sa * (long long)p.n;
  cpx<T>* const sl = smem + (size_t)xw * p.LS;
  const bool blu = p.chirp != nullptr;

  cpx<T> v[E];
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const int j = t + m * TPL;
    v[m] = (active && j < p.n) ? base[(long long)j * p.sa] : mk<T>((T)0, (T)0);
  }
  // forward DFT of the n points held in v (zero beyond n), in place
  auto dft = [&](const bool presync) {
    if (blu) {
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int j = t + m * TPL;
        if (j < p.n) v[m] = cmul(v[m], p.chirp[j]);
      }
    }
    if (presync) fft_line<T, L, -1, SyncBlock, true>(v, t, sl, p.tw);
    else fft_line<T, L, -1, SyncBlock, false>(v, t, sl, p.tw);
    if (blu) {
#pragma unroll
      for (int m = 0; m < E; ++m) {
        v[m] = cmul(v[m], p.bhat[t + m * TPL]);
        v[m].y = -v[m].y;
      }
      fft_line<T, L, -1, SyncBlock, true>(v, t, sl, p.tw);
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int j = t + m * TPL;
        v[m].y = -v[m].y;
        v[m] = j < p.n ? cmul(v[m], p.chirp[j]) : mk<T>((T)0, (T)0);
      }
    }
  };
  if (p.mode == 1) {
    dft(false);
    const long long e0 = inner + outer * p.sa * (long long)p.n;
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int j = t + m * TPL;
      if (active && j < p.n) {
        const long long sidx = (e0 + (long long)j * p.sa) % p.nspatial;
        v[m] = cmul(p.D[sidx * p.dcols + p.dcol], v[m]);
      }
      v[m].y = -v[m].y;
    }
    dft(true);
#pragma unroll
    for (int m = 0; m < E; ++m) v[m].y = -v[m].y;
  } else {
    if (p.inverse) {
#pragma unroll
      for (int m = 0; m < E; ++m) v[m].y = -v[m].y;
    }
    dft(false);
    if (p.inverse) {
#pragma unroll
      for (int m = 0; m < E; ++m) v[m].y = -v[m].y;
    }
  }
  if (active) {
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int j = t + m * TPL;
      if (j < p.n) base[(long long)j * p.sa] = v[m];
    }
  }
}

// geometry of one launch (host): lines per CTA and the padded shared-memory stride
template <typename T, int L>
void gen_fft_geometry(long long sa, int* W, int* LS, int* threads, size_t* smem) {
  using Cfg = LineCfg<T, L>;
  constexpr int TPL = Cfg::TPL;
  int w = 128 / TPL;                    // at least 128 threads per CTA
  if (w < 1) w = 1;
  if (sa > 1) {                         // strided axis: up to 64 bytes of adjacent lines per row, at most 512 threads
    int want = 64 / (int)sizeof(cpx<T>);
    while (want > 1 && want * TPL > 512) want >>= 1;
    if (w < want) w = want;
  }
  int ls = Cfg::PADN;
  if (w > 1) ls |= 1;                   // odd stride between lines: neighbouring lines start in different banks
  while ((size_t)w * ls * sizeof(cpx<T>) > (size_t)200 * 1024 && w > 1) w >>= 1;
  *W = w;
  *LS = ls;
  *threads = w * TPL;
  *smem = Cfg::E < L ? (size_t)w * ls * sizeof(cpx<T>) : 0;
}

template <typename T, int L>
int launch_gen_fft(GenFftParams<T> p, cudaStream_t st);

}  // namespace ggp
