// One FFT line per group of TPL = N/E threads: every thread keeps E elements of the line in
// registers (positions t + m*TPL, m = 0..E-1, coalesced across t), Stockham auto-sort passes of
// radix E (last pass: the remainder radix), data exchanged between passes through a padded
// shared-memory line.  First-pass inputs and last-pass outputs stay in registers, so
//   * global loads/stores fuse into the first / last pass (no staging sweep),
//   * a point-wise stage fuses between an inverse and a forward transform without touching smem.
// Natural order in, natural order out -- tables and HBM layouts need no permutation.
#pragma once
#include "cplx.cuh"

// tuning knobs: default elements per thread (= radix of a full pass) for fp32 / fp64 lines
#ifndef GGP_E32
#define GGP_E32 16
#endif
#ifndef GGP_E64
#define GGP_E64 8
#endif
// fp64 lines: form the twiddle powers w^2 .. w^(R-1) from w instead of reading them (see Passes::run)
#ifndef GGP_TW_POWERS64
#define GGP_TW_POWERS64 1
#endif

namespace ggp {

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// Elements per thread.  fp32: radix-16 passes (32 data registers); fp64: radix-8 (32 data
// registers).  Long lines use a wider radix so that a line never needs more than 256 threads.
template <typename T>
struct IsPacked {
  static constexpr bool value = false;
};
template <>
struct IsPacked<f2> {
  static constexpr bool value = true;
};
template <typename T>
__host__ __device__ constexpr int default_E(int N) {
  // fp32 (and the packed pairs): radix 16, radix 32 only where a line would otherwise need more than 512
  // threads; fp64: radix 8, radix 16 where a line would otherwise need more than 256 -- never wider than 64 data registers per component
  // (radix-32 fp64 = 128 data registers spills), so the longest fp64 lines take 512 threads instead.
  const bool f32 = sizeof(T) == 4 || IsPacked<T>::value;
  int e = f32 ? GGP_E32 : GGP_E64;
  const int emax = f32 ? 32 : 16;
  // fp32 lines may take up to 512 threads: radix 32 for N = 8192 needed 151 registers in the row kernel (one
  // 256-thread CTA = 8 warps per SM); radix 16 with 512 threads keeps 64 registers and 32 warps per SM
  // (8192^2: row 626 -> 467 us, strided 1230 -> 1067 us, profiles/r01_notes.md session 4)
  while (N / e > (f32 ? 512 : 256) && e < emax) e *= 2;
  return e < N ? e : N;
}

// EO > 0 overrides the default elements-per-thread / radix for one kernel (e.g. the stochastic fp64 row kernel, which is
// bound by the latency of its Philox / sincos chains and wants twice the warps: KCfg::row_E in kernels.cuh)
template <typename T, int N, int EO = 0>
struct LineCfg {
  static constexpr int E = EO > 0 ? (EO < N ? EO : N) : default_E<T>(N);
  static constexpr int TPL = N / E;         // threads per line
  static constexpr int LOGE = ilog2(E);
  // padded length of one line in shared memory: one pad element per E elements keeps the
  // scattered Stockham writes (stride R) and the strided reads conflict-free
  static constexpr int PADN = N + N / E;
  __host__ __device__ static constexpr int pad(int i) { return i + (i >> LOGE); }
};

struct SyncBlock {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
};
struct SyncWarp {
  static __device__ __forceinline__ void sync() { __syncwarp(); }
};

// Twiddle table layout (built on the host, ggp_api.cu): for every pass after the first, in order,
// a block of (R-1)*NS entries  tw[off + (r-1)*NS + k] = exp(-2 pi i r k / (NS*R)),  k < NS, 1 <= r < R,
// so that consecutive lanes (consecutive k) read consecutive addresses.
template <typename T, int N, int EO = 0>
__host__ __device__ constexpr int twiddle_count(int NS = 1) {
  constexpr int E = LineCfg<T, N, EO>::E;
  if (NS >= N) return 0;
  const int R = (N / NS >= E) ? E : N / NS;
  return (NS > 1 ? (R - 1) * NS : 0) + twiddle_count<T, N, EO>(NS * R);
}

// Long lines (N >= 4096): the late passes (NS >= 256) need (R-1)*NS twiddles -- a table as large as the line
// itself, re-read by every CTA for every tile (4096: 30 KB per transform next to 64 KB of tile data; ncu r01r:
// 45 % of the strided kernel's stall samples wait for these loads, L1 hit rate 38 %).  With FACT the twiddle is
// formed on the fly instead,  w_N^j = A[j >> 6] * B[j & 63]  (two tables of N/64 and 64 entries, both exact to the
// plan's precision, one extra complex multiply), and only the early passes' small blocks stay a table.
// Compact table layout (host, ggp_api.cu):  twc[0..63] = B[j] = w_N^j,  twc[64 + j] = A[j] = w_N^(64 j).
constexpr int TWF_LO = 64;
constexpr int TWF_MIN_NS = 256;
template <typename T, int N>
__host__ __device__ constexpr bool tw_factorized() {
  return N >= 4096 && !TwT<T>::split;
}
template <typename T, int N>
__host__ __device__ constexpr int twiddle_compact_count() {
  return tw_factorized<T, N>() ? TWF_LO + N / TWF_LO : 0;
}
// entries of the passes with NS < TWF_MIN_NS: a prefix of the table (the blocks are stored in pass order)
template <typename T, int N>
__host__ __device__ constexpr int twiddle_small_count(int NS = 1) {
  constexpr int E = default_E<T>(N);
  if (NS >= N || NS >= TWF_MIN_NS) return 0;
  const int R = (N / NS >= E) ? E : N / NS;
  return (NS > 1 ? (R - 1) * NS : 0) + twiddle_small_count<T, N>(NS * R);
}

// Stockham pass (NS = product of the radices of the earlier passes) and all later passes.
// DIR = -1 forward (e^{-i k x}), +1 inverse (unnormalised); TWOFF = offset of this pass' twiddles.
template <typename T, int N, int DIR, typename SYNC, int NS, int TWOFF = 0, bool FACT = false, int EO = 0>
struct Passes {
  using Cfg = LineCfg<T, N, EO>;
  static constexpr int E = Cfg::E;
  static constexpr int TPL = Cfg::TPL;
  static constexpr int R = (N / NS >= E) ? E : N / NS;
  static constexpr int NB = E / R;  // butterflies per thread in this pass
  static constexpr bool LASTP = (NS * R == N);

  static __device__ __forceinline__ void run(cpx<T> (&v)[E], const int t, cpx<T>* __restrict__ line,
                                             const typename TwT<T>::type* __restrict__ tw,
                                             const typename TwT<T>::type* __restrict__ twc = nullptr) {
    if constexpr (NS > 1) {
      if constexpr (TPL % E == 0) {
        // pad(t + m*TPL) = pad(t) + m*(TPL + TPL/E): one base, immediate offsets
        const cpx<T>* rd = line + Cfg::pad(t);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = rd[m * (TPL + TPL / E)];
      } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = line[Cfg::pad(t + m * TPL)];
      }
      if constexpr (!LASTP) SYNC::sync();  // everybody has read before anybody overwrites
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      cpx<T> a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = v[q + r * NB];
      const int b = t + q * TPL;
      const int k = b & (NS - 1);
      if constexpr (NS > 1) {
        static_assert(DIR < 0, "the inverse transform runs the forward code on conjugated data");
        if constexpr (FACT && NS >= TWF_MIN_NS && tw_factorized<T, N>()) {
          const int ks = k * (N / (NS * R));  // w_{NS R}^(r k) = w_N^(r ks)
#pragma unroll
          for (int r = 1; r < R; ++r) {
            const int j = r * ks;
            a[r] = cmul(a[r], cmul(twc[TWF_LO + (j >> 6)], twc[j & (TWF_LO - 1)]));
          }
        } else if constexpr (GGP_TW_POWERS64 && sizeof(T) == 8 && N >= 512 && R >= 4 && !TwT<T>::split) {
          // fp64 lines of 512 points and more: only w = w^1 is read; w^2 .. w^(R-1) are formed by squaring / multiplying
          // (depth log2 R): R-2 fewer 16-byte table reads per butterfly for 4 flops each.  Measured (r02x, one call):
          // 1024^2 ComplexF64 Kerr step 32.8 -> 29.5 us, C3 (1024^2, two components) unchanged; on 256-point lines
          // (C4), whose tables stay in L1 / shared memory, it is 3 % SLOWER -- hence the length condition
          cpx<T> w[R];
          w[1] = tw[TWOFF + k];
#pragma unroll
          for (int r = 2; r < R; ++r) w[r] = (r & 1) ? cmul(w[r - 1], w[1]) : cmul(w[r / 2], w[r / 2]);
#pragma unroll
          for (int r = 1; r < R; ++r) a[r] = cmul(a[r], w[r]);
        } else {
          const typename TwT<T>::type* twk = tw + TWOFF + k;
#pragma unroll
          for (int r = 1; r < R; ++r) a[r] = TwT<T>::mul(a[r], twk[(r - 1) * NS]);
        }
      }
      Dft<T, R, DIR>::run(a);
      if constexpr (LASTP) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[q + r * NB] = a[r];
      } else {
        const int base = (b - k) * R + k;
        if constexpr (NS % E == 0) {
          // pad(base + r*NS) = pad(base) + r*(NS + NS/E)
          cpx<T>* wr = line + Cfg::pad(base);
#pragma unroll
          for (int r = 0; r < R; ++r) wr[r * (NS + NS / E)] = a[r];
        } else if constexpr (NS == 1 && R == E) {
          // first pass: base = b*E, pad(b*E + r) = b*(E+1) + r
          cpx<T>* wr = line + b * (E + 1);
#pragma unroll
          for (int r = 0; r < R; ++r) wr[r] = a[r];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) line[Cfg::pad(base + r * NS)] = a[r];
        }
      }
    }
    if constexpr (!LASTP) {
      SYNC::sync();
      Passes<T, N, DIR, SYNC, NS * R, TWOFF + (NS > 1 ? (R - 1) * NS : 0), FACT, EO>::run(v, t, line, tw, twc);
    }
  }
};

// Transform one line held in registers.  PRESYNC: the shared line may still be read by a previous
// transform of the same thread group, so synchronise before the first scatter.
template <typename T, int N, int DIR, typename SYNC, bool PRESYNC, bool FACT = false, int EO = 0>
__device__ __forceinline__ void fft_line(cpx<T> (&v)[LineCfg<T, N, EO>::E], const int t, cpx<T>* __restrict__ line,
                                         const typename TwT<T>::type* __restrict__ tw,
                                         const typename TwT<T>::type* __restrict__ twc = nullptr) {
  if (PRESYNC && LineCfg<T, N, EO>::E < N) SYNC::sync();
  Passes<T, N, DIR, SYNC, 1, 0, FACT, EO>::run(v, t, line, tw, twc);
}

}  // namespace ggp
