// Persistent, TMA-fed version of the strided-axis kernel (forward FFT -> x exp_D -> inverse FFT).
//
// One CTA per tile of W adjacent fast-axis positions x the whole line (N points along the strided
// axis).  The grid is sized to the machine (resident CTAs x SMs) and every CTA walks tiles
// tile = blockIdx.x + i*gridDim.x.  While tile i is being transformed out of registers, the TMA
// engine (cp.async.bulk.tensor, box = [W complex] x [256 rows]) already fills the staging buffer
// with tile i+1 and signals an mbarrier -- the load phase, which dominated the first version of this
// kernel (ncu: stall_long_scoreboard, profiles/r01_notes.md), overlaps the butterflies.
// The exp_D lines of the tile are pulled towards L2 with prefetch.global.L2 at tile start.
#pragma once
#include <cuda.h>
#include "kernels.cuh"

namespace ggp {

template <typename T>
struct alignas(64) StrTmaParams {
  CUtensorMap map[2];   // one per field component: rank-4 tensor [2*n1 reals, n2, n3, batch]
  CUtensorMap dmap[4];  // exp_D planes as [2*n1, n2, n3, 1] (used when stage_d)
  int stage_d;          // 1: the exp_D tile is staged through shared memory by TMA as well
  int nplanes;          // planes of exp_D (1, M or M*M)
  cpx<T>* u[2];
  const typename TwT<T>::type* tw;
  long long ls;        // stride (elements) between consecutive points of a line
  long long ntx;       // tiles of W along the fast axis
  long long ntiles;
  long long no1, s1, s2, ts1;
  const cpx<T>* D[4];
  int dkind;
  int mode;
  int ax;              // 1: lines along n2, 2: lines along n3
  int W, logW, LS;
};

template <typename T, int N, int M>
__global__ void __launch_bounds__(KCfg<T, N>::STR_THREADS, KCfg<T, N>::str_min_blocks(M))
    str_tma_kernel(const __grid_constant__ StrTmaParams<T> p) {
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL;
  constexpr int ROWS = N < 256 ? N : 256;  // TMA box rows (boxDim <= 256)
  extern __shared__ __align__(128) unsigned char smem_tma_raw[];
  // [ staging: M * N * W complex | exp_D staging: nplanes * N * W (optional) | exchange: W * M * LS | mbarriers ]
  cpx<T>* stage = reinterpret_cast<cpx<T>*>(smem_tma_raw);
  cpx<T>* stage_d = stage + (size_t)M * N * p.W;
  cpx<T>* exch = stage_d + (p.stage_d ? (size_t)p.nplanes * N * p.W : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(exch + (size_t)p.W * M * p.LS);
  uint64_t* full_d = full + 1;

  const int W = p.W;
  const int xw = threadIdx.x & (W - 1), t = threadIdx.x >> p.logW;
  const long long mstride = (long long)TPL * p.ls;
  cpx<T>* sl = exch + (size_t)xw * M * p.LS;
  const uint32_t tile_bytes = (uint32_t)(M * N * W * sizeof(cpx<T>));

  auto issue = [&](long long tile) {
    const long long xt = tile % p.ntx, o = tile / p.ntx;
    const int o1 = (int)(o % p.no1), o2 = (int)(o / p.no1);
    mbar_expect_tx(full, tile_bytes);
#pragma unroll 1
    for (int c = 0; c < M; ++c)
#pragma unroll 1
      for (int r0 = 0; r0 < N; r0 += ROWS) {
        cpx<T>* dst = stage + ((size_t)c * N + r0) * W;
        if (p.ax == 1)
          tma_load_4d(dst, &p.map[c], full, (int)(2 * xt * W), r0, o1, o2);
        else
          tma_load_4d(dst, &p.map[c], full, (int)(2 * xt * W), o1, r0, o2);
      }
  };

  const bool use_d = p.mode == 1 && p.dkind != KIND_NONE;
  const bool staged_d = use_d && p.stage_d && p.dkind != KIND_SEP;
  auto issue_d = [&](long long tile) {
    const long long xt = tile % p.ntx, o = tile / p.ntx;
    const int o1 = (int)(o % p.no1);
    mbar_expect_tx(full_d, (uint32_t)(p.nplanes * N * W * sizeof(cpx<T>)));
#pragma unroll 1
    for (int pl = 0; pl < p.nplanes; ++pl)
#pragma unroll 1
      for (int r0 = 0; r0 < N; r0 += ROWS) {
        cpx<T>* dst = stage_d + ((size_t)pl * N + r0) * W;
        if (p.ax == 1)
          tma_load_4d(dst, &p.dmap[pl], full_d, (int)(2 * xt * W), r0, o1, 0);
        else
          tma_load_4d(dst, &p.dmap[pl], full_d, (int)(2 * xt * W), o1, r0, 0);
      }
  };

  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(full_d, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && (long long)blockIdx.x < p.ntiles) {
    issue(blockIdx.x);
    if (staged_d) issue_d(blockIdx.x);
  }

  uint32_t parity = 0;
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long xt = tile % p.ntx, o = tile / p.ntx;
    const long long o1 = o % p.no1, o2 = o / p.no1;
    const long long off = xt * W + xw + o1 * p.s1 + o2 * p.s2 + (long long)t * p.ls;
    const long long toff = xt * W + xw + o1 * p.ts1 + (long long)t * p.ls;
    if (use_d && !staged_d && p.dkind != KIND_SEP) {
      for (int pl = 0; pl < p.nplanes; ++pl)
#pragma unroll
        for (int m = 0; m < E; ++m) prefetch_l2(p.D[pl] + toff + m * mstride);
    }

    mbar_wait(full, parity);
    parity ^= 1;
    cpx<T> v[M][E];
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) v[c][m] = stage[((size_t)c * N + t + m * TPL) * W + xw];
    __syncthreads();  // staging consumed: the next tile may land
    if (threadIdx.x == 0 && tile + gridDim.x < p.ntiles) issue(tile + gridDim.x);

    const int it0 = p.mode == 2 ? 1 : 0, it1 = p.mode == 0 ? 0 : 1;
#pragma unroll 1
    for (int it = it0; it <= it1; ++it) {
      fft_fwd_all<T, N, M, SyncBlock>(v, t, sl, p.LS, p.tw, it == 1);
      if (it == 0 && p.mode == 1) {
        if (staged_d) {
          // exp_D of this tile sits in shared memory [plane][row][W]
          mbar_wait(full_d, parity ^ 1);
          const cpx<T>* dpl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) dpl[i] = stage_d + (size_t)i * N * W;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            cpx<T> f[M];
#pragma unroll
            for (int c = 0; c < M; ++c) f[c] = v[c][m];
            disp_point<T, M>(f, dpl, p.dkind, (long long)(t + m * TPL) * W + xw);
#pragma unroll
            for (int c = 0; c < M; ++c) v[c][m] = f[c];
          }
          __syncthreads();  // exp_D staging consumed
          if (threadIdx.x == 0 && tile + gridDim.x < p.ntiles) issue_d(tile + gridDim.x);
        } else if (p.dkind == KIND_SEP) {
          const cpx<T> dperp = p.D[0][toff - (long long)t * p.ls];
          const cpx<T>* dline = p.D[1] + t;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const cpx<T> d = cmul(dperp, dline[m * TPL]);
#pragma unroll
            for (int c = 0; c < M; ++c) v[c][m] = cmul(d, v[c][m]);
          }
        } else {
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
      }
    }
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) p.u[c][off + m * mstride] = v[c][m];
  }
}

// smem bytes needed by str_tma_kernel for a given W
template <typename T, int N>
inline size_t str_tma_smem(int M, int W, int LS, int dplanes) {
  return sizeof(cpx<T>) * ((size_t)M * N * W + (size_t)dplanes * N * W + (size_t)W * M * LS) + 16;
}

template <typename T, int N>
int launch_str_tma(int M, StrTmaParams<T> p, long long nfast, long long ngroups_other, int sm_count, cudaStream_t st);

}  // namespace ggp
