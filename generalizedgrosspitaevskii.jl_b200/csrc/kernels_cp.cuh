// Component-parallel kernels for two-component fields (BASELINE config C3, the exciton-polariton pair).
//
// The two-component variants of row_kernel / str_kernel (kernels.cuh) give one thread BOTH components of its 16 / 8
// points: registers for one component at a time, the other parked in shared memory, the two transforms one after the
// other.  At C3's size (1024^2) those grids are about one wave of CTAs, so a launch lasts as long as ONE CTA's chain
// of dependent phases -- load, inverse FFT, park, load, inverse FFT, half-steps, FFT, store, reload, FFT, store -- and
// the SMs sit at 24 warps (profiles/r02_notes.md: C3 c64 contiguous-axis kernel 57 us against 14.6 us for the
// one-component kernel of the same grid).  Here a line gets TWICE the threads instead: thread group c transforms
// component c, both groups run concurrently, and the only coupling -- the point-wise half-step (G, exp_V and the noise
// amplitude see both components of a point) and a matrix-valued exp_D -- goes through one exchange of the line in
// shared memory.  Same arithmetic per component as the kernels it replaces; half the dependent chain per thread.
#pragma once
#include "kernels.cuh"

// register budgets (per thread) the launch bounds are derived from.  Measured on C3 (A/B r02q, one gpurun call):
// contiguous-axis kernel c128 63.8 us at 64 registers (184 B of spills, 4 CTAs of 256 threads per SM), 59.4 at 80 (3 CTAs);
// c64 36.3 at 64, 43.3 at 80; strided kernel c128 63.5 at 64 (2 CTAs of 512 threads), 71.6 at 80 (one CTA per SM)
#ifndef GGP_CP_ROW_BUDGET64
#define GGP_CP_ROW_BUDGET64 80
#endif
#ifndef GGP_CP_BUDGET
#define GGP_CP_BUDGET 64
#endif
#ifndef GGP_CP_PREFETCH_D
#define GGP_CP_PREFETCH_D 0
#endif
#ifndef GGP_CP_BP64
#define GGP_CP_BP64 2
#endif

namespace ggp {

template <typename T, int N>
struct CpCfg {
  using K = KCfg<T, N>;
  static constexpr int E = K::E, TPL = K::TPL;
  // contiguous-axis kernel: 2 * TPL threads per line, at least 128 threads per CTA
  static constexpr int ROW_TPLINE = 2 * TPL;
  static constexpr int ROW_THREADS = ROW_TPLINE >= 128 ? ROW_TPLINE : 128;
  static constexpr int ROW_LPC = ROW_THREADS / ROW_TPLINE;
  static constexpr int ROW_BUDGET = sizeof(T) == 8 ? GGP_CP_ROW_BUDGET64 : GGP_CP_BUDGET;
  static constexpr int ROW_MINB = 65536 / (ROW_THREADS * ROW_BUDGET) < 1 ? 1 : 65536 / (ROW_THREADS * ROW_BUDGET);
  using RowSync = typename std::conditional<(ROW_TPLINE <= 32), SyncWarp, SyncBlock>::type;
  // strided kernel: W columns x 2 components x TPL threads; 32-byte row pieces where a CTA of <= 512 threads allows
  __host__ __device__ static constexpr int str_w() {
    int w = K::WMIN;
    while (w > 1 && w * 2 * TPL > 512) w >>= 1;
    return w;
  }
  static constexpr int STR_W = str_w();
  static constexpr int STR_THREADS = STR_W * 2 * TPL;
  static constexpr int STR_MINB = 65536 / (STR_THREADS * GGP_CP_BUDGET) < 1 ? 1 : 65536 / (STR_THREADS * GGP_CP_BUDGET);
  static constexpr bool OK = K::USES_SMEM && K::data_regs(1) <= 32 && STR_THREADS <= 1024 && ROW_THREADS <= 1024;
};

// contiguous-axis lines, two components: [inverse FFT_x] -> trailing V/2 of step n -> leading V/2 of step n+1 -> [forward FFT_x]
template <typename T, int N, int PWV>
__global__ void __launch_bounds__(CpCfg<T, N>::ROW_THREADS, CpCfg<T, N>::ROW_MINB) row_cp_kernel(const RowParams<T> p) {
  using C = CpCfg<T, N>;
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL, LS = K::row_ls();
  using SYNC = typename C::RowSync;
  extern __shared__ __align__(16) unsigned char smem_cp_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_cp_raw);

  const int slot = threadIdx.x / C::ROW_TPLINE;
  const int c = (threadIdx.x / TPL) & 1, t = threadIdx.x % TPL;
  const long long lrel = (long long)blockIdx.x * C::ROW_LPC + slot;
  const long long line = lrel + p.line0;
  const bool active = lrel < p.nlines;
  const long long goff = line * N + t;
  const long long soff = (line % p.lines_per_image) * N + t;
  cpx<T>* const mine = smem + (size_t)(slot * 2 + c) * LS;
  cpx<T>* const partner = smem + (size_t)(slot * 2 + (c ^ 1)) * LS;
  cpx<T>* const uc = p.u[c];

  if (p.pdl_pos == 0) pdl_launch_dependents();
  pdl_wait();
  cpx<T> a[E];
#pragma unroll
  for (int m = 0; m < E; ++m) a[m] = active ? uc[goff + m * TPL] : mk<T>((T)0, (T)0);
  if (p.pdl_pos == 1) pdl_launch_dependents();
  if (p.flags & 1) {
    conj_all<T, E>(a);
    fft_line<T, N, -1, SYNC, true>(a, t, mine, p.tw);
    conj_all<T, E>(a);
  }
  if (p.hs[0].apply || p.hs[1].apply) {
    // The half-step couples the components of a point.  Every thread publishes its component (positions t + m*TPL:
    // consecutive lanes, consecutive addresses); of each pair of points (2mm, 2mm+1) group c then takes point 2mm + c:
    // it picks up the other component, applies both half-steps to the pair of components, keeps its own component and
    // hands the other one back through the partner's line.  Per thread: E/2 points x 2 components -- the point-wise
    // work of a line is split between the groups, not duplicated.
    SYNC::sync();  // the last pass of the transforms may still be reading the lines
#pragma unroll
    for (int m = 0; m < E; ++m) mine[t + m * TPL] = a[m];
    SYNC::sync();
    constexpr int BP = sizeof(T) == 8 ? GGP_CP_BP64 : 4;   // points whose pump profiles are fetched together
    constexpr int HP = E / 2;
    bool pre = false;
    if constexpr (PWV != PW_KERR) pre = p.pw.pump && !p.pw.pump_const && !p.pw.pump_dense;
    if (active) {
#pragma unroll
      for (int b0 = 0; b0 < HP; b0 += BP) {
        cpx<T> sp[BP][2];
        if constexpr (PWV != PW_KERR) {
          if (pre) {
#pragma unroll
            for (int b = 0; b < BP; ++b) {
              if (b0 + b < HP) {
                const long long si = soff + (long long)(2 * (b0 + b) + c) * TPL;
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  sp[b][j] = (p.pw.pump == 2 && p.pw.pump_zero[j]) ? mk<T>((T)0, (T)0)
                                                                   : ldg_nc_ordered(p.pw.S[p.pw.pump == 1 ? 0 : j] + si);
              }
            }
          }
        }
#pragma unroll
        for (int b = 0; b < BP; ++b) {
          if (b0 + b < HP) {
            const int m0 = 2 * (b0 + b), m1 = m0 + 1;
            const int pos = t + (m0 + c) * TPL;
            const cpx<T> own = c ? a[m1] : a[m0];
            const cpx<T> o = partner[pos];
            cpx<T> f[2];
            f[0] = c ? o : own;
            f[1] = c ? own : o;
#pragma unroll 1
            for (int h = 0; h < 2; ++h)
              if (p.hs[h].apply)
                half_step_point<T, 2, PWV>(f, p.pw, p.hs[h], soff + (long long)(m0 + c) * TPL, goff + (long long)(m0 + c) * TPL,
                                           nullptr, pre ? sp[b] : nullptr);
            const cpx<T> mineNew = c ? f[1] : f[0];
            partner[pos] = c ? f[0] : f[1];
            if (c) a[m1] = mineNew;
            else a[m0] = mineNew;
          }
        }
      }
    }
    SYNC::sync();
#pragma unroll
    for (int mm = 0; mm < HP; ++mm) {
      // the point of each pair that the partner group updated
      const cpx<T> y = mine[t + (2 * mm + (c ^ 1)) * TPL];
      if (c) a[2 * mm] = y;
      else a[2 * mm + 1] = y;
    }
  }
  if (p.pdl_pos == 2) pdl_launch_dependents();
  if (p.flags & 2) fft_line<T, N, -1, SYNC, true>(a, t, mine, p.tw);  // PRESYNC: everybody has picked up its pair
  if (p.pdl_pos == 3) pdl_launch_dependents();
  if (active) {
#pragma unroll
    for (int m = 0; m < E; ++m) uc[goff + m * TPL] = a[m];
  }
}

// strided-axis lines, two components.  mode 0: forward only, 1: forward -> x exp_D -> inverse, 2: inverse only.
// Virtual column v = c * W + x (x fastest across lanes: the W adjacent columns of component c are one row piece).
template <typename T, int N>
__global__ void __launch_bounds__(CpCfg<T, N>::STR_THREADS, CpCfg<T, N>::STR_MINB) str_cp_kernel(const __grid_constant__ StrParams<T> p) {
  using C = CpCfg<T, N>;
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL, W = C::STR_W, LOGW = ilog2(C::STR_W);
  using Tw = typename TwT<T>::type;
  extern __shared__ __align__(16) unsigned char smem_cps_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_cps_raw);
  const int LS = p.LS;

  const int v = threadIdx.x & (2 * W - 1), xw = v & (W - 1), c = v >> LOGW, t = threadIdx.x >> (LOGW + 1);
  const long long g = blockIdx.x;
  const long long xt = g % p.ntx, o = g / p.ntx;
  const long long o1 = o % p.no1, o2 = o / p.no1;
  const long long off = xt * W + xw + o1 * p.s1 + o2 * p.s2 + (long long)t * p.ls;
  const long long toff = xt * W + xw + o1 * p.ts1 + p.tbase + (long long)t * p.ls;
  const long long mstride = (long long)TPL * p.ls;
  cpx<T>* const mine = smem + (size_t)v * LS;
  const cpx<T>* const other = smem + (size_t)(v ^ W) * LS;
  cpx<T>* const uc = p.u[c];

  // twiddle table into shared memory (constant data: before the wait on the previous kernel)
  const Tw* twp = p.tw;
  const Tw* twc = p.tw + K::TW_COMPACT_OFF;
  if (p.tw_smem) {
    unsigned char* const after_lines = smem_cps_raw + (((size_t)2 * W * LS * sizeof(cpx<T>) + 15) & ~(size_t)15);
    constexpr int NCH1 = (int)((size_t)(K::FACT ? K::TW_SMALL : K::TW_COUNT) * sizeof(Tw) / 16);
    constexpr int NCH = (int)(K::TW_BYTES / 16);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(after_lines);
    const char* g1 = reinterpret_cast<const char*>(p.tw);
    const char* g2 = reinterpret_cast<const char*>(p.tw + K::TW_COMPACT_OFF) - (size_t)NCH1 * 16;
    for (int i = threadIdx.x; i < NCH; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * i),
                   "l"((K::FACT && i >= NCH1 ? g2 : g1) + 16 * (size_t)i));
    asm volatile("cp.async.commit_group;");
    twp = reinterpret_cast<const Tw*>(after_lines);
    twc = twp + K::TW_SMALL;
  }
  const bool sep = p.mode == 1 && p.dkind == KIND_SEP;
  cpx<T> dperp = mk<T>((T)1, (T)0);
  if (sep) dperp = p.D[0][toff - (long long)t * p.ls];
  if (p.pdl_pos == 0) pdl_launch_dependents();
  pdl_wait();

  cpx<T> a[E];
#pragma unroll
  for (int m = 0; m < E; ++m) a[m] = uc[off + m * mstride];
  if (p.tw_smem) asm volatile("cp.async.wait_all;" ::: "memory");  // published by the barriers of the transform
  if (p.pdl_pos == 1) pdl_launch_dependents();

  if (p.mode != 2) fft_line<T, N, -1, SyncBlock, true, K::FACT>(a, t, mine, twp, twc);
  if (p.mode == 1) {
    if (sep) {
      const cpx<T>* dline = p.D[1] + t;
#pragma unroll
      for (int m = 0; m < E; ++m) a[m] = cmul(cmul(dperp, dline[m * TPL]), a[m]);
    } else if (p.dkind == KIND_SCALAR) {
#pragma unroll
      for (int m = 0; m < E; ++m) a[m] = cmul(ldg_nc_ordered(p.D[0] + toff + m * mstride), a[m]);
    } else if (p.dkind == KIND_DIAG) {
#pragma unroll
      for (int m = 0; m < E; ++m)
        a[m] = cmul(p.Daos ? ldg_nc_ordered(p.Daos + (toff + m * mstride) * p.dcols + c) : ldg_nc_ordered(p.D[c] + toff + m * mstride), a[m]);
    } else if (p.dkind == KIND_FULL) {
      // u~_c <- D_c0 u~_0 + D_c1 u~_1: this thread needs the other component of its points (planes [col * 2 + row]).
      // The table entries of the first batch of points are requested BEFORE the exchange, so that their round trip to
      // L2 / DRAM overlaps the two barriers (the kernel is bound by the latency of these loads: 35 % of its warp samples)
      constexpr int DB = GGP_DBATCH < E ? GGP_DBATCH : E;
      auto load_d = [&](cpx<T> (&d)[DB][2], const int m0) {
#pragma unroll
        for (int b = 0; b < DB; ++b)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            d[b][j] = p.Daos ? ldg_nc_ordered(p.Daos + (toff + (m0 + b) * mstride) * p.dcols + (j * 2 + c))
                             : ldg_nc_ordered(p.D[j * 2 + c] + toff + (m0 + b) * mstride);
      };
      cpx<T> d[DB][2];
      if (GGP_CP_PREFETCH_D) load_d(d, 0);
      __syncthreads();  // the last pass may still be reading the lines
#pragma unroll
      for (int m = 0; m < E; ++m) mine[t + m * TPL] = a[m];
      __syncthreads();
#pragma unroll
      for (int m0 = 0; m0 < E; m0 += DB) {
        if (!GGP_CP_PREFETCH_D || m0 > 0) load_d(d, m0);
#pragma unroll
        for (int b = 0; b < DB; ++b) {
          const int m = m0 + b;
          const cpx<T> oth = other[t + m * TPL];
          const cpx<T> f0 = c ? oth : a[m], f1 = c ? a[m] : oth;
          a[m] = cmul(d[b][0], f0) + cmul(d[b][1], f1);
        }
      }
    }
  }
  if (p.pdl_pos == 2) pdl_launch_dependents();
  if (p.mode != 0) {
    conj_all<T, E>(a);
    fft_line<T, N, -1, SyncBlock, true, K::FACT>(a, t, mine, twp, twc);   // PRESYNC: everybody has picked up its pair
    conj_all<T, E>(a);
  }
  if (p.pdl_pos == 3) pdl_launch_dependents();
#pragma unroll
  for (int m = 0; m < E; ++m) uc[off + m * mstride] = a[m];
}

}  // namespace ggp
