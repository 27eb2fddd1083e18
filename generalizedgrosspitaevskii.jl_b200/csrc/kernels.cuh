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
#include "tma.cuh"

#define GGP_MAX_PEERS 8
#ifndef GGP_ROW_PARK
#define GGP_ROW_PARK 1
#endif
#ifndef GGP_PARK_BUDGET
#define GGP_PARK_BUDGET 80
#endif
#ifndef GGP_STOCH_BUDGET
#define GGP_STOCH_BUDGET 128
#endif
#ifndef GGP_TW_UNROLL_E
#define GGP_TW_UNROLL_E 4
#endif
#ifndef GGP_ROW_MIN_THREADS
#define GGP_ROW_MIN_THREADS 128
#endif
#ifndef GGP_ROW_BUDGET
#define GGP_ROW_BUDGET 64
#endif
#ifndef GGP_DBATCH
#define GGP_DBATCH 2
#endif
#ifndef GGP_STR_D2_BLOCKS
#define GGP_STR_D2_BLOCKS 2
#endif

namespace ggp {

template <typename T>
struct RowParams {
  cpx<T>* u[2];
  const typename TwT<T>::type* tw;
  long long nlines;           // (nspatial / N) * nbatch  (or the lines of one chunk of them)
  long long line0;            // first line of this launch (chunked launches of slab plans; 0 otherwise)
  long long lines_per_image;  // nspatial / N ; spatial table line = line % lines_per_image
  PointwiseParams<T> pw;
  HalfStep<T> hs[2];  // trailing half-step of step n, leading half-step of step n+1
  int flags;          // bit 0: inverse FFT_x first, bit 1: forward FFT_x last
  int pdl_pos;        // where this grid lets its dependent launch: 0 at its start, 1 after its loads, 2 after the
                      // half-steps / exp_D multiply, 3 before its final stores (see launch_pdl in inst.cu)
  const void* tw2;    // twiddles of the packed two-line kernels (packed.cuh), fp32 plans only
};

template <typename T>
struct alignas(64) StrParams {
  // TMA staging (tma != 0): the tile [W complex] x [N rows] of every component is fetched by the TMA engine
  // (boxes of <= 256 rows) into the shared-memory lines that afterwards serve as the exchange buffer, and the
  // result leaves the same way.  rank-4 tensor [2*n1 reals, n2, n3, batch]; ax = which dimension the line runs along.
  CUtensorMap map[2];
  int tma, ax, mbar_off;  // mbar_off: byte offset of the mbarrier in dynamic shared memory (set by the launcher)
  cpx<T>* u[2];
  const typename TwT<T>::type* tw;
  long long ls;      // stride (elements) between consecutive points of a line
  long long ntx;     // tiles of W along the fast axis
  long long no1, s1, s2;  // remaining dims: offset = (o % no1) * s1 + (o / no1) * s2
  long long ts1;     // table stride of remaining dim 1 (tables do not have the batch dim)
  long long tbase;   // table offset of the first group of a chunked launch (0 otherwise)
  const cpx<T>* D[4];
  const typename TwT<T>::type* Dsp[2];  // KIND_SEP factors (D_perp, D_line) as hi + lo pairs (fp32 plans)
  int dkind;
  int mode;  // 0 forward only, 1 forward -> x D -> inverse, 2 inverse only
  int W, logW, LS;
  const void* tw2;  // twiddles of the packed two-column kernel (packed.cuh), fp32 plans only
  // Slab decomposition, fused transpose: when scatter != 0 the transformed line is not written back in place
  // but straight into the slabs of the ranks that own it after the all-to-all transpose -- peer memory mapped
  // with CUDA IPC, plain stores over NVLink (own rank: local HBM).  Line point j goes to rank j >> dst_shift,
  // element  dst_base + x + o1*dst_s1 + o2*dst_s2 + (j & mask)*dst_ls  of dst[rank][component].
  cpx<T>* dst[GGP_MAX_PEERS][2];
  long long dst_ls, dst_s1, dst_s2, dst_base;
  int scatter, dst_shift;
  int dl_smem;  // KIND_SEP: D_line is staged in shared memory behind the exchange lines (set by the launcher)
  int pdl_pos;  // as RowParams::pdl_pos
  int slab;     // slab-decomposed plan (LDG loads, peer stores): the launcher picks 128-byte tiles
  int pf_dist;  // > 0: while waiting for its own tile a CTA prefetches tile (blockIdx + pf_dist) into L2 -- set by the
                // launcher to the number of resident CTAs when the state does not fit the L2 (see str_kernel)
  // two-component diagonal / matrix tables, array-of-structs as the host lays them out ([point][entry], entries
  // column-major m11 m21 m12 m22): one tile row reads W * dcols contiguous complex numbers (128 B for W = 2 fp64 matrix
  // entries) instead of a 32-byte piece from each of four planes a full array apart.  The strided kernel of a state
  // that does not fit the L2 next to its table (C3: 32 MB + 64 MB + pump) was DRAM-bound on those pieces at
  // 1.4 TB/s -- every piece its own DRAM page (ncu r02h).  Daos == nullptr: planes D[].
  const cpx<T>* Daos;
  int dcols;
  const double2* Dq[2];  // quirk Q6 (ggp_desc.mixed_precision_tables): Float64 copies of (D_perp, D_line) of a ComplexF32 plan
  int occ1;     // launcher hint: keep this grid at ONE CTA per SM (an NVLink-bound scatter pass of a slab plan that shares
                // the SMs with the next chunk's local kernels on a second stream, see slab_iry in ggp_api.cu)
  int tw_smem;  // twiddle table staged in shared memory although the compile-time default (str_tw_smem) says no:
                // set by the launcher when the geometry chosen at run time (W, CTAs per SM) leaves room for it
};

template <typename T>
struct OneDParams {
  cpx<T>* u[2];
  const typename TwT<T>::type* tw;
  long long nlines;
  PointwiseParams<T> pw;
  const HalfStep<T>* hs;  // 2 * nsteps entries (device)
  int nsteps;
  const cpx<T>* D[4];
  int dkind;
};

// elements per thread of the row kernel where it differs from the default: the stochastic fp64 kernel on short lines is
// bound by fixed-latency chains (Philox rounds, polynomial sincos, DFMA) at 16 warps per SM with radix 8 / 128 registers;
// radix 4 fits 64 registers = 32 warps per SM (C4, 256^2 x 1024 trajectories: row kernel 1.56 -> 1.34 ms).  The strided
// kernel keeps radix 8 (it is near its HBM time there and one more pass costs more than the warps gain: 0.54 -> 0.73 ms).
template <typename T>
__host__ __device__ constexpr int row_E(int N, int M, int pwv) {
#ifdef GGP_NO_ROW_E4
  return 0;
#else
  return (sizeof(T) == 8 && pw_is_stoch(pwv) && M == 1 && N >= 64 && N <= 256) ? 4 : 0;
#endif
}

template <typename T, int N, int EO = 0>
struct KCfg {
  using L = LineCfg<T, N, EO>;
  static constexpr int E = L::E;
  static constexpr int TPL = L::TPL;
  static constexpr int G = 128 / (int)sizeof(cpx<T>);  // lanes served by one shared-memory wavefront
  static constexpr int ROW_THREADS = TPL >= GGP_ROW_MIN_THREADS ? TPL : GGP_ROW_MIN_THREADS;
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
  // strided kernel: W adjacent fast-axis positions per CTA (coalescing width): one 32-byte sector per row
  // (4 complex64 / 2 complex128), halved for the long lines so that a CTA never exceeds 512 threads -- a
  // 1024-thread CTA fills the register file alone and runs its load / transform / store phases in
  // lock-step with nothing to overlap them (4096^2 c64: 219 us per pass before, profiles/r01_notes.md).
  static constexpr int WMIN = sizeof(T) == 4 ? 4 : 2;
  static constexpr int WCAP = (sizeof(T) == 4 && TPL == 512) ? 2 : ((512 / TPL) < 1 ? 1 : (512 / TPL));
  static constexpr int WWANT = (256 / TPL) < WMIN ? WMIN : ((256 / TPL) > 32 ? 32 : (256 / TPL));
  static constexpr int WDEF = WWANT < WCAP ? WWANT : WCAP;
  static constexpr int STR_THREADS = WDEF * TPL;
  // registers: with at most 32 words of field data per thread the kernels fit 64 registers, and the launch
  // bounds ask for as many CTAs per SM as that allows; wider data (M = 2, long fp64 lines) takes what it needs
  __host__ __device__ static constexpr int data_regs(int M) { return M * E * (int)sizeof(cpx<T>) / 4; }
  // Two components in the strided kernel: ONE component in registers at a time, the other parked in its (otherwise
  // idle) exchange line -- as in the row kernel (row_parked_body).  With both components in registers the kernel needs
  // 64 data registers + the exp_D entries: the fp32 instantiation took 242 registers = ONE 8-warp CTA per SM (C3 in
  // ComplexF32: 57 us per launch).  Parked: 32 data registers, ~120 in all, two CTAs per SM, table loads batched
  // (37 us).  (No TMA staging on this path: the dense tile of
  // component 1 would overlap the exchange lines component 0 is transformed in.)
  __host__ __device__ static constexpr bool str_parked(int M) {
#ifdef GGP_NO_STR_PARK
    return false;
#else
    // fp32 only: for fp64 the parked and the two-components-in-registers kernels measured the same (C3: 72.5 vs
    // 71.8 us per launch, A/B r02j), and the latter keeps its TMA staging
    return M == 2 && USES_SMEM && sizeof(T) == 4 && data_regs(1) <= 32;
#endif
  }
  __host__ __device__ static constexpr int str_min_blocks(int M) {
    if (str_parked(M)) return 1024 / STR_THREADS >= 2 ? 2 : 1;
    // two-component fp64 tiles (64 data registers, 256 threads): left alone the compiler takes 206 registers = ONE
    // 8-warp CTA per SM (C3, ncu r01t: issue-active 11.5 %); capped at 128 two CTAs fit
    if (sizeof(T) == 8 && data_regs(M) == 64 && STR_THREADS <= 256) return GGP_STR_D2_BLOCKS;
    return data_regs(M) <= 32 ? 1024 / STR_THREADS : 1;
  }
  // Tuning headroom: the fp32 kernels that fit 64 registers may be launched with twice the tile width (1024
  // threads, one CTA per SM; GGP_STR_WIDE=1).  Not the default: at 4096^2 the wide tile is slower (225 vs 208 us)
  // and the DRAM traffic is the same 2x of the algorithmic bytes either way (ncu, profiles/r01_notes.md session 4).
  __host__ __device__ static constexpr int str_max_threads(int M) {
    return (data_regs(M) <= 32 && STR_THREADS < 1024) ? 1024 : STR_THREADS;
  }
  __host__ __device__ static constexpr int str_bound_blocks(int M) {  // min CTAs per SM stated in the launch bounds
    return str_max_threads(M) > STR_THREADS ? 1 : str_min_blocks(M);
  }
  // strided kernel: the twiddle table lives in shared memory (copied with cp.async at kernel start) whenever
  // that does not cost a resident CTA.  Read from global memory at the point of use -- the 64-register budget
  // leaves no room to batch the loads -- the ~60 twiddle loads per thread were serialised L1/L2 round trips and
  // the critical path of the CTA (ncu r01n, 4096^2: 65 % of the stall samples long_scoreboard on them).
  static constexpr int TW_COUNT = twiddle_count<T, N>();
  // long lines: the strided kernel forms the late passes' twiddles from the compact tables (fft_line.cuh, FACT) and
  // stages only the early passes' blocks + the compact tables: 3 KB instead of 33 KB for N = 4096
  static constexpr bool FACT = tw_factorized<T, N>();
  static constexpr int TW_SMALL = twiddle_small_count<T, N>();
  static constexpr int TW_COMPACT = twiddle_compact_count<T, N>();
  static constexpr int TW_COMPACT_OFF = (TW_COUNT + 1) & ~1;  // global table: [pass blocks | pad | compact tables]
  static constexpr int TW_STAGED = FACT ? TW_SMALL + TW_COMPACT : TW_COUNT;
  static_assert(!FACT || TW_SMALL % 2 == 0, "16-byte staging chunks");
  static constexpr size_t TW_BYTES = ((size_t)TW_STAGED * sizeof(typename TwT<T>::type) + 15) & ~(size_t)15;
  __host__ __device__ static constexpr size_t str_lines_bytes(int M, int W, int LS) {
    return (((size_t)W * M * LS * sizeof(cpx<T>)) + 15) & ~(size_t)15;
  }
  __host__ __device__ static constexpr int str_ls_c(int W) {  // constexpr twin of str_ls
    int lpg = W >= G ? 1 : G / W;
    int ls = L::PADN;
    while (ls % G != lpg % G) ++ls;
    return ls;
  }
  // (only together with room for a staged D_line: with the twiddles alone in shared memory the L1 that is left
  //  cannot hold D_line any more and its loads become the serialised round trips -- 4096^2 c64: 211 -> 248 us)
  __host__ __device__ static constexpr bool str_tw_smem(int M) {
    return USES_SMEM && TW_COUNT > 0 &&
           (str_lines_bytes(M, WDEF, str_ls_c(WDEF)) + TW_BYTES + (size_t)N * sizeof(cpx<T>) + 1024) *
                   (size_t)str_min_blocks(M) <= (size_t)226 * 1024;
  }
  // The register budget is stated explicitly for every variant: left to itself with "at least one CTA" the
  // compiler spends 144 registers on the stochastic fp64 kernel (3 CTAs per SM instead of 4 at 128) and 190 on
  // the two-component fp64 one (2 instead of 3 at 168) -- measured 17 % slower on C4 and C3.
  // Two components, deterministic half-steps: the row kernel keeps ONE component in registers at a time and parks
  // the other in its (otherwise idle) exchange line -- half the registers, twice the resident CTAs, and one copy
  // of the transform code instead of two (C3, ncu r01t: 168 registers = 12 warps per SM, 18 % of the stall samples
  // `no_instructions`).  The stochastic variant keeps both components in registers (Philox words per element).
  __host__ __device__ static constexpr bool row_parked(int M, int pwv) {
    return GGP_ROW_PARK && M == 2 && !pw_is_stoch(pwv) && USES_SMEM && data_regs(1) <= 32;
  }
  __host__ __device__ static constexpr int row_min_blocks(int M, int pwv) {
    if (row_parked(M, pwv)) {  // one component in registers at a time (row_parked_body)
      const int bp = 65536 / (ROW_THREADS * GGP_PARK_BUDGET);
      return bp < 1 ? 1 : bp;
    }
    const int dr = data_regs(M);
    const int budget = dr <= 32 ? ((pw_is_stoch(pwv) && EO == 0) ? GGP_STOCH_BUDGET : GGP_ROW_BUDGET) : ((dr <= 64 && sizeof(T) == 8) ? (pw_is_stoch(pwv) ? 255 : 168) : 255);
    const int b = 65536 / (ROW_THREADS * budget);
    return b < 1 ? 1 : b;
  }
  using RowSync = typename std::conditional<(TPL <= 32), SyncWarp, SyncBlock>::type;
};

template <typename T, int E>
__device__ __forceinline__ void conj_all(cpx<T> (&v)[E]) {
#pragma unroll
  for (int m = 0; m < E; ++m) v[m].y = -v[m].y;
}

// The forward transform is the only FFT code in a kernel; the inverse runs the same instructions
// on conjugated data (ifft(x) = conj(fft(conj(x)))), selected by a loop that is NOT unrolled.
// This halves the instruction footprint -- the first version of these kernels was bound by
// instruction-cache misses (ncu: stall_no_instructions dominant, profiles/r01_notes.md).
template <typename T, int N, int M, typename SYNC, bool FACT = false, int EO = 0>
__device__ __forceinline__ void fft_fwd_all(cpx<T> (&v)[M][LineCfg<T, N, EO>::E], const int t, cpx<T>* sl, const int LS,
                                            const typename TwT<T>::type* __restrict__ tw, const bool inverse,
                                            const typename TwT<T>::type* __restrict__ twc = nullptr) {
#pragma unroll
  for (int c = 0; c < M; ++c) {
    if (inverse) conj_all<T, LineCfg<T, N, EO>::E>(v[c]);
    fft_line<T, N, -1, SYNC, true, FACT, EO>(v[c], t, sl + c * LS, tw, twc);
    if (inverse) conj_all<T, LineCfg<T, N, EO>::E>(v[c]);
  }
}

template <typename T, int N, int M, int PWV, int EO = 0>
__device__ __forceinline__ void half_steps(cpx<T> (&v)[M][LineCfg<T, N, EO>::E], const PointwiseParams<T>& pw,
                                           const HalfStep<T>* hs, const int nh, const long long sidx0,
                                           const long long gidx0, const long long stride) {
  constexpr int E = LineCfg<T, N, EO>::E;
  if constexpr (pw_is_stoch(PWV)) {
    const HalfStep<T>& href = hs[0].apply ? hs[0] : hs[nh - 1];
    // pair index = (64-bit half-step counter + 1) >> 1
    const unsigned long long pair = ((((unsigned long long)href.ctr_hi << 32) | href.ctr) + 1ull) >> 1;
    if constexpr (PWV == PW_TW && E <= GGP_TW_UNROLL_E) {
      // both half-steps of the pair in line, element by element; the Philox words of an element are produced right
      // before they are used and dropped afterwards (held for all elements they were spilled: ncu r02s)
#pragma unroll
      for (int m = 0; m < E; ++m) {
        uint4 rnd[M];
        cpx<T> f[M];
#pragma unroll
        for (int c = 0; c < M; ++c) {
          rnd[c] = philox_for<T>(gidx0 + m * stride + pw.elem_offset, (uint32_t)pair, (uint32_t)(pair >> 32), c, pw.seed_lo,
                                 pw.seed_hi);
          f[c] = v[c][m];
        }
        if (hs[0].apply) half_step_point<T, M, PWV>(f, pw, hs[0], sidx0 + m * stride, gidx0 + m * stride, rnd);
        if (nh > 1 && hs[1].apply) half_step_point<T, M, PWV>(f, pw, hs[1], sidx0 + m * stride, gidx0 + m * stride, rnd);
#pragma unroll
        for (int c = 0; c < M; ++c) v[c][m] = f[c];
      }
      return;
    }
    // All Philox words of this thread first (E*M independent chains: good ILP), one call per element and
    // component serving both half-steps of the pair; then the half-steps as in the deterministic variant.
    uint4 rnd[E][M];
    if (PWV == PW_TW || pw.noise == NOISE_PHILOX) {
#pragma unroll
      for (int m = 0; m < E; ++m)
#pragma unroll
        for (int c = 0; c < M; ++c)
          rnd[m][c] = philox_for<T>(gidx0 + m * stride + pw.elem_offset, (uint32_t)pair, (uint32_t)(pair >> 32), c,
                                    pw.seed_lo, pw.seed_hi);
    }
    {
#pragma unroll 1
      for (int h = 0; h < nh; ++h) {
        if (!hs[h].apply) continue;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          cpx<T> f[M];
#pragma unroll
          for (int c = 0; c < M; ++c) f[c] = v[c][m];
          half_step_point<T, M, PWV>(f, pw, hs[h], sidx0 + m * stride, gidx0 + m * stride, rnd[m]);
#pragma unroll
          for (int c = 0; c < M; ++c) v[c][m] = f[c];
        }
      }
    }
  } else if constexpr (PWV == PW_KERR) {
    // Real diagonal nonlinearity only: u_i <- cis(-dt G_i(|u|^2)) u_i is a pure phase, |u_j| is invariant under
    // it, so the trailing half-step of step n and the leading half-step of step n+1 (same G, same |u|) are
    // exactly ONE rotation by the summed angle -- half the sincos work of the fused kernel.
    int napply = 0;
    for (int h = 0; h < nh; ++h) napply += hs[h].apply ? 1 : 0;
    if (napply) {
      const T dts = pw.dt * (T)napply;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        T n2[M];
#pragma unroll
        for (int j = 0; j < M; ++j) n2[j] = cabs2(v[j][m]);
#pragma unroll
        for (int i = 0; i < M; ++i) {
          T gre = pw.nl_c_re[i];
#pragma unroll
          for (int j = 0; j < M; ++j) gre += pw.nl_g_re[i][j] * n2[j];
          T sn, cm1;
          sincosm1_t(-dts * gre, &sn, &cm1);
          v[i][m] = rotate_m1(v[i][m], cm1, sn);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int h = 0; h < nh; ++h) {
      if (!hs[h].apply) continue;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        cpx<T> f[M];
#pragma unroll
        for (int c = 0; c < M; ++c) f[c] = v[c][m];
        half_step_point<T, M, PWV>(f, pw, hs[h], sidx0 + m * stride, gidx0 + m * stride, nullptr);
#pragma unroll
        for (int c = 0; c < M; ++c) v[c][m] = f[c];
      }
    }
  }
}

// Row kernel body for two components with one component in registers at a time (KCfg::row_parked).
//   phase 1: component 0: load, inverse FFT_x, park in its own exchange line (every thread parks and later reloads
//            ITS OWN elements, positions t + m*TPL); component 1: load, inverse FFT_x, stays in registers
//   point-wise: both half-steps element by element, component 0 read from / written back to the parked line
//   phase 2: component 1: forward FFT_x, store; component 0: reload, forward FFT_x, store
// The component loops are NOT unrolled: one copy of the transform code serves both.
template <typename T, int N, int PWV>
__device__ __forceinline__ void row_parked_body(const RowParams<T>& p, cpx<T>* sl, const int t, const bool active,
                                                const long long goff, const long long soff) {
  using K = KCfg<T, N>;
  using L = LineCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL, LS = K::row_ls();
  using SYNC = typename K::RowSync;
  cpx<T> a[E];
  cpx<T>* const park = sl + L::pad(t);                  // pad(t + m*TPL) = pad(t) + m*(TPL + TPL/E) when TPL % E == 0
  constexpr bool LIN = (TPL % E == 0);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int m = 0; m < E; ++m) a[m] = active ? p.u[c][goff + m * TPL] : mk<T>((T)0, (T)0);
    if (p.flags & 1) {
      conj_all<T, E>(a);
      fft_line<T, N, -1, SYNC, true>(a, t, sl + c * LS, p.tw);
      conj_all<T, E>(a);
    }
    if (c == 0) {
      SYNC::sync();  // the last pass of the transform may still be reading this line
#pragma unroll
      for (int m = 0; m < E; ++m) {
        if constexpr (LIN) park[m * (TPL + TPL / E)] = a[m];
        else sl[L::pad(t + m * TPL)] = a[m];
      }
    }
  }
  if (p.pdl_pos == 1) pdl_launch_dependents();
  if (active) {
#pragma unroll
    for (int m = 0; m < E; ++m) {
      cpx<T> f[2];
      if constexpr (LIN) f[0] = park[m * (TPL + TPL / E)];
      else f[0] = sl[L::pad(t + m * TPL)];
      f[1] = a[m];
#pragma unroll 1
      for (int h = 0; h < 2; ++h)
        if (p.hs[h].apply) half_step_point<T, 2, PWV>(f, p.pw, p.hs[h], soff + m * TPL, goff + m * TPL, nullptr);
      if constexpr (LIN) park[m * (TPL + TPL / E)] = f[0];
      else sl[L::pad(t + m * TPL)] = f[0];
      a[m] = f[1];
    }
  }
  if (p.pdl_pos >= 2) pdl_launch_dependents();
#pragma unroll 1
  for (int c = 1; c >= 0; --c) {
    if (c == 0) {
#pragma unroll
      for (int m = 0; m < E; ++m) {
        if constexpr (LIN) a[m] = park[m * (TPL + TPL / E)];
        else a[m] = sl[L::pad(t + m * TPL)];
      }
    }
    if (p.flags & 2) fft_line<T, N, -1, SYNC, true>(a, t, sl + c * LS, p.tw);  // PRESYNC: everybody has reloaded
    if (active) {
#pragma unroll
      for (int m = 0; m < E; ++m) p.u[c][goff + m * TPL] = a[m];
    }
  }
}

// flags: bit 0 = PRE (inverse FFT_x before the half-steps), bit 1 = POST (forward FFT_x after)
template <typename T, int N, int M, int PWV>
__global__ void __launch_bounds__(KCfg<T, N, row_E<T>(N, M, PWV)>::ROW_THREADS, KCfg<T, N, row_E<T>(N, M, PWV)>::row_min_blocks(M, PWV))
    row_kernel(const RowParams<T> p) {
  constexpr int EO = row_E<T>(N, M, PWV);
  using K = KCfg<T, N, EO>;
  constexpr int E = K::E, TPL = K::TPL, LPC = K::LPC, LS = K::row_ls();
  using SYNC = typename K::RowSync;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_raw);

  const int grp = threadIdx.x / TPL, t = threadIdx.x % TPL;
  const long long lrel = (long long)blockIdx.x * LPC + grp;
  const long long line = lrel + p.line0;
  const bool active = lrel < p.nlines;
  const long long goff = line * N + t;
  const long long soff = (line % p.lines_per_image) * N + t;
  cpx<T>* sl = smem + (size_t)grp * M * LS;

  if (p.pdl_pos == 0) pdl_launch_dependents();
  pdl_wait();
  if constexpr (K::row_parked(M, PWV)) {
    row_parked_body<T, N, PWV>(p, sl, t, active, goff, soff);
  } else {
    cpx<T> v[M][E];
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) v[c][m] = active ? p.u[c][goff + m * TPL] : mk<T>((T)0, (T)0);
    if (p.pdl_pos == 1) pdl_launch_dependents();

#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      if (p.flags & (1 << it)) fft_fwd_all<T, N, M, SYNC, false, EO>(v, t, sl, LS, p.tw, it == 0);
      if (it == 0 && active) half_steps<T, N, M, PWV, EO>(v, p.pw, p.hs, 2, soff, goff, TPL);
      if (it == 0 && p.pdl_pos == 2) pdl_launch_dependents();
    }
    if (p.pdl_pos == 3) pdl_launch_dependents();
    if (active) {
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int m = 0; m < E; ++m) p.u[c][goff + m * TPL] = v[c][m];
    }
  }
}

// mode 0: forward only, 1: forward -> x D -> inverse, 2: inverse only
// WT > 0: the tile width is a compile-time constant (the default geometry KCfg::WDEF): shared-memory addresses of
// the tile pick-up / write-back become one base register + immediate offsets instead of a shift and two adds per
// element (ncu r01x, 2048^2: 64 of those per thread and direction).  WT = 0: width from the parameters.
template <typename T, int N, int M, int WT = 0>
__global__ void __launch_bounds__(KCfg<T, N>::str_max_threads(M), KCfg<T, N>::str_bound_blocks(M))
    str_kernel(const __grid_constant__ StrParams<T> p) {
  using K = KCfg<T, N>;
  constexpr int E = K::E, TPL = K::TPL;
  const int W = WT ? WT : p.W, logW = WT ? ilog2(WT) : p.logW, LS = WT ? K::str_ls_c(WT ? WT : 1) : p.LS;
  constexpr int ROWS = N < 256 ? N : 256;  // rows per TMA box (boxDim <= 256)
  extern __shared__ __align__(128) unsigned char smem_str_raw[];
  unsigned char* const smem_raw = smem_str_raw;
  cpx<T>* smem = reinterpret_cast<cpx<T>*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + p.mbar_off);
  if (p.tma) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  const int xw = threadIdx.x & (W - 1), t = threadIdx.x >> logW;
  const long long g = blockIdx.x;
  const long long xt = g % p.ntx, o = g / p.ntx;
  const long long o1 = o % p.no1, o2 = o / p.no1;
  const long long off = xt * W + xw + o1 * p.s1 + o2 * p.s2 + (long long)t * p.ls;
  const long long toff = xt * W + xw + o1 * p.ts1 + p.tbase + (long long)t * p.ls;
  const long long mstride = (long long)TPL * p.ls;
  cpx<T>* sl = smem + (size_t)xw * M * LS;

  // Separable exp_D: the per-column factor is fetched now and D_line is copied into shared memory with
  // cp.async while the field loads and the forward transform run -- both are constant tables, so this happens
  // BEFORE the wait on the previous kernel.  (Read at the point of use, the 16 D_line loads per thread were
  // the largest stall of this kernel: ncu r01m, 4096^2: long_scoreboard 41 % of the samples.)
  // shared memory: [exchange lines | twiddle table (TW_SMEM) | D_line (dl_smem) | mbarrier (tma)]
  constexpr bool TW_SMEM = K::str_tw_smem(M);
  using Tw = typename TwT<T>::type;
  unsigned char* const after_lines = smem_raw + K::str_lines_bytes(M, W, LS);
  const Tw* twp = p.tw;
  const Tw* twc = p.tw + K::TW_COMPACT_OFF;
  const bool tws = K::TW_STAGED > 0 && (TW_SMEM || p.tw_smem);
  if (tws) {
    // FACT: [early passes' blocks | compact tables] are two ranges of the global table, contiguous in shared memory
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
  cpx<T>* sdl = reinterpret_cast<cpx<T>*>(after_lines + (tws ? K::TW_BYTES : 0));
  cpx<T> dperp = mk<T>((T)1, (T)0);
  if (sep) {
    if constexpr (!TwT<T>::split) {
      if (p.dl_smem) {
        constexpr int NCHUNK = N * (int)sizeof(cpx<T>) / 16;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sdl);
        for (int i = threadIdx.x; i < NCHUNK; i += blockDim.x)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * i),
                       "l"(reinterpret_cast<const char*>(p.D[1]) + 16 * (size_t)i));
        asm volatile("cp.async.commit_group;");
      }
      dperp = p.D[0][toff - (long long)t * p.ls];
    }
  }
  if (p.pdl_pos == 0) pdl_launch_dependents();
  pdl_wait();
  if constexpr (K::str_parked(M)) {
    // ---- two components, one in registers at a time (see KCfg::str_parked) ----
    cpx<T> a[E];
    const bool sepq = sep;
    auto store = [&](const int c) {
      if (p.scatter) {
        const long long dbase = p.dst_base + xt * W + xw + o1 * p.dst_s1 + o2 * p.dst_s2;
        const int mask = (1 << p.dst_shift) - 1;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int j = t + m * TPL;
          p.dst[j >> p.dst_shift][c][dbase + (long long)(j & mask) * p.dst_ls] = a[m];
        }
      } else {
#pragma unroll
        for (int m = 0; m < E; ++m) p.u[c][off + m * mstride] = a[m];
      }
    };
    if (tws || (sepq && p.dl_smem)) asm volatile("cp.async.wait_all;" ::: "memory");  // published by the transform's barriers
    if (p.mode != 1) {
      // forward-only / inverse-only (middle axis of a 3-D grid): the components are independent
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int m = 0; m < E; ++m) a[m] = p.u[c][off + m * mstride];
        if (p.mode == 2) conj_all<T, E>(a);
        fft_line<T, N, -1, SyncBlock, true, K::FACT>(a, t, sl + c * LS, twp, twc);
        if (p.mode == 2) conj_all<T, E>(a);
        if (c == 0 && p.pdl_pos == 2) pdl_launch_dependents();
        if (c == 1 && p.pdl_pos == 3) pdl_launch_dependents();
        store(c);
      }
      return;
    }
    using L = LineCfg<T, N>;
    constexpr bool LIN = (TPL % E == 0);               // pad(t + m*TPL) = pad(t) + m*(TPL + TPL/E)
    cpx<T>* const park = sl + L::pad(t);               // component 0's line; every thread parks and reloads ITS OWN elements
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int m = 0; m < E; ++m) a[m] = p.u[c][off + m * mstride];
      fft_line<T, N, -1, SyncBlock, true, K::FACT>(a, t, sl + c * LS, twp, twc);
      if (c == 0) {
        __syncthreads();                               // the last pass may still be reading this line
#pragma unroll
        for (int m = 0; m < E; ++m) {
          if constexpr (LIN) park[m * (TPL + TPL / E)] = a[m];
          else sl[L::pad(t + m * TPL)] = a[m];
        }
      }
    }
    if (p.pdl_pos == 1 || p.pdl_pos == 2) pdl_launch_dependents();
    // k-space multiply  u~ <- exp_D[k] (x) u~  on both components of a point
    if (sepq) {
      const cpx<T>* dline = p.dl_smem ? (sdl + t) : (p.D[1] + t);
#pragma unroll
      for (int m = 0; m < E; ++m) {
        cpx<T> f0;
        if constexpr (LIN) f0 = park[m * (TPL + TPL / E)];
        else f0 = sl[L::pad(t + m * TPL)];
        const cpx<T> d = cmul(dperp, dline[m * TPL]);
        f0 = cmul(d, f0);
        a[m] = cmul(d, a[m]);
        if constexpr (LIN) park[m * (TPL + TPL / E)] = f0;
        else sl[L::pad(t + m * TPL)] = f0;
      }
    } else {
      // full tables: the entries of DB elements are requested together (ldg_nc_ordered) -- left to the scheduler they
      // are loaded one element at a time, E dependent round trips to L2 / DRAM per thread (ncu r02h)
      constexpr int DB = GGP_DBATCH < E ? GGP_DBATCH : E;
      const int npl = p.dkind == KIND_SCALAR ? 1 : (p.dkind == KIND_DIAG ? 2 : 4);
#pragma unroll
      for (int m0 = 0; m0 < E; m0 += DB) {
        cpx<T> d[DB][4];
#pragma unroll
        for (int b = 0; b < DB; ++b)
#pragma unroll
          for (int pl = 0; pl < 4; ++pl)
            if (pl < npl)
              d[b][pl] = p.Daos ? ldg_nc_ordered(p.Daos + (toff + (m0 + b) * mstride) * p.dcols + pl)
                                : ldg_nc_ordered(p.D[pl] + toff + (m0 + b) * mstride);
#pragma unroll
        for (int b = 0; b < DB; ++b) {
          const int m = m0 + b;
          cpx<T> f0, f1 = a[m];
          if constexpr (LIN) f0 = park[m * (TPL + TPL / E)];
          else f0 = sl[L::pad(t + m * TPL)];
          if (p.dkind == KIND_FULL) {               // planes [col * 2 + row]
            const cpx<T> r0 = cmul(d[b][0], f0) + cmul(d[b][2], f1);
            const cpx<T> r1 = cmul(d[b][1], f0) + cmul(d[b][3], f1);
            f0 = r0;
            f1 = r1;
          } else if (p.dkind == KIND_DIAG) {
            f0 = cmul(d[b][0], f0);
            f1 = cmul(d[b][1], f1);
          } else {
            f0 = cmul(d[b][0], f0);
            f1 = cmul(d[b][0], f1);
          }
          if constexpr (LIN) park[m * (TPL + TPL / E)] = f0;
          else sl[L::pad(t + m * TPL)] = f0;
          a[m] = f1;
        }
      }
    }
#pragma unroll 1
    for (int c = 1; c >= 0; --c) {
      if (c == 0) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          if constexpr (LIN) a[m] = park[m * (TPL + TPL / E)];
          else a[m] = sl[L::pad(t + m * TPL)];
        }
      }
      conj_all<T, E>(a);
      fft_line<T, N, -1, SyncBlock, true, K::FACT>(a, t, sl + c * LS, twp, twc);   // PRESYNC: everybody has reloaded
      conj_all<T, E>(a);
      if (c == 0 && p.pdl_pos == 3) pdl_launch_dependents();
      store(c);
    }
    return;
  }
  cpx<T> v[M][E];
  if (p.tma) {
    // one thread asks the TMA engine for the whole tile, dense [component][row][W] at the start of the
    // exchange lines; everybody picks its elements up with conflict-free LDS -- the load/store unit sees 2
    // wavefronts per 32 elements instead of one per 32-byte row piece (8 with W = 4, 16 with W = 2)
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, (uint32_t)(M * N * W * sizeof(cpx<T>)));
#pragma unroll 1
      for (int c = 0; c < M; ++c)
#pragma unroll 1
        for (int r0 = 0; r0 < N; r0 += ROWS)
          tma_load_4d(smem + ((size_t)c * N + r0) * W, &p.map[c], bar, (int)(2 * xt * W), p.ax == 1 ? r0 : (int)o1,
                      p.ax == 1 ? (int)o1 : r0, (int)o2);
    }
    // States larger than the L2: the tile that the CTA taking over this SM slot will want (blockIdx + number of
    // resident CTAs) is pulled into L2 now, while this CTA has nothing to do but wait for its own tile -- that
    // later load then streams from L2 instead of paying one DRAM page miss per 16/32-byte row piece
    // (ncu r01u, 4096^2: 26 % of the stall samples sit on the tile's mbarrier).
    if (p.pf_dist) {
      const long long g2 = g + p.pf_dist;
      if (g2 < (long long)gridDim.x) {
        const long long xt2 = g2 % p.ntx, oo = g2 / p.ntx;
        const long long base = xt2 * W + (oo % p.no1) * p.s1 + (oo / p.no1) * p.s2;
#pragma unroll 1
        for (int c = 0; c < M; ++c)
#pragma unroll 1
          for (int r = threadIdx.x; r < N; r += blockDim.x) prefetch_l2(p.u[c] + base + (long long)r * p.ls);
      }
    }
    mbar_wait(bar, 0);
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) v[c][m] = smem[(((size_t)c * N + t + m * TPL) << logW) + xw];
  } else {
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) v[c][m] = p.u[c][off + m * mstride];
  }
  if (tws || (sep && p.dl_smem)) asm volatile("cp.async.wait_all;" ::: "memory");  // published by the barriers of the transform

  if (p.pdl_pos == 1) pdl_launch_dependents();
  const int it0 = p.mode == 2 ? 1 : 0, it1 = p.mode == 0 ? 0 : 1;
#pragma unroll 1
  for (int it = it0; it <= it1; ++it) {
    fft_fwd_all<T, N, M, SyncBlock, K::FACT>(v, t, sl, LS, twp, it == 1, twc);
    if (it == it0 && p.pdl_pos == 2) pdl_launch_dependents();
    if (it == 0 && p.mode == 1) {
      if (p.dkind == KIND_SEP) {
        if constexpr (TwT<T>::split) {
          // fp32 with -DGGP_SPLIT_TWIDDLES: both factors carry their rounding residual (hi + lo) and are applied
          // one after the other with compensated products (see TwT in cplx.cuh)
          const typename TwT<T>::type dp = p.Dsp[0][toff - (long long)t * p.ls];
          const typename TwT<T>::type* dline = p.Dsp[1] + t;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const typename TwT<T>::type dl = dline[m * TPL];
#pragma unroll
            for (int c = 0; c < M; ++c) v[c][m] = TwT<T>::mul(TwT<T>::mul(v[c][m], dl), dp);
          }
        } else if (sizeof(T) == 4 && p.Dq[0]) {
          // quirk Q6: the reference holds a ComplexF64 table for this ComplexF32 problem and forms exp_D[k] * u~[k] in
          // ComplexF64 before the store rounds it (src/misc.jl:14-17, src/kernels.jl:44-53): same here, on the
          // otherwise idle fp64 pipe
          const double2 dp = p.Dq[0][toff - (long long)t * p.ls];
          const double2* dline = p.Dq[1] + t;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const double2 dl = dline[m * TPL];
            const double dre = dp.x * dl.x - dp.y * dl.y, dim = dp.x * dl.y + dp.y * dl.x;
#pragma unroll
            for (int c = 0; c < M; ++c) {
              const double vx = (double)v[c][m].x, vy = (double)v[c][m].y;
              v[c][m] = mk<T>((T)(dre * vx - dim * vy), (T)(dre * vy + dim * vx));
            }
          }
        } else if (p.dl_smem) {
          const cpx<T>* dline = sdl + t;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const cpx<T> d = cmul(dperp, dline[m * TPL]);
#pragma unroll
            for (int c = 0; c < M; ++c) v[c][m] = cmul(d, v[c][m]);
          }
        } else {
          const cpx<T>* dline = p.D[1] + t;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const cpx<T> d = cmul(dperp, dline[m * TPL]);
#pragma unroll
            for (int c = 0; c < M; ++c) v[c][m] = cmul(d, v[c][m]);
          }
        }
      } else {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          cpx<T> f[M];
#pragma unroll
          for (int c = 0; c < M; ++c) f[c] = v[c][m];
          if (M == 2 && p.Daos) {
            const cpx<T>* q = p.Daos + (toff + m * mstride) * p.dcols;
            const cpx<T>* planes[4] = {q, q + 1, q + 2, q + 3};
            disp_point<T, M>(f, planes, p.dkind, 0);
          } else {
            disp_point<T, M>(f, p.D, p.dkind, toff + m * mstride);
          }
#pragma unroll
          for (int c = 0; c < M; ++c) v[c][m] = f[c];
        }
      }
    }
  }
  if (p.pdl_pos == 3) pdl_launch_dependents();
  if (p.scatter) {
    const long long dbase = p.dst_base + xt * W + xw + o1 * p.dst_s1 + o2 * p.dst_s2;
    const int mask = (1 << p.dst_shift) - 1;
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int j = t + m * TPL;
        p.dst[j >> p.dst_shift][c][dbase + (long long)(j & mask) * p.dst_ls] = v[c][m];
      }
    return;
  }
  if (p.tma) {
    __syncthreads();  // the last pass has read its inputs from the exchange lines
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int m = 0; m < E; ++m) smem[(((size_t)c * N + t + m * TPL) << logW) + xw] = v[c][m];
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 1
      for (int c = 0; c < M; ++c)
#pragma unroll 1
        for (int r0 = 0; r0 < N; r0 += ROWS)
          tma_store_4d(&p.map[c], smem + ((size_t)c * N + r0) * W, (int)(2 * xt * W), p.ax == 1 ? r0 : (int)o1,
                       p.ax == 1 ? (int)o1 : r0, (int)o2);
      tma_store_commit_and_wait_read();
    }
    return;
  }
#pragma unroll
  for (int c = 0; c < M; ++c)
#pragma unroll
    for (int m = 0; m < E; ++m) p.u[c][off + m * mstride] = v[c][m];
}

template <typename T, int N, int M, int PWV>
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

  pdl_launch_dependents();
  pdl_wait();
  cpx<T> v[M][E];
#pragma unroll
  for (int c = 0; c < M; ++c)
#pragma unroll
    for (int m = 0; m < E; ++m) v[c][m] = active ? p.u[c][goff + m * TPL] : mk<T>((T)0, (T)0);

  // half-step 0, then per step: [FFT x D x inverse FFT], then the trailing half-step of this step
  // together with the leading half-step of the next one (they share one Philox call per element)
  if (active) half_steps<T, N, M, PWV>(v, p.pw, p.hs, 1, t, goff, TPL);
#pragma unroll 1
  for (int st = 0; st < p.nsteps; ++st) {
    if (p.dkind != KIND_NONE) {
#pragma unroll 1
      for (int it = 0; it < 2; ++it) {
        fft_fwd_all<T, N, M, SYNC>(v, t, sl, LS, p.tw, it == 1);
        if (it == 0) {
#pragma unroll
          for (int m = 0; m < E; ++m) {
            cpx<T> f[M];
#pragma unroll
            for (int c = 0; c < M; ++c) f[c] = v[c][m];
            disp_point<T, M>(f, p.D, p.dkind, t + m * TPL);
#pragma unroll
            for (int c = 0; c < M; ++c) v[c][m] = f[c];
          }
        }
      }
    }
    if (active) half_steps<T, N, M, PWV>(v, p.pw, p.hs + 2 * st + 1, st + 1 < p.nsteps ? 2 : 1, t, goff, TPL);
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
int launch_row(int M, int pwv, const RowParams<T>& p, cudaStream_t st);
template <typename T, int N>
int launch_str(int M, StrParams<T> p, long long nfast, long long ngroups_other, cudaStream_t st);
template <typename T, int N>
int launch_oned(int M, int pwv, const OneDParams<T>& p, cudaStream_t st);
// geometry of the strided kernels for a fast axis of nfast points: W (coalescing width), padded
// line stride LS, threads per CTA, and whether the line needs the shared exchange buffer
template <typename T, int N>
void str_query(int M, int ax, int slab, long long nfast, int* W, int* LS, int* threads, int* uses_smem);
// elements per thread (= radix schedule of the twiddle table) the row kernel uses for (N, M, variant); 0: the default
template <typename T>
inline int row_E_of(int N, int M, int pwv) { return row_E<T>(N, M, pwv); }

}  // namespace ggp
