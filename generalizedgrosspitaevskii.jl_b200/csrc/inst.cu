// One translation unit per (GGP_T, GGP_N): explicit instantiation of the launchers, so the
// library builds in parallel.  Compiled with -DGGP_T=float|double -DGGP_N=<line length>.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "kernels.cuh"
#include "kernels_cp.cuh"
#include "generic.cuh"
#ifdef GGP_TMA
#include "str_tma.cuh"
#endif
#ifdef GGP_PACKED
#include "packed.cuh"
#endif

namespace ggp {

template <typename KernelT>
static int set_smem(KernelT k, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

// Launch with programmatic stream serialization (see pdl_wait() in cplx.cuh): the next kernel's CTAs are scheduled
// while this grid is still running, with their parameters, index arithmetic and table staging done, and block in
// griddepcontrol.wait until this grid has completed.  WHERE a grid lets its dependent go matters:
//  * grids of at least one wave trigger at their very start (pdl_pos 0): the dependent's CTAs slip into the SM slots
//    this grid frees while it drains (C2 chained 63.7 -> 61.8 us/step, 4096^2 313.9 -> 308.2);
//  * grids below one wave (fewer CTAs than SM slots, by the occupancy calculator) trigger right before their final
//    stores (pdl_pos 3): triggered at the start, the dependent's
//    CTAs land next to this grid's on whatever SM has room and unbalance it (1024^2 chained 22.5 -> 26.7 us/step);
//    triggered late, only the launch latency and the prologue overlap this grid's tail: 1024^2 22.5 -> 17.9 us/step,
//    512^2 14.4 -> 10.5 (profiles/r01_notes.md, session 4).
//  * in between (more than one wave, but fewer than 148 x 1024 threads, e.g. C3's register-heavy kernels): plain launch,
//    either trigger position measured slower (C3 125.8 -> 131.3 us/step with pos 3).
// GGP_NO_PDL=1 switches it off, GGP_PDL_POS=0..3 forces a trigger position.  Kernels without a pdl_pos parameter
// (oned_kernel) keep the old rule: programmatic launch only for grids of at least one wave.
static bool pdl_small_grid(unsigned grid, unsigned block) { return (unsigned long long)grid * block < 148ull * 1024ull; }
// the whole grid is resident at once (fewer CTAs than SM slots for this kernel, block size and shared memory)
template <typename KernelT>
static bool pdl_below_one_wave(KernelT k, unsigned grid, unsigned block, size_t smem) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  // one occupancy query per (kernel, block size, shared memory) and host thread, not per launch
  struct Entry { const void* k; unsigned block; size_t smem; int per_sm; };
  thread_local std::vector<Entry> cache;
  int per_sm = -1;
  for (const Entry& e : cache)
    if (e.k == (const void*)k && e.block == block && e.smem == smem) per_sm = e.per_sm;
  if (per_sm < 0) {
    per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, (int)block, smem) != cudaSuccess) {
      cudaGetLastError();
      per_sm = 0;
    }
    cache.push_back(Entry{(const void*)k, block, smem, per_sm});
  }
  if (per_sm < 1) return false;
  return (unsigned long long)grid <= (unsigned long long)per_sm * (unsigned)sms;
}
// trigger position and whether to use the programmatic launch at all, for kernels with a pdl_pos parameter
static int pdl_pos_for(bool below_one_wave, unsigned grid, unsigned block, bool* use) {
  static const int v = getenv("GGP_PDL_POS") ? atoi(getenv("GGP_PDL_POS")) : -1;
  *use = below_one_wave || !pdl_small_grid(grid, block);   // in between (a few waves of few threads): plain launch, as measured
  if (v >= 0) return v > 3 ? 3 : v;
  return below_one_wave ? 3 : 0;
}

template <typename P>
static int launch_pdl(void (*k)(const P), unsigned grid, unsigned block, size_t smem, cudaStream_t st, const P& p,
                      unsigned cluster = 1, bool has_pos = false) {
  static const int mode = getenv("GGP_NO_PDL") ? 0 : (getenv("GGP_PDL") ? 2 : 1);
  const bool pdl = mode == 2 || (mode == 1 && (has_pos || !pdl_small_grid(grid, block)));  // has_pos: the caller decided
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1 && grid % cluster == 0) {
    // co-schedule the CTAs of adjacent fast-axis tiles: their 32-byte row pieces share DRAM pages
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cluster;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return (int)cudaLaunchKernelEx(&cfg, k, p);
}

// Wave fit.  A grid of G CTAs at R resident CTAs per SM runs ceil(G / (R * SMs)) waves, the last of them partly empty
// (C2's contiguous-axis kernel: 2048 CTAs at 8 per SM = 1184 + 864).  If one resident CTA fewer per SM needs no more
// waves, the waves are fuller and every CTA shares its SM with fewer others: 2048 CTAs at 7 per SM = 1036 + 1012.
// Occupancy is lowered by asking for more dynamic shared memory than the kernel uses.  MEASURED (r02, profiles/r02_notes.md):
// it does not pay -- C2's kernel 31.3 us at 8 CTAs per SM, 32.5 at 7, 34.5 at 6, 42.2 at 3: the kernel is bound by
// latency and wants MORE resident warps, not fuller waves -- so the rule is OFF by default.  GGP_WAVE_FIT=1 switches it
// on, GGP_ROW_OCC=<n> forces n CTAs per SM for the contiguous-axis kernel.
template <typename KernelT>
static size_t wave_fit_smem(KernelT k, unsigned grid, unsigned block, size_t smem) {
  static const int mode = getenv("GGP_WAVE_FIT") ? atoi(getenv("GGP_WAVE_FIT")) : 0;
  static const int forced = getenv("GGP_ROW_OCC") ? atoi(getenv("GGP_ROW_OCC")) : 0;
  if (!mode && !forced) return smem;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  struct Entry { const void* k; unsigned grid, block; size_t smem, out; };
  thread_local std::vector<Entry> cache;
  for (const Entry& e : cache)
    if (e.k == (const void*)k && e.grid == grid && e.block == block && e.smem == smem) return e.out;
  size_t out = smem;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, (int)block, smem) != cudaSuccess) {
    cudaGetLastError();
    per_sm = 0;
  }
  int want = per_sm;
  if (forced > 0 && forced < per_sm) {
    want = forced;
  } else if (mode && per_sm > 2) {
    const unsigned long long slots = (unsigned long long)per_sm * sms;
    const unsigned long long waves = (grid + slots - 1) / slots;
    // only for grids of a few waves whose last wave is less than 80 % full
    if (waves >= 2 && waves <= 4 && (unsigned long long)grid * 10 < (waves - 1) * slots * 10 + slots * 8) {
      int r = per_sm;
      while (r > 2 && (unsigned long long)(r - 1) * sms * waves >= grid) --r;
      want = r;
    }
  }
  if (want > 0 && want < per_sm) {
    // shared memory per SM 228 KB, 1 KB of it reserved per resident CTA: `want` CTAs fit, `want + 1` do not
    const size_t s = (size_t)233472 / want - 1024;
    if (s > smem && s <= (size_t)227 * 1024) out = s & ~(size_t)15;
  }
  cache.push_back(Entry{(const void*)k, grid, block, smem, out});
  return out;
}

template <typename T, int N, int M, int PWV>
static int launch_row_mp(const RowParams<T>& p, cudaStream_t st) {
  using K = KCfg<T, N, row_E<T>(N, M, PWV)>;
  if constexpr (M == 2 && !pw_is_stoch(PWV) && CpCfg<T, N>::OK) {
    // two components: component-parallel kernel (kernels_cp.cuh); GGP_NO_CP=1 keeps the one-thread-both-components kernel
    static const bool nocp = getenv("GGP_NO_CP") != nullptr;
    if (!nocp) {
      using C = CpCfg<T, N>;
      const size_t smem_cp = (size_t)C::ROW_LPC * 2 * KCfg<T, N>::row_ls() * sizeof(cpx<T>);
      const unsigned grid_cp = (unsigned)((p.nlines + C::ROW_LPC - 1) / C::ROW_LPC);
      auto kc = row_cp_kernel<T, N, PWV>;
      int ec = set_smem(kc, smem_cp);
      if (ec) return ec;
      RowParams<T> qc = p;
      bool usec = false;
      qc.pdl_pos = pdl_pos_for(pdl_below_one_wave(kc, grid_cp, C::ROW_THREADS, smem_cp), grid_cp, C::ROW_THREADS, &usec);
      return launch_pdl<RowParams<T>>(kc, grid_cp, C::ROW_THREADS, smem_cp, st, qc, 1, usec);
    }
  }
  size_t smem = K::USES_SMEM ? (size_t)K::LPC * M * K::row_ls() * sizeof(cpx<T>) : 0;
  const unsigned grid = (unsigned)((p.nlines + K::LPC - 1) / K::LPC);
  auto k = row_kernel<T, N, M, PWV>;
  smem = wave_fit_smem(k, grid, K::ROW_THREADS, smem);
  int e = set_smem(k, smem);
  if (e) return e;
  RowParams<T> q = p;
  bool use = false;
  q.pdl_pos = pdl_pos_for(pdl_below_one_wave(k, grid, K::ROW_THREADS, smem), grid, K::ROW_THREADS, &use);
  return launch_pdl<RowParams<T>>(k, grid, K::ROW_THREADS, smem, st, q, 1, use);
}

template <typename T, int N, int M>
static int launch_row_m(int pwv, const RowParams<T>& p, cudaStream_t st) {
  if (pwv == PW_KERR) return launch_row_mp<T, N, M, PW_KERR>(p, st);
  if (pwv == PW_DET) return launch_row_mp<T, N, M, PW_DET>(p, st);
  if (pwv == PW_STOCH) return launch_row_mp<T, N, M, PW_STOCH>(p, st);
  if (pwv == PW_DENSE) return launch_row_mp<T, N, M, PW_DENSE>(p, st);
  if (pwv == PW_TW) {   // instantiated for one component (the Truncated-Wigner ensembles); two components: general variant
    if constexpr (M == 1) return launch_row_mp<T, N, M, PW_TW>(p, st);
    else return launch_row_mp<T, N, M, PW_STOCH>(p, st);
  }
  return launch_row_mp<T, N, M, PW_FIELD>(p, st);
}

template <typename T, int N>
int launch_row(int M, int pwv, const RowParams<T>& p, cudaStream_t st) {
  if (M == 1) return launch_row_m<T, N, 1>(pwv, p, st);
  if (M == 2) return launch_row_m<T, N, 2>(pwv, p, st);
  return (int)cudaErrorInvalidValue;
}

template <typename T, int N>
void str_query(int M, int ax, int slab, long long nfast, int* W_, int* LS_, int* threads, int* uses_smem) {
  using K = KCfg<T, N>;
  const int wmax = K::str_max_threads(M) / K::TPL;
  int W = K::WDEF;
  // opt-in: one 1024-thread CTA per SM with twice the tile width -- measured slower (4096^2: 225 vs 208 us per
  // pass, same DRAM traffic; profiles/r01_notes.md session 4)
  if (wmax > W && getenv("GGP_STR_WIDE")) W = wmax;
  // 2048-point fp32 lines (128 threads per line), TMA-staged tiles: 256-thread CTAs of two columns, four per SM, instead
  // of 512-thread CTAs of four columns, two per SM -- while one CTA waits for its tile (22 % of the warp samples of the
  // four-column kernel sat on the tile's mbarrier, ncu r02u) three others compute instead of one.  C2 (2048^2), one
  // gpurun call, both widths with compile-time addresses: cold step 62.45 -> 60.5 us, chained 53.2 -> 52.7.  Not for
  // 1024-point lines (128-thread CTAs: 24.5 -> 25.5 us) and not without TMA (16-byte row pieces with plain loads).
  if (sizeof(T) == 4 && M == 1 && K::TPL == 128 && K::WDEF == 4 && !slab && !getenv("GGP_NO_TMA")) W = 2;
  // the same for 256-point fp64 lines (32 threads per line; C4): 128-thread CTAs of four columns (64-byte rows), eight
  // per SM, instead of 256-thread CTAs of eight columns: strided kernel 0.539 -> 0.51 ms per 1024 trajectories (r02z A/B)
  if (sizeof(T) == 8 && M == 1 && K::TPL == 32 && K::WDEF == 8 && !slab && !getenv("GGP_NO_TMA")) W = 4;
  // z axis of a 3-D grid: consecutive points of a line are a whole xy-plane apart (8 MB at 1024^3), so every row
  // piece of the tile is its own DRAM page and TLB entry; 64-byte pieces instead of 32 halve that cost
  // (1024^3 c64: z pass 17.2 -> 6.7 ms, step 29.4 -> 18.5 ms; 512^3: neutral; 128-byte pieces: slower again)
  if (ax == 2) {
    const int w64 = 64 / (int)sizeof(cpx<T>);
    if (w64 > W && w64 <= wmax) W = w64;
  }
  // slab-decomposed plans: the tile is read with plain coalesced loads and half of it leaves as peer stores over
  // NVLink -- 128-byte row pieces (one full line per load / per NVLink write) instead of 32/64
  // (2 x B200, 512^3: 1.72 -> 1.28 ms/step; 1024^3: 16.7 -> 11.0 ms/step)
  if (slab) {
    int w128 = 128 / (int)sizeof(cpx<T>);
    while (w128 > wmax) w128 >>= 1;
    if (w128 > W) W = w128;
  }
  if (const char* e = getenv(ax == 2 ? "GGP_STR_WZ" : "GGP_STR_W")) {  // tuning knobs (z axis of 3-D grids: GGP_STR_WZ)
    const int w = atoi(e);
    if (w >= 1 && w <= wmax && (w & (w - 1)) == 0) W = w;
  }
  // the exchange lines of the whole tile must fit one CTA's shared memory
  while (W > 1 && K::str_lines_bytes(M, W, K::str_ls(W)) > (size_t)200 * 1024) W >>= 1;
  while (W > nfast) W >>= 1;
  *W_ = W;
  *LS_ = K::str_ls(W);
  *threads = W * K::TPL;
  *uses_smem = K::USES_SMEM ? 1 : 0;
}

#ifdef GGP_TMA
template <typename T, int N, int M>
static int launch_str_tma_m(StrTmaParams<T> p, long long nfast, long long nother, int sm_count, cudaStream_t st) {
  using K = KCfg<T, N>;
  int W, LS, threads, us;
  str_query<T, N>(M, p.ax, 0, nfast, &W, &LS, &threads, &us);
  p.W = W;
  p.logW = ilog2(W);
  p.LS = LS;
  p.ntx = nfast / W;
  p.ntiles = p.ntx * nother;
  const size_t smem = str_tma_smem<T, N>(M, W, LS, p.stage_d ? p.nplanes : 0);
  auto k = str_tma_kernel<T, N, M>;
  int e = set_smem(k, smem);
  if (e) return e;
  int per_sm = 0;
  cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem);
  if (ce != cudaSuccess) return (int)ce;
  if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;
  long long grid = (long long)per_sm * sm_count;
  if (grid > p.ntiles) grid = p.ntiles;
  k<<<(unsigned)grid, threads, smem, st>>>(p);
  return (int)cudaGetLastError();
}

template <typename T, int N>
int launch_str_tma(int M, StrTmaParams<T> p, long long nfast, long long nother, int sm_count, cudaStream_t st) {
  if (M == 1) return launch_str_tma_m<T, N, 1>(p, nfast, nother, sm_count, st);
  if (M == 2) return launch_str_tma_m<T, N, 2>(p, nfast, nother, sm_count, st);
  return (int)cudaErrorInvalidValue;
}

#endif

template <typename T, int N, int M>
static int launch_str_m(StrParams<T> p, long long nfast, long long nother, cudaStream_t st) {
  using K = KCfg<T, N>;
  if constexpr (M == 2 && CpCfg<T, N>::OK && !TwT<T>::split) {
    // two components: component-parallel kernel (kernels_cp.cuh) for everything but the slab plans' scatter passes
    static const bool nocp = getenv("GGP_NO_CP") != nullptr;
    using C = CpCfg<T, N>;
    if (!nocp && !p.scatter && !p.slab && !p.Dq[0] && nfast % C::STR_W == 0) {
      p.W = C::STR_W;
      p.logW = ilog2(C::STR_W);
      p.LS = K::str_ls(2 * C::STR_W);
      p.ntx = nfast / C::STR_W;
      p.tma = 0;
      p.dl_smem = 0;
      size_t smem_cp = ((size_t)2 * C::STR_W * p.LS * sizeof(cpx<T>) + 15) & ~(size_t)15;
      p.tw_smem = 0;
      if (K::TW_STAGED > 0 && (smem_cp + K::TW_BYTES + 1024) * (size_t)C::STR_MINB <= (size_t)227 * 1024 - 1024 &&
          !getenv("GGP_NO_TW_SMEM_RT")) {
        p.tw_smem = 1;
        smem_cp += K::TW_BYTES;
      }
      const long long grid_cp = p.ntx * nother;
      if (grid_cp > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
      auto kc = str_cp_kernel<T, N>;
      int ec = set_smem(kc, smem_cp);
      if (ec) return ec;
      bool usec = false;
      p.pdl_pos = pdl_pos_for(pdl_below_one_wave(kc, (unsigned)grid_cp, (unsigned)C::STR_THREADS, smem_cp), (unsigned)grid_cp,
                              (unsigned)C::STR_THREADS, &usec);
      return launch_pdl<StrParams<T>>(kc, (unsigned)grid_cp, (unsigned)C::STR_THREADS, smem_cp, st, p, 1, usec);
    }
  }
  int W, LS_, threads_, us_;
  str_query<T, N>(M, p.ax, p.slab, nfast, &W, &LS_, &threads_, &us_);
  p.W = W;
  p.logW = ilog2(W);
  p.LS = K::str_ls(W);
  p.ntx = nfast / W;
  size_t smem = K::USES_SMEM ? K::str_lines_bytes(M, W, p.LS) : 0;
  // CTAs per SM the register file allows with this geometry (the launch bounds cap the registers accordingly)
  int nb = K::str_min_blocks(M);
  if (W * K::TPL != K::STR_THREADS && K::data_regs(M) <= 32) nb = 1024 / (W * K::TPL) < 1 ? 1 : 1024 / (W * K::TPL);
  const size_t per_sm = 227 * 1024 - 1024;
  const size_t dl_bytes = (((size_t)N * sizeof(cpx<T>) + 15) & ~(size_t)15);
  p.tw_smem = 0;
  if (K::str_tw_smem(M)) {
    smem += K::TW_BYTES;  // twiddle table behind the exchange lines (compile-time decision)
  } else if (K::USES_SMEM && K::TW_COUNT > 0 && !getenv("GGP_NO_TW_SMEM_RT") &&
             (smem + K::TW_BYTES + dl_bytes + 1024) * (size_t)nb <= per_sm) {
    p.tw_smem = 1;  // ... or decided here for a geometry chosen at run time (wide tiles of the long fp32 lines)
    smem += K::TW_BYTES;
  }
  // separable exp_D: stage D_line behind that if it does not cost a resident CTA
  p.dl_smem = 0;
  if (K::USES_SMEM && p.mode == 1 && p.dkind == KIND_SEP && !getenv("GGP_NO_DL_SMEM")) {
    const size_t with_dl = smem + dl_bytes;
    if ((with_dl + 1024) * (size_t)nb <= per_sm) {
      p.dl_smem = 1;
      smem = with_dl;
    }
  }
  if (p.occ1 && smem < (size_t)116 * 1024) smem = (size_t)116 * 1024;   // more than half an SM's shared memory: one CTA per SM
  if (!K::USES_SMEM || K::str_parked(M)) p.tma = 0;   // (parked two-component path: plain loads, see KCfg::str_parked)
  if (p.tma) {
    smem = (smem + 15) & ~(size_t)15;
    p.mbar_off = (int)smem;
    smem += 16;
  }
  const long long grid = p.ntx * nother;
  if (grid > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  // next-tile L2 prefetch (str_kernel): opt-in, GGP_STR_PF = distance in CTAs.  Measured SLOWER (4096^2: 228 -> 257 us
  // per step with a distance of one wave, 8192^2: 1278 -> 1379): the prefetches compete with the demand loads for
  // the same DRAM pages instead of hiding them (profiles/r01_notes.md, session 4)
  p.pf_dist = 0;
  {
    static const int pf_env = getenv("GGP_STR_PF") ? atoi(getenv("GGP_STR_PF")) : 0;
    if (p.tma && !p.scatter && pf_env > 0) p.pf_dist = pf_env;
  }
  // the default geometry runs the instantiation with a compile-time tile width (GGP_STR_WRT=1: always the generic one)
  static const bool wrt = getenv("GGP_STR_WRT") != nullptr;
  auto k = (W == K::WDEF && !wrt) ? str_kernel<T, N, M, K::WDEF> : str_kernel<T, N, M, 0>;
  if constexpr (M == 1 && K::WDEF >= 4) {
    // half the default width, also with compile-time addresses (GGP_STR_W=<WDEF/2>): four smaller CTAs per SM hide the
    // wait for the tile better when the state comes from HBM (2048^2 cold: 40.1 -> 37.1 us per launch, r02y)
    if (W == K::WDEF / 2 && !wrt) k = str_kernel<T, N, M, K::WDEF / 2>;
  }
  int e = set_smem(k, smem);
  if (e) return e;
  bool use = false;
  p.pdl_pos = pdl_pos_for(pdl_below_one_wave(k, (unsigned)grid, (unsigned)(W * K::TPL), smem), (unsigned)grid,
                          (unsigned)(W * K::TPL), &use);
  static const unsigned cl = getenv("GGP_STR_CLUSTER") ? (unsigned)atoi(getenv("GGP_STR_CLUSTER")) : 1u;
  return launch_pdl<StrParams<T>>(k, (unsigned)grid, (unsigned)(W * K::TPL), smem, st, p, cl, use);
}

template <typename T, int N>
int launch_str(int M, StrParams<T> p, long long nfast, long long nother, cudaStream_t st) {
  if (M == 1) return launch_str_m<T, N, 1>(p, nfast, nother, st);
  if (M == 2) return launch_str_m<T, N, 2>(p, nfast, nother, st);
  return (int)cudaErrorInvalidValue;
}

template <typename T, int N, int M, int PWV>
static int launch_oned_mp(const OneDParams<T>& p, cudaStream_t st) {
  using K = KCfg<T, N>;
  const size_t smem = K::USES_SMEM ? (size_t)K::LPC * M * K::row_ls() * sizeof(cpx<T>) : 0;
  const unsigned grid = (unsigned)((p.nlines + K::LPC - 1) / K::LPC);
  auto k = oned_kernel<T, N, M, PWV>;
  int e = set_smem(k, smem);
  if (e) return e;
  return launch_pdl<OneDParams<T>>(k, grid, K::ROW_THREADS, smem, st, p);
}

template <typename T, int N, int M>
static int launch_oned_m(int pwv, const OneDParams<T>& p, cudaStream_t st) {
  if (pwv == PW_KERR) return launch_oned_mp<T, N, M, PW_KERR>(p, st);
  if (pwv == PW_DET) return launch_oned_mp<T, N, M, PW_DET>(p, st);
  if (pwv == PW_STOCH) return launch_oned_mp<T, N, M, PW_STOCH>(p, st);
  if (pwv == PW_DENSE) return launch_oned_mp<T, N, M, PW_DENSE>(p, st);
  if (pwv == PW_TW) return launch_oned_mp<T, N, M, PW_STOCH>(p, st);   // (1-D whole-step kernel: the general stochastic variant)
  return launch_oned_mp<T, N, M, PW_FIELD>(p, st);
}

template <typename T, int N>
int launch_oned(int M, int pwv, const OneDParams<T>& p, cudaStream_t st) {
  if (M == 1) return launch_oned_m<T, N, 1>(pwv, p, st);
  if (M == 2) return launch_oned_m<T, N, 2>(pwv, p, st);
  return (int)cudaErrorInvalidValue;
}

#if defined(GGP_PACKED) && GGP_N >= GGP_PACKED_MIN_N
template <int N>
int launch_row2(const RowParams<float>& p, cudaStream_t st) {
  using K = KCfg<f2, N>;
  const size_t smem = K::USES_SMEM ? (size_t)K::LPC * K::row_ls() * sizeof(cpx<f2>) : 0;
  const long long pairs = p.nlines / 2;
  const unsigned grid = (unsigned)((pairs + K::LPC - 1) / K::LPC);
  auto k = row2_kernel<N>;
  int e = set_smem(k, smem);
  if (e) return e;
  k<<<grid, K::ROW_THREADS, smem, st>>>(p);
  return (int)cudaGetLastError();
}
template <int N>
int launch_str2(StrParams<float> p, long long nfast, long long nother, cudaStream_t st) {
  using K = KCfg<f2, N>;
  int W = K::WDEF;  // column pairs per CTA
  if (const char* e = getenv("GGP_STR_W")) {
    const int w = atoi(e) / 2;
    if (w >= 1 && w <= K::WDEF && (w & (w - 1)) == 0) W = w;
  }
  while (2 * W > nfast) W >>= 1;
  if (W < 1) return (int)cudaErrorInvalidConfiguration;
  p.W = W;
  p.logW = ilog2(W);
  p.LS = K::str_ls(W);
  p.ntx = nfast / (2 * W);
  const size_t smem = K::USES_SMEM ? (size_t)W * p.LS * sizeof(cpx<f2>) : 0;
  const long long grid = p.ntx * nother;
  if (grid > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  auto k = str2_kernel<N>;
  int e = set_smem(k, smem);
  if (e) return e;
  k<<<(unsigned)grid, W * K::TPL, smem, st>>>(p);
  return (int)cudaGetLastError();
}
template int launch_row2<GGP_N>(const RowParams<float>&, cudaStream_t);
template int launch_str2<GGP_N>(StrParams<float>, long long, long long, cudaStream_t);
#endif

// generic (unfused) line transform of the generic plan: plain for n == L, Bluestein for n < L (generic.cuh)
template <typename T, int L>
int launch_gen_fft(GenFftParams<T> p, cudaStream_t st) {
  int threads = 0;
  size_t smem = 0;
  gen_fft_geometry<T, L>(p.sa, &p.W, &p.LS, &threads, &smem);
  const long long grid = (p.nlines + p.W - 1) / p.W;
  if (grid > 0x7fffffffLL || grid < 1) return (int)cudaErrorInvalidConfiguration;
  auto k = gen_fft_kernel<T, L>;
  int e = set_smem(k, smem);
  if (e) return e;
  k<<<(unsigned)grid, threads, smem, st>>>(p);
  return (int)cudaGetLastError();
}
template int launch_gen_fft<GGP_T, GGP_N>(GenFftParams<GGP_T>, cudaStream_t);

template int launch_row<GGP_T, GGP_N>(int, int, const RowParams<GGP_T>&, cudaStream_t);
template int launch_str<GGP_T, GGP_N>(int, StrParams<GGP_T>, long long, long long, cudaStream_t);
template int launch_oned<GGP_T, GGP_N>(int, int, const OneDParams<GGP_T>&, cudaStream_t);
#ifdef GGP_TMA
template int launch_str_tma<GGP_T, GGP_N>(int, StrTmaParams<GGP_T>, long long, long long, int, cudaStream_t);
#endif
template void str_query<GGP_T, GGP_N>(int, int, int, long long, int*, int*, int*, int*);

}  // namespace ggp
