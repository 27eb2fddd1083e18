// One translation unit per (GGP_T, GGP_N): explicit instantiation of the launchers, so the
// library builds in parallel.  Compiled with -DGGP_T=float|double -DGGP_N=<line length>.
#include "kernels.cuh"

namespace ggp {

template <typename KernelT>
static int set_smem(KernelT k, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

template <typename T, int N, int M>
static int launch_row_m(bool pre, bool post, const RowParams<T>& p, cudaStream_t st) {
  using K = KCfg<T, N>;
  const size_t smem = K::USES_SMEM ? (size_t)K::LPC * M * K::row_ls() * sizeof(cpx<T>) : 0;
  const unsigned grid = (unsigned)((p.nlines + K::LPC - 1) / K::LPC);
  int e = 0;
#define GGP_ROW(PRE, POST)                                                        \
  {                                                                               \
    auto k = row_kernel<T, N, M, PRE, POST>;                                      \
    if ((e = set_smem(k, smem))) return e;                                        \
    k<<<grid, K::ROW_THREADS, smem, st>>>(p);                                     \
  }
  if (pre && post) GGP_ROW(true, true)
  else if (pre) GGP_ROW(true, false)
  else if (post) GGP_ROW(false, true)
  else GGP_ROW(false, false)
#undef GGP_ROW
  return (int)cudaGetLastError();
}

template <typename T, int N>
int launch_row(int M, bool pre, bool post, const RowParams<T>& p, cudaStream_t st) {
  if (M == 1) return launch_row_m<T, N, 1>(pre, post, p, st);
  if (M == 2) return launch_row_m<T, N, 2>(pre, post, p, st);
  return (int)cudaErrorInvalidValue;
}

template <typename T, int N, int M>
static int launch_str_m(int mode, StrParams<T> p, long long nfast, long long nother, cudaStream_t st) {
  using K = KCfg<T, N>;
  int W = K::WDEF;
  while (W > nfast) W >>= 1;
  p.W = W;
  p.logW = ilog2(W);
  p.LS = K::str_ls(W);
  p.ntx = nfast / W;
  const size_t smem = K::USES_SMEM ? (size_t)W * M * p.LS * sizeof(cpx<T>) : 0;
  const long long grid = p.ntx * nother;
  if (grid > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  int e = 0;
#define GGP_STR(MODE)                                                             \
  {                                                                               \
    auto k = str_kernel<T, N, M, MODE>;                                           \
    if ((e = set_smem(k, smem))) return e;                                        \
    k<<<(unsigned)grid, W * K::TPL, smem, st>>>(p);                               \
  }
  if (mode == 0) GGP_STR(0)
  else if (mode == 1) GGP_STR(1)
  else GGP_STR(2)
#undef GGP_STR
  return (int)cudaGetLastError();
}

template <typename T, int N>
int launch_str(int M, int mode, StrParams<T> p, long long nfast, long long nother, cudaStream_t st) {
  if (M == 1) return launch_str_m<T, N, 1>(mode, p, nfast, nother, st);
  if (M == 2) return launch_str_m<T, N, 2>(mode, p, nfast, nother, st);
  return (int)cudaErrorInvalidValue;
}

template <typename T, int N, int M>
static int launch_oned_m(const OneDParams<T>& p, cudaStream_t st) {
  using K = KCfg<T, N>;
  const size_t smem = K::USES_SMEM ? (size_t)K::LPC * M * K::row_ls() * sizeof(cpx<T>) : 0;
  const unsigned grid = (unsigned)((p.nlines + K::LPC - 1) / K::LPC);
  auto k = oned_kernel<T, N, M>;
  int e = set_smem(k, smem);
  if (e) return e;
  k<<<grid, K::ROW_THREADS, smem, st>>>(p);
  return (int)cudaGetLastError();
}

template <typename T, int N>
int launch_oned(int M, const OneDParams<T>& p, cudaStream_t st) {
  if (M == 1) return launch_oned_m<T, N, 1>(p, st);
  if (M == 2) return launch_oned_m<T, N, 2>(p, st);
  return (int)cudaErrorInvalidValue;
}

template int launch_row<GGP_T, GGP_N>(int, bool, bool, const RowParams<GGP_T>&, cudaStream_t);
template int launch_str<GGP_T, GGP_N>(int, int, StrParams<GGP_T>, long long, long long, cudaStream_t);
template int launch_oned<GGP_T, GGP_N>(int, const OneDParams<GGP_T>&, cudaStream_t);

}  // namespace ggp
