// fp32 kernels on PACKED pairs of lines: every thread carries two lines (two adjacent rows in the
// contiguous-axis kernel, two adjacent columns in the strided kernel) in cpx<f2> registers, so each
// butterfly add / multiply / fma is ONE Blackwell fp32x2 instruction (FADD2 / FMUL2 / FFMA2) for both
// lines, and each shared-memory exchange moves both lines with one 128-bit access.  The scalar fp32
// kernels are bound by instruction issue (profiles/r01_notes.md); this halves their FP issue slots.
//
// Covered: M = 1, ComplexF32; row kernel with the PW_KERR half-step (C1/C2/C5 class); strided kernel with
// exp_D of kind none / scalar / separable.  Everything else keeps the scalar kernels.
#pragma once
#include "kernels.cuh"

#ifndef GGP_PACKED_MIN_N
#define GGP_PACKED_MIN_N 64
#endif

namespace ggp {

__device__ __forceinline__ cpx<f2> pack2(const float2 a, const float2 b) {
  return mk<f2>(mkf2(a.x, b.x), mkf2(a.y, b.y));
}
__device__ __forceinline__ cpx<f2> dup2(const cpx<float> w) { return mk<f2>(mkf2(w.x, w.x), mkf2(w.y, w.y)); }

static __device__ __noinline__ void sincos2_slow(const f2 a, f2* s, f2* c) {
  float s0, c0, s1, c1;
  sincosf(a.v.x, &s0, &c0);
  sincosf(a.v.y, &s1, &c1);
  *s = mkf2(s0, s1);
  *c = mkf2(c0, c1);
}

// sin/cos of two small angles at once (same polynomial pair as sincos_t<float>); falls back per lane
__device__ __forceinline__ void sincos2(const f2 a, f2* s, f2* c) {
  if (fmaxf(fabsf(a.v.x), fabsf(a.v.y)) <= 0.78539816f) {
    const f2 z = a * a;
    f2 ps = fma_(z, cst<f2>(-1.9515295891e-4), cst<f2>(8.3321608736e-3));
    ps = fma_(z, ps, cst<f2>(-1.6666654611e-1));
    *s = fma_(a * z, ps, a);
    f2 pc = fma_(z, cst<f2>(2.443315711809948e-5), cst<f2>(-1.388731625493765e-3));
    pc = fma_(z, pc, cst<f2>(4.166664568298827e-2));
    *c = fma_(z * z, pc, fma_(z, cst<f2>(-0.5), cst<f2>(1.0)));
  } else {
    sincos2_slow(a, s, c);
  }
}

// ---- contiguous-axis kernel, two rows per thread group -----------------------------------------
template <int N>
__global__ void __launch_bounds__(KCfg<f2, N>::ROW_THREADS) row2_kernel(const RowParams<float> p) {
  using K = KCfg<f2, N>;
  constexpr int E = K::E, TPL = K::TPL, LPC = K::LPC, LS = K::row_ls();
  using SYNC = typename K::RowSync;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<f2>* smem = reinterpret_cast<cpx<f2>*>(smem_raw);

  const int grp = threadIdx.x / TPL, t = threadIdx.x % TPL;
  const long long pair = (long long)blockIdx.x * LPC + grp;
  const bool active = 2 * pair < p.nlines;
  const long long goff = 2 * pair * N + t;
  cpx<f2>* sl = smem + (size_t)grp * LS;
  const float2* ua = reinterpret_cast<const float2*>(p.u[0]) + goff;
  const float2* ub = ua + N;

  cpx<f2> v[1][E];
#pragma unroll
  for (int m = 0; m < E; ++m)
    v[0][m] = active ? pack2(ua[m * TPL], ub[m * TPL]) : mk<f2>(mkf2(0.f, 0.f), mkf2(0.f, 0.f));

  const cpx<f2>* tw = reinterpret_cast<const cpx<f2>*>(p.tw2);
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {
    if (p.flags & (1 << it)) fft_fwd_all<f2, N, 1, SYNC>(v, t, sl, LS, tw, it == 0);
    if (it == 0) {
      // PW_KERR: u <- cis(-dt*(c + g|u|^2)) u.  |u| is invariant under that pure phase, so the trailing
      // half-step of step n and the leading half-step of step n+1 are ONE rotation by the summed angle.
      const int napply = (p.hs[0].apply ? 1 : 0) + (p.hs[1].apply ? 1 : 0);
      if (napply) {
        const f2 cdt = cst<f2>(-(double)napply * p.pw.dt * p.pw.nl_c_re[0]);
        const f2 gdt = cst<f2>(-(double)napply * p.pw.dt * p.pw.nl_g_re[0][0]);
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const f2 ang = fma_(gdt, cabs2(v[0][m]), cdt);
          f2 s, c;
          sincos2(ang, &s, &c);
          v[0][m] = cmul(mk<f2>(c, s), v[0][m]);
        }
      }
    }
  }
  if (active) {
    float2* wa = reinterpret_cast<float2*>(p.u[0]) + goff;
    float2* wb = wa + N;
#pragma unroll
    for (int m = 0; m < E; ++m) {
      wa[m * TPL] = make_float2(v[0][m].x.v.x, v[0][m].y.v.x);
      wb[m * TPL] = make_float2(v[0][m].x.v.y, v[0][m].y.v.y);
    }
  }
}

// ---- strided-axis kernel, two adjacent columns per thread group ----------------------------------
// p.W / p.logW count column PAIRS here.
template <int N>
__global__ void __launch_bounds__(KCfg<f2, N>::STR_THREADS, KCfg<f2, N>::STR_THREADS <= 256 ? 2 : 1)
    str2_kernel(const StrParams<float> p) {
  using K = KCfg<f2, N>;
  constexpr int E = K::E, TPL = K::TPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<f2>* smem = reinterpret_cast<cpx<f2>*>(smem_raw);

  const int xp = threadIdx.x & (p.W - 1), t = threadIdx.x >> p.logW;
  const long long g = blockIdx.x;
  const long long xt = g % p.ntx, o = g / p.ntx;
  const long long o1 = o % p.no1, o2 = o / p.no1;
  const long long x0 = xt * (2 * p.W) + 2 * xp;
  const long long off = x0 + o1 * p.s1 + o2 * p.s2 + (long long)t * p.ls;
  const long long toff = x0 + o1 * p.ts1 + (long long)t * p.ls;
  const long long mstride = (long long)TPL * p.ls;
  cpx<f2>* sl = smem + (size_t)xp * p.LS;
  const cpx<f2>* tw = reinterpret_cast<const cpx<f2>*>(p.tw2);

  cpx<f2> v[1][E];
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const float4 q = *reinterpret_cast<const float4*>(p.u[0] + off + m * mstride);
    v[0][m] = mk<f2>(mkf2(q.x, q.z), mkf2(q.y, q.w));
  }

  const int it0 = p.mode == 2 ? 1 : 0, it1 = p.mode == 0 ? 0 : 1;
#pragma unroll 1
  for (int it = it0; it <= it1; ++it) {
    fft_fwd_all<f2, N, 1, SyncBlock>(v, t, sl, p.LS, tw, it == 1);
    if (it == 0 && p.mode == 1) {
      if (p.dkind == KIND_SEP) {
        const float4 q = *reinterpret_cast<const float4*>(p.D[0] + (toff - (long long)t * p.ls));
        const cpx<f2> dperp = mk<f2>(mkf2(q.x, q.z), mkf2(q.y, q.w));
        const cpx<float>* dline = p.D[1] + t;
#pragma unroll
        for (int m = 0; m < E; ++m) v[0][m] = cmul(cmul(dperp, dup2(dline[m * TPL])), v[0][m]);
      } else if (p.dkind == KIND_SCALAR) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const float4 q = *reinterpret_cast<const float4*>(p.D[0] + toff + m * mstride);
          v[0][m] = cmul(mk<f2>(mkf2(q.x, q.z), mkf2(q.y, q.w)), v[0][m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < E; ++m)
    *reinterpret_cast<float4*>(p.u[0] + off + m * mstride) =
        make_float4(v[0][m].x.v.x, v[0][m].y.v.x, v[0][m].x.v.y, v[0][m].y.v.y);
}

// launchers (instantiated in the float translation units only)
template <int N>
int launch_row2(const RowParams<float>& p, cudaStream_t st);
template <int N>
int launch_str2(StrParams<float> p, long long nfast, long long ngroups_other, cudaStream_t st);

}  // namespace ggp
