// The fused real-space half-step of the reference's muladd_kernel! (src/kernels.jl:37-54), for the
// registered forms:
//   u_i <- sum_j E_ij (u_j + (dt/2) a_now S_j) + (dt/2) a_next S_i - i sqrt(dt) eta_i xi_i
//   E    = cis(-dt G(u)) (x) exp_V[r],   G_i = c_i + sum_j g_ij |u_j|^2   evaluated on the PRE-update field
// and the k-space multiply  u~ <- exp_D[k] (x) u~  (src/strang_splitting.jl:73-74).
// `dt` here is the half-step (reference passes dt/2, src/strang_splitting.jl:87,89).
#pragma once
#include "cplx.cuh"

namespace ggp {

// KIND_SEP: a scalar exp_D that factorises as D_perp[other axes] * D_line[strided axis] (true whenever the
// dispersion is a sum over axes, e.g. |k|^2/2): the strided kernel then reads one D_perp value per
// thread and a cache-resident vector instead of a full-grid table.
enum { KIND_NONE = 0, KIND_SCALAR = 1, KIND_DIAG = 2, KIND_FULL = 3, KIND_SEP = 4 };
enum { NOISE_OFF = 0, NOISE_HOST = 1, NOISE_PHILOX = 2 };

// per-half-step scalars
template <typename T>
struct HalfStep {
  cpx<T> fnow;    // (dt/2) * a_now   (src/kernels.jl:46)
  cpx<T> fnext;   // (dt/2) * a_next  (src/kernels.jl:45)
  const void* xi[2];  // host-fed noise (test mode), one array per component
  // dense time-dependent pump (GGP_PUMP_DENSE): F_now / F_next evaluated on the grid by the host for THIS half-step
  // (the reference's two pump buffers, src/misc.jl:22-42); fnow = fnext = dt/4
  const void* pd_now[2];
  const void* pd_next[2];
  // PW_TW: fnow * S and fnext * S for a pump that is constant in space (per component), formed once on the host
  cpx<T> aS[2], bS[2];
  uint32_t ctr;   // global half-step counter (Philox), low word
  uint32_t ctr_hi;  // bits 32.. of the pair index (folded into the Philox key)
  int apply;      // 0: skip this half-step
};

template <typename T>
struct PointwiseParams {
  T dt;        // half step
  T sqrt_dt;   // sqrt(half step)
  const cpx<T>* expV[4];
  const cpx<T>* S[2];
  int vkind;   // KIND_*
  int pump;    // 0 none, 1 scalar broadcast (S[0] for every component), 2 per component
  int pump_const;      // the pump profile is the same at every grid point (e.g. examples/truncated_wigner.jl:37-39):
  cpx<T> S_const[2];   // its value rides in the parameters and the table is not read
  int pump_zero[2];    // per-component pump (pump == 2): this component's profile is identically zero, skip its loads
  int pump_dense;      // profiles come from HalfStep::pd_now / pd_next instead of S (variants PW_DENSE / PW_FIELD)
  int nl;      // 0 none, 1 real coefficients, 2 complex coefficients
  T nl_c_re[2], nl_c_im[2];
  T nl_g_re[2][2], nl_g_im[2][2];
  int noise;   // NOISE_*
  int noise_real;
  cpx<T> eta[2];
  // GGP_NOISE_FIELD: eta_i(u, r) = P[k1] (eta_i + sum_j alpha_ij |u_j|), |u_j| of the pre-update field; P indexed by
  // the first grid index only (the reference's `point`, quirk Q2); nprof == nullptr: P = 1
  int noise_field;
  cpx<T> alpha[2][2];
  const cpx<T>* nprof;
  int n1;
  uint32_t seed_lo, seed_hi;
  long long elem_offset;  // global element index of this plan's element 0 (batch_offset * nspatial)
  // PW_TW: products formed once on the host -- -dt * c_i, -dt * g_ij (real coefficients) and -i sqrt(dt) eta_i
  T nl_cd[2], nl_gd[2][2];
  cpx<T> eta_s[2];
};

// Rotation by the nonlinear phase a = -dt*G, returned as (sin a, cos a - 1).  The phase of one half-step is
// small in every physical run (|dt*G| << 1), so the fast path is the minimax polynomial pair on
// [-pi/4, pi/4] (Cephes sinf/cosf coefficients, no range reduction, ~12 FMAs); anything larger takes the
// library routine (warp-divergent, rare).  cos a - 1 instead of cos a: for small a, fl(cos a) sits within
// half an ulp of 1 -- a modulus error of up to 3e-8 that is the SAME every step wherever |u|^2 is steady and
// therefore accumulates linearly (2e-8 per step measured on the fp32 path); u + (cm1, s) (x) u with cm1
// carried to full relative precision leaves only data-dependent rounding noise.
__device__ __forceinline__ void sincosm1_t(float a, float* s, float* cm1) {
  if (fabsf(a) <= 0.78539816f) {
    const float z = a * a;
    *s = fmaf(a * z, fmaf(z, fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), a);
    *cm1 = fmaf(z * z, fmaf(z, fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                -0.5f * z);
  } else {
    float c;
    sincosf(a, s, &c);
    *cm1 = c - 1.0f;
  }
}
// The coefficients live in constant memory: a 64-bit immediate does not fit a DFMA, so as literals every coefficient
// cost two UMOVs per use -- 24 of the 37 instructions of one evaluation (SASS of the stochastic fp64 row kernel, C4:
// UMOV was 7.8 % of all executed instructions, ncu r02o); as c[bank][offset] operands they cost none.
static __constant__ double kSin64[6] = {1.58962301576546568060e-10, -2.50507477628578072866e-8, 2.75573136213857245213e-6,
                                        -1.98412698295895385996e-4, 8.33333333332211858878e-3, -1.66666666666666307295e-1};
static __constant__ double kCos64[6] = {-1.13585365213876817300e-11, 2.08757008419747316778e-9, -2.75573141792967388112e-7,
                                        2.48015872888517045348e-5, -1.38888888888730564116e-3, 4.16666666666665929218e-2};
__device__ __forceinline__ void sincosm1_t(double a, double* s, double* cm1) {
  if (fabs(a) <= 0.78539816339744830962) {
    const double z = a * a;
    double ps = kSin64[0];
    ps = fma(ps, z, kSin64[1]);
    ps = fma(ps, z, kSin64[2]);
    ps = fma(ps, z, kSin64[3]);
    ps = fma(ps, z, kSin64[4]);
    ps = fma(ps, z, kSin64[5]);
    *s = fma(a * z, ps, a);
    double pc = kCos64[0];
    pc = fma(pc, z, kCos64[1]);
    pc = fma(pc, z, kCos64[2]);
    pc = fma(pc, z, kCos64[3]);
    pc = fma(pc, z, kCos64[4]);
    pc = fma(pc, z, kCos64[5]);
    *cm1 = fma(z * z, pc, -0.5 * z);
  } else {
    double c;
    sincos(a, s, &c);
    *cm1 = c - 1.0;
  }
}
// u + (cm1 + i s) u  =  (cos a + i sin a) u
template <typename T>
__device__ __forceinline__ cpx<T> rotate_m1(cpx<T> u, T cm1, T s) {
#ifdef GGP_OLD_ROT
  return cmul(mk<T>(cm1 + (T)1, s), u);
#else
  return mk<T>(fma_(cm1, u.x, fnma_(s, u.y, u.x)), fma_(cm1, u.y, fma_(s, u.x, u.y)));
#endif
}
__device__ __forceinline__ float expm1_t(float a) { return expm1f(a); }
__device__ __forceinline__ double expm1_t(double a) { return expm1(a); }

// Philox4x32-10 (Salmon et al., SC'11), counter = (element index lo, hi, half-step, component).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Noise stream.  One Philox call per (element, step pair, component) feeds BOTH half-steps that are
// applied back to back around a step boundary: half-step counter c (0, 1, 2, ...) belongs to pair
// (c+1)>>1 and uses words (x,y) if (c+1) is even, (z,w) otherwise -- the trailing half-step of step n and
// the leading half-step of step n+1 (the two the fused kernels apply together) share a call.
// The counter is the GLOBAL element index, so the stream does not depend on how trajectories are sharded.
// `pair` = (half-step counter + 1) >> 1 as a 64-bit number: low word in the counter, high word XORed into the key
// (zero for the first 2^33 half-steps, so earlier streams are unchanged).
template <typename T>
__device__ __forceinline__ uint4 philox_for(long long gidx, uint32_t pair_lo, uint32_t pair_hi, int comp, uint32_t k0,
                                            uint32_t k1) {
  return philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)((unsigned long long)gidx >> 32), pair_lo, (uint32_t)comp),
                       k0, k1 ^ pair_hi);
}

// xi with <|xi|^2> = 1 (complex prototype) or <xi^2> = 1 (real prototype): Box-Muller in fp32 with the
// fast MUFU paths (|error| ~ 1e-6, irrelevant for a random sample; the exact-parity route for stochastic
// runs is the host-fed buffer).
template <typename T>
__device__ __forceinline__ cpx<T> normal_from(const uint4 r, uint32_t ctr, int real_proto) {
  const bool second = ((ctr + 1u) & 1u) != 0;
  const uint32_t w0 = second ? r.z : r.x, w1 = second ? r.w : r.y;
  const float u1 = ((float)(w0 >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  const float u2 = ((float)(w1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  const float l = -__logf(u1);
  if (real_proto) {
    const float rad = sqrtf(2.0f * l);
    return mk<T>((T)(rad * c), (T)0);
  }
  const float rad = sqrtf(l);
  return mk<T>((T)(rad * c), (T)(rad * s));
}

// Compile-time variants of the half-step (each kernel is instantiated for the three of them, so the
// 16-fold unrolled per-point code only contains what the problem needs -- instruction-cache
// footprint, see profiles/r01_notes.md):
//   PW_KERR   real diagonal nonlinearity only (C1, C2, C5: scalar Kerr GPE)
//   PW_DET    + complex nonlinearity, potential table (any kind), separable pump   (C3, bistability)
//   PW_STOCH  + position noise (host-fed or Philox), constant amplitude            (C4, windowed FT)
//   PW_FIELD  + field- / position-dependent noise amplitude (GGP_NOISE_FIELD); its own variant because the extra
//             live values cost the constant-amplitude kernel 4 % when they shared one (C4: 8.52 -> 8.88 ms/step)
//   PW_DENSE  PW_DET + dense time-dependent pump (profiles per half-step, GGP_PUMP_DENSE); with noise: PW_FIELD
//   PW_TW     PW_STOCH for the Truncated-Wigner shape (C4, examples/truncated_wigner.jl): real nonlinearity coefficients
//             (or none), no potential table, pump none or constant in space, constant noise amplitude, in-kernel Philox.
//             Everything PW_STOCH decides per point at run time is fixed here: no flag tests, no table pointers, and
//             the two half-steps of a pair are unrolled (ncu r02o: ISETP + BRA + LDC were 16 % of the executed instructions)
enum { PW_KERR = 0, PW_DET = 1, PW_STOCH = 2, PW_FIELD = 3, PW_DENSE = 4, PW_TW = 5 };
__host__ __device__ constexpr bool pw_is_stoch(int pwv) { return pwv == PW_STOCH || pwv == PW_FIELD || pwv == PW_TW; }
__host__ __device__ constexpr bool pw_has_dense(int pwv) { return pwv == PW_DENSE || pwv == PW_FIELD; }

// One real-space half-step at one grid point.  sidx: index into the spatial tables; gidx: local
// element index (spatial + batch) for noise.
template <typename T, int M, int PWV>
__device__ __forceinline__ void half_step_point(cpx<T> (&f)[M], const PointwiseParams<T>& p, const HalfStep<T>& h,
                                                const long long sidx, const long long gidx, const uint4* rnd,
                                                const cpx<T>* spre = nullptr) {
  // spre: the separable pump profile of this point, already loaded by the caller (one value per component) -- the
  // component-parallel kernels fetch the profiles of a batch of points together instead of one dependent load per
  // point and half-step (kernels_cp.cuh)
  if constexpr (PWV == PW_KERR) {
    T n2[M];
#pragma unroll
    for (int j = 0; j < M; ++j) n2[j] = cabs2(f[j]);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      T gre = p.nl_c_re[i];
#pragma unroll
      for (int j = 0; j < M; ++j) gre += p.nl_g_re[i][j] * n2[j];
      T s, cm1;
      sincosm1_t(-p.dt * gre, &s, &cm1);
      f[i] = rotate_m1(f[i], cm1, s);
    }
    return;
  }
  if constexpr (PWV == PW_TW) {
    // u_i <- cis(-dt G_i) (u_i + fnow S) + fnext S - i sqrt(dt) eta_i xi_i  with G real, S constant in space (zero
    // without a pump); every product of plan / half-step constants arrives ready-made (nl_cd, nl_gd, aS, bS, eta_s)
    T n2[M];
#pragma unroll
    for (int j = 0; j < M; ++j) n2[j] = cabs2(f[j]);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      T arg = p.nl_cd[i];
#pragma unroll
      for (int j = 0; j < M; ++j) arg = fma_(p.nl_gd[i][j], n2[j], arg);
      T s, cm1;
      sincosm1_t(arg, &s, &cm1);
      const cpx<T> r = rotate_m1(f[i] + h.aS[i], cm1, s) + h.bS[i];
      const cpx<T> xi = normal_from<T>(rnd[i], h.ctr, p.noise_real);
      const cpx<T> e = p.eta_s[i];
      f[i] = mk<T>(fma_(e.x, xi.x, fnma_(e.y, xi.y, r.x)), fma_(e.x, xi.y, fma_(e.y, xi.x, r.y)));
    }
    return;
  }
  // field-dependent noise amplitude: |u_j| of the PRE-update field (src/kernels.jl:40-42)
  cpx<T> etav[M];
  if constexpr (PWV == PW_FIELD) {
    T av[M];
#pragma unroll
    for (int j = 0; j < M; ++j) av[j] = sqrt(cabs2(f[j]));
#pragma unroll
    for (int i = 0; i < M; ++i) {
      etav[i] = p.eta[i];
#pragma unroll
      for (int j = 0; j < M; ++j)
        etav[i] = mk<T>(fma_(p.alpha[i][j].x, av[j], etav[i].x), fma_(p.alpha[i][j].y, av[j], etav[i].y));
      if (p.nprof) etav[i] = cmul(p.nprof[sidx % p.n1], etav[i]);
    }
  }
  // nonlinear phase on the pre-update field, kept as ph = cis(-dt G) - 1
  cpx<T> ph[M];
  if (p.nl) {
    T n2[M];
#pragma unroll
    for (int j = 0; j < M; ++j) n2[j] = cabs2(f[j]);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      T gre = p.nl_c_re[i];
#pragma unroll
      for (int j = 0; j < M; ++j) gre += p.nl_g_re[i][j] * n2[j];
      T s, cm1;
      sincosm1_t(-p.dt * gre, &s, &cm1);
      ph[i] = mk<T>(cm1, s);
      if (p.nl == 2) {  // |cis(-dt G)| = exp(dt Im G):  e (1 + ph) - 1 = em1 + e ph
        T gim = p.nl_c_im[i];
#pragma unroll
        for (int j = 0; j < M; ++j) gim += p.nl_g_im[i][j] * n2[j];
        const T em1 = expm1_t(p.dt * gim);
        ph[i] = mk<T>(fma_(em1, cm1, cm1 + em1), fma_(em1, s, s));
      }
    }
  }
  // w = f + (dt/2) a_now S
  cpx<T> w[M], sv[M];
  if (p.pump) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      // dense pump (its own compile-time variants: carried by PW_DET behind a run-time flag the extra route cost the
      // register-bound two-component row kernels 10 - 22 %, ncu / A-B r02f)
      if (pw_has_dense(PWV) && p.pump_dense) {
        sv[j] = ((const cpx<T>*)h.pd_now[p.pump == 1 ? 0 : j])[sidx];
      } else if (spre) {
        sv[j] = spre[j];
      } else {
        sv[j] = p.pump_const ? p.S_const[p.pump == 1 ? 0 : j]
                             : ((p.pump == 2 && p.pump_zero[j]) ? mk<T>((T)0, (T)0) : p.S[p.pump == 1 ? 0 : j][sidx]);
      }
      w[j] = f[j] + cmul(h.fnow, sv[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < M; ++j) w[j] = f[j];
  }
  // res = E w
  cpx<T> res[M];
  if (p.vkind == KIND_FULL) {
    // only a scalar-returning nonlinearity (ph[0] == ph[i]) or none can meet a full V (src/kernels.jl:9)
#pragma unroll
    for (int i = 0; i < M; ++i) {
      cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
      for (int j = 0; j < M; ++j) acc = acc + cmul(p.expV[j * M + i][sidx], w[j]);
      res[i] = p.nl ? rotate_m1(acc, ph[0].x, ph[0].y) : acc;
    }
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      cpx<T> e = w[i];
      if (p.vkind == KIND_SCALAR) e = cmul(p.expV[0][sidx], e);
      if (p.vkind == KIND_DIAG) e = cmul(p.expV[i][sidx], e);
      res[i] = p.nl ? rotate_m1(e, ph[i].x, ph[i].y) : e;
    }
  }
  if (p.pump) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      // dense pump: F_next is its own profile, loaded here (no extra live values for the other pump kinds)
      if (pw_has_dense(PWV) && p.pump_dense) sv[i] = ((const cpx<T>*)h.pd_next[p.pump == 1 ? 0 : i])[sidx];
      res[i] = res[i] + cmul(h.fnext, sv[i]);
    }
  }
  if (pw_is_stoch(PWV) && p.noise) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      cpx<T> xi;
      if (p.noise == NOISE_HOST) {
        xi = p.noise_real ? mk<T>(((const T*)h.xi[i])[gidx], (T)0) : ((const cpx<T>*)h.xi[i])[gidx];
      } else {
        xi = normal_from<T>(rnd[i], h.ctr, p.noise_real);
      }
      // -i sqrt(dt) eta xi
      cpx<T> ex;
      if constexpr (PWV == PW_FIELD) ex = cmul(etav[i], xi);
      else ex = cmul(p.eta[i], xi);
      res[i] = res[i] + mk<T>(p.sqrt_dt * ex.y, -p.sqrt_dt * ex.x);
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) f[i] = res[i];
}

// k-space multiply at one point: f <- D[k] (x) f.   planes: [col*M + row] for FULL.
template <typename T, int M>
__device__ __forceinline__ void disp_point(cpx<T> (&f)[M], const cpx<T>* const* D, const int kind, const long long idx) {
  if (kind == KIND_SCALAR) {
    const cpx<T> d = D[0][idx];
#pragma unroll
    for (int i = 0; i < M; ++i) f[i] = cmul(d, f[i]);
  } else if (kind == KIND_DIAG) {
#pragma unroll
    for (int i = 0; i < M; ++i) f[i] = cmul(D[i][idx], f[i]);
  } else if (kind == KIND_FULL) {
    cpx<T> r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      cpx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
      for (int j = 0; j < M; ++j) acc = acc + cmul(D[j * M + i][idx], f[j]);
      r[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) f[i] = r[i];
  }
}

}  // namespace ggp
