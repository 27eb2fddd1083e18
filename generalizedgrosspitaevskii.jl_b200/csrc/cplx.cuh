// Complex arithmetic and in-register radix-R DFT butterflies (R = 2..32) for sm_100a.
// No tensor cores: the FFT is butterfly work on the fp32/fp64 pipes (BASELINE.json north_star).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ggp {

// Programmatic dependent launch (sm_90+): every kernel of the step chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, announces at its very start that the next kernel may
// be scheduled (its CTAs take the SM slots this grid frees while it drains, with their parameters and index
// arithmetic done), and blocks in pdl_wait() -- until the previous grid has completed and flushed -- right
// before its first read of the field.  Without the launch attribute both are no-ops.
#ifdef GGP_NO_PDL_ASM
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// f2: two fp32 values in one 64-bit register pair, operated on with Blackwell's packed fp32x2
// instructions (add/sub/mul/fma .f32x2 -> SASS FADD2/FMUL2/FFMA2).  A cpx<f2> is TWO complex numbers
// (lane 0 = line A, lane 1 = line B) that go through identical arithmetic: the FFT of two lines at
// half the instruction-issue cost (the fp32 kernels are issue-bound, profiles/r01_notes.md).
struct alignas(8) f2 {
  float2 v;
};
__device__ __forceinline__ f2 mkf2(float a, float b) {
  f2 r;
  r.v = make_float2(a, b);
  return r;
}
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return f2{__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return f2{__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<f2*>(&r);
}
__device__ __forceinline__ f2 operator-(f2 a) { return mkf2(0.f, 0.f) - a; }

// a*b + c and c - a*b for every arithmetic type (scalar types: the compiler contracts to FMA itself)
__device__ __forceinline__ float fma_(float a, float b, float c) { return a * b + c; }
__device__ __forceinline__ double fma_(double a, double b, double c) { return a * b + c; }
__device__ __forceinline__ f2 fma_(f2 a, f2 b, f2 c) { return f2{__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ float fnma_(float a, float b, float c) { return c - a * b; }
__device__ __forceinline__ double fnma_(double a, double b, double c) { return c - a * b; }
__device__ __forceinline__ f2 fnma_(f2 a, f2 b, f2 c) {  // c - a*b = fma(a, b, -c) negated ... kept as sub of a product
  return c - a * b;
}
// compile-time constant of type T
template <typename T>
__device__ __forceinline__ T cst(double c) {
  return (T)c;
}
template <>
__device__ __forceinline__ f2 cst<f2>(double c) {
  return mkf2((float)c, (float)c);
}
// low part of a constant: c - fl32(c).  The radix butterflies' internal twiddles (1/sqrt2, cos/sin of
// multiples of pi/16) all round to fp32 values whose modulus is BELOW one (-0.7e-8 ... -2.9e-8), which
// shrinks every transform by ~3e-8 and makes a ComplexF32 run lose norm at ~1e-7 per step -- the dominant,
// systematic part of the fp32 error (tools/c64_error_probe.py).  Multiplying by (hi + lo) with one extra
// FMA per product removes that bias.  fp64 does not need it.
template <typename T>
__device__ __forceinline__ T cst_lo(double c) {
  return (T)0;
}
template <>
__device__ __forceinline__ float cst_lo<float>(double c) {
  return (float)(c - (double)(float)c);
}
template <>
__device__ __forceinline__ f2 cst_lo<f2>(double c) {
  const float l = (float)(c - (double)(float)c);
  return mkf2(l, l);
}
template <typename T>
struct Compensate {
  static constexpr bool value = false;
};
#ifndef GGP_FAST_CONST
template <>
struct Compensate<float> {
  static constexpr bool value = true;
};
template <>
struct Compensate<f2> {
  static constexpr bool value = true;
};
#endif
// x * c for a compile-time constant c, bias-free in fp32
template <typename T>
__device__ __forceinline__ T mulc(T x, double c) {
  if constexpr (Compensate<T>::value) return fma_(x, cst_lo<T>(c), x * cst<T>(c));
  return x * cst<T>(c);
}

// scalar type behind T and number of lines it carries
template <typename T>
struct Lanes {
  using scalar = T;
  static constexpr int N = 1;
};
template <>
struct Lanes<f2> {
  using scalar = float;
  static constexpr int N = 2;
};

template <typename T>
struct alignas(2 * sizeof(T)) cpx {
  T x, y;
};

template <typename T>
__host__ __device__ __forceinline__ cpx<T> mk(T a, T b) {
  cpx<T> r;
  r.x = a;
  r.y = b;
  return r;
}
template <typename T>
__device__ __forceinline__ cpx<T> operator+(cpx<T> a, cpx<T> b) {
  return mk<T>(a.x + b.x, a.y + b.y);
}
template <typename T>
__device__ __forceinline__ cpx<T> operator-(cpx<T> a, cpx<T> b) {
  return mk<T>(a.x - b.x, a.y - b.y);
}
template <typename T>
__device__ __forceinline__ cpx<T> cmul(cpx<T> a, cpx<T> b) {
  return mk<T>(fnma_(a.y, b.y, a.x * b.x), fma_(a.y, b.x, a.x * b.y));
}
// a * conj(b)
template <typename T>
__device__ __forceinline__ cpx<T> cmulc(cpx<T> a, cpx<T> b) {
  return mk<T>(fma_(a.y, b.y, a.x * b.x), fnma_(a.x, b.y, a.y * b.x));
}
template <typename T>
__device__ __forceinline__ cpx<T> cscale(cpx<T> a, T s) {
  return mk<T>(a.x * s, a.y * s);
}
template <typename T>
__device__ __forceinline__ T cabs2(cpx<T> a) {
  return fma_(a.y, a.y, a.x * a.x);
}
// Table loads written as volatile asm: the front end keeps them where they are written, ahead of the arithmetic that
// consumes them, so a batch of exp_D entries is requested together instead of one element at a time (parked strided
// kernel, kernels.cuh).
__device__ __forceinline__ cpx<double> ldg_nc_ordered(const cpx<double>* q) {
  cpx<double> r;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(q));
  return r;
}
__device__ __forceinline__ cpx<float> ldg_nc_ordered(const cpx<float>* q) {
  cpx<float> r;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(q));
  return r;
}

// Inter-pass twiddle factors and the factors of a separable exp_D.  Default: a plain complex number of the
// plan's precision.  -DGGP_SPLIT_TWIDDLES (make SPLIT_TW=1): fp32 plans keep the double-precision value as
// hi + lo (one 16-byte entry, one LDG.128) and multiply with four extra FMAs.  Measured (tools/
// c64_error_probe.py, 256^2, 1000 steps): relative L2 distance to the fp64 run 5.0e-5 -> 4.1e-5 only -- the
// fp32 error that grows linearly with the step count is the round-off of the butterflies themselves, which
// repeats from step to step because the field changes slowly -- at the price of a 1.6x slower strided
// kernel (doubled twiddle traffic through L1).  Hence off by default; profiles/r01_notes.md.
template <typename T>
struct TwT {
  using type = cpx<T>;
  static constexpr bool split = false;
  static __device__ __forceinline__ cpx<T> mul(cpx<T> a, const type w) { return cmul(a, w); }
  static __host__ type make(long double c, long double s) { return mk<T>((T)c, (T)s); }
};
#ifdef GGP_SPLIT_TWIDDLES
template <>
struct TwT<float> {
  using type = float4;  // (hi.re, hi.im, lo.re, lo.im)
  static constexpr bool split = true;
  static __device__ __forceinline__ cpx<float> mul(cpx<float> a, const float4 w) {
    return mk<float>(fmaf(a.x, w.x, fmaf(-a.y, w.y, fmaf(a.x, w.z, -a.y * w.w))),
                     fmaf(a.x, w.y, fmaf(a.y, w.x, fmaf(a.x, w.w, a.y * w.z))));
  }
  static __host__ float4 make(long double c, long double s) {
    const float ch = (float)c, sh = (float)s;
    return make_float4(ch, sh, (float)(c - (long double)ch), (float)(s - (long double)sh));
  }
};
#endif

// multiply by  s*i  where s = DIR (DIR=-1: forward transform e^{-i..}, DIR=+1: inverse)
template <typename T, int DIR>
__device__ __forceinline__ cpx<T> mul_si(cpx<T> a) {
  return DIR < 0 ? mk<T>(a.y, -a.x) : mk<T>(-a.y, a.x);
}

// cos(2*pi*k/32), exact to double precision, usable in constant expressions
__host__ __device__ constexpr double cos32(int k) {
  k &= 31;
  if (k > 16) k = 32 - k;
  switch (k) {
    case 0: return 1.0;
    case 1: return 0.98078528040323044913;
    case 2: return 0.92387953251128675613;
    case 3: return 0.83146961230254523708;
    case 4: return 0.70710678118654752440;
    case 5: return 0.55557023301960222474;
    case 6: return 0.38268343236508977173;
    case 7: return 0.19509032201612826785;
    case 8: return 0.0;
    case 9: return -0.19509032201612826785;
    case 10: return -0.38268343236508977173;
    case 11: return -0.55557023301960222474;
    case 12: return -0.70710678118654752440;
    case 13: return -0.83146961230254523708;
    case 14: return -0.92387953251128675613;
    case 15: return -0.98078528040323044913;
    default: return -1.0;
  }
}
__host__ __device__ constexpr double sin32(int k) { return cos32(k - 8); }

// In-register DFT of R points, natural order in and out:  V[r] = sum_j v[j] * exp(DIR*2*pi*i*j*r/R).
// Recursive decimation in time; after full unrolling every twiddle is an immediate.
template <typename T, int R, int DIR>
struct Dft {
  static __device__ __forceinline__ void run(cpx<T> (&v)[R]) {
    constexpr int H = R / 2;
    cpx<T> e[H], o[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      e[j] = v[2 * j];
      o[j] = v[2 * j + 1];
    }
    Dft<T, H, DIR>::run(e);
    Dft<T, H, DIR>::run(o);
#pragma unroll
    for (int k = 0; k < H; ++k) {
      // v[k] = e + w^k o,  v[k+H] = e - w^k o,  w = exp(DIR*2*pi*i/R); the trivial rotations are written out
      // so that no explicit negation is ever needed (a negation is a real instruction for the packed type)
      const cpx<T> ek = e[k], ok = o[k];
      if (k == 0) {
        v[k] = ek + ok;
        v[k + H] = ek - ok;
      } else if (4 * k == R) {  // w^k = DIR*i :  t = (-DIR*o.y, DIR*o.x)
        if (DIR < 0) {
          v[k] = mk<T>(ek.x + ok.y, ek.y - ok.x);
          v[k + H] = mk<T>(ek.x - ok.y, ek.y + ok.x);
        } else {
          v[k] = mk<T>(ek.x - ok.y, ek.y + ok.x);
          v[k + H] = mk<T>(ek.x + ok.y, ek.y - ok.x);
        }
      } else if (8 * k == R) {  // (1 + DIR*i)/sqrt2
        const T p = mulc<T>(ok.x + ok.y, 0.70710678118654752440), q = mulc<T>(ok.y - ok.x, 0.70710678118654752440);
        if (DIR < 0) {  // t = (p, q)
          v[k] = mk<T>(ek.x + p, ek.y + q);
          v[k + H] = mk<T>(ek.x - p, ek.y - q);
        } else {        // t = (-q, p)
          v[k] = mk<T>(ek.x - q, ek.y + p);
          v[k + H] = mk<T>(ek.x + q, ek.y - p);
        }
      } else if (8 * k == 3 * R) {  // (-1 + DIR*i)/sqrt2
        const T p = mulc<T>(ok.x + ok.y, 0.70710678118654752440), q = mulc<T>(ok.y - ok.x, 0.70710678118654752440);
        if (DIR < 0) {  // t = (q, -p)
          v[k] = mk<T>(ek.x + q, ek.y - p);
          v[k + H] = mk<T>(ek.x - q, ek.y + p);
        } else {        // t = (-p, -q)
          v[k] = mk<T>(ek.x - p, ek.y - q);
          v[k + H] = mk<T>(ek.x + p, ek.y + q);
        }
      } else {
        const double cd = cos32(k * (32 / R)), sd = DIR * sin32(k * (32 / R));
        const T c = cst<T>(cd), sn = cst<T>(sd), msn = cst<T>(-sd);
        cpx<T> t;
        if constexpr (Compensate<T>::value) {
          const T cl = cst_lo<T>(cd), sl = cst_lo<T>(sd), msl = cst_lo<T>(-sd);
          t = mk<T>(fma_(ok.x, c, fma_(ok.y, msn, fma_(ok.y, msl, ok.x * cl))),
                    fma_(ok.y, c, fma_(ok.x, sn, fma_(ok.x, sl, ok.y * cl))));
        } else {
          t = mk<T>(fma_(ok.y, msn, ok.x * c), fma_(ok.y, c, ok.x * sn));
        }
        v[k] = ek + t;
        v[k + H] = ek - t;
      }
    }
  }
};
template <typename T, int DIR>
struct Dft<T, 1, DIR> {
  static __device__ __forceinline__ void run(cpx<T> (&)[1]) {}
};
template <typename T, int DIR>
struct Dft<T, 2, DIR> {
  static __device__ __forceinline__ void run(cpx<T> (&v)[2]) {
    cpx<T> a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
  }
};

}  // namespace ggp
