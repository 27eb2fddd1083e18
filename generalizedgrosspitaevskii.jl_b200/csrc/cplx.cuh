// Complex arithmetic and in-register radix-R DFT butterflies (R = 2..32) for sm_100a.
// No tensor cores: the FFT is butterfly work on the fp32/fp64 pipes (BASELINE.json north_star).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ggp {

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
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename T>
__device__ __forceinline__ cpx<T> cmulc(cpx<T> a, cpx<T> b) {
  return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T>
__device__ __forceinline__ cpx<T> cscale(cpx<T> a, T s) {
  return mk<T>(a.x * s, a.y * s);
}
template <typename T>
__device__ __forceinline__ T cabs2(cpx<T> a) {
  return a.x * a.x + a.y * a.y;
}
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
      cpx<T> t;
      if (k == 0) {
        t = o[k];
      } else if (4 * k == R) {
        t = mul_si<T, DIR>(o[k]);
      } else if (8 * k == R) {  // (1 + s i)/sqrt2
        const T h = (T)0.70710678118654752440;
        t = DIR < 0 ? mk<T>((o[k].x + o[k].y) * h, (o[k].y - o[k].x) * h)
                    : mk<T>((o[k].x - o[k].y) * h, (o[k].y + o[k].x) * h);
      } else if (8 * k == 3 * R) {  // (-1 + s i)/sqrt2
        const T h = (T)0.70710678118654752440;
        t = DIR < 0 ? mk<T>((o[k].y - o[k].x) * h, -(o[k].x + o[k].y) * h)
                    : mk<T>(-(o[k].x + o[k].y) * h, (o[k].x - o[k].y) * h);
      } else {
        const T c = (T)cos32(k * (32 / R));
        const T s = (T)(DIR * sin32(k * (32 / R)));
        t = mk<T>(o[k].x * c - o[k].y * s, o[k].x * s + o[k].y * c);
      }
      v[k] = e[k] + t;
      v[k + H] = e[k] - t;
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
