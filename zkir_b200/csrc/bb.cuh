// BabyBear (p = 2^31 - 2^27 + 1) arithmetic for sm_100a kernels: Montgomery form, R = 2^32, values kept in [0, p).
// Constants were checked numerically.  The reference's only field type is Mersenne31
// (zkir-spec/src/field.rs:23-189), which has 2-adicity 1 and cannot carry a radix-2 NTT; its `#[repr(transparent)]`
// canonical-u32 convention is what the C ABI keeps on the wire.
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

#define BB_P 0x78000001u
#define BB_PINV 0x88000001u /* p^-1 mod 2^32 */
#define BB_ONE 0x0ffffffeu  /* R mod p */
#define BB_R2 0x45dddde3u   /* R^2 mod p */

#ifdef __CUDACC__
#define BB_HD __host__ __device__ __forceinline__
#define BB_ALIGN16 __align__(16)
#else
#define BB_HD inline
#define BB_ALIGN16 alignas(16)
#endif

// canonical -> Montgomery at compile time / on the host
BB_HD constexpr u32 bb_to_mont_c(u32 a) { return (u32)((((u64)a) << 32) % BB_P); }

BB_HD u32 bb_add(u32 a, u32 b) {
  u32 s = a + b, t = s - BB_P;
  return s < t ? s : t;  // min: if s < p the subtraction wraps to a huge value
}
BB_HD u32 bb_sub(u32 a, u32 b) {
  u32 d = a - b, t = d + BB_P;
  return d < t ? d : t;
}
BB_HD u32 bb_neg(u32 a) { return a ? BB_P - a : 0; }
BB_HD u32 bb_dbl(u32 a) { return bb_add(a, a); }
// Montgomery product a*b/R mod p, inputs in [0,p) (one of them may be any u32), output in [0,p)
BB_HD u32 bb_mul(u32 a, u32 b) {
  u64 x = (u64)a * b;
  u32 lo = (u32)x, hi = (u32)(x >> 32);
  u32 m = lo * BB_PINV;
#ifdef __CUDA_ARCH__
  u32 t = __umulhi(m, BB_P);
#else
  u32 t = (u32)(((u64)m * BB_P) >> 32);
#endif
  u32 r = hi - t, r2 = r + BB_P;
  return r < r2 ? r : r2;
}
BB_HD u32 bb_sqr(u32 a) { return bb_mul(a, a); }
BB_HD u32 bb_to_mont(u32 a) { return bb_mul(a, BB_R2); }
BB_HD u32 bb_from_mont(u32 a) { return bb_mul(a, 1u); }
BB_HD u32 bb_pow(u32 a, u64 e) {
  u32 r = BB_ONE;
  while (e) { if (e & 1) r = bb_mul(r, a); a = bb_sqr(a); e >>= 1; }
  return r;
}
BB_HD u32 bb_inv(u32 a) { return bb_pow(a, BB_P - 2); }

// a/2
BB_HD u32 bb_halve(u32 a) { return (a >> 1) + ((a & 1u) ? (BB_P + 1) / 2 : 0u); }

// ---- compile-time helpers on canonical values, and Shoup multiplication by a constant
constexpr u32 c_mul(u32 a, u32 b) { return (u32)(((u64)a * b) % BB_P); }
constexpr u32 c_pow(u32 a, u64 e) {
  u32 r = 1;
  while (e) { if (e & 1) r = c_mul(r, a); a = c_mul(a, a); e >>= 1; }
  return r;
}
constexpr u32 c_shoup(u32 w) { return (u32)((((u64)w) << 32) / BB_P); }  // floor(w * 2^32 / p)
#ifdef __CUDACC__
// x * w mod p for a constant w (canonical) with wq = c_shoup(w); x may be ANY u32 (canonical or Montgomery
// representative, reduced or not); result in [0, p).  mul.hi + 2 mul.lo: 4 issue slots of the integer-multiply pipe
// against 5 for a Montgomery product (profiles/r01_microbench_b200.txt).
__device__ __forceinline__ u32 shoup_mul(u32 x, u32 w, u32 wq) {
  const u32 q = __umulhi(x, wq);
  const u32 r = x * w - q * BB_P;  // in [0, 2p)
  return min(r, r - BB_P);
}
#endif

// wrapper with operators, used to instantiate the generated AIR
struct Fm {
  u32 v;
  BB_HD Fm() : v(0) {}
  BB_HD explicit Fm(u32 x) : v(x) {}
};
BB_HD Fm operator+(Fm a, Fm b) { return Fm(bb_add(a.v, b.v)); }
BB_HD Fm operator-(Fm a, Fm b) { return Fm(bb_sub(a.v, b.v)); }
BB_HD Fm operator*(Fm a, Fm b) { return Fm(bb_mul(a.v, b.v)); }

// ---- F_p[X]/(X^4 - 11), coefficients in Montgomery form
struct BB_ALIGN16 E4 {
  u32 c[4];
};
#define BB_W11 bb_to_mont_c(11u)
BB_HD E4 e4_zero() { E4 r; r.c[0] = r.c[1] = r.c[2] = r.c[3] = 0; return r; }
BB_HD E4 e4_one() { E4 r = e4_zero(); r.c[0] = BB_ONE; return r; }
BB_HD E4 e4_from_base(u32 a) { E4 r = e4_zero(); r.c[0] = a; return r; }
BB_HD E4 e4_add(E4 a, E4 b) { E4 r; for (int i = 0; i < 4; i++) r.c[i] = bb_add(a.c[i], b.c[i]); return r; }
BB_HD E4 e4_sub(E4 a, E4 b) { E4 r; for (int i = 0; i < 4; i++) r.c[i] = bb_sub(a.c[i], b.c[i]); return r; }
BB_HD E4 e4_mulb(E4 a, u32 b) { E4 r; for (int i = 0; i < 4; i++) r.c[i] = bb_mul(a.c[i], b); return r; }
BB_HD E4 e4_mul(E4 a, E4 b) {
  E4 r;
  u32 t0 = bb_add(bb_add(bb_mul(a.c[1], b.c[3]), bb_mul(a.c[2], b.c[2])), bb_mul(a.c[3], b.c[1]));
  u32 t1 = bb_add(bb_mul(a.c[2], b.c[3]), bb_mul(a.c[3], b.c[2]));
  u32 t2 = bb_mul(a.c[3], b.c[3]);
  r.c[0] = bb_add(bb_mul(a.c[0], b.c[0]), bb_mul(t0, BB_W11));
  r.c[1] = bb_add(bb_add(bb_mul(a.c[0], b.c[1]), bb_mul(a.c[1], b.c[0])), bb_mul(t1, BB_W11));
  r.c[2] = bb_add(bb_add(bb_add(bb_mul(a.c[0], b.c[2]), bb_mul(a.c[1], b.c[1])), bb_mul(a.c[2], b.c[0])), bb_mul(t2, BB_W11));
  r.c[3] = bb_add(bb_add(bb_mul(a.c[0], b.c[3]), bb_mul(a.c[1], b.c[2])), bb_add(bb_mul(a.c[2], b.c[1]), bb_mul(a.c[3], b.c[0])));
  return r;
}
// inverse through the tower F_p < F_p[Z]/(Z^2-11) < F_p[X]/(X^2-Z):  a = A + X*B,  a^-1 = (A - X*B) / (A^2 - Z*B^2)
BB_HD E4 e4_inv(E4 a) {
  const u32 W = BB_W11;
  u32 A0 = a.c[0], A1 = a.c[2], B0 = a.c[1], B1 = a.c[3];  // A = A0 + A1 Z, B = B0 + B1 Z
  // A^2 = (A0^2 + 11 A1^2) + 2 A0 A1 Z ;  B^2 likewise ;  Z*B^2 = 11*B2_1 + B2_0 Z
  u32 a2_0 = bb_add(bb_sqr(A0), bb_mul(W, bb_sqr(A1))), a2_1 = bb_dbl(bb_mul(A0, A1));
  u32 b2_0 = bb_add(bb_sqr(B0), bb_mul(W, bb_sqr(B1))), b2_1 = bb_dbl(bb_mul(B0, B1));
  u32 d0 = bb_sub(a2_0, bb_mul(W, b2_1)), d1 = bb_sub(a2_1, b2_0);  // D = d0 + d1 Z
  u32 nrm = bb_sub(bb_sqr(d0), bb_mul(W, bb_sqr(d1)));               // D * conj(D)
  u32 ni = bb_inv(nrm);
  u32 i0 = bb_mul(d0, ni), i1 = bb_neg(bb_mul(d1, ni));              // D^-1 = i0 + i1 Z
  // (A - X B) * (i0 + i1 Z):  A*Dinv = (A0 i0 + 11 A1 i1) + (A0 i1 + A1 i0) Z ; same for B
  E4 r;
  r.c[0] = bb_add(bb_mul(A0, i0), bb_mul(W, bb_mul(A1, i1)));
  r.c[2] = bb_add(bb_mul(A0, i1), bb_mul(A1, i0));
  r.c[1] = bb_neg(bb_add(bb_mul(B0, i0), bb_mul(W, bb_mul(B1, i1))));
  r.c[3] = bb_neg(bb_add(bb_mul(B0, i1), bb_mul(B1, i0)));
  return r;
}

// ext4 wrapper with operators, used to instantiate the generated AIR (X type of air_generated.h)
struct Xm { E4 v; };
BB_HD Xm operator+(Xm a, Xm b) { Xm r; r.v = e4_add(a.v, b.v); return r; }
BB_HD Xm operator-(Xm a, Xm b) { Xm r; r.v = e4_sub(a.v, b.v); return r; }
BB_HD Xm operator*(Xm a, Xm b) { Xm r; r.v = e4_mul(a.v, b.v); return r; }
BB_HD Xm operator*(Xm a, Fm b) { Xm r; r.v = e4_mulb(a.v, b.v); return r; }

// ---- lazy ext4 accumulator for inner products  sum_j x_j * v_j  (x_j in E4, v_j in F_p, all < p, any representation).
// Each coordinate is a 64-bit integer kept below p*2^32: a product (< p^2) is added with one mad.wide, and after at
// most TWO products the high word is reduced by one conditional subtraction of p (2*p^2 + p*2^32 < 2^64, and
// 1.94*p*2^32 - p*2^32 < p*2^32).  One Montgomery reduction at the end returns the same value a chain of
// bb_mul/bb_add would: 1.5 instructions per term instead of 7, 2 integer-multiply issue slots instead of 5.
struct Acc4 {
  u64 c[4];
};
BB_HD Acc4 acc4_zero() { Acc4 a; a.c[0] = a.c[1] = a.c[2] = a.c[3] = 0; return a; }
BB_HD void acc4_fix(Acc4& a) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    u32 hi = (u32)(a.c[k] >> 32), h2 = hi - BB_P;
    hi = hi < h2 ? hi : h2;
    a.c[k] = (a.c[k] & 0xffffffffull) | ((u64)hi << 32);
  }
}
BB_HD void acc4_mac(Acc4& a, const E4& x, u32 v) {  // caller: acc4_fix after every second call at the latest
#pragma unroll
  for (int k = 0; k < 4; k++) a.c[k] += (u64)x.c[k] * v;
}
BB_HD u32 bb_redc64(u64 x) {  // x < p * 2^32  ->  x / 2^32 mod p
  u32 lo = (u32)x, hi = (u32)(x >> 32);
  u32 m = lo * BB_PINV;
#ifdef __CUDA_ARCH__
  u32 t = __umulhi(m, BB_P);
#else
  u32 t = (u32)(((u64)m * BB_P) >> 32);
#endif
  u32 r = hi - t, r2 = r + BB_P;
  return r < r2 ? r : r2;
}
BB_HD E4 acc4_finish(Acc4 a) {
  acc4_fix(a);
  E4 r;
#pragma unroll
  for (int k = 0; k < 4; k++) r.c[k] = bb_redc64(a.c[k]);
  return r;
}

