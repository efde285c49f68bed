// tb200_mont.cuh -- exact Montgomery arithmetic with R = 2^62 for sm_100a.
//
// The reference evaluates u = (a*b + ((a*b*k) mod R) * q) / R with 31-bit halves and eleven
// 64-bit multiplies (csrc/ops/cuda/mont_scalar_kernel.cuh:9-58).  The same integer is obtained
// here from 64x64->128 products: with x = a*b = xh*2^62 + xl (xl the low 62 bits) and
// s = (x*k) mod 2^62, xl + (s*q mod 2^62) is either 0 (iff xl == 0) or exactly 2^62, hence
//     u = floor(x / 2^62) + floor(s*q / 2^62) + (xl != 0).
// floor(s*q / 2^62) = umulhi(s, 4q) because 4q < 2^64.  Values are bit-identical to the
// reference for every signed a, b it can meet (oracle/ckks_oracle.c orc_mm_halves == orc_mm_closed).
#pragma once
#include <cstdint>

#include "tb200_platform.h"

typedef int64_t i64;
typedef uint64_t u64;

#define TB_MASK62 0x3FFFFFFFFFFFFFFFull

// per-prime constants, one 64-byte record per prime (global index order of CkksConfig.q)
struct __align__(16) TbPrime {
  i64 q;         // prime
  i64 q2;        // 2q
  u64 q4;        // 4q  (umulhi(s, q4) == (s*q) >> 62)
  u64 k;         // -q^-1 mod 2^62   (mont_context.py:38-39)
  i64 Rs;        // R^2 mod q
  i64 Rs_scale;  // R^2 * 2^scale_bits mod q
  i64 Ninv;      // N^-1 * R mod q
  i64 pad;
};

// unsigned x unsigned (both < 2^63): used where the operands are known non-negative.
__device__ __forceinline__ i64 tb_mm_uu(u64 a, u64 b, u64 q4, u64 k) {
  const u64 lo = a * b;
  const u64 hi = __umul64hi(a, b);
  const u64 s = (lo * k) & TB_MASK62;
  const u64 t = __umul64hi(s, q4);
  return (i64)(((hi << 2) | (lo >> 62)) + t + ((lo & TB_MASK62) != 0 ? 1ull : 0ull));
}

// signed x signed: the general closed form (a, b as the reference's int64 operands).
__device__ __forceinline__ i64 tb_mm_ss(i64 a, i64 b, u64 q4, u64 k) {
  const u64 lo = (u64)a * (u64)b;
  const i64 hi = __mul64hi(a, b);
  const u64 s = (lo * k) & TB_MASK62;
  const u64 t = __umul64hi(s, q4);
  const i64 xh = (i64)(((u64)hi << 2) | (lo >> 62));  // floor(a*b / 2^62), arithmetic
  return xh + (i64)t + ((lo & TB_MASK62) != 0 ? 1 : 0);
}

// multiplicand b given pre-shifted (b4 = b << 2, b < 2^62), a unsigned: twiddle multiplies.
// lo4 = low64(a*b4) = 4*((a*b) mod 2^62) so s = (lo4*k) >> 2 and (xl != 0) == (lo4 != 0).
__device__ __forceinline__ i64 tb_mm_u4(u64 a, u64 b4, u64 q4, u64 k) {
  const u64 lo4 = a * b4;
  const u64 hi = __umul64hi(a, b4);
  const u64 s = (lo4 * k) >> 2;
  const u64 t = __umul64hi(s, q4);
  return (i64)(hi + t + (lo4 != 0 ? 1ull : 0ull));
}

// same with a SIGNED a.  The reference's lazy arithmetic lets negative residues travel through its
// butterflies (mont_sub keeps negatives: public/key-switch keys b = e - a*s hold them, and so do the
// key products entering the inverse NTT), and its Montgomery product is exact for signed operands,
// so the butterflies must be too.  b4 < 2^63, hence the signed high product is well defined.
__device__ __forceinline__ i64 tb_mm_s4(i64 a, u64 b4, u64 q4, u64 k) {
  const u64 lo4 = (u64)a * b4;
  const i64 hi = __mul64hi(a, (i64)b4);
  const u64 s = (lo4 * k) >> 2;
  const u64 t = __umul64hi(s, q4);
  return hi + (i64)t + (lo4 != 0 ? 1 : 0);
}

// mont_scalar_kernel.cuh:87-126: MR(a) = (a + ((a*k) mod R) * q) / R, a signed.
// a = ah*2^62 + al (arithmetic): al + (s*q mod 2^62) is 0 or 2^62 as above.
__device__ __forceinline__ i64 tb_mr(i64 a, u64 q4, u64 k) {
  const u64 s = ((u64)a * k) & TB_MASK62;
  const u64 t = __umul64hi(s, q4);
  return (a >> 62) + (i64)t + (((u64)a & TB_MASK62) != 0 ? 1 : 0);
}

// mont_scalar_kernel.cuh:60-85,128-145 (signed compares, exactly as the reference)
__device__ __forceinline__ i64 tb_cs1(i64 x, i64 q) { return (x < q) ? x : x - q; }
__device__ __forceinline__ i64 tb_cs2(i64 x, i64 q2) { return (x < q2) ? x : x - q2; }
__device__ __forceinline__ i64 tb_add(i64 a, i64 b, i64 q2) { return tb_cs2(a + b, q2); }
__device__ __forceinline__ i64 tb_sub(i64 a, i64 b, i64 q2) { return tb_cs2(a - b, q2); }
__device__ __forceinline__ i64 tb_signed(i64 a, i64 q) { return (a <= (q >> 1)) ? a : a - q; }
