// tb200_csprng.cuh -- the reference's CSPRNG operators: ChaCha20 blocks from per-block
// state rows, uniform integers below q, CDT discrete Gaussian, randomised rounding.
//
// Reference: csrc/csprng/cuda/chacha20_cuda.{h,cu}, randint_cuda.cu, discrete_gaussian_cuda.{h,cu},
// randround_cuda.cu; state layout tiberate/rng/csprng/csprng.py:113-178 (one row of 16 int64 per block,
// one 32-bit word each: constants, key, 64-bit counter in words 12-13, nonce).
//
// One thread per state row.  The reference stages the 16 words of a row in shared memory as int64 and
// masks after every addition; here the block lives in sixteen 32-bit registers (uint32 arithmetic is
// the mod 2^32 arithmetic), rows are read with 16-byte loads, only the two counter words are written
// back, and the four samples of a row leave as two 16-byte stores.
#pragma once
#include "tb200_mont.cuh"

#define TB_RNG_MAXQ 128  // reference: LUT_SIZE 128 entries of __constant__ memory (randint_cuda.cu:10-11)
struct TbRngTable {
  u64 v[TB_RNG_MAXQ];
};

namespace tbrng {

__device__ __forceinline__ unsigned rotl(unsigned x, int n) { return (x << n) | (x >> (32 - n)); }

#define TB_QR(a, b, c, d) \
  a += b;                 \
  d = rotl(d ^ a, 16);    \
  c += d;                 \
  b = rotl(b ^ c, 12);    \
  a += b;                 \
  d = rotl(d ^ a, 8);     \
  c += d;                 \
  b = rotl(b ^ c, 7);

// x <- ChaCha20 block of the state row (chacha20_cuda.h:16-39, chacha20_cuda.cu:24-33); steps the row's
// counter by `step` (chacha20_cuda.cu:35-38) when STEP.
template <bool STEP>
__device__ __forceinline__ void block_from_row(i64* row, unsigned (&x)[16], i64 step) {
  unsigned s[16];
  const longlong2* r2 = reinterpret_cast<const longlong2*>(row);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const longlong2 v = r2[i];
    s[2 * i] = (unsigned)v.x;
    s[2 * i + 1] = (unsigned)v.y;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[i];
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    TB_QR(x[0], x[4], x[8], x[12])
    TB_QR(x[1], x[5], x[9], x[13])
    TB_QR(x[2], x[6], x[10], x[14])
    TB_QR(x[3], x[7], x[11], x[15])
    TB_QR(x[0], x[5], x[10], x[15])
    TB_QR(x[1], x[6], x[11], x[12])
    TB_QR(x[2], x[7], x[8], x[13])
    TB_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] += s[i];
  if (STEP) {
    // the stored words are int64: word12 += step; word13 += word12 >> 32; word12 &= 2^32 - 1
    longlong2 c = r2[6];
    c.x += step;
    c.y += c.x >> 32;
    c.x &= 0xffffffffll;
    reinterpret_cast<longlong2*>(row)[6] = c;
  }
}

// floor(X p / 2^128), X = (hi, lo) (randint_cuda.cu:56-81: the same floor through a 32-bit carry chain)
__device__ __forceinline__ i64 scale128(u64 lo, u64 hi, u64 p) {
  const u64 alpha = __umul64hi(p, lo);
  const u64 pl = hi * p;  // low 64 bits of hi * p
  const u64 ph = __umul64hi(hi, p);
  return (i64)(ph + ((pl + alpha) < pl ? 1ull : 0ull));
}

// CDT binary search (discrete_gaussian_cuda.cu:51-100); lut = lows[size] then highs[size]
__device__ __forceinline__ i64 cdt_sample(u64 lo, u64 hi, const u64* lut, int size, int depth) {
  const i64 sign = (i64)(hi & 1ull);
  hi >>= 1;
  int jump = 1, cur = 0, counter = 0;
  for (int j = 0; j < depth; ++j) {
    const u64 yh = lut[counter + cur + size], yl = lut[counter + cur];
    const int ge = (hi > yh) | ((hi == yh) & (lo >= yl));
    cur = 2 * cur + ge;
    counter += jump;
    jump *= 2;
  }
  return (sign * 2 - 1) * (i64)cur;
}

__device__ __forceinline__ u64 combine(unsigned high, unsigned low) { return ((u64)high << 32) | (u64)low; }

}  // namespace tbrng

// chacha20 operator: out[n][16] = blocks (one 32-bit word per int64), states stepped
__global__ void __launch_bounds__(128) k_rng_chacha20(i64* states, i64* out, long n, i64 step) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  unsigned x[16];
  tbrng::block_from_row<true>(states + r * 16, x, step);
  longlong2* o = reinterpret_cast<longlong2*>(out + r * 16);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    longlong2 v;
    v.x = (i64)x[2 * i];
    v.y = (i64)x[2 * i + 1];
    o[i] = v;
  }
}

// randint_fast: states [C][L][16] -> out [C][4L] = floor(X q_c / 2^128) + shift
__global__ void __launch_bounds__(128) k_rng_randint_fast(i64* states, i64* out, long L, TbRngTable q, i64 shift, i64 step) {
  const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (l >= L) return;
  unsigned x[16];
  tbrng::block_from_row<true>(states + ((long)c * L + l) * 16, x, step);
  const u64 p = q.v[c];
  i64 s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    s[j] = tbrng::scale128(tbrng::combine(x[4 * j], x[4 * j + 1]), tbrng::combine(x[4 * j + 2], x[4 * j + 3]), p) + shift;
  longlong2* o = reinterpret_cast<longlong2*>(out + ((long)c * L + l) * 4);
  longlong2 v;
  v.x = s[0];
  v.y = s[1];
  o[0] = v;
  v.x = s[2];
  v.y = s[3];
  o[1] = v;
}

// discrete_gaussian_fast: states [n][16] -> out [4n]
__global__ void __launch_bounds__(128) k_rng_gaussian_fast(i64* states, i64* out, long n, TbRngTable lut, int size, int depth,
                                                           i64 step) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  unsigned x[16];
  tbrng::block_from_row<true>(states + r * 16, x, step);
  i64 s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    s[j] = tbrng::cdt_sample(tbrng::combine(x[4 * j], x[4 * j + 1]), tbrng::combine(x[4 * j + 2], x[4 * j + 3]), lut.v,
                             size, depth);
  longlong2* o = reinterpret_cast<longlong2*>(out + r * 4);
  longlong2 v;
  v.x = s[0];
  v.y = s[1];
  o[0] = v;
  v.x = s[2];
  v.y = s[3];
  o[1] = v;
}

// the two-step variants work in place on random words [C][L][16] (or [n][16]): the sample of words
// 4j..4j+3 replaces word 4j (randint_cuda.cu:94-133, discrete_gaussian_cuda.cu:108-160)
__global__ void __launch_bounds__(128) k_rng_randint_inplace(i64* words, long L, TbRngTable q) {
  const long l = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (l >= L) return;
  i64* w = words + ((long)c * L + l) * 16;
  const u64 p = q.v[c];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    w[4 * j] = tbrng::scale128(tbrng::combine((unsigned)w[4 * j], (unsigned)w[4 * j + 1]),
                               tbrng::combine((unsigned)w[4 * j + 2], (unsigned)w[4 * j + 3]), p);
}
__global__ void __launch_bounds__(128) k_rng_gaussian_inplace(i64* words, long n, TbRngTable lut, int size, int depth) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  i64* w = words + r * 16;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    w[4 * j] = tbrng::cdt_sample(tbrng::combine((unsigned)w[4 * j], (unsigned)w[4 * j + 1]),
                                 tbrng::combine((unsigned)w[4 * j + 2], (unsigned)w[4 * j + 3]), lut.v, size, depth);
}

// randround (randround_cuda.cu:4-36): words[i] <- sign(c) (floor|c| + [words[i] < rn(frac 2^32)])
__global__ void __launch_bounds__(256) k_rng_randround(const double* coef, i64* words, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double c = coef[i];
  const double a = fabs(c), integ = floor(a);
  const i64 ifrac = __double2ll_rn((a - integ) * 4294967296.0);
  const i64 rnd = words[i] < ifrac ? 1 : 0;
  const i64 mag = (i64)integ + rnd;
  words[i] = (__double_as_longlong(c) < 0) ? -mag : mag;  // signbit(c)
}
