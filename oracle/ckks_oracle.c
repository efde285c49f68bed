/*
 * ckks_oracle.c -- CPU restatement of the tiberate-fhe CKKS RNS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, and there only as the checker or the reported
 * CPU baseline, never as the thing shipped.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout).  The reference has no CPU path and its tests hold no golden
 * vectors for this path, so the pins are
 *   (1) the big-int identities in tests/test_oracle_*.py and
 *   (2) fixtures produced by the reference's own CUDA extension on a B200
 *       (tests/golden/make_ref_golden.py -> tests/golden/ref_*.npz).
 *
 * All arithmetic is signed 64-bit with two's-complement wrap-around, exactly like the
 * int64 instantiation of the reference kernels (R = 2^62, halves of 31 bits).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* thread count of the parallel loops (bench.py's CPU arm: torchrun caps OMP_NUM_THREADS at 1) */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

typedef int64_t i64;
typedef uint64_t u64;
typedef __int128 i128;
typedef unsigned __int128 u128;

#define NBITS 62
#define HALF 31
#define FB_MASK ((((i64)1) << NBITS) - 1)
#define LB_MASK ((((i64)1) << HALF) - 1)

/* wrap-around helpers (signed overflow is UB in C; do it in u64) */
static inline i64 wmul(i64 a, i64 b) { return (i64)((u64)a * (u64)b); }
static inline i64 wadd(i64 a, i64 b) { return (i64)((u64)a + (u64)b); }
static inline i64 wsub(i64 a, i64 b) { return (i64)((u64)a - (u64)b); }
static inline i64 wshl(i64 a, int s) { return (i64)((u64)a << s); }

/* ---------------------------------------------------------------------------------
 * Scalar core.  csrc/ops/cuda/mont_scalar_kernel.cuh:9-58 (mont_mult_scalar):
 * operands split into 31-bit halves, s = (x*k) mod R, u = (x + s*q) / R.
 * "halves" is the literal sequence of int64 operations; "closed" is the exact
 * big-integer value it equals.  tests/test_oracle_scalar.py checks they agree.
 * --------------------------------------------------------------------------------- */
i64 orc_mm_halves(i64 a, i64 b, i64 q, i64 k) {
  const i64 ql = q & LB_MASK, qh = q >> HALF, kl = k & LB_MASK, kh = k >> HALF;
  const i64 al = a & LB_MASK, ah = a >> HALF;
  const i64 bl = b & LB_MASK, bh = b >> HALF;
  const i64 alpha = wmul(ah, bh);
  const i64 beta = wadd(wmul(ah, bl), wmul(al, bh));
  const i64 gamma = wmul(al, bl);
  const i64 gammal = gamma & LB_MASK, gammah = gamma >> HALF;
  const i64 betal = beta & LB_MASK, betah = beta >> HALF;
  i64 upper = wmul(gammal, kh);
  upper = wadd(upper, wmul(wadd(gammah, betal), kl));
  upper = wshl(upper, HALF);
  i64 s = wadd(upper, wmul(gammal, kl));
  s &= FB_MASK;
  const i64 sl = s & LB_MASK, sh = s >> HALF;
  const i64 sqb = wadd(wmul(sh, ql), wmul(sl, qh));
  const i64 sqbl = sqb & LB_MASK, sqbh = sqb >> HALF;
  i64 carry = wadd(gamma, wmul(sl, ql)) >> HALF;
  carry = wadd(wadd(carry, betal), sqbl) >> HALF;
  return wadd(wadd(wadd(wadd(alpha, betah), sqbh), carry), wmul(sh, qh));
}

i64 orc_mm_closed(i64 a, i64 b, i64 q, i64 k) {
  const i128 x = (i128)a * (i128)b;
  const u64 s = ((u64)x * (u64)k) & (u64)FB_MASK;
  const i128 t = x + (i128)((u128)s * (u128)(u64)q);
  return (i64)(t >> NBITS);
}

/* mont_scalar_kernel.cuh:87-126 (mont_reduce_scalar) */
i64 orc_mr_halves(i64 a, i64 q, i64 k) {
  const i64 ql = q & LB_MASK, qh = q >> HALF, kl = k & LB_MASK, kh = k >> HALF;
  const i64 xl = a & LB_MASK, xh = a >> HALF;
  const i64 xkb = wadd(wmul(xh, kl), wmul(xl, kh));
  i64 s = wadd(wshl(xkb, HALF), wmul(xl, kl));
  s &= FB_MASK;
  const i64 sl = s & LB_MASK, sh = s >> HALF;
  const i64 sqb = wadd(wmul(sh, ql), wmul(sl, qh));
  const i64 sqbl = sqb & LB_MASK, sqbh = sqb >> HALF;
  i64 carry = wadd(a, wmul(sl, ql)) >> HALF;
  carry = wadd(carry, sqbl) >> HALF;
  return wadd(wadd(sqbh, carry), wmul(sh, qh));
}

i64 orc_mr_closed(i64 a, i64 q, i64 k) {
  const u64 s = ((u64)a * (u64)k) & (u64)FB_MASK;
  const i128 t = (i128)a + (i128)((u128)s * (u128)(u64)q);
  return (i64)(t >> NBITS);
}

/* mont_scalar_kernel.cuh:60-85,128-145 */
static inline i64 cs1(i64 x, i64 q) { return (x < q) ? x : wsub(x, q); }        /* reduce_2q  */
static inline i64 cs2(i64 x, i64 q2) { return (x < q2) ? x : wsub(x, q2); }     /* add/sub tail */
static inline i64 madd(i64 a, i64 b, i64 q2) { return cs2(wadd(a, b), q2); }
static inline i64 msub(i64 a, i64 b, i64 q2) { return cs2(wsub(a, b), q2); }
static inline i64 msigned(i64 a, i64 q) { return (a <= (q >> 1)) ? a : wsub(a, q); }

#define MM orc_mm_halves
#define MR orc_mr_halves

/* vector forms of the scalar core for the tests */
void orc_vec_mm(i64 *out, const i64 *a, const i64 *b, size_t n, i64 q, i64 k, int closed) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i)
    out[i] = closed ? orc_mm_closed(a[i], b[i], q, k) : orc_mm_halves(a[i], b[i], q, k);
}
void orc_vec_mr(i64 *out, const i64 *a, size_t n, i64 q, i64 k, int closed) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i)
    out[i] = closed ? orc_mr_closed(a[i], q, k) : orc_mr_halves(a[i], q, k);
}

/* ---------------------------------------------------------------------------------
 * Pointwise family.  a/b/out are [C][N] row-major, q[C]/k[C] the constants of each
 * row's prime (the caller resolves the reference's right-aligned const-pool indexing,
 * SURVEY.md appendix A.0).  mont_cuda.cu:11-794, mont_extra_cuda.cu:12-368.
 * --------------------------------------------------------------------------------- */
#define ROWLOOP                                             \
  _Pragma("omp parallel for collapse(2) schedule(static)") \
  for (long i = 0; i < C; ++i)                              \
    for (long j = 0; j < N; ++j)

/* mont_cuda.cu:11-36 */
void orc_mont_mult(i64 *out, const i64 *a, const i64 *b, long C, long N, const i64 *q, const i64 *k) {
  ROWLOOP out[i * N + j] = MM(a[i * N + j], b[i * N + j], q[i], k[i]);
}
/* mont_cuda.cu:85-109 (enter_scalar), :151-173 (enter_Rs), :211-234 (Rs_scale), :272-294
 * (legacy): all are a <- MM(a, s_i) with a per-row scalar s_i. */
void orc_mont_enter_scalar(i64 *a, const i64 *s, long C, long N, const i64 *q, const i64 *k) {
  ROWLOOP a[i * N + j] = MM(a[i * N + j], s[i], q[i], k[i]);
}
/* mont_cuda.cu:345-364 */
void orc_mont_reduce(i64 *a, long C, long N, const i64 *q, const i64 *k) {
  ROWLOOP a[i * N + j] = MR(a[i * N + j], q[i], k[i]);
}
/* mont_cuda.cu:402-422 */
void orc_reduce_2q(i64 *a, long C, long N, const i64 *q) {
  ROWLOOP a[i * N + j] = cs1(a[i * N + j], q[i]);
}
/* mont_cuda.cu:642-661 */
void orc_make_signed(i64 *a, long C, long N, const i64 *q) {
  ROWLOOP a[i * N + j] = msigned(a[i * N + j], q[i]);
}
/* mont_cuda.cu:693-712 */
void orc_make_unsigned(i64 *a, long C, long N, const i64 *q) {
  ROWLOOP a[i * N + j] = wadd(a[i * N + j], q[i]);
}
/* mont_cuda.cu:453-476 / :519-537 (legacy: same value with an explicit 2q tensor) */
void orc_mont_add(i64 *out, const i64 *a, const i64 *b, long C, long N, const i64 *q) {
  ROWLOOP out[i * N + j] = madd(a[i * N + j], b[i * N + j], 2 * q[i]);
}
/* mont_cuda.cu:577-600 */
void orc_mont_sub(i64 *out, const i64 *a, const i64 *b, long C, long N, const i64 *q) {
  ROWLOOP out[i * N + j] = msub(a[i * N + j], b[i * N + j], 2 * q[i]);
}
/* mont_cuda.cu:744-760: out[i][j] = a[j] + q_i */
void orc_tile_unsigned(i64 *out, const i64 *a, long C, long N, const i64 *q) {
  ROWLOOP out[i * N + j] = wadd(a[j], q[i]);
}
/* mont_extra_cuda.cu:161-189 / :230-258 */
void orc_mont_add_reduce_2q(i64 *out, const i64 *a, const i64 *b, long C, long N, const i64 *q) {
  ROWLOOP out[i * N + j] = cs1(madd(a[i * N + j], b[i * N + j], 2 * q[i]), q[i]);
}
void orc_mont_sub_reduce_2q(i64 *out, const i64 *a, const i64 *b, long C, long N, const i64 *q) {
  ROWLOOP out[i * N + j] = cs1(msub(a[i * N + j], b[i * N + j], 2 * q[i]), q[i]);
}
/* mont_extra_cuda.cu:83-118: acc = 0; acc = CS2(acc + in[k]) for k ascending; in is [K][C][N] */
void orc_mont_reduce_add_many_3d(i64 *out, const i64 *in, long K, long C, long N, const i64 *q) {
  ROWLOOP {
    i64 acc = 0;
    for (long kk = 0; kk < K; ++kk) acc = madd(acc, in[(kk * C + i) * N + j], 2 * q[i]);
    out[i * N + j] = acc;
  }
}
/* mont_extra_cuda.cu:12-42: pairs are added first, then folded into acc */
void orc_mont_add_many_3d(i64 *out, const i64 *in, long K, long C, long N, const i64 *q) {
  ROWLOOP {
    i64 acc = 0;
    const i64 q2 = 2 * q[i];
    for (long kk = 0; kk < K / 2; ++kk)
      acc = madd(acc, madd(in[((2 * kk) * C + i) * N + j], in[((2 * kk + 1) * C + i) * N + j], q2), q2);
    if (K & 1) acc = madd(acc, in[((K - 1) * C + i) * N + j], q2);
    out[i * N + j] = acc;
  }
}
/* he_fused_cuda.cu:12-51: CS1(MR(CS2(MM(ct, Rs) + pt))) */
void orc_pc_add_fused(i64 *out, const i64 *ct, const i64 *pt, long C, long N, const i64 *q, const i64 *k,
                      const i64 *Rs) {
  ROWLOOP {
    i64 x = MM(ct[i * N + j], Rs[i], q[i], k[i]);
    x = madd(x, pt[i * N + j], 2 * q[i]);
    x = MR(x, q[i], k[i]);
    out[i * N + j] = cs1(x, q[i]);
  }
}

/* ---------------------------------------------------------------------------------
 * NTT.  psi/ipsi are the COMPACT bit-reversed tables [C][N] (entry m+i is the twiddle of
 * group i of the stage with m groups), i.e. the reference's expanded [P][logN][N/2]
 * tensors de-duplicated (tiberate/context/ntt_context.py:88-141,200-205).
 * --------------------------------------------------------------------------------- */

/* ntt_radix2_cuda.cu:10-47 applied for level = 0..logN-1 (:75-78). Cooley-Tukey,
 * natural in -> bit-reversed out, V = MM(S, O); a[e] = CS2(U+V); a[o] = CS2(U+2q-V). */
void orc_ntt_stages(i64 *a, long C, long N, const i64 *psi, const i64 *q, const i64 *k) {
#pragma omp parallel for schedule(dynamic, 1)
  for (long i = 0; i < C; ++i) {
    i64 *x = a + i * N;
    const i64 *w = psi + i * N;
    const i64 qi = q[i], ki = k[i], q2 = 2 * qi;
    long t = N;
    for (long m = 1; m < N; m <<= 1) {
      t >>= 1;
      for (long g = 0; g < m; ++g) {
        const i64 S = w[m + g];
        i64 *lo = x + 2 * g * t, *hi = lo + t;
        for (long j = 0; j < t; ++j) {
          const i64 U = lo[j];
          const i64 V = MM(S, hi[j], qi, ki);
          lo[j] = cs2(wadd(U, V), q2);
          hi[j] = cs2(wsub(wadd(U, q2), V), q2);
        }
      }
    }
  }
}

/* intt_radix2_cuda.cu:10-48 applied for level = 0..logN-1 (:82-85). Gentleman-Sande,
 * a[e] = CS2(U+V); a[o] = MM(S, CS2(U+2q-V)).  No N^-1 here (see epilogues). */
void orc_intt_stages(i64 *a, long C, long N, const i64 *ipsi, const i64 *q, const i64 *k) {
#pragma omp parallel for schedule(dynamic, 1)
  for (long i = 0; i < C; ++i) {
    i64 *x = a + i * N;
    const i64 *w = ipsi + i * N;
    const i64 qi = q[i], ki = k[i], q2 = 2 * qi;
    long t = 1;
    for (long h = N >> 1; h >= 1; h >>= 1) {
      for (long g = 0; g < h; ++g) {
        const i64 S = w[h + g];
        i64 *lo = x + 2 * g * t, *hi = lo + t;
        for (long j = 0; j < t; ++j) {
          const i64 U = lo[j], V = hi[j];
          const i64 O = cs2(wsub(wadd(U, q2), V), q2);
          hi[j] = MM(S, O, qi, ki);
          lo[j] = cs2(wadd(U, V), q2);
        }
      }
      t <<= 1;
    }
  }
}

/* mont_used_in_ntt.cuh:36-59 (mode 0: MM(x,Ninv)), :134-170 (1: + MR), :172-206 (2: + CS1),
 * :208-248 (3: + make_signed) */
void orc_intt_epilogue(i64 *a, long C, long N, const i64 *Ninv, const i64 *q, const i64 *k, int mode) {
  ROWLOOP {
    i64 x = MM(a[i * N + j], Ninv[i], q[i], k[i]);
    if (mode >= 1) x = MR(x, q[i], k[i]);
    if (mode >= 2) x = cs1(x, q[i]);
    if (mode >= 3) x = msigned(x, q[i]);
    a[i * N + j] = x;
  }
}

/* ---------------------------------------------------------------------------------
 * Fused HE kernels (he_fused_cuda.cu).
 * --------------------------------------------------------------------------------- */

/* he_fused_cuda.cu:99-142 (exact) / :190-230 (non-exact); in place on the kept rows. */
void orc_rescale(i64 *a, const i64 *scales, const i64 *rescaler, i64 round_at, int exact, long C, long N,
                 const i64 *q, const i64 *k) {
  ROWLOOP {
    const i64 r = rescaler[j];
    i64 x = wsub(a[i * N + j], r);
    x = MM(x, scales[i], q[i], k[i]);
    if (exact) x = wadd(x, (r > round_at) ? 1 : 0);
    a[i * N + j] = cs1(x, q[i]);
  }
}

/* he_fused_cuda.cu:276-312: out[i][j] = MM(state[0][j], Rs_i) (+)_{k>=1} MM(state[k][j], l_enter[k-1][i]),
 * (+) = CS2 add.  l_enter is [alpha-1][C] already offset to this level. */
void orc_extend(i64 *out, const i64 *state, long alpha, const i64 *l_enter, long C, long N, const i64 *Rs,
                const i64 *q, const i64 *k) {
  ROWLOOP {
    i64 x = MM(state[j], Rs[i], q[i], k[i]);
    for (long kk = 0; kk + 1 < alpha; ++kk) {
      const i64 y = MM(state[(kk + 1) * N + j], l_enter[kk * C + i], q[i], k[i]);
      x = madd(x, y, 2 * q[i]);
    }
    out[i * N + j] = x;
  }
}

/* he_fused_cuda.cu:361-391: out[i][perm[j] % N] = CS1(sign * a[i][j] + q_i) */
void orc_codec_rotate(i64 *out, const i64 *a, const i64 *perm, long C, long N, const i64 *q) {
  ROWLOOP {
    const i64 p = perm[j];
    const i64 folded = p % N;
    const i64 sign = ((p / N) & 1) ? -1 : 1;
    i64 x = wmul(a[i * N + j], sign);
    x = wadd(x, q[i]);
    out[i * N + folded] = cs1(x, q[i]);
  }
}

/* he_fused_cuda.cu:433-469: p is [K][N] (special limbs, canonical in); pir_sp[k*K+row] =
 * (P_k^-1 mod P_row) * R mod P_row.  In place, rows K-2 .. 0, each using the already
 * updated rows above it. */
void orc_chain_backward(i64 *p, long K, long N, const i64 *pir_sp, const i64 *qsp, const i64 *ksp) {
#pragma omp parallel for schedule(static)
  for (long j = 0; j < N; ++j) {
    for (long row = K - 2; row >= 0; --row) {
      i64 x = p[row * N + j];
      for (long kk = K - 1; kk > row; --kk) {
        const i64 s = msub(x, p[kk * N + j], 2 * qsp[row]);
        x = MM(s, pir_sp[kk * K + row], qsp[row], ksp[row]);
      }
      p[row * N + j] = x;
    }
  }
}

/* he_fused_cuda.cu:471-519: x = MM(c, Rs); for k = K-1..0: x = MM(CS2(x - MM(p_k, Rs)), PiR[k][i]);
 * out = CS1(MR(x)).  pir is [K][C] with pir[k*C+i] = (P_k^-1 mod q_i) * R mod q_i. */
void orc_divide_by_p(i64 *out, const i64 *c, const i64 *p, long K, const i64 *pir, long C, long N,
                     const i64 *Rs, const i64 *q, const i64 *k) {
  ROWLOOP {
    i64 x = MM(c[i * N + j], Rs[i], q[i], k[i]);
    for (long kk = K - 1; kk >= 0; --kk) {
      const i64 pe = MM(p[kk * N + j], Rs[i], q[i], k[i]);
      x = msub(x, pe, 2 * q[i]);
      x = MM(x, pir[kk * C + i], q[i], k[i]);
    }
    x = MR(x, q[i], k[i]);
    out[i * N + j] = cs1(x, q[i]);
  }
}

/* ---------------------------------------------------------------------------------
 * Table helpers (not reference code paths; plain modular arithmetic used to build the
 * psi power tables quickly: tiberate/context/ntt_context.py:32-36).
 * --------------------------------------------------------------------------------- */
void orc_pow_series(i64 *out, long n, i64 base, i64 q) {
  u64 x = 1;
  for (long i = 0; i < n; ++i) {
    out[i] = (i64)x;
    x = (u64)(((u128)x * (u128)(u64)base) % (u128)(u64)q);
  }
}
void orc_mulmod_vec(i64 *out, const i64 *a, const i64 *b, long n, i64 q) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    i128 x = ((i128)a[i] * (i128)b[i]) % (i128)q;
    if (x < 0) x += q;
    out[i] = (i64)x;
  }
}
