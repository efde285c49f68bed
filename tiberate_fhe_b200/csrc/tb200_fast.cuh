// tb200_fast.cuh -- the internal ("mod-q") transform path of the fused engine calls.
//
// Inside cc_mult+relinearize, switch_key and rotate_single every NTT-domain intermediate is consumed
// by a step that canonicalises: the reference ends each chain with intt_radix2_exit_reduce
// (mont_used_in_ntt.cuh:172-206) or the ModDown (he_fused_cuda.cu:471-519), whose outputs depend only
// on the residue class of their inputs.  So between the (exact, integer) ModUp digits and those
// canonicalising steps only values modulo q matter, and the butterflies need not reproduce the
// reference's lazy Montgomery representatives.  This path uses Harvey butterflies with Shoup
// twiddles (w, w' = floor(w 2^64 / q)):  V = O*w - umulhi(O, w')*q  in [0, 2q) for ANY 64-bit O,
//   * "small" primes (q < 2^42, the 40-bit scale primes = 34 of 39 limbs at logN16): no conditional
//     subtraction at all -- forward values grow by 4q per stage (< 72q < 2^49 after 17 stages),
//     inverse values double per stage (< 2^(2+17) q < 2^61);
//   * other primes (the 60-bit base / special primes): one conditional subtraction per butterfly,
//     values in [0, 4q) (forward) / [0, 2q) (inverse).
// 10 IMAD + ~12 ALU instructions per butterfly instead of 20 + 35 for the exact one.
// The exposed ops (tb200_ntt / tb200_intt / cc_mult_triplet) keep the exact butterflies.
#pragma once
#include "tb200_ntt.cuh"

struct __align__(16) TbTw2 {
  u64 w;   // plain twiddle psi^brev(k) mod q (canonical)
  u64 ws;  // floor(w * 2^64 / q)
};

// per-prime constants of the fast path
struct __align__(16) TbFastPrime {
  u64 q, q2;
  u64 Rm, Rm_s;      // R mod q and its Shoup companion              (enter: x -> x R)
  u64 ex, ex_s;      // N^-1 R^-1 mod q                             (exit of the inverse transform)
  u64 off;           // q << (62 - bitlen(q)): multiple of q in [2^61, 2^62) making signed digits non-negative
  int small;         // q < 2^42
  int f64;           // small prime whose butterflies run on the FP64 pipe (see FastF64Pol)
  double qd, qinv;   // q and 1/q as doubles
  double exd, nid;   // ex and N^-1 mod q, centred into (-q/2, q/2]
  double Rcd, cPd;   // R mod q and P mod q (P = product of the special primes) centred
  u64 c96;           // 2^96 mod q (tb_montred_wide)
  double Rid;        // R^-1 mod q centred (k_fast_fwd_B_tensor: Montgomery products on the FP64 pipe)
};

namespace tb {

__device__ __forceinline__ u64 shoup(u64 x, u64 w, u64 ws, u64 q) {
  const u64 h = __umul64hi(x, ws);
  return x * w - h * q;  // in [0, 2q)
}

// Shoup product with an approximate quotient (the low x low partial product and the carries of the
// cross terms are dropped): the quotient is short by at most 2, so the result lies in [0, 4q).
// Three 32-bit multiplies instead of the four (plus carry chain) of an exact umulhi.
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
  const unsigned xh = (unsigned)(x >> 32), xl = (unsigned)x;
  const unsigned wh = (unsigned)(ws >> 32), wl = (unsigned)ws;
  const u64 h = (u64)xh * wh + __umulhi(xh, wl) + __umulhi(xl, wh);
  return x * w - h * q;
}

__device__ __forceinline__ TbTw2 load_tw2(const TbTw2* p) {
#ifndef TB200_HOST_EMU
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
  TbTw2 r;
  r.w = v.x;
  r.ws = v.y;
  return r;
#else
  return *p;
#endif
}

// q < 2^42: no reductions.  Forward: values < B0 + 4q * stages (lazy Shoup quotient).  Inverse: inputs < 4q, the bound
// doubles per stage; `off` = q << (2 + stage) keeps U - V non-negative and is a multiple of q.
struct FastSmallPol {
  u64 q, q2;
  int logN;
  typedef TbTw2 TW;
  typedef TbTw2 TWS;
  static __device__ __forceinline__ TW load(const TWS* t) { return load_tw2(t); }
  __device__ __forceinline__ void ct(i64& U, i64& O, TW S, int) const {
    const u64 u = (u64)U, v = shoup_lazy((u64)O, S.w, S.ws, q);  // v < 4q
    U = (i64)(u + v);
    O = (i64)(u + 2 * q2 - v);
  }
  __device__ __forceinline__ void gs(i64& U, i64& V, TW S, int mlog) const {
    const u64 u = (u64)U, v = (u64)V;
    const u64 off = q << (2 + (logN - 1 - mlog));
    U = (i64)(u + v);
    V = (i64)shoup_lazy(u + off - v, S.w, S.ws, q);  // < 4q <= bound of the next stage
  }
};

// q < 2^42 on the FP64 pipe.  B200 issues 64 DFMA/clk/SM on a pipe the integer butterflies leave idle;
// a modular product of an exact-integer double a (|a| <= 2^52) with a centred twiddle |w| <= q/2 is
//     h = a*w (rounded), l = fma(a, w, -h) (the exact rounding error),
//     c = (h/q + M) - M  with M = 1.5 * 2^52  (round-to-integer on the FP64 pipe: |h/q| <= 2^51, so
//                                              h/q + M lies in [2^52, 2^53] where doubles are integers),
//     r = fma(-c, q, h) + l            -- exact: h - c q and l are integers far below 2^53 --
// giving r = a w - c q with |r| < 1.1 q: 6 FP64 instructions, signed residues, no conditional
// subtraction, nothing on the XU pipe (rint / int<->double conversions would go there).  Measured
// (tools/ubench_modmul.cu): 1.33 T butterflies/s against 0.78 T/s for the integer Shoup butterfly.
// Registers hold the doubles' bit patterns in the i64 tile, so the tiling and the shared-memory
// exchanges are shared; the twiddles come from a table of centred doubles in the same layout.
// Bounds: forward values grow by < 1.1q per stage (< 2^49 after 17 stages); inverse values double per
// stage, so the inverse kernels renormalise once per pass (8-9 stages: < 2^52).
#define TB_F64_MAGIC 6755399441055744.0        // 1.5 * 2^52
#define TB_F64_MAGIC_BITS 0x4338000000000000ll  // its bit pattern
struct FastF64Pol {
  double q, qinv;
  typedef double TW;
  typedef double TWS;
  static __device__ __forceinline__ TW load(const TWS* t) { return __ldg(t); }
  static __device__ __forceinline__ double round_int(double x_plus_magic) { return __dadd_rn(x_plus_magic, -TB_F64_MAGIC); }
  // exact integer (|x| < 2^51) <-> double without the conversion unit
  // (the magic's low word is zero: one 32-bit add on the high word)
  static __device__ __forceinline__ double from_int(i64 x) {
    const u64 b = ((u64)((unsigned)((u64)x >> 32) + 0x43380000u) << 32) | (unsigned)x;
    return __dadd_rn(__longlong_as_double((i64)b), -TB_F64_MAGIC);
  }
  static __device__ __forceinline__ i64 to_int(double d) {
    const u64 b = (u64)__double_as_longlong(__dadd_rn(d, TB_F64_MAGIC));
    return (i64)(((u64)((unsigned)(b >> 32) - 0x43380000u) << 32) | (unsigned)b);
  }
  __device__ __forceinline__ double mulmod(double a, double w) const {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double c = round_int(__fma_rn(h, qinv, TB_F64_MAGIC));
    return __dadd_rn(__fma_rn(-c, q, h), l);
  }
  __device__ __forceinline__ double reduce(double a) const {  // -> [-q/2 - 1, q/2 + 1]
    return __fma_rn(-round_int(__fma_rn(a, qinv, TB_F64_MAGIC)), q, a);
  }
  __device__ __forceinline__ void ct(i64& U, i64& O, TW w, int) const {
    const double u = __longlong_as_double(U), v = mulmod(__longlong_as_double(O), w);
    U = __double_as_longlong(__dadd_rn(u, v));
    O = __double_as_longlong(__dadd_rn(u, -v));
  }
  __device__ __forceinline__ void gs(i64& U, i64& V, TW w, int) const {
    const double u = __longlong_as_double(U), v = __longlong_as_double(V);
    U = __double_as_longlong(__dadd_rn(u, v));
    V = __double_as_longlong(mulmod(__dadd_rn(u, -v), w));
  }
};

// The reference's lazy Montgomery butterfly (bfly_ct: V = MM(S, O); CS2(U + V), CS2(U + 2q - V)) with its
// EXACT representatives, on the FP64 pipe, for q < 2^42 and |values| < 2^50.
// MM(a, b) = floor(a b / 2^62) + floor(s q / 2^62) + [a b mod 2^62 != 0]  (tb200_mont.cuh) lies in
// [xh, xh + q] with xh = floor(a b / 2^62) and is congruent to r = a b 2^-62 mod q, so it IS the canonical
// residue r whenever xh < r < q + min(xh, 0): with |xh| <= |a| q 2^-62 + 1 (a few thousand for lazy inputs
// against q ~ 2^40) that is all but ~2^-20 of the cases; r comes from one error-free FP64 product with the
// plain twiddle w = S 2^-62 mod q.  The remaining cases take the integer Montgomery product (cold branch,
// twiddle fetched from the integer table at the same index).  Bit-identical to ExactPol by construction.
// out of line: 64 inlined copies of the integer product per tile would cost registers and I-cache for a
// branch taken about once per million butterflies
__device__ __noinline__ double exact_mm_cold(double O, const u64* s4, u64 q4, u64 k) {
  return FastF64Pol::from_int(tb_mm_s4(FastF64Pol::to_int(O), __ldg(s4), q4, k));
}

struct ExactF64Pol {
  FastF64Pol f;
  PrimeRegs p;
  double q2, xbs;       // 2q, q * 2^-62
  const double* twd;    // this prime's rows of the double / integer twiddle tables (same layout)
  const u64* psi4;
  struct TW {
    double w;
    const u64* s4;
  };
  typedef double TWS;
  __device__ __forceinline__ TW load(const TWS* t) const {
    TW r;
    r.w = __ldg(t);
    r.s4 = psi4 + (t - twd);
    return r;
  }
  // MM(O, S) as a double; w = centred S 2^-62 mod q
  __device__ __forceinline__ double mm(double O, double w, const u64* s4) const {
    double r = f.mulmod(O, w);
    r = r < 0.0 ? __dadd_rn(r, f.q) : r;
    r = r >= f.q ? __dadd_rn(r, -f.q) : r;
    const double xb = __fma_rn(O < 0.0 ? -O : O, xbs, 2.0);
    if (!(r > xb && r < __dadd_rn(f.q, -xb)))  // rare: the representative depends on floor(O S / 2^62)
      r = exact_mm_cold(O, s4, p.q4, p.k);
    return r;
  }
  __device__ __forceinline__ void ct(i64& Ub, i64& Ob, TW S, int) const {
    const double U = __longlong_as_double(Ub);
    const double V = mm(__longlong_as_double(Ob), S.w, S.s4);
    const double a = __dadd_rn(U, V), b = __dadd_rn(__dadd_rn(U, q2), -V);
    Ub = __double_as_longlong(a >= q2 ? __dadd_rn(a, -q2) : a);
    Ob = __double_as_longlong(b >= q2 ? __dadd_rn(b, -q2) : b);
  }
};

// The same exact representatives at the cost of a mod-q butterfly, for tiles whose inputs all lie in [0, 2q)
// (every value the reference's own operators hand to a forward transform).  With U in [0, 2q) and
// V = MM(S, O) in [0, q + small], CS2(U + V) and CS2(U + 2q - V) ARE reductions modulo 2q, outputs stay in
// [0, 2q), and V depends on O only through O mod q (outside the ~2^-20 band above).  So the lazy value after any
// number of stages is T mod 2q, where T is the same butterfly network evaluated WITHOUT the conditional
// subtractions: U' = U + V, O' = U - V, |T| < 2q + stages * (q + small) < 2^46 -- exact in a double.  The two
// CS2 per butterfly (compare + predicated add each) and the two-sided canonicalisation disappear; what is
// left is the error-free product, one sign fix and one band test: 11 FP64 instructions instead of 21.  The
// kernels reduce T modulo 2q once per pass.
// The band cases are not resolved inside the butterfly: the policy only records the smallest |V| it met, and a
// tile that saw one inside the band (about 3 % of the 4096-point tiles at logN16) is transformed again with
// ExactF64Pol from its untouched input.  The hot loop is then branch-free.
struct ExactSumPol {
  FastF64Pol f;
  double q2, inv2q, xbmax;  // 2q, 1/(2q), band below which MM's representative depends on floor(O S / 2^62)
  unsigned* minhi;          // smallest high word of |V| seen by this thread (non-negative doubles order like
                            // their bit patterns: two integer instructions instead of an FP64 min)
  typedef double TW;
  typedef double TWS;
  static __device__ __forceinline__ TW load(const TWS* t) { return __ldg(t); }
  __device__ __forceinline__ void ct(i64& Ub, i64& Ob, TW w, int) const {
    const double U = __longlong_as_double(Ub), O = __longlong_as_double(Ob);
    double V = f.mulmod(O, w);  // centred residue of O S 2^-62
    const unsigned hi = (unsigned)((u64)__double_as_longlong(V) >> 32);
    *minhi = min(*minhi, hi & 0x7fffffffu);
    V = (hi >> 31) ? __dadd_rn(V, f.q) : V;  // sign bit from the integer side
    Ub = __double_as_longlong(__dadd_rn(U, V));
    Ob = __double_as_longlong(__dadd_rn(U, -V));
  }
  // Gentleman-Sande (inverse) butterfly, same idea: lo = CS2(U + V) is (U + V) mod 2q, and hi = MM(S, CS2(U + 2q - V))
  // depends on U - V only modulo q outside the band, where it is the canonical residue.  The lo path doubles per
  // stage (|T| < 2q 2^8 per pass), the hi output restarts in [0, q).
  __device__ __forceinline__ void gs(i64& Ub, i64& Vb, TW w, int) const {
    const double U = __longlong_as_double(Ub), V = __longlong_as_double(Vb);
    double r = f.mulmod(__dadd_rn(U, -V), w);
    const unsigned hi = (unsigned)((u64)__double_as_longlong(r) >> 32);
    *minhi = min(*minhi, hi & 0x7fffffffu);
    r = (hi >> 31) ? __dadd_rn(r, f.q) : r;
    Ub = __double_as_longlong(__dadd_rn(U, V));
    Vb = __double_as_longlong(r);
  }
  // MM(y, c R) for a lazy y in [0, 2q) and a canonical constant: the canonical residue of y c outside the band
  __device__ __forceinline__ double scale(double y, double c_centred) const {
    double r = f.mulmod(y, c_centred);
    const unsigned hi = (unsigned)((u64)__double_as_longlong(r) >> 32);
    *minhi = min(*minhi, hi & 0x7fffffffu);
    return (hi >> 31) ? __dadd_rn(r, f.q) : r;
  }
  // T -> T mod 2q in [0, 2q), as a double / as an integer
  __device__ __forceinline__ double finish_d(i64 Tb) const {
    const double T = __longlong_as_double(Tb);
    const double y = __fma_rn(-FastF64Pol::round_int(__fma_rn(T, inv2q, TB_F64_MAGIC)), q2, T);
    return y < 0.0 ? __dadd_rn(y, q2) : y;
  }
  __device__ __forceinline__ i64 finish(i64 Tb) const {
    const double T = __longlong_as_double(Tb);
    double y = __fma_rn(-FastF64Pol::round_int(__fma_rn(T, inv2q, TB_F64_MAGIC)), q2, T);
    y = y < 0.0 ? __dadd_rn(y, q2) : y;
    return FastF64Pol::to_int(y);
  }
};

// any q < 2^60: Harvey butterflies with the lazy Shoup quotient (short by at most 2: products in [0, 4q)).
// Forward: inputs below 8q; U is brought below 4q, outputs below 8q < 2^63.  Inverse: values below 4q.
struct FastBigPol {
  u64 q, q2;
  typedef TbTw2 TW;
  typedef TbTw2 TWS;
  static __device__ __forceinline__ TW load(const TWS* t) { return load_tw2(t); }
  __device__ __forceinline__ void ct(i64& U, i64& O, TW S, int) const {
    const u64 q4 = q2 + q2;
    u64 u = (u64)U;
    u = (u >= q4) ? u - q4 : u;
    const u64 v = shoup_lazy((u64)O, S.w, S.ws, q);  // < 4q for any 64-bit O
    U = (i64)(u + v);
    O = (i64)(u + q4 - v);
  }
  __device__ __forceinline__ void gs(i64& U, i64& V, TW S, int) const {
    const u64 q4 = q2 + q2;
    const u64 u = (u64)U, v = (u64)V;
    u64 a = u + v;
    a = (a >= q4) ? a - q4 : a;
    U = (i64)a;
    V = (i64)shoup_lazy(u + q4 - v, S.w, S.ws, q);
  }
  // forward output (< 8q) -> [0, 2q)
  __device__ __forceinline__ i64 reduce2q(i64 x) const {
    const u64 q4 = q2 + q2;
    u64 y = (u64)x;
    y = (y >= q4) ? y - q4 : y;
    y = (y >= q2) ? y - q2 : y;
    return (i64)y;
  }
};

}  // namespace tb
