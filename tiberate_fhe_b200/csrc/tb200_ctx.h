// tb200_ctx.h -- host side of the context: every prime-dependent constant the kernels consume.
//
// Restates (in C++ with 128-bit integers) what the reference derives in Python:
//   tiberate/context/mont_context.py:26-57        R = 2^62, R^2 mod q, k = (R R^-1 - 1)/q
//   tiberate/context/ntt_context.py:21-85,277-298 psi root (smallest x >= 2 rule), power tables in
//                                                 bit-reversed order, twiddles entered into
//                                                 Montgomery form with mont_enter_Rs (lazy), N^-1 R
//   tiberate/context/rns_partition.py:7-66        digit groups (single device; limb sharding is
//                                                 layered on top by the host, see dist.py)
//   tiberate/context/ntt_context.py:497-534       Y_scalar / L_scalar / L_enter
//   tiberate/ckks_engine.py:114-143,201-239       rescale scales, P_k^-1 R tables
#pragma once
#include <string>
#include <vector>

#include "tb200_ks_core.cuh"

typedef __int128 i128;
typedef unsigned __int128 u128;

static inline u64 h_mulmod(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
static inline u64 h_powmod(u64 a, u64 e, u64 m) {
  u64 r = 1 % m;
  a %= m;
  while (e) {
    if (e & 1) r = h_mulmod(r, a, m);
    a = h_mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}
static inline u64 h_invmod_prime(u64 a, u64 p) { return h_powmod(a % p, p - 2, p); }
// -q^-1 mod 2^62 by Newton iteration (q odd)
static inline u64 h_neg_inv_pow2(u64 q) {
  u64 x = q;  // correct to 3 bits
  for (int i = 0; i < 6; ++i) x *= 2 - q * x;
  return (0 - x) & TB_MASK62;
}
// host copy of the exact Montgomery product (same closed form as tb_mm_ss)
static inline i64 h_mm(i64 a, i64 b, i64 q, u64 k) {
  const i128 x = (i128)a * (i128)b;
  const u64 s = ((u64)x * k) & TB_MASK62;
  const i128 t = x + (i128)((u128)s * (u128)(u64)q);
  return (i64)(t >> 62);
}
static inline u64 h_shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static inline int h_bitlen(u64 x) {
  int n = 0;
  while (x) {
    ++n;
    x >>= 1;
  }
  return n;
}
static inline int h_bitrev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) {
    r = (r << 1) | (x & 1);
    x >>= 1;
  }
  return r;
}

struct tb200_ctx {
  int device = 0, logN = 0, N = 0, P = 0, K = 0, LA = 0, LB = 0, scale_bits = 40, num_ord = 0;
  int num_levels = 0;  // levels 0..num_ord-1 (level = number of dropped scale primes)
  int chunk = 4;
  int rank = 0, world = 1;       // RNS-limb sharding: this context holds the primes owned by `rank`
  std::vector<int> ord_gid;      // global id of every local ordinary prime (ascending)
  std::vector<i64> qg;           // the global prime chain
  std::vector<int> lstart;       // per global level: local index of the first ordinary prime still alive
  std::vector<i64> q;
  std::vector<u64> k;
  std::vector<TbPrime> primes;
  std::vector<u64> psi, ipsi;  // [P][N] lazy Montgomery twiddles (unshifted), host copies
  std::vector<TbKsLevel> ks;   // per level
  // device
  TbPrime* d_primes = nullptr;
  u64 *d_psi4 = nullptr, *d_ipsi4 = nullptr;
  i64* d_rescale = nullptr;  // [num_ord][P]: row l, column g = (q_l^-1 mod q_g) R mod q_g
  i64* d_pir = nullptr;      // [K][P]
  i64* d_pir_sp = nullptr;   // [K][K]  (k, row)
  i64* d_lenter = nullptr;   // all groups' L_enter blocks
  TbKsLevel* d_ks = nullptr; // [num_ord]
  // fast (mod-q) path tables
  int fast = 1;
  int fused_core = 1;        // FP64 limbs: pass B + key inner product + inverse pass B' in one kernel
  int fused_moddown = 0;     // ModDown + tail inside the exit of inverse pass A' of the ordinary limbs (measured slower)
  int fused_tensor = 1;      // cc_mult: pass B of the last operand + tensor product in one kernel on the FP64 limbs
  int sum_ntt = 1;           // exposed forward NTT: deferred-reduction kernels for lazy inputs on the FP64 limbs
  int stream_ws = 1;         // engine-call scratch leased from the stream-ordered pool on the caller's stream
  int side_rows = 0;         // the 60-bit limb rows of a key switch run on a forked stream beside the FP64 rows
  cudaStream_t side = nullptr;            // created on first use (device of the context)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int f64_eighths = 8;       // share (in eighths) of the small-prime limbs whose butterflies use the FP64 pipe
  std::vector<TbFastPrime> fps;
  TbFastPrime* d_fp = nullptr;
  TbTw2 *d_tw = nullptr, *d_itw = nullptr;
  double *d_twd = nullptr, *d_itwd = nullptr;  // centred double twiddles, same layout
  u64* d_resc3 = nullptr;    // [num_ord][P][3]: (q_l^-1 R mod q_g, Shoup companion, offset) per (level l, prime g)
  i64* d_cP = nullptr;       // [2][P]: P mod q, P R mod q (P = product of the special primes): relinearisation tail in the MAC
  u64* d_bn = nullptr;       // ModDown: [(K+1)][P][2] (-B_k mod q, Shoup) k < K, then (B_{K-1}, Shoup)
  double* d_lenterd = nullptr; // L_{k-1} mod q_g centred doubles, same indexing as d_lenter
  u64* d_lenter2 = nullptr;  // (L_{k-1} R mod q_g, Shoup companion) pairs, same indexing as d_lenter
  i64* ws = nullptr;         // engine workspace
  size_t ws_elems = 0;
  TbDevFast devf() const {
    TbDevFast d;
    d.fp = d_fp;
    d.tw = d_tw;
    d.itw = d_itw;
    d.twd = d_twd;
    d.itwd = d_itwd;
    d.logN = logN;
    d.LA = LA;
    d.LB = LB;
    d.P = P;
    return d;
  }
  TbDev dev() const {
    TbDev d;
    d.pr = d_primes;
    d.psi4 = d_psi4;
    d.ipsi4 = d_ipsi4;
    d.logN = logN;
    d.LA = LA;
    d.LB = LB;
    d.P = P;
    d.fp = d_fp;
    d.twd = d_twd;
    d.itwd = d_itwd;
    d.x64 = fast;
    return d;
  }
};
