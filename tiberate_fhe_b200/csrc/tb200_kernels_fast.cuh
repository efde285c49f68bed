// tb200_kernels_fast.cuh -- kernels of the internal mod-q path (see tb200_fast.cuh for why it is
// allowed and what it saves).  Same tiling as the exact transforms in tb200_kernels.cuh.
#pragma once
#include "tb200_fast.cuh"
#include "tb200_kernels.cuh"

struct TbDevFast {
  const TbFastPrime* fp;  // [P]
  const TbTw2* tw;        // [P][N] forward twiddles (plain + Shoup)
  const TbTw2* itw;       // [P][N] inverse twiddles
  const double *twd, *itwd;  // the same twiddles as centred doubles (FP64 butterfly policy)
  int logN, LA, LB, P;
};

// i64 tile <-> bit patterns of exact-integer doubles (FP64 butterfly policy)
__device__ __forceinline__ void tile_to_f64(i64 (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = __double_as_longlong(tb::FastF64Pol::from_int(x[i]));
}
// reduce to [0, q) as integers (forward pass B output) or to [-q/2-1, q/2+1] kept as doubles (the
// intermediate between the two inverse passes)
template <bool CANONICAL_INT>
__device__ __forceinline__ void tile_f64_reduce(i64 (&x)[16], const tb::FastF64Pol& p) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double r = p.reduce(__longlong_as_double(x[i]));
    if constexpr (CANONICAL_INT) {
      r = r < 0.0 ? __dadd_rn(r, p.q) : r;
      x[i] = tb::FastF64Pol::to_int(r);
    } else {
      x[i] = __double_as_longlong(r);
    }
  }
}

#define TB_FPRO_ENTER 0          // x = (a + q) * R                         (enter_ntt_radix2, mod q)
#define TB_FPRO_RESCALE_ENTER 1  // x = ((a - r) q_l^-1 + [r > q_l/2]) * R  (rescale + enter, mod q)
#define TB_FPRO_EXTEND 2         // x = sum_k d_k * (L_{k-1} R)             (ModUp extend, mod q)

struct TbFwdAArgs {
  TbView src, dst;
  const u64* resc;        // RESCALE_ENTER: per limb row (c1, c1', off) triples for this level
  i64 round_at;           // RESCALE_ENTER: q_l / 2
  const TbKsLevel* lv;    // EXTEND
  const u64* lenter2;     // EXTEND: (C, C') pairs, indexed (G.lenter_off + (k-1) P + prime)
  const double* lenterd;  // EXTEND, FP64 limbs: L_{k-1} mod q centred (no Montgomery factor), same indexing
  int prime0, LW, ngroups;
  int skip_own;           // EXTEND: the (group, own limb) pairs were pre-filled by k_fast_own_fill
  int row_shift;          // ENTER / RESCALE_ENTER: source row of limb 0 of this launch (row split of one tensor)
  int sel;                // EXTEND: 0 = every group, 1 = the groups this rank owns, 2 = the other ranks' groups
  int nsel;               // EXTEND: number of selected groups (grid.z = batch * nsel)
};

template <int PRO>
__device__ __forceinline__ i64 fast_prologue(const TbFwdAArgs& a, const TbFastPrime& P, int bt, int gi, int limb,
                                             unsigned col_off, int g, int nP) {
  if constexpr (PRO == TB_FPRO_ENTER) {
    const i64 v = a.src.row(bt, limb + a.row_shift)[col_off];
    return (i64)(P.small ? tb::shoup_lazy((u64)(v + (i64)P.q), P.Rm, P.Rm_s, P.q)
                         : tb::shoup((u64)(v + (i64)P.q), P.Rm, P.Rm_s, P.q));
  } else if constexpr (PRO == TB_FPRO_RESCALE_ENTER) {
    const i64 v = a.src.row(bt, limb + a.row_shift + 1)[col_off];
    const i64 r = a.src.row(bt, 0)[col_off];
    const u64* c = a.resc + 3 * limb;
    const u64 t = (u64)(v - r + (i64)c[2]);
    u64 x = P.small ? tb::shoup_lazy(t, c[0], c[1], P.q) : tb::shoup(t, c[0], c[1], P.q);
    x += (r > a.round_at) ? P.Rm : 0ull;  // small primes: < 5q, others: < 3q
    return (i64)x;
  } else {
    return 0;  // EXTEND is handled by extend_prologue<ALPHA> (all 16 residues of a thread at once)
  }
}

// bound (in units of 4q) of the lazy 60-bit extension sum before term k is added: 1 after the first term, +1 per
// term, back to 1 by a reduction whenever it has reached 4
__host__ __device__ constexpr int ext_terms(int k) { return k <= 4 ? k : (k - 5) % 3 + 2; }

// ModUp extend for the 16 residues of a thread, ALPHA digits each: loads first, then the Shoup sums.
template <int ALPHA>
__device__ __forceinline__ void extend_prologue(i64 (&x)[16], const TbFwdAArgs& a, const TbFastPrime& P,
                                                const TbKsGroup& G, int bt, int g, int nP, int tr, int f0, int LB,
                                                unsigned c0) {
  u64 C[ALPHA], Cs[ALPHA];
  C[0] = P.Rm;
  Cs[0] = P.Rm_s;
  const u64* le = a.lenter2 + 2 * (G.lenter_off + g);
#pragma unroll
  for (int k = 1; k < ALPHA; ++k) {
    C[k] = le[0];
    Cs[k] = le[1];
    le += 2 * nP;
  }
  const i64* base = a.src.row(bt, G.state_row0) + c0;
  const unsigned rs = (unsigned)a.src.rs;  // state rows are dense: row stride N < 2^31 / alpha
#pragma unroll
  for (int h = 0; h < 16; h += 4) {
    i64 d[4][ALPHA];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < ALPHA; ++k) d[i][k] = base[(unsigned)k * rs + ((unsigned)tb::tile_x(tr, h + i, f0) << LB)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      u64 v;
      if (P.small) {  // lazy quotients: every term < 4q, sum < 4 ALPHA q
        v = tb::shoup_lazy((u64)(d[i][0] + (i64)P.off), C[0], Cs[0], P.q);
#pragma unroll
        for (int k = 1; k < ALPHA; ++k) v += tb::shoup_lazy((u64)(d[i][k] + (i64)P.off), C[k], Cs[k], P.q);
      } else {
        // 60-bit limbs (q < 2^60): lazy quotients as well -- every term below 4q, so a sum of four stays below
        // 16q < 2^64; a longer sum is brought below 4q before the term that would overflow (ext_terms: the
        // bound in units of 4q).  Pass A takes any U below 8q.
        const u64 q4 = P.q2 + P.q2, q8 = q4 + q4;
        v = tb::shoup_lazy((u64)(d[i][0] + (i64)P.off), C[0], Cs[0], P.q);
#pragma unroll
        for (int k = 1; k < ALPHA; ++k) {
          if (ext_terms(k) == 4) {
            v = (v >= q8) ? v - q8 : v;
            v = (v >= q4) ? v - q4 : v;
          }
          v += tb::shoup_lazy((u64)(d[i][k] + (i64)P.off), C[k], Cs[k], P.q);
        }
        if (ext_terms(ALPHA) > 2) v = (v >= q8) ? v - q8 : v;
      }
      x[h + i] = (i64)v;
    }
  }
}

// The same sums on the FP64 pipe, for a target limb that takes the FP64 butterflies -- WITHOUT the
// Montgomery factor: x = d_0 + sum_{k>=1} d_k L_{k-1} mod q, so the first digit costs no product.
// (The key residues are in Montgomery form, so the FP64 key inner product x (key R) carries the factor
// again and the usual exit N^-1 R^-1 applies.)  A digit of a 40-bit source prime is a
// signed integer below 2^42 (a Montgomery product of k_digits) and enters FastF64Pol::mulmod directly;
// a digit of a 60-bit source prime is split as hi * 2^30 + lo.  Every term is below max(2^42, 1.1 q)
// in magnitude, so |x| < 2^47; x stays a double for the stages.
// Digit-major: one constant live at a time, 8 loads in flight per thread.
__device__ __forceinline__ void extend_prologue_f64(i64 (&x)[16], const TbFwdAArgs& a, const TbFastPrime& P,
                                                    const TbKsGroup& G, int bt, int g, int nP, int tr, int f0, int LB,
                                                    unsigned c0) {
  const tb::FastF64Pol pol{P.qd, P.qinv};
  double v[16];
  const double* le = a.lenterd + (G.lenter_off + g);
  const i64* row = a.src.row(bt, G.state_row0) + c0;
  // software pipeline over half-tiles: while the 8 products of one half run on the FP64 pipe the 8 digit
  // loads of the next half (same digit, or the next digit's first half) are in flight
  auto load8 = [&](i64 (&d)[8], const i64* r, int h) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = r[(unsigned)tb::tile_x(tr, h + i, f0) << LB];
  };
  auto acc8 = [&](const i64 (&d)[8], int h, int k, bool wide, double C, double C30) {
    if (wide) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double hi = tb::FastF64Pol::from_int(d[i] >> 30), lo = tb::FastF64Pol::from_int(d[i] & 0x3fffffffll);
        const double t = __dadd_rn(pol.mulmod(hi, C30), k == 0 ? lo : pol.mulmod(lo, C));
        v[h + i] = k == 0 ? t : __dadd_rn(v[h + i], t);
      }
    } else if (k == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[h + i] = tb::FastF64Pol::from_int(d[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[h + i] = __dadd_rn(v[h + i], pol.mulmod(tb::FastF64Pol::from_int(d[i]), C));
    }
  };
  i64 d0[8], d1[8];
  load8(d0, row, 0);
  for (int k = 0; k < G.alpha; ++k) {
    const bool wide = (G.wide_mask >> k) & 1;
    double C = 1.0;
    if (k > 0) {
      C = le[0];
      le += nP;
    }
    const double C30 = wide ? pol.mulmod(C, 1073741824.0) : 0.0;
    load8(d1, row, 8);
    acc8(d0, 0, k, wide, C, C30);
    row += a.src.rs;
    if (k + 1 < G.alpha) load8(d0, row, 0);
    acc8(d1, 8, k, wide, C, C30);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = __double_as_longlong(v[i]);
}

// forward pass A with a fused prologue.  EXTEND: grid.z = batch * ngroups, dst batch index = grid.z.
#ifndef TB_EXT_MINB
#define TB_EXT_MINB 3
#endif
#ifndef TB_F64_MINB
#define TB_F64_MINB 4
#endif
#ifndef TB_F64A_MINB
#define TB_F64A_MINB 3
#endif
// BIG: logN >= 12, where LB = 8 and the column width is 2^(12 - LA): strides, shared-memory slots and
// load/store offsets become immediates (ncu: half of the instructions of the runtime-LB build were
// address arithmetic).
// F64ONLY: every limb row of the launch takes the FP64 route (launcher splits the rows as for pass B).
template <int LA, int PRO, bool BIG, bool F64ONLY = false>
__global__ void __launch_bounds__(256, F64ONLY ? (PRO == TB_FPRO_EXTEND ? TB_F64A_MINB : TB_F64_MINB)
                                                : (PRO == TB_FPRO_EXTEND ? TB_EXT_MINB : 3))
    k_fast_fwd_A(TbDevFast c, TbFwdAArgs a) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int LB = BIG ? 8 : c.LB, LW = BIG ? 12 - LA : a.LW;
  const int W = 1 << LW;
  const int col = threadIdx.x & (W - 1), tr = threadIdx.x >> LW;
  const int limb = blockIdx.y, g = a.prime0 + limb;
  const TbFastPrime P = c.fp[g];
  int bt = blockIdx.z, gi = 0;
  int dz = blockIdx.z;  // batch index of the destination rows
  if constexpr (PRO == TB_FPRO_EXTEND) {
    gi = blockIdx.z % a.nsel;
    bt = blockIdx.z / a.nsel;
    if (a.sel == 1) gi = a.lv->own[gi];
    if (a.sel == 2) gi = a.lv->foreign[gi];
    dz = bt * a.ngroups + gi;
  }
  const unsigned c0 = blockIdx.x * W + col;
  i64* d = a.dst.row(dz, limb) + c0;
  constexpr int f0 = tb::fwd_field<LA>(0);
  i64 x[16];
  if constexpr (PRO == TB_FPRO_EXTEND) {
    const TbKsGroup& G = a.lv->g[gi];
    if (a.skip_own && g >= G.src_prime0 && g < G.src_prime0 + G.alpha) return;  // CTA-uniform
    if (F64ONLY || P.f64) {
      extend_prologue_f64(x, a, P, G, bt, g, c.P, tr, f0, LB, c0);
    } else if constexpr (!F64ONLY) {
      switch (G.alpha) {
#define XCASE(n) \
  case n:        \
    extend_prologue<n>(x, a, P, G, bt, g, c.P, tr, f0, LB, c0); \
    break;
        XCASE(1) XCASE(2) XCASE(3) XCASE(4) XCASE(5) XCASE(6) XCASE(7) XCASE(8)
#undef XCASE
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      x[i] = fast_prologue<PRO>(a, P, bt, gi, limb, ((unsigned)tb::tile_x(tr, i, f0) << LB) + c0, g, c.P);
  }
  auto slot = [&](int lx) { return tb::pad16((lx << LW) | col); };
  const TbTw2* tw = c.tw + ((long)g << c.logN);
  if (F64ONLY || P.f64) {  // integer prologues give lazy non-negative values < 2^48: exact doubles
    if constexpr (PRO != TB_FPRO_EXTEND) tile_to_f64(x);
    tb::tile_fwd<LA>(x, sm, tr, 0, LA - 1, c.twd + ((long)g << c.logN), tb::FastF64Pol{P.qd, P.qinv}, slot);
    // stored as doubles (|x| < 2^49): pass B of the same limb takes the FP64 route as well
  } else if constexpr (!F64ONLY) {
    if (P.small)
      tb::tile_fwd<LA>(x, sm, tr, 0, LA - 1, tw, tb::FastSmallPol{P.q, P.q2, c.logN}, slot);
    else
      tb::tile_fwd<LA>(x, sm, tr, 0, LA - 1, tw, tb::FastBigPol{P.q, P.q2}, slot);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) d[(unsigned)tb::tile_x(tr, i, 0) << LB] = x[i];
}

// forward pass B; outputs: small primes < 72q, other primes reduced to [0, 2q).
// A CTA owns one (limb, 4096-residue tile) and can walk over `bper` consecutive batch entries (the
// launcher uses bper = 1: see launch_fast_B for the measurement).
// F64ONLY: every limb row of the launch takes the FP64 butterflies (the launcher splits the rows); without
// the integer policies the kernel fits 64 registers and a fourth CTA per SM.
template <int LB, bool F64ONLY>
__global__ void __launch_bounds__(256, F64ONLY ? TB_F64_MINB : 3) k_fast_fwd_B(TbDevFast c, TbView src, TbView dst,
                                                                               int prime0, int batch, int bper,
                                                                               const TbKsLevel* skip_lv, int dbl_out) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb;
  if (skip_lv) {  // key-switch extension whose (group, own limb) pairs were pre-filled (k_fast_own_fill)
    const TbKsGroup& G = skip_lv->g[blockIdx.z % skip_lv->ngroups];
    if (g >= G.src_prime0 && g < G.src_prime0 + G.alpha) return;
  }
  const TbFastPrime P = c.fp[g];
  const long e0 = (long)blockIdx.x * nt * 16;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  const TbTw2* tw = c.tw + ((long)g << c.logN);
  {
    const int z = blockIdx.z;  // one batch entry per CTA (bper == 1, see launch_fast_B)
    const i64* s = src.row(z, limb) + e0;
    i64* d = dst.row(z, limb) + e0;
    i64 x[16];
    // round-0 layout: 16 consecutive threads read 16 consecutive residues (one 128-byte line)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = s[(blk << LB) | tb::tile_x(lt, i, f0)];
    if (F64ONLY || P.f64) {
      const tb::FastF64Pol pol{P.qd, P.qinv};
      tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, c.twd + ((long)g << c.logN), pol, slot);  // doubles from pass A
      // [0, q) integers, or -- feeding the FP64 key inner product -- doubles in [-q/2 - 1, q/2 + 1]
      if (dbl_out)
        tile_f64_reduce<false>(x, pol);
      else
        tile_f64_reduce<true>(x, pol);
    } else if constexpr (!F64ONLY) {
      if (P.small) {
        tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, tw, tb::FastSmallPol{P.q, P.q2, c.logN}, slot);
      } else {
        const tb::FastBigPol bp{P.q, P.q2};
        tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, tw, bp, slot);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = bp.reduce2q(x[i]);
      }
    }
    // final layout (field 0): a thread owns 16 consecutive residues.  Storing them directly would make
    // every warp instruction touch 32 different 128-byte lines (ncu: the L1TEX data pipe was 79 % busy),
    // so transpose through shared memory and store 512 contiguous bytes per warp instruction.
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[slot(tb::tile_x(lt, i, 0))] = x[i];
    __syncthreads();
    longlong2* dv = reinterpret_cast<longlong2*>(d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = 2 * (i * nt + tid);
      longlong2 v;
      v.x = sm[tb::pad16(e)];
      v.y = sm[tb::pad16(e + 1)];
      dv[i * nt + tid] = v;
    }
  }
}

// Forward pass B of the LAST operand of a ciphertext product fused with the tensor product (ckks_engine.py:1669-1675),
// FP64 limbs only: the transformed y1 tile never goes to HBM -- it meets the three operands transformed before it
// (x0, x1, y0: canonical Montgomery-form integers) in the coalesced layout and the thread writes d0 = x0 y0,
// d1 = x0 y1 + x1 y0, d2 = x1 y1 (Montgomery products a b R^-1 as two error-free FP64 products; canonical integers,
// which the consumers -- key inner product, own-limb reads, inverse pass B' -- take like the lazy values of k_tensor).
// The tensor product alone is HBM-bound (7 limb passes); fused, the transform's FP64 work runs under that traffic and
// one write + one read of y1 disappear.
template <int LB>
__global__ void __launch_bounds__(256, 3) k_fast_fwd_B_tensor(TbDevFast c, TbView src, TbView x0, TbView x1, TbView y0,
                                                              TbView d0, TbView d1, TbView d2, int prime0) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb, z = blockIdx.z;
  const TbFastPrime P = c.fp[g];
  const long e0 = (long)blockIdx.x * nt * 16;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  const tb::FastF64Pol pol{P.qd, P.qinv};
  const i64* s = src.row(z, limb) + e0;
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[(blk << LB) | tb::tile_x(lt, i, f0)];
  tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, c.twd + ((long)g << c.logN), pol, slot);
  tile_f64_reduce<false>(x, pol);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) sm[slot(tb::tile_x(lt, i, 0))] = x[i];
  __syncthreads();
  const i64 *p0 = x0.row(z, limb) + e0, *p1 = x1.row(z, limb) + e0, *q0 = y0.row(z, limb) + e0;
  i64 *o0 = d0.row(z, limb) + e0, *o1 = d1.row(z, limb) + e0, *o2 = d2.row(z, limb) + e0;
  auto canon = [&](double r) {  // |r| < 1.2 q -> canonical integer
    r = r < 0.0 ? __dadd_rn(r, pol.q) : r;
    r = r >= pol.q ? __dadd_rn(r, -pol.q) : r;
    return tb::FastF64Pol::to_int(r);
  };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = 2 * (i * nt + tid);
    const longlong2 a0 = *reinterpret_cast<const longlong2*>(p0 + e);
    const longlong2 a1 = *reinterpret_cast<const longlong2*>(p1 + e);
    const longlong2 b0 = *reinterpret_cast<const longlong2*>(q0 + e);
    const double b1x = pol.mulmod(__longlong_as_double(sm[tb::pad16(e)]), P.Rid);       // y1 R^-1
    const double b1y = pol.mulmod(__longlong_as_double(sm[tb::pad16(e + 1)]), P.Rid);
    const double b0x = pol.mulmod(tb::FastF64Pol::from_int(b0.x), P.Rid), b0y = pol.mulmod(tb::FastF64Pol::from_int(b0.y), P.Rid);
    const double a0x = tb::FastF64Pol::from_int(a0.x), a0y = tb::FastF64Pol::from_int(a0.y);
    const double a1x = tb::FastF64Pol::from_int(a1.x), a1y = tb::FastF64Pol::from_int(a1.y);
    longlong2 o;
    o.x = canon(pol.mulmod(a0x, b0x));
    o.y = canon(pol.mulmod(a0y, b0y));
    *reinterpret_cast<longlong2*>(o0 + e) = o;
    o.x = canon(pol.reduce(__dadd_rn(pol.mulmod(a0x, b1x), pol.mulmod(a1x, b0x))));
    o.y = canon(pol.reduce(__dadd_rn(pol.mulmod(a0y, b1y), pol.mulmod(a1y, b0y))));
    *reinterpret_cast<longlong2*>(o1 + e) = o;
    o.x = canon(pol.mulmod(a1x, b1x));
    o.y = canon(pol.mulmod(a1y, b1y));
    *reinterpret_cast<longlong2*>(o2 + e) = o;
  }
}

// inverse pass B'.  Inputs: lazy residues in (-2q, 2q) (negatives are lifted by 2q first).
// IN_MODE 1 (the exposed tb200_intt): FP64 limbs accept any |x| < 2^51 and are reduced first;
// IN_MODE 2 (after k_fast_mac): FP64 limbs arrive as doubles (|x| < 16 q), reduced first.
template <int LB, bool F64ONLY, int IN_MODE>
__global__ void __launch_bounds__(256, F64ONLY ? TB_F64_MINB : 3) k_fast_inv_B(TbDevFast c, TbView src, TbView dst,
                                                                               int prime0, int batch, int bper) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb;
  const TbFastPrime P = c.fp[g];
  const long e0 = (long)blockIdx.x * nt * 16;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  const TbTw2* tw = c.itw + ((long)g << c.logN);
  {
    const int z = blockIdx.z;  // one batch entry per CTA (bper == 1, see launch_fast_B)
    const i64* s = src.row(z, limb) + e0;
    i64* d = dst.row(z, limb) + e0;
    i64 x[16];
    // coalesced 128-bit loads (512 contiguous bytes per warp instruction), transposed through shared
    // memory into the field-0 layout (16 consecutive residues per thread)
    const longlong2* sv = reinterpret_cast<const longlong2*>(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const longlong2 v = sv[i * nt + tid];
      const int e = 2 * (i * nt + tid);
      if (IN_MODE != 0 && (F64ONLY || P.f64)) {
        sm[tb::pad16(e)] = v.x;
        sm[tb::pad16(e + 1)] = v.y;
      } else {
        sm[tb::pad16(e)] = v.x < 0 ? v.x + (i64)P.q2 : v.x;
        sm[tb::pad16(e + 1)] = v.y < 0 ? v.y + (i64)P.q2 : v.y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[slot(tb::tile_x(lt, i, 0))];
    if (F64ONLY || P.f64) {  // inputs in [0, 4q): after the LB stages < 2^(LB+2) q < 2^52; renormalised before the store
      const tb::FastF64Pol pol{P.qd, P.qinv};
      if constexpr (IN_MODE != 2) tile_to_f64(x);
      if constexpr (IN_MODE != 0) tile_f64_reduce<false>(x, pol);
      tb::tile_inv<LB, true>(x, sm, lt, tile, c.logN - 1, c.itwd + ((long)g << c.logN), pol, slot);
      tile_f64_reduce<false>(x, pol);  // stays double for pass A'
    } else if constexpr (!F64ONLY) {
      if (P.small)
        tb::tile_inv<LB, true>(x, sm, lt, tile, c.logN - 1, tw, tb::FastSmallPol{P.q, P.q2, c.logN}, slot);
      else
        tb::tile_inv<LB, true>(x, sm, lt, tile, c.logN - 1, tw, tb::FastBigPol{P.q, P.q2}, slot);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) d[(blk << LB) | tb::tile_x(lt, i, f0)] = x[i];
  }
}

// ModDown fused into the exit of inverse pass A' (MD instantiation): the ordinary limbs of the two key sums
// never return to HBM in coefficient form.  The special limbs were transformed and chain-reduced before
// (k_chain_backward, exact); here y = c B_{K-1} - sum_k p_k B_k mod q (he_fused_cuda.cu:471-519 expanded, see
// k_fast_divide_by_p) + the relinearisation / switch-key tail, written to the caller's output rows.
struct TbMdArgs {
  const i64* p;       // special limbs: element (z, k, j) at p + z * pbs + k * N + j   (z = ciphertext * 2 + half)
  long pbs;
  const u64* bn;      // [(K+1)][P][2]
  TbView add0, add1, out0, out1;
  int K, tail;        // tail as in ks_finish
};
__device__ __forceinline__ i64 tb_moddown_exit(i64 cv, const TbFastPrime& P, const u64* b, int nP, const TbMdArgs& m,
                                               const i64* pz, unsigned pos, int N) {
  const u64* bk = b + 2 * (long)m.K * nP;
  u64 x = tb::shoup((u64)(cv + (i64)P.q), bk[0], bk[1], P.q);
  for (int k = 0; k < m.K; ++k) {
    const u64* bb = b + 2 * (long)k * nP;
    x += tb::shoup((u64)(pz[(long)k * N + pos] + (i64)P.off), bb[0], bb[1], P.q);
    x = (x >= P.q2) ? x - P.q2 : x;
  }
  return (i64)(x >= P.q ? x - P.q : x);
}

// inverse pass A' + exit: y = CS1(x * N^-1 R^-1)  == intt_radix2_exit_reduce of the reference (canonical).
template <int LA, bool BIG, bool MD = false>
__global__ void __launch_bounds__(256, 3) k_fast_inv_A(TbDevFast c, TbView src, TbView dst, int prime0, int LWr,
                                                       int mac_chain, TbMdArgs md) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int LB = BIG ? 8 : c.LB, LW = BIG ? 12 - LA : LWr;
  const int W = 1 << LW;
  const int col = threadIdx.x & (W - 1), tr = threadIdx.x >> LW;
  const int limb = blockIdx.y, g = prime0 + limb;
  const TbFastPrime P = c.fp[g];
  const unsigned c0 = blockIdx.x * W + col;
  const i64* s = src.row(blockIdx.z, limb) + c0;
  i64* d = dst.row(blockIdx.z, limb) + c0;
  // MD: z = ciphertext * 2 + half; the result goes to that half's output rows, with the tail applied
  const i64 *pz = nullptr, *av = nullptr;
  const u64* bn = nullptr;
  int tail = 0;
  if constexpr (MD) {
    const int bt = blockIdx.z >> 1, h = blockIdx.z & 1;
    d = (h ? md.out1 : md.out0).row(bt, limb) + c0;
    av = (h ? md.add1 : md.add0).row(bt, limb) + c0;
    tail = (md.tail == 2 && h == 1) ? 0 : md.tail;
    pz = md.p + (long)blockIdx.z * md.pbs + c0;
    bn = md.bn + 2 * g;
  }
  const int Nn = 1 << c.logN;
  auto finish = [&](i64 y, unsigned pos) {  // canonical residue -> stored value
    if constexpr (MD) {
      y = tb_moddown_exit(y, P, bn, c.P, md, pz, pos, Nn);
      if (tail == 1) y = tb_cs1(av[pos] + y, (i64)P.q);
      if (tail == 2) y = tb_cs1(tb_add(av[pos], y, (i64)P.q2), (i64)P.q);
    }
    d[pos] = y;
  };
  constexpr int f0 = tb::fwd_field<LA>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[(unsigned)tb::tile_x(tr, i, 0) << LB];
  auto slot = [&](int lx) { return tb::pad16((lx << LW) | col); };
  const TbTw2* tw = c.itw + ((long)g << c.logN);
  if (P.f64) {  // inputs |x| <= q/2 + 1 (renormalised by inverse pass B'): < 2^(LA-1) q after the LA stages
    const tb::FastF64Pol pol{P.qd, P.qinv};
    tb::tile_inv<LA>(x, sm, tr, 0, LA - 1, c.itwd + ((long)g << c.logN), pol, slot);
    const double exd = P.exd;
    (void)mac_chain;  // (kept in the signature: the input of this pass is the same either way)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double r = pol.mulmod(__longlong_as_double(x[i]), exd);  // x N^-1 R^-1, |r| < 1.1 q
      r = r < 0.0 ? __dadd_rn(r, pol.q) : r;
      r = r >= pol.q ? __dadd_rn(r, -pol.q) : r;
      finish(tb::FastF64Pol::to_int(r), (unsigned)tb::tile_x(tr, i, f0) << LB);
    }
    return;
  }
  if (P.small)
    tb::tile_inv<LA>(x, sm, tr, 0, LA - 1, tw, tb::FastSmallPol{P.q, P.q2, c.logN}, slot);
  else
    tb::tile_inv<LA>(x, sm, tr, 0, LA - 1, tw, tb::FastBigPol{P.q, P.q2}, slot);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const u64 y = tb::shoup((u64)x[i], P.ex, P.ex_s, P.q);
    finish((i64)(y >= P.q ? y - P.q : y), (unsigned)tb::tile_x(tr, i, f0) << LB);
  }
}

// Relinearisation shortcut.  The mixed-radix digits of a digit group reconstruct the key-switch input modulo
// each of the group's OWN primes, so the extension of group g at its own limb t is the input itself, and
// its forward transform is the NTT-domain d2 limb the tensor product already produced: those L of the
// beta (L+K) (group, limb) pairs skip ModUp and both forward passes.  d2 carries the Montgomery factor;
// FP64 limbs keep their extensions without it and canonical (extend_prologue_f64), other limbs with it.
__global__ void __launch_bounds__(256) k_fast_own_fill(TbDev c, TbDevFast f, const TbKsLevel* lv, TbView d2, i64* ext,
                                                       int level, int N, int rowsE, int row0) {
  const int r = row0 + blockIdx.y, bt = blockIdx.z, g = level + r;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const int ng = lv->ngroups;
  int gi = 0;
  while (gi < ng && !(g >= lv->g[gi].src_prime0 && g < lv->g[gi].src_prime0 + lv->g[gi].alpha)) ++gi;
  if (gi == ng) return;
  longlong2 v = *reinterpret_cast<const longlong2*>(d2.row(bt, r) + j);
  if (f.fp[g].f64) {
    const TbPrime& P = c.pr[g];
    v.x = tb_mr(v.x, P.q4, P.k);
    v.y = tb_mr(v.y, P.q4, P.k);
    v.x = v.x < 0 ? v.x + P.q : v.x;
    v.y = v.y < 0 ? v.y + P.q : v.y;
    v.x = v.x >= P.q ? v.x - P.q : v.x;
    v.y = v.y >= P.q ? v.y - P.q : v.y;
    v.x = __double_as_longlong(tb::FastF64Pol::from_int(v.x));  // FP64 limbs: the extensions are doubles
    v.y = __double_as_longlong(tb::FastF64Pol::from_int(v.y));
  }
  *reinterpret_cast<longlong2*>(ext + (((long)bt * ng + gi) * rowsE + r) * N + j) = v;
}

// Montgomery reduction of a signed 128-bit T (|T| < 2^124): (T + ((T k) mod 2^62) q) / 2^62
__device__ __forceinline__ i64 tb_montred128(__int128 T, u64 q4, u64 k) {
  const u64 lo = (u64)T;
  const i64 hi = (i64)(T >> 64);
  const u64 s = (lo * k) & TB_MASK62;
  const u64 t = __umul64hi(s, q4);
  return (i64)(((u64)hi << 2) | (lo >> 62)) + (i64)t + ((lo & TB_MASK62) != 0 ? 1 : 0);
}

// Montgomery reduction of any signed 128-bit T (|T| < 2^127), result in [0, 2q): the bits above 2^96 are folded
// first, T = th 2^96 + tl  ->  tl + th (2^96 mod q), |.| < 2^97, so the reduced value lies in (-2^35, 2^35 + q].
__device__ __forceinline__ i64 tb_montred_wide(__int128 T, const TbPrime& P, u64 c96) {
  const i64 th = (i64)(T >> 96);
  const __int128 tl = T - ((__int128)th << 96);  // [0, 2^96)
  i64 r = tb_montred128(tl + (__int128)th * (i64)c96, P.q4, P.k) + P.q2;
  return r >= P.q2 ? r - P.q2 : r;
}

// x in (-2q, 4q) -> [0, 2q)   (key residues may be negative: mont_sub keeps negatives)
__device__ __forceinline__ i64 tb_norm2q(i64 x, i64 q2) {
  x = (x >= q2) ? x - q2 : x;
  return (x < 0) ? x + q2 : x;
}

// key inner product, mod q: small primes accumulate the 128-bit products over the digit groups and
// reduce once; other primes reduce per term.  Output lazy residues in [0, 2q + small).
// Relinearisation tail fused in (nadd0 != nullptr): the NTT-domain d0 / d1 limbs, times P = prod of the
// special primes, are added to the ordinary limbs of the two sums -- ModDown then returns ks + d exactly
// ((x + P d - [x]_P) / P = (x - [x]_P) / P + d), so d0 and d1 need no inverse transform of their own.
// cP: [2][nP] = P mod q (limbs whose sums carry no Montgomery factor: FP64 limbs) and P R mod q (others).
// Galois automorphism X -> X^g in the NTT domain: slot i of the transform holds the evaluation at
// psi^(2 brev(i) + 1), and (sigma_g a)(psi^e) = a(psi^(e g)), so slot i of sigma_g(a) is slot
// brev(((2 brev(i) + 1) g mod 2N - 1) / 2) of a -- a pure permutation, no signs.  Aligned blocks of 2^k slots
// map to aligned blocks (the low bits of brev(i) fix e modulo a power of two), in particular an aligned pair to
// an aligned pair: gathered reads stay as coalesced as straight ones.
__device__ __forceinline__ unsigned tb_ntt_slot_perm(unsigned i, unsigned g, int logN) {
  const unsigned bi = __brev(i) >> (32 - logN);
  const unsigned e = ((2u * bi + 1u) * g) & ((2u << logN) - 1u);
  return __brev((e - 1u) >> 1) >> (32 - logN);
}

// galois != 0 (hoisted rotations): the extension limbs are read through the NTT-domain automorphism above.
__global__ void __launch_bounds__(256, 4) k_fast_mac(TbDev c, TbDevFast f, const TbKsLevel* lv, TbKskDev key,
                                                  const i64* ext, i64* acc, int level, int N, int rowsE, int nb,
                                                  const i64* nadd0, const i64* nadd1, const i64* cP, int row0,
                                                  unsigned galois) {
  // grid.x = nb * tiles with the batch index fastest: the nb ciphertexts of a chunk read the same key
  // tile back to back, so the key is fetched from HBM once per chunk (L2 serves the rest)
  const int t = row0 + blockIdx.y, bt = blockIdx.x % nb;
  const TbPrime& P = c.pr[level + t];
  const int small = f.fp[level + t].small;
  const int j = ((blockIdx.x / nb) * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const int ng = lv->ngroups;
  // source pair of the extension limbs: (j, j + 1) itself, or its image under the automorphism
  int sj = j;
  bool swp = false;
  if (galois != 0) {
    const unsigned pj = tb_ntt_slot_perm((unsigned)j, galois, f.logN);
    sj = (int)(pj & ~1u);
    swp = (pj & 1u) != 0;
  }
  auto ldext = [&](const i64* p) {
    longlong2 v = *reinterpret_cast<const longlong2*>(p);
    if (swp) {
      const i64 tmp = v.x;
      v.x = v.y;
      v.y = tmp;
    }
    return v;
  };
  longlong2 o0, o1;
  if (f.fp[level + t].f64) {
    // FP64 limbs: extensions are doubles (k_fast_fwd_B dbl_out / k_fast_own_fill), the key residues
    // (|k| < 2^51, Montgomery form) are converted on the fly; every term is an error-free modular product,
    // |sum| < 12 q.  16 instructions per (residue, group) for both key halves against ~45 integer ones.
    const TbFastPrime F = f.fp[level + t];
    const tb::FastF64Pol pol{F.qd, F.qinv};
    double s0x = 0.0, s0y = 0.0, s1x = 0.0, s1y = 0.0;
    if (nadd0 != nullptr && t < lv->L) {  // + P d0, P d1 (NTT domain, Montgomery form like the key products)
      const long at = ((long)bt * lv->L + t) * N + j;
      const longlong2 u0 = *reinterpret_cast<const longlong2*>(nadd0 + at);
      const longlong2 u1 = *reinterpret_cast<const longlong2*>(nadd1 + at);
      s0x = pol.mulmod(tb::FastF64Pol::from_int(u0.x), F.cPd);
      s0y = pol.mulmod(tb::FastF64Pol::from_int(u0.y), F.cPd);
      s1x = pol.mulmod(tb::FastF64Pol::from_int(u1.x), F.cPd);
      s1y = pol.mulmod(tb::FastF64Pol::from_int(u1.y), F.cPd);
    }
    const i64* ep = ext + (((long)bt * ng) * rowsE + t) * N + sj;
    const long estride = (long)rowsE * N, koff = (long)(level + t) * key.rs + j;
#pragma unroll 2
    for (int gi = 0; gi < ng; ++gi) {
      const int gid = lv->g[gi].gid;
      const longlong2 e = ldext(ep);
      const longlong2 kb = *reinterpret_cast<const longlong2*>(key.b[gid] + koff);
      const longlong2 ka = *reinterpret_cast<const longlong2*>(key.a[gid] + koff);
      ep += estride;
      const double ex = __longlong_as_double(e.x), ey = __longlong_as_double(e.y);
      s0x = __dadd_rn(s0x, pol.mulmod(ex, tb::FastF64Pol::from_int(kb.x)));
      s0y = __dadd_rn(s0y, pol.mulmod(ey, tb::FastF64Pol::from_int(kb.y)));
      s1x = __dadd_rn(s1x, pol.mulmod(ex, tb::FastF64Pol::from_int(ka.x)));
      s1y = __dadd_rn(s1y, pol.mulmod(ey, tb::FastF64Pol::from_int(ka.y)));
    }
    o0.x = __double_as_longlong(s0x);
    o0.y = __double_as_longlong(s0y);
    o1.x = __double_as_longlong(s1x);
    o1.y = __double_as_longlong(s1y);
    *reinterpret_cast<longlong2*>(acc + (((long)bt * 2 + 0) * rowsE + t) * N + j) = o0;
    *reinterpret_cast<longlong2*>(acc + (((long)bt * 2 + 1) * rowsE + t) * N + j) = o1;
    return;
  }
  // relinearisation tail: issued first so that these loads overlap the group loop
  const bool has_add = nadd0 != nullptr && t < lv->L;
  i64 add0x = 0, add0y = 0, add1x = 0, add1y = 0;
  if (has_add) {
    const i64 cp = cP[(f.fp[level + t].f64 ? 0 : f.P) + level + t];
    const long at = ((long)bt * lv->L + t) * N + j;
    const longlong2 u0 = *reinterpret_cast<const longlong2*>(nadd0 + at);
    const longlong2 u1 = *reinterpret_cast<const longlong2*>(nadd1 + at);
    add0x = tb_mm_ss(u0.x, cp, P.q4, P.k);
    add0y = tb_mm_ss(u0.y, cp, P.q4, P.k);
    add1x = tb_mm_ss(u1.x, cp, P.q4, P.k);
    add1y = tb_mm_ss(u1.y, cp, P.q4, P.k);
  }
  if (small) {
    // lazy residue (< 2^49) times key residue (|k| < 2^42): |term| < 2^91, the sum over <= 32 groups fits 128 bits.
    // Two digit groups per iteration so that six 16-byte loads are in flight per thread (the kernel
    // is latency-bound: ncu long_scoreboard 5.6 stalls per issue before the unroll)
    __int128 a0x = 0, a0y = 0, a1x = 0, a1y = 0;
    const i64* ep = ext + (((long)bt * ng) * rowsE + t) * N + sj;
    const long estride = (long)rowsE * N, koff = (long)(level + t) * key.rs + j;
    int gi = 0;
    for (; gi + 2 <= ng; gi += 2) {
      const int g0 = lv->g[gi].gid, g1 = lv->g[gi + 1].gid;
      const longlong2 e0 = ldext(ep);
      const longlong2 e1 = ldext(ep + estride);
      const longlong2 kb0 = *reinterpret_cast<const longlong2*>(key.b[g0] + koff);
      const longlong2 ka0 = *reinterpret_cast<const longlong2*>(key.a[g0] + koff);
      const longlong2 kb1 = *reinterpret_cast<const longlong2*>(key.b[g1] + koff);
      const longlong2 ka1 = *reinterpret_cast<const longlong2*>(key.a[g1] + koff);
      ep += 2 * estride;
      a0x += (__int128)e0.x * kb0.x + (__int128)e1.x * kb1.x;
      a0y += (__int128)e0.y * kb0.y + (__int128)e1.y * kb1.y;
      a1x += (__int128)e0.x * ka0.x + (__int128)e1.x * ka1.x;
      a1y += (__int128)e0.y * ka0.y + (__int128)e1.y * ka1.y;
    }
    if (gi < ng) {
      const int gid = lv->g[gi].gid;
      const longlong2 e = ldext(ep);
      const longlong2 kb = *reinterpret_cast<const longlong2*>(key.b[gid] + koff);
      const longlong2 ka = *reinterpret_cast<const longlong2*>(key.a[gid] + koff);
      a0x += (__int128)e.x * kb.x;
      a0y += (__int128)e.y * kb.y;
      a1x += (__int128)e.x * ka.x;
      a1y += (__int128)e.y * ka.y;
    }
    o0.x = tb_montred128(a0x, P.q4, P.k) + P.q;
    o0.y = tb_montred128(a0y, P.q4, P.k) + P.q;
    o1.x = tb_montred128(a1x, P.q4, P.k) + P.q;
    o1.y = tb_montred128(a1y, P.q4, P.k) + P.q;
  } else {
    // 60-bit limbs: extension residues in [0, 2q) (< 2^61) times key residues (|k| < 2^62): |term| < 2^123, the sum
    // over <= 16 groups fits 128 bits; one reduction per sum (tb_montred_wide) instead of one Montgomery product
    // and one normalisation per term.
    __int128 a0x = 0, a0y = 0, a1x = 0, a1y = 0;
    const i64* ep = ext + (((long)bt * ng) * rowsE + t) * N + sj;
    const long estride = (long)rowsE * N, koff = (long)(level + t) * key.rs + j;
#pragma unroll 2
    for (int gi = 0; gi < ng; ++gi) {
      const int gid = lv->g[gi].gid;
      const longlong2 e = ldext(ep);
      const longlong2 kb = *reinterpret_cast<const longlong2*>(key.b[gid] + koff);
      const longlong2 ka = *reinterpret_cast<const longlong2*>(key.a[gid] + koff);
      ep += estride;
      a0x += (__int128)e.x * kb.x;
      a0y += (__int128)e.y * kb.y;
      a1x += (__int128)e.x * ka.x;
      a1y += (__int128)e.y * ka.y;
    }
    const u64 c96 = f.fp[level + t].c96;
    o0.x = tb_montred_wide(a0x, P, c96);
    o0.y = tb_montred_wide(a0y, P, c96);
    o1.x = tb_montred_wide(a1x, P, c96);
    o1.y = tb_montred_wide(a1y, P, c96);
  }
  if (has_add) {
    o0.x += add0x;
    o0.y += add0y;
    o1.x += add1x;
    o1.y += add1y;
    if (!small) {  // the 60-bit inverse butterflies expect [0, 2q)
      o0.x = tb_norm2q(o0.x, P.q2);
      o0.y = tb_norm2q(o0.y, P.q2);
      o1.x = tb_norm2q(o1.x, P.q2);
      o1.y = tb_norm2q(o1.y, P.q2);
    }
  }
  *reinterpret_cast<longlong2*>(acc + (((long)bt * 2 + 0) * rowsE + t) * N + j) = o0;
  *reinterpret_cast<longlong2*>(acc + (((long)bt * 2 + 1) * rowsE + t) * N + j) = o1;
}

// ModDown, mod q (he_fused_cuda.cu:471-519 composes x <- (x - p_k) P_k^-1 for k = K-1..0; expanded:
// x = c B_{K-1} - sum_k p_k B_k with B_k = prod_{j<=k} P_j^-1).  The special limbs p_k are the exact
// integers produced by the chain-backward step (representative-sensitive, kept exact); the result is
// canonical, hence identical to the reference's.  bn: [(2K+1)][P][2] = (-B_k mod q, Shoup) for k < K,
// then (B_{K-1}, Shoup), then (P_k^-1 mod q, Shoup) for k < K.  TAIL as in k_divide_by_p.
template <int TAIL>
__global__ void __launch_bounds__(256) k_fast_divide_by_p(TbDevFast f, TbView cc, TbView p, TbView add, TbView out,
                                                          const u64* bn, int K, int prime0, int N) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const int g = prime0 + r;
  const TbFastPrime P = f.fp[g];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const u64* b = bn + 2 * g;
  const longlong2 cv = *reinterpret_cast<const longlong2*>(cc.row(bt, r) + j);
  const u64* bk = b + 2 * (long)K * f.P;
  i64 y0, y1;
  if (P.small) {
    // 40-bit limbs: the reference's own composition x <- (x - p_k) P_k^-1, k = K-1 .. 0, with lazy Shoup quotients:
    // K products instead of the K + 1 of the expanded form; every intermediate is below 4q, `2 off - p_k` is a
    // non-negative representative of -p_k (off: multiple of q in [2^61, 2^62), |p_k| < 2^62); one exact reduction
    // at the end on the FP64 pipe.  The (first four) special limbs are loaded before the chain starts.
    longlong2 pv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < K) pv[k] = *reinterpret_cast<const longlong2*>(p.row(bt, k) + j);
    const u64* pi = b + 2 * (long)(K + 1) * f.P;  // rows K+1 .. 2K: (P_k^-1 mod q, Shoup)
    u64 x0 = (u64)(cv.x + (i64)P.q), x1 = (u64)(cv.y + (i64)P.q);
    const u64 off2 = P.off << 1;
    for (int k = K - 1; k >= 4; --k) {  // more than four special primes: the upper ones straight from memory
      const longlong2 pk = *reinterpret_cast<const longlong2*>(p.row(bt, k) + j);
      const u64* bb = pi + 2 * (long)k * f.P;
      x0 = tb::shoup_lazy(x0 + (off2 - (u64)pk.x), bb[0], bb[1], P.q);
      x1 = tb::shoup_lazy(x1 + (off2 - (u64)pk.y), bb[0], bb[1], P.q);
    }
#pragma unroll
    for (int k = 3; k >= 0; --k)
      if (k < K) {
        const u64* bb = pi + 2 * (long)k * f.P;
        x0 = tb::shoup_lazy(x0 + (off2 - (u64)pv[k].x), bb[0], bb[1], P.q);
        x1 = tb::shoup_lazy(x1 + (off2 - (u64)pv[k].y), bb[0], bb[1], P.q);
      }
    const tb::FastF64Pol pol{P.qd, P.qinv};
    double r0 = pol.reduce(tb::FastF64Pol::from_int((i64)x0)), r1 = pol.reduce(tb::FastF64Pol::from_int((i64)x1));
    r0 = r0 < 0.0 ? __dadd_rn(r0, pol.q) : r0;
    r1 = r1 < 0.0 ? __dadd_rn(r1, pol.q) : r1;
    r0 = r0 >= pol.q ? __dadd_rn(r0, -pol.q) : r0;
    r1 = r1 >= pol.q ? __dadd_rn(r1, -pol.q) : r1;
    y0 = tb::FastF64Pol::to_int(r0);
    y1 = tb::FastF64Pol::to_int(r1);
  } else {
    u64 x0 = tb::shoup((u64)(cv.x + (i64)P.q), bk[0], bk[1], P.q);
    u64 x1 = tb::shoup((u64)(cv.y + (i64)P.q), bk[0], bk[1], P.q);
    for (int k = 0; k < K; ++k) {
      const longlong2 pv = *reinterpret_cast<const longlong2*>(p.row(bt, k) + j);
      const u64* bb = b + 2 * (long)k * f.P;
      x0 += tb::shoup((u64)(pv.x + (i64)P.off), bb[0], bb[1], P.q);
      x1 += tb::shoup((u64)(pv.y + (i64)P.off), bb[0], bb[1], P.q);
      x0 = (x0 >= P.q2) ? x0 - P.q2 : x0;
      x1 = (x1 >= P.q2) ? x1 - P.q2 : x1;
    }
    y0 = (i64)(x0 >= P.q ? x0 - P.q : x0);
    y1 = (i64)(x1 >= P.q ? x1 - P.q : x1);
  }
  if constexpr (TAIL != 0) {
    const longlong2 av = *reinterpret_cast<const longlong2*>(add.row(bt, r) + j);
    if constexpr (TAIL == 1) {
      y0 = tb_cs1(av.x + y0, (i64)P.q);
      y1 = tb_cs1(av.y + y1, (i64)P.q);
    } else {
      y0 = tb_cs1(tb_add(av.x, y0, (i64)P.q2), (i64)P.q);
      y1 = tb_cs1(tb_add(av.y, y1, (i64)P.q2), (i64)P.q);
    }
  }
  longlong2 o;
  o.x = y0;
  o.y = y1;
  *reinterpret_cast<longlong2*>(out.row(bt, r) + j) = o;
}
