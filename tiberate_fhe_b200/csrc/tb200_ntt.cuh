// tb200_ntt.cuh -- register/shared-memory tile NTT used by the two-pass negacyclic transforms.
//
// The butterfly DAG is exactly the reference's radix-2 one (csrc/ops/cuda/ntt_radix2_cuda.cu:10-47,
// intt_radix2_cuda.cu:10-48; index/twiddle formulas in SURVEY.md appendix A.1-A.3), so every lazy
// [0,2q) representative is bit-identical; only the schedule differs: instead of logN launches that
// each stream the tensor through HBM, a length-2^logN transform is split in two passes
//     pass A: the LA stages whose butterfly distance is >= 2^LB  (column transforms, strided rows)
//     pass B: the LB stages whose distance is < 2^LB             (contiguous 2^LB blocks)
// and inside a pass a CTA keeps a 2^LT-point sub-transform on chip: each thread owns 16 residues
// in registers and runs up to 4 stages ("a round") before exchanging through shared memory.
//
// Local index x in [0, 2^LT); thread t in [0, 2^(LT-4)); register i in [0,16).  In a round whose
// 4-bit register field starts at local bit f:   x = ((t >> f) << (f+4)) | (i << f) | (t & (2^f-1)).
// A stage with distance 2^d (f <= d < f+4) pairs registers i and i ^ (1 << (d-f)); its twiddle is
// table[m + (tile << (LT-1-d)) + (x >> (d+1))] with m the number of butterfly groups of that stage.
#pragma once
#include "tb200_mont.cuh"

namespace tb {

__device__ __forceinline__ int tile_x(int t, int i, int f) {
  return ((t >> f) << (f + 4)) | (i << f) | (t & ((1 << f) - 1));
}

// shared-memory element slot with one pad word per 16 (conflict-free for every round pattern)
__device__ __forceinline__ int pad16(int e) { return e + (e >> 4); }

struct PrimeRegs {
  i64 q, q2;
  u64 q4, k;
};

// Cooley-Tukey butterfly (ntt_radix2_cuda.cu:36-46): V = MM(S, O); lo = CS2(U+V); hi = CS2(U+2q-V)
__device__ __forceinline__ void bfly_ct(i64& U, i64& O, u64 S4, const PrimeRegs& p) {
  const i64 V = tb_mm_s4(O, S4, p.q4, p.k);
  const i64 a = U + V;
  const i64 b = U + p.q2 - V;
  U = tb_cs2(a, p.q2);
  O = tb_cs2(b, p.q2);
}

// Gentleman-Sande butterfly (intt_radix2_cuda.cu:36-47): lo = CS2(U+V); hi = MM(S, CS2(U+2q-V))
__device__ __forceinline__ void bfly_gs(i64& U, i64& V, u64 S4, const PrimeRegs& p) {
  const i64 a = U + V;
  const i64 o = tb_cs2(U + p.q2 - V, p.q2);
  U = tb_cs2(a, p.q2);
  V = tb_mm_s4(o, S4, p.q4, p.k);
}

// Butterfly policy of the exposed, bit-exact transforms: the reference's lazy Montgomery butterflies.
struct ExactPol {
  PrimeRegs p;
  typedef u64 TW;   // twiddle as the butterflies take it
  typedef u64 TWS;  // table element
  static __device__ __forceinline__ TW load(const TWS* t) { return __ldg(t); }
  __device__ __forceinline__ void ct(i64& U, i64& O, TW S, int /*mlog*/) const { bfly_ct(U, O, S, p); }
  __device__ __forceinline__ void gs(i64& U, i64& V, TW S, int /*mlog*/) const { bfly_gs(U, V, S, p); }
};

// field start of forward round r (r = 0 handles the largest distances)
template <int LT>
__host__ __device__ constexpr int fwd_field(int r) {
  return (LT - 4 * (r + 1)) > 0 ? (LT - 4 * (r + 1)) : 0;
}
template <int LT>
__host__ __device__ constexpr int num_rounds() {
  return (LT + 3) / 4;
}

// One forward round: stages with local distance bits top .. top-ns+1 (descending).
// mlog(d) = log2(number of groups) of the stage with local distance bit d.
// PERM (only meaningful for the field-0 round, where every thread owns (8 >> b) CONSECUTIVE twiddles of
// stage b): the table stores those stages "transposed" inside each tile -- entry (t, g) at g * T + t
// with T = 2^(LT-4) threads per tile -- so that for a fixed g the lanes of a warp read consecutive
// 16-byte records (4 L1 wavefronts per warp instruction instead of 32).
template <int LT, int R, class POL, bool PERM = false>
__device__ __forceinline__ void fwd_round(i64 (&x)[16], int t, int tile, int mlog_of_d0,
                                          const typename POL::TWS* __restrict__ tw, const POL& p) {
  constexpr int f = fwd_field<LT>(R);
  constexpr int top = LT - 1 - 4 * R;
  constexpr int ns = (LT - 4 * R) > 4 ? 4 : (LT - 4 * R);
  constexpr int T = LT > 4 ? (1 << (LT - 4)) : 1;
#pragma unroll
  for (int s = 0; s < ns; ++s) {
    const int d = top - s;   // local distance bit
    const int b = d - f;     // register bit
    const int mlog = mlog_of_d0 - d;  // groups = 2^mlog
    const bool perm = PERM && f == 0;
    const int base = (1 << mlog) + (tile << (LT - 1 - d)) + (perm ? t : ((t >> f) << (3 - b)));
    const int gs = perm ? T : 1;
#pragma unroll
    for (int g = 0; g < (8 >> b); ++g) {       // distinct twiddles: register bits above b
      const typename POL::TW S4 = p.load(tw + base + g * gs);
#pragma unroll
      for (int l = 0; l < (1 << b); ++l) {     // register bits below b
        const int i = (g << (b + 1)) | l;
        p.ct(x[i], x[i | (1 << b)], S4, mlog);
      }
    }
  }
}

// One inverse round: the same field as forward round R, stages ascending in distance.
template <int LT, int R, class POL, bool PERM = false>
__device__ __forceinline__ void inv_round(i64 (&x)[16], int t, int tile, int mlog_of_d0,
                                          const typename POL::TWS* __restrict__ tw, const POL& p) {
  constexpr int f = fwd_field<LT>(R);
  constexpr int top = LT - 1 - 4 * R;
  constexpr int ns = (LT - 4 * R) > 4 ? 4 : (LT - 4 * R);
  constexpr int T = LT > 4 ? (1 << (LT - 4)) : 1;
#pragma unroll
  for (int s = ns - 1; s >= 0; --s) {
    const int d = top - s;
    const int b = d - f;
    const int mlog = mlog_of_d0 - d;
    const bool perm = PERM && f == 0;
    const int base = (1 << mlog) + (tile << (LT - 1 - d)) + (perm ? t : ((t >> f) << (3 - b)));
    const int gs = perm ? T : 1;
#pragma unroll
    for (int g = 0; g < (8 >> b); ++g) {
      const typename POL::TW S4 = p.load(tw + base + g * gs);
#pragma unroll
      for (int l = 0; l < (1 << b); ++l) {
        const int i = (g << (b + 1)) | l;
        p.gs(x[i], x[i | (1 << b)], S4, mlog);
      }
    }
  }
}

// Exchange registers through shared memory: written with field fw, read back with field fr.
// SlotFn maps a local index x to the CTA's shared-memory slot (already padded).
template <class SlotFn>
__device__ __forceinline__ void exchange(i64 (&x)[16], i64* sm, int t, int fw, int fr, SlotFn slot) {
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) sm[slot(tile_x(t, i, fw))] = x[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = sm[slot(tile_x(t, i, fr))];
}

// Full forward tile: rounds 0..NR-1.  On entry x is laid out with field fwd_field<LT>(0);
// on exit with field fwd_field<LT>(NR-1) (== 0 unless LT == 4k where it is also 0).
template <int LT, bool PERM = false, class POL, class SlotFn>
__device__ __forceinline__ void tile_fwd(i64 (&x)[16], i64* sm, int t, int tile, int mlog_of_d0,
                                         const typename POL::TWS* __restrict__ tw, const POL& p, SlotFn slot) {
  fwd_round<LT, 0, POL, PERM>(x, t, tile, mlog_of_d0, tw, p);
  if constexpr (num_rounds<LT>() > 1) {
    exchange(x, sm, t, fwd_field<LT>(0), fwd_field<LT>(1), slot);
    fwd_round<LT, 1, POL, PERM>(x, t, tile, mlog_of_d0, tw, p);
  }
  if constexpr (num_rounds<LT>() > 2) {
    exchange(x, sm, t, fwd_field<LT>(1), fwd_field<LT>(2), slot);
    fwd_round<LT, 2, POL, PERM>(x, t, tile, mlog_of_d0, tw, p);
  }
}

// Full inverse tile: rounds NR-1..0.  Entry layout: field fwd_field<LT>(NR-1); exit: fwd_field<LT>(0).
template <int LT, bool PERM = false, class POL, class SlotFn>
__device__ __forceinline__ void tile_inv(i64 (&x)[16], i64* sm, int t, int tile, int mlog_of_d0,
                                         const typename POL::TWS* __restrict__ tw, const POL& p, SlotFn slot) {
  if constexpr (num_rounds<LT>() > 2) {
    inv_round<LT, 2, POL, PERM>(x, t, tile, mlog_of_d0, tw, p);
    exchange(x, sm, t, fwd_field<LT>(2), fwd_field<LT>(1), slot);
  }
  if constexpr (num_rounds<LT>() > 1) {
    inv_round<LT, 1, POL, PERM>(x, t, tile, mlog_of_d0, tw, p);
    exchange(x, sm, t, fwd_field<LT>(1), fwd_field<LT>(0), slot);
  }
  inv_round<LT, 0, POL, PERM>(x, t, tile, mlog_of_d0, tw, p);
}

}  // namespace tb
