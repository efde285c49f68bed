// tb200_ks_core.cuh -- the key-switch core of the mod-q path as ONE kernel per (limb, tile, ciphertext):
//
//     for every digit group g:   E_g  = forward pass B of the ModUp extension (pass A left it in HBM)
//                                acc += E_g (x) key_g.{b, a}          (key inner product, both halves)
//     (+ P d0, P d1 in a relinearisation);   inverse pass B' of both sums.
//
// Before this fusion pass B wrote the transformed extensions back to HBM (beta (L+K) limb passes), the
// key inner product read them again and wrote the two sums, and inverse pass B' read those: ~950 of the
// ~3200 limb passes of DRAM traffic per HMult at logN16 (ncu, profiles/r01e).  Here the transformed
// extension never leaves the registers: a thread keeps its 16 residues of the tile plus the 2 x 16
// running sums, the tile of the NEXT digit group is fetched by the TMA unit (cp.async.bulk, mbarrier
// completion) into a double-buffered shared-memory stage while the current group's butterflies run,
// and the key residues are read with coalesced 128-bit loads (the batch index is the fastest grid
// dimension, so the ciphertexts of a chunk hit the same key tile in L2).
//
// The reference computes the same sum as beta x (ntt_radix2 + 2 mont_mult) + torch.stack +
// mont_reduce_add_many_3d (tiberate/ckks_engine.py:1365-1400, :1316-1324), then intt_radix2_exit_reduce.
// All values are residues mod q on the FP64 pipe (tb200_fast.cuh, FastF64Pol) or lazy integers
// (FastBigPol / FastSmallPol); the results of the chain are canonicalised by the following steps, so they
// are bit-identical to the reference's.
#pragma once
#include "tb200_kernels_fast.cuh"

#ifndef TB200_HOST_EMU
__device__ __forceinline__ unsigned tb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tb_mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tb_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tb_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one elected thread: arm the barrier with the byte count, then start the bulk copy global -> shared
__device__ __forceinline__ void tb_bulk_load(void* dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tb_smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tb_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(tb_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tb_mbar_wait(u64* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TB_DONE;\n"
      "bra TB_WAIT;\n"
      "TB_DONE:\n"
      "}" ::"r"(tb_smem_addr(bar)),
      "r"(parity)
      : "memory");
}
#define TB_DYN_SHARED(type, name) extern __shared__ __align__(128) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#else
// host emulation (tests/emu): the elected thread copies synchronously, the wait is a CTA barrier
static inline void tb_mbar_init(u64*, unsigned) {}
static inline void tb_mbar_init_fence() {}
static inline void tb_bulk_load(void* dst, const void* src, unsigned bytes, u64*) { std::memcpy(dst, src, bytes); }
static inline void tb_mbar_wait(u64*, unsigned) { emu_syncthreads(); }
#define TB_DYN_SHARED(type, name) static type name[2 * TB_TILE]
#endif

struct TbKsCoreArgs {
  const TbKsLevel* lv;
  TbKskDev key;
  const i64* ext;   // [nb][ng][rowsE][N]: pass-A output of every (group, limb) (own pairs: final values)
  i64* acc;         // [nb][2][rowsE][N]: inverse pass B' output (input of inverse pass A')
  const i64 *nadd0, *nadd1;  // relinearisation: NTT-domain d0 / d1, dense [nb][L][N], or null
  int p0;           // prime index of row 0 (level start)
  int row0;         // first limb row handled by this launch
  int N, rowsE, nb;
  int skip_own;     // the (group, own limb) extensions are not in ext: they are the NTT-domain key-switch input
  const i64* own;   // skip_own: that input, dense [nb][L][N], Montgomery form (relinearisation: d2)
  const TbPrime* pr;
};

#define TB_KSCORE_STAGE_BYTES (2 * TB_TILE * 8)

// F64 limbs.  grid = (nb * tiles, rows, 1), batch index fastest.
template <int LB>
__global__ void __launch_bounds__(256, 2) k_fast_ks_core_f64(TbDevFast f, TbKsCoreArgs a) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  TB_KERNEL_SHARED u64 bar[2];
  TB_DYN_SHARED(i64, stage);  // [2][tile]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int TILE = nt * 16;
  const int bt = blockIdx.x % a.nb;
  const long e0 = (long)(blockIdx.x / a.nb) * TILE;
  const int t = a.row0 + blockIdx.y, g = a.p0 + t;
  const TbFastPrime P = f.fp[g];
  const tb::FastF64Pol pol{P.qd, P.qinv};
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  const int N = a.N, ng = a.lv->ngroups;
  const long estride = (long)a.rowsE * N;
  const i64* ep = a.ext + (((long)bt * ng) * a.rowsE + t) * N + e0;  // group 0 of this (ciphertext, limb, tile)
  const long koff = (long)g * a.key.rs + e0;
  const double* twd = f.twd + ((long)g << f.logN);

  // per-group metadata once per CTA (reading it from the level table inside the loop put two dependent
  // global loads in front of every key access: ncu long_scoreboard)
  TB_KERNEL_SHARED const i64* s_kb[TB_MAXG];
  TB_KERNEL_SHARED const i64* s_ka[TB_MAXG];
  TB_KERNEL_SHARED int s_own[TB_MAXG];
  for (int i = tid; i < ng; i += nt) {
    const TbKsGroup& G = a.lv->g[i];
    s_own[i] = (a.skip_own && g >= G.src_prime0 && g < G.src_prime0 + G.alpha) ? 1 : 0;
    s_kb[i] = a.key.b[G.gid] + koff;
    s_ka[i] = a.key.a[G.gid] + koff;
  }
  auto is_own = [&](int gi) { return s_own[gi] != 0; };
  auto next_transformed = [&](int gi) {  // first group > gi whose extension still needs pass B
    ++gi;
    while (gi < ng && is_own(gi)) ++gi;
    return gi;
  };
  if (tid == 0) {
    tb_mbar_init(&bar[0], 1);
    tb_mbar_init(&bar[1], 1);
    tb_mbar_init_fence();
  }
  __syncthreads();  // barriers and the group metadata are visible
  int s = 0;
  unsigned ph0 = 0, ph1 = 0;
  {
    const int g1 = next_transformed(-1);
    if (tid == 0 && g1 < ng) tb_bulk_load(stage, ep + g1 * estride, (unsigned)TILE * 8u, &bar[0]);
  }

  double acc0[16], acc1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc0[i] = acc1[i] = 0.0;
  if (a.nadd0 != nullptr && t < a.lv->L) {  // + P d0, P d1 (NTT domain, Montgomery form like the key products)
    const long at = ((long)bt * a.lv->L + t) * N + e0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = 2 * (i * nt + tid);
      const longlong2 u0 = *reinterpret_cast<const longlong2*>(a.nadd0 + at + e);
      const longlong2 u1 = *reinterpret_cast<const longlong2*>(a.nadd1 + at + e);
      acc0[2 * i] = pol.mulmod(tb::FastF64Pol::from_int(u0.x), P.cPd);
      acc0[2 * i + 1] = pol.mulmod(tb::FastF64Pol::from_int(u0.y), P.cPd);
      acc1[2 * i] = pol.mulmod(tb::FastF64Pol::from_int(u1.x), P.cPd);
      acc1[2 * i + 1] = pol.mulmod(tb::FastF64Pol::from_int(u1.y), P.cPd);
    }
  }

  for (int gi = 0; gi < ng; ++gi) {
    i64 x[16];
    if (is_own(gi)) {
      // The digits of a group reconstruct the key-switch input modulo the group's own primes, so the
      // transformed extension at an own limb is the NTT-domain input itself (the tensor product's d2 limb):
      // leave Montgomery form, canonicalise, read in the coalesced layout directly.
      const i64* src = a.own + ((long)bt * a.lv->L + t) * N + e0;
      const TbPrime& Q = a.pr[g];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        longlong2 v = *reinterpret_cast<const longlong2*>(src + 2 * (i * nt + tid));
        v.x = tb_mr(v.x, Q.q4, Q.k);
        v.y = tb_mr(v.y, Q.q4, Q.k);
        v.x = v.x < 0 ? v.x + Q.q : v.x;
        v.y = v.y < 0 ? v.y + Q.q : v.y;
        v.x = v.x >= Q.q ? v.x - Q.q : v.x;
        v.y = v.y >= Q.q ? v.y - Q.q : v.y;
        x[2 * i] = __double_as_longlong(tb::FastF64Pol::from_int(v.x));
        x[2 * i + 1] = __double_as_longlong(tb::FastF64Pol::from_int(v.y));
      }
    } else {
      const int g2 = next_transformed(gi);
      // the other stage was consumed before the barriers of the previous transformed group
      if (tid == 0 && g2 < ng) tb_bulk_load(stage + (s ^ 1) * TILE, ep + g2 * estride, (unsigned)TILE * 8u, &bar[s ^ 1]);
      tb_mbar_wait(&bar[s], s ? ph1 : ph0);
      if (s) ph1 ^= 1; else ph0 ^= 1;
      const i64* st = stage + s * TILE;
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = st[(blk << LB) | tb::tile_x(lt, i, f0)];
      s ^= 1;
      tb::tile_fwd<LB, true>(x, sm, lt, tile, f.logN - 1, twd, pol, slot);
      tile_f64_reduce<false>(x, pol);
      // field-0 layout (16 consecutive residues per thread) -> coalesced layout through shared memory
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 16; ++i) sm[slot(tb::tile_x(lt, i, 0))] = x[i];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = 2 * (i * nt + tid);
        x[2 * i] = sm[tb::pad16(e)];
        x[2 * i + 1] = sm[tb::pad16(e + 1)];
      }
    }
    const i64* kbp = s_kb[gi];
    const i64* kap = s_ka[gi];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = 2 * (i * nt + tid);
      const longlong2 kb = *reinterpret_cast<const longlong2*>(kbp + e);
      const longlong2 ka = *reinterpret_cast<const longlong2*>(kap + e);
      const double ex = __longlong_as_double(x[2 * i]), ey = __longlong_as_double(x[2 * i + 1]);
      acc0[2 * i] = __dadd_rn(acc0[2 * i], pol.mulmod(ex, tb::FastF64Pol::from_int(kb.x)));
      acc0[2 * i + 1] = __dadd_rn(acc0[2 * i + 1], pol.mulmod(ey, tb::FastF64Pol::from_int(kb.y)));
      acc1[2 * i] = __dadd_rn(acc1[2 * i], pol.mulmod(ex, tb::FastF64Pol::from_int(ka.x)));
      acc1[2 * i + 1] = __dadd_rn(acc1[2 * i + 1], pol.mulmod(ey, tb::FastF64Pol::from_int(ka.y)));
    }
  }

  // inverse pass B' of both sums (|sum| < 12 q: reduced first), output doubles for inverse pass A'
  const double* itwd = f.itwd + ((long)g << f.logN);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    i64 x[16];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = 2 * (i * nt + tid);
      sm[tb::pad16(e)] = __double_as_longlong(h == 0 ? acc0[2 * i] : acc1[2 * i]);
      sm[tb::pad16(e + 1)] = __double_as_longlong(h == 0 ? acc0[2 * i + 1] : acc1[2 * i + 1]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[slot(tb::tile_x(lt, i, 0))];
    tile_f64_reduce<false>(x, pol);
    tb::tile_inv<LB, true>(x, sm, lt, tile, f.logN - 1, itwd, pol, slot);
    tile_f64_reduce<false>(x, pol);
    i64* d = a.acc + (((long)bt * 2 + h) * a.rowsE + t) * N + e0;
#pragma unroll
    for (int i = 0; i < 16; ++i) d[(blk << LB) | tb::tile_x(lt, i, f0)] = x[i];
  }
}
