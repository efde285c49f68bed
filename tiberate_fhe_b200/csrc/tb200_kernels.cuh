// tb200_kernels.cuh -- sm_100a kernels of the CKKS RNS hot path (device code only).
//
// Conventions shared by all kernels
//   * residues are int64; a "limb" is one row [N] of a polynomial; limb r of a tensor whose first
//     row belongs to global prime `prime0` uses TbPrime[prime0 + r];
//   * grid = (tiles along N, limbs, batch entries); one CTA never mixes primes, so the per-prime
//     constants are loaded once per CTA into registers;
//   * every kernel is HBM-streaming or INT-pipe bound: loads are coalesced (a warp touches whole
//     128-byte lines), the pointwise kernels use 128-bit vector accesses, and the NTT kernels stage
//     their tile in shared memory (tb200_ntt.cuh).
#pragma once
#include "tb200_fast.cuh"
#include "tb200_ntt.cuh"
#include "tb200_platform.h"

#define TB_MAXA 8          // max primes in one digit group (= max number of special primes)
#define TB_MAXG 32         // max digit groups (TB200_MAX_GROUPS)
#define TB_TILE 4096       // residues per NTT CTA tile
#define TB_SMEM_SLOTS (TB_TILE + TB_TILE / 16)

struct TbDev {
  const TbPrime* pr;   // [P]
  const u64* psi4;     // [P][N] forward twiddles, lazy Montgomery form, pre-shifted by 2 (the last-round
                       // stages of pass B stored transposed per tile, see tb::fwd_round<PERM>)
  const u64* ipsi4;    // [P][N] inverse twiddles
  int logN, LA, LB, P;
  // forward transforms of the 40-bit limbs on the FP64 pipe with the reference's exact representatives
  // (ExactF64Pol): enabled per context, per prime by TbFastPrime.f64
  const TbFastPrime* fp;
  const double* twd;
  const double* itwd;  // inverse twiddles, centred doubles (deferred-reduction inverse kernels)
  int x64;
};

// strided view of a batched polynomial
struct TbView {
  i64* p;
  long bs, rs;  // batch / row stride in elements
  __device__ __forceinline__ i64* row(int b, int r) const { return p + (long)b * bs + (long)r * rs; }
};

__device__ __forceinline__ tb::PrimeRegs load_prime(const TbPrime* pr, int g) {
  tb::PrimeRegs p;
  p.q = pr[g].q;
  p.q2 = pr[g].q2;
  p.q4 = pr[g].q4;
  p.k = pr[g].k;
  return p;
}

// =====================================================================================
// Pointwise family
// =====================================================================================
struct TbPwArgs {
  TbView a, b, out;
  const i64* scal;                      // per-row scalar (device) or null
  const i64 *ql, *qh, *kl, *kh, *two_q; // explicit constants (legacy ops) or null
  int prime0, N;
};

// Slots of the reference's right-aligned 64-entry constant pool below its first prime are zero padding
// (constant_mem_context.py:100-123).  A call whose rows + sp_prime_len exceed the number of primes reads them
// (decrypt_triplet's s2 = mont_mult(sk, sk) does, ckks_engine.py:669, at levels < K), and with q = k = 0 the
// reference's product degenerates to floor(a b / 2^62) (mont_scalar_kernel.cuh:9-58 with s = 0) -- without
// the "+ [low bits != 0]" carry that the closed form adds for a real prime.
__device__ __forceinline__ i64 tb_mm_zero_slot(i64 a, i64 b) {
  const u64 lo = (u64)a * (u64)b;
  return (i64)(((u64)__mul64hi(a, b) << 2) | (lo >> 62));
}

template <int OP>
__device__ __forceinline__ i64 pw_apply(i64 a, i64 b, i64 s, i64 q, i64 q2, u64 q4, u64 k, i64 Rs, i64 Rss) {
  if (q4 == 0) {  // zero-padding slot of the reference's pool: every constant is 0
    if constexpr (OP == 0) return tb_mm_zero_slot(a, b);
    if constexpr (OP == 5 || OP == 14) return tb_mm_zero_slot(a, s);
    if constexpr (OP == 6 || OP == 7) return 0;
    if constexpr (OP == 8) return a >> 62;
    if constexpr (OP == 13) return b >> 62;
  }
  if constexpr (OP == 0) return tb_mm_ss(a, b, q4, k);
  if constexpr (OP == 1) return tb_add(a, b, q2);
  if constexpr (OP == 2) return tb_sub(a, b, q2);
  if constexpr (OP == 3) return tb_cs1(tb_add(a, b, q2), q);
  if constexpr (OP == 4) return tb_cs1(tb_sub(a, b, q2), q);
  if constexpr (OP == 5) return tb_mm_ss(a, s, q4, k);
  if constexpr (OP == 6) return tb_mm_ss(a, Rs, q4, k);
  if constexpr (OP == 7) return tb_mm_ss(a, Rss, q4, k);
  if constexpr (OP == 8) return tb_mr(a, q4, k);
  if constexpr (OP == 9) return tb_cs1(a, q);
  if constexpr (OP == 10) return tb_signed(a, q);
  if constexpr (OP == 11) return a + q;
  if constexpr (OP == 12) return a + q;
  if constexpr (OP == 13) return tb_cs1(tb_mr(tb_add(tb_mm_ss(a, Rs, q4, k), b, q2), q4, k), q);
  if constexpr (OP == 14) return tb_cs1(tb_mm_ss(a, s, q4, k), q);
  if constexpr (OP == 15) return a;  // plain copy (internal)
  return 0;
}

template <int OP>
__global__ void __launch_bounds__(256) k_pointwise(TbDev c, TbPwArgs g) {
  constexpr bool HAS_B = (OP <= 4) || (OP == 13);
  const int r = blockIdx.y, bt = blockIdx.z;
  i64 q, q2, Rs = 0, Rss = 0;
  u64 q4, k;
  if (g.ql != nullptr || g.two_q != nullptr) {
    if (g.two_q != nullptr) {
      q2 = g.two_q[r];
      q = q2 >> 1;
    } else {
      q = g.ql[r] + (g.qh[r] << 31);
      q2 = q << 1;
    }
    q4 = (u64)q << 2;
    k = (g.kl != nullptr) ? (u64)(g.kl[r] + (g.kh[r] << 31)) : 0ull;
  } else if (g.prime0 + r < 0) {
    q = q2 = 0;
    q4 = k = 0;
  } else {
    const TbPrime& P = c.pr[g.prime0 + r];
    q = P.q;
    q2 = P.q2;
    q4 = P.q4;
    k = P.k;
    Rs = P.Rs;
    Rss = P.Rs_scale;
  }
  const i64 s = (g.scal != nullptr) ? g.scal[r] : 0;
  const i64* pa = g.a.row(bt, r);
  const i64* pb = HAS_B ? g.b.row(bt, r) : nullptr;
  i64* po = g.out.row(bt, r);
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j < g.N) {
    const longlong2 va = *reinterpret_cast<const longlong2*>(pa + j);
    longlong2 vb;
    vb.x = vb.y = 0;
    if constexpr (HAS_B) vb = *reinterpret_cast<const longlong2*>(pb + j);
    longlong2 vo;
    vo.x = pw_apply<OP>(va.x, vb.x, s, q, q2, q4, k, Rs, Rss);
    vo.y = pw_apply<OP>(va.y, vb.y, s, q, q2, q4, k, Rs, Rss);
    *reinterpret_cast<longlong2*>(po + j) = vo;
  }
}

// ---- packed wire format of the scale-prime limbs ------------------------------------------------------------
// The scale primes lie within 2^32 of 2^40, on both sides, so a canonical residue needs 41 bits; the int64 layout
// moves 64.  On the host <-> device link (PCIe, the bound of every end-to-end figure) a limb travels as
//     N x 5 bytes  (bits 0..39 of residue j at byte 5 j, little endian)  +  N / 8 bytes  (bit 40 of residue j = bit
//     j % 8 of byte j / 8)                                                           = 5.125 bytes per residue.
// A thread converts 8 residues = five 8-byte words + one byte <-> eight int64.
__global__ void __launch_bounds__(256) k_unpack41(const unsigned char* src, long src_bs, long src_rs, TbView dst, int N) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (j >= N) return;
  const unsigned char* row = src + (long)bt * src_bs + (long)r * src_rs;
  const u64* p = reinterpret_cast<const u64*>(row + (long)j * 5);
  u64 w[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) w[i] = p[i];
  const u64 hb = row[(long)N * 5 + (j >> 3)];
  i64* d = dst.row(bt, r) + j;
  const u64 M = (1ull << 40) - 1;
  u64 v[8];
  v[0] = w[0] & M;
  v[1] = ((w[0] >> 40) | (w[1] << 24)) & M;
  v[2] = (w[1] >> 16) & M;
  v[3] = ((w[1] >> 56) | (w[2] << 8)) & M;
  v[4] = ((w[2] >> 32) | (w[3] << 32)) & M;
  v[5] = (w[3] >> 8) & M;
  v[6] = ((w[3] >> 48) | (w[4] << 16)) & M;
  v[7] = w[4] >> 24;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    longlong2 o;
    o.x = (i64)(v[i] | (((hb >> i) & 1ull) << 40));
    o.y = (i64)(v[i + 1] | (((hb >> (i + 1)) & 1ull) << 40));
    *reinterpret_cast<longlong2*>(d + i) = o;
  }
}
__global__ void __launch_bounds__(256) k_pack41(TbView src, unsigned char* dst, long dst_bs, long dst_rs, int N) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (j >= N) return;
  const i64* s = src.row(bt, r) + j;
  const u64 M = (1ull << 40) - 1;
  u64 v[8], hb = 0;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const longlong2 t = *reinterpret_cast<const longlong2*>(s + i);
    hb |= (((u64)t.x >> 40) & 1ull) << i;
    hb |= (((u64)t.y >> 40) & 1ull) << (i + 1);
    v[i] = (u64)t.x & M;
    v[i + 1] = (u64)t.y & M;
  }
  unsigned char* row = dst + (long)bt * dst_bs + (long)r * dst_rs;
  u64* p = reinterpret_cast<u64*>(row + (long)j * 5);
  p[0] = v[0] | (v[1] << 40);
  p[1] = (v[1] >> 24) | (v[2] << 16) | (v[3] << 56);
  p[2] = (v[3] >> 8) | (v[4] << 32);
  p[3] = (v[4] >> 32) | (v[5] << 8) | (v[6] << 48);
  p[4] = (v[6] >> 16) | (v[7] << 24);
  row[(long)N * 5 + (j >> 3)] = (unsigned char)hb;
}

// mont_(reduce_)add_many_3d: in [K][rows][N] dense -> out [rows][N]
__global__ void __launch_bounds__(256) k_add_many(TbDev c, const i64* in, i64* out, int K, int rows, int N,
                                                  int prime0, int pairwise) {
  const int r = blockIdx.y;
  const i64 q2 = c.pr[prime0 + r].q2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const long rn = (long)rows * N;
  const i64* p = in + (long)r * N + j;
  i64 acc = 0;
  if (pairwise) {
    for (int kk = 0; kk < K / 2; ++kk) acc = tb_add(acc, tb_add(p[(2 * kk) * rn], p[(2 * kk + 1) * rn], q2), q2);
    if (K & 1) acc = tb_add(acc, p[(long)(K - 1) * rn], q2);
  } else {
    for (int kk = 0; kk < K; ++kk) acc = tb_add(acc, p[(long)kk * rn], q2);
  }
  out[(long)r * N + j] = acc;
}

// =====================================================================================
// NTT passes.  grid = (tiles, limbs, batch); 256 threads when N >= 4096.
// =====================================================================================
// prologue / epilogue selectors
#define TB_PRO_NONE 0
#define TB_PRO_ENTER 1   // x <- MM(x, R^2)            (enter_ntt_radix2)
#define TB_EPI_NINV 0    // x <- MM(x, N^-1 R)         (intt_radix2)
#define TB_EPI_EXIT 1    // ... ; MR                   (intt_radix2_exit)
#define TB_EPI_EXIT_REDUCE 2   // ... ; CS1            (intt_radix2_exit_reduce)
#define TB_EPI_EXIT_SIGNED 3   // ... ; make_signed    (intt_radix2_exit_reduce_signed)

// forward pass A: stages 0..LA-1 on a tile of 2^LA rows x W columns (row = index >> LB).
__device__ __noinline__ double exact_enter_cold(i64 x, i64 Rs, u64 q4, u64 k) {
  return tb::FastF64Pol::from_int(tb_mm_ss(x, Rs, q4, k));
}
// CTA-uniform: may this tile take ExactF64Pol?  (prime on the FP64 pipe and every residue below 2^50 in magnitude)
__device__ __forceinline__ bool tile_fits_f64(const TbDev& c, int g, const i64 (&x)[16]) {
  if (!c.x64 || !c.fp[g].f64) return false;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 16; ++i) bad |= (x[i] >= (1ll << 50)) | (x[i] <= -(1ll << 50));
  return !tb_block_any(bad);  // barrier + OR over the CTA (no shared flag written by several threads)
}
// CTA-uniform: every residue of the tile in [0, hi)?  (hi = 2q: the lazy range the ExactSumPol shortcut needs)
__device__ __forceinline__ bool tile_in_range(const i64 (&x)[16], i64 hi) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 16; ++i) bad |= (x[i] < 0) | (x[i] >= hi);
  return !tb_block_any(bad);
}
__device__ __forceinline__ tb::ExactSumPol exact_sum_policy(const TbDev& c, int g, unsigned* minhi) {
  const TbFastPrime& F = c.fp[g];
  tb::ExactSumPol pol;
  pol.f = tb::FastF64Pol{F.qd, F.qinv};
  pol.q2 = F.qd + F.qd;
  pol.inv2q = 0.5 * F.qinv;
  pol.xbmax = pol.q2 * F.qd * 2.168404344971008868e-19 + 4.0;  // |O| < 2q: floor(O S / 2^62) < 2q q 2^-62 + 1
  pol.minhi = minhi;
  return pol;
}
__device__ __forceinline__ tb::ExactF64Pol exact_f64_policy(const TbDev& c, int g, const tb::PrimeRegs& p, const u64* psi4) {
  const TbFastPrime& F = c.fp[g];
  tb::ExactF64Pol pol;
  pol.f = tb::FastF64Pol{F.qd, F.qinv};
  pol.p = p;
  pol.q2 = F.qd + F.qd;
  pol.xbs = F.qd * 2.168404344971008868e-19;  // 2^-62
  pol.twd = c.twd + ((long)g << c.logN);
  pol.psi4 = psi4 + ((long)g << c.logN);
  return pol;
}

// CTA index into the "left for the generic kernel" flag array of the deferred-reduction kernels below
__device__ __forceinline__ long tile_flag_index() {
  return ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
}

// todo != nullptr: second launch after k_ntt_fwd_A_sum -- only the tiles that kernel flagged are transformed
template <int LA, int PRO>
__global__ void __launch_bounds__(256, 3) k_ntt_fwd_A(TbDev c, TbView src, TbView dst, int prime0, int LW,
                                                      const unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  if (todo != nullptr && todo[tile_flag_index()] == 0) return;
  const int W = 1 << LW;
  const int col = threadIdx.x & (W - 1), tr = threadIdx.x >> LW;
  const int limb = blockIdx.y, g = prime0 + limb;
  const tb::PrimeRegs p = load_prime(c.pr, g);
  const i64* s = src.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  i64* d = dst.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  constexpr int f0 = tb::fwd_field<LA>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[(long)tb::tile_x(tr, i, f0) << c.LB];
  auto slot = [&](int lx) { return tb::pad16((lx << LW) | col); };
  if (tile_fits_f64(c, g, x)) {
    const tb::ExactF64Pol pol = exact_f64_policy(c, g, p, c.psi4);
    const i64 Rs = c.pr[g].Rs;
    const double Rc = c.fp[g].Rcd;  // R mod q, centred
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double v = tb::FastF64Pol::from_int(x[i]);
      if constexpr (PRO == TB_PRO_ENTER) {  // MM(x, R^2) with its exact representative (see ExactF64Pol)
        double r = pol.f.mulmod(v, Rc);
        r = r < 0.0 ? __dadd_rn(r, pol.f.q) : r;
        r = r >= pol.f.q ? __dadd_rn(r, -pol.f.q) : r;
        const double xb = __fma_rn(v < 0.0 ? -v : v, pol.xbs, 2.0);
        if (!(r > xb && r < __dadd_rn(pol.f.q, -xb))) r = exact_enter_cold(x[i], Rs, p.q4, p.k);
        v = r;
      }
      x[i] = __double_as_longlong(v);
    }
    tb::tile_fwd<LA>(x, sm, tr, 0, LA - 1, pol.twd, pol, slot);
#pragma unroll
    for (int i = 0; i < 16; ++i) d[(long)tb::tile_x(tr, i, 0) << c.LB] = tb::FastF64Pol::to_int(__longlong_as_double(x[i]));
    return;
  }
  if constexpr (PRO == TB_PRO_ENTER) {
    const i64 Rs = c.pr[g].Rs;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = tb_mm_ss(x[i], Rs, p.q4, p.k);
  }
  tb::tile_fwd<LA>(x, sm, tr, 0, LA - 1, c.psi4 + ((long)g << c.logN), tb::ExactPol{p}, slot);
#pragma unroll
  for (int i = 0; i < 16; ++i) d[(long)tb::tile_x(tr, i, 0) << c.LB] = x[i];
}

// forward pass B: stages LA..logN-1 on contiguous 2^LB blocks; a CTA owns TE = blockDim*16 residues.
template <int LB>
__global__ void __launch_bounds__(256, 3) k_ntt_fwd_B(TbDev c, TbView src, TbView dst, int prime0,
                                                      const unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  if (todo != nullptr && todo[tile_flag_index()] == 0) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb;
  const tb::PrimeRegs p = load_prime(c.pr, g);
  const long e0 = (long)blockIdx.x * nt * 16;
  const i64* s = src.row(blockIdx.z, limb) + e0;
  i64* d = dst.row(blockIdx.z, limb) + e0;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) sm[tb::pad16(i * nt + tid)] = s[i * nt + tid];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = sm[slot(tb::tile_x(lt, i, f0))];
  if (tile_fits_f64(c, g, x)) {
    const tb::ExactF64Pol pol = exact_f64_policy(c, g, p, c.psi4);
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __double_as_longlong(tb::FastF64Pol::from_int(x[i]));
    tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, pol.twd, pol, slot);
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = tb::FastF64Pol::to_int(__longlong_as_double(x[i]));
  } else {
    tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, c.psi4 + ((long)g << c.logN), tb::ExactPol{p}, slot);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) sm[slot(tb::tile_x(lt, i, 0))] = x[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i * nt + tid] = sm[tb::pad16(i * nt + tid)];
}

// ---- deferred-reduction forward transforms (ExactSumPol) ---------------------------------------------------
// Same grids as k_ntt_fwd_A / k_ntt_fwd_B.  A tile is transformed here iff its prime takes the FP64 route, every
// input lies in the lazy range and no product fell into the representative band; otherwise the CTA leaves the
// tile untouched and sets todo[tile], and the generic kernel (launched right after with the same flag array)
// transforms it.  Without the integer and per-butterfly-test code paths these kernels fit 64 registers.
template <int LA, int PRO>
__global__ void __launch_bounds__(256, 4) k_ntt_fwd_A_sum(TbDev c, TbView src, TbView dst, int prime0, int LW,
                                                          unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int W = 1 << LW;
  const int col = threadIdx.x & (W - 1), tr = threadIdx.x >> LW;
  const int limb = blockIdx.y, g = prime0 + limb;
  const TbFastPrime& F = c.fp[g];
  if (!c.x64 || !F.f64) {
    if (threadIdx.x == 0) todo[tile_flag_index()] = 1;
    return;
  }
  const i64* s = src.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  i64* d = dst.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  constexpr int f0 = tb::fwd_field<LA>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[(long)tb::tile_x(tr, i, f0) << c.LB];
  unsigned minhi = 0x7fffffffu;
  const tb::ExactSumPol sp = exact_sum_policy(c, g, &minhi);
  // |V| < xbmax is implied by hiword(|V|) <= hiword(xbmax): test the (slightly wider) word-granular band
  const unsigned band_hi = (unsigned)((u64)__double_as_longlong(sp.xbmax) >> 32) + 1u;
  // enter: non-negative inputs below 2^50 -- MM(x, R^2) then lies in [0, q + small] and equals the canonical
  // residue outside the band (tested on |r| like the butterflies); plain: lazy values in [0, 2q)
  bool bad = false;
  const i64 hi = PRO == TB_PRO_ENTER ? (1ll << 50) : (i64)(2 * F.q);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bad |= (x[i] < 0) | (x[i] >= hi);
    double v = tb::FastF64Pol::from_int(x[i]);
    if constexpr (PRO == TB_PRO_ENTER) {
      double r = sp.f.mulmod(v, F.Rcd);
      const double xb = __fma_rn(v, sp.f.q * 2.168404344971008868e-19, 2.0);  // floor(x R^2 / 2^62) <= x q 2^-62 + 1
      bad |= (r < 0.0 ? -r : r) < xb;
      v = r < 0.0 ? __dadd_rn(r, sp.f.q) : r;
    }
    x[i] = __double_as_longlong(v);
  }
  auto slot = [&](int lx) { return tb::pad16((lx << LW) | col); };
  tb::tile_fwd<LA>(x, sm, tr, 0, LA - 1, c.twd + ((long)g << c.logN), sp, slot);
  if (tb_block_any(bad | (minhi < band_hi))) {
    if (threadIdx.x == 0) todo[tile_flag_index()] = 1;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) d[(long)tb::tile_x(tr, i, 0) << c.LB] = sp.finish(x[i]);
}

template <int LB>
__global__ void __launch_bounds__(256, 4) k_ntt_fwd_B_sum(TbDev c, TbView src, TbView dst, int prime0,
                                                          unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb;
  const TbFastPrime& F = c.fp[g];
  if (!c.x64 || !F.f64) {
    if (tid == 0) todo[tile_flag_index()] = 1;
    return;
  }
  const long e0 = (long)blockIdx.x * nt * 16;
  const i64* s = src.row(blockIdx.z, limb) + e0;
  i64* d = dst.row(blockIdx.z, limb) + e0;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[(blk << LB) | tb::tile_x(lt, i, f0)];
  unsigned minhi = 0x7fffffffu;
  const tb::ExactSumPol sp = exact_sum_policy(c, g, &minhi);
  // |V| < xbmax is implied by hiword(|V|) <= hiword(xbmax): test the (slightly wider) word-granular band
  const unsigned band_hi = (unsigned)((u64)__double_as_longlong(sp.xbmax) >> 32) + 1u;
  bool bad = false;
  const i64 hi = (i64)(2 * F.q);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bad |= (x[i] < 0) | (x[i] >= hi);
    x[i] = __double_as_longlong(tb::FastF64Pol::from_int(x[i]));
  }
  tb::tile_fwd<LB, true>(x, sm, lt, tile, c.logN - 1, c.twd + ((long)g << c.logN), sp, slot);
  if (tb_block_any(bad | (minhi < band_hi))) {
    if (tid == 0) todo[tile_flag_index()] = 1;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = sp.finish(x[i]);
  // field-0 layout -> coalesced 128-bit stores through shared memory (as k_fast_fwd_B)
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

// inverse pass B': inverse stages 0..LB-1 (distances 1..2^(LB-1)) on contiguous blocks.
template <int LB>
__global__ void __launch_bounds__(256) k_ntt_inv_B(TbDev c, TbView src, TbView dst, int prime0, const unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  if (todo != nullptr && todo[tile_flag_index()] == 0) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb;
  const tb::PrimeRegs p = load_prime(c.pr, g);
  const long e0 = (long)blockIdx.x * nt * 16;
  const i64* s = src.row(blockIdx.z, limb) + e0;
  i64* d = dst.row(blockIdx.z, limb) + e0;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) sm[tb::pad16(i * nt + tid)] = s[i * nt + tid];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = sm[slot(tb::tile_x(lt, i, 0))];
  tb::tile_inv<LB, true>(x, sm, lt, tile, c.logN - 1, c.ipsi4 + ((long)g << c.logN), tb::ExactPol{p}, slot);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) sm[slot(tb::tile_x(lt, i, f0))] = x[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i * nt + tid] = sm[tb::pad16(i * nt + tid)];
}

template <int EPI>
__device__ __forceinline__ i64 intt_epilogue(i64 x, i64 Ninv, const tb::PrimeRegs& p) {
  x = tb_mm_ss(x, Ninv, p.q4, p.k);
  if constexpr (EPI >= TB_EPI_EXIT) x = tb_mr(x, p.q4, p.k);
  if constexpr (EPI >= TB_EPI_EXIT_REDUCE) x = tb_cs1(x, p.q);
  if constexpr (EPI >= TB_EPI_EXIT_SIGNED) x = tb_signed(x, p.q);
  return x;
}

// inverse pass A': inverse stages LB..logN-1 on 2^LA rows x W columns, then the N^-1 epilogue.
template <int LA, int EPI>
__global__ void __launch_bounds__(256) k_ntt_inv_A(TbDev c, TbView src, TbView dst, int prime0, int LW,
                                                   const unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  if (todo != nullptr && todo[tile_flag_index()] == 0) return;
  const int W = 1 << LW;
  const int col = threadIdx.x & (W - 1), tr = threadIdx.x >> LW;
  const int limb = blockIdx.y, g = prime0 + limb;
  const tb::PrimeRegs p = load_prime(c.pr, g);
  const i64* s = src.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  i64* d = dst.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  constexpr int f0 = tb::fwd_field<LA>(0);
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s[(long)tb::tile_x(tr, i, 0) << c.LB];
  auto slot = [&](int lx) { return tb::pad16((lx << LW) | col); };
  tb::tile_inv<LA>(x, sm, tr, 0, LA - 1, c.ipsi4 + ((long)g << c.logN), tb::ExactPol{p}, slot);
  const i64 Ninv = c.pr[g].Ninv;
#pragma unroll
  for (int i = 0; i < 16; ++i) d[(long)tb::tile_x(tr, i, f0) << c.LB] = intt_epilogue<EPI>(x[i], Ninv, p);
}

// ---- deferred-reduction inverse transforms (ExactSumPol::gs), all four exits --------------------------------
// As the forward pair: FP64 limbs with lazy inputs in [0, 2q) only; other tiles are flagged for the generic kernels.
template <int LB>
__global__ void __launch_bounds__(256, 4) k_ntt_inv_B_sum(TbDev c, TbView src, TbView dst, int prime0,
                                                          unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int limb = blockIdx.y, g = prime0 + limb;
  const TbFastPrime& F = c.fp[g];
  if (!c.x64 || !F.f64) {
    if (tid == 0) todo[tile_flag_index()] = 1;
    return;
  }
  const long e0 = (long)blockIdx.x * nt * 16;
  const i64* s = src.row(blockIdx.z, limb) + e0;
  i64* d = dst.row(blockIdx.z, limb) + e0;
  const int blk = tid >> (LB - 4), lt = tid & ((1 << (LB - 4)) - 1);
  const int tile = (int)(e0 >> LB) + blk;
  auto slot = [&](int lx) { return tb::pad16((blk << LB) | lx); };
  constexpr int f0 = tb::fwd_field<LB>(0);
  // coalesced 128-bit loads, transposed through shared memory into the field-0 layout (as k_fast_inv_B)
  const longlong2* sv = reinterpret_cast<const longlong2*>(s);
  bool bad = false;
  const i64 hi = (i64)(2 * F.q);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const longlong2 v = sv[i * nt + tid];
    const int e = 2 * (i * nt + tid);
    bad |= (v.x < 0) | (v.x >= hi) | (v.y < 0) | (v.y >= hi);
    sm[tb::pad16(e)] = v.x;
    sm[tb::pad16(e + 1)] = v.y;
  }
  __syncthreads();
  i64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = __double_as_longlong(tb::FastF64Pol::from_int(sm[slot(tb::tile_x(lt, i, 0))]));
  unsigned minhi = 0x7fffffffu;
  const tb::ExactSumPol sp = exact_sum_policy(c, g, &minhi);
  const unsigned band_hi = (unsigned)((u64)__double_as_longlong(sp.xbmax) >> 32) + 1u;
  tb::tile_inv<LB, true>(x, sm, lt, tile, c.logN - 1, c.itwd + ((long)g << c.logN), sp, slot);
  if (tb_block_any(bad | (minhi < band_hi))) {
    if (tid == 0) todo[tile_flag_index()] = 1;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) d[(blk << LB) | tb::tile_x(lt, i, f0)] = sp.finish(x[i]);
}

template <int LA, int EPI>
__global__ void __launch_bounds__(256, 4) k_ntt_inv_A_sum(TbDev c, TbView src, TbView dst, int prime0, int LW,
                                                          unsigned char* todo) {
  TB_KERNEL_SHARED i64 sm[TB_SMEM_SLOTS];
  const int W = 1 << LW;
  const int col = threadIdx.x & (W - 1), tr = threadIdx.x >> LW;
  const int limb = blockIdx.y, g = prime0 + limb;
  const TbFastPrime& F = c.fp[g];
  if (!c.x64 || !F.f64) {
    if (threadIdx.x == 0) todo[tile_flag_index()] = 1;
    return;
  }
  const i64* s = src.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  i64* d = dst.row(blockIdx.z, limb) + (long)blockIdx.x * W + col;
  constexpr int f0 = tb::fwd_field<LA>(0);
  i64 x[16];
  bool bad = false;
  const i64 hi = (i64)(2 * F.q);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const i64 v = s[(long)tb::tile_x(tr, i, 0) << c.LB];
    bad |= (v < 0) | (v >= hi);
    x[i] = __double_as_longlong(tb::FastF64Pol::from_int(v));
  }
  unsigned minhi = 0x7fffffffu;
  const tb::ExactSumPol sp = exact_sum_policy(c, g, &minhi);
  const unsigned band_hi = (unsigned)((u64)__double_as_longlong(sp.xbmax) >> 32) + 1u;
  auto slot = [&](int lx) { return tb::pad16((lx << LW) | col); };
  tb::tile_inv<LA>(x, sm, tr, 0, LA - 1, c.itwd + ((long)g << c.logN), sp, slot);
  // epilogue (intt_epilogue): MM(y, N^-1 R) is the canonical residue of y N^-1 outside the band; MR of a canonical
  // residue is the canonical residue of its product with R^-1; then CS1 / make_signed act on canonical values
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double v = sp.scale(sp.finish_d(x[i]), F.nid);  // EPI NINV: stays in Montgomery form
    if constexpr (EPI >= TB_EPI_EXIT) {
      // MR(v) = v R^-1 mod q, canonical for canonical v: one more exact product (R^-1 = ex N mod q is not stored;
      // v N^-1 R^-1 = y ex N^-1 ... use the stored exit constant on the pre-scale value instead)
      v = sp.f.mulmod(sp.finish_d(x[i]), F.exd);
      v = v < 0.0 ? __dadd_rn(v, sp.f.q) : v;
      v = v >= sp.f.q ? __dadd_rn(v, -sp.f.q) : v;
      if constexpr (EPI >= TB_EPI_EXIT_SIGNED) v = (v <= (double)((i64)F.q >> 1)) ? v : __dadd_rn(v, -sp.f.q);
    }
    r[i] = v;
  }
  if (tb_block_any(bad | (minhi < band_hi))) {
    if (threadIdx.x == 0) todo[tile_flag_index()] = 1;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) d[(long)tb::tile_x(tr, i, f0) << c.LB] = tb::FastF64Pol::to_int(r[i]);
}

// =====================================================================================
// Fused HE kernels
// =====================================================================================

// rescale (he_fused_cuda.cu:99-142 / :190-230): out[r][j] = CS1(MM(in[r][j] - resc[j], scale_r) + rnd)
// `scales` is indexed by row; resc is the dropped limb.
__global__ void __launch_bounds__(256) k_rescale(TbDev c, TbView in, TbView resc, TbView out, const i64* scales,
                                                 int prime0, int N, i64 round_at, int exact) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const TbPrime& P = c.pr[prime0 + r];
  const i64 sc = scales[r];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const longlong2 va = *reinterpret_cast<const longlong2*>(in.row(bt, r) + j);
  const longlong2 vr = *reinterpret_cast<const longlong2*>(resc.row(bt, 0) + j);
  longlong2 vo;
  i64 x = tb_mm_ss(va.x - vr.x, sc, P.q4, P.k);
  if (exact) x += (vr.x > round_at) ? 1 : 0;
  vo.x = tb_cs1(x, P.q);
  x = tb_mm_ss(va.y - vr.y, sc, P.q4, P.k);
  if (exact) x += (vr.y > round_at) ? 1 : 0;
  vo.y = tb_cs1(x, P.q);
  *reinterpret_cast<longlong2*>(out.row(bt, r) + j) = vo;
}

// ModUp "extend" as the reference op (he_fused_cuda.cu:276-312): one digit group.
__global__ void __launch_bounds__(256) k_extend_op(TbDev c, const i64* state, long state_stride, int alpha,
                                                   const i64* l_enter, long le_stride, long le_off, i64* out,
                                                   long out_stride, int prime0, int N) {
  const int r = blockIdx.y;
  const TbPrime& P = c.pr[prime0 + r];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  i64 x = tb_mm_ss(state[j], P.Rs, P.q4, P.k);
  for (int kk = 0; kk + 1 < alpha; ++kk) {
    const i64 y = tb_mm_ss(state[(long)(kk + 1) * state_stride + j], l_enter[(long)kk * le_stride + le_off + r],
                           P.q4, P.k);
    x = tb_add(x, y, P.q2);
  }
  out[(long)r * out_stride + j] = x;
}

// Galois automorphism in the coefficient domain (he_fused_cuda.cu:361-391), explicit perm table.
__global__ void __launch_bounds__(256) k_codec_rotate_op(TbView a, const i64* perm, const i64* two_q, TbView out,
                                                         int N) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const i64 q = two_q[r] >> 1;
  const i64 pm = perm[j];
  const i64 folded = pm % N;
  const i64 sign = ((pm / N) & 1) ? -1 : 1;
  i64 x = a.row(bt, r)[j] * sign;
  x = x + q;
  out.row(bt, r)[folded] = tb_cs1(x, q);
}

// Same automorphism with perm[j] = galois*j mod 2N computed on the fly (engine layer).
__global__ void __launch_bounds__(256) k_automorphism(TbDev c, TbView a, TbView out, int prime0, i64 galois) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const int N = 1 << c.logN;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const i64 q = c.pr[prime0 + r].q;
  const unsigned pm = (unsigned)(((u64)galois * (u64)j) & (u64)(2 * N - 1));
  const int folded = pm & (N - 1);
  i64 x = a.row(bt, r)[j];
  if (pm & (unsigned)N) x = -x;
  x = x + q;
  out.row(bt, r)[folded] = tb_cs1(x, q);
}

// ---- key-switch tables -------------------------------------------------------------------------
struct TbKsGroup {       // one digit group at one level
  int alpha;             // primes alive in the group
  int state_row0;        // first row of the group's digits inside the digit-state buffer
  int gid;               // global group id (index into the key-switch key)
  int src_row0;          // owner only: row of the group's first prime inside the (local) input tensor, else -1
  int src_prime0;        // owner only: index of that prime in the context's prime table
  int wide_mask;         // bit k: the group's k-th alive prime is >= 2^42 (its digits do not fit a double)
  long lenter_off;       // offset of this group's L_enter block [(alpha-1)][P] in TbKsTables.lenter
  i64 Y[TB_MAXA];        // Y[i] = (L_i^-1 mod m_{i+1}) R mod m_{i+1}
  i64 Lsc[TB_MAXA][TB_MAXA];  // Lsc[i][j] = L_i R mod m_j for j >= i+2
};
struct TbKsLevel {
  int ngroups, L;        // alive digit groups (all ranks'), local ordinary rows at this level
  int nown;              // groups whose digits this rank computes
  int state_rows;        // rows of the digit-state buffer (= world * seg_rows; owner-major layout)
  int seg_rows;          // rows of one rank's segment
  int pad;
  int own[TB_MAXG];      // indices into g[] of the owned groups
  int nforeign;          // groups whose digits arrive from other ranks (limb sharding), indices in foreign[]
  int foreign[TB_MAXG];
  TbKsGroup g[TB_MAXG];
};

// ModUp digits (pre_extend, ckks_engine.py:889-921): one thread per (coefficient pair, group).
// a: [L][N] canonical coefficient domain; state: [L][N] (digit rows, same row numbering).
// W residues per thread (W = 2: 128-bit accesses), A = compile-time bound on the group size: the A source limbs
// are loaded before the dependent chain of Montgomery products starts.
template <int A, int W>
__device__ __forceinline__ void digits_body(const TbDev& c, const TbKsGroup& G, const TbView& a, const TbView& st, int bt,
                                            int j) {
  const int alpha = G.alpha;
  i64 in[A][W], s[A][W];
#pragma unroll
  for (int i = 0; i < A; ++i)
    if (i < alpha) {
      if constexpr (W == 2) {
        const longlong2 v = *reinterpret_cast<const longlong2*>(a.row(bt, G.src_row0 + i) + j);
        in[i][0] = v.x;
        in[i][1] = v.y;
      } else {
        in[i][0] = a.row(bt, G.src_row0 + i)[j];
      }
    }
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int w = 0; w < W; ++w) s[i][w] = in[0][w];
#pragma unroll
  for (int i = 0; i < A - 1; ++i) {
    if (i + 1 < alpha) {
      const TbPrime& P1 = c.pr[G.src_prime0 + i + 1];
      i64 Y[W];
#pragma unroll
      for (int w = 0; w < W; ++w) s[i + 1][w] = Y[w] = tb_mm_ss(in[i + 1][w] - s[i + 1][w], G.Y[i], P1.q4, P1.k);
#pragma unroll
      for (int jj = i + 2; jj < A; ++jj) {
        if (jj < alpha) {
          const TbPrime& Pj = c.pr[G.src_prime0 + jj];
#pragma unroll
          for (int w = 0; w < W; ++w) s[jj][w] += tb_mm_ss(Y[w], G.Lsc[i][jj], Pj.q4, Pj.k);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < A; ++i)
    if (i < alpha) {
      if constexpr (W == 2) {
        longlong2 o;
        o.x = s[i][0];
        o.y = s[i][1];
        *reinterpret_cast<longlong2*>(st.row(bt, G.state_row0 + i) + j) = o;
      } else {
        st.row(bt, G.state_row0 + i)[j] = s[i][0];
      }
    }
}
__global__ void __launch_bounds__(256) k_digits(TbDev c, const TbKsLevel* lv, TbView a, TbView st, int N) {
  const TbKsGroup& G = lv->g[lv->own[blockIdx.y]];
  const int bt = blockIdx.z;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  if (G.alpha <= 4) {  // CTA-uniform
    digits_body<4, 2>(c, G, a, st, bt, j);
  } else {
    digits_body<TB_MAXA, 1>(c, G, a, st, bt, j);
    digits_body<TB_MAXA, 1>(c, G, a, st, bt, j + 1);
  }
}

// ModUp extend for every group of the level (he_fused_cuda.cu:298-311):
// ext[bt][gi][t][j], t over the L+K limbs of the extended basis (prime level + t).
__global__ void __launch_bounds__(256) k_extend_all(TbDev c, const TbKsLevel* lv, const i64* lenter, TbView st,
                                                    i64* ext, int level, int N, int rowsE) {
  const int t = blockIdx.y;
  const int gi = blockIdx.z % lv->ngroups, bt = blockIdx.z / lv->ngroups;
  const TbKsGroup& G = lv->g[gi];
  const TbPrime& P = c.pr[level + t];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const i64* le = lenter + G.lenter_off + (level + t);
  longlong2 d = *reinterpret_cast<const longlong2*>(st.row(bt, G.state_row0) + j);
  i64 x0 = tb_mm_ss(d.x, P.Rs, P.q4, P.k);
  i64 x1 = tb_mm_ss(d.y, P.Rs, P.q4, P.k);
  for (int kk = 1; kk < G.alpha; ++kk) {
    d = *reinterpret_cast<const longlong2*>(st.row(bt, G.state_row0 + kk) + j);
    const i64 s = le[(long)(kk - 1) * c.P];
    x0 = tb_add(x0, tb_mm_ss(d.x, s, P.q4, P.k), P.q2);
    x1 = tb_add(x1, tb_mm_ss(d.y, s, P.q4, P.k), P.q2);
  }
  longlong2 o;
  o.x = x0;
  o.y = x1;
  *reinterpret_cast<longlong2*>(ext + (((long)bt * lv->ngroups + gi) * rowsE + t) * N + j) = o;
}

// key inner product (ckks_engine.py:1392-1397 + 1316-1324): for both key halves,
// acc = fold_g CS2(acc + MM(E_g, key_g)), groups in storage (ascending global id) order.
struct TbKskDev {
  const i64* b[TB_MAXG];
  const i64* a[TB_MAXG];
  long rs;
};
__global__ void __launch_bounds__(256) k_mac(TbDev c, const TbKsLevel* lv, TbKskDev key, const i64* ext, i64* acc,
                                             int level, int N, int rowsE) {
  const int t = blockIdx.y, bt = blockIdx.z;
  const TbPrime& P = c.pr[level + t];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const int ng = lv->ngroups;
  i64 a0x = 0, a0y = 0, a1x = 0, a1y = 0;
  for (int gi = 0; gi < ng; ++gi) {
    const int gid = lv->g[gi].gid;
    const longlong2 e = *reinterpret_cast<const longlong2*>(ext + (((long)bt * ng + gi) * rowsE + t) * N + j);
    const longlong2 kb = *reinterpret_cast<const longlong2*>(key.b[gid] + (long)(level + t) * key.rs + j);
    const longlong2 ka = *reinterpret_cast<const longlong2*>(key.a[gid] + (long)(level + t) * key.rs + j);
    a0x = tb_add(a0x, tb_mm_ss(e.x, kb.x, P.q4, P.k), P.q2);
    a0y = tb_add(a0y, tb_mm_ss(e.y, kb.y, P.q4, P.k), P.q2);
    a1x = tb_add(a1x, tb_mm_ss(e.x, ka.x, P.q4, P.k), P.q2);
    a1y = tb_add(a1y, tb_mm_ss(e.y, ka.y, P.q4, P.k), P.q2);
  }
  longlong2 o;
  o.x = a0x;
  o.y = a0y;
  *reinterpret_cast<longlong2*>(acc + (((long)bt * 2 + 0) * rowsE + t) * N + j) = o;
  o.x = a1x;
  o.y = a1y;
  *reinterpret_cast<longlong2*>(acc + (((long)bt * 2 + 1) * rowsE + t) * N + j) = o;
}

// ModDown step 1 (he_fused_cuda.cu:433-469): make the K special limbs mutually consistent, in place.
// pir_sp[k*K + row] = (P_k^-1 mod P_row) R mod P_row.  p: [K][N] rows.
__global__ void __launch_bounds__(256) k_chain_backward(TbDev c, TbView p, const i64* pir_sp, int K, int sp0, int N) {
  const int bt = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  i64 v[TB_MAXA];
#pragma unroll
  for (int i = 0; i < TB_MAXA; ++i)
    if (i < K) v[i] = p.row(bt, i)[j];
#pragma unroll
  for (int row = TB_MAXA - 2; row >= 0; --row) {
    if (row <= K - 2) {
      const TbPrime& P = c.pr[sp0 + row];
      i64 x = v[row];
#pragma unroll
      for (int kk = TB_MAXA - 1; kk > row; --kk) {
        if (kk <= K - 1) {
          const i64 sdiff = tb_sub(x, v[kk], P.q2);
          x = tb_mm_ss(sdiff, pir_sp[kk * K + row], P.q4, P.k);
        }
      }
      v[row] = x;
    }
  }
#pragma unroll
  for (int i = 0; i < TB_MAXA; ++i)
    if (i < K - 1) p.row(bt, i)[j] = v[i];
}

// ModDown step 2 (he_fused_cuda.cu:471-519) with the callers' tails fused:
//   TAIL 0: out = y                      (create_switcher output)
//   TAIL 1: out = CS1(add + y)           (relinearize :1717-1722: plain add, reduce_2q)
//   TAIL 2: out = CS1(CS2(add + y))      (switch_key :1409: mont_add_reduce_2q)
// where y = CS1(MR(x)).  pir[k*P + g] = (P_k^-1 mod q_g) R mod q_g.
template <int TAIL>
__global__ void __launch_bounds__(256) k_divide_by_p(TbDev c, TbView cc, TbView p, TbView add, TbView out,
                                                     const i64* pir, int K, int prime0, int N) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const int g = prime0 + r;
  const TbPrime& P = c.pr[g];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  i64 x = tb_mm_ss(cc.row(bt, r)[j], P.Rs, P.q4, P.k);
  for (int kk = K - 1; kk >= 0; --kk) {
    const i64 pe = tb_mm_ss(p.row(bt, kk)[j], P.Rs, P.q4, P.k);
    x = tb_sub(x, pe, P.q2);
    x = tb_mm_ss(x, pir[(long)kk * c.P + g], P.q4, P.k);
  }
  x = tb_cs1(tb_mr(x, P.q4, P.k), P.q);
  if constexpr (TAIL == 1) x = tb_cs1(add.row(bt, r)[j] + x, P.q);
  if constexpr (TAIL == 2) x = tb_cs1(tb_add(add.row(bt, r)[j], x, P.q2), P.q);
  out.row(bt, r)[j] = x;
}

// tensor product (ckks_engine.py:1669-1675): d0 = x0 y0, d1 = CS2(x0 y1 + x1 y0), d2 = x1 y1
__global__ void __launch_bounds__(256) k_tensor(TbDev c, TbView x0, TbView x1, TbView y0, TbView y1, TbView d0,
                                                TbView d1, TbView d2, int prime0, int N) {
  const int r = blockIdx.y, bt = blockIdx.z;
  const TbPrime& P = c.pr[prime0 + r];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= N) return;
  const longlong2 a0 = *reinterpret_cast<const longlong2*>(x0.row(bt, r) + j);
  const longlong2 a1 = *reinterpret_cast<const longlong2*>(x1.row(bt, r) + j);
  const longlong2 b0 = *reinterpret_cast<const longlong2*>(y0.row(bt, r) + j);
  const longlong2 b1 = *reinterpret_cast<const longlong2*>(y1.row(bt, r) + j);
  longlong2 o;
  o.x = tb_mm_ss(a0.x, b0.x, P.q4, P.k);
  o.y = tb_mm_ss(a0.y, b0.y, P.q4, P.k);
  *reinterpret_cast<longlong2*>(d0.row(bt, r) + j) = o;
  o.x = tb_add(tb_mm_ss(a0.x, b1.x, P.q4, P.k), tb_mm_ss(a1.x, b0.x, P.q4, P.k), P.q2);
  o.y = tb_add(tb_mm_ss(a0.y, b1.y, P.q4, P.k), tb_mm_ss(a1.y, b0.y, P.q4, P.k), P.q2);
  *reinterpret_cast<longlong2*>(d1.row(bt, r) + j) = o;
  o.x = tb_mm_ss(a1.x, b1.x, P.q4, P.k);
  o.y = tb_mm_ss(a1.y, b1.y, P.q4, P.k);
  *reinterpret_cast<longlong2*>(d2.row(bt, r) + j) = o;
}
