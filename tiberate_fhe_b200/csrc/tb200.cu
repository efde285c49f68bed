// tb200.cu -- context construction, launchers and the C ABI declared in include/tb200.h.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/tb200.h"
#include "tb200_ctx.h"
#include "tb200_csprng.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) return fail((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
// Optional per-kernel timing (bench.py roofline): when enabled every launch is bracketed by CUDA
// events on the launching stream; tb200_prof_collect() later sums the elapsed time per kernel name.
struct ProfRec {
  const char* name;
#ifndef TB200_HOST_EMU
  cudaEvent_t e0, e1;
#endif
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;  // launches may come from several host threads
#ifndef TB200_HOST_EMU
#define PROF_BEGIN(name_, stream)                                 \
  ProfRec pr_;                                                    \
  if (g_prof_on) {                                                \
    pr_.name = name_;                                             \
    cudaEventCreate(&pr_.e0);                                     \
    cudaEventCreate(&pr_.e1);                                     \
    cudaEventRecord(pr_.e0, (cudaStream_t)(stream));              \
  }
#define PROF_END(stream)                             \
  if (g_prof_on) {                                   \
    cudaEventRecord(pr_.e1, (cudaStream_t)(stream)); \
    std::lock_guard<std::mutex> lk_(g_prof_mu);      \
    g_prof.push_back(pr_);                           \
  }
#else
#define PROF_BEGIN(name_, stream)
#define PROF_END(stream)
#endif
#define LAUNCH(kernel, grid, block, stream, ...) LAUNCHN(#kernel, kernel, grid, block, stream, __VA_ARGS__)
#define LAUNCHN(kname, kernel, grid, block, stream, ...)                 \
  do {                                                                   \
    PROF_BEGIN(kname, stream)                                            \
    TB_LAUNCH(kernel, grid, block, (cudaStream_t)(stream), __VA_ARGS__); \
    PROF_END(stream)                                                     \
    g_launches.fetch_add(1, std::memory_order_relaxed);                  \
  } while (0)
#define LAUNCH_DYN(kname, kernel, grid, block, smem, stream, ...)                       \
  do {                                                                                  \
    PROF_BEGIN(kname, stream)                                                           \
    TB_LAUNCH_DYN(kernel, grid, block, smem, (cudaStream_t)(stream), __VA_ARGS__);      \
    PROF_END(stream)                                                                    \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
  } while (0)
#define POST()                                                                     \
  do {                                                                             \
    cudaError_t e_ = cudaPeekAtLastError();                                        \
    if (e_ != cudaSuccess) return fail((int)e_, "launch: %s", cudaGetErrorString(e_)); \
  } while (0)

// cudaSetDevice for the duration of one entry point; the caller's current device is restored on return
// (the reference's launchers leave it changed).
struct DevGuard {
  int prev = -1;
  cudaError_t set(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev == dev) {
      prev = -1;
      return cudaSuccess;
    }
    return cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define SET_DEVICE(dev) \
  DevGuard dev_guard_;  \
  CK(dev_guard_.set(dev))

extern "C" const char* tb200_last_error(void) { return g_err; }
extern "C" const char* tb200_version(void) { return "tb200 0.1 (sm_100a)"; }
extern "C" int64_t tb200_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" void tb200_prof_enable(int on) {
  g_prof_on = on != 0;
}
// Synchronises, then writes "name\tlaunches\ttotal_ms\n" lines into buf; returns bytes written.
extern "C" int tb200_prof_collect(char* buf, int cap) {
  int n = 0;
#ifndef TB200_HOST_EMU
  cudaDeviceSynchronize();
  std::vector<const char*> names;
  std::vector<double> ms;
  std::vector<long> cnt;
  for (auto& r : g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    size_t i = 0;
    for (; i < names.size(); ++i)
      if (strcmp(names[i], r.name) == 0) break;
    if (i == names.size()) {
      names.push_back(r.name);
      ms.push_back(0);
      cnt.push_back(0);
    }
    ms[i] += t;
    cnt[i] += 1;
  }
  for (size_t i = 0; i < names.size() && n < cap - 128; ++i)
    n += snprintf(buf + n, cap - n, "%s\t%ld\t%.6f\n", names[i], cnt[i], ms[i]);
#endif
  g_prof.clear();
  if (cap > 0) buf[n < cap ? n : cap - 1] = 0;
  return n;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
template <class T>
static cudaError_t upload(T** dptr, const std::vector<T>& h) {
  cudaError_t e = cudaMalloc((void**)dptr, h.size() * sizeof(T) + 16);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" void tb200_ctx_destroy(tb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_primes);
  cudaFree(c->d_psi4);
  cudaFree(c->d_ipsi4);
  cudaFree(c->d_rescale);
  cudaFree(c->d_pir);
  cudaFree(c->d_pir_sp);
  cudaFree(c->d_lenter);
  cudaFree(c->d_ks);
  cudaFree(c->d_fp);
  cudaFree(c->d_tw);
  cudaFree(c->d_itw);
  cudaFree(c->d_twd);
  cudaFree(c->d_itwd);
  cudaFree(c->d_resc3);
  cudaFree(c->d_lenter2);
  cudaFree(c->d_lenterd);
  cudaFree(c->d_cP);
  cudaFree(c->d_bn);
  cudaFree(c->ws);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
}

// Ownership of the ordinary primes when a ring is sharded by RNS limb over `world` ranks
// (tiberate/context/rns_partition.py:34-52): scale-prime group g (K consecutive primes) belongs to
// rank (np-1-g) mod world, the base prime to rank 0; the special primes are replicated.
static int tb_group_owner(int gid, int np, int world) { return gid < np ? (np - 1 - gid) % world : 0; }

static tb200_ctx* ctx_create_impl(int device, int logN, int Pg, int num_special, const int64_t* qg, int scale_bits,
                                  int rank, int world) {
  if (logN < 8 || logN > 17 || num_special < 1 || num_special > TB_MAXA || Pg < num_special + 1 || !qg) {
    fail(TB200_EINVAL, "ctx_create: need 8 <= logN <= 17, 1 <= K <= %d, P >= K+1", TB_MAXA);
    return nullptr;
  }
  if (world < 1 || rank < 0 || rank >= world) {
    fail(TB200_EINVAL, "ctx_create: bad rank %d / world %d", rank, world);
    return nullptr;
  }
  const int N = 1 << logN, K = num_special, no_g = Pg - K, ns = no_g - 1;
  const int np = (ns + K - 1) / K;  // scale-prime groups; + 1 base group
  if (np + 1 > TB_MAXG) {
    fail(TB200_EINVAL, "ctx_create: %d digit groups exceed TB200_MAX_GROUPS", np + 1);
    return nullptr;
  }
  for (int g = 0; g < Pg; ++g) {
    const u64 qi = (u64)qg[g];
    if (qg[g] <= 2 || (qi - 1) % (2ull * N) != 0 || (qi >> 60) > 0) {
      fail(TB200_EINVAL, "ctx_create: prime %d (%lld) must be = 1 mod 2N and < 2^60", g, (long long)qg[g]);
      return nullptr;
    }
  }
  // local prime list: owned ordinary primes (ascending global id) followed by the special primes
  auto group_of = [&](int p) { return p < ns ? p / K : np; };
  std::vector<int> ord_gid;
  for (int p = 0; p < no_g; ++p)
    if (tb_group_owner(group_of(p), np, world) == rank) ord_gid.push_back(p);
  std::vector<int64_t> qloc;
  for (int p : ord_gid) qloc.push_back(qg[p]);
  for (int kk = 0; kk < K; ++kk) qloc.push_back(qg[no_g + kk]);
  const int64_t* q = qloc.data();
  const int no = (int)ord_gid.size(), P = no + K;
  if (cudaSetDevice(device) != cudaSuccess) {
    fail(TB200_ENODEV, "ctx_create: cudaSetDevice(%d) failed", device);
    return nullptr;
  }
#ifndef TB200_HOST_EMU
  {  // keep freed workspace blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
#endif
  tb200_ctx* c = new tb200_ctx();
  c->device = device;
  c->logN = logN;
  c->N = N;
  c->P = P;
  c->K = K;
  c->num_ord = no;
  c->num_levels = no_g;
  c->rank = rank;
  c->world = world;
  c->ord_gid = ord_gid;
  c->qg.assign(qg, qg + Pg);
  c->lstart.resize(no_g + 1);
  for (int l = 0; l <= no_g; ++l) {
    int n = 0;
    for (int p : ord_gid) n += p < l ? 1 : 0;
    c->lstart[l] = n;
  }
  c->scale_bits = scale_bits;
  c->LB = logN >= 12 ? 8 : logN / 2;
  c->LA = logN - c->LB;
  c->q.assign(q, q + P);
  c->k.resize(P);
  c->primes.resize(P);
  const u128 Rbig = (u128)1 << 62;
  for (int g = 0; g < P; ++g) {
    const u64 qi = (u64)q[g];
    TbPrime& p = c->primes[g];
    c->k[g] = h_neg_inv_pow2(qi);
    const u64 Rm = (u64)(Rbig % qi);
    p.q = (i64)qi;
    p.q2 = (i64)(2 * qi);
    p.q4 = 4 * qi;
    p.k = c->k[g];
    p.Rs = (i64)h_mulmod(Rm, Rm, qi);
    p.Rs_scale = (i64)h_mulmod((u64)p.Rs, h_powmod(2, (u64)scale_bits, qi), qi);
    p.Ninv = (i64)h_mulmod(h_invmod_prime((u64)N, qi), Rm, qi);
    p.pad = 0;
  }
  // twiddles (ntt_context.py:21-85 + psi_enter :277-281)
  c->psi.resize((size_t)P * N);
  c->ipsi.resize((size_t)P * N);
  std::vector<u64> psi4((size_t)P * N), ipsi4((size_t)P * N), series(N);
  std::vector<TbTw2> tw2((size_t)P * N), itw2((size_t)P * N);
  std::vector<TbFastPrime> fps(P);
  for (int g = 0; g < P; ++g) {
    const u64 qi = (u64)q[g];
    const u64 e = (qi - 1) / (2ull * N);
    u64 w = 0;
    for (u64 x = 2; x < (u64)N; ++x) {
      w = h_powmod(x, e, qi);
      if (h_powmod(w, (u64)N, qi) != 1) break;
    }
    for (int dir = 0; dir < 2; ++dir) {
      const u64 base = dir == 0 ? w : h_invmod_prime(w, qi);
      u64 acc = 1;
      for (int i = 0; i < N; ++i) {
        series[i] = acc;
        acc = h_mulmod(acc, base, qi);
      }
      std::vector<u64>& tab = dir == 0 ? c->psi : c->ipsi;
      std::vector<u64>& tab4 = dir == 0 ? psi4 : ipsi4;
      std::vector<TbTw2>& t2 = dir == 0 ? tw2 : itw2;
      for (int i = 0; i < N; ++i) {
        const u64 plain = series[h_bitrev(i, logN)];
        const i64 m = h_mm((i64)plain, c->primes[g].Rs, (i64)qi, c->k[g]);
        tab[(size_t)g * N + i] = (u64)m;
        tab4[(size_t)g * N + i] = (u64)m << 2;
        t2[(size_t)g * N + i].w = plain;
        t2[(size_t)g * N + i].ws = h_shoup(plain, qi);
      }
    }
    // pass B's last round: store the per-thread-consecutive twiddle runs transposed inside each tile
    // (see fwd_round<PERM>): entry (t, gg) of stage d moves from t * G + gg to gg * T + t
    {
      const int LBv = c->LB, NRb = (LBv + 3) / 4, ns_last = LBv - 4 * (NRb - 1), T = 1 << (LBv - 4);
      std::vector<TbTw2> tmp;
      std::vector<u64> tmp4;
      for (int dir = 0; dir < 2; ++dir) {
        std::vector<TbTw2>& t2 = dir == 0 ? tw2 : itw2;
        std::vector<u64>& t4 = dir == 0 ? psi4 : ipsi4;  // the exact-path device tables use the same layout
        for (int d = 0; d < ns_last && d < 3; ++d) {
          const int m = 1 << (logN - 1 - d), per_tile = 1 << (LBv - 1 - d), G = 8 >> d;
          tmp.assign(t2.begin() + (size_t)g * N + m, t2.begin() + (size_t)g * N + 2 * m);
          tmp4.assign(t4.begin() + (size_t)g * N + m, t4.begin() + (size_t)g * N + 2 * m);
          for (int tile0 = 0; tile0 < m; tile0 += per_tile)
            for (int t = 0; t < T; ++t)
              for (int gg = 0; gg < G; ++gg) {
                t2[(size_t)g * N + m + tile0 + gg * T + t] = tmp[tile0 + t * G + gg];
                t4[(size_t)g * N + m + tile0 + gg * T + t] = tmp4[tile0 + t * G + gg];
              }
        }
      }
    }
    {
      TbFastPrime& f = fps[g];
      const u64 Rm = (u64)(Rbig % qi);
      f.q = qi;
      f.q2 = 2 * qi;
      f.Rm = Rm;
      f.Rm_s = h_shoup(Rm, qi);
      f.ex = h_mulmod(h_invmod_prime((u64)N, qi), h_invmod_prime(Rm, qi), qi);
      f.ex_s = h_shoup(f.ex, qi);
      f.off = qi << (62 - h_bitlen(qi));
      f.small = (qi >> 42) == 0 ? 1 : 0;
      f.f64 = f.small;  // default: every 40-bit limb on the FP64 pipe (tb200_ctx_set_f64_share)
      f.qd = (double)qi;
      f.qinv = 1.0 / (double)qi;
      f.exd = f.ex > qi / 2 ? -(double)(qi - f.ex) : (double)f.ex;
      f.Rcd = Rm > qi / 2 ? -(double)(qi - Rm) : (double)Rm;
      f.c96 = (u64)(((unsigned __int128)1 << 96) % qi);
      {
        const u64 ri = h_invmod_prime(Rm, qi);
        f.Rid = ri > qi / 2 ? -(double)(qi - ri) : (double)ri;
      }
      {
        u64 pp = 1 % qi;
        for (int kk = 0; kk < K; ++kk) pp = h_mulmod(pp, (u64)q[no + kk] % qi, qi);
        f.cPd = pp > qi / 2 ? -(double)(qi - pp) : (double)pp;
      }
      {
        const u64 ni = h_invmod_prime((u64)N, qi);
        f.nid = ni > qi / 2 ? -(double)(qi - ni) : (double)ni;
      }
    }
  }
  // rescale scales and P_k^-1 tables
  std::vector<i64> resc((size_t)no_g * P, 0), pir((size_t)K * P, 0), pirsp((size_t)K * K, 0);
  for (int l = 0; l < no_g; ++l)
    for (int g = 0; g < no; ++g) {
      if (ord_gid[g] <= l) continue;
      const u64 qq = (u64)q[g], Rm = (u64)(Rbig % qq);
      resc[(size_t)l * P + g] = (i64)h_mulmod(h_invmod_prime((u64)qg[l] % qq, qq), Rm, qq);
    }
  std::vector<u64> resc3((size_t)no_g * P * 3, 0);
  for (int l = 0; l < no_g; ++l)
    for (int g = 0; g < no; ++g) {
      if (ord_gid[g] <= l) continue;
      const u64 qq = (u64)q[g], ql = (u64)qg[l], Rm = (u64)(Rbig % qq);
      const u64 c1 = h_mulmod(h_invmod_prime(ql % qq, qq), Rm, qq);
      u64* r3 = &resc3[((size_t)l * P + g) * 3];
      r3[0] = c1;
      r3[1] = h_shoup(c1, qq);
      r3[2] = qq * ((ql + qq - 1) / qq + 1);  // multiple of q_g above q_l (+ q_g of slack for tiny negatives)
    }
  for (int kk = 0; kk < K; ++kk)
    for (int g = 0; g < no + kk; ++g) {
      const u64 qq = (u64)q[g], Rm = (u64)(Rbig % qq);
      pir[(size_t)kk * P + g] = (i64)h_mulmod(h_invmod_prime((u64)q[no + kk] % qq, qq), Rm, qq);
    }
  for (int kk = 0; kk < K; ++kk)
    for (int row = 0; row < kk; ++row) pirsp[(size_t)kk * K + row] = pir[(size_t)kk * P + no + row];
  std::vector<i64> cP((size_t)2 * P, 0);
  for (int g = 0; g < P; ++g) {
    const u64 qq = (u64)q[g];
    u64 pp = 1 % qq;
    for (int kk = 0; kk < K; ++kk) pp = h_mulmod(pp, (u64)q[no + kk] % qq, qq);
    cP[g] = (i64)pp;
    cP[(size_t)P + g] = (i64)h_mulmod(pp, (u64)(Rbig % qq), qq);
  }
  // ModDown in product form: B_k = prod_{j<=k} P_j^-1 mod q_g
  // rows [0, K): (-B_k, Shoup); row K: (B_{K-1}, Shoup); rows K+1+k: (P_k^-1, Shoup) for the composed form
  std::vector<u64> bn((size_t)(2 * K + 1) * P * 2, 0);
  for (int g = 0; g < no; ++g) {
    const u64 qq = (u64)q[g];
    u64 B = 1;
    for (int kk = 0; kk < K; ++kk) {
      const u64 pinv = h_invmod_prime((u64)q[no + kk] % qq, qq);
      bn[((size_t)(K + 1 + kk) * P + g) * 2] = pinv;
      bn[((size_t)(K + 1 + kk) * P + g) * 2 + 1] = h_shoup(pinv, qq);
      B = h_mulmod(B, pinv, qq);
      const u64 neg = (qq - B) % qq;
      bn[((size_t)kk * P + g) * 2] = neg;
      bn[((size_t)kk * P + g) * 2 + 1] = h_shoup(neg, qq);
    }
    bn[((size_t)K * P + g) * 2] = B;
    bn[((size_t)K * P + g) * 2 + 1] = h_shoup(B, qq);
  }
  // digit groups per level
  std::vector<i64> lenter;
  std::vector<u64> lenter2;
  std::vector<double> lenterd;  // L_{k-1} mod q_g centred, as doubles (FP64 extend: no Montgomery factor)
  c->ks.resize(no_g);
  for (int l = 0; l < no_g; ++l) {
    TbKsLevel& lv = c->ks[l];
    memset(&lv, 0, sizeof(lv));
    lv.L = no - c->lstart[l];
    // rows of the digit-state buffer: owner-major, every rank's segment padded to the largest one, so
    // that ONE in-place all-gather completes it (for world == 1: the global row order)
    std::vector<int> owned_rows(world, 0);
    for (int gid = 0; gid <= np; ++gid) {
      const int lo = gid < np ? gid * K : ns;
      const int hi = gid < np ? ((gid + 1) * K < ns ? (gid + 1) * K : ns) : ns + 1;
      for (int p = lo; p < hi; ++p)
        if (p >= l) owned_rows[tb_group_owner(gid, np, world)]++;
    }
    int seg = 0;
    for (int r = 0; r < world; ++r) seg = owned_rows[r] > seg ? owned_rows[r] : seg;
    lv.seg_rows = seg;
    lv.state_rows = seg * world;
    std::vector<int> fill(world, 0);
    for (int gid = 0; gid <= np; ++gid) {
      const int lo = gid < np ? gid * K : ns;
      const int hi = gid < np ? ((gid + 1) * K < ns ? (gid + 1) * K : ns) : ns + 1;
      std::vector<int> alive;
      for (int p = lo; p < hi; ++p)
        if (p >= l) alive.push_back(p);
      if (alive.empty()) continue;
      const int owner = tb_group_owner(gid, np, world);
      TbKsGroup& G = lv.g[lv.ngroups];
      const int alpha = (int)alive.size();
      G.alpha = alpha;
      G.gid = gid;
      G.state_row0 = owner * seg + fill[owner];
      fill[owner] += alpha;
      G.src_prime0 = -1;
      G.src_row0 = -1;
      G.wide_mask = 0;
      for (int k = 0; k < alpha; ++k)
        if (((u64)qg[alive[k]] >> 42) != 0) G.wide_mask |= 1 << k;
      if (owner == rank) {
        int li = 0;
        while (ord_gid[li] != alive[0]) ++li;
        G.src_prime0 = li;
        G.src_row0 = li - c->lstart[l];
        lv.own[lv.nown++] = lv.ngroups;
      } else {
        lv.foreign[lv.nforeign++] = lv.ngroups;
      }
      lv.ngroups++;
      G.lenter_off = (long)lenter.size();
      // L_i = m_0 ... m_i reduced modulo whatever prime is needed
      auto Lmod = [&](int i, u64 m) {
        u64 r = 1 % m;
        for (int t = 0; t <= i; ++t) r = h_mulmod(r, (u64)qg[alive[t]] % m, m);
        return r;
      };
      for (int i = 0; i + 1 < alpha; ++i) {
        const u64 m1 = (u64)qg[alive[i + 1]];
        G.Y[i] = (i64)h_mulmod(h_invmod_prime(Lmod(i, m1), m1), (u64)(Rbig % m1), m1);
        for (int j = i + 2; j < alpha; ++j) {
          const u64 mj = (u64)qg[alive[j]];
          G.Lsc[i][j] = (i64)h_mulmod(Lmod(i, mj), (u64)(Rbig % mj), mj);
        }
      }
      for (int i = 0; i + 1 < alpha; ++i)
        for (int g = 0; g < P; ++g) {
          const u64 qq = (u64)q[g];
          lenter.push_back((i64)h_mulmod(Lmod(i, qq), (u64)c->primes[g].Rs, qq));
          const u64 C = h_mulmod(Lmod(i, qq), (u64)(Rbig % qq), qq);
          lenter2.push_back(C);
          lenter2.push_back(h_shoup(C, qq));
          const u64 Lq = Lmod(i, qq);
          lenterd.push_back(Lq > qq / 2 ? -(double)(qq - Lq) : (double)Lq);
        }
    }
  }
  // FP64 butterfly policy: the same tables as doubles centred into (-q/2, q/2]
  std::vector<double> twd((size_t)P * N), itwd((size_t)P * N);
  for (int g = 0; g < P; ++g) {
    const u64 qi = (u64)q[g];
    for (int i = 0; i < N; ++i) {
      const size_t at = (size_t)g * N + i;
      const u64 a = tw2[at].w, b = itw2[at].w;
      twd[at] = a > qi / 2 ? -(double)(qi - a) : (double)a;
      itwd[at] = b > qi / 2 ? -(double)(qi - b) : (double)b;
    }
  }
  if (lenter.empty()) lenter.push_back(0);
  if (lenter2.empty()) lenter2.push_back(0);
  if (lenterd.empty()) lenterd.push_back(0.0);
  bool ok = upload(&c->d_primes, c->primes) == cudaSuccess && upload(&c->d_psi4, psi4) == cudaSuccess &&
            upload(&c->d_ipsi4, ipsi4) == cudaSuccess && upload(&c->d_rescale, resc) == cudaSuccess &&
            upload(&c->d_pir, pir) == cudaSuccess && upload(&c->d_pir_sp, pirsp) == cudaSuccess &&
            upload(&c->d_lenter, lenter) == cudaSuccess && upload(&c->d_ks, c->ks) == cudaSuccess &&
            ((c->fps = fps), upload(&c->d_fp, fps)) == cudaSuccess && upload(&c->d_tw, tw2) == cudaSuccess &&
            upload(&c->d_itw, itw2) == cudaSuccess && upload(&c->d_twd, twd) == cudaSuccess &&
            upload(&c->d_itwd, itwd) == cudaSuccess && upload(&c->d_resc3, resc3) == cudaSuccess &&
            upload(&c->d_lenter2, lenter2) == cudaSuccess && upload(&c->d_lenterd, lenterd) == cudaSuccess && upload(&c->d_cP, cP) == cudaSuccess && upload(&c->d_bn, bn) == cudaSuccess;
  if (!ok) {
    fail(TB200_ENOMEM, "ctx_create: device allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    tb200_ctx_destroy(c);
    return nullptr;
  }
  return c;
}

extern "C" tb200_ctx* tb200_ctx_create(int device, int logN, int num_primes, int num_special, const int64_t* q,
                                       int scale_bits) {
  return ctx_create_impl(device, logN, num_primes, num_special, q, scale_bits, 0, 1);
}
extern "C" tb200_ctx* tb200_ctx_create_sharded(int device, int logN, int num_primes, int num_special,
                                               const int64_t* q, int scale_bits, int rank, int world) {
  return ctx_create_impl(device, logN, num_primes, num_special, q, scale_bits, rank, world);
}

extern "C" int tb200_ctx_get_prime_consts(const tb200_ctx* c, int64_t* out) {
  if (!c || !out) return fail(TB200_EINVAL, "null");
  memcpy(out, c->primes.data(), sizeof(TbPrime) * c->P);
  return 0;
}
extern "C" int tb200_ctx_get_twiddles(const tb200_ctx* c, int inverse, int prime, int64_t* out) {
  if (!c || !out || prime < 0 || prime >= c->P) return fail(TB200_EINVAL, "bad prime index");
  const std::vector<u64>& t = inverse ? c->ipsi : c->psi;
  memcpy(out, t.data() + (size_t)prime * c->N, sizeof(u64) * c->N);
  return 0;
}
extern "C" int tb200_ctx_info(const tb200_ctx* c, int32_t* out) {
  if (!c || !out) return fail(TB200_EINVAL, "null");
  out[0] = c->logN;
  out[1] = c->N;
  out[2] = c->P;
  out[3] = c->K;
  out[4] = c->LA;
  out[5] = c->LB;
  out[6] = c->device;
  out[7] = c->ks[0].ngroups;
  return 0;
}
extern "C" int tb200_ctx_set_chunk(tb200_ctx* c, int chunk) {
  if (!c || chunk < 1) return fail(TB200_EINVAL, "chunk must be >= 1");
  c->chunk = chunk;
  return 0;
}

extern "C" int tb200_ctx_set_fast(tb200_ctx* c, int on) {
  if (!c) return fail(TB200_EINVAL, "null context");
  c->fast = on != 0;
  return 0;
}

// Share of the small-prime limbs (in eighths, 0..8) whose butterflies run on the FP64 pipe; the others
// stay on the integer pipes, so co-resident CTAs of both kinds keep both pipes busy.
extern "C" int tb200_ctx_set_f64_share(tb200_ctx* c, int eighths) {
  if (!c || eighths < 0 || eighths > 8) return fail(TB200_EINVAL, "f64 share must be 0..8 (eighths)");
  SET_DEVICE(c->device);
  c->f64_eighths = eighths;
  for (int g = 0; g < c->P; ++g) c->fps[g].f64 = (c->fps[g].small && ((g * 5) & 7) < eighths) ? 1 : 0;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(c->d_fp, c->fps.data(), sizeof(TbFastPrime) * c->P, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int tb200_ctx_set_tuning(tb200_ctx* c, int knob, int value) {
  if (!c) return fail(TB200_EINVAL, "null context");
  switch (knob) {
    case TB200_TUNE_FUSED_CORE:
      c->fused_core = value != 0;
      return 0;
    case TB200_TUNE_SIDE_ROWS:
      c->side_rows = value;  // bit 0: key-switch rows, bit 1: the 60-bit rows of the input transforms of cc_mult
      return 0;
    case TB200_TUNE_FUSED_MODDOWN:
      c->fused_moddown = value != 0;
      return 0;
    case TB200_TUNE_STREAM_WS:
      c->stream_ws = value != 0;
      return 0;
    case TB200_TUNE_SUM_NTT:
      c->sum_ntt = value != 0;
      return 0;
    case TB200_TUNE_FUSED_TENSOR:
      c->fused_tensor = value != 0;
      return 0;
  }
  return fail(TB200_EINVAL, "unknown tuning knob %d", knob);
}

static int ws_reserve(tb200_ctx* c, size_t elems) {
  if (elems <= c->ws_elems) return 0;
  // grow-only; the previous buffer may still be in use by queued kernels of the caller's stream
  cudaDeviceSynchronize();
  cudaFree(c->ws);
  c->ws = nullptr;
  c->ws_elems = 0;
  if (cudaMalloc((void**)&c->ws, elems * sizeof(i64)) != cudaSuccess) {
    cudaGetLastError();
    return fail(TB200_ENOMEM, "workspace of %zu MiB could not be allocated", elems * 8 >> 20);
  }
  c->ws_elems = elems;
  return 0;
}

// Scratch of one engine call.  Default: leased from the device's stream-ordered memory pool on the CALLER'S
// stream (cudaMallocAsync at entry, cudaFreeAsync at exit): calls issued on different streams or from
// different host threads get different blocks, nothing synchronises the device, and the whole call is a
// sequence of stream operations that a CUDA graph can capture.  TB200_TUNE_STREAM_WS = 0 selects the
// grow-only workspace owned by the context (one stream at a time; growth synchronises the device).
struct WsLease {
  tb200_ctx* c;
  cudaStream_t st;
  i64* p = nullptr;
  bool leased = false;
  WsLease(tb200_ctx* c_, tb200_stream st_) : c(c_), st((cudaStream_t)st_) {}
  WsLease(const WsLease&) = delete;
  WsLease& operator=(const WsLease&) = delete;
  int reserve(size_t elems) {
    if (!c->stream_ws) {
      int rc = ws_reserve(c, elems);
      p = c->ws;
      return rc;
    }
    if (cudaMallocAsync((void**)&p, (elems ? elems : 1) * sizeof(i64), st) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(TB200_ENOMEM, "workspace of %zu MiB could not be allocated", elems * 8 >> 20);
    }
    leased = true;
    return 0;
  }
  ~WsLease() {
    if (leased && p) cudaFreeAsync(p, st);
  }
};


// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static inline TbView view(const tb200_poly* p) {
  TbView v;
  v.p = p ? p->ptr : nullptr;
  v.bs = p ? (long)p->batch_stride : 0;
  v.rs = p ? (long)p->row_stride : 0;
  return v;
}
static inline TbView dense(i64* p, long rows, long N) {
  TbView v;
  v.p = p;
  v.bs = rows * N;
  v.rs = N;
  return v;
}
static int check_poly(const tb200_ctx* c, const tb200_poly* p, const char* name, bool allow_broadcast = false) {
  if (!p || !p->ptr) return fail(TB200_EINVAL, "%s: null polynomial", name);
  if (((uintptr_t)p->ptr & 15) != 0) return fail(TB200_EINVAL, "%s: pointer must be 16-byte aligned", name);
  if (!(allow_broadcast && p->row_stride == 0) && (p->row_stride < c->N || (p->row_stride & 1)))
    return fail(TB200_EINVAL, "%s: row_stride %lld must be even and >= N", name, (long long)p->row_stride);
  if (p->batch_stride & 1) return fail(TB200_EINVAL, "%s: batch_stride must be even", name);
  return 0;
}
#define CHECK_POLY(p) \
  do {                \
    int rc_ = check_poly(c, p, #p); \
    if (rc_) return rc_;            \
  } while (0)
#define CHECK_ROWS(prime0, rows)                                                                   \
  if ((rows) < 1 || (prime0) < 0 || (prime0) + (rows) > c->P)                                       \
  return fail(TB200_EINVAL, "rows %d at prime0 %d outside the %d primes of the context", rows, prime0, c->P)

static inline dim3 grid_pw(const tb200_ctx* c, int rows, int batch, int per_thread) {
  return dim3((unsigned)((c->N / per_thread + 255) / 256), (unsigned)rows, (unsigned)batch);
}

// ------------------------------------------------------------------------------------------------
// op layer
// ------------------------------------------------------------------------------------------------
template <int OP>
static void launch_pw(const tb200_ctx* c, const TbPwArgs& a, int rows, int batch, tb200_stream st) {
  auto kfn = k_pointwise<OP>;
  LAUNCHN(OP == 0 ? "k_pointwise<mult>" : (OP == 15 ? "k_pointwise<copy>" : "k_pointwise<other>"), kfn,
          grid_pw(c, rows, batch, 2), dim3(c->N / 2 < 256 ? c->N / 2 : 256), st, c->dev(), a);
}

extern "C" int tb200_pointwise(tb200_ctx* c, int op, int rows, int batch, int prime0, const tb200_poly* a,
                               const tb200_poly* b, const int64_t* scal, const tb200_explicit_consts* ec,
                               const tb200_poly* out, tb200_stream st) {
  if (!c) return fail(TB200_EINVAL, "null context");
  if (op < 0 || op > 14) return fail(TB200_EINVAL, "unknown pointwise op %d", op);
  if (batch < 1) return fail(TB200_EINVAL, "batch must be >= 1");
  const bool has_b = op <= 4 || op == 13;
  const bool needs_scal = op == 5 || op == 14;
  const bool explicit_c = ec && (ec->ql || ec->two_q);
  // prime0 in [P - 64, 0): the leading rows address the zero padding of the reference's 64-slot pool
  if (!explicit_c && !(prime0 < 0 && prime0 >= c->P - 64 && rows >= 1 && prime0 + rows <= c->P))
    CHECK_ROWS(prime0, rows);
  if (rows < 1) return fail(TB200_EINVAL, "rows must be >= 1");
  {
    int rc = check_poly(c, a, "a", op == TB200_TILE_UNSIGNED);
    if (rc) return rc;
  }
  CHECK_POLY(out);
  if (has_b) CHECK_POLY(b);
  if (needs_scal && !scal) return fail(TB200_EINVAL, "op %d needs a per-row scalar tensor", op);
  if (explicit_c && !ec->two_q && !ec->qh) return fail(TB200_EINVAL, "explicit constants need ql and qh");
  const bool needs_k = op == 0 || (op >= 5 && op <= 8) || op >= 13;
  if (explicit_c && needs_k && !(ec->kl && ec->kh) ) return fail(TB200_EINVAL, "op %d needs kl/kh", op);
  if (explicit_c && (op == 6 || op == 7 || op == 13))
    return fail(TB200_EINVAL, "op %d takes its constants from the context only", op);
  SET_DEVICE(c->device);
  TbPwArgs g;
  g.a = view(a);
  g.b = has_b ? view(b) : view(a);
  g.out = view(out);
  g.scal = scal;
  g.ql = explicit_c ? ec->ql : nullptr;
  g.qh = explicit_c ? ec->qh : nullptr;
  g.kl = explicit_c ? ec->kl : nullptr;
  g.kh = explicit_c ? ec->kh : nullptr;
  g.two_q = explicit_c ? ec->two_q : nullptr;
  g.prime0 = prime0;
  g.N = c->N;
  switch (op) {
#define PWCASE(n) \
  case n:         \
    launch_pw<n>(c, g, rows, batch, st); \
    break;
    PWCASE(0) PWCASE(1) PWCASE(2) PWCASE(3) PWCASE(4) PWCASE(5) PWCASE(6) PWCASE(7) PWCASE(8) PWCASE(9) PWCASE(10)
    PWCASE(11) PWCASE(12) PWCASE(13) PWCASE(14)
#undef PWCASE
  }
  POST();
  return 0;
}

extern "C" int tb200_add_many(tb200_ctx* c, int pairwise, int K, int rows, int prime0, const int64_t* in,
                              int64_t* out, tb200_stream st) {
  if (!c || !in || !out || K < 1) return fail(TB200_EINVAL, "add_many: bad arguments");
  CHECK_ROWS(prime0, rows);
  SET_DEVICE(c->device);
  LAUNCH(k_add_many, grid_pw(c, rows, 1, 1), dim3(256), st, c->dev(), (const i64*)in, (i64*)out, K, rows, c->N,
         prime0, pairwise);
  POST();
  return 0;
}

// ---- packed wire format (scale-prime limbs as 41 bits per residue) -------------------------------------------
static int check_pack(const tb200_ctx* c, int rows, int batch, int prime0, const void* bytes, int64_t bs, int64_t rs) {
  if (!c) return fail(TB200_EINVAL, "null context");
  if (rows < 1 || prime0 < 0 || prime0 + rows > c->P) return fail(TB200_EINVAL, "pack41: rows outside the context's primes");
  if (batch < 1 || !bytes) return fail(TB200_EINVAL, "pack41: bad batch / null buffer");
  if (c->N < 64) return fail(TB200_EINVAL, "pack41: N must be >= 64");
  if (((uintptr_t)bytes & 7) || (bs & 7) || (rs & 7) || rs < (int64_t)c->N * 5 + c->N / 8)
    return fail(TB200_EINVAL, "pack41: packed buffer needs 8-byte alignment and a row pitch >= 5 N + N / 8 (multiple of 8)");
  for (int r = 0; r < rows; ++r)
    if (((u64)c->q[prime0 + r] >> 41) != 0)
      return fail(TB200_EINVAL, "pack41: prime %d has more than 41 bits; such limbs travel as int64", prime0 + r);
  return 0;
}
extern "C" int tb200_unpack41(tb200_ctx* c, int rows, int batch, int prime0, const uint8_t* src, int64_t src_batch_stride,
                              int64_t src_row_stride, const tb200_poly* dst, tb200_stream st) {
  int rc = check_pack(c, rows, batch, prime0, src, src_batch_stride, src_row_stride);
  if (rc) return rc;
  CHECK_POLY(dst);
  SET_DEVICE(c->device);
  LAUNCH(k_unpack41, dim3((unsigned)((c->N / 8 + 255) / 256), (unsigned)rows, (unsigned)batch),
         dim3(c->N / 8 < 256 ? c->N / 8 : 256), st, (const unsigned char*)src, (long)src_batch_stride,
         (long)src_row_stride, view(dst), c->N);
  POST();
  return 0;
}
extern "C" int tb200_pack41(tb200_ctx* c, int rows, int batch, int prime0, const tb200_poly* src, uint8_t* dst,
                            int64_t dst_batch_stride, int64_t dst_row_stride, tb200_stream st) {
  int rc = check_pack(c, rows, batch, prime0, dst, dst_batch_stride, dst_row_stride);
  if (rc) return rc;
  CHECK_POLY(src);
  SET_DEVICE(c->device);
  LAUNCH(k_pack41, dim3((unsigned)((c->N / 8 + 255) / 256), (unsigned)rows, (unsigned)batch),
         dim3(c->N / 8 < 256 ? c->N / 8 : 256), st, view(src), (unsigned char*)dst, (long)dst_batch_stride,
         (long)dst_row_stride, c->N);
  POST();
  return 0;
}

// ---- NTT ----------------------------------------------------------------------------------------
static inline int ntt_lw(const tb200_ctx* c) {  // log2 of the column width of a pass-A tile
  int lw = 12 - c->LA;
  if (lw > c->LB) lw = c->LB;
  return lw;
}

// flags != nullptr: the deferred-reduction kernel first (it flags the tiles it leaves), then the generic kernel on
// the flagged tiles only
template <int PRO>
static int launch_fwd_A(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0, tb200_stream st,
                        unsigned char* flags = nullptr) {
  const int lw = ntt_lw(c);
  const dim3 grid((unsigned)(1 << (c->LB - lw)), (unsigned)rows, (unsigned)batch), block(1u << (c->LA - 4 + lw));
  switch (c->LA) {
#define ACASE(n)                                            \
  case n: {                                                 \
    if (flags) {                                            \
      auto ksum = k_ntt_fwd_A_sum<n, PRO>;                  \
      LAUNCHN("k_ntt_fwd_A_sum", ksum, grid, block, st, c->dev(), src, dst, prime0, lw, flags); \
    }                                                       \
    auto kfn = k_ntt_fwd_A<n, PRO>;                         \
    LAUNCHN("k_ntt_fwd_A", kfn, grid, block, st, c->dev(), src, dst, prime0, lw, (const unsigned char*)flags); \
  } break;
    ACASE(4) ACASE(5) ACASE(6) ACASE(7) ACASE(8) ACASE(9)
#undef ACASE
    default:
      return fail(TB200_EINVAL, "unsupported LA %d", c->LA);
  }
  return 0;
}
template <int EPI>
static int launch_inv_A(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0, tb200_stream st,
                        unsigned char* flags = nullptr) {
  const int lw = ntt_lw(c);
  const dim3 grid((unsigned)(1 << (c->LB - lw)), (unsigned)rows, (unsigned)batch), block(1u << (c->LA - 4 + lw));
  switch (c->LA) {
#define ACASE(n)                                            \
  case n: {                                                 \
    if (flags) {                                            \
      auto ksum = k_ntt_inv_A_sum<n, EPI>;                  \
      LAUNCHN("k_ntt_inv_A_sum", ksum, grid, block, st, c->dev(), src, dst, prime0, lw, flags); \
    }                                                       \
    auto kfn = k_ntt_inv_A<n, EPI>;                         \
    LAUNCHN("k_ntt_inv_A", kfn, grid, block, st, c->dev(), src, dst, prime0, lw, (const unsigned char*)flags); \
  } break;
    ACASE(4) ACASE(5) ACASE(6) ACASE(7) ACASE(8) ACASE(9)
#undef ACASE
    default:
      return fail(TB200_EINVAL, "unsupported LA %d", c->LA);
  }
  return 0;
}
static int launch_B(const tb200_ctx* c, bool inverse, TbView src, TbView dst, int rows, int batch, int prime0,
                    tb200_stream st, unsigned char* flags = nullptr) {
  const int te = c->N < TB_TILE ? c->N : TB_TILE;
  const dim3 grid((unsigned)(c->N / te), (unsigned)rows, (unsigned)batch), block((unsigned)(te / 16));
  switch (c->LB) {
#define BCASE(n)                                                 \
  case n: {                                                      \
    if (inverse) {                                               \
      if (flags) {                                               \
        auto ksum = k_ntt_inv_B_sum<n>;                          \
        LAUNCHN("k_ntt_inv_B_sum", ksum, grid, block, st, c->dev(), src, dst, prime0, flags); \
      }                                                          \
      auto kfn = k_ntt_inv_B<n>;                                 \
      LAUNCHN("k_ntt_inv_B", kfn, grid, block, st, c->dev(), src, dst, prime0, (const unsigned char*)flags);  \
    } else {                                                     \
      if (flags) {                                               \
        auto ksum = k_ntt_fwd_B_sum<n>;                          \
        LAUNCHN("k_ntt_fwd_B_sum", ksum, grid, block, st, c->dev(), src, dst, prime0, flags); \
      }                                                          \
      auto kfn = k_ntt_fwd_B<n>;                                 \
      LAUNCHN("k_ntt_fwd_B", kfn, grid, block, st, c->dev(), src, dst, prime0, (const unsigned char*)flags);  \
    }                                                            \
  } break;
    BCASE(4) BCASE(5) BCASE(6) BCASE(7) BCASE(8)
#undef BCASE
    default:
      return fail(TB200_EINVAL, "unsupported LB %d", c->LB);
  }
  return 0;
}

// number of CTA tiles of one pass (= bytes of one flag array)
static size_t ntt_tiles(const tb200_ctx* c, int rows, int batch) {
  const int te = c->N < TB_TILE ? c->N : TB_TILE;
  return (size_t)(c->N / te) * rows * batch;
}
// flags: 2 * ntt_tiles zeroed bytes (one array per pass) or nullptr (generic kernels only)
static int ntt_forward(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0, bool enter,
                       tb200_stream st, unsigned char* flags = nullptr) {
  int rc = enter ? launch_fwd_A<TB_PRO_ENTER>(c, src, dst, rows, batch, prime0, st, flags)
                 : launch_fwd_A<TB_PRO_NONE>(c, src, dst, rows, batch, prime0, st, flags);
  if (rc) return rc;
  return launch_B(c, false, dst, dst, rows, batch, prime0, st, flags ? flags + ntt_tiles(c, rows, batch) : nullptr);
}
static size_t ntt_tiles(const tb200_ctx* c, int rows, int batch);
// flags: 2 * ntt_tiles zeroed bytes (deferred-reduction kernels first, the generic ones on the flagged tiles) or nullptr
static int ntt_inverse(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0, int mode,
                       tb200_stream st, unsigned char* flags = nullptr) {
  int rc = launch_B(c, true, src, dst, rows, batch, prime0, st, flags);
  if (rc) return rc;
  unsigned char* fa = flags ? flags + ntt_tiles(c, rows, batch) : nullptr;
  switch (mode) {
    case 0:
      return launch_inv_A<TB_EPI_NINV>(c, dst, dst, rows, batch, prime0, st, fa);
    case 1:
      return launch_inv_A<TB_EPI_EXIT>(c, dst, dst, rows, batch, prime0, st, fa);
    case 2:
      return launch_inv_A<TB_EPI_EXIT_REDUCE>(c, dst, dst, rows, batch, prime0, st, fa);
    case 3:
      return launch_inv_A<TB_EPI_EXIT_SIGNED>(c, dst, dst, rows, batch, prime0, st, fa);
  }
  return fail(TB200_EINVAL, "intt mode %d", mode);
}

extern "C" int tb200_ntt(tb200_ctx* c, int rows, int batch, int prime0, const tb200_poly* a, int enter,
                         tb200_stream st) {
  if (!c) return fail(TB200_EINVAL, "null context");
  CHECK_ROWS(prime0, rows);
  CHECK_POLY(a);
  if (batch < 1) return fail(TB200_EINVAL, "batch must be >= 1");
  SET_DEVICE(c->device);
  // 40-bit limbs with lazy inputs take the deferred-reduction kernels (same bits, tb200_fast.cuh: ExactSumPol)
  WsLease ws(c, st);
  unsigned char* flags = nullptr;
  int rc = 0;
  if (c->fast && c->sum_ntt) {
    const size_t nbytes = 2 * ntt_tiles(c, rows, batch);
    if ((rc = ws.reserve((nbytes + 7) / 8))) return rc;
    flags = reinterpret_cast<unsigned char*>(ws.p);
    CK(cudaMemsetAsync(flags, 0, nbytes, (cudaStream_t)st));
  }
  rc = ntt_forward(c, view(a), view(a), rows, batch, prime0, enter != 0, st, flags);
  if (rc) return rc;
  POST();
  return 0;
}
static int fast_inverse_exit(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0,
                             tb200_stream st, int mac_chain = 0, bool wide_in = false);
// inverse pass B' input modes: lazy integers in (-2q, 2q); any |x| < 2^51 on the FP64 limbs (exposed intt);
// doubles on the FP64 limbs (output of the FP64 key inner product)
enum { TB_INV_IN_LAZY = 0, TB_INV_IN_WIDE = 1, TB_INV_IN_DOUBLE = 2 };
extern "C" int tb200_intt(tb200_ctx* c, int rows, int batch, int prime0, const tb200_poly* a, int mode,
                          tb200_stream st) {
  if (!c) return fail(TB200_EINVAL, "null context");
  CHECK_ROWS(prime0, rows);
  CHECK_POLY(a);
  if (batch < 1 || mode < 0 || mode > 3) return fail(TB200_EINVAL, "bad batch/mode");
  SET_DEVICE(c->device);
  // intt_radix2_exit_reduce returns canonical residues, so the mod-q transforms give the same bits
  // (domain: |x| < 2^51 on the FP64 limbs, (-2q, 2q) elsewhere -- every lazy value the ops produce;
  // tb200_ctx_set_fast(ctx, 0) selects the reference's own butterflies for anything wider)
  int rc = 0;
  if (c->fast && mode == 2) {
    rc = fast_inverse_exit(c, view(a), view(a), rows, batch, prime0, st, 0, true);
  } else {
    // the lazy exits (stay in Montgomery form / exit / signed): deferred-reduction kernels on the 40-bit limbs
    WsLease ws(c, st);
    unsigned char* flags = nullptr;
    if (c->fast && c->sum_ntt) {
      const size_t nbytes = 2 * ntt_tiles(c, rows, batch);
      if ((rc = ws.reserve((nbytes + 7) / 8))) return rc;
      flags = reinterpret_cast<unsigned char*>(ws.p);
      CK(cudaMemsetAsync(flags, 0, nbytes, (cudaStream_t)st));
    }
    rc = ntt_inverse(c, view(a), view(a), rows, batch, prime0, mode, st, flags);
  }
  if (rc) return rc;
  POST();
  return 0;
}

// ---- fast (mod-q) transforms ------------------------------------------------------------------------
template <int PRO, bool F64ONLY>
static int launch_fast_fwd_A_rows(const tb200_ctx* c, TbFwdAArgs a, int rows, int gridz, tb200_stream st) {
  const int lw = ntt_lw(c);
  a.LW = lw;
  const dim3 grid((unsigned)(1 << (c->LB - lw)), (unsigned)rows, (unsigned)gridz), block(1u << (c->LA - 4 + lw));
  const bool big = c->LB == 8 && lw == 12 - c->LA;  // logN >= 12: compile-time strides
  if constexpr (F64ONLY) {
    switch (c->LA) {
#define FCASE(n)                                                 \
  case n: {                                                      \
    auto kfn = k_fast_fwd_A<n, PRO, true, true>;                 \
    LAUNCHN("k_fast_fwd_A", kfn, grid, block, st, c->devf(), a); \
  } break;
      FCASE(4) FCASE(5) FCASE(6) FCASE(7) FCASE(8) FCASE(9)
#undef FCASE
      default:
        return fail(TB200_EINVAL, "unsupported LA %d (LB %d)", c->LA, c->LB);
    }
    return 0;
  }
  switch (c->LA + (big ? 100 : 0)) {
#define ACASE(n, B)                                                     \
  case n + (B ? 100 : 0): {                                             \
    auto kfn = k_fast_fwd_A<n, PRO, B>;                                 \
    LAUNCHN("k_fast_fwd_A", kfn, grid, block, st, c->devf(), a);        \
  } break;
    ACASE(4, false) ACASE(5, false) ACASE(6, false)
    ACASE(4, true) ACASE(5, true) ACASE(6, true) ACASE(7, true) ACASE(8, true) ACASE(9, true)
#undef ACASE
    default:
      return fail(TB200_EINVAL, "unsupported LA %d (LB %d)", c->LA, c->LB);
  }
  return 0;
}
static int f64_prefix(const tb200_ctx* c, int prime0, int rows);
static TbView rows_from(TbView v, int r);
// The ModUp extend launch is split like pass B: the FP64 rows run an instantiation without the integer
// prologue and butterflies (64 registers, 4 CTAs/SM), the remaining rows the generic kernel.
template <int PRO>
static int launch_fast_fwd_A(const tb200_ctx* c, TbFwdAArgs a, int rows, int gridz, tb200_stream st) {
  if constexpr (PRO == TB_FPRO_EXTEND) {
    const int lw = ntt_lw(c);
    const int nf = (c->LB == 8 && lw == 12 - c->LA) ? f64_prefix(c, a.prime0, rows) : 0;
    if (nf > 0) {
      int rc = launch_fast_fwd_A_rows<PRO, true>(c, a, nf, gridz, st);
      if (rc || nf == rows) return rc;
      a.prime0 += nf;
      a.dst = rows_from(a.dst, nf);
      return launch_fast_fwd_A_rows<PRO, false>(c, a, rows - nf, gridz, st);
    }
  }
  return launch_fast_fwd_A_rows<PRO, false>(c, a, rows, gridz, st);
}
static int launch_fast_inv_A(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0,
                             int mac_chain, tb200_stream st, const TbMdArgs* md = nullptr) {
  const int lw = ntt_lw(c);
  const dim3 grid((unsigned)(1 << (c->LB - lw)), (unsigned)rows, (unsigned)batch), block(1u << (c->LA - 4 + lw));
  const bool big = c->LB == 8 && lw == 12 - c->LA;
  TbMdArgs m0;
  memset(&m0, 0, sizeof(m0));
  switch (c->LA + (big ? 100 : 0)) {
#define ACASE(n, B)                                                                                       \
  case n + (B ? 100 : 0): {                                                                               \
    if (md) {                                                                                             \
      auto kfn = k_fast_inv_A<n, B, true>;                                                                \
      LAUNCHN("k_fast_inv_A_moddown", kfn, grid, block, st, c->devf(), src, dst, prime0, lw, mac_chain, *md); \
    } else {                                                                                              \
      auto kfn = k_fast_inv_A<n, B, false>;                                                               \
      LAUNCHN("k_fast_inv_A", kfn, grid, block, st, c->devf(), src, dst, prime0, lw, mac_chain, m0);      \
    }                                                                                                     \
  } break;
    ACASE(4, false) ACASE(5, false) ACASE(6, false)
    ACASE(4, true) ACASE(5, true) ACASE(6, true) ACASE(7, true) ACASE(8, true) ACASE(9, true)
#undef ACASE
    default:
      return fail(TB200_EINVAL, "unsupported LA %d (LB %d)", c->LA, c->LB);
  }
  return 0;
}
// leading rows of [prime0, prime0 + rows) that take the FP64 butterflies, provided no later row does
// (prime order is [40-bit scale primes..., base, special...], so with every small prime on the FP64
// pipe this is the whole scale-prime prefix); 0 when the FP64 rows are scattered
static int f64_prefix(const tb200_ctx* c, int prime0, int rows) {
  int nf = 0;
  while (nf < rows && c->fps[prime0 + nf].f64) ++nf;
  for (int r = nf; r < rows; ++r)
    if (c->fps[prime0 + r].f64) return 0;
  return nf;
}
static TbView rows_from(TbView v, int r) {
  v.p += (long)r * v.rs;
  return v;
}
static int launch_fast_B_rows(const tb200_ctx* c, bool inverse, bool f64only, TbView src, TbView dst, int rows,
                              int batch, int prime0, tb200_stream st, int in_mode = 0,
                              const TbKsLevel* skip_lv = nullptr, int dbl_out = 0) {
  const int te = c->N < TB_TILE ? c->N : TB_TILE;
  // A CTA can walk over `bper` batch entries of one (limb, tile) to keep its twiddles in L1.  Measured
  // on B200 (logN16, chunk 16): bper = 8..16 is 12 % SLOWER than one entry per CTA at 2 or 3 CTAs/SM
  // (the per-CTA serialisation of load latency outweighs the L1 hits), so one entry per CTA it is.
  const int bper = 1;
  const int gz = (batch + bper - 1) / bper;
  const dim3 grid((unsigned)(c->N / te), (unsigned)rows, (unsigned)gz), block((unsigned)(te / 16));
  switch (c->LB * 2 + (f64only ? 1 : 0)) {
#define BCASE(n, F)                                                                              \
  case n * 2 + (F ? 1 : 0): {                                                                    \
    if (inverse && in_mode == TB_INV_IN_WIDE) {                                                  \
      auto kfn = k_fast_inv_B<n, F, TB_INV_IN_WIDE>;                                             \
      LAUNCHN("k_fast_inv_B", kfn, grid, block, st, c->devf(), src, dst, prime0, batch, bper);   \
    } else if (inverse && in_mode == TB_INV_IN_DOUBLE) {                                         \
      auto kfn = k_fast_inv_B<n, F, TB_INV_IN_DOUBLE>;                                           \
      LAUNCHN("k_fast_inv_B", kfn, grid, block, st, c->devf(), src, dst, prime0, batch, bper);   \
    } else if (inverse) {                                                                        \
      auto kfn = k_fast_inv_B<n, F, TB_INV_IN_LAZY>;                                             \
      LAUNCHN("k_fast_inv_B", kfn, grid, block, st, c->devf(), src, dst, prime0, batch, bper);   \
    } else {                                                                                     \
      auto kfn = k_fast_fwd_B<n, F>;                                                             \
      LAUNCHN("k_fast_fwd_B", kfn, grid, block, st, c->devf(), src, dst, prime0, batch, bper, skip_lv, dbl_out); \
    }                                                                                            \
  } break;
    BCASE(4, false) BCASE(5, false) BCASE(6, false) BCASE(7, false) BCASE(8, false)
    BCASE(4, true) BCASE(5, true) BCASE(6, true) BCASE(7, true) BCASE(8, true)
#undef BCASE
    default:
      return fail(TB200_EINVAL, "unsupported LB %d", c->LB);
  }
  return 0;
}
// FP64 rows and integer rows go to separate launches: the FP64-only kernels need 64 registers (4 CTAs per SM)
static int launch_fast_B(const tb200_ctx* c, bool inverse, TbView src, TbView dst, int rows, int batch, int prime0,
                         tb200_stream st, int in_mode = 0, const TbKsLevel* skip_lv = nullptr, int dbl_out = 0) {
  const int nf = f64_prefix(c, prime0, rows);
  int rc = 0;
  if (nf > 0 && (rc = launch_fast_B_rows(c, inverse, true, src, dst, nf, batch, prime0, st, in_mode, skip_lv, dbl_out)))
    return rc;
  if (nf < rows)
    rc = launch_fast_B_rows(c, inverse, false, rows_from(src, nf), rows_from(dst, nf), rows - nf, batch, prime0 + nf, st,
                            in_mode, skip_lv, dbl_out);
  return rc;
}
// Fork / join of the library-owned side stream around a section of the caller's stream (event pair; both are
// ordinary stream operations, so the section stays capturable in a CUDA graph).
static int side_stream_fork(tb200_ctx* c, tb200_stream st) {
  if (!c->side) {
    CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  CK(cudaEventRecord(c->ev_fork, (cudaStream_t)st));
  CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
  return 0;
}
static int side_stream_join(tb200_ctx* c, tb200_stream st) {
  CK(cudaEventRecord(c->ev_join, c->side));
  CK(cudaStreamWaitEvent((cudaStream_t)st, c->ev_join, 0));
  return 0;
}
// pass B of every digit group + key inner product + inverse pass B' for `rows` FP64 limb rows
static int launch_ks_core(const tb200_ctx* c, const TbKsCoreArgs& a, int rows, tb200_stream st) {
  const int te = c->N < TB_TILE ? c->N : TB_TILE;
  const dim3 grid((unsigned)((c->N / te) * a.nb), (unsigned)rows, 1u), block((unsigned)(te / 16));
  const size_t smem = (size_t)2 * te * 8;
  switch (c->LB) {
#define KCASE(n)                                                                                         \
  case n: {                                                                                              \
    auto kfn = k_fast_ks_core_f64<n>;                                                                    \
    TB_SET_MAX_DYN_SMEM(kfn, TB_KSCORE_STAGE_BYTES);                                                     \
    LAUNCH_DYN("k_fast_ks_core", kfn, grid, block, smem, st, c->devf(), a);                              \
  } break;
    KCASE(4) KCASE(5) KCASE(6) KCASE(7) KCASE(8)
#undef KCASE
    default:
      return fail(TB200_EINVAL, "unsupported LB %d", c->LB);
  }
  return 0;
}
// forward transform with the "enter" (x R) or "rescale + enter" prologue, mod q; dst dense or strided
// st_int: stream of the launches over the non-FP64 rows (default: st); they only depend on each other
// skip_b_f64 (out, optional): pass B of the leading FP64 rows is left to the caller (k_fast_fwd_B_tensor); receives
// their count (0: nothing was skipped)
static int fast_forward_enter(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0,
                              int rescale_level, tb200_stream st, tb200_stream st_int = nullptr,
                              int* skip_b_f64 = nullptr) {
  if (!st_int) st_int = st;
  if (skip_b_f64) *skip_b_f64 = 0;
  TbFwdAArgs a;
  memset(&a, 0, sizeof(a));
  a.src = src;
  a.dst = dst;
  a.prime0 = prime0;
  a.ngroups = 1;
  int rc;
  // the FP64 rows take the instantiation without the integer butterflies (64 registers, 4 CTAs per SM), the
  // remaining rows the generic one -- as for pass B and the ModUp launch
  const int nf = (c->LB == 8 && ntt_lw(c) == 12 - c->LA) ? f64_prefix(c, prime0, rows) : 0;
  if (rescale_level >= 0) {
    a.resc = c->d_resc3 + ((size_t)rescale_level * c->P + rescale_level + 1) * 3;
    a.round_at = (i64)(c->q[rescale_level] / 2);
    if (nf > 0) {
      if ((rc = launch_fast_fwd_A_rows<TB_FPRO_RESCALE_ENTER, true>(c, a, nf, batch, st))) return rc;
      if (nf < rows) {
        // rows nf..: the dropped limb stays row 0 of the source, so the body rows are addressed through `row_shift`
        TbFwdAArgs b = a;
        b.prime0 += nf;
        b.dst = rows_from(b.dst, nf);
        b.resc += 3 * (size_t)nf;
        b.row_shift = nf;
        if ((rc = launch_fast_fwd_A_rows<TB_FPRO_RESCALE_ENTER, false>(c, b, rows - nf, batch, st_int))) return rc;
      }
    } else if ((rc = launch_fast_fwd_A<TB_FPRO_RESCALE_ENTER>(c, a, rows, batch, st))) {
      return rc;
    }
  } else if (nf > 0) {
    if ((rc = launch_fast_fwd_A_rows<TB_FPRO_ENTER, true>(c, a, nf, batch, st))) return rc;
    if (nf < rows) {
      TbFwdAArgs b = a;
      b.prime0 += nf;
      b.dst = rows_from(b.dst, nf);
      b.row_shift = nf;
      if ((rc = launch_fast_fwd_A_rows<TB_FPRO_ENTER, false>(c, b, rows - nf, batch, st_int))) return rc;
    }
  } else if ((rc = launch_fast_fwd_A<TB_FPRO_ENTER>(c, a, rows, batch, st))) {
    return rc;
  }
  if (skip_b_f64 && nf > 0) {
    *skip_b_f64 = nf;
    if (nf < rows)
      return launch_fast_B_rows(c, false, false, rows_from(dst, nf), rows_from(dst, nf), rows - nf, batch, prime0 + nf, st_int);
    return 0;
  }
  if (st_int != st && nf > 0) {  // pass B split by hand: FP64 rows on st, the rest on st_int
    if ((rc = launch_fast_B_rows(c, false, true, dst, dst, nf, batch, prime0, st))) return rc;
    if (nf < rows)
      rc = launch_fast_B_rows(c, false, false, rows_from(dst, nf), rows_from(dst, nf), rows - nf, batch, prime0 + nf, st_int);
    return rc;
  }
  return launch_fast_B(c, false, dst, dst, rows, batch, prime0, st);
}
// mac_chain: the input is the output of k_fast_mac, whose FP64 limbs are doubles
static int fast_inverse_exit(const tb200_ctx* c, TbView src, TbView dst, int rows, int batch, int prime0,
                             tb200_stream st, int mac_chain, bool wide_in) {
  int rc = launch_fast_B(c, true, src, dst, rows, batch, prime0, st,
                         mac_chain ? TB_INV_IN_DOUBLE : (wide_in ? TB_INV_IN_WIDE : TB_INV_IN_LAZY));
  if (rc) return rc;
  return launch_fast_inv_A(c, dst, dst, rows, batch, prime0, mac_chain, st);
}

// ---- fused HE ops ---------------------------------------------------------------------------------
extern "C" int tb200_rescale_rows(tb200_ctx* c, int rows, int prime0, const tb200_poly* a, const int64_t* scales,
                                  const int64_t* rescaler, int64_t round_at, int exact, tb200_stream st) {
  if (!c || !scales || !rescaler) return fail(TB200_EINVAL, "rescale_rows: null argument");
  CHECK_ROWS(prime0, rows);
  CHECK_POLY(a);
  if (((uintptr_t)rescaler & 15) != 0) return fail(TB200_EINVAL, "rescaler must be 16-byte aligned");
  SET_DEVICE(c->device);
  TbView r;
  r.p = (i64*)rescaler;
  r.bs = 0;
  r.rs = 0;
  LAUNCH(k_rescale, grid_pw(c, rows, 1, 2), dim3(c->N / 2 < 256 ? c->N / 2 : 256), st, c->dev(), view(a), r, view(a),
         (const i64*)scales, prime0, c->N, (i64)round_at, exact);
  POST();
  return 0;
}

extern "C" int tb200_extend(tb200_ctx* c, int rows, int prime0, int alpha, const int64_t* state, int64_t state_stride,
                            const int64_t* l_enter, int64_t le_stride, int64_t le_off, int64_t* out,
                            int64_t out_stride, tb200_stream st) {
  if (!c || !state || !out || alpha < 1 || (alpha > 1 && !l_enter)) return fail(TB200_EINVAL, "extend: bad arguments");
  CHECK_ROWS(prime0, rows);
  SET_DEVICE(c->device);
  LAUNCH(k_extend_op, grid_pw(c, rows, 1, 1), dim3(256), st, c->dev(), (const i64*)state, (long)state_stride, alpha,
         (const i64*)l_enter, (long)le_stride, (long)le_off, (i64*)out, (long)out_stride, prime0, c->N);
  POST();
  return 0;
}

extern "C" int tb200_codec_rotate(tb200_ctx* c, int rows, const tb200_poly* a, const int64_t* perm,
                                  const int64_t* two_q, const tb200_poly* out, tb200_stream st) {
  if (!c || !perm || !two_q || rows < 1) return fail(TB200_EINVAL, "codec_rotate: bad arguments");
  CHECK_POLY(a);
  CHECK_POLY(out);
  if (a->ptr == out->ptr) return fail(TB200_EINVAL, "codec_rotate cannot run in place");
  SET_DEVICE(c->device);
  LAUNCH(k_codec_rotate_op, grid_pw(c, rows, 1, 1), dim3(256), st, view(a), (const i64*)perm, (const i64*)two_q,
         view(out), c->N);
  POST();
  return 0;
}

static int moddown(tb200_ctx* c, int lvl_, int batch, TbView cc, TbView p, TbView add, TbView out, int tail,
                   tb200_stream st) {
  const int level = c->lstart[lvl_];  // index of the first alive ordinary prime in the (local) prime table
  const int L = c->num_ord - level;
  LAUNCH(k_chain_backward, grid_pw(c, 1, batch, 1), dim3(256), st, c->dev(), p, (const i64*)c->d_pir_sp, c->K,
         c->num_ord, c->N);
  if (c->fast) {
    const dim3 g2 = grid_pw(c, L, batch, 2);
    const dim3 b2(c->N / 2 < 256 ? c->N / 2 : 256);
    if (tail == 0) {
      LAUNCH(k_fast_divide_by_p<0>, g2, b2, st, c->devf(), cc, p, add, out, (const u64*)c->d_bn, c->K, level, c->N);
    } else if (tail == 1) {
      LAUNCH(k_fast_divide_by_p<1>, g2, b2, st, c->devf(), cc, p, add, out, (const u64*)c->d_bn, c->K, level, c->N);
    } else {
      LAUNCH(k_fast_divide_by_p<2>, g2, b2, st, c->devf(), cc, p, add, out, (const u64*)c->d_bn, c->K, level, c->N);
    }
    return 0;
  }
  const dim3 grid = grid_pw(c, L, batch, 1);
  if (tail == 0) {
    LAUNCH(k_divide_by_p<0>, grid, dim3(256), st, c->dev(), cc, p, add, out, (const i64*)c->d_pir, c->K, level, c->N);
  } else if (tail == 1) {
    LAUNCH(k_divide_by_p<1>, grid, dim3(256), st, c->dev(), cc, p, add, out, (const i64*)c->d_pir, c->K, level, c->N);
  } else {
    LAUNCH(k_divide_by_p<2>, grid, dim3(256), st, c->dev(), cc, p, add, out, (const i64*)c->d_pir, c->K, level, c->N);
  }
  return 0;
}

extern "C" int tb200_divide_by_p(tb200_ctx* c, int level, const tb200_poly* cc, const tb200_poly* p,
                                 const tb200_poly* out, tb200_stream st) {
  if (!c || level < 0 || level >= c->num_levels) return fail(TB200_EINVAL, "divide_by_p: bad level");
  CHECK_POLY(cc);
  CHECK_POLY(p);
  CHECK_POLY(out);
  SET_DEVICE(c->device);
  int rc = moddown(c, level, 1, view(cc), view(p), view(cc), view(out), 0, st);
  if (rc) return rc;
  POST();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// engine layer
// ------------------------------------------------------------------------------------------------
static int check_level(const tb200_ctx* c, int level, int batch) {
  if (!c) return fail(TB200_EINVAL, "null context");
  if (level < 0 || level >= c->num_levels) return fail(TB200_EINVAL, "level %d outside [0, %d)", level, c->num_levels);
  if (batch < 1) return fail(TB200_EINVAL, "batch must be >= 1");
  return 0;
}
#define REQUIRE_UNSHARDED()                                                                                   \
  if (c->world > 1)                                                                                           \
  return fail(TB200_EINVAL, "this entry point needs an unsharded context; limb-sharded contexts support the " \
                            "op layer and tb200_ks_digits / tb200_ks_finish")
static inline TbView shift(TbView v, long b) {
  v.p += b * v.bs;
  return v;
}

static int rescale_impl(tb200_ctx* c, int level, int batch, TbView in, TbView out, int exact, tb200_stream st) {
  const int L = c->num_ord - level - 1;  // rows kept
  TbView body = in;
  body.p += in.rs;  // rows 1..L
  LAUNCH(k_rescale, grid_pw(c, L, batch, 2), dim3(c->N / 2 < 256 ? c->N / 2 : 256), st, c->dev(), body, in, out,
         (const i64*)(c->d_rescale + (size_t)level * c->P + level + 1), level + 1, c->N, (i64)(c->q[level] / 2), exact);
  return 0;
}

extern "C" int tb200_rescale(tb200_ctx* c, int level, int batch, const tb200_poly* in0, const tb200_poly* in1,
                             const tb200_poly* out0, const tb200_poly* out1, int exact, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  if (level + 1 >= c->num_ord) return fail(TB200_EINVAL, "rescale: level %d is the last level", level);
  CHECK_POLY(in0);
  CHECK_POLY(out0);
  SET_DEVICE(c->device);
  rescale_impl(c, level, batch, view(in0), view(out0), exact, st);
  if (in1) {
    CHECK_POLY(in1);
    CHECK_POLY(out1);
    rescale_impl(c, level, batch, view(in1), view(out1), exact, st);
  }
  POST();
  return 0;
}

static int make_key(const tb200_ctx* c, int level, const tb200_ksk* k, TbKskDev* out) {
  if (!k) return fail(TB200_EINVAL, "null key-switch key");
  if (k->row_stride < c->N || (k->row_stride & 1)) return fail(TB200_EINVAL, "ksk row_stride");
  memset(out, 0, sizeof(*out));
  out->rs = (long)k->row_stride;
  const TbKsLevel& lv = c->ks[level];
  for (int gi = 0; gi < lv.ngroups; ++gi) {
    const int gid = lv.g[gi].gid;
    if (gid >= k->num_groups || !k->b[gid] || !k->a[gid])
      return fail(TB200_EINVAL, "key-switch key lacks digit group %d needed at level %d", gid, level);
    if ((((uintptr_t)k->b[gid]) | ((uintptr_t)k->a[gid])) & 15) return fail(TB200_EINVAL, "ksk alignment");
    out->b[gid] = (const i64*)k->b[gid];
    out->a[gid] = (const i64*)k->a[gid];
  }
  return 0;
}

// workspace elements needed by the key switch of one ciphertext at `level` (state | ext | acc)
static size_t ks_ws_elems(const tb200_ctx* c, int level) {
  const size_t L = c->num_ord - c->lstart[level], E = L + c->K, ng = c->ks[level].ngroups;
  return ((size_t)c->ks[level].state_rows + ng * E + 2 * E) * (size_t)c->N;
}

// key switch, part 1: mixed-radix digits of the digit groups this rank owns -> rows of the state buffer
static int ks_digits(tb200_ctx* c, int level, int nb, TbView a, TbView state, tb200_stream st) {
  const TbKsLevel& lv = c->ks[level];
  if (lv.nown > 0)
    LAUNCH(k_digits, dim3((unsigned)((c->N / 2 + 255) / 256), (unsigned)lv.nown, (unsigned)nb),
           dim3(c->N / 2 < 256 ? c->N / 2 : 256), st, c->dev(), c->d_ks + level, a, state, c->N);
  return 0;
}

// key switch, part 2: from the complete digit state to the (local) output rows.
// tail: 0 -> out0 = ks0 ; 1 -> out = CS1(add + ks) for both ; 2 -> out0 = CS1(CS2(add0 + ks0)), out1 = ks1
// Relinearisation extras of the mod-q path: own_prefilled = the caller already wrote the (group, own limb)
// extensions in NTT form (k_fast_own_fill); nadd0/nadd1 = NTT-domain d0/d1 (dense [nb][L][N]) folded into
// the key inner product instead of being added after ModDown (k_fast_mac).
struct TbRelinExtra {
  bool own_prefilled;
  const i64 *nadd0, *nadd1;
  const i64* own_ntt;  // NTT-domain key-switch input (dense [nb][L][N]): the fused core reads the own limbs from it
};
// rows [0, nf) of a key switch at this level go through the fused core kernel
static int ks_core_rows(const tb200_ctx* c, int p0, int E) { return (c->fast && c->fused_core) ? f64_prefix(c, p0, E) : 0; }
// ModUp of the selected digit groups (sel: 0 every group, 1 the groups this rank owns, 2 the other ranks' groups):
// digits -> extended limbs, with forward pass A fused in on the mod-q path.  Splitting by ownership lets a limb-
// sharded key switch extend its own groups while the all-gather of the others is in flight (dist.py).
// This rank's share [s0, s1) of the K special limbs when their key sums are sharded as well (tb200_ks_core_sp):
// equal segments of ceil(K / world) limbs in rank order, the last ones possibly short or empty.
static void sp_share(const tb200_ctx* c, int* s0, int* s1, int* seg) {
  const int sg = (c->K + c->world - 1) / c->world;
  *seg = sg;
  *s0 = c->rank * sg < c->K ? c->rank * sg : c->K;
  *s1 = (c->rank + 1) * sg < c->K ? (c->rank + 1) * sg : c->K;
}
static int ks_modup(tb200_ctx* c, int level, int nb, TbView state, i64* ws, tb200_stream st, bool own_prefilled, int sel,
                    bool sp_own_only = false) {
  const int N = c->N, p0 = c->lstart[level], L = c->num_ord - p0, E = L + c->K;
  const TbKsLevel& lv = c->ks[level];
  const int ng = lv.ngroups;
  const TbKsLevel* dlv = c->d_ks + level;
  i64* ext = ws;
  int rc = 0;
  if (!c->fast) {  // exact op kernels: one launch over every group, issued with the last selection
    if (sel == 1 && lv.nforeign > 0) return 0;
    LAUNCH(k_extend_all, dim3((unsigned)((N / 2 + 255) / 256), (unsigned)E, (unsigned)(ng * nb)),
           dim3(N / 2 < 256 ? N / 2 : 256), st, c->dev(), dlv, (const i64*)c->d_lenter, state, ext, p0, N, E);
    return 0;
  }
  const int nsel = sel == 0 ? ng : (sel == 1 ? lv.nown : lv.nforeign);
  if (nsel == 0) return 0;
  TbFwdAArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.src = state;
  fa.dst = dense(ext, E, N);
  fa.lv = dlv;
  fa.lenter2 = c->d_lenter2;
  fa.lenterd = c->d_lenterd;
  fa.prime0 = p0;
  fa.ngroups = ng;
  fa.skip_own = own_prefilled ? 1 : 0;
  fa.sel = sel;
  fa.nsel = nsel;
  if (sp_own_only) {  // the ordinary rows, then only this rank's share of the special rows
    int s0, s1, seg;
    sp_share(c, &s0, &s1, &seg);
    if (L > 0 && (rc = launch_fast_fwd_A<TB_FPRO_EXTEND>(c, fa, L, nb * nsel, st))) return rc;
    if (s1 > s0) {
      TbFwdAArgs fb = fa;
      fb.prime0 += L + s0;
      fb.dst = rows_from(fb.dst, L + s0);
      if ((rc = launch_fast_fwd_A<TB_FPRO_EXTEND>(c, fb, s1 - s0, nb * nsel, st))) return rc;
    }
    return 0;
  }
  const int nf = ks_core_rows(c, p0, E);
  const bool split_a = c->LB == 8 && ntt_lw(c) == 12 - c->LA && nf > 0;  // FP64-only pass A instantiation exists
  if (split_a) {
    if ((rc = launch_fast_fwd_A_rows<TB_FPRO_EXTEND, true>(c, fa, nf, nb * nsel, st))) return rc;
    if (nf < E) {
      TbFwdAArgs fb = fa;
      fb.prime0 += nf;
      fb.dst = rows_from(fb.dst, nf);
      if ((rc = launch_fast_fwd_A_rows<TB_FPRO_EXTEND, false>(c, fb, E - nf, nb * nsel, st))) return rc;
    }
    return 0;
  }
  return launch_fast_fwd_A<TB_FPRO_EXTEND>(c, fa, E, nb * nsel, st);
}

// modup_done: the caller already ran ks_modup for every group (split call of a limb-sharded key switch)
static int ks_finish(tb200_ctx* c, int level, int nb, TbView state, const TbKskDev& key, TbView add0, TbView add1,
                     TbView out0, TbView out1, int tail, i64* ws, tb200_stream st, const TbRelinExtra* ex = nullptr,
                     bool modup_done = false) {
  const bool own_prefilled = ex && ex->own_prefilled;
  const int N = c->N, p0 = c->lstart[level], L = c->num_ord - p0, E = L + c->K;
  const TbKsLevel& lv = c->ks[level];
  const int ng = lv.ngroups;
  const TbKsLevel* dlv = c->d_ks + level;
  i64* ext = ws;
  i64* acc = ext + (size_t)nb * ng * E * N;
  const TbDev d = c->dev();
  int rc = 0;
  if (!modup_done && (rc = ks_modup(c, level, nb, state, ws, st, own_prefilled, 0))) return rc;
  if (c->fast) {
    // FP64 limbs: pass B of every group + key inner product + inverse pass B' in one kernel (tb200_ks_core.cuh);
    // the remaining limbs (60-bit base / special primes) take the three separate kernels on their rows
    // (optionally on a forked stream: they load the integer pipes, the FP64 rows the FP64 pipe).
    const int nf = ks_core_rows(c, p0, E);
    tb200_stream sst = st;
    if (nf > 0 && nf < E && (c->side_rows & 1) && !g_prof_on) {
      if ((rc = side_stream_fork(c, st))) return rc;
      sst = (tb200_stream)c->side;
    }
    if (nf > 0) {
      TbKsCoreArgs ka;
      memset(&ka, 0, sizeof(ka));
      ka.lv = dlv;
      ka.key = key;
      ka.ext = ext;
      ka.acc = acc;
      ka.nadd0 = ex ? ex->nadd0 : nullptr;
      ka.nadd1 = ex ? ex->nadd1 : nullptr;
      ka.p0 = p0;
      ka.row0 = 0;
      ka.N = N;
      ka.rowsE = E;
      ka.nb = nb;
      ka.skip_own = own_prefilled ? 1 : 0;
      ka.own = ex ? ex->own_ntt : nullptr;
      ka.pr = c->d_primes;
      if ((rc = launch_ks_core(c, ka, nf, st))) return rc;
    }
    if (nf < E) {
      const TbView er = rows_from(dense(ext, E, N), nf);
      if ((rc = launch_fast_B(c, false, er, er, E - nf, nb * ng, p0 + nf, sst, 0, own_prefilled ? dlv : nullptr, 1)))
        return rc;
      // key inner product, 128-bit accumulation over the groups
      LAUNCH(k_fast_mac, dim3((unsigned)(((N / 2 + 255) / 256) * nb), (unsigned)(E - nf), 1u),
             dim3(N / 2 < 256 ? N / 2 : 256), sst, d, c->devf(), dlv, key, (const i64*)ext, acc, p0, N, E, nb,
             ex ? ex->nadd0 : (const i64*)nullptr, ex ? ex->nadd1 : (const i64*)nullptr, (const i64*)c->d_cP, nf, 0u);
      const TbView ar = rows_from(dense(acc, E, N), nf);
      if ((rc = launch_fast_B(c, true, ar, ar, E - nf, nb * 2, p0 + nf, sst, TB_INV_IN_DOUBLE))) return rc;
    }
    if (sst != st && (rc = side_stream_join(c, st))) return rc;
    if (ex && ex->nadd0) tail = 0;  // already inside the sums
    if (c->fused_moddown) {
      // special limbs first (inverse pass A' + exit, then the exact chain-backward step), then the ordinary
      // limbs with ModDown and the tail fused into the exit: they go straight to the output rows
      const TbView sp = rows_from(dense(acc, E, N), L);
      if ((rc = launch_fast_inv_A(c, sp, sp, c->K, nb * 2, p0 + L, 1, st))) return rc;
      LAUNCH(k_chain_backward, grid_pw(c, 1, nb * 2, 1), dim3(256), st, c->dev(), sp, (const i64*)c->d_pir_sp, c->K,
             c->num_ord, c->N);
      TbMdArgs md;
      memset(&md, 0, sizeof(md));
      md.p = acc + (size_t)L * N;
      md.pbs = (long)E * N;
      md.bn = c->d_bn;
      md.add0 = add0;
      md.add1 = add1;
      md.out0 = out0;
      md.out1 = out1;
      md.K = c->K;
      md.tail = tail;
      return launch_fast_inv_A(c, dense(acc, E, N), dense(acc, E, N), L, nb * 2, p0, 1, st, &md);
    }
    // inverse pass A' + exit: back to coefficients, canonical
    if ((rc = launch_fast_inv_A(c, dense(acc, E, N), dense(acc, E, N), E, nb * 2, p0, 1, st))) return rc;
  } else {
    if ((rc = ntt_forward(c, dense(ext, E, N), dense(ext, E, N), E, nb * ng, p0, false, st))) return rc;
    LAUNCH(k_mac, dim3((unsigned)((N / 2 + 255) / 256), (unsigned)E, (unsigned)nb), dim3(N / 2 < 256 ? N / 2 : 256), st,
           d, dlv, key, (const i64*)ext, acc, p0, N, E);
    if ((rc = ntt_inverse(c, dense(acc, E, N), dense(acc, E, N), E, nb * 2, p0, 2, st))) return rc;
  }
  // ModDown (+ fused tail)
  for (int h = 0; h < 2; ++h) {
    TbView cc;
    cc.p = acc + (size_t)h * E * N;
    cc.bs = 2L * E * N;
    cc.rs = N;
    TbView p = cc;
    p.p += (size_t)L * N;
    const int t = (tail == 2 && h == 1) ? 0 : tail;
    if ((rc = moddown(c, level, nb, cc, p, h == 0 ? add0 : add1, h == 0 ? out0 : out1, t, st))) return rc;
  }
  return 0;
}

// key switch of `nb` polynomials (nb <= chunk) on an unsharded context. a: coefficient canonical [L][N].
static int keyswitch_chunk(tb200_ctx* c, int level, int nb, TbView a, const TbKskDev& key, TbView add0, TbView add1,
                           TbView out0, TbView out1, int tail, i64* ws, tb200_stream st,
                           const TbRelinExtra* ex = nullptr) {
  const int S = c->ks[level].state_rows;
  TbView state = dense(ws, S, c->N);
  int rc = ks_digits(c, level, nb, a, state, st);
  if (rc) return rc;
  return ks_finish(c, level, nb, state, key, add0, add1, out0, out1, tail, ws + (size_t)nb * S * c->N, st, ex);
}

extern "C" int tb200_ks_state_info(const tb200_ctx* c, int level, int32_t* out) {
  if (!c || !out || level < 0 || level >= c->num_levels) return fail(TB200_EINVAL, "ks_state_info: bad arguments");
  const TbKsLevel& lv = c->ks[level];
  out[0] = lv.state_rows;
  out[1] = lv.seg_rows * c->rank;
  out[2] = lv.seg_rows;
  out[3] = c->num_ord - c->lstart[level];
  return 0;
}
extern "C" int tb200_ctx_local_primes(const tb200_ctx* c, int32_t* out) {
  if (!c || !out) return fail(TB200_EINVAL, "null");
  for (int i = 0; i < c->num_ord; ++i) out[i] = c->ord_gid[i];
  for (int kk = 0; kk < c->K; ++kk) out[c->num_ord + kk] = (int)c->qg.size() - c->K + kk;
  return 0;
}
extern "C" int tb200_ks_digits(tb200_ctx* c, int level, int batch, const tb200_poly* a, const tb200_poly* state,
                               tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  if (c->ks[level].nown > 0) CHECK_POLY(a);
  CHECK_POLY(state);
  SET_DEVICE(c->device);
  if ((rc = ks_digits(c, level, batch, view(a), view(state), st))) return rc;
  POST();
  return 0;
}
extern "C" int tb200_ks_finish(tb200_ctx* c, int level, int batch, const tb200_poly* state, const tb200_ksk* ksk,
                               const tb200_poly* add0, const tb200_poly* add1, const tb200_poly* out0,
                               const tb200_poly* out1, int tail, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  if (tail < 0 || tail > 2) return fail(TB200_EINVAL, "ks_finish: tail must be 0, 1 or 2");
  CHECK_POLY(state);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  if (tail != 0) CHECK_POLY(add0);
  if (tail == 1) CHECK_POLY(add1);
  TbKskDev key;
  if ((rc = make_key(c, level, ksk, &key))) return rc;
  SET_DEVICE(c->device);
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve(ks_ws_elems(c, level) * ch))) return rc;
  for (int b0 = 0; b0 < batch; b0 += ch) {
    const int nb = batch - b0 < ch ? batch - b0 : ch;
    rc = ks_finish(c, level, nb, shift(view(state), b0), key, shift(view(add0 ? add0 : out0), b0),
                   shift(view(add1 ? add1 : out1), b0), shift(view(out0), b0), shift(view(out1), b0), tail, ws.p, st);
    if (rc) return rc;
  }
  POST();
  return 0;
}

// The same in two calls, for overlapping the digit all-gather of a limb-sharded key switch with the ModUp of the
// groups this rank owns: tb200_ks_modup(which = 1) while the collective runs, tb200_ks_modup(which = 2) after it,
// then tb200_ks_core.  The extended limbs live in the context workspace between the calls: batch <= chunk.
extern "C" int tb200_ks_modup(tb200_ctx* c, int level, int batch, const tb200_poly* state, int which, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  const bool sp_own_only = (which & 4) != 0;  // + 4: of the special rows only this rank's share (tb200_ks_core_sp)
  which &= ~4;
  if (which < 0 || which > 2) return fail(TB200_EINVAL, "ks_modup: which must be 0 (all), 1 (own) or 2 (foreign), + 4");
  if (sp_own_only && !c->fast) return fail(TB200_EINVAL, "ks_modup: the special-limb share needs the mod-q path");
  if (batch > c->chunk) return fail(TB200_EINVAL, "ks_modup: batch %d exceeds the chunk %d", batch, c->chunk);
  CHECK_POLY(state);
  SET_DEVICE(c->device);
  if ((rc = ws_reserve(c, ks_ws_elems(c, level) * (size_t)batch))) return rc;
  if ((rc = ks_modup(c, level, batch, view(state), c->ws, st, false, which, sp_own_only))) return rc;
  POST();
  return 0;
}
extern "C" int tb200_ks_core(tb200_ctx* c, int level, int batch, const tb200_poly* state, const tb200_ksk* ksk,
                             const tb200_poly* add0, const tb200_poly* add1, const tb200_poly* out0,
                             const tb200_poly* out1, int tail, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  if (tail < 0 || tail > 2) return fail(TB200_EINVAL, "ks_core: tail must be 0, 1 or 2");
  if (batch > c->chunk) return fail(TB200_EINVAL, "ks_core: batch %d exceeds the chunk %d", batch, c->chunk);
  CHECK_POLY(state);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  if (tail != 0) CHECK_POLY(add0);
  if (tail == 1) CHECK_POLY(add1);
  TbKskDev key;
  if ((rc = make_key(c, level, ksk, &key))) return rc;
  SET_DEVICE(c->device);
  if (c->ws_elems < ks_ws_elems(c, level) * (size_t)batch) return fail(TB200_EINVAL, "ks_core: call tb200_ks_modup first");
  rc = ks_finish(c, level, batch, view(state), key, view(add0 ? add0 : out0), view(add1 ? add1 : out1), view(out0),
                 view(out1), tail, c->ws, st, nullptr, true);
  if (rc) return rc;
  POST();
  return 0;
}

// ---- limb-sharded key switch with the special limbs sharded as well ---------------------------------------------
// tb200_ks_finish / tb200_ks_core replicate the K special limbs: every rank extends, transforms and multiplies all of
// them (the reference does the same, rns_partition.py:34-52).  They are 60-bit limbs (2.5x the cost of a scale limb),
// so at logN17 (73 + 6 limbs) they cap the speed-up at 1.66x / 2.5x on 2 / 4 GPUs.  Here every rank computes the key
// sums of its share of them only and ONE more all-gather (2 K N words per polynomial, issued by the caller while the
// ordinary limbs are still being processed) completes them before ModDown:
//   tb200_ks_modup(which | 4)   extend + pass A of the ordinary rows and of the rank's share of the special rows
//   tb200_ks_core_sp            pass B + key inner product + inverse transform of that share -> its segment of `sp`
//   -- all-gather of the segments (asynchronous) --
//   tb200_ks_core_ord           the same for the local ordinary rows (sums stay in the context workspace)
//   tb200_ks_moddown            chain-backward on the complete special limbs + ModDown + tail
// sp: dense [world * ceil(K / world)][batch][2][N] (special limb, polynomial, key half); rank r owns rows
// [r seg, (r + 1) seg).  Results are the same bits as tb200_ks_finish (every step is the same arithmetic).
static int ks_sums(tb200_ctx* c, int level, int nb, const TbKskDev& key, i64* ws, int part, i64* sp, tb200_stream st) {
  const int N = c->N, p0 = c->lstart[level], L = c->num_ord - p0, E = L + c->K;
  const TbKsLevel& lv = c->ks[level];
  const int ng = lv.ngroups;
  const TbKsLevel* dlv = c->d_ks + level;
  i64* ext = ws;
  i64* acc = ext + (size_t)nb * ng * E * N;
  int s0, s1, seg, rc = 0;
  sp_share(c, &s0, &s1, &seg);
  const int r0 = part == 0 ? 0 : L + s0, r1 = part == 0 ? L : L + s1;
  if (r1 <= r0) return 0;
  int nf = part == 0 ? ks_core_rows(c, p0, E) : 0;  // leading rows of the range that go through the fused FP64 core
  if (nf > L) nf = L;
  if (nf > 0) {
    TbKsCoreArgs ka;
    memset(&ka, 0, sizeof(ka));
    ka.lv = dlv;
    ka.key = key;
    ka.ext = ext;
    ka.acc = acc;
    ka.p0 = p0;
    ka.N = N;
    ka.rowsE = E;
    ka.nb = nb;
    ka.pr = c->d_primes;
    if ((rc = launch_ks_core(c, ka, nf, st))) return rc;
  }
  const int i0 = r0 > nf ? r0 : nf;
  if (i0 < r1) {
    const TbView er = rows_from(dense(ext, E, N), i0);
    if ((rc = launch_fast_B(c, false, er, er, r1 - i0, nb * ng, p0 + i0, st, 0, nullptr, 1))) return rc;
    LAUNCH(k_fast_mac, dim3((unsigned)(((N / 2 + 255) / 256) * nb), (unsigned)(r1 - i0), 1u),
           dim3(N / 2 < 256 ? N / 2 : 256), st, c->dev(), c->devf(), dlv, key, (const i64*)ext, acc, p0, N, E, nb,
           (const i64*)nullptr, (const i64*)nullptr, (const i64*)c->d_cP, i0, 0u);
    const TbView ar = rows_from(dense(acc, E, N), i0);
    if ((rc = launch_fast_B(c, true, ar, ar, r1 - i0, nb * 2, p0 + i0, st, TB_INV_IN_DOUBLE))) return rc;
  }
  // inverse pass A' + exit; the special share goes straight to its segment of sp ([k][polynomial][half][N])
  const TbView src = rows_from(dense(acc, E, N), r0);
  TbView dst = src;
  if (part == 1) {
    dst.p = sp + (size_t)s0 * nb * 2 * N;
    dst.bs = N;
    dst.rs = (long)nb * 2 * N;
  }
  return launch_fast_inv_A(c, src, dst, r1 - r0, nb * 2, p0 + r0, 1, st);
}
static int ks_sp_check(tb200_ctx* c, int level, int batch, const char* what) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  if (!c->fast) return fail(TB200_EINVAL, "%s: needs the mod-q path (tb200_ctx_set_fast)", what);
  if (batch > c->chunk) return fail(TB200_EINVAL, "%s: batch %d exceeds the chunk %d", what, batch, c->chunk);
  if (c->ws_elems < ks_ws_elems(c, level) * (size_t)batch) return fail(TB200_EINVAL, "%s: call tb200_ks_modup first", what);
  return 0;
}
extern "C" int tb200_ks_sp_info(const tb200_ctx* c, int32_t* out) {
  if (!c || !out) return fail(TB200_EINVAL, "ks_sp_info: null argument");
  int s0, s1, seg;
  sp_share(c, &s0, &s1, &seg);
  out[0] = seg * c->world;  // rows of the sp buffer
  out[1] = seg;             // rows per rank segment (segment r starts at row r * seg)
  out[2] = s0;
  out[3] = s1;
  return 0;
}
extern "C" int tb200_ks_core_sp(tb200_ctx* c, int level, int batch, const tb200_ksk* ksk, int64_t* sp, tb200_stream st) {
  int rc = ks_sp_check(c, level, batch, "ks_core_sp");
  if (rc) return rc;
  if (!sp || ((uintptr_t)sp & 15)) return fail(TB200_EINVAL, "ks_core_sp: sp must be a 16-byte aligned device pointer");
  TbKskDev key;
  if ((rc = make_key(c, level, ksk, &key))) return rc;
  SET_DEVICE(c->device);
  if ((rc = ks_sums(c, level, batch, key, c->ws, 1, sp, st))) return rc;
  POST();
  return 0;
}
extern "C" int tb200_ks_core_ord(tb200_ctx* c, int level, int batch, const tb200_ksk* ksk, tb200_stream st) {
  int rc = ks_sp_check(c, level, batch, "ks_core_ord");
  if (rc) return rc;
  TbKskDev key;
  if ((rc = make_key(c, level, ksk, &key))) return rc;
  SET_DEVICE(c->device);
  if ((rc = ks_sums(c, level, batch, key, c->ws, 0, nullptr, st))) return rc;
  POST();
  return 0;
}
extern "C" int tb200_ks_moddown(tb200_ctx* c, int level, int batch, int64_t* sp, const tb200_poly* add0,
                                const tb200_poly* add1, const tb200_poly* out0, const tb200_poly* out1, int tail,
                                tb200_stream st) {
  int rc = ks_sp_check(c, level, batch, "ks_moddown");
  if (rc) return rc;
  if (tail < 0 || tail > 2) return fail(TB200_EINVAL, "ks_moddown: tail must be 0, 1 or 2");
  if (!sp || ((uintptr_t)sp & 15)) return fail(TB200_EINVAL, "ks_moddown: sp must be a 16-byte aligned device pointer");
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  if (tail != 0) CHECK_POLY(add0);
  if (tail == 1) CHECK_POLY(add1);
  SET_DEVICE(c->device);
  const int N = c->N, p0 = c->lstart[level], L = c->num_ord - p0, E = L + c->K, ng = c->ks[level].ngroups;
  if (L < 1) {  // a rank without ordinary limbs at this level only contributed its special share
    POST();
    return 0;
  }
  i64* acc = c->ws + (size_t)batch * ng * E * N;
  for (int h = 0; h < 2; ++h) {
    TbView cc;
    cc.p = acc + (size_t)h * E * N;
    cc.bs = 2L * E * N;
    cc.rs = N;
    TbView p;
    p.p = sp + (size_t)h * N;
    p.bs = 2L * N;
    p.rs = (long)batch * 2 * N;
    const int t = (tail == 2 && h == 1) ? 0 : tail;
    const tb200_poly* add = h == 0 ? add0 : add1;
    const tb200_poly* out = h == 0 ? out0 : out1;
    if ((rc = moddown(c, level, batch, cc, p, view(add ? add : out), view(out), t, st))) return rc;
  }
  POST();
  return 0;
}

extern "C" int tb200_keyswitch(tb200_ctx* c, int level, int batch, const tb200_poly* a, const tb200_ksk* ksk,
                               const tb200_poly* out0, const tb200_poly* out1, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  CHECK_POLY(a);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  TbKskDev key;
  if ((rc = make_key(c, level, ksk, &key))) return rc;
  SET_DEVICE(c->device);
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve(ks_ws_elems(c, level) * ch))) return rc;
  for (int b0 = 0; b0 < batch; b0 += ch) {
    const int nb = batch - b0 < ch ? batch - b0 : ch;
    rc = keyswitch_chunk(c, level, nb, shift(view(a), b0), key, shift(view(a), b0), shift(view(a), b0),
                         shift(view(out0), b0), shift(view(out1), b0), 0, ws.p, st);
    if (rc) return rc;
  }
  POST();
  return 0;
}

extern "C" int tb200_switch_key(tb200_ctx* c, int level, int batch, const tb200_poly* c0, const tb200_poly* c1,
                                const tb200_ksk* ksk, const tb200_poly* out0, const tb200_poly* out1,
                                tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  CHECK_POLY(c0);
  CHECK_POLY(c1);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  TbKskDev key;
  if ((rc = make_key(c, level, ksk, &key))) return rc;
  SET_DEVICE(c->device);
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve(ks_ws_elems(c, level) * ch))) return rc;
  for (int b0 = 0; b0 < batch; b0 += ch) {
    const int nb = batch - b0 < ch ? batch - b0 : ch;
    rc = keyswitch_chunk(c, level, nb, shift(view(c1), b0), key, shift(view(c0), b0), shift(view(c0), b0),
                         shift(view(out0), b0), shift(view(out1), b0), 2, ws.p, st);
    if (rc) return rc;
  }
  POST();
  return 0;
}

extern "C" int tb200_rotate(tb200_ctx* c, int level, int batch, int64_t galois, const tb200_poly* c0,
                            const tb200_poly* c1, const tb200_ksk* rotk, const tb200_poly* out0,
                            const tb200_poly* out1, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  if (rotk) REQUIRE_UNSHARDED();  // the automorphism alone is limb-local; the key switch is dist.py's on sharded contexts
  CHECK_POLY(c0);
  CHECK_POLY(c1);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  if (!(galois & 1) || galois < 1 || galois >= 2 * (int64_t)c->N)
    return fail(TB200_EINVAL, "galois element must be odd and in [1, 2N)");
  if (c0->ptr == out0->ptr || c1->ptr == out1->ptr) return fail(TB200_EINVAL, "rotate cannot run in place");
  SET_DEVICE(c->device);
  const int N = c->N, L = c->num_ord - c->lstart[level];
  if (!rotk) {
    if (L > 0) {
      const dim3 gridb((unsigned)((N + 255) / 256), (unsigned)L, (unsigned)batch);
      LAUNCH(k_automorphism, gridb, dim3(256), st, c->dev(), view(c0), view(out0), c->lstart[level], (i64)galois);
      LAUNCH(k_automorphism, gridb, dim3(256), st, c->dev(), view(c1), view(out1), c->lstart[level], (i64)galois);
    }
    POST();
    return 0;
  }
  TbKskDev key;
  if ((rc = make_key(c, level, rotk, &key))) return rc;
  const int ch = batch < c->chunk ? batch : c->chunk;
  const size_t rot_elems = 2 * (size_t)L * N;  // rotated c0, c1 per ciphertext
  WsLease ws(c, st);
  if ((rc = ws.reserve((ks_ws_elems(c, level) + rot_elems) * ch))) return rc;
  for (int b0 = 0; b0 < batch; b0 += ch) {
    const int nb = batch - b0 < ch ? batch - b0 : ch;
    i64* r0 = ws.p;
    i64* r1 = r0 + (size_t)nb * L * N;
    i64* ksws = r1 + (size_t)nb * L * N;
    const dim3 gridb((unsigned)((N + 255) / 256), (unsigned)L, (unsigned)nb);
    LAUNCH(k_automorphism, gridb, dim3(256), st, c->dev(), shift(view(c0), b0), dense(r0, L, N), level, (i64)galois);
    LAUNCH(k_automorphism, gridb, dim3(256), st, c->dev(), shift(view(c1), b0), dense(r1, L, N), level, (i64)galois);
    rc = keyswitch_chunk(c, level, nb, dense(r1, L, N), key, dense(r0, L, N), dense(r0, L, N), shift(view(out0), b0),
                         shift(view(out1), b0), 2, ksws, st);
    if (rc) return rc;
  }
  POST();
  return 0;
}

// Hoisted rotations: `nrot` rotations of the SAME ciphertext(s) share the ModUp digits, the extension and its
// forward transform; per rotation only the key inner product (reading the extension through the NTT-domain
// automorphism), the inverse transform, ModDown and the coefficient-domain automorphism of c0 remain.
//   out_r = (sigma_r(c0) + ks0_r, ks1_r),  ks_r = ModDown(sum_g sigma_r(NTT(ext_g)) (x) rotk_r[g]).
// An extension beyond the reference (it rotates one key at a time, ckks_engine.py:1804-1840, 1908-1926): the result
// decrypts to the same rotation as rotate_single but is NOT bit-identical to it (the digits are taken before the
// automorphism, so negated coefficients lift to other representatives); oracle/engine.py: rotate_hoisted restates
// it, tests compare bit for bit with that and decrypt.
extern "C" int tb200_rotate_hoisted(tb200_ctx* c, int level, int batch, int nrot, const int64_t* galois,
                                    const tb200_poly* c0, const tb200_poly* c1, const tb200_ksk* const* rotks,
                                    const tb200_poly* out0, const tb200_poly* out1, int64_t rot_stride,
                                    tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  if (nrot < 1 || !galois || !rotks) return fail(TB200_EINVAL, "rotate_hoisted: need nrot >= 1, galois and keys");
  if (!c->fast) return fail(TB200_EINVAL, "rotate_hoisted runs on the mod-q path (tb200_ctx_set_fast(ctx, 1))");
  CHECK_POLY(c0);
  CHECK_POLY(c1);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  if (rot_stride & 1) return fail(TB200_EINVAL, "rotate_hoisted: rot_stride must be even");
  std::vector<TbKskDev> keys(nrot);
  for (int r = 0; r < nrot; ++r) {
    if (!(galois[r] & 1) || galois[r] < 1 || galois[r] >= 2 * (int64_t)c->N)
      return fail(TB200_EINVAL, "galois element %d must be odd and in [1, 2N)", r);
    if ((rc = make_key(c, level, rotks[r], &keys[r]))) return rc;
  }
  SET_DEVICE(c->device);
  const int N = c->N, L = c->num_ord - level, E = L + c->K;
  const TbKsLevel& lv = c->ks[level];
  const int ng = lv.ngroups, S = lv.state_rows;
  const TbKsLevel* dlv = c->d_ks + level;
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve((ks_ws_elems(c, level) + (size_t)L * N) * ch))) return rc;
  for (int b0 = 0; b0 < batch; b0 += ch) {
    const int nb = batch - b0 < ch ? batch - b0 : ch;
    i64* state = ws.p;
    i64* ext = state + (size_t)nb * S * N;
    i64* acc = ext + (size_t)nb * ng * E * N;
    i64* r0 = acc + (size_t)nb * 2 * E * N;
    // once: digits, ModUp + forward transform of every group over all E limbs
    if ((rc = ks_digits(c, level, nb, shift(view(c1), b0), dense(state, S, N), st))) return rc;
    if ((rc = ks_modup(c, level, nb, dense(state, S, N), ext, st, false, 0))) return rc;
    if ((rc = launch_fast_B(c, false, dense(ext, E, N), dense(ext, E, N), E, nb * ng, level, st, 0, nullptr, 1))) return rc;
    for (int r = 0; r < nrot; ++r) {
      LAUNCH(k_fast_mac, dim3((unsigned)(((N / 2 + 255) / 256) * nb), (unsigned)E, 1u), dim3(N / 2 < 256 ? N / 2 : 256), st,
             c->dev(), c->devf(), dlv, keys[r], (const i64*)ext, acc, level, N, E, nb, (const i64*)nullptr,
             (const i64*)nullptr, (const i64*)c->d_cP, 0, (unsigned)galois[r]);
      if ((rc = fast_inverse_exit(c, dense(acc, E, N), dense(acc, E, N), E, nb * 2, level, st, 1))) return rc;
      const dim3 gridb((unsigned)((N + 255) / 256), (unsigned)L, (unsigned)nb);
      LAUNCH(k_automorphism, gridb, dim3(256), st, c->dev(), shift(view(c0), b0), dense(r0, L, N), level, (i64)galois[r]);
      TbView o0 = shift(view(out0), b0), o1 = shift(view(out1), b0);
      o0.p += (long)r * rot_stride;
      o1.p += (long)r * rot_stride;
      for (int h = 0; h < 2; ++h) {
        TbView cc;
        cc.p = acc + (size_t)h * E * N;
        cc.bs = 2L * E * N;
        cc.rs = N;
        TbView p = cc;
        p.p += (size_t)L * N;
        if ((rc = moddown(c, level, nb, cc, p, dense(r0, L, N), h == 0 ? o0 : o1, h == 0 ? 2 : 0, st))) return rc;
      }
    }
  }
  POST();
  return 0;
}

// forward transforms + tensor product of one chunk; x: 4 polys [4][nb][L][N] workspace (x0,x1,y0,y1)
static int mult_front(tb200_ctx* c, int level, int nb, TbView a0, TbView a1, TbView b0, TbView b1, int pre_rescale,
                      i64* x, TbView d0, TbView d1, TbView d2, bool fast, tb200_stream st) {
  const int N = c->N;
  const int lvl = c->lstart[level + (pre_rescale ? 1 : 0)];  // index of the first alive prime in the (local) table
  const int L = c->num_ord - lvl;
  const size_t pe = (size_t)nb * L * N;
  TbView in[4] = {a0, a1, b0, b1};
  // squaring (both operands are the same tensors): the transformed operands are identical, so two of the
  // four rescale + forward transforms suffice and the tensor product reads them twice
  const bool square = a0.p == b0.p && a1.p == b1.p && a0.bs == b0.bs && a1.bs == b1.bs && a0.rs == b0.rs &&
                      a1.rs == b1.rs;
  // TB200_TUNE_SIDE_ROWS bit 1: the 60-bit rows of the input transforms (a third of a wave per launch) run on the
  // forked side stream under the FP64 rows' launches
  tb200_stream sst = st;
  if (fast && (c->side_rows & 2) && !g_prof_on) {
    int rc = side_stream_fork(c, st);
    if (rc) return rc;
    sst = (tb200_stream)c->side;
  }
  // the last operand's pass B runs fused with the tensor product on the FP64 rows (k_fast_fwd_B_tensor)
  const bool fuse = fast && !square && c->fused_tensor && c->LB == 8;
  int nfused = 0;
  for (int i = 0; i < (square ? 2 : 4); ++i) {
    TbView xi = dense(x + i * pe, L, N);
    if (fast) {
      int rc = fast_forward_enter(c, in[i], xi, L, nb, lvl, pre_rescale ? level : -1, st, sst,
                                  (fuse && i == 3) ? &nfused : nullptr);
      if (rc) return rc;
    } else if (pre_rescale) {
      rescale_impl(c, level, nb, in[i], xi, 1, st);
      int rc = ntt_forward(c, xi, xi, L, nb, lvl, true, st);
      if (rc) return rc;
    } else {
      int rc = ntt_forward(c, in[i], xi, L, nb, lvl, true, st);
      if (rc) return rc;
    }
  }
  if (sst != st) {
    int rc = side_stream_join(c, st);
    if (rc) return rc;
  }
  if (nfused > 0) {
    const int te = N < TB_TILE ? N : TB_TILE;
    auto kfn = k_fast_fwd_B_tensor<8>;
    LAUNCHN("k_fast_fwd_B_tensor", kfn, dim3((unsigned)(N / te), (unsigned)nfused, (unsigned)nb), dim3((unsigned)(te / 16)), st,
            c->devf(), dense(x + 3 * pe, L, N), dense(x, L, N), dense(x + pe, L, N), dense(x + 2 * pe, L, N), d0, d1, d2, lvl);
    if (nfused == L) return 0;
  }
  const int r0 = nfused;  // rows left to the stand-alone tensor product (60-bit limbs; every row without the fusion)
  LAUNCH(k_tensor, grid_pw(c, L - r0, nb, 2), dim3(N / 2 < 256 ? N / 2 : 256), st, c->dev(),
         rows_from(dense(x, L, N), r0), rows_from(dense(x + pe, L, N), r0),
         rows_from(dense(x + (square ? 0 : 2) * pe, L, N), r0), rows_from(dense(x + (square ? 1 : 3) * pe, L, N), r0),
         rows_from(d0, r0), rows_from(d1, r0), rows_from(d2, r0), lvl + r0, N);
  return 0;
}

extern "C" int tb200_cc_mult_triplet(tb200_ctx* c, int level, int batch, const tb200_poly* a0, const tb200_poly* a1,
                                     const tb200_poly* b0, const tb200_poly* b1, const tb200_poly* d0,
                                     const tb200_poly* d1, const tb200_poly* d2, int pre_rescale, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  if (pre_rescale) REQUIRE_UNSHARDED();  // the tensor product is limb-local; the rescale needs the dropped limb (dist.py)
  if (pre_rescale && level + 1 >= c->num_ord) return fail(TB200_EINVAL, "cc_mult: no level left to rescale");
  CHECK_POLY(a0);
  CHECK_POLY(a1);
  CHECK_POLY(b0);
  CHECK_POLY(b1);
  CHECK_POLY(d0);
  CHECK_POLY(d1);
  CHECK_POLY(d2);
  SET_DEVICE(c->device);
  const int lvl = level + (pre_rescale ? 1 : 0), L = c->num_ord - c->lstart[lvl];
  if (L < 1) {  // a rank that owns no limb at this level
    POST();
    return 0;
  }
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve(4 * (size_t)ch * L * c->N))) return rc;
  for (int b = 0; b < batch; b += ch) {
    const int nb = batch - b < ch ? batch - b : ch;
    rc = mult_front(c, level, nb, shift(view(a0), b), shift(view(a1), b), shift(view(b0), b), shift(view(b1), b),
                    pre_rescale, ws.p, shift(view(d0), b), shift(view(d1), b), shift(view(d2), b), false, st);
    if (rc) return rc;
  }
  POST();
  return 0;
}

// relinearize one chunk: d (dense [3][nb][L][N], NTT+Montgomery, destroyed) -> out
// dcoef: scratch of nb * L * N words for the coefficient form of d2 (may not alias d)
static int relin_chunk(tb200_ctx* c, int lvl, int nb, i64* d, i64* dcoef, const TbKskDev& key, TbView out0, TbView out1,
                       i64* ksws, tb200_stream st) {
  const int N = c->N, L = c->num_ord - lvl;
  const size_t pe = (size_t)nb * L * N;
  const bool reuse = c->fast && c->world == 1;
  if (reuse) {
    // NTT-domain d2 = forward transform of every group's extension at its own limbs: the fused core reads
    // those limbs from d2 itself, the rows outside it get them from k_fast_own_fill
    const int nf = ks_core_rows(c, lvl, L + c->K) < L ? ks_core_rows(c, lvl, L + c->K) : L;
    if (nf < L) {
      i64* ext = ksws + (size_t)nb * c->ks[lvl].state_rows * N;  // where keyswitch_chunk / ks_finish put it
      LAUNCH(k_fast_own_fill, dim3((unsigned)((N / 2 + 255) / 256), (unsigned)(L - nf), (unsigned)nb),
             dim3(N / 2 < 256 ? N / 2 : 256), st, c->dev(), c->devf(), (const TbKsLevel*)(c->d_ks + lvl),
             dense(d + 2 * pe, L, N), ext, lvl, N, L + c->K, nf);
    }
    // d2 back to coefficients, out of place (dcoef): its NTT form stays for the core; d0 / d1 stay in the NTT
    // domain and enter the key inner product
    int rc = fast_inverse_exit(c, dense(d + 2 * pe, L, N), dense(dcoef, L, N), L, nb, lvl, st);
    if (rc) return rc;
    const TbRelinExtra ex = {true, d, d + pe, d + 2 * pe};
    return keyswitch_chunk(c, lvl, nb, dense(dcoef, L, N), key, dense(d, L, N), dense(d + pe, L, N), out0, out1,
                           1, ksws, st, &ex);
  }
  int rc = c->fast ? fast_inverse_exit(c, dense(d, L, N), dense(d, L, N), L, 3 * nb, lvl, st)
                   : ntt_inverse(c, dense(d, L, N), dense(d, L, N), L, 3 * nb, lvl, 2, st);
  if (rc) return rc;
  return keyswitch_chunk(c, lvl, nb, dense(d + 2 * pe, L, N), key, dense(d, L, N), dense(d + pe, L, N), out0, out1, 1,
                         ksws, st);
}

extern "C" int tb200_relinearize(tb200_ctx* c, int level, int batch, const tb200_poly* d0, const tb200_poly* d1,
                                 const tb200_poly* d2, const tb200_ksk* evk, const tb200_poly* out0,
                                 const tb200_poly* out1, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  CHECK_POLY(d0);
  CHECK_POLY(d1);
  CHECK_POLY(d2);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  TbKskDev key;
  if ((rc = make_key(c, level, evk, &key))) return rc;
  SET_DEVICE(c->device);
  const int N = c->N, L = c->num_ord - level;
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve((4 * (size_t)L * N + ks_ws_elems(c, level)) * ch))) return rc;
  for (int b = 0; b < batch; b += ch) {
    const int nb = batch - b < ch ? batch - b : ch;
    const size_t pe = (size_t)nb * L * N;
    i64* d = ws.p;
    const tb200_poly* src[3] = {d0, d1, d2};
    for (int i = 0; i < 3; ++i) {  // copy (the reference mutates its triplet; we do not)
      TbPwArgs g;
      memset(&g, 0, sizeof(g));
      g.a = shift(view(src[i]), b);
      g.b = g.a;
      g.out = dense(d + i * pe, L, N);
      g.prime0 = level;
      g.N = N;
      launch_pw<15>(c, g, L, nb, st);
    }
    rc = relin_chunk(c, level, nb, d, d + 3 * pe, key, shift(view(out0), b), shift(view(out1), b), d + 4 * pe, st);
    if (rc) return rc;
  }
  POST();
  return 0;
}

extern "C" int tb200_cc_mult_relin(tb200_ctx* c, int level, int batch, const tb200_poly* a0, const tb200_poly* a1,
                                   const tb200_poly* b0, const tb200_poly* b1, const tb200_ksk* evk,
                                   const tb200_poly* out0, const tb200_poly* out1, int pre_rescale, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  if (pre_rescale && level + 1 >= c->num_ord) return fail(TB200_EINVAL, "cc_mult: no level left to rescale");
  CHECK_POLY(a0);
  CHECK_POLY(a1);
  CHECK_POLY(b0);
  CHECK_POLY(b1);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  const int lvl = level + (pre_rescale ? 1 : 0);
  TbKskDev key;
  if ((rc = make_key(c, lvl, evk, &key))) return rc;
  SET_DEVICE(c->device);
  const int N = c->N, L = c->num_ord - lvl;
  const int ch = batch < c->chunk ? batch : c->chunk;
  // workspace: x[4] | d[3] | key-switch
  WsLease ws(c, st);
  if ((rc = ws.reserve((7 * (size_t)L * N + ks_ws_elems(c, lvl)) * ch))) return rc;
  for (int b = 0; b < batch; b += ch) {
    const int nb = batch - b < ch ? batch - b : ch;
    const size_t pe = (size_t)nb * L * N;
    i64* x = ws.p;
    i64* d = x + 4 * pe;
    i64* ksws = d + 3 * pe;
    rc = mult_front(c, level, nb, shift(view(a0), b), shift(view(a1), b), shift(view(b0), b), shift(view(b1), b),
                    pre_rescale, x, dense(d, L, N), dense(d + pe, L, N), dense(d + 2 * pe, L, N), c->fast != 0, st);
    if (rc) return rc;
    rc = relin_chunk(c, lvl, nb, d, x, key, shift(view(out0), b), shift(view(out1), b), ksws, st);  // x is free again
    if (rc) return rc;
  }
  POST();
  return 0;
}

extern "C" int tb200_pc_mult(tb200_ctx* c, int level, int batch, const tb200_poly* pt, const tb200_poly* c0,
                             const tb200_poly* c1, const tb200_poly* out0, const tb200_poly* out1, int post_rescale,
                             tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  if (post_rescale && level + 1 >= c->num_ord) return fail(TB200_EINVAL, "pc_mult: no level left to rescale");
  CHECK_POLY(pt);
  CHECK_POLY(c0);
  CHECK_POLY(c1);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  SET_DEVICE(c->device);
  const int N = c->N, L = c->num_ord - level;
  const int ch = batch < c->chunk ? batch : c->chunk;
  WsLease ws(c, st);
  if ((rc = ws.reserve(2 * (size_t)ch * L * N))) return rc;
  for (int b = 0; b < batch; b += ch) {
    const int nb = batch - b < ch ? batch - b : ch;
    const size_t pe = (size_t)nb * L * N;
    const tb200_poly* in[2] = {c0, c1};
    const tb200_poly* out[2] = {out0, out1};
    for (int h = 0; h < 2; ++h) {
      TbView x = dense(ws.p + h * pe, L, N);
      rc = c->fast ? fast_forward_enter(c, shift(view(in[h]), b), x, L, nb, level, -1, st)
                   : ntt_forward(c, shift(view(in[h]), b), x, L, nb, level, true, st);
      if (rc) return rc;
      TbPwArgs g;
      memset(&g, 0, sizeof(g));
      TbView ptv = view(pt);
      if (ptv.bs != 0) ptv = shift(ptv, b);
      g.a = ptv;  // mont_mult(pt_, ct): pt is the first operand (ckks_engine.py:2569)
      g.b = x;
      g.out = x;
      g.prime0 = level;
      g.N = N;
      launch_pw<0>(c, g, L, nb, st);
      TbView dst = post_rescale ? x : shift(view(out[h]), b);
      rc = c->fast ? fast_inverse_exit(c, x, dst, L, nb, level, st) : ntt_inverse(c, x, dst, L, nb, level, 2, st);
      if (rc) return rc;
      if (post_rescale) rescale_impl(c, level, nb, x, shift(view(out[h]), b), 1, st);
    }
  }
  POST();
  return 0;
}

extern "C" int tb200_cc_addsub(tb200_ctx* c, int level, int batch, int sub, const tb200_poly* a0, const tb200_poly* a1,
                               const tb200_poly* b0, const tb200_poly* b1, const tb200_poly* out0,
                               const tb200_poly* out1, tb200_stream st) {
  int rc = check_level(c, level, batch);
  if (rc) return rc;
  REQUIRE_UNSHARDED();
  CHECK_POLY(a0);
  CHECK_POLY(a1);
  CHECK_POLY(b0);
  CHECK_POLY(b1);
  CHECK_POLY(out0);
  CHECK_POLY(out1);
  SET_DEVICE(c->device);
  const int L = c->num_ord - level;
  const tb200_poly* A[2] = {a0, a1};
  const tb200_poly* B[2] = {b0, b1};
  const tb200_poly* O[2] = {out0, out1};
  for (int h = 0; h < 2; ++h) {
    TbPwArgs g;
    memset(&g, 0, sizeof(g));
    g.a = view(A[h]);
    g.b = view(B[h]);
    g.out = view(O[h]);
    g.prime0 = level;
    g.N = c->N;
    if (sub)
      launch_pw<4>(c, g, L, batch, st);
    else
      launch_pw<3>(c, g, L, batch, st);
  }
  POST();
  return 0;
}

// ---- CSPRNG operators ---------------------------------------------------------------------------------
static int rng_table(TbRngTable* t, const uint64_t* host, int n, const char* what) {
  if (!host || n < 1 || n > TB_RNG_MAXQ) return fail(TB200_EINVAL, "%s: table must hold 1..%d words", what, TB_RNG_MAXQ);
  memset(t, 0, sizeof(*t));
  for (int i = 0; i < n; ++i) t->v[i] = host[i];
  return 0;
}
#define RNG_ENTER(ptr, cnt, what)                                                                 \
  if (!(ptr) || (cnt) < 1) return fail(TB200_EINVAL, what ": null buffer or empty");               \
  if (((uintptr_t)(ptr)&15) != 0) return fail(TB200_EINVAL, what ": buffers must be 16-byte aligned"); \
  SET_DEVICE(device);

extern "C" int tb200_chacha20(int device, int64_t* states, int64_t n_rows, int64_t* out, int64_t step, tb200_stream st) {
  RNG_ENTER(states, n_rows, "chacha20");
  if (!out || ((uintptr_t)out & 15) != 0) return fail(TB200_EINVAL, "chacha20: bad output buffer");
  LAUNCH(k_rng_chacha20, dim3((unsigned)((n_rows + 127) / 128)), dim3(128), st, (i64*)states, (i64*)out, (long)n_rows,
         (i64)step);
  POST();
  return 0;
}
extern "C" int tb200_randint_fast(int device, int64_t* states, int channels, int64_t L, const uint64_t* q_host,
                                  int64_t shift, int64_t step, int64_t* out, tb200_stream st) {
  RNG_ENTER(states, L, "randint_fast");
  if (!out || ((uintptr_t)out & 15) != 0) return fail(TB200_EINVAL, "randint_fast: bad output buffer");
  TbRngTable q;
  int rc = rng_table(&q, q_host, channels, "randint_fast");
  if (rc) return rc;
  LAUNCH(k_rng_randint_fast, dim3((unsigned)((L + 127) / 128), (unsigned)channels), dim3(128), st, (i64*)states,
         (i64*)out, (long)L, q, (i64)shift, (i64)step);
  POST();
  return 0;
}
extern "C" int tb200_discrete_gaussian_fast(int device, int64_t* states, int64_t n_rows, const uint64_t* lut_host,
                                            int size, int depth, int64_t step, int64_t* out, tb200_stream st) {
  RNG_ENTER(states, n_rows, "discrete_gaussian_fast");
  if (!out || ((uintptr_t)out & 15) != 0) return fail(TB200_EINVAL, "discrete_gaussian_fast: bad output buffer");
  if (depth < 1 || depth > 6 || size != (1 << depth) - 1) return fail(TB200_EINVAL, "discrete_gaussian_fast: bad tree");
  TbRngTable lut;
  int rc = rng_table(&lut, lut_host, 2 * size, "discrete_gaussian_fast");
  if (rc) return rc;
  LAUNCH(k_rng_gaussian_fast, dim3((unsigned)((n_rows + 127) / 128)), dim3(128), st, (i64*)states, (i64*)out,
         (long)n_rows, lut, size, depth, (i64)step);
  POST();
  return 0;
}
extern "C" int tb200_randint(int device, int64_t* words, int channels, int64_t L, const uint64_t* q_host,
                             tb200_stream st) {
  RNG_ENTER(words, L, "randint");
  TbRngTable q;
  int rc = rng_table(&q, q_host, channels, "randint");
  if (rc) return rc;
  LAUNCH(k_rng_randint_inplace, dim3((unsigned)((L + 127) / 128), (unsigned)channels), dim3(128), st, (i64*)words, (long)L,
         q);
  POST();
  return 0;
}
extern "C" int tb200_discrete_gaussian(int device, int64_t* words, int64_t n_rows, const uint64_t* lut_host, int size,
                                       int depth, tb200_stream st) {
  RNG_ENTER(words, n_rows, "discrete_gaussian");
  if (depth < 1 || depth > 6 || size != (1 << depth) - 1) return fail(TB200_EINVAL, "discrete_gaussian: bad tree");
  TbRngTable lut;
  int rc = rng_table(&lut, lut_host, 2 * size, "discrete_gaussian");
  if (rc) return rc;
  LAUNCH(k_rng_gaussian_inplace, dim3((unsigned)((n_rows + 127) / 128)), dim3(128), st, (i64*)words, (long)n_rows, lut,
         size, depth);
  POST();
  return 0;
}
extern "C" int tb200_randround(int device, const double* coef, int64_t* words, int64_t n, tb200_stream st) {
  if (!coef || !words || n < 1) return fail(TB200_EINVAL, "randround: null buffer or empty");
  SET_DEVICE(device);
  LAUNCH(k_rng_randround, dim3((unsigned)((n + 255) / 256)), dim3(256), st, coef, (i64*)words, (long)n);
  POST();
  return 0;
}
