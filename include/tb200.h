/*
 * tb200.h -- C ABI of libtb200.so: a B200 (sm_100a) backend for the CKKS RNS polynomial hot path
 * behind tiberate-fhe's CkksEngine.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer named `*_dev` / of type
 * tb200_poly is DEVICE memory of the context's GPU holding int64 residues in the reference's
 * layout: a polynomial is int64 [limbs, N] row-major, row i = residues modulo
 * prime `prime0 + i` in CkksConfig.q order [scale primes..., base prime, special primes...].
 * All launches go to the caller's stream; functions return 0 on success, a negative TB200_E*
 * code on bad arguments, a positive cudaError_t on a CUDA failure (tb200_last_error() has text).
 *
 * Two layers:
 *  (1) op layer  -- one entry per operator of the reference's torch op libraries
 *      (csrc/ops/mont.cpp:138-170, mont_extra.cpp:67-88, ntt_radix2.cpp:29-41,
 *      intt_radix2.cpp:48-68, he_fused.cpp:80-110); `prime0` replaces the reference's
 *      "sp_prime_len + right-aligned constant pool" convention: prime0 = P - rows - sp_prime_len
 *      (SURVEY.md appendix A.0).  tiberate_fhe_b200/wrapper/ maps the reference signatures on it.
 *  (2) engine layer -- fused, batched entry points for the hot CkksEngine methods
 *      (tiberate/ckks_engine.py: rescale :1520, cc_mult :1640, relinearize :1695,
 *      create_switcher :1201, switch_key :1403, rotate_single :1804, pc_mult :2542,
 *      cc_add_double :1932), bit-exact with the op-by-op reference sequence.
 */
#ifndef TB200_H
#define TB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb200_ctx tb200_ctx;
typedef void* tb200_stream; /* cudaStream_t */

#define TB200_EINVAL (-1)   /* bad argument (shape / level / alignment) */
#define TB200_ENOMEM (-2)   /* workspace allocation failed */
#define TB200_ENODEV (-3)   /* no usable CUDA device */

#define TB200_MAX_GROUPS 32 /* digit groups of a key-switch key (logN17 preset has 13) */

/* A (batched) polynomial: element (b, row, j) lives at ptr[b*batch_stride + row*row_stride + j]. */
typedef struct {
  int64_t* ptr;
  int64_t batch_stride; /* in elements; ignored when batch == 1 */
  int64_t row_stride;   /* in elements; >= N */
} tb200_poly;

/* A key-switch key (KeySwitchKey.data, ckks_engine.py:822-860): per GLOBAL digit-group id g the
 * pair (b_g, a_g), each int64 [P, N] NTT+Montgomery at level 0. Unused ids may be NULL. */
typedef struct {
  int32_t num_groups;
  int32_t reserved;
  int64_t row_stride;
  const int64_t* b[TB200_MAX_GROUPS];
  const int64_t* a[TB200_MAX_GROUPS];
} tb200_ksk;

/* ---- context (replaces NTTContext + the __constant__ pool: ntt_context.py:151-361,
 *      constant_mem_context.py:126-295; no 64-prime cap, not process-global) ------------------- */
tb200_ctx* tb200_ctx_create(int device, int logN, int num_primes, int num_special,
                            const int64_t* q /*[num_primes]*/, int scale_bits);
/* RNS-limb sharding (BASELINE.json configs[3]): the context of `rank` holds the ordinary
 * primes of the digit groups that rank owns (rns_partition.py:34-52: scale group g -> rank (np-1-g) mod
 * world, base prime -> rank 0) followed by the replicated special primes; tensors passed to it hold the
 * corresponding LOCAL rows.  Such a context supports the op layer and tb200_ks_digits / tb200_ks_finish. */
tb200_ctx* tb200_ctx_create_sharded(int device, int logN, int num_primes, int num_special, const int64_t* q,
                                    int scale_bits, int rank, int world);
/* global prime index of every local row [num_local_primes] */
int tb200_ctx_local_primes(const tb200_ctx*, int32_t* out);
void tb200_ctx_destroy(tb200_ctx*);
const char* tb200_last_error(void);
const char* tb200_version(void);
/* host copies of derived constants (for tests and for the Python context mirror) */
int tb200_ctx_get_prime_consts(const tb200_ctx*, int64_t* out /*[num_primes][8]*/);
int tb200_ctx_get_twiddles(const tb200_ctx*, int inverse, int prime, int64_t* out /*[N], lazy Montgomery form*/);
int tb200_ctx_info(const tb200_ctx*, int32_t* out /*[8]: logN,N,P,K,LA,LB,device,num_groups*/);
/* max ciphertexts processed per internal pass by the engine layer (workspace = chunk * ~730 limb rows) */
int tb200_ctx_set_chunk(tb200_ctx*, int chunk);
/* 1 (default): the fused engine calls run their internal transforms on the mod-q path (error-free FP64
 * products for primes below 2^42, Harvey/Shoup integer butterflies otherwise, fused ModUp prologue, key
 * inner product without per-term reductions); 0: they are composed from the exact op-layer kernels with
 * the reference's own lazy Montgomery butterflies.  Outputs are bit-identical either way (every internal
 * chain ends in a canonicalising step); the switch exists for A/B tests and measurements.  The mod-q path
 * expects key-switching key residues below 2^51 in magnitude on primes below 2^42 and below 2^62 on larger
 * primes (whose 128-bit key sums then hold 16 digit groups; with residues below 2^61, all 32) -- the
 * reference's keys are lazy Montgomery residues in (-2q, 2q). */
int tb200_ctx_set_fast(tb200_ctx*, int on);
/* Mod-q path only: share (in eighths, 0..8; default 8 = all) of the limbs of primes below 2^42 whose
 * arithmetic runs on the FP64 pipe (exact-integer doubles, 6 DFMA-class instructions per modular product)
 * while the remaining limbs use the integer pipes; results are bit-identical for every share. */
int tb200_ctx_set_f64_share(tb200_ctx*, int eighths);
/* Mod-q path only: scheduling knobs for A/B measurements; results are bit-identical for every setting.
 *   TB200_TUNE_FUSED_CORE (default 1): FP64 limbs run pass B of every digit group, the key inner product
 *     and inverse pass B as one kernel (no HBM round trip of the transformed extensions); 0: three kernels.
 *   TB200_TUNE_SIDE_ROWS (bit mask): the launches over the 60-bit limb rows (integer pipes) are forked onto a
 *     library-owned stream and joined again, beside the launches over the FP64 limb rows.  Bit 0: inside a key
 *     switch (measured no gain: the first kernel fills the GPU); bit 1: the input transforms of cc_mult, whose
 *     60-bit launches are a third of a wave each; 0: one stream.
 *   TB200_TUNE_FUSED_MODDOWN (default 0: measured 63 us against 36 + 20 us separately, B200 logN16): ModDown and the relinearisation / switch-key tail run inside the
 *     exit of inverse pass A of the ordinary limbs (after the special limbs were transformed and
 *     chain-reduced); 0: separate kernels over the coefficient-domain sums.
 *   TB200_TUNE_STREAM_WS (default 1): the scratch of an engine call is leased from the device's stream-ordered
 *     memory pool on the caller's stream (see "Streams, threads and CUDA graphs" below); 0: one grow-only
 *     workspace per context (single stream, growth synchronises the device).
 *   TB200_TUNE_FUSED_TENSOR (default 1): in cc_mult the forward pass B of the last operand and the tensor product
 *     run as one kernel on the FP64 limbs (the transformed operand never goes to HBM); 0: separate kernels.
 *   TB200_TUNE_SUM_NTT (default 1): tb200_ntt runs the deferred-reduction kernels (the reference's lazy
 *     representatives from one reduction modulo 2q per pass) on tiles of 40-bit limbs whose inputs are lazy
 *     values in [0, 2q), the generic exact kernels on the rest; 0: generic kernels only. */
enum tb200_tuning { TB200_TUNE_FUSED_CORE = 0, TB200_TUNE_SIDE_ROWS = 1, TB200_TUNE_FUSED_MODDOWN = 2,
                    TB200_TUNE_STREAM_WS = 3, TB200_TUNE_SUM_NTT = 4, TB200_TUNE_FUSED_TENSOR = 5 };
int tb200_ctx_set_tuning(tb200_ctx*, int knob, int value);

/* ---- packed wire format ------------------------------------------------------------------------------
 * Canonical residues of the scale primes (all within 2^32 of 2^40, i.e. below 2^41) as 41 bits each: a limb is
 * 5 N bytes (bits 0..39 of residue j at byte 5 j, little endian) followed by N / 8 bytes (bit 40 of residue j = bit
 * j % 8 of byte j / 8) -- 5.125 bytes per residue instead of the 8 of the int64 layout: what crosses the host <->
 * device link (-36 % bytes on those limbs).  The `rows` limbs starting at prime index prime0 must all be below 2^41
 * (wider limbs travel as int64).  Packed buffers: device memory, 8-byte aligned, strides in BYTES (multiples of 8,
 * row pitch >= 5 N + N / 8), N >= 64.  tb200_pack41 expects canonical inputs (every engine call returns canonical
 * residues); the reference has no equivalent (its ciphertexts leave the device only through pickle,
 * tiberate/typing.py:283-290). */
int tb200_unpack41(tb200_ctx*, int rows, int batch, int prime0, const uint8_t* src, int64_t src_batch_stride,
                   int64_t src_row_stride, const tb200_poly* dst, tb200_stream);
int tb200_pack41(tb200_ctx*, int rows, int batch, int prime0, const tb200_poly* src, uint8_t* dst,
                 int64_t dst_batch_stride, int64_t dst_row_stride, tb200_stream);

/* ---- op layer: pointwise Montgomery family (mont_cuda.cu, mont_extra_cuda.cu) ---------------- */
enum tb200_pw_op {
  TB200_MONT_MULT = 0,          /* out = MM(a, b)                      mont_cuda.cu:11-36    */
  TB200_MONT_ADD = 1,           /* out = CS2(a + b)                    :453-476              */
  TB200_MONT_SUB = 2,           /* out = CS2(a - b)                    :577-600              */
  TB200_MONT_ADD_REDUCE_2Q = 3, /* out = CS1(CS2(a + b))               mont_extra_cuda.cu:161 */
  TB200_MONT_SUB_REDUCE_2Q = 4, /* out = CS1(CS2(a - b))               :230                  */
  TB200_MONT_ENTER_SCALAR = 5,  /* out = MM(a, scal[row])              mont_cuda.cu:85-109   */
  TB200_MONT_ENTER_RS = 6,      /* out = MM(a, R^2)                    :151-173              */
  TB200_MONT_ENTER_RS_SCALE = 7,/* out = MM(a, R^2 2^scale_bits)       :211-234              */
  TB200_MONT_REDUCE = 8,        /* out = MR(a)                         :345-364              */
  TB200_REDUCE_2Q = 9,          /* out = CS1(a)                        :402-422              */
  TB200_MAKE_SIGNED = 10,       /* out = a <= q/2 ? a : a - q          :642-661              */
  TB200_MAKE_UNSIGNED = 11,     /* out = a + q                         :693-712              */
  TB200_TILE_UNSIGNED = 12,     /* out[r][j] = a[j] + q_r              :744-760              */
  TB200_PC_ADD_FUSED = 13,      /* out = CS1(MR(CS2(MM(a,R^2) + b)))   he_fused_cuda.cu:12-51 */
  TB200_MONT_ENTER_SCALAR_REDUCE_2Q = 14 /* out = CS1(MM(a, scal[row])) mont_extra_cuda.cu:299 */
};
/* Generic launcher.  b / scal may be NULL when the op does not use them.  When `explicit_consts`
 * is non-NULL it points to DEVICE int64 arrays {ql, qh, kl, kh}[rows] (legacy ops that receive
 * their constants as tensors: mont_enter :272-339, mont_add_legacy :519-571, tile_unsigned);
 * entries kl/kh may be NULL for ops that only need q. Otherwise row r uses prime prime0 + r. */
typedef struct {
  const int64_t* ql;
  const int64_t* qh;
  const int64_t* kl;
  const int64_t* kh;
  const int64_t* two_q; /* alternative to ql/qh: 2q per row */
} tb200_explicit_consts;
int tb200_pointwise(tb200_ctx*, int op, int rows, int batch, int prime0, const tb200_poly* a,
                    const tb200_poly* b, const int64_t* scal_dev, const tb200_explicit_consts* ec,
                    const tb200_poly* out, tb200_stream);
/* out[c][n] = fold_k CS2(acc + in[k][c][n]); pairwise != 0 selects mont_add_many_3d's pair-first
 * order (mont_extra_cuda.cu:12-42), 0 mont_reduce_add_many_3d (:83-118). in: [K][rows][N] dense. */
int tb200_add_many(tb200_ctx*, int pairwise, int K, int rows, int prime0, const int64_t* in_dev,
                   int64_t* out_dev, tb200_stream);

/* ---- op layer: NTT (ntt_radix2_cuda.cu, intt_radix2_cuda.cu, mont_used_in_ntt.cuh) ------------ */
/* forward, in place on `rows` rows: enter != 0 -> enter_ntt_radix2 (MM by R^2 first, :98-136),
 * else ntt_radix2 (:49-96).  Outputs are the reference's lazy [0, 2q) representatives bit for bit (for
 * primes below 2^42 and |x| < 2^50 they are computed on the FP64 pipe, other tiles on the integer one). */
int tb200_ntt(tb200_ctx*, int rows, int batch, int prime0, const tb200_poly* a, int enter, tb200_stream);
/* inverse, in place: mode 0 intt_radix2 (x N^-1, stays Montgomery), 1 _exit (+MR), 2 _exit_reduce
 * (+CS1, canonical), 3 _exit_reduce_signed (+centre). intt_radix2_cuda.cu:51-278.
 * Mode 2 returns canonical residues, which do not depend on the lazy representatives in between, and runs
 * on the mod-q kernels while tb200_ctx_set_fast is on (the default).  Input domain of that route:
 * |x| < 2^51 on primes below 2^42, (-2q, 2q) on larger primes -- every value the operators produce;
 * tb200_ctx_set_fast(ctx, 0) selects the reference's own butterflies for anything wider. */
int tb200_intt(tb200_ctx*, int rows, int batch, int prime0, const tb200_poly* a, int mode, tb200_stream);

/* ---- op layer: fused HE kernels (he_fused_cuda.cu) ------------------------------------------- */
/* :99-270. In place on a [rows][N]; row r uses prime prime0 + r; scales_dev[rows]; rescaler_dev[N]. */
int tb200_rescale_rows(tb200_ctx*, int rows, int prime0, const tb200_poly* a, const int64_t* scales_dev,
                       const int64_t* rescaler_dev, int64_t round_at, int exact, tb200_stream);
/* :276-355. state [alpha][N] -> out [rows][N], row r = prime prime0 + r;
 * l_enter_dev [alpha-1][l_enter_stride], read at column l_enter_offset + r. */
int tb200_extend(tb200_ctx*, int rows, int prime0, int alpha, const int64_t* state_dev, int64_t state_stride,
                 const int64_t* l_enter_dev, int64_t l_enter_stride, int64_t l_enter_offset,
                 int64_t* out_dev, int64_t out_stride, tb200_stream);
/* :361-427. out[r][perm[j] % N] = CS1(+-a[r][j] + q_r); perm_dev int64 [N]; two_q_dev [rows]. */
int tb200_codec_rotate(tb200_ctx*, int rows, const tb200_poly* a, const int64_t* perm_dev,
                       const int64_t* two_q_dev, const tb200_poly* out, tb200_stream);
/* :433-584. c [rows][N] (ordinary limbs of the level), p [K][N] (special limbs, modified in place by
 * the chain-backward step exactly as the reference does); out [rows][N]. Constants from the context. */
int tb200_divide_by_p(tb200_ctx*, int level, const tb200_poly* c, const tb200_poly* p, const tb200_poly* out,
                      tb200_stream);

/* ---- engine layer (batched; `level` is the level of the INPUT ciphertexts) --------------------- */
/* rescale :1520-1618.  in: [L+1][N] at `level`, out: [L][N] at level+1 (fresh buffers, not views). */
int tb200_rescale(tb200_ctx*, int level, int batch, const tb200_poly* in0, const tb200_poly* in1,
                  const tb200_poly* out0, const tb200_poly* out1, int exact, tb200_stream);
/* create_switcher :1201-1363 on a coefficient-domain canonical polynomial a [L][N] at `level`. */
int tb200_keyswitch(tb200_ctx*, int level, int batch, const tb200_poly* a, const tb200_ksk* ksk,
                    const tb200_poly* out0, const tb200_poly* out1, tb200_stream);
/* Key switch in two halves, the seam where limb-sharded ranks exchange the ModUp digits:
 *   tb200_ks_state_info: out[4] = {rows of the digit-state buffer, first row of this rank's segment,
 *                        rows per segment, local ordinary rows L at this level};
 *   tb200_ks_digits: mixed-radix digits (ckks_engine.py:889-921) of the digit groups this rank owns,
 *                    written into its segment of `state` [state_rows][N] (a: local rows, canonical);
 *   -- sharded: ONE in-place all-gather of the segments (ncclAllGather over NVLink) --
 *   tb200_ks_finish: extend to the local ordinary + special limbs, NTT, key inner product with the
 *                    local key rows, inverse NTT, ModDown; tail 0: out = (ks0, ks1); 1: out = CS1(add +
 *                    ks) for both (relinearize); 2: out0 = CS1(CS2(add0 + ks0)), out1 = ks1 (switch_key).
 * On an unsharded context digits + finish == tb200_keyswitch. */
int tb200_ks_state_info(const tb200_ctx*, int level, int32_t* out);
int tb200_ks_digits(tb200_ctx*, int level, int batch, const tb200_poly* a, const tb200_poly* state, tb200_stream);
int tb200_ks_finish(tb200_ctx*, int level, int batch, const tb200_poly* state, const tb200_ksk* ksk,
                    const tb200_poly* add0, const tb200_poly* add1, const tb200_poly* out0, const tb200_poly* out1,
                    int tail, tb200_stream);
/* tb200_ks_finish in two calls, so that the ModUp of the digit groups this rank owns runs while the all-gather
 * of the other ranks' digits is still in flight (tiberate_fhe_b200/dist.py): tb200_ks_modup(which = 1 own /
 * 2 foreign / 0 all; + 4: see tb200_ks_core_sp) extends + forward-transforms (pass A) the selected groups into the context workspace,
 * tb200_ks_core does the rest.  batch <= chunk (the extended limbs live in the workspace between the calls). */
int tb200_ks_modup(tb200_ctx*, int level, int batch, const tb200_poly* state, int which, tb200_stream);
int tb200_ks_core(tb200_ctx*, int level, int batch, const tb200_poly* state, const tb200_ksk* ksk,
                  const tb200_poly* add0, const tb200_poly* add1, const tb200_poly* out0, const tb200_poly* out1,
                  int tail, tb200_stream);
/* Limb-sharded key switch with the special limbs sharded as well.  tb200_ks_finish / tb200_ks_core replicate the
 * K special limbs on every rank, as the reference does (tiberate/context/rns_partition.py:34-52: every device
 * holds the special group; ckks_engine.py:1365-1400 extends and multiplies all of them per device).  Here a rank
 * computes the key sums of ITS SHARE of the special limbs only (ceil(K / world) limbs per rank, rank order), and one
 * more all-gather of 2 K N words per polynomial -- issued by the caller while the ordinary limbs are processed --
 * completes them before ModDown:
 *   tb200_ks_modup(which + 4)  extend + pass A of the local ordinary rows and of the share of the special rows
 *   tb200_ks_core_sp           pass B, key inner product, inverse transform of the share -> the rank's segment of sp
 *   (all-gather of the segments)
 *   tb200_ks_core_ord          the same for the local ordinary rows (the sums stay in the context workspace)
 *   tb200_ks_moddown           chain-backward on the complete special limbs, ModDown, tail (as tb200_ks_finish)
 * sp: dense int64 [rows][batch][2][N] (special limb, polynomial, key half) with rows = world * ceil(K / world);
 * tb200_ks_sp_info: out[0] = rows, out[1] = rows per segment, out[2], out[3] = [s0, s1) the special limbs of this rank.
 * batch <= chunk.  The outputs are the same bits as tb200_ks_finish. */
int tb200_ks_sp_info(const tb200_ctx*, int32_t* out);
int tb200_ks_core_sp(tb200_ctx*, int level, int batch, const tb200_ksk* ksk, int64_t* sp, tb200_stream);
int tb200_ks_core_ord(tb200_ctx*, int level, int batch, const tb200_ksk* ksk, tb200_stream);
int tb200_ks_moddown(tb200_ctx*, int level, int batch, int64_t* sp, const tb200_poly* add0, const tb200_poly* add1,
                     const tb200_poly* out0, const tb200_poly* out1, int tail, tb200_stream);
/* cc_mult (+relinearize) :1640-1732.  a*, b*: [L_in][N] at `level`; with pre_rescale the product
 * lives at level+1 and has L_in-1 rows.  out0/out1 canonical coefficient domain. */
int tb200_cc_mult_relin(tb200_ctx*, int level, int batch, const tb200_poly* a0, const tb200_poly* a1,
                        const tb200_poly* b0, const tb200_poly* b1, const tb200_ksk* evk,
                        const tb200_poly* out0, const tb200_poly* out1, int pre_rescale, tb200_stream);
/* cc_mult(post_relin=False): the NTT+Montgomery triplet d0,d1,d2 (:1664-1687). */
int tb200_cc_mult_triplet(tb200_ctx*, int level, int batch, const tb200_poly* a0, const tb200_poly* a1,
                          const tb200_poly* b0, const tb200_poly* b1, const tb200_poly* d0,
                          const tb200_poly* d1, const tb200_poly* d2, int pre_rescale, tb200_stream);
/* relinearize :1695-1732 on an NTT+Montgomery triplet (inputs are NOT modified). */
int tb200_relinearize(tb200_ctx*, int level, int batch, const tb200_poly* d0, const tb200_poly* d1,
                      const tb200_poly* d2, const tb200_ksk* evk, const tb200_poly* out0,
                      const tb200_poly* out1, tb200_stream);
/* rotate_single :1804-1840: Galois automorphism X -> X^galois (galois = 3^delta mod 2N, odd) on both
 * polynomials, then switch_key with rotk.  With ksk == NULL only the automorphism is applied. */
int tb200_rotate(tb200_ctx*, int level, int batch, int64_t galois, const tb200_poly* c0, const tb200_poly* c1,
                 const tb200_ksk* rotk, const tb200_poly* out0, const tb200_poly* out1, tb200_stream);
/* Hoisted rotations (extension; the reference rotates one key at a time): `nrot` rotations of the same
 * ciphertext(s) share the ModUp digits, the extension and its forward transform; per rotation only the key inner
 * product (reading the extension through the NTT-domain automorphism), the inverse transform and ModDown remain.
 * galois[r] / rotks[r]: Galois element and rotation key of rotation r; rotation r is written at
 * out0->ptr + r * rot_stride (same batch / row strides for every rotation).  The results decrypt like
 * tb200_rotate's but are not bit-identical to them (the digits are taken before the automorphism);
 * oracle/engine.py: rotate_hoisted is the bit-exact restatement. */
int tb200_rotate_hoisted(tb200_ctx*, int level, int batch, int nrot, const int64_t* galois, const tb200_poly* c0,
                         const tb200_poly* c1, const tb200_ksk* const* rotks, const tb200_poly* out0,
                         const tb200_poly* out1, int64_t rot_stride, tb200_stream);
/* switch_key :1403-1420 (ct.c0 + ks0, ks1). */
int tb200_switch_key(tb200_ctx*, int level, int batch, const tb200_poly* c0, const tb200_poly* c1,
                     const tb200_ksk* ksk, const tb200_poly* out0, const tb200_poly* out1, tb200_stream);
/* pc_mult :2542-2580 with the cached NTT+Montgomery plaintext pt [L][N] (batch stride 0 broadcasts it). */
int tb200_pc_mult(tb200_ctx*, int level, int batch, const tb200_poly* pt, const tb200_poly* c0,
                  const tb200_poly* c1, const tb200_poly* out0, const tb200_poly* out1, int post_rescale,
                  tb200_stream);
/* cc_add_double :1932-1956 / cc_sub (mont_sub_reduce_2q) */
int tb200_cc_addsub(tb200_ctx*, int level, int batch, int sub, const tb200_poly* a0, const tb200_poly* a1,
                    const tb200_poly* b0, const tb200_poly* b1, const tb200_poly* out0, const tb200_poly* out1,
                    tb200_stream);

/* ---- CSPRNG operators (torch namespace tiberate_csprng_ops) ----------------------
 * State rows: 16 int64 per ChaCha20 block, one 32-bit word each -- constants, key, 64-bit counter in
 * words 12-13, nonce (tiberate/rng/csprng/csprng.py:113-178).  All buffers are device pointers of
 * contiguous int64 on `device`; q / lut are HOST pointers (the reference passes numpy addresses,
 * csprng.py:249-252, discrete_gaussian_sampler.py:104-106) of at most 128 words. */
/* chacha20 (csrc/csprng/chacha20.cpp:11-32): out[n][16] = blocks of the current states; every row's
 * counter then advances by `step`. */
int tb200_chacha20(int device, int64_t* states, int64_t n_rows, int64_t* out, int64_t step, tb200_stream);
/* randint_fast (randint.cpp:23-35, cuda/randint_cuda.cu:23-87): states [channels][L][16] ->
 * out [channels][4L] = floor(X * q[c] / 2^128) + shift from 128 random bits each; states stepped. */
int tb200_randint_fast(int device, int64_t* states, int channels, int64_t L, const uint64_t* q_host,
                       int64_t shift, int64_t step, int64_t* out, tb200_stream);
/* discrete_gaussian_fast (discrete_gaussian.cpp, cuda/discrete_gaussian_cuda.cu:19-101): states [n][16]
 * -> out [4n]; lut_host = lows[size] then highs[size] of the CDT binary search tree. */
int tb200_discrete_gaussian_fast(int device, int64_t* states, int64_t n_rows, const uint64_t* lut_host, int size,
                                 int depth, int64_t step, int64_t* out, tb200_stream);
/* randint / discrete_gaussian (the two-step variants): in place on random words [channels][L][16] /
 * [n][16]; the sample of words 4j..4j+3 replaces word 4j. */
int tb200_randint(int device, int64_t* words, int channels, int64_t L, const uint64_t* q_host, tb200_stream);
int tb200_discrete_gaussian(int device, int64_t* words, int64_t n_rows, const uint64_t* lut_host, int size, int depth,
                            tb200_stream);
/* randround (randround.cpp:10-19, cuda/randround_cuda.cu:4-36): words[i] <- sign(c_i) * (floor|c_i| +
 * [words[i] < rn(frac(|c_i|) * 2^32)]), words[i] a 32-bit random word. */
int tb200_randround(int device, const double* coef, int64_t* words, int64_t n, tb200_stream);

/* number of kernel launches issued by this library since process start (bench.py gpu_launches) */
int64_t tb200_launch_count(void);
/* per-kernel timing for bench.py's roofline: while enabled every launch is bracketed by CUDA events;
 * collect() synchronises and writes "kernel\tlaunches\ttotal_ms\n" lines (returns bytes written). */
void tb200_prof_enable(int on);
int tb200_prof_collect(char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* TB200_H */
