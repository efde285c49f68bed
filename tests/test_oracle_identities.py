"""Big-integer identities that pin the oracle from first principles (SURVEY.md 8c pin 1)."""

import ctypes

import numpy as np
import pytest

from oracle import call, lib
from oracle.context import PRESETS, OracleContext, R, bit_reverse, toy_primes
from oracle.engine import OracleEngine


@pytest.mark.parametrize("q", [PRESETS[16]["q"][0], PRESETS[16]["q"][-1], PRESETS[14]["q"][-2]])
def test_montgomery_halves_equal_closed_form_and_bigint(q):
    L = lib()
    k = (R * pow(R, -1, q) - 1) // q
    rng = np.random.default_rng(q % 1000)
    n = 400_000
    a = rng.integers(-2 * q, 2 * q, size=n, dtype=np.int64)
    b = rng.integers(-q // 2, 2 * q, size=n, dtype=np.int64)
    a[:6] = [0, 1, -1, q, 2 * q - 1, -(2 * q - 1)]
    b[:6] = [0, q - 1, 2 * q - 1, q, 1, 2 * q - 1]
    o1, o2 = np.empty_like(a), np.empty_like(a)
    call("orc_vec_mm", o1, a, b, ctypes.c_size_t(n), ctypes.c_int64(q), ctypes.c_int64(k), ctypes.c_int(0))
    call("orc_vec_mm", o2, a, b, ctypes.c_size_t(n), ctypes.c_int64(q), ctypes.c_int64(k), ctypes.c_int(1))
    assert np.array_equal(o1, o2)
    for x, y, o in list(zip(a.tolist(), b.tolist(), o1.tolist()))[:2000]:
        s = (x * y * k) % R
        assert o == (x * y + s * q) >> 62
        assert (o * R - x * y) % q == 0
    call("orc_vec_mr", o1, a, ctypes.c_size_t(n), ctypes.c_int64(q), ctypes.c_int64(k), ctypes.c_int(0))
    call("orc_vec_mr", o2, a, ctypes.c_size_t(n), ctypes.c_int64(q), ctypes.c_int64(k), ctypes.c_int(1))
    assert np.array_equal(o1, o2)
    for x, o in list(zip(a.tolist(), o1.tolist()))[:2000]:
        assert o == (x + ((x * k) % R) * q) >> 62
    assert L.orc_mm_halves(5, 7, q, k) == L.orc_mm_closed(5, 7, q, k)


def test_ntt_is_evaluation_at_odd_powers_of_psi_and_inverts():
    ctx = OracleContext(6, toy_primes(6, 3, 2), 2)
    eng = OracleEngine(ctx)
    rng = np.random.default_rng(0)
    pr = list(range(ctx.P))
    a = eng.uniform(rng, pr)
    A = eng.enter_ntt(a, pr)
    for g in range(ctx.P):
        q, w = ctx.q[g], ctx.psi_root[g]
        assert pow(w, ctx.N, q) == q - 1  # primitive 2N-th root
        for i in range(ctx.N):
            e = pow(w, 2 * bit_reverse(i, ctx.logN) + 1, q)
            v = sum(int(a[g, j]) * pow(e, j, q) for j in range(ctx.N)) % q
            assert (int(A[g, i]) - v * R) % q == 0 and 0 <= A[g, i] < 2 * q
    assert np.array_equal(eng.intt(A, pr, 2), a)
    signed = eng.intt(A, pr, 3)
    assert np.array_equal(signed % ctx.qa[:, None], a)


def test_preset_psi_roots_match_survey_appendix_d():
    want = {14: [81696219706, 641000223749548346, 77965612450023209],
            15: [993612494692, 1100972123716672435, 741224627014235163],
            16: [58415410147, 987813353222176621, 1050720516549580945]}
    from oracle.context import primitive_root_2N

    for logN, w in want.items():
        q, K = PRESETS[logN]["q"], PRESETS[logN]["K"]
        got = [primitive_root_2N(q[0], 1 << logN), primitive_root_2N(q[-K - 1], 1 << logN),
               primitive_root_2N(q[-1], 1 << logN)]
        assert got == w


def _negacyclic(a, b, N):
    out = [0] * N
    for i in range(N):
        for j in range(N):
            k = i + j
            if k < N:
                out[k] += int(a[i]) * int(b[j])
            else:
                out[k - N] -= int(a[i]) * int(b[j])
    return out


def test_keyswitch_mult_rotate_rescale_decrypt_correctly():
    """Semantic pin: with valid keys, relinearised products / rotations decrypt to the right
    polynomial (error = a few noise units), and rescale divides by the dropped prime."""
    logN = 6
    ctx = OracleContext(logN, toy_primes(logN, 7, 3), 3)
    eng = OracleEngine(ctx)
    N = ctx.N
    rng = np.random.default_rng(1)
    sk, _ = eng.gen_secret(rng)
    pk = eng.gen_public(rng, sk, True)
    evk = eng.gen_evk(rng, sk)
    scale = 1 << 40
    m1 = rng.integers(-50, 50, size=N)
    m2 = rng.integers(-50, 50, size=N)
    ct1 = eng.encrypt_poly(rng, m1 * scale, pk, 0)
    ct2 = eng.encrypt_poly(rng, m2 * scale, pk, 0)
    assert max(abs(x - int(y) * scale) for x, y in zip(eng.decrypt_poly(ct1, sk, 0), m1)) < 200
    res, lvl = eng.cc_mult(ct1, ct2, evk, 0, pre_rescale=False)
    exp = _negacyclic(m1, m2, N)
    d = eng.decrypt_poly(res, sk, lvl)
    assert max(abs(x / scale / scale - e) for x, e in zip(d, exp)) < 1e-6
    d = eng.decrypt_poly(eng.rescale(res, 0), sk, 1)
    assert max(abs(x * ctx.q[0] / scale / scale - e) for x, e in zip(d, exp)) < 1e-6
    for level in (0, 2):
        delta = 3
        rotk = eng.gen_rotk(rng, sk, delta)
        ct = eng.encrypt_poly(rng, m1 * scale, pk, level)
        r = eng.rotate_single(ct, rotk, delta, level)
        want = eng.rotate_plain((m1 * scale)[None, :], delta)[0]
        assert max(abs(x - int(e)) for x, e in zip(eng.decrypt_poly(r, sk, level), want)) < 500


def test_partition_matches_survey_table():
    from oracle.context import Partition

    p = Partition(35, 4)  # logN16: 34 scale + base, K = 4
    sizes = [len(pr) for _, pr in p.level_groups(0)]
    assert sizes == [4] * 8 + [2, 1]
    assert [len(pr) for _, pr in p.level_groups(3)] == [1] + [4] * 7 + [2, 1]
    p2 = Partition(17, 2, num_devices=2)  # rns_partition.py defaults
    assert p2.part_allocations[0][-2:] == [8, 9] and p2.part_allocations[1][-1] == 9
    assert sorted(p2.part_allocations[0][:-2] + p2.part_allocations[1][:-1]) == list(range(8))


def test_hoisted_rotations_decrypt_to_the_rotated_message():
    """rotate_hoisted (an extension beyond the reference, restated in oracle/engine.py) shares the ModUp between
    rotations; its outputs need not equal rotate_single's bits but must decrypt to the same rotated message."""
    logN = 6
    ctx = OracleContext(logN, toy_primes(logN, 7, 3), 3)
    eng = OracleEngine(ctx)
    N = ctx.N
    rng = np.random.default_rng(5)
    sk, _ = eng.gen_secret(rng)
    pk = eng.gen_public(rng, sk, True)
    scale = 1 << 40
    m = rng.integers(-50, 50, size=N)
    deltas = [1, 3, 7]
    rotks = {d: eng.gen_rotk(rng, sk, d) for d in deltas}
    for level in (0, 2):
        ct = eng.encrypt_poly(rng, m * scale, pk, level)
        outs = eng.rotate_hoisted(ct, rotks, deltas, level)
        for d, out in zip(deltas, outs):
            want = eng.rotate_plain((m * scale)[None, :], d)[0]
            got = eng.decrypt_poly(out, sk, level)
            assert max(abs(x - int(e)) for x, e in zip(got, want)) < 2000, (level, d)
            single = eng.decrypt_poly(eng.rotate_single(ct, rotks[d], d, level), sk, level)
            assert max(abs(x - y) for x, y in zip(got, single)) < 4000
