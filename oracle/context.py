"""Oracle restatement of the reference's prime-dependent context (TEST INFRASTRUCTURE ONLY).

Follows, with Python big ints:
  tiberate/context/mont_context.py:26-57      Montgomery constants (R = 2^62)
  tiberate/context/ntt_context.py:21-85       psi root choice, power series, bit reversal
  tiberate/context/ntt_context.py:277-298     psi_enter (twiddles put in Montgomery form *lazily*
                                              with mont_enter_Rs), Ninv, Rs_scale
  tiberate/context/rns_partition.py:7-186     digit groups ("partitions") per level / device
  tiberate/context/ntt_context.py:497-534     Y_scalar / L_scalar / L_enter
  tiberate/ckks_engine.py:114-143,185-256     rescale scales, PiR, mont_PR
"""

from __future__ import annotations

import math

import numpy as np

from . import _c, call, lib

R_BITS = 62
R = 1 << R_BITS

# Appendix D of SURVEY.md: prime chains produced by the reference's own generate_primes.py +
# CkksConfig defaults (scale_bits=40, 128-bit post-quantum / uniform).  tests/golden/ctx_*.json
# re-derives them through the reference code itself.
PRESETS = {
    14: dict(K=1, q=[1099510054913, 1099515691009, 1099508121601, 1099515789313, 1099507695617,
                     1099516280833, 1099506515969, 1152921504606748673, 1152921504606683137]),
    15: dict(K=2, q=[1099510054913, 1099515691009, 1099507695617, 1099516280833, 1099506515969,
                     1099520606209, 1099504549889, 1099523555329, 1099503894529, 1099527946241,
                     1099503370241, 1099529060353, 1099498258433, 1099531223041, 1099469684737,
                     1099532009473, 1152921504606584833, 1152921504598720513, 1152921504597016577]),
    16: dict(K=4, q=[1099510054913, 1099515691009, 1099507695617, 1099516870657, 1099506515969,
                     1099521458177, 1099503894529, 1099522375681, 1099490000897, 1099523555329,
                     1099489607681, 1099525128193, 1099486855169, 1099526176769, 1099484889089,
                     1099529060353, 1099480956929, 1099535220737, 1099469684737, 1099536138241,
                     1099468767233, 1099537580033, 1099461820417, 1099538104321, 1099457495041,
                     1099540725761, 1099455004673, 1099540856833, 1099454218241, 1099591974913,
                     1099453431809, 1099629723649, 1099451465729, 1099630510081, 1152921504606584833,
                     1152921504598720513, 1152921504597016577, 1152921504595968001,
                     1152921504592822273]),
}


def bit_reverse(x: int, nbits: int) -> int:
    r = 0
    for _ in range(nbits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def primitive_root_2N(q: int, N: int) -> int:
    """ntt_context.py:21-29: g = x^((q-1)/2N) for the smallest x >= 2 with g^N != 1."""
    e = (q - 1) // (2 * N)
    g = None
    for x in range(2, N):
        g = pow(x, e, q)
        if pow(g, N, q) != 1:
            break
    return g


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def find_ntt_primes(bits: int, N: int, count: int, skip=()) -> list[int]:
    """Small test helper (not reference code): `count` primes q = 1 mod 2N just below 2^bits."""
    out, c = [], ((1 << bits) - 1) // (2 * N)
    while len(out) < count:
        q = c * 2 * N + 1
        if q not in skip and is_prime(q):
            out.append(q)
        c -= 1
    return out


def toy_primes(logN: int, num_scales: int, K: int, scale_bits: int = 40, big_bits: int = 60) -> list[int]:
    """A small prime chain [scale primes..., base, special...] shaped like the reference's."""
    N = 1 << logN
    small = find_ntt_primes(scale_bits, N, num_scales)
    big = find_ntt_primes(big_bits, N, 1 + K)
    return small + big


class Partition:
    """rns_partition.py:7-186 restated.  Group g < num_partitions holds the scale primes
    [g*K, (g+1)*K); group num_partitions is the base prime; group num_partitions+1 the special
    primes.  Device d owns groups num_partitions-1-d, -d-num_devices, ... (ascending), device 0
    also the base group, every device the special group."""

    def __init__(self, num_ordinary: int, K: int, num_devices: int = 1):
        self.num_ordinary, self.K, self.num_devices = num_ordinary, K, num_devices
        ns = num_ordinary - 1
        self.num_partitions = -(-ns // K)
        self.groups = [list(range(i * K, min((i + 1) * K, ns))) for i in range(self.num_partitions)]
        self.groups.append([ns])
        self.groups.append(list(range(num_ordinary, num_ordinary + K)))
        self.part_allocations = []
        for d in range(num_devices):
            al = sorted(range(self.num_partitions - d - 1, -1, -num_devices))
            if d == 0:
                al.append(self.num_partitions)
            al.append(self.num_partitions + 1)
            self.part_allocations.append(al)
        # device-local prime lists at level 0 (with special): rnsPart.d_special
        self.d_special = [[p for g in al for p in self.groups[g]] for al in self.part_allocations]

    def dest_with_special(self, level: int, dev: int = 0) -> list[int]:
        return [p for p in self.d_special[dev] if p >= level]

    def start(self, level: int, dev: int = 0) -> int:
        """rnsPart.diff[level][dev]: rows dropped from the front of the device's level-0 tensor."""
        return len(self.d_special[dev]) - len(self.dest_with_special(level, dev))

    def level_groups(self, level: int, dev: int = 0):
        """[(global_group_id, [global prime ids still alive])] of the ORDINARY groups of `dev`."""
        out = []
        for g in self.part_allocations[dev][:-1]:
            alive = [p for p in self.groups[g] if p >= level]
            if alive:
                out.append((g, alive))
        return out

    def storage_order(self, level: int):
        """Global group ids in accumulation order (ckks_engine.py:175-183 stor_ids)."""
        ids = sorted(g for d in range(self.num_devices) for g, _ in self.level_groups(level, d))
        return ids


class OracleContext:
    def __init__(self, logN: int, q: list[int], K: int, scale_bits: int = 40, num_devices: int = 1):
        self.logN, self.N = logN, 1 << logN
        self.q = [int(x) for x in q]
        self.P, self.K = len(q), K
        self.num_ordinary = self.P - K
        self.num_scales = self.num_ordinary - 1
        self.num_levels = self.num_scales  # CkksEngine.num_levels (ckks_engine.py:102-104)
        self.scale_bits = scale_bits
        N = self.N
        for qi in self.q:
            assert (qi - 1) % (2 * N) == 0 and 4 * qi < R, qi

        # mont_context.py:26-57
        self.Rinv = [pow(R, -1, qi) for qi in self.q]
        self.k = [(R * ri - 1) // qi for ri, qi in zip(self.Rinv, self.q)]
        self.Rs = [R * R % qi for qi in self.q]
        self.Rs_scale = [rs * (1 << scale_bits) % qi for rs, qi in zip(self.Rs, self.q)]
        self.Ninv = [pow(N, -1, qi) * R % qi for qi in self.q]  # ntt_context.py:283-287

        self.qa, self.ka = _c(self.q), _c(self.k)
        self.Rsa, self.Ninva = _c(self.Rs), _c(self.Ninv)

        # ntt_context.py:49-85: psi power series, bit-reversed; :277-281 psi_enter = MM(., Rs) lazily
        brev = np.array([bit_reverse(i, logN) for i in range(N)], dtype=np.int64)
        self.psi_root = [primitive_root_2N(qi, N) for qi in self.q]
        psi_plain = np.empty((self.P, N), dtype=np.int64)
        ipsi_plain = np.empty((self.P, N), dtype=np.int64)
        tmp = np.empty(N, dtype=np.int64)
        for g, (qi, w) in enumerate(zip(self.q, self.psi_root)):
            call("orc_pow_series", tmp, N, _q(w), _q(qi))
            psi_plain[g] = tmp[brev]
            call("orc_pow_series", tmp, N, _q(pow(w, -1, qi)), _q(qi))
            ipsi_plain[g] = tmp[brev]
        self.psi_plain, self.ipsi_plain = psi_plain, ipsi_plain
        self.psi = psi_plain.copy()
        self.ipsi = ipsi_plain.copy()
        call("orc_mont_enter_scalar", self.psi, self.Rsa, self.P, N, self.qa, self.ka)
        call("orc_mont_enter_scalar", self.ipsi, self.Rsa, self.P, N, self.qa, self.ka)

        self.part = Partition(self.num_ordinary, K, num_devices)

        # ckks_engine.py:114-143: scales[level][i] = (q_level^-1 mod q_g) * R mod q_g, g = level+1..
        self.rescale_scales = []
        for level in range(self.num_levels):
            m0 = self.q[level]
            self.rescale_scales.append(
                _c([pow(m0, -1, self.q[g]) * R % self.q[g] for g in range(level + 1, self.num_ordinary)])
            )

        # ckks_engine.py:201-207: PiR[k][g] = (P_k^-1 mod q_g) * R mod q_g for g < index of special prime k
        self.sp = list(range(self.num_ordinary, self.P))
        self.PiR = np.zeros((K, self.P), dtype=np.int64)
        for kk, pk in enumerate(self.sp):
            for g in range(pk):
                self.PiR[kk, g] = pow(self.q[pk], -1, self.q[g]) * R % self.q[g]

        # ckks_engine.py:241-256
        Pprod = math.prod(self.q[g] for g in self.sp)
        self.mont_PR = _c([Pprod * R % self.q[g] for g in range(self.num_ordinary)])

    # ---- row -> constants --------------------------------------------------------------
    def rows(self, primes):
        idx = np.asarray(list(primes), dtype=np.int64)
        return self.qa[idx].copy(), self.ka[idx].copy()

    def level_primes(self, level: int, with_special: bool):
        hi = self.P if with_special else self.num_ordinary
        return list(range(level, hi))

    # ---- key-switch scalars (ntt_context.py:497-534) ------------------------------------
    def group_scalars(self, primes: list[int]):
        """For a digit group with global prime ids `primes` (= m_0..m_{alpha-1}):
        Y[i] = (L_i^-1 mod m_{i+1}) R mod m_{i+1};  Lsc[i][j-(i+2)] = L_i R mod m_j, j >= i+2;
        Lenter[i][g] = L_i R^2 mod q_g for every global prime g;   L_i = m_0 ... m_i."""
        m = [self.q[p] for p in primes]
        alpha = len(m)
        L = [m[0]]
        for i in range(1, alpha - 1):
            L.append(L[-1] * m[i])
        Y, Lsc = [], []
        for i in range(alpha - 1):
            Y.append(pow(L[i], -1, m[i + 1]) * R % m[i + 1])
            Lsc.append([L[i] * R % m[j] for j in range(i + 2, alpha)])
        Lenter = np.zeros((max(alpha - 1, 0), self.P), dtype=np.int64)
        for i in range(alpha - 1):
            for g in range(self.P):
                Lenter[i, g] = L[i] * self.Rs[g] % self.q[g]
        return Y, Lsc, Lenter


def _q(x):
    import ctypes

    return ctypes.c_int64(int(x))


_ = lib  # keep the import (forces the build on first use)
