"""Csprng: mirror of tiberate/rng/csprng/csprng.py (same constructor arguments, attributes and methods,
same state-tensor layout, so states can be copied between the two) on the libtb200 CSPRNG kernels.

Layout (csprng.py:113-178): per device a state tensor [(share + repeats) * L, 16] int64, L = N / 4, one
ChaCha20 block state per row (words 0-3 constants, 4-11 key, 12-13 counter, 14-15 nonce); channels
[0, share) have device-unique counters, the trailing `num_repeating_channels` channels have the same
counters on every device and therefore produce identical numbers everywhere; after each use a row's
counter advances by `inc` = the total number of rows of all devices.

Deviation kept deliberately: the reference ignores its `seed` / `nonce` arguments (csprng.py:214-221
pass `seed=None` down) and always draws from os.urandom; here explicit values are honoured, None draws
from os.urandom as the reference does.
"""

from __future__ import annotations

import math
import os

import numpy as np
import torch

from ..wrapper import csprng_ops

_SIGMA_WORDS = (1634760805, 857760878, 2036477234, 1797285236)  # "expand 32-byte k" (csprng.py:101-118)


def build_cdt_tree(sigma: float = 3.2, security_bits: int = 128):
    """Cumulative distribution table of the half Gaussian as a binary search tree
    (tiberate/rng/csprng/discrete_gaussian_sampler.py:9-112): 2^ceil(log2(6 sigma)) sampling points,
    128-bit fixed point, probability at 0 halved; node order = breadth first.  Returns
    (uint64 array = lows then highs, size, depth)."""
    import mpmath as mpm

    with mpm.workprec(security_bits * 2):
        power = math.ceil(math.log2(6 * sigma))
        n = 2 ** power
        s, two = mpm.mpf(str(sigma)), mpm.mpf("2")
        norm = s * mpm.sqrt(two * mpm.pi)
        prob = [mpm.exp(-mpm.mpf(str(x)) ** 2 / (two * s ** 2)) / norm for x in range(n)]
        prob[0] /= 2
        cdt, acc = [0], mpm.mpf(0)
        for p in prob:
            acc = acc + p
            cdt.append(int(acc * two ** mpm.mpf(str(security_bits))))
    order = []
    for depth in range(power):
        nodes = 2 ** depth
        order += list(range(n // nodes // 2, n, n // nodes))
    m64 = (1 << 64) - 1
    table = [cdt[i] & m64 for i in order] + [(cdt[i] >> 64) & m64 for i in order]
    return np.ascontiguousarray(table, dtype=np.uint64), len(order), power


class Csprng:
    def __init__(self, num_coefs=2 ** 15, num_channels=[8], num_repeating_channels=2, sigma=3.2, devices=None,
                 seed=None, nonce=None):
        self.num_coefs = num_coefs
        self.num_channels = num_channels
        self.num_repeating_channels = num_repeating_channels
        self.sigma = sigma
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        self.devices = devices
        self.num_devices = len(devices)
        if len(num_channels) == 1:
            self.shares = [num_channels[0]] * self.num_devices
        elif len(num_channels) == self.num_devices:
            self.shares = list(num_channels)
        else:
            raise Exception("There was a contradicting mismatch between num_channels, and devices.")
        self.total_num_channels = sum(self.shares)
        self.L = num_coefs // 4
        self.btree, self.btree_size, self.tree_depth = build_cdt_tree(sigma=sigma)
        self.btree_ptr = self.btree.__array_interface__["data"][0]
        self.start_ind = [0]
        for s in self.shares[:-1]:
            self.start_ind.append(self.start_ind[-1] + s * self.L)  # device-unique counter ranges
        self.inc = (self.total_num_channels + num_repeating_channels) * self.L
        self.repeating_start = self.total_num_channels * self.L
        self.states, self.channeled_states, self.counters = [], [], []
        for dev_id, dev in enumerate(devices):
            rows = (self.shares[dev_id] + num_repeating_channels) * self.L
            st = torch.zeros((rows, 16), dtype=torch.int64, device=dev)
            self.states.append(st)
            self.channeled_states.append(st.view(self.shares[dev_id] + num_repeating_channels, self.L, 16))
            own = torch.arange(self.start_ind[dev_id], self.start_ind[dev_id] + self.shares[dev_id] * self.L,
                               dtype=torch.int64, device=dev)
            rep = torch.arange(self.repeating_start, self.inc, dtype=torch.int64, device=dev)
            self.counters.append(torch.cat([own, rep]))
        self.refresh(seed, nonce)

    # NOTE on start_ind: the reference computes start_ind = [0] + [s * L for s in shares[:-1]] (not a
    # running sum), which makes the ranges of devices >= 2 overlap when there are 3+ devices; with one or
    # two devices both formulas agree.
    def refresh(self, seed=None, nonce=None):
        def words(n, given):
            if given is None:
                return [int.from_bytes(os.urandom(4), "big") for _ in range(n)]
            given = [int(v) & 0xFFFFFFFF for v in given]
            if len(given) != n:
                raise ValueError(f"expected {n} 32-bit words")
            return given

        key, non = words(8, seed), words(2, nonce)
        self.key = [torch.tensor(key, dtype=torch.int64, device=d) for d in self.devices]
        self.nonce = [torch.tensor(non, dtype=torch.int64, device=d) for d in self.devices]
        for dev_id in range(self.num_devices):
            self.initialize_states(dev_id)

    def initialize_states(self, dev_id, seed=None, nonce=None):
        st = self.states[dev_id]
        st.zero_()
        st[:, 12] = self.counters[dev_id]
        st[:, 0:4] = torch.tensor(_SIGMA_WORDS, dtype=torch.int64, device=st.device)[None, :]
        st[:, 4:12] = self.key[dev_id][None, :]
        st[:, 14:] = self.nonce[dev_id][None, :]

    def _target(self, devi, share, repeats):
        start = self.shares[devi] - share
        return self.channeled_states[devi][start : self.shares[devi] + repeats]

    def randbytes(self, shares=None, repeats=0, reshape=False):
        if shares is None:
            shares = self.shares
        target = [self._target(d, shares[d], repeats).view(-1, 16) for d in range(self.num_devices)]
        out = csprng_ops.chacha20(target, self.inc)
        if reshape:
            out = [rb.view(-1, self.L, 16) for rb in out]
        return out

    def randint(self, amax=3, shift=0, repeats=0):
        """Uniform integers in [shift, amax_c + shift) per channel; amax: scalar or per-device lists of
        per-channel bounds whose last `repeats` entries address the repeating channels."""
        if not isinstance(amax, (list, tuple)):
            amax = [[amax] for _ in self.shares]
        q = [np.ascontiguousarray(a, dtype=np.uint64) for a in amax]
        target = [self._target(d, len(amax[d]) - repeats, repeats) for d in range(self.num_devices)]
        return csprng_ops.randint_fast(target, [a.__array_interface__["data"][0] for a in q], shift, self.inc)

    def discrete_gaussian(self, non_repeats=0, repeats=1):
        shares = list(non_repeats) if isinstance(non_repeats, (list, tuple)) else [non_repeats] * self.num_devices
        target = [self._target(d, shares[d], repeats).view(-1, 16) for d in range(self.num_devices)]
        out = csprng_ops.discrete_gaussian_fast(target, self.btree_ptr, self.btree_size, self.tree_depth, self.inc)
        return [o.view(-1, self.num_coefs) for o in out]

    def randround(self, coef):
        """Randomly round the float64 tensor `coef` (on the first device); returns int64."""
        rows = self.num_coefs // 16
        words = csprng_ops.chacha20((self.states[0][:rows],), self.inc)[0].ravel()
        csprng_ops.randround([coef], [words])
        return words
