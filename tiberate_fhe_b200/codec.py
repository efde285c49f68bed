"""Plaintext encoding / decoding: the canonical-embedding codec of
tiberate/utils/encoding.py -- slot permutations (:48-201), negacyclic FFT via twist / skew (:163-199),
encode with randomised rounding (:315-338), decode (:341-362), padding (:8-40).

Float64 work on the device through torch.fft (cuFFT); none of it is on the hot path.  The operations
and their order are the reference's, so equal inputs give bit-equal outputs on the same GPU library.
The slot permutations depend on the reference's particular cycle enumeration (a conjugating permutation
is not unique), which `_cycles` restates.
"""

from __future__ import annotations

import numpy as np
import torch

_perm_cache: dict = {}
_twist_cache: dict = {}


def padding(m, num_slots: int) -> torch.Tensor:
    if isinstance(m, (int, float)):
        m = [m]
    if isinstance(m, torch.Tensor):
        if m.dim() != 1:
            raise AssertionError(f"Input tensor should be 1D, but got {m.dim()}D.")
        return torch.cat((m.clone(), torch.zeros(num_slots - m.shape[0], device=m.device)))
    arr = np.asarray(m)
    return torch.tensor(np.pad(arr, (0, num_slots - len(arr)), constant_values=0))


def _circular_shift(N, shift=1):
    half = np.arange(N // 2)
    return np.concatenate([np.roll(half, shift), np.roll(half, -shift) + N // 2])


def _cycles(perm):
    """Cycles in the reference's enumeration (encoding.py:136-158): seeds in increasing index order, each
    cycle listed as perm[seed], perm[perm[seed]], ..., seed."""
    perm = [int(v) for v in perm]
    seen = [False] * len(perm)
    out = []
    for seed in range(len(perm)):
        if seen[seed]:
            continue
        cyc, cur = [], perm[seed]
        while not seen[cur]:
            seen[cur] = True
            cyc.append(cur)
            cur = perm[cur]
        out.append(cyc)
    return out


def _conjugate(p, q):
    """r with r[q-cycle position] = p-cycle position, cycles paired in enumeration order (:105-133)."""
    pc, qc = _cycles(p), _cycles(q)
    if [len(c) for c in pc] != [len(c) for c in qc]:
        raise AssertionError("Cycle structures of permutations must match for a conjugate to exist!!!")
    pe = np.array([i for c in pc for i in c])
    qe = np.array([i for c in qc for i in c])
    r = np.zeros_like(np.asarray(p))
    r[qe] = pe
    return r


def prepost_perms(N: int, device):
    key = (N, str(device))
    if key not in _perm_cache:
        circ = _circular_shift(N)
        canon = (3 * np.arange(2 * N)) % (2 * N)          # canon_permutation(N, k=1): p = 3
        fold = (canon[1::2] - 1) // 2                     # fold_permutation
        post = _conjugate(circ, fold)
        pre = np.arange(N)[np.argsort(post)][: N // 2]   # inverse_permutation(post)[:N/2]
        _perm_cache[key] = (torch.from_numpy(pre).to(device), torch.from_numpy(post).to(device))
    return _perm_cache[key]


def _twister(N, device, sign):
    key = (N, str(device), sign)
    if key not in _twist_cache:
        expr = sign * 1j * torch.pi * torch.arange(N, device=device, dtype=torch.float64) / N
        _twist_cache[key] = torch.exp(expr)
    return _twist_cache[key]


def encode(m, rng=None, scale=2 ** 40, deviation=1.0, device="cuda:0", norm="forward", return_without_scaling=False):
    """m: num_slots values -> N real polynomial coefficients (float64 when return_without_scaling, else
    int64 after rng.randround)."""
    N = len(m) * 2
    pre, _ = prepost_perms(N, device)
    mm = m * deviation
    if isinstance(mm, torch.Tensor):
        mm = mm.to(device, copy=True)
    permed = torch.zeros((N,), dtype=mm.dtype, device=mm.device)
    permed[pre] = mm
    mm = permed + permed.conj().flip(0)
    poly = (torch.fft.fft(mm, norm=norm) * _twister(N, device, -1)).real
    if return_without_scaling:
        return poly
    return rng.randround(poly * np.float64(scale))


def decode(m, scale=2 ** 40, correction=1.0, norm="forward", return_without_scaling=False):
    N = len(m)
    device = m.device.type + ":" + str(m.device.index)
    _, post = prepost_perms(N, device)
    mm = torch.fft.ifft(m * _twister(N, device, +1), norm=norm)
    if not return_without_scaling:
        mm = mm / scale * correction
    out = torch.zeros_like(mm)
    out[post] = mm
    return out
