"""Seeded inputs shared by the golden-vector generator (run against the reference's own CUDA
extension on a B200) and by the checkers (oracle on CPU, libtb200 on GPU).  Everything is a pure
function of (seed, primes, N) so that fixtures only need to store hashes of the OUTPUTS."""

from __future__ import annotations

import hashlib

import numpy as np


def uniform(seed: int, primes, N: int) -> np.ndarray:
    """[len(primes), N] canonical residues."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.integers(0, int(q), size=N, dtype=np.int64) for q in primes])


def lazy_signed(seed: int, primes, N: int) -> np.ndarray:
    """Residues in (-q/2, 2q): what the reference's lazy arithmetic leaves in keys / intermediates
    (mont_sub keeps negatives, MM outputs reach 1.5q)."""
    rng = np.random.default_rng(seed)
    rows = []
    for q in primes:
        q = int(q)
        rows.append(rng.integers(-(q // 2), 2 * q, size=N, dtype=np.int64))
    return np.stack(rows)


def ciphertext(seed: int, primes, N: int):
    return [uniform(seed, primes, N), uniform(seed + 1, primes, N)]


def ksk(seed: int, primes_all, N: int, num_groups: int):
    """Synthetic key-switch key: per digit group (b, a) over all P primes; b carries lazy/negative
    residues like a real key (b = e - a*s through mont_sub), a is uniform."""
    return [(lazy_signed(seed + 10 * g, primes_all, N), uniform(seed + 10 * g + 1, primes_all, N))
            for g in range(num_groups)]


def digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a, dtype=np.int64)
    return hashlib.sha256(a.tobytes()).hexdigest()


def record(a: np.ndarray) -> dict:
    a = np.ascontiguousarray(a, dtype=np.int64)
    return {"shape": list(a.shape), "sha256": digest(a), "head": a.reshape(-1)[:4].tolist(),
            "tail": a.reshape(-1)[-4:].tolist()}


# the cases: name -> (levels to run at); kept identical in generator and checkers
LEVEL_CASES = {14: [0, 3, 6], 15: [0, 1, 8, 15], 16: [0, 17, 33]}
ROT_DELTA = 5
