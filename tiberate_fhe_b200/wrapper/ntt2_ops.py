"""tiberate/libs/wrapper/ntt2_ops.py mirror (schemas: csrc/ops/ntt_radix2.cpp:29-36,
intt_radix2.cpp:48-61).  The even/odd index tables and the expanded psi tensors the reference
passes are accepted and ignored: libtb200 derives the same butterfly network and the same (compact)
twiddles from the primes (tests check the tables are identical, tests/golden_check.py)."""

from __future__ import annotations

from . import context_for


def ntt_radix2(a, even, odd, psi, sp_prime_len):
    """Stages only; transforms the first a.size(0) - sp_prime_len rows (ntt_radix2_cuda.cu:63)."""
    for ai in a:
        ctx = context_for(ai)
        rows = ai.size(0) - sp_prime_len
        if rows > 0:
            ctx.ntt(ai, ctx.P - ai.size(0), False, rows=rows)


def enter_ntt_radix2(a, even, odd, psi, sp_prime_len):
    for ai in a:
        ctx = context_for(ai)
        ctx.ntt(ai, ctx.prime0_for(ai.size(0), sp_prime_len), True)


def intt_radix2(a, ieven, iodd, ipsi, sp_prime_len):
    """Stays in Montgomery form; first a.size(0) - sp_prime_len rows (intt_radix2_cuda.cu:65)."""
    for ai in a:
        ctx = context_for(ai)
        rows = ai.size(0) - sp_prime_len
        if rows > 0:
            ctx.intt(ai, ctx.P - ai.size(0), 0, rows=rows)


def _exit(a, sp_prime_len, mode):
    for ai in a:
        ctx = context_for(ai)
        ctx.intt(ai, ctx.prime0_for(ai.size(0), sp_prime_len), mode)


def intt_radix2_exit(a, ieven, iodd, ipsi, sp_prime_len):
    _exit(a, sp_prime_len, 1)


def intt_radix2_exit_reduce(a, ieven, iodd, ipsi, sp_prime_len):
    _exit(a, sp_prime_len, 2)


def intt_radix2_exit_reduce_signed(a, ieven, iodd, ipsi, sp_prime_len):
    _exit(a, sp_prime_len, 3)
