"""tiberate/libs/wrapper/csprng_ops.py mirror (schemas: csrc/csprng/chacha20.cpp:35-37,
randint.cpp:37-42, discrete_gaussian.cpp, randround.cpp:21-23): same names, per-device tensor lists,
numpy host addresses for the q / CDT tables, same in-place vs functional behaviour."""

from __future__ import annotations

import torch

from .._native import get_lib


def _dev(t) -> int:
    if not t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous():
        raise RuntimeError("csprng tensors must be contiguous int64 CUDA tensors")  # macros.h:3-12 CHECK_INPUT
    return t.device.index or 0


def _st(t) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def chacha20(input: list[torch.Tensor], step: int) -> list[torch.Tensor]:
    """Blocks of the current states [n, 16]; the states are stepped in place."""
    lib, outs = get_lib(), []
    for s in input:
        d = _dev(s)
        out = torch.empty_like(s)
        lib.check(lib.tb200_chacha20(d, s.data_ptr(), s.numel() // 16, out.data_ptr(), int(step), _st(s)), "chacha20")
        outs.append(out)
    return outs


def randint_fast(input: list[torch.Tensor], q_ptrs: list[int], shift: int, step: int) -> list[torch.Tensor]:
    """states [C, L, 16] -> [C, 4L] uniform in [shift, q_c + shift)."""
    lib, outs = get_lib(), []
    for s, qp in zip(input, q_ptrs):
        d = _dev(s)
        C_, L = s.size(0), s.size(1)
        out = s.new_empty((C_, L * 4))
        lib.check(lib.tb200_randint_fast(d, s.data_ptr(), C_, L, int(qp), int(shift), int(step), out.data_ptr(), _st(s)),
                  "randint_fast")
        outs.append(out)
    return outs


def discrete_gaussian_fast(input: list[torch.Tensor], btree_ptr: int, btree_size: int, depth: int,
                           step: int) -> list[torch.Tensor]:
    """states [n, 16] -> [4n] discrete Gaussian samples."""
    lib, outs = get_lib(), []
    for s in input:
        d = _dev(s)
        n = s.numel() // 16
        out = s.new_empty((n * 4,))
        lib.check(lib.tb200_discrete_gaussian_fast(d, s.data_ptr(), n, int(btree_ptr), int(btree_size), int(depth),
                                                   int(step), out.data_ptr(), _st(s)), "discrete_gaussian_fast")
        outs.append(out)
    return outs


def randint(input: list[torch.Tensor], q_ptrs: list[int]) -> None:
    """In place on random words [C, L, 16]: word 4j of every row becomes the sample of words 4j..4j+3."""
    lib = get_lib()
    for w, qp in zip(input, q_ptrs):
        d = _dev(w)
        lib.check(lib.tb200_randint(d, w.data_ptr(), w.size(0), w.size(1), int(qp), _st(w)), "randint")


def discrete_gaussian(input: list[torch.Tensor], btree_ptr: int, btree_size: int, depth: int) -> None:
    lib = get_lib()
    for w in input:
        d = _dev(w)
        lib.check(lib.tb200_discrete_gaussian(d, w.data_ptr(), w.numel() // 16, int(btree_ptr), int(btree_size),
                                              int(depth), _st(w)), "discrete_gaussian")


def randround(input: list[torch.Tensor], rand_bytes: list[torch.Tensor]) -> None:
    """rand_bytes[i] (32-bit random words) <- randomly rounded input[i] (float64), in place."""
    lib = get_lib()
    for c, w in zip(input, rand_bytes):
        d = _dev(w)
        if c.dtype != torch.float64 or c.numel() < w.numel():
            raise RuntimeError("randround: input must be a float64 tensor at least as long as rand_bytes")
        c = c.contiguous()  # the reference reads `input` through a strided accessor (e.g. the .real view of an FFT)
        lib.check(lib.tb200_randround(d, c.data_ptr(), w.data_ptr(), w.numel(), _st(w)), "randround")
