"""The op layer registered with the PyTorch dispatcher (torch.ops.tb200_*).

The reference's boundary is a set of torch custom operators declared with TORCH_LIBRARY_FRAGMENT
(csrc/ops/mont.cpp:138-170, mont_extra.cpp:67-88, ntt_radix2.cpp:29-41, intt_radix2.cpp:48-68,
he_fused.cpp:80-110) and loaded by tiberate/libs/__init__.py:17-37.  `register()` declares the same
operator schemas -- same names, argument lists, mutability annotations and return types -- under our own
namespaces (two libraries defining `tiberate_*` cannot be loaded into one process, and the same-process
A/B test needs the reference's extension loaded too):

    torch.ops.tb200_mont_ops.*   torch.ops.tb200_ntt2_ops.*   torch.ops.tb200_he_ops.*

and binds their CUDA kernels to the extern "C" launchers of libtb200 (through wrapper/*.py, which resolve the
reference's right-aligned constant-pool indexing and forward raw pointers + the current stream).  Callers
that hold `torch.ops.tiberate_mont_ops.mont_mult` can therefore switch by changing the namespace only.
"""

from __future__ import annotations

SCHEMAS = {
    "mont_ops": {
        "mont_mult": "(Tensor[] a, Tensor[] b, int sp_prime_len) -> Tensor[]",
        "mont_enter_scalar": "(Tensor[](a!) a, Tensor[] b, int sp_prime_len) -> ()",
        "mont_enter_Rs": "(Tensor[](a!) a, int sp_prime_len) -> ()",
        "mont_enter_Rs_scale": "(Tensor[](a!) a, int sp_prime_len) -> ()",
        "mont_enter": "(Tensor[](a!) a, Tensor[] Rs, Tensor[] ql, Tensor[] qh, Tensor[] kl, Tensor[] kh) -> ()",
        "mont_reduce": "(Tensor[](a!) a, int sp_prime_len) -> ()",
        "mont_add": "(Tensor[] a, Tensor[] b, int sp_prime_len) -> Tensor[]",
        "mont_add_legacy": "(Tensor[] a, Tensor[] b, Tensor[] _2q) -> Tensor[]",
        "mont_sub": "(Tensor[] a, Tensor[] b, int sp_prime_len) -> Tensor[]",
        "reduce_2q": "(Tensor[](a!) a, int sp_prime_len) -> ()",
        "make_signed": "(Tensor[](a!) a, int sp_prime_len) -> ()",
        "make_unsigned": "(Tensor[](a!) a, int sp_prime_len) -> ()",
        "tile_unsigned": "(Tensor[] a, Tensor[] _2q) -> Tensor[]",
        "mont_add_many_3d": "(Tensor[] input, int sp_prime_len) -> Tensor[]",
        "mont_reduce_add_many_3d": "(Tensor[] input, int sp_prime_len) -> Tensor[]",
        "mont_add_reduce_2q": "(Tensor[] a, Tensor[] b, int sp_prime_len) -> Tensor[]",
        "mont_sub_reduce_2q": "(Tensor[] a, Tensor[] b, int sp_prime_len) -> Tensor[]",
        "mont_enter_scalar_reduce_2q": "(Tensor[] a, Tensor[] b, int sp_prime_len) -> Tensor[]",
    },
    "ntt2_ops": {
        name: "(Tensor[](a!) a, Tensor[] even, Tensor[] odd, Tensor[] psi, int sp_prime_len) -> ()"
        for name in ("ntt_radix2", "enter_ntt_radix2", "intt_radix2", "intt_radix2_exit", "intt_radix2_exit_reduce",
                     "intt_radix2_exit_reduce_signed")
    },
    "he_ops": {
        "pc_add_fused": "(Tensor[] ct, Tensor[] pt, int sp_prime_len) -> Tensor[]",
        "rescale_exact_rounding_fused":
            "(Tensor[](a!) a, Tensor[] scales, Tensor[] rescaler, int round_at, int sp_prime_len) -> ()",
        "rescale_non_exact_rounding_fused": "(Tensor[](a!) a, Tensor[] scales, Tensor[] rescaler, int sp_prime_len) -> ()",
        "switch_key_switch_later_part_extend":
            "(int rns_len, Tensor state, Tensor l_enter, int l_enter_start_offset, int sp_prime_len) -> Tensor",
        "codec_rotate_make_unsigned_reduce_2q": "(Tensor[] a, Tensor[] perm, Tensor[] _2q) -> Tensor[]",
        "create_switcher_divide_by_p": "(Tensor[] c, Tensor[] p, Tensor[][] PiRi) -> Tensor[]",
    },
}

_libs = []


def namespace(module: str) -> str:
    return f"tb200_{module}"


def register() -> dict:
    """Idempotent.  Returns {module: [operator names]}."""
    if _libs:
        return {m: list(ops) for m, ops in SCHEMAS.items()}
    import torch

    from . import wrapper

    for module, ops in SCHEMAS.items():
        lib = torch.library.Library(namespace(module), "DEF")
        impl_mod = getattr(wrapper, module)
        for name, schema in ops.items():
            lib.define(name + schema)
            fn = getattr(impl_mod, name)
            mutating = "(a!)" in schema

            def kernel(*args, _fn=fn, _mut=mutating):
                out = _fn(*args)
                return None if _mut else out

            # every tensor argument lives on the device of the current context; there is no CPU kernel
            lib.impl(name, kernel, "CUDA")
        _libs.append(lib)
    return {m: list(ops) for m, ops in SCHEMAS.items()}
