"""CSPRNG operators: oracle pins (RFC 8439 known answer, the reference's CDT tree),
the kernels compiled for the host against the oracle (CPU), the CUDA kernels against the oracle and
the Csprng class against the oracle's restatement of the reference class (GPU)."""

import json
import os
import sys

import numpy as np
import pytest

import csprng_checks as cc  # noqa: I001
from oracle import csprng as oc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_oracle_rfc8439_block_known_answer():
    st = cc.make_states(1, cc.KEY, [0x4A000000, 0])
    st[0, 12], st[0, 13] = 1, 0x09000000
    out = oc.chacha20_block(st)[0]
    assert [hex(v) for v in out[:4]] == ["0xe4e7f110", "0x15593bd1", "0x1fdd0f50", "0xc47120a3"]
    assert [hex(v) for v in out[12:]] == ["0xd19c12b5", "0xb94e16de", "0xe883d0cb", "0x4e3c50a2"]


def test_oracle_cdt_tree_matches_reference_fixture():
    """tests/golden/ref_cdt_sigma3.2.json was produced by the reference's own
    build_CDT_binary_search_tree (tests/golden/make_ref_golden_csprng.py)."""
    with open(os.path.join(HERE, "golden", "ref_cdt_sigma3.2.json")) as f:
        g = json.load(f)
    lut, size, depth = oc.build_cdt_tree()
    assert (size, depth) == (g["size"], g["depth"])
    assert [int(v) for v in lut] == [int(v) for v in g["lut"]]
    from tiberate_fhe_b200.rng.csprng import build_cdt_tree

    lut2, s2, d2 = build_cdt_tree()
    assert (s2, d2) == (size, depth) and np.array_equal(lut2, lut)


def test_oracle_randint_is_the_reference_carry_chain():
    """randint_cuda.cu:56-81 computes floor(X p / 2^128) with an explicit 32-bit carry chain; restate
    that chain literally and compare with the big-integer floor on edge values."""
    M = 0xFFFFFFFF
    rng = np.random.default_rng(1)
    cases = [(0, 0, 0, 0, 3), (M, M, M, M, (1 << 64) - 1), (M, M, M, M, 3), (1, 0, 0, 0, (1 << 63))]
    for _ in range(300):
        w = [int(v) for v in rng.integers(0, 1 << 32, 4)]
        cases.append((*w, int(rng.integers(2, 1 << 62))))
    for x0, x1, x2, x3, p in cases:
        x_low = (x0 << 32) | x1
        alpha = (p * x_low) >> 64
        pl, ph, xhh, xhl = p & M, p >> 32, x2, x3
        plxhl, plxhh, phxhl, phxhh = pl * xhl, pl * xhh, ph * xhl, ph * xhh
        carry = ((plxhl & M) + (alpha & M)) >> 32
        carry = (carry + (plxhl >> 32) + (alpha >> 32) + (phxhl & M) + (plxhh & M)) >> 32
        sample = (carry + (phxhl >> 32) + (plxhh >> 32) + phxhh) & ((1 << 64) - 1)
        X = (((x2 << 32) | x3) << 64) | x_low
        assert sample == (X * p) >> 128


@pytest.mark.parametrize("check", cc.ALL, ids=lambda f: f.__name__)
def test_emulated_kernels(emu_lib, check):
    check(emu_lib, cc.Buf(False))


def test_oracle_csprng_golden():
    """The oracle's Csprng restatement against outputs of the reference's Csprng + CUDA extension
    (fixture generated on a B200 by tests/golden/make_ref_golden_csprng.py)."""
    path = os.path.join(HERE, "golden", "ref_csprng.json")
    if not os.path.exists(path):
        pytest.skip("fixture not generated yet")
    import golden_csprng

    golden_csprng.check_against(path, golden_csprng.run_oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("check", cc.ALL, ids=lambda f: f.__name__)
def test_gpu_kernels(check):
    from tiberate_fhe_b200._native import get_lib

    check(get_lib(), cc.Buf(True))


@pytest.mark.gpu
def test_gpu_csprng_class_matches_oracle_and_golden():
    import golden_csprng

    path = os.path.join(HERE, "golden", "ref_csprng.json")
    ours = golden_csprng.run_tb200()
    assert ours == golden_csprng.run_oracle(), "Csprng (libtb200) vs the oracle's restatement"
    if os.path.exists(path):
        golden_csprng.check_against(path, lambda: ours)


def test_codec_permutations_and_encode_match_reference_fixture():
    """the slot permutations depend on the reference's cycle enumeration; fixture from the
    reference's own prepost_perms (tests/golden/make_ref_golden_csprng.py codec)."""
    import hashlib

    import torch

    from tiberate_fhe_b200 import codec

    with open(os.path.join(HERE, "golden", "ref_codec.json")) as f:
        g = json.load(f)
    for logN, (hpre, hpost) in g["perms"].items():
        pre, post = codec.prepost_perms(1 << int(logN), "cpu")
        assert hashlib.sha256(pre.numpy().astype(np.int64).tobytes()).hexdigest() == hpre, f"pre_perm logN{logN}"
        assert hashlib.sha256(post.numpy().astype(np.int64).tobytes()).hexdigest() == hpost, f"post_perm logN{logN}"
    N = 1 << 10
    gen = torch.Generator().manual_seed(3)
    m = torch.randn(N // 2, generator=gen, dtype=torch.float64) + 1j * torch.randn(N // 2, generator=gen,
                                                                                  dtype=torch.float64)
    enc = codec.encode(m, device="cpu", deviation=1.25, return_without_scaling=True)
    assert np.allclose(enc[:8].numpy(), g["encode_logN10"], rtol=1e-12, atol=1e-12)
    # decode(encode(m)) = m  (canonical embedding round trip, no scaling)
    class _Cpu:  # decode() derives the device name from the tensor; CPU tensors have no index
        pass
    dec = torch.fft.ifft(enc * codec._twister(N, "cpu", +1), norm="forward")
    back = torch.zeros_like(dec)
    back[codec.prepost_perms(N, "cpu")[1]] = dec
    assert torch.allclose(back[: N // 2], m * 1.25, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1], ids=["left", "right"])
def test_gpu_constant_pool_round_trip(layout):
    """The reference's only live unit test (tests/test_constant_mem.py:21-74) against our const_pool mirror:
    mixed int64 / int32 entries at aligned offsets, left and right gravity, read back unchanged."""
    import torch

    from tiberate_fhe_b200.wrapper import const_pool

    torch.manual_seed(42)
    dev = torch.device("cuda:0")
    entries = [torch.randint(1, 1_000_000, (32,), dtype=dt, device=dev)
               for dt in (torch.int64, torch.int32, torch.int64, torch.int32, torch.int64, torch.int32, torch.int64)]
    offsets, off = [], 0
    for t in entries:
        off = (off + 7) // 8 * 8
        offsets.append(off)
        off += t.numel() * t.element_size()
    const_pool.upload_tensor_list(entries, offsets, layout, 0)
    dummy = torch.empty(0, device=dev)
    for t, o in zip(entries, offsets):
        got = const_pool.read_constant_chunk(dummy, o, t.numel(), t.dtype, layout)
        assert got.dtype == t.dtype and torch.equal(got, t)
    with pytest.raises(RuntimeError):
        const_pool.upload_tensor_list([torch.zeros(600, dtype=torch.int64, device=dev)], [0], layout, 0)
