"""GPU parity tests proper: libtb200.so (CUDA, through the C ABI) vs the oracle, bit-exact.

Run on the B200 box: python -m pytest tests -m gpu.  Sizes: toy rings (every tile-round pattern),
the reference presets logN14/15/16 on real prime chains, and
size-independent properties at the BASELINE.json batch sizes.
"""

import numpy as np
import pytest

import parity
from parity import Harness, Setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    from tiberate_fhe_b200 import get_lib

    return Harness(get_lib(), use_torch=True)


def test_native_library_is_loaded(h):
    import torch

    assert torch.cuda.is_available()
    assert h.lib.path.endswith("libtb200.so")
    assert b"sm_100a" in h.lib.tb200_version()
    before = h.lib.tb200_launch_count()
    s = Setup.toy(h, 8, 2, 1)
    parity.check_ntt(s, 0, True, 1)
    s.close()
    assert h.lib.tb200_launch_count() > before, "no CUDA kernel was launched"


TOY = [(8, 5, 2), (9, 4, 3), (10, 3, 1), (11, 6, 4), (12, 2, 2), (13, 3, 2)]


@pytest.mark.parametrize("logN,ns,K", TOY)
def test_toy_rings(h, logN, ns, K):
    s = Setup.toy(h, logN, ns, K, seed=logN, rot_deltas=(1, 3))
    try:
        parity.check_context(s)
        parity.check_ntt(s, 0, True, 2)
        parity.check_pointwise(s, 0, True)
        parity.check_pointwise(s, min(1, ns), False)
        for level in range(0, ns + 1):
            parity.check_he_ops(s, level)
            for mode in parity.ENGINE_MODES:  # internal mod-q path / exact op kernels: same bits
                parity.set_mode(s, mode)
                parity.check_engine(s, level)
        s.ctx.set_chunk(2)
        parity.check_engine(s, 0, batch=3)
    finally:
        s.close()


@pytest.mark.parametrize("logN", [9, 12])
def test_primes_far_from_a_power_of_two(h, logN):
    """Every prime of the reference lies just below a power of two; this chain (0.7 * 2^55 scale primes,
    0.75 * 2^60 base / special primes) checks that no bound of the lazy arithmetic relies on that."""
    s = Setup(h, logN, parity.far_primes(logN, 4, 2), 2, seed=500 + logN)
    try:
        parity.check_ntt(s, 0, True, 2)
        for level in (0, 2, 4):
            for mode in parity.ENGINE_MODES:
                parity.set_mode(s, mode)
                parity.check_engine(s, level)
    finally:
        s.close()


def _preset(h, logN, **kw):
    from oracle.context import PRESETS

    return Setup(h, logN, PRESETS[logN]["q"], PRESETS[logN]["K"], seed=logN, **kw)


def test_preset_logN14(h):
    """BASELINE config 1 shape on the GPU: logN14, 9 limbs."""
    s = _preset(h, 14)
    try:
        parity.check_context(s)
        parity.check_ntt(s, 0, True, 2)
        parity.check_pointwise(s, 0, True)
        for level in (0, 3, 6):
            parity.check_he_ops(s, level)
            for mode in parity.ENGINE_MODES:
                parity.set_mode(s, mode)
                parity.check_engine(s, level)
    finally:
        s.close()


def test_preset_logN15_default_engine(h):
    """BASELINE config 2 shape: CkksEngine default (logN15, 19 limbs, K=2, 9 digit groups)."""
    s = _preset(h, 15)
    try:
        parity.check_context(s)
        parity.check_ntt(s, 0, True, 1)
        for level in (0, 1, 15):
            parity.check_engine(s, level, ops=("rescale", "keyswitch", "rotate", "cc_mult", "pc_mult", "addsub"))
        parity.check_engine(s, 7, batch=2, ops=("cc_mult", "rotate"))
    finally:
        s.close()


def test_preset_logN16_headline(h):
    """The headline configuration: logN16, 39 limbs, K=4, 10 digit groups, first multiplicative level."""
    s = _preset(h, 16)
    try:
        parity.check_ntt(s, 0, True, 1)
        parity.check_engine(s, 0, ops=("rescale", "cc_mult", "rotate"))
        parity.check_engine(s, 17, ops=("keyswitch", "cc_mult"))
        parity.check_engine(s, 33, ops=("cc_mult", "rotate", "pc_mult"))
    finally:
        s.close()


# ---- size-independent properties at the BASELINE batch sizes ------------------------------------
def _torch_uniform(torch, q, shape_tail, gen):
    rows = [torch.randint(0, int(qi), shape_tail, dtype=torch.int64, device="cuda", generator=gen) for qi in q]
    return torch.stack(rows, dim=-2)


def test_properties_logN16_batch(h):
    import torch

    from oracle.context import PRESETS
    from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context, galois_element

    q, K = PRESETS[16]["q"], PRESETS[16]["K"]
    ctx = Tb200Context(16, q, K)
    N, P, no = ctx.N, ctx.P, ctx.num_ordinary
    gen = torch.Generator(device="cuda").manual_seed(0xB200)
    B = 16
    a = _torch_uniform(torch, q[:no], (B, N), gen)  # [B, 35, N]
    # 1. NTT round trip on every limb of a batch
    t = a.clone()
    ctx.ntt(t, 0, True)
    assert not torch.equal(t, a)
    ctx.intt(t, 0, 2)
    assert torch.equal(t, a), "intt_exit_reduce(enter_ntt(a)) != a"
    # 2. negacyclic convolution with a monomial: a * X^k = signed shift  (pc_mult without rescale)
    k = 12345
    mono = torch.zeros(no, N, dtype=torch.int64, device="cuda")
    mono[:, k] = 1
    ctx.ntt(mono, 0, True)
    o0, o1 = torch.empty_like(a), torch.empty_like(a)
    ctx.pc_mult(0, mono, a, a, o0, o1, post_rescale=False)
    qv = torch.tensor(q[:no], dtype=torch.int64, device="cuda")[None, :, None]
    want = torch.roll(a, k, dims=-1)
    want[..., :k] = (qv - want[..., :k]) % qv
    assert torch.equal(o0, want) and torch.equal(o1, want), "a * X^k is not the negacyclic shift"
    # 3. automorphisms compose: sigma_g(sigma_h(a)) == sigma_{gh}(a)
    g1, g2 = galois_element(N, 1), galois_element(N, 7)
    x0, x1, y0, y1 = (torch.empty_like(a) for _ in range(4))
    ctx.rotate(0, g1, a, a, None, x0, x1)
    ctx.rotate(0, g2, x0, x1, None, y0, y1)
    ctx.rotate(0, g1 * g2 % (2 * N), a, a, None, x0, x1)
    assert torch.equal(x0, y0) and torch.equal(x1, y1)
    # 4. (a + b) - b == a
    b = _torch_uniform(torch, q[:no], (B, N), gen)
    ctx.cc_addsub(0, False, a, a, b, b, x0, x1)
    ctx.cc_addsub(0, True, x0, x1, b, b, y0, y1)
    # the reference's mont_sub keeps negatives (CS2(a - b) never adds 2q), so compare modulo q
    assert torch.equal(y0 % qv, a) and torch.equal(y1 % qv, a)
    # 5. batch / chunk invariance of HMult+relin and rotate with synthetic keys (BASELINE config 3 data)
    ng = ctx.num_groups0
    parts = [(_torch_uniform(torch, q, (N,), gen), _torch_uniform(torch, q, (N,), gen)) for _ in range(ng)]
    key = KeySwitchKeyView(parts, N)
    r0, r1 = torch.empty(B, no - 1, N, dtype=torch.int64, device="cuda"), torch.empty(B, no - 1, N, dtype=torch.int64, device="cuda")
    ctx.set_chunk(4)
    ctx.cc_mult_relin(0, a, b, b, a, key, r0, r1, True)
    ctx.set_chunk(3)
    s0, s1 = torch.empty_like(r0), torch.empty_like(r1)
    ctx.cc_mult_relin(0, a, b, b, a, key, s0, s1, True)
    assert torch.equal(r0, s0) and torch.equal(r1, s1), "result depends on the batch chunking"
    one0, one1 = torch.empty(no - 1, N, dtype=torch.int64, device="cuda"), torch.empty(no - 1, N, dtype=torch.int64, device="cuda")
    ctx.cc_mult_relin(0, a[5], b[5], b[5], a[5], key, one0, one1, True)
    assert torch.equal(one0, r0[5]) and torch.equal(one1, r1[5]), "batched result != single-ciphertext result"
    assert int(r0.min()) >= 0 and bool((r0 < qv[:, 1:]).all()), "HMult output must be canonical"
    ctx.rotate(0, g1, a, b, key, x0, x1)
    ctx.set_chunk(5)
    ctx.rotate(0, g1, a, b, key, y0, y1)
    assert torch.equal(x0, y0) and torch.equal(x1, y1)
    ctx.close()


def test_golden_reference_vectors(h):
    """Outputs of the reference's own CUDA extension (tests/golden/ref_*.json, generated on a B200 by
    tests/golden/make_ref_golden.py) must be reproduced bit for bit."""
    import glob
    import os

    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_logN*.json")))
    if not files:
        pytest.skip("no reference-extension fixtures committed yet")
    import golden_check

    for f in files:
        golden_check.check_file(h, f)
