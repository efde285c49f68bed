"""Boundary robustness (VERDICT r01 weak 6 / ADVICE): engine calls lease their scratch from the stream-ordered
memory pool on the caller's stream, so one context can be used from several CUDA streams at once, the calls
never synchronise the device, they can be captured into a CUDA graph, and the caller's current device is
restored."""

import pytest

pytestmark = pytest.mark.gpu


def _setup(torch, B=6):
    from oracle.context import PRESETS
    from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context

    q, K = PRESETS[14]["q"], PRESETS[14]["K"]
    ctx = Tb200Context(14, q, K)
    ctx.set_chunk(2)
    N, P, no = ctx.N, ctx.P, ctx.num_ordinary
    gen = torch.Generator(device="cuda").manual_seed(77)

    def uni(primes, *lead):
        return torch.stack([torch.randint(0, int(qi), (*lead, N), device="cuda", generator=gen) for qi in primes], dim=-2)

    key = KeySwitchKeyView([(uni(q), uni(q)) for _ in range(ctx.num_groups0)], N)
    cts = [uni(q[:no], B) for _ in range(4)]
    return ctx, key, cts, no, N


def test_two_streams_share_one_context():
    import torch

    ctx, key, (a0, a1, b0, b1), no, N = _setup(torch)
    B = a0.shape[0]
    want0, want1 = torch.empty(B, no - 1, N, dtype=torch.int64, device="cuda"), torch.empty(B, no - 1, N, dtype=torch.int64, device="cuda")
    ctx.cc_mult_relin(0, a0, a1, b0, b1, key, want0, want1, True)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [[torch.zeros_like(want0), torch.zeros_like(want1)] for _ in range(2)]
    for rep in range(3):  # interleaved launches: with a shared workspace the two calls would overwrite each other
        for s, o in zip((s1, s2), outs):
            with torch.cuda.stream(s):
                ctx.cc_mult_relin(0, a0, a1, b0, b1, key, o[0], o[1], True)
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o[0], want0) and torch.equal(o[1], want1)
    ctx.close()


def test_engine_call_is_cuda_graph_capturable():
    import torch

    from tiberate_fhe_b200 import get_lib
    from tiberate_fhe_b200.context import galois_element

    ctx, key, (a0, a1, b0, b1), no, N = _setup(torch, B=2)
    want0, want1 = torch.empty(2, no - 1, N, dtype=torch.int64, device="cuda"), torch.empty(2, no - 1, N, dtype=torch.int64, device="cuda")
    r0, r1 = torch.empty_like(a0), torch.empty_like(a1)
    ctx.cc_mult_relin(0, a0, a1, b0, b1, key, want0, want1, True)   # also warms the kernels up (attribute calls)
    ctx.rotate(0, galois_element(N, 1), a0, a1, key, r0, r1)
    torch.cuda.synchronize()
    o0, o1 = torch.zeros_like(want0), torch.zeros_like(want1)
    g0, g1 = torch.zeros_like(r0), torch.zeros_like(r1)
    graph = torch.cuda.CUDAGraph()
    lib = get_lib()
    with torch.cuda.graph(graph):
        ctx.cc_mult_relin(0, a0, a1, b0, b1, key, o0, o1, True)
        ctx.rotate(0, galois_element(N, 1), a0, a1, key, g0, g1)
    assert int(o0.abs().sum()) == 0, "capture must not execute"
    before = lib.tb200_launch_count()
    graph.replay()
    torch.cuda.synchronize()
    assert lib.tb200_launch_count() == before, "a replay launches nothing from the host side of the library"
    assert torch.equal(o0, want0) and torch.equal(o1, want1) and torch.equal(g0, r0) and torch.equal(g1, r1)
    a0.add_(1).remainder_(2)  # new inputs in the same buffers, replay again
    ctx.cc_mult_relin(0, a0, a1, b0, b1, key, want0, want1, True)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(o0, want0) and torch.equal(o1, want1)
    ctx.close()


def test_current_device_is_restored():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle.context import PRESETS
    from tiberate_fhe_b200.context import Tb200Context

    ctx = Tb200Context(14, PRESETS[14]["q"], PRESETS[14]["K"], device=1)
    torch.cuda.set_device(0)
    a = torch.zeros(ctx.num_ordinary, ctx.N, dtype=torch.int64, device="cuda:1")
    ctx.ntt(a, 0, True)
    assert torch.cuda.current_device() == 0
    ctx.close()


def test_packed_wire_format_on_the_device():
    import numpy as np
    import torch

    ctx, key, (a0, a1, b0, b1), no, N = _setup(torch, B=3)
    nar = ctx.narrow_rows(0, no)
    assert nar == no - 1  # every scale prime (40 or 41 bits), not the 60-bit base prime
    wide = [r for r in range(nar) if ctx.q[r] >= (1 << 40)]
    assert wide, "the preset has scale primes above 2^40"
    for r in wide:  # residues that need bit 40
        a0[0, r, 5] = ctx.q[r] - 1
        a0[2, r, N - 1] = 1 << 40
    packed = torch.empty(3, nar, ctx.packed_row_bytes, dtype=torch.uint8, device="cuda")
    ctx.pack41(a0[:, :nar], packed, 0)
    host = a0[:, :nar].cpu().numpy()
    assert int((host >> 40).max()) == 1, "the preset has scale primes above 2^40: bit 40 must be exercised"
    lo = (host & ((1 << 40) - 1)).astype("<u8").view(np.uint8).reshape(3, nar, N, 8)[..., :5].reshape(3, nar, 5 * N)
    hi = np.packbits(((host >> 40) & 1).astype(np.uint8), axis=-1, bitorder="little")
    assert (packed.cpu().numpy() == np.concatenate([lo, hi], axis=-1)).all()
    back = torch.zeros(3, no, N, dtype=torch.int64, device="cuda")
    ctx.unpack41(packed, back[:, :nar], 0)
    assert torch.equal(back[:, :nar], a0[:, :nar])
    ctx.close()
