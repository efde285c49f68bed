"""tiberate_fhe_b200.CkksEngine generating its own keys and ciphertexts (own CSPRNG
kernels + op sequencing) against the UNMODIFIED reference engine on its own CUDA extension, both fed
the same ChaCha20 key / nonce: secret, public, evaluation and rotation keys, encrypt at two levels and
decrypt must be bit-identical.  Then a self-contained homomorphic round trip on our engine alone.
Skipped when baseline/_ref is absent (git-ignored reference install)."""

import pytest

pytestmark = pytest.mark.gpu

KEY = [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0, 0x082EFA98, 0xEC4E6C89]
NONCE = [0x452821E6, 0x38D01377]


def _tensors(x, out):
    import torch

    if isinstance(x, torch.Tensor):
        out.append(x)
    elif isinstance(x, (list, tuple)):
        for y in x:
            _tensors(y, out)
    elif hasattr(x, "data"):
        _tensors(x.data, out)
    return out


def _message(N, device):
    import torch

    g = torch.Generator().manual_seed(99)
    return torch.randint(-(1 << 20), 1 << 20, (N,), generator=g, dtype=torch.int64).to(device)


def _run(engine, rec):
    import torch

    rec["sk"] = engine.sk
    rec["pk"] = engine.pk
    rec["evk"] = engine.evk
    rec["rotk1"] = engine.rotk[1]
    rec["rotk5"] = engine.rotk[5]
    m = _message(1 << engine_logN(engine), "cuda:0")
    for level in (0, 3):
        ct = engine.encrypt([m.clone()], level=level)
        rec[f"encrypt_l{level}"] = ct
        # the reference's encrypt() tags its coefficient-domain output NTT|MONTGOMERY (ckks_engine.py:621)
        plain = type(ct)(data=ct.data, level=ct.level)
        rec[f"decrypt_l{level}"] = engine.decrypt(plain)
    torch.cuda.synchronize()
    return rec


def engine_logN(engine):
    return engine.ckksCfg.logN if hasattr(engine, "ckksCfg") else engine.logN


@pytest.mark.parametrize("preset", ["logN14", "logN15", "logN16"])
def test_keys_and_ciphertexts_match_the_reference_engine(preset):
    import torch

    from baseline import ref_harness

    if not ref_harness.available():
        pytest.skip("baseline/_ref (reference install) not present")
    ref_harness.load()
    from tiberate import CkksEngine as RefEngine
    from tiberate import Preset

    import tiberate_fhe_b200 as tb

    ref = RefEngine(getattr(Preset, preset), devices=["cuda:0"])
    ref.rng.key = [torch.tensor(KEY, dtype=torch.int64, device="cuda:0")]
    ref.rng.nonce = [torch.tensor(NONCE, dtype=torch.int64, device="cuda:0")]
    ref.rng.initialize_states(0)
    want = _run(ref, {})

    ours = tb.CkksEngine(int(preset[4:]), devices=["cuda:0"], seed=KEY, nonce=NONCE)
    assert ours.ctx.q == [int(v) for v in ref.ckksCfg.q]
    got = _run(ours, {})
    for name in want:
        a, b = _tensors(want[name], []), _tensors(got[name], [])
        assert len(a) == len(b) and len(a) > 0, name
        for i, (x, y) in enumerate(zip(a, b)):
            assert x.shape == y.shape, f"{name}[{i}] shape {tuple(x.shape)} vs {tuple(y.shape)}"
            assert torch.equal(x, y), f"{name}[{i}] differs from the reference engine"
    assert torch.equal(ref.rng.states[0], ours.rng.states[0]), "CSPRNG consumption differs"


@pytest.mark.parametrize("preset", ["logN14", "logN15", "logN16"])
def test_readme_scenario_matches_the_reference_engine(preset):
    """encodecrypt -> pc_mult -> pc_add -> cc_mult -> rescale -> cc_add -> rotate_single ->
    decryptcode, float messages in, float messages out, on the reference engine (own extension) and on
    ours (own codec, CSPRNG, key generation and fused kernels) from the same ChaCha20 key: every
    ciphertext tensor and the decoded message must be identical."""
    import numpy as np
    import torch

    from baseline import ref_harness

    if not ref_harness.available():
        pytest.skip("baseline/_ref (reference install) not present")
    ref_harness.load()
    from tiberate import CkksEngine as RefEngine
    from tiberate import Preset
    from tiberate.typing import Plaintext as RefPlaintext

    import tiberate_fhe_b200 as tb
    from tiberate_fhe_b200.typing import Plaintext

    ref = RefEngine(getattr(Preset, preset), devices=["cuda:0"])
    ref.rng.key = [torch.tensor(KEY, dtype=torch.int64, device="cuda:0")]
    ref.rng.nonce = [torch.tensor(NONCE, dtype=torch.int64, device="cuda:0")]
    ref.rng.initialize_states(0)
    ours = tb.CkksEngine(int(preset[4:]), devices=["cuda:0"], seed=KEY, nonce=NONCE)
    data = torch.randn(ours.num_slots, generator=torch.Generator().manual_seed(1234), dtype=torch.float64)
    half = data[: ours.num_slots // 3]  # exercises padding

    def scenario(engine, PT):
        rec = {}
        _ = engine.sk, engine.pk, engine.evk
        rotk1 = engine.rotk[1]
        ct = engine.encodecrypt(data)
        rec["encodecrypt"] = ct
        rec["encodecrypt_padded_l2"] = engine.encodecrypt(half, level=2)
        pt = PT(data)
        ct2 = engine.pc_mult(pt, ct)
        rec["pc_mult"] = ct2
        ct3 = engine.pc_add(pt, ct2)
        rec["pc_add"] = ct3
        ct4 = engine.cc_mult(ct3, ct3)
        rec["cc_mult_relin"] = ct4
        rec["rescale"] = engine.rescale(ct4)
        ct5 = engine.cc_add(ct4, ct4)
        rec["cc_add"] = ct5
        ct6 = engine.rotate_single(ct5, rotk1)
        rec["rotate_single"] = ct6
        trip = engine.cc_mult(ct6, ct6, post_relin=False)
        rec["triplet"] = trip
        rec["encode_plain"] = engine.encode(half, level=1)
        rec["level_up"] = engine.level_up(ct, 3)
        rec["rotate_offset_5"] = engine.rotate_offset(ct5, 5)  # makes rotk[4] on the way: same CSPRNG draws
        rec["negate"] = engine.negate(ct5)
        dec = [engine.decryptcode(ct6, is_real=True), engine.decryptcode(trip), engine.decryptcode(ct)]
        torch.cuda.synchronize()
        return rec, [np.asarray(d) for d in dec]

    want, wdec = scenario(ref, RefPlaintext)
    got, gdec = scenario(ours, Plaintext)
    for name in want:
        a, b = _tensors(want[name], []), _tensors(got[name], [])
        assert len(a) == len(b) and len(a) > 0, name
        for i, (x, y) in enumerate(zip(a, b)):
            assert x.shape == y.shape and torch.equal(x, y), f"{name}[{i}] differs from the reference engine"
    for i, (x, y) in enumerate(zip(wdec, gdec)):
        assert x.shape == y.shape and np.array_equal(x, y), f"decryptcode output {i} differs"
    assert torch.equal(ref.rng.states[0], ours.rng.states[0]), "CSPRNG consumption differs"
    # and the numbers mean what they should: ((d*d + d)^2 * 2) rotated by one slot
    d = data.numpy()
    expect = np.roll(2 * (d * d + d) ** 2, -1)
    err = np.abs(gdec[0] - expect)
    assert min(np.abs(gdec[0] - np.roll(expect, s)).max() for s in (0, 1, 2)) < 1e-3 * np.abs(expect).max(), err.max()
    assert np.abs(gdec[2].real - d).max() < 1e-6


def test_self_contained_round_trip():
    """keygen -> encrypt -> cc_mult+relin -> rotate -> decrypt on our engine alone: the decrypted
    integers are (m * m rotated by 1 slot step in coefficient-embedding terms) up to noise; checked
    through the oracle-free invariant  Dec(Enc(m)) ~ m  and  Dec(Enc(a) + Enc(b)) ~ a + b."""
    import torch

    import tiberate_fhe_b200 as tb

    eng = tb.CkksEngine(14, devices=["cuda:0"], seed=KEY, nonce=NONCE)
    N = eng.N
    a, b = _message(N, "cuda:0"), torch.flip(_message(N, "cuda:0"), dims=[0])
    ca, cb = eng.encrypt([a]), eng.encrypt([b])
    scale = float(1 << 40)
    # decrypt returns round(value * scale / q_level ...) in the base prime's signed range: the encoded
    # integer message times scale, divided by the first prime of the level (engine final_scalar)
    da = eng.decrypt(ca)[0].reshape(-1).double()
    q0 = float(eng.ctx.q[0])
    assert torch.allclose(da, a.double() * scale / q0, atol=64.0), "Dec(Enc(a))"
    dsum = eng.decrypt(eng.cc_add(ca, cb))[0].reshape(-1).double()
    assert torch.allclose(dsum, (a + b).double() * scale / q0, atol=128.0), "Dec(Enc(a) + Enc(b))"
    # multiplication.  The library keeps ciphertexts at scale^2 (encode at scale, encrypt multiplies by
    # scale again) and cc_mult rescales both inputs first, so messages must themselves sit at ~scale:
    # Dec(Enc(a) * Enc(x)) = a x scale^2 / (q0^2 q1) for the negacyclic product a x.
    va = torch.randint(-512, 512, (N,), generator=torch.Generator().manual_seed(7), dtype=torch.int64).to("cuda:0")
    am = va * (1 << 40)
    xm = torch.zeros(N, dtype=torch.int64, device="cuda:0")
    xm[1] = 3 << 40  # 3 X at scale
    prod = eng.cc_mult(eng.encrypt([am]), eng.encrypt([xm]))  # pre-rescale + relinearize with the engine's own evk
    assert prod.level == 1
    dp = eng.decrypt(prod)[0].reshape(-1).double()
    want = 3.0 * torch.roll(va.double(), 1)
    want[0] = -want[0]
    q1 = float(eng.ctx.q[1])
    est = want * scale ** 4 / (q0 * q0 * q1)
    assert (dp - est).abs().max() < 1.0e-6 * est.abs().max(), "Dec(Enc(a) * Enc(3X))"
    # rotation by one slot = automorphism X -> X^3 on the coefficients
    rot = eng.rotate_single(eng.encrypt([am]), eng.rotk[1])
    dr = eng.decrypt(rot)[0].reshape(-1).double()
    idx = (torch.arange(N, device="cuda:0") * 3) % (2 * N)
    wr = torch.zeros(N, dtype=torch.float64, device="cuda:0")
    wr[idx % N] = torch.where(idx >= N, -va.double(), va.double())
    assert (dr - wr * scale * scale / q0).abs().max() < 1.0e-6 * scale, "Dec(rotate(Enc(a)))"


def test_two_engines_with_different_prime_chains_share_a_gpu():
    """key generation / encryption / decryption must use the engine's own context, not the
    device-current one.  Engine B (another prime chain, same N) is created after engine A; A must still
    encrypt and decrypt correctly, and so must B."""
    import torch

    import tiberate_fhe_b200 as tb
    from oracle.context import toy_primes

    a = tb.CkksEngine(14, devices=["cuda:0"], seed=list(range(8)), nonce=[1, 2])
    b = tb.CkksEngine(dict(logN=14, q=toy_primes(14, 5, 2), num_special_primes=2), devices=["cuda:0"],
                      seed=list(range(8, 16)), nonce=[3, 4])
    assert a.ctx.q != b.ctx.q and a.N == b.N
    gen = torch.Generator().manual_seed(9)
    m = torch.randn(a.num_slots, generator=gen, dtype=torch.float64)
    for eng in (a, b, a):
        ct = eng.encodecrypt(m)
        sq = eng.cc_mult(ct, ct)
        dec = torch.as_tensor(eng.decryptcode(sq, is_real=True), dtype=torch.float64)[: m.numel()]
        assert (dec - m * m).abs().max().item() < 1e-4, "engine mixed its parameter set with another engine's"
