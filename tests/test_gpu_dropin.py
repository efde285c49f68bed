"""Drop-in check (BASELINE.json configs[1]): the UNMODIFIED reference CkksEngine (pip-installed
under baseline/_ref with its own CUDA extension) runs its README scenario twice from the same CSPRNG
state -- once on its own operators, once with tiberate_fhe_b200 installed behind
tiberate.libs.wrapper (Montgomery, NTT, fused HE, CSPRNG and constant-pool operators alike) -- and every
integer tensor (keys, ciphertexts after encodecrypt, pc_mult,
pc_add, cc_mult+relin, rescale, cc_add, rotate_single) must be bit-identical; decrypted values too.
Skipped when baseline/_ref is absent (it is git-ignored; `pip install --target baseline/_ref` +
`python baseline/ref_harness.py`).
"""

import pytest

pytestmark = pytest.mark.gpu


def _flatten(x, out):
    import torch

    if isinstance(x, torch.Tensor):
        out.append(x.detach().clone())
    elif isinstance(x, (list, tuple)):
        for y in x:
            _flatten(y, out)
    elif hasattr(x, "data") and not isinstance(x, (int, float, str)):
        _flatten(x.data, out)
    return out


@pytest.mark.parametrize("fused", [False, True], ids=["ops", "fused"])
@pytest.mark.parametrize("preset", ["logN14", "logN15", "logN16"])
def test_reference_engine_on_tb200_backend_is_bit_identical(preset, fused):
    import torch

    from baseline import ref_harness

    if not ref_harness.available():
        pytest.skip("baseline/_ref (reference install) not present")
    ref_harness.load()
    from tiberate import CkksEngine, Preset
    from tiberate.typing import Plaintext

    from tiberate_fhe_b200 import backend, get_lib

    engine = CkksEngine(getattr(Preset, preset), devices=["cuda:0"])
    states0 = [s.clone() for s in engine.rng.states]
    gen = torch.Generator().manual_seed(1234)
    data = torch.randn(engine.num_slots, generator=gen, dtype=torch.float64)

    def scenario():
        for s, s0 in zip(engine.rng.states, states0):
            s.copy_(s0)
        for name in ("sk", "pk", "evk", "gk"):
            setattr(engine, f"_CkksEngine__{name}", None)
        engine._CkksEngine__rotk = {}
        rec = {}
        rec["sk"] = engine.sk
        rec["pk"] = engine.pk
        rec["evk"] = engine.evk
        rotk1 = engine.rotk[1]
        rec["rotk1"] = rotk1
        ct = engine.encodecrypt(data)
        rec["encodecrypt"] = ct.clone()
        pt = Plaintext(data)
        ct2 = engine.pc_mult(pt, ct)
        rec["pc_mult"] = ct2.clone()
        ct3 = engine.pc_add(pt, ct2)
        rec["pc_add"] = ct3.clone()
        ct4 = engine.cc_mult(ct3, ct3)
        rec["cc_mult_relin"] = ct4.clone()
        rec["rescale"] = engine.rescale(ct4).clone()
        ct5 = engine.cc_add(ct4, ct4)
        rec["cc_add"] = ct5.clone()
        ct6 = engine.rotate_single(ct5, rotk1)
        rec["rotate_single"] = ct6.clone()
        rec["triplet"] = engine.cc_mult(ct6, ct6, post_relin=False).clone()
        dec = engine.decryptcode(ct6, is_real=True)
        torch.cuda.synchronize()
        import numpy as np

        dec = dec.detach().cpu().numpy() if isinstance(dec, torch.Tensor) else np.asarray(dec)
        return {k: _flatten(v, []) for k, v in rec.items()}, torch.from_numpy(np.array(dec, dtype=np.float64))

    ref_out, ref_dec = scenario()
    lib = get_lib()
    before = lib.tb200_launch_count()
    backend.install_as_tiberate_backend(engine, fused=fused)
    try:
        our_out, our_dec = scenario()
        if fused:  # the hot methods are the fused entry points and still hand out the reference's own types
            from tiberate.typing import Ciphertext as RefCiphertext

            r = engine.rescale(engine.encodecrypt(data))
            assert isinstance(r, RefCiphertext) and r.data[0][0].storage_offset() == engine.ckksCfg.N, \
                "rescale must return storage-offset views like the reference (data[1:])"
            assert "rescale" in engine.__dict__ and "cc_mult" in engine.__dict__
    finally:
        backend.uninstall()
    assert "rescale" not in engine.__dict__
    assert lib.tb200_launch_count() - before > 100, "the tb200 operators were not used"
    for name in ref_out:
        a, b = ref_out[name], our_out[name]
        assert len(a) == len(b) and len(a) > 0, name
        for i, (x, y) in enumerate(zip(a, b)):
            assert x.shape == y.shape, (name, i)
            if not torch.equal(x, y):
                bad = (x != y).nonzero()
                raise AssertionError(f"{name}[{i}]: {bad.shape[0]} residues differ, first at {bad[0].tolist()}: "
                                     f"reference {x[tuple(bad[0])].item()} ours {y[tuple(bad[0])].item()}")
    assert torch.equal(ref_dec, our_dec), "decrypted values differ"
    # and the decrypted result is the expected function of the data within the reference's error bars
    want = data * data + data
    want = want * want
    want = want + want
    got = ref_dec[: data.numel()].double()
    errs = [(got - torch.roll(want, sh)).abs().max().item() for sh in (-1, 1)]  # rotk[1]: one slot
    assert min(errs) < 1e-3 * want.abs().max().item(), errs


def test_torch_ops_namespace_calls_the_same_kernels():
    """torch.ops.tb200_* (tiberate_fhe_b200/torchops.py) against the wrapper functions they bind."""
    import torch

    from tiberate_fhe_b200 import Tb200Context, get_lib, torchops, wrapper
    from tiberate_fhe_b200.presets import PRESETS

    torchops.register()
    q, K = PRESETS[14]["q"], PRESETS[14]["K"]
    ctx = Tb200Context(14, q, K)
    wrapper.set_context(ctx)
    lib = get_lib()
    gen = torch.Generator(device="cuda").manual_seed(3)
    a = torch.stack([torch.randint(0, int(qi), (ctx.N,), device="cuda", generator=gen) for qi in q])
    b = torch.stack([torch.randint(0, int(qi), (ctx.N,), device="cuda", generator=gen) for qi in q])
    before = lib.tb200_launch_count()
    got = torch.ops.tb200_mont_ops.mont_mult([a], [b], 0)[0]
    assert torch.equal(got, wrapper.mont_ops.mont_mult([a], [b], 0)[0])
    x, y = a.clone(), a.clone()
    assert torch.ops.tb200_ntt2_ops.enter_ntt_radix2([x], [], [], [], 0) is None
    wrapper.ntt2_ops.enter_ntt_radix2([y], None, None, None, 0)
    assert torch.equal(x, y) and not torch.equal(x, a)
    torch.ops.tb200_ntt2_ops.intt_radix2_exit_reduce([x], [], [], [], 0)
    assert torch.equal(x, a)
    got = torch.ops.tb200_he_ops.pc_add_fused([a], [b], 0)[0]
    assert torch.equal(got, wrapper.he_ops.pc_add_fused([a], [b], 0)[0])
    assert lib.tb200_launch_count() - before >= 7
    ctx.close()
