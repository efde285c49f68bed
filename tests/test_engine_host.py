"""Host-side logic of the engine mirror that needs no GPU: rotation-offset
decomposition and the on-disk format of the data structures."""

import os
import pickle

import pytest
import torch

from tiberate_fhe_b200.keygen import decompose_rot_offsets, decompose_with_power_of_2
from tiberate_fhe_b200.typing import FLAGS, Ciphertext, DataStruct, PublicKey, RotationKey


def test_power_of_two_decomposition():
    assert decompose_with_power_of_2(5, 16) == [1, 4]
    assert decompose_with_power_of_2(-3, 16) == [1, 4, 8]  # 13
    assert decompose_with_power_of_2(0, 16) == []
    with pytest.raises(AssertionError):
        decompose_with_power_of_2(3, 12)


def test_rotation_offsets_prefer_existing_keys_and_never_exceed_binary_length():
    slots = 8192
    assert decompose_rot_offsets(5, slots, {1: None}) == [1, 4]
    assert decompose_rot_offsets(7, slots, {1: None, 6: None}) == [1, 6]          # 2 steps instead of 3
    assert decompose_rot_offsets(96, slots, {}) == [32, 64]
    assert decompose_rot_offsets(100, slots, {100: None}) == [100]
    assert decompose_rot_offsets(-1, slots, {}) == [1 << i for i in range(13)]   # no negative steps: binary of 8191
    for off in range(1, 200):
        path = decompose_rot_offsets(off, slots, {3: None, 17: None})
        assert sum(path) == off and len(path) <= bin(off).count("1")


def test_data_structures_round_trip_through_the_safe_file_format(tmp_path):
    pk = PublicKey(data=[[torch.arange(6).reshape(2, 3)], [torch.ones(2, 3, dtype=torch.int64)]],
                   flags=FLAGS.NTT_STATE | FLAGS.MONTGOMERY_STATE, level=0, logN=4)
    rk = RotationKey(data=[pk, pk], flags=FLAGS.INCLUDE_SPECIAL, level=0, delta=5)
    ct = Ciphertext(data=[[torch.arange(8).reshape(2, 4)], [torch.arange(8).reshape(2, 4) * 3]], level=3, logN=2)
    p1, p2 = os.path.join(tmp_path, "rk.tb200"), os.path.join(tmp_path, "ct.tb200")
    rk.save(p1)
    ct.save(p2)
    rk2 = RotationKey.load(p1)
    assert isinstance(rk2, RotationKey) and rk2.delta == 5 and rk2.has_flag(FLAGS.INCLUDE_SPECIAL)
    assert isinstance(rk2.data[1], PublicKey) and rk2.data[1].has_flag(FLAGS.NTT_STATE)
    assert torch.equal(rk2.data[0].data[0][0], pk.data[0][0]) and rk2.data[0].misc["logN"] == 4
    ct2 = DataStruct.load(p2)
    assert isinstance(ct2, Ciphertext) and ct2.level == 3 and torch.equal(ct2.data[1][0], ct.data[1][0])
    with pytest.raises(TypeError):
        RotationKey.load(p2)


def test_loading_never_unpickles_arbitrary_objects(tmp_path):
    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))

    p = os.path.join(tmp_path, "evil.tb200")
    with open(p, "wb") as f:
        pickle.dump({"format": "tb200-datastruct-1", "payload": Evil()}, f)
    with pytest.raises(Exception):
        DataStruct.load(p)


def test_reference_pickles_load_without_executing_code(tmp_path):
    """files written by the reference's DataStruct.save (pickle of its own classes,
    tiberate/typing.py:283-290; fixtures made by tests/golden/make_ref_pickles.py from /root/reference) are read
    through an allow-list unpickler into this package's classes; anything else in the stream is refused."""
    import os
    import pickle

    import pytest
    import torch

    from tiberate_fhe_b200 import typing as T

    g = os.path.join(os.path.dirname(__file__), "golden")
    want = torch.load(os.path.join(g, "ref_pickles_expected.pt"), weights_only=True)
    ct = T.load_reference_pickle(os.path.join(g, "ref_ciphertext.pkl"))
    assert type(ct) is T.Ciphertext and ct.level == 2 and ct.misc["logN"] == 4 and ct.misc["note"] == "from the reference"
    assert ct.misc["absent"] is None and not ct.has_flag(T.FLAGS.NTT_STATE)
    assert torch.equal(ct.data[0][0], want["ct"][0]) and torch.equal(ct.data[1][0], want["ct"][1])
    rk = T.load_reference_pickle(os.path.join(g, "ref_rotation_key.pkl"))
    assert type(rk) is T.RotationKey and rk.delta == 3 and len(rk.data) == 2
    assert rk.has_flag(T.FLAGS.INCLUDE_SPECIAL) and rk.has_flag(T.FLAGS.NTT_STATE) and rk.has_flag(T.FLAGS.MONTGOMERY_STATE)
    for part, (b, a) in zip(rk.data, want["rk"]):
        assert type(part) is T.PublicKey and torch.equal(part.data[0][0], b) and torch.equal(part.data[1][0], a)
    # round trip into the package's own (pickle-free) format
    p = tmp_path / "ct.tb200"
    ct.save(str(p))
    assert torch.equal(T.Ciphertext.load(str(p)).data[1][0], want["ct"][1])

    class Evil:
        def __reduce__(self):
            import os as _os

            return (_os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))

    bad = tmp_path / "evil.pkl"
    with open(bad, "wb") as f:
        pickle.dump(Evil(), f)
    with pytest.raises(pickle.UnpicklingError):
        T.load_reference_pickle(str(bad))
    assert not (tmp_path / "pwned").exists()
