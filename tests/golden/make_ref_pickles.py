"""Writes tests/golden/ref_*.pkl with the REFERENCE's own classes and its own DataStruct.save
(tiberate/typing.py:283-290), imported from /root/reference in the build container (needs no GPU: only
typing.py is loaded, under a synthetic `tiberate` package so that tiberate/__init__ does not pull the CUDA
ops in).  tests/test_engine_host.py reads the files back through tiberate_fhe_b200.typing.load_reference_pickle.

  python tests/golden/make_ref_pickles.py
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TIBERATE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "baseline", "stubs"))


def main():
    import torch

    try:
        import loguru  # noqa: F401
    except ImportError:  # the reference logs through loguru; a stand-in is enough for typing.py
        m = types.ModuleType("loguru")
        m.logger = types.SimpleNamespace(debug=print, info=print, warning=print, error=print)
        sys.modules["loguru"] = m
    pkg = types.ModuleType("tiberate")
    pkg.__path__ = [os.path.join(REF, "tiberate")]
    sys.modules["tiberate"] = pkg
    import importlib

    T = importlib.import_module("tiberate.typing")
    g = torch.Generator().manual_seed(7)
    N, L, P = 16, 3, 5

    def t(rows):
        return torch.randint(0, 1 << 40, (rows, N), dtype=torch.int64, generator=g)

    ct = T.Ciphertext(data=[[t(L)], [t(L)]], level=2, logN=4, creator_hash="abc", misc={"note": "from the reference"})
    ct.save(os.path.join(HERE, "ref_ciphertext.pkl"))
    flags = T.FLAGS.INCLUDE_SPECIAL | T.FLAGS.MONTGOMERY_STATE | T.FLAGS.NTT_STATE
    parts = [T.PublicKey(data=[[t(P)], [t(P)]], flags=flags, level=0, logN=4) for _ in range(2)]
    rk = T.RotationKey(data=parts, flags=flags, level=0, delta=3, logN=4)
    rk.save(os.path.join(HERE, "ref_rotation_key.pkl"))
    torch.save({"ct": [x[0] for x in ct.data], "rk": [[p.data[0][0], p.data[1][0]] for p in parts]},
               os.path.join(HERE, "ref_pickles_expected.pt"))
    print("wrote", [f for f in os.listdir(HERE) if f.startswith("ref_") and (f.endswith(".pkl") or f.endswith(".pt"))])


if __name__ == "__main__":
    main()
