"""One CSPRNG scenario run by three implementations -- the reference's Csprng on its own CUDA extension
(run_reference, GPU box only, via baseline/_ref), the oracle's restatement (run_oracle, CPU) and
tiberate_fhe_b200.rng.Csprng on libtb200 (run_tb200, GPU) -- each returning {name: {sha256, head}}."""

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if HERE not in sys.path:
    sys.path.insert(0, HERE)

N, CH, REP = 4096, 3, 2
KEY = [0x9E3779B9, 0x7F4A7C15, 0xF39CC060, 0x5CEDC834, 0x1082276B, 0xF3A27251, 0xF86C6A11, 0xD0C18E95]
NONCE = [0x2545F491, 0x4F6CDD1D]
Q = [(1 << 40) - 87, (1 << 60) - 93, 1000003]


def _digest(a):
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int64)
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "shape": list(a.shape),
            "head": [int(v) for v in a.ravel()[:6]]}


def _coef():
    i = np.arange(N, dtype=np.float64)
    return np.sin(i * 0.37) * 1.0e5 + (i % 7) * 0.125 - 3.0


def _scenario(rng, to_np, coef_dev):
    out = {}
    out["randint_all"] = _digest(to_np(rng.randint([list(Q)], shift=0, repeats=0)[0]))
    out["ternary_rep"] = _digest(to_np(rng.randint(3, -1, repeats=1)[0]))
    out["randint_mixed"] = _digest(to_np(rng.randint([[Q[1], Q[2], 5, 7]], shift=0, repeats=2)[0]))
    out["gauss_rep"] = _digest(to_np(rng.discrete_gaussian(non_repeats=0, repeats=1)[0]))
    out["gauss_mixed"] = _digest(to_np(rng.discrete_gaussian(non_repeats=2, repeats=2)[0]))
    out["bytes_all"] = _digest(to_np(rng.randbytes()[0]))
    out["bytes_part"] = _digest(to_np(rng.randbytes(shares=[1], repeats=1)[0]))
    out["randround"] = _digest(to_np(rng.randround(coef_dev)))
    out["states_end"] = _digest(to_np(rng.states[0] if isinstance(rng.states, list) else rng.states))
    return out


def run_oracle():
    from oracle.csprng import OracleCsprng

    class Adapter(OracleCsprng):  # list-per-device calling convention of the reference class
        def randint(self, amax=3, shift=0, repeats=0):
            if isinstance(amax, (list, tuple)):
                amax = amax[0]
            return [OracleCsprng.randint(self, amax, shift, repeats)]

        def discrete_gaussian(self, non_repeats=0, repeats=1):
            return [OracleCsprng.discrete_gaussian(self, non_repeats, repeats)]

        def randbytes(self, shares=None, repeats=0):
            return [OracleCsprng.randbytes(self, None if shares is None else shares[0], repeats)]

    rng = Adapter(N, CH, REP, key=KEY, nonce=NONCE)
    return _scenario(rng, np.asarray, _coef())


def run_tb200():
    import torch

    from tiberate_fhe_b200.rng import Csprng

    rng = Csprng(N, [CH], REP, devices=["cuda:0"], seed=KEY, nonce=NONCE)
    return _scenario(rng, lambda t: t.cpu().numpy(), torch.from_numpy(_coef()).to("cuda:0"))


def run_reference():
    import contextlib
    import io

    import torch

    from baseline import ref_harness

    with contextlib.redirect_stdout(io.StringIO()):
        ref_harness.load()
        from tiberate.rng.csprng.csprng import Csprng
    rng = Csprng(N, [CH], REP, devices=["cuda:0"])
    rng.key = [torch.tensor(KEY, dtype=torch.int64, device="cuda:0")]
    rng.nonce = [torch.tensor(NONCE, dtype=torch.int64, device="cuda:0")]
    rng.initialize_states(0)
    return _scenario(rng, lambda t: t.cpu().numpy(), torch.from_numpy(_coef()).to("cuda:0"))


def check_against(path, runner):
    with open(path) as f:
        want = json.load(f)["cases"]
    got = runner()
    assert set(got) == set(want)
    for k in want:
        assert got[k] == want[k], f"csprng golden case {k}: {got[k]} != {want[k]}"
