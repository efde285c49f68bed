"""Small GPU workload for compute-sanitizer (memcheck / racecheck): toy rings through every kernel class of
the hot path -- exact and mod-q transforms, the fused key-switch core (TMA-staged tiles + mbarriers), ModUp,
ModDown, automorphism, batched + chunked calls -- each checked against the oracle.

  compute-sanitizer --tool memcheck  python tests/sanitizer_case.py
  compute-sanitizer --tool racecheck python tests/sanitizer_case.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import parity  # noqa: E402
from parity import Harness, Setup  # noqa: E402


def main():
    from tiberate_fhe_b200 import get_lib

    h = Harness(get_lib(), use_torch=True)
    for logN, ns, K in ((12, 3, 2), (13, 2, 2), (9, 4, 3)):
        s = Setup.toy(h, logN, ns, K, seed=logN, rot_deltas=(1,))
        try:
            parity.check_ntt(s, 0, True, 2)
            parity.check_pointwise(s, 0, True)
            parity.check_he_ops(s, 0)
            for mode in parity.ENGINE_MODES:
                parity.set_mode(s, mode)
                parity.check_engine(s, 0, ops=("rescale", "keyswitch", "rotate", "cc_mult", "pc_mult"))
            parity.set_mode(s, parity.ENGINE_MODES[-1])
            s.ctx.set_chunk(2)
            parity.check_engine(s, 1, batch=3, ops=("cc_mult", "rotate"))
        finally:
            s.close()
        print("ok", logN, flush=True)


if __name__ == "__main__":
    main()
