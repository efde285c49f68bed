"""Loads the UNMODIFIED reference (pip-installed under baseline/_ref, built for sm_100) so that it
can be driven through its own public API: used by tests/golden/make_ref_golden.py (golden vectors
produced by the reference's CUDA extension on a B200) and by baseline/bench_reference_ext.py.

Nothing here is on the product path.  `prepare()` is run once in the build container (CPU only):
it puts the built extension where the reference's loader looks for it and pre-computes the prime
tables with the reference's own generator (sequentially: its joblib path cannot re-import the
package inside worker processes in this image).
"""

from __future__ import annotations

import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
PKG = os.path.join(REF, "tiberate")


def available() -> bool:
    return os.path.isdir(PKG) and bool(glob.glob(os.path.join(PKG, "libs", "torchops", "_ops*.so")))


def _paths():
    for p in (os.path.join(HERE, "stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)


def prepare(logNs=(14, 15, 16)) -> None:
    """Fix the install layout and pre-generate the prime pickles (needs no GPU)."""
    # scikit-build's wheel.install-dir put the extension under tiberate/tiberate/libs/...
    for sub in ("torchops", "utils"):
        src = os.path.join(PKG, "tiberate", "libs", sub)
        dst = os.path.join(PKG, "libs", sub)
        if os.path.isdir(src):
            os.makedirs(dst, exist_ok=True)
            for f in glob.glob(os.path.join(src, "*.so")):
                if not os.path.exists(os.path.join(dst, os.path.basename(f))):
                    shutil.copy2(f, dst)
    _paths()
    import pickle
    import types

    # import the pure-Python prime generator without triggering tiberate/__init__ (which loads CUDA ops)
    if "tiberate" not in sys.modules:
        m = types.ModuleType("tiberate")
        m.__path__ = [PKG]
        sys.modules["tiberate"] = m
        fake = True
    else:
        fake = False
    try:
        import tiberate.config.ckks_config  # noqa: F401  (must precede generate_primes: circular import)
    except Exception:
        pass
    from tiberate.utils import generate_primes as gp

    cache = os.path.join(PKG, "utils", "scale_primes.pkl")
    if not os.path.exists(cache):
        res = {}
        for logN in (12, 13, 14, 15, 16, 17):
            if logN not in logNs:
                continue
            N = 2 ** logN
            res[(40, N)] = gp.pgen_pseq(40, N, 64 if logN < 16 else 128)
        with open(cache, "wb") as f:
            pickle.dump(res, f)
    gp.generate_message_primes()  # writes message_special_primes.pkl beside the module
    if fake:
        for k in [k for k in sys.modules if k == "tiberate" or k.startswith("tiberate.")]:
            del sys.modules[k]


def load():
    """import tiberate (the reference) with the stubs on sys.path; needs a CUDA device."""
    if not available():
        raise RuntimeError("baseline/_ref is not prepared (see DESIGN.md: reference install)")
    _paths()
    import tiberate

    return tiberate


if __name__ == "__main__":
    prepare()
    print("prepared:", available())
