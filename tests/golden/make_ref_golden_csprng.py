"""Generates the CSPRNG fixtures from the UNMODIFIED reference.
  python tests/golden/make_ref_golden_csprng.py cdt      (anywhere /root/reference is mounted, CPU)
      -> tests/golden/ref_cdt_sigma3.2.json from tiberate/rng/csprng/discrete_gaussian_sampler.py
  python tests/golden/make_ref_golden_csprng.py codec    (anywhere /root/reference is mounted, CPU)
      -> tests/golden/ref_codec.json: sha256 of the reference's slot permutations (utils/encoding.py
         prepost_perms) for logN 3..16 and of a CPU encode / decode of a fixed message at logN 10
  python tests/golden/make_ref_golden_csprng.py csprng   (GPU box, needs baseline/_ref)
      -> gpurun_out/ref_csprng.json (copy to tests/golden/) from the reference Csprng + its CUDA extension
"""

import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    what = sys.argv[1]
    if what == "cdt":
        spec = importlib.util.spec_from_file_location(
            "dgs", "/root/reference/tiberate/rng/csprng/discrete_gaussian_sampler.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        btree, _, size, depth = m.build_CDT_binary_search_tree(security_bits=128, sigma=3.2)
        lut = [str(int(v)) for v in btree.T.ravel()]
        with open(os.path.join(HERE, "ref_cdt_sigma3.2.json"), "w") as f:
            json.dump({"source": "reference build_CDT_binary_search_tree(128, 3.2); lows then highs", "size": int(size),
                       "depth": int(depth), "lut": lut}, f, indent=0)
    elif what == "codec":
        import hashlib
        import types

        import numpy as np
        import torch

        vd, vdc = types.ModuleType("vdtoys"), types.ModuleType("vdtoys.cache")

        class CachedDict(dict):  # the 30-line stand-in for the reference's missing dependency
            def __init__(self, f):
                super().__init__()
                self.f = f

            def __missing__(self, k):
                self[k] = v = self.f(*k) if isinstance(k, tuple) else self.f(k)
                return v

        vdc.CachedDict, vd.cache = CachedDict, vdc
        tbm, trm = types.ModuleType("tiberate"), types.ModuleType("tiberate.rng")
        trm.RandNumGen = object
        sys.modules.update({"vdtoys": vd, "vdtoys.cache": vdc, "tiberate": tbm, "tiberate.rng": trm})
        spec = importlib.util.spec_from_file_location("refenc", "/root/reference/tiberate/utils/encoding.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        out = {"source": "reference tiberate/utils/encoding.py on CPU", "perms": {}}
        for logN in range(3, 17):
            pre, post = ref.prepost_perms(1 << logN, device="cpu")
            out["perms"][str(logN)] = [hashlib.sha256(np.ascontiguousarray(t.numpy(), dtype=np.int64).tobytes()).hexdigest()
                                       for t in (pre, post)]
        N = 1 << 10
        g = torch.Generator().manual_seed(3)
        m = torch.randn(N // 2, generator=g, dtype=torch.float64) + 1j * torch.randn(N // 2, generator=g, dtype=torch.float64)
        enc = ref.encode(m, device="cpu", deviation=1.25, return_without_scaling=True)
        out["encode_logN10"] = [float(v) for v in enc[:8]]
        with open(os.path.join(HERE, "ref_codec.json"), "w") as f:
            json.dump(out, f, indent=0)
    elif what == "csprng":
        import golden_csprng

        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ref_csprng.json"), "w") as f:
            json.dump({"source": "reference Csprng + tiberate_csprng_ops (sm_100 build) on a B200; scenario in "
                                 "tests/golden/golden_csprng.py", "cases": golden_csprng.run_reference()}, f, indent=1)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
