"""Generates the CSPRNG fixtures from the UNMODIFIED reference.
  python tests/golden/make_ref_golden_csprng.py cdt      (anywhere /root/reference is mounted, CPU)
      -> tests/golden/ref_cdt_sigma3.2.json from tiberate/rng/csprng/discrete_gaussian_sampler.py
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
