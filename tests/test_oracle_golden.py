"""The oracle against the outputs of the reference's own CUDA extension (fixtures generated on a
B200 by tests/golden/make_ref_golden.py).  This is the pin that makes the oracle trustworthy."""

import glob
import os

import pytest

import golden_check

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_logN*.json")))


@pytest.mark.skipif(not FILES, reason="no reference-extension fixtures committed yet")
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_reference_extension(path):
    n = golden_check.check_file_oracle(path)
    assert n >= 20
