"""The CUDA library must load without a GPU and export every symbol include/tb200.h declares
(no compute calls here)."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from tiberate_fhe_b200 import _native

    assert sorted(_native.SIGNATURES) == _declared()


def test_cuda_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    if not os.path.exists(g.LIB):
        g.build()
    from tiberate_fhe_b200 import _native

    lib = _native.Lib(g.LIB)  # dlopen; getattr on each symbol happens in the constructor
    assert b"sm_100a" in lib.tb200_version()
    assert lib.tb200_launch_count() == 0
    # argument validation happens before any CUDA call
    assert lib.tb200_ctx_create(0, 3, 1, 1, None, 40) is None
    assert b"logN" in lib.tb200_last_error()


def test_library_contains_sm_100a_code_only():
    import shutil
    import subprocess

    import __graft_entry__ as g

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(g.LIB) or not os.path.exists(cuobjdump):
        pytest.skip("library or cuobjdump missing")
    out = subprocess.run([cuobjdump, "-lelf", g.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tiberate_fhe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_torch_dispatcher_registration_mirrors_the_reference_schemas():
    """SURVEY 8b / VERDICT r01: the op layer is also reachable as torch custom ops under our own namespaces,
    with the reference's schemas (names, argument lists, mutability, returns).  Registration needs no GPU."""
    import torch

    from tiberate_fhe_b200 import backend, torchops

    ops = torchops.register()
    assert torchops.register() == ops  # idempotent
    for module, names in ops.items():
        assert sorted(names) == sorted(backend._HOT[module]), f"{module}: registered ops != the re-pointed wrapper set"
        ns = getattr(torch.ops, torchops.namespace(module))
        for name in names:
            schema = str(getattr(ns, name).default._schema)
            assert schema == f"{torchops.namespace(module)}::{name}{torchops.SCHEMAS[module][name]}"
    # there is no CPU kernel: a host tensor is refused by the dispatcher, not silently computed elsewhere
    import pytest

    with pytest.raises(NotImplementedError):
        z = torch.zeros(2, 8, dtype=torch.int64)
        torch.ops.tb200_mont_ops.mont_mult([z], [z], 0)
