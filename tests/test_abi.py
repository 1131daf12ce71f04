"""The C-ABI library loads on a GPU-less box and exports every symbol include/roboschool_b200.h
declares; compute entry points fail loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "roboschool_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_\w+)\s*\(", src)))


def test_exports_every_declared_symbol():
    from roboschool_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from roboschool_b200 import _lib
    from roboschool_b200.envspec import load_env_model
    from roboschool_b200.world import BatchedWorld
    assert _lib.lib().rs_device_count() == 0
    with pytest.raises(_lib.RoboschoolB200Error, match="no CUDA device"):
        BatchedWorld(load_env_model("RoboschoolHopper-v1"), 4, 0)
    from roboschool_b200.vec_env import VecEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VecEnv("RoboschoolHopper-v1", 4)


def test_product_does_not_import_oracle():
    """Nothing under roboschool_b200/ may reference oracle/ or tests/ (the oracle is the checker only)."""
    pkg = os.path.join(ROOT, "roboschool_b200")
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "liboracle" not in txt and "host_emul" not in txt.replace("tests/host_emul", ""), f
