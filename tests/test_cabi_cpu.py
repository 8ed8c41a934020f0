"""CPU-side checks of the C-ABI boundary: the library builds, loads without a GPU and
exports every symbol include/nabo_b200.h declares; argument validation that needs no
device; the product refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

from nabo_b200 import _lib, build


@pytest.fixture(scope="module")
def L():
    build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(L):
    declared = _lib.declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(L, name), "libnabo_b200.so does not export %s" % name
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert L.nabo_abi_version() == 1


def test_argument_validation_without_device(L):
    # invalid arguments are rejected before any CUDA call
    rc = L.nabo_knn(None, 1, None, 1, 4, 0, 1, 1, 0, 0.25, None, 0, 0, 0, None, None, None, 0, None, None)
    assert rc == -1 and b"bad sizes" in L.nabo_last_error()
    rc = L.nabo_knn(None, 1, None, 1, 4, 4, 2, 1, 7, 0.25, None, 0, 0, 0, None, None, None, 0, None, None)
    assert rc == -1
    rc = L.nabo_merge_topk(None, None, 0, 1, 1, None, None, None)
    assert rc == -1
    assert L.nabo_scores_workspace_bytes(1000, 30, 5000) > 4 * 1000 * 30 * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nabo_b200 import core
    with pytest.raises(RuntimeError, match="no CUDA device"):
        core.knn(np.zeros((4, 3)), np.zeros((5, 3)), 2, mode="exact")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        core.euclidean_dist(np.zeros((4, 3)), np.zeros((5, 3)))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, fs in os.walk(os.path.join(root, "nabo_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_weight_lut_matches_python_round():
    from nabo_b200 import core
    lut = core.snn_weight_lut(11)
    assert lut[0] == 0.0 and lut[1] == round(1 / 19, 2) and lut[11] == round(11 / 9, 2)
    with pytest.raises(ZeroDivisionError):
        core.snn_weight_lut(2)
    assert core.fix_weight(11) == 0.5 / (20 - 0.5)
