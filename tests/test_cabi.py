"""CPU: the C-ABI shared library loads and exports every symbol include/apples_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from tests import util
from apples_b200 import _lib


def _header_symbols():
    src = open(os.path.join(util.ROOT, 'include', 'apples_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(apples_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol(lib):
    for name in _header_symbols():
        assert hasattr(lib, name), name


def test_row_geometry(lib):
    assert lib.apples_words_per_row(1) == 4 and lib.apples_words_per_row(128) == 4 and lib.apples_words_per_row(129) == 8
    assert lib.apples_words_per_row(5000) == 160 and lib.apples_words_per_row(1620) == 52
    assert lib.apples_aa_row_bytes(1638) == 1648
    from apples_b200 import fasta
    for L in (1, 31, 32, 33, 1500, 1620, 5000, 65535):
        assert lib.apples_words_per_row(L) == fasta.words_per_row(L)
        assert lib.apples_aa_row_bytes(L) == fasta.aa_row_bytes(L)


def test_params_struct_layout():
    p = _lib.make_params('BME', 'HYBRID', True, 7, 0.45, 0.01)
    assert ctypes.sizeof(p) == 32
    assert (p.method, p.criterion, p.negative_branch, p.base_observation_threshold) == (2, 2, 1, 7)
    assert _lib.make_params('nonsense', 'nonsense').method == _lib.OLS  # PoolQueryWorker.py:110-111


def test_no_device_is_an_error_not_a_fallback(lib):
    """Without a CUDA device context creation fails; the product has no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    h = ctypes.c_void_p()
    assert lib.apples_ctx_create(0, ctypes.byref(h)) != 0
    from apples_b200.placer import GpuPlacer
    from apples_b200.tree import BackboneTree
    with pytest.raises(RuntimeError):
        GpuPlacer(BackboneTree.from_newick('((A:1,B:1):1,C:1,D:1);'))
