"""The C-ABI library loads on a CPU-only box and exports every symbol include/cgpcm_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import cgpcm_b200
from cgpcm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'cgpcm_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cgpcm_[a-z_0-9]+)\s*\(', src)))


def test_header_symbols_are_exported():
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for nm in names:
        assert hasattr(L, nm), nm
    assert sorted(_lib.EXPORTS) == names


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason='checks the behaviour on a box without a GPU')
def test_no_cpu_fallback():
    with pytest.raises(cgpcm_b200.CgpcmError):
        cgpcm_b200.Engine(8, 8)
    with pytest.raises(cgpcm_b200.CgpcmError):
        cgpcm_b200.bvn_cdf(np.zeros(3), np.zeros(3), np.full(3, .5))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'cgpcm_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f
