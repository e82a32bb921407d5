"""The C-ABI shared library loads on a GPU-less host and exports every entry point include/akshar_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'akshar_b200.h'), encoding='utf-8').read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(akshar_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    from akshar_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n


def test_binding_lists_the_same_symbols():
    from akshar_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared()


def test_host_only_calls(lib):
    lib.akshar_version.restype = ctypes.c_int
    assert lib.akshar_version() >= 100
    lib.akshar_status_str.restype = ctypes.c_char_p
    assert lib.akshar_status_str(0) == b'ok'
    lib.akshar_workspace_bytes.restype = ctypes.c_size_t
    lib.akshar_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    a, b = lib.akshar_workspace_bytes(1 << 20, 1000), lib.akshar_workspace_bytes(1 << 24, 1000)
    assert 0 < a < b


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'akshar_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                txt = open(os.path.join(dirpath, f), encoding='utf-8').read()
                assert 'akshar_oracle' not in txt and 'oracle/' not in txt, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import akshar_b200
    with pytest.raises(RuntimeError):
        akshar_b200.normalize_text('Hello')
    with pytest.raises(RuntimeError):
        akshar_b200.aksharTokenizer().tokenize('नमस्ते')
