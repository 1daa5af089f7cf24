"""CPU-only: libpmgrav.so loads without a GPU and exports every symbol include/pmgrav.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'pmgrav.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pm_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from concept_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in pmgrav.h but not exported'
    assert sorted(_lib.SIGNATURES) == declared, 'ctypes signature table out of sync with pmgrav.h'
    assert lib.pm_version().startswith(b'pmgrav')
    assert lib.pm_launch_count() == 0


def test_kick_params_struct_layout_matches_header():
    import ctypes
    from concept_b200._lib import KickParams
    # 4 ints + 4 doubles, natural alignment
    assert ctypes.sizeof(KickParams) == 4*4 + 4*8
    assert KickParams.contribution.offset == 16


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from concept_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path/'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'concept_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn), encoding='utf-8').read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
                assert 'pm_oracle' not in src, fn
