"""The C-ABI library builds, loads without a GPU, and exports every symbol that
include/tonic_b200.h declares (no compute calls here)."""

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'tonic_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from tonic_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert lib.tb_version() == 1


def test_ctypes_prototypes_cover_the_header():
    from tonic_b200 import _lib
    assert set(_lib.exported_symbols()) == set(declared_symbols())
    _lib.load()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'tonic_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert '/root/reference' not in src, f
