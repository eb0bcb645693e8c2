"""The C-ABI library builds, loads and exports every symbol declared in include/tce_b200.h (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tce_b200.h")


@pytest.fixture(scope="module")
def lib():
    from tce_rl_b200 import _build, _lib
    _build.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tce_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in tce_b200.h but not exported"


def test_binding_covers_header(lib):
    from tce_rl_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_binding_argument_counts_match_header():
    """Every ctypes signature has exactly as many arguments as the declaration in the header (a short argtypes list
    makes ctypes silently pass garbage)."""
    from tce_rl_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    seen = 0
    for m in re.finditer(r"\b(?:int|size_t|void|const char \*)\s*\**(tce_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert len(_lib.SIGNATURES[name][1]) == n, name
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_version_and_errors(lib):
    assert lib.tce_version() >= 100
    assert lib.tce_strerror(0) == b"ok"
    assert b"unsupported" in lib.tce_strerror(-2)
    # argument validation happens before any CUDA call
    assert lib.tce_gae(None, None, None, None, 1.0, 0.95, 1, None, None, 4, 10, None) == -1
    assert lib.tce_prodmp_tables_create(None, None, None) == -1


def test_mp_cfg_layout():
    from tce_rl_b200._lib import MpCfg
    assert ctypes.sizeof(MpCfg) == 8 * 4 + 8 * 8


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tce_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "from .. import oracle" not in src and "/root/reference" not in src
