"""CPU: the C-ABI library builds, loads and exports every function include/asz_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "asz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asz_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from alphasnake_zero_b200 import build, _lib
    build.build()
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), "libasz_b200.so does not export %s" % n
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) == bound, "ctypes binding and header disagree: %s" % (set(names) ^ bound)


def test_version_and_error_string():
    from alphasnake_zero_b200 import _lib
    L = _lib.lib()
    assert L.asz_version() == 2
    assert isinstance(L.asz_last_error(), bytes)


def test_engine_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alphasnake_zero_b200.engine import Engine, AszError
    with pytest.raises(AszError):
        Engine(games=4)
    # and the C ABI itself fails loudly as well
    from alphasnake_zero_b200 import _lib
    L = _lib.lib()
    cfg = _lib.Config(11, 4, 1, 0.15, 4, 0, 0, 0, 100.0, 0, 0, 0)
    h = C.c_void_p()
    rc = L.asz_engine_create(C.byref(h), C.byref(cfg))
    assert rc != 0 and b"no CUDA device" in L.asz_last_error()


def test_config_validation():
    from alphasnake_zero_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    for bad in (dict(side=10), dict(snakes=9), dict(games=0), dict(health_dec=-1)):
        kw = dict(side=11, snakes=4, health_dec=1, games=4)
        kw.update(bad)
        cfg = _lib.Config(kw["side"], kw["snakes"], kw["health_dec"], 0.15, kw["games"], 0, 0, 0, 100.0, 0, 0, 0)
        assert L.asz_engine_create(C.byref(h), C.byref(cfg)) == -1
