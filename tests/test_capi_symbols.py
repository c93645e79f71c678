"""The C-ABI library loads and exports every symbol include/bags_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from parity_common import PKG, ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bags_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bags_[a-z0-9_]+)\s*\(", text)) - {"bags_alloc_fn"})


def test_library_exists_and_loads():
    import bags_capi
    L = bags_capi.lib()
    assert L.bags_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    import bags_capi
    raw = ctypes.CDLL(bags_capi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in bags_b200.h but not exported"
        assert s in bags_capi.SIGNATURES, f"{s} has no ctypes signature in bags_capi.SIGNATURES"
    for s in bags_capi.SIGNATURES:
        assert s in syms, f"{s} bound in bags_capi but not declared in the header"


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "bags_b200.h")).read()
    assert "torch" not in text.replace("no torch", "").lower() or "at::" not in text
    assert "std::" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def test_state_size_functions():
    import bags_capi
    L = bags_capi.lib()
    g1, g2 = L.bags_geom_bytes(1000), L.bags_geom_bytes(2000)
    assert 0 < g1 < g2
    assert L.bags_geom_bytes(1_000_000) < 150 * 1_000_000          # < 150 B / Gaussian of state
    i = L.bags_image_bytes(1920, 1080)
    assert 8 * 1920 * 1080 <= i < 9 * 1920 * 1080
    b1, b2 = L.bags_binning_bytes(0, 1_000_000, 1920, 1080), L.bags_binning_bytes(0, 2_000_000, 1920, 1080)
    assert 16 * 1_000_000 <= b1 < b2
    assert L.bags_backward_scratch_bytes(1000) >= 64 * 1000
    assert L.bags_sort_scratch_bytes(10_000_000) < 64 << 20


def test_invalid_arguments_fail_loudly_without_gpu():
    import bags_capi
    L = bags_capi.lib()
    rc = L.bags_sort_pairs_u32(-1, None, None, None, None, 0, 32, None, None)
    assert rc == -1
    assert b"invalid" in L.bags_last_error()
    with pytest.raises(bags_capi.BagsError):
        bags_capi.check(rc, "bags_sort_pairs_u32")
    rc = L.bags_mark_visible(5, None, None, None, None, None)
    assert rc == -1


def test_missing_library_is_an_error_not_a_fallback(monkeypatch):
    import bags_capi
    monkeypatch.setattr(bags_capi, "_lib", None)
    monkeypatch.setattr(bags_capi, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(bags_capi.BagsError):
        bags_capi.lib()
