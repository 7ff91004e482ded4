"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and its host-side integer contract is bit-exact."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from albatross_b200 import capi
from oracle.oracle import Restate, group_keys
from tests.helpers import features

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "albatross_b200.h")).read()
    return sorted(set(re.findall(r"AB_API[^;(]*?\b(ab_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 40
    lib = capi.lib()
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/albatross_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == names
    assert lib.ab_version() == 1


def test_no_device_fails_loudly():
    """No CPU fallback: without a CUDA device ab_create must return AB_ERR_CUDA."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.AbError) as e:
        capi.Handle(0)
    assert e.value.status == 2
    assert "no CPU fallback" in str(e.value) or "cuda" in str(e.value).lower()


def test_null_arguments_are_rejected():
    lib = capi.lib()
    assert lib.ab_matrix_dims(None, None, None) == 1
    assert b"requirement failed" in lib.ab_last_error()
    n = C.c_int64()
    assert lib.ab_factor_rows(None, C.byref(n)) == 1


def test_group_indexers_bit_exact(golden):
    _, ref = golden
    x = ref["gp_x1"]
    for tag, gk, ga in (("grp", 1, 8), ("grp2", 2, 3.7)):
        keys, offsets, indices = capi.group_indexers(group_keys(x, gk, ga))
        assert np.array_equal(keys, ref[f"{tag}_keys"])
        assert np.array_equal(offsets, ref[f"{tag}_offsets"])
        assert np.array_equal(indices, ref[f"{tag}_indices"])
    # leave-one-out grouper, empty input, negative and repeated keys
    k, o, i = capi.group_indexers(np.arange(17))
    assert np.array_equal(k, np.arange(17)) and np.array_equal(o, np.arange(18))
    assert np.array_equal(i, np.arange(17))
    k, o, i = capi.group_indexers(np.array([], dtype=np.int64))
    assert len(k) == 0 and np.array_equal(o, [0]) and len(i) == 0
    rng = np.random.default_rng(0)
    item_keys = rng.integers(-5, 6, size=1000)
    got = capi.group_indexers(item_keys)
    want = Restate.group_indexers(item_keys)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    x = features(500, 1, 3)
    got = capi.group_indexers(group_keys(x, 1, 8))
    want = Restate.group_indexers(group_keys(x, 1, 8))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_bench_programs_match_the_test_menu():
    """capi.bench_program (what the timing tools and bench.py feed the library) builds the same postfix
    programs as the oracle-side menu used by the parity tests."""
    from oracle.oracle import menu_program

    def same(a, b):
        return list(a[0]) == list(b[0]) and [float(v) for v in a[1]] == [float(v) for v in b[1]]

    assert same(capi.bench_program("se_noise"), menu_program(6, [1.0, 1.0, 0.1]))
    assert same(capi.bench_program("se_m52"), menu_program(7, [2.0, 1.5, 3.0, 0.7]))
    assert same(capi.bench_program("se_m52_noise"), menu_program(8, [2.0, 1.5, 3.0, 0.7, 0.1]))
    with pytest.raises(ValueError):
        capi.bench_program("nope")
