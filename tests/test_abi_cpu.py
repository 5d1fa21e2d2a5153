"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/kyber_b200.h declares, and refuses to work without a GPU (no CPU fallback)."""
import os
import re

import pytest

from helpers import ROOT


def _pkg():
    import importlib

    return importlib.import_module("kyber-rs_b200")


def test_library_exports_every_declared_symbol():
    kb = _pkg()
    if not os.path.exists(kb.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    L = kb.load_library()
    header = open(os.path.join(ROOT, "include", "kyber_b200.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", header))
    declared -= {"kb_last_error"} - set(kb.binding.EXPORTS)
    assert declared == set(kb.binding.EXPORTS), declared ^ set(kb.binding.EXPORTS)
    for sym in declared:
        assert hasattr(L, sym), sym


def test_header_is_plain_c_and_the_c_harness_compiles(tmp_path):
    """include/kyber_b200.h is valid C99 and C++17 on its own, and the plain-C caller of tests/c compiles against it
    (it is linked and run by the -m gpu suite)."""
    import subprocess

    hdr = os.path.join(ROOT, "include", "kyber_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-x", "c++", hdr])
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-c", os.path.join(ROOT, "tests", "c", "abi_check.c"),
                           "-o", str(tmp_path / "abi_check.o")])


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    kb = _pkg()
    with pytest.raises(kb.KBError):
        kb.Context(0)
    with pytest.raises(kb.KBError):
        kb.host.Point.base().mul(kb.host.Scalar.one())


def test_product_does_not_touch_the_oracle():
    """Nothing under kyber-rs_b200/ may import, link or execute oracle/ (tier rule ③)."""
    pkg = os.path.join(ROOT, "kyber-rs_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="replace").read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(d, f)


def test_shard_ranges():
    kb = _pkg()
    for n in (0, 1, 7, 8, 1000, 2**20):
        for world in (1, 2, 3, 8):
            spans = [kb.sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_rust_sys_matches_header():
    """The (uncompiled) Rust -sys crate declares exactly the symbols of include/kyber_b200.h."""
    kb = _pkg()
    rs = open(os.path.join(ROOT, "kyber-rs_b200", "rust", "kyber-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (kb_[a-z0-9_]+)\(", rs))
    assert declared == set(kb.binding.EXPORTS), declared ^ set(kb.binding.EXPORTS)
