"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares;
without a GPU the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hmgb200 as hmg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = hmg.load()
    names = declared("hmg.h") + declared("hmg_introspect.h")
    assert len(names) > 40
    for name in names:
        assert hasattr(lib, name), name
    bound = set(hmg.PROTOTYPES) | set(hmg._lib.HOST_PROTOTYPES)
    assert set(names) == bound           # the ctypes table and the headers agree


def test_version():
    assert hmg.load().hmg_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mesh, sigma = hmg.inputs.checkerboard_problem(2, 2)
    with pytest.raises(hmg.HmgError, match="no CUDA device"):
        hmg.ImplicitFineGrid(mesh, 2, sigma)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the product may import, include or open it."""
    pkg = os.path.join(ROOT, "homogenization.jl_b200")
    pat = re.compile(r"(from|import)\s+oracle|#include[^\n]*oracle|oracle[/.]\w")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".jl")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f
    assert not pat.search(open(os.path.join(ROOT, "hmgb200.py")).read())
