"""The C-ABI library loads on a machine without a GPU and exports every symbol include/mrd_b200.h
declares (no compute calls here)."""

import ctypes
import os
import re
from importlib import import_module

import mrd_b200  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mrd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 24
    raw = ctypes.CDLL(import_module("multimodal-rare-disease_b200._lib").library_path())
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/mrd_b200.h but not exported"


def test_binding_covers_header(lib):
    sigs = import_module("multimodal-rare-disease_b200._lib").SIGNATURES
    assert set(sigs) == set(_declared())


def test_abi_version_and_error_string(lib):
    assert lib.mrd_abi_version() == 1
    assert isinstance(lib.mrd_last_error(), bytes)
