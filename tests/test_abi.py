"""CPU-side checks of the drop-in boundary: libdogstep.so loads and exports every symbol that
include/dogstep.h declares; the product package never touches the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "exploring-muzero-on-dog_b200")


def _declared():
    names = []
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(dogstep_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(PKG, "libdogstep.so"))
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.dogstep_version() >= 100


def test_invalid_arguments_are_rejected_without_a_gpu():
    import __graft_entry__ as g
    g.build()
    from exploring_muzero_on_dog_b200 import _lib
    lib = _lib.lib()
    cfg = _lib.MadnCfg(7, 0xF, 10, 0)
    st = _lib.MadnDetState()
    assert lib.dogstep_madn_det_valid_action(ctypes.byref(st), ctypes.c_int64(4), ctypes.byref(cfg), None, None) == -1
    cfg = _lib.MadnCfg(4, 0xF, 13, 0)
    assert lib.dogstep_madn_det_valid_action(ctypes.byref(st), ctypes.c_int64(4), ctypes.byref(cfg), None, None) == -2
    cfg = _lib.MadnCfg(4, 0xF, 10, 0)
    assert lib.dogstep_madn_det_valid_action(ctypes.byref(st), ctypes.c_int64(4), ctypes.byref(cfg), None, None) == -1  # null leaves


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt or "_oracle" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_cpu_tensors_fail_loudly():
    import pytest
    import torch
    from exploring_muzero_on_dog_b200 import _lib
    with pytest.raises(_lib.DogstepError):
        _lib.ptr(torch.zeros(4))


def test_entry_points_follow_the_device_of_their_pointers():
    """_lib._Entry: pointers of two GPUs in one call raise; the device is read off tagged structs and device pointers"""
    import pytest
    from exploring_muzero_on_dog_b200 import _lib
    a, b = _lib._DevPtr(16), _lib._DevPtr(32)
    a.dev, b.dev = 0, 1
    calls = []
    e = _lib._Entry(lambda *args: calls.append(args) or 0, "fake")
    with pytest.raises(_lib.DogstepError, match="different GPUs"):
        e(a, b)
    st = _lib.MadnDetState()
    st.dev = 1
    with pytest.raises(_lib.DogstepError, match="different GPUs"):
        e(ctypes.byref(st), a)
    assert e(ctypes.c_int64(3), None) == 0 and len(calls) == 1   # nothing device-side: passed straight through
