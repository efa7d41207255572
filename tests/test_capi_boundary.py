"""CPU tests of the drop-in boundary: libdtr_b200.so loads, exports every symbol include/dtr_b200.h
declares, the ctypes mirror binds all of them, and -- with no GPU -- creation fails loudly
instead of falling back to any CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dtr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dtr_b200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from dtrenderer_b200 import api
    lib = ctypes.CDLL(api.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dtr_b200.h but not exported"


def test_python_binding_covers_the_header(built):
    from dtrenderer_b200 import api
    api.load_library()
    assert sorted(s[0] for s in api.SYMBOLS) == _declared()


def test_product_library_does_not_link_the_oracle(built):
    """The oracle is test infrastructure: the shipped module must not reference it."""
    from dtrenderer_b200 import api
    out = subprocess.run(["nm", "-D", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "dtro_" not in out
    ldd = subprocess.run(["ldd", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "libdtr_oracle" not in ldd and "libdtr_ref" not in ldd
    for root, _, files in os.walk(os.path.join(ROOT, "dtrenderer_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("the CPU oracle", ""), f"{f} mentions the oracle"


def test_sm100a_cubin_is_embedded(built):
    from dtrenderer_b200 import api
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_create_fails_loudly_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dtrenderer_b200 import api
    with pytest.raises(api.DtrError):
        api.Renderer(64, 64, 1, 0)
    lib = api.load_library()
    h = ctypes.c_void_p()
    assert lib.dtr_b200_create(0, 64, 64, 1, ctypes.byref(h)) == -2  # DTR_B200_ERR_CUDA
    assert lib.dtr_b200_last_error(None)
    assert lib.dtr_b200_create(0, 0, 64, 1, ctypes.byref(h)) == -1   # DTR_B200_ERR_ARG comes first
