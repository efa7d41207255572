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


def test_band_partition_of_the_c_abi_matches_the_python_mirror(built):
    """dtr_b200_band_rows is THE partition (dtr_b200_gather_bands uses it); multigpu.band_rows must agree."""
    from dtrenderer_b200 import api, multigpu
    lib = api.load_library()
    th = lib.dtr_b200_tile_height()
    assert th in (24, 32)
    y0, y1 = ctypes.c_int(), ctypes.c_int()
    for h in (1, 31, 32, 33, 600, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                assert lib.dtr_b200_band_rows(h, world, r, ctypes.byref(y0), ctypes.byref(y1)) == 0
                assert (y0.value, y1.value) == multigpu.band_rows(h, world, r, th)
                rows.append((y0.value, y1.value))
            assert rows[0][0] == 0 and rows[-1][1] == h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert lib.dtr_b200_band_rows(1080, 2, 2, ctypes.byref(y0), ctypes.byref(y1)) == -1


def test_band_exchange_entry_points_fail_loudly_without_a_context(built):
    from dtrenderer_b200 import api
    lib = api.load_library()
    assert lib.dtr_b200_gather_bands(None, 0, 0) == -1
    assert lib.dtr_b200_band_barrier(None) == -1
    assert lib.dtr_b200_band_comm_init(None, None, 2, 0) == -1


def test_opaque_stage_selector_rejects_bad_arguments(built):
    """dtr_b200_set_opaque_stage: no context -> ERR_ARG; the header's three modes are what the binding names;
    without a context nothing ran, so dtr_b200_last_pass_deferred is 0."""
    import re
    from dtrenderer_b200 import api
    lib = api.load_library()
    for mode in (0, 1, 2, 3, -1):
        assert lib.dtr_b200_set_opaque_stage(None, mode) == -1
    assert lib.dtr_b200_last_pass_deferred(None) == 0
    hdr = open(os.path.join(ROOT, "include", "dtr_b200.h")).read()
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"DTR_B200_(OPAQUE_\w+)\s*=\s*(\d+)", hdr))
    assert enum == {"OPAQUE_SINGLE_KERNEL": api.OPAQUE_SINGLE_KERNEL, "OPAQUE_TWO_KERNELS": api.OPAQUE_TWO_KERNELS,
                    "OPAQUE_ONE_KERNEL": api.OPAQUE_ONE_KERNEL}


def test_library_has_no_link_time_nccl_dependency(built):
    """NCCL is bound with dlopen at the first band call; a host without NCCL can still load the module."""
    from dtrenderer_b200 import api
    ldd = subprocess.run(["ldd", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in ldd
