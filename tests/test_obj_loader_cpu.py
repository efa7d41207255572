"""CPU tests of the host-side .obj loader (dtrenderer_b200/host/DTRAssetB200.h):
its number parsers against the reference's own Dqn_StrToF32 / Dqn_StrToI64 (exported from the
unmodified reference build, oracle/_ref), and its output against a hand-built DTRMesh
(DTRendererAsset.cpp:190-612: layout of the model block, 0-based indices, parser quirks)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dtrenderer_b200 import scenes
from oracle import dtro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_DIR = os.path.join(ROOT, "tests", "_shim")


@pytest.fixture(scope="module")
def shim(built):
    so = os.path.join(SHIM_DIR, "libobjloader_shim.so")
    src = os.path.join(SHIM_DIR, "obj_loader_shim.cpp")
    hdr = os.path.join(ROOT, "dtrenderer_b200", "host", "DTRAssetB200.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall",
                               "-I", os.path.join(ROOT, "dtrenderer_b200", "host"), "-I", os.path.join(ROOT, "include"),
                               "-o", so, src])
    lib = C.CDLL(so)
    lib.objshim_strtof32.restype = C.c_float
    lib.objshim_strtof32.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    lib.objshim_strtoi64.restype = C.c_int64
    lib.objshim_strtoi64.argtypes = [C.c_char_p, C.c_int]
    lib.objshim_load.restype = C.c_int
    lib.objshim_load.argtypes = [C.c_char_p, C.c_long, C.POINTER(C.c_long), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def load(shim, text):
    raw = text.encode("ascii")
    counts = (C.c_long * 6)()
    if not shim.objshim_load(raw, len(raw), counts, None, None, None, None, None):
        return None
    nV, nT, nN, nF, block, layout_ok = list(counts)
    v, t, n = np.zeros((nV, 4), np.float32), np.zeros((nT, 3), np.float32), np.zeros((nN, 3), np.float32)
    f, fc = np.zeros((nF, 9), np.int32), np.zeros((nF, 3), np.uint32)
    assert shim.objshim_load(raw, len(raw), counts, v.ctypes.data, t.ctypes.data, n.ctypes.data, f.ctypes.data, fc.ctypes.data)
    return dict(vertexes=v, texUV=t, normals=n, faces=f, counts=fc, block=block, layout_ok=bool(layout_ok))


def test_number_parsers_match_the_reference(shim):
    if not dtro.available("reference"):
        pytest.skip("oracle/_ref not built")
    ref = dtro._load("reference")
    if not hasattr(ref, "dtro_ref_strtof32"):
        pytest.skip("oracle/_ref predates the parser exports")
    ref.dtro_ref_strtof32.restype = C.c_float
    ref.dtro_ref_strtof32.argtypes = [C.c_char_p, C.c_int]
    ref.dtro_ref_strtoi64.restype = C.c_int64
    ref.dtro_ref_strtoi64.argtypes = [C.c_char_p, C.c_int]
    rng = np.random.default_rng(17)
    cases = ["0", "1", "-1", "0.5", "-0.333333", "123.456", "1.000000", "0.000001", "12345678", "3.14159265", "-0.0",
             "1e-3", "2.5e-2", "7e+2", "-4.25e+1", ".5", "5.", "0.1234567890", "99999.99999"]
    for _ in range(3000):
        digits = int(rng.integers(1, 9))
        s = "".join(str(int(d)) for d in rng.integers(0, 10, digits))
        if rng.random() < 0.8:
            k = int(rng.integers(0, len(s) + 1))
            s = s[:k] + "." + s[k:]
        if rng.random() < 0.5:
            s = "-" + s
        if rng.random() < 0.15:
            s += "e" + str(rng.choice(["-", "+"])) + str(int(rng.integers(0, 6)))
        cases.append(s)
    ok = C.c_int(0)
    for s in cases:
        b = s.encode()
        mine = shim.objshim_strtof32(b, len(b), C.byref(ok))
        theirs = ref.dtro_ref_strtof32(b, len(b))
        assert ok.value == 1
        assert np.float32(mine).view(np.uint32) == np.float32(theirs).view(np.uint32), s
    for s in ["0", "7", "123456789", "-42", "+17", "12/3", "", "x1"]:
        b = s.encode()
        assert shim.objshim_strtoi64(b, len(b)) == ref.dtro_ref_strtoi64(b, len(b)), s


def _obj_text(mesh, decimals=6):
    """The synthetic mesh written the way exporters write .obj files."""
    lines = ["# synthetic sphere", "g default"]
    lines += ["v " + " ".join(f"{x:.{decimals}f}" for x in v[:3]) for v in mesh["vertexes"]]
    lines += ["vt " + " ".join(f"{x:.{decimals}f}" for x in t[:2]) for t in mesh["texUV"]]
    lines += ["vn " + " ".join(f"{x:.{decimals}f}" for x in n) for n in mesh["normals"]]
    lines.append("s off")
    for f in mesh["faces"]:
        lines.append("f " + " ".join(f"{f[k] + 1}/{f[3 + k] + 1}/{f[6 + k] + 1}" for k in range(3)))
    return "\n".join(lines) + "\n"


def test_loader_reproduces_a_hand_built_mesh(shim):
    mesh = scenes.uv_sphere(12, 6)
    got = load(shim, _obj_text(mesh))
    assert got is not None and got["layout_ok"]
    nV, nT, nN, nF = len(mesh["vertexes"]), len(mesh["texUV"]), len(mesh["normals"]), len(mesh["faces"])
    assert got["vertexes"].shape[0] == nV and got["texUV"].shape[0] == nT and got["normals"].shape[0] == nN
    assert got["block"] == 16 * nV + 12 * nT + 12 * nN + 48 * nF + 36 * nF
    assert np.array_equal(got["faces"], mesh["faces"]) and np.all(got["counts"] == 3)
    assert np.all(got["vertexes"][:, 3] == 1.0)  # w defaults to 1 (DTRendererAsset.cpp:268)
    # values: what Dqn_StrToF32's algorithm makes of the six-decimal text -- int(all digits) * 0.1f^6
    def parse(x):
        s = f"{x:.6f}"
        raw = np.float32(int(s.replace("-", "").replace(".", "")))
        shift = np.float32(1.0)
        for _ in range(6):
            shift = np.float32(shift * np.float32(0.1))
        r = np.float32(raw * shift)
        return -r if s.startswith("-") else r
    want = np.array([[parse(x) for x in v[:3]] for v in mesh["vertexes"]], np.float32)
    assert np.array_equal(got["vertexes"][:, :3].view(np.uint32), want.view(np.uint32))
    assert np.allclose(got["vertexes"][:, :3], mesh["vertexes"][:, :3], atol=2e-6)
    assert np.allclose(got["normals"], mesh["normals"], atol=2e-6) and np.allclose(got["texUV"][:, :2], mesh["texUV"][:, :2], atol=2e-6)


def test_loader_quirks_and_rejections(shim):
    # an omitted attribute is skipped; a slash-less face is ONE vertex whose v/vt/vn are the three numbers
    got = load(shim, "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nf 1 2 3 2 3 1 3 1 2\n")
    assert got is not None
    assert got["counts"].tolist() == [[3, 0, 3], [3, 3, 3]]
    assert got["faces"][0].tolist() == [0, 1, 2, -1, -1, -1, 0, 0, 0]
    assert got["faces"][1].tolist() == [0, 1, 2, 1, 2, 0, 2, 0, 1]
    # unknown statements and comments are skipped, CRLF line ends are fine, a number list runs over line ends
    got = load(shim, "mtllib x.mtl\r\n# c\r\nusemtl a\r\nv 1 2\r\n3\r\nvt 0.5 0.25\r\n")
    assert got is not None and got["vertexes"].tolist() == [[1, 2, 3, 1]]
    # (25 * 0.1f * 0.1f is not 0.25f: the reference's parser multiplies 0.1f up once per decimal)
    assert got["texUV"][0, 0] == np.float32(0.5) and got["texUV"][0, 1] == np.float32(np.float32(25) * np.float32(np.float32(0.1) * np.float32(0.1)))
    # relative (negative) indices are not rejected by the reference's scanner, they are misread -- '-' is skipped as a
    # separator and the attribute types shift; the restatement misreads them the same way
    got = load(shim, "f -1/-1/-1 2/2/2 3/3/3\n")
    assert got is not None and got["faces"].tolist() == [[0, 1, 2, 0, 1, 2, 0, 1, 2]]
    # what the reference asserts on fails the load
    for bad in ["v 1 2 3 4\n", "f 0/1/1 2/2/2 3/3/3\n", "p 1\n", "l 1 2\n", "vx 1 2 3\n", "v 1.0 abc 2\n", "f 1/1/1 2/2/2\n"]:
        assert load(shim, bad) is None, bad
