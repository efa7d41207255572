"""ctypes binding of oracle/dtro.h -- TEST INFRASTRUCTURE ONLY.

Loads either checker behind the same C interface:
  kind="reference": oracle/_ref/libdtr_ref.so          (unmodified reference, markers off)
  kind="reference_markers": oracle/_ref/libdtr_ref_markers.so  (reference default build)
  kind="port":      oracle/libdtr_oracle.so            (plain-C restatement)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (dtrenderer_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "reference": os.path.join(_HERE, "_ref", "libdtr_ref.so"),
    "reference_markers": os.path.join(_HERE, "_ref", "libdtr_ref_markers.so"),
    "port": os.path.join(_HERE, "libdtr_oracle.so"),
    # the reference's own types and call signatures driving the CUDA back end (drop-in proof)
    "reference_api_b200": os.path.join(_HERE, "_ref", "libdtr_ref_b200.so"),
}
_LIBS = {}

SHADE_FULLBRIGHT, SHADE_FLAT, SHADE_GOURAUD = 0, 1, 2

# stbtt_packedchar (external/stb_truetype.h:522-527) == dtro_packedchar == dtr_b200_packedchar
PACKEDCHAR = np.dtype([("x0", "<u2"), ("y0", "<u2"), ("x1", "<u2"), ("y1", "<u2"), ("xoff", "<f4"), ("yoff", "<f4"),
                       ("xadvance", "<f4"), ("xoff2", "<f4"), ("yoff2", "<f4")])
assert PACKEDCHAR.itemsize == 28

_f = C.POINTER(C.c_float)
_u8 = C.POINTER(C.c_uint8)
_i32 = C.POINTER(C.c_int32)


def premultiply_bitmap(rgba, kind="port"):
    """DTRAsset_LoadBitmap's premultiply pass on a copy of u8[..., 4] straight-alpha texels."""
    lib = _load(kind)
    a = np.ascontiguousarray(rgba, dtype=np.uint8).copy()
    lib.dtro_premultiply_bitmap.argtypes = [C.c_void_p, C.c_int]
    lib.dtro_premultiply_bitmap(a.ctypes.data_as(C.c_void_p), a.size // 4)
    return a


def available(kind):
    return os.path.exists(_PATHS[kind])


def _load(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    lib = C.CDLL(_PATHS[kind])
    lib.dtro_kind.restype = C.c_char_p
    lib.dtro_create.restype = C.c_void_p
    lib.dtro_create.argtypes = [C.c_int, C.c_int]
    lib.dtro_destroy.argtypes = [C.c_void_p]
    lib.dtro_color.restype = C.POINTER(C.c_uint32)
    lib.dtro_color.argtypes = [C.c_void_p]
    lib.dtro_zbuffer.restype = _f
    lib.dtro_zbuffer.argtypes = [C.c_void_p]
    lib.dtro_reset_z.argtypes = [C.c_void_p]
    lib.dtro_counter.restype = C.c_uint64
    lib.dtro_counter.argtypes = [C.c_void_p, C.c_int]
    lib.dtro_reset_counters.argtypes = [C.c_void_p]
    lib.dtro_clear.argtypes = [C.c_void_p, _f]
    lib.dtro_triangle.argtypes = [C.c_void_p, _f, _f, _f]
    lib.dtro_triangles.argtypes = [C.c_void_p, C.c_int, _f, _f, _f]
    lib.dtro_textured_triangle.argtypes = [C.c_void_p, _f, _f, _u8, C.c_int, C.c_int, _f, _f]
    lib.dtro_mesh.argtypes = [C.c_void_p, _f, C.c_int, _f, C.c_int, _f, C.c_int, _i32, C.c_int,
                              _u8, C.c_int, C.c_int, C.c_int, _f, _f, _f, _f]
    lib.dtro_rectangle.argtypes = [C.c_void_p, _f, _f, _f, _f]
    lib.dtro_bitmap.argtypes = [C.c_void_p, _u8, C.c_int, C.c_int, _f, _f, _f]
    if hasattr(lib, "dtro_text"):
        lib.dtro_text.argtypes = [C.c_void_p, _u8, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, _f, C.c_char_p, _f, C.c_int]
    lib.dtro_line.argtypes = [C.c_void_p, _i32, _i32, _f]
    _LIBS[kind] = lib
    return lib


def _fa(x, n=None):
    a = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def _fp(a):
    return a.ctypes.data_as(_f)


def transform7(rotation=0.0, anchor=(0.5, 0.5, 0.5), scale=(1.0, 1.0, 1.0)):
    """DTRRenderTransform flattened (DTRendererRender.h:28-33)."""
    return np.array([rotation, *anchor, *scale], dtype=np.float32)


DEFAULT_TRANSFORM = transform7()
DEFAULT_TRIANGLE_TRANSFORM = transform7(anchor=(0.33, 0.33, 0.33))  # DTRendererRender.h:41-47


class Oracle:
    """One headless render target driven through the reference's draw-call semantics."""

    def __init__(self, width, height, kind="port"):
        self.lib = _load(kind)
        self.kind = kind
        self.width, self.height = width, height
        self.ctx = self.lib.dtro_create(width, height)
        if not self.ctx:
            raise MemoryError("dtro_create failed")

    def close(self):
        if self.ctx:
            self.lib.dtro_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    # ---- buffers ---------------------------------------------------------------------------
    def color(self):
        n = self.width * self.height
        return np.ctypeslib.as_array(self.lib.dtro_color(self.ctx), shape=(n,)).reshape(
            self.height, self.width)

    def zbuffer(self):
        n = self.width * self.height
        return np.ctypeslib.as_array(self.lib.dtro_zbuffer(self.ctx), shape=(n,)).reshape(
            self.height, self.width)

    def reset_z(self):
        self.lib.dtro_reset_z(self.ctx)

    def counters(self):
        return (int(self.lib.dtro_counter(self.ctx, 0)), int(self.lib.dtro_counter(self.ctx, 1)))

    def reset_counters(self):
        self.lib.dtro_reset_counters(self.ctx)

    # ---- draw calls (DTRendererRender.h:91-98) --------------------------------------------
    def clear(self, rgb):
        self.lib.dtro_clear(self.ctx, _fp(_fa(rgb, 3)))

    def triangle(self, p, color, transform=DEFAULT_TRIANGLE_TRANSFORM):
        self.lib.dtro_triangle(self.ctx, _fp(_fa(p, 9)), _fp(_fa(color, 4)), _fp(_fa(transform, 7)))

    def triangles(self, p, color, transform=DEFAULT_TRIANGLE_TRANSFORM):
        p = _fa(p)
        n = p.size // 9
        self.lib.dtro_triangles(self.ctx, n, _fp(p), _fp(_fa(color, 4 * n)), _fp(_fa(transform, 7)))

    def textured_triangle(self, p, uv, tex, color, transform=DEFAULT_TRIANGLE_TRANSFORM):
        tex = np.ascontiguousarray(tex, dtype=np.uint8)
        h, w = tex.shape[:2]
        self.lib.dtro_textured_triangle(self.ctx, _fp(_fa(p, 9)), _fp(_fa(uv, 6)),
                                        tex.ctypes.data_as(_u8), w, h, _fp(_fa(color, 4)),
                                        _fp(_fa(transform, 7)))

    def mesh(self, mesh, tex, light_mode, light_vector, light_color, pos=(0, 0, 0),
             transform=DEFAULT_TRANSFORM):
        v = _fa(mesh["vertexes"])
        t = _fa(mesh["texUV"])
        n = _fa(mesh["normals"])
        f = np.ascontiguousarray(mesh["faces"], dtype=np.int32).reshape(-1)
        tex = np.ascontiguousarray(tex, dtype=np.uint8)
        h, w = tex.shape[:2]
        self.lib.dtro_mesh(self.ctx, _fp(v), v.size // 4, _fp(t), t.size // 3, _fp(n), n.size // 3,
                           f.ctypes.data_as(_i32), f.size // 9, tex.ctypes.data_as(_u8), w, h,
                           int(light_mode), _fp(_fa(light_vector, 3)), _fp(_fa(light_color, 4)),
                           _fp(_fa(pos, 3)), _fp(_fa(transform, 7)))

    def rectangle(self, mn, mx, color, transform=DEFAULT_TRANSFORM):
        self.lib.dtro_rectangle(self.ctx, _fp(_fa(mn, 2)), _fp(_fa(mx, 2)), _fp(_fa(color, 4)),
                                _fp(_fa(transform, 7)))

    def bitmap(self, tex, pos, transform=DEFAULT_TRANSFORM, color=(1, 1, 1, 1)):
        tex = np.ascontiguousarray(tex, dtype=np.uint8)
        h, w = tex.shape[:2]
        self.lib.dtro_bitmap(self.ctx, tex.ctypes.data_as(_u8), w, h, _fp(_fa(pos, 2)),
                             _fp(_fa(transform, 7)), _fp(_fa(color, 4)))

    def text(self, font, pos, text, color, length=-1):
        """font = (atlas u8[h, w], packedchars structured array PACKEDCHAR, cpMin, cpMax)."""
        atlas, chars, cp_min, cp_max = font
        atlas = np.ascontiguousarray(atlas, dtype=np.uint8)
        chars = np.ascontiguousarray(chars, dtype=PACKEDCHAR)
        self.lib.dtro_text(self.ctx, atlas.ctypes.data_as(C.POINTER(C.c_uint8)), atlas.shape[1], atlas.shape[0],
                           chars.ctypes.data_as(C.c_void_p), cp_min, cp_max, _fp(_fa(pos, 2)),
                           text.encode("latin-1") if isinstance(text, str) else text, _fp(_fa(color, 4)), length)

    def line(self, a, b, color):
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        self.lib.dtro_line(self.ctx, a.ctypes.data_as(_i32), b.ctypes.data_as(_i32),
                           _fp(_fa(color, 4)))
