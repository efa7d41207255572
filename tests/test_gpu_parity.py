"""GPU parity: the CUDA back end, driven through the C ABI, against the CPU oracle on the same
seeded scenes.  Bar (BASELINE.json north_star): coverage and depth decisions bit-exact, depth
values bit-exact (tolerance 1e-6 relative is the reporting threshold only), colour within 1/255
per channel -- in practice we require bit-exact colour too and report the looser bar on failure."""
import numpy as np
import pytest

from dtrenderer_b200 import scenes

pytestmark = pytest.mark.gpu


def _oracle(w, h):
    from oracle import dtro
    kind = "reference" if dtro.available("reference") else "port"
    return dtro.Oracle(w, h, kind)


def _render_gpu(w, h, scene, frames=1):
    from dtrenderer_b200 import api
    r = api.Renderer(w, h, frames, 0)
    r.begin_frame(0)
    scenes.replay(scene, r)
    col, z = r.end_frame(0)
    return r, col, z


def _check(w, h, scene):
    o = _oracle(w, h)
    o.reset_counters()
    scenes.replay(scene, o)
    r, col, z = _render_gpu(w, h, scene)
    zo = o.zbuffer()
    cov_gpu, cov_ref = z != np.float32(-3.4028234663852886e38), zo != np.float32(-3.4028234663852886e38)
    assert np.array_equal(cov_gpu, cov_ref), f"coverage differs on {(cov_gpu != cov_ref).sum()} px"
    assert np.array_equal(z.view(np.uint32), zo.view(np.uint32)), \
        f"depth differs on {(z.view(np.uint32) != zo.view(np.uint32)).sum()} px"
    co = o.color()
    if not np.array_equal(col, co):
        d = np.abs(col.view(np.uint8).astype(int) - co.view(np.uint8).astype(int))
        raise AssertionError(f"colour differs on {(col != co).sum()} px, max channel delta {d.max()}")
    sp, tris = o.counters()
    st = r.stats()
    assert st["setPixels"] == sp, (st, sp)
    assert st["triangles"] == tris, (st, tris)
    return r


def test_cfg1_flat_alpha_triangles_rect_bitmap(built):
    _check(800, 600, scenes.cfg1_scene(800, 600))


def test_cfg2_gouraud_mesh_1080p(built):
    _check(1920, 1080, scenes.mesh_scene(1920, 1080))


def test_cfg3_textured_mesh_overlays(built):
    _check(1280, 720, scenes.mesh_scene(1280, 720, textured=True, tex_size=256, overlays=8))


@pytest.mark.parametrize("mode", [scenes.SHADE_FULLBRIGHT, scenes.SHADE_FLAT, scenes.SHADE_GOURAUD])
def test_shading_modes(built, mode):
    _check(640, 480, scenes.mesh_scene(640, 480, textured=True, tex_size=64, light_mode=mode))


def test_cfg4_small_triangles(built):
    _check(1024, 768, scenes.fill_scene(1024, 768, 20000))


def test_odd_resolution_partial_tiles(built):
    _check(333, 217, scenes.cfg1_scene(333, 217) + scenes.mesh_scene(333, 217, textured=True, tex_size=32)[1:])
