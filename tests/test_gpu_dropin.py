"""GPU drop-in proof (-m gpu): the reference's OWN types and draw-call signatures (DTRRenderContext,
DTRRenderTransform, DTRMesh, DqnV*) drive the CUDA back end through the host-side mirror
dtrenderer_b200/host/DTRRenderB200.h, inside the same headless harness that drives the unmodified
reference.  oracle/_ref/libdtr_ref_b200.so (built where the reference headers exist, shipped as a
binary) must produce the frames oracle/_ref/libdtr_ref.so does."""
import numpy as np
import pytest

from dtrenderer_b200 import scenes
from golden_scenes import DIGESTS, SCENES, digest

pytestmark = pytest.mark.gpu

NAMES = ["cfg1_800x600", "edges_320x200", "lines_256x192", "text_320x200", "odd_333x217", "flat_640x480",
         "cfg3_textured_overlays_720p", "view_37_of_4096_640x360"]


@pytest.mark.parametrize("name", NAMES)
def test_reference_api_on_b200_matches_reference(built, name):
    from oracle import dtro
    if not (dtro.available("reference_api_b200") and dtro.available("reference")):
        pytest.skip("oracle/_ref drop-in harness not built (needs the reference headers at build time)")
    w, h, make = SCENES[name]
    scene = make()
    ref = dtro.Oracle(w, h, "reference")
    ref.reset_counters()
    scenes.replay(scene, ref)
    gpu = dtro.Oracle(w, h, "reference_api_b200")
    assert gpu.lib.dtro_kind() == b"reference-api-on-b200"
    gpu.reset_counters()
    scenes.replay(scene, gpu)
    assert np.array_equal(gpu.zbuffer().view(np.uint32), ref.zbuffer().view(np.uint32))
    assert np.array_equal(gpu.color(), ref.color())
    assert gpu.counters() == ref.counters()
    assert digest(gpu.color()) == DIGESTS[name]["color_sha256"]


def test_two_frames_keep_colour_and_reset_depth(built):
    """Frame protocol of DTR_Update: depth is reset every frame, colour persists until cleared."""
    from oracle import dtro
    if not (dtro.available("reference_api_b200") and dtro.available("reference")):
        pytest.skip("oracle/_ref drop-in harness not built")
    w, h = 320, 200
    outs = []
    for kind in ("reference", "reference_api_b200"):
        o = dtro.Oracle(w, h, kind)
        scenes.replay(scenes.cfg1_scene(w, h), o)
        o.color()  # present frame 1
        o.reset_z()  # frame 2: no clear, draws on top of frame 1's colour
        o.triangle(np.asarray((20, 20, 3, 300, 40, 3, 150, 190, 3), np.float32), (0.2, 0.4, 0.9, 0.6))
        outs.append((o.color().copy(), o.zbuffer().copy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
