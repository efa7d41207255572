#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libdtr_ref.so, built by `make -C oracle ref` from /root/reference/src).

The reference repository holds no golden vectors, known-answer tests or fixtures for this path
(SURVEY.md §4/§8c), so parity is pinned by outputs of the reference itself, produced here:

  digests.json      sha256 of the colour and depth planes + the reference's own work counters
                    (DTRDebugCounter_SetPixels / _RenderTriangle) for every scene in SCENES
  small_*.npz       full colour+depth planes of small scenes, so a failure can be localised

Run in the authoring container (needs /root/reference):   python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dtrenderer_b200 import scenes  # noqa: E402


def line_scene(w, h, seed=11, n=40):
    """DTRRender_Line in all octants, incl. partly off-screen (SetPixel rejects those pixels)."""
    rng = np.random.default_rng(seed)
    cmds = [("clear", dict(rgb=(0.1, 0.2, 0.3)))]
    for _ in range(n):
        a = rng.integers(-20, [w + 20, h + 20]).astype(np.int32)
        b = rng.integers(-20, [w + 20, h + 20]).astype(np.int32)
        col = (*rng.random(3).tolist(), float(rng.choice([1.0, 0.5])))
        cmds.append(("line", dict(a=a, b=b, color=col)))
    cmds.append(("line", dict(a=np.array([5, 5], np.int32), b=np.array([5, 5], np.int32), color=(1, 1, 1, 1))))
    cmds.append(("line", dict(a=np.array([5, 7], np.int32), b=np.array([50, 7], np.int32), color=(1, 1, 1, 1))))
    cmds.append(("line", dict(a=np.array([9, 3], np.int32), b=np.array([9, 60], np.int32), color=(1, 0, 1, 0.3))))
    return cmds


def edge_scene(w, h):
    """Edge cases the reference's code paths distinguish: degenerate (zero-area) triangles,
    fully off-screen and partly clipped primitives, both windings, shared edges with equal depth
    (strict `>` keeps the first), translucent over opaque, clear in the middle of a frame."""
    T = scenes.DEFAULT_TRIANGLE_TRANSFORM
    cmds = [("clear", dict(rgb=(0.2, 0.2, 0.2)))]
    tri = lambda p, c, t=T: ("triangle", dict(p=np.asarray(p, np.float32), color=c, transform=t))  # noqa: E731
    cmds += [
        tri((10, 10, 5, 10, 10, 5, 10, 10, 5), (1, 0, 0, 1)),                    # point
        tri((10, 10, 5, 60, 10, 5, 110, 10, 5), (1, 0, 0, 1)),                   # collinear
        tri((-200, -200, 1, -100, -200, 1, -150, -100, 1), (0, 1, 0, 1)),        # off-screen
        tri((w - 30, h - 30, 9, w + 50, h - 20, 9, w - 10, h + 40, 9), (0, 0, 1, 1)),  # clipped top-right
        tri((-20, 20, 9, 40, 10, 9, 10, 70, 9), (0, 1, 1, 1)),                   # clipped left
        tri((100, 50, 20, 160, 50, 20, 100, 110, 20), (1, 1, 0, 1)),             # CCW
        tri((100, 110, 20, 160, 50, 20, 160, 110, 20), (1, 0, 1, 1)),            # shares the diagonal, same z
        tri((100, 50, 20, 100, 110, 20, 160, 50, 20), (0.3, 0.9, 0.3, 1)),       # same triangle again: z tie loses
        tri((90, 40, 30, 170, 45, 30, 120, 120, 30), (1, 1, 1, 0.4)),            # translucent in front
        tri((90, 40, 10, 170, 45, 10, 120, 120, 10), (0, 0, 0, 1)),              # behind: rejected by z
    ]
    cmds.append(("clear", dict(rgb=(0.0, 0.3, 0.0))))                            # colour only, z stays
    cmds += [
        tri((95, 45, 25, 165, 50, 25, 125, 115, 25), (1, 0.5, 0, 1)),            # z-tested against the OLD depth
        tri((200, 20, 1, 280, 30, 200, 230, 150, 90), (0.7, 0.7, 1, 0.9),
            scenes.transform7(-1.1, (0.1, 0.9, 0), (0.7, 1.6, 1))),              # rotated/scaled: sequential path
    ]
    cmds.append(("rectangle", dict(mn=(-10, -10), mx=(30, 25), color=(1, 1, 1, 0.5), transform=scenes.DEFAULT_TRANSFORM)))
    cmds.append(("rectangle", dict(mn=(w - 40, h - 20), mx=(w + 30, h + 30), color=(1, 0, 0, 1),
                                   transform=scenes.transform7(0.9, (0.5, 0.5, 0), (1, 1, 1)))))
    tex = scenes.random_texture(17, 9, 5, opaque=False)
    cmds.append(("bitmap", dict(tex=tex, pos=(w - 60.5, 3.25), transform=scenes.transform7(2.0, (0.2, 0.7, 0), (5, 3, 1)),
                                color=(1, 1, 1, 1))))
    cmds.append(("bitmap", dict(tex=tex, pos=(-5, h - 12), transform=scenes.DEFAULT_TRANSFORM, color=(0.5, 1, 0.25, 0.75))))
    uvtex = scenes.random_texture(32, 16, 6, opaque=False)
    cmds.append(("textured_triangle", dict(p=np.asarray((40, 130, 50, 150, 140, 60, 90, 190, 70), np.float32),
                                           uv=np.asarray((0, 0, 0.999, 0.1, 0.4, 0.999), np.float32), tex=uvtex,
                                           color=(1, 0.9, 0.8, 0.9), transform=T)))
    return cmds


# name -> (width, height, scene factory).  Sizes keep the whole CPU suite at a few minutes.
SCENES = {
    "cfg1_800x600": (800, 600, lambda: scenes.cfg1_scene(800, 600)),
    "cfg2_gouraud_1080p": (1920, 1080, lambda: scenes.mesh_scene(1920, 1080)),
    "cfg3_textured_overlays_720p": (1280, 720, lambda: scenes.mesh_scene(1280, 720, textured=True, tex_size=256, overlays=8)),
    "flat_640x480": (640, 480, lambda: scenes.mesh_scene(640, 480, textured=True, tex_size=64, light_mode=scenes.SHADE_FLAT)),
    "fullbright_640x480": (640, 480, lambda: scenes.mesh_scene(640, 480, textured=True, tex_size=64,
                                                               light_mode=scenes.SHADE_FULLBRIGHT)),
    "cfg4_fill_20k_1024x768": (1024, 768, lambda: scenes.fill_scene(1024, 768, 20000)),
    "odd_333x217": (333, 217, lambda: scenes.cfg1_scene(333, 217) + scenes.mesh_scene(333, 217, textured=True, tex_size=32)[1:]),
    "edges_320x200": (320, 200, lambda: edge_scene(320, 200)),
    "lines_256x192": (256, 192, lambda: line_scene(256, 192)),
    "text_320x200": (320, 200, lambda: scenes.text_scene(320, 200)),
    "view_37_of_4096_640x360": (640, 360, lambda: [("clear", dict(rgb=(0.5, 0, 1)))] + [
        ("mesh", dict(mesh=scenes.uv_sphere(), tex=scenes.random_texture(128, 128, 1, True),
                      light_mode=scenes.SHADE_GOURAUD, light_vector=(1, -1, 1), light_color=(1, 1, 1, 1),
                      pos=(0.1, -0.05, 0.0), transform=scenes.view_transforms(4096)[37]))]),
}
SMALL = ["edges_320x200", "lines_256x192", "odd_333x217", "text_320x200"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def render(kind, name):
    from oracle import dtro
    w, h, make = SCENES[name]
    o = dtro.Oracle(w, h, kind)
    o.reset_counters()
    scenes.replay(make(), o)
    sp, tr = o.counters()
    return o.color().copy(), o.zbuffer().copy(), sp, tr


def main():
    from oracle import dtro
    if not dtro.available("reference"):
        raise SystemExit("oracle/_ref/libdtr_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for name in SCENES:
        col, z, sp, tr = render("reference", name)
        out[name] = {"color_sha256": digest(col), "depth_sha256": digest(z), "setPixels": sp, "triangles": tr,
                     "width": SCENES[name][0], "height": SCENES[name][1]}
        if name in SMALL:
            np.savez_compressed(os.path.join(HERE, f"small_{name}.npz"), color=col, depth=z)
        print(name, out[name]["color_sha256"][:12], sp, tr)
    json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libdtr_ref.so (unmodified reference, "
               "DTR_DEBUG_RENDER off)", "scenes": out}, open(os.path.join(HERE, "digests.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
