"""Seeded synthetic scenes for the five BASELINE.json configs.

A scene is a list of draw commands ``(name, kwargs)`` whose names are the renderer's draw
calls (DTRendererRender.h:91-98: clear / triangle / triangles / textured_triangle / mesh /
rectangle / bitmap).  ``replay(scene, target)`` issues them in order on any object exposing
methods of those names, so the same bytes drive the CUDA back end, the reference build and the
C restatement.  All inputs respect the reference's armed asserts: w == 1, |p| <= 0.9 so that
1 - z_view > 0, uv in [0, 0.999], colours in [0, 1], premultiplied texels (r,g,b <= a).
"""
import numpy as np

SHADE_FULLBRIGHT, SHADE_FLAT, SHADE_GOURAUD = 0, 1, 2


def transform7(rotation=0.0, anchor=(0.5, 0.5, 0.5), scale=(1.0, 1.0, 1.0)):
    """DTRRenderTransform flattened: rotation, anchor.xyz, scale.xyz (DTRendererRender.h:28-33)."""
    return np.array([rotation, *anchor, *scale], dtype=np.float32)


DEFAULT_TRANSFORM = transform7()
DEFAULT_TRIANGLE_TRANSFORM = transform7(anchor=(0.33, 0.33, 0.33))


def uv_sphere(n_lon=50, n_lat=25, radius=0.8):
    """UV sphere, n_lon*n_lat*2 faces (2 500 by default) incl. the degenerate pole triangles.

    Returns the flattened DTRMesh layout (DTRendererAsset.h:27-41): vertexes f32[nV,4] (w = 1),
    texUV f32[nT,3], normals f32[nN,3] (= normalised position), faces i32[nF,9] =
    (v0 v1 v2 t0 t1 t2 n0 n1 n2).
    """
    lat = np.linspace(-0.5 * np.pi, 0.5 * np.pi, n_lat + 1)
    lon = np.linspace(0.0, 2.0 * np.pi, n_lon + 1)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    x = np.cos(la) * np.sin(lo)
    y = np.sin(la)
    z = np.cos(la) * np.cos(lo)
    n = np.stack([x, y, z], -1).reshape(-1, 3)
    pos = (n * radius).astype(np.float32)
    vertexes = np.concatenate([pos, np.ones((pos.shape[0], 1), np.float32)], 1)
    normals = n.astype(np.float32)
    s = np.broadcast_to(np.linspace(0, 1, n_lon + 1)[None, :], la.shape)
    t = np.broadcast_to(np.linspace(0, 1, n_lat + 1)[:, None], la.shape)
    uv = np.stack([s * 0.999, t * 0.999, np.zeros_like(s)], -1).reshape(-1, 3).astype(np.float32)
    faces = []
    w = n_lon + 1
    for i in range(n_lat):
        for j in range(n_lon):
            a, b, c, d = i * w + j, i * w + j + 1, (i + 1) * w + j, (i + 1) * w + j + 1
            faces.append((a, b, c))
            faces.append((b, d, c))
    f = np.asarray(faces, np.int32)
    faces9 = np.concatenate([f, f, f], 1)
    return {"vertexes": vertexes, "texUV": uv, "normals": normals, "faces": faces9}


def random_texture(w, h, seed=1, opaque=True):
    """RGBA8 texture u8[h,w,4] in the DTRBitmap byte order (R,G,B,A), premultiplied in sRGB
    space like DTRAsset_LoadBitmap's output (DTRendererAsset.cpp:825-841)."""
    rng = np.random.default_rng(seed)
    a = np.full((h, w), 255, np.uint32) if opaque else rng.integers(0, 256, (h, w), np.uint32)
    rgb = rng.integers(0, 256, (h, w, 3), np.uint32)
    rgb = (rgb * a[..., None]) // 255  # r,g,b <= a
    return np.concatenate([rgb, a[..., None]], -1).astype(np.uint8)


WHITE_TEXTURE = np.full((1, 1, 4), 255, np.uint8)  # DTRRender_Mesh always samples mesh->tex (:1563)


def mesh_scene(width, height, textured=False, tex_size=1024, rotation_deg=30.0, overlays=0,
               light_mode=SHADE_GOURAUD, seed=1, clear=(0.5, 0.0, 1.0)):
    """cfg 2 (Gouraud sphere) / cfg 3 (textured + alpha overlay quads) style frame."""
    mesh = uv_sphere()
    tex = random_texture(tex_size, tex_size, seed, opaque=True) if textured else WHITE_TEXTURE
    cmds = [("clear", dict(rgb=clear)),
            ("mesh", dict(mesh=mesh, tex=tex, light_mode=light_mode, light_vector=(1, -1, 1),
                          light_color=(1, 1, 1, 1), pos=(0, 0, 0),
                          transform=transform7(rotation_deg, (0, 1, 0), (1, 1, 1))))]
    rng = np.random.default_rng(seed + 100)
    for _ in range(overlays):
        x0 = float(rng.integers(0, width - width // 4))
        y0 = float(rng.integers(0, height - height // 4))
        x1 = x0 + float(rng.integers(width // 16, width // 4))
        y1 = y0 + float(rng.integers(height // 16, height // 4))
        col = (*rng.random(3).astype(np.float32).tolist(), 0.5)
        cmds.append(("rectangle", dict(mn=(x0, y0), mx=(x1, y1), color=col,
                                       transform=DEFAULT_TRANSFORM)))
    return cmds


def view_transforms(n_views):
    """cfg 5: n rotations i*360/n degrees about +Y, expressed through DTRRender_Mesh's own
    transform argument (rotation in DEGREES, axis = transform.anchor; DTRendererRender.cpp:1410)."""
    return [transform7(np.float32(i * 360.0 / n_views), (0, 1, 0), (1, 1, 1))
            for i in range(n_views)]


def small_triangles(width, height, n, seed=7, min_edge=2, max_edge=16):
    """cfg 4: n small random opaque FullBright triangles, integer vertices, z in [0,255]."""
    rng = np.random.default_rng(seed)
    cx = rng.integers(0, width, n).astype(np.float32)
    cy = rng.integers(0, height, n).astype(np.float32)
    ext = rng.integers(min_edge, max_edge + 1, (n, 3, 2)).astype(np.float32)
    sign = rng.integers(0, 2, (n, 3, 2)).astype(np.float32) * 2 - 1
    xy = np.stack([cx, cy], -1)[:, None, :] + np.floor(ext * sign * 0.5)
    z = (rng.random((n, 3)) * 255.0).astype(np.float32)
    p = np.concatenate([xy, z[..., None]], -1).astype(np.float32).reshape(n, 9)
    color = np.concatenate([rng.random((n, 3)), np.ones((n, 1))], 1).astype(np.float32)
    return p, color


def fill_scene(width, height, n, seed=7):
    p, color = small_triangles(width, height, n, seed)
    return [("clear", dict(rgb=(0.0, 0.0, 0.0))),
            ("triangles", dict(p=p, color=color, transform=DEFAULT_TRIANGLE_TRANSFORM))]


def cfg1_scene(width=800, height=600, seed=3):
    """cfg 0/'cfg 1' of SURVEY §8d: flat + alpha-blended triangles (incl. the six demo
    triangles of DTRenderer.cpp:1001-1008 and one rotated), rectangles, a bilinear bitmap."""
    cmds = [("clear", dict(rgb=(0.5, 0.0, 1.0)))]
    w, h = float(width), float(height)
    red, red_t = (0.8, 0.0, 0.0, 1.0), (1.0, 0.0, 0.0, 0.5)
    tris = [
        ((w * .25, h * .25, 0, w * .75, h * .25, 0, w * .5, h * .75, 0), red),
        ((w * .10, h * .10, 10, w * .40, h * .15, 10, w * .20, h * .50, 10), red_t),
        ((w * .60, h * .60, 20, w * .90, h * .65, 5, w * .70, h * .95, 40), red_t),
        ((w * .05, h * .70, 1, w * .30, h * .72, 1, w * .15, h * .98, 1), red),
        ((w * .55, h * .05, 30, w * .95, h * .10, 30, w * .80, h * .45, 30), red_t),
        ((w * .45, h * .40, 50, w * .65, h * .42, 50, w * .50, h * .62, 50), red),
    ]
    for p, c in tris:
        cmds.append(("triangle", dict(p=np.asarray(p, np.float32), color=c,
                                      transform=DEFAULT_TRIANGLE_TRANSFORM)))
    cmds.append(("triangle", dict(p=np.asarray((w * .3, h * .3, 60, w * .7, h * .35, 60,
                                                w * .45, h * .8, 60), np.float32),
                                  color=(0.1, 0.9, 0.2, 0.7),
                                  transform=transform7(0.6, (0.33, 0.33, 0.33), (1.2, 0.8, 1.0)))))
    cmds.append(("rectangle", dict(mn=(w * .05, h * .05), mx=(w * .35, h * .30),
                                   color=(0.0, 1.0, 1.0, 0.6), transform=DEFAULT_TRANSFORM)))
    cmds.append(("rectangle", dict(mn=(w * .55, h * .55), mx=(w * .85, h * .80),
                                   color=(0.0, 1.0, 1.0, 1.0),
                                   transform=transform7(0.4, (0.5, 0.5, 0.5), (1.0, 1.0, 1.0)))))
    tex = random_texture(96, 64, seed, opaque=False)
    cmds.append(("bitmap", dict(tex=tex, pos=(w * .35, h * .30),
                                transform=transform7(0.3, (0.5, 0.5, 0.5), (2.0, 2.0, 1.0)),
                                color=(1.0, 1.0, 1.0, 1.0))))
    cmds.append(("bitmap", dict(tex=tex, pos=(w * .02, h * .80), transform=DEFAULT_TRANSFORM,
                                color=(0.9, 0.8, 1.0, 0.8))))
    return cmds


# stbtt_packedchar (external/stb_truetype.h:522-527): what DTRFont::atlas holds per codepoint
PACKEDCHAR = np.dtype([("x0", "<u2"), ("y0", "<u2"), ("x1", "<u2"), ("y1", "<u2"), ("xoff", "<f4"), ("yoff", "<f4"),
                       ("xadvance", "<f4"), ("xoff2", "<f4"), ("yoff2", "<f4")])


def synthetic_font(seed=1, cp_min=32, cp_max=127, cell=(12, 16), cols=16):
    """A DTRFont-shaped font without a .ttf: (atlas u8[h, w], packedchars[cp_max - cp_min], cp_min,
    cp_max).  Every codepoint gets a random-size glyph of random 8-bit coverage (a third of the
    texels 0) in its own atlas cell, and stb_truetype-style metrics: fractional offsets and advance,
    yoff negative (stb's y axis points down).  One spare row below every glyph keeps the reference's
    off-by-one row read (DTRendererRender.cpp:253) inside the atlas."""
    rng = np.random.default_rng(seed)
    n = cp_max - cp_min
    rows = (n + cols - 1) // cols
    cw, ch = cell
    atlas = np.zeros((rows * ch + 2, cols * cw), np.uint8)
    chars = np.zeros(n, PACKEDCHAR)
    for i in range(n):
        gx, gy = (i % cols) * cw + 1, (i // cols) * ch + 1
        w, h = int(rng.integers(2, cw - 2)), int(rng.integers(3, ch - 3))
        cov = rng.integers(0, 256, (h + 1, w), np.uint8)
        cov[rng.random((h + 1, w)) < 0.33] = 0
        atlas[gy:gy + h + 1, gx:gx + w] = cov
        xoff, yoff = float(rng.uniform(-1.0, 2.0)), float(rng.uniform(-(h + 2.0), -1.0))
        chars[i] = (gx, gy, gx + w, gy + h, xoff, yoff, float(rng.uniform(w, w + 3.0)), xoff + w, yoff + h)
    return atlas, chars, cp_min, cp_max


def text_scene(width, height, seed=4):
    """DTRRender_Text over a cleared frame and a translucent rectangle: strings in several colours,
    partly translucent, some running off the right / bottom / top edges."""
    rng = np.random.default_rng(seed)
    font = synthetic_font(seed)
    scene = [("clear", dict(rgb=(0.1, 0.2, 0.3))),
             ("rectangle", dict(mn=(width * 0.2, height * 0.2), mx=(width * 0.8, height * 0.7), color=(0.9, 0.8, 0.1, 0.6)))]
    words = ["DTRenderer", "B200 back end", "shaded Gpixels/s: 62.2", "The quick brown fox_jumps|over {lazy} dogs 0123456789"]
    for i in range(14):
        pos = (float(rng.uniform(-20, width - 30)), float(rng.uniform(-4, height + 6)))
        color = (*rng.random(3).tolist(), float(rng.choice([1.0, rng.random()])))
        scene.append(("text", dict(font=font, pos=pos, text=words[i % len(words)], color=color)))
    return scene


def replay(scene, target):
    """Issue the scene's draw calls, in order, on ``target``."""
    for name, kw in scene:
        getattr(target, name)(**kw)
